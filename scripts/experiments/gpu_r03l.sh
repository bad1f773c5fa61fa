#!/bin/bash
timeout 600 python -m pytest tests -m gpu -x -q -k "advection or Advection or fourier or predicted or example" 2>&1 | tail -4
