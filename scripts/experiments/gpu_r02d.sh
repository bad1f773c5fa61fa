#!/bin/bash
tag=${1:-r02d}
mkdir -p gpurun_out
for shp in 32,33 64,17 128,9; do
  echo "== sine shape $shp"
  MGB_SHAPE_7=$shp python scripts/solve_timeline.py cfg5 2>&1 | head -12
done > gpurun_out/${tag}_sine_shapes.txt 2>&1
cat gpurun_out/${tag}_sine_shapes.txt
MGB_HEAT1D_SINE=0 python scripts/solve_timeline.py cfg5 2>&1 | head -8 | tee gpurun_out/${tag}_timeline_node.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py --no-cpu > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
cut -c1-3000 gpurun_out/${tag}_bench.json; tail -5 gpurun_out/${tag}_bench.err
