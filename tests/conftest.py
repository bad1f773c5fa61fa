import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run with -m gpu on the B200 box)")
    config.addinivalue_line("markers", "slow: long-running CPU test")


def pytest_collection_modifyitems(config, items):
    """Without a CUDA device the GPU tests are skipped instead of failing one by one (`pytest tests` on a CPU box stays
    readable).  With a device they always run: a missing libmgrit_b200.so then fails loudly, as it must."""
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason='needs a CUDA device (run with -m gpu on the B200 box)')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)
