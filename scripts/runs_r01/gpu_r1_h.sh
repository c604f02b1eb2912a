set -x
timeout 900 python -m pytest tests/test_gpu_slab.py -x -q 2>&1 | tail -15
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
