python -m pytest tests/test_gpu_masks.py -m gpu -x -q 2>&1 | tail -3
python tools/bench_overlap.py 2>&1 | tail -2
