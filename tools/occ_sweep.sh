python -m pytest tests/test_gpu_fit.py tests/test_gpu_shapes.py tests/test_gpu_retention.py -m gpu -q 2>&1 | tail -2 | cut -c1-300
python tools/fit_iters.py 512 6 2>&1 | tail -1 | sed "s/^/split: /"
AGYM_FIT_HEAVY=96 python tools/fit_iters.py 512 6 2>&1 | tail -1 | sed "s/^/split heavy96: /"
AGYM_FIT_HEAVY=32 python tools/fit_iters.py 512 6 2>&1 | tail -1 | sed "s/^/split heavy32: /"
