python -m pytest tests/test_gpu_fit.py -m gpu -q 2>&1 | tail -1 | cut -c1-300
python tools/fit_iters.py 512 6 2>&1 | tail -1 | sed "s/^/cur: /"
