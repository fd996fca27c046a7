python -m pytest tests/test_gpu_fit.py tests/test_gpu_shapes.py -m gpu -q 2>&1 | tail -3
for h in 100000 32 16 64 24; do
  AGYM_FIT_HEAVY=$h python bench.py --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('heavy=$h', 'step_ms', round(d['ms_per_step'],1), 'K6_ms', round(d['roofline_kernels']['bucket_kernel + fit_kernel (K6)']['ms'],1), 'value', round(d['value']))
"
done
