"""Scratch: do K sub-shards of the resident runs on K CUDA streams hide the tail of the fit grid?  usage: subshard_test.py R iters K"""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import bench
import auction_gym_b200 as ag
from auction_gym_b200 import _lib
R, N, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
T = bench.WORKLOAD["T"]
engs, streams = [], []
for k in range(K):
    r0, rn = k * R // K, R // K
    e = bench.make_engine(ag, _lib, rn, T, True, 0, r0)
    e.set_allocator_state(bench.initial_m(r0, rn))
    engs.append(e); streams.append(torch.cuda.Stream())
def run(n0, n):
    torch.cuda.synchronize(); t0 = time.time()
    for it in range(n0, n0 + n):
        for e, s in zip(engs, streams):
            with torch.cuda.stream(s):
                e.clear_iteration(); e.simulate(0, it, T); e.update_allocators(want_info=False)
    torch.cuda.synchronize(); return time.time() - t0
run(0, 5)
dt = run(5, N)
print(f"R={R} K={K}: {dt / N * 1e3:.1f} ms per iteration over iterations 5..{5 + N - 1}; checksum {sum(float(e.m.double().sum()) for e in engs):.3f}")
