"""Scratch: time / profile the staged resolution kernel K4 (+K5) at the bench shape (512 runs x 10 000 rounds, P = 2)."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import bench
import auction_gym_b200 as ag
from auction_gym_b200 import _lib
R = int(sys.argv[1]) if len(sys.argv) > 1 else 512
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
T = bench.WORKLOAD["T"]
eng = bench.make_engine(ag, _lib, R, T, True, 0, 0)
eng.set_allocator_state(bench.initial_m(0, R))
eng.clear_iteration()
b = eng.staged_round(0, 0, T)
flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device="cuda")
stream = torch.cuda.current_stream()
import os
for acc, variant in [(False, 0), (True, 0)]:
    for _ in range(3):
        eng.k4_resolve(0, 0, T, b, acc)
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(stream); eng.k4_resolve(0, 0, T, b, acc); e.record(stream)
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    by = R * T * 36
    print(f"accumulate={acc} variant={variant}: {np.mean(ts) * 1e3:.1f} us (min {np.min(ts) * 1e3:.1f}), {by / np.mean(ts) / 1e6:.0f} GB/s algorithmic")
