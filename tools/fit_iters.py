"""Scratch: per-iteration K6 time at the bench shape (the row distribution over items concentrates as the allocators learn)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import auction_gym_b200 as ag
from auction_gym_b200 import _lib
from oracle import auction_oracle as ao
R = int(sys.argv[1]) if len(sys.argv) > 1 else 512
N = int(sys.argv[2]) if len(sys.argv) > 2 else 6
T, A, I, D, Do, P = 10000, 64, 64, 5, 4, 2
E, V = ao.make_catalog(np.random.default_rng(0), A, I, D)
eng = ag.Engine(R=R, A=A, I=I, D=D, Do=Do, P=P, mechanism=0, E=E, V=V, n_items=[I] * A, alloc_kind=[1] * A, bidder_kind=[0] * A,
                precision=_lib.FP32, rounds_capacity=T)
eng.set_allocator_state(torch.randn(R, A, I, Do + 1, generator=torch.Generator().manual_seed(0)))
out = []
for it in range(N):
    eng.clear_iteration()
    eng.simulate(1, it, T)
    torch.cuda.synchronize(); t0 = time.time()
    info = eng.update_allocators(want_info=True, fit_mode=int(__import__('os').environ.get('FIT_MODE', '0')))
    torch.cuda.synchronize(); dt = time.time() - t0
    out.append(f"{dt * 1e3:.0f}ms/{info[..., 1].mean().item():.0f}ep")
print(" ".join(out), "checksum", f"{eng.m.double().sum().item():.4f}")
