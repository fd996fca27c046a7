"""Scratch: time K6 (update_allocators) at the synthetic shape from the same starting state."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import auction_gym_b200 as ag
from auction_gym_b200 import _lib
from oracle import auction_oracle as ao
R = int(sys.argv[1]) if len(sys.argv) > 1 else 128
T = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
A = int(sys.argv[3]) if len(sys.argv) > 3 else 64
I = int(sys.argv[4]) if len(sys.argv) > 4 else 64
ME = int(sys.argv[5]) if len(sys.argv) > 5 else 0
REPS = int(sys.argv[6]) if len(sys.argv) > 6 else 3
D, Do, P = 5, 4, 2
rng = np.random.default_rng(0)
E, V = ao.make_catalog(rng, A, I, D)
eng = ag.Engine(R=R, A=A, I=I, D=D, Do=Do, P=P, mechanism=0, E=E, V=V, n_items=[I]*A, alloc_kind=[1]*A, bidder_kind=[0]*A, precision=_lib.FP32, rounds_capacity=T)
m0 = torch.randn(R, A, I, Do+1, generator=torch.Generator().manual_seed(0))
eng.set_allocator_state(m0)
eng.simulate(1, 0, T)
torch.cuda.synchronize()
res = []
for rep in range(REPS):
    eng.set_allocator_state(m0, torch.ones_like(m0))
    torch.cuda.synchronize(); t0 = time.time()
    info = eng.update_allocators(max_epochs=ME, fit_mode=int(__import__('os').environ.get('FIT_MODE','0')))
    torch.cuda.synchronize(); dt = time.time() - t0
    inf = info.cpu().numpy()
    res.append(dt)
ep = inf[..., 1].sum()
print(f"R={R} T={T} A={A} I={I}: fit {min(res)*1e3:.1f} ms; fits {R*A}; epochs mean {inf[...,1].mean():.0f}; fit-epochs/s {ep/min(res):.3e}; SM-cycles per fit-epoch {148*1.965e9*min(res)/ep:.0f}; m checksum {eng.m.double().sum().item():.6f}")
