"""Scratch: a few launches of the fused round-loop kernel and the staged K4 at the bench shape (for ncu)."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import auction_gym_b200 as ag
from auction_gym_b200 import _lib
from oracle import auction_oracle as ao
R = int(sys.argv[1]) if len(sys.argv) > 1 else 512
T = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
A = I = 64; D, Do, P = 5, 4, 2
E, V = ao.make_catalog(np.random.default_rng(0), A, I, D)
eng = ag.Engine(R=R, A=A, I=I, D=D, Do=Do, P=P, mechanism=0, E=E, V=V, n_items=[I]*A, alloc_kind=[1]*A, bidder_kind=[0]*A, precision=_lib.FP32, rounds_capacity=T)
eng.set_allocator_state(torch.randn(R, A, I, Do+1, generator=torch.Generator().manual_seed(0)))
for it in range(3):
    eng.clear_iteration(); eng.simulate(1, it, T)
b = eng.staged_round(1, 0, T)
for acc in (False, True, False, True):
    eng.k4_resolve(1, 0, T, b, acc)
torch.cuda.synchronize()
print("done")
