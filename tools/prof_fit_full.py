"""Scratch: three iterations at the full bench shape; the third allocator-fit launch is the one to capture with ncu."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import auction_gym_b200 as ag  # noqa: E402
from auction_gym_b200 import _lib  # noqa: E402
from oracle import auction_oracle as ao  # noqa: E402

R, T, A, I, D, Do, P = (int(sys.argv[1]) if len(sys.argv) > 1 else 512), 10000, 64, 64, 5, 4, 2
E, V = ao.make_catalog(np.random.default_rng(0), A, I, D)
eng = ag.Engine(R=R, A=A, I=I, D=D, Do=Do, P=P, mechanism=0, E=E, V=V, n_items=[I] * A, alloc_kind=[1] * A, bidder_kind=[0] * A,
                precision=_lib.FP32, rounds_capacity=T)
eng.set_allocator_state(torch.randn(R, A, I, Do + 1, generator=torch.Generator().manual_seed(0)))
for it in range(3):
    eng.clear_iteration()
    eng.simulate(1, it, T)
    eng.update_allocators(want_info=False)
torch.cuda.synchronize()
print("done")
