"""Scratch: time the fused round-loop kernel at the bench shape."""
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
for kind, name in ((1, "ts"), (0, "oracle")):
    eng = ag.Engine(R=R, A=A, I=I, D=D, Do=Do, P=P, mechanism=0, E=E, V=V, n_items=[I]*A, alloc_kind=[kind]*A, bidder_kind=[0]*A, precision=_lib.FP32, rounds_capacity=T)
    import os
    if os.environ.get('SIM_G'): eng.set_option('sim_g', float(os.environ['SIM_G']))
    for kv in filter(None, os.environ.get('SIM_OPTS', '').split(',')): eng.set_option(kv.split('=')[0], float(kv.split('=')[1]))
    if kind: eng.set_allocator_state(torch.randn(R, A, I, Do+1, generator=torch.Generator().manual_seed(0)))
    ts = []
    for it in range(6):
        eng.clear_iteration()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); eng.simulate(1, it, T); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    acc, rev = eng.metrics()
    print(f"{name}: {min(ts[2:]):.3f} ms  -> {R*T/min(ts[2:])*1e3:.3e} opp/s   (revenue mean {rev.mean():.3f}, welfare {acc[...,1].sum(1).mean():.3f})")
    eng.close()
