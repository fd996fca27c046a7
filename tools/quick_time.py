"""Scratch timing of the main kernels at the synthetic shape (not the bench)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import auction_gym_b200 as ag
from auction_gym_b200 import _lib
from oracle import auction_oracle as ao

R = int(sys.argv[1]) if len(sys.argv) > 1 else 128
T = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
A = I = 64; D, Do, P = 5, 4, 2
rng = np.random.default_rng(0)
E, V = ao.make_catalog(rng, A, I, D)
def ev(fn, n=3):
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(n):
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    return min(ts)
for kind, name in ((_lib.ALLOC_ORACLE, "oracle"), (_lib.ALLOC_TS, "ts")):
    eng = ag.Engine(R=R, A=A, I=I, D=D, Do=Do, P=P, mechanism=0, E=E, V=V, n_items=[I]*A, alloc_kind=[kind]*A,
                    bidder_kind=[0]*A, precision=_lib.FP32, rounds_capacity=T)
    if kind != 0:
        eng.set_allocator_state(torch.randn(R, A, I, Do+1))
    def rounds():
        eng.clear_iteration(); eng.simulate(1, 0, T)
    ms = ev(rounds)
    print(f"{name}: fused rounds R={R} T={T}: {ms:.2f} ms -> {R*T/ms/1e3:.3e} opp/s", flush=True)
    if kind != 0:
        for it in range(3):
            eng.clear_iteration(); eng.simulate(1, it, T)
            torch.cuda.synchronize(); t0 = time.time()
            info = eng.update_allocators()
            torch.cuda.synchronize(); dt = time.time() - t0
            inf = info.cpu().numpy()
            print(f"  iter {it}: fit {dt*1e3:.1f} ms; epochs mean {inf[...,1].mean():.0f} max {inf[...,1].max():.0f}; rows mean {inf[...,3].mean():.0f} max {inf[...,3].max():.0f}; stop<0 frac {(inf[...,0]<0).mean():.3f}", flush=True)
    # staged K4
    eng.clear_iteration()
    b = eng.staged_round(1, 0, T)
    for acc in (0, 1):
        ms4 = ev(lambda: eng.k4_resolve(1, 0, T, b, accumulate=bool(acc)))
        print(f"  K4 accumulate={acc}: {ms4:.3f} ms -> {R*T*36/ms4/1e6:.1f} GB/s algorithmic, {R*T/ms4/1e3:.3e} opp/s", flush=True)
    eng.close()
