"""Scratch: where the end-to-end step of bench.py spends its time (CUDA events around each phase)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import auction_gym_b200 as ag
from auction_gym_b200 import _lib
from oracle import auction_oracle as ao
R, T, A, I, D, Do, P = 512, 10000, 64, 64, 5, 4, 2
K = Do + 1
dev = torch.device("cuda", 0)
E, V = ao.make_catalog(np.random.default_rng(0), A, I, D)
eng = ag.Engine(R=R, A=A, I=I, D=D, Do=Do, P=P, mechanism=0, E=E, V=V, n_items=[I] * A, alloc_kind=[1] * A, bidder_kind=[0] * A,
                precision=_lib.FP32, rounds_capacity=T)
m_host = torch.randn((R, A, I, K), generator=torch.Generator().manual_seed(0)).pin_memory()
q_host = torch.ones((R, A, I, K)).pin_memory()
mp_host = m_host.clone().pin_memory()
acc_host = torch.empty((R, A, _lib.NUM_METRICS), dtype=torch.float64).pin_memory()
rev_host = torch.empty((R,), dtype=torch.float64).pin_memory()
stream = torch.cuda.current_stream(dev)
def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(stream); return e
for it in range(8):
    t0 = time.perf_counter()
    e0 = ev()
    eng.set_allocator_state(m_host, q_host, mp_host, non_blocking=True)
    e1 = ev()
    eng.clear_iteration(); eng.simulate(0, it, T)
    e2 = ev()
    eng.update_allocators(want_info=False)
    e3 = ev()
    m_host.copy_(eng.m, non_blocking=True); q_host.copy_(eng.q, non_blocking=True); mp_host.copy_(eng.m_prev, non_blocking=True)
    acc_host.copy_(eng.acc, non_blocking=True); rev_host.copy_(eng.revenue, non_blocking=True)
    e4 = ev()
    t_queued = time.perf_counter() - t0
    stream.synchronize()
    wall = time.perf_counter() - t0
    print(f"it {it}: h2d+sigma {e0.elapsed_time(e1):6.2f}  rounds {e1.elapsed_time(e2):6.2f}  fit {e2.elapsed_time(e3):7.2f}  d2h {e3.elapsed_time(e4):6.2f}  "
          f"total {e0.elapsed_time(e4):7.2f} ms   host: queued after {t_queued * 1e3:6.2f} ms, wall {wall * 1e3:7.2f} ms")
