"""Scratch: per-iteration K6 time at the bench shape (the row distribution over items concentrates as the allocators learn).
usage: fit_iters.py [R] [N] ; env FIT_MODE=0|1, FIT_OPTS="fit_ncap=1.2,fit_warp=0" """
import os, sys, time
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
for kv in filter(None, os.environ.get("FIT_OPTS", "").split(",")):
    k, v = kv.split("="); eng.set_option(k, float(v))
eng.set_allocator_state(torch.randn(R, A, I, Do + 1, generator=torch.Generator().manual_seed(0)))
mode = int(os.environ.get("FIT_MODE", "0"))
tot = 0.0
prev_ep = None
for it in range(N):
    eng.clear_iteration()
    eng.simulate(1, it, T)
    meta = eng.fit_meta[:, :T].to(torch.int64) & 0xFFFFFFFF
    key = ((meta >> 12) & 0xFFF) * 64 + (meta & 0xFFF)                      # (agent, item) of every winner record
    onehot = torch.zeros(R, A * 64, device=meta.device).scatter_(1, key, 1.0)
    n_active = onehot.view(R, A, 64).sum(-1)                                 # distinct items per fit
    torch.cuda.synchronize(); t0 = time.time()
    info = eng.update_allocators(want_info=True, fit_mode=mode)
    torch.cuda.synchronize(); dt = time.time() - t0
    tot += dt
    n = info[..., 3]
    ep = info[..., 1].flatten().double()
    corr = float(torch.corrcoef(torch.stack([ep, prev_ep]))[0, 1]) if prev_ep is not None else float("nan")
    prev_ep = ep
    print(f"it {it:3d}  {dt * 1e3:7.1f} ms  epochs mean {info[..., 1].mean().item():7.0f} max {info[..., 1].max().item():6.0f}  "
          f"rows mean {n.mean().item():5.0f} p99 {n.flatten().kthvalue(int(0.99 * n.numel())).values.item():4.0f} max {n.max().item():4.0f}  "
          f"active items mean {n_active.mean().item():5.1f} max {n_active.max().item():3.0f} >12 {(n_active > 12).float().mean().item():.3f} "
          f">19 {(n_active > 19).float().mean().item():.3f}  epoch corr with previous {corr:.2f}", flush=True)
print(f"total {tot * 1e3:.0f} ms over {N} iterations; checksum {eng.m.double().sum().item():.4f}")
