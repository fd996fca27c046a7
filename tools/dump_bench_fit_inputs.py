"""Writes tests/golden/bench_fit_inputs.npz (run on a GPU box: `python tools/dump_bench_fit_inputs.py`).

The fit inputs -- won rows, m, q (prev_iter_m == m at the start of every fit) -- of 32 fits per iteration among global runs 0-7
(the fits at the 32 evenly spaced quantiles of the engine's own epoch counts, so that the sample mean tracks the population's), for
iterations 0 .. N-1 of bench.py's learning trajectory (same seed, same per-run initial state, same Philox keys as the
bench's own runs 0 and 1), plus the epochs the engine's fit ran for each of them.  `bench.py --impl reference` and the
`cpu_baseline` leg fit exactly these inputs with the unmodified reference / the numpy port, so the CPU arms time the
same iterations of the same trajectory as the GPU arm.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import auction_gym_b200 as ag  # noqa: E402
from auction_gym_b200 import _lib  # noqa: E402

N_IT = int(sys.argv[1]) if len(sys.argv) > 1 else 26
N_FITS, R = 32, 8
w = bench.WORKLOAD
A, I, D, Do, P, T = w["A"], w["I"], w["D"], w["Do"], w["P"], w["T"]
eng = bench.make_engine(ag, _lib, R=R, T=T, learnt=True, device=0, run_offset=0)
eng.set_allocator_state(bench.initial_m(0, R))
out = {"n_iterations": N_IT, "fits_per_iteration": N_FITS, "seed": bench.SEED, "T": T}
# Which fits: a sample STRATIFIED by the engine's own epoch count.  The epochs of the R * A = 512 fits of the first 8 runs spread
# over a factor of 4 within an iteration, so 32 arbitrary fits miss the population mean by several per cent; the 32 fits at the
# (k + 0.5) / 31 quantiles of that distribution track it (and so do the first 16: the mean fit and the 15 odd quantiles, symmetric about the median).
# Fit 0 is the one whose epoch count is nearest the population mean (the one-fit-per-iteration cpu_baseline leg times only it).
order = list(range(1, N_FITS - 1, 2)) + list(range(0, N_FITS - 1, 2))  # 15 odd quantiles (symmetric about the median), then the 16 even ones
for it in range(N_IT):
    eng.clear_iteration()
    eng.simulate(bench.SEED, it, T)
    meta = eng.fit_meta[:R, :T].cpu().numpy().view(np.uint32)
    ctx = eng.fit_ctx[:R, :T].cpu().numpy()
    m, q, mp = (t[:R].cpu().numpy().copy() for t in (eng.m, eng.q, eng.m_prev))
    assert np.array_equal(m, mp)
    info = eng.update_allocators(want_info=True).cpu().numpy()
    ep = info[..., 1].reshape(-1)
    by_epochs = np.argsort(ep, kind="stable")
    picks = [int(np.argmin(np.abs(ep - ep.mean())))] + [int(by_epochs[int((k + 0.5) * len(ep) / (N_FITS - 1))]) for k in order]
    for j, f in enumerate(picks):
        r, a = divmod(f, A)
        valid, agent, item, click = meta[r] >> 31, (meta[r] >> 12) & 0xFFF, meta[r] & 0xFFF, (meta[r] >> 30) & 1
        sel = (valid == 1) & (agent == a)
        k = f"it{it}_f{j}_"
        out[k + "X"] = np.concatenate([ctx[r][sel], np.ones((sel.sum(), 1), np.float32)], axis=1)
        out[k + "items"] = item[sel].astype(np.uint8)
        out[k + "y"] = click[sel].astype(np.uint8)
        out[k + "m"], out[k + "q"], out[k + "m_prev"] = m[r, a], q[r, a], mp[r, a]
        out[k + "run"], out[k + "agent"] = r, a
        out[k + "engine_epochs"] = int(info[r, a, 1])
    print(f"iteration {it}: engine epochs mean over all {R * A} fits {ep.mean():.0f}, over the dumped {N_FITS}: {ep[picks].mean():.0f}, "
          f"over the first 16 of them: {ep[picks[:16]].mean():.0f}", flush=True)
# prev_iter_m == m: drop the copy
for k in [k for k in out if k.endswith("_m_prev")]:
    del out[k]
path = os.path.join(ROOT, "tests", "golden", "bench_fit_inputs.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path) / 1e6, "MB")
