"""Writes tests/golden/bench_fit_inputs.npz (run on a GPU box: `python tools/dump_bench_fit_inputs.py`).

The fit inputs -- won rows, m, q (prev_iter_m == m at the start of every fit) -- of global runs 0-1, agents 0-15, for
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
RUNS, AGENTS, R = 2, 16, 8
w = bench.WORKLOAD
A, I, D, Do, P, T = w["A"], w["I"], w["D"], w["Do"], w["P"], w["T"]
eng = bench.make_engine(ag, _lib, R=R, T=T, learnt=True, device=0, run_offset=0)
eng.set_allocator_state(bench.initial_m(0, R))
out = {"n_iterations": N_IT, "fits_per_iteration": RUNS * AGENTS, "seed": bench.SEED, "T": T}
for it in range(N_IT):
    eng.clear_iteration()
    eng.simulate(bench.SEED, it, T)
    meta = eng.fit_meta[:RUNS, :T].cpu().numpy().view(np.uint32)
    ctx = eng.fit_ctx[:RUNS, :T].cpu().numpy()
    m, q, mp = (t[:RUNS].cpu().numpy().copy() for t in (eng.m, eng.q, eng.m_prev))
    assert np.array_equal(m, mp)
    info = eng.update_allocators(want_info=True).cpu().numpy()
    for r in range(RUNS):
        valid, agent, item, click = meta[r] >> 31, (meta[r] >> 12) & 0xFFF, meta[r] & 0xFFF, (meta[r] >> 30) & 1
        for a in range(AGENTS):
            sel = (valid == 1) & (agent == a)
            k = f"it{it}_f{r * AGENTS + a}_"
            out[k + "X"] = np.concatenate([ctx[r][sel], np.ones((sel.sum(), 1), np.float32)], axis=1)
            out[k + "items"] = item[sel].astype(np.uint8)
            out[k + "y"] = click[sel].astype(np.uint8)
            out[k + "m"], out[k + "q"], out[k + "m_prev"] = m[r, a], q[r, a], mp[r, a]
            out[k + "run"], out[k + "agent"] = r, a
            out[k + "engine_epochs"] = int(info[r, a, 1])
    ep = info[..., 1]
    print(f"iteration {it}: engine epochs mean over all {R * A} fits {ep.mean():.0f}, over the dumped {RUNS * AGENTS}: {ep[:RUNS, :AGENTS].mean():.0f}", flush=True)
# prev_iter_m == m: drop the copy
for k in [k for k in out if k.endswith("_m_prev")]:
    del out[k]
path = os.path.join(ROOT, "tests", "golden", "bench_fit_inputs.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path) / 1e6, "MB")
