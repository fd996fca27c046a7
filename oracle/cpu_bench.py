"""TEST / BENCH INFRASTRUCTURE ONLY -- times the CPU port (oracle) of the hot path on a bounded sample.

Used by bench.py for the ``cpu_baseline`` object and for ``--impl reference`` (the reference itself is
pure Python under /root/reference and cannot travel to the GPU box, so the port in
oracle/auction_oracle.py + oracle/fit_oracle.py -- pinned against the reference's outputs in
tests/golden -- is what gets timed; ``kind`` is therefore "port").

One "iteration of one run" of the SP_Truthful_TS-shaped workload costs the CPU
    T rounds of simulate_opportunity  +  A allocator fits (Agent.update),
so the opportunities/s of one core is  T / (T / round_rate + A * seconds_per_fit).
"""
from __future__ import annotations

import time

import numpy as np

from . import auction_oracle as ao
from . import fit_oracle as fo


def make_case(A, I, D, Do, P, seed=0):
    rng = np.random.default_rng(seed)
    E, V = ao.make_catalog(rng, A, I, D)
    return {"A": A, "I": I, "D": D, "Do": Do, "P": P, "mechanism": ao.MECH_SECOND, "embedding_var": 1.0,
            "n_items": np.full(A, I, np.int32), "E": E, "V": V,
            "m": rng.standard_normal((A, I, Do + 1)).astype(np.float32), "q": np.ones((A, I, Do + 1), np.float32),
            "alloc_kind": np.full(A, ao.ALLOC_TS, np.int32), "bidder_kind": np.full(A, ao.BID_TRUTHFUL, np.int32),
            "bidder_f": np.zeros((A, 4))}


def sample(args):
    """One bounded sample on one core. args = (A, I, D, Do, P, T, n_rounds, n_fits, seed)."""
    A, I, D, Do, P, T, n_rounds, n_fits, seed = args
    case = make_case(A, I, D, Do, P, seed)
    rng = np.random.default_rng(seed + 1)
    nz = ao.draw_replay_inputs(rng, n_rounds, A, P, D, I, Do, 1.0, want_eps=True)
    t0 = time.perf_counter()
    ao.simulate_rounds_scalar(case, nz["ctx"], nz["parts"], nz["u"], nz["ts_eps"])
    t_rounds = time.perf_counter() - t0
    # rows for the fits: one full iteration of T rounds through the vectorised oracle (not timed)
    nz = ao.draw_replay_inputs(rng, T, A, P, D, I, Do, 1.0, want_eps=True)
    rec, _ = ao.simulate_rounds(case, nz["ctx"], nz["parts"], nz["u"], nz["ts_eps"])
    obs = np.concatenate([nz["ctx"][:, :Do], np.ones((T, 1))], axis=1)
    fit_s, fit_epochs = [], []
    for a in range(n_fits):
        won = (nz["parts"] == a) & (rec["won"] == 1)
        t_idx, s_idx = np.nonzero(won)
        t0 = time.perf_counter()
        r = fo.fit_allocator(obs[t_idx], rec["item"][t_idx, s_idx], rec["outcome"][t_idx, s_idx], case["m"][a], case["q"][a], case["m"][a])
        fit_s.append(time.perf_counter() - t0)
        fit_epochs.append(r["n_epochs"])
    round_rate = n_rounds / t_rounds
    fit_mean = float(np.mean(fit_s)) if fit_s else 0.0
    iter_seconds = T / round_rate + A * fit_mean
    return {"round_rate": round_rate, "fit_seconds": fit_mean, "fit_epochs": float(np.mean(fit_epochs)) if fit_epochs else 0.0,
            "opp_per_s": T / iter_seconds, "round_only_opp_per_s": round_rate, "cpu_seconds": t_rounds + sum(fit_s)}


def run(A, I, D, Do, P, T, n_rounds, n_fits, workers=1, seed=0):
    """`workers` independent samples in parallel processes (runs are independent: one process per core is the
    best case for the reference).  Returns the aggregate and the per-core mean."""
    jobs = [(A, I, D, Do, P, T, n_rounds, n_fits, seed + 17 * w) for w in range(workers)]
    t0 = time.perf_counter()
    if workers == 1:
        res = [sample(jobs[0])]
    else:
        import multiprocessing as mp
        from concurrent.futures import ProcessPoolExecutor

        with ProcessPoolExecutor(max_workers=workers, mp_context=mp.get_context("spawn")) as ex:
            res = list(ex.map(sample, jobs))
    wall = time.perf_counter() - t0
    per_core = float(np.mean([r["opp_per_s"] for r in res]))
    return {"workers": workers, "wall_seconds": wall, "per_core_opp_per_s": per_core, "aggregate_opp_per_s": per_core * workers,
            "round_only_per_core": float(np.mean([r["round_only_opp_per_s"] for r in res])),
            "fit_seconds": float(np.mean([r["fit_seconds"] for r in res])), "fit_epochs": float(np.mean([r["fit_epochs"] for r in res]))}
