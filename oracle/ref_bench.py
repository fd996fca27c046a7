"""TEST / BENCH INFRASTRUCTURE ONLY -- times the UNMODIFIED reference (oracle/_ref, see oracle/make_ref.py) and the numpy
port on the same bounded sample of the bench workload, on this box's host cores.

One iteration of one run of the SP_Truthful_TS-shaped workload costs a CPU
    T rounds of Auction.simulate_opportunity (reference src/Auction.py:28-74)
  + A allocator fits, PyTorchLogisticRegressionAllocator.update (src/BidderAllocation.py:29-65),
so one core delivers  T / (T / round_rate + A * seconds_per_fit)  opportunities/s, and independent runs scale over the cores
(one process per core, one torch thread each: the best case for the reference).

SAME WORK AS THE GPU ARM: the fits are not drawn fresh (a fresh allocator, m ~ N(0, 1), q = 1, needs ~8 000 epochs; an allocator
that has learnt for ten iterations needs ~3 000).  They are the fit inputs -- rows, m, q, prev_iter_m -- of iterations
W .. W+K-1 of the bench's own learning trajectory, dumped from the engine by tools/dump_bench_fit_inputs.py into
tests/golden/bench_fit_inputs.npz (global runs 0-1, agents 0-15 of every iteration).  Both JSON lines print
``fit_epochs_mean`` so that equal work can be checked from the outside.
"""
from __future__ import annotations

import json
import os
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
FIT_INPUTS = os.path.join(os.path.dirname(HERE), "tests", "golden", "bench_fit_inputs.npz")


def reference_available():
    from . import make_ref

    return make_ref.ref_src() is not None


def synthetic_config(A, I, D, Do, P, T, learnt=True):
    """config/SP_Truthful_TS.json (or SP_Oracle.json) at the synthetic shape: one agent entry with num_copies = A."""
    alloc = ({"type": "PyTorchLogisticRegressionAllocator", "kwargs": {"embedding_size": Do, "num_items": I}} if learnt
             else {"type": "OracleAllocator", "kwargs": {}})
    return {"random_seed": 0, "num_runs": 1, "num_iter": 1, "rounds_per_iter": T, "num_participants_per_round": P,
            "embedding_size": D, "embedding_var": 1.0, "obs_embedding_size": Do, "allocation": "SecondPrice",
            "agents": [{"name": "Truthful Learnt" if learnt else "Truthful Oracle", "num_copies": A, "num_items": I, "allocator": alloc,
                        "bidder": {"type": "TruthfulBidder", "kwargs": {}}}],
            "output_dir": "results/bench/"}


def load_fit_inputs(iteration, which):
    """Fit inputs number ``which`` (0 .. n-1) of ``iteration`` from the committed dump; iterations beyond the dump use its last."""
    z = np.load(FIT_INPUTS)
    n_it, n_fit = int(z["n_iterations"]), int(z["fits_per_iteration"])
    it = min(int(iteration), n_it - 1)
    k = f"it{it}_f{which % n_fit}_"
    return {"X": z[k + "X"], "items": z[k + "items"].astype(np.int64), "y": z[k + "y"].astype(np.float32), "m0": z[k + "m"], "q0": z[k + "q"],
            "m_prev": z[k + "m"],  # prev_iter_m == m at the start of every fit (update_prior, Models.py:47-48)
            "iteration": it, "run": int(z[k + "run"]), "agent": int(z[k + "agent"])}


def _reference_sample(A, I, D, Do, P, T, n_rounds, fits, learnt, seed):
    """The unmodified reference on one core: n_rounds of simulate_opportunity at the synthetic shape, then the given fits."""
    import contextlib
    import io
    import warnings

    import torch

    torch.set_num_threads(1)
    warnings.filterwarnings("ignore")  # deprecation chatter of the reference's torch idioms on torch 2.x
    from . import ref_harness as rh

    ref = rh.load_reference()
    main, BA = ref["main"], ref["BidderAllocation"]
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "cfg.json")
        cfg = synthetic_config(A, I, D, Do, P, T, learnt)
        cfg["random_seed"] = seed
        with open(path, "w") as f:
            json.dump(cfg, f)
        rng, config, agent_configs, a2i, a2v, num_runs, max_slots, es, ev, oes = main.parse_config(path)   # main.py:24-74
    agents = main.instantiate_agents(rng, agent_configs, a2v, a2i)                                           # main.py:77-95
    auction, _, _, _ = main.instantiate_auction(rng, config, a2i, a2v, agents, max_slots, es, ev, oes)      # main.py:98-109
    for _ in range(20):  # numba JIT of Models.sigmoid, torch warm-up
        auction.simulate_opportunity()
    t0 = time.perf_counter()
    for _ in range(n_rounds):  # main.py:116-117
        auction.simulate_opportunity()
    t_rounds = time.perf_counter() - t0
    fit_s, fit_ep = [], []
    for fi in fits:
        alloc = BA.PyTorchLogisticRegressionAllocator(rng, embedding_size=Do, num_items=I)
        rm = alloc.response_model
        with torch.no_grad():
            rm.m.copy_(torch.from_numpy(np.asarray(fi["m0"], np.float32)))
        rm.q = torch.from_numpy(np.array(fi["q0"], np.float32))
        rm.prev_iter_m = torch.from_numpy(np.array(fi["m_prev"], np.float32))
        steps = {"n": 0}
        orig_step = torch.optim.Adam.step

        def counting_step(self, *a, _o=orig_step, **k):
            steps["n"] += 1
            return _o(self, *a, **k)

        torch.optim.Adam.step = counting_step
        try:
            with contextlib.redirect_stdout(io.StringIO()):  # "Stopping at Epoch ..." (BidderAllocation.py:54)
                t0 = time.perf_counter()
                alloc.update(np.asarray(fi["X"], np.float64), np.asarray(fi["items"]), np.asarray(fi["y"], np.float64), 0, False, None, None, "bench")
                fit_s.append(time.perf_counter() - t0)
        finally:
            torch.optim.Adam.step = orig_step
        fit_ep.append(steps["n"])
    return t_rounds, fit_s, fit_ep


def _port_sample(A, I, D, Do, P, T, n_rounds, fits, learnt, seed):
    """The numpy port (oracle/auction_oracle.py, oracle/fit_oracle.py) on the same sample."""
    from . import auction_oracle as ao
    from . import cpu_bench
    from . import fit_oracle as fo

    case = cpu_bench.make_case(A, I, D, Do, P, seed)
    if not learnt:
        case["alloc_kind"][:] = ao.ALLOC_ORACLE
    nz = ao.draw_replay_inputs(np.random.default_rng(seed + 1), n_rounds, A, P, D, I, Do, 1.0, want_eps=learnt)
    t0 = time.perf_counter()
    ao.simulate_rounds_scalar(case, nz["ctx"], nz["parts"], nz["u"], nz.get("ts_eps"))
    t_rounds = time.perf_counter() - t0
    fit_s, fit_ep = [], []
    for fi in fits:
        t0 = time.perf_counter()
        r = fo.fit_allocator(fi["X"], fi["items"], fi["y"], fi["m0"], fi["q0"], fi["m_prev"])
        fit_s.append(time.perf_counter() - t0)
        fit_ep.append(r["n_epochs"])
    return t_rounds, fit_s, fit_ep


def sample(job):
    """One bounded sample on one core.  job = dict(kind, A, I, D, Do, P, T, n_rounds, iteration, fit_ids, learnt, seed)."""
    j = job
    fits = [load_fit_inputs(j["iteration"], w) for w in j["fit_ids"]] if j["learnt"] else []
    fn = _reference_sample if j["kind"] == "reference" else _port_sample
    t_rounds, fit_s, fit_ep = fn(j["A"], j["I"], j["D"], j["Do"], j["P"], j["T"], j["n_rounds"], fits, j["learnt"], j["seed"])
    round_rate = j["n_rounds"] / t_rounds
    fit_mean = float(np.mean(fit_s)) if fit_s else 0.0
    return {"round_rate": round_rate, "fit_seconds": fit_s, "fit_epochs": fit_ep, "fit_rows": [len(f["y"]) for f in fits],
            "opp_per_s": j["T"] / (j["T"] / round_rate + j["A"] * fit_mean), "cpu_seconds": t_rounds + sum(fit_s)}


def run(kind, A, I, D, Do, P, T, iteration, n_rounds, fits_per_worker, workers, learnt=True, seed=0, pool=None):
    """``workers`` independent samples in parallel processes, each timing its own fits of ``iteration``."""
    jobs = [dict(kind=kind, A=A, I=I, D=D, Do=Do, P=P, T=T, n_rounds=n_rounds, iteration=iteration, learnt=learnt, seed=seed + 17 * w,
                 fit_ids=list(range(w * fits_per_worker, (w + 1) * fits_per_worker))) for w in range(workers)]
    t0 = time.perf_counter()
    res = [sample(jobs[0])] if (workers == 1 and pool is None) else list(pool.map(sample, jobs))
    wall = time.perf_counter() - t0
    ep = [e for r in res for e in r["fit_epochs"]]
    fs = [s for r in res for s in r["fit_seconds"]]
    per_core = float(np.mean([r["opp_per_s"] for r in res]))
    return {"kind": kind, "workers": workers, "wall_seconds": wall, "per_core_opp_per_s": per_core, "aggregate_opp_per_s": per_core * workers,
            "round_only_per_core": float(np.mean([r["round_rate"] for r in res])), "fit_seconds_mean": float(np.mean(fs)) if fs else 0.0,
            "fit_epochs_mean": float(np.mean(ep)) if ep else 0.0, "fits_timed": len(ep), "iteration": iteration}


def make_pool(workers):
    import multiprocessing as mp
    from concurrent.futures import ProcessPoolExecutor

    return ProcessPoolExecutor(max_workers=workers, mp_context=mp.get_context("spawn"))
