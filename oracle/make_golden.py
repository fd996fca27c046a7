"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container (the reference is mounted at /root/reference):

    python -m oracle.make_golden            # all cases
    python -m oracle.make_golden rounds     # only the round-loop cases
    python -m oracle.make_golden fit        # only the allocator-fit cases (slow: ~1 min)

Every fixture stores the replay inputs (host-drawn contexts, participants, noise, uniforms, the
catalog and the learnt state) next to what the reference produced from them, so the GPU box --
which has no reference tree -- can check both the oracle restatement and the CUDA path.
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import re
import sys

import numpy as np

from . import auction_oracle as ao
from . import ref_harness as rh

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

ALLOC_NAMES = {ao.ALLOC_ORACLE: "OracleAllocator", ao.ALLOC_TS: "PyTorchLogisticRegressionAllocator",
               ao.ALLOC_MAP: "PyTorchLogisticRegressionAllocator"}


def agent_cfg(name, n_items, alloc_kind, bidder, Do):
    if alloc_kind == ao.ALLOC_ORACLE:
        alloc = {"type": "OracleAllocator", "kwargs": {}}
    else:
        kw = {"embedding_size": Do, "num_items": n_items}
        if alloc_kind == ao.ALLOC_MAP:
            kw["thompson_sampling"] = False
        alloc = {"type": "PyTorchLogisticRegressionAllocator", "kwargs": kw}
    return {"name": name, "num_items": n_items, "allocator": alloc, "bidder": bidder}


def bidder_cfg(kind, prev_gamma=1.0, sigma=0.02):
    if kind == ao.BID_TRUTHFUL:
        return {"type": "TruthfulBidder", "kwargs": {}}
    if kind == ao.BID_GAUSS:
        return {"type": "ValueLearningBidder", "kwargs": {"gamma_sigma": sigma, "init_gamma": prev_gamma, "inference": "\"search\""}}
    if kind == ao.BID_GAUSS_CLIP:
        return {"type": "EmpiricalShadedBidder", "kwargs": {"gamma_sigma": sigma, "init_gamma": prev_gamma}}
    raise ValueError(kind)


def build_case(seed, A, n_items, D, Do, P, mechanism, alloc_kinds, bidder_kinds, T,
               embedding_var=1.0, prev_gamma=1.0, sigma=0.02, q_spread=False, dup_agents=None,
               bidder_variants=None):
    """Draw a catalog + learnt state + replay noise; returns (case dict for the oracle, cfg for the reference)."""
    rng = np.random.default_rng(seed)
    n_items = np.asarray(n_items if np.ndim(n_items) else [n_items] * A, np.int32)
    I = int(n_items.max())
    E, V = ao.make_catalog(rng, A, I, D, embedding_var)
    if dup_agents:
        for dst, src in dup_agents:
            E[dst], V[dst] = E[src], V[src]
    m = rng.standard_normal((A, I, Do + 1)).astype(np.float32)
    q = np.ones((A, I, Do + 1), np.float32)
    if q_spread:
        q = (1.0 + 30.0 * rng.random((A, I, Do + 1)) ** 3).astype(np.float32)
    bidder_f = np.zeros((A, 4), np.float64)
    bidder_f[:, 0], bidder_f[:, 1] = prev_gamma, sigma
    if bidder_variants is not None:
        bidder_f[:, :2] = np.asarray(bidder_variants, np.float64)
    case = {
        "A": A, "I": I, "D": D, "Do": Do, "P": P, "mechanism": mechanism, "embedding_var": embedding_var,
        "n_items": n_items, "E": E, "V": V, "m": m, "q": q,
        "alloc_kind": np.asarray(alloc_kinds, np.int32), "bidder_kind": np.asarray(bidder_kinds, np.int32),
        "bidder_f": bidder_f,
    }
    noise = ao.draw_replay_inputs(rng, T, A, P, D, I, Do, embedding_var,
                                  want_eps=any(k == ao.ALLOC_TS for k in alloc_kinds),
                                  want_gamma=any(k != ao.BID_TRUTHFUL for k in bidder_kinds))
    cfg = {
        "num_participants_per_round": P, "embedding_size": D, "embedding_var": embedding_var,
        "obs_embedding_size": Do, "allocation": "FirstPrice" if mechanism == ao.MECH_FIRST else "SecondPrice",
        "num_iter": 1, "rounds_per_iter": T, "output_dir": "/tmp/agym_golden/",
        "agents": [agent_cfg(f"agent {a}", int(n_items[a]), int(alloc_kinds[a]),
                             bidder_cfg(int(bidder_kinds[a]), float(bidder_f[a, 0]), float(bidder_f[a, 1])), Do)
                   for a in range(A)],
    }
    return case, noise, cfg


def run_reference(case, noise, cfg):
    """Drive the unmodified reference with the replay noise; returns (rec, metrics)."""
    import torch

    ref = rh.load_reference()
    A = case["A"]
    names = [ac["name"] for ac in cfg["agents"]]
    E = {names[a]: case["E"][a, : case["n_items"][a]].copy() for a in range(A)}
    V = {names[a]: case["V"][a, : case["n_items"][a]].copy() for a in range(A)}
    rng = rh.ReplayRNG(noise["ctx"], noise["parts"], noise["u"], noise.get("gamma_z"))
    auction, agents, _ = rh.build_reference_auction(cfg, E, V, rng, ref)
    for a, ag in enumerate(agents):
        if case["alloc_kind"][a] != ao.ALLOC_ORACLE:
            nI = int(case["n_items"][a])
            rm = ag.allocator.response_model
            with torch.no_grad():
                rm.m.copy_(torch.from_numpy(case["m"][a, :nI].copy()))
            rm.prev_iter_m = rm.m.detach().clone()
            rm.q = torch.from_numpy(case["q"][a, :nI].copy())
    rh.wrap_bid_slots(agents, rng)
    T = noise["parts"].shape[0]
    rec = rh.run_reference_rounds(auction, agents, rng, T, noise.get("ts_eps"))
    met = rh.reference_metrics(auction, agents)
    return rec, met, auction, agents


def save_case(name, case, noise, rec, met, extra=None):
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    out = {}
    for k, v in case.items():
        out["case_" + k] = np.asarray(v)
    for k, v in noise.items():
        out["in_" + k] = v
    for k, v in rec.items():
        out["ref_" + k] = v
    for k, v in met.items():
        out["met_" + k] = np.asarray(v)
    for k, v in (extra or {}).items():
        out[k] = np.asarray(v)
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


def round_cases():
    O, TS, MAP = ao.ALLOC_ORACLE, ao.ALLOC_TS, ao.ALLOC_MAP
    TR, GA, GC = ao.BID_TRUTHFUL, ao.BID_GAUSS, ao.BID_GAUSS_CLIP
    S, F = ao.MECH_SECOND, ao.MECH_FIRST
    return {
        # config/SP_Oracle.json shape
        "rounds_sp_oracle": dict(seed=11, A=6, n_items=12, D=5, Do=4, P=2, mechanism=S, alloc_kinds=[O] * 6, bidder_kinds=[TR] * 6, T=1500),
        # config/SP_Truthful_TS.json shape (pre-fit state and a spread-out posterior)
        "rounds_sp_ts": dict(seed=12, A=6, n_items=12, D=5, Do=4, P=2, mechanism=S, alloc_kinds=[TS] * 6, bidder_kinds=[TR] * 6, T=600),
        "rounds_sp_ts_q": dict(seed=13, A=6, n_items=12, D=5, Do=4, P=2, mechanism=S, alloc_kinds=[TS] * 6, bidder_kinds=[TR] * 6, T=400, q_spread=True),
        # config/FP_DM_Oracle.json shape before the first bidder fit (Gaussian gamma)
        "rounds_fp_gauss": dict(seed=14, A=6, n_items=12, D=5, Do=4, P=2, mechanism=F, alloc_kinds=[O] * 6, bidder_kinds=[GA] * 6, T=1500),
        # config/FP_DR_TS.json / FP_IPS_TS.json shape before the first fit
        "rounds_fp_ts_gauss": dict(seed=15, A=3, n_items=12, D=5, Do=4, P=2, mechanism=F, alloc_kinds=[TS] * 3, bidder_kinds=[GA] * 3, T=400, sigma=0.3),
        # participant-count edge cases
        "rounds_sp_p3": dict(seed=16, A=6, n_items=12, D=5, Do=4, P=3, mechanism=S, alloc_kinds=[O] * 6, bidder_kinds=[TR] * 6, T=600),
        "rounds_fp_pA": dict(seed=17, A=6, n_items=12, D=5, Do=4, P=6, mechanism=F, alloc_kinds=[O, TS, O, MAP, TS, O], bidder_kinds=[GA, TR, GA, GA, TR, TR], T=300, sigma=0.1),
        "rounds_sp_p1": dict(seed=18, A=4, n_items=6, D=5, Do=4, P=1, mechanism=S, alloc_kinds=[O] * 4, bidder_kinds=[TR] * 4, T=200),
        "rounds_fp_p1": dict(seed=19, A=4, n_items=6, D=5, Do=4, P=1, mechanism=F, alloc_kinds=[O] * 4, bidder_kinds=[GA] * 4, T=200),
        # exact ties: identical catalogs for agents 0/1/2 (equal truthful bids) and gamma clipped to 0 (all bids 0)
        "rounds_sp_ties": dict(seed=20, A=4, n_items=8, D=5, Do=4, P=3, mechanism=S, alloc_kinds=[O] * 4, bidder_kinds=[TR] * 4, T=400, dup_agents=[(1, 0), (2, 0)]),
        "rounds_fp_ties": dict(seed=21, A=4, n_items=8, D=5, Do=4, P=3, mechanism=F, alloc_kinds=[O] * 4, bidder_kinds=[GC] * 4, T=300,
                               bidder_variants=[(-1.0, 0.02), (-1.0, 0.02), (0.5, 0.4), (-1.0, 0.02)]),
        # ragged catalogs, mixed allocators, other embedding sizes
        "rounds_sp_ragged": dict(seed=22, A=5, n_items=[3, 12, 1, 7, 33], D=7, Do=3, P=2, mechanism=S, alloc_kinds=[O, TS, TS, MAP, TS], bidder_kinds=[TR] * 5, T=500, embedding_var=0.7, q_spread=True),
        # synthetic scale-out shape (BASELINE.json configs[4]): 64 agents x 64 items
        "rounds_sp_ts_64x64": dict(seed=23, A=64, n_items=64, D=5, Do=4, P=2, mechanism=S, alloc_kinds=[TS] * 64, bidder_kinds=[TR] * 64, T=48, q_spread=True),
        "rounds_sp_oracle_64x64": dict(seed=24, A=64, n_items=64, D=5, Do=4, P=2, mechanism=S, alloc_kinds=[O] * 64, bidder_kinds=[TR] * 64, T=400),
    }


def make_round_goldens():
    for name, kw in round_cases().items():
        case, noise, cfg = build_case(**kw)
        rec, met, _, _ = run_reference(case, noise, cfg)
        save_case(name, case, noise, rec, met, extra={"cfg_json": json.dumps(cfg)})


def make_fit_goldens():
    """allocator.update (BidderAllocation.py:29-65) on rows the reference itself logged."""
    import torch

    TS, TR = ao.ALLOC_TS, ao.BID_TRUTHFUL
    for name, kw in {
        # reference shape: ~T/A*... won rows per agent
        "fit_ref_shape": dict(seed=31, A=6, n_items=12, D=5, Do=4, P=2, mechanism=ao.MECH_SECOND, alloc_kinds=[TS] * 6, bidder_kinds=[TR] * 6, T=6000),
        # synthetic shape: few rows per agent, many items
        "fit_64x64": dict(seed=32, A=64, n_items=64, D=5, Do=4, P=2, mechanism=ao.MECH_SECOND, alloc_kinds=[TS] * 64, bidder_kinds=[TR] * 64, T=5000),
    }.items():
        case, noise, cfg = build_case(**kw)
        # the round loop of this case is not stored (eps would be large): run it without TS noise patch
        noise["ts_eps"] = np.random.default_rng(kw["seed"] + 1000).standard_normal(
            (kw["T"], kw["P"], case["I"], case["Do"] + 1)).astype(np.float32)
        rec, met, auction, agents = run_reference(case, noise, cfg)
        out = {}
        fit_agents = [0, 1, 2] if kw["A"] == 6 else [0, 5, 9, 33]
        for it in range(2):  # two consecutive iterations: the second one sees the Laplace prior
            for a in fit_agents:
                ag = agents[a]
                rm = ag.allocator.response_model
                won = np.array([o.won for o in ag.logs], bool)
                X = np.array([o.context for o in ag.logs])[won]
                items = np.array([o.item for o in ag.logs])[won]
                y = np.array([o.outcome for o in ag.logs])[won].astype(np.float32)
                m0 = rm.m.detach().numpy().copy()
                q0 = rm.q.numpy().copy()
                mp = rm.prev_iter_m.numpy().copy()
                losses = []
                orig_loss = rm.loss

                def rec_loss(pred, lab, _o=orig_loss, _l=losses):
                    v = _o(pred, lab)
                    _l.append(float(v.item()))
                    return v

                rm.loss = rec_loss
                buf = io.StringIO()
                torch.manual_seed(0)
                with contextlib.redirect_stdout(buf):
                    ag.update(iteration=it)
                rm.loss = orig_loss
                mt = re.search(r"Stopping at Epoch (\d+)", buf.getvalue())
                stop = int(mt.group(1)) if mt else -1
                pre = f"it{it}_a{a}_"
                out[pre + "X"], out[pre + "items"], out[pre + "y"] = X.astype(np.float32), items.astype(np.int32), y
                out[pre + "m0"], out[pre + "q0"], out[pre + "m_prev"] = m0, q0, mp
                out[pre + "m1"] = rm.m.detach().numpy().copy()
                out[pre + "q1"] = rm.q.numpy().copy()
                out[pre + "stop_epoch"] = stop
                out[pre + "n_epochs"] = len(losses)
                out[pre + "losses_head"] = np.asarray(losses[:64])
                out[pre + "losses_tail"] = np.asarray(losses[-128:])
                print(f"{name} it{it} agent {a}: rows {len(y)}, stop epoch {stop}, final loss {losses[-1]:.6f}")
            if it == 0:
                # second iteration: new rounds with the fitted state, fresh noise
                for ag in agents:
                    ag.clear_utility()
                    ag.clear_logs()
                rng2 = np.random.default_rng(kw["seed"] + 2000)
                noise2 = ao.draw_replay_inputs(rng2, kw["T"], kw["A"], kw["P"], kw["D"], case["I"], case["Do"], 1.0, want_eps=True)
                rr = rh.ReplayRNG(noise2["ctx"], noise2["parts"], noise2["u"])
                auction.rng = rr
                for ag in agents:
                    ag.bid = type(ag).bid.__get__(ag)
                rh.wrap_bid_slots(agents, rr)
                rh.run_reference_rounds(auction, agents, rr, kw["T"], noise2["ts_eps"])
        out["fit_agents"] = np.asarray(fit_agents)
        os.makedirs(GOLDEN_DIR, exist_ok=True)
        path = os.path.join(GOLDEN_DIR, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("all", "rounds"):
        make_round_goldens()
    if what in ("all", "fit"):
        make_fit_goldens()


def make_bidder_goldens():
    """ValueLearningBidder('search'): the win-rate fit (Bidder.py:210-260) on rows the reference logged, then a second
    iteration whose bids come from the 128-point grid search (Bidder.py:180-196)."""
    import torch

    O, GA = ao.ALLOC_ORACLE, ao.BID_GAUSS
    kw = dict(seed=41, A=6, n_items=12, D=5, Do=4, P=2, mechanism=ao.MECH_FIRST, alloc_kinds=[O] * 6, bidder_kinds=[GA] * 6, T=3000)
    case, noise, cfg = build_case(**kw)
    torch.manual_seed(7)  # PyTorchWinRateEstimator init (Models.py:55-58) comes from torch's global generator
    rec, met, auction, agents = run_reference(case, noise, cfg)
    out = {}
    fit_agents = [0, 1, 2, 3, 4, 5]
    w_after = np.zeros((6, 4), np.float32)
    for a in fit_agents:
        ag = agents[a]
        lin = ag.bidder.winrate_model.model[0]
        w0 = np.concatenate([lin.weight.detach().numpy().ravel(), lin.bias.detach().numpy()]).astype(np.float32)
        won = np.array([o.won for o in ag.logs], bool)
        est = np.array([o.estimated_CTR for o in ag.logs])
        val = np.array([o.value for o in ag.logs])
        gam = np.array(ag.bidder.gammas)
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            ag.update(iteration=0)
        mt = re.search(r"Stopping at Epoch (\d+)", buf.getvalue())
        w1 = np.concatenate([lin.weight.detach().numpy().ravel(), lin.bias.detach().numpy()]).astype(np.float32)
        w_after[a] = w1
        pre = f"a{a}_"
        out[pre + "est"], out[pre + "value"], out[pre + "gamma"], out[pre + "won"] = est, val, gam, won
        out[pre + "w0"], out[pre + "w1"] = w0, w1
        out[pre + "stop_epoch"] = int(mt.group(1)) if mt else -1
        print(f"bidder fit agent {a}: rows {len(won)}, wins {won.sum()}, stop epoch {out[pre + 'stop_epoch']}, w1 {w1}")
        assert ag.bidder.model_initialised
    out["fit_agents"] = np.asarray(fit_agents)
    path = os.path.join(GOLDEN_DIR, "bidfit_winrate.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")
    # second iteration: post-fit bids by grid search, replayed
    for ag in agents:
        ag.clear_utility()
        ag.clear_logs()
    auction.clear_revenue()
    T2 = 400
    rng2 = np.random.default_rng(4242)
    noise2 = ao.draw_replay_inputs(rng2, T2, 6, 2, 5, grid=128)
    rr = rh.ReplayRNG(noise2["ctx"], noise2["parts"], noise2["u"], None, noise2["grid_u"])
    auction.rng = rr
    for ag in agents:
        ag.bid = type(ag).bid.__get__(ag)
        ag.bidder.rng = rr
    rh.wrap_bid_slots(agents, rr)
    rec2 = rh.run_reference_rounds(auction, agents, rr, T2, None)
    met2 = rh.reference_metrics(auction, agents)
    case2 = dict(case)
    case2["bidder_kind"] = np.full(6, ao.BID_SEARCH, np.int32)
    case2["winrate_w"] = w_after
    save_case("rounds_fp_search", case2, noise2, rec2, met2)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "bidder":
    make_bidder_goldens()
