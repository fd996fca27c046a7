"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/retention.npz by running the UNMODIFIED reference with
``Agent(memory=...)`` over several iterations (src/Agent.py:124-129, src/main.py:112-155).

    python -m oracle.make_golden_retention

Model state is held fixed across iterations (``allocator.update`` / ``bidder.update`` are replaced by recorders), so the
fixture pins exactly what retention changes: which records ``agent.logs`` holds at the end of every iteration, what
the metric getters return from them, and which rows ``Agent.update`` hands to the allocator (won rows) and to the bidder
(all rows) -- src/Agent.py:79-118.
"""
import json

import numpy as np

from . import auction_oracle as ao
from . import make_golden as mg
from . import ref_harness as rh

N_ITER, T_ITER = 4, 200
MEMORY = [50, 300, 0, 130, 30]


def main():
    import torch

    O, TS, MAP = ao.ALLOC_ORACLE, ao.ALLOC_TS, ao.ALLOC_MAP
    TR, GA, GC = ao.BID_TRUTHFUL, ao.BID_GAUSS, ao.BID_GAUSS_CLIP
    case, noise, cfg = mg.build_case(seed=41, A=5, n_items=8, D=5, Do=4, P=3, mechanism=ao.MECH_FIRST, alloc_kinds=[TS, O, TS, O, MAP],
                                     bidder_kinds=[TR, GA, GA, TR, GC], T=N_ITER * T_ITER, sigma=0.15, q_spread=True)
    for ac, mem in zip(cfg["agents"], MEMORY):
        if mem:
            ac["memory"] = mem
    ref = rh.load_reference()
    A = case["A"]
    names = [ac["name"] for ac in cfg["agents"]]
    E = {names[a]: case["E"][a, : case["n_items"][a]].copy() for a in range(A)}
    V = {names[a]: case["V"][a, : case["n_items"][a]].copy() for a in range(A)}
    sl = lambda i: slice(i * T_ITER, (i + 1) * T_ITER)  # noqa: E731
    rng = rh.ReplayRNG(noise["ctx"][sl(0)], noise["parts"][sl(0)], noise["u"][sl(0)], noise["gamma_z"][sl(0)])
    auction, agents, _ = rh.build_reference_auction(cfg, E, V, rng, ref)
    assert [ag.memory for ag in agents] == MEMORY
    for a, ag in enumerate(agents):
        if case["alloc_kind"][a] != O:
            rm = ag.allocator.response_model
            with torch.no_grad():
                rm.m.copy_(torch.from_numpy(case["m"][a].copy()))
            rm.prev_iter_m = rm.m.detach().clone()
            rm.q = torch.from_numpy(case["q"][a].copy())
    rh.wrap_bid_slots(agents, rng)
    seen = {}
    for a, ag in enumerate(agents):  # record what Agent.update passes on instead of fitting (state stays fixed)
        ag.allocator.update = lambda contexts, items, outcomes, *r, _a=a: seen.__setitem__(("alloc", _a), (contexts, items, outcomes))
        ag.bidder.update = lambda contexts, values, bids, prices, outcomes, ests, won, *r, _a=a: seen.__setitem__(
            ("bid", _a), (values, bids, prices, outcomes, ests, won))
    out = {"memory": np.asarray(MEMORY, np.int32), "n_iter": N_ITER, "t_iter": T_ITER, "cfg_json": json.dumps(cfg)}
    for k, v in case.items():
        out["case_" + k] = np.asarray(v)
    for k, v in noise.items():
        out["in_" + k] = v
    for it in range(N_ITER):
        rng.ctx, rng.parts, rng.u, rng.gamma_z = noise["ctx"][sl(it)], noise["parts"][sl(it)], noise["u"][sl(it)], noise["gamma_z"][sl(it)]
        rng.t = -1
        rec = rh.run_reference_rounds(auction, agents, rng, T_ITER, noise["ts_eps"][sl(it)])
        for ag in agents:  # src/main.py:128-129
            ag.update(iteration=it)
        met = rh.reference_metrics(auction, agents)
        for k, v in rec.items():
            out[f"it{it}_ref_{k}"] = v
        for k, v in met.items():
            out[f"it{it}_met_{k}"] = np.asarray(v)
        for a, ag in enumerate(agents):
            p = f"it{it}_a{a}_"
            cx, items, y = seen[("alloc", a)]
            out[p + "fit_ctx"], out[p + "fit_items"], out[p + "fit_y"] = np.asarray(cx, np.float64).reshape(len(items), -1), np.asarray(items), np.asarray(y)
            values, bids, prices, outcomes, ests, won = seen[("bid", a)]
            out[p + "values"], out[p + "bids"], out[p + "prices"] = np.asarray(values), np.asarray(bids), np.asarray(prices)
            out[p + "outcomes"], out[p + "ests"], out[p + "won"] = np.asarray(outcomes), np.asarray(ests, np.float64), np.asarray(won)
            out[p + "gammas"] = np.asarray(getattr(ag.bidder, "gammas", []), np.float64)
            out[p + "propensities"] = np.asarray(getattr(ag.bidder, "propensities", []), np.float64)
            out[p + "mean_gamma"] = np.float64(np.mean(ag.bidder.gammas)) if len(getattr(ag.bidder, "gammas", [])) else np.float64("nan")
        for ag in agents:  # src/main.py:151-155
            ag.clear_utility()
            ag.clear_logs()
        auction.clear_revenue()
        for a, ag in enumerate(agents):
            out[f"it{it}_a{a}_kept"] = len(ag.logs)
    path = mg.GOLDEN_DIR + "/retention.npz"
    np.savez_compressed(path, **out)
    import os
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB); kept per iteration:",
          [[int(out[f'it{it}_a{a}_kept']) for a in range(A)] for it in range(N_ITER)])


if __name__ == "__main__":
    main()
