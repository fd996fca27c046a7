"""TEST INFRASTRUCTURE ONLY (build container) -- golden rounds with SEVERAL SLOTS per auction.

The reference's driver pins ``max_slots = 1`` ("Multi-slot is currently not fully supported", src/main.py:36-37), but
``Auction.simulate_opportunity`` (src/Auction.py:30,60-74) and both mechanisms (src/AuctionAllocation.py:19-23,33-35) are
written for ``num_slots`` slots.  This script drives the UNMODIFIED reference with ``max_slots`` = 2 / 3 passed to its own
``instantiate_auction`` (src/main.py:98-109) under oracle/ref_harness.py, with host-drawn slot counts and one click uniform
per slot, and writes ``tests/golden/rounds_*_slots.npz``.   python -m oracle.make_golden_multislot
"""
import json

import numpy as np

from . import auction_oracle as ao
from . import make_golden as mg
from . import ref_harness as rh


def cases():
    O, TS = ao.ALLOC_ORACLE, ao.ALLOC_TS
    TR, GA = ao.BID_TRUTHFUL, ao.BID_GAUSS
    S, F = ao.MECH_SECOND, ao.MECH_FIRST
    return {
        "rounds_sp_slots": dict(kw=dict(seed=31, A=6, n_items=12, D=5, Do=4, P=4, mechanism=S, alloc_kinds=[O] * 6, bidder_kinds=[TR] * 6, T=500), max_slots=3),
        # P - 1 = 2 < max_slots = 3: the slot without a runner-up is dropped by the reference's zip (Auction.py:68)
        "rounds_fp_slots": dict(kw=dict(seed=32, A=5, n_items=8, D=5, Do=4, P=3, mechanism=F, alloc_kinds=[O, TS, O, TS, O], bidder_kinds=[GA, TR, GA, GA, TR], T=400, sigma=0.1),
                                max_slots=3),
        # equal bids across slots (identical catalogs), two slots
        "rounds_sp_slots_ties": dict(kw=dict(seed=33, A=4, n_items=8, D=5, Do=4, P=4, mechanism=S, alloc_kinds=[O] * 4, bidder_kinds=[TR] * 4, T=300, dup_agents=[(1, 0), (2, 0)]),
                                     max_slots=2),
    }


def run_reference(case, noise, cfg, max_slots):
    import torch

    ref = rh.load_reference()
    A = case["A"]
    names = [ac["name"] for ac in cfg["agents"]]
    E = {names[a]: case["E"][a, : case["n_items"][a]].copy() for a in range(A)}
    V = {names[a]: case["V"][a, : case["n_items"][a]].copy() for a in range(A)}
    rng = rh.ReplayRNG(noise["ctx"], noise["parts"], noise["u"], noise.get("gamma_z"), num_slots=noise["num_slots"])
    auction, agents, _ = rh.build_reference_auction(cfg, E, V, rng, ref, max_slots=max_slots)
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
    return rec, met


def main():
    for name, c in cases().items():
        case, noise, cfg = mg.build_case(**c["kw"])
        rng = np.random.default_rng(c["kw"]["seed"] + 1000)
        T = noise["parts"].shape[0]
        noise["num_slots"] = rng.integers(1, c["max_slots"] + 1, size=T).astype(np.int32)
        noise["u"] = rng.random((T, c["max_slots"]))
        case["max_slots"] = c["max_slots"]
        rec, met = run_reference(case, noise, cfg, c["max_slots"])
        # the oracle restatement must agree before the fixture is written
        orec, omet = ao.simulate_rounds(case, noise["ctx"], noise["parts"], noise["u"], noise.get("ts_eps"), noise.get("gamma_z"), None, num_slots=noise["num_slots"])
        for k in ("item", "won", "outcome"):
            assert np.array_equal(orec[k], rec[k]), (name, k)
        for k in ("price", "second", "bid"):
            np.testing.assert_allclose(orec[k], rec[k], rtol=2e-6, atol=1e-15, err_msg=f"{name} {k}")
        d = ao.derived_metrics(omet["acc"])
        for k in ("net", "gross", "overbid_regret", "underbid_regret"):
            np.testing.assert_allclose(d[k], met[k], rtol=2e-6, atol=2e-7, err_msg=f"{name} {k}")  # sums of float32-derived differences
        np.testing.assert_allclose(omet["revenue"], met["revenue"], rtol=1e-9)
        mg.save_case(name, case, noise, rec, met, extra={"cfg_json": json.dumps(cfg)})


if __name__ == "__main__":
    main()
