"""Helpers shared by the -m gpu parity tests: build an Engine from an oracle `case` dict."""
import numpy as np

import auction_gym_b200 as ag
from auction_gym_b200 import _lib
from oracle import auction_oracle as ao

# oracle kind -> ABI kind (the numbering is the same by construction; keep the mapping explicit)
ALLOC = {ao.ALLOC_ORACLE: _lib.ALLOC_ORACLE, ao.ALLOC_TS: _lib.ALLOC_TS, ao.ALLOC_MAP: _lib.ALLOC_MAP}
BID = {ao.BID_TRUTHFUL: _lib.BID_TRUTHFUL, ao.BID_GAUSS: _lib.BID_GAUSS, ao.BID_GAUSS_CLIP: _lib.BID_GAUSS_CLIP,
       ao.BID_SEARCH: _lib.BID_SEARCH, ao.BID_BANDIT: _lib.BID_BANDIT, ao.BID_POLICY: _lib.BID_POLICY}


def engine_from_case(case, R=1, precision=_lib.FP64, run_offset=0, rounds_capacity=0):
    A, I = int(case["A"]), int(case["I"])
    eng = ag.Engine(R=R, A=A, I=I, D=int(case["D"]), Do=int(case["Do"]), P=int(case["P"]), mechanism=int(case["mechanism"]),
                    E=case["E"], V=case["V"], n_items=case["n_items"],
                    alloc_kind=[ALLOC[int(k)] for k in case["alloc_kind"]],
                    bidder_kind=[BID[int(k)] for k in case["bidder_kind"]],
                    embedding_var=float(case["embedding_var"]), precision=precision, run_offset=run_offset,
                    rounds_capacity=rounds_capacity)
    if eng.any_learnt:
        m = np.broadcast_to(case["m"], (R,) + case["m"].shape)
        q = np.broadcast_to(case["q"], (R,) + case["q"].shape)
        eng.set_allocator_state(np.ascontiguousarray(m), np.ascontiguousarray(q))
    if eng.any_shaded:
        fitted = np.isin(np.asarray(case["bidder_kind"]), [ao.BID_SEARCH, ao.BID_BANDIT, ao.BID_POLICY])  # fixtures hold fitted models
        eng.set_bidder_state(case["bidder_f"][:, 0][None, :], case["bidder_f"][:, 1][None, :],
                             initialised=fitted[None, :].astype(np.float64),
                             winrate_w=case["winrate_w"][None] if "winrate_w" in case else None,
                             policy_w=case["policy_w"][None] if "policy_w" in case else None)
    return eng


def log_to_numpy(out, run=0):
    rec = {k: v[run].cpu().numpy() for k, v in out.items()}
    return rec


def replay_case(eng, inp, run0=0):
    kw = {}
    if "ts_eps" in inp and eng.any_learnt:
        kw["ts_eps"] = inp["ts_eps"][None]
    if "gamma_z" in inp:
        kw["gamma_z"] = inp["gamma_z"][None]
    if "grid_u" in inp:
        kw["grid_u"] = inp["grid_u"][None]
    return eng.replay(inp["ctx"][None], inp["parts"][None], inp["u"][None], run0=run0, **kw)
