"""GPU parity of the round loop (K1-K5), through the C ABI.

Replay mode feeds the identical host-drawn contexts / participants / noise to the CUDA path, to the
oracle restatement, and (when the fixture was made) to the unmodified reference.  Bar: items, winners,
won flags and click outcomes bit-exact; prices / utilities / regrets within the stated tolerance
(FP64 mode: 1e-11 rel on the float64 path, 2e-6 rel where a float32 CTR estimate enters; FP32 mode:
1e-5 rel, BASELINE.json north_star).
"""
import numpy as np
import pytest

from oracle import auction_oracle as ao
from tests import parity
from tests.conftest import load_golden, multislot_golden_names, round_golden_names

pytestmark = pytest.mark.gpu


def _gpu():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from tests import gpu_util

    return gpu_util


@pytest.mark.parametrize("name", round_golden_names())
def test_replay_fp64_matches_oracle_and_reference(name):
    gu = _gpu()
    from auction_gym_b200 import _lib

    case, inp, ref, met = load_golden(name)
    eng = gu.engine_from_case(case, R=1, precision=_lib.FP64)
    out = gu.replay_case(eng, inp)
    got = gu.log_to_numpy(out)
    assert np.array_equal(got["agent"], inp["parts"])
    rec, m = ao.simulate_rounds(case, inp["ctx"], inp["parts"], inp["u"], inp.get("ts_eps"), inp.get("gamma_z"), inp.get("grid_u"))
    learnt = bool((case["alloc_kind"] != ao.ALLOC_ORACLE).any())
    net = bool(np.isin(case["bidder_kind"], [ao.BID_BANDIT, ao.BID_POLICY]).any())  # float32 policy net decides gamma
    gtol = dict(gamma_rtol=2e-6, prop_rtol=2e-4) if net else {}
    est_rtol = parity.RTOL_F32_EST if (learnt or net) else parity.RTOL_F64
    # CUDA vs oracle
    rep = parity.compare_rounds(got, rec, rec, rtol=parity.RTOL_F64, est_rtol=est_rtol, what=f"{name} cuda-vs-oracle", **gtol)
    # CUDA vs the unmodified reference's own outputs
    ref = dict(ref)
    ref["winner"] = np.where(ref["won"].any(axis=1), ref["won"].argmax(axis=1), rec["winner"])
    parity.compare_rounds(got, ref, rec, rtol=parity.RTOL_F64, est_rtol=est_rtol, what=f"{name} cuda-vs-reference", **gtol)
    acc, rev = eng.metrics()
    if rep["near_tie_rounds"] == 0:
        parity.compare_metrics(acc[0], rev[0], met, rtol=2e-6 if (learnt or net) else 1e-10, what=f"{name} metrics-vs-reference")
        np.testing.assert_allclose(acc[0], m["acc"], rtol=2e-6 if (learnt or net) else 1e-10, atol=1e-9)
    eng.close()


@pytest.mark.parametrize("G", ["4", "8", "16", "32"])
@pytest.mark.parametrize("name", ["rounds_sp_ts_64x64", "rounds_sp_oracle_64x64", "rounds_fp_pA", "rounds_sp_ragged", "rounds_fp_search", "rounds_fp_bandit"])
def test_replay_is_exact_for_every_lane_group_width(name, G):
    """The launcher picks the lane-group width from the participants per round (sim_group_width; 4 to 32 lanes);
    every width must reproduce the reference's discrete decisions -- forced here through the "sim_g" option (agym_set_option)."""
    gu = _gpu()
    from auction_gym_b200 import _lib

    case, inp, ref, met = load_golden(name)
    eng = gu.engine_from_case(case, R=1, precision=_lib.FP64)
    eng.set_option("sim_g", int(G))
    got = gu.log_to_numpy(gu.replay_case(eng, inp))
    rec, m = ao.simulate_rounds(case, inp["ctx"], inp["parts"], inp["u"], inp.get("ts_eps"), inp.get("gamma_z"), inp.get("grid_u"))
    learnt = bool((case["alloc_kind"] != ao.ALLOC_ORACLE).any())
    net = bool(np.isin(case["bidder_kind"], [ao.BID_BANDIT, ao.BID_POLICY]).any())
    gtol = dict(gamma_rtol=2e-6, prop_rtol=2e-4) if net else {}
    est_rtol = parity.RTOL_F32_EST if (learnt or net) else parity.RTOL_F64
    ref = dict(ref)
    ref["winner"] = np.where(ref["won"].any(axis=1), ref["won"].argmax(axis=1), rec["winner"])
    rep = parity.compare_rounds(got, ref, rec, rtol=parity.RTOL_F64, est_rtol=est_rtol, what=f"{name} G={G} cuda-vs-reference", **gtol)
    assert rep["near_tie_rounds"] == 0
    acc, rev = eng.metrics()
    parity.compare_metrics(acc[0], rev[0], met, rtol=2e-6 if (learnt or net) else 1e-10, what=f"{name} G={G}")
    eng.close()


@pytest.mark.parametrize("name", round_golden_names())
def test_replay_fp32_within_tolerance(name):
    gu = _gpu()
    from auction_gym_b200 import _lib

    case, inp, ref, met = load_golden(name)
    eng = gu.engine_from_case(case, R=1, precision=_lib.FP32)
    got = gu.log_to_numpy(gu.replay_case(eng, inp))
    rec, m = ao.simulate_rounds(case, inp["ctx"], inp["parts"], inp["u"], inp.get("ts_eps"), inp.get("gamma_z"), inp.get("grid_u"))
    T = inp["parts"].shape[0]
    # float32 decisions: an arg-max may flip only where the float64 margin is below float32 resolution,
    # a click only where u is within float32 resolution of the CTR
    margins = {"item_margin": rec["item_margin"], "bid_margin": rec["bid_margin"]}
    item_bad = got["item"] != rec["item"]
    assert not (item_bad & (margins["item_margin"] > 1e-5)).any(), name
    tainted = item_bad.any(axis=1)
    n_gamma_flip = 0
    if (case["bidder_kind"] == ao.BID_SEARCH).any():
        # the 128-point grid search is an arg-max over a nearly flat utility curve: in float32 it may land on a
        # neighbouring grid point of (almost) equal utility; such rounds are counted, bounded, and excluded
        gflip = (np.abs(np.nan_to_num(got["gamma"]) - np.nan_to_num(rec["gamma"])) > 1e-6).any(axis=1)
        n_gamma_flip = int(gflip.sum())
        assert n_gamma_flip <= 0.05 * T, f"{name}: {n_gamma_flip} float32 grid-search flips"
        tainted |= gflip
    win_bad = (got["winner"] != rec["winner"]) & ~tainted
    assert not (win_bad & (margins["bid_margin"] > 1e-5)).any(), name
    tainted |= win_bad
    ar = np.arange(T)
    near_click = np.abs(inp["u"] - rec["true_ctr"][ar, rec["winner"]]) < 1e-6
    out_bad = (got["outcome"] != rec["outcome"]).any(axis=1) & ~tainted
    assert not (out_bad & ~near_click).any(), name
    tainted |= out_bad
    assert tainted.sum() - n_gamma_flip <= max(2, T // 100), f"{name}: {tainted.sum()} float32 near-ties"
    ok = ~tainted
    for k in ("est", "value", "bid", "true_ctr", "best_ev", "price", "second"):
        np.testing.assert_allclose(got[k][ok], rec[k][ok], rtol=1e-5, atol=1e-8, err_msg=f"{name}: {k}")
    if not tainted.any():
        acc, rev = eng.metrics()
        np.testing.assert_allclose(acc[0], m["acc"], rtol=2e-5, atol=1e-5)
        np.testing.assert_allclose(rev[0], m["revenue"], rtol=1e-5)
    eng.close()


def test_replay_multi_run_indexing():
    """Three runs with different learnt state and different noise in one launch == three single-run launches."""
    gu = _gpu()
    from auction_gym_b200 import _lib

    case, inp, ref, met = load_golden("rounds_sp_ts_q")
    T = 120
    R = 3
    rng = np.random.default_rng(5)
    ms = rng.standard_normal((R,) + case["m"].shape).astype(np.float32)
    qs = (1 + 5 * rng.random((R,) + case["q"].shape)).astype(np.float32)
    sl = lambda a, r: a[r * T:(r + 1) * T]  # noqa: E731
    eng = gu.engine_from_case(case, R=R, precision=_lib.FP64)
    eng.set_allocator_state(ms, qs)
    out = eng.replay(np.stack([sl(inp["ctx"], r) for r in range(R)]), np.stack([sl(inp["parts"], r) for r in range(R)]),
                     np.stack([sl(inp["u"], r) for r in range(R)]), ts_eps=np.stack([sl(inp["ts_eps"], r) for r in range(R)]))
    acc, rev = eng.metrics()
    for r in range(R):
        c = dict(case)
        c["m"], c["q"] = ms[r], qs[r]
        rec, m = ao.simulate_rounds(c, sl(inp["ctx"], r), sl(inp["parts"], r), sl(inp["u"], r), sl(inp["ts_eps"], r))
        got = gu.log_to_numpy(out, r)
        parity.compare_rounds(got, rec, rec, rtol=parity.RTOL_F64, est_rtol=parity.RTOL_F32_EST, what=f"run {r}")
        np.testing.assert_allclose(acc[r], m["acc"], rtol=2e-6, atol=1e-9)
        np.testing.assert_allclose(rev[r], m["revenue"], rtol=2e-6)
    eng.close()


def test_abi_error_paths():
    gu = _gpu()
    import auction_gym_b200 as ag
    from auction_gym_b200 import _lib

    case, inp, ref, met = load_golden("rounds_sp_oracle")
    with pytest.raises(ag.AgymError):
        ag.Engine(R=1, A=2, I=2, D=5, Do=4, P=3, mechanism=0, E=np.zeros((2, 2, 6)), V=np.ones((2, 2)), n_items=[2, 2],
                  alloc_kind=[0, 0], bidder_kind=[0, 0])  # P > A
    eng = gu.engine_from_case(case)
    with pytest.raises(ag.AgymError, match="bad run range"):
        eng.replay(inp["ctx"][None, :4], inp["parts"][None, :4], inp["u"][None, :4], run0=1)
    eng.close()


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
@pytest.mark.parametrize("name", multislot_golden_names())
def test_replay_multislot_matches_oracle_and_reference(name, precision):
    """Several slots per round (Auction.py:30,60-74; AuctionAllocation.py:19-23,33-35) in the fused kernel against the oracle
    and against the UNMODIFIED reference run with max_slots 2 / 3 (tests/golden/rounds_*_slots.npz): winners of every slot,
    per-slot clicks, slot-by-slot charging, the log overwritten by the last slot's price, S winner-log rows per round."""
    gu = _gpu()
    import auction_gym_b200 as ag
    from auction_gym_b200 import _lib

    case, inp, ref, met = load_golden(name)
    S = int(case["max_slots"])
    A, I = int(case["A"]), int(case["I"])
    eng = ag.Engine(R=1, A=A, I=I, D=int(case["D"]), Do=int(case["Do"]), P=int(case["P"]), mechanism=int(case["mechanism"]), E=case["E"], V=case["V"],
                    n_items=case["n_items"], alloc_kind=[gu.ALLOC[int(k)] for k in case["alloc_kind"]],
                    bidder_kind=[gu.BID[int(k)] for k in case["bidder_kind"]], embedding_var=float(case["embedding_var"]),
                    precision=_lib.FP64 if precision == "fp64" else _lib.FP32, max_slots=S)
    learnt = bool((case["alloc_kind"] != ao.ALLOC_ORACLE).any())
    if learnt:
        eng.set_allocator_state(case["m"][None], case["q"][None])
    if eng.any_shaded:
        eng.set_bidder_state(case["bidder_f"][:, 0][None, :], case["bidder_f"][:, 1][None, :])
    kw = {}
    if learnt:
        kw["ts_eps"] = inp["ts_eps"][None]
    if "gamma_z" in inp:
        kw["gamma_z"] = inp["gamma_z"][None]
    got = gu.log_to_numpy(eng.replay(inp["ctx"][None], inp["parts"][None], inp["u"][None], num_slots=inp["num_slots"][None], **kw))
    rec, m = ao.simulate_rounds(case, inp["ctx"], inp["parts"], inp["u"], inp.get("ts_eps"), inp.get("gamma_z"), None, num_slots=inp["num_slots"])
    T, P = inp["parts"].shape
    exact = precision == "fp64"
    # float32: a rank may flip only where two DIFFERENT bids are within float32 resolution (exactly equal bids stay equal)
    srt = np.sort(rec["bid"], axis=1)
    gaps = np.diff(srt, axis=1)
    near = ((gaps > 0) & (gaps < 1e-5 * np.abs(srt[:, 1:]))).any(axis=1)
    ok = np.ones(T, bool) if exact else ~near & (rec["item_margin"] > 1e-5).all(axis=1)
    assert ok.mean() > 0.97
    for want, who in ((rec, "oracle"), (ref, "reference")):
        for k in ("item", "won", "outcome"):
            assert np.array_equal(got[k][ok], np.asarray(want[k])[ok]), f"{name} {who}: {k}"
        tol = (parity.RTOL_F32_EST if learnt else parity.RTOL_F64) if exact else 1e-5
        for k in ("bid", "price", "second", "est", "true_ctr", "best_ev", "value"):
            np.testing.assert_allclose(got[k][ok], np.asarray(want[k])[ok], rtol=tol, atol=1e-12 if exact else 1e-8, err_msg=f"{name} {who}: {k}")
    assert np.array_equal(got["winner"][ok], rec["winner"][ok])
    acc, rev = eng.metrics()
    if ok.all():
        parity.compare_metrics(acc[0], rev[0], met, rtol=(2e-6 if learnt else 1e-10) if exact else 2e-5, atol=(2e-7 if learnt else 1e-9) if exact else 1e-5,
                               what=f"{name} metrics-vs-reference")
    # the winner log: S rows per round, row k = the winner of slot k (agent, item, click), the others invalid
    if learnt:
        meta = eng.fit_meta[0, :T * S].cpu().numpy().view(np.uint32).reshape(T, S)
        valid = (meta >> 31).astype(bool)
        assert np.array_equal(valid.sum(axis=1)[ok], rec["n_charged"][ok])
        for t in np.nonzero(ok)[0][:64]:
            for s in range(P):
                if rec["won"][t, s]:
                    k = rec["rank"][t, s]
                    assert (meta[t, k] >> 12) & 0xFFF == inp["parts"][t, s] and meta[t, k] & 0xFFF == rec["item"][t, s]
                    assert bool((meta[t, k] >> 30) & 1) == bool(rec["outcome"][t, s])
    eng.close()


def test_multislot_production_rounds_and_fit():
    """Production mode with max_slots = 3: the slot count is uniform on {1, 2, 3}, revenue grows with it, and the allocator fit
    consumes the S-rows-per-round winner log (more rows than rounds)."""
    gu = _gpu()
    import auction_gym_b200 as ag
    from auction_gym_b200 import _lib

    case, inp, ref, met = load_golden("rounds_sp_ts")
    A, I, T = int(case["A"]), int(case["I"]), 6000
    out = {}
    for S in (1, 3):
        eng = ag.Engine(R=2, A=A, I=I, D=int(case["D"]), Do=int(case["Do"]), P=4, mechanism=_lib.SECOND_PRICE, E=case["E"], V=case["V"],
                        n_items=case["n_items"], alloc_kind=[_lib.ALLOC_TS] * A, bidder_kind=[_lib.BID_TRUTHFUL] * A, precision=_lib.FP32, max_slots=S)
        eng.set_allocator_state(np.broadcast_to(case["m"], (2,) + case["m"].shape).copy())
        log = eng.simulate(5, 0, T, ("won", "price"))
        won = log["won"].cpu().numpy()
        acc, rev = eng.metrics()
        info = eng.update_allocators(max_epochs=50).cpu().numpy()
        out[S] = (won.sum(axis=2), rev, info[..., 3].sum(axis=1))
        eng.close()
    n1, rev1, rows1 = out[1]
    n3, rev3, rows3 = out[3]
    assert (n1 == 1).all() and (rows1 == T).all()
    frac = np.bincount(n3.ravel(), minlength=4)[1:] / n3.size
    assert np.abs(frac - 1 / 3).max() < 0.03                      # rng.integers(1, max_slots + 1), Auction.py:30
    assert (rows3 == n3.sum(axis=1)).all() and (rows3 > 1.8 * T).all()  # one fit row per charged slot
    assert (rev3 > 1.5 * rev1).all()
