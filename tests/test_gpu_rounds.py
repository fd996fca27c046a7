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
from tests.conftest import load_golden, round_golden_names

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


@pytest.mark.parametrize("G", ["8", "16", "32"])
@pytest.mark.parametrize("name", ["rounds_sp_ts_64x64", "rounds_sp_oracle_64x64", "rounds_fp_pA", "rounds_sp_ragged", "rounds_fp_search", "rounds_fp_bandit"])
def test_replay_is_exact_for_every_lane_group_width(name, G):
    """The launcher picks the lane-group width from the catalog width and the launch size (8 lanes for large launches);
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
