"""Shared comparison rules for the parity tests (oracle vs golden, CUDA vs oracle, CUDA vs golden).

Bar (BASELINE.json north_star): participants / items / winners / click outcomes bit-exact;
prices, utilities, regrets within a stated floating-point tolerance.  A discrete decision may only
differ on a NEAR-TIE: a round whose arg-max margin (relative gap between the best and the
runner-up score) is below ``TIE_MARGIN`` -- the float32 CTR estimate differs in the last ulp between
numpy, torch and CUDA (hard part 2), which can only flip an arg-max there.
Near-tie rounds are counted and reported, never silently dropped: the tests bound how many there
are.
"""
import numpy as np

TIE_MARGIN = 2e-6           # relative score gap below which an item / winner flip is a near-tie
RTOL_F64 = 1e-11            # float64 arithmetic (Oracle path, prices, utilities)
RTOL_F32_EST = 2e-6         # float32 CTR estimates (learnt path): a few ulp


def compare_rounds(got, want, margins, *, rtol, est_rtol, max_tie_frac=0.01, what="", gamma_rtol=None, prop_rtol=None):
    """Compare per-(round, slot) records.

    got / want: dicts with item, winner, outcome, won, est, value, bid, true_ctr, best_ev, price, second.
    margins: dict(item_margin [T,P], bid_margin [T]) from the oracle.
    Returns a small report dict; raises AssertionError on a real mismatch.
    """
    T, P = want["item"].shape
    item_bad = got["item"] != want["item"]
    # an item flip is admissible only on a near-tie of that participant's arg-max
    real_item_bad = item_bad & ~(margins["item_margin"] < TIE_MARGIN)
    assert not real_item_bad.any(), f"{what}: {real_item_bad.sum()} item mismatches away from ties, first at {np.argwhere(real_item_bad)[:3]}"
    tainted = item_bad.any(axis=1)  # a flipped item changes that round's bid / price legitimately
    win_bad = (np.asarray(got["winner"]) != np.asarray(want["winner"])) & ~tainted
    real_win_bad = win_bad & ~(margins["bid_margin"] < TIE_MARGIN)
    assert not real_win_bad.any(), f"{what}: {real_win_bad.sum()} winner mismatches away from ties, first at {np.argwhere(real_win_bad)[:3]}"
    tainted |= win_bad
    n_tie = int(tainted.sum())
    assert n_tie <= max(1, int(max_tie_frac * T)), f"{what}: too many near-tie rounds ({n_tie} of {T})"
    ok = ~tainted
    for k in ("won", "outcome"):
        assert np.array_equal(np.asarray(got[k])[ok], np.asarray(want[k])[ok]), f"{what}: {k} differs"
    for k, tol in (("value", rtol), ("true_ctr", rtol), ("best_ev", rtol), ("est", est_rtol),
                   ("bid", est_rtol), ("price", est_rtol), ("second", est_rtol)):
        np.testing.assert_allclose(np.asarray(got[k])[ok], np.asarray(want[k])[ok], rtol=tol, atol=tol * 1e-3,
                                   err_msg=f"{what}: {k}")
    for k, tol in (("gamma", gamma_rtol), ("propensity", prop_rtol)):
        if k in got and k in want:
            g, w = np.asarray(got[k])[ok], np.asarray(want[k])[ok]
            assert np.array_equal(np.isnan(g), np.isnan(w)), f"{what}: {k} NaN pattern"
            np.testing.assert_allclose(np.nan_to_num(g), np.nan_to_num(w), rtol=tol or max(rtol, 1e-9), atol=1e-300, err_msg=f"{what}: {k}")
    return {"rounds": T, "near_tie_rounds": n_tie, "min_item_margin": float(margins["item_margin"].min()),
            "min_bid_margin": float(np.min(margins["bid_margin"]))}


def compare_metrics(acc, revenue, met, *, rtol, atol=1e-9, what=""):
    """Per-agent accumulator block [A, M] + revenue vs the reference's getters (tests/golden met_*)."""
    from oracle import auction_oracle as ao

    d = ao.derived_metrics(acc)
    for k in ("net", "gross", "alloc_regret", "estim_regret", "overbid_regret", "underbid_regret"):
        np.testing.assert_allclose(d[k], met[k], rtol=rtol, atol=atol, err_msg=f"{what}: {k}")
    has = met["n_logs"] > 0
    np.testing.assert_allclose(d["ctr_rmse"][has], met["ctr_rmse"][has], rtol=max(rtol, 1e-7), atol=atol, err_msg=f"{what}: ctr_rmse")
    np.testing.assert_allclose(d["best_ev_mean"][has], met["best_ev_mean"][has], rtol=rtol, atol=atol, err_msg=f"{what}: best_ev_mean")
    won = met["n_won"] > 0
    np.testing.assert_allclose(d["ctr_bias"][won], met["ctr_bias"][won], rtol=max(rtol, 1e-7), atol=atol, err_msg=f"{what}: ctr_bias")
    assert np.array_equal(acc[:, ao.M_NPART], met["n_logs"]), f"{what}: participation counts"
    assert np.array_equal(acc[:, ao.M_NWON], met["n_won"]), f"{what}: win counts"
    np.testing.assert_allclose(revenue, met["revenue"], rtol=rtol, atol=atol, err_msg=f"{what}: revenue")
