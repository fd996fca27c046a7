"""GPU parity of the bidder fits (K7) and of post-fit bidding, through the C ABI.

Built so far: ValueLearningBidder(inference='search') -- the win-rate fit (Bidder.py:210-260) and the 128-point grid
search at bid time (Bidder.py:180-196; replay parity is in test_gpu_rounds.py on tests/golden/rounds_fp_search.npz).
Tolerance of the fit: like the allocator fit this is an Adam + plateau-scheduler trajectory with an early stop, so the
bar is |dw| <= 1e-2 on weights of magnitude 1..10, stop epoch +-1 % (+8); against the fit oracle on a fixed epoch budget
the agreement is ~1e-4.
"""
import numpy as np
import pytest

from oracle import auction_oracle as ao
from oracle import fit_oracle as fo
from tests.conftest import GOLDEN_DIR

pytestmark = pytest.mark.gpu


def _gpu():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


def _engine(nA, T, w0, R=1):
    import torch

    import auction_gym_b200 as ag
    from auction_gym_b200 import _lib

    E, V = ao.make_catalog(np.random.default_rng(0), nA, 4, 5)
    eng = ag.Engine(R=R, A=nA, I=4, D=5, Do=4, P=2, mechanism=_lib.FIRST_PRICE, E=E, V=V, n_items=[4] * nA,
                    alloc_kind=[_lib.ALLOC_ORACLE] * nA, bidder_kind=[_lib.BID_SEARCH] * nA, rounds_capacity=T)
    eng.set_bidder_state(1.0, 0.02, initialised=0.0, winrate_w=w0)
    return eng, torch


def _fill(eng, torch, run, records):
    """records: list of (agent, est, value, gamma, won) -> slot 0 of consecutive rounds (slot 1 left invalid)."""
    n = len(records)
    rows = np.zeros((n, 2, 5), np.float32)
    meta = np.zeros((n, 2), np.uint32)
    for t, (a, e, v, g, won) in enumerate(records):
        rows[t, 0] = (e, v, g, 1.0, 0.0)
        meta[t, 0] = (1 << 31) | (int(won) << 30) | int(a)
    eng.bid_rows[run, :n].copy_(torch.from_numpy(rows))
    eng.bid_meta[run, :n].copy_(torch.from_numpy(meta.view(np.int32)))
    return n


def test_winrate_fit_matches_reference_and_oracle():
    _gpu()
    z = np.load(f"{GOLDEN_DIR}/bidfit_winrate.npz")
    agents = [int(a) for a in z["fit_agents"]]
    recs = []
    per = {a: list(zip(z[f"a{a}_est"], z[f"a{a}_value"], z[f"a{a}_gamma"], z[f"a{a}_won"])) for a in agents}
    for k in range(max(len(v) for v in per.values())):  # interleave the agents as a real iteration would
        for j, a in enumerate(agents):
            if k < len(per[a]):
                recs.append((j,) + tuple(per[a][k]))
    w0 = np.stack([z[f"a{a}_w0"] for a in agents])[None]
    eng, torch = _engine(len(agents), len(recs), w0)
    n = _fill(eng, torch, 0, recs)
    eng._check(eng.lib.agym_set_rounds_in_iteration(eng.handle, n))
    info = eng.update_bidders().cpu().numpy()[0, :, 0]
    w1 = eng.bidder_w.cpu().numpy()[0, :, :4]
    init = eng.bidder_d.cpu().numpy()[0, :, 2]
    for j, a in enumerate(agents):
        ref_w, ref_stop = z[f"a{a}_w1"], int(z[f"a{a}_stop_epoch"])
        what = f"agent {a}: rows {int(info[j, 3])}, stop cuda {int(info[j, 0])} / reference {ref_stop}, w {w1[j]} vs {ref_w}"
        assert info[j, 3] == len(per[a]) and init[j] == 1.0, what
        if ref_stop >= 0:
            assert abs(info[j, 0] - ref_stop) <= max(8, 0.015 * ref_stop), what
        else:
            assert info[j, 0] == -1 and info[j, 1] == 32768, what
        np.testing.assert_allclose(w1[j], ref_w, atol=1e-2, rtol=0, err_msg=what)
        # what the model is used for: P(win) on a gamma grid (Bidder.py:187-189)
        g = np.linspace(0.1, 1.0, 64)
        x = np.stack([np.full(64, 0.12), np.full(64, 1.1), g], axis=1).astype(np.float32)
        np.testing.assert_allclose(ao.winrate32(w1[j], x), ao.winrate32(ref_w, x), atol=2e-3, err_msg=what)
    eng.close()


def test_winrate_fit_fixed_budget_against_oracle_and_no_win_fallback():
    _gpu()
    z = np.load(f"{GOLDEN_DIR}/bidfit_winrate.npz")
    a = 4
    rows = list(zip(z[f"a{a}_est"], z[f"a{a}_value"], z[f"a{a}_gamma"], z[f"a{a}_won"]))
    recs = [(0,) + r for r in rows] + [(1,) + r[:3] + (False,) for r in rows[:50]]  # agent 1 loses every auction
    w0 = np.stack([z[f"a{a}_w0"], z[f"a{a}_w0"], z[f"a{a}_w0"]])[None]
    eng, torch = _engine(3, len(recs), w0)
    eng.bidder_d[0, 1, 2] = 1.0  # agent 1 had a fitted model before
    n = _fill(eng, torch, 0, recs)
    eng._check(eng.lib.agym_set_rounds_in_iteration(eng.handle, n))
    info = eng.update_bidders(max_epochs=600).cpu().numpy()[0, :, 0]
    orc = fo.fit_winrate(z[f"a{a}_est"], z[f"a{a}_value"], z[f"a{a}_gamma"], z[f"a{a}_won"], z[f"a{a}_w0"], max_epochs=600)
    w1 = eng.bidder_w.cpu().numpy()[0, :, :4]
    init = eng.bidder_d.cpu().numpy()[0, :, 2]
    assert info[0, 1] == orc["n_epochs"] == 600
    np.testing.assert_allclose(w1[0], orc["w"], atol=2e-4)
    np.testing.assert_allclose(info[0, 2], orc["final_loss"], rtol=1e-5)
    assert init.tolist() == [1.0, 0.0, 0.0]                      # Bidder.py:213-216 fallback for the agent that won nothing
    assert np.array_equal(w1[1], z[f"a{a}_w0"]) and info[1, 3] == 50 and info[1, 1] == 0
    assert np.array_equal(w1[2], z[f"a{a}_w0"]) and info[2, 3] == 0  # an agent that never participated: untouched
    eng.close()


def test_policy_learning_bidder_fit_matches_reference():
    """PolicyLearningBidder.update (Bidder.py:369-431): initialise_policy then the PPO fit, on rows the reference logged.
    The PPO objective has flat directions (the reference's own parameters move by 1e-2..2e-1 between equivalent runs of
    the oracle), so the bar is on what the policy computes -- mu and sigma on the logged contexts -- not on raw weights."""
    _gpu()
    import torch

    import auction_gym_b200 as ag
    from auction_gym_b200 import _lib
    from oracle import policy_oracle as po

    z = np.load(f"{GOLDEN_DIR}/bidfit_ppo.npz")
    nA = 3
    per = {a: list(zip(z[f"a{a}_est"], z[f"a{a}_value"], z[f"a{a}_gamma"], z[f"a{a}_prop"], z[f"a{a}_utility"], z[f"a{a}_won"])) for a in range(nA)}
    recs = []
    for k in range(max(len(v) for v in per.values())):
        for a in range(nA):
            if k < len(per[a]):
                recs.append((a,) + tuple(per[a][k]))
    T = len(recs)
    E, V = ao.make_catalog(np.random.default_rng(0), nA, 4, 5)
    eng = ag.Engine(R=1, A=nA, I=4, D=5, Do=4, P=2, mechanism=_lib.FIRST_PRICE, E=E, V=V, n_items=[4] * nA,
                    alloc_kind=[_lib.ALLOC_ORACLE] * nA, bidder_kind=[_lib.BID_BANDIT] * nA, rounds_capacity=T,
                    bidder_fit=[_lib.BFIT_PL_PPO] * nA)
    eng.set_bidder_state(1.0, 0.02, initialised=0.0, policy_w=np.stack([z[f"a{a}_theta0"] for a in range(nA)])[None])
    rows = np.zeros((T, 2, 5), np.float32)
    meta = np.zeros((T, 2), np.uint32)
    for t, (a, e, v, g, pr, u, won) in enumerate(recs):
        rows[t, 0] = (e, v, g, pr, -u if won else 0.0)  # utility = value * click - price with click = 0
        meta[t, 0] = (1 << 31) | (int(won) << 30) | int(a)
    eng.bid_rows[0, :T].copy_(torch.from_numpy(rows))
    eng.bid_meta[0, :T].copy_(torch.from_numpy(meta.view(np.int32)))
    eng._check(eng.lib.agym_set_rounds_in_iteration(eng.handle, T))
    info = eng.update_bidders().cpu().numpy()[0]
    th = eng.bidder_w.cpu().numpy()[0, :, 4:16]
    assert (eng.bidder_d.cpu().numpy()[0, :, 2] == 1.0).all()
    for a in range(nA):
        X = np.stack([z[f"a{a}_est"], z[f"a{a}_value"]], axis=1).astype(np.float32)
        orc = po.fit_imitation(z[f"a{a}_theta0"], X, z[f"a{a}_gamma"])
        assert info[a, 1, 3] == len(X) and info[a, 2, 3] == len(X)
        assert info[a, 1, 1] == orc["n_epochs"], (info[a, 1], orc["n_epochs"])
        np.testing.assert_allclose(info[a, 1, 2], orc["final_loss"], rtol=2e-3, atol=1e-7)
        got, ref = po.forward(th[a], X), po.forward(z[f"a{a}_theta1"], X)
        np.testing.assert_allclose(got["mu"].mean(), ref["mu"].mean(), atol=2e-3, err_msg=f"agent {a} mean mu")
        np.testing.assert_allclose(got["mu"], ref["mu"], atol=1e-2, err_msg=f"agent {a} mu")
        np.testing.assert_allclose(got["sigma"], ref["sigma"], atol=1e-2, err_msg=f"agent {a} sigma")
    # re-running on an initialised policy skips initialise_policy (Bidder.py:381-382)
    info2 = eng.update_bidders(max_epochs=50).cpu().numpy()[0]
    assert (info2[:, 1, 1] == 0).all() and (info2[:, 2, 1] == 50).all()
    eng.close()


def test_policy_fit_fixed_budget_against_oracle():
    """Every deterministic policy loss with a fixed epoch budget against the hand-differentiated oracle (whose gradients are
    pinned to torch autograd in tests/test_oracle_golden.py)."""
    _gpu()
    import torch

    import auction_gym_b200 as ag
    from auction_gym_b200 import _lib
    from oracle import policy_oracle as po

    z = np.load(f"{GOLDEN_DIR}/bidfit_ppo.npz")
    a = 1
    X = np.stack([z[f"a{a}_est"], z[f"a{a}_value"]], axis=1).astype(np.float32)
    g, pr, u, won = z[f"a{a}_gamma"], z[f"a{a}_prop"], z[f"a{a}_utility"], z[f"a{a}_won"]
    kinds = [("REINFORCE", _lib.BFIT_PL_REINFORCE), ("REINFORCE_offpolicy", _lib.BFIT_PL_OFFPOLICY), ("TRPO", _lib.BFIT_PL_TRPO), ("PPO", _lib.BFIT_PL_PPO)]
    nA, T = len(kinds), len(X)
    E, V = ao.make_catalog(np.random.default_rng(0), nA, 4, 5)
    eng = ag.Engine(R=1, A=nA, I=4, D=5, Do=4, P=4, mechanism=_lib.FIRST_PRICE, E=E, V=V, n_items=[4] * nA,
                    alloc_kind=[_lib.ALLOC_ORACLE] * nA, bidder_kind=[_lib.BID_BANDIT] * nA, rounds_capacity=T,
                    bidder_fit=[k for _, k in kinds])
    th0 = z[f"a{a}_theta_imit"]
    eng.set_bidder_state(1.0, 0.02, initialised=1.0, policy_w=np.broadcast_to(th0, (1, nA, 12)))
    rows = np.zeros((T, 4, 5), np.float32)
    meta = np.zeros((T, 4), np.uint32)
    for s in range(nA):  # the same rows for every agent, one agent per slot
        rows[:, s] = np.stack([X[:, 0], X[:, 1], g, pr, np.where(won, -u, 0.0)], axis=1)
        meta[:, s] = (1 << 31) | (won.astype(np.uint32) << 30) | s
    eng.bid_rows[0, :T].copy_(torch.from_numpy(rows))
    eng.bid_meta[0, :T].copy_(torch.from_numpy(meta.view(np.int32)))
    eng._check(eng.lib.agym_set_rounds_in_iteration(eng.handle, T))
    info = eng.update_bidders(max_epochs=300).cpu().numpy()[0]
    th = eng.bidder_w.cpu().numpy()[0, :, 4:16]
    for s, (name, _) in enumerate(kinds):
        orc = po.fit_policy_ppo(th0, X, g, pr, u, loss_name=name, max_epochs=300)
        assert info[s, 2, 1] == 300 and info[s, 1, 1] == 0
        np.testing.assert_allclose(th[s], orc["theta"], atol=2e-3, err_msg=name)
        np.testing.assert_allclose(info[s, 2, 2], orc["final_loss"], rtol=2e-3, atol=1e-6, err_msg=name)
    eng.close()


@pytest.mark.parametrize("kind_name", ["DR", "VL_POLICY"])
def test_stochastic_policy_fits_against_oracle_with_the_same_noise(kind_name):
    """Doubly Robust (Models.py:198-218) and ValueLearning 'policy' (Bidder.py:292-302) draw fresh rsample noise every epoch.
    The kernel's noise is Philox keyed by (seed, run, iteration, epoch, row), restated in oracle/philox_oracle.py, so a
    fixed budget of epochs can be compared value for value (the device uses __logf / __sincosf: ~1e-6 per draw)."""
    _gpu()
    import torch

    import auction_gym_b200 as ag
    from auction_gym_b200 import _lib
    from oracle import philox_oracle as ph
    from oracle import policy_oracle as po

    z = np.load(f"{GOLDEN_DIR}/bidfit_ppo.npz")
    zw = np.load(f"{GOLDEN_DIR}/bidfit_winrate.npz")
    a = 2
    n = 900
    X = np.stack([z[f"a{a}_est"], z[f"a{a}_value"]], axis=1).astype(np.float32)[:n]
    g, pr, u, won = (z[f"a{a}_{k}"][:n] for k in ("gamma", "prop", "utility", "won"))
    dr = kind_name == "DR"
    E, V = ao.make_catalog(np.random.default_rng(0), 2, 4, 5)
    eng = ag.Engine(R=2, A=2, I=4, D=5, Do=4, P=2, mechanism=_lib.FIRST_PRICE, E=E, V=V, n_items=[4, 4],
                    alloc_kind=[_lib.ALLOC_ORACLE] * 2, bidder_kind=[_lib.BID_BANDIT if dr else _lib.BID_POLICY] * 2, rounds_capacity=n,
                    bidder_fit=[_lib.BFIT_DR if dr else _lib.BFIT_VL_POLICY] * 2, run_offset=5)
    th0, w0 = z[f"a{a}_theta_imit"], zw["a4_w0"]
    eng.set_bidder_state(1.0, 0.02, initialised=1.0, winrate_w=w0, policy_w=th0)
    rows = np.zeros((2, n, 2, 5), np.float32)
    meta = np.zeros((2, n, 2), np.uint32)
    run, agent = 1, 1  # only (run 1, agent 1) has rows: exercises the run / agent indexing of the noise counters
    rows[run, :, 0] = np.stack([X[:, 0], X[:, 1], g, pr, np.where(won, -u, 0.0)], axis=1)
    meta[run, :, 0] = (1 << 31) | (won.astype(np.uint32) << 30) | agent
    eng.bid_rows[:, :n].copy_(torch.from_numpy(rows))
    eng.bid_meta[:, :n].copy_(torch.from_numpy(meta.view(np.int32)))
    eng._check(eng.lib.agym_set_rounds_in_iteration(eng.handle, n))
    seed, it, budget = 77, 3, 120
    info = eng.update_bidders(seed=seed, iteration=it, max_epochs=budget).cpu().numpy()[run, agent]
    th = eng.bidder_w.cpu().numpy()[run, agent, 4:16]
    ww = eng.bidder_w.cpu().numpy()[run, agent, 0:4]
    # oracle: win-rate fit with the bidder's scheduler settings, then the policy fit with the same noise
    wr = fo.fit_winrate(X[:, 0], X[:, 1], g, won, w0, max_epochs=budget, **(dict(patience=256, factor=0.2, stop_after=1024) if dr else {}))
    np.testing.assert_allclose(ww, wr["w"], atol=2e-4)
    W = ao.winrate32(wr["w"], np.stack([X[:, 0], X[:, 1], g], axis=1))
    uhat = (W * (X[:, 0] * X[:, 1] - X[:, 0] * X[:, 1] * g)).astype(np.float32)
    key = ph.make_key(seed, 5 + run)
    noise = lambda e: ph.normal_x(np.arange(n, dtype=np.uint32), np.uint32(e), np.uint32((6 << 16) | agent), np.uint32(it), key)  # noqa: E731
    lp = np.maximum(pr.astype(np.float32), np.float32(1e-15))
    if dr:
        lg = lambda t, eps: po.policy_loss_grad(t, X, g, lp, u, "Doubly Robust", utility_estimates=uhat, winrate_w=wr["w"], eps=eps)  # noqa: E731
        orc = po.run_fit(th0, lg, lr=7e-3, weight_decay=1e-4, max_epochs=budget, stop_after=512,
                         plateau=dict(patience=100, factor=0.2, min_lr=1e-8, threshold=5e-3), noise=noise)
    else:
        lg = lambda t, eps: po.policy_loss_grad(t, X, None, None, None, "DM", winrate_w=wr["w"], eps=eps)  # noqa: E731
        orc = po.run_fit(th0, lg, lr=2e-3, weight_decay=1e-6, max_epochs=budget, stop_after=256,
                         plateau=dict(patience=100, factor=0.1, min_lr=1e-7), noise=noise)
    assert info[2, 1] == orc["n_epochs"] == budget and info[2, 3] == n
    np.testing.assert_allclose(th, orc["theta"], atol=3e-3, err_msg=kind_name)
    np.testing.assert_allclose(info[2, 2], orc["final_loss"], rtol=5e-3, atol=1e-5, err_msg=kind_name)
    # rows of the other (run, agent) pairs are empty: their state is untouched
    assert np.array_equal(eng.bidder_w.cpu().numpy()[0, 0, 4:16], th0)
    eng.close()


def test_empirical_shaded_bidder_update_matches_reference():
    """EmpiricalShadedBidder.update (Bidder.py:60-125) on rows the reference logged: prev_gamma moves to the same bucket
    centre (the device log is float32, the reference's lists float64: equal up to float32 resolution of the bucket edges)."""
    _gpu()
    import torch

    import auction_gym_b200 as ag
    from auction_gym_b200 import _lib
    from oracle.empirical_oracle import fit_empirical

    z = np.load(f"{GOLDEN_DIR}/bidfit_empirical.npz")
    nA = 4
    n = max(len(z[f"a{a}_gamma"]) for a in range(nA))
    E, V = ao.make_catalog(np.random.default_rng(0), nA, 4, 5)
    eng = ag.Engine(R=1, A=nA, I=4, D=5, Do=4, P=4, mechanism=_lib.FIRST_PRICE, E=E, V=V, n_items=[4] * nA,
                    alloc_kind=[_lib.ALLOC_ORACLE] * nA, bidder_kind=[_lib.BID_GAUSS_CLIP] * nA, rounds_capacity=n,
                    bidder_fit=[_lib.BFIT_EMPIRICAL] * nA)
    eng.set_bidder_state([0.8, 0.6, 0.95, 0.5], [0.1, 0.15, 0.05, 0.3])
    rows = np.zeros((n, 4, 5), np.float32)
    meta = np.zeros((n, 4), np.uint32)
    for a in range(nA):  # agent a in slot a
        k = len(z[f"a{a}_gamma"])
        won, click = z[f"a{a}_won"], z[f"a{a}_outcome"].astype(bool)
        rows[:k, a] = np.stack([np.full(k, 0.1), z[f"a{a}_value"], z[f"a{a}_gamma"], np.ones(k), z[f"a{a}_price"]], axis=1)
        meta[:k, a] = (1 << 31) | (won.astype(np.uint32) << 30) | ((won & click).astype(np.uint32) << 29) | a
    eng.bid_rows[0, :n].copy_(torch.from_numpy(rows))
    eng.bid_meta[0, :n].copy_(torch.from_numpy(meta.view(np.int32)))
    eng._check(eng.lib.agym_set_rounds_in_iteration(eng.handle, n))
    info = eng.update_bidders().cpu().numpy()[0, :, 0]
    prev = eng.bidder_d.cpu().numpy()[0, :, 0]
    for a in range(nA):
        g_ref = float(z[f"a{a}_best_gamma"])
        g_orc, b_orc, nb = fit_empirical(z[f"a{a}_gamma"], z[f"a{a}_utility"])
        assert g_orc == g_ref and info[a, 3] == len(z[f"a{a}_gamma"]) and info[a, 1] == nb
        assert abs(prev[a] - g_ref) < 1e-6, (a, prev[a], g_ref, info[a])
    eng.close()


@pytest.mark.parametrize("kind_name", ["DR", "VL_POLICY"])
def test_stochastic_policy_fits_match_the_reference_run_full_trajectory(kind_name):
    """DoublyRobustBidder.update (Bidder.py:477-615) and ValueLearningBidder('policy').update (Bidder.py:210-325) to their own
    stopping rules, against the UNMODIFIED reference fed the device's Philox noise (tests/golden/bidfit_stochastic.npz, written by
    oracle/make_golden_stochastic_fits.py): the win-rate model (deterministic fit) and the policy the stochastic fit lands on.
    The policy fit is a noisy descent with plateau scheduling, so the bar is on what the policy computes -- mu and sigma of the
    shading distribution over the logged contexts -- not on the raw weights (flat directions)."""
    _gpu()
    import torch

    import auction_gym_b200 as ag
    from auction_gym_b200 import _lib

    z = np.load(f"{GOLDEN_DIR}/bidfit_stochastic.npz")
    n = len(z["est"])
    dr = kind_name == "DR"
    run, agent, seed, it = int(z["run"]), int(z["agent"]), int(z["seed"]), int(z["iteration"])
    E, V = ao.make_catalog(np.random.default_rng(0), 2, 4, 5)
    eng = ag.Engine(R=2, A=2, I=4, D=5, Do=4, P=2, mechanism=_lib.FIRST_PRICE, E=E, V=V, n_items=[4, 4],
                    alloc_kind=[_lib.ALLOC_ORACLE] * 2, bidder_kind=[_lib.BID_BANDIT if dr else _lib.BID_POLICY] * 2, rounds_capacity=n,
                    bidder_fit=[_lib.BFIT_DR if dr else _lib.BFIT_VL_POLICY] * 2, run_offset=int(z["run_offset"]))
    eng.set_bidder_state(1.0, 0.02, initialised=1.0, winrate_w=z["w0"], policy_w=z["theta0"])
    rows = np.zeros((2, n, 2, 5), np.float32)
    meta = np.zeros((2, n, 2), np.uint32)
    won = z["won"]
    rows[run, :, 0] = np.stack([z["est"], z["value"], z["gamma"], z["prop"], np.where(won, -z["utility"], 0.0)], axis=1)
    meta[run, :, 0] = (1 << 31) | (won.astype(np.uint32) << 30) | agent
    eng.bid_rows[:, :n].copy_(torch.from_numpy(rows))
    eng.bid_meta[:, :n].copy_(torch.from_numpy(meta.view(np.int32)))
    eng._check(eng.lib.agym_set_rounds_in_iteration(eng.handle, n))
    info = eng.update_bidders(seed=seed, iteration=it).cpu().numpy()[run, agent]
    ww = eng.bidder_w.cpu().numpy()[run, agent, 0:4]
    th = eng.bidder_w.cpu().numpy()[run, agent, 4:16]
    ref_w, ref_stops = z[f"{kind_name}_w1"], z[f"{kind_name}_stops"]
    X = np.stack([z["est"], z["value"]], axis=1).astype(np.float32)
    mu, sg = ao.bandit_mu_sigma32(th, X)
    what = (f"{kind_name}: win-rate stop cuda {int(info[0, 0])} / reference {int(ref_stops[0])}, policy stop cuda {int(info[2, 0])} / reference {int(ref_stops[1])}; "
            f"|dw| {np.abs(ww - ref_w).max():.2e}; mu dev max {np.abs(mu - z[f'{kind_name}_mu']).max():.2e} mean {abs(mu.mean() - z[f'{kind_name}_mu'].mean()):.2e}; "
            f"sigma dev max {np.abs(sg - z[f'{kind_name}_sigma']).max():.2e}")
    print(what)
    # the win-rate fit is deterministic (Bidder.py:239-260 / 518-538): same bar as test_winrate_fit_matches_reference_and_oracle
    assert abs(info[0, 0] - ref_stops[0]) <= max(8, 0.015 * ref_stops[0]), what
    np.testing.assert_allclose(ww, ref_w, atol=1e-2, rtol=0, err_msg=what)
    g = np.linspace(0.1, 1.0, 64)
    xg = np.stack([np.full(64, 0.12), np.full(64, 1.1), g], axis=1).astype(np.float32)
    np.testing.assert_allclose(ao.winrate32(ww, xg), ao.winrate32(ref_w, xg), atol=2e-3, err_msg=what)
    # the policy: the same noise stream, but the device draws it with __logf / __sincosf (~1e-6 per normal) and sums in another
    # order, and a noisy loss decides the plateau scheduler and the stop rule: the trajectories separate slowly
    # (measured on B200: identical policy stop epochs -- 2612 and 1504 --, mu within 8e-5, sigma within 4e-6 of the reference)
    assert info[2, 3] == n and abs(info[2, 0] - ref_stops[1]) <= max(8, 0.02 * ref_stops[1]), what
    np.testing.assert_allclose(mu, z[f"{kind_name}_mu"], atol=1e-3, err_msg=what)
    np.testing.assert_allclose(sg, z[f"{kind_name}_sigma"], atol=1e-4, err_msg=what)
    eng.close()
