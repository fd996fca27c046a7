"""CPU: the oracle restatement against outputs of the UNMODIFIED reference (tests/golden/*.npz,
written by oracle/make_golden.py).  This is what pins the oracle (the reference has no tests)."""
import numpy as np
import pytest

from oracle import auction_oracle as ao
from tests import parity
from tests.conftest import load_golden, multislot_golden_names, round_golden_names


@pytest.mark.parametrize("name", round_golden_names())
def test_round_loop_matches_reference(name):
    case, inp, ref, met = load_golden(name)
    rec, m = ao.simulate_rounds(case, inp["ctx"], inp["parts"], inp["u"], inp.get("ts_eps"), inp.get("gamma_z"), inp.get("grid_u"))
    ref = dict(ref)
    ref["winner"] = np.where(ref["won"].any(axis=1), ref["won"].argmax(axis=1), rec["winner"])
    learnt = bool((case["alloc_kind"] != ao.ALLOC_ORACLE).any())
    net = bool(np.isin(case["bidder_kind"], [ao.BID_BANDIT, ao.BID_POLICY]).any())  # float32 policy net decides gamma
    gtol = dict(gamma_rtol=2e-6, prop_rtol=2e-4) if net else {}
    rep = parity.compare_rounds(rec, ref, rec, rtol=parity.RTOL_F64,
                                est_rtol=parity.RTOL_F32_EST if (learnt or net) else parity.RTOL_F64, what=name, **gtol)
    if rep["near_tie_rounds"] == 0:
        parity.compare_metrics(m["acc"], m["revenue"], met, rtol=2e-6 if (learnt or net) else 1e-10, what=name)


@pytest.mark.parametrize("name", multislot_golden_names())
def test_multislot_rounds_match_reference(name):
    """Several slots per round (Auction.py:30,60-74 with max_slots 2 / 3 passed to the reference's own instantiate_auction):
    winners per slot, per-slot clicks, slot-by-slot charging, the log overwritten by the last slot's set_price."""
    case, inp, ref, met = load_golden(name)
    assert len(multislot_golden_names()) >= 3 and int(case["max_slots"]) >= 2
    rec, m = ao.simulate_rounds(case, inp["ctx"], inp["parts"], inp["u"], inp.get("ts_eps"), inp.get("gamma_z"), None, num_slots=inp["num_slots"])
    for k in ("item", "won", "outcome"):
        assert np.array_equal(rec[k], ref[k]), k
    learnt = bool((case["alloc_kind"] != ao.ALLOC_ORACLE).any())
    tol = parity.RTOL_F32_EST if learnt else parity.RTOL_F64
    for k in ("bid", "price", "second", "est", "true_ctr", "best_ev", "value"):
        np.testing.assert_allclose(rec[k], ref[k], rtol=tol, atol=1e-15, err_msg=k)
    assert (rec["won"].sum(axis=1) == np.minimum(inp["num_slots"], inp["parts"].shape[1] - 1)).all()  # the zip at Auction.py:68
    parity.compare_metrics(m["acc"], m["revenue"], met, rtol=2e-6 if learnt else 1e-10, atol=2e-7 if learnt else 1e-9, what=name)


@pytest.mark.parametrize("name", ["rounds_sp_oracle", "rounds_fp_gauss", "rounds_sp_ts", "rounds_fp_pA", "rounds_sp_p1", "rounds_fp_ties"])
def test_scalar_port_equals_vectorised(name):
    case, inp, ref, met = load_golden(name)
    T = min(200, inp["parts"].shape[0])
    sl = {k: v[:T] for k, v in inp.items()}
    rec_v, m_v = ao.simulate_rounds(case, sl["ctx"], sl["parts"], sl["u"], sl.get("ts_eps"), sl.get("gamma_z"), sl.get("grid_u"))
    rec_s, m_s = ao.simulate_rounds_scalar(case, sl["ctx"], sl["parts"], sl["u"], sl.get("ts_eps"), sl.get("gamma_z"), sl.get("grid_u"))
    assert np.array_equal(rec_s["winner"], rec_v["winner"])
    assert np.array_equal(rec_s["item"], rec_v["item"])
    assert np.array_equal(rec_s["outcome"], rec_v["outcome"].max(axis=1))
    np.testing.assert_allclose(m_s["acc"], m_v["acc"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(m_s["revenue"], m_v["revenue"], rtol=1e-12)


def test_tie_break_is_lowest_slot():
    # AuctionAllocation.py:18-23 on exact ties (SURVEY.md section 4 edge cases)
    bids = np.array([[0.4, 0.9, 0.9], [0.0, 0.0, 0.0], [0.5, 0.5, 0.1]])
    w, price, second, valid = ao.resolve(bids, ao.MECH_FIRST)
    assert w.tolist() == [1, 0, 0]
    assert price.tolist() == [0.9, 0.0, 0.5] and second.tolist() == [0.9, 0.0, 0.5]
    w, price, second, valid = ao.resolve(bids, ao.MECH_SECOND)
    assert price.tolist() == [0.9, 0.0, 0.5]
    # golden: the reference itself on exact ties
    for name in ("rounds_sp_ties", "rounds_fp_ties"):
        case, inp, ref, met = load_golden(name)
        b = ref["bid"]
        top = b.max(axis=1, keepdims=True)
        tied = (b == top).sum(axis=1) > 1
        assert tied.sum() > 20, "fixture should contain exact ties"
        assert np.array_equal(ref["won"][tied].argmax(axis=1), (b == top)[tied].argmax(axis=1))


def test_single_participant_charges_nobody():
    for name in ("rounds_sp_p1", "rounds_fp_p1"):
        case, inp, ref, met = load_golden(name)
        assert ref["won"].sum() == 0 and met["revenue"] == 0.0
        rec, m = ao.simulate_rounds(case, inp["ctx"], inp["parts"], inp["u"], None, inp.get("gamma_z"))
        assert rec["won"].sum() == 0 and m["revenue"] == 0.0


def test_policy_oracle_gradients_match_torch_autograd():
    """oracle/policy_oracle.py's hand-written gradients against torch autograd of the unmodified reference's losses
    (tests/golden/policy_grad.npz, written by oracle/make_golden_policy.py)."""
    from oracle import policy_oracle as po
    from tests.conftest import GOLDEN_DIR

    z = np.load(f"{GOLDEN_DIR}/policy_grad.npz")
    X, g, lp, u, uh, ww = (z[k] for k in ("X", "gammas", "logging_prop", "utility", "utility_estimates", "winrate_w"))
    for c in range(4):
        th, eps = z[f"c{c}_theta"], z[f"c{c}_eps"]
        for name in ("REINFORCE", "REINFORCE_offpolicy", "TRPO", "PPO", "Doubly Robust", "DM", "imitation"):
            key = name.replace(" ", "_")
            if name == "imitation":
                L, G = po.imitation_loss_grad(th, X, g)
            else:
                L, G = po.policy_loss_grad(th, X, g, lp, u, name, utility_estimates=uh, winrate_w=ww, eps=eps)
            rl, rg = z[f"c{c}_{key}_loss"], z[f"c{c}_{key}_grad"]
            assert abs(L - rl) <= 1e-5 * max(abs(rl), 1e-6), (c, name, L, rl)
            assert np.abs(G - rg).max() <= 1e-5 * max(np.abs(rg).max(), 1e-6), (c, name)


def test_imitation_fit_matches_reference():
    """initialise_policy (Models.py:110-133) is deterministic and well conditioned: 16 384 Adam epochs land within 1e-5."""
    from oracle import policy_oracle as po
    from tests.conftest import GOLDEN_DIR

    z = np.load(f"{GOLDEN_DIR}/bidfit_ppo.npz")
    a = 0
    X = np.stack([z[f"a{a}_est"], z[f"a{a}_value"]], axis=1).astype(np.float32)
    r = po.fit_imitation(z[f"a{a}_theta0"], X, z[f"a{a}_gamma"])
    assert r["n_epochs"] == 16384
    np.testing.assert_allclose(r["theta"], z[f"a{a}_theta_imit"], atol=1e-5)


def test_empirical_bidder_update_matches_reference():
    from oracle.empirical_oracle import fit_empirical
    from tests.conftest import GOLDEN_DIR

    z = np.load(f"{GOLDEN_DIR}/bidfit_empirical.npz")
    for a in range(4):
        g, _, _ = fit_empirical(z[f"a{a}_gamma"], z[f"a{a}_utility"])
        assert g == float(z[f"a{a}_best_gamma"])


def test_log_retention_matches_reference():
    """Agent(memory=...) (Agent.py:124-129): what agent.logs holds at the end of each of four iterations, the getters'
    values over those records and the rows Agent.update hands on, against the unmodified reference
    (tests/golden/retention.npz, oracle/make_golden_retention.py)."""
    from oracle import retention_oracle as ro
    from tests import retention_util as ru

    case, memory, inputs, ref = ru.load_retention()
    out = ro.simulate_iterations(case, inputs, memory)
    kept = np.zeros(len(memory), int)
    for it, (o, r) in enumerate(zip(out, ref)):
        rr = dict(r["rec"])
        rr["winner"] = np.where(rr["won"].any(axis=1), rr["won"].argmax(axis=1), o["rec"]["winner"])
        rep = parity.compare_rounds(o["rec"], rr, o["rec"], rtol=parity.RTOL_F64, est_rtol=parity.RTOL_F32_EST, what=f"it{it}")
        assert rep["near_tie_rounds"] == 0
        ru.check_logs_against_reference(o["logs"], r["agents"], case, parity.RTOL_F32_EST, what=f"it{it}")
        parity.compare_metrics(o["acc"], o["revenue"], r["met"], rtol=2e-6, atol=1e-7, what=f"it{it}")
        for a in range(len(memory)):
            n_new = int((inputs[it]["parts"] == a).sum())
            assert len(o["logs"][a]["won"]) == kept[a] + n_new  # kept records + this iteration's
            kept[a] = min(kept[a] + n_new, memory[a])
            assert kept[a] == int(r["agents"][a]["kept"])
            g = r["agents"][a]["mean_gamma"]
            if not np.isnan(g):
                np.testing.assert_allclose(o["acc"][a, ao.M_GAMMA] / o["acc"][a, ao.M_NPART], g, rtol=1e-9)
    assert kept.tolist() == [50, 300, 0, 130, 30]


def test_log_retention_oracle_limits():
    """Size-independent properties of Agent(memory=...): memory 0 is the plain per-iteration reset, a memory larger than
    everything ever logged makes the log-derived accumulators cumulative, memory 1 keeps exactly the last record."""
    from oracle import retention_oracle as ro
    from tests import retention_util as ru

    case, _, inputs, _ = ru.load_retention()
    A = int(case["A"])
    plain = [ao.simulate_rounds(case, nz["ctx"], nz["parts"], nz["u"], nz.get("ts_eps"), nz.get("gamma_z"))[1]["acc"] for nz in inputs]
    none = ro.simulate_iterations(case, inputs, np.zeros(A, int))
    keep_all = ro.simulate_iterations(case, inputs, np.full(A, 10**6))
    one = ro.simulate_iterations(case, inputs, np.ones(A, int))
    logged = [ao.M_ALLOC_REG, ao.M_ESTIM_REG, ao.M_OVERBID, ao.M_UNDERBID, ao.M_SQERR, ao.M_BIAS, ao.M_NPART, ao.M_NWON, ao.M_BEST_EV, ao.M_GAMMA]
    cum = np.zeros_like(plain[0])
    for it in range(len(inputs)):
        np.testing.assert_allclose(none[it]["acc"], plain[it], rtol=1e-12, atol=1e-12)
        cum += plain[it]
        np.testing.assert_allclose(keep_all[it]["acc"][:, logged], cum[:, logged], rtol=1e-10, atol=1e-10)
        for o in (none, keep_all, one):  # utilities always restart (Agent.py:120-122)
            np.testing.assert_allclose(o[it]["acc"][:, [ao.M_NET, ao.M_GROSS]], plain[it][:, [ao.M_NET, ao.M_GROSS]], rtol=1e-12, atol=1e-12)
        if it > 0:
            extra = one[it]["acc"][:, ao.M_NPART] - plain[it][:, ao.M_NPART]
            took_part_before = np.array([(np.concatenate([nz["parts"].ravel() for nz in inputs[:it]]) == a).any() for a in range(A)])
            assert np.array_equal(extra, took_part_before.astype(float))
            for a in range(A):  # ... and it is the most recent one
                last = one[it - 1]["logs"][a]
                assert one[it]["logs"][a]["best_ev"][0] == last["best_ev"][-1]
