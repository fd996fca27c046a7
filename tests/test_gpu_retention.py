"""GPU: log retention across iterations (Agent(memory=...), Agent.py:124-129) through the C ABI, against the unmodified
reference's four-iteration run in tests/golden/retention.npz and the oracle restatement (oracle/retention_oracle.py).

Model state is held fixed (reset after every update) exactly as the fixture was generated, so every iteration's discrete
decisions are bit-comparable; what is under test is which records survive an iteration boundary, where they sit in the
logs, what the accumulators restart from and which rows the two fits see.
"""
import numpy as np
import pytest

from oracle import auction_oracle as ao
from oracle import fit_oracle as fo
from oracle import retention_oracle as ro
from oracle.empirical_oracle import fit_empirical
from tests import parity
from tests import retention_util as ru

pytestmark = pytest.mark.gpu


def _gpu():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from tests import gpu_util

    return gpu_util


def _engine(gu, case, memory, R, precision, rounds_capacity):
    import auction_gym_b200 as ag
    from auction_gym_b200 import _lib

    A = int(case["A"])
    fit = [_lib.BFIT_EMPIRICAL if k == ao.BID_GAUSS_CLIP else _lib.BFIT_NONE for k in case["bidder_kind"]]
    eng = ag.Engine(R=R, A=A, I=int(case["I"]), D=int(case["D"]), Do=int(case["Do"]), P=int(case["P"]), mechanism=int(case["mechanism"]),
                    E=case["E"], V=case["V"], n_items=case["n_items"], alloc_kind=[gu.ALLOC[int(k)] for k in case["alloc_kind"]],
                    bidder_kind=[gu.BID[int(k)] for k in case["bidder_kind"]], embedding_var=float(case["embedding_var"]),
                    precision=precision, rounds_capacity=rounds_capacity, bidder_fit=fit, memory=memory)
    return eng


def _reset_state(eng, case, R):
    eng.set_allocator_state(np.ascontiguousarray(np.broadcast_to(case["m"], (R,) + case["m"].shape)),
                            np.ascontiguousarray(np.broadcast_to(case["q"], (R,) + case["q"].shape)))
    eng.set_bidder_state(case["bidder_f"][:, 0][None, :], case["bidder_f"][:, 1][None, :])


@pytest.mark.parametrize("capacity", [200, 0])  # logs sized up front / grown on demand (retained rows move with them)
def test_retention_matches_reference_over_four_iterations(capacity):
    gu = _gpu()
    from auction_gym_b200 import _lib

    case, memory, inputs, ref = ru.load_retention()
    A, P, Do, R = int(case["A"]), int(case["P"]), int(case["Do"]), 2
    out = ro.simulate_iterations(case, inputs, memory)
    eng = _engine(gu, case, memory, R, _lib.FP64, capacity)
    B = int(memory.sum())
    assert eng.log_base == B and int(eng.lib.agym_retained_capacity(eng.handle)) == (B if capacity else 0)
    off = np.concatenate([[0], np.cumsum(memory)[:-1]])
    _reset_state(eng, case, R)
    for it, (nz, o, r) in enumerate(zip(inputs, out, ref)):
        T = nz["parts"].shape[0]
        kw = dict(ts_eps=np.stack([nz["ts_eps"]] * R), gamma_z=np.stack([nz["gamma_z"]] * R))
        got = eng.replay(np.stack([nz["ctx"]] * R), np.stack([nz["parts"]] * R), np.stack([nz["u"]] * R), **kw)
        rep = parity.compare_rounds(gu.log_to_numpy(got), o["rec"], o["rec"], rtol=parity.RTOL_F64, est_rtol=parity.RTOL_F32_EST, what=f"it{it}")
        assert rep["near_tie_rounds"] == 0
        assert eng.rounds_in_iteration == T
        # ---- the getters (Agent.py:96-118) over kept + new records
        acc, rev = eng.metrics()
        np.testing.assert_allclose(acc[0], acc[1], rtol=1e-12, atol=1e-12)  # FP64 atomics: order differs between runs
        np.testing.assert_allclose(acc[0], o["acc"], rtol=2e-6, atol=1e-5, err_msg=f"it{it}")
        parity.compare_metrics(acc[0], rev[0], r["met"], rtol=2e-6, atol=1e-5, what=f"it{it} vs reference")
        # ---- Agent.update (Agent.py:79-94): allocator on the won records, bidder on all of them
        info = eng.update_allocators(max_epochs=100).cpu().numpy()
        binfo = eng.update_bidders(want_info=True).cpu().numpy()
        m1, q1 = eng.m.cpu().numpy(), eng.q.cpu().numpy()
        prev_gamma = eng.bidder_d.cpu().numpy()[..., 0]
        assert np.array_equal(m1[0], m1[1]) and np.array_equal(prev_gamma[0], prev_gamma[1])
        for a in range(A):
            rf, lg = r["agents"][a], o["logs"][a]
            what = f"it{it} agent {a}"
            if case["alloc_kind"][a] != ao.ALLOC_ORACLE:
                assert info[0, a, 3] == len(rf["fit_items"]), what  # rows the reference's allocator.update received
                orc = fo.fit_allocator(rf["fit_ctx"], rf["fit_items"], rf["fit_y"], case["m"][a], case["q"][a], case["m"][a], max_epochs=100)
                np.testing.assert_allclose(m1[0, a], orc["m"], atol=3e-4, err_msg=what)
                np.testing.assert_allclose(q1[0, a], orc["q"], rtol=2e-4, err_msg=what)
                np.testing.assert_allclose(info[0, a, 2], orc["final_loss"], rtol=2e-5, err_msg=what)
            if case["bidder_kind"][a] == ao.BID_GAUSS_CLIP:
                assert binfo[0, a, 0, 3] == len(rf["won"]), what   # rows the reference's bidder.update received
                won = rf["won"].astype(bool)
                util = np.where(won, rf["values"] * rf["outcomes"] - rf["prices"], 0.0)  # Bidder.py:62-64
                g, _, _ = fit_empirical(rf["gammas"], util)
                np.testing.assert_allclose(prev_gamma[0, a], g, atol=2e-6, err_msg=what)
        _reset_state(eng, case, R)
        # ---- iteration boundary: Agent.clear_utility / clear_logs, Auction.clear_revenue (main.py:151-155)
        eng.clear_iteration()
        assert eng.rounds_in_iteration == 0
        acc, rev = eng.metrics()
        bid_rows, bid_meta = eng.bid_rows.cpu().numpy(), eng.bid_meta.cpu().numpy().view(np.uint32)
        fit_ctx, fit_meta = eng.fit_ctx.cpu().numpy(), eng.fit_meta.cpu().numpy().view(np.uint32)
        held = (bid_meta[0, :B, 0] >> 31).astype(bool)
        assert (rev == 0).all() and np.array_equal(bid_meta[0, :B], bid_meta[1, :B])
        assert np.array_equal(bid_rows[0, :B, 0][held], bid_rows[1, :B, 0][held], equal_nan=True)
        assert (bid_meta[0, :B, 1:] == 0).all()  # retained records sit in slot 0
        for a in range(A):
            kept = ro.keep_last(o["logs"][a], int(memory[a]))
            k = 0 if kept is None else len(kept["won"])
            what = f"after it{it} agent {a}"
            assert k == int(r["agents"][a]["kept"]), what
            want = ro.metric_sums(kept) if k else np.zeros(ao.NUM_METRICS)
            np.testing.assert_allclose(acc[0, a], want, rtol=2e-6, atol=1e-5, err_msg=what)  # net / gross restart at 0
            rows = slice(off[a], off[a] + k)
            mt = bid_meta[0, rows, 0]
            assert (mt >> 31).all() and ((mt & 0xFFF) == a).all(), what
            assert (bid_meta[0, off[a] + k:off[a] + memory[a], 0] == 0).all() and (fit_meta[0, off[a] + k:off[a] + memory[a]] == 0).all(), what
            if k == 0:
                continue
            won = kept["won"].astype(bool)
            assert np.array_equal((mt >> 30) & 1, won) and np.array_equal((mt >> 29) & 1, kept["outcome"].astype(bool) & won), what
            assert np.array_equal((mt >> 12) & 0xFFF, kept["item"]), what
            np.testing.assert_allclose(bid_rows[0, rows, 0, 0], kept["est"], rtol=1e-6, err_msg=what)
            np.testing.assert_allclose(bid_rows[0, rows, 0, 1], kept["value"], rtol=1e-6, err_msg=what)
            np.testing.assert_allclose(bid_rows[0, rows, 0, 2], kept["gamma"], rtol=1e-6, err_msg=what)  # NaN == NaN (truthful)
            np.testing.assert_allclose(bid_rows[0, rows, 0, 4], kept["price"], rtol=1e-6, err_msg=what)
            fm = fit_meta[0, rows]
            assert np.array_equal(fm >> 31, won), what  # the winner log holds the won records only (Agent.py:91)
            assert np.array_equal(fm[won] & 0xFFF, kept["item"][won]) and ((fm[won] >> 12) & 0xFFF == a).all(), what
            np.testing.assert_allclose(fit_ctx[0, rows][won], kept["ctx"][won], rtol=1e-6, err_msg=what)
        if capacity == 0 and it == 1:  # grow the logs between iterations: the retained rows move to the new buffers
            eng.reserve_rounds(3 * T)
    eng.close()


def test_retention_production_mode_and_full_clear():
    """In-kernel noise: the accumulators after the boundary equal the sums over the retained rows' summands, an agent with
    memory >= everything it ever logged keeps all of it, and agym_clear_iteration drops every retained record."""
    gu = _gpu()
    from auction_gym_b200 import _lib

    case, memory, inputs, ref = ru.load_retention()
    A, P, R, T = int(case["A"]), int(case["P"]), 3, 500
    memory = np.array([40, 5000, 0, 7, 300], np.int32)
    eng = _engine(gu, case, memory, R, _lib.FP32, T)
    _reset_state(eng, case, R)
    B = int(memory.sum())
    off = np.concatenate([[0], np.cumsum(memory)[:-1]])
    total = np.zeros((R, A))
    for it in range(3):
        eng.simulate(7, it, T)
        acc, _ = eng.metrics()
        new = acc[..., _lib.M_NPART] - np.minimum(total, memory[None, :])
        total += new
        assert (new.sum(axis=1) == T * P).all()
        eng.clear_iteration()
        acc2, rev2 = eng.metrics()
        assert (acc2[..., _lib.M_NET] == 0).all() and (acc2[..., _lib.M_GROSS] == 0).all() and (rev2 == 0).all()
        np.testing.assert_array_equal(acc2[..., _lib.M_NPART], np.minimum(total, memory[None, :]))
        terms = eng.terms.cpu().numpy()
        meta = eng.bid_meta.cpu().numpy().view(np.uint32)
        for a in range(A):
            rows = slice(off[a], off[a] + memory[a])
            valid = (meta[:, rows, 0] >> 31).astype(bool)
            assert np.array_equal(valid.sum(axis=1), acc2[:, a, _lib.M_NPART])
            s = (terms[:, rows, 0] * valid[..., None]).sum(axis=1)  # [R, 8]
            cols = [_lib.M_ALLOC_REGRET, _lib.M_ESTIM_REGRET, _lib.M_OVERBID_REGRET, _lib.M_UNDERBID_REGRET, _lib.M_SQERR, _lib.M_BIAS,
                    _lib.M_GAMMA, _lib.M_BEST_EV]
            np.testing.assert_allclose(acc2[:, a, cols], s, rtol=1e-12, atol=1e-12)
            won = ((meta[:, rows, 0] >> 30) & 1).astype(bool) & valid
            np.testing.assert_array_equal(won.sum(axis=1), acc2[:, a, _lib.M_NWON])
    # agent 1 (memory 5000) kept everything it ever logged
    assert (total[:, 1] == eng.acc.cpu().numpy()[:, 1, _lib.M_NPART]).all() and (total[:, 1] > 3 * T * P / A * 0.8).all()
    eng._check(eng.lib.agym_clear_iteration(eng.handle, eng._stream()))
    assert (eng.acc.cpu().numpy() == 0).all() and (eng.bid_meta[:, :B].cpu().numpy() == 0).all() and (eng.fit_meta[:, :B].cpu().numpy() == 0).all()
    eng.close()


def test_agent_surface_with_memory():
    """The reference surface: Agent(memory=...) + agent.clear_logs() per agent (main.py:151-155) keeps agent.logs[-memory:]."""
    _gpu()
    import auction_gym_b200 as ag
    from auction_gym_b200 import _lib

    rng = np.random.default_rng(3)
    A, I, D, Do = 3, 4, 5, 4
    E, V = ao.make_catalog(rng, A, I, D)
    names = [f"agent {a}" for a in range(A)]
    mems = [5, 0, 1000]
    agents = [ag.Agent(rng, names[a], I, V[a], ag.OracleAllocator(rng), ag.TruthfulBidder(rng), memory=mems[a]) for a in range(A)]
    auction = ag.Auction(rng, ag.SecondPrice(), agents, {n: E[a] for a, n in enumerate(names)}, {n: V[a] for a, n in enumerate(names)},
                         1, D, 1.0, Do, 2, precision=_lib.FP64)
    counts = np.zeros(A, int)
    for it in range(3):
        for _ in range(12):
            auction.simulate_opportunity()
        logs = [a.logs for a in agents]
        for a, agent in enumerate(agents):
            new = len(logs[a]) - min(counts[a], mems[a])
            counts[a] = min(counts[a], mems[a]) + new
            assert len(logs[a]) == counts[a]
            want = sum(o.best_expected_value - o.true_CTR * o.value for o in logs[a])
            np.testing.assert_allclose(agent.get_allocation_regret(), want, rtol=1e-9, atol=1e-12)
            np.testing.assert_allclose(agent.get_overbid_regret(), sum((o.price - o.second_price) * o.won for o in logs[a]), rtol=1e-9, atol=1e-12)
        for agent in agents:
            agent.update(iteration=it)
        for agent in agents:
            agent.clear_utility()
            agent.clear_logs()
        auction.clear_revenue()
        for a, agent in enumerate(agents):
            kept = logs[a][-mems[a]:] if mems[a] else []
            assert len(agent.logs) == len(kept)
            assert [o.best_expected_value for o in agent.logs] == [o.best_expected_value for o in kept]
            np.testing.assert_allclose(agent.get_allocation_regret(), sum(o.best_expected_value - o.true_CTR * o.value for o in kept), rtol=1e-9, atol=1e-12)
            assert agent.net_utility == 0.0
