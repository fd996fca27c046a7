"""GPU: production mode (in-kernel Philox noise) -- determinism, chunking, run sharding, the staged
pipeline against the fused kernel, and distribution-level parity with the oracle (the numpy PCG64
stream cannot be reproduced in parallel, SURVEY.md section 7 hard part 4, so exactness lives in replay
mode and production mode is checked statistically)."""
import numpy as np
import pytest

from oracle import auction_oracle as ao
from tests.conftest import load_golden

pytestmark = pytest.mark.gpu

FIELDS = ("agent", "item", "winner", "outcome", "won", "price", "bid", "est", "true_ctr", "ctx")


def _gpu():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from tests import gpu_util

    return gpu_util


def _np(out):
    return {k: v.cpu().numpy() for k, v in out.items()}


@pytest.mark.parametrize("name", ["rounds_sp_ts", "rounds_fp_gauss", "rounds_fp_pA", "rounds_sp_ragged"])
def test_deterministic_chunked_and_sharded(name):
    gu = _gpu()
    from auction_gym_b200 import _lib

    case, *_ = load_golden(name)
    R, T, seed = 4, 600, 1234
    full = gu.engine_from_case(case, R=R, precision=_lib.FP32)
    a = _np(full.simulate(seed, 3, T, FIELDS))
    acc_a, rev_a = full.metrics()
    # same call again -> identical
    full.clear_iteration()
    b = _np(full.simulate(seed, 3, T, FIELDS))
    for k in FIELDS:
        assert np.array_equal(a[k], b[k], equal_nan=True), k
    # two half-iterations == one full (the Philox counter is the round index inside the iteration)
    full.clear_iteration()
    c1 = _np(full.simulate(seed, 3, 250, FIELDS))
    c2 = _np(full.simulate(seed, 3, T - 250, FIELDS))
    for k in FIELDS:
        assert np.array_equal(a[k], np.concatenate([c1[k], c2[k]], axis=1), equal_nan=True), k
    acc_c, rev_c = full.metrics()
    np.testing.assert_allclose(acc_c, acc_a, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(rev_c, rev_a, rtol=1e-12)
    # runs sharded over two "ranks" reproduce the single-device job (run_offset keys the RNG)
    for off in (0, 2):
        sh = gu.engine_from_case(case, R=2, precision=_lib.FP32, run_offset=off)
        s = _np(sh.simulate(seed, 3, T, FIELDS))
        for k in FIELDS:
            assert np.array_equal(a[k][off:off + 2], s[k], equal_nan=True), (k, off)
        acc_s, rev_s = sh.metrics()
        np.testing.assert_allclose(acc_s, acc_a[off:off + 2], rtol=1e-12, atol=1e-12)
        sh.close()
    # a different seed / iteration gives a different stream
    full.clear_iteration()
    d = _np(full.simulate(seed + 1, 3, T, FIELDS))
    assert not np.array_equal(a["agent"], d["agent"])
    full.close()


@pytest.mark.parametrize("name,T", [("rounds_sp_ts", 512), ("rounds_fp_gauss", 512), ("rounds_sp_oracle_64x64", 512), ("rounds_sp_p3", 512),
                                    ("rounds_fp_ts_gauss", 512), ("rounds_fp_gauss", 2052), ("rounds_sp_ts", 1536), ("rounds_sp_oracle_64x64", 4000)])
def test_staged_pipeline_equals_fused(name, T):
    """K1 -> K2 -> K3 -> K4 through HBM reproduces the fused FP32 kernel exactly (same Philox counters)."""
    gu = _gpu()
    from auction_gym_b200 import _lib

    case, *_ = load_golden(name)
    R, seed = 3, 99
    eng = gu.engine_from_case(case, R=R, precision=_lib.FP32)
    f = _np(eng.simulate(seed, 1, T, FIELDS + ("value", "best_ev", "second", "gamma")))
    acc_f, rev_f = eng.metrics()
    eng.clear_iteration()
    s = _np(eng.staged_round(seed, 1, T))
    acc_s, rev_s = eng.metrics()
    sh = lambda x: x.reshape(R, T, *x.shape[1:])  # noqa: E731
    assert np.array_equal(sh(s["parts"]), f["agent"])
    np.testing.assert_array_equal(sh(s["ctx"]).astype(np.float64), f["ctx"])
    assert np.array_equal(sh(s["item"]), f["item"])
    for k in ("est", "true_ctr", "best_ev", "value", "bid"):
        np.testing.assert_array_equal(sh(s[k]).astype(np.float64), f[k], err_msg=k)
    assert np.array_equal(sh(s["winner"]), f["winner"])
    np.testing.assert_array_equal(sh(s["price"]).astype(np.float64), f["price"][:, :, 0])
    assert np.array_equal(sh(s["outcome"]), f["outcome"].max(axis=2))
    cols = [c for c in range(_lib.NUM_METRICS) if c != _lib.M_BIAS]  # K4's contract carries no estimate
    np.testing.assert_allclose(acc_s[..., cols], acc_f[..., cols], rtol=5e-6, atol=1e-5)  # P == 2: K4 sums float partials
    np.testing.assert_allclose(rev_s, rev_f, rtol=1e-6)
    eng.close()


def test_participants_and_context_distribution():
    gu = _gpu()
    from auction_gym_b200 import _lib

    case, *_ = load_golden("rounds_sp_p3")  # A = 6, P = 3
    R, T = 8, 20000
    eng = gu.engine_from_case(case, R=R, precision=_lib.FP32)
    o = _np(eng.simulate(7, 0, T, ("agent", "ctx")))
    ag = o["agent"].reshape(-1, 3)
    assert (np.sort(ag, axis=1)[:, 1:] != np.sort(ag, axis=1)[:, :-1]).all(), "participants must be distinct"
    n = ag.shape[0]
    for s in range(3):  # every slot uniform over agents
        cnt = np.bincount(ag[:, s], minlength=6)
        assert np.abs(cnt / n - 1 / 6).max() < 5 * np.sqrt((1 / 6) * (5 / 6) / n)
    pair = np.bincount(ag[:, 0] * 6 + ag[:, 1], minlength=36).reshape(6, 6)
    assert np.diag(pair).sum() == 0
    off = pair[~np.eye(6, dtype=bool)] / n
    assert np.abs(off - 1 / 30).max() < 5 * np.sqrt((1 / 30) * (29 / 30) / n)
    ctx = o["ctx"].reshape(-1, eng.D)
    assert np.abs(ctx.mean(axis=0)).max() < 5 / np.sqrt(n)
    assert np.abs(ctx.std(axis=0) - 1.0).max() < 0.01
    assert np.abs(np.corrcoef(ctx.T) - np.eye(eng.D)).max() < 0.01
    # tails: Box-Muller on 24-bit uniforms must still reach |z| > 4
    assert (np.abs(ctx) > 4.0).mean() == pytest.approx(6.3e-5, rel=0.5)
    eng.close()


@pytest.mark.parametrize("name,precision", [("rounds_sp_oracle", 0), ("rounds_sp_oracle", 1), ("rounds_sp_ts_q", 0), ("rounds_fp_gauss", 0)])
def test_statistical_parity_with_oracle(name, precision):
    """Per-round revenue / welfare / surplus / regrets of the production kernel vs the oracle driven by
    numpy's own generator: means must agree within 5 standard errors over 64 runs x 3000 rounds."""
    gu = _gpu()
    case, *_ = load_golden(name)
    R, T = 64, 3000
    eng = gu.engine_from_case(case, R=R, precision=precision)
    eng.simulate(2024, 0, T)
    acc, rev = eng.metrics()
    rng = np.random.default_rng(77)
    ref_acc, ref_rev = [], []
    learnt = bool((case["alloc_kind"] == ao.ALLOC_TS).any())
    shaded = bool((case["bidder_kind"] != ao.BID_TRUTHFUL).any())
    for r in range(R):
        nz = ao.draw_replay_inputs(rng, T, int(case["A"]), int(case["P"]), int(case["D"]), int(case["I"]), int(case["Do"]),
                                   float(case["embedding_var"]), want_eps=learnt, want_gamma=shaded)
        _, m = ao.simulate_rounds(case, nz["ctx"], nz["parts"], nz["u"], nz.get("ts_eps"), nz.get("gamma_z"))
        ref_acc.append(m["acc"])
        ref_rev.append(m["revenue"])
    ref_acc, ref_rev = np.stack(ref_acc), np.asarray(ref_rev)

    def close(x, y, what):
        se = np.sqrt(x.var(ddof=1) / len(x) + y.var(ddof=1) / len(y))
        assert abs(x.mean() - y.mean()) <= 5 * se + 1e-12, f"{what}: {x.mean()} vs {y.mean()} (se {se})"

    close(rev, ref_rev, "revenue")
    for col, nm in ((ao.M_NET, "surplus"), (ao.M_GROSS, "welfare"), (ao.M_ALLOC_REG, "allocation regret"),
                    (ao.M_ESTIM_REG, "estimation regret"), (ao.M_OVERBID, "overbid regret"), (ao.M_UNDERBID, "underbid regret"),
                    (ao.M_SQERR, "sq err"), (ao.M_BEST_EV, "best ev"), (ao.M_NWON, "wins"), (ao.M_GAMMA, "gamma")):
        close(acc[:, :, col].sum(axis=1), ref_acc[:, :, col].sum(axis=1), nm)
    # per-agent welfare too (agent identity matters, not only totals)
    for a in range(int(case["A"])):
        close(acc[:, a, ao.M_GROSS], ref_acc[:, a, ao.M_GROSS], f"welfare of agent {a}")
    eng.close()


def test_device_philox_matches_restatement():
    """The in-kernel Philox4x32-10 + Box-Muller against oracle/philox_oracle.py (itself pinned to Random123's known-answer
    vectors): contexts of the staged K1 kernel, value for value."""
    gu = _gpu()
    from auction_gym_b200 import _lib
    from oracle import philox_oracle as ph

    case, *_ = load_golden("rounds_sp_oracle")
    R, T, seed, it = 3, 64, 0x1234567890ABCDEF, 5
    eng = gu.engine_from_case(case, R=R, precision=_lib.FP32, run_offset=7)
    b = eng.staged_round(seed, it, T, accumulate=False)
    ctx = b["ctx"].cpu().numpy().reshape(R, T, eng.D)
    for r in range(R):
        key = ph.make_key(seed, 7 + r)
        t = np.arange(T, dtype=np.uint32)
        w0 = ph.philox4x32_10(t, np.uint32(it), np.uint32(0), np.uint32(0), key)   # purpose 0 (context), block 0: components 0..3
        w1 = ph.philox4x32_10(t, np.uint32(it), np.uint32(0), np.uint32(1), key)   # block 1: component 4
        n0, n1 = ph.box_muller(w0[0], w0[1])
        n2, n3 = ph.box_muller(w0[2], w0[3])
        n4, _ = ph.box_muller(w1[0], w1[1])
        want = np.stack([n0, n1, n2, n3, n4], axis=1)
        np.testing.assert_allclose(ctx[r], want, rtol=2e-5, atol=2e-6)
    eng.close()


def test_packed_state_follows_every_write_of_the_learnt_state():
    """Production mode of the standard shape reads the library's packed copy of {m, 1/q} (DESIGN.md, round loop).  The copy must
    follow the learnt state through both doors: a host write (set_allocator_state -> agym_refresh_sigma) and the allocator update
    (agym_update_allocators) -- an engine whose state went through them simulates exactly what a fresh engine holding the same
    state simulates."""
    gu = _gpu()
    from auction_gym_b200 import _lib

    case, *_ = load_golden("rounds_sp_ts_64x64")
    R, T, seed = 2, 400, 77
    rng = np.random.default_rng(5)
    eng = gu.engine_from_case(case, R=R, precision=_lib.FP32, rounds_capacity=T)
    a = _np(eng.simulate(seed, 0, T, FIELDS))
    m2 = (np.broadcast_to(case["m"], (R,) + case["m"].shape) + 0.7 * rng.standard_normal((R,) + case["m"].shape)).astype(np.float32)
    q2 = (1.0 + 3.0 * rng.random((R,) + case["m"].shape)).astype(np.float32)
    eng.clear_iteration()
    eng.set_allocator_state(m2, q2)
    b = _np(eng.simulate(seed, 0, T, FIELDS))
    assert not np.array_equal(a["item"], b["item"])  # the new state is the one that bids
    fresh = gu.engine_from_case(case, R=R, precision=_lib.FP32, rounds_capacity=T)
    fresh.set_allocator_state(m2, q2)
    c = _np(fresh.simulate(seed, 0, T, FIELDS))
    for k in FIELDS:
        assert np.array_equal(b[k], c[k], equal_nan=True), k
    # ... and through the update: iteration 1 from the fitted state
    eng.update_allocators(max_epochs=60, want_info=False)
    m3, q3 = eng.m.cpu().numpy().copy(), eng.q.cpu().numpy().copy()
    assert not np.array_equal(m3, m2)
    eng.clear_iteration()
    d = _np(eng.simulate(seed, 1, T, FIELDS))
    fresh.clear_iteration()
    fresh.set_allocator_state(m3, q3)
    e = _np(fresh.simulate(seed, 1, T, FIELDS))
    for k in FIELDS:
        assert np.array_equal(d[k], e[k], equal_nan=True), k
    eng.close()
    fresh.close()
