"""GPU: the bench workload at BASELINE.json's full single-GPU size (512 runs x 10 000 rounds x 64 agents x 64 items),
checked through size-independent properties of the domain instead of an oracle run (which would take CPU-hours):

  * every round has exactly one winner and P distinct participants: sum_a wins = T, sum_a participations = T * P per run
  * money is conserved: sum_a (gross - net) == revenue per run (winners pay exactly what the auctioneer books,
    Agent.py:72-74 vs Auction.py:74)
  * second price + truthful bidders: zero overbid and underbid regret
  * 0 <= click-through <= 1 in aggregate: gross utility <= wins * max item value
  * the allocator update only ever increases q (Models.py:45), keeps prev_iter_m == m and sigma == 1 / sqrt(q)
    (Models.py:47-48), and leaves items that logged no row bit-identical
  * the winner log has one valid record per round whose agent / item fields are in range
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_bench_shape_invariants():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import auction_gym_b200 as ag
    from auction_gym_b200 import _lib
    from oracle import auction_oracle as ao

    R, T, A, I, D, Do, P = 512, 10000, 64, 64, 5, 4, 2
    E, V = ao.make_catalog(np.random.default_rng(0), A, I, D)
    eng = ag.Engine(R=R, A=A, I=I, D=D, Do=Do, P=P, mechanism=_lib.SECOND_PRICE, E=E, V=V, n_items=[I] * A,
                    alloc_kind=[_lib.ALLOC_TS] * A, bidder_kind=[_lib.BID_TRUTHFUL] * A, rounds_capacity=T)
    m0 = torch.randn(R, A, I, Do + 1, generator=torch.Generator().manual_seed(3))
    eng.set_allocator_state(m0)
    eng.simulate(11, 0, T)
    acc, rev = eng.metrics()
    assert np.array_equal(acc[..., _lib.M_NWON].sum(axis=1), np.full(R, T))
    assert np.array_equal(acc[..., _lib.M_NPART].sum(axis=1), np.full(R, T * P))
    np.testing.assert_allclose((acc[..., _lib.M_GROSS] - acc[..., _lib.M_NET]).sum(axis=1), rev, rtol=1e-9)
    assert (acc[..., _lib.M_OVERBID_REGRET] == 0).all() and (acc[..., _lib.M_UNDERBID_REGRET] == 0).all()
    assert (acc[..., _lib.M_GROSS] <= acc[..., _lib.M_NWON] * V.max() + 1e-9).all() and (acc[..., _lib.M_GROSS] >= 0).all()
    assert (acc[..., _lib.M_ALLOC_REGRET] >= -1e-6).all()       # best expected value >= value of the chosen item
    # participation is uniform over agents: each agent is in P/A of the rounds
    part = acc[..., _lib.M_NPART].sum(axis=0) / (R * T * P)
    assert np.abs(part - 1 / A).max() < 5 * np.sqrt((1 / A) / (R * T * P))
    meta = eng.fit_meta.cpu().numpy().view(np.uint32)
    assert ((meta >> 31) == 1).all()
    assert ((meta >> 12) & 0xFFF).max() < A and (meta & 0xFFF).max() < I
    wins_from_log = np.stack([np.bincount((meta[r] >> 12) & 0xFFF, minlength=A) for r in range(0, R, 37)])
    assert np.array_equal(wins_from_log, acc[::37, :, _lib.M_NWON])
    # one allocator update (capped epochs: the properties do not depend on where Adam stops)
    q0 = eng.q.clone()
    info = eng.update_allocators(max_epochs=40).cpu().numpy()
    assert np.array_equal(info[..., 3], acc[..., _lib.M_NWON])   # rows per fit == wins of that agent
    assert (info[..., 1] == 40).all()
    m1, q1, mp1, s1 = eng.m, eng.q, eng.m_prev, eng.sigma
    assert bool((q1 >= q0).all()) and bool(torch.equal(mp1, m1))
    torch.testing.assert_close(s1, 1.0 / torch.sqrt(q1), rtol=1e-6, atol=0)
    run = 5
    used = np.zeros((A, I), bool)
    used[(meta[run] >> 12) & 0xFFF, meta[run] & 0xFFF] = True
    m1r, q1r = m1[run].cpu().numpy(), q1[run].cpu().numpy()
    assert np.array_equal(m1r[~used], m0[run].numpy()[~used]) and (q1r[~used] == 1).all()
    assert (np.abs(m1r[used] - m0[run].numpy()[used]).max(axis=1) > 0).all() and (q1r[used] > 1).any()
    eng.close()


def test_two_independent_fit_kernels_agree_at_the_bench_shape():
    """The warp-per-fit kernel (default) and the CTA kernel are separate implementations of the same optimiser.  On the
    bench shape with the reference's full epoch budget they must reach the same optimum: final losses within 5e-4 rel
    (items whose rows are all non-clicks have no finite optimum -- their intercept drifts until the stop rule fires),
    stop epochs within the +-2 % (+-25) spread of the reference's own trajectory, parameters
    within the fit bar (|dm| <= 1e-2 on the items that have rows), and each of them bit-reproducibly."""
    import torch

    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import auction_gym_b200 as ag
    from auction_gym_b200 import _lib
    from oracle import auction_oracle as ao

    R, T, A, I, D, Do, P = 8, 10000, 64, 64, 5, 4, 2
    E, V = ao.make_catalog(np.random.default_rng(0), A, I, D)
    m0 = torch.randn(R, A, I, Do + 1, generator=torch.Generator().manual_seed(5))
    out = {}
    for name, warp in (("warp", None), ("warp_again", None), ("cta", 0)):
        eng = ag.Engine(R=R, A=A, I=I, D=D, Do=Do, P=P, mechanism=_lib.SECOND_PRICE, E=E, V=V, n_items=[I] * A,
                        alloc_kind=[_lib.ALLOC_TS] * A, bidder_kind=[_lib.BID_TRUTHFUL] * A, rounds_capacity=T)
        if warp is not None:
            eng.set_option("fit_warp", warp)
        eng.set_allocator_state(m0)
        for it in range(2):  # second iteration: rows concentrated on few items per agent
            eng.clear_iteration()
            eng.simulate(3, it, T)
            info = eng.update_allocators().cpu().numpy()
            if it == 0:
                first = (eng.m.cpu().numpy().copy(), info.copy())
                if name == "cta":  # continue from the warp kernel's state so that the second iteration sees the same rows
                    eng.set_allocator_state(out["warp"]["m1"], out["warp"]["q1"])
                else:
                    out.setdefault(name, {})["m1"], out[name]["q1"] = eng.m.cpu().numpy().copy(), eng.q.cpu().numpy().copy()
        meta = eng.fit_meta.cpu().numpy().view(np.uint32)
        out.setdefault(name, {}).update(first=first, m=eng.m.cpu().numpy(), q=eng.q.cpu().numpy(), info=info, meta=meta)
        eng.close()
    w, w2, c = out["warp"], out["warp_again"], out["cta"]
    assert np.array_equal(w["m"], w2["m"]) and np.array_equal(w["q"], w2["q"]) and np.array_equal(w["info"], w2["info"], equal_nan=True)
    for a_, b_ in ((w["first"], c["first"]), ((w["m"], w["info"]), (c["m"], c["info"]))):
        (ma, ia), (mb, ib) = a_, b_
        assert np.array_equal(ia[..., 3], ib[..., 3])                                    # same rows per fit
        np.testing.assert_allclose(ia[..., 2], ib[..., 2], rtol=5e-4)                     # same optimum
        assert (np.abs(ia[..., 0] - ib[..., 0]) <= np.maximum(25, 0.02 * ib[..., 0])).mean() > 0.97   # stop epochs
        assert np.abs(ma - mb).max() <= 2e-2 and np.mean(np.abs(ma - mb) > 1e-2) < 1e-4
    assert np.array_equal(w["meta"], c["meta"])  # the second iteration really saw the same rows
