"""GPU: the less-travelled code paths -- other embedding sizes (generic DMAX = 32 kernels, K = 9 / 33 fit variants), ragged
catalogs, many participants, mixed allocators and bidders in one auction, the dense and the sparse fit kernels on the same
data, and the shared-memory overflow path of the fit (rows beyond the staged capacity) -- replayed against the oracle."""
import zlib

import numpy as np
import pytest

from oracle import auction_oracle as ao
from oracle import fit_oracle as fo
from tests import parity

pytestmark = pytest.mark.gpu


def _gpu():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from tests import gpu_util

    return gpu_util


def _case(seed, A, n_items, D, Do, P, mech, alloc, bid, sigma=0.1):
    rng = np.random.default_rng(seed)
    n_items = np.asarray(n_items if np.ndim(n_items) else [n_items] * A, np.int32)
    I = int(n_items.max())
    E, V = ao.make_catalog(rng, A, I, D, 0.8)
    bf = np.zeros((A, 4))
    bf[:, 0], bf[:, 1] = 0.9, sigma
    return {"A": A, "I": I, "D": D, "Do": Do, "P": P, "mechanism": mech, "embedding_var": 0.8, "n_items": n_items, "E": E, "V": V,
            "m": rng.standard_normal((A, I, Do + 1)).astype(np.float32), "q": (1 + 9 * rng.random((A, I, Do + 1))).astype(np.float32),
            "alloc_kind": np.asarray(alloc, np.int32), "bidder_kind": np.asarray(bid, np.int32), "bidder_f": bf}, rng


SHAPES = {
    "D12_Do9_P5": dict(A=9, n_items=[20, 3, 7, 40, 1, 12, 33, 5, 16], D=12, Do=9, P=5, mech=ao.MECH_FIRST,
                       alloc=[1, 0, 2, 1, 1, 0, 1, 2, 1], bid=[1, 0, 2, 0, 1, 1, 0, 2, 0]),
    "D20_Do20_P2": dict(A=4, n_items=9, D=20, Do=20, P=2, mech=ao.MECH_SECOND, alloc=[1, 1, 0, 2], bid=[0, 0, 0, 0]),
    "D3_Do1_P17": dict(A=20, n_items=6, D=3, Do=1, P=17, mech=ao.MECH_FIRST, alloc=[1] * 10 + [0] * 10, bid=[1] * 20),
    "A130_I100": dict(A=130, n_items=100, D=5, Do=4, P=2, mech=ao.MECH_SECOND, alloc=[1] * 130, bid=[0] * 130),
}


@pytest.mark.parametrize("name", list(SHAPES))
def test_odd_shapes_replay_and_fit(name):
    gu = _gpu()
    from auction_gym_b200 import _lib

    kw = SHAPES[name]
    case, rng = _case(zlib.crc32(name.encode()) % 1000, **kw)  # a fixed seed per shape (str hashes change per process)
    T = 700
    nz = ao.draw_replay_inputs(rng, T, case["A"], case["P"], case["D"], case["I"], case["Do"], 0.8,
                               want_eps=True, want_gamma=True)
    eng = gu.engine_from_case(case, R=1, precision=_lib.FP64)
    got = gu.log_to_numpy(gu.replay_case(eng, nz))
    rec, m = ao.simulate_rounds(case, nz["ctx"], nz["parts"], nz["u"], nz["ts_eps"], nz["gamma_z"])
    rep = parity.compare_rounds(got, rec, rec, rtol=parity.RTOL_F64, est_rtol=parity.RTOL_F32_EST, what=name)
    acc, rev = eng.metrics()
    if rep["near_tie_rounds"] == 0:
        # the signed accumulators (estimation error, regrets) are sums of up to T terms built on float32 CTR estimates that
        # agree to 4e-7 each and can cancel to ~0, so the meaningful bound on them is absolute: T * 4e-7 * |value| ~ 1e-5
        # (a 40-seed sweep of these shapes, tools/shape_sweep.py, saw at most 1.1e-7)
        np.testing.assert_allclose(acc[0], m["acc"], rtol=2e-6, atol=1e-5)
        np.testing.assert_allclose(rev[0], m["revenue"], rtol=2e-6)
    # allocator fit of every learnt agent on the rows the round loop logged, fixed epoch budget
    info = eng.update_allocators(max_epochs=150).cpu().numpy()[0]
    Do = case["Do"]
    obs = np.concatenate([nz["ctx"][:, :Do], np.ones((T, 1))], axis=1)
    m1, q1 = eng.m.cpu().numpy()[0], eng.q.cpu().numpy()[0]
    checked = 0
    for a in range(case["A"]):
        if case["alloc_kind"][a] == ao.ALLOC_ORACLE:
            continue
        # only compared when no arg-max flipped anywhere, so the device logged exactly the oracle's won rows
        t_idx, s_idx = np.nonzero((nz["parts"] == a) & (rec["won"] == 1))
        if rep["near_tie_rounds"] or len(t_idx) < 2 or checked >= 12:
            continue
        nI = int(case["n_items"][a])
        orc = fo.fit_allocator(obs[t_idx], rec["item"][t_idx, s_idx], rec["outcome"][t_idx, s_idx], case["m"][a, :nI], case["q"][a, :nI],
                               case["m"][a, :nI], max_epochs=150)
        assert info[a, 3] == len(t_idx) and info[a, 1] == 150
        np.testing.assert_allclose(m1[a, :nI], orc["m"], atol=3e-4, err_msg=f"{name} agent {a}")
        np.testing.assert_allclose(q1[a, :nI], orc["q"], rtol=2e-4, err_msg=f"{name} agent {a}")
        checked += 1
    assert checked > 0 or (case["alloc_kind"] == ao.ALLOC_ORACLE).all()
    eng.close()


@pytest.mark.parametrize("ncap", ["0.3", "1.5"])
@pytest.mark.parametrize("shape", [(64, 64, 6000), (6, 12, 3000)])  # sparse-kernel regime, dense-kernel regime
def test_fit_overflow_rows_and_both_kernels(shape, ncap):
    """The "fit_ncap" option shrinks the rows staged in shared memory so most rows take the global overflow path."""
    gu = _gpu()
    from auction_gym_b200 import _lib

    A, I, T = shape
    case, rng = _case(7, A=A, n_items=I, D=5, Do=4, P=2, mech=ao.MECH_SECOND, alloc=[1] * A, bid=[0] * A)
    nz = ao.draw_replay_inputs(rng, T, A, 2, 5, I, 4, 0.8, want_eps=True)
    eng = gu.engine_from_case(case, R=2, precision=_lib.FP64)
    eng.set_option("fit_ncap", float(ncap))
    eng.replay(np.stack([nz["ctx"]] * 2), np.stack([nz["parts"]] * 2), np.stack([nz["u"]] * 2), ts_eps=np.stack([nz["ts_eps"]] * 2))
    info = eng.update_allocators(max_epochs=120).cpu().numpy()
    rec, _ = ao.simulate_rounds(case, nz["ctx"], nz["parts"], nz["u"], nz["ts_eps"])
    obs = np.concatenate([nz["ctx"][:, :4], np.ones((T, 1))], axis=1)
    m1, q1 = eng.m.cpu().numpy(), eng.q.cpu().numpy()
    assert np.array_equal(m1[0], m1[1]) and np.array_equal(q1[0], q1[1])  # identical runs -> identical fits
    for a in range(0, A, max(1, A // 6)):
        t_idx, s_idx = np.nonzero((nz["parts"] == a) & (rec["won"] == 1))
        orc = fo.fit_allocator(obs[t_idx], rec["item"][t_idx, s_idx], rec["outcome"][t_idx, s_idx], case["m"][a], case["q"][a], case["m"][a],
                               max_epochs=120)
        assert info[0, a, 3] == len(t_idx)
        np.testing.assert_allclose(m1[0, a], orc["m"], atol=3e-4, err_msg=f"agent {a}")
        np.testing.assert_allclose(q1[0, a], orc["q"], rtol=2e-4, err_msg=f"agent {a}")
    eng.close()
