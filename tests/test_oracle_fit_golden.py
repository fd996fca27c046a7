"""CPU pins of the fit oracles (oracle/fit_oracle.py) against outputs of the UNMODIFIED reference
(tests/golden/fit_*.npz, bidfit_winrate.npz: written by oracle/make_golden.py / make_golden_policy.py driving
/root/reference/src/BidderAllocation.py:29-65 and src/Bidder.py:210-260 under oracle/ref_harness.py).

Tolerances are the reference's own reproducibility floor (permuting its training rows moves the
fitted m by 8.8e-4 and the stop epoch by 5-11; an independent float32 restatement lands within +-18 epochs, |dm| 6e-3,
q 6e-4): |dm| <= 1e-2, q <= 1e-3 rel, stop epoch +-1 % with a floor of 20 epochs.
"""
import os

import numpy as np
import pytest

from oracle import auction_oracle as ao
from oracle import fit_oracle as fo
from tests.conftest import GOLDEN_DIR


def _cases():
    out = []
    for name in ("fit_ref_shape", "fit_64x64"):
        z = np.load(f"{GOLDEN_DIR}/{name}.npz")
        for it in (0, 1):
            for a in z["fit_agents"]:
                out.append((name, it, int(a)))
    return out


@pytest.mark.parametrize("name,it,agent", _cases())
def test_allocator_fit_oracle_matches_the_reference(name, it, agent):
    z = np.load(f"{GOLDEN_DIR}/{name}.npz")
    p = f"it{it}_a{agent}_"
    orc = fo.fit_allocator(z[p + "X"], z[p + "items"], z[p + "y"], z[p + "m0"], z[p + "q0"], z[p + "m_prev"])
    ref_stop = int(z[p + "stop_epoch"])
    what = f"{name} it{it} agent {agent}: stop oracle {orc['stop_epoch']} / reference {ref_stop}"
    assert abs(orc["stop_epoch"] - ref_stop) <= max(20, 0.01 * ref_stop), what
    assert orc["n_epochs"] == orc["stop_epoch"] + 1
    np.testing.assert_allclose(orc["m"], z[p + "m1"], atol=1e-2, rtol=0, err_msg=what)
    np.testing.assert_allclose(orc["q"], z[p + "q1"], rtol=1e-3, err_msg=what)
    np.testing.assert_allclose(orc["final_loss"], z[p + "losses_tail"][-1], rtol=1e-4, err_msg=what)
    # items without rows: untouched bit for bit (Adam sees a zero gradient, the Laplace sum is empty)
    unused = np.setdiff1d(np.arange(z[p + "m0"].shape[0]), np.unique(z[p + "items"]))
    assert np.array_equal(orc["m"][unused], z[p + "m0"][unused]) and np.array_equal(orc["q"][unused], z[p + "q0"][unused])


def test_allocator_fit_oracle_first_losses_match_the_reference():
    """The first epochs are not chaotic yet: the loss trajectory must agree to float32 rounding."""
    z = np.load(f"{GOLDEN_DIR}/fit_64x64.npz")
    p = f"it0_a{int(z['fit_agents'][0])}_"
    head = z[p + "losses_head"]
    orc = fo.fit_allocator(z[p + "X"], z[p + "items"], z[p + "y"], z[p + "m0"], z[p + "q0"], z[p + "m_prev"], max_epochs=len(head),
                           return_losses=True)
    np.testing.assert_allclose(orc["losses"], head, rtol=2e-6)


@pytest.mark.parametrize("agent", [0, 1, 2, 3, 4, 5])
def test_winrate_fit_oracle_matches_the_reference(agent):
    z = np.load(f"{GOLDEN_DIR}/bidfit_winrate.npz")
    a = agent
    orc = fo.fit_winrate(z[f"a{a}_est"], z[f"a{a}_value"], z[f"a{a}_gamma"], z[f"a{a}_won"], z[f"a{a}_w0"])
    ref_w, ref_stop = z[f"a{a}_w1"], int(z[f"a{a}_stop_epoch"])
    what = f"agent {a}: stop oracle {orc['stop_epoch']} / reference {ref_stop}, w {orc['w']} vs {ref_w}"
    if ref_stop >= 0:
        assert abs(orc["stop_epoch"] - ref_stop) <= max(8, 0.015 * ref_stop), what
    else:
        assert orc["stop_epoch"] == -1 and orc["n_epochs"] == 8192 * 4, what
    np.testing.assert_allclose(orc["w"], ref_w, atol=1e-2, rtol=0, err_msg=what)
    g = np.linspace(0.1, 1.0, 64)
    x = np.stack([np.full(64, 0.12), np.full(64, 1.1), g], axis=1).astype(np.float32)
    np.testing.assert_allclose(ao.winrate32(orc["w"], x), ao.winrate32(ref_w, x), atol=2e-3, err_msg=what)


@pytest.mark.parametrize("name", ["fit_ref_shape", "fit_64x64"])
def test_newton_restatement_reaches_the_optimum_of_its_objective(name):
    """oracle/fit_oracle.py::fit_allocator_newton restates the OPT-IN mode AGYM_FIT_NEWTON, which is not a reference algorithm:
    the check is optimality on the reference-run fit inputs (gradient ~ 0, objective not above the reference's end point)."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    for a in z["fit_agents"]:
        p = f"it1_a{int(a)}_"
        X, items, y = z[p + "X"], z[p + "items"], z[p + "y"]
        r = fo.fit_allocator_newton(X, items, y, z[p + "m0"], z[p + "q0"], z[p + "m_prev"])
        obj, grad = fo.allocator_objective(X, items, y, r["m"], z[p + "q0"], z[p + "m_prev"], prior_on_intercept=True)
        obj_ref, _ = fo.allocator_objective(X, items, y, z[p + "m1"], z[p + "q0"], z[p + "m_prev"], prior_on_intercept=True)
        assert np.abs(grad).max() < 5e-4 and obj <= obj_ref and r["passes"].max() <= 50, (name, p)
        unused = np.setdiff1d(np.arange(len(z[p + "m0"])), np.unique(items))
        assert np.array_equal(r["m"][unused], z[p + "m0"][unused]) and np.array_equal(r["q"][unused], z[p + "q0"][unused])
