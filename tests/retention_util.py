"""Loader and comparison helpers for tests/golden/retention.npz (written by oracle/make_golden_retention.py from the
unmodified reference with Agent(memory=...), 4 iterations, model state held fixed)."""
import numpy as np

from oracle import auction_oracle as ao
from tests.conftest import GOLDEN_DIR

REC_KEYS = ("item", "est", "value", "bid", "true_ctr", "best_ev", "price", "second", "outcome", "won", "gamma", "propensity")
ROW_KEYS = ("fit_ctx", "fit_items", "fit_y", "values", "bids", "prices", "outcomes", "ests", "won", "gammas", "propensities", "mean_gamma", "kept")


def load_retention():
    z = np.load(f"{GOLDEN_DIR}/retention.npz", allow_pickle=False)
    case = {k[5:]: (z[k].item() if z[k].ndim == 0 else z[k]) for k in z.files if k.startswith("case_")}
    n_iter, T = int(z["n_iter"]), int(z["t_iter"])
    memory = z["memory"]
    inputs, ref = [], []
    for it in range(n_iter):
        sl = slice(it * T, (it + 1) * T)
        inputs.append({k[3:]: z[k][sl] for k in z.files if k.startswith("in_")})
        e = {"rec": {k: z[f"it{it}_ref_{k}"] for k in REC_KEYS},
             "met": {k[len(f"it{it}_met_"):]: z[k] for k in z.files if k.startswith(f"it{it}_met_")},
             "agents": [{k: z[f"it{it}_a{a}_{k}"] for k in ROW_KEYS} for a in range(int(case["A"]))]}
        ref.append(e)
    return case, memory, inputs, ref


def check_logs_against_reference(logs, ref_agents, case, est_rtol, what=""):
    """logs: per agent dict of 1-D arrays (+ ctx [n, Do]) = agent.logs at the end of an iteration (kept + new);
    ref_agents: what the reference's Agent.update handed to allocator.update / bidder.update in that iteration."""
    Do = int(case["Do"])
    for a, (lg, rf) in enumerate(zip(logs, ref_agents)):
        w = f"{what} agent {a}"
        won = lg["won"].astype(bool)
        assert len(won) == len(rf["won"]), w
        assert np.array_equal(won, rf["won"].astype(bool)), w
        # bidder.update sees every record (Agent.py:94)
        np.testing.assert_allclose(lg["value"], rf["values"], rtol=1e-12, err_msg=w)
        np.testing.assert_allclose(lg["est"], rf["ests"], rtol=est_rtol, err_msg=w)
        np.testing.assert_allclose(lg["price"], rf["prices"], rtol=est_rtol, atol=1e-12, err_msg=w)
        assert np.array_equal(lg["outcome"].astype(bool), rf["outcomes"].astype(bool)), w
        if len(rf["gammas"]):  # shaded bidders keep gammas / propensities in step with the logs (Bidder.py:149-153,327-333)
            np.testing.assert_allclose(lg["gamma"], rf["gammas"], rtol=1e-9, err_msg=w)
        if len(rf["propensities"]):
            np.testing.assert_allclose(lg["propensity"], rf["propensities"], rtol=1e-6, err_msg=w)
        # allocator.update sees the won records (Agent.py:91)
        assert np.array_equal(lg["item"][won], rf["fit_items"]), w
        assert np.array_equal(lg["outcome"][won].astype(bool), rf["fit_y"].astype(bool)), w
        if case["alloc_kind"][a] != ao.ALLOC_ORACLE and won.any():
            np.testing.assert_allclose(lg["ctx"][won], rf["fit_ctx"][:, :Do], rtol=1e-6, err_msg=w)
