"""TEST INFRASTRUCTURE ONLY -- CPU restatement of log retention across iterations (``Agent(memory=...)``).

Part of ``oracle/`` (the checker): only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import
it; the product path (``auction_gym_b200/``) never does.

What the reference does (cited lines are of the unmodified reference):
  * ``Agent.clear_logs`` (src/Agent.py:124-129): ``self.logs = self.logs[-self.memory:]`` when ``memory`` is set, else
    ``[]``; the bidder's ``gammas`` / ``propensities`` lists are cut the same way (src/Bidder.py:149-153,327-333,
    433-439,617-623);
  * ``Agent.clear_utility`` (src/Agent.py:120-122) zeroes net / gross utility regardless of ``memory``;
  * everything that reads ``self.logs`` in the next iteration therefore sees the kept records first: the metric getters
    (src/Agent.py:96-118, src/main.py:142-148), ``allocator.update`` on the won rows and ``bidder.update`` on all rows
    (src/Agent.py:79-94).

Pinned against the unmodified reference by ``oracle/make_golden_retention.py`` -> ``tests/golden/retention.npz``
(``tests/test_oracle_golden.py``).
"""
import numpy as np

from . import auction_oracle as ao

FIELDS = ("item", "est", "value", "bid", "true_ctr", "best_ev", "price", "second", "outcome", "won", "gamma", "propensity")


def agent_records(rec, parts, ctx_obs, a):
    """This iteration's records of agent ``a`` in time order: dict of 1-D arrays + ``ctx`` [n, Do]."""
    t_idx, s_idx = np.nonzero(parts == a)  # row-major = time order; an agent holds at most one slot per round
    out = {k: np.asarray(rec[k])[t_idx, s_idx] for k in FIELDS}
    out["ctx"] = ctx_obs[t_idx]
    return out


def concat(old, new):
    if old is None:
        return new
    return {k: np.concatenate([old[k], new[k]], axis=0) for k in new}


def keep_last(records, memory):
    """Agent.py:124-129."""
    if not memory:
        return None
    return {k: v[-memory:] for k, v in records.items()}


def metric_sums(records, net=0.0, gross=0.0):
    """The 12 accumulators over one agent's records (same formulas as auction_oracle.accumulate_metrics)."""
    acc = np.zeros(ao.NUM_METRICS)
    r = records
    won = r["won"].astype(bool)
    tv = r["true_ctr"] * r["value"]
    acc[ao.M_NET], acc[ao.M_GROSS] = net, gross
    acc[ao.M_ALLOC_REG] = np.sum(r["best_ev"] - tv)
    acc[ao.M_ESTIM_REG] = np.sum(r["est"] * r["value"] - tv)
    acc[ao.M_OVERBID] = np.sum(np.where(won, r["price"] - r["second"], 0.0))
    acc[ao.M_UNDERBID] = np.sum(np.where(~won & (r["price"] < tv), r["price"] - r["bid"], 0.0))
    acc[ao.M_SQERR] = np.sum((r["true_ctr"] - r["est"]) ** 2)
    acc[ao.M_BIAS] = np.sum(np.where(won, r["est"] / r["true_ctr"], 0.0))
    acc[ao.M_NPART] = len(won)
    acc[ao.M_NWON] = won.sum()
    acc[ao.M_BEST_EV] = np.sum(r["best_ev"])
    acc[ao.M_GAMMA] = np.sum(np.nan_to_num(r["gamma"], nan=0.0))
    return acc


def simulate_iterations(case, inputs, memory):
    """Run ``len(inputs)`` iterations with fixed model state (no fits) and log retention.

    inputs: list of replay-input dicts (ctx, parts, u, optional ts_eps / gamma_z / grid_u), one per iteration.
    Returns a list (one entry per iteration) of dicts:
      rec      the round records of that iteration alone (auction_oracle.simulate_rounds)
      logs     per agent: the records the reference's ``agent.logs`` holds at the end of the iteration (kept + new)
      acc      [A, 12] accumulators as the getters read them at the end of the iteration
      revenue  scalar
    """
    A, Do = int(case["A"]), int(case["Do"])
    kept = [None] * A
    out = []
    for nz in inputs:
        rec, met = ao.simulate_rounds(case, nz["ctx"], nz["parts"], nz["u"], nz.get("ts_eps"), nz.get("gamma_z"), nz.get("grid_u"))
        ctx_obs = np.asarray(nz["ctx"])[:, :Do]
        logs, acc = [], np.zeros((A, ao.NUM_METRICS))
        for a in range(A):
            cur = concat(kept[a], agent_records(rec, nz["parts"], ctx_obs, a))
            logs.append(cur)
            acc[a] = metric_sums(cur, met["acc"][a, ao.M_NET], met["acc"][a, ao.M_GROSS])  # utilities restart at 0 (Agent.py:120-122)
            kept[a] = keep_last(cur, int(memory[a]))
        out.append({"rec": rec, "logs": logs, "acc": acc, "revenue": met["revenue"]})
    return out
