"""``Agent`` with the reference's constructor, fields and methods (src/Agent.py:8-129).

An agent is a descriptor until it is handed to an ``Auction``; from then on its utilities, regrets and
logs are views of the engine's per-(run, agent) accumulators and device log.  For an auction that holds
several runs the scalar properties return arrays of length ``num_runs``.
"""
import numpy as np

from . import _lib
from .impression import ImpressionOpportunity


class Agent:
    """An agent representing an advertiser (Agent.py:8-27)."""

    def __init__(self, rng, name, num_items, item_values, allocator, bidder, memory=0):
        self.rng = rng
        self.name = name
        self.num_items = num_items
        self.item_values = item_values
        self.allocator = allocator
        self.bidder = bidder
        self.memory = memory
        self._auction = None
        self._index = None

    # ------------------------------------------------------------------ engine plumbing
    def _attach(self, auction, index):
        self._auction, self._index = auction, index
        self.allocator._attach(auction, index)
        self.bidder._attach(auction, index)

    def _col(self, col):
        if self._auction is None or self._auction.engine is None:
            return 0.0
        v = self._auction.engine.acc[:, self._index, col].cpu().numpy()
        return float(v[0]) if len(v) == 1 else v

    # ------------------------------------------------------------------ reference surface
    @property
    def net_utility(self):  # Agent.py:20, :73
        return self._col(_lib.M_NET)

    @property
    def gross_utility(self):  # Agent.py:21, :74
        return self._col(_lib.M_GROSS)

    @property
    def logs(self):
        """List of ImpressionOpportunity of this agent for the current iteration (run 0), oldest first.
        Only available when the auction keeps the detailed log (always for simulate_opportunity())."""
        if self._auction is None:
            return []
        return self._auction._agent_logs(self._index)

    def select_item(self, context, _eps=None):
        """Agent.py:29-42 for one context: arg-max of estimated CTR x item value (first maximum); a Thompson-sampling
        allocator chooses on the sampled estimates and reports the MAP estimate of the chosen item."""
        from .allocators import OracleAllocator

        if isinstance(self.allocator, OracleAllocator):
            estim_CTRs = self.allocator.estimate_CTR(context)
        else:
            estim_CTRs = self.allocator.estimate_CTR(context, _eps=_eps)
        best_item = int(np.argmax(estim_CTRs * np.asarray(self.item_values)))
        if not isinstance(self.allocator, OracleAllocator) and self.allocator.thompson_sampling:
            estim_CTRs = self.allocator.estimate_CTR(context, sample=False)
        return best_item, estim_CTRs[best_item]

    def bid(self, context, _eps=None):
        """Agent.py:44-68 for one context: (bid, chosen item).  A query: the engine keeps its own per-round log when it
        simulates (agym_simulate_rounds), so nothing is appended to ``logs`` here."""
        best_item, estimated_CTR = self.select_item(context, _eps=_eps)
        value = self.item_values[best_item]
        bid = self.bidder.bid(value, context, estimated_CTR)
        return bid, best_item

    def update(self, iteration, plot=False, figsize=(8, 5), fontsize=14):
        """Agent.py:79-94: allocator.update on the won rows, bidder.update on all rows.  The engine fits
        every (run, agent) in one launch; the first agent.update() of an iteration triggers it."""
        if self._auction is not None:
            self._auction._update_models()

    def get_allocation_regret(self):  # Agent.py:96-98
        return self._col(_lib.M_ALLOC_REGRET)

    def get_estimation_regret(self):  # Agent.py:100-102
        return self._col(_lib.M_ESTIM_REGRET)

    def get_overbid_regret(self):  # Agent.py:104-106
        return self._col(_lib.M_OVERBID_REGRET)

    def get_underbid_regret(self):  # Agent.py:108-112
        return self._col(_lib.M_UNDERBID_REGRET)

    def get_CTR_RMSE(self):  # Agent.py:114-115
        with np.errstate(invalid="ignore", divide="ignore"):
            return np.sqrt(np.asarray(self._col(_lib.M_SQERR)) / np.asarray(self._col(_lib.M_NPART)))[()]

    def get_CTR_bias(self):  # Agent.py:117-118 (mean over won rows; NaN when the agent never won)
        with np.errstate(invalid="ignore", divide="ignore"):
            return (np.asarray(self._col(_lib.M_BIAS)) / np.asarray(self._col(_lib.M_NWON)))[()]

    def get_mean_best_expected_value(self):  # main.py:147
        with np.errstate(invalid="ignore", divide="ignore"):
            return (np.asarray(self._col(_lib.M_BEST_EV)) / np.asarray(self._col(_lib.M_NPART)))[()]

    def get_mean_gamma(self):  # main.py:142-145
        with np.errstate(invalid="ignore", divide="ignore"):
            return (np.asarray(self._col(_lib.M_GAMMA)) / np.asarray(self._col(_lib.M_NPART)))[()]

    def clear_utility(self):  # Agent.py:120-122
        if self._auction is not None and self._auction.engine is not None:
            self._auction.engine.acc[:, self._index, [_lib.M_NET, _lib.M_GROSS]] = 0.0

    def clear_logs(self):  # Agent.py:124-129
        if self._auction is not None:
            self._auction._clear_agent_logs(self._index)
        self.bidder.clear_logs(memory=self.memory)


def materialise_logs(cols, agent_index, D_ctx):
    """Turn SoA log columns (numpy, [T, P]) into the reference's list of records for one agent."""
    out = []
    t_idx, s_idx = np.nonzero(cols["agent"] == agent_index)
    for t, s in zip(t_idx, s_idx):
        out.append(ImpressionOpportunity(
            context=np.concatenate((cols["ctx"][t][:D_ctx[agent_index]], [1.0])), item=int(cols["item"][t, s]),
            value=float(cols["value"][t, s]), bid=float(cols["bid"][t, s]), best_expected_value=float(cols["best_ev"][t, s]),
            true_CTR=float(cols["true_ctr"][t, s]), estimated_CTR=float(cols["est"][t, s]), price=float(cols["price"][t, s]),
            second_price=float(cols["second"][t, s]), outcome=bool(cols["outcome"][t, s]), won=bool(cols["won"][t, s])))
    return out
