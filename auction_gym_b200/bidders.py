"""Bidders -- same class names and constructor signatures as the reference's src/Bidder.py.

Like the allocators these are descriptors: ``bid`` for a whole batch of opportunities is computed by the
engine (agym_simulate_rounds / agym_k3_bids); the per-(run, agent) state (prev_gamma, gamma_sigma,
``model_initialised``, win-rate and policy weights) lives in the engine's ``bidder_d`` / ``bidder_w``.
"""
import numpy as np

from . import _lib


class Bidder:
    """Bidder base class (Bidder.py:15-25)."""

    kind = _lib.BID_TRUTHFUL
    needs_fit = False   # update() does something in the reference
    fit_built = False   # ... and the engine has that update (agym_update_bidders)
    fit_kind = _lib.BFIT_NONE

    def __init__(self, rng):
        self.rng = rng
        self.truthful = False  # Bidder.py:19
        self._auction = None
        self._index = None

    def _attach(self, auction, index):
        self._auction, self._index = auction, index

    def update(self, contexts, values, bids, prices, outcomes, estimated_CTRs, won_mask, iteration, plot, figsize, fontsize, name):
        if self._auction is not None:
            self._auction._update_models()

    def bid(self, value, context, estimated_CTR):
        """Bidder.bid (Bidder.py:34-35,47-58,171-208,348-367,455-475) for one opportunity: a T = 1 launch of the staged bid
        kernel (agym_k3_bids) with this agent's current state (shading noise from Philox, keyed by the auction seed, the
        iteration and a query counter)."""
        au = self._auction
        if au is None:
            raise RuntimeError("bid needs the agent to be part of an Auction (the bidder state lives in its engine)")
        if au.engine is None:
            au._build()
        au._queries += 1
        b, g, p = au.engine.bid_one(self._index, estimated_CTR, value, seed=au.seed + 0x9E3779B97F4A7C15 * au._queries, iteration=au.iteration)
        self._last_gamma, self._last_propensity = g, p
        return b

    def clear_logs(self, memory):
        pass

    # state the engine needs at bid time: (prev_gamma, gamma_sigma)
    def _gamma_params(self):
        return 1.0, 0.0

    @property
    def gammas(self):
        """Shading factors logged this iteration (Bidder.py:44,164,342,448); needs the detailed log."""
        if self._auction is None:
            return []
        return self._auction._agent_log_column(self._index, "gamma")

    @property
    def propensities(self):
        if self._auction is None:
            return []
        return self._auction._agent_log_column(self._index, "propensity")


class TruthfulBidder(Bidder):
    """bid = value * estimated CTR (Bidder.py:28-35)."""

    kind = _lib.BID_TRUTHFUL

    def __init__(self, rng):
        super().__init__(rng)
        self.truthful = True

    def bid(self, value, context, estimated_CTR):
        return value * estimated_CTR


class _ShadedBidder(Bidder):
    def __init__(self, rng, gamma_sigma, init_gamma=1.0):
        super().__init__(rng)
        self.gamma_sigma = float(gamma_sigma)
        self.prev_gamma = float(init_gamma)

    def _gamma_params(self):
        return self.prev_gamma, self.gamma_sigma

    @property
    def model_initialised(self):  # Bidder.py:168,345,452 -- per run once attached
        if self._auction is None or self._auction.engine is None:
            return False
        v = self._auction.engine.bidder_d[:, self._index, 2].cpu().numpy() != 0
        return bool(v[0]) if len(v) == 1 else v


class EmpiricalShadedBidder(_ShadedBidder):
    """One global gamma ~ N(prev_gamma, sigma) clipped to [0, 1] (Bidder.py:38-58).  Its update
    (Bidder.py:60-125: bucketised search for the gamma with the best lower confidence bound) runs in agym_update_bidders."""

    kind = _lib.BID_GAUSS_CLIP
    needs_fit = True
    fit_built = True
    fit_kind = _lib.BFIT_EMPIRICAL


class ValueLearningBidder(_ShadedBidder):
    """Win-rate model + grid search / learnt policy over gamma (Bidder.py:156-333)."""

    needs_fit = True

    def __init__(self, rng, gamma_sigma, init_gamma=1.0, inference="search"):
        assert inference in ["search", "policy"]
        super().__init__(rng, gamma_sigma, init_gamma)
        self.inference = inference
        self.kind = _lib.BID_SEARCH if inference == "search" else _lib.BID_POLICY
        self.fit_built = True
        self.fit_kind = _lib.BFIT_VL_SEARCH if inference == "search" else _lib.BFIT_VL_POLICY



class PolicyLearningBidder(_ShadedBidder):
    """Contextual-bandit policy over gamma trained with REINFORCE / TRPO / PPO losses (Bidder.py:336-439)."""

    kind = _lib.BID_BANDIT
    needs_fit = True

    fit_built = True
    _LOSSES = {"REINFORCE": _lib.BFIT_PL_REINFORCE, "REINFORCE_offpolicy": _lib.BFIT_PL_OFFPOLICY, "TRPO": _lib.BFIT_PL_TRPO,
               "PPO": _lib.BFIT_PL_PPO}

    def __init__(self, rng, gamma_sigma, loss, init_gamma=1.0):
        super().__init__(rng, gamma_sigma, init_gamma)
        if loss not in self._LOSSES:
            raise ValueError(f"unknown PolicyLearningBidder loss {loss!r} (Models.py:173-196: {sorted(self._LOSSES)})")
        self.loss = loss
        self.fit_kind = self._LOSSES[loss]


class DoublyRobustBidder(_ShadedBidder):
    """Win-rate model + doubly-robust policy learning (Bidder.py:442-623)."""

    kind = _lib.BID_BANDIT
    needs_fit = True

    fit_built = True
    fit_kind = _lib.BFIT_DR

    def __init__(self, rng, gamma_sigma, init_gamma=1.0):
        super().__init__(rng, gamma_sigma, init_gamma)


def gaussian_propensity(prev_gamma, sigma, gamma):
    """Bidder.py:178 -- density of the logging policy at the drawn gamma."""
    return np.exp(-((prev_gamma - gamma) / sigma) ** 2 / 2) / (sigma * np.sqrt(2 * np.pi))
