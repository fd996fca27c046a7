"""TEST INFRASTRUCTURE ONLY -- harness that drives the UNMODIFIED reference in replay mode.

This file is part of ``oracle/`` (the checker).  Nothing in the product path
(``auction_gym_b200/``) may import it.  The reference tree it drives is
``oracle/_ref/src`` (the verbatim, git-ignored copy written by ``oracle/make_ref.py``,
which travels to the GPU box) or ``/root/reference/src`` (build container only).  The
``-m gpu`` tests never need it: they use the committed fixtures in ``tests/golden/``
that ``oracle/make_golden.py`` produced with this harness.

What it does:
  * stubs ``matplotlib`` / ``seaborn`` (not installed) and lets
    ``ReduceLROnPlateau`` swallow the removed ``verbose=`` kwarg
    (reference ``src/Bidder.py:243,286,392,521,578``);
  * imports the reference modules by bare name from ``/root/reference/src``
    exactly as the reference's own ``main.py`` does (``src/main.py:12-16``);
  * provides ``ReplayRNG``, a duck-typed stand-in for ``numpy.random.Generator``
    that serves host-drawn arrays at the five call sites on the hot path
    (``src/Auction.py:30,33,42,65``; ``src/Bidder.py:177,185,354,461``);
  * patches ``torch.normal`` so the Thompson draw of ``src/Models.py:31`` becomes
    ``mean + eps[t, slot] * std`` with host-drawn ``eps``.

The click rule in replay mode is ``outcome = (u < p)`` on every side (SURVEY.md
App. C) -- the replay RNG replaces numpy's own binomial.
"""
from __future__ import annotations

import importlib
import os
import sys
from unittest import mock

import numpy as np

from . import make_ref as _make_ref

REF_SRC = os.environ.get("AGYM_REF_SRC") or _make_ref.ref_src() or "/root/reference/src"
REF_CONFIG = os.path.join(os.path.dirname(REF_SRC), "config")

_ref_modules = None


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_SRC, "Auction.py"))


def load_reference():
    """Import the unmodified reference modules (cached).  Returns a namespace dict."""
    global _ref_modules
    if _ref_modules is not None:
        return _ref_modules
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REF_SRC}")
    plt = mock.MagicMock()
    plt.subplots.side_effect = lambda *a, **k: (mock.MagicMock(), mock.MagicMock())
    mpl = mock.MagicMock()
    mpl.pyplot = plt
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pyplot", plt)
    sys.modules.setdefault("seaborn", mock.MagicMock())
    import torch

    base = torch.optim.lr_scheduler.ReduceLROnPlateau
    if not getattr(base, "_agym_compat", False):

        class _PlateauCompat(base):
            _agym_compat = True

            def __init__(self, *a, verbose=False, **k):
                super().__init__(*a, **k)

        torch.optim.lr_scheduler.ReduceLROnPlateau = _PlateauCompat
    # The product package also ships modules called Auction / Agent / ...; make sure the
    # bare names resolve to the reference for the lifetime of this harness.
    saved = {}
    names = ["Impression", "Models", "AuctionAllocation", "BidderAllocation", "Bidder", "Agent", "Auction", "main"]
    for n in names:
        if n in sys.modules:
            saved[n] = sys.modules.pop(n)
    sys.path.insert(0, REF_SRC)
    try:
        mods = {n: importlib.import_module(n) for n in names}
    finally:
        sys.path.remove(REF_SRC)
        for n in names:
            # keep the reference modules reachable only through the returned dict
            sys.modules.pop(n, None)
        sys.modules.update(saved)
    quiet = lambda it, **k: it  # noqa: E731 - silence tqdm bars
    for n in ("BidderAllocation", "Bidder", "Models", "main"):
        if hasattr(mods[n], "tqdm"):
            mods[n].tqdm = quiet
    _ref_modules = mods
    return mods


class ReplayRNG:
    """Serves pre-drawn arrays to the reference's rng call sites, indexed by round ``t``.

    ctx        [T, D]     value returned by ``rng.normal(0, sd, size=D)``      (Auction.py:33)
    parts      [T, P]     value returned by ``rng.choice(A, P, replace=False)`` (Auction.py:42)
    gamma_z    [T, P]     standard normals; scalar ``rng.normal(mu, sd)`` -> mu + sd*z[t, slot]
                          (Bidder.py:177,354,461)
    grid_u     [T, P, G]  uniforms in [0,1); ``rng.uniform(lo, hi, size=G)`` -> lo + (hi-lo)*u
                          (Bidder.py:185)
    u          [T]        uniforms; ``rng.binomial(1, p)`` -> (u[t] < p)           (Auction.py:65)
    """

    def __init__(self, ctx, parts, u, gamma_z=None, grid_u=None, num_slots=None):
        self.ctx = np.asarray(ctx, dtype=np.float64)
        self.parts = np.asarray(parts, dtype=np.int64)
        self.u = np.asarray(u, dtype=np.float64)
        self.gamma_z = None if gamma_z is None else np.asarray(gamma_z, dtype=np.float64)
        self.grid_u = None if grid_u is None else np.asarray(grid_u, dtype=np.float64)
        self.num_slots = None if num_slots is None else np.asarray(num_slots, dtype=np.int64)  # [T] for max_slots > 1 (Auction.py:30)
        self.t = -1
        self.slot = 0  # participant slot inside the current round; advanced by the bid wrapper

    def integers(self, lo, hi):
        # Auction.py:30 draws the slot count BEFORE the context, i.e. before `t` advances
        return 1 if self.num_slots is None else int(self.num_slots[self.t + 1])

    def normal(self, mu, sd, size=None):
        if size is not None:
            self.t += 1
            self.slot = 0
            return self.ctx[self.t].copy()
        return mu + sd * self.gamma_z[self.t, self.slot]

    def choice(self, n, k, replace=False):
        return self.parts[self.t].copy()

    def uniform(self, lo, hi, size=None):
        return lo + (hi - lo) * self.grid_u[self.t, self.slot, :size].copy()

    def binomial(self, n, p):
        p = np.asarray(p)
        u = self.u[self.t]
        return (u[: len(p)] < p).astype(np.int64) if np.ndim(u) else (u < p).astype(np.int64)


class patched_ts_noise:
    """Context manager: torch.normal(mean, std) -> mean + eps[t, slot] * std  (Models.py:31)."""

    def __init__(self, rng: ReplayRNG, eps):
        import torch

        self.torch = torch
        self.rng = rng
        self.eps = None if eps is None else torch.from_numpy(np.asarray(eps, dtype=np.float32))

    def __enter__(self):
        self._orig = self.torch.normal

        def fake_normal(mean=0.0, std=1.0, *a, **k):
            e = self.eps[self.rng.t, self.rng.slot]
            e = e[: std.shape[0], : std.shape[1]]
            return mean + e * std

        self.torch.normal = fake_normal
        return self

    def __exit__(self, *exc):
        self.torch.normal = self._orig
        return False


def build_reference_auction(cfg: dict, E: dict, V: dict, rng, ref=None, max_slots=1):
    """Instantiate the reference's agents + auction for ``cfg`` (same schema as config/*.json)
    through the reference's own helpers (src/main.py:77-109)."""
    ref = ref or load_reference()
    main = ref["main"]
    import copy

    agent_configs = []
    n = 0
    for ac in cfg["agents"]:
        if "num_copies" in ac:
            for _ in range(ac["num_copies"]):
                c = copy.deepcopy(ac)
                c["name"] += f" {n + 1}"
                agent_configs.append(c)
                n += 1
        else:
            agent_configs.append(ac)
            n += 1
    agents = main.instantiate_agents(rng, agent_configs, V, E)
    auction, num_iter, rounds_per_iter, output_dir = main.instantiate_auction(
        rng, cfg, E, V, agents, max_slots, cfg["embedding_size"], cfg["embedding_var"], cfg["obs_embedding_size"]
    )
    return auction, agents, agent_configs


def wrap_bid_slots(agents, rng: ReplayRNG):
    """Advance ``rng.slot`` after each agent.bid so per-slot noise is addressed by participant slot."""
    for ag in agents:
        orig = ag.bid

        def wrapped(context, _orig=orig):
            out = _orig(context)
            rng.slot += 1
            return out

        ag.bid = wrapped


def run_reference_rounds(auction, agents, rng: ReplayRNG, T: int, ts_eps=None):
    """Run T rounds of the unmodified ``Auction.simulate_opportunity`` (src/Auction.py:28-74)
    and return per-round / per-slot records read back from ``agent.logs``."""
    P = rng.parts.shape[1]
    A = len(agents)
    rec = {
        "item": np.zeros((T, P), np.int32),
        "est": np.zeros((T, P), np.float64),
        "value": np.zeros((T, P), np.float64),
        "bid": np.zeros((T, P), np.float64),
        "true_ctr": np.zeros((T, P), np.float64),
        "best_ev": np.zeros((T, P), np.float64),
        "price": np.zeros((T, P), np.float64),
        "second": np.zeros((T, P), np.float64),
        "outcome": np.zeros((T, P), np.uint8),
        "won": np.zeros((T, P), np.uint8),
        "gamma": np.full((T, P), np.nan, np.float64),
        "propensity": np.full((T, P), np.nan, np.float64),
    }
    with patched_ts_noise(rng, ts_eps):
        for t in range(T):
            n_before = [len(a.logs) for a in agents]
            g_before = [len(getattr(a.bidder, "gammas", [])) for a in agents]
            auction.simulate_opportunity()
            for s, a_idx in enumerate(rng.parts[t]):
                ag = agents[a_idx]
                assert len(ag.logs) == n_before[a_idx] + 1
                opp = ag.logs[-1]
                rec["item"][t, s] = opp.item
                rec["est"][t, s] = float(opp.estimated_CTR)
                rec["value"][t, s] = opp.value
                rec["bid"][t, s] = opp.bid
                rec["true_ctr"][t, s] = opp.true_CTR
                rec["best_ev"][t, s] = opp.best_expected_value
                rec["price"][t, s] = opp.price
                rec["second"][t, s] = opp.second_price
                rec["outcome"][t, s] = int(opp.outcome)
                rec["won"][t, s] = int(opp.won)
                gs = getattr(ag.bidder, "gammas", None)
                if gs is not None and len(gs) == g_before[a_idx] + 1:
                    rec["gamma"][t, s] = float(gs[-1])
                    rec["propensity"][t, s] = float(ag.bidder.propensities[-1]) if hasattr(ag.bidder, "propensities") else np.nan
    return rec


def reference_metrics(auction, agents):
    """Per-agent metrics exactly as src/main.py:131-148 reads them (before clear_*)."""
    import warnings

    out = {k: np.zeros(len(agents), np.float64) for k in
           ("net", "gross", "alloc_regret", "estim_regret", "overbid_regret", "underbid_regret",
            "ctr_rmse", "ctr_bias", "best_ev_mean", "n_logs", "n_won")}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for i, ag in enumerate(agents):
            out["net"][i] = ag.net_utility
            out["gross"][i] = ag.gross_utility
            out["n_logs"][i] = len(ag.logs)
            out["n_won"][i] = sum(1 for o in ag.logs if o.won)
            if len(ag.logs) == 0:
                for k in ("alloc_regret", "estim_regret", "overbid_regret", "underbid_regret"):
                    out[k][i] = 0.0
                out["ctr_rmse"][i] = out["ctr_bias"][i] = out["best_ev_mean"][i] = np.nan
                continue
            out["alloc_regret"][i] = ag.get_allocation_regret()
            out["estim_regret"][i] = ag.get_estimation_regret()
            out["overbid_regret"][i] = ag.get_overbid_regret()
            out["underbid_regret"][i] = ag.get_underbid_regret()
            out["ctr_rmse"][i] = ag.get_CTR_RMSE()
            out["ctr_bias"][i] = ag.get_CTR_bias()
            out["best_ev_mean"][i] = np.mean([o.best_expected_value for o in ag.logs])
    out["revenue"] = np.float64(auction.revenue)
    return out
