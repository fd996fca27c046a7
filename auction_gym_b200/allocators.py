"""CTR allocators -- same class names and constructor signatures as the reference's
src/BidderAllocation.py, so ``eval(f"{type}(rng=rng, ...)")`` on the strings in config/*.json keeps
working (src/main.py:85).  The objects are descriptors: the arithmetic (estimate_CTR over a batch of
opportunities, Thompson draw, per-iteration fit) runs in the engine; once an agent is attached to an
``Auction`` the learnt parameters are views of the device state.
"""
import numpy as np

from . import _lib


class Allocator:
    """Base class for an allocator (BidderAllocation.py:11-18)."""

    kind = None

    def __init__(self, rng):
        self.rng = rng
        self._auction = None
        self._index = None

    def _attach(self, auction, index):
        self._auction, self._index = auction, index

    def _estimate(self, context, sample, _eps=None):
        """One context through agym_estimate_ctr (the engine is built on first use, as simulate_opportunity would)."""
        au = self._auction
        if au is None:
            raise RuntimeError("estimate_CTR needs the agent to be part of an Auction (the catalog and the learnt state live in its engine)")
        if au.engine is None:
            au._build()
        au._queries += 1
        return au.engine.estimate_ctr(self._index, context, sample=sample, eps=_eps, seed=au.seed, iteration=au.iteration, query=au._queries)

    def update(self, contexts, items, outcomes, iteration, plot, figsize, fontsize, name):
        """The reference passes the agent's logged rows; here the rows already live on the device, so the
        arguments are ignored and the batched fit of the attached auction runs (once per iteration)."""
        if self._auction is not None:
            self._auction._update_models()


class ResponseModelView:
    """Read/write view of one agent's learnt parameters (Models.py:18-26: ``m``, ``q``, ``prev_iter_m``)
    inside the engine's ``[R, A, I, K]`` state.  ``run`` selects the replica (0 for a single-run auction)."""

    def __init__(self, allocator, run=0):
        self._a, self._run = allocator, run

    def _slice(self, t):
        a = self._a
        return t[self._run, a._index, :a.num_items]

    @property
    def m(self):
        a = self._a
        return self._slice(a._auction.engine.m) if a._auction is not None else a._init_m

    @property
    def q(self):
        a = self._a
        return self._slice(a._auction.engine.q) if a._auction is not None else a._init_q

    @property
    def prev_iter_m(self):
        a = self._a
        return self._slice(a._auction.engine.m_prev) if a._auction is not None else a._init_m


class PyTorchLogisticRegressionAllocator(Allocator):
    """Per-item Bayesian logistic regression with Thompson sampling (BidderAllocation.py:21-68,
    Models.py:18-48).  Initial state as Models.py:21-24: m ~ N(0, 1), q = 1, prev_iter_m = m.
    The reference draws m from torch's unseeded global generator; here it comes from the ``rng`` that
    is passed in, so a config seed reproduces a run."""

    def __init__(self, rng, embedding_size, num_items, thompson_sampling=True):
        super().__init__(rng)
        self.embedding_size = int(embedding_size)
        self.num_items = int(num_items)
        self.thompson_sampling = bool(thompson_sampling)
        self.kind = _lib.ALLOC_TS if self.thompson_sampling else _lib.ALLOC_MAP
        draw = rng.standard_normal if hasattr(rng, "standard_normal") else np.random.default_rng().standard_normal
        self._init_m = np.asarray(draw((self.num_items, self.embedding_size + 1)), np.float32)
        self._init_q = np.ones((self.num_items, self.embedding_size + 1), np.float32)
        self.response_model = ResponseModelView(self)

    def estimate_CTR(self, context, sample=True, _eps=None):
        """BidderAllocation.py:67-68: float32 estimates of every item for one observed context [Do + 1]; a Thompson draw when
        ``thompson_sampling and sample``.  ``_eps`` ([num_items, Do + 1] standard normals) replays a given draw (tests)."""
        est = self._estimate(context, bool(self.thompson_sampling and sample), _eps)
        return est[:self.num_items].astype(np.float32)


# the stale name used in the reference's own comment (src/main.py:16) and in BASELINE.json
LogisticTSAllocator = PyTorchLogisticRegressionAllocator


class OracleAllocator(Allocator):
    """Acts on the true P(click) (BidderAllocation.py:71-82)."""

    kind = _lib.ALLOC_ORACLE

    def __init__(self, rng):
        super().__init__(rng)
        self.item_embeddings = None

    def update_item_embeddings(self, item_embeddings):
        self.item_embeddings = item_embeddings

    def estimate_CTR(self, context):
        """BidderAllocation.py:81-82: sigmoid(item_embeddings @ context) for one true context [D + 1], float64."""
        return self._estimate(context, False)[:len(self.item_embeddings)]
