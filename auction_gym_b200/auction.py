"""``Auction`` with the reference's constructor and methods (src/Auction.py:9-77) on top of the engine.

``simulate_opportunity()`` keeps its one-round-at-a-time meaning (notebooks call it in a Python loop);
``simulate_rounds(T)`` is the batched form that the driver uses: one fused kernel launch simulates T
rounds of every resident run.  ``num_runs`` independent replicas of the auction (main.py:186) can live in
one object; the reference surface then reports arrays of length ``num_runs`` instead of scalars.
"""
import numpy as np

from . import _lib
from .agent import materialise_logs
from .allocators import OracleAllocator
from .engine import Engine

_LOG_FIELDS = ("agent", "item", "est", "value", "bid", "true_ctr", "best_ev", "price", "second", "gamma", "propensity",
               "outcome", "won", "ctx")


class Auction:
    """Base class for auctions (Auction.py:9-26)."""

    def __init__(self, rng, allocation, agents, agent2items, agents2item_values, max_slots, embedding_size, embedding_var,
                 obs_embedding_size, num_participants_per_round, *, num_runs=1, run_offset=0, device=0, precision=None,
                 seed=None, init_seed=None, rounds_capacity=0, per_run_init=None):
        if int(max_slots) < 1:
            raise ValueError("max_slots must be >= 1")
        self.rng = rng
        self.allocation = allocation
        self.agents = agents
        self.max_slots = max_slots
        self.agent2items = agent2items
        self.agents2item_values = agents2item_values
        self.embedding_size = embedding_size
        self.embedding_var = embedding_var
        self.obs_embedding_size = obs_embedding_size
        self.num_participants_per_round = num_participants_per_round
        self.num_runs = int(num_runs)
        self.run_offset = int(run_offset)
        self.device = device
        self.precision = _lib.FP32 if precision is None else precision
        if seed is None:
            seed = int(rng.integers(0, 2**62)) if hasattr(rng, "integers") and hasattr(rng, "bit_generator") else 0
        self.seed = int(seed)
        self.init_seed = self.seed if init_seed is None else int(init_seed)
        # Initial model state.  per_run_init = True (the batched driver, any shard): every draw is seeded by the GLOBAL run
        # index, so run r starts from the same state whichever rank owns it.  False (the reference's own calling pattern, one
        # run driven from Python): the allocator's constructor draw, from the rng the caller passed, as Models.py:21-24 does.
        self.per_run_init = bool(self.num_runs > 1 or self.run_offset) if per_run_init is None else bool(per_run_init)
        self.iteration = 0
        self.fit_mode = _lib.FIT_ADAM_REF  # allocator fit arithmetic (include/agym.h: agym_fit_mode); FIT_NEWTON is opt-in, a different algorithm
        self.engine = None
        self._rounds_capacity = int(rounds_capacity)
        self._models_updated = False
        self._queries = 0       # per-context host queries so far (estimate_CTR / select_item / bid): their Philox counter
        self._cleared = set()
        self._log_chunks = []   # detailed log of run 0 for the current iteration (list of dicts of numpy arrays)
        self._log_cache = None
        self._retained = {}     # agent index -> (records, gammas, propensities) kept by clear_logs (Agent(memory=...))
        for i, ag in enumerate(self.agents):
            ag._attach(self, i)

    # ------------------------------------------------------------------ engine construction
    def _build(self):
        A = len(self.agents)
        D, Do = int(self.embedding_size), int(self.obs_embedding_size)
        n_items = np.array([int(ag.num_items) for ag in self.agents], np.int32)
        I = int(n_items.max())
        E = np.zeros((A, I, D + 1))
        V = np.ones((A, I))
        for a, ag in enumerate(self.agents):
            e = np.asarray(self.agent2items[ag.name], np.float64)
            if e.shape != (n_items[a], D + 1):
                raise ValueError(f"item embeddings of {ag.name!r} have shape {e.shape}, expected {(n_items[a], D + 1)}")
            E[a, :n_items[a]] = e
            V[a, :n_items[a]] = np.asarray(self.agents2item_values[ag.name], np.float64)
        eng = Engine(R=self.num_runs, A=A, I=I, D=D, Do=Do, P=int(self.num_participants_per_round),
                     mechanism=self.allocation.code, E=E, V=V, n_items=n_items,
                     alloc_kind=[ag.allocator.kind for ag in self.agents], bidder_kind=[ag.bidder.kind for ag in self.agents],
                     embedding_var=float(self.embedding_var), precision=self.precision, device=self.device,
                     run_offset=self.run_offset, rounds_capacity=self._rounds_capacity,
                     bidder_fit=[ag.bidder.fit_kind for ag in self.agents], max_slots=int(self.max_slots),
                     memory=[int(ag.memory or 0) for ag in self.agents])
        if eng.any_learnt:
            m = np.zeros((self.num_runs, A, I, Do + 1), np.float32)
            q = np.ones_like(m)
            for r in range(self.num_runs):
                # one independent initialisation per run (the reference re-instantiates its agents per run, main.py:188)
                rr = np.random.default_rng([self.init_seed, self.run_offset + r, 0x6d30]) if self.per_run_init else None
                for a, ag in enumerate(self.agents):
                    if isinstance(ag.allocator, OracleAllocator):
                        continue
                    if ag.allocator.embedding_size != Do:
                        raise ValueError(f"{ag.name!r}: allocator embedding_size {ag.allocator.embedding_size} != obs_embedding_size {Do}")
                    m[r, a, :n_items[a]] = ag.allocator._init_m if rr is None else rr.standard_normal((n_items[a], Do + 1))
            eng.set_allocator_state(m, q)
        if eng.any_shaded:
            pg = np.array([ag.bidder._gamma_params()[0] for ag in self.agents])
            sg = np.array([ag.bidder._gamma_params()[1] for ag in self.agents])
            # PyTorchWinRateEstimator = Linear(3, 1): torch's default init is U(-1/sqrt(3), 1/sqrt(3)) for weight and bias
            # (Models.py:55-58); one independent draw per (run, agent), from the seed instead of torch's global generator
            # (keyed by the global run index: a shard starts its runs from the same weights as a single-device job)
            wr = np.stack([np.random.default_rng([self.init_seed, self.run_offset + r, 0x7772]).uniform(-1, 1, (A, 4))
                           for r in range(self.num_runs)]) / np.sqrt(3.0)
            # policy nets (Models.py:71-77,97-101): Linear layers with fan-in 2 -> U(-1/sqrt(2), 1/sqrt(2))
            pw = np.stack([np.random.default_rng([self.init_seed, self.run_offset + r, 0x706f]).uniform(-1, 1, (A, 12))
                           for r in range(self.num_runs)]) / np.sqrt(2.0)
            eng.set_bidder_state(pg[None, :], sg[None, :], winrate_w=wr, policy_w=pw)
        self.engine = eng
        self.D_ctx = [D if isinstance(ag.allocator, OracleAllocator) else Do for ag in self.agents]  # Auction.py:46-49

    # ------------------------------------------------------------------ round loop
    def simulate_opportunity(self):
        """One auction round (Auction.py:28-74) for every resident run; the detailed log is kept."""
        self.simulate_rounds(1, keep_logs=True)

    def simulate_rounds(self, T, keep_logs=False):
        """T rounds in one fused launch (replaces the Python loop at src/main.py:116-117)."""
        if self.engine is None:
            self._build()
        if self._cleared:  # an iteration boundary was only partly crossed (some agents cleared their logs, not all)
            raise NotImplementedError("clear_logs() must be called for every agent before the next round (the batched "
                                      "logs are rewound once, when the last agent has cleared)")
        self._models_updated = False
        out = self.engine.simulate(self.seed, self.iteration, int(T), _LOG_FIELDS if keep_logs else None)
        if keep_logs:
            self._log_chunks.append({k: v[0].cpu().numpy() for k, v in out.items()})
            self._log_cache = None

    @property
    def revenue(self):  # Auction.py:16,74
        if self.engine is None:
            return 0.0
        v = self.engine.revenue.cpu().numpy()
        return float(v[0]) if len(v) == 1 else v

    @revenue.setter
    def revenue(self, value):
        if self.engine is not None:
            self.engine.revenue.fill_(float(value))

    def clear_revenue(self):  # Auction.py:76-77
        if self.engine is not None:
            self.engine.revenue.zero_()

    # ------------------------------------------------------------------ per-iteration model updates
    def _update_models(self):
        if self._models_updated or self.engine is None:
            return
        unsupported = sorted({type(ag.bidder).__name__ for ag in self.agents if ag.bidder.needs_fit and not ag.bidder.fit_built})
        if unsupported:
            raise _lib.AgymError(f"bidder update for {unsupported} (K7, src/Bidder.py:60-147,278-316,369-431,477-615) is not built yet; "
                                 "see DESIGN.md 'not yet built'")
        self.engine.update_allocators(want_info=False, fit_mode=self.fit_mode)
        info = self.engine.update_bidders(self.seed, self.iteration)
        if info is not None and bool((info[..., 1] > 0).logical_and(info[..., 2].isnan()).any()):
            raise _lib.AgymError("NaN loss in a bidder fit (the reference prints 'NAN DETECTED!' and exits, Bidder.py:412-419,598-605)")
        self._models_updated = True

    # ------------------------------------------------------------------ logs
    def _log_columns(self):
        if self._log_cache is None and self._log_chunks:
            self._log_cache = {k: np.concatenate([c[k] for c in self._log_chunks], axis=0) for k in self._log_chunks[0]}
        return self._log_cache

    def _agent_logs(self, index):
        kept = self._retained.get(index, ([], [], []))[0]
        cols = self._log_columns()
        if cols is None or index in self._cleared:  # clear_logs() of this agent already ran (Agent.py:124-129)
            return list(kept)
        return list(kept) + materialise_logs(cols, index, self.D_ctx)

    def _agent_log_column(self, index, name):
        kept = self._retained.get(index, ([], [], []))[1 if name == "gamma" else 2]
        cols = self._log_columns()
        if cols is None or index in self._cleared:
            return list(kept)
        return list(kept) + list(cols[name][cols["agent"] == index])

    def _roll_host_logs(self, index):
        """Host view of Agent.clear_logs / Bidder.clear_logs (Agent.py:124-129, Bidder.py:149-153): keep the last `memory`."""
        mem = int(self.agents[index].memory or 0)
        if mem and (self._log_chunks or index in self._retained):
            self._retained[index] = (self._agent_logs(index)[-mem:], self._agent_log_column(index, "gamma")[-mem:],
                                     self._agent_log_column(index, "propensity")[-mem:])
        else:
            self._retained.pop(index, None)

    def _clear_agent_logs(self, index):
        if self.engine is None:
            return
        self._roll_host_logs(index)
        if not self.agents[index].memory:
            keep = [_lib.M_NET, _lib.M_GROSS]
            cols = [c for c in range(_lib.NUM_METRICS) if c not in keep]
            self.engine.acc[:, index, cols] = 0.0
        self._cleared.add(index)
        if len(self._cleared) == len(self.agents):  # every agent crossed the iteration boundary
            if self.engine.retention:  # the kept records' sums replace the log-derived accumulators of every agent
                self.engine.retain_logs()
            else:
                self.engine._check(self.engine.lib.agym_set_rounds_in_iteration(self.engine.handle, 0))
            self._log_chunks, self._log_cache = [], None
            self._cleared = set()
            self.iteration += 1

    # ------------------------------------------------------------------ checkpoint
    def save_checkpoint(self, path, **extra):
        """Learnt state of every resident run at an iteration boundary -> ``path`` (.npz); ``extra`` arrays ride along
        (the driver stores the metrics collected so far).  The reference has no counterpart."""
        if self.engine is None:
            self._build()
        d = self.engine.save_state()
        d.update(iteration=np.int64(self.iteration), seed=np.int64(self.seed), run_offset=np.int64(self.run_offset), **extra)
        tmp = str(path) + ".tmp.npz"
        np.savez(tmp, **d)
        import os

        os.replace(tmp, str(path))

    def load_checkpoint(self, path):
        """Resume from ``save_checkpoint``: same config, same seed, same shard.  Returns the stored dict."""
        if self.engine is None:
            self._build()
        d = dict(np.load(str(path)))
        if int(d["seed"]) != self.seed or int(d["run_offset"]) != self.run_offset:
            raise _lib.AgymError(f"checkpoint of seed {int(d['seed'])} / first run {int(d['run_offset'])} does not belong to this auction "
                                 f"(seed {self.seed}, first run {self.run_offset})")
        self.engine.load_state(d)
        self.iteration = int(d["iteration"])
        self._models_updated = False
        return d

    def end_iteration(self):
        """Batched equivalent of main.py:151-155 for all agents: clear utilities, logs and revenue."""
        if self.engine is None:
            return
        for index in range(len(self.agents)):
            self._roll_host_logs(index)
        self.engine.clear_iteration()
        self._log_chunks, self._log_cache = [], None
        self._cleared = set()
        self.iteration += 1
