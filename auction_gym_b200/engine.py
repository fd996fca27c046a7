"""Device-resident batch of auction runs: the host side of the C ABI.

``Engine`` owns (through torch) every device buffer the library borrows -- learnt allocator state
``[R, A, I, K]``, bidder state, metric accumulators, the winner log that feeds the allocator fit -- and
forwards to ``libagym.so``.  The reference-named classes (``Auction``, ``Agent`` ... in this package)
and the batched driver (``main.py``) sit on top of it.  torch is used for memory and streams only.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import AgymError, ReplayInputs, RoundLog, Shape

_LOG_DTYPES = {
    "agent": torch.int32, "item": torch.int32, "est": torch.float64, "value": torch.float64, "bid": torch.float64,
    "true_ctr": torch.float64, "best_ev": torch.float64, "price": torch.float64, "second": torch.float64,
    "gamma": torch.float64, "propensity": torch.float64, "outcome": torch.uint8, "won": torch.uint8,
}


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class Engine:
    """R independent runs of one auction configuration on one GPU.

    Parameters mirror what ``parse_config`` / ``instantiate_*`` of the reference produce
    (src/main.py:24-109): the catalog ``E [A, I, D+1]`` / ``V [A, I]``, per-agent item counts and
    allocator / bidder kinds (``_lib.ALLOC_*`` / ``_lib.BID_*``).
    """

    def __init__(self, *, R, A, I, D, Do, P, mechanism, E, V, n_items, alloc_kind, bidder_kind, embedding_var=1.0,
                 precision=_lib.FP32, device=0, run_offset=0, rounds_capacity=0, bidder_fit=None, memory=None, max_slots=1):
        if not torch.cuda.is_available():
            raise AgymError("CUDA device required: the AuctionGym B200 engine has no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device("cuda", device)
        self.R, self.A, self.I, self.D, self.Do, self.P = int(R), int(A), int(I), int(D), int(Do), int(P)
        self.K = self.Do + 1
        self.max_slots = max(1, int(max_slots))  # slots per round are uniform in [1, max_slots] (Auction.py:30)
        self.mechanism, self.precision = int(mechanism), int(precision)
        self.shape = Shape(self.R, self.A, self.I, self.D, self.Do, self.P, self.mechanism, self.precision,
                           int(run_offset), self.max_slots, float(embedding_var))
        self.handle = C.c_void_p()
        rc = self.lib.agym_create(C.byref(self.shape), device, C.byref(self.handle))
        if rc != 0:
            raise AgymError(f"agym_create failed ({rc}): {self.lib.agym_last_error(None).decode()}")
        self.n_items = np.ascontiguousarray(n_items, np.int32)
        self.alloc_kind = np.ascontiguousarray(alloc_kind, np.int32)
        self.bidder_kind = np.ascontiguousarray(bidder_kind, np.int32)
        assert self.n_items.shape == (self.A,) and self.alloc_kind.shape == (self.A,) and self.bidder_kind.shape == (self.A,)
        self._check(self.lib.agym_set_agents(self.handle, self.n_items.ctypes.data, self.alloc_kind.ctypes.data,
                                             self.bidder_kind.ctypes.data))
        if bidder_fit is not None:  # which Bidder.update each agent runs (only needed for AGYM_BID_BANDIT agents)
            self.bidder_fit = np.ascontiguousarray(bidder_fit, np.int32)
            assert self.bidder_fit.shape == (self.A,)
            self._check(self.lib.agym_set_bidder_fits(self.handle, self.bidder_fit.ctypes.data))
        E = np.ascontiguousarray(E, np.float64)
        V = np.ascontiguousarray(V, np.float64)
        assert E.shape == (self.A, self.I, self.D + 1) and V.shape == (self.A, self.I), (E.shape, V.shape)
        self._check(self.lib.agym_set_catalog(self.handle, E.ctypes.data, V.ctypes.data))
        self.any_learnt = bool((self.alloc_kind != _lib.ALLOC_ORACLE).any())
        self.any_shaded = bool((self.bidder_kind != _lib.BID_TRUTHFUL).any())
        dev = self.device
        # metric accumulators (Agent.net_utility ... / Auction.revenue)
        self.acc = torch.zeros((self.R, self.A, _lib.NUM_METRICS), dtype=torch.float64, device=dev)
        self.revenue = torch.zeros((self.R,), dtype=torch.float64, device=dev)
        self._check(self.lib.agym_bind_metrics(self.handle, _ptr(self.acc), _ptr(self.revenue)))
        # learnt allocator state (Models.py:21-24): m ~ N(0,1) is drawn by the caller; q = 1
        st = (self.R, self.A, self.I, self.K)
        if self.any_learnt:
            self.m = torch.zeros(st, dtype=torch.float32, device=dev)
            self.q = torch.ones(st, dtype=torch.float32, device=dev)
            self.m_prev = torch.zeros(st, dtype=torch.float32, device=dev)
            self.sigma = torch.ones(st, dtype=torch.float32, device=dev)
            self._check(self.lib.agym_bind_allocator_state(self.handle, _ptr(self.m), _ptr(self.q), _ptr(self.m_prev), _ptr(self.sigma)))
        else:
            self.m = self.q = self.m_prev = self.sigma = None
        self.bidder_d = torch.zeros((self.R, self.A, _lib.BIDDER_D), dtype=torch.float64, device=dev)
        self.bidder_w = torch.zeros((self.R, self.A, _lib.BIDDER_W), dtype=torch.float32, device=dev)
        self._check(self.lib.agym_bind_bidder_state(self.handle, _ptr(self.bidder_d), _ptr(self.bidder_w)))
        self.fit_ctx = self.fit_meta = self.workspace = None
        self.bid_rows = self.bid_meta = self.bidder_workspace = None
        self.learning_bidders = bool(np.isin(self.bidder_kind, [_lib.BID_SEARCH, _lib.BID_BANDIT, _lib.BID_POLICY]).any()) or \
            (bidder_fit is not None and bool((np.asarray(bidder_fit) != _lib.BFIT_NONE).any()))
        # log retention across iterations (Agent(memory=...), Agent.py:124-129): rows reserved at the head of the logs
        self.memory = np.zeros(self.A, np.int32) if memory is None else np.ascontiguousarray(memory, np.int32)
        assert self.memory.shape == (self.A,) and (self.memory >= 0).all()
        self.log_base = int(self.memory.sum())
        self.retention = self.log_base > 0
        self.terms = None
        self.rounds_capacity = 0
        if rounds_capacity:
            self.reserve_rounds(rounds_capacity)

    # ------------------------------------------------------------------ plumbing
    def _check(self, rc):
        if rc != 0:
            msg = self.lib.agym_last_error(self.handle)
            raise AgymError(f"libagym error {rc}: {msg.decode() if msg else '?'}")

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def launch_count(self):
        """Kernels of libagym launched through this engine so far."""
        return int(self.lib.agym_launch_count(self.handle))

    def set_option(self, name, value):
        """Kernel-selection override (include/agym.h: agym_set_option) -- tests and experiments only."""
        self._check(self.lib.agym_set_option(self.handle, name.encode(), float(value)))

    def close(self):
        if getattr(self, "handle", None) and self.handle.value:
            self.lib.agym_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reserve_rounds(self, T):
        """Bind the winner log (fit_ctx / fit_meta), the bid log and the fit workspaces for up to T rounds per iteration
        (plus the ``log_base`` rows that hold retained records when some agent has ``memory``)."""
        T = int(T)
        if T <= self.rounds_capacity:
            return
        need_bid = self.learning_bidders or self.retention
        if not self.any_learnt and not need_bid:
            self.rounds_capacity = T
            return
        done = int(self.lib.agym_rounds_in_iteration(self.handle))
        if done:  # growing in the middle of an iteration (simulate_opportunity() one round at a time): amortise
            T = max(T, 2 * self.rounds_capacity)
        dev, cap = self.device, T * self.max_slots + self.log_base  # the winner log holds max_slots rows per round
        filled = self.log_base + done * self.max_slots  # retained rows + rounds recorded so far move to the new buffers

        def grown(old, shape, dtype, zero):
            new = (torch.zeros if zero else torch.empty)(shape, dtype=dtype, device=dev)
            if old is not None and filled:
                new[:, :filled] = old[:, :filled]
            return new

        if need_bid:  # per-(round, slot) bid records for the bidder fits / retention
            self.bid_rows = grown(self.bid_rows, (self.R, cap, self.P, _lib.BID_ROW), torch.float32, False)
            self.bid_meta = grown(self.bid_meta, (self.R, cap, self.P), torch.int32, True)
            self._check(self.lib.agym_bind_bid_log(self.handle, _ptr(self.bid_rows), _ptr(self.bid_meta), cap))
        if self.learning_bidders:
            nb = int(self.lib.agym_bidder_workspace_bytes(self.handle, cap))
            self.bidder_workspace = torch.empty((nb,), dtype=torch.uint8, device=dev)
            self._check(self.lib.agym_bind_bidder_workspace(self.handle, _ptr(self.bidder_workspace), nb))
        if self.any_learnt:
            self.fit_ctx = grown(self.fit_ctx, (self.R, cap, max(self.Do, 1)), torch.float32, False)
            self.fit_meta = grown(self.fit_meta, (self.R, cap), torch.int32, True)
            self._check(self.lib.agym_bind_fit_log(self.handle, _ptr(self.fit_ctx), _ptr(self.fit_meta), cap))
            nbytes = int(self.lib.agym_workspace_bytes(self.handle, cap))
            self.workspace = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
            self._check(self.lib.agym_bind_workspace(self.handle, _ptr(self.workspace), nbytes))
        if self.retention:
            self.terms = grown(self.terms, (self.R, cap, self.P, _lib.TERM_ROW), torch.float64, False)
            self._check(self.lib.agym_set_log_retention(self.handle, self.memory.ctypes.data, _ptr(self.terms)))
        self.rounds_capacity = T

    # ------------------------------------------------------------------ state
    def set_allocator_state(self, m, q=None, m_prev=None, non_blocking=False):
        """Upload learnt state ``[R, A, I, K]`` (numpy or torch, host or device); sigma is refreshed."""
        if not self.any_learnt:
            return
        m, q, m_prev = (np.array(v, dtype=np.float32) if isinstance(v, np.ndarray) else v for v in (m, q, m_prev))
        self.m.copy_(torch.as_tensor(m, dtype=torch.float32).reshape(self.m.shape), non_blocking=non_blocking)
        if q is not None:
            self.q.copy_(torch.as_tensor(q, dtype=torch.float32).reshape(self.q.shape), non_blocking=non_blocking)
        self.m_prev.copy_(self.m if m_prev is None else torch.as_tensor(m_prev, dtype=torch.float32).reshape(self.m.shape),
                          non_blocking=non_blocking)
        self._check(self.lib.agym_refresh_sigma(self.handle, self._stream()))

    def set_bidder_state(self, prev_gamma, gamma_sigma, initialised=None, winrate_w=None, policy_w=None):
        """Per-(run, agent) bidder state; arguments broadcast to ``[R, A]`` (Bidder.py:41-45,159-169,339-346,445-453)."""
        bd = np.zeros((self.R, self.A, _lib.BIDDER_D), np.float64)
        bd[..., 0] = np.broadcast_to(np.asarray(prev_gamma, np.float64), (self.R, self.A))
        bd[..., 1] = np.broadcast_to(np.asarray(gamma_sigma, np.float64), (self.R, self.A))
        if initialised is not None:
            bd[..., 2] = np.broadcast_to(np.asarray(initialised, np.float64), (self.R, self.A))
        self.bidder_d.copy_(torch.from_numpy(bd))
        if winrate_w is not None or policy_w is not None:
            bw = np.zeros((self.R, self.A, _lib.BIDDER_W), np.float32)
            if winrate_w is not None:
                bw[..., 0:4] = np.broadcast_to(np.asarray(winrate_w, np.float32), (self.R, self.A, 4))
            if policy_w is not None:
                bw[..., 4:16] = np.broadcast_to(np.asarray(policy_w, np.float32), (self.R, self.A, 12))
            self.bidder_w.copy_(torch.from_numpy(bw))

    # ------------------------------------------------------------------ checkpoint
    def save_state(self):
        """Everything that survives an iteration boundary, as host numpy arrays: the learnt allocator state (Models.py:21-24
        m, q, prev_iter_m), the bidder state (prev_gamma, gamma_sigma, model_initialised, win-rate and policy weights) and,
        with ``Agent(memory=...)``, the retained log rows plus the accumulators that restart from them.  Call it between
        iterations (after clear_iteration): the rounds of an unfinished iteration are not part of the state."""
        if self.rounds_in_iteration:
            raise AgymError("save_state: call it at an iteration boundary (after clear_iteration / end_iteration)")
        d = {"shape": np.array([self.R, self.A, self.I, self.D, self.Do, self.P, self.log_base], np.int64),
             "bidder_d": self.bidder_d.cpu().numpy(), "bidder_w": self.bidder_w.cpu().numpy()}
        if self.any_learnt:
            d.update(m=self.m.cpu().numpy(), q=self.q.cpu().numpy(), m_prev=self.m_prev.cpu().numpy())
        if self.retention:
            d.update(acc=self.acc.cpu().numpy(), revenue=self.revenue.cpu().numpy())
            for k in ("fit_ctx", "fit_meta", "bid_rows", "bid_meta", "terms"):
                t = getattr(self, k)
                if t is not None:
                    d["log_" + k] = t[:, :self.log_base].cpu().numpy()
        return d

    def load_state(self, d):
        """Inverse of save_state on an engine of the same shape and configuration."""
        want = np.array([self.R, self.A, self.I, self.D, self.Do, self.P, self.log_base], np.int64)
        if not np.array_equal(np.asarray(d["shape"]), want):
            raise AgymError(f"load_state: checkpoint shape {list(np.asarray(d['shape']))} != engine shape {list(want)}")
        self.bidder_d.copy_(torch.from_numpy(np.asarray(d["bidder_d"])))
        self.bidder_w.copy_(torch.from_numpy(np.asarray(d["bidder_w"])))
        if self.any_learnt:
            self.set_allocator_state(np.asarray(d["m"]), np.asarray(d["q"]), np.asarray(d["m_prev"]))
        if self.retention:
            if self.rounds_capacity == 0:
                self.reserve_rounds(1)
            self._check(self.lib.agym_clear_iteration(self.handle, self._stream()))
            self.acc.copy_(torch.from_numpy(np.asarray(d["acc"])))
            self.revenue.copy_(torch.from_numpy(np.asarray(d["revenue"])))
            for k in ("fit_ctx", "fit_meta", "bid_rows", "bid_meta", "terms"):
                t = getattr(self, k)
                if t is not None and "log_" + k in d:
                    t[:, :self.log_base].copy_(torch.from_numpy(np.asarray(d["log_" + k])))

    # ------------------------------------------------------------------ round loop
    def _alloc_log(self, n_runs, T, fields):
        log = RoundLog()
        out = {}
        for name in fields:
            if name == "winner":
                t = torch.empty((n_runs, T), dtype=torch.int32, device=self.device)
            elif name == "ctx":
                t = torch.empty((n_runs, T, self.D), dtype=torch.float64, device=self.device)
            else:
                t = torch.empty((n_runs, T, self.P), dtype=_LOG_DTYPES[name], device=self.device)
            out[name] = t
            setattr(log, name, t.data_ptr())
        return log, out

    def simulate(self, seed, iteration, T, log_fields=None):
        """T rounds for every run with in-kernel Philox noise (replaces src/main.py:116-117).

        Returns the dict of device tensors of the detailed log when ``log_fields`` is given."""
        T = int(T)
        self.reserve_rounds(int(self.lib.agym_rounds_in_iteration(self.handle)) + T)
        log, out = (None, None)
        if log_fields:
            log, out = self._alloc_log(self.R, T, log_fields)
        self._check(self.lib.agym_simulate_rounds(self.handle, C.c_uint64(int(seed) & (2**64 - 1)), int(iteration), T,
                                                  C.byref(log) if log is not None else None, self._stream()))
        return out

    def replay(self, ctx, parts, u, ts_eps=None, gamma_z=None, grid_u=None, run0=0, log_fields=tuple(_lib._LOG_FIELDS[:-1]), num_slots=None):
        """Replay mode: host-drawn noise for runs ``[run0, run0 + n_runs)``; inputs have a leading run axis.
        With ``max_slots`` > 1: ``u`` is [n_runs, T, max_slots] and ``num_slots`` [n_runs, T] (Auction.py:30,65)."""
        dev = self.device
        ctx = torch.as_tensor(np.ascontiguousarray(ctx, np.float64)).to(dev)
        n_runs, T = ctx.shape[0], ctx.shape[1]
        parts_t = torch.as_tensor(np.ascontiguousarray(parts, np.int32)).to(dev)
        u_t = torch.as_tensor(np.ascontiguousarray(u, np.float64)).to(dev)
        assert ctx.shape == (n_runs, T, self.D) and parts_t.shape == (n_runs, T, self.P)
        assert u_t.shape == ((n_runs, T) if self.max_slots == 1 else (n_runs, T, self.max_slots)), u_t.shape
        keep = [ctx, parts_t, u_t]
        rin = ReplayInputs(ctx.data_ptr(), parts_t.data_ptr(), None, None, None, u_t.data_ptr(), 0, 0, None)
        if num_slots is not None:
            ns = torch.as_tensor(np.ascontiguousarray(num_slots, np.int32)).to(dev)
            assert ns.shape == (n_runs, T) and self.max_slots > 1
            keep.append(ns)
            rin.num_slots = ns.data_ptr()
        if ts_eps is not None:
            e = torch.as_tensor(np.ascontiguousarray(ts_eps, np.float32)).to(dev)
            assert e.shape == (n_runs, T, self.P, self.I, self.K), e.shape
            keep.append(e)
            rin.ts_eps = e.data_ptr()
        if gamma_z is not None:
            g = torch.as_tensor(np.ascontiguousarray(gamma_z, np.float64)).to(dev)
            keep.append(g)
            rin.gamma_z = g.data_ptr()
        if grid_u is not None:
            gu = torch.as_tensor(np.ascontiguousarray(grid_u, np.float64)).to(dev)
            keep.append(gu)
            rin.grid_u = gu.data_ptr()
            rin.grid_n = gu.shape[-1]
        self.reserve_rounds(int(self.lib.agym_rounds_in_iteration(self.handle)) + T)
        log, out = self._alloc_log(n_runs, T, log_fields)
        self._check(self.lib.agym_replay_rounds(self.handle, int(run0), int(n_runs), T, C.byref(rin), C.byref(log), self._stream()))
        torch.cuda.current_stream(dev).synchronize()  # inputs in `keep` must outlive the kernel
        return out

    def clear_iteration(self):
        """Agent.clear_utility / clear_logs + Auction.clear_revenue for every run (Agent.py:120-129, Auction.py:76)."""
        if self.retention:
            self.acc[:, :, [_lib.M_NET, _lib.M_GROSS]] = 0.0
            self.revenue.zero_()
            self.retain_logs()
        else:
            self._check(self.lib.agym_clear_iteration(self.handle, self._stream()))

    def retain_logs(self):
        """Agent.clear_logs for every agent with ``memory`` (Agent.py:124-129): keep the last memory[a] records, rewind
        the logs, restart the log-derived accumulators from the kept records.  Utilities and revenue are untouched."""
        if not self.retention:
            raise AgymError("retain_logs: no agent has memory > 0")
        if self.rounds_capacity == 0:
            self.reserve_rounds(1)
        self._check(self.lib.agym_retain_logs(self.handle, self._stream()))

    @property
    def rounds_in_iteration(self):
        return int(self.lib.agym_rounds_in_iteration(self.handle))

    # ------------------------------------------------------------------ updates
    def update_allocators(self, max_epochs=0, want_info=True, fit_mode=0):
        """Agent.update -> allocator.update for every (run, learnt agent) (BidderAllocation.py:29-65)."""
        if not self.any_learnt:
            return None
        info = torch.zeros((self.R, self.A, 4), dtype=torch.float32, device=self.device) if want_info else None
        self._check(self.lib.agym_update_allocators(self.handle, int(fit_mode), int(max_epochs), _ptr(info), self._stream()))
        return info

    def update_bidders(self, seed=0, iteration=0, max_epochs=0, want_info=True):
        """Agent.update -> bidder.update for every (run, agent) whose bidder learns (Bidder.py:210-325,369-431,477-615).
        Returns fit_info [R, A, 3, 4]: per stage (win-rate, initialise_policy, policy) {stop epoch, epochs, loss, rows}."""
        if not self.learning_bidders:
            return None
        info = torch.zeros((self.R, self.A, 3, 4), dtype=torch.float32, device=self.device) if want_info else None
        self._check(self.lib.agym_update_bidders(self.handle, C.c_uint64(int(seed) & (2**64 - 1)), int(iteration), int(max_epochs),
                                                 _ptr(info), self._stream()))
        return info

    # ------------------------------------------------------------------ staged kernels
    def staged_round(self, seed, iteration, T, accumulate=True):
        """K1 -> K2 -> K3 -> K4 with the intermediates in HBM; returns them as device tensors."""
        T = int(T)
        N, P, dev = self.R * T, self.P, self.device
        f32, u8 = torch.float32, torch.uint8
        s = C.c_uint64(int(seed) & (2**64 - 1))
        b = {
            "ctx": torch.empty((N, self.D), dtype=f32, device=dev), "parts": torch.empty((N, P), dtype=u8, device=dev),
            "item": torch.empty((N, P), dtype=u8, device=dev), "est": torch.empty((N, P), dtype=f32, device=dev),
            "true_ctr": torch.empty((N, P), dtype=f32, device=dev), "best_ev": torch.empty((N, P), dtype=f32, device=dev),
            "value": torch.empty((N, P), dtype=f32, device=dev), "bid": torch.empty((N, P), dtype=f32, device=dev),
            "gamma": torch.empty((N, P), dtype=f32, device=dev), "propensity": torch.empty((N, P), dtype=f32, device=dev),
            "winner": torch.empty((N,), dtype=u8, device=dev), "price": torch.empty((N,), dtype=f32, device=dev),
            "second": torch.empty((N,), dtype=f32, device=dev), "outcome": torch.empty((N,), dtype=u8, device=dev),
        }
        st = self._stream()
        it = int(iteration)
        self._check(self.lib.agym_k1_contexts(self.handle, s, it, T, _ptr(b["ctx"]), _ptr(b["parts"]), st))
        self._check(self.lib.agym_k2_allocate(self.handle, s, it, T, _ptr(b["ctx"]), _ptr(b["parts"]), _ptr(b["item"]),
                                              _ptr(b["est"]), _ptr(b["true_ctr"]), _ptr(b["best_ev"]), _ptr(b["value"]), st))
        self._check(self.lib.agym_k3_bids(self.handle, s, it, T, _ptr(b["parts"]), _ptr(b["est"]), _ptr(b["value"]),
                                          _ptr(b["bid"]), _ptr(b["gamma"]), _ptr(b["propensity"]), st))
        self.k4_resolve(s, it, T, b, accumulate)
        return b

    def k4_resolve(self, seed, iteration, T, b, accumulate=True):
        s = seed if isinstance(seed, C.c_uint64) else C.c_uint64(int(seed) & (2**64 - 1))
        self._check(self.lib.agym_k4_resolve(self.handle, s, int(iteration), int(T), _ptr(b["bid"]), _ptr(b["true_ctr"]),
                                             _ptr(b["value"]), _ptr(b["parts"]), _ptr(b["winner"]), _ptr(b["price"]),
                                             _ptr(b["second"]), _ptr(b["outcome"]), 1 if accumulate else 0, self._stream()))

    # ------------------------------------------------------------------ one context at a time (the reference's per-call surface)
    def estimate_ctr(self, agent, context, sample=False, eps=None, seed=0, iteration=0, query=0, run=0):
        """Allocator.estimate_CTR for one context (include/agym.h: agym_estimate_ctr): numpy float64 [I]."""
        ctx = np.ascontiguousarray(context, dtype=np.float64)
        need = (self.D if self.alloc_kind[agent] == _lib.ALLOC_ORACLE else self.Do) + 1
        if ctx.shape != (need,):
            raise ValueError(f"context has shape {ctx.shape}, expected ({need},) (the trailing 1 included, Auction.py:33-49)")
        e = None if eps is None else np.ascontiguousarray(eps, dtype=np.float32)
        if e is not None and e.shape != (self.I, self.Do + 1):
            raise ValueError(f"eps has shape {e.shape}, expected {(self.I, self.Do + 1)}")
        out = np.empty(self.I, np.float64)
        self._check(self.lib.agym_estimate_ctr(self.handle, int(run), int(agent), ctx.ctypes.data, 1 if sample else 0,
                                               None if e is None else e.ctypes.data, C.c_uint64(int(seed) & (2**64 - 1)), int(iteration),
                                               int(query), out.ctypes.data, self._stream()))
        return out

    def bid_one(self, agent, estimated_ctr, value, seed=0, iteration=0, run=0):
        """Bidder.bid for one (value, estimated CTR) of one agent through the staged K3 kernel (a T = 1 launch):
        returns (bid, gamma, propensity); gamma / propensity are NaN for a TruthfulBidder."""
        R, P, dev = self.R, self.P, self.device
        others = [a for a in range(self.A) if a != agent][:P - 1]
        parts = torch.tensor([[agent] + others] * R, dtype=torch.uint8, device=dev)
        est = torch.zeros((R, P), dtype=torch.float32, device=dev)
        val = torch.zeros((R, P), dtype=torch.float32, device=dev)
        est[:, 0], val[:, 0] = float(estimated_ctr), float(value)
        bid, gamma, prop = (torch.empty((R, P), dtype=torch.float32, device=dev) for _ in range(3))
        self._check(self.lib.agym_k3_bids(self.handle, C.c_uint64(int(seed) & (2**64 - 1)), int(iteration), 1, _ptr(parts), _ptr(est), _ptr(val),
                                          _ptr(bid), _ptr(gamma), _ptr(prop), self._stream()))
        return float(bid[run, 0]), float(gamma[run, 0]), float(prop[run, 0])

    # ------------------------------------------------------------------ K8: the metric gather (NCCL through the C ABI)
    def comm_init(self, rank=None, world=None):
        """Create this engine's NCCL communicator (include/agym.h: agym_comm_init).  Under torch.distributed the unique id is
        broadcast from rank 0; with world == 1 no process group is needed."""
        import ctypes

        if world is None:
            import torch.distributed as dist

            rank, world = (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)
        buf = ctypes.create_string_buffer(128)
        if rank == 0:
            rc = self.lib.agym_nccl_unique_id(buf)
            if rc != 0:
                raise AgymError(f"agym_nccl_unique_id failed ({rc}): {self.lib.agym_last_error(None).decode()}")
        ident = [buf.raw]
        if world > 1:
            import torch.distributed as dist

            dist.broadcast_object_list(ident, src=0)
        self._check(self.lib.agym_comm_init(self.handle, ident[0], int(rank), int(world)))
        self.comm_world = int(world)
        self.gathered_acc = torch.empty((self.comm_world, self.R, self.A, _lib.NUM_METRICS), dtype=torch.float64, device=self.device)
        self.gathered_revenue = torch.empty((self.comm_world, self.R), dtype=torch.float64, device=self.device)

    def gather_metrics(self):
        """All-gather of every rank's accumulator block and revenue (agym_gather_metrics_nccl) on the current stream:
        returns device tensors [world, R, A, NUM_METRICS] and [world, R]."""
        if not getattr(self, "comm_world", 0):
            raise AgymError("gather_metrics: call comm_init first")
        self._check(self.lib.agym_gather_metrics_nccl(self.handle, _ptr(self.gathered_acc), _ptr(self.gathered_revenue), self._stream()))
        return self.gathered_acc, self.gathered_revenue

    def gather_block(self, block):
        """All-gather of a float64 device tensor the caller kept (e.g. the metric blocks of every iteration of a job) through
        agym_gather_block_nccl on the current stream: returns [world, *block.shape]."""
        if not getattr(self, "comm_world", 0):
            raise AgymError("gather_block: call comm_init first")
        if block.dtype != torch.float64 or not block.is_contiguous() or not block.is_cuda:
            raise AgymError("gather_block: needs a contiguous float64 device tensor")
        out = torch.empty((self.comm_world,) + tuple(block.shape), dtype=torch.float64, device=block.device)
        self._check(self.lib.agym_gather_block_nccl(self.handle, _ptr(block), _ptr(out), block.numel(), self._stream()))
        return out

    # ------------------------------------------------------------------ results
    def metrics(self):
        """(acc [R, A, NUM_METRICS], revenue [R]) as numpy arrays (one D2H copy each)."""
        return self.acc.cpu().numpy(), self.revenue.cpu().numpy()
