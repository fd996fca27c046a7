"""ctypes binding of libagym.so (the C ABI declared in include/agym.h).

There is no CPU fallback: if the shared library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AGYM_LIB_PATH") or os.path.join(_HERE, "libagym.so")  # the override is for A/B builds of the library

NUM_METRICS = 12
BIDDER_D = 4
BIDDER_W = 16
BID_ROW = 5
TERM_ROW = 8         # AGYM_TERM_ROW
FIT_ADAM_REF, FIT_ADAM_FAST, FIT_NEWTON = 0, 1, 2  # FIT_NEWTON: opt-in, a different algorithm (include/agym.h)
(BFIT_NONE, BFIT_VL_SEARCH, BFIT_VL_POLICY, BFIT_PL_REINFORCE, BFIT_PL_OFFPOLICY, BFIT_PL_TRPO, BFIT_PL_PPO, BFIT_DR, BFIT_EMPIRICAL) = range(9)
ABI_VERSION = 2

# enum mirrors (include/agym.h)
SECOND_PRICE, FIRST_PRICE = 0, 1
ALLOC_ORACLE, ALLOC_TS, ALLOC_MAP = 0, 1, 2
BID_TRUTHFUL, BID_GAUSS, BID_GAUSS_CLIP, BID_SEARCH, BID_BANDIT, BID_POLICY = range(6)
FP32, FP64 = 0, 1
(M_NET, M_GROSS, M_ALLOC_REGRET, M_ESTIM_REGRET, M_OVERBID_REGRET, M_UNDERBID_REGRET, M_SQERR, M_BIAS, M_NPART,
 M_NWON, M_BEST_EV, M_GAMMA) = range(NUM_METRICS)


class AgymError(RuntimeError):
    pass


class Shape(C.Structure):
    _fields_ = [("R", C.c_int32), ("A", C.c_int32), ("I", C.c_int32), ("D", C.c_int32), ("Do", C.c_int32),
                ("P", C.c_int32), ("mechanism", C.c_int32), ("precision", C.c_int32), ("run_offset", C.c_int32),
                ("max_slots", C.c_int32), ("embedding_var", C.c_double)]


_LOG_FIELDS = ["agent", "item", "est", "value", "bid", "true_ctr", "best_ev", "price", "second", "gamma",
               "propensity", "outcome", "won", "winner", "ctx"]


class RoundLog(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in _LOG_FIELDS]


class ReplayInputs(C.Structure):
    _fields_ = [("ctx", C.c_void_p), ("parts", C.c_void_p), ("ts_eps", C.c_void_p), ("gamma_z", C.c_void_p),
                ("grid_u", C.c_void_p), ("u", C.c_void_p), ("grid_n", C.c_int32), ("reserved", C.c_int32), ("num_slots", C.c_void_p)]


# name -> (restype, argtypes); every symbol include/agym.h declares
_H = C.c_void_p
SIGNATURES = {
    "agym_abi_version": (C.c_int, []),
    "agym_last_error": (C.c_char_p, [_H]),
    "agym_create": (C.c_int, [C.POINTER(Shape), C.c_int, C.POINTER(_H)]),
    "agym_destroy": (C.c_int, [_H]),
    "agym_set_agents": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p]),
    "agym_set_catalog": (C.c_int, [_H, C.c_void_p, C.c_void_p]),
    "agym_bind_allocator_state": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "agym_refresh_sigma": (C.c_int, [_H, C.c_void_p]),
    "agym_bind_bidder_state": (C.c_int, [_H, C.c_void_p, C.c_void_p]),
    "agym_bind_metrics": (C.c_int, [_H, C.c_void_p, C.c_void_p]),
    "agym_bind_fit_log": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_int64]),
    "agym_workspace_bytes": (C.c_size_t, [_H, C.c_int64]),
    "agym_bind_workspace": (C.c_int, [_H, C.c_void_p, C.c_size_t]),
    "agym_simulate_rounds": (C.c_int, [_H, C.c_uint64, C.c_int32, C.c_int64, C.POINTER(RoundLog), C.c_void_p]),
    "agym_replay_rounds": (C.c_int, [_H, C.c_int32, C.c_int32, C.c_int64, C.POINTER(ReplayInputs), C.POINTER(RoundLog), C.c_void_p]),
    "agym_rounds_in_iteration": (C.c_int64, [_H]),
    "agym_set_rounds_in_iteration": (C.c_int, [_H, C.c_int64]),
    "agym_clear_iteration": (C.c_int, [_H, C.c_void_p]),
    "agym_set_log_retention": (C.c_int, [_H, C.c_void_p, C.c_void_p]),
    "agym_retained_capacity": (C.c_int64, [_H]),
    "agym_retain_logs": (C.c_int, [_H, C.c_void_p]),
    "agym_estimate_ctr": (C.c_int, [_H, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_uint64, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p]),
    "agym_nccl_unique_id": (C.c_int, [C.c_char_p]),
    "agym_comm_init": (C.c_int, [_H, C.c_char_p, C.c_int32, C.c_int32]),
    "agym_gather_metrics_nccl": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p]),
    "agym_gather_block_nccl": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "agym_launch_count": (C.c_uint64, [_H]),
    "agym_set_option": (C.c_int, [_H, C.c_char_p, C.c_double]),
    "agym_update_allocators": (C.c_int, [_H, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "agym_bind_bid_log": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_int64]),
    "agym_bidder_workspace_bytes": (C.c_size_t, [_H, C.c_int64]),
    "agym_bind_bidder_workspace": (C.c_int, [_H, C.c_void_p, C.c_size_t]),
    "agym_set_bidder_fits": (C.c_int, [_H, C.c_void_p]),
    "agym_update_bidders": (C.c_int, [_H, C.c_uint64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "agym_k1_contexts": (C.c_int, [_H, C.c_uint64, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "agym_k2_allocate": (C.c_int, [_H, C.c_uint64, C.c_int32, C.c_int64] + [C.c_void_p] * 8),
    "agym_k3_bids": (C.c_int, [_H, C.c_uint64, C.c_int32, C.c_int64] + [C.c_void_p] * 7),
    "agym_k4_resolve": (C.c_int, [_H, C.c_uint64, C.c_int32, C.c_int64] + [C.c_void_p] * 8 + [C.c_int32, C.c_void_p]),
}

_lib = None


def load():
    """Load libagym.so (built in-tree by __graft_entry__.build()); raises if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise AgymError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the export is missing
        fn.restype = res
        fn.argtypes = args
    v = lib.agym_abi_version()
    if v != ABI_VERSION:
        raise AgymError(f"libagym ABI version {v}, binding expects {ABI_VERSION}")
    _lib = lib
    return lib


def check(rc, handle=None):
    if rc != 0:
        msg = load().agym_last_error(handle)
        raise AgymError(f"libagym error {rc}: {msg.decode() if msg else '?'}")
