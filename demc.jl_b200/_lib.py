"""ctypes binding of libdemcz_cuda.so (include/demcz.h).  Fails loudly when the CUDA library is
missing: there is no CPU fallback on the product path."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdemcz_cuda.so")

OK = 0
ERR = {-1: "INVALID", -2: "CUDA", -3: "NCCL", -4: "OOM", -5: "NOT_PD", -6: "UNSUPPORTED"}


class DemczError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libdemcz_cuda: {ERR.get(code, code)}: {msg}")
        self.code = code


class Opts(C.Structure):
    """demcz_opts (include/demcz.h) — mirror of DEMCopt, src/DEMC.jl:24-39."""
    _fields_ = [
        ("struct_size", C.c_int32), ("N", C.c_int32), ("K", C.c_int32), ("Ngeneration", C.c_int32),
        ("Nblocks", C.c_int32), ("d", C.c_int32),
        ("block_ptr", C.POINTER(C.c_int32)), ("block_idx", C.POINTER(C.c_int32)), ("eps_scale", C.POINTER(C.c_double)),
        ("gamma", C.c_double), ("T0", C.c_double), ("TN", C.c_double),
        ("anneal", C.c_int32), ("autostop", C.c_int32), ("autostop_every", C.c_int32), ("autostop_Rhat", C.c_double),
        ("adapt", C.c_int32), ("min_gamma", C.c_double), ("max_gamma", C.c_double), ("adapt_every", C.c_int32),
        ("seed", C.c_uint64), ("gen_offset", C.c_int64),
        ("store_chain", C.c_int32), ("init_mode", C.c_int32), ("archive_capacity", C.c_int64),
        ("trace", C.c_int32), ("profile", C.c_int32),
        ("rank", C.c_int32), ("world", C.c_int32), ("N_global", C.c_int64), ("chain_offset", C.c_int64),
        ("sync_mode", C.c_int32), ("exchange_layout", C.c_int32),
        ("push_engine", C.c_int32), ("thin", C.c_int32), ("moments", C.c_int32), ("persistent", C.c_int32),
        ("snooker_prob", C.c_double), ("snooker_gamma", C.c_double),
    ]


class RunInfo(C.Structure):
    _fields_ = [
        ("generations_done", C.c_int64), ("generations_total", C.c_int64), ("stopped_by_rhat", C.c_int32),
        ("rhat_valid", C.c_int32), ("rhat_max", C.c_double), ("final_gamma", C.c_double), ("M", C.c_int64),
        ("proposals", C.c_int64), ("accepted", C.c_int64), ("step_launches", C.c_int64),
        ("step_kernel_ms", C.c_double), ("device_ms", C.c_double), ("exchange_ms", C.c_double), ("exchanges", C.c_int64),
    ]


EXPORTS = [
    "demcz_version", "demcz_create", "demcz_destroy", "demcz_last_error", "demcz_set_likelihood_mvnormal",
    "demcz_set_likelihood_linreg", "demcz_set_likelihood_logistic", "demcz_set_archive", "demcz_set_state", "demcz_set_temperature_table",
    "demcz_run", "demcz_get_kernel_info", "demcz_set_profile", "demcz_get_chain", "demcz_get_state", "demcz_get_archive", "demcz_get_moments", "demcz_get_archive_moments",
    "demcz_get_history_moments", "demcz_get_archive_checksum", "demcz_archive_view_rule", "demcz_get_rhat",
    "demcz_get_accept_counts", "demcz_eval_logdensity", "demcz_get_trace", "demcz_comm_unique_id", "demcz_comm_init", "demcz_comm_peer_info", "demcz_comm_open_peers",
]

_lib = None


def load():
    """Load libdemcz_cuda.so (built in-tree by __graft_entry__.build() / csrc/Makefile)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `make -C {_HERE}/csrc` "
                          "(python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, dp, i64 = C.c_void_p, C.POINTER(C.c_double), C.c_int64
    lib.demcz_version.restype = C.c_int
    lib.demcz_last_error.restype = C.c_char_p
    lib.demcz_last_error.argtypes = [vp]
    lib.demcz_create.argtypes = [C.POINTER(vp), C.POINTER(Opts), C.c_int]
    lib.demcz_destroy.argtypes = [vp]
    lib.demcz_destroy.restype = None
    lib.demcz_set_likelihood_mvnormal.argtypes = [vp, C.c_int32, dp, dp, C.c_int32]
    lib.demcz_set_likelihood_linreg.argtypes = [vp, i64, C.c_int32, dp, dp]
    lib.demcz_set_likelihood_logistic.argtypes = [vp, i64, C.c_int32, dp, dp]
    lib.demcz_set_archive.argtypes = [vp, dp, i64]
    lib.demcz_set_state.argtypes = [vp, dp, dp]
    lib.demcz_set_temperature_table.argtypes = [vp, dp, i64]
    lib.demcz_run.argtypes = [vp, i64, C.POINTER(RunInfo)]
    lib.demcz_set_profile.argtypes = [vp, C.c_int32]
    lib.demcz_get_kernel_info.argtypes = [vp, C.c_char_p, i64]
    lib.demcz_get_chain.argtypes = [vp, dp, dp]
    lib.demcz_get_state.argtypes = [vp, dp, dp]
    lib.demcz_get_archive.argtypes = [vp, dp, C.POINTER(i64)]
    lib.demcz_get_moments.argtypes = [vp, dp, dp]
    lib.demcz_get_archive_moments.argtypes = [vp, i64, i64, dp, dp]
    lib.demcz_get_history_moments.argtypes = [vp, dp, dp, C.POINTER(i64), C.c_int32]
    lib.demcz_get_archive_checksum.argtypes = [vp, i64, i64, C.POINTER(C.c_uint64)]
    lib.demcz_archive_view_rule.argtypes = [i64, i64, i64, C.c_int32, i64, i64, C.POINTER(i64), C.POINTER(i64)]
    lib.demcz_get_rhat.argtypes = [vp, dp]
    lib.demcz_get_accept_counts.argtypes = [vp, C.POINTER(i64)]
    lib.demcz_eval_logdensity.argtypes = [vp, dp, i64, dp]
    lib.demcz_get_trace.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    lib.demcz_comm_unique_id.argtypes = [C.POINTER(C.c_uint8)]
    lib.demcz_comm_init.argtypes = [vp, C.POINTER(C.c_uint8)]
    lib.demcz_comm_peer_info.argtypes = [vp, C.POINTER(C.c_uint8)]
    lib.demcz_comm_open_peers.argtypes = [vp, C.POINTER(C.c_uint8)]
    _lib = lib
    return lib
