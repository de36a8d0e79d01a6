"""ctypes binding of the C-ABI in include/bae_b200.h (libbae_b200.so, built in-tree for sm_100a).

There is no CPU fallback: if the shared library is missing, or no sm_100 device is present when a
handle is created, this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# BAE_B200_LIB selects another build of the same library (A/B experiments: tools/); the default is the in-tree build
LIB_PATH = os.environ.get("BAE_B200_LIB") or os.path.join(_HERE, "lib", "libbae_b200.so")

EXPORTS = [
    "bae_last_error", "bae_version", "bae_create", "bae_destroy", "bae_w_size", "bae_n_trainable",
    "bae_upload_data", "bae_set_ladder", "bae_init_chain", "bae_get_state", "bae_set_state", "bae_step",
    "bae_swap_local", "bae_run", "bae_synchronize", "bae_eval", "bae_debug_draws", "bae_debug_swap_uniform",
    "bae_read_traces", "bae_read_traces_all", "bae_swap_counts", "bae_exchange_pack", "bae_exchange_unpack", "bae_exchange_finish",
    "bae_launch_count", "bae_last_step_ms", "bae_gemm_path", "bae_stream", "bae_debug_buffer", "bae_debug_trace", "bae_debug_profiler", "bae_debug_phase_times", "bae_exchange_table", "bae_exchange_local", "bae_host_sweep", "bae_last_sweep", "bae_forward_all", "bae_shard_config", "bae_swap_exchange", "bae_plan_exchange",
    "bae_exchange_vectors_sent", "bae_step_kernel", "bae_eval_bn_stats", "bae_mma_kind",
]


class BaeError(RuntimeError):
    pass


class Config(C.Structure):
    _fields_ = [
        ("dims", C.c_int32 * 4), ("n_train", C.c_int32), ("n_test", C.c_int32), ("n_chains", C.c_int32),
        ("n_rungs", C.c_int32), ("chain_id_base", C.c_int32), ("samples", C.c_int32),
        ("pt_switch_step", C.c_int32), ("swap_interval", C.c_int32), ("use_langevin", C.c_int32),
        ("lr", C.c_float), ("step_w", C.c_float), ("step_eta", C.c_float), ("l_prob", C.c_float),
        ("sigma_squared", C.c_float), ("nu_1", C.c_float), ("nu_2", C.c_float),
        ("seed", C.c_uint64), ("gemm_path", C.c_int32), ("drift_mode", C.c_int32),
        ("act_group", C.c_int32), ("reserved", C.c_int32),
    ]


class Trace(C.Structure):
    _fields_ = [
        ("likelihood_proposal", C.c_double), ("likelihood", C.c_double), ("sum_value", C.c_double),
        ("log_u", C.c_double), ("prior_proposal", C.c_double), ("diff_prop", C.c_double),
        ("mse_train", C.c_double), ("mse_test", C.c_double), ("mse_train_proposal", C.c_double),
        ("mse_test_proposal", C.c_double), ("eta", C.c_double), ("weights", C.c_float * 13),
        ("accepted", C.c_int32), ("langevin", C.c_int32), ("step", C.c_int32),
    ]


class ChainScalars(C.Structure):
    _fields_ = [
        ("eta", C.c_double), ("likelihood", C.c_double), ("prior_current", C.c_double),
        ("last_sse_train", C.c_double), ("mse_train", C.c_double), ("mse_test", C.c_double),
        ("temperature", C.c_double), ("adapttemp", C.c_double),
        ("adam_t", C.c_int32), ("w_is_alias", C.c_int32), ("init_count", C.c_int32),
        ("num_accepted", C.c_int32), ("langevin_count", C.c_int32), ("step", C.c_int32),
        ("prop_sse_train", C.c_double),
    ]


_lib = None


def load():
    """Load libbae_b200.so (raises BaeError when it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BaeError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    P = C.POINTER
    lib.bae_last_error.restype = C.c_char_p
    lib.bae_create.argtypes = [P(Config), C.c_int, P(C.c_void_p)]
    lib.bae_destroy.argtypes = [C.c_void_p]
    lib.bae_w_size.argtypes = [C.c_void_p]
    lib.bae_n_trainable.argtypes = [C.c_void_p]
    lib.bae_upload_data.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.bae_set_ladder.argtypes = [C.c_void_p, C.c_void_p]
    lib.bae_init_chain.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.bae_get_state.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, P(ChainScalars)]
    lib.bae_set_state.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, P(ChainScalars)]
    lib.bae_step.argtypes = [C.c_void_p, C.c_int]
    lib.bae_swap_local.argtypes = [C.c_void_p]
    lib.bae_run.argtypes = [C.c_void_p, C.c_int]
    lib.bae_synchronize.argtypes = [C.c_void_p]
    lib.bae_eval.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, P(C.c_double), P(C.c_double), C.c_void_p, C.c_void_p]
    lib.bae_debug_draws.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.bae_debug_swap_uniform.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, P(C.c_double)]
    lib.bae_read_traces.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.bae_read_traces_all.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.bae_swap_counts.argtypes = [C.c_void_p, P(C.c_int64), P(C.c_int64)]
    lib.bae_last_sweep.argtypes = [C.c_void_p, C.c_void_p]
    lib.bae_shard_config.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.bae_swap_exchange.argtypes = [C.c_void_p, C.c_void_p]
    lib.bae_plan_exchange.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
    lib.bae_exchange_vectors_sent.argtypes = [C.c_void_p]
    lib.bae_exchange_vectors_sent.restype = C.c_int64
    lib.bae_eval_bn_stats.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.bae_forward_all.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.bae_exchange_table.argtypes = [C.c_void_p, C.c_void_p]
    lib.bae_exchange_local.argtypes = [C.c_void_p, C.c_void_p]
    lib.bae_host_sweep.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p, P(C.c_int)]
    lib.bae_exchange_pack.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.bae_exchange_unpack.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.bae_exchange_finish.argtypes = [C.c_void_p]
    lib.bae_launch_count.argtypes = [C.c_void_p]
    lib.bae_launch_count.restype = C.c_int64
    lib.bae_last_step_ms.argtypes = [C.c_void_p, P(C.c_float)]
    lib.bae_gemm_path.argtypes = [C.c_void_p]
    lib.bae_step_kernel.argtypes = [C.c_void_p]
    lib.bae_mma_kind.argtypes = [C.c_void_p]
    lib.bae_debug_buffer.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_void_p, C.c_int64, P(C.c_int), C.c_int]
    lib.bae_debug_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    lib.bae_debug_profiler.argtypes = [C.c_int]
    if hasattr(lib, "bae_debug_phase_times"):   # (an older build selected through BAE_B200_LIB for an A/B run may lack it)
        lib.bae_debug_phase_times.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.bae_stream.argtypes = [C.c_void_p]
    lib.bae_stream.restype = C.c_void_p
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        raise BaeError(f"bae error {rc}: {load().bae_last_error().decode()}")
