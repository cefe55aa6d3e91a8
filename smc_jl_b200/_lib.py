"""ctypes binding of libsmcb200.so (C ABI declared in include/smcb200.h).

The library must be present: importing this module without it raises ImportError (there is no
CPU fallback).  Loading works without a GPU (so the symbol table can be checked on CPU); creating
a context without a GPU fails with SMCB200_ERR_CUDA.
"""
import ctypes as C
import os
import re

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SMCB200_LIB", os.path.join(_HERE, "libsmcb200.so"))   # env override: developer experiments only
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "smcb200.h")

OK, ERR_NAN_ESS, ERR_BAD_RESAMPLER, ERR_BAD_ARGUMENT, ERR_NOT_POSDEF, ERR_CUDA, ERR_NCCL, ERR_UNSUPPORTED, ERR_NOT_READY = range(9)


class StageConfig(C.Structure):
    _fields_ = [
        ("phi_n1", C.c_double), ("phi_n", C.c_double), ("threshold_ratio", C.c_double), ("target", C.c_double),
        ("alpha", C.c_double), ("tempering_target", C.c_double), ("prior_weight", C.c_double),
        ("log_prob_old_data", C.c_double),
        ("n_mh_steps", C.c_int32), ("n_blocks", C.c_int32), ("resample_method", C.c_int32), ("adaptive", C.c_int32),
        ("has_old_data", C.c_int32), ("reserved", C.c_int32),
        ("seed", C.c_uint64), ("stage", C.c_uint32), ("reserved2", C.c_uint32),
    ]


class StageState(C.Structure):
    _fields_ = [
        ("c", C.c_double), ("accept", C.c_double), ("ess_prev", C.c_double), ("phi_prop", C.c_double),
        ("j", C.c_int64), ("resampled_last_period", C.c_int32), ("reserved", C.c_int32),
    ]


class StageResult(C.Structure):
    _fields_ = [
        ("phi_n", C.c_double), ("ess", C.c_double), ("sum_weights", C.c_double), ("c", C.c_double), ("accept", C.c_double),
        ("resampled", C.c_int32), ("status", C.c_int32),
        ("ms_correct", C.c_float), ("ms_resample", C.c_float), ("ms_moments", C.c_float), ("ms_mutate", C.c_float),
    ]


def declared_symbols():
    """Every function name declared in include/smcb200.h."""
    txt = open(HEADER_PATH).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(smcb200_[a-z0-9_]+)\s*\(", txt)))


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libsmcb200.so is missing (%s). Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C smc_jl_b200/csrc`. There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, d, i32, i64, u32, u64 = C.c_void_p, C.c_double, C.c_int32, C.c_int64, C.c_uint32, C.c_uint64
    pd, pi32, pi64 = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_int64)
    sig = {
        "smcb200_abi_version": (i32, []),
        "smcb200_status_string": (C.c_char_p, [i32]),
        "smcb200_create": (i32, [C.POINTER(vp), i32]),
        "smcb200_destroy": (i32, [vp]),
        "smcb200_last_error": (C.c_char_p, [vp]),
        "smcb200_comm_unique_id": (i32, [vp]),
        "smcb200_comm_init": (i32, [vp, i32, i32, vp]),
        "smcb200_cloud_create": (i32, [vp, i64, i32]),
        "smcb200_cloud_shard": (i32, [vp, pi64, pi64]),
        "smcb200_cloud_upload": (i32, [vp, vp, i64, i64]),
        "smcb200_cloud_download": (i32, [vp, vp, i64, i64]),
        "smcb200_cloud_read_column": (i32, [vp, i32, vp]),
        "smcb200_cloud_write_column": (i32, [vp, i32, vp]),
        "smcb200_set_parameters": (i32, [vp, i32, vp, vp, vp, vp, vp, vp]),
        "smcb200_set_likelihood": (i32, [vp, i32, i32, vp, i32, vp, i64]),
        "smcb200_evaluate": (i32, [vp, i32]),
        "smcb200_initial_draw": (i32, [vp, vp, u64, i32]),
        "smcb200_correct": (i32, [vp, d, d, d, d, vp, vp, vp]),
        "smcb200_ess_at": (i32, [vp, vp, i32, d, vp]),
        "smcb200_solve_adaptive_phi": (i32, [vp, vp, i32, pi64, pd, d, d, d, i32, pd]),
        "smcb200_resample": (i32, [vp, i32, u64, u32, d, vp]),
        "smcb200_resample_weights": (i32, [vp, vp, i64, i32, u64, u32, d, vp, vp]),
        "smcb200_resample_weights_n": (i32, [vp, vp, i64, i64, i32, u64, u32, d, vp, vp]),
        "smcb200_moments": (i32, [vp, vp, vp]),
        "smcb200_moments_onepass": (i32, [vp, vp, vp]),
        "smcb200_run_stages": (i32, [vp, C.POINTER(StageConfig), C.POINTER(StageState), vp, i32, i32, i32, vp, vp, i64,
                                     C.POINTER(StageResult), C.POINTER(i32)]),
        "smcb200_fp64_peak": (i32, [vp, i32, pd]),
        "smcb200_mutate": (i32, [vp, vp, vp, i32, i32, vp, vp, vp, d, d, d, d, i32, i32, u64, u32, pd]),
        "smcb200_stage": (i32, [vp, C.POINTER(StageConfig), C.POINTER(StageState), vp, i32, vp, vp, C.POINTER(StageResult)]),
        "smcb200_stage_host": (i32, [vp, vp, i64, C.POINTER(StageConfig), C.POINTER(StageState), vp, i32,
                                     C.POINTER(StageResult)]),
        "smcb200_kernel_launches": (i64, [vp]),
        "smcb200_last_kernel_ms": (i32, [vp, i32, C.POINTER(C.c_float)]),
        "smcb200_timer_start": (i32, [vp]),
        "smcb200_timer_stop": (i32, [vp, C.POINTER(C.c_float)]),
        "smcb200_debug_math": (i32, [vp, i32, vp, i64, u64, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)   # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    return L, sorted(sig)


lib, BOUND_SYMBOLS = _load()


def ptr(a):
    """void* of a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    assert isinstance(a, np.ndarray)
    return a.ctypes.data_as(C.c_void_p)
