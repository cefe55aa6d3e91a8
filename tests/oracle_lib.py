"""ctypes binding of the CPU oracle (oracle/libsmc_oracle.so).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_SO = os.path.join(ORACLE_DIR, "libsmc_oracle.so")

f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build(force=False):
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("smc_oracle.c", "as_model.c", "normal_table.h")]
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"], env={**os.environ, "CC": "gcc"})
    return _SO


class StageIO(C.Structure):
    _fields_ = [
        ("phi_n1", C.c_double), ("phi_n", C.c_double),
        ("threshold_ratio", C.c_double), ("target", C.c_double), ("alpha", C.c_double),
        ("tempering_target", C.c_double), ("pw", C.c_double), ("log_prob_old_data", C.c_double),
        ("n_mh_steps", C.c_int), ("n_blocks", C.c_int), ("resample_method", C.c_int),
        ("adaptive", C.c_int), ("has_old", C.c_int), ("nthreads", C.c_int),
        ("seed", C.c_uint64), ("stage", C.c_uint32),
        ("c", C.c_double), ("accept", C.c_double), ("ess_prev", C.c_double),
        ("resampled_last", C.c_int), ("j", C.c_int64), ("phi_prop", C.c_double),
        ("ess", C.c_double), ("sum_w", C.c_double), ("phi_out", C.c_double),
        ("resampled", C.c_int), ("status", C.c_int),
    ]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(build())
    d, i64, i32, u64, u32, vp = C.c_double, C.c_int64, C.c_int32, C.c_uint64, C.c_uint32, C.c_void_p
    sig = {
        "orc_exp": (d, [d]), "orc_log": (d, [d]), "orc_sincos2pi_v": (None, [d, f64p]),
        "orc_philox4x32_10": (None, [np.ctypeslib.ndpointer(np.uint32), np.ctypeslib.ndpointer(np.uint32),
                                     np.ctypeslib.ndpointer(np.uint32)]),
        "orc_normal_pair": (None, [u64, u32, u32, u32, f64p]),
        "orc_normal_quad": (None, [u64, u32, u32, u32, f64p]),
        "orc_uniform": (d, [u64, u32, u32, u32, u32, C.c_int]),
        "orc_canon_sum": (d, [f64p, i64]), "orc_canon_sumsq": (d, [f64p, i64]),
        "orc_canon_sum_generic": (d, [f64p, i64, C.c_int, C.c_int]),
        "orc_cumsum": (None, [f64p, i64, f64p]),
        "orc_correct": (C.c_int, [f64p, i64, C.c_int, d, d, d, d, vp, vp, f64p]),
        "orc_compute_ess": (d, [f64p, f64p, f64p, i64, d, d, f64p]),
        "orc_solve_adaptive_phi": (C.c_int, [f64p, i64, C.c_int, f64p, C.c_int, C.POINTER(i64), C.POINTER(d), d, d, d,
                                            C.c_int, C.POINTER(d), C.POINTER(i64)]),
        "orc_resample": (C.c_int, [f64p, i64, C.c_int, u64, u32, d, i64p, vp]),
        "orc_resample_n": (C.c_int, [f64p, i64, i64, C.c_int, u64, u32, d, i64p, vp]),
        "orc_gather": (None, [f64p, f64p, i64, C.c_int, i64p]),
        "orc_update_c": (d, [d, d, d]),
        "orc_moments": (None, [f64p, i64, C.c_int, f64p, f64p]),
        "orc_moments_shifted": (None, [f64p, i64, C.c_int, f64p, f64p, f64p]),
        "orc_mean_accept": (d, [f64p, i64, C.c_int, C.c_int]),
        "orc_cholesky": (C.c_int, [f64p, C.c_int, f64p]),
        "orc_model_create": (vp, [C.c_int]), "orc_model_free": (None, [vp]),
        "orc_model_set_params": (C.c_int, [vp, i32p, f64p, f64p, i32p, f64p, f64p]),
        "orc_model_set_gaussreg": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, f64p]),
        "orc_model_set_as": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, f64p]),
        "orc_as_loglik": (d, [f64p, f64p, C.c_int, C.c_int]),
        "orc_logprior": (d, [vp, f64p]), "orc_loglik": (d, [vp, C.c_int, f64p]),
        "orc_loglik_lineq_direct": (d, [f64p, f64p, f64p, C.c_int, C.c_int]),
        "orc_loglik_linreg_direct": (d, [f64p, f64p, f64p, C.c_int, C.c_int, d]),
        "orc_proposal_create": (vp, [C.c_int, C.c_int, f64p, f64p, C.c_int, i32p, i32p, i32p, d, C.POINTER(C.c_int)]),
        "orc_proposal_free": (None, [vp]),
        "orc_generate_blocks": (None, [C.c_int, C.c_int, u64, u32, i32p, i32p]),
        "orc_proposal_densities": (None, [f64p, f64p, f64p, C.c_int, d, f64p, f64p, d, C.POINTER(d), C.POINTER(d)]),
        "orc_mutate": (None, [vp, vp, f64p, i64, i64, d, d, d, C.c_int, C.c_int, C.c_int, u64, u32, C.c_int]),
        "orc_max_threads": (C.c_int, []),
        "orc_initialize_likelihoods": (None, [vp, f64p, i64]),
        "orc_evaluate": (None, [vp, f64p, i64]),
        "orc_initial_draw": (C.c_int, [vp, f64p, i64, i64, f64p, u64, C.c_int]),
        "orc_stage": (C.c_int, [vp, f64p, f64p, i64, f64p, C.c_int, C.POINTER(StageIO), vp, vp, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Model:
    """Oracle-side model: priors + Gaussian-regression likelihood slots (0 = data, 1 = old data)."""

    def __init__(self, spec):
        """spec: smc_jl_b200.model.ModelSpec (plain arrays; shared description of the model)."""
        self.spec = spec
        L = lib()
        self.h = L.orc_model_create(spec.d)
        assert self.h
        st = L.orc_model_set_params(self.h, spec.fixed, spec.lo, spec.hi, spec.kind, spec.p1, spec.p2)
        assert st == 0
        for slot, lk in enumerate(spec.liks):
            if lk is None:
                continue
            if lk.kind == 2:    # An-Schorfheide DSGE
                st = L.orc_model_set_as(self.h, slot, lk.data.shape[1], lk.n_presample, lk.eqdata)
            else:
                st = L.orc_model_set_gaussreg(self.h, slot, lk.neq, lk.k, lk.stride, lk.coef_off, lk.sig_off,
                                              np.ascontiguousarray(lk.eqdata))
            assert st == 0

    def __del__(self):
        try:
            lib().orc_model_free(self.h)
        except Exception:
            pass

    def logprior(self, theta):
        return lib().orc_logprior(self.h, np.ascontiguousarray(theta, dtype=np.float64))

    def loglik(self, theta, slot=0):
        return lib().orc_loglik(self.h, slot, np.ascontiguousarray(theta, dtype=np.float64))


def cloud_f(a):
    """particles (N x (d+5)) -> flat column-major buffer (the Julia matrix bytes)."""
    return np.array(np.asarray(a, dtype=np.float64).T, order="C", copy=True).ravel()


def cloud_m(buf, N, d):
    return buf.reshape(d + 5, N).T.copy()
