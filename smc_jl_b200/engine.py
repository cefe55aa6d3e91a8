"""Engine -- thin object wrapper over the C ABI (include/smcb200.h), one per GPU.

Status codes are converted to the exception classes the Julia shim would raise
(SURVEY 8(b) / INTEGRATION.md).
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import StageConfig, StageResult, StageState, lib, ptr
from .model import ModelSpec

# :polyalgo (src/resample.jl:73-75) is `StatsBase.sample(1:n, Weights(w), n)`: n i.i.d. categorical draws through an alias
# table (or direct sampling for small n) -- the same joint distribution as the multinomial resampler's inverse-CDF draws,
# on a random stream that is unpinned either way.  It is served by the multinomial kernel.
RESAMPLERS = {"systematic": 0, "multinomial": 1, "polyalgo": 1}


class NotPosDefError(np.linalg.LinAlgError):
    pass


def _raise(status, handle):
    msg = lib.smcb200_last_error(handle).decode() if handle else ""
    base = lib.smcb200_status_string(status).decode()
    text = "%s%s" % (base, (": " + msg) if msg and msg != base else "")
    if status == _lib.ERR_NAN_ESS:
        raise AssertionError(text)
    if status in (_lib.ERR_BAD_RESAMPLER, _lib.ERR_BAD_ARGUMENT):
        raise ValueError(text)
    if status == _lib.ERR_NOT_POSDEF:
        raise NotPosDefError(text)
    if status == _lib.ERR_UNSUPPORTED:
        raise NotImplementedError(text)
    raise RuntimeError(text)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class Engine:
    def __init__(self, device=0):
        self.h = C.c_void_p()
        st = lib.smcb200_create(C.byref(self.h), int(device))
        if st:
            self.h = None
            raise RuntimeError("smcb200_create(device=%d) failed: %s -- a CUDA GPU is required, there is no CPU "
                               "fallback" % (device, lib.smcb200_status_string(st).decode()))
        self.device = int(device)
        self.n_parts = 0
        self.n_para = 0
        self.spec = None
        self.rank, self.world = 0, 1

    @staticmethod
    def unique_id():
        """128-byte communicator id (rank 0 creates it, the host distributes it to every rank)."""
        buf = C.create_string_buffer(128)
        st = lib.smcb200_comm_unique_id(buf)
        if st:
            raise RuntimeError("smcb200_comm_unique_id failed: " + lib.smcb200_status_string(st).decode())
        return buf.raw

    def comm_init(self, rank, world, comm_id):
        """Join the job's communicator (one process per GPU).  Must precede cloud_create."""
        buf = C.create_string_buffer(bytes(comm_id), 128)
        self._ck(lib.smcb200_comm_init(self.h, int(rank), int(world), buf))
        self.rank, self.world = int(rank), int(world)

    def close(self):
        if getattr(self, "h", None):
            lib.smcb200_destroy(self.h)
            self.h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, st):
        if st:
            _raise(st, self.h)

    # ---- cloud ----------------------------------------------------------------------------------
    def cloud_create(self, n_parts, n_para):
        self._ck(lib.smcb200_cloud_create(self.h, int(n_parts), int(n_para)))
        self.n_parts, self.n_para = int(n_parts), int(n_para)
        first, count = C.c_int64(), C.c_int64()
        self._ck(lib.smcb200_cloud_shard(self.h, C.byref(first), C.byref(count)))
        self.first, self.count = first.value, count.value

    def upload(self, particles):
        """particles: the GLOBAL n_parts x (n_para+5) matrix (this rank copies its shard's rows) or just this
        rank's shard; Fortran-ordered arrays are passed without a copy."""
        p = np.asfortranarray(particles, dtype=np.float64)
        if p.shape == (self.n_parts, self.n_para + 5):
            self._ck(lib.smcb200_cloud_upload(self.h, ptr(p), p.shape[0], self.first))
        else:
            assert p.shape == (self.count, self.n_para + 5), p.shape
            self._ck(lib.smcb200_cloud_upload(self.h, ptr(p), p.shape[0], 0))

    def download(self, out=None):
        """Global-shaped output (only this rank's rows are written) or, with out=None on a sharded engine,
        this rank's shard."""
        if out is None:
            rows = self.n_parts if self.world == 1 else self.count
            out = np.zeros((rows, self.n_para + 5), order="F")
        assert out.flags.f_contiguous and out.shape[1] == self.n_para + 5
        row0 = self.first if out.shape[0] == self.n_parts and self.world > 1 else (self.first if self.world == 1 else 0)
        self._ck(lib.smcb200_cloud_download(self.h, ptr(out), out.shape[0], row0))
        return out

    def read_column(self, col):
        if col < 0:
            col += self.n_para + 5
        out = np.zeros(self.count)
        self._ck(lib.smcb200_cloud_read_column(self.h, col, ptr(out)))
        return out

    def write_column(self, col, v):
        if col < 0:
            col += self.n_para + 5
        v = _f64(v)
        assert v.shape == (self.count,)
        self._ck(lib.smcb200_cloud_write_column(self.h, col, ptr(v)))

    # ---- model ----------------------------------------------------------------------------------
    def set_model(self, spec: ModelSpec):
        self.spec = spec
        self._ck(lib.smcb200_set_parameters(self.h, spec.d, ptr(_i32(spec.fixed)), ptr(_f64(spec.lo)), ptr(_f64(spec.hi)),
                                            ptr(_i32(spec.kind)), ptr(_f64(spec.p1)), ptr(_f64(spec.p2))))
        for slot, lk in enumerate(spec.liks):
            if lk is None:
                continue
            ip, dp = lk.iparams(), _f64(lk.eqdata)
            self._ck(lib.smcb200_set_likelihood(self.h, slot, lk.kind, ptr(ip), len(ip), ptr(dp), dp.size))

    def initial_draw(self, fixed_values, seed, max_tries=1000):
        """initial_draw! (src/initialization.jl:88-119) on the device."""
        fv = _f64(fixed_values)
        self._ck(lib.smcb200_initial_draw(self.h, ptr(fv), int(seed), int(max_tries)))

    def evaluate(self, mode=0):
        self._ck(lib.smcb200_evaluate(self.h, mode))

    # ---- stage operations -------------------------------------------------------------------------
    def correct(self, phi_n1, phi_n, prior_weight=0.0, log_prob_old_data=0.0, want_inc=False, want_normw=False):
        inc = np.zeros(self.count) if want_inc else None
        nw = np.zeros(self.count) if want_normw else None
        out = np.zeros(3)
        self._ck(lib.smcb200_correct(self.h, phi_n1, phi_n, prior_weight, log_prob_old_data, ptr(inc), ptr(nw), ptr(out)))
        return out, inc, nw

    def ess_at(self, phis, phi_n1):
        phis = _f64(np.atleast_1d(phis))
        out = np.zeros(len(phis))
        self._ck(lib.smcb200_ess_at(self.h, ptr(phis), len(phis), phi_n1, ptr(out)))
        return out

    def solve_adaptive_phi(self, schedule, j, phi_prop, phi_n1, tempering_target, ess_prev, resampled_last_period):
        sched = _f64(schedule)
        jj, pp, pn = C.c_int64(int(j)), C.c_double(phi_prop), C.c_double()
        self._ck(lib.smcb200_solve_adaptive_phi(self.h, ptr(sched), len(sched), C.byref(jj), C.byref(pp), phi_n1,
                                                tempering_target, ess_prev, int(bool(resampled_last_period)), C.byref(pn)))
        return pn.value, False, jj.value, pp.value

    def resample(self, method="systematic", seed=0, stage=0, u=-1.0, want_indices=False):
        if method not in RESAMPLERS:
            raise ValueError("Invalid resampler in SMC. Options are :systematic, :multinomial, or :polyalgo")
        idx = np.zeros(self.count, np.int64) if want_indices else None
        self._ck(lib.smcb200_resample(self.h, RESAMPLERS[method], seed, stage, u, ptr(idx)))
        return idx

    def resample_weights(self, weights, method="systematic", seed=0, stage=0, u=-1.0, want_cum=False, n_parts=None):
        """`resample(weights; n_parts = length(weights), method)` (src/resample.jl:23): 1-based ancestor indices."""
        if method not in RESAMPLERS:
            raise ValueError("Invalid resampler in SMC. Options are :systematic, :multinomial, or :polyalgo")
        w = _f64(weights)
        n_out = len(w) if n_parts is None else int(n_parts)
        idx = np.zeros(n_out, np.int64)
        cum = np.zeros(len(w)) if want_cum else None
        self._ck(lib.smcb200_resample_weights_n(self.h, ptr(w), len(w), n_out, RESAMPLERS[method], seed, stage, u, ptr(idx), ptr(cum)))
        return (idx, cum) if want_cum else idx

    def moments(self):
        mean = np.zeros(self.n_para)
        cov = np.zeros((self.n_para, self.n_para))
        self._ck(lib.smcb200_moments(self.h, ptr(mean), ptr(cov)))
        return mean, cov

    def moments_onepass(self):
        """The stage's own one-pass moments (shift = particle 0); equals moments() up to rounding."""
        mean = np.zeros(self.n_para)
        cov = np.zeros((self.n_para, self.n_para))
        self._ck(lib.smcb200_moments_onepass(self.h, ptr(mean), ptr(cov)))
        return mean, cov

    def mutate(self, mean_fr, cov_fr, blocks_free, blocks_all, phi_n, phi_n1, c=1.0, alpha=1.0, n_mh_steps=1,
               has_old_data=False, seed=0, stage=0):
        """blocks_*: list of 0-based index lists (generate_free_blocks / generate_all_blocks)."""
        sizes = _i32([len(b) for b in blocks_all])
        bf = _i32(np.concatenate([np.asarray(b) for b in blocks_free]))
        ba = _i32(np.concatenate([np.asarray(b) for b in blocks_all]))
        mean_fr, cov_fr = _f64(mean_fr), _f64(cov_fr)
        acc = C.c_double()
        self._ck(lib.smcb200_mutate(self.h, ptr(mean_fr), ptr(cov_fr), len(mean_fr), len(sizes), ptr(sizes), ptr(bf), ptr(ba),
                                    phi_n, phi_n1, c, alpha, n_mh_steps, int(bool(has_old_data)), seed, stage, C.byref(acc)))
        return acc.value

    def stage(self, cfg: StageConfig, state: StageState, schedule=None, want_inc=False, want_normw=False, inc_out=None,
              normw_out=None):
        """One fused stage (src/smc_main.jl:377-497).  want_inc / want_normw return the columns the reference appends to
        w_matrix / W_matrix; inc_out / normw_out let the caller supply (e.g. pinned) float64 buffers of shard length."""
        inc = inc_out if inc_out is not None else (np.zeros(self.count) if want_inc else None)
        nw = normw_out if normw_out is not None else (np.zeros(self.count) if want_normw else None)
        for b in (inc, nw):
            assert b is None or (b.dtype == np.float64 and b.shape == (self.count,) and b.flags.c_contiguous)
        sched = _f64(schedule) if schedule is not None else None
        res = StageResult()
        self._ck(lib.smcb200_stage(self.h, C.byref(cfg), C.byref(state), ptr(sched), 0 if sched is None else len(sched),
                                   ptr(inc), ptr(nw), C.byref(res)))
        return res, inc, nw

    def run_stages(self, cfg: StageConfig, state: StageState, schedule, i_first, n_stages, inc_hist=None, normw_hist=None):
        """Up to n_stages consecutive stages of the recursion in ONE call (smcb200_run_stages): stage k is the reference's loop
        index i = i_first + k.  inc_hist / normw_hist: optional float64 arrays of shape (n_stages, shard length), C order
        (row k = the w_matrix / W_matrix column of stage k); pinned memory keeps their streaming asynchronous.
        Returns the list of StageResult of the completed stages."""
        sched = _f64(schedule)
        for b in (inc_hist, normw_hist):
            assert b is None or (b.dtype == np.float64 and b.shape == (n_stages, self.count) and b.flags.c_contiguous)
        results = (StageResult * int(n_stages))()
        n_done = C.c_int32(0)
        st = lib.smcb200_run_stages(self.h, C.byref(cfg), C.byref(state), ptr(sched), len(sched), int(i_first), int(n_stages),
                                    ptr(inc_hist), ptr(normw_hist), self.count, results, C.byref(n_done))
        self.last_results = [results[k] for k in range(n_done.value)]
        self._ck(st)
        return self.last_results

    def stage_host(self, particles, cfg: StageConfig, state: StageState, schedule=None):
        """One stage on a host-resident cloud (Fortran-ordered, updated in place)."""
        assert particles.flags.f_contiguous and particles.dtype == np.float64
        sched = _f64(schedule) if schedule is not None else None
        res = StageResult()
        self._ck(lib.smcb200_stage_host(self.h, ptr(particles), particles.shape[0], C.byref(cfg), C.byref(state), ptr(sched),
                                        0 if sched is None else len(sched), C.byref(res)))
        return res

    # ---- introspection ------------------------------------------------------------------------------
    @property
    def kernel_launches(self):
        return lib.smcb200_kernel_launches(self.h)

    def last_kernel_ms(self, which):
        ms = C.c_float()
        self._ck(lib.smcb200_last_kernel_ms(self.h, which, C.byref(ms)))
        return ms.value

    def timer_start(self):
        self._ck(lib.smcb200_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float()
        self._ck(lib.smcb200_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def fp64_peak(self, iters=2048):
        """Measured FP64 FMA throughput of this GPU in TFLOP/s (denominator of the FP64 roofline)."""
        out = C.c_double()
        self._ck(lib.smcb200_fp64_peak(self.h, int(iters), C.byref(out)))
        return out.value

    def debug_math(self, op, x, seed=0):
        x = _f64(x)
        out = np.zeros_like(x)
        self._ck(lib.smcb200_debug_math(self.h, op, ptr(x), x.size, seed, ptr(out)))
        return out
