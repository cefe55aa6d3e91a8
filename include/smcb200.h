/*
 * smcb200.h -- C ABI of the B200-native SMC particle engine (libsmcb200.so).
 *
 * The reference (FRBNY-DSGE/SMC.jl v0.1.15) is pure Julia and has NO FFI of its own; the drop-in
 * boundary is its Julia API: `smc(loglikelihood, parameters, data; ...)` (src/smc_main.jl:118-161)
 * and `mutable struct Cloud` (src/particle.jl:31-41).  This header is the C ABI a thin Julia shim
 *  `ccall`s from that API's stage loop; each entry point cites the reference code it
 * replaces.  Conventions:
 *   - every function returns an int32 status (SMCB200_OK == 0); no exceptions/callbacks cross the ABI;
 *   - all calls are blocking (the context's stream is synchronised before returning);
 *   - the caller owns host buffers, the library owns device buffers, no host pointer is retained;
 *   - one context per host thread / per GPU (not re-entrant); multi-GPU = one process per GPU, the
 *     ranks of one job joined by smcb200_comm_init();
 *   - host matrices use the reference's layout: `Cloud.particles` is n_parts x (n_para+5) column-major
 *     Float64 (src/particle.jl:31-63: columns 1..n_para parameters, then loglh, logprior, old_loglh,
 *     accept, weight) -- byte-identical to the device struct-of-arrays, so upload/download are memcpys;
 *   - indices crossing the ABI are 1-based Int64 where the reference's are (ancestor indices), 0-based
 *     int32 for block index lists (documented per call).
 */
#ifndef SMCB200_H
#define SMCB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library itself is built with -fvisibility=hidden */
#endif

#define SMCB200_ABI_VERSION 2

typedef struct smcb200_ctx smcb200_ctx;

/* status codes -> the Julia exceptions of SURVEY 8(b) in the shim */
enum {
    SMCB200_OK = 0,
    SMCB200_ERR_NAN_ESS = 1,        /* check_nan_ess assertion, src/helpers.jl:270-305            */
    SMCB200_ERR_BAD_RESAMPLER = 2,  /* throw("Invalid resampler ..."), src/resample.jl:77          */
    SMCB200_ERR_BAD_ARGUMENT = 3,   /* DomainError / size errors, src/smc_main.jl:331, particle.jl */
    SMCB200_ERR_NOT_POSDEF = 4,     /* PosDefException out of MvNormal, src/mutation.jl:81          */
    SMCB200_ERR_CUDA = 5,
    SMCB200_ERR_NCCL = 6,
    SMCB200_ERR_UNSUPPORTED = 7,    /* likelihood / prior / dimension without a device kernel       */
    SMCB200_ERR_NOT_READY = 8       /* call order: cloud / parameters / likelihood not set          */
};

/* prior families (ModelConstructors priors seen in the reference's examples and fixtures) */
enum {
    SMCB200_PRIOR_NORMAL = 0,         /* Normal(mu = p1, sigma = p2)                */
    SMCB200_PRIOR_UNIFORM = 1,        /* Uniform(a = p1, b = p2)                    */
    SMCB200_PRIOR_GAMMA = 2,          /* Gamma(shape = p1, scale = p2)              */
    SMCB200_PRIOR_ROOT_INV_GAMMA = 3, /* RootInverseGamma(nu = p1, tau = p2)        */
    SMCB200_PRIOR_BETA = 4,           /* Beta(a = p1, b = p2)                       */
    SMCB200_PRIOR_INV_GAMMA = 5       /* InverseGamma(shape = p1, scale = p2)       */
};

/* likelihood families with a device functor */
enum {
    SMCB200_LIK_NONE = 0,
    SMCB200_LIK_GAUSSREG = 1, /* Gaussian regression family in centred sufficient-statistic form:
                                iparams = {n_eq, k, stride, coef_off, sig_off(-1 = sigma known)},
                                dparams = per equation {T, qscale, rss, sigma_fixed, bhat[k], U[k*k] (upper, row-major)}.
                                Covers examples/regression_model (:46-53), test/modelsetup.jl:119-138 and
                                examples/capm_model (:48-69). */
    SMCB200_LIK_AS_DSGE = 2  /* three-equation An-Schorfheide DSGE model (BASELINE config C4): the user likelihood
                                `DSGE.likelihood(m, data; sampler=false, catch_errors=true)` of
                                examples/dsge_models/small_dsge_model.jl:35-50, solved (closed-form decision rule) and
                                Kalman-filtered on the device.  n_para = 16 in DSGE.jl's AnSchorfheide order;
                                iparams = {n_periods, n_presample (filtered, not scored)},
                                dparams = data, 3 x n_periods column-major (gdp growth, inflation, nominal rate). */
};

/* :polyalgo (src/resample.jl:73-75, StatsBase.sample: i.i.d. categorical draws through an alias table) has the
 * multinomial resampler's distribution; the host shims map it to SMCB200_RESAMPLE_MULTINOMIAL. */
enum { SMCB200_RESAMPLE_SYSTEMATIC = 0, SMCB200_RESAMPLE_MULTINOMIAL = 1 };

/* ---- library / context ---------------------------------------------------------------------- */
int32_t smcb200_abi_version(void);
const char *smcb200_status_string(int32_t status);
/* Creates a context on CUDA device `device`.  Fails (SMCB200_ERR_CUDA) when there is no usable GPU:
 * there is deliberately no CPU fallback. */
int32_t smcb200_create(smcb200_ctx **ctx_out, int32_t device);
int32_t smcb200_destroy(smcb200_ctx *ctx);
const char *smcb200_last_error(const smcb200_ctx *ctx);

/* ---- multi-GPU: replaces Distributed.jl's worker fan-out (src/smc_main.jl:169-170,471-476) --- */
/* rank 0 creates a 128-byte id, the host distributes it (any transport), every rank calls comm_init.
 * Particles are then sharded in contiguous global-index ranges; see smcb200_cloud_shard(). */
int32_t smcb200_comm_unique_id(void *id128_out);
int32_t smcb200_comm_init(smcb200_ctx *ctx, int32_t rank, int32_t world, const void *id128);

/* ---- Cloud (src/particle.jl:31-63) ----------------------------------------------------------- */
/* Cloud(n_params, n_parts) (particle.jl:50-53).  n_parts is the GLOBAL particle count. */
int32_t smcb200_cloud_create(smcb200_ctx *ctx, int64_t n_parts, int32_t n_para);
/* this rank's contiguous shard [first, first+count) of the global particle index (0-based) */
int32_t smcb200_cloud_shard(const smcb200_ctx *ctx, int64_t *first, int64_t *count);
/* host <-> device copies of this rank's shard; `ld` = leading dimension (rows) of the host matrix,
 * `row0` = first host row to use (so a rank can pass the full matrix with row0 = first). */
int32_t smcb200_cloud_upload(smcb200_ctx *ctx, const double *particles, int64_t ld, int64_t row0);
int32_t smcb200_cloud_download(smcb200_ctx *ctx, double *particles, int64_t ld, int64_t row0);
/* single column accessors: get_loglh/get_weights/... and update_* (particle.jl:71-259) */
int32_t smcb200_cloud_read_column(smcb200_ctx *ctx, int32_t col, double *out);
int32_t smcb200_cloud_write_column(smcb200_ctx *ctx, int32_t col, const double *in);

/* ---- model ---------------------------------------------------------------------------------- */
/* ParameterVector: fixed flags, valuebounds, prior family + two parameters (ModelConstructors
 * `update!` bounds check and `prior`, called at src/mutation.jl:93-95). */
int32_t smcb200_set_parameters(smcb200_ctx *ctx, int32_t n_para, const int32_t *fixed, const double *lo,
                               const double *hi, const int32_t *prior_kind, const double *p1, const double *p2);
/* slot 0: loglikelihood on data; slot 1: old_loglikelihood on old_data (src/mutation.jl:96,106). */
int32_t smcb200_set_likelihood(smcb200_ctx *ctx, int32_t slot, int32_t kind, const int32_t *iparams,
                               int32_t n_iparams, const double *dparams, int64_t n_dparams);
/* stage-0 evaluators.  mode 0: loglh/logprior of the stored draws (draw_likelihood,
 * src/initialization.jl:129-139); mode 1: initialize_likelihoods! (:153-186: old_loglh <- loglh, then
 * re-evaluate loglh and logprior on the new data). */
int32_t smcb200_evaluate(smcb200_ctx *ctx, int32_t mode);
/* initial_draw! (src/initialization.jl:88-119): draws every free parameter from its prior (redrawn
 * until inside valuebounds and until the log-likelihood is finite, at most `max_tries` times),
 * fixed parameters get `fixed_values`; sets loglh, logprior, old_loglh = 0, weight = 1. */
int32_t smcb200_initial_draw(smcb200_ctx *ctx, const double *fixed_values, uint64_t seed, int32_t max_tries);

/* ---- stage operations (src/smc_main.jl:377-497) ---------------------------------------------- */
/* Correction, :400-427 + particle.jl:250-259,362-369.  inc_out / normw_out (nullable, shard length)
 * receive the columns the reference appends to w_matrix / W_matrix.  out[0] = sum of the unnormalised
 * weights, out[1] = ESS, out[2] = sum of the normalised weights. */
int32_t smcb200_correct(smcb200_ctx *ctx, double phi_n1, double phi_n, double prior_weight,
                        double log_prob_old_data, double *inc_out, double *normw_out, double out[3]);
/* compute_ESS at K trial phi (src/helpers.jl:173-181); weights are not modified. */
int32_t smcb200_ess_at(smcb200_ctx *ctx, const double *phi, int32_t K, double phi_n1, double *ess_out);
/* solve_adaptive_phi (src/helpers.jl:9-56): j (1-based) and phi_prop are in/out as in the reference. */
int32_t smcb200_solve_adaptive_phi(smcb200_ctx *ctx, const double *schedule, int32_t n_phi, int64_t *j_io,
                                   double *phi_prop_io, double phi_n1, double tempering_target, double ess_prev,
                                   int32_t resampled_last_period, double *phi_n_out);
/* Selection on the device cloud, :435-446: ancestor indices from the current weights (resample(),
 * src/resample.jl:23-71), gather of all columns, reset_weights!.  idx_out (nullable, GLOBAL length,
 * 1-based) receives the ancestors.  u_override >= 0 replaces the Philox systematic offset. */
int32_t smcb200_resample(smcb200_ctx *ctx, int32_t method, uint64_t seed, uint32_t stage, double u_override,
                         int64_t *idx_out);
/* The exported `resample(weights; method)` on a host weight vector (src/resample.jl:23): returns
 * 1-based indices; cum_out (nullable) receives cumsum(weights ./ sum(weights)). Single GPU. */
int32_t smcb200_resample_weights(smcb200_ctx *ctx, const double *weights, int64_t n, int32_t method, uint64_t seed,
                                 uint32_t stage, double u_override, int64_t *idx_out, double *cum_out);
/* `resample(weights; n_parts = n_out, method)`: n_out ancestors out of n weights, thresholds (i - 1 + u) / n_out
 * (bridge initialisation, src/smc_main.jl:262-268).  All n weights are searched (the reference's systematic
 * search stops at cumulative[n_parts], src/resample.jl:54, which makes n_parts < n unusable there). */
int32_t smcb200_resample_weights_n(smcb200_ctx *ctx, const double *weights, int64_t n, int64_t n_out, int32_t method,
                                   uint64_t seed, uint32_t stage, double u_override, int64_t *idx_out, double *cum_out);
/* weighted_mean / weighted_cov (src/particle.jl:481-486,526-532): mean[n_para], cov[n_para^2].  Two passes over the
 * cloud like StatsBase.cov (mean, then centred scatter). */
int32_t smcb200_moments(smcb200_ctx *ctx, double *mean, double *cov);
/* The same moments as the fused stage computes them for the proposal (src/smc_main.jl:457-465): ONE pass over the cloud
 * with the parameter vector of particle 0 as shift; equal to smcb200_moments up to rounding (~1e-15 relative). */
int32_t smcb200_moments_onepass(smcb200_ctx *ctx, double *mean, double *cov);
/* mutation of every particle (src/mutation.jl:56-138 through the fan-out at smc_main.jl:471-484).
 * mean_fr / cov_fr: theta_bar and R restricted to the free parameters (:462-465).  Blocks as from
 * generate_free_blocks / generate_all_blocks (src/helpers.jl:215-260) but 0-based, concatenated, with
 * block_sizes[n_blocks].  Writes theta, loglh, logprior, old_loglh, accept; returns the mean of the
 * accept column (update_acceptance_rate!, particle.jl:466-468). */
int32_t smcb200_mutate(smcb200_ctx *ctx, const double *mean_fr, const double *cov_fr, int32_t n_free,
                       int32_t n_blocks, const int32_t *block_sizes, const int32_t *blocks_free,
                       const int32_t *blocks_all, double phi_n, double phi_n1, double c, double alpha,
                       int32_t n_mh_steps, int32_t has_old_data, uint64_t seed, uint32_t stage,
                       double *mean_accept_out);

/* ---- one fused stage: the body of `while phi_n < 1` (src/smc_main.jl:377-497) ------------------ */
typedef struct {
    double phi_n1;            /* cloud.tempering_schedule[i-1]                                  */
    double phi_n;             /* fixed schedule value; ignored when adaptive != 0                */
    double threshold_ratio;   /* resample iff ESS < threshold_ratio * n_parts (:435)             */
    double target;            /* target accept rate (:453)                                       */
    double alpha;             /* mixture proportion                                              */
    double tempering_target;  /* adaptive only                                                   */
    double prior_weight;      /* tempered_update_prior_weight (:401-410)                         */
    double log_prob_old_data;
    int32_t n_mh_steps, n_blocks, resample_method, adaptive, has_old_data, reserved;
    uint64_t seed;
    uint32_t stage;           /* cloud.stage_index after the increment at :379 (2, 3, ...)       */
    uint32_t reserved2;
} smcb200_stage_config;

typedef struct {              /* carried from stage to stage by the caller                       */
    double c;                 /* cloud.c                                                         */
    double accept;            /* cloud.accept (previous stage's mean accept; initial = target)   */
    double ess_prev;          /* cloud.ESS[i-1]                                                  */
    double phi_prop;          /* adaptive only                                                   */
    int64_t j;                /* adaptive only, 1-based cursor into the proposed fixed schedule  */
    int32_t resampled_last_period;
    int32_t reserved;
} smcb200_stage_state;

typedef struct {
    double phi_n, ess, sum_weights, c, accept;
    int32_t resampled, status;
    float ms_correct, ms_resample, ms_moments, ms_mutate; /* device time of the stage's phases (CUDA events) */
} smcb200_stage_result;

/* schedule/n_phi: the proposed fixed schedule (adaptive) -- may be NULL for a fixed schedule.
 * inc_out / normw_out as in smcb200_correct (normw_out is reset to 1 on resample, :445). */
int32_t smcb200_stage(smcb200_ctx *ctx, const smcb200_stage_config *cfg, smcb200_stage_state *state,
                      const double *schedule, int32_t n_phi, double *inc_out, double *normw_out,
                      smcb200_stage_result *result);
/* The recursion `while phi_n < 1` (src/smc_main.jl:377-508) for up to n_stages consecutive stages in ONE call, without
 * the host in the loop.  cfg: as for smcb200_stage; its phi_n / stage fields are ignored, phi_n1 is read for an adaptive
 * run (cloud.tempering_schedule[i_first - 1]).  i_first: the reference's loop index `i` of the first stage to run
 * (>= 2; stage k uses phi_n1 = schedule[i-2], phi_n = schedule[i-1] on a fixed schedule and seeds its random streams
 * with stage = i, exactly as n_stages calls of smcb200_stage would).  Fixed schedule: all stages are enqueued back to
 * back -- step size, mean accept rate and the resample decision stay on the device -- and the host synchronises once;
 * adaptive schedule: one synchronisation per stage, stops after the stage that reaches phi_n = 1.
 * inc_hist / normw_hist (nullable): host matrices, column k (leading dimension ld_hist >= shard length) receives the
 * w_matrix / W_matrix column of stage k (src/smc_main.jl:419-420,445); they stream out on a copy stream behind the
 * computation (pinned memory keeps that asynchronous).  results[n_stages]; *n_done = stages completed.  The same status
 * codes as smcb200_stage; after an error results[0 .. *n_done) are valid. */
int32_t smcb200_run_stages(smcb200_ctx *ctx, const smcb200_stage_config *cfg, smcb200_stage_state *state,
                           const double *schedule, int32_t n_phi, int32_t i_first, int32_t n_stages, double *inc_hist,
                           double *normw_hist, int64_t ld_hist, smcb200_stage_result *results, int32_t *n_done);
/* Same with the Cloud living in HOST memory: upload -> stage -> download (what a caller holding a
 * Julia `Cloud` pays per call). */
int32_t smcb200_stage_host(smcb200_ctx *ctx, double *particles, int64_t ld, const smcb200_stage_config *cfg,
                           smcb200_stage_state *state, const double *schedule, int32_t n_phi,
                           smcb200_stage_result *result);

/* ---- introspection for tests / benchmarks ---------------------------------------------------- */
/* number of kernels this context has launched since creation */
int64_t smcb200_kernel_launches(const smcb200_ctx *ctx);
/* last device timings (ms) of the named kernel family measured with CUDA events on the launch stream:
 * which = 0 correct, 1 resample, 2 moments, 3 mutate */
int32_t smcb200_last_kernel_ms(const smcb200_ctx *ctx, int32_t which, float *ms_out);
/* CUDA-event stopwatch on the context's launch stream: start records an event, stop records a second
 * one, synchronises and returns the elapsed device time in milliseconds */
int32_t smcb200_timer_start(smcb200_ctx *ctx);
int32_t smcb200_timer_stop(smcb200_ctx *ctx, float *ms_out);
/* measured FP64 FMA throughput of this GPU (TFLOP/s, best of 5 launches of `iters` x 64 dependent-chain DFMA per thread
 * on every SM): the denominator of bench.py's FP64 roofline */
int32_t smcb200_fp64_peak(smcb200_ctx *ctx, int32_t iters, double *tflops_out);
/* device-side deterministic elementary functions, for parity tests: op 0 exp, 1 log, 2 sin(2 pi x),
 * 3 cos(2 pi x), 4/5 = z0/z1 of normal_pair(seed, particle = i, stage = 0, slot = x[i]) (binary64 Box-Muller,
 * prior draws), 6..9 = the four proposal normals of normal_quad (table-driven inverse CDF in binary32) of the same
 * Philox block */
int32_t smcb200_debug_math(smcb200_ctx *ctx, int32_t op, const double *x, int64_t n, uint64_t seed, double *out);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* SMCB200_H */
