// mutate_kernel.cuh -- K7: random-walk Metropolis-Hastings mutation of every particle, plus the stage-0
// likelihood/prior evaluators that share its device functors.
//
// Replaces src/mutation.jl:56-138 fanned out at src/smc_main.jl:471-484 (one Distributed.jl task
// per particle, 1 Cholesky + 4 SVD per MH step) with ONE kernel per stage:
//   * one thread per particle: the cloud is a struct-of-arrays [n_para+5][N], so a warp's loads and
//     stores of any column are 256 contiguous bytes; the particle lives in registers for all
//     n_mh_steps * n_blocks steps, so HBM traffic is 8(2d+7) B/particle/stage regardless of n_mh_steps;
//   * the scaled Cholesky factor c*L (factored ONCE per stage), the likelihood's sufficient statistics
//     and the prior table live in __constant__ memory and enter the DFMAs as constant-bank operands
//     (fully unrolled loops => immediate offsets): no shared-memory or register traffic for them;
//   * counter-based Philox4x32-10 keyed on the GLOBAL particle index => results are invariant to the
//     launch shape and to the number of GPUs.
// The kernel is FP64-pipe bound (see DESIGN.md roofline), not HBM bound, for d >~ 8.
//
// This header is included by several translation units (mut_*.cu), each instantiating the kernels of a
// few likelihood functors so that they compile in parallel.  Without relocatable device code every
// translation unit owns its copies of the __constant__ blocks below; KernelEntry carries the upload
// functions of the unit that owns the selected kernel.
#pragma once
#include "common.cuh"
#include "reduce.cuh"
#include "aslik.cuh"

#ifndef SMC_NORM_BATCH
#define SMC_NORM_BATCH 1
#endif
#ifndef SMC_F2D_INT
#define SMC_F2D_INT 0
#endif

namespace smc {
namespace {

__constant__ MutConst c_mut;
__constant__ LikSlot c_lik[2];
__constant__ PriorConst c_pri;

// binary32 -> binary64 of a proposal normal (zero or normal, never subnormal / inf / nan): exact either way; the integer
// form keeps the conversion off the XU pipe
__device__ __forceinline__ double f2d(float z)
{
#if SMC_F2D_INT
    const uint32_t zb = __float_as_uint(z), mag = zb & 0x7fffffffu;
    const uint32_t hi = (mag ? (mag >> 3) + 0x38000000u : 0u) | (zb & 0x80000000u);
    return __hiloint2double((int)hi, (int)(zb << 29));
#else
    return (double)z;
#endif
}

// ---- priors (ModelConstructors.prior: sum over free parameters) ------------------
// Families other than Normal sit behind a call so that the unrolled per-parameter code stays small
// (the instruction cache matters: this kernel is latency bound).
__device__ __noinline__ double logpdf_general(int k, double x)
{
    switch (c_pri.kind[k]) {
    case SMCB200_PRIOR_UNIFORM:
        return (x >= c_pri.p1[k] && x <= c_pri.p2[k]) ? c_pri.cst[k] : -dinf();
    case SMCB200_PRIOR_GAMMA:
        if (!(x > 0.0)) return (x == 0.0 && c_pri.a1[k] == 0.0) ? c_pri.cst[k] : -dinf();
        return fma(c_pri.a1[k], det_log(x), c_pri.cst[k]) - x * c_pri.a2[k];
    case SMCB200_PRIOR_ROOT_INV_GAMMA: {
        if (!(x > 0.0)) return -dinf();
        const double x2 = x * x;
        return fma(-c_pri.a1[k], det_log(x2), c_pri.cst[k]) - c_pri.a2[k] / x2;
    }
    case SMCB200_PRIOR_BETA:
        if (!(x > 0.0 && x < 1.0)) return -dinf();
        return fma(c_pri.a2[k], det_log(1.0 - x), fma(c_pri.a1[k], det_log(x), c_pri.cst[k]));
    case SMCB200_PRIOR_INV_GAMMA:
        if (!(x > 0.0)) return -dinf();
        return fma(-c_pri.a1[k], det_log(x), c_pri.cst[k]) - c_pri.a2[k] / x;
    }
    return dnan();
}
__device__ __forceinline__ double logpdf1(int k, double x)
{
    if (c_pri.kind[k] == SMCB200_PRIOR_NORMAL) {
        const double z = (x - c_pri.p1[k]) * c_pri.a1[k];
        return fma(-0.5 * z, z, c_pri.cst[k]);
    }
    return logpdf_general(k, x);
}

template <int D>
__device__ __forceinline__ double logprior(const double (&th)[D])
{
    if (c_pri.all_normal) {      // block-uniform: 3 FP64 instructions per parameter instead of ~14 mixed ones
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const double z = (th[k] - c_pri.p1[k]) * c_pri.a1[k];
            acc = fma(z, z, acc);
        }
        return fma(-0.5, acc, c_pri.cst_sum);
    }
    double lp = 0.0;
#pragma unroll
    for (int k = 0; k < D; ++k)
        if (!c_pri.fixed[k]) lp = lp + logpdf1(k, th[k]);
    return lp;
}
template <int D>
__device__ __forceinline__ bool in_bounds(const double (&th)[D])
{
    bool ok = true;
    if (c_pri.all_normal) {      // no fixed parameters: no per-parameter flag loads
#pragma unroll
        for (int k = 0; k < D; ++k) ok = ok && (th[k] >= c_pri.lo[k] && th[k] <= c_pri.hi[k]);
        return ok;
    }
#pragma unroll
    for (int k = 0; k < D; ++k)
        if (!c_pri.fixed[k]) ok = ok && (th[k] >= c_pri.lo[k] && th[k] <= c_pri.hi[k]);
    return ok;
}


// ---- Gaussian regression family (centred sufficient statistics) ---------------------------------
template <int NEQ_, int K_, int STRIDE_, int COEF_, int SIG_>
struct GaussReg {
    static constexpr int KIND = SMCB200_LIK_GAUSSREG;
    static constexpr int NEQ = NEQ_, K = K_, STRIDE = STRIDE_, COEF = COEF_, SIG = SIG_;
    static constexpr int KP = K * (K + 1) / 2;
    static constexpr int D_COEF = COEF + (NEQ - 1) * STRIDE + K;
    static constexpr int D_SIG = (SIG >= 0) ? SIG + (NEQ - 1) * STRIDE + 1 : 0;
    static constexpr int D = (D_COEF > D_SIG) ? D_COEF : D_SIG;
    static constexpr int MINB = (D <= 20) ? MUT_MINB20 : ((D <= 24) ? MUT_MINB20 * 4 / 5 : MUT_MINB20 * 3 / 5);

    template <int SLOT>
    static __device__ __forceinline__ double ll(const double (&th)[D])
    {
        const LikSlot& L = c_lik[SLOT];
        double ll = 0.0;
        bool bad = false;
#pragma unroll
        for (int e = 0; e < NEQ; ++e) {
            double dl[K];
#pragma unroll
            for (int j = 0; j < K; ++j) dl[j] = th[COEF + e * STRIDE + j] - L.bhat[e * K + j];
            double q = L.rss[e];
#pragma unroll
            for (int i = 0; i < K; ++i) {
                double r = 0.0;
#pragma unroll
                for (int j = i; j < K; ++j) r = fma(L.U[e * KP + i * K - (i * (i - 1)) / 2 + (j - i)], dl[j], r);
                q = fma(r, r, q);
            }
            double logs, inv_s2;
            if (SIG >= 0) {
                const double s = th[(SIG >= 0 ? SIG : 0) + e * STRIDE];
                bad = bad || !(s > 0.0);
                logs = det_log(s);
                inv_s2 = 1.0 / (s * s);
            } else {
                logs = L.logs[e];
                inv_s2 = L.inv_s2[e];
            }
            const double le = fma(-L.T[e], logs, L.cT[e]) - (0.5 * L.qscale[e] * q) * inv_s2;
            ll = ll + le;
        }
        return bad ? -dinf() : ll;
    }
};

// Quadratic form v' (c^2 Sigma_b)^{-1} v through the scaled factor embedded in parameter order: forward
// substitution row by row (each y_i one fma chain over ascending j, times the reciprocal diagonal), v is overwritten by
// y.  Entries of v outside the block must be 0 (their factor entries are 0, so the chain passes through).
// position of L[r][j] (r >= j) in the column-packed lower factor: column j is contiguous in r, so the unrolled
// mat-vec fetches two entries per LDCU.128
template <int D>
__device__ __forceinline__ constexpr int lcol(int r, int j) { return j * D - (j * (j - 1)) / 2 + (r - j); }

template <int D>
__device__ __forceinline__ double mvn_quad(int b, uint32_t mask, double (&v)[D])
{
    double q = 0.0;
#pragma unroll
    for (int i = 0; i < D; ++i) {
        if ((mask >> i) & 1u) {
            double s = v[i];
#pragma unroll
            for (int j = 0; j < i; ++j) s = fma(-c_mut.L[b][lcol<D>(i, j)], v[j], s);
            v[i] = s * c_mut.rl[b][i];
            q = fma(v[i], v[i], q);
        }
    }
    return q;
}


// One MH chain per thread.  The chain's current state lives in shared memory ([2][D][128] doubles:
// current and candidate buffer, thread-private columns => conflict-free); registers hold only the working
// vector, which is in turn the proposal increment, the candidate and (inside the likelihood) the centred
// candidate.  Accepting a move flips which buffer is current.  This keeps the kernel under ~100 registers
// (5 blocks = 20 warps per SM) -- it is FP64-latency bound, so resident warps are what buys throughput.
// BLK = 0: several blocks; 1: n_blocks == 1 (block index is a literal, so factor entries load as LDCU.128 pairs);
// 2: one block holding ALL D parameters (none fixed): the membership mask is a compile-time constant and every
// per-parameter membership branch / select disappears (config C2).

// MIX = (alpha < 1): the three-component mixture proposal of mvnormal_mixture_draw (helpers.jl:87-100) and
// the proposal densities of compute_proposal_densities (helpers.jl:128-164).

// scalars of one chain (loglh, logprior, old_loglh of the current state) + the outcome of the last step
struct ChainScal { double like, lpri, lprev; int accepted; };

// ONE Metropolis-Hastings step of one block of parameters (mutation.jl:75-134).  SMC_STEP_CALL = 1 compiles it as a real
// call: as straight-line code outside the kernel's loops ptxas rotates the uniform registers that carry the factor
// entries into the DFMAs (LDCU.128 several columns ahead) instead of funnelling a whole mat-vec through one uniform
// register quad.  Measured on B200 (tools/operand_probe.cu, profiles/r02_operand_probe.txt): the mat-vec alone gains
// (25.6 -> 29.1 TFLOP/s at 20 warps/SM), the kernel does not (0.290 -> 0.309 ms: the call's spills and the lost
// overlap across steps cost more), so the step stays inlined.
#ifndef SMC_STEP_CALL
#define SMC_STEP_CALL 0
#endif
#if SMC_STEP_CALL
#define SMC_STEP_ATTR __noinline__
#else
#define SMC_STEP_ATTR __forceinline__
#endif
template <class LIK, bool HAS_OLD, int BLK, bool MIX>
__device__ SMC_STEP_ATTR ChainScal mh_step(const double* cur, double* cand, const float4* NTAB, const MutArgs& a, uint32_t gp,
                                           uint32_t sb, int b, double phi, double omphi, ChainScal ch)
{
    const uint32_t stage = a.stage;
    const double alpha = a.alpha;
    constexpr int D = LIK::D;
    constexpr uint32_t FULL_MASK = (D >= 32) ? 0xffffffffu : ((1u << D) - 1u);
    const uint32_t mask = (BLK == 2) ? FULL_MASK : c_mut.mask[b];
    // one Philox block per (step, block): MH uniform (mutation.jl:66,133) and the mixture component
    const u32x4 r4 = rng4_keyed(a.rk, gp, stage, sb, PURP_STEP);
    const double step_prob = u01(r4.x, r4.y);
    int comp = 1;
    if (MIX) {
        const double u_mix = u01(r4.z, r4.w);
        comp = (u_mix < alpha) ? 1 : ((u_mix < alpha + (1.0 - alpha) / 2.0) ? 2 : 3);
    }
    // (1) + (2): one Philox block gives the four normals of parameters 4q .. 4q+3 (table-driven inverse
    // CDF, normal_icdf); each normal is consumed at once by its column of the proposal increment
    // s = (c L) z (every s[r] sums over ascending columns j, the oracle's order), so the normals never
    // leave registers.  The mixture path also parks them in the candidate buffer (component 2 needs z_k).
    constexpr int NQUAD = (D + 3) / 4;
#if SMC_NORM_BATCH
    // phase-batched: every Philox block of the step first (independent chains), then all table rows and
    // polynomials, then the mat-vec -- the long-latency operations (IMAD chains, uint->float conversions,
    // table loads) of different normals overlap instead of sitting in front of each column's DFMAs
    float zf[4 * NQUAD];
    {
        uint32_t rw[4 * NQUAD];
#pragma unroll
        for (int q = 0; q < NQUAD; ++q) {
            if (BLK == 2 || ((mask >> (4 * q)) & 15u)) {
                const u32x4 r = rng4_keyed(a.rk, gp, stage, (sb << 8) | (uint32_t)q, PURP_NORMAL);
                rw[4 * q] = r.x; rw[4 * q + 1] = r.y; rw[4 * q + 2] = r.z; rw[4 * q + 3] = r.w;
            } else {
                rw[4 * q] = 0u; rw[4 * q + 1] = 0u; rw[4 * q + 2] = 0u; rw[4 * q + 3] = 0u;
            }
        }
#pragma unroll
        for (int j = 0; j < 4 * NQUAD; ++j) zf[j] = (j < D) ? normal_icdf_f(rw[j], NTAB) : 0.0f;
    }
    double s[D];
#pragma unroll
    for (int k = 0; k < D; ++k) s[k] = 0.0;
#pragma unroll
    for (int j = 0; j < D; ++j) {
        const double zj = f2d(zf[j]);
        if (MIX) cand[j * MUT_THREADS] = zj;
        if (BLK == 2 || ((mask >> j) & 1u)) {
#pragma unroll
            for (int r = j; r < D; ++r) s[r] = fma(c_mut.L[b][lcol<D>(r, j)], zj, s[r]);
        }
    }
#else
    double s[D];
#pragma unroll
    for (int k = 0; k < D; ++k) s[k] = 0.0;
#pragma unroll
    for (int q = 0; q < NQUAD; ++q) {
        if (BLK == 2 || ((mask >> (4 * q)) & 15u)) {
            double z[4];
            normal_quad(rng4_keyed(a.rk, gp, stage, (sb << 8) | (uint32_t)q, PURP_NORMAL), NTAB, z[0], z[1], z[2], z[3]);
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const int j = 4 * q + jj;
                if (j < D) {
                    if (MIX) cand[j * MUT_THREADS] = z[jj];
                    if (BLK == 2 || ((mask >> j) & 1u)) {
#pragma unroll
                        for (int r = j; r < D; ++r) s[r] = fma(c_mut.L[b][lcol<D>(r, j)], z[jj], s[r]);
                    }
                }
            }
        }
    }
#endif
#pragma unroll
    for (int k = 0; k < D; ++k) {
        const double t = cur[k * MUT_THREADS];
        if (MIX) {
            // component 2: theta_old + c sqrt(Sigma_ii) z_i; component 3: theta_bar + c L z
            const double inc = (comp == 2) ? c_mut.csd[b][k] * cand[k * MUT_THREADS] : s[k];
            const double base = (comp == 3) ? c_mut.mu[k] : t;
            s[k] = ((mask >> k) & 1u) ? base + inc : t;
        } else {
            s[k] = (BLK == 2 || ((mask >> k) & 1u)) ? t + s[k] : t;          // s is now theta'
        }
        cand[k * MUT_THREADS] = s[k];
    }
    const bool ok = in_bounds<D>(s);
    double pn = logprior<D>(s);
    double ln = LIK::template ll<0>(s);
    if (ln == -dinf()) pn = -dinf();                       // mutation.jl:102-104
    double lo = HAS_OLD ? LIK::template ll<1>(s) : 0.0;    // mutation.jl:106
    if (!ok) { pn = -dinf(); ln = -dinf(); lo = -dinf(); } // ParamBoundsError, mutation.jl:112-121
    // alpha == 1: q0 - q1 == +0 exactly (symmetric proposal), see DESIGN.md
    double qdiff = 0.0;
    if (MIX) {
        // compute_proposal_densities (helpers.jl:128-164).  The first terms of q0 and q1 are the same
        // number: N(theta_old; theta', c^2 Sigma) and N(theta'; theta_old, c^2 Sigma) run sign-mirrored
        // fma chains, so one evaluation serves both.
        const double lognorm = c_mut.lognorm[b];
#pragma unroll
        for (int k = 0; k < D; ++k)
            s[k] = ((mask >> k) & 1u) ? cur[k * MUT_THREADS] - cand[k * MUT_THREADS] : 0.0;
        double ind = 1.0;                                   // diagonal component: variance Sigma_ii, no c
#pragma unroll
        for (int k = 0; k < D; ++k)
            if ((mask >> k) & 1u) {
                const double zs = s[k] * c_mut.isd[b][k];
                ind = (ind * c_mut.isdn[b][k]) * det_exp(-0.5 * (zs * zs));
            }
        const double e_sym = det_exp(-0.5 * (lognorm + mvn_quad<D>(b, mask, s)));
#pragma unroll
        for (int k = 0; k < D; ++k) s[k] = ((mask >> k) & 1u) ? cur[k * MUT_THREADS] - c_mut.mu[k] : 0.0;
        const double e_old = det_exp(-0.5 * (lognorm + mvn_quad<D>(b, mask, s)));
#pragma unroll
        for (int k = 0; k < D; ++k) s[k] = ((mask >> k) & 1u) ? cand[k * MUT_THREADS] - c_mut.mu[k] : 0.0;
        const double e_new = det_exp(-0.5 * (lognorm + mvn_quad<D>(b, mask, s)));
        const double w2 = (1.0 - alpha) / 2.0;
        double q0 = alpha * e_sym, q1 = alpha * e_sym;
        q0 = q0 + w2 * ind; q1 = q1 + w2 * ind;
        q0 = q0 + w2 * e_old; q1 = q1 + w2 * e_new;
        q0 = det_log(q0); q1 = det_log(q1);
        if (q0 == dinf() && q1 == dinf()) q0 = 0.0;
        qdiff = q0 - q1;
    }
    const double eta = det_exp(((phi * (ln - ch.like) + omphi * (lo - ch.lprev)) + (pn - ch.lpri)) + qdiff);
    ch.accepted = 0;
    if (step_prob < eta) {                                  // strict <, NaN rejects (mutation.jl:126)
        ch.like = ln; ch.lpri = pn; ch.lprev = lo;
        ch.accepted = 1;
    }
    return ch;
}

template <class LIK, bool HAS_OLD, int BLK, bool MIX>
__global__ void __launch_bounds__(MUT_THREADS, LIK::MINB)
k_mutate(double* cloud, int64_t N, int64_t index0, MutArgs a)
{
    constexpr int D = LIK::D;
    extern __shared__ double sm_state[];
    __shared__ float4 sm_tab[NORMAL_TAB_ROWS];     // inverse-normal-CDF table (random per-lane rows)
    const float4* NTAB = sm_tab;
    __shared__ bool sm_last;
    if (a.scal && a.scal[SC_STATUS] != 0.0) return;   // the stage was poisoned (NaN ESS / non-PD covariance): leave the cloud alone
    if (c_mut.status != 0) return;
    for (int k = threadIdx.x; k < NORMAL_TAB_ROWS; k += MUT_THREADS) sm_tab[k] = reinterpret_cast<const float4*>(normal_tab_dev)[k];
    __syncthreads();
    // Persistent warps: every warp fetches 32-particle work items from a global counter until the cloud is exhausted, so
    // the table above is loaded once per block, there is no block barrier in the loop and the SMs stay evenly loaded.
    const unsigned n_items = (unsigned)((N + 31) / 32);
    const int lane = threadIdx.x & 31;
    unsigned next_item = 0u;
    if (lane == 0) next_item = atomicAdd(a.work_counter, 1u);
    for (;;) {
    const unsigned item = __shfl_sync(0xffffffffu, next_item, 0);
    if (item >= n_items) break;
    if (lane == 0) next_item = atomicAdd(a.work_counter, 1u);      // the next item's ticket travels while this one is processed
    const int64_t i = (int64_t)item * 32 + lane;
    const bool live = i < N;
    int acc_cnt = 0;
    if (live) {
        const double* src = (a.scal && a.scal[SC_RESAMPLE] != 0.0) ? a.alt_in : cloud;
        double* buf0 = sm_state + threadIdx.x;
        double* buf1 = sm_state + D * MUT_THREADS + threadIdx.x;
#pragma unroll
        for (int k = 0; k < D; ++k) buf0[k * MUT_THREADS] = src[col_off(N, k) + i];
        ChainScal ch;
        ch.like = src[col_off(N, D) + i];
        ch.lpri = src[col_off(N, D + 1) + i];
        ch.lprev = src[col_off(N, D + 2) + i];
        ch.accepted = 0;
        bool flipped = false;
        const uint32_t gp = (uint32_t)(index0 + i);
        const double phi = a.scal ? a.scal[SC_PHI_N] : a.phi_n;
        const double omphi = 1.0 - phi;
        constexpr bool SINGLE = BLK != 0;
        const int nb = SINGLE ? 1 : a.n_blocks;

        for (int step = 0; step < a.n_mh_steps; ++step) {
            for (int bb = 0; bb < nb; ++bb) {
                const int b = SINGLE ? 0 : bb;
                const uint32_t sb = (uint32_t)(step * nb + b);
                const double* cur = flipped ? buf1 : buf0;
                double* cand = flipped ? buf0 : buf1;
                ch = mh_step<LIK, HAS_OLD, BLK, MIX>(cur, cand, NTAB, a, gp, sb, b, phi, omphi, ch);
                if (ch.accepted) {
                    flipped = !flipped;
                    acc_cnt += c_mut.bsize[b];
                }
            }
        }
        const double* cur = flipped ? buf1 : buf0;
#pragma unroll
        for (int k = 0; k < D; ++k) cloud[col_off(N, k) + i] = cur[k * MUT_THREADS];
        cloud[col_off(N, D) + i] = ch.like;
        cloud[col_off(N, D + 1) + i] = ch.lpri;
        cloud[col_off(N, D + 2) + i] = ch.lprev;
        cloud[col_off(N, D + 3) + i] = (double)acc_cnt / (double)a.n_free;   // particle.jl:410-418
    }
    // mean of the accept column (update_acceptance_rate!, particle.jl:466-468): the per-particle counts are integers, so
    // their sum is exact and order-free -- one warp reduction and one atomic per work item
    const int wsum = __reduce_add_sync(0xffffffffu, live ? acc_cnt : 0);
    if (lane == 0 && wsum) atomicAdd(a.acc_total, (unsigned long long)wsum);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned t = atomicInc(a.acc_counter, gridDim.x - 1);
        sm_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!sm_last) return;
    // last block: every work item is done.  accept sum = total / n_free (global after the cross-GPU exchange)
    __threadfence();
    __shared__ double sm_x[2];
    if (threadIdx.x == 0) {
        const unsigned long long tot = *reinterpret_cast<volatile unsigned long long*>(a.acc_total);
        sm_x[0] = (double)tot; sm_x[1] = (double)tot;
        *a.acc_total = 0ull; *a.work_counter = 0u;          // ready for the next launch
    }
    __syncthreads();
    if (a.pc.world > 1) peer_exchange_block(a.pc, sm_x, 1, 1, sm_x + 1);     // integer-valued doubles: exact in any order
    __syncthreads();
    if (threadIdx.x == 0) {
        const double sum = sm_x[1] / (double)a.n_free;
        *a.acc_out = sum;
        if (a.acc_mean_out) *a.acc_mean_out = sum / a.n_global;
    }
}

// stage-0 evaluators: mode 0 = draw_likelihood (initialization.jl:129-139); mode 1 =
// initialize_likelihoods! (:153-186): old_loglh <- loglh, then loglh/logprior on the new data
template <class LIK>
__global__ void __launch_bounds__(128) k_evaluate(double* __restrict__ cloud, int64_t N, int mode)
{
    constexpr int D = LIK::D;
    const int64_t i = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (i >= N) return;
    double th[D];
#pragma unroll
    for (int k = 0; k < D; ++k) th[k] = cloud[col_off(N, k) + i];
    if (mode == 1) {
        cloud[col_off(N, D + 2) + i] = cloud[col_off(N, D) + i];
        cloud[col_off(N, D) + i] = LIK::template ll<0>(th);
    } else {
        cloud[col_off(N, D) + i] = in_bounds<D>(th) ? LIK::template ll<0>(th) : -dinf();
    }
    cloud[col_off(N, D + 1) + i] = logprior<D>(th);
}

// ---- stage 0: initial_draw! / one_draw (src/initialization.jl:23-119) --------------------------------------
// Every free parameter is drawn from its prior (ModelConstructors `rand(parameters, 1)`: redrawn until strictly
// inside its valuebounds), fixed parameters take their value; the whole vector is redrawn while the
// log-likelihood is not finite (:43-60).  Randomness: a per-particle sequential Philox stream
// (counter = (global particle, 0, call number, PURP_INIT)); one block per uniform / normal (stage 0 only).
struct DrawRng {
    uint64_t seed;
    uint32_t gp, ctr;
    __device__ __forceinline__ u32x4 next() { return rng4(seed, gp, 0u, ctr++, PURP_INIT); }
    __device__ __forceinline__ double uniform() { const u32x4 r = next(); return u01(r.x, r.y); }
    __device__ __forceinline__ double uniform_open0() { const u32x4 r = next(); return u01_open0(r.x, r.y); }
    __device__ __forceinline__ double normal() { double z0, z1; normal_pair(next(), z0, z1); return z0; }
};

// Gamma(shape, 1): Marsaglia & Tsang (2000); shape < 1 through shape + 1 and U^(1/shape).  NaN after 256 rejections.
__device__ __noinline__ double draw_gamma(DrawRng& r, double shape)
{
    const bool boost = shape < 1.0;
    const double a = boost ? shape + 1.0 : shape;
    const double d = a - 1.0 / 3.0;
    const double c = 1.0 / sqrt(9.0 * d);
    double g = dnan();
    for (int it = 0; it < 256; ++it) {
        const double x = r.normal();
        const double t = 1.0 + c * x;
        if (!(t > 0.0)) continue;
        const double v = (t * t) * t;
        const double u = r.uniform_open0();
        if (det_log(u) < ((0.5 * x) * x + d) - d * v + d * det_log(v)) { g = d * v; break; }
    }
    if (boost) {
        const double u = r.uniform_open0();
        g = g * det_exp(det_log(u) / shape);
    }
    return g;
}

__device__ __noinline__ double draw_prior(DrawRng& r, int k)
{
    const double p1 = c_pri.p1[k], p2 = c_pri.p2[k];
    switch (c_pri.kind[k]) {
    case SMCB200_PRIOR_NORMAL: return p1 + p2 * r.normal();
    case SMCB200_PRIOR_UNIFORM: return p1 + (p2 - p1) * r.uniform();
    case SMCB200_PRIOR_GAMMA: return p2 * draw_gamma(r, p1);
    case SMCB200_PRIOR_ROOT_INV_GAMMA: return sqrt(c_pri.a2[k] / draw_gamma(r, 0.5 * p1));   // x^2 ~ InvGamma(nu/2, nu tau^2/2)
    case SMCB200_PRIOR_BETA: { const double x = draw_gamma(r, p1), y = draw_gamma(r, p2); return x / (x + y); }
    case SMCB200_PRIOR_INV_GAMMA: return p2 / draw_gamma(r, p1);
    }
    return dnan();
}

template <class LIK>
__global__ void __launch_bounds__(128) k_initial_draw(double* __restrict__ cloud, int64_t N, int64_t index0, const double* __restrict__ fixed_values,
                                                      uint64_t seed, int max_tries, int* __restrict__ n_failed)
{
    constexpr int D = LIK::D;
    const int64_t i = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (i >= N) return;
    DrawRng r;
    r.seed = seed; r.gp = (uint32_t)(index0 + i); r.ctr = 0u;
    double th[D];
    double ll = -dinf(), lp = -dinf();
    bool success = false;
    for (int attempt = 0; attempt < max_tries && !success; ++attempt) {
#pragma unroll 1
        for (int k = 0; k < D; ++k) {
            double x;
            if (c_pri.fixed[k]) {
                x = fixed_values[k];
            } else {
                x = dnan();
                for (int it = 0; it < 1000; ++it) {
                    x = draw_prior(r, k);
                    if (x > c_pri.lo[k] && x < c_pri.hi[k]) break;
                    x = dnan();
                }
            }
            th[k] = x;
        }
        ll = LIK::template ll<0>(th);
        lp = logprior<D>(th);
        if (!(ll - ll == 0.0)) { ll = -dinf(); lp = -dinf(); }     // -Inf, +Inf or NaN: redraw (:39-41,56-60)
        else success = true;
    }
    if (!success) atomicAdd(n_failed, 1);
#pragma unroll
    for (int k = 0; k < D; ++k) cloud[col_off(N, k) + i] = th[k];
    cloud[col_off(N, D) + i] = ll;
    cloud[col_off(N, D + 1) + i] = lp;
    cloud[col_off(N, D + 2) + i] = 0.0;        // update_old_loglh!(c, zeros)
    cloud[col_off(N, D + 3) + i] = 0.0;
    cloud[col_off(N, D + 4) + i] = 1.0;        // set_weights!(c, ones)
}

// ---- per-translation-unit registration ---------------------------------------------------------------
static int tu_upload_model(Ctx* ctx)
{
    SMC_CUDA(ctx, cudaMemcpyToSymbolAsync(c_pri, &ctx->prior, sizeof(PriorConst), 0, cudaMemcpyHostToDevice, ctx->stream));
    SMC_CUDA(ctx, cudaMemcpyToSymbolAsync(c_lik, ctx->lik_host, sizeof(LikSlot) * 2, 0, cudaMemcpyHostToDevice, ctx->stream));
    SMC_CUDA(ctx, cudaMemcpyToSymbolAsync(c_as, ctx->as_host, sizeof(ASConst) * 2, 0, cudaMemcpyHostToDevice, ctx->stream));
    return SMCB200_OK;
}

static int tu_upload_proposal(Ctx* ctx, bool from_device)
{
    if (from_device) {
        SMC_CUDA(ctx, cudaMemcpyToSymbolAsync(c_mut, ctx->mutc_dev, sizeof(MutConst), 0, cudaMemcpyDeviceToDevice, ctx->stream));
    } else {
        SMC_CUDA(ctx, cudaMemcpyToSymbolAsync(c_mut, ctx->mutc_host, sizeof(MutConst), 0, cudaMemcpyHostToDevice, ctx->stream));
    }
    return SMCB200_OK;
}

template <class LIK>
static KernelEntry make_entry()
{
    KernelEntry e;
    e.kind = LIK::KIND;
    e.neq = LIK::NEQ; e.k = LIK::K; e.stride = LIK::STRIDE; e.coef = LIK::COEF; e.sig = LIK::SIG; e.d = LIK::D;
    e.minb = LIK::MINB;
    e.mut[0][0][0] = k_mutate<LIK, false, 0, false>;
    e.mut[0][1][0] = k_mutate<LIK, false, 1, false>;
    e.mut[0][2][0] = k_mutate<LIK, false, 2, false>;
    e.mut[1][0][0] = k_mutate<LIK, true, 0, false>;
    e.mut[1][1][0] = k_mutate<LIK, true, 1, false>;
    e.mut[1][2][0] = k_mutate<LIK, true, 2, false>;
    e.mut[0][0][1] = k_mutate<LIK, false, 0, true>;
    e.mut[0][1][1] = k_mutate<LIK, false, 1, true>;
    e.mut[0][2][1] = k_mutate<LIK, false, 2, true>;
    e.mut[1][0][1] = k_mutate<LIK, true, 0, true>;
    e.mut[1][1][1] = k_mutate<LIK, true, 1, true>;
    e.mut[1][2][1] = k_mutate<LIK, true, 2, true>;
    e.eval = k_evaluate<LIK>;
    e.draw = k_initial_draw<LIK>;
    e.upload_model = tu_upload_model;
    e.upload_proposal = tu_upload_proposal;
    return e;
}
#define LINREG(K) make_entry<GaussReg<1, K, K, 0, -1>>()

// The regression sizes in between get the general kernel only (BLK = 0 handles any blocking, one block included): a third
// of the code of a full entry, so that no K <= DMAX is refused.
template <class LIK>
static KernelEntry make_entry_lite()
{
    KernelEntry e;
    e.kind = LIK::KIND;
    e.neq = LIK::NEQ; e.k = LIK::K; e.stride = LIK::STRIDE; e.coef = LIK::COEF; e.sig = LIK::SIG; e.d = LIK::D;
    e.minb = LIK::MINB;
    for (int b = 0; b < 3; ++b) {
        e.mut[0][b][0] = k_mutate<LIK, false, 0, false>;
        e.mut[1][b][0] = k_mutate<LIK, true, 0, false>;
        e.mut[0][b][1] = k_mutate<LIK, false, 0, true>;
        e.mut[1][b][1] = k_mutate<LIK, true, 0, true>;
    }
    e.eval = k_evaluate<LIK>;
    e.draw = k_initial_draw<LIK>;
    e.upload_model = tu_upload_model;
    e.upload_proposal = tu_upload_proposal;
    return e;
}
#define LINREG_LITE(K) make_entry_lite<GaussReg<1, K, K, 0, -1>>()

}  // namespace
}  // namespace smc
