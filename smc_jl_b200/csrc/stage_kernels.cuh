// stage_kernels.cuh -- the fused per-stage kernels (src/smc_main.jl:377-465 without the host in the loop):
//   k_correct_coop    ONE cooperative launch for solve_adaptive_phi (src/helpers.jl:9-56) + correction + ESS +
//                     the resample decision (src/smc_main.jl:386-435); every decision is left in device scalars
//   k_moments_mma<NT> weighted mean / covariance in ONE pass over the cloud (src/particle.jl:481-532) as a weighted SYRK
//                     on the FP64 tensor path (mma.sync.m8n8k4.f64), fragments loaded straight from the cloud's columns
//   k_moments_finish  tile trees of the one-pass sums; its last block exchanges them across GPUs, updates the step
//                     size (src/smc_main.jl:453-455) and factors the proposal covariance (src/mutation.jl:81) with
//                     the whole block -> MutConst
#pragma once
#include <cooperative_groups.h>

#include "kernels.cuh"

namespace smc {
namespace cg = cooperative_groups;

// =================================================================================================
// Transition of solve_adaptive_phi (src/helpers.jl:26-54) over one sweep of K trial phi: consumes as many of the
// K evaluations as the sequential algorithm would have made (the schedule walk, then up to log2(K+1) bisection levels),
// i.e. the result is the one of the one-evaluation-at-a-time loop, bit for bit.  One thread.  sq: S_k then Q_k.
// =================================================================================================
template <int K>
__device__ inline void phi_build_tree_k(PhiState* st)
{
    double a[K], b[K];
    a[0] = st->lo; b[0] = st->hi;
    for (int n = 0; n < K; ++n) {
        const double mid = 0.5 * (a[n] + b[n]);
        st->trial[n] = mid;
        if (2 * n + 2 < K) { a[2 * n + 1] = a[n]; b[2 * n + 1] = mid; a[2 * n + 2] = mid; b[2 * n + 2] = b[n]; }
    }
}
template <int K>
__device__ inline void phi_fill_walk_k(PhiState* st, const double* sched)
{
    st->trial[0] = st->phi_prop;
    for (int k = 1; k < K; ++k) {
        long long idx = st->j - 1 + (k - 1);
        if (idx > st->n_phi - 1) idx = st->n_phi - 1;
        st->trial[k] = sched[idx];
    }
}
template <int K>
__device__ inline void phi_transition(PhiState* st, const double* __restrict__ sched, const double* sq)
{
    constexpr int LEVELS = (K == 1) ? 1 : (K == 3) ? 2 : (K == 7) ? 3 : 4;
    static_assert(K == 1 || K == 3 || K == 7 || K == 15, "K must be 2^L - 1");
    if (st->done) return;
    bool finish = false;
    if (st->phase == 0) {
        int k = 0;
        double g;
        for (;;) {
            g = (sq[k] * sq[k]) / sq[K + k] - st->ess_bar;
            st->evals += 1; st->g_last = g;
            if (g >= 0.0 && st->j <= st->n_phi) {
                st->phi_prop = sched[st->j - 1];
                st->j += 1;
                st->phi_cur = st->phi_prop;
                if (++k == K) { phi_fill_walk_k<K>(st, sched); return; }     // more schedule points next sweep
                continue;
            }
            break;
        }
        if (st->phi_prop != 1.0 || g < 0.0) {
            st->lo = st->phi_n1; st->hi = st->phi_prop; st->phase = 1;
        } else {
            st->phi_n = 1.0; st->done = 1;
            return;
        }
    } else {
        int node = 0;
        for (int level = 0; level < LEVELS && !finish; ++level) {
            const double mid = 0.5 * (st->lo + st->hi);
            if (!(mid > st->lo && mid < st->hi)) { finish = true; break; }
            const double g = (sq[node] * sq[node]) / sq[K + node] - st->ess_bar;
            st->evals += 1; st->g_last = g;
            if (g == 0.0) { st->lo = mid; finish = true; }
            else if (g > 0.0) { st->lo = mid; node = 2 * node + 2; }      // continue in (mid, hi): right child
            else { st->hi = mid; node = 2 * node + 1; }                   // continue in (lo, mid): left child
        }
    }
    if (!finish) {
        const double mid = 0.5 * (st->lo + st->hi);
        if (mid > st->lo && mid < st->hi) { st->phi_cur = mid; phi_build_tree_k<K>(st); return; }
    }
    st->phi_n = (st->lo == st->phi_n1) ? st->hi : st->lo;
    st->done = 1;
}

// =================================================================================================
// k_correct_coop: one cooperative launch per stage for everything up to the resample decision.
// The shard's 1024-element tiles are dealt round-robin to the 256-thread groups of at most one 512-thread block per SM;
// every reduction stores the tiles' sums, crosses ONE grid barrier, and then every block finishes the canonical
// adjacent-pair tile tree itself, so all blocks agree bit for bit and run identical copies of the bisection state
// machine; on several GPUs block 0 exchanges the shard roots in place (NVLink mailboxes) and publishes the global sums.
//   adaptive: sweeps of K trial phi (compute_ESS, helpers.jl:173-181; K = 3, 7 or 15 by cloud size: each sweep costs one
//             grid barrier plus K exp per particle) drive the bisection state machine until the bracket is exhausted --
//             the three columns stay in L1/L2, there is no host poll and no extra launch;
//   pass A:   w~ = w * inc, S = sum w~                      (smc_main.jl:401-413, particle.jl:250-259)
//   pass B:   W = (w~ N) / S, Q = sum W^2, sum W, sum W/N   (particle.jl:362-369, smc_main.jl:427)
//   decision: ESS = N^2 / Q, resample iff ESS < threshold_ratio N (smc_main.jl:435); NaN ESS poisons the stage.
// =================================================================================================
struct CoopArgs {
    const double* ll; const double* old; double* w;   // loglh, old_loglh, weight columns of this shard
    double* inc_out; double* normw_out;               // nullable: w_matrix / W_matrix columns of this stage
    int64_t N; double n_global;
    int ntiles, P;                                    // tiles of this shard; P = next power of two (tile-tree width)
    CorrArgs corr;
    double threshold_ratio;
    int adaptive, solve_only, use_carry;
    const double* sched; int n_phi; double tempering_target;
    double c_in, accept_in, ess_prev_in, phi_prop_in; long long j_in; int resampled_last_in;
    double* partials;                                 // [2][NQMAX][P] (two alternating regions), entries >= ntiles stay zero
    double* scal;
    // multi-GPU: block 0 crosses the GPUs and publishes the global sums to the other blocks of its grid
    PeerCtx pc;
    // Every reduction of every launch has a sequence number (ll_step, device resident, identical on all ranks); its
    // parity selects the mailbox slot and its low 32 bits tag the words (the flag travels INSIDE each 8-byte word).
    unsigned* ticket;                                 // block arrival counter (wraps to 0 after every reduction)
    int ll_always;                                    // one GPU: reduce through the (local) mailbox as well instead of a grid barrier
    unsigned long long* ll_step;
};
constexpr int COOP_NT = 512;                         // threads per block: two 1024-element tiles in flight per block
constexpr int COOP_GROUPS = COOP_NT / W_LANES;
constexpr int COOP_KMAX = 7;
constexpr int COOP_NQMAX = 2 * COOP_KMAX;

// per-tile lane sums of the three reductions (functors: nvcc's front end aborts on the equivalent lambdas inside a
// template kernel); t = thread index inside the tile's 256-thread group
template <int K>
struct SweepSums {                 // S_k = sum x, Q_k = sum x^2 with x = w inc(phi_k), k < K   (compute_ESS, helpers.jl:173-181)
    static constexpr int NLOAD = 3 * W_R;
    const CoopArgs& a; const double* phi; double phi_n1;
    __device__ __forceinline__ void load(int64_t tile, int t, double (&b)[NLOAD]) const
    {
        const int64_t base = tile * W_TILE + t;
#pragma unroll
        for (int r = 0; r < W_R; ++r) {
            const int64_t i = base + (int64_t)r * W_LANES;
            const bool in = i < a.N;
            b[3 * r] = in ? __ldg(a.ll + i) : 0.0; b[3 * r + 1] = in ? __ldg(a.old + i) : 0.0; b[3 * r + 2] = in ? a.w[i] : 0.0;
        }
    }
    __device__ __forceinline__ void compute(int64_t tile, int t, const double (&b)[NLOAD], double (&v)[2 * K]) const
    {
#pragma unroll
        for (int q = 0; q < 2 * K; ++q) v[q] = 0.0;
        // branch-free: elements past N were loaded as l = o = w = 0, i.e. x = 0 exp(0) = +0 and v + 0 = v exactly -- the
        // W_R K exponentials of a thread are independent chains the scheduler can interleave
#pragma unroll
        for (int r = 0; r < W_R; ++r) {
            const double l = b[3 * r], o = b[3 * r + 1], wi = b[3 * r + 2];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const double x = wi * det_exp((phi_n1 - phi[k]) * o + (phi[k] - phi_n1) * l);
                v[k] = v[k] + x;
                v[K + k] = v[K + k] + x * x;
            }
        }
    }
};
struct PassASums {                 // w~ = w inc, S = sum w~   (smc_main.jl:401-413, particle.jl:250-259)
    static constexpr int NLOAD = 1;
    const CoopArgs& a; double phi_n;
    __device__ __forceinline__ void load(int64_t, int, double (&)[NLOAD]) const {}
    __device__ __forceinline__ void compute(int64_t tile, int t, const double (&)[NLOAD], double (&v)[1]) const
    {
        const int64_t base = tile * W_TILE + t;
        double x[W_R];
#pragma unroll
        for (int r = 0; r < W_R; ++r) {
            const int64_t i = base + (int64_t)r * W_LANES;
            x[r] = 0.0;
            if (i < a.N) {
                const double inc = inc_weight(__ldg(a.ll + i), __ldg(a.old + i), a.corr, phi_n);
                x[r] = a.w[i] * inc;
                a.w[i] = x[r];
                if (a.inc_out) a.inc_out[i] = inc;
            }
        }
        double acc = 0.0;
#pragma unroll
        for (int r = 0; r < W_R; ++r) acc = acc + x[r];
        v[0] = acc;
    }
};
struct PassBSums {                 // W = (w~ N) / S, Q = sum W^2, sum W, sum W / N   (particle.jl:362-369, smc_main.jl:427,438)
    static constexpr int NLOAD = 1;
    const CoopArgs& a; double S;
    __device__ __forceinline__ void load(int64_t, int, double (&)[NLOAD]) const {}
    __device__ __forceinline__ void compute(int64_t tile, int t, const double (&)[NLOAD], double (&v)[3]) const
    {
        const int64_t base = tile * W_TILE + t;
        double x[W_R], y[W_R];
#pragma unroll
        for (int r = 0; r < W_R; ++r) {
            const int64_t i = base + (int64_t)r * W_LANES;
            x[r] = 0.0; y[r] = 0.0;
            if (i < a.N) {
                x[r] = (a.w[i] * a.n_global) / S;
                y[r] = x[r] / a.n_global;                 // normalized_weights / n_parts, smc_main.jl:438
                a.w[i] = x[r];
                if (a.normw_out) a.normw_out[i] = x[r];
            }
        }
        double q = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
        for (int r = 0; r < W_R; ++r) { q = q + x[r] * x[r]; s2 = s2 + x[r]; s3 = s3 + y[r]; }
        v[0] = q; v[1] = s2; v[2] = s3;
    }
};

constexpr int COOP_RMAX = 8;                         // rounds whose warp sums are parked before the single block barrier
struct CoopSmem {
    double warp[COOP_RMAX][COOP_GROUPS][COOP_NQMAX][8];     // per-tile warp sums
    double seg[COOP_NQMAX][COOP_NT / 32];        // per-warp segment roots of the tile tree
    double res[COOP_NQMAX], tmp[COOP_NQMAX];
    double vals[16][COOP_NQMAX];                 // multi-GPU: every rank's shard roots
    unsigned long long ll_base;                  // sequence number of this launch's first reduction
    int last;
};
static_assert(COOP_NQMAX <= LL_NQ_A, "low-latency mailbox slot too small");

// tile sums of NQ quantities in the canonical order (lane l of a tile adds its 4 elements, adjacent-pair tree over the
// 256 lanes): each 256-thread group of the block takes one tile per round (tiles are dealt round-robin over the grid)
// and stores the tile's sums for the grid-wide tree
template <int NQ, class F>
__device__ __forceinline__ void coop_tile_sums(const CoopArgs& a, CoopSmem& sm, const F& f, double* region)
{
    const int tid = threadIdx.x, g = tid / W_LANES, t = tid % W_LANES, lane = tid & 31, warp = t >> 5;
    const int per_round = gridDim.x * COOP_GROUPS;
    const int rounds = (a.ntiles + per_round - 1) / per_round;
    for (int r0 = 0; r0 < rounds; r0 += COOP_RMAX) {
        const int r1 = (r0 + COOP_RMAX < rounds) ? r0 + COOP_RMAX : rounds;
        __syncthreads();                 // protects sm.warp against the previous reduction's readers
        double nxt[F::NLOAD];
        {
            const int64_t tile = ((int64_t)r0 * COOP_GROUPS + g) * gridDim.x + blockIdx.x;
            if (tile < a.ntiles) f.load(tile, t, nxt);
        }
        for (int r = r0; r < r1; ++r) {
            const int64_t tile = ((int64_t)r * COOP_GROUPS + g) * gridDim.x + blockIdx.x;
            double cur[F::NLOAD];
#pragma unroll
            for (int k = 0; k < F::NLOAD; ++k) cur[k] = nxt[k];
            const int64_t tile_n = ((int64_t)(r + 1) * COOP_GROUPS + g) * gridDim.x + blockIdx.x;
            if (r + 1 < r1 && tile_n < a.ntiles) f.load(tile_n, t, nxt);     // the next round's columns travel during this round
            double v[NQ];
            if (tile < a.ntiles) f.compute(tile, t, cur, v);
            else {
#pragma unroll
                for (int q = 0; q < NQ; ++q) v[q] = 0.0;
            }
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const double s = warp_tree(v[q]);
                if (lane == 0) sm.warp[r - r0][g][q][warp] = s;
            }
        }
        __syncthreads();
        for (int e = tid; e < (r1 - r0) * COOP_GROUPS * NQ; e += COOP_NT) {
            const int q = e % NQ, gg = (e / NQ) % COOP_GROUPS, rr = e / (NQ * COOP_GROUPS);
            const int64_t tile = ((int64_t)(r0 + rr) * COOP_GROUPS + gg) * gridDim.x + blockIdx.x;
            if (tile < a.ntiles) {
                const double* w8 = sm.warp[rr][gg][q];
                __stcg(region + (size_t)q * a.P + tile, ((w8[0] + w8[1]) + (w8[2] + w8[3])) + ((w8[4] + w8[5]) + (w8[6] + w8[7])));
            }
        }
    }
}

// thread tid's contiguous slice of the P tile sums of quantity q (M = P / COOP_NT of them), reduced in tree order
template <int M>
__device__ __forceinline__ double coop_slice(const double* q_part, int tid)
{
    double v[M];
#pragma unroll
    for (int k = 0; k < M; ++k) v[k] = __ldcg(q_part + (size_t)tid * M + k);
#pragma unroll
    for (int sft = 1; sft < M; sft <<= 1)
#pragma unroll
        for (int k = 0; k + sft < M; k += 2 * sft) v[k] = v[k] + v[k + sft];
    return v[0];
}

// One grid-wide reduction of NQ quantities with ONE grid barrier: the blocks store their tile sums, the grid synchronises
// (at most one block per SM takes part, which keeps the barrier short), and then EVERY block finishes the adjacent-pair
// tile tree itself from L2 (thread slices -> 16 warp trees -> one warp), so all blocks hold the same bits without a second
// barrier.  The two partial regions alternate with `step`: a block can only run two reductions ahead of the slowest reader
// after passing the barrier in between.  Multi-GPU: block 0 alone crosses the GPUs (NVLink mailboxes) and publishes the
// global sums behind a flag.  Result: sm.res[0 .. NQ).
// the canonical adjacent-pair tree over the P tile sums of NQ quantities by one block -> sm.tmp[0 .. NQ)
template <int NQ>
__device__ __forceinline__ void coop_tile_tree(const CoopArgs& a, CoopSmem& sm, const double* region)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int P = a.P;
    double x[NQ];                       // all loads first (independent L2 round trips), then the trees
    if (P <= COOP_NT) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) x[q] = (tid < P) ? __ldcg(region + (size_t)q * P + tid) : 0.0;
    } else if (P == 2 * COOP_NT) {
        double v[NQ][2];
#pragma unroll
        for (int q = 0; q < NQ; ++q) { v[q][0] = __ldcg(region + (size_t)q * P + 2 * tid); v[q][1] = __ldcg(region + (size_t)q * P + 2 * tid + 1); }
#pragma unroll
        for (int q = 0; q < NQ; ++q) x[q] = v[q][0] + v[q][1];
    } else {
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const double* qp = region + (size_t)q * P;
            if (P == 4 * COOP_NT) x[q] = coop_slice<4>(qp, tid);
            else if (P == 8 * COOP_NT) x[q] = coop_slice<8>(qp, tid);
            else x[q] = tree_run(qp + (size_t)tid * (P / COOP_NT), P / COOP_NT);
        }
    }
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        const double r = warp_tree(x[q]);
        if (lane == 0) sm.seg[q][warp] = r;
    }
    __syncthreads();
    if (warp < NQ) {
        double y = (lane < COOP_NT / 32) ? sm.seg[warp][lane] : 0.0;
        y = warp_tree(y);
        if (lane == 0) sm.tmp[warp] = y;
    }
    __syncthreads();
}

// One grid-wide reduction of NQ quantities.  The two partial regions alternate with `step`: a block can only run two
// reductions ahead of the slowest reader.  Result: sm.res[0 .. NQ), the same bits in every block (and on every GPU).
//   one GPU:  the blocks store their tile sums, ONE grid barrier (at most two blocks per SM take part, which keeps it
//             short), then EVERY block finishes the adjacent-pair tile tree itself from L2 -- no second barrier.
//   several:  no grid barrier at all.  The block that arrives last (atomic ticket) reduces the shard's tile sums and
//             stores the shard roots straight into every rank's low-latency mailbox over NVLink: each double travels as
//             two 8-byte words that carry the reduction's sequence tag in their upper halves, so a word is either absent
//             or complete -- no fence, no separate flag, one NVLink write latency.  All blocks of all ranks then poll
//             their own rank's mailbox and combine the world's roots in the fixed adjacent-pair order.
template <int NQ, class F>
__device__ __forceinline__ void coop_allreduce(const CoopArgs& a, cg::grid_group& grid, CoopSmem& sm, const F& f, int& step)
{
    const int tid = threadIdx.x;
    double* region = a.partials + (size_t)(step & 1) * COOP_NQMAX * a.P;
    coop_tile_sums<NQ>(a, sm, f, region);
    __threadfence();
    if (a.pc.world == 1 && !a.ll_always) {
        grid.sync();
        coop_tile_tree<NQ>(a, sm, region);
        if (tid < NQ) sm.res[tid] = sm.tmp[tid];
    } else {
        const int world = a.pc.world, rank = a.pc.rank;
        const unsigned long long seq = sm.ll_base + (unsigned long long)step;
        const int par = (int)(seq & 1ull);
        const unsigned long long tag = ((seq + 1ull) & 0xffffffffull) << 32;
        __syncthreads();
        if (tid == 0) {
            const unsigned t = atomicInc(a.ticket, gridDim.x - 1);
            sm.last = (t == gridDim.x - 1);
        }
        __syncthreads();
        if (sm.last) {                                   // block-uniform
            __threadfence();
            coop_tile_tree<NQ>(a, sm, region);
            for (int e = tid; e < world * NQ; e += COOP_NT) {
                const int r = e / NQ, q = e % NQ;
                const unsigned long long bits = (unsigned long long)__double_as_longlong(sm.tmp[q]);
                volatile unsigned long long* dst = ll_region_a(a.pc.inbox[r], world) + ((size_t)(par * 16 + rank) * LL_NQ_A + q) * 2;
                dst[0] = (bits & 0xffffffffull) | tag;
                dst[1] = (bits >> 32) | tag;
            }
        }
        if (tid < world * NQ) {
            const int r = tid / NQ, q = tid % NQ;
            volatile unsigned long long* src = ll_region_a(a.pc.inbox[rank], world) + ((size_t)(par * 16 + r) * LL_NQ_A + q) * 2;
            unsigned long long w0 = src[0], w1 = src[1];
            const long long t0 = clock64();
            while ((w0 & 0xffffffff00000000ull) != tag || (w1 & 0xffffffff00000000ull) != tag) {
                if (clock64() - t0 > 240000000000ll) { *a.pc.err = 1; break; }      // ~2 minutes: a rank died; the host reports it
                w0 = src[0]; w1 = src[1];
            }
            sm.vals[r][q] = __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
        }
        __syncthreads();
        if (tid < NQ) {
            double v[16];
            for (int r = 0; r < world; ++r) v[r] = sm.vals[r][tid];
            for (int sft = 1; sft < world; sft <<= 1)
                for (int i = 0; i + sft < world; i += 2 * sft) v[i] = v[i] + v[i + sft];
            sm.res[tid] = v[0];
        }
    }
    __syncthreads();
    ++step;
}

template <int K>
__device__ __forceinline__ void coop_solve(const CoopArgs& a, cg::grid_group& grid, CoopSmem& sm, PhiState& st, int& step, int& sweeps,
                                           double phi_n1)
{
    const int tid = threadIdx.x;
    if (tid == 0) {
        st.ess_bar = a.resampled_last_in ? a.tempering_target * a.n_global : a.tempering_target * a.ess_prev_in;   // helpers.jl:14-20
        st.phi_prop = a.phi_prop_in; st.phi_cur = a.phi_prop_in; st.phi_n1 = phi_n1; st.phi_n = 0.0;
        st.j = a.j_in; st.n_phi = a.n_phi; st.phase = 0; st.done = 0; st.evals = 0; st.g_last = 0.0;
        st.lo = 0.0; st.hi = 0.0;
        phi_fill_walk_k<K>(&st, a.sched);
    }
    __syncthreads();
    while (!st.done) {                       // block-uniform and grid-uniform: every block runs the same state machine
        double phi[K];
#pragma unroll
        for (int k = 0; k < K; ++k) phi[k] = st.trial[k];
        const SweepSums<K> f{a, phi, phi_n1};
        coop_allreduce<2 * K>(a, grid, sm, f, step);
        if (tid == 0) phi_transition<K>(&st, a.sched, sm.res);
        ++sweeps;
        __syncthreads();
    }
}

// K = trial phi per sweep of the adaptive solve; K = 0: fixed schedule only -- a leaner kernel (no solve code, half the
// registers) of which two blocks fit on an SM, so twice as many tiles are in flight during the two weight passes
#ifndef SMC_COOP_MINB_ADAPTIVE
#define SMC_COOP_MINB_ADAPTIVE 1
#endif
template <int K>
__global__ void __launch_bounds__(COOP_NT, (K == 0) ? 2 : SMC_COOP_MINB_ADAPTIVE) k_correct_coop(CoopArgs a)
{
    cg::grid_group grid = cg::this_grid();
    __shared__ CoopSmem sm;
    __shared__ PhiState st;                     // every block runs its own copy of the bisection state machine
    const int tid = threadIdx.x;
    const bool lead = (blockIdx.x == 0 && tid == 0);
    double* scal = a.scal;
    double phi_n = a.corr.phi_n;
    const double phi_n1 = a.corr.phi_n1;
    int step = 0;
    if (a.pc.world > 1 || a.ll_always) {
        if (tid == 0) sm.ll_base = *reinterpret_cast<volatile unsigned long long*>(a.ll_step);
        __syncthreads();
    }

    if (lead) {
        if (!a.use_carry) { scal[SC_C] = a.c_in; scal[SC_ACCEPT] = a.accept_in; scal[SC_STATUS] = 0.0; }
        scal[SC_EVALS] = 0.0; scal[SC_SWEEPS] = 0.0;
    }
    // ---- solve_adaptive_phi ---------------------------------------------------------------------------
    if constexpr (K > 0) if (a.adaptive) {
        int sweeps = 0;
        coop_solve<K>(a, grid, sm, st, step, sweeps, phi_n1);
        phi_n = st.phi_n;
        if (lead) {
            scal[SC_J] = (double)st.j; scal[SC_PHI_PROP] = st.phi_prop; scal[SC_EVALS] = (double)st.evals;
            scal[SC_SWEEPS] = (double)sweeps;
        }
    }
    if (lead) { scal[SC_PHI_N] = phi_n; scal[SC_PHI_N1] = phi_n1; }
    if (a.solve_only) {
        if (lead && (a.pc.world > 1 || a.ll_always)) *a.ll_step = sm.ll_base + (unsigned long long)step;     // every block read the base before its first reduction
        return;
    }

    // ---- pass A: incremental weights, unnormalised weights, S ------------------------------------------
    {
        const PassASums f{a, phi_n};
        coop_allreduce<1>(a, grid, sm, f, step);
    }
    const double S = sm.res[0];
    __syncthreads();
    // ---- pass B: normalised weights, Q = sum W^2, sum W, sum W / n_parts ----------------------------------
    {
        const PassBSums f{a, S};
        coop_allreduce<3>(a, grid, sm, f, step);
        if (lead) {
            const double n = a.n_global;
            const double ess = (n * n) / sm.res[0];                         // smc_main.jl:427
            scal[SC_S] = S; scal[SC_Q] = sm.res[0]; scal[SC_S2] = sm.res[1]; scal[SC_SRES] = sm.res[2];
            scal[SC_ESS] = ess;
            const bool nan = (ess != ess);                                  // check_nan_ess, helpers.jl:270-305
            if (nan) scal[SC_STATUS] = (double)SMCB200_ERR_NAN_ESS;
            scal[SC_RESAMPLE] = (!nan && scal[SC_STATUS] == 0.0 && ess < a.threshold_ratio * n) ? 1.0 : 0.0;   // smc_main.jl:435
            if (a.pc.world > 1 || a.ll_always) *a.ll_step = sm.ll_base + (unsigned long long)step;
        }
    }
}

// =================================================================================================
// One-pass moments on the FP64 tensor path.  With x0 = the parameter vector of global particle 0 (any point of the
// cloud's support) as shift,
//   Sw = sum w,  m_k = sum w (x_k - x0_k),  C_ab = sum (w (x_a - x0_a)) (x_b - x0_b)
//   mean_k = x0_k + m_k / Sw,   cov_ab = C_ab / Sw - (m_a / Sw)(m_b / Sw)
// which is weighted_mean / weighted_cov (src/particle.jl:481-532; StatsBase.cov(..., corrected = false)) in one sweep
// instead of two (the reference's two-pass result differs from it by rounding only: ~1e-15 relative).
// All 1 + d + d(d+1)/2 sums are the lower triangle of ONE weighted SYRK  G = sum_p (w_p y_p) y_p'  over the augmented
// vector y = (x_0 - x0_0, ..., x_{d-1} - x0_{d-1}, 1): G_ab = C_ab, G_db = m_b, G_dd = Sw.  It runs as
// mma.sync.m8n8k4.f64 (DMMA; measured 37 TFLOP/s on B200 against 32 for DFMA, one issue slot per 256 fma):
// M and N index the variables in tiles of 8, K runs over the particles four at a time; the value a lane loads for
// variable tile t -- y[8t + lane/4][p + lane%4], an 8 x 32-byte sector pattern straight from the cloud's columns -- is the
// B fragment of column tile t and, times w_p, the A fragment of row tile t.  No shared memory, no barriers: a warp owns
// M1P_SC consecutive particles, keeps the NT(NT+1)/2 lower tiles in registers and prefetches the next 16 particles
// while the tensor pipe works on the current ones.  (A warp-private cp.async ring in shared memory was measured slower:
// 8-byte LDGSTS saturate the MIO queue -- 57 vs 50 us at d = 20, N = 2^20.)
// Canonical order per quantity (DMMA accumulates k ascending with one rounding per fma -- profiles/r01_dmma_probe.txt):
// sequential fma chain acc <- fma(w_p y_a[p], y_b[p], acc) over the particles of a sub-chunk in ascending order, then the
// adjacent-pair tree over sub-chunks (zero padded to a power of two).  Mirrored by orc_moments_shifted.
// Partials layout: [1 + d + d(d+1)/2][P] with quantity 0 = Sw, 1 + k = m_k, 1 + d + a(a+1)/2 + b = C_ab.
// =================================================================================================
constexpr int M1P_WARPS = 4;               // sub-chunks per block (measured at d = 20, N = 2^20: 4 warps x 4 blocks per SM 50 us; 2 x 7: 60 us; 4 x 5 at 96 registers: 56 us)
__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// one batch = M1P_KS k-steps = 16 consecutive particles: this lane's fragment elements, straight from the columns.
// Only the last variable tile can hold the constant / the zero padding (8 (NT - 1) <= d): its lanes load nothing.
constexpr int M1P_KS = 4;
template <int NT>
__device__ __forceinline__ void m1p_load(double (&xv)[M1P_KS][NT], double (&wv)[M1P_KS], const double* const (&col)[NT], const double* wcol,
                                         int64_t p, int64_t N, bool real_last, double sh_last)
{
#pragma unroll
    for (int s = 0; s < M1P_KS; ++s) {
        // beyond N: weight 0 and the (finite) row of the last particle, i.e. fma(0 y_a, y_b, acc) = acc
        const int64_t q = p + 4 * s;
        const bool in = q < N;
        const int64_t pc = in ? q : N - 1;
        const double w = __ldg(wcol + pc);
        wv[s] = in ? w : 0.0;
#pragma unroll
        for (int t = 0; t < NT - 1; ++t) xv[s][t] = __ldg(col[t] + pc);
        xv[s][NT - 1] = real_last ? __ldg(col[NT - 1] + pc) : sh_last;
    }
}
template <int NT>
__device__ __forceinline__ void m1p_batch(double (&acc)[NT * (NT + 1) / 2][2], const double (&xv)[M1P_KS][NT], const double (&wv)[M1P_KS],
                                          const double (&sh)[NT], bool real_last, double cst_last)
{
#pragma unroll
    for (int s = 0; s < M1P_KS; ++s) {
        double bf[NT], af[NT];
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            bf[t] = xv[s][t] - sh[t];
            if (t == NT - 1) bf[t] = real_last ? bf[t] : cst_last;
            af[t] = wv[s] * bf[t];
        }
#pragma unroll
        for (int ti = 0; ti < NT; ++ti)
#pragma unroll
            for (int tj = 0; tj <= ti; ++tj) dmma884(acc[ti * (ti + 1) / 2 + tj], af[ti], bf[tj]);
    }
}
template <int NT>
__device__ __forceinline__ void m1p_subchunk(double (&acc)[NT * (NT + 1) / 2][2], const double* const (&col)[NT], const double* wcol, int64_t p,
                                             int64_t N, const double (&sh)[NT], bool real_last, double cst_last)
{
    double xa[M1P_KS][NT], wa[M1P_KS], xb[M1P_KS][NT], wb[M1P_KS];
    constexpr int B = 4 * M1P_KS;                       // particles per batch
    m1p_load<NT>(xa, wa, col, wcol, p, N, real_last, sh[NT - 1]);
#pragma unroll 1
    for (int it = 0; it < M1P_SC / (2 * B); ++it) {     // two batches per trip: the register sets swap roles, no copies
        m1p_load<NT>(xb, wb, col, wcol, p + B, N, real_last, sh[NT - 1]);
        m1p_batch<NT>(acc, xa, wa, sh, real_last, cst_last);
        m1p_load<NT>(xa, wa, col, wcol, p + 2 * B, N, real_last, sh[NT - 1]);      // (the last trip prefetches a clamped, unused batch)
        m1p_batch<NT>(acc, xb, wb, sh, real_last, cst_last);
        p += 2 * B;
    }
}

// wcol: the weight column (current buffer); x0 / x1: the parameter columns (x1 = the gather target, used when
// scal[SC_RESAMPLE] != 0); shift_base[k * shift_stride] = parameter k of global particle 0.  8 (NT - 1) <= d < 8 NT.
// The warps of a block own adjacent sub-chunks and add their sums in the canonical tree order before they leave the
// block: partials[q][block] is the tree node over M1P_WARPS M1P_SC = 1024 particles.
template <int NT>
__global__ void __launch_bounds__(32 * M1P_WARPS)
k_moments_mma(const double* __restrict__ x0, const double* __restrict__ x1, const double* __restrict__ wcol, int64_t N, int d,
              const double* __restrict__ scal, const double* __restrict__ shift_base, int64_t shift_stride,
              double* __restrict__ partials, int P)
{
    constexpr int NTILE = NT * (NT + 1) / 2;
    constexpr int REDW = NTILE * 64 + 2;
    __shared__ double red[M1P_WARPS * REDW];
    pdl_wait();
    pdl_trigger();
    if (scal && scal[SC_STATUS] != 0.0) return;
    const double* __restrict__ X = (scal && scal[SC_RESAMPLE] != 0.0) ? x1 : x0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, k = lane & 3;
    const int64_t c0 = ((int64_t)blockIdx.x * M1P_WARPS + warp) * M1P_SC;
    double acc[NTILE][2];
#pragma unroll
    for (int e = 0; e < NTILE; ++e) { acc[e][0] = 0.0; acc[e][1] = 0.0; }
    if (c0 < N) {
        // this lane's variable in tile t: a parameter (load, subtract the shift), the constant 1, or zero padding
        const double* col[NT];
        double sh[NT];
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            const int var = 8 * t + g;
            col[t] = X + col_off(N, var < d ? var : 0);
            sh[t] = (var < d) ? shift_base[(size_t)var * shift_stride] : 0.0;
        }
        const int var_last = 8 * (NT - 1) + g;
        const bool real_last = var_last < d;
        const double cst_last = (var_last == d) ? 1.0 : 0.0;
        m1p_subchunk<NT>(acc, col, wcol, c0 + k, N, sh, real_last, cst_last);
    }
    // C fragment: rows 8 ti + g, columns 8 tj + 2k, + 1  ->  red[warp][tile][row][column]
#pragma unroll
    for (int e = 0; e < NTILE; ++e) {
        red[warp * REDW + e * 64 + g * 8 + 2 * k] = acc[e][0];
        red[warp * REDW + e * 64 + g * 8 + 2 * k + 1] = acc[e][1];
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < NTILE * 64; idx += 32 * M1P_WARPS) {
        const int e = idx >> 6, r = (idx >> 3) & 7, cc = idx & 7;
        int ti = 0;
        while ((ti + 1) * (ti + 2) / 2 <= e) ++ti;
        const int tj = e - ti * (ti + 1) / 2;
        const int a = 8 * ti + r, b = 8 * tj + cc;
        if (a <= d && b <= a) {
            const int q = (a == d) ? ((b == d) ? 0 : 1 + b) : 1 + d + a * (a + 1) / 2 + b;
            static_assert(M1P_WARPS == 4, "tree over the block's sub-chunks");
            partials[(size_t)q * P + blockIdx.x] = (red[idx] + red[REDW + idx]) + (red[2 * REDW + idx] + red[3 * REDW + idx]);
        }
    }
}

// Proposal preparation with the whole block: mean, covariance, step size, per-block Cholesky -> MutConst.
// Same per-element arithmetic as build_mutconst() / cholesky_lower() (each L_ij is the same fma chain over k ascending).
struct PrepSmem {
    double cov[DMAX][DMAX + 1], S[DMAX][DMAX + 1], L[DMAX][DMAX + 1];
    double mean[DMAX], e[DMAX], logs[DMAX], diag[DMAX];
    int bad;
};
__device__ __forceinline__ void prepare_proposal_block(PrepSmem& sm, const BlockSpec& bs, double c, MutConst* out, double* scal)
{
    const int tid = threadIdx.x, nt = blockDim.x;
    const int d = bs.d;
    if (tid == 0) { out->n_blocks = bs.n_blocks; out->status = 0; sm.bad = 0; }
    if (tid < DMAX) out->mu[tid] = (tid < d) ? sm.mean[tid] : 0.0;
    __syncthreads();
    for (int b = 0; b < bs.n_blocks; ++b) {
        const int n = bs.bsize[b];
        for (int e = tid; e < PACKMAX; e += nt) out->L[b][e] = 0.0;
        if (tid < DMAX) { out->csd[b][tid] = 0.0; out->isd[b][tid] = 0.0; out->isdn[b][tid] = 0.0; out->rl[b][tid] = 0.0; }
        if (tid == 0) {
            uint32_t mask = 0;
            for (int i = 0; i < n; ++i) mask |= 1u << bs.member[b][i];
            out->mask[b] = mask; out->bsize[b] = n;
        }
        for (int e = tid; e < n * n; e += nt) {
            const int i = e / n, j = e % n;
            const int ai = bs.member[b][i], aj = bs.member[b][j];
            sm.S[i][j] = (sm.cov[ai][aj] + sm.cov[aj][ai]) / 2.0;          // R_fr = (R + R') / 2, smc_main.jl:462
        }
        __syncthreads();
        // right-looking factorisation by the whole block: column j, then the rank-one update of the trailing block.  Every
        // entry receives fma(-L_ik, L_jk, .) for k = 0, 1, ... in this order -- the chain of cholesky_lower(), bit for bit --
        // but a column costs one sqrt, one division and two barriers instead of a warp-serial dot product.
        if (tid < n) sm.diag[tid] = sm.S[tid][tid];
        for (int j = 0; j < n; ++j) {
            const double sjj = sm.S[j][j];
            if (!(sjj > 0.0)) { if (tid == 0) sm.bad = 1; break; }          // block-uniform
            const double dj = sqrt(sjj);
            if (tid == 0) sm.L[j][j] = dj;
            for (int i = j + 1 + tid; i < n; i += nt) sm.L[i][j] = sm.S[i][j] / dj;
            __syncthreads();
            const int m = n - j - 1;
            for (int e = tid; e < m * m; e += nt) {
                const int i = j + 1 + e / m, l = j + 1 + e % m;
                if (l <= i) sm.S[i][l] = fma(-sm.L[i][j], sm.L[l][j], sm.S[i][l]);
            }
            __syncthreads();
        }
        __syncthreads();
        if (sm.bad) {
            if (tid == 0) { out->status = SMCB200_ERR_NOT_POSDEF; if (scal) scal[SC_STATUS] = (double)SMCB200_ERR_NOT_POSDEF; }
            return;
        }
        for (int e = tid; e < n * n; e += nt) {
            const int i = e / n, j = e % n;
            if (j <= i) {
                const int ai = bs.member[b][i], aj = bs.member[b][j];
                out->L[b][aj * d - (aj * (aj - 1)) / 2 + (ai - aj)] = c * sm.L[i][j];
            }
        }
        if (tid < n) {
            const int ai = bs.member[b][tid];
            out->csd[b][ai] = c * sqrt(sm.diag[tid]);
            const double isd = 1.0 / sqrt(sm.diag[tid]);
            out->isd[b][ai] = isd;
            out->isdn[b][ai] = isd * 0x1.9884533d43651p-2;
            out->rl[b][ai] = 1.0 / (c * sm.L[tid][tid]);
            sm.logs[tid] = det_log(c * sm.L[tid][tid]);
        }
        __syncthreads();
        if (tid == 0) {
            double ld = 0.0;
            for (int i = 0; i < n; ++i) ld = ld + sm.logs[i];
            out->lognorm[b] = (double)n * (2.0 * 0.91893853320467274178) + 2.0 * ld;
        }
        __syncthreads();
    }
}

// One block per quantity reduces its chunk partials; the last block to finish crosses the GPUs, forms mean / cov, updates
// the step size c <- c f(accept) (smc_main.jl:453-455; both live in scal[]) and prepares the proposal.
__global__ void __launch_bounds__(256)
k_moments_finish(const double* __restrict__ partials, int P, int nq, double* __restrict__ sums_loc, double* __restrict__ sums,
                 unsigned* counter, PeerCtx pc, const double* __restrict__ shift_base, int64_t shift_stride, BlockSpec bs,
                 double target, double* scal, MutConst* out)
{
    __shared__ double smr[8];
    __shared__ bool is_last;
    __shared__ PrepSmem ps;
    pdl_wait();
    if (scal[SC_STATUS] != 0.0) return;
    const double r = tiles_tree_block<256>(partials + (size_t)blockIdx.x * P, P, smr);
    if (threadIdx.x == 0) {
        __stcg(sums_loc + blockIdx.x, r);
        __threadfence();
        const unsigned t = atomicInc(counter, gridDim.x - 1);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (pc.world > 1) peer_exchange_block(pc, sums_loc, nq, 1, sums);
    else for (int q = threadIdx.x; q < nq; q += blockDim.x) sums[q] = __ldcg(sums_loc + q);
    __syncthreads();
    const int d = bs.d;
    const double sw = sums[0];
    if ((int)threadIdx.x < d) {
        const double e = sums[1 + threadIdx.x] / sw;
        ps.e[threadIdx.x] = e;
        ps.mean[threadIdx.x] = shift_base[(size_t)threadIdx.x * shift_stride] + e;
    }
    __syncthreads();
    for (int q = threadIdx.x; q < d * (d + 1) / 2; q += blockDim.x) {
        int a = 0;
        while ((a + 1) * (a + 2) / 2 <= q) ++a;
        const int b = q - a * (a + 1) / 2;
        const double v = fma(-ps.e[a], ps.e[b], sums[1 + d + q] / sw);
        ps.cov[a][b] = v;
        ps.cov[b][a] = v;
    }
    __shared__ double c_new;
    if (threadIdx.x == 0) {
        c_new = update_step_size(scal[SC_C], scal[SC_ACCEPT], target);
        scal[SC_C] = c_new;
    }
    __syncthreads();
    prepare_proposal_block(ps, bs, c_new, out, scal);
}

// ---- FP64 peak probe (bench.py's roofline_fp64 denominator): 8 independent DFMA chains per thread ------------
__global__ void __launch_bounds__(256) k_fp64_peak(double* out, int iters, double a, double b)
{
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

}  // namespace smc
