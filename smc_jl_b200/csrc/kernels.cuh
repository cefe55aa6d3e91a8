// kernels.cuh -- K1..K6: correction/ESS, the multi-trial adaptive-phi sweep and its device state machine,
// resampling (scan + search + gather), moments, proposal preparation, and the NVLink mailbox exchange
// that carries the small cross-GPU reductions.  The streaming kernels run over struct-of-arrays columns
// (HBM-bound); every reduction follows the canonical orders of DESIGN.md "Numerical contract".
#pragma once
#include "common.cuh"
#include "reduce.cuh"

namespace smc {

// =================================================================================================
// Small cross-GPU reductions without a collective library: ONE launch per reduction.  Every rank pushes its nq
// shard-local roots straight into every peer's inbox over NVLink (plain remote stores through CUDA-IPC mappings),
// publishes an epoch flag behind a system-scope fence, waits for the peers' flags in its own inbox and combines the
// world x nq values in the fixed rank-order tree (combine != 0) or lays them out as [world][nq] (gather).
// Inbox slots alternate with the epoch's parity: a rank can only reach epoch e + 1 after it has seen every peer's
// epoch-e flag, and a peer publishes e only after it has consumed e - 1, so two slots never collide.
// A peer that never arrives (a crashed rank) trips a clock-based time-out instead of hanging the GPU.
// =================================================================================================
__global__ void __launch_bounds__(256)
k_peer_exchange(const double* __restrict__ local_src, int nq, PeerCtx pc, int combine, double* __restrict__ dst, const double* flag)
{
    pdl_wait();
    pdl_trigger();
    if (flag && *flag == 0.0) return;
    peer_exchange_block(pc, local_src, nq, combine, dst);
}

// =================================================================================================
// K1 / K2: correction and ESS  (src/smc_main.jl:400-427, src/helpers.jl:173-181)
// =================================================================================================
struct CorrArgs {
    double phi_n1, phi_n, pw, lpod, log_1m_pw;
    int mode;   // 0: prior weight 0, 1: prior weight 1, 2: in between
};

__device__ __forceinline__ double inc_weight(double ll, double old, const CorrArgs& a, double phi_n)
{
    if (a.mode == 0) return det_exp((a.phi_n1 - phi_n) * old + (phi_n - a.phi_n1) * ll);
    if (a.mode == 1) return det_exp((phi_n - a.phi_n1) * ll);
    const double inner = det_log(det_exp((old - a.lpod) + a.log_1m_pw) + a.pw);
    return det_exp((a.phi_n1 - phi_n) * inner + (phi_n - a.phi_n1) * ll);
}

// -------------------------------------------------------------------------------------------------
// K2: compute_ESS (src/helpers.jl:173-181) at ESS_K trial phi in ONE pass over the three columns.
// ESS(phi) = N^2 / sum (N x_i / S)^2 with x_i = w_i inc_i(phi), S = sum x_i -- N cancels, so one pass gives
// S_k = sum x and Q_k = sum x^2 for every trial and ESS_k = S_k^2 / Q_k (the correction step itself keeps the
// reference's normalise-then-square order).  Quantity q < K is S_q, q >= K is Q_{q-K}; each follows the canonical
// weight-sum order (4 elements per lane, 256-lane tree here, tile tree in k_tree_finalize).
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_ess_multi(const double* __restrict__ ll, const double* __restrict__ old, const double* __restrict__ w, int64_t N,
            double phi_n1, const double* __restrict__ trials, const int* __restrict__ done_flag,
            double* __restrict__ partials, int P)
{
    __shared__ double sm[2 * ESS_K][8];
    if (done_flag && *done_flag) return;
    double phi[ESS_K];
#pragma unroll
    for (int k = 0; k < ESS_K; ++k) phi[k] = trials[k];
    double aS[ESS_K], aQ[ESS_K];
#pragma unroll
    for (int k = 0; k < ESS_K; ++k) { aS[k] = 0.0; aQ[k] = 0.0; }
    const int64_t base = (int64_t)blockIdx.x * W_TILE + threadIdx.x;
#pragma unroll
    for (int r = 0; r < W_R; ++r) {
        const int64_t i = base + (int64_t)r * W_LANES;
        if (i < N) {
            const double l = ll[i], o = old[i], wi = w[i];
#pragma unroll
            for (int k = 0; k < ESS_K; ++k) {
                const double x = wi * det_exp((phi_n1 - phi[k]) * o + (phi[k] - phi_n1) * l);
                aS[k] = aS[k] + x;
                aQ[k] = aQ[k] + x * x;
            }
        } else {
#pragma unroll
            for (int k = 0; k < ESS_K; ++k) { aS[k] = aS[k] + 0.0; aQ[k] = aQ[k] + 0.0 * 0.0; }
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < ESS_K; ++k) {
        const double s = warp_tree(aS[k]), q = warp_tree(aQ[k]);
        if (lane == 0) { sm[k][warp] = s; sm[ESS_K + k][warp] = q; }
    }
    __syncthreads();
    if (threadIdx.x < 2 * ESS_K) {
        const double* v = sm[threadIdx.x];
        partials[(size_t)threadIdx.x * P + blockIdx.x] = ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
    }
}

// canonical sum of one column (optionally divided by a constant first): out = sum_i x_i / div
__global__ void __launch_bounds__(256)
k_colsum(const double* __restrict__ x, int64_t N, double div, int use_div, double* partials, int ntiles, int P,
         unsigned* counter, double* out, const double* flag = nullptr)
{
    __shared__ double sm[8];
    if (flag && *flag == 0.0) return;
    const int64_t base = (int64_t)blockIdx.x * W_TILE + threadIdx.x;
    double acc = 0.0;
#pragma unroll
    for (int r = 0; r < W_R; ++r) {
        const int64_t i = base + (int64_t)r * W_LANES;
        double v = 0.0;
        if (i < N) v = use_div ? x[i] / div : x[i];
        acc = acc + v;
    }
    double v[1] = {block_tree_256(acc, sm)};
    finish_tiles<1>(v, partials, ntiles, P, counter, out, sm);
}

// =================================================================================================
// K3: Julia-style pairwise cumsum with 16-element sequential leaves (src/resample.jl:29,47)
// =================================================================================================
__host__ __device__ inline int lvl_off(int n_level0, int k) { return 2 * n_level0 - ((2 * n_level0) >> k); }

// FINAL = false: leaf totals + in-block tree -> blocktot[b]
// FINAL = true : + top-down offsets from blockoff[b]; writes the block-local running max of the
//                cumsum (rmax), optionally the raw cumsum, and the block maximum.
template <bool FINAL>
__global__ void __launch_bounds__(SCAN_THREADS)
k_scan(const double* __restrict__ src, int div_n, double n_parts, const double* __restrict__ S_ptr, int64_t n, int B,
       double* __restrict__ blocktot, const double* __restrict__ blockoff, double* __restrict__ rmax,
       double* __restrict__ craw, double* __restrict__ bmax, const double* flag = nullptr)
{
    pdl_wait();
    pdl_trigger();
    extern __shared__ double sm_dyn[];
    if (flag && *flag == 0.0) return;
    double* tile = sm_dyn;                               // [nleaf][LEAF + 1]
    double* lv = sm_dyn + SCAN_THREADS * (LEAF + 1);     // [2*nleaf] tree levels
    double* lmax = lv + 2 * SCAN_THREADS;        // [nleaf]
    const int tid = threadIdx.x;
    const int nleaf = B / LEAF;
    const int64_t base = (int64_t)blockIdx.x * B;
    const double S = *S_ptr;
    for (int k = tid; k < B; k += SCAN_THREADS) {
        const int64_t i = base + k;
        double x = 0.0;
        if (i < n) {
            x = src[i];
            if (div_n) x = x / n_parts;
            x = x / S;
        }
        tile[(k / LEAF) * (LEAF + 1) + (k % LEAF)] = x;
    }
    __syncthreads();
    double total = 0.0;
    if (tid < nleaf) {
        double run = 0.0;
#pragma unroll
        for (int i = 0; i < LEAF; ++i) {
            run = run + tile[tid * (LEAF + 1) + i];
            tile[tid * (LEAF + 1) + i] = run;
        }
        total = run;
        lv[tid] = total;
    }
    int nlev = 0;
    while ((1 << nlev) < nleaf) ++nlev;
    for (int k = 1; k <= nlev; ++k) {
        __syncthreads();
        if (tid < (nleaf >> k))
            lv[lvl_off(nleaf, k) + tid] = lv[lvl_off(nleaf, k - 1) + 2 * tid] + lv[lvl_off(nleaf, k - 1) + 2 * tid + 1];
    }
    __syncthreads();
    if (!FINAL) {
        if (tid == 0) blocktot[blockIdx.x] = lv[lvl_off(nleaf, nlev)];
        return;
    }
    if (tid < nleaf) {
        double off = blockoff[blockIdx.x];
        for (int k = nlev - 1; k >= 0; --k) {
            const int node = tid >> k;
            if (node & 1) off = off + lv[lvl_off(nleaf, k) + node - 1];
        }
        const int64_t first = base + (int64_t)tid * LEAF;
        int64_t valid = n - first;
        if (valid > LEAF) valid = LEAF;
        double last = -dinf();
#pragma unroll
        for (int i = 0; i < LEAF; ++i) {
            const double c = off + tile[tid * (LEAF + 1) + i];
            tile[tid * (LEAF + 1) + i] = c;
            if (i < valid) last = c;    // c is non-decreasing inside a leaf
        }
        lmax[tid] = last;
    }
    // inclusive prefix max over the leaves (Hillis-Steele; max is exact, any order)
    for (int s = 1; s < nleaf; s <<= 1) {
        __syncthreads();
        double v = 0.0;
        if (tid < nleaf) v = (tid >= s) ? fmax(lmax[tid], lmax[tid - s]) : lmax[tid];
        __syncthreads();
        if (tid < nleaf) lmax[tid] = v;
    }
    __syncthreads();
    for (int k = tid; k < B; k += SCAN_THREADS) {
        const int64_t i = base + k;
        if (i < n) {
            const int leaf = k / LEAF;
            const double c = tile[leaf * (LEAF + 1) + (k % LEAF)];
            const double pm = (leaf > 0) ? lmax[leaf - 1] : -dinf();
            rmax[i] = fmax(c, pm);
            if (craw) craw[i] = c;
        }
    }
    if (tid == 0) bmax[blockIdx.x] = lmax[nleaf - 1];
}

// upper tree over the nb (power of two) block totals of this shard, single block.
// up:   levels + shard root (root_out).
// down: block offsets, top-down, starting from the shard's own offset in the cross-rank tree
//       (rank_roots = gathered shard roots; nullptr on one GPU).
__global__ void __launch_bounds__(256)
k_scan_upper_up(const double* __restrict__ blocktot, int nb, double* __restrict__ lv, double* __restrict__ root_out,
                const double* flag = nullptr)
{
    pdl_wait();
    pdl_trigger();
    if (flag && *flag == 0.0) return;
    int nlev = 0;
    while ((1 << nlev) < nb) ++nlev;
    for (int i = threadIdx.x; i < nb; i += blockDim.x) lv[i] = blocktot[i];
    for (int k = 1; k <= nlev; ++k) {
        __syncthreads();
        for (int i = threadIdx.x; i < (nb >> k); i += blockDim.x)
            lv[lvl_off(nb, k) + i] = lv[lvl_off(nb, k - 1) + 2 * i] + lv[lvl_off(nb, k - 1) + 2 * i + 1];
    }
    __syncthreads();
    if (threadIdx.x == 0) *root_out = lv[lvl_off(nb, nlev)];
}

__global__ void __launch_bounds__(256)
k_scan_upper_down(const double* __restrict__ lv, int nb, const double* __restrict__ rank_roots, int world, int rank,
                  double* __restrict__ blockoff, const double* flag = nullptr)
{
    pdl_wait();
    pdl_trigger();
    __shared__ double off0;
    if (flag && *flag == 0.0) return;
    if (threadIdx.x == 0) {
        double off = 0.0;
        if (rank_roots && world > 1) {
            // levels of the cross-rank tree over the gathered shard roots (world <= 16)
            double t[5][16];
            int wl = 0;
            while ((1 << wl) < world) ++wl;
            for (int r = 0; r < world; ++r) t[0][r] = rank_roots[r];
            for (int k = 1; k <= wl; ++k)
                for (int i = 0; i < (world >> k); ++i) t[k][i] = t[k - 1][2 * i] + t[k - 1][2 * i + 1];
            for (int k = wl - 1; k >= 0; --k) {
                const int node = rank >> k;
                if (node & 1) off = off + t[k][node - 1];
            }
        }
        off0 = off;
    }
    __syncthreads();
    int nlev = 0;
    while ((1 << nlev) < nb) ++nlev;
    for (int b = threadIdx.x; b < nb; b += blockDim.x) {
        double off = off0;
        for (int k = nlev - 1; k >= 0; --k) {
            const int node = b >> k;
            if (node & 1) off = off + lv[lvl_off(nb, k) + node - 1];
        }
        blockoff[b] = off;
    }
}

// multi-GPU: bmax [world][nb] holds every shard's inclusive prefix max of its own block maxima; fold the maxima of the
// lower ranks in, which makes it the inclusive prefix max over the global block sequence.  One block.
__global__ void __launch_bounds__(256) k_fix_rank_carry(double* __restrict__ bmax, int nb, int world, const double* flag = nullptr)
{
    __shared__ double carry[16];
    pdl_wait();
    pdl_trigger();
    if (flag && *flag == 0.0) return;
    if ((int)threadIdx.x < world) {
        double cmax = -dinf();
        for (int r = 0; r < (int)threadIdx.x; ++r) cmax = fmax(cmax, bmax[(size_t)r * nb + nb - 1]);
        carry[threadIdx.x] = cmax;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < nb * world; e += blockDim.x) bmax[e] = fmax(bmax[e], carry[e / nb]);
}

// inclusive prefix max of the block maxima (one block of 256 threads; max is exact in any order)
__global__ void __launch_bounds__(256) k_prefix_max(double* bmax, int nb, const double* flag = nullptr)
{
    pdl_wait();
    pdl_trigger();
    __shared__ double sm[256];
    __shared__ double carry;
    if (flag && *flag == 0.0) return;
    const int tid = threadIdx.x;
    if (tid == 0) carry = -dinf();
    __syncthreads();
    for (int base = 0; base < nb; base += 256) {
        const int i = base + tid;
        double v = (i < nb) ? bmax[i] : -dinf();
        sm[tid] = v;
        for (int s = 1; s < 256; s <<= 1) {
            __syncthreads();
            const double o = (tid >= s) ? sm[tid - s] : -dinf();
            __syncthreads();
            v = fmax(v, o);
            sm[tid] = v;
        }
        __syncthreads();
        v = fmax(v, carry);
        if (i < nb) bmax[i] = v;
        __syncthreads();
        if (tid == 255) carry = v;
        __syncthreads();
    }
}

// =================================================================================================
// K4: ancestor search.  Sequential semantics of src/resample.jl:51-70 ("first j >= previous ancestor
// with cum[j] > threshold") == first exceedance of the threshold in the running max of cum
// (thresholds are non-decreasing), found by a two-level binary search.  "Not found" (resample.jl:60
// returns 0) is clamped to n.
// =================================================================================================
__global__ void __launch_bounds__(256)
k_search(const double* __restrict__ rmax_single, const double* const* __restrict__ rmax_tab, int nb_rank,
         const double* __restrict__ bmax_incl, int nb, int B, int64_t n,
         int64_t n_out, int64_t out0, int method, uint64_t seed, uint32_t stage, double u, double n_parts,
         int64_t* __restrict__ idx, const double* flag = nullptr)
{
    pdl_wait();
    pdl_trigger();
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n_out) return;
    if (flag && *flag == 0.0) return;
    const int64_t gi = out0 + i;
    double t;
    if (method == SMCB200_RESAMPLE_SYSTEMATIC) {
        t = ((double)gi + u) / n_parts;                       // (i - 1 + offset) / n_parts, resample.jl:53
    } else {
        const u32x4 r = rng4(seed, (uint32_t)gi, stage, 1u, PURP_RESAMPLE);
        t = u01(r.x, r.y);                                    // offset[i], resample.jl:30
    }
    int lo = 0, hi = nb;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (bmax_incl[mid] > t) hi = mid; else lo = mid + 1;
    }
    int64_t res = n;
    if (lo < nb) {
        // running max of block `lo`: this GPU's buffer, or the owning rank's over NVLink (CUDA-IPC mapping); rm[g - base]
        // is the value at global index g
        const double* rm = rmax_single;
        int64_t base = 0;
        if (!rmax_single) {
            const int owner = lo / nb_rank;
            rm = rmax_tab[owner];
            base = (int64_t)owner * nb_rank * B;
        }
        int64_t jlo = (int64_t)lo * B, jhi = jlo + B;
        if (jhi > n) jhi = n;
        while (jlo < jhi) {
            const int64_t mid = (jlo + jhi) >> 1;
            if (rm[mid - base] > t) jhi = mid; else jlo = mid + 1;
        }
        res = jlo + 1;
        if (res > n) res = n;
    }
    idx[i] = res;
}

// =================================================================================================
// K5: gather rows by ancestor + reset weights (src/smc_main.jl:440-442, particle.jl:378-383)
// =================================================================================================
__global__ void __launch_bounds__(256)
k_gather(const double* __restrict__ src, double* __restrict__ dst, const int64_t* __restrict__ idx, int64_t N,
         int ncopy, double* __restrict__ wreset, double* __restrict__ nw_hist, const double* flag = nullptr)
{
    pdl_wait();
    pdl_trigger();
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= N) return;
    if (flag && *flag == 0.0) return;
    const int64_t a = idx[i] - 1;
    int c = 0;
    for (; c + 4 <= ncopy; c += 4) {
        const double v0 = src[col_off(N, c) + a], v1 = src[col_off(N, c + 1) + a];
        const double v2 = src[col_off(N, c + 2) + a], v3 = src[col_off(N, c + 3) + a];
        dst[col_off(N, c) + i] = v0; dst[col_off(N, c + 1) + i] = v1;
        dst[col_off(N, c + 2) + i] = v2; dst[col_off(N, c + 3) + i] = v3;
    }
    for (; c < ncopy; ++c) dst[col_off(N, c) + i] = src[col_off(N, c) + a];
    wreset[i] = 1.0;                       // reset_weights!, particle.jl:378-383
    if (nw_hist) nw_hist[i] = 1.0;         // W_matrix[:, i] .= 1, smc_main.jl:445
}

// multi-GPU gather: ancestors are GLOBAL indices; rows are read straight from the owning rank's cloud
// through CUDA-IPC peer mappings over NVLink (systematic ancestors are sorted, so a warp's reads of one
// column are near-contiguous on one or two peers) -- the "all-to-all(v)" of the stage, done by the kernel.
__global__ void __launch_bounds__(256)
k_gather_peer(double* const* __restrict__ peers, const int64_t* __restrict__ peer_cnt, int64_t per, double* __restrict__ dst,
              const int64_t* __restrict__ idx, int64_t N, int ncopy, double* __restrict__ wreset, double* __restrict__ nw_hist,
              const double* flag = nullptr)
{
    pdl_wait();
    pdl_trigger();
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= N) return;
    if (flag && *flag == 0.0) return;
    const int64_t a = idx[i] - 1;
    const int r = (int)(a / per);
    const int64_t row = a - (int64_t)r * per;
    const double* __restrict__ src = peers[r];
    const int64_t nr = peer_cnt[r];
    int c = 0;
    for (; c + 4 <= ncopy; c += 4) {
        const double v0 = src[col_off(nr, c) + row], v1 = src[col_off(nr, c + 1) + row];
        const double v2 = src[col_off(nr, c + 2) + row], v3 = src[col_off(nr, c + 3) + row];
        dst[col_off(N, c) + i] = v0; dst[col_off(N, c + 1) + i] = v1;
        dst[col_off(N, c + 2) + i] = v2; dst[col_off(N, c + 3) + i] = v3;
    }
    for (; c < ncopy; ++c) dst[col_off(N, c) + i] = src[col_off(nr, c) + row];
    wreset[i] = 1.0;
    if (nw_hist) nw_hist[i] = 1.0;
}

// =================================================================================================
// K6: moments (src/particle.jl:481-486,526-532).  Canonical order: lane l of a 2048-particle tile
// accumulates particles l, l+32, ... (64 of them) with fma, then a 32-lane adjacent-pair tree,
// then the tile tree.  One warp per tile.
// =================================================================================================
// pass 1: msum[0] = sum w, msum[1+k] = sum w x_k.  partials layout [1+d][P].
// One warp per (tile, quantity): grid = (ntiles, ceil((1+d)/4)), 4 warps per block; the weight column is
// re-read by every warp of a tile from L1/L2.
__global__ void __launch_bounds__(128)
k_moments1(const double* __restrict__ cloud, int64_t N, int d, double* __restrict__ partials, int P)
{
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.y * 4 + (threadIdx.x >> 5);
    if (q > d) return;
    const int64_t base = (int64_t)blockIdx.x * M_TILE + lane;
    const double* __restrict__ w = cloud + col_off(N, d + 4);
    const double* __restrict__ x = (q == 0) ? nullptr : cloud + col_off(N, q - 1);
    double acc = 0.0;
    if (base + (int64_t)(M_R - 1) * M_LANES < N) {          // full tile: all loads issued up front
        double wv[16], xv[16];
#pragma unroll
        for (int r0 = 0; r0 < M_R; r0 += 16) {
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                const int64_t i = base + (int64_t)(r0 + r) * M_LANES;
                wv[r] = w[i];
                xv[r] = x ? x[i] : 0.0;
            }
#pragma unroll
            for (int r = 0; r < 16; ++r) acc = x ? fma(wv[r], xv[r], acc) : acc + wv[r];
        }
    } else {
        for (int r = 0; r < M_R; ++r) {
            const int64_t i = base + (int64_t)r * M_LANES;
            if (i < N) acc = x ? fma(w[i], x[i], acc) : acc + w[i];
        }
    }
    acc = warp_tree(acc);
    if (lane == 0) partials[(size_t)q * P + blockIdx.x] = acc;
}

// pass 2: csum[a(a+1)/2 + b] = sum_i (w_i (x_ia - mean_a)) (x_ib - mean_b), b <= a.
// Canonical order per entry: inside a chunk of M2_CH = 512 consecutive particles lane l accumulates particles
// l, l + 32, ... sequentially with fma, then the adjacent-pair tree over the 32 lanes, then over chunks.
// Mapping: one block per chunk.  The chunk's d + 1 columns (86 KB at d = 20) are brought into shared memory by
// asynchronous copies that are all in flight at once (LDGSTS, no registers; two resident blocks per SM keep
// ~170 KB outstanding, which is what HBM needs), then lane = particle (conflict-free column reads) and warp g owns
// the matrix rows [rowb(g), rowb(g+1)) of the lower triangle (balanced entry counts: 55/50/48/57 at d = 20) with
// its entries in registers: 210 DFMA per particle against 21 x 4 shared loads.
constexpr int M2_G = 4;                       // row groups (warps) per block
constexpr int M2_R = M2_CH / 32;              // sequential depth per lane
template <int D>
__host__ __device__ constexpr int m2_rowb(int g)
{
    // smallest a with a(a+1)/2 >= g * E / G : balances the number of entries per group
    int a = 0;
    while (a < D && (a * (a + 1)) / 2 * M2_G < g * (D * (D + 1) / 2)) ++a;
    return g >= M2_G ? D : a;
}

__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gmem_src)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

template <int D, int GRP>
__device__ __forceinline__ void m2_group(const double* __restrict__ xs /* smem [D+1][M2_CH] */, int64_t N, int64_t c0, int64_t chunk,
                                         int lane, const double* __restrict__ mean, double* __restrict__ red)
{
    constexpr int A0 = m2_rowb<D>(GRP), A1 = m2_rowb<D>(GRP + 1);
    if (A1 <= A0) { __syncthreads(); return; }     // (keeps the block barrier below matched)
    constexpr int NR = (A1 > A0) ? (A1 - A0) : 1;
    double acc[NR][A1];
#pragma unroll
    for (int a = 0; a < NR; ++a)
#pragma unroll
        for (int b = 0; b < A1; ++b) acc[a][b] = 0.0;
    double mu[A1];
#pragma unroll
    for (int k = 0; k < A1; ++k) mu[k] = mean[k];
#pragma unroll 2
    for (int r = 0; r < M2_R; ++r) {
        const int64_t i = c0 + (int64_t)r * 32 + lane;
        if (i < N) {
            double dx[A1];
#pragma unroll
            for (int k = 0; k < A1; ++k) dx[k] = xs[k * M2_CH + r * 32 + lane] - mu[k];
            const double wi = xs[D * M2_CH + r * 32 + lane];
#pragma unroll
            for (int a = A0; a < A1; ++a) {
                const double wa = wi * dx[a];
#pragma unroll
                for (int b = 0; b <= a; ++b) acc[a - A0][b] = fma(wa, dx[b], acc[a - A0][b]);
            }
        }
    }
    // per-lane sums -> shared memory rows [entry][lane] (row stride 33: conflict-free for the row-wise tree below);
    // the rows alias the staged columns, so every warp of the block must have finished reading them
    __syncthreads();
#pragma unroll
    for (int a = A0; a < A1; ++a)
#pragma unroll
        for (int b = 0; b <= a; ++b) red[(a * (a + 1) / 2 + b) * 33 + lane] = acc[a - A0][b];
}

template <int D>
__global__ void __launch_bounds__(32 * M2_G)
k_moments2(const double* __restrict__ cloud, int64_t N, const double* __restrict__ msum,
           double* __restrict__ partials, int P)
{
    extern __shared__ double sm_m2[];          // [D + 1][M2_CH]: parameter columns, then the weight column
    __shared__ double mean[D];
    const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
    const int64_t chunk = blockIdx.x;
    const int64_t c0 = chunk * M2_CH;
    constexpr int NT = 32 * M2_G, PER = M2_CH / NT;
    if (c0 + M2_CH <= N) {                     // full chunk: (D + 1) * PER copies per thread, pointer stepping only
        const double* src = cloud + c0 + threadIdx.x;
        double* dst = sm_m2 + threadIdx.x;
#pragma unroll
        for (int k = 0; k <= D; ++k) {
            const double* col = (k < D) ? src + col_off(N, k) : src + col_off(N, D + 4);
#pragma unroll
            for (int j = 0; j < PER; ++j) cp_async8(dst + k * M2_CH + j * NT, col + j * NT);
        }
    } else {
        for (int idx = threadIdx.x; idx < (D + 1) * M2_CH; idx += NT) {
            const int k = idx / M2_CH, p = idx % M2_CH;
            const int64_t i = c0 + p;
            if (i < N) cp_async8(sm_m2 + idx, cloud + col_off(N, k < D ? k : D + 4) + i);
            else sm_m2[idx] = 0.0;
        }
    }
    if (threadIdx.x < D) mean[threadIdx.x] = msum[1 + threadIdx.x] / msum[0];
    cp_async_wait_all();
    __syncthreads();
    constexpr int E = D * (D + 1) / 2;
    double* red = sm_m2;                       // [E][33] per-lane sums, aliasing the staged columns once every warp is done
    switch (g) {     // warp-uniform: each group is compiled with static row bounds (register-resident accumulators)
    case 0: m2_group<D, 0>(sm_m2, N, c0, chunk, lane, mean, red); break;
    case 1: m2_group<D, 1>(sm_m2, N, c0, chunk, lane, mean, red); break;
    case 2: m2_group<D, 2>(sm_m2, N, c0, chunk, lane, mean, red); break;
    default: m2_group<D, 3>(sm_m2, N, c0, chunk, lane, mean, red); break;
    }
    __syncthreads();
    // adjacent-pair tree over the 32 lanes of every entry, one thread per entry (registers)
    for (int e = threadIdx.x; e < E; e += NT) {
        double v[32];
#pragma unroll
        for (int l = 0; l < 32; ++l) v[l] = red[e * 33 + l];
#pragma unroll
        for (int sft = 1; sft < 32; sft <<= 1)
#pragma unroll
            for (int l = 0; l < 32; l += 2 * sft) v[l] = v[l] + v[l + sft];
        partials[(size_t)e * P + chunk] = v[0];
    }
}

// generic-d fallback of pass 2 (one block of 4 warps per chunk, warps over entries; any d <= DMAX); same order
__global__ void __launch_bounds__(128)
k_moments2_generic(const double* __restrict__ cloud, int64_t N, int d, const double* __restrict__ msum,
                   double* __restrict__ partials, int P)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t chunk = blockIdx.x;
    const int64_t c0 = chunk * M2_CH;
    const double* w = cloud + col_off(N, d + 4);
    const double sw = msum[0];
    const int E = d * (d + 1) / 2;
    for (int e = warp; e < E; e += 4) {
        int a = 0;
        while ((a + 1) * (a + 2) / 2 <= e) ++a;
        const int b = e - a * (a + 1) / 2;
        const double ma = msum[1 + a] / sw, mb = msum[1 + b] / sw;
        const double* xa = cloud + col_off(N, a);
        const double* xb = cloud + col_off(N, b);
        double acc = 0.0;
        for (int r = 0; r < M2_R; ++r) {
            const int64_t i = c0 + (int64_t)r * 32 + lane;
            if (i < N) acc = fma(w[i] * (xa[i] - ma), xb[i] - mb, acc);
        }
        acc = warp_tree(acc);
        if (lane == 0) partials[(size_t)e * P + chunk] = acc;
    }
}

// one block per quantity: adjacent-pair tree over its tile partials
__global__ void __launch_bounds__(256) k_tree_finalize(double* partials, int ntiles, int P, double* out)
{
    __shared__ double sm[8];
    const double r = tiles_tree_256(partials + (size_t)blockIdx.x * P, ntiles, P, sm);
    if (threadIdx.x == 0) out[blockIdx.x] = r;
}

// =================================================================================================
// proposal preparation: mean, covariance, per-block Cholesky -> MutConst (staged in global memory)
// =================================================================================================
// Shared by host (explicit mean/cov through the C ABI) and device (fused stage).  cov: d x d full.
SMC_HD int build_mutconst(const double* mean, const double* cov, int d, const BlockSpec& bs, double c, MutConst* out,
                          double* work /* 2*DMAX*DMAX */)
{
    out->n_blocks = bs.n_blocks;
    out->status = 0;
    for (int k = 0; k < DMAX; ++k) out->mu[k] = (k < d) ? mean[k] : 0.0;
    double* S = work;
    double* L = work + DMAX * DMAX;
    for (int b = 0; b < bs.n_blocks; ++b) {
        const int n = bs.bsize[b];
        out->bsize[b] = n;
        uint32_t mask = 0;
        for (int e = 0; e < PACKMAX; ++e) out->L[b][e] = 0.0;
        for (int k = 0; k < DMAX; ++k) { out->csd[b][k] = 0.0; out->isd[b][k] = 0.0; out->isdn[b][k] = 0.0; out->rl[b][k] = 0.0; }
        for (int i = 0; i < n; ++i) {
            mask |= 1u << bs.member[b][i];
            for (int j = 0; j < n; ++j) {
                // R_fr = (R[f,f] + R[f,f]') / 2, src/smc_main.jl:462
                const int ai = bs.member[b][i], aj = bs.member[b][j];
                S[i * n + j] = (cov[ai * d + aj] + cov[aj * d + ai]) / 2.0;
            }
        }
        out->mask[b] = mask;
        const int st = cholesky_lower(S, n, L);
        if (st) { out->status = SMCB200_ERR_NOT_POSDEF; return SMCB200_ERR_NOT_POSDEF; }
        for (int i = 0; i < n; ++i) {
            const int ai = bs.member[b][i];
            for (int j = 0; j <= i; ++j) {
                const int aj = bs.member[b][j];
                out->L[b][aj * d - (aj * (aj - 1)) / 2 + (ai - aj)] = c * L[i * n + j];
            }
            out->csd[b][ai] = c * sqrt(S[i * n + i]);
            const double isd = 1.0 / sqrt(S[i * n + i]);
            out->isd[b][ai] = isd;
            out->isdn[b][ai] = isd * 0x1.9884533d43651p-2;
            out->rl[b][ai] = 1.0 / (c * L[i * n + i]);
        }
        // log-normaliser of N(.; ., c^2 Sigma_b): n log(2 pi) + 2 sum_i log(c L_ii)  (members ascending)
        double ld = 0.0;
        for (int i = 0; i < n; ++i) ld = ld + det_log(c * L[i * n + i]);
        out->lognorm[b] = (double)n * (2.0 * 0.91893853320467274178) + 2.0 * ld;
    }
    return 0;
}

// debug: elementwise deterministic math on the device
__global__ void k_debug_math(int op, const double* __restrict__ x, int64_t n, uint64_t seed, double* __restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s, c;
    switch (op) {
    case 0: out[i] = det_exp(x[i]); break;
    case 1: out[i] = det_log(x[i]); break;
    case 2: det_sincos2pi(x[i], s, c); out[i] = s; break;
    case 3: det_sincos2pi(x[i], s, c); out[i] = c; break;
    case 4: case 5: {
        double z0, z1;
        normal_pair(rng4(seed, (uint32_t)i, 0u, (uint32_t)x[i], PURP_NORMAL), z0, z1);
        out[i] = (op == 4) ? z0 : z1;
        break;
    }
    default: {   // 6..9: the four proposal normals of normal_quad
        double z[4];
        normal_quad(rng4(seed, (uint32_t)i, 0u, (uint32_t)x[i], PURP_NORMAL), reinterpret_cast<const float4*>(normal_tab_dev), z[0], z[1], z[2], z[3]);
        out[i] = z[(op - 6) & 3];
    }
    }
}

}  // namespace smc
