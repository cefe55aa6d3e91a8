// reduce.cuh -- the canonical reduction trees of DESIGN.md "Numerical contract" (mirrored by the oracle): adjacent-pair
// binary trees over lanes, then over tile partials that are zero padded to a power of two.
#pragma once
#include "common.cuh"

namespace smc {

// Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-serialization attribute may start
// while its predecessor in the stream drains; pdl_wait() blocks until that predecessor has completed and its writes are
// visible (a no-op for a kernel launched the ordinary way), pdl_trigger() lets the successor's blocks be scheduled early.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// =================================================================================================
// canonical 256-lane block tree + last-block tile tree
// =================================================================================================
__device__ __forceinline__ double warp_tree(double v)
{
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) v = v + __shfl_xor_sync(0xffffffffu, v, s);
    return v;
}

// adjacent-pair tree over the 256 threads of the block; result valid in thread 0. sm: 8 doubles.
__device__ __forceinline__ double block_tree_256(double v, double* sm)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_tree(v);
    __syncthreads();               // protect sm reuse across calls
    if (lane == 0) sm[warp] = v;
    __syncthreads();
    if (warp == 0) {
        double x = (lane < 8) ? sm[lane] : 0.0;
        x = x + __shfl_xor_sync(0xffffffffu, x, 1);
        x = x + __shfl_xor_sync(0xffffffffu, x, 2);
        x = x + __shfl_xor_sync(0xffffffffu, x, 4);
        v = x;
    }
    return v;
}

// adjacent-pair tree over `ntiles` tile partials (zero padded to the power of two P) by one block
// of 256 threads; result valid in thread 0.  part[] entries >= ntiles must be zero (they are never
// written after the initial memset).
__device__ __forceinline__ double tiles_tree_256(double* part, int ntiles, int P, double* sm)
{
    double x;
    if (P <= 256) {
        x = ((int)threadIdx.x < ntiles) ? __ldcg(part + threadIdx.x) : 0.0;
    } else {
        const int m = P / 256;
        double* p = part + (size_t)threadIdx.x * m;
        for (int s = 1; s < m; s <<= 1)
            for (int i = 0; i < m; i += 2 * s) __stcg(p + i, __ldcg(p + i) + __ldcg(p + i + s));
        x = __ldcg(p);
    }
    return block_tree_256(x, sm);
}

// The block that finishes last reduces the tile partials of NQ quantities into out[q].
// v[q] must be valid in thread 0.  Returns true (block-uniform) in the finishing block.
template <int NQ>
__device__ __forceinline__ bool finish_tiles(const double (&v)[NQ], double* partials, int ntiles, int P,
                                             unsigned* counter, double* out, double* sm)
{
    __shared__ bool is_last;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) __stcg(partials + (size_t)q * P + blockIdx.x, v[q]);
        __threadfence();
        const unsigned t = atomicInc(counter, gridDim.x - 1);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return false;
    __threadfence();
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        const double r = tiles_tree_256(partials + (size_t)q * P, ntiles, P, sm);
        if (threadIdx.x == 0) out[q] = r;
    }
    return true;
}

// adjacent-pair tree of 8 values
__device__ __forceinline__ double tree8(const double (&v)[8])
{
    return ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
}
// Adjacent-pair tree over m consecutive partials (m a power of two) held by ONE thread: batches of eight independent
// loads, an in-register tree per batch, and a binary-counter stack over the batches (which reproduces the tree).
__device__ __forceinline__ double tree_run(const double* p, int m)
{
    double stack[28];
    int sp = 0;
    for (int i0 = 0, bi = 0; i0 < m; i0 += 8, ++bi) {
        double v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = (i0 + k < m) ? __ldcg(p + i0 + k) : 0.0;
        double x = tree8(v);
        int k = bi;
        while (k & 1) { x = stack[--sp] + x; k >>= 1; }
        stack[sp++] = x;
    }
    return stack[0];
}
// Adjacent-pair tree over P tile partials (P a power of two; entries past the last tile are zero) by one block of T
// threads (T a power of two <= 1024); result valid in thread 0.  sm: T / 32 doubles.  Read-only on `part`.
template <int T>
__device__ __forceinline__ double tiles_tree_block(const double* part, int P, double* sm)
{
    double x;
    if (P <= T) x = ((int)threadIdx.x < P) ? __ldcg(part + threadIdx.x) : 0.0;
    else x = tree_run(part + (size_t)threadIdx.x * (P / T), P / T);
    x = warp_tree(x);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = x;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x == 0) {
        constexpr int NW = T / 32;
        double t[NW];
#pragma unroll
        for (int k = 0; k < NW; ++k) t[k] = sm[k];
#pragma unroll
        for (int sft = 1; sft < NW; sft <<= 1)
#pragma unroll
            for (int k = 0; k + sft < NW; k += 2 * sft) t[k] = t[k] + t[k + sft];
        r = t[0];
    }
    return r;
}
// Same tree by ONE warp (lane 0 holds the result): each lane takes P / 32 consecutive partials.
__device__ __forceinline__ double tiles_tree_warp(const double* part, int P)
{
    const int lane = threadIdx.x & 31;
    double x;
    if (P <= 32) x = (lane < P) ? __ldcg(part + lane) : 0.0;
    else x = tree_run(part + (size_t)lane * (P / 32), P / 32);
    return warp_tree(x);
}

// =================================================================================================
// Small cross-GPU reductions without a collective library, callable from inside any kernel.  Every rank pushes its nq
// shard-local roots straight into every peer's inbox over NVLink (plain remote stores through CUDA-IPC mappings),
// publishes an epoch flag behind a system-scope fence, waits for the peers' flags in its own inbox and combines the
// world x nq values in the fixed rank-order tree (combine != 0) or lays them out as [world][nq] (gather).
// Inbox slots alternate with the epoch's parity: a rank can only reach epoch e + 1 after it has seen every peer's
// epoch-e flag, and a peer publishes e only after it has consumed e - 1, so two slots never collide.
// A peer that never arrives (a crashed rank) trips a clock-based time-out instead of hanging the GPU.
// =================================================================================================
// Small exchanges (nq <= LL_NQ_B values, i.e. everything but the moment sums of wide models) use the low-latency form:
// every double travels as two 8-byte words whose upper halves carry the exchange's epoch tag, so a word is either absent
// or complete -- no system fence, no separate flag, one NVLink write latency per exchange.
constexpr int LL_NQ_A = 32;                      // region A: the cooperative correction kernel's own sequence (stage_kernels.cuh)
constexpr int LL_NQ_B = 256;                     // region B: exchanges numbered by pc.epoch
__device__ __forceinline__ unsigned long long* ll_region_a(double* inbox, int world)
{
    return reinterpret_cast<unsigned long long*>(inbox) + (size_t)2 * world * MB_NQ + (size_t)2 * world;
}
__device__ __forceinline__ unsigned long long* ll_region_b(double* inbox, int world)
{
    return ll_region_a(inbox, world) + (size_t)2 * 16 * LL_NQ_A * 2;
}
constexpr size_t LL_WORDS = (size_t)2 * 16 * (LL_NQ_A + LL_NQ_B) * 2;

// Called by ALL threads of ONE block (blockDim.x >= world).  local_src / dst: global or shared memory.
__device__ __forceinline__ void peer_exchange_block(const PeerCtx& pc, const double* local_src, int nq, int combine, double* dst)
{
    __shared__ unsigned long long s_epoch;
    __syncthreads();
    if (threadIdx.x == 0) { s_epoch = *pc.epoch + 1ull; *pc.epoch = s_epoch; }
    __syncthreads();
    const unsigned long long epoch = s_epoch;
    const int rank = pc.rank, world = pc.world;
    const int par = (int)(epoch & 1ull);
    if (nq <= LL_NQ_B) {
        const unsigned long long tag = (epoch & 0xffffffffull) << 32;
        for (int e = threadIdx.x; e < world * nq; e += blockDim.x) {
            const int r = e / nq, q = e % nq;
            const unsigned long long bits = (unsigned long long)__double_as_longlong(local_src[q]);
            volatile unsigned long long* w = ll_region_b(pc.inbox[r], world) + ((size_t)(par * 16 + rank) * LL_NQ_B + q) * 2;
            w[0] = (bits & 0xffffffffull) | tag;
            w[1] = (bits >> 32) | tag;
        }
        for (int q = threadIdx.x; q < nq; q += blockDim.x) {
            double v[16];
            for (int r = 0; r < world; ++r) {
                volatile unsigned long long* w = ll_region_b(pc.inbox[rank], world) + ((size_t)(par * 16 + r) * LL_NQ_B + q) * 2;
                unsigned long long w0 = w[0], w1 = w[1];
                const long long t0 = clock64();
                while ((w0 & 0xffffffff00000000ull) != tag || (w1 & 0xffffffff00000000ull) != tag) {
                    if (clock64() - t0 > 240000000000ll) { *pc.err = 1; break; }    // ~2 minutes: a peer died; the host reports it
                    w0 = w[0]; w1 = w[1];
                }
                v[r] = __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
            }
            if (combine) {
                for (int s = 1; s < world; s <<= 1)
                    for (int i = 0; i + s < world; i += 2 * s) v[i] = v[i] + v[i + s];
                dst[q] = v[0];
            } else {
                for (int r = 0; r < world; ++r) dst[(size_t)r * nq + q] = v[r];
            }
        }
        __syncthreads();
        return;
    }
    const size_t flag_off = (size_t)2 * world * MB_NQ;               // flags follow the value slots (as doubles' worth of u64)
    for (int r = 0; r < world; ++r) {
        double* slot = pc.inbox[r] + ((size_t)par * world + rank) * MB_NQ;
        for (int q = threadIdx.x; q < nq; q += blockDim.x) slot[q] = local_src[q];
    }
    __syncthreads();                     // the block's remote stores happen-before the (cumulative) fences below
    if ((int)threadIdx.x < world) {
        __threadfence_system();
        volatile unsigned long long* f = reinterpret_cast<volatile unsigned long long*>(pc.inbox[threadIdx.x] + flag_off) + (size_t)par * world + rank;
        *f = epoch;
        volatile unsigned long long* g = reinterpret_cast<volatile unsigned long long*>(pc.inbox[rank] + flag_off) + (size_t)par * world + threadIdx.x;
        const long long t0 = clock64();
        while (*g < epoch) {
            if (clock64() - t0 > 240000000000ll) { *pc.err = 1; break; }    // ~2 minutes: a peer died; the host reports it
        }
        __threadfence_system();
    }
    __syncthreads();
    const double* mine = pc.inbox[rank] + (size_t)par * world * MB_NQ;
    for (int q = threadIdx.x; q < nq; q += blockDim.x) {
        if (combine) {
            double v[16];
            for (int r = 0; r < world; ++r) v[r] = __ldcg(mine + (size_t)r * MB_NQ + q);
            for (int s = 1; s < world; s <<= 1)
                for (int i = 0; i + s < world; i += 2 * s) v[i] = v[i] + v[i + s];
            dst[q] = v[0];
        } else {
            for (int r = 0; r < world; ++r) dst[(size_t)r * nq + q] = __ldcg(mine + (size_t)r * MB_NQ + q);
        }
    }
    __syncthreads();
}

}  // namespace smc
