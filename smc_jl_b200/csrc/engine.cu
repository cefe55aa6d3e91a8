// engine.cu -- the C ABI of libsmcb200 (include/smcb200.h) and the host-side orchestration of one
// SMC stage on one GPU shard.  Host code here only sequences kernels and moves scalars; all
// per-particle arithmetic runs in the kernels of kernels.cuh / mutate_kernel.cuh (instantiated in mut_*.cu).
#include <dlfcn.h>
#include <nccl.h>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>

#include "stage_kernels.cuh"

using namespace smc;

struct smcb200_ctx : public smc::Ctx {};

namespace {

constexpr double HALF_LOG_2PI = 0.91893853320467274178;

int fail(Ctx* c, int code, const char* msg)
{
    c->err = msg;
    return code;
}

// ---- NCCL, resolved at run time (the process may already hold torch's bundled libnccl) ----------------
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi* nccl_api()
{
    static NcclApi api;
    static bool tried = false;
    if (tried) return api.lib ? &api : nullptr;
    tried = true;
    const char* names[] = {"libnccl.so.2", "libnccl.so", "/usr/lib/x86_64-linux-gnu/libnccl.so.2"};
    for (const char* n : names) {
        api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.lib) break;
    }
    if (!api.lib) return nullptr;
    api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.lib, "ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.lib, "ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.lib, "ncclCommDestroy");
    api.AllGather = (decltype(api.AllGather))dlsym(api.lib, "ncclAllGather");
    api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.lib, "ncclGetErrorString");
    if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllGather) { api.lib = nullptr; return nullptr; }
    return &api;
}
#define SMC_NCCL(ctx, call)                                                                        \
    do {                                                                                           \
        ncclResult_t r__ = (call);                                                                 \
        if (r__ != ncclSuccess) {                                                                  \
            NcclApi* a__ = nccl_api();                                                             \
            (ctx)->err = std::string(#call) + ": " + ((a__ && a__->GetErrorString) ? a__->GetErrorString(r__) : "nccl error"); \
            return SMCB200_ERR_NCCL;                                                               \
        }                                                                                          \
    } while (0)

constexpr int SL_MSUM = 16, SL_CSUM = 64, SL_SCAN_ROOT = 600, SL_ESS = 604, SL_COUNT = 640;   // slots of scal_loc
static_assert(SL_ESS + 2 * ESS_K <= SL_COUNT, "scal_loc too small");
// where a kernel should write shard-local roots, and the cross-rank tree that follows
// cross-rank tree of nq shard-local roots (combine) or their [world][nq] layout (gather into dst): one
// k_peer_exchange launch over the NVLink mailboxes; `flag` (device, nullable) predicates the launch
// Launch with the programmatic-serialization attribute: the kernel (which starts with pdl_wait()) may be scheduled while
// its predecessor in the stream drains, which hides the launch latency of the stage's chain of small dependent kernels.
template <class... KArgs, class... Args>
cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

int peer_exchange(Ctx* c, double* dst, const double* local_src, int nq, int combine, const double* flag = nullptr)
{
    if (nq > MB_NQ) { c->err = "peer_exchange: too many quantities"; return SMCB200_ERR_UNSUPPORTED; }
    SMC_CUDA(c, launch_pdl(k_peer_exchange, dim3(1), dim3(256), 0, c->stream, local_src, nq, peer_ctx(c), combine, dst, flag));
    c->launches += 1;
    SMC_CUDA(c, cudaGetLastError());
    return SMCB200_OK;
}
int reduce_ranks(Ctx* c, double* final_dst, const double* local_src, int nq)
{
    if (c->world == 1) return SMCB200_OK;
    return peer_exchange(c, final_dst, local_src, nq, 1);
}

void free_cloud(Ctx* c)
{
    cudaFree(c->cloud[0]); cudaFree(c->cloud[1]); cudaFree(c->tmp); cudaFree(c->rmax); cudaFree(c->idx);
    cudaFree(c->partials); cudaFree(c->mpartials); cudaFree(c->ess_partials); c->ess_partials = nullptr; cudaFree(c->scan_blocktot); cudaFree(c->scan_blockoff); cudaFree(c->scan_levels);
    cudaFree(c->scan_bmax); cudaFree(c->msum); cudaFree(c->csum);
    cudaFree(c->hist_scr); cudaFree(c->m1p_partials); cudaFree(c->m1p_sums); cudaFree(c->coop_partials);
    c->hist_scr = c->m1p_partials = c->m1p_sums = c->coop_partials = nullptr;
    for (int b = 0; b < 3; ++b)
        for (int r = 0; r < 16; ++r)
            if (c->ipc_open[b][r]) { cudaIpcCloseMemHandle(c->ipc_open[b][r]); c->ipc_open[b][r] = nullptr; }
    cudaFree(c->bmax_g); cudaFree(c->peer_tab); cudaFree(c->peer_cnt); cudaFree(c->rmax_tab);
    c->bmax_g = nullptr; c->peer_tab = nullptr; c->peer_cnt = nullptr; c->rmax_tab = nullptr;
    c->cloud[0] = c->cloud[1] = c->tmp = c->rmax = nullptr; c->idx = nullptr; c->partials = c->mpartials = nullptr;
    c->scan_blocktot = c->scan_blockoff = c->scan_levels = c->scan_bmax = nullptr; c->msum = c->csum = nullptr;
    c->scan_nb_cap = 0;
    c->N = c->N_global = 0;
}

// tile geometry of the canonical orders for this shard
struct Tiles { int ntiles, P; };
Tiles weight_tiles(int64_t n) { int nt = (int)((n + W_TILE - 1) / W_TILE); if (nt < 1) nt = 1; return {nt, (int)next_pow2(nt)}; }
Tiles moment_tiles(int64_t n) { int nt = (int)((n + M_TILE - 1) / M_TILE); if (nt < 1) nt = 1; return {nt, (int)next_pow2(nt)}; }
Tiles chunk_tiles(int64_t n) { int nt = (int)((n + M2_CH - 1) / M2_CH); if (nt < 1) nt = 1; return {nt, (int)next_pow2(nt)}; }
Tiles unit_tiles(int64_t n) { const int64_t u = (int64_t)M1P_SC * M1P_WARPS; int nt = (int)((n + u - 1) / u); if (nt < 1) nt = 1; return {nt, (int)next_pow2(nt)}; }   // blocks of the one-pass moments

// trial phi per sweep of the adaptive solve: a sweep costs one grid barrier + K exp per particle, so small clouds take
// wider sweeps (the result does not depend on K)
int coop_k(int64_t n) { return (n >= ((int64_t)1 << 19)) ? 3 : 7; }     // per-sweep cost ~ 6 us + K x 3.4 us x n / 2^20 (B200): 27 sweeps of 3 beat 18 of 7 from 2^19 up
// variant index: 0 = K 3, 1 = K 7, 2 = fixed schedule (no solve code)
int coop_variant(int64_t n, bool adaptive) { return !adaptive ? 2 : (coop_k(n) == 3 ? 0 : 1); }
const void* coop_kernel(int v) { return v == 0 ? (const void*)k_correct_coop<3> : v == 1 ? (const void*)k_correct_coop<7> : (const void*)k_correct_coop<0>; }
// cooperative grid of k_correct_coop for this shard: at most one block per SM (a short grid barrier), never more blocks
// than pairs of tiles
int coop_grid(const Ctx* c, int64_t n, int variant)
{
    const Tiles t = weight_tiles(n);
    int g = (t.ntiles + COOP_GROUPS - 1) / COOP_GROUPS;
    int per_sm = c->coop_blocks_per_sm[variant];
    if (per_sm > 2) per_sm = 2;                       // a short grid barrier matters more than a third block
    const int cap = (c->sm_count > 0 ? c->sm_count : 1) * (per_sm > 0 ? per_sm : 1);
    return g > cap ? cap : g;
}

int sync(Ctx* c)
{
    SMC_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->world > 1) {           // a peer that never reached an exchange trips the device-side time-out
        int e = 0;
        SMC_CUDA(c, cudaMemcpy(&e, c->mb_err, sizeof(int), cudaMemcpyDeviceToHost));
        if (e) { c->err = "a peer rank never reached a cross-GPU reduction (time-out)"; return SMCB200_ERR_NCCL; }
    }
    return SMCB200_OK;
}

// ---- correction (+ adaptive phi) --------------------------------------------------------------------
struct CorrectLaunch {
    double phi_n1 = 0, phi_n = 0, pw = 0, lpod = 0, threshold_ratio = 0;
    int adaptive = 0, solve_only = 0, use_carry = 0;
    const double* sched_dev = nullptr; int n_phi = 0; double tempering_target = 0;
    double c_in = 0, accept_in = 0, ess_prev_in = 0, phi_prop_in = 0; long long j_in = 0; int resampled_last_in = 0;
    double* inc_dev = nullptr; double* normw_dev = nullptr;
};
int launch_correct_coop(Ctx* c, const CorrectLaunch& L)
{
    const int d = c->d;
    double* cl = c->cloud[c->cur];
    const Tiles tw = weight_tiles(c->N);
    CoopArgs a;
    std::memset(&a, 0, sizeof(a));
    a.ll = cl + col_off(c->N, d); a.old = cl + col_off(c->N, d + 2); a.w = cl + col_off(c->N, d + 4);
    a.inc_out = L.inc_dev; a.normw_out = L.normw_dev;
    a.N = c->N; a.n_global = (double)c->N_global;
    a.ntiles = tw.ntiles; a.P = tw.P;
    a.corr.phi_n1 = L.phi_n1; a.corr.phi_n = L.phi_n; a.corr.pw = L.pw; a.corr.lpod = L.lpod;
    a.corr.mode = (L.pw == 0.0) ? 0 : (L.pw == 1.0 ? 1 : 2);
    a.corr.log_1m_pw = (a.corr.mode == 2) ? det_log(1.0 - L.pw) : 0.0;
    a.threshold_ratio = L.threshold_ratio;
    a.adaptive = L.adaptive; a.solve_only = L.solve_only; a.use_carry = L.use_carry;
    a.sched = L.sched_dev; a.n_phi = L.n_phi; a.tempering_target = L.tempering_target;
    a.c_in = L.c_in; a.accept_in = L.accept_in; a.ess_prev_in = L.ess_prev_in; a.phi_prop_in = L.phi_prop_in; a.j_in = L.j_in;
    a.resampled_last_in = L.resampled_last_in;
    a.partials = c->coop_partials; a.scal = c->scal; a.pc = peer_ctx(c);
    a.ticket = c->coop_ticket; a.ll_step = c->coop_seq;
    a.ll_always = c->coop_ll_single;
    void* args[] = {&a};
    const int variant = coop_variant(c->N, L.adaptive != 0);
    SMC_CUDA(c, cudaLaunchCooperativeKernel(coop_kernel(variant), dim3(coop_grid(c, c->N, variant)), dim3(COOP_NT), args, 0, c->stream));
    c->launches += 1;
    return SMCB200_OK;
}

int upload_schedule(Ctx* c, const double* sched, int n_phi)
{
    if (c->sched_cap < n_phi) {
        cudaFree(c->sched_dev);
        c->sched_dev = nullptr; c->sched_cap = 0;
        SMC_CUDA(c, cudaMalloc(&c->sched_dev, sizeof(double) * n_phi));
        c->sched_cap = n_phi;
    }
    SMC_CUDA(c, cudaMemcpyAsync(c->sched_dev, sched, sizeof(double) * n_phi, cudaMemcpyHostToDevice, c->stream));
    return SMCB200_OK;
}

// one multi-trial pass of compute_ESS: S_k, Q_k for the ESS_K trial phi in `trials` (device) -> c->ess_sq
int launch_ess_multi(Ctx* c, double phi_n1, const double* trials_dev)
{
    const int d = c->d;
    double* cl = c->cloud[c->cur];
    const Tiles t = weight_tiles(c->N);
    k_ess_multi<<<t.ntiles, 256, 0, c->stream>>>(cl + col_off(c->N, d), cl + col_off(c->N, d + 2), cl + col_off(c->N, d + 4), c->N,
                                                 phi_n1, trials_dev, nullptr, c->ess_partials, t.P);
    k_tree_finalize<<<2 * ESS_K, 256, 0, c->stream>>>(c->ess_partials, t.ntiles, t.P, c->world > 1 ? c->scal_loc + SL_ESS : c->ess_sq);
    c->launches += 2;
    SMC_CUDA(c, cudaGetLastError());
    return reduce_ranks(c, c->ess_sq, c->scal_loc + SL_ESS, 2 * ESS_K);
}

// ---- resampling on explicit buffers -----------------------------------------------------------------
struct ScanGeom { int64_t P; int B, nb; };
ScanGeom scan_geom(int64_t n)
{
    ScanGeom g;
    g.P = next_pow2(n < LEAF ? LEAF : n);
    g.B = (g.P < SCAN_TILE) ? (int)g.P : SCAN_TILE;
    g.nb = (int)(g.P / g.B);
    return g;
}
constexpr size_t SCAN_SMEM = (size_t)(SCAN_THREADS * (LEAF + 1) + 2 * SCAN_THREADS + SCAN_THREADS) * sizeof(double);

int ensure_scan_buffers(Ctx* c, int nb)
{
    if (nb <= c->scan_nb_cap) return SMCB200_OK;
    cudaFree(c->scan_blocktot); cudaFree(c->scan_blockoff); cudaFree(c->scan_levels); cudaFree(c->scan_bmax);
    c->scan_blocktot = c->scan_blockoff = c->scan_levels = c->scan_bmax = nullptr; c->scan_nb_cap = 0;
    SMC_CUDA(c, cudaMalloc(&c->scan_blocktot, sizeof(double) * nb));
    SMC_CUDA(c, cudaMalloc(&c->scan_blockoff, sizeof(double) * nb));
    SMC_CUDA(c, cudaMalloc(&c->scan_levels, sizeof(double) * 2 * nb));
    SMC_CUDA(c, cudaMalloc(&c->scan_bmax, sizeof(double) * nb));
    c->scan_nb_cap = nb;
    return SMCB200_OK;
}

double systematic_offset(uint64_t seed, uint32_t stage, double u_override)
{
    if (u_override >= 0.0) return u_override;
    const u32x4 r = rng4(seed, 0u, stage, 0u, PURP_RESAMPLE);
    return u01(r.x, r.y);
}

// src: n weights on the device (div_n: use src/n_parts, the `normalized_weights/n_parts` of smc_main.jl:438);
// rmax/craw/idx: device outputs.  `sres` = device slot holding sum(weights) ("weights ./ sum(weights)"); it is computed
// here unless have_sres.  Single-shard version (also serves the exported resample(weights) on host vectors).
// flag (device, nullable): every kernel returns at once when *flag == 0 (device-side resample decision).
int launch_resample_indices(Ctx* c, const double* src, int div_n, int64_t n, int method, uint64_t seed, uint32_t stage,
                            double u_override, double* rmax, double* craw, int64_t* idx, double* partials,
                            unsigned* counter, double* sres, bool have_sres, const double* flag, int64_t n_out = -1)
{
    if (n_out < 0) n_out = n;
    if (method != SMCB200_RESAMPLE_SYSTEMATIC && method != SMCB200_RESAMPLE_MULTINOMIAL)
        return fail(c, SMCB200_ERR_BAD_RESAMPLER, "Invalid resampler in SMC. Options are :systematic or :multinomial");
    const ScanGeom g = scan_geom(n);
    if (g.nb > (1 << 16)) return fail(c, SMCB200_ERR_UNSUPPORTED, "n_parts too large for the device scan (max 2^28)");
    int st = ensure_scan_buffers(c, g.nb);
    if (st) return st;
    const Tiles t = weight_tiles(n);
    const double nd = (double)n;
    if (!have_sres) {
        k_colsum<<<t.ntiles, 256, 0, c->stream>>>(src, n, nd, div_n, partials, t.ntiles, t.P, counter, sres, flag);
        c->launches += 1;
    }
    const double* nul = nullptr;
    double* nulw = nullptr;
    const double* const* nultab = nullptr;
    SMC_CUDA(c, launch_pdl(k_scan<false>, dim3(g.nb), dim3(SCAN_THREADS), SCAN_SMEM, c->stream, src, div_n, nd, (const double*)sres, n, g.B,
                           c->scan_blocktot, nul, nulw, nulw, nulw, flag));
    SMC_CUDA(c, launch_pdl(k_scan_upper_up, dim3(1), dim3(256), 0, c->stream, (const double*)c->scan_blocktot, g.nb, c->scan_levels,
                           c->scal_loc + SL_SCAN_ROOT, flag));
    SMC_CUDA(c, launch_pdl(k_scan_upper_down, dim3(1), dim3(256), 0, c->stream, (const double*)c->scan_levels, g.nb, nul, 1, 0,
                           c->scan_blockoff, flag));
    SMC_CUDA(c, launch_pdl(k_scan<true>, dim3(g.nb), dim3(SCAN_THREADS), SCAN_SMEM, c->stream, src, div_n, nd, (const double*)sres, n, g.B,
                           nulw, (const double*)c->scan_blockoff, rmax, craw, c->scan_bmax, flag));
    SMC_CUDA(c, launch_pdl(k_prefix_max, dim3(1), dim3(256), 0, c->stream, c->scan_bmax, g.nb, flag));
    const double u = systematic_offset(seed, stage, u_override);
    SMC_CUDA(c, launch_pdl(k_search, dim3((unsigned)((n_out + 255) / 256)), dim3(256), 0, c->stream, (const double*)rmax, nultab, 0,
                           (const double*)c->scan_bmax, g.nb, g.B, n, n_out, (int64_t)0, method, seed, stage, u, (double)n_out, idx, flag));
    c->launches += 6;
    SMC_CUDA(c, cudaGetLastError());
    return SMCB200_OK;
}

// Selection on the sharded cloud: canonical global cumsum (shard roots exchanged), this shard's running maxima stay in
// place; the per-block maxima of all shards are gathered through the mailboxes (a few hundred doubles), every rank
// searches its own outputs' ancestors in them and then in the OWNER's running-max block over NVLink, and the rows are
// pulled from the owners the same way (k_gather_peer).  No collective library call, so the whole chain can be
// predicated on the device-side resample flag.
int launch_resample_cloud_sharded(Ctx* c, int method, uint64_t seed, uint32_t stage, double u_override, bool fused, double* nw_hist)
{
    if (method != SMCB200_RESAMPLE_SYSTEMATIC && method != SMCB200_RESAMPLE_MULTINOMIAL)
        return fail(c, SMCB200_ERR_BAD_RESAMPLER, "Invalid resampler in SMC. Options are :systematic or :multinomial");
    const int d = c->d;
    double* cl = c->cloud[c->cur];
    const double* src = cl + col_off(c->N, d + 4);
    const double* flag = fused ? c->scal + SC_RESAMPLE : nullptr;
    const int B = SCAN_TILE;
    const int nb = (int)(c->per / B);                 // blocks of this shard's full subtree
    if (nb > MB_NQ) return fail(c, SMCB200_ERR_UNSUPPORTED, "multi-GPU selection supports up to 2^24 particles per GPU");
    int st = ensure_scan_buffers(c, nb); if (st) return st;
    const Tiles t = weight_tiles(c->N);
    const double nd = (double)c->N_global;
    if (!fused) {      // sum(weights) over all shards (the fused stage has it from the correction kernel)
        k_colsum<<<t.ntiles, 256, 0, c->stream>>>(src, c->N, nd, 1, c->partials + (size_t)3 * t.P, t.ntiles, t.P, c->counters + 2,
                                                  c->scal_loc + SC_SRES);
        c->launches += 1;
        st = reduce_ranks(c, c->scal + SC_SRES, c->scal_loc + SC_SRES, 1); if (st) return st;
    }
    const double* nul = nullptr;
    double* nulw = nullptr;
    SMC_CUDA(c, launch_pdl(k_scan<false>, dim3(nb), dim3(SCAN_THREADS), SCAN_SMEM, c->stream, src, 1, nd, (const double*)(c->scal + SC_SRES), c->N, B,
                           c->scan_blocktot, nul, nulw, nulw, nulw, flag));
    SMC_CUDA(c, launch_pdl(k_scan_upper_up, dim3(1), dim3(256), 0, c->stream, (const double*)c->scan_blocktot, nb, c->scan_levels,
                           c->scal_loc + SL_SCAN_ROOT, flag));
    st = peer_exchange(c, c->gath, c->scal_loc + SL_SCAN_ROOT, 1, 0, flag); if (st) return st;
    SMC_CUDA(c, launch_pdl(k_scan_upper_down, dim3(1), dim3(256), 0, c->stream, (const double*)c->scan_levels, nb, (const double*)c->gath, c->world,
                           c->rank, c->scan_blockoff, flag));
    SMC_CUDA(c, launch_pdl(k_scan<true>, dim3(nb), dim3(SCAN_THREADS), SCAN_SMEM, c->stream, src, 1, nd, (const double*)(c->scal + SC_SRES), c->N, B,
                           nulw, (const double*)c->scan_blockoff, c->rmax, nulw, c->scan_bmax, flag));
    SMC_CUDA(c, launch_pdl(k_prefix_max, dim3(1), dim3(256), 0, c->stream, c->scan_bmax, nb, flag));
    // block maxima of every shard -> [world][nb]; the maxima of the lower ranks are folded in afterwards
    st = peer_exchange(c, c->bmax_g, c->scan_bmax, nb, 0, flag); if (st) return st;
    SMC_CUDA(c, launch_pdl(k_fix_rank_carry, dim3(1), dim3(256), 0, c->stream, c->bmax_g, nb, c->world, flag));
    const double u = systematic_offset(seed, stage, u_override);
    SMC_CUDA(c, launch_pdl(k_search, dim3((unsigned)((c->N + 255) / 256)), dim3(256), 0, c->stream, nul, (const double* const*)c->rmax_tab, nb,
                           (const double*)c->bmax_g, nb * c->world, B, c->N_global, c->N, c->index0, method, seed, stage, u, nd, c->idx, flag));
    double* dst = c->cloud[c->cur ^ 1];
    SMC_CUDA(c, launch_pdl(k_gather_peer, dim3((unsigned)((c->N + 255) / 256)), dim3(256), 0, c->stream,
                           (double* const*)(c->peer_tab + (size_t)c->cur * c->world), (const int64_t*)c->peer_cnt, c->per, dst, (const int64_t*)c->idx,
                           c->N, d + 4, fused ? cl + col_off(c->N, d + 4) : dst + col_off(c->N, d + 4), nw_hist, flag));
    c->launches += 8;
    SMC_CUDA(c, cudaGetLastError());
    if (!fused) c->cur ^= 1;
    return SMCB200_OK;
}

// fused: the kernels are predicated on scal[SC_RESAMPLE], sum(weights) comes from the correction kernel, the rows are
// gathered into the other buffer and the weights are reset in the current one (the buffers do not flip: the moments and
// mutation kernels pick the gathered rows up from there and mutation writes the current buffer again).
int launch_resample_cloud(Ctx* c, int method, uint64_t seed, uint32_t stage, double u_override, bool fused, double* nw_hist)
{
    if (c->world > 1) return launch_resample_cloud_sharded(c, method, seed, stage, u_override, fused, nw_hist);
    const int d = c->d;
    double* cl = c->cloud[c->cur];
    double* dst = c->cloud[c->cur ^ 1];
    const Tiles t = weight_tiles(c->N);
    const double* flag = fused ? c->scal + SC_RESAMPLE : nullptr;
    int st = launch_resample_indices(c, cl + col_off(c->N, d + 4), 1, c->N, method, seed, stage, u_override, c->rmax, nullptr,
                                     c->idx, c->partials + (size_t)3 * t.P, c->counters + 2, c->scal + SC_SRES, fused, flag);
    if (st) return st;
    SMC_CUDA(c, launch_pdl(k_gather, dim3((unsigned)((c->N + 255) / 256)), dim3(256), 0, c->stream, (const double*)cl, dst,
                           (const int64_t*)c->idx, c->N, d + 4, fused ? cl + col_off(c->N, d + 4) : dst + col_off(c->N, d + 4), nw_hist, flag));
    c->launches += 1;
    SMC_CUDA(c, cudaGetLastError());
    if (!fused) c->cur ^= 1;
    return SMCB200_OK;
}

// ---- moments (two-pass form of the reference: the exported weighted_mean / weighted_cov) -----------------------
template <int D>
int launch_m2(Ctx* c, const double* cl, double* partials, const Tiles& t)
{
    const size_t stage = (size_t)(D + 1) * M2_CH, red = (size_t)(D * (D + 1) / 2) * 33;   // the reduction rows alias the staged columns
    const size_t smem = sizeof(double) * (stage > red ? stage : red);
    if (smem > 48 * 1024)
        SMC_CUDA(c, cudaFuncSetAttribute(k_moments2<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_moments2<D><<<t.ntiles, 32 * M2_G, smem, c->stream>>>(cl, c->N, c->msum, partials, t.P);
    return SMCB200_OK;
}

int launch_moments(Ctx* c)
{
    const int d = c->d;
    const double* cl = c->cloud[c->cur];
    const Tiles t = moment_tiles(c->N);
    const Tiles tc = chunk_tiles(c->N);
    const int E = d * (d + 1) / 2;
    double* part = c->mpartials;  // max([1+d][P_m], [E][P_c]) -- sized at cloud creation
    k_moments1<<<dim3(t.ntiles, (d + 4) / 4), 128, 0, c->stream>>>(cl, c->N, d, part, t.P);
    k_tree_finalize<<<1 + d, 256, 0, c->stream>>>(part, t.ntiles, t.P, c->world > 1 ? c->scal_loc + SL_MSUM : c->msum);
    { int st0 = reduce_ranks(c, c->msum, c->scal_loc + SL_MSUM, 1 + d); if (st0) return st0; }
    // pass 2 reuses the partial buffer with the chunk geometry: clear the zero padding it relies on
    if (tc.ntiles < tc.P || t.ntiles < t.P)
        SMC_CUDA(c, cudaMemsetAsync(part, 0, sizeof(double) * c->partials_len_m, c->stream));
    int st = SMCB200_OK;
    switch (d) {
    case 2: st = launch_m2<2>(c, cl, part, tc); break;
    case 3: st = launch_m2<3>(c, cl, part, tc); break;
    case 9: st = launch_m2<9>(c, cl, part, tc); break;
    case 16: st = launch_m2<16>(c, cl, part, tc); break;
    case 20: st = launch_m2<20>(c, cl, part, tc); break;
    default: k_moments2_generic<<<tc.ntiles, 128, 0, c->stream>>>(cl, c->N, d, c->msum, part, tc.P); break;
    }
    if (st) return st;
    k_tree_finalize<<<E, 256, 0, c->stream>>>(part, tc.ntiles, tc.P, c->world > 1 ? c->scal_loc + SL_CSUM : c->csum);
    { int st0 = reduce_ranks(c, c->csum, c->scal_loc + SL_CSUM, E); if (st0) return st0; }
    if (tc.ntiles < tc.P || t.ntiles < t.P)
        SMC_CUDA(c, cudaMemsetAsync(part, 0, sizeof(double) * c->partials_len_m, c->stream));
    c->launches += 4;
    SMC_CUDA(c, cudaGetLastError());
    return SMCB200_OK;
}

// ---- one-pass moments + step size + proposal factor (the fused stage) ---------------------------------------------
template <int NT>
int launch_mma(Ctx* c, unsigned grid, const double* x0, const double* x1, const double* wcol, const double* shift, int64_t stride, int P)
{
    SMC_CUDA(c, launch_pdl(k_moments_mma<NT>, dim3(grid), dim3(32 * M1P_WARPS), 0, c->stream, x0, x1, wcol, c->N, c->d, (const double*)c->scal,
                           shift, stride, c->m1p_partials, P));
    return SMCB200_OK;
}

int launch_moments_prepare(Ctx* c, const BlockSpec& bs, double target)
{
    const int d = c->d;
    const double* x0 = c->cloud[c->cur];
    const double* x1 = c->cloud[c->cur ^ 1];
    const double* wcol = x0 + col_off(c->N, d + 4);
    // shift = parameter vector of global particle 0 in the (pre-selection) current buffer of rank 0
    const double* shift = (c->world > 1) ? c->shift_base[c->cur] : x0;
    const int64_t stride = (c->world > 1) ? c->n_rank0 : c->N;
    const Tiles t = unit_tiles(c->N);
    const int E = d * (d + 1) / 2, nq = 1 + d + E;
    const unsigned grid = (unsigned)t.ntiles;
    int st = SMCB200_OK;
    switch ((d + 1 + 7) / 8) {      // tiles of 8 variables holding d parameters + the constant
    case 1: st = launch_mma<1>(c, grid, x0, x1, wcol, shift, stride, t.P); break;
    case 2: st = launch_mma<2>(c, grid, x0, x1, wcol, shift, stride, t.P); break;
    case 3: st = launch_mma<3>(c, grid, x0, x1, wcol, shift, stride, t.P); break;
    case 4: st = launch_mma<4>(c, grid, x0, x1, wcol, shift, stride, t.P); break;
    case 5: st = launch_mma<5>(c, grid, x0, x1, wcol, shift, stride, t.P); break;
    default: return fail(c, SMCB200_ERR_UNSUPPORTED, "one-pass moments support n_para <= 39");
    }
    if (st) return st;
    SMC_CUDA(c, launch_pdl(k_moments_finish, dim3(nq), dim3(256), 0, c->stream, (const double*)c->m1p_partials, t.P, nq, c->m1p_sums,
                           c->m1p_sums + (1 + DMAX + PACKMAX), c->counters + 6, peer_ctx(c), shift, stride, bs, target, c->scal, c->mutc_dev));
    c->launches += 2;
    SMC_CUDA(c, cudaGetLastError());
    return SMCB200_OK;
}

// ---- blocks -------------------------------------------------------------------------------------------
// generate_free_blocks (src/helpers.jl:215-231): Fisher-Yates on the Philox stream, then cld-sized blocks
void generate_blocks(int n_free, int n_blocks, uint64_t seed, uint32_t stage, int* perm, int* sizes)
{
    for (int i = 0; i < n_free; ++i) perm[i] = i;
    for (int i = n_free - 1; i >= 1; --i) {
        const u32x4 r = rng4(seed, (uint32_t)i, stage, 0u, PURP_BLOCKS);
        const double u = u01(r.x, r.y);
        const int j = (int)(u * (double)(i + 1));
        const int t = perm[i]; perm[i] = perm[j]; perm[j] = t;
    }
    const int sub = (n_free + n_blocks - 1) / n_blocks;
    const int last = n_free - sub * (n_blocks - 1);
    for (int b = 0; b < n_blocks; ++b) sizes[b] = (b < n_blocks - 1) ? sub : last;
}

int make_blockspec(Ctx* c, int n_blocks, const int32_t* sizes, const int32_t* blocks_all, BlockSpec* bs)
{
    if (n_blocks < 1 || n_blocks > NBMAX) return fail(c, SMCB200_ERR_UNSUPPORTED, "n_blocks must be in 1..8");
    std::memset(bs, 0, sizeof(*bs));
    bs->n_blocks = n_blocks; bs->d = c->d; bs->n_free = c->n_free;
    int pos = 0, total = 0;
    for (int b = 0; b < n_blocks; ++b) {
        const int n = sizes[b];
        if (n < 1 || n > DMAX) return fail(c, SMCB200_ERR_BAD_ARGUMENT, "bad block size");
        bs->bsize[b] = n;
        int tmp[DMAX];
        for (int i = 0; i < n; ++i) {
            const int a = blocks_all[pos + i];
            if (a < 0 || a >= c->d || c->prior.fixed[a]) return fail(c, SMCB200_ERR_BAD_ARGUMENT, "block index is not a free parameter");
            tmp[i] = a;
        }
        for (int i = 1; i < n; ++i) {   // ascending parameter order inside a block
            const int x = tmp[i]; int j = i - 1;
            while (j >= 0 && tmp[j] > x) { tmp[j + 1] = tmp[j]; --j; }
            tmp[j + 1] = x;
        }
        for (int i = 0; i < n; ++i) bs->member[b][i] = (int8_t)tmp[i];
        pos += n; total += n;
    }
    if (total != c->n_free) return fail(c, SMCB200_ERR_BAD_ARGUMENT, "blocks do not partition the free parameters");
    return SMCB200_OK;
}

int check_ready(Ctx* c, bool need_lik)
{
    if (!c->cloud[0]) return fail(c, SMCB200_ERR_NOT_READY, "no cloud: call smcb200_cloud_create first");
    if (need_lik) {
        if (!c->have_params) return fail(c, SMCB200_ERR_NOT_READY, "no parameters: call smcb200_set_parameters first");
        if (c->lik[0].kind == SMCB200_LIK_NONE) return fail(c, SMCB200_ERR_NOT_READY, "no likelihood: call smcb200_set_likelihood first");
    }
    return SMCB200_OK;
}

// device temporaries of one call, released on every exit path
struct DevTemps {
    std::vector<void*> p;
    ~DevTemps() { for (void* q : p) cudaFree(q); }
    template <class T> cudaError_t alloc(T** out, size_t n) { cudaError_t e = cudaMalloc(out, sizeof(T) * (n ? n : 1)); if (e == cudaSuccess) p.push_back(*out); return e; }
};

// ---- one stage, enqueued without any host synchronisation ---------------------------------------------------------------
struct StageLaunch {
    const smcb200_stage_config* cfg;
    double phi_n1, phi_n;
    uint32_t stage;
    double c, accept, ess_prev, phi_prop; long long j; int resampled_last;
    bool use_carry;
    int n_phi;                 // adaptive: schedule already on the device (c->sched_dev)
    int hist_slot;             // -1: no weight-history columns
    bool time_phases;
};

int stage_enqueue(Ctx* c, const StageLaunch& L)
{
    const smcb200_stage_config* cfg = L.cfg;
    const int d = c->d;
    double* inc_dev = nullptr; double* nw_dev = nullptr;
    if (L.hist_slot >= 0) {
        inc_dev = c->hist_scr + (size_t)L.hist_slot * 2 * c->N;
        nw_dev = inc_dev + c->N;
    }
    // ---- adaptive phi + correction + ESS + resample decision: one cooperative kernel --------------------------
    if (L.time_phases) SMC_CUDA(c, cudaEventRecord(c->ev[0], c->stream));
    CorrectLaunch cl;
    cl.phi_n1 = L.phi_n1; cl.phi_n = L.phi_n; cl.pw = cfg->prior_weight; cl.lpod = cfg->log_prob_old_data;
    cl.threshold_ratio = cfg->threshold_ratio; cl.adaptive = cfg->adaptive; cl.use_carry = L.use_carry ? 1 : 0;
    cl.sched_dev = c->sched_dev; cl.n_phi = L.n_phi; cl.tempering_target = cfg->tempering_target;
    cl.c_in = L.c; cl.accept_in = L.accept; cl.ess_prev_in = L.ess_prev; cl.phi_prop_in = L.phi_prop; cl.j_in = L.j;
    cl.resampled_last_in = L.resampled_last;
    cl.inc_dev = inc_dev; cl.normw_dev = nw_dev;
    int st = launch_correct_coop(c, cl); if (st) return st;
    if (L.time_phases) SMC_CUDA(c, cudaEventRecord(c->ev[1], c->stream));
    // ---- selection, predicated on the device-side decision ------------------------------------------------------
    if (L.time_phases) SMC_CUDA(c, cudaEventRecord(c->ev[2], c->stream));
    st = launch_resample_cloud(c, cfg->resample_method, cfg->seed, L.stage, -1.0, true, nw_dev); if (st) return st;
    if (L.time_phases) SMC_CUDA(c, cudaEventRecord(c->ev[3], c->stream));
    if (L.hist_slot >= 0) SMC_CUDA(c, cudaEventRecord(c->hist_ready[L.hist_slot], c->stream));
    // ---- moments, step size, proposal factor ---------------------------------------------------------------------
    if (L.time_phases) SMC_CUDA(c, cudaEventRecord(c->ev[4], c->stream));
    int perm[DMAX], sizes[NBMAX], ball[DMAX];
    generate_blocks(c->n_free, cfg->n_blocks, cfg->seed, L.stage, perm, sizes);
    for (int a = 0; a < c->n_free; ++a) ball[a] = c->free_idx[perm[a]];
    BlockSpec bs;
    st = make_blockspec(c, cfg->n_blocks, sizes, ball, &bs); if (st) return st;
    st = launch_moments_prepare(c, bs, cfg->target); if (st) return st;
    c->mutc_host->n_blocks = cfg->n_blocks;
    st = mutate_upload_proposal(c, true); if (st) return st;
    if (L.time_phases) SMC_CUDA(c, cudaEventRecord(c->ev[5], c->stream));
    // ---- mutation (+ mean accept) -------------------------------------------------------------------------------
    if (L.time_phases) SMC_CUDA(c, cudaEventRecord(c->ev[6], c->stream));
    st = mutate_launch(c, 0.0, cfg->alpha, cfg->n_mh_steps, cfg->has_old_data != 0, cfg->seed, L.stage, true); if (st) return st;
    if (L.time_phases) SMC_CUDA(c, cudaEventRecord(c->ev[7], c->stream));
    (void)d;
    return SMCB200_OK;
}

int check_stage_cfg(Ctx* c, const smcb200_stage_config* cfg)
{
    if (!(cfg->alpha >= 0.0 && cfg->alpha <= 1.0)) return fail(c, SMCB200_ERR_BAD_ARGUMENT, "alpha must be within [0, 1]");
    if (!(cfg->prior_weight >= 0.0 && cfg->prior_weight <= 1.0))
        return fail(c, SMCB200_ERR_BAD_ARGUMENT, "The keyword tempered_update_prior_weight must be within the interval [0, 1]");
    if (cfg->n_blocks < 1 || cfg->n_blocks > NBMAX) return fail(c, SMCB200_ERR_UNSUPPORTED, "n_blocks must be in 1..8");
    if (cfg->n_mh_steps < 1) return fail(c, SMCB200_ERR_BAD_ARGUMENT, "n_mh_steps must be >= 1");
    if (cfg->resample_method != SMCB200_RESAMPLE_SYSTEMATIC && cfg->resample_method != SMCB200_RESAMPLE_MULTINOMIAL)
        return fail(c, SMCB200_ERR_BAD_RESAMPLER, "Invalid resampler in SMC. Options are :systematic or :multinomial");
    if (!mutate_supported(c, cfg->has_old_data != 0)) return fail(c, SMCB200_ERR_UNSUPPORTED, "no device mutation kernel for this likelihood / n_para");
    return SMCB200_OK;
}

// stage summary (pinned host copy of the device scalars) -> result / state
int summary_to_result(Ctx* c, const double* h, const smcb200_stage_config* cfg, smcb200_stage_state* state, smcb200_stage_result* res)
{
    std::memset(res, 0, sizeof(*res));
    res->phi_n = h[SC_PHI_N]; res->ess = h[SC_ESS]; res->sum_weights = h[SC_S];
    res->status = (int)h[SC_STATUS];
    if (res->status == SMCB200_ERR_NAN_ESS) return fail(c, SMCB200_ERR_NAN_ESS, "No particles have non-zero weight.");
    res->resampled = h[SC_RESAMPLE] != 0.0 ? 1 : 0;
    res->c = h[SC_C]; res->accept = h[SC_ACCEPT];
    if (cfg->adaptive) {
        state->j = (int64_t)h[SC_J]; state->phi_prop = h[SC_PHI_PROP];
        state->resampled_last_period = 0;
    }
    if (res->resampled) state->resampled_last_period = 1;
    if (res->status) return fail(c, res->status, "proposal covariance is not positive definite");
    state->c = res->c; state->accept = res->accept; state->ess_prev = res->ess;
    return SMCB200_OK;
}

}  // namespace

// =====================================================================================================
// C ABI
// =====================================================================================================
extern "C" {

int32_t smcb200_abi_version(void) { return SMCB200_ABI_VERSION; }

const char* smcb200_status_string(int32_t s)
{
    switch (s) {
    case SMCB200_OK: return "ok";
    case SMCB200_ERR_NAN_ESS: return "No particles have non-zero weight (ESS is NaN)";
    case SMCB200_ERR_BAD_RESAMPLER: return "Invalid resampler in SMC. Options are :systematic, :multinomial";
    case SMCB200_ERR_BAD_ARGUMENT: return "bad argument";
    case SMCB200_ERR_NOT_POSDEF: return "proposal covariance is not positive definite";
    case SMCB200_ERR_CUDA: return "CUDA error";
    case SMCB200_ERR_NCCL: return "NCCL error";
    case SMCB200_ERR_UNSUPPORTED: return "unsupported likelihood / prior / dimension (no device kernel)";
    case SMCB200_ERR_NOT_READY: return "call order: cloud / parameters / likelihood not set";
    }
    return "unknown status";
}

int32_t smcb200_create(smcb200_ctx** out, int32_t device)
{
    if (!out) return SMCB200_ERR_BAD_ARGUMENT;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return SMCB200_ERR_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return SMCB200_ERR_CUDA;
    smcb200_ctx* c = new (std::nothrow) smcb200_ctx();
    if (!c) return SMCB200_ERR_BAD_ARGUMENT;
    c->device = device;
    bool ok = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaMalloc(&c->counters, sizeof(unsigned) * 16) == cudaSuccess;
    ok = ok && cudaMemset(c->counters, 0, sizeof(unsigned) * 16) == cudaSuccess;
    ok = ok && cudaMalloc(&c->ess_sq, sizeof(double) * 2 * ESS_K) == cudaSuccess;
    ok = ok && cudaMalloc(&c->scal, sizeof(double) * SC_COUNT) == cudaSuccess;
    ok = ok && cudaMemset(c->scal, 0, sizeof(double) * SC_COUNT) == cudaSuccess;
    ok = ok && cudaMallocHost(&c->h_scal, sizeof(double) * SC_COUNT) == cudaSuccess;
    ok = ok && cudaMalloc(&c->scal_loc, sizeof(double) * SL_COUNT) == cudaSuccess;
    ok = ok && cudaMemset(c->scal_loc, 0, sizeof(double) * SL_COUNT) == cudaSuccess;
    ok = ok && cudaMalloc(&c->gath, sizeof(double) * SL_COUNT * 16) == cudaSuccess;
    ok = ok && cudaMalloc(&c->phi_state, sizeof(PhiState)) == cudaSuccess;
    ok = ok && cudaMallocHost(&c->h_phi_state, sizeof(PhiState)) == cudaSuccess;
    ok = ok && cudaMalloc(&c->mutc_dev, sizeof(MutConst) + sizeof(double) * (DMAX + 3 * DMAX * DMAX)) == cudaSuccess;
    ok = ok && cudaMallocHost(&c->mutc_host, sizeof(MutConst)) == cudaSuccess;
    ok = ok && cudaMalloc(&c->status_dev, sizeof(int)) == cudaSuccess;
    ok = ok && cudaMemset(c->status_dev, 0, sizeof(int)) == cudaSuccess;
    ok = ok && cudaMallocHost(&c->h_status, 2 * sizeof(int)) == cudaSuccess;
    if (ok) c->h_status[1] = 0;
    ok = ok && cudaMallocHost(&c->h_moments, sizeof(double) * (1 + DMAX + PACKMAX)) == cudaSuccess;
    for (int i = 0; i < 8 && ok; ++i) ok = cudaEventCreate(&c->ev[i]) == cudaSuccess;
    for (int i = 0; i < 2 && ok; ++i) ok = cudaEventCreate(&c->tev[i]) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; i < HIST_RING && ok; ++i) {
        ok = cudaEventCreateWithFlags(&c->hist_ready[i], cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&c->hist_copied[i], cudaEventDisableTiming) == cudaSuccess;
    }
    ok = ok && cudaMallocHost(&c->h_summary, sizeof(double) * SC_COUNT * SUMMARY_RING) == cudaSuccess;
    ok = ok && cudaMalloc(&c->coop_seq, sizeof(unsigned long long)) == cudaSuccess;       // sequence number of the next cross-GPU reduction
    // developer switch.  One GPU, measured: the grid-barrier reduction (every block finishes the tile tree itself) beats the
    // mailbox form (last block reduces, all poll) -- adaptive solve 0.399 vs 0.422 ms at N = 2^20 -- so it stays the default
    c->coop_ll_single = std::getenv("SMCB200_COOP_MAILBOX") ? 1 : 0;
    ok = ok && cudaMalloc(&c->coop_ticket, sizeof(unsigned)) == cudaSuccess;
    ok = ok && cudaMemset(c->coop_ticket, 0, sizeof(unsigned)) == cudaSuccess;
    ok = ok && cudaMemset(c->coop_seq, 0, sizeof(unsigned long long)) == cudaSuccess;
    ok = ok && cudaMalloc(&c->acc_total, sizeof(unsigned long long)) == cudaSuccess;
    ok = ok && cudaMemset(c->acc_total, 0, sizeof(unsigned long long)) == cudaSuccess;
    ok = ok && cudaMalloc(&c->mb_epoch_dev, sizeof(unsigned long long)) == cudaSuccess;
    ok = ok && cudaMemset(c->mb_epoch_dev, 0, sizeof(unsigned long long)) == cudaSuccess;
    ok = ok && cudaMalloc(&c->mb_err, sizeof(int)) == cudaSuccess;
    ok = ok && cudaMemset(c->mb_err, 0, sizeof(int)) == cudaSuccess;
    if (ok) {
        int coop = 0;
        ok = cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device) == cudaSuccess && coop != 0;
        ok = ok && cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device) == cudaSuccess;
        ok = ok && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c->coop_blocks_per_sm[0], k_correct_coop<3>, COOP_NT, 0) == cudaSuccess;
        ok = ok && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c->coop_blocks_per_sm[1], k_correct_coop<7>, COOP_NT, 0) == cudaSuccess;
        ok = ok && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c->coop_blocks_per_sm[2], k_correct_coop<0>, COOP_NT, 0) == cudaSuccess;
        ok = ok && c->coop_blocks_per_sm[0] >= 1 && c->coop_blocks_per_sm[1] >= 1 && c->coop_blocks_per_sm[2] >= 1;
    }
    ok = ok && cudaFuncSetAttribute(k_scan<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SCAN_SMEM) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(k_scan<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SCAN_SMEM) == cudaSuccess;
    if (!ok) { smcb200_destroy(c); return SMCB200_ERR_CUDA; }
    std::memset(c->mutc_host, 0, sizeof(MutConst));
    *out = c;
    return SMCB200_OK;
}

int32_t smcb200_destroy(smcb200_ctx* c)
{
    if (!c) return SMCB200_OK;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    free_cloud(c);
    if (c->nccl_comm && nccl_api()) { nccl_api()->CommDestroy((ncclComm_t)c->nccl_comm); c->nccl_comm = nullptr; }
    cudaFree(c->scal_loc); cudaFree(c->gath);
    cudaFree(c->counters); cudaFree(c->scal); cudaFreeHost(c->h_scal); cudaFree(c->phi_state); cudaFreeHost(c->h_phi_state);
    cudaFree(c->mutc_dev); cudaFreeHost(c->mutc_host); cudaFree(c->status_dev); cudaFreeHost(c->h_status);
    cudaFreeHost(c->h_moments); cudaFree(c->sched_dev);
    cudaFree(c->as_data[0]); cudaFree(c->as_data[1]); cudaFree(c->ess_sq);
    for (int r = 0; r < 16; ++r) if (c->mbox_open[r]) cudaIpcCloseMemHandle(c->mbox_open[r]);
    cudaFree(c->mbox); cudaFree(c->mbox_tab); cudaFree(c->mb_err); cudaFree(c->mb_epoch_dev); cudaFree(c->acc_total); cudaFree(c->coop_seq); cudaFree(c->coop_ticket);
    cudaFreeHost(c->h_summary);
    for (int i = 0; i < HIST_RING; ++i) {
        if (c->hist_ready[i]) cudaEventDestroy(c->hist_ready[i]);
        if (c->hist_copied[i]) cudaEventDestroy(c->hist_copied[i]);
    }
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    for (int i = 0; i < 8; ++i) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    for (int i = 0; i < 2; ++i) if (c->tev[i]) cudaEventDestroy(c->tev[i]);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return SMCB200_OK;
}

const char* smcb200_last_error(const smcb200_ctx* c) { return c ? c->err.c_str() : "null context"; }

int32_t smcb200_comm_unique_id(void* id)
{
    NcclApi* nc = nccl_api();
    if (!nc || !id) return SMCB200_ERR_NCCL;
    ncclUniqueId u;
    if (nc->GetUniqueId(&u) != ncclSuccess) return SMCB200_ERR_NCCL;
    std::memcpy(id, &u, sizeof(u));
    return SMCB200_OK;
}

int32_t smcb200_comm_init(smcb200_ctx* c, int32_t rank, int32_t world, const void* id)
{
    if (!c) return SMCB200_ERR_BAD_ARGUMENT;
    if (c->cloud[0]) return fail(c, SMCB200_ERR_BAD_ARGUMENT, "smcb200_comm_init must precede smcb200_cloud_create");
    if (world == 1 && rank == 0) { c->rank = 0; c->world = 1; return SMCB200_OK; }
    if (world < 1 || world > 16 || (world & (world - 1)) || rank < 0 || rank >= world || !id)
        return fail(c, SMCB200_ERR_BAD_ARGUMENT, "world must be a power of two <= 16 and 0 <= rank < world");
    NcclApi* nc = nccl_api();
    if (!nc) return fail(c, SMCB200_ERR_NCCL, "libnccl.so.2 could not be loaded");
    cudaSetDevice(c->device);
    ncclUniqueId u;
    std::memcpy(&u, id, sizeof(u));
    ncclComm_t comm;
    SMC_NCCL(c, nc->CommInitRank(&comm, world, u, rank));
    c->nccl_comm = comm; c->rank = rank; c->world = world;
    return SMCB200_OK;
}

// ---- cloud --------------------------------------------------------------------------------------------
int32_t smcb200_cloud_create(smcb200_ctx* c, int64_t n_parts, int32_t n_para)
{
    if (!c || n_parts < 1 || n_para < 1) return c ? fail(c, SMCB200_ERR_BAD_ARGUMENT, "n_parts, n_para must be positive") : SMCB200_ERR_BAD_ARGUMENT;
    if (n_para > DMAX) return fail(c, SMCB200_ERR_UNSUPPORTED, "n_para > 32 has no device kernels");
    if (n_parts > ((int64_t)1 << 28)) return fail(c, SMCB200_ERR_UNSUPPORTED, "n_parts > 2^28");
    cudaSetDevice(c->device);
    free_cloud(c);
    c->N_global = n_parts; c->d = n_para;
    // contiguous shard of the zero-padded power-of-two index space (keeps every canonical tree shard-aligned)
    const int64_t P2 = next_pow2(n_parts);
    const int64_t per = P2 / c->world;
    if (c->world > 1 && (per % SCAN_TILE) != 0)
        return fail(c, SMCB200_ERR_UNSUPPORTED, "multi-GPU needs at least 4096 (padded) particles per rank");
    int64_t first = per * c->rank, last = first + per;
    if (first > n_parts) first = n_parts;
    if (last > n_parts) last = n_parts;
    c->index0 = first; c->N = last - first; c->per = per;
    if (c->N < 1) return fail(c, SMCB200_ERR_BAD_ARGUMENT, "empty shard: fewer particles than ranks");
    const size_t cols = (size_t)n_para + 5;
    SMC_CUDA(c, cudaMalloc(&c->cloud[0], sizeof(double) * cols * c->N));
    SMC_CUDA(c, cudaMalloc(&c->cloud[1], sizeof(double) * cols * c->N));
    SMC_CUDA(c, cudaMalloc(&c->tmp, sizeof(double) * c->N));
    SMC_CUDA(c, cudaMalloc(&c->rmax, sizeof(double) * c->N));
    SMC_CUDA(c, cudaMalloc(&c->idx, sizeof(int64_t) * c->N));
    const Tiles tw = weight_tiles(c->N), tm = moment_tiles(c->N);
    const int E = n_para * (n_para + 1) / 2;
    const size_t len = (size_t)6 * tw.P;
    const Tiles tch = chunk_tiles(c->N), tun = unit_tiles(c->N);
    size_t lm = (size_t)(n_para + 1) * tm.P;
    if ((size_t)E * tch.P > lm) lm = (size_t)E * tch.P;
    c->partials_len_m = lm;
    c->partials_len = len;
    SMC_CUDA(c, cudaMalloc(&c->partials, sizeof(double) * len));
    SMC_CUDA(c, cudaMemset(c->partials, 0, sizeof(double) * len));
    SMC_CUDA(c, cudaMalloc(&c->mpartials, sizeof(double) * lm));
    SMC_CUDA(c, cudaMemset(c->mpartials, 0, sizeof(double) * lm));
    SMC_CUDA(c, cudaMalloc(&c->ess_partials, sizeof(double) * 2 * ESS_K * (size_t)tw.P));
    SMC_CUDA(c, cudaMemset(c->ess_partials, 0, sizeof(double) * 2 * ESS_K * (size_t)tw.P));
    SMC_CUDA(c, cudaMalloc(&c->msum, sizeof(double) * (1 + DMAX)));
    SMC_CUDA(c, cudaMalloc(&c->csum, sizeof(double) * PACKMAX));
    SMC_CUDA(c, cudaMemset(c->cloud[0], 0, sizeof(double) * cols * c->N));
    // fused-stage buffers: weight-history ring, accept-column tile sums, one-pass moment partials, cooperative-grid partials
    SMC_CUDA(c, cudaMalloc(&c->hist_scr, sizeof(double) * (size_t)HIST_RING * 2 * c->N));
    const size_t nq1 = (size_t)1 + n_para + E;
    c->m1p_P = tun.P; c->m1p_len = nq1 * tun.P;
    SMC_CUDA(c, cudaMalloc(&c->m1p_partials, sizeof(double) * c->m1p_len));
    SMC_CUDA(c, cudaMemset(c->m1p_partials, 0, sizeof(double) * c->m1p_len));
    SMC_CUDA(c, cudaMalloc(&c->m1p_sums, sizeof(double) * 2 * (1 + DMAX + PACKMAX)));
    c->coop_partials_len = (size_t)2 * COOP_NQMAX * tw.P;
    SMC_CUDA(c, cudaMalloc(&c->coop_partials, sizeof(double) * c->coop_partials_len));
    SMC_CUDA(c, cudaMemset(c->coop_partials, 0, sizeof(double) * c->coop_partials_len));
    c->cur = 0;
    if (c->world == 1 && !c->mbox) {
        // one GPU: the cooperative correction kernel still reduces through the low-latency mailbox (its own)
        const size_t bytes = sizeof(double) * 2 * MB_NQ + sizeof(unsigned long long) * 2 + sizeof(unsigned long long) * LL_WORDS;
        SMC_CUDA(c, cudaMalloc(&c->mbox, bytes));
        SMC_CUDA(c, cudaMemset(c->mbox, 0, bytes));
        SMC_CUDA(c, cudaMalloc(&c->mbox_tab, sizeof(double*)));
        SMC_CUDA(c, cudaMemcpy(c->mbox_tab, &c->mbox, sizeof(double*), cudaMemcpyHostToDevice));
    }
    if (c->world > 1) {
        // peers' cloud buffers and running-max columns: exchange CUDA-IPC handles + shard sizes through the communicator
        NcclApi* nc = nccl_api();
        const int W = c->world;
        struct Info { cudaIpcMemHandle_t h[3]; int64_t count; int64_t pad; };
        Info mine;
        std::vector<Info> all(W);
        DevTemps tmp;
        Info *all_dev = nullptr, *mine_dev = nullptr;
        SMC_CUDA(c, cudaIpcGetMemHandle(&mine.h[0], c->cloud[0]));
        SMC_CUDA(c, cudaIpcGetMemHandle(&mine.h[1], c->cloud[1]));
        SMC_CUDA(c, cudaIpcGetMemHandle(&mine.h[2], c->rmax));
        mine.count = c->N; mine.pad = 0;
        SMC_CUDA(c, tmp.alloc(&all_dev, (size_t)W));
        SMC_CUDA(c, tmp.alloc(&mine_dev, 1));
        SMC_CUDA(c, cudaMemcpyAsync(mine_dev, &mine, sizeof(Info), cudaMemcpyHostToDevice, c->stream));
        SMC_NCCL(c, nc->AllGather(mine_dev, all_dev, sizeof(Info), ncclChar, (ncclComm_t)c->nccl_comm, c->stream));
        SMC_CUDA(c, cudaMemcpyAsync(all.data(), all_dev, sizeof(Info) * W, cudaMemcpyDeviceToHost, c->stream));
        SMC_CUDA(c, cudaStreamSynchronize(c->stream));
        std::vector<double*> tab(3 * W);
        std::vector<int64_t> cnt(W);
        for (int r = 0; r < W; ++r) {
            cnt[r] = all[r].count;
            for (int b = 0; b < 3; ++b) {
                if (r == c->rank) { tab[b * W + r] = (b < 2) ? c->cloud[b] : c->rmax; continue; }
                void* p = nullptr;
                SMC_CUDA(c, cudaIpcOpenMemHandle(&p, all[r].h[b], cudaIpcMemLazyEnablePeerAccess));
                c->ipc_open[b][r] = p;
                tab[b * W + r] = (double*)p;
            }
        }
        c->shift_base[0] = tab[0]; c->shift_base[1] = tab[W]; c->n_rank0 = cnt[0];
        SMC_CUDA(c, cudaMalloc(&c->peer_tab, sizeof(double*) * 2 * W));
        SMC_CUDA(c, cudaMalloc(&c->rmax_tab, sizeof(double*) * W));
        SMC_CUDA(c, cudaMalloc(&c->peer_cnt, sizeof(int64_t) * W));
        SMC_CUDA(c, cudaMemcpy(c->peer_tab, tab.data(), sizeof(double*) * 2 * W, cudaMemcpyHostToDevice));
        SMC_CUDA(c, cudaMemcpy(c->rmax_tab, tab.data() + 2 * W, sizeof(double*) * W, cudaMemcpyHostToDevice));
        SMC_CUDA(c, cudaMemcpy(c->peer_cnt, cnt.data(), sizeof(int64_t) * W, cudaMemcpyHostToDevice));
        // reduction mailboxes (allocated once per context, zeroed: epoch 0 = nothing published)
        if (!c->mbox) {
            const size_t bytes = sizeof(double) * 2 * W * MB_NQ + sizeof(unsigned long long) * 2 * W +
                                 sizeof(unsigned long long) * LL_WORDS;      // value slots, flags, low-latency words
            SMC_CUDA(c, cudaMalloc(&c->mbox, bytes));
            SMC_CUDA(c, cudaMemset(c->mbox, 0, bytes));
            cudaIpcMemHandle_t mh, *mh_all_dev = nullptr, *mh_dev = nullptr;
            std::vector<cudaIpcMemHandle_t> mh_all(W);
            SMC_CUDA(c, cudaIpcGetMemHandle(&mh, c->mbox));
            SMC_CUDA(c, tmp.alloc(&mh_all_dev, (size_t)W));
            SMC_CUDA(c, tmp.alloc(&mh_dev, 1));
            SMC_CUDA(c, cudaMemcpyAsync(mh_dev, &mh, sizeof(mh), cudaMemcpyHostToDevice, c->stream));
            SMC_NCCL(c, nc->AllGather(mh_dev, mh_all_dev, sizeof(mh), ncclChar, (ncclComm_t)c->nccl_comm, c->stream));
            SMC_CUDA(c, cudaMemcpyAsync(mh_all.data(), mh_all_dev, sizeof(mh) * W, cudaMemcpyDeviceToHost, c->stream));
            SMC_CUDA(c, cudaStreamSynchronize(c->stream));
            std::vector<double*> mt(W);
            for (int r = 0; r < W; ++r) {
                if (r == c->rank) { mt[r] = c->mbox; continue; }
                void* p = nullptr;
                SMC_CUDA(c, cudaIpcOpenMemHandle(&p, mh_all[r], cudaIpcMemLazyEnablePeerAccess));
                c->mbox_open[r] = p;
                mt[r] = (double*)p;
            }
            SMC_CUDA(c, cudaMalloc(&c->mbox_tab, sizeof(double*) * W));
            SMC_CUDA(c, cudaMemcpy(c->mbox_tab, mt.data(), sizeof(double*) * W, cudaMemcpyHostToDevice));
        }
        SMC_CUDA(c, cudaMalloc(&c->bmax_g, sizeof(double) * (size_t)(per / SCAN_TILE) * W));
    }
    return SMCB200_OK;
}

int32_t smcb200_cloud_shard(const smcb200_ctx* c, int64_t* first, int64_t* count)
{
    if (!c || !c->cloud[0]) return SMCB200_ERR_NOT_READY;
    if (first) *first = c->index0;
    if (count) *count = c->N;
    return SMCB200_OK;
}

int32_t smcb200_cloud_upload(smcb200_ctx* c, const double* p, int64_t ld, int64_t row0)
{
    int st = check_ready(c, false); if (st) return st;
    if (!p || ld < row0 + c->N) return fail(c, SMCB200_ERR_BAD_ARGUMENT, "update_cloud!: host matrix is too small");
    cudaSetDevice(c->device);
    SMC_CUDA(c, cudaMemcpy2DAsync(c->cloud[c->cur], sizeof(double) * c->N, p + row0, sizeof(double) * ld, sizeof(double) * c->N,
                                  (size_t)c->d + 5, cudaMemcpyHostToDevice, c->stream));
    return sync(c);
}

int32_t smcb200_cloud_download(smcb200_ctx* c, double* p, int64_t ld, int64_t row0)
{
    int st = check_ready(c, false); if (st) return st;
    if (!p || ld < row0 + c->N) return fail(c, SMCB200_ERR_BAD_ARGUMENT, "host matrix is too small");
    cudaSetDevice(c->device);
    SMC_CUDA(c, cudaMemcpy2DAsync(p + row0, sizeof(double) * ld, c->cloud[c->cur], sizeof(double) * c->N, sizeof(double) * c->N,
                                  (size_t)c->d + 5, cudaMemcpyDeviceToHost, c->stream));
    return sync(c);
}

int32_t smcb200_cloud_read_column(smcb200_ctx* c, int32_t col, double* out)
{
    int st = check_ready(c, false); if (st) return st;
    if (col < 0 || col >= c->d + 5 || !out) return fail(c, SMCB200_ERR_BAD_ARGUMENT, "bad column");
    SMC_CUDA(c, cudaMemcpyAsync(out, c->cloud[c->cur] + col_off(c->N, col), sizeof(double) * c->N, cudaMemcpyDeviceToHost, c->stream));
    return sync(c);
}

int32_t smcb200_cloud_write_column(smcb200_ctx* c, int32_t col, const double* in)
{
    int st = check_ready(c, false); if (st) return st;
    if (col < 0 || col >= c->d + 5 || !in) return fail(c, SMCB200_ERR_BAD_ARGUMENT, "bad column");
    SMC_CUDA(c, cudaMemcpyAsync(c->cloud[c->cur] + col_off(c->N, col), in, sizeof(double) * c->N, cudaMemcpyHostToDevice, c->stream));
    return sync(c);
}

// ---- model --------------------------------------------------------------------------------------------
int32_t smcb200_set_parameters(smcb200_ctx* c, int32_t d, const int32_t* fixed, const double* lo, const double* hi,
                               const int32_t* kind, const double* p1, const double* p2)
{
    if (!c) return SMCB200_ERR_BAD_ARGUMENT;
    if (d < 1 || d > DMAX) return fail(c, SMCB200_ERR_UNSUPPORTED, "n_para must be in 1..32");
    if (c->cloud[0] && d != c->d) return fail(c, SMCB200_ERR_BAD_ARGUMENT, "n_para does not match the cloud");
    PriorConst& P = c->prior;
    std::memset(&P, 0, sizeof(P));
    c->n_free = 0;
    for (int k = 0; k < DMAX; ++k) P.fixed[k] = 1;
    for (int k = 0; k < d; ++k) {
        P.fixed[k] = fixed[k] ? 1 : 0;
        P.lo[k] = lo[k]; P.hi[k] = hi[k]; P.kind[k] = kind[k]; P.p1[k] = p1[k]; P.p2[k] = p2[k];
        if (!P.fixed[k]) c->free_idx[c->n_free++] = k;
        const double a = p1[k], b = p2[k];
        if (P.fixed[k]) continue;
        // per-parameter constants with the host libm (the oracle does the same with the same glibc)
        switch (kind[k]) {
        case SMCB200_PRIOR_NORMAL: P.cst[k] = -std::log(b) - HALF_LOG_2PI; P.a1[k] = 1.0 / b; break;
        case SMCB200_PRIOR_UNIFORM: P.cst[k] = -std::log(b - a); break;
        case SMCB200_PRIOR_GAMMA: P.cst[k] = -std::lgamma(a) - a * std::log(b); P.a1[k] = a - 1.0; P.a2[k] = 1.0 / b; break;
        case SMCB200_PRIOR_ROOT_INV_GAMMA:
            P.cst[k] = std::log(2.0) - std::lgamma(0.5 * a) + 0.5 * a * std::log(0.5 * a * b * b);
            P.a1[k] = 0.5 * (a + 1.0); P.a2[k] = 0.5 * a * b * b; break;
        case SMCB200_PRIOR_BETA: P.cst[k] = std::lgamma(a + b) - std::lgamma(a) - std::lgamma(b); P.a1[k] = a - 1.0; P.a2[k] = b - 1.0; break;
        case SMCB200_PRIOR_INV_GAMMA: P.cst[k] = a * std::log(b) - std::lgamma(a); P.a1[k] = a + 1.0; P.a2[k] = b; break;
        default: return fail(c, SMCB200_ERR_UNSUPPORTED, "unknown prior family");
        }
    }
    if (c->n_free == 0) return fail(c, SMCB200_ERR_BAD_ARGUMENT, "All model parameters are fixed!");   // smc_main.jl:237
    P.all_normal = 1; P.cst_sum = 0.0;
    for (int k = 0; k < d; ++k) {
        if (P.fixed[k] || P.kind[k] != SMCB200_PRIOR_NORMAL) P.all_normal = 0;
        P.cst_sum = P.cst_sum + P.cst[k];
    }
    if (!c->cloud[0]) c->d = d;
    c->have_params = true;
    cudaSetDevice(c->device);
    return mutate_upload_model(c);
}

int32_t smcb200_set_likelihood(smcb200_ctx* c, int32_t slot, int32_t kind, const int32_t* ip, int32_t n_ip, const double* dp,
                               int64_t n_dp)
{
    if (!c || slot < 0 || slot > 1) return c ? fail(c, SMCB200_ERR_BAD_ARGUMENT, "slot must be 0 or 1") : SMCB200_ERR_BAD_ARGUMENT;
    if (kind == SMCB200_LIK_AS_DSGE) {
        if (n_ip < 2 || !ip || !dp) return fail(c, SMCB200_ERR_BAD_ARGUMENT, "AS_DSGE needs iparams {n_periods, n_presample}");
        const int T = ip[0], npre = ip[1];
        if (T < 1 || npre < 0 || n_dp != (int64_t)3 * T) return fail(c, SMCB200_ERR_BAD_ARGUMENT, "AS_DSGE data must be 3 x n_periods");
        int has_nan = 0;
        for (int64_t i = 0; i < n_dp; ++i) {
            if (dp[i] != dp[i]) { has_nan = 1; continue; }      // NaN = missing observation (dropped from that period's update)
            if (!(dp[i] - dp[i] == 0.0)) return fail(c, SMCB200_ERR_BAD_ARGUMENT, "AS_DSGE: infinite observations");
        }
        cudaSetDevice(c->device);
        if (c->as_data[slot]) { cudaFree(c->as_data[slot]); c->as_data[slot] = nullptr; }
        SMC_CUDA(c, cudaMalloc(&c->as_data[slot], sizeof(double) * 3 * (size_t)T));
        SMC_CUDA(c, cudaMemcpy(c->as_data[slot], dp, sizeof(double) * 3 * (size_t)T, cudaMemcpyHostToDevice));
        c->as_host[slot].data = c->as_data[slot]; c->as_host[slot].T = T; c->as_host[slot].npre = npre;
        c->as_host[slot].has_nan = has_nan; c->as_host[slot].pad = 0;
        LikDesc L; L.kind = kind;
        c->lik[slot] = L;
        return mutate_upload_model(c);
    }
    if (kind != SMCB200_LIK_GAUSSREG) return fail(c, SMCB200_ERR_UNSUPPORTED, "likelihood family has no device functor");
    if (n_ip < 5 || !ip || !dp) return fail(c, SMCB200_ERR_BAD_ARGUMENT, "GAUSSREG needs 5 iparams");
    LikDesc L; L.kind = kind; L.neq = ip[0]; L.k = ip[1]; L.stride = ip[2]; L.coef_off = ip[3]; L.sig_off = ip[4];
    const int k = L.k, kp = k * (k + 1) / 2;
    if (L.neq < 1 || L.neq > EQMAX || k < 1 || k > DMAX || L.neq * k > 2 * DMAX || L.neq * kp > PACKMAX)
        return fail(c, SMCB200_ERR_UNSUPPORTED, "GAUSSREG shape out of range");
    const int64_t per = 4 + k + (int64_t)k * k;
    if (n_dp != per * L.neq) return fail(c, SMCB200_ERR_BAD_ARGUMENT, "GAUSSREG dparams length mismatch");
    LikSlot& S = c->lik_host[slot];
    std::memset(&S, 0, sizeof(S));
    for (int e = 0; e < L.neq; ++e) {
        const double* p = dp + per * e;
        S.T[e] = p[0]; S.qscale[e] = p[1]; S.rss[e] = p[2];
        S.cT[e] = -p[0] * HALF_LOG_2PI;
        if (L.sig_off < 0) { S.logs[e] = std::log(p[3]); S.inv_s2[e] = 1.0 / (p[3] * p[3]); }
        for (int j = 0; j < k; ++j) S.bhat[e * k + j] = p[4 + j];
        const double* U = p + 4 + k;
        for (int i = 0; i < k; ++i)
            for (int j = i; j < k; ++j) S.U[e * kp + i * k - (i * (i - 1)) / 2 + (j - i)] = U[i * k + j];
    }
    c->lik[slot] = L;
    cudaSetDevice(c->device);
    return mutate_upload_model(c);
}

int32_t smcb200_evaluate(smcb200_ctx* c, int32_t mode)
{
    int st = check_ready(c, true); if (st) return st;
    cudaSetDevice(c->device);
    st = evaluate_launch(c, mode); if (st) return st;
    return sync(c);
}

int32_t smcb200_initial_draw(smcb200_ctx* c, const double* fixed_values, uint64_t seed, int32_t max_tries)
{
    int st = check_ready(c, true); if (st) return st;
    if (max_tries < 1) return fail(c, SMCB200_ERR_BAD_ARGUMENT, "max_tries must be >= 1");
    cudaSetDevice(c->device);
    // the fixed parameters' values travel through the scratch column `tmp` (DMAX doubles; a tiny cloud gets its own buffer);
    // the device status word counts the particles that exhausted max_tries
    double* fv_dev = c->tmp;
    double* fv_alloc = nullptr;
    if (c->N < DMAX + 1) { SMC_CUDA(c, cudaMalloc(&fv_alloc, sizeof(double) * (DMAX + 1))); fv_dev = fv_alloc; }
    double fv[DMAX] = {0};
    for (int k = 0; k < c->d; ++k) {
        if (c->prior.fixed[k]) {
            if (!fixed_values) { cudaFree(fv_alloc); return fail(c, SMCB200_ERR_BAD_ARGUMENT, "fixed parameters need fixed_values"); }
            fv[k] = fixed_values[k];
        }
    }
    SMC_CUDA(c, cudaMemcpyAsync(fv_dev, fv, sizeof(double) * DMAX, cudaMemcpyHostToDevice, c->stream));
    SMC_CUDA(c, cudaMemsetAsync(c->status_dev, 0, sizeof(int), c->stream));
    st = initial_draw_launch(c, fv_dev, seed, max_tries, c->status_dev);
    if (st) { cudaFree(fv_alloc); return st; }
    SMC_CUDA(c, cudaMemcpyAsync(c->h_status, c->status_dev, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    st = sync(c);
    const int n_failed = *c->h_status;
    cudaMemsetAsync(c->status_dev, 0, sizeof(int), c->stream);
    cudaStreamSynchronize(c->stream);
    cudaFree(fv_alloc);
    if (st) return st;
    if (n_failed) return fail(c, SMCB200_ERR_BAD_ARGUMENT, "initial_draw!: some particles found no finite log-likelihood within max_tries prior draws");
    return SMCB200_OK;
}

// ---- stage operations -----------------------------------------------------------------------------------
int32_t smcb200_correct(smcb200_ctx* c, double phi_n1, double phi_n, double pw, double lpod, double* inc_out, double* normw_out,
                        double out[3])
{
    int st = check_ready(c, false); if (st) return st;
    if (!(pw >= 0.0 && pw <= 1.0))
        return fail(c, SMCB200_ERR_BAD_ARGUMENT, "The keyword tempered_update_prior_weight must be within the interval [0, 1]");
    cudaSetDevice(c->device);
    CorrectLaunch L;
    L.phi_n1 = phi_n1; L.phi_n = phi_n; L.pw = pw; L.lpod = lpod; L.threshold_ratio = 0.0;
    L.inc_dev = inc_out ? c->hist_scr : nullptr;
    L.normw_dev = normw_out ? c->hist_scr + c->N : nullptr;
    SMC_CUDA(c, cudaEventRecord(c->ev[0], c->stream));
    st = launch_correct_coop(c, L); if (st) return st;
    SMC_CUDA(c, cudaEventRecord(c->ev[1], c->stream));
    SMC_CUDA(c, cudaMemcpyAsync(c->h_scal, c->scal, sizeof(double) * SC_COUNT, cudaMemcpyDeviceToHost, c->stream));
    if (inc_out) SMC_CUDA(c, cudaMemcpyAsync(inc_out, L.inc_dev, sizeof(double) * c->N, cudaMemcpyDeviceToHost, c->stream));
    if (normw_out) SMC_CUDA(c, cudaMemcpyAsync(normw_out, L.normw_dev, sizeof(double) * c->N, cudaMemcpyDeviceToHost, c->stream));
    st = sync(c); if (st) return st;
    cudaEventElapsedTime(&c->last_ms[0], c->ev[0], c->ev[1]);
    if (out) { out[0] = c->h_scal[SC_S]; out[1] = c->h_scal[SC_ESS]; out[2] = c->h_scal[SC_S2]; }
    if (c->h_scal[SC_ESS] != c->h_scal[SC_ESS]) return fail(c, SMCB200_ERR_NAN_ESS, "No particles have non-zero weight.");
    return SMCB200_OK;
}

int32_t smcb200_ess_at(smcb200_ctx* c, const double* phi, int32_t K, double phi_n1, double* ess_out)
{
    int st = check_ready(c, false); if (st) return st;
    if (K < 0 || (K > 0 && (!phi || !ess_out))) return fail(c, SMCB200_ERR_BAD_ARGUMENT, "bad phi vector");
    cudaSetDevice(c->device);
    for (int k0 = 0; k0 < K; k0 += ESS_K) {
        double tr[ESS_K];
        for (int k = 0; k < ESS_K; ++k) tr[k] = phi[(k0 + k < K) ? k0 + k : K - 1];
        // the trial vector travels through the state block's trial[] array
        SMC_CUDA(c, cudaMemcpyAsync(c->phi_state->trial, tr, sizeof(tr), cudaMemcpyHostToDevice, c->stream));
        st = launch_ess_multi(c, phi_n1, c->phi_state->trial); if (st) return st;
        double sq[2 * ESS_K];
        SMC_CUDA(c, cudaMemcpyAsync(sq, c->ess_sq, sizeof(sq), cudaMemcpyDeviceToHost, c->stream));
        st = sync(c); if (st) return st;
        for (int k = 0; k < ESS_K && k0 + k < K; ++k) ess_out[k0 + k] = (sq[k] * sq[k]) / sq[ESS_K + k];
    }
    return SMCB200_OK;
}

int32_t smcb200_solve_adaptive_phi(smcb200_ctx* c, const double* sched, int32_t n_phi, int64_t* j_io, double* phi_prop_io,
                                   double phi_n1, double tempering_target, double ess_prev, int32_t resampled_last, double* phi_n_out)
{
    int st = check_ready(c, false); if (st) return st;
    if (!sched || n_phi < 1 || !j_io || !phi_prop_io || !phi_n_out) return fail(c, SMCB200_ERR_BAD_ARGUMENT, "null argument");
    cudaSetDevice(c->device);
    st = upload_schedule(c, sched, n_phi); if (st) return st;
    // one cooperative launch: schedule walk + bisection on the device, no host polling
    CorrectLaunch L;
    L.phi_n1 = phi_n1; L.adaptive = 1; L.solve_only = 1; L.use_carry = 1;      // (carry: leaves c / accept / status alone)
    L.sched_dev = c->sched_dev; L.n_phi = n_phi; L.tempering_target = tempering_target;
    L.ess_prev_in = ess_prev; L.phi_prop_in = *phi_prop_io; L.j_in = *j_io; L.resampled_last_in = resampled_last;
    st = launch_correct_coop(c, L); if (st) return st;
    SMC_CUDA(c, cudaMemcpyAsync(c->h_scal, c->scal, sizeof(double) * SC_COUNT, cudaMemcpyDeviceToHost, c->stream));
    st = sync(c); if (st) return st;
    const double phi_n = c->h_scal[SC_PHI_N];
    if (!(phi_n == phi_n)) return fail(c, SMCB200_ERR_NAN_ESS, "solve_adaptive_phi did not converge (NaN ESS?)");
    *j_io = (int64_t)c->h_scal[SC_J]; *phi_prop_io = c->h_scal[SC_PHI_PROP]; *phi_n_out = phi_n;
    return SMCB200_OK;
}

int32_t smcb200_resample(smcb200_ctx* c, int32_t method, uint64_t seed, uint32_t stage, double u_override, int64_t* idx_out)
{
    int st = check_ready(c, false); if (st) return st;
    cudaSetDevice(c->device);
    SMC_CUDA(c, cudaEventRecord(c->ev[2], c->stream));
    st = launch_resample_cloud(c, method, seed, stage, u_override, false, nullptr); if (st) return st;
    SMC_CUDA(c, cudaEventRecord(c->ev[3], c->stream));
    if (idx_out) SMC_CUDA(c, cudaMemcpyAsync(idx_out, c->idx, sizeof(int64_t) * c->N, cudaMemcpyDeviceToHost, c->stream));
    st = sync(c); if (st) return st;
    cudaEventElapsedTime(&c->last_ms[1], c->ev[2], c->ev[3]);
    return SMCB200_OK;
}

int32_t smcb200_resample_weights(smcb200_ctx* c, const double* weights, int64_t n, int32_t method, uint64_t seed, uint32_t stage,
                                 double u_override, int64_t* idx_out, double* cum_out)
{
    return smcb200_resample_weights_n(c, weights, n, n, method, seed, stage, u_override, idx_out, cum_out);
}

int32_t smcb200_resample_weights_n(smcb200_ctx* c, const double* weights, int64_t n, int64_t n_out, int32_t method, uint64_t seed,
                                   uint32_t stage, double u_override, int64_t* idx_out, double* cum_out)
{
    if (!c || !weights || n < 1 || n_out < 1 || !idx_out) return c ? fail(c, SMCB200_ERR_BAD_ARGUMENT, "null argument") : SMCB200_ERR_BAD_ARGUMENT;
    cudaSetDevice(c->device);
    DevTemps tmp;                       // released on every return path
    double *w = nullptr, *r = nullptr, *cr = nullptr, *part = nullptr;
    int64_t* idx = nullptr;
    const Tiles t = weight_tiles(n);
    SMC_CUDA(c, tmp.alloc(&w, (size_t)n));
    SMC_CUDA(c, tmp.alloc(&r, (size_t)n));
    SMC_CUDA(c, tmp.alloc(&cr, (size_t)n));
    SMC_CUDA(c, tmp.alloc(&idx, (size_t)n_out));
    SMC_CUDA(c, tmp.alloc(&part, (size_t)t.P));
    SMC_CUDA(c, cudaMemsetAsync(part, 0, sizeof(double) * t.P, c->stream));
    SMC_CUDA(c, cudaMemcpyAsync(w, weights, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
    int st = launch_resample_indices(c, w, 0, n, method, seed, stage, u_override, r, cr, idx, part, c->counters + 4, c->scal_loc + SC_SRES,
                                     false, nullptr, n_out);
    if (st) { cudaStreamSynchronize(c->stream); return st; }
    SMC_CUDA(c, cudaMemcpyAsync(idx_out, idx, sizeof(int64_t) * n_out, cudaMemcpyDeviceToHost, c->stream));
    if (cum_out) SMC_CUDA(c, cudaMemcpyAsync(cum_out, cr, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    SMC_CUDA(c, cudaStreamSynchronize(c->stream));
    return SMCB200_OK;
}

int32_t smcb200_moments(smcb200_ctx* c, double* mean, double* cov)
{
    int st = check_ready(c, false); if (st) return st;
    cudaSetDevice(c->device);
    const int d = c->d, E = d * (d + 1) / 2;
    SMC_CUDA(c, cudaEventRecord(c->ev[4], c->stream));
    st = launch_moments(c); if (st) return st;
    SMC_CUDA(c, cudaEventRecord(c->ev[5], c->stream));
    SMC_CUDA(c, cudaMemcpyAsync(c->h_moments, c->msum, sizeof(double) * (1 + d), cudaMemcpyDeviceToHost, c->stream));
    SMC_CUDA(c, cudaMemcpyAsync(c->h_moments + 1 + DMAX, c->csum, sizeof(double) * E, cudaMemcpyDeviceToHost, c->stream));
    st = sync(c); if (st) return st;
    cudaEventElapsedTime(&c->last_ms[2], c->ev[4], c->ev[5]);
    const double sw = c->h_moments[0];
    if (mean) for (int k = 0; k < d; ++k) mean[k] = c->h_moments[1 + k] / sw;
    if (cov)
        for (int a = 0; a < d; ++a)
            for (int b = 0; b <= a; ++b) {
                const double v = c->h_moments[1 + DMAX + a * (a + 1) / 2 + b] / sw;
                cov[a * d + b] = v; cov[b * d + a] = v;
            }
    return SMCB200_OK;
}

int32_t smcb200_moments_onepass(smcb200_ctx* c, double* mean, double* cov)
{
    int st = check_ready(c, false); if (st) return st;
    if (!c->have_params) return fail(c, SMCB200_ERR_NOT_READY, "no parameters: call smcb200_set_parameters first");
    cudaSetDevice(c->device);
    const int d = c->d, E = d * (d + 1) / 2, nq = 1 + d + E;
    // the stage's moment pass on the current cloud (no selection pending, step size untouched): a one-block spec that
    // holds every free parameter keeps k_moments_finish's proposal preparation well defined
    int sizes[1] = {c->n_free}, ball[DMAX];
    for (int a = 0; a < c->n_free; ++a) ball[a] = c->free_idx[a];
    BlockSpec bs;
    st = make_blockspec(c, 1, sizes, ball, &bs); if (st) return st;
    double keep[SC_COUNT];
    SMC_CUDA(c, cudaMemcpyAsync(keep, c->scal, sizeof(keep), cudaMemcpyDeviceToHost, c->stream));
    SMC_CUDA(c, cudaStreamSynchronize(c->stream));
    double z[SC_COUNT]; std::memcpy(z, keep, sizeof(z));
    z[SC_RESAMPLE] = 0.0; z[SC_STATUS] = 0.0; z[SC_C] = 1.0; z[SC_ACCEPT] = 0.25;
    SMC_CUDA(c, cudaMemcpyAsync(c->scal, z, sizeof(z), cudaMemcpyHostToDevice, c->stream));
    SMC_CUDA(c, cudaEventRecord(c->ev[4], c->stream));
    st = launch_moments_prepare(c, bs, 0.25); if (st) return st;
    SMC_CUDA(c, cudaEventRecord(c->ev[5], c->stream));
    double sums[1 + DMAX + PACKMAX], shift[DMAX];
    SMC_CUDA(c, cudaMemcpyAsync(sums, c->m1p_sums + (1 + DMAX + PACKMAX), sizeof(double) * nq, cudaMemcpyDeviceToHost, c->stream));
    const double* sb = (c->world > 1) ? c->shift_base[c->cur] : c->cloud[c->cur];
    const int64_t stride = (c->world > 1) ? c->n_rank0 : c->N;
    SMC_CUDA(c, cudaMemcpy2DAsync(shift, sizeof(double), sb, sizeof(double) * stride, sizeof(double), d, cudaMemcpyDeviceToHost, c->stream));
    SMC_CUDA(c, cudaMemcpyAsync(c->scal, keep, sizeof(keep), cudaMemcpyHostToDevice, c->stream));
    st = sync(c); if (st) return st;
    cudaEventElapsedTime(&c->last_ms[2], c->ev[4], c->ev[5]);
    const double sw = sums[0];
    double e[DMAX];
    for (int k = 0; k < d; ++k) { e[k] = sums[1 + k] / sw; if (mean) mean[k] = shift[k] + e[k]; }
    if (cov)
        for (int a = 0; a < d; ++a)
            for (int b = 0; b <= a; ++b) {
                const double v = fma(-e[a], e[b], sums[1 + d + a * (a + 1) / 2 + b] / sw);
                cov[a * d + b] = v; cov[b * d + a] = v;
            }
    return SMCB200_OK;
}

int32_t smcb200_mutate(smcb200_ctx* c, const double* mean_fr, const double* cov_fr, int32_t n_free, int32_t n_blocks,
                       const int32_t* block_sizes, const int32_t* blocks_free, const int32_t* blocks_all, double phi_n,
                       double phi_n1, double cc, double alpha, int32_t n_mh_steps, int32_t has_old, uint64_t seed, uint32_t stage,
                       double* mean_accept_out)
{
    (void)phi_n1; (void)blocks_free;
    int st = check_ready(c, true); if (st) return st;
    if (!mean_fr || !cov_fr || !block_sizes || !blocks_all) return fail(c, SMCB200_ERR_BAD_ARGUMENT, "null argument");
    if (n_free != c->n_free) return fail(c, SMCB200_ERR_BAD_ARGUMENT, "n_free does not match the ParameterVector");
    if (!(alpha >= 0.0 && alpha <= 1.0)) return fail(c, SMCB200_ERR_BAD_ARGUMENT, "alpha must be within [0, 1]");
    if (n_mh_steps < 1) return fail(c, SMCB200_ERR_BAD_ARGUMENT, "n_mh_steps must be >= 1");
    cudaSetDevice(c->device);
    BlockSpec bs;
    st = make_blockspec(c, n_blocks, block_sizes, blocks_all, &bs); if (st) return st;
    // scatter the free-parameter moments back to full parameter order (fixed rows/cols are never read)
    const int d = c->d;
    std::vector<double> mean(d, 0.0), cov((size_t)d * d, 0.0), work(2 * DMAX * DMAX);
    for (int a = 0; a < n_free; ++a) {
        mean[c->free_idx[a]] = mean_fr[a];
        for (int b = 0; b < n_free; ++b) cov[(size_t)c->free_idx[a] * d + c->free_idx[b]] = cov_fr[(size_t)a * n_free + b];
    }
    st = build_mutconst(mean.data(), cov.data(), d, bs, cc, c->mutc_host, work.data());
    if (st) return fail(c, SMCB200_ERR_NOT_POSDEF, "proposal covariance is not positive definite");
    st = mutate_upload_proposal(c, false); if (st) return st;
    SMC_CUDA(c, cudaEventRecord(c->ev[6], c->stream));
    st = mutate_launch(c, phi_n, alpha, n_mh_steps, has_old != 0, seed, stage, false); if (st) return st;
    SMC_CUDA(c, cudaEventRecord(c->ev[7], c->stream));
    SMC_CUDA(c, cudaMemcpyAsync(c->h_scal, c->scal, sizeof(double) * SC_COUNT, cudaMemcpyDeviceToHost, c->stream));
    st = sync(c); if (st) return st;
    cudaEventElapsedTime(&c->last_ms[3], c->ev[6], c->ev[7]);
    if (mean_accept_out) *mean_accept_out = c->h_scal[SC_ACCEPT];
    return SMCB200_OK;
}

int32_t smcb200_stage(smcb200_ctx* c, const smcb200_stage_config* cfg, smcb200_stage_state* state, const double* sched, int32_t n_phi,
                      double* inc_out, double* normw_out, smcb200_stage_result* res)
{
    int st = check_ready(c, true); if (st) return st;
    if (!cfg || !state || !res) return fail(c, SMCB200_ERR_BAD_ARGUMENT, "null argument");
    cudaSetDevice(c->device);
    std::memset(res, 0, sizeof(*res));
    st = check_stage_cfg(c, cfg); if (st) return st;
    if (cfg->adaptive) {
        if (!sched || n_phi < 1) return fail(c, SMCB200_ERR_BAD_ARGUMENT, "adaptive tempering needs the proposed fixed schedule");
        st = upload_schedule(c, sched, n_phi); if (st) return st;
    }
    StageLaunch L;
    L.cfg = cfg; L.phi_n1 = cfg->phi_n1; L.phi_n = cfg->phi_n; L.stage = cfg->stage;
    L.c = state->c; L.accept = state->accept; L.ess_prev = state->ess_prev; L.phi_prop = state->phi_prop; L.j = state->j;
    L.resampled_last = state->resampled_last_period; L.use_carry = false; L.n_phi = n_phi;
    L.hist_slot = (inc_out || normw_out) ? 0 : -1; L.time_phases = true;
    st = stage_enqueue(c, L); if (st) return st;
    SMC_CUDA(c, cudaMemcpyAsync(c->h_summary, c->scal, sizeof(double) * SC_COUNT, cudaMemcpyDeviceToHost, c->stream));
    if (L.hist_slot >= 0) {
        // the w_matrix / W_matrix columns leave on the copy stream while moments and mutation run
        SMC_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->hist_ready[0], 0));
        if (inc_out) SMC_CUDA(c, cudaMemcpyAsync(inc_out, c->hist_scr, sizeof(double) * c->N, cudaMemcpyDeviceToHost, c->copy_stream));
        if (normw_out) SMC_CUDA(c, cudaMemcpyAsync(normw_out, c->hist_scr + c->N, sizeof(double) * c->N, cudaMemcpyDeviceToHost, c->copy_stream));
        SMC_CUDA(c, cudaStreamSynchronize(c->copy_stream));
    }
    st = sync(c); if (st) return st;
    st = summary_to_result(c, c->h_summary, cfg, state, res); if (st) return st;
    cudaEventElapsedTime(&res->ms_correct, c->ev[0], c->ev[1]);
    cudaEventElapsedTime(&res->ms_resample, c->ev[2], c->ev[3]);
    cudaEventElapsedTime(&res->ms_moments, c->ev[4], c->ev[5]);
    cudaEventElapsedTime(&res->ms_mutate, c->ev[6], c->ev[7]);
    c->last_ms[0] = res->ms_correct; c->last_ms[1] = res->ms_resample; c->last_ms[2] = res->ms_moments; c->last_ms[3] = res->ms_mutate;
    return SMCB200_OK;
}

// The recursion `while phi_n < 1` (src/smc_main.jl:377-508) for up to n_stages stages in ONE call.  Fixed schedule: every
// stage is enqueued back to back (the step size, the mean accept rate and the resample decision are carried on the
// device), the stage summaries and the w/W history columns stream to the host behind the computation, and the host
// synchronises once.  Adaptive schedule: one synchronisation per stage (the host has to see phi_n to stop at 1).
int32_t smcb200_run_stages(smcb200_ctx* c, const smcb200_stage_config* cfg0, smcb200_stage_state* state, const double* sched,
                           int32_t n_phi, int32_t i_first, int32_t n_stages, double* inc_hist, double* normw_hist, int64_t ld_hist,
                           smcb200_stage_result* results, int32_t* n_done)
{
    int st = check_ready(c, true); if (st) return st;
    if (!cfg0 || !state || !sched || !results || !n_done || n_phi < 2 || i_first < 2 || n_stages < 1)
        return fail(c, SMCB200_ERR_BAD_ARGUMENT, "bad argument");
    if ((inc_hist || normw_hist) && ld_hist < c->N) return fail(c, SMCB200_ERR_BAD_ARGUMENT, "history matrices are too small");
    cudaSetDevice(c->device);
    *n_done = 0;
    st = check_stage_cfg(c, cfg0); if (st) return st;
    const bool adaptive = cfg0->adaptive != 0;
    if (!adaptive && i_first - 1 + n_stages > n_phi) return fail(c, SMCB200_ERR_BAD_ARGUMENT, "the fixed schedule has fewer stages");
    st = upload_schedule(c, sched, n_phi); if (st) return st;
    const bool hist = inc_hist || normw_hist;
    smcb200_stage_config cfg = *cfg0;
    double phi_n1 = cfg0->phi_n1;
    int done = 0;
    while (done < n_stages) {
        const int batch = adaptive ? 1 : ((n_stages - done < SUMMARY_RING) ? n_stages - done : SUMMARY_RING);
        for (int k = 0; k < batch; ++k) {
            const int s = done + k, i = i_first + s;
            StageLaunch L;
            L.cfg = &cfg;
            L.stage = (uint32_t)i;
            if (adaptive) { L.phi_n1 = phi_n1; L.phi_n = 0.0; }
            else { L.phi_n1 = sched[i - 2]; L.phi_n = sched[i - 1]; }
            L.c = state->c; L.accept = state->accept; L.ess_prev = state->ess_prev; L.phi_prop = state->phi_prop; L.j = state->j;
            L.resampled_last = state->resampled_last_period;
            L.use_carry = (s > 0) && !adaptive;        // later stages of a fixed-schedule batch: c / accept live on the device
            L.n_phi = n_phi;
            L.hist_slot = hist ? (s % HIST_RING) : -1;
            L.time_phases = false;
            if (hist && s >= HIST_RING) SMC_CUDA(c, cudaStreamWaitEvent(c->stream, c->hist_copied[L.hist_slot], 0));
            st = stage_enqueue(c, L); if (st) return st;
            SMC_CUDA(c, cudaMemcpyAsync(c->h_summary + (size_t)k * SC_COUNT, c->scal, sizeof(double) * SC_COUNT, cudaMemcpyDeviceToHost, c->stream));
            if (hist) {
                const double* src = c->hist_scr + (size_t)L.hist_slot * 2 * c->N;
                SMC_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->hist_ready[L.hist_slot], 0));
                if (inc_hist) SMC_CUDA(c, cudaMemcpyAsync(inc_hist + (size_t)s * ld_hist, src, sizeof(double) * c->N, cudaMemcpyDeviceToHost, c->copy_stream));
                if (normw_hist) SMC_CUDA(c, cudaMemcpyAsync(normw_hist + (size_t)s * ld_hist, src + c->N, sizeof(double) * c->N, cudaMemcpyDeviceToHost, c->copy_stream));
                SMC_CUDA(c, cudaEventRecord(c->hist_copied[L.hist_slot], c->copy_stream));
            }
        }
        st = sync(c); if (st) return st;
        for (int k = 0; k < batch; ++k) {
            st = summary_to_result(c, c->h_summary + (size_t)k * SC_COUNT, &cfg, state, &results[done + k]);
            if (st) { if (hist) cudaStreamSynchronize(c->copy_stream); *n_done = done + k; return st; }
        }
        done += batch;
        if (adaptive) {
            phi_n1 = results[done - 1].phi_n;
            if (phi_n1 >= 1.0) break;
        }
    }
    if (hist) SMC_CUDA(c, cudaStreamSynchronize(c->copy_stream));
    *n_done = done;
    return SMCB200_OK;
}

int32_t smcb200_stage_host(smcb200_ctx* c, double* particles, int64_t ld, const smcb200_stage_config* cfg, smcb200_stage_state* state,
                           const double* sched, int32_t n_phi, smcb200_stage_result* res)
{
    int st = smcb200_cloud_upload(c, particles, ld, 0); if (st) return st;
    st = smcb200_stage(c, cfg, state, sched, n_phi, nullptr, nullptr, res); if (st) return st;
    return smcb200_cloud_download(c, particles, ld, 0);
}

int32_t smcb200_fp64_peak(smcb200_ctx* c, int32_t iters, double* tflops_out)
{
    if (!c || iters < 1 || !tflops_out) return SMCB200_ERR_BAD_ARGUMENT;
    cudaSetDevice(c->device);
    const int blocks = c->sm_count * 8, threads = 256;
    DevTemps tmp;
    double* out = nullptr;
    SMC_CUDA(c, tmp.alloc(&out, (size_t)blocks * threads));
    k_fp64_peak<<<blocks, threads, 0, c->stream>>>(out, 16, 0.999999, 1e-9);          // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        SMC_CUDA(c, cudaEventRecord(c->tev[0], c->stream));
        k_fp64_peak<<<blocks, threads, 0, c->stream>>>(out, iters, 0.999999, 1e-9);
        SMC_CUDA(c, cudaEventRecord(c->tev[1], c->stream));
        SMC_CUDA(c, cudaEventSynchronize(c->tev[1]));
        float ms = 0.f;
        SMC_CUDA(c, cudaEventElapsedTime(&ms, c->tev[0], c->tev[1]));
        if (ms < best) best = ms;
    }
    c->launches += 6;
    SMC_CUDA(c, cudaGetLastError());
    const double flops = 2.0 * 64.0 * (double)iters * (double)blocks * (double)threads;    // 64 DFMA per thread and iteration
    *tflops_out = flops / ((double)best * 1e-3) / 1e12;
    return SMCB200_OK;
}

int64_t smcb200_kernel_launches(const smcb200_ctx* c) { return c ? c->launches : 0; }

int32_t smcb200_last_kernel_ms(const smcb200_ctx* c, int32_t which, float* ms)
{
    if (!c || which < 0 || which > 3 || !ms) return SMCB200_ERR_BAD_ARGUMENT;
    *ms = c->last_ms[which];
    return SMCB200_OK;
}

int32_t smcb200_timer_start(smcb200_ctx* c)
{
    if (!c) return SMCB200_ERR_BAD_ARGUMENT;
    cudaSetDevice(c->device);
    SMC_CUDA(c, cudaEventRecord(c->tev[0], c->stream));
    return SMCB200_OK;
}

int32_t smcb200_timer_stop(smcb200_ctx* c, float* ms)
{
    if (!c || !ms) return SMCB200_ERR_BAD_ARGUMENT;
    cudaSetDevice(c->device);
    SMC_CUDA(c, cudaEventRecord(c->tev[1], c->stream));
    SMC_CUDA(c, cudaEventSynchronize(c->tev[1]));
    SMC_CUDA(c, cudaEventElapsedTime(ms, c->tev[0], c->tev[1]));
    return SMCB200_OK;
}

int32_t smcb200_debug_math(smcb200_ctx* c, int32_t op, const double* x, int64_t n, uint64_t seed, double* out)
{
    if (!c || !x || !out || n < 1) return SMCB200_ERR_BAD_ARGUMENT;
    cudaSetDevice(c->device);
    double *dx = nullptr, *dout = nullptr;
    SMC_CUDA(c, cudaMalloc(&dx, sizeof(double) * n));
    SMC_CUDA(c, cudaMalloc(&dout, sizeof(double) * n));
    SMC_CUDA(c, cudaMemcpyAsync(dx, x, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
    k_debug_math<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(op, dx, n, seed, dout);
    c->launches += 1;
    cudaMemcpyAsync(out, dout, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream);
    int st = sync(c);
    cudaFree(dx); cudaFree(dout);
    return st;
}

}  // extern "C"
