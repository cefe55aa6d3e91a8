// common.cuh -- internal declarations of libsmcb200 (context, constants of the canonical orders).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/smcb200.h"
#include "detmath.cuh"

namespace smc {

// ---- canonical reduction orders (DESIGN.md "Numerical contract"; mirrored by the oracle) --------
constexpr int W_LANES = 256, W_R = 4, W_TILE = W_LANES * W_R;      // weight-type sums
constexpr int M_LANES = 32, M_R = 64, M_TILE = M_LANES * M_R;      // moment-type sums (pass 1)
constexpr int M2_CH = 512;                                         // scatter-matrix chunk (sequential fma)
constexpr int M1P_SC = 256;                                        // one-pass moments: sub-chunk accumulated by one warp (sequential fma)
constexpr int LEAF = 16;                                           // cumsum leaf (sequential)
constexpr int SCAN_THREADS = 256, SCAN_TILE = SCAN_THREADS * LEAF; // cumsum block tile

constexpr int DMAX = 32;               // max n_para with a device mutation kernel
constexpr int NBMAX = 8;               // max n_blocks
constexpr int PACKMAX = DMAX * (DMAX + 1) / 2;
constexpr int EQMAX = 8;
constexpr int HIST_RING = 4;         // device ring of weight-history column pairs in flight to the host
constexpr int SUMMARY_RING = 512;
constexpr int MB_NQ = 4096;             // doubles per rank and mailbox slot (>= PACKMAX + DMAX + 1)

// device scalar slots (ctx->scal): the stage's sums, decisions and the state carried from stage to stage.  Flags and
// counters are stored as doubles so that ONE small copy brings the whole stage summary to the host.
enum {
    SC_S = 0, SC_Q = 1, SC_S2 = 2, SC_SRES = 3,   // sum w~, sum W^2, sum W, sum W / n_parts   (global)
    SC_ACC = 4,                                    // sum of the accept column (global)
    SC_ESS = 5, SC_RESAMPLE = 6,                   // ESS of this stage; 1.0 when ESS < threshold_ratio * n_parts
    SC_PHI_N = 7, SC_PHI_N1 = 8,
    SC_C = 9, SC_ACCEPT = 10,                      // step size in force; mean accept of the latest mutation
    SC_ESS_PREV = 11, SC_PHI_PROP = 12, SC_J = 13, SC_RESAMPLED_LAST = 14,
    SC_STATUS = 15,                                // 0 or an SMCB200_ERR_* code raised on the device (poisons the stage)
    SC_EVALS = 16, SC_SWEEPS = 17,
    SC_COUNT = 32
};

constexpr int ESS_K = 15;   // trial phi per pass of the adaptive-phi solve: the 15 nodes of a depth-4 bisection tree
struct PhiState {           // adaptive-phi state machine, lives in device memory
    double lo, hi, phi_prop, ess_bar, phi_cur, phi_n1, phi_n, g_last;
    double trial[ESS_K];    // phase 0: phi_prop followed by the next schedule points; phase 1: bisection-tree nodes (heap order)
    long long j;
    int phase;              // 0 walk schedule, 1 bisect
    int done;
    int evals;
    int n_phi;
};

struct PriorConst {
    double lo[DMAX], hi[DMAX], p1[DMAX], p2[DMAX], cst[DMAX], a1[DMAX], a2[DMAX];
    int32_t kind[DMAX], fixed[DMAX];
    double cst_sum;            // all_normal: sum of the Normal log-normalisers (index order)
    int32_t all_normal, pad;   // every parameter free with a Normal prior: fused log-prior  -0.5 sum z^2 + cst_sum
};
struct LikSlot {
    double T[EQMAX], qscale[EQMAX], rss[EQMAX], cT[EQMAX], logs[EQMAX], inv_s2[EQMAX];
    double bhat[2 * DMAX];
    double U[PACKMAX];
};
struct MutConst {
    double L[NBMAX][PACKMAX];   // c * L_b embedded in parameter order, lower, packed by COLUMNS of the d x d matrix:
                                // L[r][j] at [j d - j(j-1)/2 + (r - j)]
    double csd[NBMAX][DMAX];    // c * sqrt(Sigma_ii)
    double isd[NBMAX][DMAX];    // 1 / sqrt(Sigma_ii), WITHOUT c (diagonal mixture density, helpers.jl:146)
    double isdn[NBMAX][DMAX];   // isd / sqrt(2 pi)
    double rl[NBMAX][DMAX];     // 1 / (c L_ii): reciprocal diagonal of the scaled factor (forward substitutions)
    double mu[DMAX];            // theta_bar
    double lognorm[NBMAX];      // n_b log(2 pi) + log det(c^2 Sigma_b)
    uint32_t mask[NBMAX];
    int32_t bsize[NBMAX];
    int32_t n_blocks, status;
};
struct BlockSpec {              // host-built, passed by value to the proposal-preparation kernel
    int32_t n_blocks, d, n_free, pad;
    int32_t bsize[NBMAX];
    int8_t member[NBMAX][DMAX]; // ascending parameter indices of block b
};
struct LikDesc {                // host copy of a likelihood slot's shape
    int kind = 0, neq = 0, k = 0, stride = 0, coef_off = 0, sig_off = -1;
};
struct ASConst {                // An-Schorfheide DSGE likelihood slot: 3 x T data (column-major) in global memory
    const double* data;
    int32_t T, npre;
    int32_t has_nan, pad;        // some observation is missing (NaN): the filter masks those slots period by period
};
struct PeerCtx {                 // the NVLink mailboxes of this rank's communicator (world == 1: unused)
    double* const* inbox;        // device table [world]: every rank's inbox
    unsigned long long* epoch;   // device-resident exchange counter (identical on all ranks: same sequence of exchanges)
    int* err;
    int rank, world;
};
struct MutArgs {
    double phi_n, alpha;
    int n_mh_steps, n_blocks, n_free;
    uint64_t seed;
    uint32_t stage;
    uint32_t rk[20];           // Philox round keys of `seed` (k0 + r W0, k1 + r W1): the kernel xors them in as constant-bank operands
    // fused stage: device-side inputs (all nullable)
    const double* scal;        // SC_* scalars: phi_n = scal[SC_PHI_N], the kernel returns at once when scal[SC_STATUS] != 0,
                               // rows are read from `alt_in` when scal[SC_RESAMPLE] != 0 (the stage resampled into it)
    const double* alt_in;
    // persistent warps: 32-particle work items handed out by a device counter
    unsigned* work_counter;
    // mean of the accept column (update_acceptance_rate!, particle.jl:466-468) fused into the kernel: exact integer total
    unsigned long long* acc_total;
    unsigned* acc_counter;     // block ticket (the last block finalises)
    double* acc_out;           // sum of the accept column (global)
    double* acc_mean_out;      // nullable: global sum / n_global -> scal[SC_ACCEPT]
    double n_global;
    PeerCtx pc;                // cross-GPU exchange of the shard roots inside the kernel's last block
};
struct Ctx;
struct KernelEntry {            // one likelihood functor's kernels + the constant-memory uploaders of its translation unit
    int kind, neq, k, stride, coef, sig, d, minb;
    void (*mut[2][3][2])(double*, int64_t, int64_t, MutArgs);   // [has_old][0 blocks / 1 single block / 2 single full block][mixture]
    const void* mut_fn(int has_old, int blk, int mix) const { return (const void*)mut[has_old][blk][mix]; }
    void (*eval)(double*, int64_t, int);
    void (*draw)(double*, int64_t, int64_t, const double*, uint64_t, int, int*);   // initial_draw!
    int (*upload_model)(Ctx*);
    int (*upload_proposal)(Ctx*, bool);
};

struct Ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    // sharding: contiguous ranges of the zero-padded power-of-two index space, `per` positions per rank
    int rank = 0, world = 1;
    void* nccl_comm = nullptr;
    int64_t N_global = 0, N = 0, index0 = 0, per = 0;
    double* scal_loc = nullptr;    // [256] local roots before the cross-rank tree
    double* gath = nullptr;        // [world][256] gathered roots
    double* bmax_g = nullptr;      // [world][nb_local] block maxima of the global cumsum's running max (world > 1)
    double** rmax_tab = nullptr;   // device table [world]: every rank's running-max column (CUDA IPC)
    const double* shift_base[2] = {nullptr, nullptr};   // rank 0's cloud buffers (moment shift = its particle 0)
    int64_t n_rank0 = 0;
    // small-reduction mailboxes: every rank owns an inbox [2 parities][world][MB_NQ] doubles + [2][world] epoch flags that
    // its peers write directly over NVLink (CUDA IPC mappings); see k_peer_exchange
    double* mbox = nullptr;
    double** mbox_tab = nullptr;   // device table [world]: every rank's inbox
    void* mbox_open[16] = {};      // opened peer mappings (host)
    unsigned long long* mb_epoch_dev = nullptr;   // device-resident exchange counter (kernels exchange without the host)
    int* mb_err = nullptr;         // device flag: a peer never showed up (time-out)
    double** peer_tab = nullptr;   // device table [2][world] of peers' cloud buffers (CUDA IPC)
    int64_t* peer_cnt = nullptr;   // device [world] particles held by each rank
    void* ipc_open[3][16] = {};    // opened peer mappings (host): cloud buffer 0 / 1, running-max column
    int d = 0;
    // device buffers
    double* cloud[2] = {nullptr, nullptr};
    int cur = 0;
    double* tmp = nullptr;         // N
    double* hist_scr = nullptr;    // [HIST_RING][2][N] incremental / normalised weight columns of recent stages
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t hist_ready[8]{}, hist_copied[8]{};
    unsigned long long* acc_total = nullptr;         // accept-column integer total of the running mutation kernel
    double* m1p_partials = nullptr; size_t m1p_len = 0; int m1p_P = 0;   // one-pass moments: [1 + d + E][P_chunks]
    double* m1p_sums = nullptr;    // [1 + d + E] shard-local roots, then global
    int coop_blocks_per_sm[3] = {0, 0, 0};    // occupancy of k_correct_coop<3 / 7 / 15> (queried once)
    int sm_count = 0;
    double* coop_partials = nullptr; size_t coop_partials_len = 0;
    unsigned long long* coop_seq = nullptr;          // device: sequence number of the correction kernel's next cross-GPU reduction
    unsigned* coop_ticket = nullptr;
    int coop_ll_single = 0;
    double* h_summary = nullptr;   // pinned ring of stage summaries [SUMMARY_RING][SC_COUNT]
    double* rmax = nullptr;        // N   running max of the cumsum
    int64_t* idx = nullptr;        // N
    double* partials = nullptr;    // tile partial sums, weight-type orders  [6][P_w]
    double* mpartials = nullptr;   // tile partial sums, moment-type orders  [max(1+d, E)][P_m]
    size_t partials_len = 0, partials_len_m = 0;
    unsigned* counters = nullptr;  // last-block counters
    double* scal = nullptr;        // SC_COUNT device scalars
    double* h_scal = nullptr;      // pinned mirror
    double* ess_partials = nullptr;  // [2 * ESS_K][P_w] tile partials of the multi-trial ESS pass
    double* ess_sq = nullptr;        // [2 * ESS_K] S_k then Q_k
    PhiState* phi_state = nullptr;
    PhiState* h_phi_state = nullptr;
    double* sched_dev = nullptr; int sched_cap = 0;
    double* scan_blocktot = nullptr; double* scan_blockoff = nullptr; double* scan_levels = nullptr; double* scan_bmax = nullptr;
    int64_t scan_nb_cap = 0;
    double* msum = nullptr;        // [1 + d] Sw, sum w x_k
    double* csum = nullptr;        // [d(d+1)/2] packed lower scatter sums
    double* h_moments = nullptr;   // pinned
    MutConst* mutc_dev = nullptr;  // staging copy in global memory
    MutConst* mutc_host = nullptr; // pinned
    int* status_dev = nullptr; int* h_status = nullptr;
    // model
    bool have_params = false;
    PriorConst prior{};
    LikDesc lik[2];
    LikSlot lik_host[2]{};
    ASConst as_host[2]{};
    double* as_data[2] = {nullptr, nullptr};   // device copies of the An-Schorfheide data
    int n_free = 0;
    int free_idx[DMAX]{};
    // bookkeeping
    long long launches = 0;
    float last_ms[4] = {0, 0, 0, 0};
    cudaEvent_t ev[8]{};
    cudaEvent_t tev[2]{};
    std::string err;
};

#define SMC_CUDA(ctx, call)                                                                        \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e__);                      \
            return SMCB200_ERR_CUDA;                                                               \
        }                                                                                          \
    } while (0)

inline int64_t next_pow2(int64_t n) { int64_t p = 1; while (p < n) p <<= 1; return p; }
inline int ilog2(int64_t n) { int k = 0; while ((int64_t(1) << k) < n) ++k; return k; }

// column pointers of the struct-of-arrays cloud
__host__ __device__ inline size_t col_off(int64_t N, int j) { return (size_t)j * (size_t)N; }

// mutation kernel launch shape
#ifndef SMC_MUT_THREADS
#define SMC_MUT_THREADS 128
#endif
#ifndef SMC_MUT_WARPS_PER_SM
#define SMC_MUT_WARPS_PER_SM 20
#endif
constexpr int MUT_THREADS = SMC_MUT_THREADS;
constexpr int MUT_MINB20 = SMC_MUT_WARPS_PER_SM * 32 / MUT_THREADS;   // resident blocks asked of ptxas for D <= 20

// ---- mutation dispatch (mutate.cu) ---------------------------------------------------------------
int mutate_upload_model(Ctx* ctx);                      // priors + likelihood slots -> __constant__
int mutate_upload_proposal(Ctx* ctx, bool from_device); // MutConst -> __constant__
bool mutate_supported(const Ctx* ctx, bool has_old);
int mutate_launch(Ctx* ctx, double phi_n, double alpha, int n_mh_steps, bool has_old, uint64_t seed, uint32_t stage, bool fused);
PeerCtx peer_ctx(const Ctx* ctx);
int evaluate_launch(Ctx* ctx, int mode);
int initial_draw_launch(Ctx* ctx, const double* fixed_values_dev, uint64_t seed, int max_tries, int* n_failed_dev);

}  // namespace smc
