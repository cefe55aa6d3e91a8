// mutate.cu -- dispatch of the mutation / evaluation / initial-draw kernels (mutate_kernel.cuh) over the
// likelihood functors compiled in the mut_*.cu translation units.
#include "common.cuh"

namespace smc {

void register_linreg_small(std::vector<KernelEntry>&);
void register_linreg_mid(std::vector<KernelEntry>&);
void register_linreg_20(std::vector<KernelEntry>&);
void register_linreg_large(std::vector<KernelEntry>&);
void register_linreg_fill_a(std::vector<KernelEntry>&);
void register_linreg_fill_b(std::vector<KernelEntry>&);
void register_linreg_fill_c(std::vector<KernelEntry>&);
void register_linreg_fill_d(std::vector<KernelEntry>&);
void register_linreg_fill_e(std::vector<KernelEntry>&);
void register_linreg_fill_f(std::vector<KernelEntry>&);
void register_equations(std::vector<KernelEntry>&);
void register_dsge(std::vector<KernelEntry>&);

static const std::vector<KernelEntry>& table()
{
    static const std::vector<KernelEntry> t = [] {
        std::vector<KernelEntry> v;
        register_linreg_20(v);
        register_equations(v);
        register_dsge(v);
#ifndef SMC_FAST_BUILD   /* developer builds keep only the kernels of the benchmark / smoke configurations */
        register_linreg_small(v);
        register_linreg_mid(v);
        register_linreg_large(v);
        register_linreg_fill_a(v);
        register_linreg_fill_b(v);
        register_linreg_fill_c(v);
        register_linreg_fill_d(v);
        register_linreg_fill_e(v);
        register_linreg_fill_f(v);
#endif
        return v;
    }();
    return t;
}

static const KernelEntry* find_entry(const Ctx* ctx)
{
    const LikDesc& l = ctx->lik[0];
    for (const auto& e : table()) {
        if (e.kind != l.kind || e.d != ctx->d) continue;
        if (l.kind == SMCB200_LIK_AS_DSGE) return &e;
        if (e.neq == l.neq && e.k == l.k && e.stride == l.stride && e.coef == l.coef_off && e.sig == l.sig_off) return &e;
    }
    return nullptr;
}

bool mutate_supported(const Ctx* ctx, bool has_old)
{
    const KernelEntry* e = find_entry(ctx);
    if (!e) return false;
    if (has_old) {
        const LikDesc &a = ctx->lik[0], &b = ctx->lik[1];
        if (b.kind != a.kind || b.neq != a.neq || b.k != a.k || b.stride != a.stride || b.coef_off != a.coef_off ||
            b.sig_off != a.sig_off)
            return false;
    }
    return true;
}

// priors + likelihood slots -> the __constant__ blocks of every translation unit (set-up time only)
int mutate_upload_model(Ctx* ctx)
{
    std::vector<int (*)(Ctx*)> done;
    for (const auto& e : table()) {
        bool seen = false;
        for (auto f : done) seen = seen || (f == e.upload_model);
        if (seen) continue;
        done.push_back(e.upload_model);
        const int st = e.upload_model(ctx);
        if (st) return st;
    }
    SMC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SMCB200_OK;
}

int mutate_upload_proposal(Ctx* ctx, bool from_device)
{
    const KernelEntry* e = find_entry(ctx);
    if (!e) {
        ctx->err = "no device mutation kernel for this likelihood / n_para";
        return SMCB200_ERR_UNSUPPORTED;
    }
    return e->upload_proposal(ctx, from_device);
}

PeerCtx peer_ctx(const Ctx* ctx)
{
    PeerCtx pc;
    pc.inbox = ctx->mbox_tab; pc.epoch = ctx->mb_epoch_dev; pc.err = ctx->mb_err; pc.rank = ctx->rank; pc.world = ctx->world;
    return pc;
}

// fused != 0: phi_n, the resample flag (which buffer holds the rows) and the poison status come from the device scalars
int mutate_launch(Ctx* ctx, double phi_n, double alpha, int n_mh_steps, bool has_old, uint64_t seed, uint32_t stage, bool fused)
{
    const KernelEntry* e = find_entry(ctx);
    if (!e || !mutate_supported(ctx, has_old)) {
        ctx->err = "no device mutation kernel for this likelihood / n_para";
        return SMCB200_ERR_UNSUPPORTED;
    }
    MutArgs a;
    a.phi_n = phi_n; a.alpha = alpha; a.n_mh_steps = n_mh_steps; a.n_blocks = ctx->mutc_host->n_blocks; a.n_free = ctx->n_free;
    a.seed = seed; a.stage = stage;
    philox_round_keys(seed, a.rk);
    a.scal = fused ? ctx->scal : nullptr;
    a.alt_in = ctx->cloud[ctx->cur ^ 1];
    a.work_counter = ctx->counters + 7; a.acc_total = ctx->acc_total; a.acc_counter = ctx->counters + 5;
    a.acc_out = ctx->scal + SC_ACC; a.acc_mean_out = ctx->scal + SC_ACCEPT; a.n_global = (double)ctx->N_global;
    a.pc = peer_ctx(ctx);
    const bool single = (a.n_blocks == 1);
    // persistent blocks: as many as stay resident (the kernel's launch bound), never more than there is work
    unsigned grid = (unsigned)((ctx->N + MUT_THREADS - 1) / MUT_THREADS);
    const unsigned resident = (unsigned)(e->minb * (ctx->sm_count > 0 ? ctx->sm_count : 1));
    if (grid > resident) grid = resident;
    const size_t smem = sizeof(double) * 2 * (size_t)e->d * MUT_THREADS;
    // one block that holds every parameter (none fixed): compile-time membership mask
    const bool full = single && ctx->n_free == e->d;      // (a single block always holds all free parameters)
    auto kern = e->mut[has_old ? 1 : 0][single ? (full ? 2 : 1) : 0][alpha < 1.0 ? 1 : 0];
    static std::vector<const void*> configured;           // per-kernel attributes are set once
    bool seen = false;
    for (const void* k : configured) seen = seen || (k == (const void*)kern);
    if (!seen) {
        if (smem > 48 * 1024)
            SMC_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        // the chain state lives in shared memory: ask for just enough carve-out to keep LIK::MINB blocks resident and leave
        // the rest of the 256 KB to L1 (the Kalman-filter functor keeps a few spilled registers there)
        cudaFuncAttributes fa;
        SMC_CUDA(ctx, cudaFuncGetAttributes(&fa, kern));
        const size_t need = (size_t)e->minb * (smem + fa.sharedSizeBytes + 1024);
        int pct = (int)((need * 100 + 228 * 1024 - 1) / (228 * 1024));
        if (pct > 100) pct = 100;
        SMC_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
        configured.push_back((const void*)kern);
    }
    kern<<<grid, MUT_THREADS, smem, ctx->stream>>>(ctx->cloud[ctx->cur], ctx->N, ctx->index0, a);
    ctx->launches++;
    SMC_CUDA(ctx, cudaGetLastError());
    return SMCB200_OK;
}

int evaluate_launch(Ctx* ctx, int mode)
{
    const KernelEntry* e = find_entry(ctx);
    if (!e) {
        ctx->err = "no device likelihood functor for this likelihood / n_para";
        return SMCB200_ERR_UNSUPPORTED;
    }
    const unsigned grid = (unsigned)((ctx->N + 127) / 128);
    e->eval<<<grid, 128, 0, ctx->stream>>>(ctx->cloud[ctx->cur], ctx->N, mode);
    ctx->launches++;
    SMC_CUDA(ctx, cudaGetLastError());
    return SMCB200_OK;
}

int initial_draw_launch(Ctx* ctx, const double* fixed_values_dev, uint64_t seed, int max_tries, int* n_failed_dev)
{
    const KernelEntry* e = find_entry(ctx);
    if (!e) {
        ctx->err = "no device likelihood functor for this likelihood / n_para";
        return SMCB200_ERR_UNSUPPORTED;
    }
    const unsigned grid = (unsigned)((ctx->N + 127) / 128);
    e->draw<<<grid, 128, 0, ctx->stream>>>(ctx->cloud[ctx->cur], ctx->N, ctx->index0, fixed_values_dev, seed, max_tries, n_failed_dev);
    ctx->launches++;
    SMC_CUDA(ctx, cudaGetLastError());
    return SMCB200_OK;
}

}  // namespace smc
