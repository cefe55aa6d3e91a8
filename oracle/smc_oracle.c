/*
 * smc_oracle.c -- CPU restatement of the SMC.jl per-stage hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (smc_jl_b200/, include/) may include,
 * link or call this file; only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
 * legs do.  The product path is the CUDA library and fails loudly without it.
 *
 * What it restates (all citations relative to /root/reference, SMC.jl v0.1.15 @ de6c318):
 *   correction / weights / ESS      src/smc_main.jl:400-427, src/particle.jl:250-259,362-369
 *   adaptive tempering              src/helpers.jl:9-56 (solve_adaptive_phi), :173-181 (compute_ESS)
 *   selection                       src/resample.jl:23-71 (systematic, multinomial), src/smc_main.jl:435-446
 *   step-size adaptation            src/smc_main.jl:453-455
 *   moments                         src/particle.jl:481-486,526-532, src/smc_main.jl:457-465
 *   blocks                          src/helpers.jl:215-260
 *   mutation                        src/mutation.jl:56-138, src/helpers.jl:87-100 (mixture draw), :128-164 (densities)
 *   write-back / accept mean        src/particle.jl:410-418,466-468
 *   stage-0 evaluators              src/initialization.jl:23-63,153-186
 *
 * The reference itself (Julia + unvendored ModelConstructors/Distributions/PDMats/Roots/StatsBase)
 * cannot run in this image; this restatement is pinned against the reference's own goldens in
 * tests/test_oracle_golden.py (see tests/golden/make_golden.py for provenance).
 *
 * Numerical contract (shared with the CUDA engine; written independently in both):
 *   - IEEE binary64 everywhere, no contraction except the explicit FMA() calls below
 *     (build with -ffp-contract=off).
 *   - exp/log/sincos are the fixed polynomial algorithms below (<= ~1 ulp of libm), so that the
 *     device and this file agree bit-for-bit; the reference's libm differs from them by <= 2 ulp.
 *   - Reductions use fixed, shard-invariant orders: canon_sum (strided-sequential then adjacent-
 *     pair binary tree) and the pairwise cumsum (Julia-style pairwise accumulate: 16-element
 *     sequential leaves, tree totals, top-down offsets).  Julia's own `sum`/`cumsum` orders are
 *     SIMD-/length-dependent and not bit-portable; ours are a fixed member of the same family.
 *   - Randomness: Philox4x32-10 keyed by the engine seed, counter = (global particle, stage,
 *     slot, purpose).  The reference's dSFMT streams are not reproducible => RNG-dependent
 *     parity is oracle<->device only ("parity unpinned" against the reference for those).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define FMA(a, b, c) __builtin_fma((a), (b), (c))
#define ORC_API __attribute__((visibility("default")))

typedef int64_t i64;
typedef int32_t i32;

/* ------------------------------------------------------------------------------------------ */
/* bit helpers                                                                                */
/* ------------------------------------------------------------------------------------------ */
static inline uint64_t d2u(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
static inline double u2d(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }

/* ------------------------------------------------------------------------------------------ */
/* deterministic elementary functions                                                         */
/* ------------------------------------------------------------------------------------------ */
ORC_API double orc_exp(double x)
{
    if (x != x) return x;
    if (x > 709.782712893384) return INFINITY;
    if (x < -745.1332191019412) return 0.0;
    const double LOG2E = 1.4426950408889634074;
    const double MAGIC = 6755399441055744.0; /* 1.5 * 2^52: round-to-nearest-integer by addition */
    const double LN2_HI = 6.93147180369123816490e-01;
    const double LN2_LO = 1.90821492927058770002e-10;
    double t = FMA(x, LOG2E, MAGIC);
    double kd = t - MAGIC;
    int k = (int)kd;
    double r = FMA(-kd, LN2_HI, x);
    r = FMA(-kd, LN2_LO, r);
    /* Taylor to degree 13 on |r| <= 0.347: truncation 4e-18 */
    double p = 1.6059043836821613e-10;          /* 1/13! */
    p = FMA(p, r, 2.08767569878681e-09);        /* 1/12! */
    p = FMA(p, r, 2.505210838544172e-08);       /* 1/11! */
    p = FMA(p, r, 2.755731922398589e-07);       /* 1/10! */
    p = FMA(p, r, 2.7557319223985893e-06);      /* 1/9!  */
    p = FMA(p, r, 2.48015873015873e-05);        /* 1/8!  */
    p = FMA(p, r, 1.984126984126984e-04);       /* 1/7!  */
    p = FMA(p, r, 1.388888888888889e-03);       /* 1/6!  */
    p = FMA(p, r, 8.333333333333333e-03);       /* 1/5!  */
    p = FMA(p, r, 4.1666666666666664e-02);      /* 1/4!  */
    p = FMA(p, r, 1.6666666666666666e-01);      /* 1/3!  */
    p = FMA(p, r, 0.5);
    p = FMA(p, r, 1.0);
    p = FMA(p, r, 1.0);
    if (k > 1023) { p *= 0x1p1023; k -= 1023; }
    if (k < -1021) { p *= 0x1p-1000; k += 1000; }
    return p * u2d((uint64_t)(k + 1023) << 52);
}

ORC_API double orc_log(double x)
{
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
    const double Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01,
                 Lg3 = 2.857142874366239149e-01, Lg4 = 2.222219843214978396e-01,
                 Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
                 Lg7 = 1.479819860511658591e-01;
    uint64_t ix = d2u(x);
    int32_t hx = (int32_t)(ix >> 32);
    uint32_t lx = (uint32_t)ix;
    int k = 0;
    if (hx < 0x00100000) {
        if (((hx & 0x7fffffff) | lx) == 0) return -INFINITY;
        if (hx < 0) return NAN;
        k -= 54; x *= 0x1p54; ix = d2u(x); hx = (int32_t)(ix >> 32);
    }
    if (hx >= 0x7ff00000) return x + x;
    k += (hx >> 20) - 1023;
    hx &= 0x000fffff;
    int32_t i = (hx + 0x95f64) & 0x100000;
    ix = ((uint64_t)(uint32_t)(hx | (i ^ 0x3ff00000)) << 32) | (ix & 0xffffffffu);
    x = u2d(ix);
    k += (i >> 20);
    double dk = (double)k;
    double f = x - 1.0;
    double s = f / (2.0 + f);
    double z = s * s;
    double w = z * z;
    double t1 = w * FMA(w, FMA(w, Lg6, Lg4), Lg2);
    double t2 = z * FMA(w, FMA(w, FMA(w, Lg7, Lg5), Lg3), Lg1);
    double R = t2 + t1;
    double hfsq = 0.5 * f * f;
    double u = FMA(s, hfsq + R, dk * ln2_lo);
    return FMA(dk, ln2_hi, -((hfsq - u) - f));
}

/* sin and cos of 2*pi*u for u in [0,1), by octant reduction (exact) + fdlibm kernels */
static inline void orc_sincos2pi(double u, double *sn, double *cs)
{
    const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03,
                 S3 = -1.98412698298579493134e-04, S4 = 2.75573137070700676789e-06,
                 S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
    const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03,
                 C3 = 2.48015872894767294178e-05, C4 = -2.75573143513906633035e-07,
                 C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
    const double PIO4 = 7.85398163397448278999e-01;
    double t = u * 8.0;
    int o = (int)t;               /* octant 0..7 */
    double f = t - (double)o;     /* exact */
    double g = (o & 1) ? (1.0 - f) : f;
    double x = g * PIO4;
    double z = x * x;
    double ps = FMA(z, FMA(z, FMA(z, FMA(z, FMA(z, S6, S5), S4), S3), S2), S1);
    double s = FMA(x * z, ps, x);
    double pc = FMA(z, FMA(z, FMA(z, FMA(z, FMA(z, C6, C5), C4), C3), C2), C1);
    double c = FMA(z * z, pc, FMA(-0.5, z, 1.0));
    double sp = (o & 1) ? c : s;  /* sin, cos of the angle inside the quadrant */
    double cp = (o & 1) ? s : c;
    switch (o >> 1) {
    case 0: *sn = sp;  *cs = cp;  break;
    case 1: *sn = cp;  *cs = -sp; break;
    case 2: *sn = -sp; *cs = -cp; break;
    default: *sn = -cp; *cs = sp; break;
    }
}
ORC_API void orc_sincos2pi_v(double u, double *out) { orc_sincos2pi(u, out, out + 1); }

/* ------------------------------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon et al. 2011), counter-based                                          */
/* ------------------------------------------------------------------------------------------ */
ORC_API void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* purposes (counter word 3) */
enum { PURP_STEP = 1, PURP_NORMAL = 2, PURP_RESAMPLE = 3, PURP_BLOCKS = 4, PURP_INIT = 5 };

static inline void rng4(uint64_t seed, uint32_t particle, uint32_t stage, uint32_t slot, uint32_t purpose,
                        uint32_t out[4])
{
    uint32_t ctr[4] = {particle, stage, slot, purpose};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    orc_philox4x32_10(ctr, key, out);
}
/* [0,1) with 53 bits */
static inline double u01(uint32_t hi, uint32_t lo) { return (double)((((uint64_t)hi << 32) | lo) >> 11) * 0x1p-53; }
/* (0,1] with 53 bits */
static inline double u01_open0(uint32_t hi, uint32_t lo) { return (double)(((((uint64_t)hi << 32) | lo) >> 11) + 1) * 0x1p-53; }

/* Box-Muller pair from one Philox block */
static inline void normal_pair(const uint32_t r[4], double *z0, double *z1)
{
    double u1 = u01_open0(r[0], r[1]);
    double u2 = u01(r[2], r[3]);
    double rad = sqrt(-2.0 * orc_log(u1));
    double sn, cs;
    orc_sincos2pi(u2, &sn, &cs);
    *z0 = rad * cs;
    *z1 = rad * sn;
}
/* Proposal normals: four N(0,1) variates from one Philox block, one 32-bit word each, through a table-driven
 * inverse normal CDF in binary32 (table: oracle/normal_table.h, written by tools/make_normal_table.py together with the
 * engine's copy; max |z - Phi^-1(p)| = 5e-7).  Word r: sign = bit 31; v = (r << 1) | 1, p = v / 2^33 in (0, 1/2);
 * f = (float)v selects the row (exponent, top three mantissa bits), x = 1 + (low 20 mantissa bits) / 2^23, and
 * z = ((b3 x + b2) x + b1) x + b0 with three binary32 fma.  Exact integer steps + IEEE binary32 conversion / fma only
 * => the device produces the same bits (DESIGN.md "Numerical contract"). */
#include "normal_table.h"
static const float ORC_NORMAL_TAB[4 * SMC_NORMAL_TABLE_ROWS] = {SMC_NORMAL_TABLE_VALUES};
static inline double normal_icdf(uint32_t r)
{
    uint32_t v = (r << 1) | 1u;
    float f = (float)v;                     /* round to nearest even */
    uint32_t fb; memcpy(&fb, &f, 4);
    const int SH = 23 - SMC_NORMAL_TABLE_LOG2SUB;
    const float *c = ORC_NORMAL_TAB + 4 * ((fb >> SH) - (127u << SMC_NORMAL_TABLE_LOG2SUB));
    uint32_t xb = (fb & ((1u << SH) - 1u)) | 0x3f800000u;
    float x; memcpy(&x, &xb, 4);
    float z = __builtin_fmaf(__builtin_fmaf(__builtin_fmaf(c[3], x, c[2]), x, c[1]), x, c[0]);
    uint32_t zb; memcpy(&zb, &z, 4);
    zb ^= (r & 0x80000000u);
    memcpy(&z, &zb, 4);
    return (double)z;
}
static inline void normal_quad(const uint32_t r[4], double z[4])
{
    for (int k = 0; k < 4; ++k) z[k] = normal_icdf(r[k]);
}
ORC_API void orc_normal_quad(uint64_t seed, uint32_t particle, uint32_t stage, uint32_t slot, double *out)
{
    uint32_t r[4];
    rng4(seed, particle, stage, slot, PURP_NORMAL, r);
    normal_quad(r, out);
}
ORC_API void orc_normal_pair(uint64_t seed, uint32_t particle, uint32_t stage, uint32_t slot, double *out)
{
    uint32_t r[4];
    rng4(seed, particle, stage, slot, PURP_NORMAL, r);
    normal_pair(r, out, out + 1);
}
ORC_API double orc_uniform(uint64_t seed, uint32_t particle, uint32_t stage, uint32_t slot, uint32_t purpose, int which)
{
    uint32_t r[4];
    rng4(seed, particle, stage, slot, purpose, r);
    return which ? u01(r[2], r[3]) : u01(r[0], r[1]);
}

/* ------------------------------------------------------------------------------------------ */
/* canonical reductions                                                                       */
/* ------------------------------------------------------------------------------------------ */
static double tree_inplace(double *v, i64 n_pow2)
{
    for (i64 s = 1; s < n_pow2; s <<= 1)
        for (i64 i = 0; i + s < n_pow2; i += 2 * s) v[i] = v[i] + v[i + s];
    return v[0];
}
static i64 next_pow2(i64 n) { i64 p = 1; while (p < n) p <<= 1; return p; }

/* number of tiles (padded to a power of two) used by the canonical order for n elements */
static i64 canon_ntiles_pow2(i64 n, int lanes, int R)
{
    i64 tile = (i64)lanes * R;
    i64 nt = (n + tile - 1) / tile;
    if (nt < 1) nt = 1;
    return next_pow2(nt);
}

/* generic canonical sum of term(i), i in [0,n): lane l of tile t accumulates elements
 * t*tile + r*lanes + l for r = 0..R-1 sequentially, then adjacent-pair tree over lanes,
 * then adjacent-pair tree over tiles (zero padded to a power of two). */
typedef double (*term_fn)(i64 i, void *ctx);
static double canon_sum_fn(term_fn f, void *ctx, i64 n, int lanes, int R)
{
    i64 tile = (i64)lanes * R;
    i64 P = canon_ntiles_pow2(n, lanes, R);
    double *tiles = (double *)calloc((size_t)P, sizeof(double));
    double *lane = (double *)malloc(sizeof(double) * (size_t)lanes);
    i64 nt = (n + tile - 1) / tile;
    for (i64 t = 0; t < nt; ++t) {
        for (int l = 0; l < lanes; ++l) {
            double acc = 0.0;
            for (int r = 0; r < R; ++r) {
                i64 i = t * tile + (i64)r * lanes + l;
                double x = (i < n) ? f(i, ctx) : 0.0;
                acc = (r == 0) ? x : acc + x;
            }
            lane[l] = acc;
        }
        tiles[t] = tree_inplace(lane, lanes);
    }
    double tot = tree_inplace(tiles, P);
    free(tiles); free(lane);
    return tot;
}
static double term_plain(i64 i, void *ctx) { return ((const double *)ctx)[i]; }
static double term_square(i64 i, void *ctx) { double x = ((const double *)ctx)[i]; return x * x; }

/* weights-style canonical sum: lanes = 256, R = 4 */
#define W_LANES 256
#define W_R 4
ORC_API double orc_canon_sum(const double *x, i64 n) { return canon_sum_fn(term_plain, (void *)x, n, W_LANES, W_R); }
ORC_API double orc_canon_sumsq(const double *x, i64 n) { return canon_sum_fn(term_square, (void *)x, n, W_LANES, W_R); }
ORC_API double orc_canon_sum_generic(const double *x, i64 n, int lanes, int R) { return canon_sum_fn(term_plain, (void *)x, n, lanes, R); }

/* Julia-style pairwise inclusive cumsum (cf. Base.accumulate_pairwise!, leaf < 128) on a
 * zero-padded power-of-two length with fixed 16-element sequential leaves:
 *   total(leaf)   = sequential sum of the leaf
 *   total(node)   = total(left) + total(right)
 *   offset(root)  = 0; offset(left) = offset(node); offset(right) = offset(node) + total(left)
 *   c[i]          = offset(leaf(i)) + running_sum_within_leaf(i)
 */
#define LEAF 16
static double cs_total(const double *x, i64 n, i64 lo, i64 len)
{
    if (len == LEAF) {
        double s = 0.0;
        for (i64 i = 0; i < LEAF; ++i) {
            double v = (lo + i < n) ? x[lo + i] : 0.0;
            s = (i == 0) ? v : s + v;
        }
        return s;
    }
    return cs_total(x, n, lo, len / 2) + cs_total(x, n, lo + len / 2, len / 2);
}
/* memoised variant to keep it O(n): totals per level */
ORC_API void orc_cumsum(const double *x, i64 n, double *c)
{
    i64 P = next_pow2(n < LEAF ? LEAF : n);
    i64 nl = P / LEAF;
    /* level 0: leaf totals; level k: nl >> k nodes */
    int nlev = 0; while (((i64)1 << nlev) < nl) ++nlev;
    double **tot = (double **)malloc(sizeof(double *) * (size_t)(nlev + 1));
    tot[0] = (double *)malloc(sizeof(double) * (size_t)nl);
    for (i64 b = 0; b < nl; ++b) tot[0][b] = cs_total(x, n, b * LEAF, LEAF);
    for (int k = 1; k <= nlev; ++k) {
        i64 m = nl >> k;
        tot[k] = (double *)malloc(sizeof(double) * (size_t)m);
        for (i64 b = 0; b < m; ++b) tot[k][b] = tot[k - 1][2 * b] + tot[k - 1][2 * b + 1];
    }
    /* top-down offsets */
    double *off = (double *)malloc(sizeof(double) * (size_t)nl);
    double *nxt = (double *)malloc(sizeof(double) * (size_t)nl);
    off[0] = 0.0;
    for (int k = nlev; k >= 1; --k) {
        i64 m = nl >> k; /* nodes at level k */
        for (i64 b = 0; b < m; ++b) {
            nxt[2 * b] = off[b];
            nxt[2 * b + 1] = off[b] + tot[k - 1][2 * b];
        }
        double *tmp = off; off = nxt; nxt = tmp;
    }
    for (i64 b = 0; b < nl; ++b) {
        double run = 0.0;
        for (i64 i = 0; i < LEAF; ++i) {
            i64 g = b * LEAF + i;
            if (g >= n) break;
            run = (i == 0) ? x[g] : run + x[g];
            c[g] = off[b] + run;
        }
    }
    for (int k = 0; k <= nlev; ++k) free(tot[k]);
    free(tot); free(off); free(nxt);
}

/* ------------------------------------------------------------------------------------------ */
/* Cloud column helpers: particles is column-major N x (d+5) (src/particle.jl:31-63)           */
/* ------------------------------------------------------------------------------------------ */
#define COL(cloud, N, j) ((cloud) + (size_t)(j) * (size_t)(N))
#define C_LOGLH(d) ((d) + 0)
#define C_LOGPRIOR(d) ((d) + 1)
#define C_OLDLOGLH(d) ((d) + 2)
#define C_ACCEPT(d) ((d) + 3)
#define C_WEIGHT(d) ((d) + 4)

/* ------------------------------------------------------------------------------------------ */
/* Correction (src/smc_main.jl:400-427)                                                       */
/* ------------------------------------------------------------------------------------------ */
static inline double inc_weight(double loglh, double old, double phi_n1, double phi_n, double pw,
                                double log_prob_old, double log_1m_pw)
{
    if (pw == 0.0) return orc_exp((phi_n1 - phi_n) * old + (phi_n - phi_n1) * loglh);
    if (pw == 1.0) return orc_exp((phi_n - phi_n1) * loglh);
    double inner = orc_log(orc_exp((old - log_prob_old) + log_1m_pw) + pw);
    return orc_exp((phi_n1 - phi_n) * inner + (phi_n - phi_n1) * loglh);
}

/* out[0] = sum of unnormalised weights, out[1] = ESS, out[2] = canonical sum of the normalised
 * weights.  inc_out / normw_out (nullable) receive the columns appended to w_matrix / W_matrix. */
ORC_API int orc_correct(double *cloud, i64 N, int d, double phi_n1, double phi_n, double pw,
                        double log_prob_old, double *inc_out, double *normw_out, double *out)
{
    const double *ll = COL(cloud, N, C_LOGLH(d)), *old = COL(cloud, N, C_OLDLOGLH(d));
    double *w = COL(cloud, N, C_WEIGHT(d));
    double l1 = (pw > 0.0 && pw < 1.0) ? orc_log(1.0 - pw) : 0.0;
    for (i64 i = 0; i < N; ++i) {
        double inc = inc_weight(ll[i], old[i], phi_n1, phi_n, pw, log_prob_old, l1);
        if (inc_out) inc_out[i] = inc;
        w[i] = w[i] * inc;                              /* update_weights!, particle.jl:250-259 */
    }
    double S = orc_canon_sum(w, N);
    double n = (double)N;
    for (i64 i = 0; i < N; ++i) w[i] = (w[i] * n) / S;   /* normalize_weights!, particle.jl:362-369 */
    if (normw_out) memcpy(normw_out, w, sizeof(double) * (size_t)N);
    double Q = orc_canon_sumsq(w, N);
    out[0] = S;
    out[1] = (n * n) / Q;                               /* smc_main.jl:427 */
    out[2] = orc_canon_sum(w, N);
    return (out[1] != out[1]) ? 1 : 0;                  /* NaN ESS => check_nan_ess would assert */
}

/* compute_ESS (src/helpers.jl:173-181); tmp is scratch of length N */
ORC_API double orc_compute_ess(const double *loglh, const double *w, const double *old, i64 N,
                               double phi_n, double phi_n1, double *tmp)
{
    for (i64 i = 0; i < N; ++i) {
        double inc = orc_exp((phi_n1 - phi_n) * old[i] + (phi_n - phi_n1) * loglh[i]);
        tmp[i] = w[i] * inc;
    }
    /* ESS = N^2 / sum (N x_i / S)^2 with S = sum x_i: N cancels, so the trial-phi evaluations of the adaptive
     * solve use the one-pass form S^2 / sum x_i^2 (orc_correct keeps the reference's normalise-then-square order) */
    double S = orc_canon_sum(tmp, N);
    return (S * S) / orc_canon_sumsq(tmp, N);
}

/* solve_adaptive_phi (src/helpers.jl:9-56).  j is the 1-based schedule cursor as in the reference.
 * Root: arithmetic bisection of g on [phi_n1, phi_prop] until the bracket has no interior double
 * (Roots.fzero(...; xtol = 0.) semantics; any point of the last bracket is a valid answer). */
struct ess_ctx { const double *ll, *w, *old; i64 N; double phi_n1, ess_bar; double *tmp; i64 evals; };
static double gfun(struct ess_ctx *c, double phi)
{
    c->evals++;
    return orc_compute_ess(c->ll, c->w, c->old, c->N, phi, c->phi_n1, c->tmp) - c->ess_bar;
}
ORC_API int orc_solve_adaptive_phi(const double *cloud, i64 N, int d, const double *sched, int n_phi,
                                   i64 *j_io, double *phi_prop_io, double phi_n1, double tempering_target,
                                   double ess_prev, int resampled_last, double *phi_n_out, i64 *evals_out)
{
    struct ess_ctx c;
    c.ll = COL(cloud, N, C_LOGLH(d)); c.w = COL(cloud, N, C_WEIGHT(d)); c.old = COL(cloud, N, C_OLDLOGLH(d));
    c.N = N; c.phi_n1 = phi_n1; c.evals = 0;
    c.ess_bar = resampled_last ? tempering_target * (double)N : tempering_target * ess_prev;
    c.tmp = (double *)malloc(sizeof(double) * (size_t)N);
    i64 j = *j_io;
    double phi_prop = *phi_prop_io;
    double gp = gfun(&c, phi_prop);
    while (gp >= 0.0 && j <= n_phi) {
        phi_prop = sched[j - 1];
        j += 1;
        gp = gfun(&c, phi_prop);
    }
    double phi_n;
    if (phi_prop != 1.0 || gp < 0.0) {
        double lo = phi_n1, hi = phi_prop;
        for (;;) {
            double mid = 0.5 * (lo + hi);
            if (!(mid > lo && mid < hi)) break;
            double gm = gfun(&c, mid);
            if (gm == 0.0) { lo = mid; break; }
            if (gm > 0.0) lo = mid; else hi = mid;
        }
        phi_n = (lo == phi_n1) ? hi : lo;
    } else {
        phi_n = 1.0;
    }
    free(c.tmp);
    *j_io = j; *phi_prop_io = phi_prop; *phi_n_out = phi_n;
    if (evals_out) *evals_out = c.evals;
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Selection (src/resample.jl:23-71)                                                          */
/* ------------------------------------------------------------------------------------------ */
enum { RESAMPLE_SYSTEMATIC = 0, RESAMPLE_MULTINOMIAL = 1 };

/* weights_in = normalized_weights / n_parts as passed at smc_main.jl:438.  idx is 1-based.
 * cum_out (nullable) receives cumsum(weights ./ sum(weights)).
 * Deviation: where resample.jl:60 would return 0 ("no index found", then crash in the gather)
 * we return N.  u_override >= 0 replaces the Philox draw of the systematic offset. */
ORC_API int orc_resample_n(const double *weights_in, i64 N, i64 n_out, int method, uint64_t seed, uint32_t stage,
                           double u_override, i64 *idx, double *cum_out);
ORC_API int orc_resample(const double *weights_in, i64 N, int method, uint64_t seed, uint32_t stage,
                         double u_override, i64 *idx, double *cum_out)
{
    return orc_resample_n(weights_in, N, N, method, seed, stage, u_override, idx, cum_out);
}
/* resample(weights; n_parts = n_out, method): n_out ancestors out of N weights (bridge initialisation,
 * smc_main.jl:262-268).  Thresholds are (i - 1 + u) / n_out.  Deviation: the reference's systematic search
 * only scans cumulative[1:n_parts] (resample.jl:54, so with n_parts < length(weights) the tail of the old cloud
 * is unreachable and the last thresholds find nothing); here all N weights are searched. */
ORC_API int orc_resample_n(const double *weights_in, i64 N, i64 n_out, int method, uint64_t seed, uint32_t stage,
                           double u_override, i64 *idx, double *cum_out)
{
    double *x = (double *)malloc(sizeof(double) * (size_t)N);
    double *cum = cum_out ? cum_out : (double *)malloc(sizeof(double) * (size_t)N);
    double S = orc_canon_sum(weights_in, N);
    for (i64 i = 0; i < N; ++i) x[i] = weights_in[i] / S;
    orc_cumsum(x, N, cum);
    double n = (double)n_out;
    if (method == RESAMPLE_SYSTEMATIC) {
        double offset = (u_override >= 0.0) ? u_override : orc_uniform(seed, 0u, stage, 0u, PURP_RESAMPLE, 0);
        i64 start = 1;
        for (i64 i = 1; i <= n_out; ++i) {
            double threshold = ((double)(i - 1) + offset) / n;
            i64 found = 0;
            for (i64 j = start; j <= N; ++j)
                if (cum[j - 1] > threshold) { found = j; break; }
            if (found == 0) found = N;
            idx[i - 1] = found;
            start = found;
        }
    } else if (method == RESAMPLE_MULTINOMIAL) {
        /* findfirst(x -> offset[i] < x, cumulative): first exceedance == upper bound on the running max */
        double *m = (double *)malloc(sizeof(double) * (size_t)N);
        double mx = -INFINITY;
        for (i64 i = 0; i < N; ++i) { if (cum[i] > mx) mx = cum[i]; m[i] = mx; }
        for (i64 i = 0; i < n_out; ++i) {
            double off = orc_uniform(seed, (uint32_t)i, stage, 1u, PURP_RESAMPLE, 0);
            i64 lo = 0, hi = N; /* first k with m[k] > off */
            while (lo < hi) { i64 mid = (lo + hi) >> 1; if (m[mid] > off) hi = mid; else lo = mid + 1; }
            idx[i] = (lo >= N) ? N : lo + 1;
        }
        free(m);
    } else {
        free(x); if (!cum_out) free(cum);
        return 2;
    }
    free(x); if (!cum_out) free(cum);
    return 0;
}

/* gather all columns by ancestor and reset weights (smc_main.jl:440-445, particle.jl:378-383) */
ORC_API void orc_gather(const double *src, double *dst, i64 N, int d, const i64 *idx)
{
    for (int j = 0; j < d + 4; ++j) {
        const double *s = COL(src, N, j); double *t = COL(dst, N, j);
        for (i64 i = 0; i < N; ++i) t[i] = s[idx[i] - 1];
    }
    double *w = COL(dst, N, C_WEIGHT(d));
    for (i64 i = 0; i < N; ++i) w[i] = 1.0;
}

/* ------------------------------------------------------------------------------------------ */
/* Step-size adaptation (src/smc_main.jl:453-455)                                             */
/* ------------------------------------------------------------------------------------------ */
ORC_API double orc_update_c(double c, double accept, double target)
{
    double e = orc_exp(16.0 * (accept - target));
    return c * (0.95 + 0.10 * e / (1.0 + e));
}

/* ------------------------------------------------------------------------------------------ */
/* Moments (src/particle.jl:481-486,526-532): lanes = 32, R = 64 canonical order               */
/* ------------------------------------------------------------------------------------------ */
#define M_LANES 32
#define M_R 64
#define M2_CH 512
#define MAX_D_MOM 64
struct wx_ctx { const double *w, *x, *y; double mx, my, sw; };
ORC_API void orc_moments(const double *cloud, i64 N, int d, double *mean, double *cov)
{
    const double *w = COL(cloud, N, C_WEIGHT(d));
    i64 tile = (i64)M_LANES * M_R;
    i64 P = canon_ntiles_pow2(N, M_LANES, M_R);
    i64 nt = (N + tile - 1) / tile;
    int nq = 1 + d;
    double *tiles = (double *)calloc((size_t)(P * nq), sizeof(double));
    double lane[M_LANES];
    /* pass 1: Sw and sum_i w_i x_ik, accumulated with fma in the canonical order */
    for (int q = 0; q < nq; ++q) {
        const double *x = (q == 0) ? NULL : COL(cloud, N, q - 1);
        for (i64 t = 0; t < nt; ++t) {
            for (int l = 0; l < M_LANES; ++l) {
                double acc = 0.0;
                for (int r = 0; r < M_R; ++r) {
                    i64 i = t * tile + (i64)r * M_LANES + l;
                    if (i < N) acc = x ? FMA(w[i], x[i], acc) : acc + w[i];
                }
                lane[l] = acc;
            }
            tiles[(size_t)q * P + t] = tree_inplace(lane, M_LANES);
        }
    }
    double sw = tree_inplace(tiles, P);
    for (int k = 0; k < d; ++k) mean[k] = tree_inplace(tiles + (size_t)(k + 1) * P, P) / sw;
    free(tiles);
    /* pass 2: centred weighted scatter, lower triangle, then / Sw; symmetric by construction.
     * Canonical order: inside each chunk of M2_CH consecutive particles lane l accumulates particles
     * l, l + 32, ... sequentially with fma, then the adjacent-pair tree over the 32 lanes, then over
     * chunks (zero padded to a power of two). */
    i64 nch = (N + M2_CH - 1) / M2_CH;
    i64 Pc = next_pow2(nch < 1 ? 1 : nch);
    double *tl = (double *)calloc((size_t)Pc, sizeof(double));
    for (int a = 0; a < d; ++a) {
        const double *xa = COL(cloud, N, a);
        for (int b = 0; b <= a; ++b) {
            const double *xb = COL(cloud, N, b);
            memset(tl, 0, sizeof(double) * (size_t)Pc);
            for (i64 c = 0; c < nch; ++c) {
                for (int l = 0; l < M_LANES; ++l) {
                    double acc = 0.0;
                    for (int r = 0; r < M2_CH / M_LANES; ++r) {
                        i64 i = c * M2_CH + (i64)r * M_LANES + l;
                        if (i < N) {
                            double da = xa[i] - mean[a], db = xb[i] - mean[b];
                            acc = FMA(w[i] * da, db, acc);
                        }
                    }
                    lane[l] = acc;
                }
                tl[c] = tree_inplace(lane, M_LANES);
            }
            double v = tree_inplace(tl, Pc) / sw;
            cov[(size_t)a * d + b] = v;
            cov[(size_t)b * d + a] = v;
        }
    }
    free(tl);
}

/* One-pass form of the same moments, as the engine's fused stage computes them: with a shift x0 (the parameter vector of
 * particle 0 -- any point of the cloud's support keeps the sums well conditioned)
 *   Sw = sum w,  m_k = sum w (x_k - x0_k),  C_ab = sum (w (x_a - x0_a)) (x_b - x0_b)
 *   mean_k = x0_k + m_k / Sw,  cov_ab = C_ab / Sw - (m_a / Sw)(m_b / Sw)
 * Equal to orc_moments (the reference's two-pass form, src/particle.jl:481-532) up to rounding (~1e-15 relative).
 * Canonical order of every sum (the engine runs them as one FP64 tensor-core SYRK, whose accumulator is a sequential fma
 * chain over the particles): sub-chunks of M1P_SC consecutive particles accumulated one particle after the other in
 * ascending order, then the adjacent-pair tree over sub-chunks (zero padded to a power of two). */
#define M1P_SC 256
ORC_API void orc_moments_shifted(const double *cloud, i64 N, int d, const double *shift, double *mean, double *cov)
{
    const double *w = COL(cloud, N, C_WEIGHT(d));
    i64 nch = (N + M1P_SC - 1) / M1P_SC;
    i64 Pc = next_pow2(nch < 1 ? 1 : nch);
    double *tl = (double *)calloc((size_t)Pc, sizeof(double));
    int E = d * (d + 1) / 2, nq = 1 + d + E;
    double *sums = (double *)malloc(sizeof(double) * (size_t)nq);
    for (int q = 0; q < nq; ++q) {
        int a = 0, b = 0;
        if (q > d) { int e = q - 1 - d; while ((a + 1) * (a + 2) / 2 <= e) ++a; b = e - a * (a + 1) / 2; }
        else if (q >= 1) a = q - 1;
        const double *xa = COL(cloud, N, a), *xb = COL(cloud, N, b);
        memset(tl, 0, sizeof(double) * (size_t)Pc);
        for (i64 c = 0; c < nch; ++c) {
            double acc = 0.0;
            i64 i1 = (c + 1) * M1P_SC < N ? (c + 1) * M1P_SC : N;
            for (i64 i = c * M1P_SC; i < i1; ++i) {
                if (q == 0) acc = acc + w[i];
                else if (q <= d) acc = FMA(w[i], xa[i] - shift[a], acc);
                else acc = FMA(w[i] * (xa[i] - shift[a]), xb[i] - shift[b], acc);
            }
            tl[c] = acc;
        }
        sums[q] = tree_inplace(tl, Pc);
    }
    double sw = sums[0];
    double e[MAX_D_MOM];
    for (int k = 0; k < d; ++k) { e[k] = sums[1 + k] / sw; mean[k] = shift[k] + e[k]; }
    for (int a = 0; a < d; ++a)
        for (int b = 0; b <= a; ++b) {
            double v = FMA(-e[a], e[b], sums[1 + d + a * (a + 1) / 2 + b] / sw);
            cov[(size_t)a * d + b] = v;
            cov[(size_t)b * d + a] = v;
        }
    free(tl); free(sums);
}

/* mean of the accept column (update_acceptance_rate!, particle.jl:466-468).  Every entry is (integer count) / n_free, so
 * the engine sums the integer counts exactly (order-free) and divides once: mean = (sum_i count_i / n_free) / N -- equal to
 * the reference's floating-point mean of the column up to rounding. */
ORC_API double orc_mean_accept(const double *cloud, i64 N, int d, int n_free)
{
    const double *a = COL(cloud, N, C_ACCEPT(d));
    long long tot = 0;
    for (i64 i = 0; i < N; ++i) tot += llround(a[i] * (double)n_free);
    return ((double)tot / (double)n_free) / (double)N;
}

/* Lower Cholesky, row by row; returns 0 or (1 + failing row) if not positive definite
 * (the reference would throw PosDefException from MvNormal at mutation.jl:81). */
ORC_API int orc_cholesky(const double *A, int n, double *L)
{
    memset(L, 0, sizeof(double) * (size_t)n * n);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j <= i; ++j) {
            double s = A[(size_t)i * n + j];
            for (int k = 0; k < j; ++k) s = FMA(-L[(size_t)i * n + k], L[(size_t)j * n + k], s);
            if (i == j) {
                if (!(s > 0.0)) return 1 + i;
                L[(size_t)i * n + i] = sqrt(s);
            } else {
                L[(size_t)i * n + j] = s / L[(size_t)j * n + j];
            }
        }
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Model: priors (ModelConstructors.prior / update! semantics) + likelihoods     */
/* ------------------------------------------------------------------------------------------ */
enum { PRIOR_NORMAL = 0, PRIOR_UNIFORM = 1, PRIOR_GAMMA = 2, PRIOR_ROOT_INV_GAMMA = 3, PRIOR_BETA = 4, PRIOR_INV_GAMMA = 5 };
enum { LIK_NONE = 0, LIK_GAUSSREG = 1, LIK_AS_DSGE = 2 };
#define MAX_D 64
#define MAX_EQ 8

typedef struct {
    int neq, k, stride, coef_off, sig_off; /* sig_off < 0: sigma known */
    double T[MAX_EQ], qscale[MAX_EQ], rss[MAX_EQ], cT[MAX_EQ], logs[MAX_EQ], inv_s2[MAX_EQ];
    double *bhat[MAX_EQ]; /* k */
    double *U[MAX_EQ];    /* k x k upper, row-major full */
} gaussreg;

typedef struct {
    int d;
    int fixed[MAX_D], kind[MAX_D];
    double lo[MAX_D], hi[MAX_D], p1[MAX_D], p2[MAX_D], cst[MAX_D], a1[MAX_D], a2[MAX_D];
    int all_normal; double cst_sum;   /* every parameter free with a Normal prior: fused log-prior (see orc_logprior) */
    int lik_kind[2];
    gaussreg gr[2];
    struct { int T, npre; double *data; } as[2];   /* An-Schorfheide DSGE likelihood (as_model.c) */
} orc_model;

ORC_API orc_model *orc_model_create(int d)
{
    if (d > MAX_D) return NULL;
    orc_model *m = (orc_model *)calloc(1, sizeof(orc_model));
    m->d = d;
    return m;
}
ORC_API void orc_model_free(orc_model *m)
{
    if (!m) return;
    for (int s = 0; s < 2; ++s)
        for (int e = 0; e < MAX_EQ; ++e) { free(m->gr[s].bhat[e]); free(m->gr[s].U[e]); }
    free(m->as[0].data); free(m->as[1].data);
    free(m);
}

#define HALF_LOG_2PI 0.91893853320467274178
/* Per-parameter constants are computed with the host libm (same glibc as the engine's host side). */
ORC_API int orc_model_set_params(orc_model *m, const i32 *fixed, const double *lo, const double *hi,
                                 const i32 *kind, const double *p1, const double *p2)
{
    for (int k = 0; k < m->d; ++k) {
        m->fixed[k] = fixed[k]; m->lo[k] = lo[k]; m->hi[k] = hi[k];
        m->kind[k] = kind[k]; m->p1[k] = p1[k]; m->p2[k] = p2[k];
        double a = p1[k], b = p2[k];
        switch (kind[k]) {
        case PRIOR_NORMAL:   /* Normal(mu = a, sigma = b) */
            m->cst[k] = -log(b) - HALF_LOG_2PI; m->a1[k] = 1.0 / b; m->a2[k] = 0.0; break;
        case PRIOR_UNIFORM:  /* Uniform(a, b) */
            m->cst[k] = -log(b - a); m->a1[k] = 0.0; m->a2[k] = 0.0; break;
        case PRIOR_GAMMA:    /* Gamma(shape = a, scale = b) */
            m->cst[k] = -lgamma(a) - a * log(b); m->a1[k] = a - 1.0; m->a2[k] = 1.0 / b; break;
        case PRIOR_ROOT_INV_GAMMA: /* RootInverseGamma(nu = a, tau = b) */
            m->cst[k] = log(2.0) - lgamma(0.5 * a) + 0.5 * a * log(0.5 * a * b * b);
            m->a1[k] = 0.5 * (a + 1.0); m->a2[k] = 0.5 * a * b * b; break;
        case PRIOR_BETA:     /* Beta(a, b) */
            m->cst[k] = lgamma(a + b) - lgamma(a) - lgamma(b); m->a1[k] = a - 1.0; m->a2[k] = b - 1.0; break;
        case PRIOR_INV_GAMMA: /* InverseGamma(shape = a, scale = b) */
            m->cst[k] = a * log(b) - lgamma(a); m->a1[k] = a + 1.0; m->a2[k] = b; break;
        default: return 3;
        }
    }
    m->all_normal = 1; m->cst_sum = 0.0;
    for (int k = 0; k < m->d; ++k) {
        if (m->fixed[k] || m->kind[k] != PRIOR_NORMAL) m->all_normal = 0;
        m->cst_sum = m->cst_sum + (m->fixed[k] ? 0.0 : m->cst[k]);
    }
    return 0;
}

static inline double logpdf1(const orc_model *m, int k, double x)
{
    switch (m->kind[k]) {
    case PRIOR_NORMAL: { double z = (x - m->p1[k]) * m->a1[k]; return FMA(-0.5 * z, z, m->cst[k]); }
    case PRIOR_UNIFORM: return (x >= m->p1[k] && x <= m->p2[k]) ? m->cst[k] : -INFINITY;
    case PRIOR_GAMMA:
        if (!(x > 0.0)) return (x == 0.0 && m->a1[k] == 0.0) ? m->cst[k] : -INFINITY;
        return FMA(m->a1[k], orc_log(x), m->cst[k]) - x * m->a2[k];
    case PRIOR_ROOT_INV_GAMMA: {
        if (!(x > 0.0)) return -INFINITY;
        double x2 = x * x;
        return FMA(-m->a1[k], orc_log(x2), m->cst[k]) - m->a2[k] / x2; }
    case PRIOR_BETA:
        if (!(x > 0.0 && x < 1.0)) return -INFINITY;
        return FMA(m->a2[k], orc_log(1.0 - x), FMA(m->a1[k], orc_log(x), m->cst[k]));
    case PRIOR_INV_GAMMA:
        if (!(x > 0.0)) return -INFINITY;
        return FMA(-m->a1[k], orc_log(x), m->cst[k]) - m->a2[k] / x;
    }
    return NAN;
}
/* sum over FREE parameters in index order */
ORC_API double orc_logprior(const orc_model *m, const double *theta)
{
    if (m->all_normal) {   /* same sum with the squares accumulated first: -0.5 sum_k z_k^2 + sum_k cst_k */
        double acc = 0.0;
        for (int k = 0; k < m->d; ++k) { double z = (theta[k] - m->p1[k]) * m->a1[k]; acc = FMA(z, z, acc); }
        return FMA(-0.5, acc, m->cst_sum);
    }
    double lp = 0.0;
    for (int k = 0; k < m->d; ++k) if (!m->fixed[k]) lp = lp + logpdf1(m, k, theta[k]);
    return lp;
}
static inline int in_bounds(const orc_model *m, const double *theta)
{
    for (int k = 0; k < m->d; ++k)
        if (!m->fixed[k] && !(theta[k] >= m->lo[k] && theta[k] <= m->hi[k])) return 0;
    return 1;
}

/* Gaussian regression family in centred sufficient-statistic form.  Equation e has coefficients
 * theta[coef_off + e*stride + (0..k-1)] and (optionally) sigma = theta[sig_off + e*stride]:
 *   ll_e = cT_e - T_e*log(sigma_e) - 0.5*qscale_e*(rss_e + |U_e (b - bhat_e)|^2) / sigma_e^2
 * eq layout in `eqdata` per equation: [T, qscale, rss, sigma_fixed, bhat(k), U(k*k row-major upper)] */
ORC_API int orc_model_set_gaussreg(orc_model *m, int slot, int neq, int k, int stride, int coef_off,
                                   int sig_off, const double *eqdata)
{
    if (neq > MAX_EQ || slot < 0 || slot > 1) return 3;
    gaussreg *g = &m->gr[slot];
    g->neq = neq; g->k = k; g->stride = stride; g->coef_off = coef_off; g->sig_off = sig_off;
    size_t per = 4 + (size_t)k + (size_t)k * k;
    for (int e = 0; e < neq; ++e) {
        const double *p = eqdata + per * e;
        g->T[e] = p[0]; g->qscale[e] = p[1]; g->rss[e] = p[2];
        g->cT[e] = -p[0] * HALF_LOG_2PI;
        if (sig_off < 0) { g->logs[e] = log(p[3]); g->inv_s2[e] = 1.0 / (p[3] * p[3]); }
        free(g->bhat[e]); free(g->U[e]);
        g->bhat[e] = (double *)malloc(sizeof(double) * (size_t)k);
        g->U[e] = (double *)malloc(sizeof(double) * (size_t)k * k);
        memcpy(g->bhat[e], p + 4, sizeof(double) * (size_t)k);
        memcpy(g->U[e], p + 4 + k, sizeof(double) * (size_t)k * k);
    }
    m->lik_kind[slot] = LIK_GAUSSREG;
    return 0;
}
static double gaussreg_ll(const gaussreg *g, const double *theta)
{
    double ll = 0.0;
    double dl[MAX_D];
    for (int e = 0; e < g->neq; ++e) {
        const double *b = theta + g->coef_off + e * g->stride;
        int k = g->k;
        for (int j = 0; j < k; ++j) dl[j] = b[j] - g->bhat[e][j];
        double q = g->rss[e];
        for (int i = 0; i < k; ++i) {
            double r = 0.0;
            for (int j = i; j < k; ++j) r = FMA(g->U[e][(size_t)i * k + j], dl[j], r);
            q = FMA(r, r, q);
        }
        double logs, inv_s2;
        if (g->sig_off >= 0) {
            double s = theta[g->sig_off + e * g->stride];
            if (!(s > 0.0)) return -INFINITY;  /* log of a non-positive sigma: DomainError => -Inf (mutation.jl:112-121) */
            logs = orc_log(s);
            inv_s2 = 1.0 / (s * s);
        } else { logs = g->logs[e]; inv_s2 = g->inv_s2[e]; }
        double le = FMA(-g->T[e], logs, g->cT[e]) - (0.5 * g->qscale[e] * q) * inv_s2;
        ll = ll + le;
    }
    return ll;
}
/* An-Schorfheide DSGE likelihood (config C4): data 3 x T column-major, first npre periods unscored */
double orc_as_loglik(const double *th, const double *data, int T, int npre);   /* as_model.c */
ORC_API int orc_model_set_as(orc_model *m, int slot, int T, int npre, const double *data)
{
    if (slot < 0 || slot > 1 || m->d != 16 || T < 1 || npre < 0) return 3;
    free(m->as[slot].data);
    m->as[slot].data = (double *)malloc(sizeof(double) * 3 * (size_t)T);
    memcpy(m->as[slot].data, data, sizeof(double) * 3 * (size_t)T);
    m->as[slot].T = T; m->as[slot].npre = npre;
    m->lik_kind[slot] = LIK_AS_DSGE;
    return 0;
}
ORC_API double orc_loglik(const orc_model *m, int slot, const double *theta)
{
    if (m->lik_kind[slot] == LIK_GAUSSREG) return gaussreg_ll(&m->gr[slot], theta);
    if (m->lik_kind[slot] == LIK_AS_DSGE) return orc_as_loglik(theta, m->as[slot].data, m->as[slot].T, m->as[slot].npre);
    return NAN;
}

/* Direct (per-observation) forms of the reference's example likelihoods -- used by the tests to
 * validate the sufficient-statistic form above and the golden (theta -> loglh) rows.
 *  test/modelsetup.jl:119-138 (3-equation model; data n_eq x T, X n_eq x T, column-major like Julia) */
ORC_API double orc_loglik_lineq_direct(const double *p, const double *data, const double *X, int neq, int T)
{
    double det = 1.0;
    for (int i = 0; i < neq; ++i) det *= p[3 * i + 2] * p[3 * i + 2];
    double term1 = -(double)neq / 2.0 * log(2.0 * M_PI) - 0.5 * log(det);
    double lp = 0.0;
    for (int t = 0; t < T; ++t) {
        double q = 0.0;
        for (int i = 0; i < neq; ++i) {
            double e = data[(size_t)t * neq + i] - p[3 * i] - p[3 * i + 1] * X[(size_t)t * neq + i];
            q += e * (1.0 / (p[3 * i + 2] * p[3 * i + 2])) * e;
        }
        lp += term1 - 0.5 * q;
    }
    return lp;
}
/* examples/regression_model/estimate_regression.jl:46-53 generalised to k regressors (X is T x k row-major) */
ORC_API double orc_loglik_linreg_direct(const double *beta, const double *y, const double *X, int T, int k, double sigma2)
{
    double term1 = -((double)T / 2.0) * log(2.0 * M_PI) - ((double)T / 2.0) * log(sigma2);
    double ss = 0.0;
    for (int t = 0; t < T; ++t) {
        double e = y[t];
        for (int j = 0; j < k; ++j) e -= beta[j] * X[(size_t)t * k + j];
        ss += e * e;
    }
    return term1 - (1.0 / (2.0 * sigma2)) * ss;
}

/* ------------------------------------------------------------------------------------------ */
/* Proposal preparation: blocks -> per-block Cholesky factors                                  */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    int n_blocks;
    int d;
    int bsize[MAX_D];
    int member[MAX_D][MAX_D];   /* sorted ascending parameter indices (0-based) */
    int mfree[MAX_D][MAX_D];    /* matching indices into the free set */
    double *Lc[MAX_D];          /* c * L_b, bsize x bsize lower, row-major */
    double *sd[MAX_D];          /* sqrt(diag Sigma_b) (WITHOUT c, helpers.jl:146) */
    double *mu[MAX_D];          /* theta_bar restricted to the block */
    double logdet_c[MAX_D];     /* log det(c^2 Sigma_b) = 2 sum log(c L_ii) */
    double c;
} orc_proposal;

ORC_API void orc_proposal_free(orc_proposal *p)
{
    if (!p) return;
    for (int b = 0; b < MAX_D; ++b) { free(p->Lc[b]); free(p->sd[b]); free(p->mu[b]); }
    free(p);
}
static void isort2(int *a, int *b, int n)
{
    for (int i = 1; i < n; ++i) {
        int x = a[i], y = b[i], j = i - 1;
        while (j >= 0 && a[j] > x) { a[j + 1] = a[j]; b[j + 1] = b[j]; --j; }
        a[j + 1] = x; b[j + 1] = y;
    }
}
/* mean_fr (n_free), cov_fr (n_free x n_free) as at smc_main.jl:462-465; blocks as produced by
 * generate_free_blocks / generate_all_blocks (0-based here).  Within a block the engine works in
 * ascending parameter order -- a relabelling of the proposal's normal draws that leaves the proposal
 * distribution N(theta_b, c^2 Sigma_bb) unchanged (the reference keeps the shuffled order). */
ORC_API orc_proposal *orc_proposal_create(int d, int n_free, const double *mean_fr, const double *cov_fr,
                                          int n_blocks, const i32 *block_sizes, const i32 *blocks_free,
                                          const i32 *blocks_all, double c, int *status)
{
    orc_proposal *p = (orc_proposal *)calloc(1, sizeof(orc_proposal));
    p->n_blocks = n_blocks; p->d = d; p->c = c;
    int pos = 0;
    *status = 0;
    for (int b = 0; b < n_blocks; ++b) {
        int n = block_sizes[b];
        p->bsize[b] = n;
        for (int i = 0; i < n; ++i) { p->member[b][i] = blocks_all[pos + i]; p->mfree[b][i] = blocks_free[pos + i]; }
        pos += n;
        isort2(p->member[b], p->mfree[b], n);
        double *S = (double *)malloc(sizeof(double) * (size_t)n * n);
        double *L = (double *)malloc(sizeof(double) * (size_t)n * n);
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j)   /* (R + R') / 2, smc_main.jl:462 (exact no-op for symmetric input) */
                S[(size_t)i * n + j] = (cov_fr[(size_t)p->mfree[b][i] * n_free + p->mfree[b][j]] +
                                        cov_fr[(size_t)p->mfree[b][j] * n_free + p->mfree[b][i]]) / 2.0;
        int st = orc_cholesky(S, n, L);
        if (st) { *status = 4; free(S); free(L); orc_proposal_free(p); return NULL; }
        p->Lc[b] = (double *)malloc(sizeof(double) * (size_t)n * n);
        p->sd[b] = (double *)malloc(sizeof(double) * (size_t)n);
        p->mu[b] = (double *)malloc(sizeof(double) * (size_t)n);
        double ld = 0.0;
        for (int i = 0; i < n; ++i) {
            for (int j = 0; j < n; ++j) p->Lc[b][(size_t)i * n + j] = c * L[(size_t)i * n + j];
            p->sd[b][i] = sqrt(S[(size_t)i * n + i]);
            p->mu[b][i] = mean_fr[p->mfree[b][i]];
            ld = ld + orc_log(p->Lc[b][(size_t)i * n + i]);
        }
        p->logdet_c[b] = 2.0 * ld;
        free(S); free(L);
    }
    return p;
}

/* Fisher-Yates shuffle of 0..n-1 driven by Philox (generate_free_blocks, helpers.jl:215-231) */
ORC_API void orc_generate_blocks(int n_free, int n_blocks, uint64_t seed, uint32_t stage, i32 *perm, i32 *sizes)
{
    for (int i = 0; i < n_free; ++i) perm[i] = i;
    for (int i = n_free - 1; i >= 1; --i) {
        double u = orc_uniform(seed, (uint32_t)i, stage, 0u, PURP_BLOCKS, 0);
        int j = (int)(u * (double)(i + 1));
        int t = perm[i]; perm[i] = perm[j]; perm[j] = t;
    }
    int sub = (n_free + n_blocks - 1) / n_blocks;
    int last = n_free - sub * (n_blocks - 1);
    for (int b = 0; b < n_blocks; ++b) sizes[b] = (b < n_blocks - 1) ? sub : last;
}

/* Gaussian log-density N(x; mu, c^2 Sigma_b) through the scaled Cholesky factor */
static double mvn_logpdf(const double *Lc, int n, double logdet, const double *x, const double *mu)
{
    double y[MAX_D];
    double q = 0.0;
    for (int i = 0; i < n; ++i) {
        double s = x[i] - mu[i];
        for (int j = 0; j < i; ++j) s = FMA(-Lc[(size_t)i * n + j], y[j], s);
        y[i] = s * (1.0 / Lc[(size_t)i * n + i]);      /* reciprocal diagonal (the device precomputes it once per stage) */
        q = FMA(y[i], y[i], q);
    }
    return -0.5 * (((double)n * (2.0 * HALF_LOG_2PI) + logdet) + q);
}
/* compute_proposal_densities (helpers.jl:128-164) */
ORC_API void orc_proposal_densities(const double *Lc, const double *sd, const double *mu, int n, double logdet,
                                    const double *para_draw, const double *para_subset, double alpha,
                                    double *q0_out, double *q1_out)
{
    double q0 = alpha * orc_exp(mvn_logpdf(Lc, n, logdet, para_subset, para_draw));
    double q1 = alpha * orc_exp(mvn_logpdf(Lc, n, logdet, para_draw, para_subset));
    double ind = 1.0;
    for (int i = 0; i < n; ++i) {
        /* zstat = (theta_i - vartheta_i) / Sigma_ii^(1/2); ind_pdf *= exp(-zstat^2 / 2) / (Sigma_ii^(1/2) sqrt(2 pi)),
         * helpers.jl:145-148, with the reciprocal standard deviation formed once */
        double isd = 1.0 / sd[i];
        double zs = (para_subset[i] - para_draw[i]) * isd;
        ind = (ind * (isd * 0x1.9884533d43651p-2)) * orc_exp(-0.5 * (zs * zs));
    }
    q0 += (1.0 - alpha) / 2.0 * ind;
    q1 += (1.0 - alpha) / 2.0 * ind;
    q0 += (1.0 - alpha) / 2.0 * orc_exp(mvn_logpdf(Lc, n, logdet, para_subset, mu));
    q1 += (1.0 - alpha) / 2.0 * orc_exp(mvn_logpdf(Lc, n, logdet, para_draw, mu));
    q0 = orc_log(q0);
    q1 = orc_log(q1);
    if (q0 == INFINITY && q1 == INFINITY) q0 = 0.0;
    *q0_out = q0; *q1_out = q1;
}

/* ------------------------------------------------------------------------------------------ */
/* Mutation (src/mutation.jl:56-138) for one particle                                          */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    double phi_n, phi_n1, alpha;
    int n_mh_steps, n_free, has_old;
    uint64_t seed;
    uint32_t stage;
} mut_cfg;

static void mutate_one(const orc_model *m, const orc_proposal *pr, const mut_cfg *cfg, double *cloud, i64 N, i64 i,
                       i64 global_index)
{
    int d = m->d;
    double para[MAX_D], pnew[MAX_D], z[MAX_D], draw[MAX_D], sub[MAX_D];
    for (int k = 0; k < d; ++k) para[k] = COL(cloud, N, k)[i];
    double like = COL(cloud, N, C_LOGLH(d))[i];
    double logprior = COL(cloud, N, C_LOGPRIOR(d))[i];
    double like_prev = COL(cloud, N, C_OLDLOGLH(d))[i];
    double accept = 0.0;
    uint32_t gp = (uint32_t)global_index;
    for (int step = 0; step < cfg->n_mh_steps; ++step) {
        for (int b = 0; b < pr->n_blocks; ++b) {
            uint32_t sb = (uint32_t)(step * pr->n_blocks + b);
            int n = pr->bsize[b];
            const int *mem = pr->member[b];
            uint32_t r4[4];
            rng4(cfg->seed, gp, cfg->stage, sb, PURP_STEP, r4);
            double step_prob = u01(r4[0], r4[1]);
            double u_mix = u01(r4[2], r4[3]);
            /* normals are indexed by PARAMETER index: Philox block (sb << 8) | (index >> 2) gives the four
             * proposal normals of parameters 4q .. 4q+3 */
            for (int q = 0; q < n; ++q) {
                int k = mem[q];
                double zz[4];
                uint32_t rr[4];
                rng4(cfg->seed, gp, cfg->stage, (sb << 8) | (uint32_t)(k >> 2), PURP_NORMAL, rr);
                normal_quad(rr, zz);
                z[q] = zz[k & 3];
            }
            for (int q = 0; q < n; ++q) sub[q] = para[mem[q]];
            /* mvnormal_mixture_draw (helpers.jl:87-100) */
            int comp = 1;
            if (cfg->alpha < 1.0) {
                if (u_mix < cfg->alpha) comp = 1;
                else if (u_mix < cfg->alpha + (1.0 - cfg->alpha) / 2.0) comp = 2;
                else comp = 3;
            }
            const double *Lc = pr->Lc[b];
            for (int q = 0; q < n; ++q) {
                double s = 0.0;
                if (comp == 2) {
                    /* Diagonal(diag(c^2 Sigma)): std_i = c * sqrt(Sigma_ii) */
                    s = (pr->c * pr->sd[b][q]) * z[q];
                } else {
                    for (int j = 0; j <= q; ++j) s = FMA(Lc[(size_t)q * n + j], z[j], s);
                }
                double base = (comp == 3) ? pr->mu[b][q] : sub[q];
                draw[q] = base + s;
            }
            double q0, q1;
            orc_proposal_densities(Lc, pr->sd[b], pr->mu[b], n, pr->logdet_c[b], draw, sub, cfg->alpha, &q0, &q1);
            for (int k = 0; k < d; ++k) pnew[k] = para[k];
            for (int q = 0; q < n; ++q) pnew[mem[q]] = draw[q];
            double prior_new, like_new, like_old_data;
            if (!in_bounds(m, pnew)) {
                prior_new = like_new = like_old_data = -INFINITY;     /* ParamBoundsError, mutation.jl:112-121 */
            } else {
                prior_new = orc_logprior(m, pnew);
                like_new = orc_loglik(m, 0, pnew);
                if (like_new == -INFINITY) prior_new = -INFINITY;     /* mutation.jl:102-104 */
                like_old_data = cfg->has_old ? orc_loglik(m, 1, pnew) : 0.0;   /* :106 */
            }
            double eta = orc_exp(((cfg->phi_n * (like_new - like) + (1.0 - cfg->phi_n) * (like_old_data - like_prev)) +
                                  (prior_new - logprior)) + (q0 - q1));
            if (step_prob < eta) {
                for (int k = 0; k < d; ++k) para[k] = pnew[k];
                like = like_new; logprior = prior_new; like_prev = like_old_data;
                accept += (double)n;
            }
        }
    }
    for (int k = 0; k < d; ++k) COL(cloud, N, k)[i] = para[k];
    COL(cloud, N, C_LOGLH(d))[i] = like;
    COL(cloud, N, C_LOGPRIOR(d))[i] = logprior;
    COL(cloud, N, C_OLDLOGLH(d))[i] = like_prev;
    COL(cloud, N, C_ACCEPT(d))[i] = accept / (double)cfg->n_free;
}

/* all particles; static contiguous chunks = the @distributed split (smc_main.jl:471-476).
 * index0 = global index of local particle 0 (shards). */
ORC_API void orc_mutate(const orc_model *m, const orc_proposal *pr, double *cloud, i64 N, i64 index0,
                        double phi_n, double phi_n1, double alpha, int n_mh_steps, int n_free, int has_old,
                        uint64_t seed, uint32_t stage, int nthreads)
{
    mut_cfg cfg = {phi_n, phi_n1, alpha, n_mh_steps, n_free, has_old, seed, stage};
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(static) num_threads(nthreads)
#endif
    for (i64 i = 0; i < N; ++i) mutate_one(m, pr, &cfg, cloud, N, i, index0 + i);
}

ORC_API int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------------ */
/* Stage-0 evaluators (src/initialization.jl:23-63,153-186)                                    */
/* ------------------------------------------------------------------------------------------ */
/* initialize_likelihoods!: old_loglh <- loglh; loglh, logprior re-evaluated on the new data */
ORC_API void orc_initialize_likelihoods(const orc_model *m, double *cloud, i64 N)
{
    int d = m->d;
    double th[MAX_D];
    for (i64 i = 0; i < N; ++i) {
        for (int k = 0; k < d; ++k) th[k] = COL(cloud, N, k)[i];
        COL(cloud, N, C_OLDLOGLH(d))[i] = COL(cloud, N, C_LOGLH(d))[i];
        COL(cloud, N, C_LOGLH(d))[i] = orc_loglik(m, 0, th);
        COL(cloud, N, C_LOGPRIOR(d))[i] = orc_logprior(m, th);
    }
}
/* evaluate loglh/logprior of given draws (draw_likelihood, initialization.jl:129-139) */
ORC_API void orc_evaluate(const orc_model *m, double *cloud, i64 N)
{
    int d = m->d;
    double th[MAX_D];
    for (i64 i = 0; i < N; ++i) {
        for (int k = 0; k < d; ++k) th[k] = COL(cloud, N, k)[i];
        COL(cloud, N, C_LOGLH(d))[i] = in_bounds(m, th) ? orc_loglik(m, 0, th) : -INFINITY;
        COL(cloud, N, C_LOGPRIOR(d))[i] = orc_logprior(m, th);
    }
}

/* initial_draw! / one_draw (src/initialization.jl:23-119): every free parameter from its prior (redrawn
 * until strictly inside its valuebounds, ModelConstructors `rand(parameters, 1)`), fixed parameters keep
 * their value; the whole vector is redrawn while the log-likelihood is not finite (:43-60).  The
 * reference's dSFMT stream is not reproducible: randomness is a per-particle sequential Philox stream,
 * counter = (global particle, 0, call number, PURP_INIT), one block per uniform / normal.
 * Gamma(shape, 1): Marsaglia & Tsang (2000), shape < 1 through shape + 1 and U^(1/shape). */
typedef struct { uint64_t seed; uint32_t gp, ctr; } draw_rng;
static double dr_uniform(draw_rng *r) { uint32_t x[4]; rng4(r->seed, r->gp, 0u, r->ctr++, PURP_INIT, x); return u01(x[0], x[1]); }
static double dr_uniform_open0(draw_rng *r) { uint32_t x[4]; rng4(r->seed, r->gp, 0u, r->ctr++, PURP_INIT, x); return u01_open0(x[0], x[1]); }
static double dr_normal(draw_rng *r) { uint32_t x[4]; double z0, z1; rng4(r->seed, r->gp, 0u, r->ctr++, PURP_INIT, x); normal_pair(x, &z0, &z1); return z0; }
static double dr_gamma(draw_rng *r, double shape)
{
    int boost = shape < 1.0;
    double a = boost ? shape + 1.0 : shape;
    double d = a - 1.0 / 3.0;
    double c = 1.0 / sqrt(9.0 * d);
    double g = NAN;
    for (int it = 0; it < 256; ++it) {
        double x = dr_normal(r);
        double t = 1.0 + c * x;
        if (!(t > 0.0)) continue;
        double v = (t * t) * t;
        double u = dr_uniform_open0(r);
        if (orc_log(u) < ((0.5 * x) * x + d) - d * v + d * orc_log(v)) { g = d * v; break; }
    }
    if (boost) {
        double u = dr_uniform_open0(r);
        g = g * orc_exp(orc_log(u) / shape);
    }
    return g;
}
static double dr_prior(const orc_model *m, draw_rng *r, int k)
{
    double p1 = m->p1[k], p2 = m->p2[k];
    switch (m->kind[k]) {
    case PRIOR_NORMAL: return p1 + p2 * dr_normal(r);
    case PRIOR_UNIFORM: return p1 + (p2 - p1) * dr_uniform(r);
    case PRIOR_GAMMA: return p2 * dr_gamma(r, p1);
    case PRIOR_ROOT_INV_GAMMA: return sqrt(m->a2[k] / dr_gamma(r, 0.5 * p1));
    case PRIOR_BETA: { double x = dr_gamma(r, p1); double y = dr_gamma(r, p2); return x / (x + y); }
    case PRIOR_INV_GAMMA: return p2 / dr_gamma(r, p1);
    }
    return NAN;
}
/* returns the number of particles that found no finite log-likelihood within max_tries */
ORC_API int orc_initial_draw(const orc_model *m, double *cloud, i64 N, i64 index0, const double *fixed_values,
                             uint64_t seed, int max_tries)
{
    int d = m->d, failed = 0;
    for (i64 i = 0; i < N; ++i) {
        draw_rng r = {seed, (uint32_t)(index0 + i), 0u};
        double th[MAX_D], ll = -INFINITY, lp = -INFINITY;
        int success = 0;
        for (int attempt = 0; attempt < max_tries && !success; ++attempt) {
            for (int k = 0; k < d; ++k) {
                double x;
                if (m->fixed[k]) x = fixed_values[k];
                else {
                    x = NAN;
                    for (int it = 0; it < 1000; ++it) {
                        x = dr_prior(m, &r, k);
                        if (x > m->lo[k] && x < m->hi[k]) break;
                        x = NAN;
                    }
                }
                th[k] = x;
            }
            ll = orc_loglik(m, 0, th);
            lp = orc_logprior(m, th);
            if (!(ll - ll == 0.0)) { ll = -INFINITY; lp = -INFINITY; }
            else success = 1;
        }
        if (!success) ++failed;
        for (int k = 0; k < d; ++k) COL(cloud, N, k)[i] = th[k];
        COL(cloud, N, C_LOGLH(d))[i] = ll;
        COL(cloud, N, C_LOGPRIOR(d))[i] = lp;
        COL(cloud, N, C_OLDLOGLH(d))[i] = 0.0;
        COL(cloud, N, C_ACCEPT(d))[i] = 0.0;
        COL(cloud, N, C_WEIGHT(d))[i] = 1.0;
    }
    return failed;
}

/* ------------------------------------------------------------------------------------------ */
/* One full stage (src/smc_main.jl:377-497), fixed or adaptive schedule                        */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    /* in */
    double phi_n1, phi_n;           /* phi_n ignored when adaptive */
    double threshold_ratio, target, alpha, tempering_target, pw, log_prob_old_data;
    int n_mh_steps, n_blocks, resample_method, adaptive, has_old, nthreads;
    uint64_t seed;
    uint32_t stage;                 /* = cloud.stage_index after increment (2, 3, ...) */
    /* in/out */
    double c, accept, ess_prev;
    int resampled_last;
    i64 j; double phi_prop;
    /* out */
    double ess, sum_w, phi_out;
    int resampled, status;
} orc_stage_io;

ORC_API int orc_stage(const orc_model *m, double *cloud, double *scratch /* N*(d+5) */, i64 N,
                      const double *sched, int n_phi, orc_stage_io *io, double *inc_out, double *normw_out,
                      double *mean_out, double *cov_out)
{
    int d = m->d;
    double phi_n = io->phi_n;
    if (io->adaptive) {
        orc_solve_adaptive_phi(cloud, N, d, sched, n_phi, &io->j, &io->phi_prop, io->phi_n1, io->tempering_target,
                               io->ess_prev, io->resampled_last, &phi_n, NULL);
        io->resampled_last = 0;
    }
    io->phi_out = phi_n;
    double out[3];
    int st = orc_correct(cloud, N, d, io->phi_n1, phi_n, io->pw, io->log_prob_old_data, inc_out, normw_out, out);
    io->sum_w = out[0]; io->ess = out[1];
    if (st) { io->status = 1; return 1; }
    io->resampled = 0;
    /* shift of the one-pass moments: the parameter vector of particle 0 BEFORE selection */
    double shift[MAX_D];
    for (int k = 0; k < d; ++k) shift[k] = COL(cloud, N, k)[0];
    if (io->ess < io->threshold_ratio * (double)N) {
        double *wn = (double *)malloc(sizeof(double) * (size_t)N);
        i64 *idx = (i64 *)malloc(sizeof(i64) * (size_t)N);
        const double *w = COL(cloud, N, C_WEIGHT(d));
        for (i64 i = 0; i < N; ++i) wn[i] = w[i] / (double)N;
        orc_resample(wn, N, io->resample_method, io->seed, io->stage, -1.0, idx, NULL);
        orc_gather(cloud, scratch, N, d, idx);
        memcpy(cloud, scratch, sizeof(double) * (size_t)N * (size_t)(d + 5));
        free(wn); free(idx);
        io->resampled = 1; io->resampled_last = 1;
        if (normw_out) for (i64 i = 0; i < N; ++i) normw_out[i] = 1.0;   /* W_matrix[:, i] .= 1, smc_main.jl:445 */
    }
    io->c = orc_update_c(io->c, io->accept, io->target);
    double *mean = mean_out ? mean_out : (double *)malloc(sizeof(double) * (size_t)d);
    double *cov = cov_out ? cov_out : (double *)malloc(sizeof(double) * (size_t)d * d);
    orc_moments_shifted(cloud, N, d, shift, mean, cov);
    int n_free = 0; int freeidx[MAX_D];
    for (int k = 0; k < d; ++k) if (!m->fixed[k]) freeidx[n_free++] = k;
    double mean_fr[MAX_D];
    double *cov_fr = (double *)malloc(sizeof(double) * (size_t)n_free * n_free);
    for (int a = 0; a < n_free; ++a) {
        mean_fr[a] = mean[freeidx[a]];
        for (int b = 0; b < n_free; ++b)
            cov_fr[(size_t)a * n_free + b] = (cov[(size_t)freeidx[a] * d + freeidx[b]] + cov[(size_t)freeidx[b] * d + freeidx[a]]) / 2.0;
    }
    i32 perm[MAX_D], sizes[MAX_D], ball[MAX_D];
    orc_generate_blocks(n_free, io->n_blocks, io->seed, io->stage, perm, sizes);
    for (int a = 0; a < n_free; ++a) ball[a] = freeidx[perm[a]];
    int pst = 0;
    orc_proposal *pr = orc_proposal_create(d, n_free, mean_fr, cov_fr, io->n_blocks, sizes, perm, ball, io->c, &pst);
    free(cov_fr);
    if (!mean_out) free(mean);
    if (!cov_out) free(cov);
    if (!pr) { io->status = pst; return pst; }
    orc_mutate(m, pr, cloud, N, 0, phi_n, io->phi_n1, io->alpha, io->n_mh_steps, n_free, io->has_old, io->seed,
               io->stage, io->nthreads);
    orc_proposal_free(pr);
    io->accept = orc_mean_accept(cloud, N, d, n_free);
    io->ess_prev = io->ess;
    io->status = 0;
    return 0;
}
