// detmath.cuh -- deterministic binary64 elementary functions + Philox4x32-10 for the SMC engine.
//
// Every operation is an explicit IEEE-754 binary64 +,-,*,/,sqrt or fma, so host code, device
// code and the CPU oracle (oracle/smc_oracle.c, written independently against the same
// algorithm description in DESIGN.md "Numerical contract") produce identical bits.  The library
// is compiled with -fmad=false (device) and -ffp-contract=off (host): the ONLY fused operations
// are the fma() calls written here.
#pragma once
#include <cstdint>
#include <cstring>
#include <cmath>

#include "normal_table.h"

#if defined(__CUDACC__)
#define SMC_HD __host__ __device__ __forceinline__
#else
#define SMC_HD inline
#endif

namespace smc {

// Polynomial / reduction constants.  On the device they live in __constant__ memory (one LDCU.128
// brings two of them into uniform registers; as immediates each would cost two UMOVs per use);
// on the host they are the same literals.
#define SMC_DM_COEFS(X)                                                                               \
    X(1.4426950408889634074) X(6755399441055744.0) X(6.93147180369123816490e-01)                      \
    X(1.90821492927058770002e-10) /* 0-3: log2(e), 1.5*2^52, ln2_hi, ln2_lo */                        \
    X(1.6059043836821613e-10) X(2.08767569878681e-09) X(2.505210838544172e-08)                        \
    X(2.755731922398589e-07) X(2.7557319223985893e-06) X(2.48015873015873e-05)                        \
    X(1.984126984126984e-04) X(1.388888888888889e-03) X(8.333333333333333e-03)                        \
    X(4.1666666666666664e-02) X(1.6666666666666666e-01) /* 4-14: 1/13! .. 1/3! */                     \
    X(6.666666666666735130e-01) X(3.999999999940941908e-01) X(2.857142874366239149e-01)               \
    X(2.222219843214978396e-01) X(1.818357216161805012e-01) X(1.531383769920937332e-01)               \
    X(1.479819860511658591e-01) /* 15-21: Lg1..Lg7 */                                                 \
    X(-1.66666666666666324348e-01) X(8.33333333332248946124e-03) X(-1.98412698298579493134e-04)       \
    X(2.75573137070700676789e-06) X(-2.50507602534068634195e-08) X(1.58969099521155010221e-10)        \
    /* 22-27: S1..S6 */                                                                               \
    X(4.16666666666666019037e-02) X(-1.38888888888741095749e-03) X(2.48015872894767294178e-05)        \
    X(-2.75573143513906633035e-07) X(2.08757232129817482790e-09) X(-1.13596475577881948265e-11)       \
    /* 28-33: C1..C6 */                                                                               \
    X(7.85398163397448278999e-01) /* 34: pi/4 */
#define SMC_DM_VAL(v) v,
#if defined(__CUDACC__)
static __constant__ double dm_coef_dev[] = {SMC_DM_COEFS(SMC_DM_VAL)};
#endif
static constexpr double dm_coef_host[] = {SMC_DM_COEFS(SMC_DM_VAL)};
#if defined(__CUDA_ARCH__)
#define DMC(i) dm_coef_dev[i]
#else
#define DMC(i) dm_coef_host[i]
#endif

SMC_HD double bits_to_double(uint64_t u)
{
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double x; std::memcpy(&x, &u, 8); return x;
#endif
}
SMC_HD uint64_t double_to_bits(double x)
{
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(x);
#else
    uint64_t u; std::memcpy(&u, &x, 8); return u;
#endif
}
SMC_HD double dinf()
{
    return bits_to_double(0x7ff0000000000000ull);
}
SMC_HD double dnan()
{
    return bits_to_double(0x7ff8000000000000ull);
}

// exp(x): k = rint(x*log2(e)); r = x - k*ln2 (Cody-Waite, 2 fma); degree-13 Taylor in Horner
// form; scale by 2^k with at most one rounding.
SMC_HD double det_exp(double x)
{
    // straight-line code (selects instead of early returns: independent calls interleave on the device); every finite x
    // inside the range takes exactly the operations of the textbook form, p * 1.0 included
    const bool over = x > 709.782712893384, under = x < -745.1332191019412;
    const double xc = (over || under || x != x) ? 0.0 : x;
    const double t = fma(xc, DMC(0), DMC(1));
    const double kd = t - DMC(1);
    int k = (int)kd;
    double r = fma(-kd, DMC(2), xc);
    r = fma(-kd, DMC(3), r);
    double p = DMC(4);
    p = fma(p, r, DMC(5));
    p = fma(p, r, DMC(6));
    p = fma(p, r, DMC(7));
    p = fma(p, r, DMC(8));
    p = fma(p, r, DMC(9));
    p = fma(p, r, DMC(10));
    p = fma(p, r, DMC(11));
    p = fma(p, r, DMC(12));
    p = fma(p, r, DMC(13));
    p = fma(p, r, DMC(14));
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const bool big = k > 1023, small = k < -1021;
    p = p * (big ? 0x1p1023 : (small ? 0x1p-1000 : 1.0));
    k = k - (big ? 1023 : (small ? -1000 : 0));
    double res = p * bits_to_double((uint64_t)(k + 1023) << 52);
    res = over ? dinf() : res;
    res = under ? 0.0 : res;
    return (x != x) ? x : res;
}

// log(x): fdlibm-style reduction x = 2^k (1+f), s = f/(2+f), degree-14 odd series in s.
// Core for positive, finite, NORMAL x given as bits (k0 = exponent adjustment already applied).
SMC_HD double det_log_core(uint64_t ix, int k)
{
    int32_t hx = (int32_t)(ix >> 32);
    k += (hx >> 20) - 1023;
    hx &= 0x000fffff;
    const int32_t i = (hx + 0x95f64) & 0x100000;
    ix = ((uint64_t)(uint32_t)(hx | (i ^ 0x3ff00000)) << 32) | (ix & 0xffffffffull);
    const double x = bits_to_double(ix);
    k += (i >> 20);
    const double dk = (double)k;
    const double f = x - 1.0;
    const double s = f / (2.0 + f);
    const double z = s * s;
    const double w = z * z;
    const double t1 = w * fma(w, fma(w, DMC(20), DMC(18)), DMC(16));
    const double t2 = z * fma(w, fma(w, fma(w, DMC(21), DMC(19)), DMC(17)), DMC(15));
    const double R = t2 + t1;
    const double hfsq = 0.5 * f * f;
    const double u = fma(s, hfsq + R, dk * DMC(3));
    return fma(dk, DMC(2), -((hfsq - u) - f));
}
SMC_HD double det_log(double x)
{
    uint64_t ix = double_to_bits(x);
    const int32_t hx = (int32_t)(ix >> 32);
    const uint32_t lx = (uint32_t)ix;
    int k = 0;
    if (hx < 0x00100000) {
        if (((hx & 0x7fffffff) | lx) == 0) return -dinf();
        if (hx < 0) return dnan();
        k -= 54; x *= 0x1p54; ix = double_to_bits(x);
    }
    if ((int32_t)(ix >> 32) >= 0x7ff00000) return x + x;
    return det_log_core(ix, k);
}
// same value as det_log(x) for positive finite normal x (no special-case branches)
SMC_HD double det_log_normal(double x) { return det_log_core(double_to_bits(x), 0); }

// sin(2*pi*u), cos(2*pi*u) for u in [0,1): exact octant split, fdlibm kernels on [0, pi/4].
SMC_HD void det_sincos2pi(double u, double& sn, double& cs)
{
    const double t = u * 8.0;
    const int o = (int)t;
    const double f = t - (double)o;
    const double g = (o & 1) ? (1.0 - f) : f;
    const double x = g * DMC(34);
    const double z = x * x;
    const double ps = fma(z, fma(z, fma(z, fma(z, fma(z, DMC(27), DMC(26)), DMC(25)), DMC(24)), DMC(23)), DMC(22));
    const double s = fma(x * z, ps, x);
    const double pc = fma(z, fma(z, fma(z, fma(z, fma(z, DMC(33), DMC(32)), DMC(31)), DMC(30)), DMC(29)), DMC(28));
    const double c = fma(z * z, pc, fma(-0.5, z, 1.0));
    const double sp = (o & 1) ? c : s;
    const double cp = (o & 1) ? s : c;
    const int q = o >> 1;
    sn = (q == 0) ? sp : (q == 1) ? cp : (q == 2) ? -sp : -cp;
    cs = (q == 0) ? cp : (q == 1) ? -sp : (q == 2) ? -cp : sp;
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10, counter = (particle, stage, slot, purpose), key = seed
// ---------------------------------------------------------------------------------------------
struct u32x4 { uint32_t x, y, z, w; };

SMC_HD void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo)
{
    const uint64_t p = (uint64_t)a * b;      // one IMAD.WIDE.U32 on the device
    lo = (uint32_t)p; hi = (uint32_t)(p >> 32);
}

SMC_HD u32x4 philox4x32_10(u32x4 c, uint32_t k0, uint32_t k1)
{
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int r = 0; r < 10; ++r) {
        uint32_t h0, l0, h1, l1;
        mulhilo(0xD2511F53u, c.x, h0, l0);
        mulhilo(0xCD9E8D57u, c.z, h1, l1);
        u32x4 n;
        n.x = h1 ^ c.y ^ k0; n.y = l1; n.z = h0 ^ c.w ^ k1; n.w = l0;
        c = n;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return c;
}

// the same ten rounds with the key schedule precomputed (rk[2r] = k0 + r W0, rk[2r + 1] = k1 + r W1): bit-identical
SMC_HD void philox_round_keys(uint64_t seed, uint32_t* rk)
{
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) { rk[2 * r] = k0; rk[2 * r + 1] = k1; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
}
template <class KEYS>
SMC_HD u32x4 philox4x32_10_keyed(u32x4 c, const KEYS& rk)
{
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int r = 0; r < 10; ++r) {
        uint32_t h0, l0, h1, l1;
        mulhilo(0xD2511F53u, c.x, h0, l0);
        mulhilo(0xCD9E8D57u, c.z, h1, l1);
        u32x4 n;
        n.x = h1 ^ c.y ^ rk[2 * r]; n.y = l1; n.z = h0 ^ c.w ^ rk[2 * r + 1]; n.w = l0;
        c = n;
    }
    return c;
}
template <class KEYS>
SMC_HD u32x4 rng4_keyed(const KEYS& rk, uint32_t particle, uint32_t stage, uint32_t slot, uint32_t purpose)
{
    u32x4 c; c.x = particle; c.y = stage; c.z = slot; c.w = purpose;
    return philox4x32_10_keyed(c, rk);
}

enum : uint32_t { PURP_STEP = 1, PURP_NORMAL = 2, PURP_RESAMPLE = 3, PURP_BLOCKS = 4, PURP_INIT = 5 };

SMC_HD u32x4 rng4(uint64_t seed, uint32_t particle, uint32_t stage, uint32_t slot, uint32_t purpose)
{
    u32x4 c; c.x = particle; c.y = stage; c.z = slot; c.w = purpose;
    return philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
}
// 53-bit uniforms: [0,1) and (0,1]
SMC_HD double u01(uint32_t hi, uint32_t lo) { return (double)((((uint64_t)hi << 32) | lo) >> 11) * 0x1p-53; }
SMC_HD double u01_open0(uint32_t hi, uint32_t lo) { return (double)(((((uint64_t)hi << 32) | lo) >> 11) + 1) * 0x1p-53; }

// Box-Muller pair from one Philox block
SMC_HD void normal_pair(u32x4 r, double& z0, double& z1)
{
    const double u1 = u01_open0(r.x, r.y);
    const double u2 = u01(r.z, r.w);
    const double rad = sqrt(-2.0 * det_log_normal(u1));   // u1 in [2^-53, 1]: always normal
    double sn, cs;
    det_sincos2pi(u2, sn, cs);
    z0 = rad * cs;
    z1 = rad * sn;
}

// ---------------------------------------------------------------------------------------------
// Proposal normals: FOUR N(0,1) variates from ONE Philox block, one 32-bit word each, by a table-driven inverse
// normal CDF evaluated in binary32 (tools/make_normal_table.py writes the table and reports its accuracy:
// max |z - Phi^-1| = 5e-7, |z| <= 6.34).  The random-walk proposal only needs a symmetric, well-distributed
// increment (the Metropolis ratio is computed in binary64 from the point actually proposed), so the normals carry
// float precision.  Word r: sign = bit 31; v = (r << 1) | 1 is odd, p = v / 2^33 in (0, 1/2); the binary32 image
// f = (float)v selects the segment (exponent, top three mantissa bits) and the remaining 20 mantissa bits give
// x in [1, 1.125); z = ((b3 x + b2) x + b1) x + b0 with three fma; the sign bit is XORed in, so the map is exactly
// antisymmetric (z -> -z is measure preserving).  ~11 instructions per normal on sm_100a (the binary32 Box-Muller
// it replaces took ~55) and every step is an exact integer operation or an IEEE binary32 conversion / fma, so the
// oracle reproduces the same bits.
// ---------------------------------------------------------------------------------------------
SMC_HD float bits_to_float(uint32_t u)
{
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float x; std::memcpy(&x, &u, 4); return x;
#endif
}
SMC_HD uint32_t float_to_bits(float x)
{
#if defined(__CUDA_ARCH__)
    return __float_as_uint(x);
#else
    uint32_t u; std::memcpy(&u, &x, 4); return u;
#endif
}
constexpr int NORMAL_TAB_ROWS = SMC_NORMAL_TABLE_ROWS;
static constexpr float normal_tab_host[4 * NORMAL_TAB_ROWS] = {SMC_NORMAL_TABLE_VALUES};
#if defined(__CUDACC__)
static __device__ __align__(16) const float normal_tab_dev[4 * NORMAL_TAB_ROWS] = {SMC_NORMAL_TABLE_VALUES};
#endif
// tab: [NORMAL_TAB_ROWS] rows of (b0, b1, b2, b3); 16-byte aligned (shared memory inside the mutation kernel)
SMC_HD float normal_icdf_f(uint32_t r, const float4* tab)
{
    const uint32_t v = (r << 1) | 1u;
#if defined(__CUDA_ARCH__)
    const float f = __uint2float_rn(v);
#else
    const float f = (float)v;
#endif
    const uint32_t fb = float_to_bits(f);
    constexpr int SH = 23 - SMC_NORMAL_TABLE_LOG2SUB;
    const float4 c = tab[(fb >> SH) - (127u << SMC_NORMAL_TABLE_LOG2SUB)];
    const float x = bits_to_float((fb & ((1u << SH) - 1u)) | 0x3f800000u);
    const float z = fmaf(fmaf(fmaf(c.w, x, c.z), x, c.y), x, c.x);
    return bits_to_float(float_to_bits(z) ^ (r & 0x80000000u));
}
SMC_HD double normal_icdf(uint32_t r, const float4* tab) { return (double)normal_icdf_f(r, tab); }
SMC_HD void normal_quad(u32x4 r, const float4* tab, double& z0, double& z1, double& z2, double& z3)
{
    z0 = normal_icdf(r.x, tab); z1 = normal_icdf(r.y, tab); z2 = normal_icdf(r.z, tab); z3 = normal_icdf(r.w, tab);
}

// Lower Cholesky, row by row, explicit fma order.  Returns 0 or 1 + failing row.
SMC_HD int cholesky_lower(const double* A, int n, double* L)
{
    for (int i = 0; i < n * n; ++i) L[i] = 0.0;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j <= i; ++j) {
            double s = A[i * n + j];
            for (int k = 0; k < j; ++k) s = fma(-L[i * n + k], L[j * n + k], s);
            if (i == j) {
                if (!(s > 0.0)) return 1 + i;
                L[i * n + i] = sqrt(s);
            } else {
                L[i * n + j] = s / L[j * n + j];
            }
        }
    return 0;
}

// c <- c * (0.95 + 0.10 e^{16(a-t)} / (1 + e^{16(a-t)}))      (src/smc_main.jl:453-455)
SMC_HD double update_step_size(double c, double accept, double target)
{
    const double e = det_exp(16.0 * (accept - target));
    return c * (0.95 + 0.10 * e / (1.0 + e));
}

}  // namespace smc
