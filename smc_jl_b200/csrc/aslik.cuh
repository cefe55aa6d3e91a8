// aslik.cuh -- K8: device log-likelihood of the three-equation An-Schorfheide DSGE model (BASELINE
// config C4; in the reference it is the user likelihood `DSGE.likelihood(m, data; ...)` passed to smc(),
// examples/dsge_models/small_dsge_model.jl:35-50 -- DSGE.jl itself is not part of the reference tree).
//
// One thread per particle, everything in registers:
//   1. decision rule in closed form: the model has ONE predetermined endogenous variable, so instead of a
//      QZ decomposition (gensys) the persistence a_R is the root inside the unit circle of a cubic
//      (bisection, 64 fixed steps) and the shock loadings follow from a 2x2 solve; "exactly one stable
//      root" <=> gensys' existence + uniqueness, anything else returns -Inf like catch_errors=true;
//   2. 6-state form (y, pi, R, y_lag, g, z): the reference's 8 states minus the two expectation states
//      (unobserved, no feedback); stationary covariance in closed form (no Lyapunov iteration);
//   3. Kalman filter with the transition matrix's sparsity (10 non-zeros of 36) and the selection-type
//      measurement matrix unrolled at compile time; 3x3 innovation covariance as L D L' (three reciprocals, no
//      square roots); covariance update P - H D^-1 H'.  About 400 FP64 instructions per period instead of ~4 500 for dense 8-state
//      algebra; a dense DMMA formulation would spend ~10x the flops on structural zeros, and B200's FP64
//      tensor rate is no higher than its FP64 FMA rate, so tensor cores are deliberately not used here.
// The operation order is fixed (explicit fma) and mirrored by the oracle (oracle/as_model.c).
#pragma once
#include "common.cuh"

namespace smc {
namespace {

__constant__ ASConst c_as[2];

namespace as {
enum { SY = 0, SPI = 1, SR = 2, SYL = 3, SG = 4, SZ = 5, NS = 6 };

// structural non-zeros of the transition matrix
__device__ __forceinline__ constexpr bool tnz(int i, int k)
{
    return (i == SY && (k == SR || k == SG || k == SZ)) || (i == SPI && (k == SR || k == SZ)) ||
           (i == SR && (k == SR || k == SZ)) || (i == SYL && k == SY) || (i == SG && k == SG) || (i == SZ && k == SZ);
}
__device__ __forceinline__ constexpr int lo(int i, int j) { return i >= j ? i * (i + 1) / 2 + j : j * (j + 1) / 2 + i; }

__device__ __forceinline__ double cubic(double c3, double c2, double c1, double c0, double x)
{
    return fma(fma(fma(c3, x, c2), x, c1), x, c0);
}

__device__ __forceinline__ bool stable_root(double c3, double c2, double c1, double c0, double& root)
{
    const double p0 = cubic(c3, c2, c1, c0, 0.0), p1 = cubic(c3, c2, c1, c0, 1.0), pm = cubic(c3, c2, c1, c0, -1.0);
    double lo_, hi_;
    if (p0 == 0.0) { lo_ = 0.0; hi_ = 0.0; }
    else if ((p0 < 0.0) != (p1 < 0.0)) { lo_ = 0.0; hi_ = 1.0; }
    else if ((p0 < 0.0) != (pm < 0.0)) { lo_ = -1.0; hi_ = 0.0; }
    else return false;
    const bool neg_lo = cubic(c3, c2, c1, c0, lo_) < 0.0;
#pragma unroll 1
    for (int it = 0; it < 64; ++it) {
        const double mid = 0.5 * (lo_ + hi_);
        const bool neg_mid = cubic(c3, c2, c1, c0, mid) < 0.0;
        if (neg_mid == neg_lo) lo_ = mid; else hi_ = mid;
    }
    const double r = 0.5 * (lo_ + hi_);
    const double q1 = fma(c3, r, c2);
    const double q0 = fma(q1, r, c1);
    const double disc = fma(q1, q1, -4.0 * c3 * q0);
    if (disc < 0.0) {
        if (q0 / c3 < 1.0) return false;
    } else {
        const double sq = sqrt(disc);
        const double t = -0.5 * (q1 + (q1 >= 0.0 ? sq : -sq));
        const double x1 = t / c3;
        const double x2 = (t != 0.0) ? q0 / t : 0.0;
        if (fabs(x1) < 1.0 || fabs(x2) < 1.0) return false;
    }
    if (!(fabs(r) < 1.0)) return false;
    root = r;
    return true;
}
}  // namespace as

// theta = (tau, kappa, psi1, psi2, rA, pi*, gammaQ, rho_R, rho_g, rho_z, sigma_R, sigma_g, sigma_z, e_y, e_pi, e_R)
// NANS: the data hold missing observations (instantiated separately so that complete data run exactly the code they always did)
template <bool NANS>
__device__ __noinline__ double as_loglik_t(int slot, const double* __restrict__ th)
{
    using namespace as;
    const double tau = th[0], kap = th[1], psi1 = th[2], psi2 = th[3], rA = th[4], pistar = th[5], gamQ = th[6];
    const double rhoR = th[7], rhog = th[8], rhoz = th[9], sigR = th[10], sigg = th[11], sigz = th[12];
    const double ey = th[13], epi = th[14], eR = th[15];
    // ---- 1. decision rule ------------------------------------------------------------------------
    const double beta = 1.0 / (1.0 + rA / 400.0);
    const double b = (1.0 + beta) + kap / tau;
    const double h = 1.0 - rhoR;
    const double k = h / tau;
    const double c3 = beta;
    const double c2 = -((b + rhoR * beta) + (k * psi2) * beta);
    const double c1 = (1.0 + rhoR * b) + k * (psi1 * kap + psi2);
    const double c0 = -rhoR;
    double lam;
    if (!(tau > 0.0) || !stable_root(c3, c2, c1, c0, lam)) return -dinf();
    const double Dl = fma(fma(beta, lam, -b), lam, 1.0);
    const double a_y = -((lam / tau) * (1.0 - beta * lam)) / Dl;
    const double a_p = -((lam * kap) / tau) / Dl;
    const double a_R = lam;
    const double m = (a_y - 1.0 / tau) + a_p / tau;
    const double cpr = kap * m + beta * a_p;
    const double B_Rr = 1.0 / (1.0 - h * (psi1 * cpr + psi2 * m));
    const double B_yr = m * B_Rr;
    const double B_pr = cpr * B_Rr;
    const double A11 = (1.0 - rhoz) - (m * h) * psi2;
    const double A12 = -(rhoz / tau) - (m * h) * psi1;
    const double A21 = -kap - ((beta * a_p) * h) * psi2;
    const double A22 = (1.0 - beta * rhoz) - ((beta * a_p) * h) * psi1;
    const double r1 = rhoz / tau;
    const double det = A11 * A22 - A12 * A21;
    const double B_yz = (r1 * A22) / det;
    const double B_pz = -(A21 * r1) / det;
    const double B_Rz = h * (psi2 * B_yz + psi1 * B_pz);
    // ---- 2. state space ----------------------------------------------------------------------------
    double Tm[NS][NS];
#pragma unroll
    for (int i = 0; i < NS; ++i)
#pragma unroll
        for (int j = 0; j < NS; ++j) Tm[i][j] = 0.0;
    Tm[SY][SR] = a_y;  Tm[SY][SG] = rhog; Tm[SY][SZ] = B_yz * rhoz;
    Tm[SPI][SR] = a_p; Tm[SPI][SZ] = B_pz * rhoz;
    Tm[SR][SR] = a_R;  Tm[SR][SZ] = B_Rz * rhoz;
    Tm[SYL][SY] = 1.0; Tm[SG][SG] = rhog; Tm[SZ][SZ] = rhoz;
    const double qz = sigz * sigz, qg = sigg * sigg, qr = sigR * sigR;
    const double Iz[NS] = {B_yz, B_pz, B_Rz, 0.0, 0.0, 1.0};
    const double Ig[NS] = {1.0, 0.0, 0.0, 0.0, 1.0, 0.0};
    const double Ir[NS] = {B_yr, B_pr, B_Rr, 0.0, 0.0, 0.0};
    double RQR[21];
#pragma unroll
    for (int i = 0; i < NS; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j)
            RQR[lo(i, j)] = fma(Ir[i] * qr, Ir[j], fma(Ig[i] * qg, Ig[j], (Iz[i] * qz) * Iz[j]));
    // stationary covariance
    const double Sgg = qg / (1.0 - rhog * rhog);
    const double Szz = qz / (1.0 - rhoz * rhoz);
    const double SRz = (B_Rz * Szz) / (1.0 - a_R * rhoz);
    const double SRR = (fma(((2.0 * a_R) * B_Rz) * rhoz, SRz, fma(B_Rz * B_Rz, Szz, (B_Rr * B_Rr) * qr))) / (1.0 - a_R * a_R);
    const double VRz = rhoz * SRz;
    const double M[5][4] = {{a_y, 1.0, B_yz, B_yr}, {a_p, 0.0, B_pz, B_pr}, {a_R, 0.0, B_Rz, B_Rr}, {0.0, 1.0, 0.0, 0.0}, {0.0, 0.0, 1.0, 0.0}};
    constexpr int map5[5] = {SY, SPI, SR, SG, SZ};
    double P[21];
#pragma unroll
    for (int a = 0; a < 5; ++a)
#pragma unroll
        for (int c = 0; c <= a; ++c) {
            double v = (M[a][0] * SRR) * M[c][0];
            v = fma(M[a][1] * Sgg, M[c][1], v);
            v = fma(M[a][2] * Szz, M[c][2], v);
            v = fma(M[a][3] * qr, M[c][3], v);
            v = fma(VRz, fma(M[a][0], M[c][2], M[a][2] * M[c][0]), v);
            P[lo(map5[a], map5[c])] = v;
        }
    P[lo(SYL, SYL)] = P[lo(SY, SY)];
#pragma unroll
    for (int a = 0; a < 5; ++a) {
        const int i = map5[a];
        double v = 0.0;
#pragma unroll
        for (int kk = 0; kk < NS; ++kk)
            if (tnz(i, kk) && kk != SYL) v = fma(Tm[i][kk], P[lo(SY, kk)], v);
        P[lo(SYL, i)] = v;
    }
    // ---- 3. Kalman filter ----------------------------------------------------------------------------
    const double D0 = gamQ, D1 = pistar, D2 = (pistar + rA) + 4.0 * gamQ;
    const double E0 = ey * ey, E1 = epi * epi, E2 = eR * eR;
    const double* __restrict__ data = c_as[slot].data;
    const int T = c_as[slot].T, npre = c_as[slot].npre;
    double x[NS] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    double ll = 0.0;
    bool bad = false;
#pragma unroll 1
    for (int t = 0; t < T; ++t) {
        const double y0 = __ldg(data + 3 * t), y1 = __ldg(data + 3 * t + 1), y2 = __ldg(data + 3 * t + 2);
        double xn[NS], TP[NS][NS], Pn[21];
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            double v = 0.0;
#pragma unroll
            for (int kk = 0; kk < NS; ++kk)
                if (tnz(i, kk)) v = fma(Tm[i][kk], x[kk], v);
            xn[i] = v;
#pragma unroll
            for (int j = 0; j < NS; ++j) {
                double w = 0.0;
#pragma unroll
                for (int kk = 0; kk < NS; ++kk)
                    if (tnz(i, kk)) w = fma(Tm[i][kk], P[lo(kk, j)], w);
                TP[i][j] = w;
            }
        }
#pragma unroll
        for (int i = 0; i < NS; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) {
                double w = RQR[lo(i, j)];
#pragma unroll
                for (int kk = 0; kk < NS; ++kk)
                    if (tnz(j, kk)) w = fma(TP[i][kk], Tm[j][kk], w);
                Pn[lo(i, j)] = w;
            }
        double PZ[NS][3];
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            PZ[i][0] = (Pn[lo(i, SY)] - Pn[lo(i, SYL)]) + Pn[lo(i, SZ)];
            PZ[i][1] = 4.0 * Pn[lo(i, SPI)];
            PZ[i][2] = 4.0 * Pn[lo(i, SR)];
        }
        double F00 = ((PZ[SY][0] - PZ[SYL][0]) + PZ[SZ][0]) + E0;
        double F10 = 4.0 * PZ[SPI][0];
        double F11 = 4.0 * PZ[SPI][1] + E1;
        double F20 = 4.0 * PZ[SR][0];
        double F21 = 4.0 * PZ[SR][1];
        double F22 = 4.0 * PZ[SR][2] + E2;
        double n0 = (y0 - ((xn[SY] - xn[SYL]) + xn[SZ])) - D0;
        double n1 = (y1 - 4.0 * xn[SPI]) - D1;
        double n2 = (y2 - 4.0 * xn[SR]) - D2;
        // Missing observations (NaN; DSGE.jl's filter drops those rows for the period): a missing slot k becomes an inert
        // pivot -- F_kk = 1, F_kj = 0, innovation 0, gain column 0 -- so the factorisation below passes through it with exact
        // zeros and ones and the remaining slots see precisely the operations of the reduced system (what the oracle runs).
        // Warp-uniform: the data are the same for every particle.
        double nobs_c = 3.0 * 1.8378770664093453;
        if (NANS) {
            const bool m0 = y0 != y0, m1 = y1 != y1, m2 = y2 != y2;
            if (m0 | m1 | m2) {
                if (m0) { F00 = 1.0; F10 = 0.0; F20 = 0.0; n0 = 0.0; }
                if (m1) { F11 = 1.0; F10 = 0.0; F21 = 0.0; n1 = 0.0; }
                if (m2) { F22 = 1.0; F20 = 0.0; F21 = 0.0; n2 = 0.0; }
#pragma unroll
                for (int i = 0; i < NS; ++i) {
                    if (m0) PZ[i][0] = 0.0;
                    if (m1) PZ[i][1] = 0.0;
                    if (m2) PZ[i][2] = 0.0;
                }
                nobs_c = (double)(3 - (int)m0 - (int)m1 - (int)m2) * 1.8378770664093453;
            }
        }
        // F = L D L' (unit lower L): three reciprocals, no square roots; d_k > 0 <=> F positive definite
        const double d0 = F00, r0 = 1.0 / d0;
        const double l10 = F10 * r0, l20 = F20 * r0;
        const double d1 = fma(-l10, F10, F11), r1 = 1.0 / d1;
        const double t21 = fma(-l20, F10, F21);
        const double l21 = t21 * r1;
        const double d2 = fma(-l21, t21, fma(-l20, F20, F22)), r2 = 1.0 / d2;
        if (!(d0 > 0.0 && d1 > 0.0 && d2 > 0.0)) bad = true;
        const double w0 = n0;
        const double w1 = fma(-l10, w0, n1);
        const double w2 = fma(-l21, w1, fma(-l20, w0, n2));
        const double v0 = w0 * r0, v1 = w1 * r1, v2 = w2 * r2;
        if (t >= npre) {
            const double logdet = det_log((d0 * d1) * d2);
            const double quad = fma(w2, v2, fma(w1, v1, w0 * v0));
            ll = ll + -0.5 * ((nobs_c + logdet) + quad);
        }
        // update: H = PZ L^{-T}; x += H D^{-1} w; P -= H D^{-1} H'
        double H[NS][3], HR[NS][3];
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            H[i][0] = PZ[i][0];
            H[i][1] = fma(-l10, H[i][0], PZ[i][1]);
            H[i][2] = fma(-l21, H[i][1], fma(-l20, H[i][0], PZ[i][2]));
            HR[i][0] = H[i][0] * r0; HR[i][1] = H[i][1] * r1; HR[i][2] = H[i][2] * r2;
            x[i] = fma(H[i][2], v2, fma(H[i][1], v1, fma(H[i][0], v0, xn[i])));
        }
#pragma unroll
        for (int i = 0; i < NS; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j)
                P[lo(i, j)] = fma(-HR[i][2], H[j][2], fma(-HR[i][1], H[j][1], fma(-HR[i][0], H[j][0], Pn[lo(i, j)])));
    }
    if (bad) return -dinf();
    if (!(ll == ll)) return -dinf();
    return ll;
}

__device__ __forceinline__ double as_loglik(int slot, const double* __restrict__ th)
{
    return c_as[slot].has_nan ? as_loglik_t<true>(slot, th) : as_loglik_t<false>(slot, th);
}

struct ASLik {
    static constexpr int KIND = SMCB200_LIK_AS_DSGE;
    static constexpr int NEQ = 0, K = 0, STRIDE = 0, COEF = 0, SIG = -1;
    static constexpr int D = 16;
    static constexpr int MINB = 4;     // resident blocks of 128 threads: 128 registers + ~130 B of L1-resident spills beat 2 blocks x 214 registers (measured 6.4 vs 7.6 ms)
    template <int SLOT>
    static __device__ __forceinline__ double ll(const double (&th)[D]) { return as_loglik(SLOT, th); }
};

}  // namespace
}  // namespace smc
