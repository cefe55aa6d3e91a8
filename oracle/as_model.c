/*
 * as_model.c -- CPU oracle of the An-Schorfheide (2007) three-equation DSGE log-likelihood, the user
 * likelihood of BASELINE config C4 (examples/dsge_models/small_dsge_model.jl:10-50 ->
 * DSGE.likelihood(m, data; sampler=false, catch_errors=true, use_chand_recursion=true)).
 *
 * TEST INFRASTRUCTURE ONLY (see smc_oracle.c header).  The model itself lives in DSGE.jl /
 * StateSpaceRoutines.jl, which are NOT in /root/reference (unvendored, unpinned dependency); its results
 * are pinned by the (theta -> loglh / old_loglh) columns of the clouds the reference stores under
 * test/reference/solve_adaptive_phi.jld2 and test/save/output_data/an_schorfheide/ss0/estimate/raw/ (the .jld2 clouds)
 * (2 000 rows, extracted to tests/golden/as_clouds.npz; tests/test_oracle_golden.py checks this file
 * against them).  Equations: SURVEY.md Appendix A.
 *
 * Algorithm (published model, restated):
 *   1. The canonical system has one predetermined endogenous variable (R_{t-1}) and two jump variables, so
 *      the decision rule is (y, pi, R)_t = a R_{t-1} + B (g_t, z_t, eps_R,t)'.  a_R is the root inside the
 *      unit circle of  (x - rho_R) D(x) + ((1 - rho_R)/tau) x (psi1 kappa + psi2 (1 - beta x)) = 0,
 *      D(x) = beta x^2 - (1 + beta + kappa/tau) x + 1.  Exactly one such root <=> a unique stable solution
 *      (gensys eu = [1, 1]); otherwise the likelihood is -Inf (DSGE.jl catch_errors).
 *   2. State (y, pi, R, y_lag, g, z) -- the reference's 8-state vector without the two expectation states,
 *      which neither feed back nor are observed.  Stationary covariance in closed form.
 *   3. Kalman filter over all periods; the first `npre` (presample) periods are filtered but not scored.
 *      The 3x3 innovation covariance is factored as L D L' (no square roots); the covariance update is
 *      P - H D^{-1} H', H = P Z' L^{-T}.
 * Every product/sum below is written in a fixed order with explicit fma (the device kernel
 * smc_jl_b200/csrc/aslik.cuh was written independently against this same description).
 */
#include <math.h>
#include <stdint.h>

#define FMA(a, b, c) __builtin_fma((a), (b), (c))
#define ORC_API __attribute__((visibility("default")))

double orc_log(double x);   /* smc_oracle.c: deterministic log */

enum { SY = 0, SPI = 1, SR = 2, SYL = 3, SG = 4, SZ = 5, NS = 6 };

static double cubic(double c3, double c2, double c1, double c0, double x)
{
    return FMA(FMA(FMA(c3, x, c2), x, c1), x, c0);
}

/* root inside the unit circle, if it is the only one; returns 0 on failure */
static int stable_root(double c3, double c2, double c1, double c0, double *root)
{
    const double p0 = cubic(c3, c2, c1, c0, 0.0), p1 = cubic(c3, c2, c1, c0, 1.0), pm = cubic(c3, c2, c1, c0, -1.0);
    double lo, hi;
    if (p0 == 0.0) { lo = 0.0; hi = 0.0; }
    else if ((p0 < 0.0) != (p1 < 0.0)) { lo = 0.0; hi = 1.0; }
    else if ((p0 < 0.0) != (pm < 0.0)) { lo = -1.0; hi = 0.0; }
    else return 0;
    int neg_lo = cubic(c3, c2, c1, c0, lo) < 0.0;
    for (int it = 0; it < 64; ++it) {
        const double mid = 0.5 * (lo + hi);
        const int neg_mid = cubic(c3, c2, c1, c0, mid) < 0.0;
        if (neg_mid == neg_lo) lo = mid; else hi = mid;
    }
    const double r = 0.5 * (lo + hi);
    /* deflate to c3 x^2 + q1 x + q0 and require both remaining roots outside (or on) the unit circle */
    const double q1 = FMA(c3, r, c2);
    const double q0 = FMA(q1, r, c1);
    const double disc = FMA(q1, q1, -4.0 * c3 * q0);
    if (disc < 0.0) {
        if (q0 / c3 < 1.0) return 0;
    } else {
        const double sq = sqrt(disc);
        const double t = -0.5 * (q1 + (q1 >= 0.0 ? sq : -sq));
        const double x1 = t / c3;
        const double x2 = (t != 0.0) ? q0 / t : 0.0;
        if (fabs(x1) < 1.0 || fabs(x2) < 1.0) return 0;
    }
    if (!(fabs(r) < 1.0)) return 0;
    *root = r;
    return 1;
}

/* structural non-zero pattern of the transition matrix (rows/cols in state order) */
static const int TNZ[NS][NS] = {
    /* y   */ {0, 0, 1, 0, 1, 1},
    /* pi  */ {0, 0, 1, 0, 0, 1},
    /* R   */ {0, 0, 1, 0, 0, 1},
    /* yl  */ {1, 0, 0, 0, 0, 0},
    /* g   */ {0, 0, 0, 0, 1, 0},
    /* z   */ {0, 0, 0, 0, 0, 1},
};

/* theta = (tau, kappa, psi1, psi2, rA, pi*, gammaQ, rho_R, rho_g, rho_z, sigma_R, sigma_g, sigma_z, e_y, e_pi, e_R);
 * data: 3 x T column-major (rows gdp growth, inflation, nominal rate). */
ORC_API double orc_as_loglik(const double *th, const double *data, int T, int npre)
{
    const double tau = th[0], kap = th[1], psi1 = th[2], psi2 = th[3], rA = th[4], pistar = th[5], gamQ = th[6];
    const double rhoR = th[7], rhog = th[8], rhoz = th[9], sigR = th[10], sigg = th[11], sigz = th[12];
    const double ey = th[13], epi = th[14], eR = th[15];
    /* ---- 1. decision rule --------------------------------------------------------------------- */
    const double beta = 1.0 / (1.0 + rA / 400.0);
    const double b = (1.0 + beta) + kap / tau;
    const double h = 1.0 - rhoR;
    const double k = h / tau;
    const double c3 = beta;
    const double c2 = -((b + rhoR * beta) + (k * psi2) * beta);
    const double c1 = (1.0 + rhoR * b) + k * (psi1 * kap + psi2);
    const double c0 = -rhoR;
    double lam;
    if (!(tau > 0.0) || !stable_root(c3, c2, c1, c0, &lam)) return -INFINITY;
    const double Dl = FMA(FMA(beta, lam, -b), lam, 1.0);
    const double a_y = -((lam / tau) * (1.0 - beta * lam)) / Dl;
    const double a_p = -((lam * kap) / tau) / Dl;
    const double a_R = lam;
    const double m = (a_y - 1.0 / tau) + a_p / tau;
    /* monetary policy shock */
    const double cpr = kap * m + beta * a_p;
    const double B_Rr = 1.0 / (1.0 - h * (psi1 * cpr + psi2 * m));
    const double B_yr = m * B_Rr;
    const double B_pr = cpr * B_Rr;
    /* technology: Rz = h (psi2 Y + psi1 P) substituted into the Euler and Phillips equations -> 2x2 */
    const double A11 = (1.0 - rhoz) - (m * h) * psi2;
    const double A12 = -(rhoz / tau) - (m * h) * psi1;
    const double A21 = -kap - ((beta * a_p) * h) * psi2;
    const double A22 = (1.0 - beta * rhoz) - ((beta * a_p) * h) * psi1;
    const double r1 = rhoz / tau;
    const double det = A11 * A22 - A12 * A21;
    const double B_yz = (r1 * A22) / det;
    const double B_pz = -(A21 * r1) / det;
    const double B_Rz = h * (psi2 * B_yz + psi1 * B_pz);
    /* ---- 2. state space ------------------------------------------------------------------------- */
    double Tm[NS][NS] = {{0}};
    Tm[SY][SR] = a_y;  Tm[SY][SG] = rhog; Tm[SY][SZ] = B_yz * rhoz;
    Tm[SPI][SR] = a_p; Tm[SPI][SZ] = B_pz * rhoz;
    Tm[SR][SR] = a_R;  Tm[SR][SZ] = B_Rz * rhoz;
    Tm[SYL][SY] = 1.0; Tm[SG][SG] = rhog; Tm[SZ][SZ] = rhoz;
    const double qz = sigz * sigz, qg = sigg * sigg, qr = sigR * sigR;
    /* impact columns: z_sh -> (B_yz, B_pz, B_Rz, 0, 0, 1); g_sh -> (1, 0, 0, 0, 1, 0); rm_sh -> (B_yr, B_pr, B_Rr, 0, 0, 0) */
    const double Iz[NS] = {B_yz, B_pz, B_Rz, 0.0, 0.0, 1.0};
    const double Ig[NS] = {1.0, 0.0, 0.0, 0.0, 1.0, 0.0};
    const double Ir[NS] = {B_yr, B_pr, B_Rr, 0.0, 0.0, 0.0};
    double RQR[NS][NS];
    for (int i = 0; i < NS; ++i)
        for (int j = 0; j <= i; ++j) {
            const double v = FMA(Ir[i] * qr, Ir[j], FMA(Ig[i] * qg, Ig[j], (Iz[i] * qz) * Iz[j]));
            RQR[i][j] = v; RQR[j][i] = v;
        }
    /* stationary covariance: g independent of (R, z); (R, z) triangular */
    const double Sgg = qg / (1.0 - rhog * rhog);
    const double Szz = qz / (1.0 - rhoz * rhoz);
    const double SRz = (B_Rz * Szz) / (1.0 - a_R * rhoz);
    const double SRR = (FMA(((2.0 * a_R) * B_Rz) * rhoz, SRz, FMA(B_Rz * B_Rz, Szz, (B_Rr * B_Rr) * qr))) / (1.0 - a_R * a_R);
    /* v = (R_{t-1}, g_t, z_t, eps_R): Var = diag(SRR, Sgg, Szz, qr), Cov(R_{t-1}, z_t) = rho_z SRz */
    const double VRz = rhoz * SRz;
    /* (y, pi, R, g, z) = M v */
    const double M[5][4] = {{a_y, 1.0, B_yz, B_yr}, {a_p, 0.0, B_pz, B_pr}, {a_R, 0.0, B_Rz, B_Rr}, {0.0, 1.0, 0.0, 0.0}, {0.0, 0.0, 1.0, 0.0}};
    static const int map5[5] = {SY, SPI, SR, SG, SZ};
    double P[NS][NS] = {{0}};
    for (int a = 0; a < 5; ++a)
        for (int c = 0; c <= a; ++c) {
            /* sum_k M[a][k] V[k][k] M[c][k] + VRz (M[a][0] M[c][2] + M[a][2] M[c][0]) */
            double v = (M[a][0] * SRR) * M[c][0];
            v = FMA(M[a][1] * Sgg, M[c][1], v);
            v = FMA(M[a][2] * Szz, M[c][2], v);
            v = FMA(M[a][3] * qr, M[c][3], v);
            v = FMA(VRz, FMA(M[a][0], M[c][2], M[a][2] * M[c][0]), v);
            P[map5[a]][map5[c]] = v; P[map5[c]][map5[a]] = v;
        }
    /* y_lag,t = y_{t-1}: Cov(y_{t-1}, x_t) = sum_k T[x][k] P[y][k] */
    P[SYL][SYL] = P[SY][SY];
    for (int a = 0; a < 5; ++a) {
        const int i = map5[a];
        double v = 0.0;
        for (int kk = 0; kk < NS; ++kk) if (TNZ[i][kk] && kk != SYL) v = FMA(Tm[i][kk], P[SY][kk], v);
        P[SYL][i] = v; P[i][SYL] = v;
    }
    /* ---- 3. Kalman filter ----------------------------------------------------------------------- */
    const double D0 = gamQ, D1 = pistar, D2 = (pistar + rA) + 4.0 * gamQ;
    const double E0 = ey * ey, E1 = epi * epi, E2 = eR * eR;
    double x[NS] = {0, 0, 0, 0, 0, 0};
    double ll = 0.0;
    int bad = 0;
    for (int t = 0; t < T; ++t) {
        /* predict */
        double xn[NS], TP[NS][NS], Pn[NS][NS];
        for (int i = 0; i < NS; ++i) {
            double v = 0.0;
            for (int kk = 0; kk < NS; ++kk) if (TNZ[i][kk]) v = FMA(Tm[i][kk], x[kk], v);
            xn[i] = v;
            for (int j = 0; j < NS; ++j) {
                double w = 0.0;
                for (int kk = 0; kk < NS; ++kk) if (TNZ[i][kk]) w = FMA(Tm[i][kk], P[kk][j], w);
                TP[i][j] = w;
            }
        }
        for (int i = 0; i < NS; ++i)
            for (int j = 0; j <= i; ++j) {
                double w = RQR[i][j];
                for (int kk = 0; kk < NS; ++kk) if (TNZ[j][kk]) w = FMA(TP[i][kk], Tm[j][kk], w);
                Pn[i][j] = w; Pn[j][i] = w;
            }
        /* innovation */
        double PZ[NS][3];
        for (int i = 0; i < NS; ++i) {
            PZ[i][0] = (Pn[i][SY] - Pn[i][SYL]) + Pn[i][SZ];
            PZ[i][1] = 4.0 * Pn[i][SPI];
            PZ[i][2] = 4.0 * Pn[i][SR];
        }
        const double F00 = ((PZ[SY][0] - PZ[SYL][0]) + PZ[SZ][0]) + E0;
        const double F10 = 4.0 * PZ[SPI][0];
        const double F11 = 4.0 * PZ[SPI][1] + E1;
        const double F20 = 4.0 * PZ[SR][0];
        const double F21 = 4.0 * PZ[SR][1];
        const double F22 = 4.0 * PZ[SR][2] + E2;
        const double *yt = data + (long)3 * t;
        const double n0 = (yt[0] - ((xn[SY] - xn[SYL]) + xn[SZ])) - D0;
        const double n1 = (yt[1] - 4.0 * xn[SPI]) - D1;
        const double n2 = (yt[2] - 4.0 * xn[SR]) - D2;
        /* missing observations (NaN): DSGE.jl's filter drops those rows of the measurement equation for the period --
         * the reduced system below (observed indices in ascending order); a period without any observation only predicts */
        int obs[3], m = 0;
        for (int k = 0; k < 3; ++k) if (yt[k] == yt[k]) obs[m++] = k;
        if (m < 3) {
            const double Ffull[3][3] = {{F00, F10, F20}, {F10, F11, F21}, {F20, F21, F22}};
            const double nfull[3] = {n0, n1, n2};
            double Lr[3][3] = {{0}}, dr[3] = {1, 1, 1}, rr[3] = {1, 1, 1}, wr[3] = {0, 0, 0}, vr[3] = {0, 0, 0};
            /* F_red = L D L' by the same elimination as the full 3 x 3 case: for a < m
             *   t_ab = F_ab - sum_{c<b} l_ac t'_bc ... written out for m <= 2 */
            if (m >= 1) { dr[0] = Ffull[obs[0]][obs[0]]; rr[0] = 1.0 / dr[0]; wr[0] = nfull[obs[0]]; }
            if (m == 2) {
                const double f10 = Ffull[obs[1]][obs[0]], f11 = Ffull[obs[1]][obs[1]];
                Lr[1][0] = f10 * rr[0];
                dr[1] = FMA(-Lr[1][0], f10, f11); rr[1] = 1.0 / dr[1];
                wr[1] = FMA(-Lr[1][0], wr[0], nfull[obs[1]]);
            }
            for (int a2 = 0; a2 < m; ++a2) { if (!(dr[a2] > 0.0)) bad = 1; vr[a2] = wr[a2] * rr[a2]; }
            if (t >= npre) {
                const double det = (m == 0) ? 1.0 : (m == 1 ? dr[0] : dr[0] * dr[1]);
                const double quad = (m == 0) ? 0.0 : (m == 1 ? wr[0] * vr[0] : FMA(wr[1], vr[1], wr[0] * vr[0]));
                ll = ll + -0.5 * (((double)m * 1.8378770664093453 + orc_log(det)) + quad);
            }
            double Hr[NS][2], HRr[NS][2];
            for (int i = 0; i < NS; ++i) {
                double xi = xn[i];
                if (m >= 1) { Hr[i][0] = PZ[i][obs[0]]; HRr[i][0] = Hr[i][0] * rr[0]; }
                if (m == 2) { Hr[i][1] = FMA(-Lr[1][0], Hr[i][0], PZ[i][obs[1]]); HRr[i][1] = Hr[i][1] * rr[1]; }
                if (m >= 1) xi = FMA(Hr[i][0], vr[0], xi);
                if (m == 2) xi = FMA(Hr[i][1], vr[1], FMA(Hr[i][0], vr[0], xn[i]));
                x[i] = xi;
            }
            for (int i = 0; i < NS; ++i)
                for (int j = 0; j <= i; ++j) {
                    double w = Pn[i][j];
                    if (m >= 1) w = FMA(-HRr[i][0], Hr[j][0], w);
                    if (m == 2) w = FMA(-HRr[i][1], Hr[j][1], FMA(-HRr[i][0], Hr[j][0], Pn[i][j]));
                    P[i][j] = w; P[j][i] = w;
                }
            continue;
        }
        /* F = L D L' (unit lower L, no square roots): d_k > 0 <=> F positive definite */
        const double d0 = F00, r0 = 1.0 / d0;
        const double l10 = F10 * r0, l20 = F20 * r0;
        const double d1 = FMA(-l10, F10, F11), r1 = 1.0 / d1;
        const double t21 = FMA(-l20, F10, F21);
        const double l21 = t21 * r1;
        const double d2 = FMA(-l21, t21, FMA(-l20, F20, F22)), r2 = 1.0 / d2;
        if (!(d0 > 0.0 && d1 > 0.0 && d2 > 0.0)) bad = 1;
        const double w0 = n0;
        const double w1 = FMA(-l10, w0, n1);
        const double w2 = FMA(-l21, w1, FMA(-l20, w0, n2));
        const double v0 = w0 * r0, v1 = w1 * r1, v2 = w2 * r2;
        if (t >= npre) {
            const double logdet = orc_log((d0 * d1) * d2);
            const double quad = FMA(w2, v2, FMA(w1, v1, w0 * v0));
            ll = ll + -0.5 * ((3.0 * 1.8378770664093453 + logdet) + quad);
        }
        /* update: H = PZ L^{-T}; x += H D^{-1} w; P -= H D^{-1} H' */
        double H[NS][3], HR[NS][3];
        for (int i = 0; i < NS; ++i) {
            H[i][0] = PZ[i][0];
            H[i][1] = FMA(-l10, H[i][0], PZ[i][1]);
            H[i][2] = FMA(-l21, H[i][1], FMA(-l20, H[i][0], PZ[i][2]));
            HR[i][0] = H[i][0] * r0; HR[i][1] = H[i][1] * r1; HR[i][2] = H[i][2] * r2;
            x[i] = FMA(H[i][2], v2, FMA(H[i][1], v1, FMA(H[i][0], v0, xn[i])));
        }
        for (int i = 0; i < NS; ++i)
            for (int j = 0; j <= i; ++j) {
                const double w = FMA(-HR[i][2], H[j][2], FMA(-HR[i][1], H[j][1], FMA(-HR[i][0], H[j][0], Pn[i][j])));
                P[i][j] = w; P[j][i] = w;
            }
    }
    if (bad) return -INFINITY;           /* innovation covariance not positive definite: an error in the reference => -Inf */
    if (!(ll == ll)) return -INFINITY;   /* NaN (non-PD innovation covariance etc.): an error in the reference => -Inf */
    return ll;
}
