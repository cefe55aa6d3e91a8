#!/usr/bin/env python
"""Prototype (numpy) of the reduced An-Schorfheide solver + structured 6-state Kalman filter that
oracle/smc_oracle.c (orc AS likelihood) and csrc/aslik.cuh implement; checked here against the
reference-produced (theta -> loglh) rows in tests/golden/as_clouds.npz and, when run in the build
container with scipy, against a QZ-based gensys of the full 8-state canonical form.

Development tool only (not imported by the package).  Usage:  python tools/as_reduced_prototype.py
"""
import math
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def cubic_stable_root(c3, c2, c1, c0):
    """Unique root inside the unit circle of c3 x^3 + c2 x^2 + c1 x + c0, or None when the number of roots
    with modulus < 1 is not exactly one (indeterminacy / non-existence)."""
    def p(x):
        return ((c3 * x + c2) * x + c1) * x + c0
    p0, p1, pm1 = p(0.0), p(1.0), p(-1.0)
    if p0 == 0.0:
        lo = hi = 0.0
    elif (p0 < 0.0) != (p1 < 0.0):
        lo, hi = 0.0, 1.0
    elif (p0 < 0.0) != (pm1 < 0.0):
        lo, hi = -1.0, 0.0
    else:
        return None
    flo = p(lo)
    for _ in range(64):
        mid = 0.5 * (lo + hi)
        fm = p(mid)
        if (fm < 0.0) == (flo < 0.0):
            lo, flo = mid, fm
        else:
            hi = mid
    r = 0.5 * (lo + hi)
    # deflate: c3 x^2 + q1 x + q0
    q1 = c2 + c3 * r
    q0 = c1 + q1 * r
    disc = q1 * q1 - 4.0 * c3 * q0
    if disc < 0.0:
        if q0 / c3 < 1.0:      # |x|^2 = q0 / c3
            return None
    else:
        sq = math.sqrt(disc)
        t = -0.5 * (q1 + (sq if q1 >= 0.0 else -sq))
        x1 = t / c3
        x2 = q0 / t if t != 0.0 else 0.0
        if abs(x1) < 1.0 or abs(x2) < 1.0:
            return None
    if not abs(r) < 1.0:
        return None
    return r


def solve_reduced(p):
    tau, kap, psi1, psi2, rA, pistar, gamQ, rhoR, rhog, rhoz, sigR, sigg, sigz, ey, epi, eR = p
    beta = 1.0 / (1.0 + rA / 400.0)
    b = 1.0 + beta + kap / tau
    k = (1.0 - rhoR) / tau
    c3 = beta
    c2 = -(b + rhoR * beta + k * psi2 * beta)
    c1 = 1.0 + rhoR * b + k * (psi1 * kap + psi2)
    c0 = -rhoR
    lam = cubic_stable_root(c3, c2, c1, c0)
    if lam is None:
        return None
    D = (beta * lam - b) * lam + 1.0
    a_y = -(lam / tau) * (1.0 - beta * lam) / D
    a_pi = -(lam * kap / tau) / D
    a_R = lam
    m = a_y - 1.0 / tau + a_pi / tau
    h = (1.0 - rhoR)
    # monetary shock loadings
    cpi_r = kap * m + beta * a_pi
    B_Rr = 1.0 / (1.0 - h * (psi1 * cpi_r + psi2 * m))
    B_yr = m * B_Rr
    B_pr = cpi_r * B_Rr
    # technology loadings: 3x3 system in (Y, P, Rz)
    #   (1-rz) Y - (rz/tau) P - m Rz = rz/tau ;  -kap Y + (1 - beta rz) P - beta a_pi Rz = 0 ;  -h psi2 Y - h psi1 P + Rz = 0
    A = np.array([[1.0 - rhoz, -rhoz / tau, -m], [-kap, 1.0 - beta * rhoz, -beta * a_pi], [-h * psi2, -h * psi1, 1.0]])
    rhs = np.array([rhoz / tau, 0.0, 0.0])
    B_yz, B_pz, B_Rz = np.linalg.solve(A, rhs)
    return dict(a_y=a_y, a_pi=a_pi, a_R=a_R, B_yr=B_yr, B_pr=B_pr, B_Rr=B_Rr, B_yz=B_yz, B_pz=B_pz, B_Rz=B_Rz)


def state_space6(p, s):
    tau, kap, psi1, psi2, rA, pistar, gamQ, rhoR, rhog, rhoz, sigR, sigg, sigz, ey, epi, eR = p
    y, pi, R, y1, g, z = range(6)
    T = np.zeros((6, 6)); Rm = np.zeros((6, 3))
    T[y, R] = s["a_y"]; T[y, g] = rhog; T[y, z] = s["B_yz"] * rhoz
    T[pi, R] = s["a_pi"]; T[pi, z] = s["B_pz"] * rhoz
    T[R, R] = s["a_R"]; T[R, z] = s["B_Rz"] * rhoz
    T[y1, y] = 1.0; T[g, g] = rhog; T[z, z] = rhoz
    Rm[:, 0] = [s["B_yz"], s["B_pz"], s["B_Rz"], 0, 0, 1]       # z_sh
    Rm[:, 1] = [1, 0, 0, 0, 1, 0]                                # g_sh
    Rm[:, 2] = [s["B_yr"], s["B_pr"], s["B_Rr"], 0, 0, 0]        # rm_sh
    Q = np.diag([sigz ** 2, sigg ** 2, sigR ** 2])
    Z = np.zeros((3, 6)); Z[0, y] = 1; Z[0, y1] = -1; Z[0, z] = 1; Z[1, pi] = 4; Z[2, R] = 4
    Dv = np.array([gamQ, pistar, pistar + rA + 4 * gamQ])
    E = np.diag([ey ** 2, epi ** 2, eR ** 2])
    return T, Rm, Q, Z, Dv, E


def stationary_cov(p, s, T):
    """Closed-form solution of P = T P T' + R Q R' (g is independent of (R, z); (R, z) is triangular)."""
    tau, kap, psi1, psi2, rA, pistar, gamQ, rhoR, rhog, rhoz, sigR, sigg, sigz, ey, epi, eR = p
    Sgg = sigg ** 2 / (1.0 - rhog * rhog)
    Szz = sigz ** 2 / (1.0 - rhoz * rhoz)
    aR, BRz, BRr = s["a_R"], s["B_Rz"], s["B_Rr"]
    SRz = BRz * Szz / (1.0 - aR * rhoz)
    SRR = (2.0 * aR * BRz * rhoz * SRz + BRz * BRz * Szz + BRr * BRr * sigR ** 2) / (1.0 - aR * aR)
    # v = (R_{t-1}, g_t, z_t, eps_R): covariance V
    V = np.zeros((4, 4))
    V[0, 0] = SRR; V[1, 1] = Sgg; V[2, 2] = Szz; V[3, 3] = sigR ** 2; V[0, 2] = V[2, 0] = rhoz * SRz
    M = np.array([[s["a_y"], 1.0, s["B_yz"], s["B_yr"]],
                  [s["a_pi"], 0.0, s["B_pz"], s["B_pr"]],
                  [s["a_R"], 0.0, s["B_Rz"], s["B_Rr"]],
                  [0.0, 1.0, 0.0, 0.0],
                  [0.0, 0.0, 1.0, 0.0]])
    P5 = M @ V @ M.T                                             # (y, pi, R, g, z)
    idx5 = [0, 1, 2, 4, 5]
    P = np.zeros((6, 6))
    for a, ia in enumerate(idx5):
        for b_, ib in enumerate(idx5):
            P[ia, ib] = P5[a, b_]
    # y1_t = y_{t-1}: Cov(y_{t-1}, x_t) = sum_k T[x,k] P[y,k] over the lag-free states
    P[3, 3] = P5[0, 0]
    for ia in idx5:
        cv = sum(T[ia, k] * P[0, k] for k in idx5)
        P[3, ia] = P[ia, 3] = cv
    return P


def loglik(p, data, npre=2):
    s = solve_reduced(p)
    if s is None:
        return -np.inf
    T, Rm, Q, Z, Dv, E = state_space6(p, s)
    RQR = Rm @ Q @ Rm.T
    P = stationary_cov(p, s, T)
    x = np.zeros(6); ll = 0.0
    for t in range(data.shape[1]):
        x = T @ x; P = T @ P @ T.T + RQR
        nu = data[:, t] - Z @ x - Dv
        F = Z @ P @ Z.T + E
        Fi = np.linalg.inv(F)
        if t >= npre:
            ll += -0.5 * (3 * math.log(2 * math.pi) + math.log(np.linalg.det(F)) + nu @ Fi @ nu)
        K = P @ Z.T @ Fi
        x = x + K @ nu; P = P - K @ Z @ P
    return ll


def main():
    g = np.load(os.path.join(HERE, "..", "tests", "golden", "as_clouds.npz"))
    data = g["data"]
    for name, col, T_ in (("prior_draws", 16, 230), ("cloud600", 16, 230), ("cloud600", 18, 115), ("cloud1000", 16, 115)):
        P = g[name]
        err = []
        for r in range(0, P.shape[0], 7):
            ll = loglik(P[r, :16], data[:, :T_])
            err.append(abs(ll - P[r, col]) / max(1.0, abs(P[r, col])))
        err = np.array(err)
        print("%-12s col %d T=%d: rows %d  median rel %.2e  max rel %.2e" % (name, col, T_, len(err), np.median(err), err.max()))
    # stationary covariance really solves the Lyapunov equation
    p = g["cloud600"][0, :16]
    s = solve_reduced(p)
    T, Rm, Q, Z, Dv, E = state_space6(p, s)
    P0 = stationary_cov(p, s, T)
    print("lyapunov residual", np.abs(P0 - (T @ P0 @ T.T + Rm @ Q @ Rm.T)).max())


if __name__ == "__main__":
    main()
