#!/usr/bin/env python
"""Independent check of the An-Schorfheide oracle against the textbook route the reference takes through DSGE.jl:
canonical 8-state form Gamma0 s_t = Gamma1 s_{t-1} + Psi eps_t + Pi eta_t, Sims' gensys by QZ
(scipy.linalg.ordqz), existence / uniqueness from the SVD rank conditions, dense 8-state Kalman filter with a
Lyapunov-equation initial covariance.  Used by tests/test_oracle_golden.py to pin (a) the determinacy decision of the
cubic-root solver (exactly one root inside the unit circle <=> gensys eu = [1, 1]; anything else => -Inf) on prior draws
that include the indeterminacy region psi_1 < 1, and (b) the likelihood values of the reduced 6-state filter.
Development / test tool (needs scipy); not imported by the package."""
import numpy as np
import scipy.linalg as sl


def eqcond(p):
    tau, kap, psi1, psi2, rA, pistar, gamQ, rhoR, rhog, rhoz, sigR, sigg, sigz, ey, epi, eR = p
    y, pi, R, y1, g, z, Ey, Epi = range(8)
    zsh, gsh, rmsh = range(3)
    G0 = np.zeros((8, 8)); G1 = np.zeros((8, 8)); Psi = np.zeros((8, 3)); Pi = np.zeros((8, 2))
    beta = 1 / (1 + rA / 400)
    G0[0, y] = 1; G0[0, R] = 1 / tau; G0[0, g] = -(1 - rhog); G0[0, z] = -rhoz / tau; G0[0, Ey] = -1; G0[0, Epi] = -1 / tau
    G0[1, y] = -kap; G0[1, pi] = 1; G0[1, g] = kap; G0[1, Epi] = -beta
    G0[2, y] = -(1 - rhoR) * psi2; G0[2, pi] = -(1 - rhoR) * psi1; G0[2, R] = 1; G0[2, g] = (1 - rhoR) * psi2
    G1[2, R] = rhoR; Psi[2, rmsh] = 1
    G0[3, y1] = 1; G1[3, y] = 1
    G0[4, g] = 1; G1[4, g] = rhog; Psi[4, gsh] = 1
    G0[5, z] = 1; G1[5, z] = rhoz; Psi[5, zsh] = 1
    G0[6, y] = 1; G1[6, Ey] = 1; Pi[6, 0] = 1
    G0[7, pi] = 1; G1[7, Epi] = 1; Pi[7, 1] = 1
    return G0, G1, Psi, Pi


def gensys(G0, G1, Psi, Pi, div=1.0 + 1e-8):
    """Sims (2002).  Returns (T, R, exist, unique)."""
    n = G0.shape[0]
    AA, BB, al, be, Q, Z = sl.ordqz(G0, G1, sort=lambda a, b: np.abs(b) <= div * np.abs(a), output="complex")
    q = Q.conj().T
    nunstab = int(np.sum(~(np.abs(be) <= div * np.abs(al))))
    ns = n - nunstab
    q1, q2 = q[:ns, :], q[ns:, :]
    eps = 1e-9
    etawt = q2 @ Pi
    u, d, vh = np.linalg.svd(etawt, full_matrices=False)
    big = d > eps
    ueta, deta, veta = u[:, big], np.diag(d[big]), vh.conj().T[:, big]
    exist = int(np.sum(big)) >= nunstab               # the unstable block can be spanned by the expectational errors
    etawt1 = q1 @ Pi
    u1, d1, vh1 = np.linalg.svd(etawt1, full_matrices=False)
    big1 = d1 > eps
    ueta1, deta1, veta1 = u1[:, big1], np.diag(d1[big1]), vh1.conj().T[:, big1]
    loose = veta1 - veta @ veta.conj().T @ veta1
    unique = np.linalg.norm(loose) < eps * n if loose.size else True
    if not (exist and unique):
        return None, None, exist, unique
    tmat = np.hstack([np.eye(ns), -(ueta @ np.linalg.solve(deta, veta.conj().T) @ veta1 @ deta1 @ ueta1.conj().T).conj().T])
    G0n = np.vstack([tmat @ AA, np.hstack([np.zeros((nunstab, ns)), np.eye(nunstab)])])
    G1n = np.vstack([tmat @ BB, np.zeros((nunstab, n))])
    G0I = np.linalg.inv(G0n)
    G1n = G0I @ G1n
    impact = G0I @ np.vstack([tmat @ q @ Psi, np.zeros((nunstab, Psi.shape[1]))])
    return np.real(Z @ G1n @ Z.conj().T), np.real(Z @ impact), exist, unique


def loglik8(p, data, npre=2):
    """8-state dense Kalman filter on the gensys solution; -inf without a unique stable solution."""
    T, R, ex, un = gensys(*eqcond(p))
    if T is None:
        return -np.inf
    tau, kap, psi1, psi2, rA, pistar, gamQ, rhoR, rhog, rhoz, sigR, sigg, sigz, ey, epi, eR = p
    ZZ = np.zeros((3, 8)); ZZ[0, 0] = 1; ZZ[0, 3] = -1; ZZ[0, 5] = 1; ZZ[1, 1] = 4; ZZ[2, 2] = 4
    DD = np.array([gamQ, pistar, pistar + rA + 4 * gamQ])
    EE = np.diag([ey ** 2, epi ** 2, eR ** 2]); QQ = np.diag([sigz ** 2, sigg ** 2, sigR ** 2])
    RQR = R @ QQ @ R.T
    P = sl.solve_discrete_lyapunov(T, RQR); s = np.zeros(8); ll = 0.0
    for t in range(data.shape[1]):
        s = T @ s; P = T @ P @ T.T + RQR
        ok = ~np.isnan(data[:, t])                      # missing observations: drop their rows for this period (DSGE.jl's filter)
        if not ok.any():
            continue
        Z, D, E = ZZ[ok], DD[ok], EE[np.ix_(ok, ok)]
        nu = data[ok, t] - Z @ s - D
        F = Z @ P @ Z.T + E
        Fi = np.linalg.inv(F)
        if t >= npre:
            ll += -0.5 * (ok.sum() * np.log(2 * np.pi) + np.log(np.linalg.det(F)) + nu @ Fi @ nu)
        K = P @ Z.T @ Fi
        s = s + K @ nu; P = P - K @ Z @ P
    return ll
