"""Oracle restatement of the synthetic initial fields (test infrastructure only).

  TGV2D  L/benchmarks/TaylorGreenVortex2D.cpp:36-60  (t = 0, L = 2 pi, horizontal velocity 0)
  TGV3D  L/benchmarks/TaylorGreenVortex3D.cpp:38-72
  f init L/solver/CFDSolver.cpp:1104-1129 -> BGKStandard::getEquilibriumDistribution
         (L/collision/BGKStandard.cpp:23-41; scaled e, scaled cs2, physical u)
  f,g init (compressible) L/solver/CompressibleCFDSolver.h:790-901 (own quartic formula,
         g = feq * T * (2 Cv - dim))
"""
import numpy as np


def tgv2d(x, cs=None, init_rho_analytically=False, L=2 * np.pi):
    k = 2 * np.pi / L
    u = np.stack([np.sin(x[:, 0] * k) * np.cos(x[:, 1] * k),
                  -np.cos(x[:, 0] * k) * np.sin(x[:, 1] * k)])
    if init_rho_analytically:
        p = 1.0 / 4. * (np.cos(2 * (x[:, 0] * k)) + np.cos(2 * x[:, 1] * k))
        rho = 1.0 + p / (cs * cs)
    else:
        rho = np.ones(x.shape[0])
    return rho, u


def tgv3d(x, cs=None, init_rho_analytically=True, compressible=False, density_numerator=1.0):
    u = np.stack([np.sin(x[:, 0]) * np.cos(x[:, 1]) * np.cos(x[:, 2]),
                  -np.cos(x[:, 0]) * np.sin(x[:, 1]) * np.cos(x[:, 2]),
                  np.zeros(x.shape[0])])
    if init_rho_analytically:
        p = 1.0 / 16. * (np.cos(2 * x[:, 0]) + np.cos(2 * x[:, 1])) * (np.cos(2 * x[:, 2]) + 2)
        rho = 1.0 + p * density_numerator if compressible else 1.0 + p / (cs * cs)
    else:
        rho = np.ones(x.shape[0])
    T = np.ones(x.shape[0])
    return rho, u, T


def equilibrium_init(e, w, cs2, rho, u):
    """f = feq(rho, u), legacy BGKStandard formula, vectorised over DoFs.  Returns (Q, N)."""
    Q = e.shape[0]
    uu = -(u * u).sum(axis=0) / (2 * cs2)
    f = np.empty((Q, rho.shape[0]))
    for i in range(Q):
        pref = w[i] * rho
        if i == 0:
            f[i] = pref * (1 + uu)
            continue
        mixed = (e[i][:, None] * u).sum(axis=0) / cs2
        f[i] = pref * (1 + mixed * (1 + 0.5 * mixed) + uu)
    return f


def quartic_equilibrium_init(e_scaled, w, cs2_scaled, scaling, rho, u, T, gamma):
    """CompressibleCFDSolver::calcQuarticEquilibrium + initializeDistributions (:790-901),
    literal loop nest, vectorised over DoFs.  u is the physical velocity (divided by scaling
    inside, :885).  Returns f, g of shape (Q, N)."""
    Q, dim = e_scaled.shape
    e = e_scaled / scaling
    cs2 = cs2_scaled / (scaling * scaling)
    v = u / scaling
    eye = np.eye(dim)
    uu_term = np.zeros_like(rho)
    for j in range(dim):
        uu_term += -(v[j] * v[j]) / (2.0 * cs2)
    T1 = cs2 * (T - 1)
    f = np.empty((Q, rho.shape[0]))
    for i in range(Q):
        ue = np.zeros_like(rho)
        for j in range(dim):
            ue += (v[j] * e[i][j]) / cs2
        fe = w[i] * rho * (1 + ue * (1 + 0.5 * ue) + uu_term)
        for a in range(dim):
            for b in range(dim):
                fe = fe + rho * w[i] / (2.0 * cs2) * ((T - 1) * eye[a][b] * e[i][a] * e[i][b] - cs2 * eye[a][b] * (T - 1))
                for c in range(dim):
                    fe = fe + w[i] * rho / (6. * cs2 * cs2 * cs2) * (
                        v[a] * v[b] * v[c] + T1 * (eye[a][b] * v[c] + eye[b][c] * v[a] + eye[a][c] * v[b])
                    ) * (e[i][a] * e[i][b] * e[i][c] - cs2 * (e[i][c] * eye[a][b] + e[i][b] * eye[a][c] + e[i][a] * eye[b][c]))
                    for d in range(dim):
                        power4 = e[i][a] * e[i][b] * e[i][c] * e[i][d]
                        power2 = (e[i][a] * e[i][b] * eye[c][d] + e[i][a] * e[i][c] * eye[b][d]
                                  + e[i][a] * e[i][d] * eye[b][c] + e[i][b] * e[i][c] * eye[a][d]
                                  + e[i][b] * e[i][d] * eye[a][c] + e[i][c] * e[i][d] * eye[a][b])
                        power0 = eye[a][b] * eye[c][d] + eye[a][c] * eye[b][d] + eye[a][d] * eye[b][c]
                        u4 = v[a] * v[b] * v[c] * v[d]
                        u2 = (v[a] * v[b] * eye[c][d] + v[a] * v[c] * eye[b][d] + v[a] * v[d] * eye[b][c]
                              + v[b] * v[c] * eye[a][d] + v[b] * v[d] * eye[a][c] + v[c] * v[d] * eye[a][b])
                        fe = fe + w[i] * rho / (24. * cs2 ** 4) * (power4 - cs2 * power2 + cs2 * cs2 * power0) * (
                            u4 + T1 * (u2 + T1 * power0))
        f[i] = fe
    C_v = 1. / (gamma - 1.0)
    g = f * T * (2.0 * C_v - dim)
    return f, g
