"""Oracle restatement of the velocity stencils on the hot path (test infrastructure only).

Each function returns ``(e, w, cs2, max_speed)`` with ``e`` of shape (Q, D) *scaled*
directions, ``w`` the weights, ``cs2`` the scaled speed of sound squared and
``max_speed`` = ``getMaxParticleVelocityMagnitude()`` (used only by the dt formula).

Follows (L = src/library/natrium):
  D2Q9    L/stencils/D2Q9.cpp:28-61,   D2Q9.h:101-103
  D3Q19   L/stencils/D3Q19.cpp:25-67,  D3Q19.h:154-155
  D3Q15   L/stencils/D3Q15.cpp:25-73,  D3Q15.h:141-143
  D2Q25H  L/stencils/D2Q25H.cpp:25-79, D2Q25H.h:149-150
  D3Q45   L/stencils/D3Q45.cpp:25-154, D3Q45.h:211-212
"""
import math

import numpy as np


def d2q9(scaling=1.0):
    s = scaling
    e = np.array([[0, 0], [s, 0], [0, s], [-s, 0], [0, -s],
                  [s, s], [-s, s], [-s, -s], [s, -s]], dtype=np.float64)
    w = np.array([4. / 9.] + [1. / 9.] * 4 + [1. / 36.] * 4)
    return e, w, s * s / 3., math.sqrt(2) * s


def d3q19(scaling=1.0):
    s = scaling
    e = np.array([[0, 0, 0],
                  [s, 0, 0], [0, 0, s], [-s, 0, 0], [0, 0, -s], [0, -s, 0], [0, s, 0],
                  [s, 0, s], [-s, 0, s], [-s, 0, -s], [s, 0, -s],
                  [s, -s, 0], [s, s, 0], [-s, s, 0], [-s, -s, 0],
                  [0, -s, s], [0, s, s], [0, s, -s], [0, -s, -s]], dtype=np.float64)
    w = np.array([1. / 3.] + [1. / 18.] * 6 + [1. / 36.] * 12)
    return e, w, s * s / 3., math.sqrt(2) * s


def d3q15(scaling=1.0):
    s = scaling
    e = np.array([[0, 0, 0],
                  [s, 0, 0], [-s, 0, 0], [0, s, 0], [0, -s, 0], [0, 0, s], [0, 0, -s],
                  [s, s, s], [-s, -s, -s], [s, s, -s], [-s, -s, s],
                  [s, -s, s], [-s, s, -s], [s, -s, -s], [-s, s, s]], dtype=np.float64)
    w = np.array([2. / 9.] + [1. / 9.] * 6 + [1. / 72.] * 8)
    return e, w, s * s / 3., math.sqrt(3) * s


def d2q25h(scaling=1.0):
    s = scaling
    r = (math.sqrt(5.) - math.sqrt(2.)) / math.sqrt(3.)
    w_0 = (-3 - 3 * r * r * r * r + 54 * r * r) / (75 * r * r)
    w_m = (9 * r * r * r * r - 6 - 27 * r * r) / (300 * r * r * (r * r - 1))
    w_n = (9 - 6 * r * r * r * r - 27 * r * r) / (300 * (1 - r * r))
    w_0n, w_0m, w_mm, w_mn, w_nn = w_0 * w_n, w_0 * w_m, w_m * w_m, w_m * w_n, w_n * w_n
    w = np.array([w_0 * w_0] + [w_0m] * 4 + [w_mm] * 4 + [w_0n] * 4 + [w_nn] * 4 + [w_mn] * 8)
    c_m = s * math.sqrt(5. - math.sqrt(10.)) / math.sqrt(3.)
    c_n = math.sqrt(5. + math.sqrt(10.)) * s / math.sqrt(3.)
    e = np.array([[0, 0], [c_m, 0], [0, c_m], [-c_m, 0], [0, -c_m],
                  [c_m, c_m], [-c_m, c_m], [-c_m, -c_m], [c_m, -c_m],
                  [c_n, 0], [0, c_n], [-c_n, 0], [0, -c_n],
                  [c_n, c_n], [-c_n, c_n], [-c_n, -c_n], [c_n, -c_n],
                  [c_m, c_n], [c_m, -c_n], [-c_m, -c_n], [-c_m, c_n],
                  [c_n, c_m], [c_n, -c_m], [-c_n, -c_m], [-c_n, c_m]], dtype=np.float64)
    return e, w, s * s / 3., math.sqrt(2) * s


_D3Q45_A, _D3Q45_B = 0.06386083877343968, 1.2239121278243665
_D3Q45_C, _D3Q45_D = 1.5766994272507744, 0.5069610024977665
_D3Q45_E, _D3Q45_F = 2.9239876105912574, 0.4744978678080795
_D3Q45_G = 1.7320508075688787
_D3Q45_H, _D3Q45_I, _D3Q45_J = 2.403092127540177, 0.8892242114059369, 1.5602655313772367
_D3Q45_K, _D3Q45_L = 2.7367507163016924, 0.14279717659756475
_D3Q45_M, _D3Q45_N = 3.5256070994177073, 1.1335992635264445


def d3q45(scaling=1.0):
    A, B, C, D, E, F, G = _D3Q45_A, _D3Q45_B, _D3Q45_C, _D3Q45_D, _D3Q45_E, _D3Q45_F, _D3Q45_G
    H, I, J, K, L, M, N = _D3Q45_H, _D3Q45_I, _D3Q45_J, _D3Q45_K, _D3Q45_L, _D3Q45_M, _D3Q45_N
    raw = np.array([
        [0, 0, 0],
        [A, -B, -B], [-B, A, -B], [-B, -B, A],
        [C, -D, -D], [-D, C, -D], [-D, -D, C],
        [D, D, -C], [D, -C, D], [-C, D, D],
        [B, B, -A], [B, -A, B], [-A, B, B],
        [E, F, F], [F, E, F], [F, F, E],
        [G, G, G],
        [H, I, -J], [H, -J, I], [J, -I, -H], [J, -H, -I], [I, H, -J], [I, -J, H],
        [-I, J, -H], [-I, -H, J], [-J, H, I], [-J, I, H], [-H, J, -I], [-H, -I, J],
        [-G, -G, -G],
        [-F, -F, -E], [-F, -E, -F], [-E, -F, -F],
        [K, K, -L], [K, -L, K], [-L, K, K],
        [-M, N, N], [N, -M, N], [N, N, -M],
        [-N, -N, M], [-N, M, -N], [M, -N, -N],
        [L, -K, -K], [-K, L, -K], [-K, -K, L]], dtype=np.float64)
    e = scaling * raw / math.sqrt(3)
    w = np.array([0.20740740740740618] + [0.05787037037037047] * 12
                 + [0.00462962962962958] * 20 + [0.0004629629629629939] * 12)
    return e, w, scaling * scaling / 3., math.sqrt(2) * scaling


STENCILS = {"D2Q9": d2q9, "D3Q19": d3q19, "D3Q15": d3q15, "D2Q25H": d2q25h, "D3Q45": d3q45}


def make(name, scaling=1.0):
    return STENCILS[name](scaling)
