// entropic.cuh -- device-side entropic collision family (legacy CollisionModel interface of the reference).
//
//   nb_collide_kbc_d2q9          L/collision/KBCStandard.cpp:88-497   (KBC_C variant)
//   nb_collide_kbc_d3q15         L/collision/KBCStandard.cpp:499-1028
//   nb_collide_mrt_entropic_d3q19  L/collision/MRTEntropic.cpp:167-305
//
// Register-resident like collide.cuh: one DoF's populations per thread.  `v` carries the *scaled*
// macroscopic velocity the reference stores (input when in_init, output otherwise); cP.tau_legacy is
// nu/(dt*cs2_scaled) (L/collision/CollisionModel.h:152-157).  Reference quirks that a drop-in has to
// reproduce are marked QUIRK.
#pragma once
#include "collide.cuh"

// split f = k + s + h, entropic stabiliser gamma, relax: shared tail of both KBC variants
template <int Q>
__device__ __forceinline__ void nb_kbc_relax(double (&f)[Q], const double (&k)[Q], const double (&s)[Q],
                                             const double (&seq)[Q], const double (&feq)[Q], bool ratio_first)
{
    double ds[Q], dh[Q];
    double sum_s = 0.0, sum_h = 0.0;
#pragma unroll
    for (int q = 0; q < Q; q++) {
        const double h = f[q] - k[q] - s[q];
        const double heq = feq[q] - k[q] - seq[q];
        ds[q] = s[q] - seq[q];
        dh[q] = h - heq;
        const double a = ds[q] * dh[q] / feq[q];
        const double b = dh[q] * dh[q] / feq[q];
        sum_s = q == 0 ? a : sum_s + a;
        sum_h = q == 0 ? b : sum_h + b;
    }
    const double beta = 1. / (cP.tau_legacy + 0.5) / 2;
    // D2Q9 writes (..)*(sum_s/sum_h) (:427), D3Q15 writes (..)*sum_s/sum_h (:991)
    double gamma = ratio_first ? 1. / beta - (2 - 1. / beta) * (sum_s / sum_h) : 1. / beta - (2 - 1. / beta) * sum_s / sum_h;
    if (sum_h < 1e-16) gamma = 2;      // BGK fallback
#pragma unroll
    for (int q = 0; q < Q; q++) f[q] = f[q] - beta * (2 * ds[q] + gamma * dh[q]);
}

__device__ __forceinline__ void nb_collide_kbc_d2q9(double (&f)[9], double& rho, double (&v)[3], bool in_init)
{
    const double scaling = cP.scaling;
    rho = f[0] + f[1] + f[2] + f[3] + f[4] + f[5] + f[6] + f[7] + f[8];
    if (rho < 1e-10) return;
    if (!in_init) {
        v[0] = scaling / rho * (f[1] + f[5] + f[8] - f[3] - f[6] - f[7]);
        v[1] = scaling / rho * (f[2] + f[5] + f[6] - f[4] - f[7] - f[8]);
    }
    const double ux = v[0] / scaling, uy = v[1] / scaling;
    double T = f[1] + f[2] + f[3] + f[4] + 2 * (f[5] + f[6] + f[7] + f[8]);
    T = T / rho;
    double N = f[1] - f[2] + f[3] - f[4];
    N = N / rho;
    double Pi_xy = f[5] - f[6] + f[7] - f[8];
    Pi_xy = Pi_xy / rho;
    double k[9], s[9], seq[9], feq[9];
    k[0] = rho; k[1] = 0.5 * rho * ux; k[3] = 0.5 * rho * -ux; k[2] = 0.5 * rho * uy; k[4] = 0.5 * rho * -uy;
    k[5] = k[6] = k[7] = k[8] = 0;
    s[0] = -rho * T;
    s[1] = 0.5 * rho * 0.5 * (T + N); s[2] = 0.5 * rho * 0.5 * (T - N);
    s[3] = 0.5 * rho * 0.5 * (T + N); s[4] = 0.5 * rho * 0.5 * (T - N);
    s[5] = 0.25 * rho * Pi_xy; s[6] = 0.25 * rho * -Pi_xy; s[7] = 0.25 * rho * Pi_xy; s[8] = 0.25 * rho * -Pi_xy;
    // product-form entropic equilibrium (:250-271)
    const double r3 = sqrt(3.0);
    const double uxi = ux * r3, uyi = uy * r3;
    const double sx = sqrt(1 + uxi * uxi), sy = sqrt(1 + uyi * uyi);
    const double pre = rho * (2 - sx) * (2 - sy);
    const double px = (2 * uxi / r3 + sx) / (1 - uxi / r3);
    const double py = (2 * uyi / r3 + sy) / (1 - uyi / r3);
    feq[0] = 4. / 9. * pre;
    feq[1] = 1. / 9. * pre * px; feq[2] = 1. / 9. * pre * py;
    feq[3] = 1. / 9. * pre / px; feq[4] = 1. / 9. * pre / py;
    feq[5] = 1. / 36. * pre * px * py; feq[6] = 1. / 36. * pre / px * py;
    feq[7] = 1. / 36. * pre / px / py; feq[8] = 1. / 36. * pre * px / py;
    T = feq[1] + feq[2] + feq[3] + feq[4] + 2 * (feq[5] + feq[6] + feq[7] + feq[8]);
    T = T / rho;
    N = feq[1] - feq[2] + feq[3] - feq[4];
    N = N / rho;
    Pi_xy = feq[5] - feq[6] + feq[7] - feq[8];
    Pi_xy = Pi_xy / rho;
    seq[0] = -rho * T;
    seq[1] = 0.5 * rho * 0.5 * (T + N); seq[3] = 0.5 * rho * 0.5 * (T + N);
    seq[2] = 0.5 * rho * 0.5 * (T - N); seq[4] = 0.5 * rho * 0.5 * (T - N);
    seq[5] = 0.25 * rho * Pi_xy; seq[6] = 0.25 * rho * -Pi_xy; seq[7] = 0.25 * rho * Pi_xy; seq[8] = 0.25 * rho * -Pi_xy;
    nb_kbc_relax<9>(f, k, s, seq, feq, true);
}

__device__ __forceinline__ void nb_collide_kbc_d3q15(double (&f)[15], double& rho, double (&v)[3], bool in_init)
{
    const double scaling = cP.scaling;
    const double cs2 = cP.cs2 * scaling * scaling;     // QUIRK (:507): the scaled speed of sound ...
    const double prefactor = scaling / cs2;
    rho = f[0] + f[1] + f[2] + f[3] + f[4] + f[5] + f[6] + f[7] + f[8] + f[9] + f[10] + f[11] + f[12] + f[13] + f[14];
    if (rho < 1e-10) return;
    if (!in_init) {
        v[0] = scaling / rho * (f[1] - f[2] + f[7] - f[8] + f[9] - f[10] + f[11] - f[12] + f[13] - f[14]);
        v[1] = scaling / rho * (f[3] - f[4] + f[7] - f[8] + f[9] - f[10] - f[11] + f[12] - f[13] + f[14]);
        v[2] = scaling / rho * (f[5] - f[6] + f[7] - f[8] - f[9] + f[10] + f[11] - f[12] - f[13] + f[14]);
    }
    const double ux = v[0] / scaling, uy = v[1] / scaling, uz = v[2] / scaling;
    double T = f[1] + f[2] + f[3] + f[4] + f[5] + f[6] + 3 * f[7] + 3 * f[8] + 3 * f[9] + 3 * f[10] + 3 * f[11] + 3 * f[12]
        + 3 * f[13] + 3 * f[14];
    T /= rho;
    double N_xz = f[1] + f[2] - f[5] - f[6];
    N_xz /= rho;
    double N_yz = f[3] + f[4] - f[5] - f[6];
    N_yz /= rho;
    double Q_xyz = f[7] - f[8] - f[9] + f[10] - f[11] + f[12] + f[13] - f[14];
    Q_xyz /= rho;
    double k[15], s[15], seq[15], feq[15];
    k[0] = rho;
    k[1] = rho / 6 * (3 * ux); k[2] = rho / 6 * (3 * -ux);
    k[3] = rho / 6 * (3 * uy); k[4] = rho / 6 * (3 * -uy);
    k[5] = rho / 6 * (3 * uz); k[6] = rho / 6 * (3 * -uz);
#pragma unroll
    for (int q = 7; q < 15; q++) k[q] = 0;
    s[0] = rho * -T;
    s[1] = 1. / 6. * rho * (2 * N_xz - N_yz + T); s[2] = s[1];
    s[3] = 1. / 6. * rho * (-N_xz + 2 * N_yz + T); s[4] = s[3];
    s[5] = 1. / 6. * rho * (-N_xz - N_yz + T); s[6] = s[5];
    s[7] = 1. / 8. * rho * Q_xyz;
    s[8] = -s[7]; s[9] = -s[7]; s[10] = s[7]; s[11] = -s[7]; s[12] = s[7]; s[13] = s[7]; s[14] = -s[7];
    // second-order polynomial equilibrium; QUIRK (:749-750): u^2 term = unscaled u^2 over the scaled cs2
    const double uSquareTerm = -(ux * ux + uy * uy + uz * uz) / (2 * cs2);
    double weighting = 2. / 9. * rho, mixedTerm;
    feq[0] = weighting * (1 + uSquareTerm);
    weighting = 1. / 9. * rho;
    mixedTerm = prefactor * (v[0]);
    feq[1] = weighting * (1 + mixedTerm * (1 + 0.5 * mixedTerm) + uSquareTerm);
    feq[2] = weighting * (1 - mixedTerm * (1 - 0.5 * mixedTerm) + uSquareTerm);
    mixedTerm = prefactor * (v[1]);
    feq[3] = weighting * (1 + mixedTerm * (1 + 0.5 * mixedTerm) + uSquareTerm);
    feq[4] = weighting * (1 - mixedTerm * (1 - 0.5 * mixedTerm) + uSquareTerm);
    mixedTerm = prefactor * (v[2]);
    feq[5] = weighting * (1 + mixedTerm * (1 + 0.5 * mixedTerm) + uSquareTerm);
    feq[6] = weighting * (1 - mixedTerm * (1 - 0.5 * mixedTerm) + uSquareTerm);
    weighting = 1. / 72. * rho;
    mixedTerm = prefactor * (v[0] + v[1] + v[2]);
    feq[7] = weighting * (1 + mixedTerm * (1 + 0.5 * mixedTerm) + uSquareTerm);
    feq[8] = weighting * (1 - mixedTerm * (1 - 0.5 * mixedTerm) + uSquareTerm);
    mixedTerm = prefactor * (v[0] + v[1] - v[2]);
    feq[9] = weighting * (1 + mixedTerm * (1 + 0.5 * mixedTerm) + uSquareTerm);
    feq[10] = weighting * (1 - mixedTerm * (1 - 0.5 * mixedTerm) + uSquareTerm);
    mixedTerm = prefactor * (v[0] - v[1] + v[2]);
    feq[11] = weighting * (1 + mixedTerm * (1 + 0.5 * mixedTerm) + uSquareTerm);
    feq[12] = weighting * (1 - mixedTerm * (1 - 0.5 * mixedTerm) + uSquareTerm);
    mixedTerm = prefactor * (v[0] - v[1] - v[2]);
    feq[13] = weighting * (1 + mixedTerm * (1 + 0.5 * mixedTerm) + uSquareTerm);
    feq[14] = weighting * (1 - mixedTerm * (1 - 0.5 * mixedTerm) + uSquareTerm);
    T = feq[1] + feq[2] + feq[3] + feq[4] + feq[5] + feq[6] + 3 * feq[7] + 3 * feq[8] + 3 * feq[9] + 3 * feq[10] + 3 * feq[11]
        + 3 * feq[12] + 3 * feq[13] + 3 * feq[14];
    T /= rho;
    N_xz = feq[1] + feq[2] - feq[5] - feq[6];
    N_xz /= rho;
    N_yz = feq[3] + feq[4] - feq[5] - feq[6];
    N_yz /= rho;
    Q_xyz = feq[7] - feq[8] - feq[9] + feq[10] - feq[11] + feq[12] + feq[13] - feq[14];
    Q_xyz /= rho;
    seq[0] = rho * -T;
    seq[1] = 1. / 6. * rho * (2 * N_xz - N_yz + T); seq[2] = seq[1];
    seq[3] = 1. / 6. * rho * (-N_xz + 2 * N_yz + T); seq[4] = seq[3];
    seq[5] = 1. / 6. * rho * (-N_xz - N_yz + T); seq[6] = seq[5];
    seq[7] = 1. / 8. * rho * Q_xyz;
    seq[8] = -seq[7]; seq[9] = -seq[7]; seq[10] = seq[7]; seq[11] = -seq[7]; seq[12] = seq[7]; seq[13] = seq[7];
    seq[14] = -seq[7];
    nb_kbc_relax<15>(f, k, s, seq, feq, false);
}

// moment matrix of MRTEntropic.cpp:172-198 and its inverse :200-218, filled by nb200_set_collision
struct NbMrtTables {
    double tm[19][19];
    double invm[19][19];
};
__constant__ NbMrtTables cM;

__device__ __forceinline__ void nb_collide_mrt_entropic_d3q19(double (&f)[19], double& rho, double (&v)[3], bool in_init)
{
    double m[19];
#pragma unroll
    for (int p = 0; p < 19; p++) {
        double acc = 0;
#pragma unroll
        for (int q = 0; q < 19; q++) acc += cM.tm[p][q] * f[q];
        m[p] = acc;
    }
    rho = m[0];
    const double jx = m[3], jy = m[5], jz = m[7];
    if (!in_init) {
        v[0] = cP.scaling / rho * jx;
        v[1] = cP.scaling / rho * jy;
        v[2] = cP.scaling / rho * jz;
    }
    const double w = -1. / (cP.tau_legacy + 0.5);
    const double meq1 = -11 * rho + 19. / rho * (jx * jx + jy * jy + jz * jz);
    const double meq9 = 1. / rho * (2 * jx * jx - (jy * jy + jz * jz));
    const double meq11 = 1. / rho * (jy * jy - jz * jz);
    const double meq13 = 1. / rho * jx * jy;
    const double meq14 = 1. / rho * jy * jz;
    const double meq15 = 1. / rho * jx * jz;
    m[1] = m[1] + w * (m[1] - meq1);
    m[9] = m[9] + w * (m[9] - meq9);
    m[11] = m[11] + w * (m[11] - meq11);
    m[13] = m[13] + w * (m[13] - meq13);
    m[14] = m[14] + w * (m[14] - meq14);
    m[15] = m[15] + w * (m[15] - meq15);
    m[2] = -7. / 38 * rho - 11. / 38 * m[1];
    m[4] = -2. / 3. * jx;
    m[6] = -2. / 3. * jy;
    m[8] = -2. / 3. * jz;
    m[10] = -1. / 2. * m[9];
    m[12] = -1. / 2. * m[11];
    m[16] = 0; m[17] = 0; m[18] = 0;
#pragma unroll
    for (int p = 0; p < 19; p++) {
        double acc = 0;
#pragma unroll
        for (int q = 0; q < 19; q++) acc += cM.invm[p][q] * m[q];
        f[p] = acc;
    }
}

// One entry point for the f-only collisions.  KIND: NB_EQ_BGK / NB_EQ_QUARTIC (collision_advanced BGK),
// NB_KIND_KBC, NB_KIND_MRT_ENTROPIC.  v: scaled velocity (in when in_init, out otherwise).
// rho_prev: density stored by the previous call.  Returns true if the reference would throw
// CollisionException for this DoF.
template <int D, int Q, int KIND>
__device__ __forceinline__ bool nb_collide_f(double (&f)[Q], double& rho, double (&v)[3], bool in_init, double rho_prev)
{
    if constexpr (KIND == NB_KIND_KBC && D == 2 && Q == 9) {
        nb_collide_kbc_d2q9(f, rho, v, in_init);
        return rho < 1e-10;
    } else if constexpr (KIND == NB_KIND_KBC && D == 3 && Q == 15) {
        nb_collide_kbc_d3q15(f, rho, v, in_init);
        return rho < 1e-10;
    } else if constexpr (KIND == NB_KIND_MRT_ENTROPIC && D == 3 && Q == 19) {
        // QUIRK (MRTEntropic.cpp:229-232): the guard reads the density of the previous call
        if (rho_prev < 1e-10) { rho = rho_prev; return true; }
        nb_collide_mrt_entropic_d3q19(f, rho, v, in_init);
        return false;
    } else if constexpr (KIND == NB_KIND_REGULARIZED) {
        nb_collide_adv<D, Q, NB_EQ_BGK, 1, false>(f, rho, v, in_init);
        return rho < 1e-10;
    } else if constexpr (KIND == NB_KIND_MRT && Q <= NB_MRT_MAXQ) {
        nb_collide_adv<D, Q, NB_EQ_BGK, 2, false>(f, rho, v, in_init);
        return rho < 1e-10;
    } else {
        double u[3];
        nb_collide_bgk<D, Q, KIND == NB_EQ_QUARTIC ? NB_EQ_QUARTIC : NB_EQ_BGK>(f, rho, u, in_init ? v : nullptr);
        if (!in_init) {
#pragma unroll
            for (int j = 0; j < D; j++) v[j] = u[j] * cP.scaling;
        }
        return rho < 1e-10;
    }
}

// Same with the external-force hooks compiled in (stand-alone collide kernels only: a forced problem runs stream
// and collide as two kernels).  Entropic kinds have no force hooks in the reference's collideAll.
template <int D, int Q, int KIND>
__device__ __forceinline__ bool nb_collide_f_forced(double (&f)[Q], double& rho, double (&v)[3], bool in_init)
{
    if constexpr (KIND == NB_KIND_REGULARIZED) nb_collide_adv<D, Q, NB_EQ_BGK, 1, true>(f, rho, v, in_init);
    else if constexpr (KIND == NB_KIND_MRT && Q <= NB_MRT_MAXQ) nb_collide_adv<D, Q, NB_EQ_BGK, 2, true>(f, rho, v, in_init);
    else nb_collide_adv<D, Q, KIND == NB_EQ_QUARTIC ? NB_EQ_QUARTIC : NB_EQ_BGK, 0, true>(f, rho, v, in_init);
    return rho < 1e-10;
}
