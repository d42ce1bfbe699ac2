/*
 * entropic_oracle.c -- CPU restatement of NATriuM's entropic collision family (legacy CollisionModel
 * interface).  TEST INFRASTRUCTURE ONLY (see natrium_oracle.c).  L = src/library/natrium.
 *
 *   orc_collide_kbc_d2q9        L/collision/KBCStandard.cpp:88-497   (KBC_C variant, the one #defined there)
 *   orc_collide_kbc_d3q15       L/collision/KBCStandard.cpp:499-1028
 *   orc_collide_mrt_entropic_d3q19  L/collision/MRTEntropic.cpp:167-305
 *
 * tau is the legacy relaxation parameter nu/(dt*cs2_scaled) (L/collision/CollisionModel.h:152-157);
 * beta = 1/(tau+0.5)/2.  Populations are addressed as f[q*stride + i]; velocities are the *scaled*
 * macroscopic velocities the reference stores (u[d*n + i]).  Reference quirks are kept on purpose and
 * marked QUIRK.  Every expression keeps the reference's order of operations (-ffp-contract=off).
 *
 * Return value: 0, or 1 if a density below 1e-10 was met (the reference throws CollisionException).
 */
#include <math.h>
#include <stdint.h>

int orc_entropic_available(void) { return 1; }

int orc_collide_kbc_d2q9(int64_t n, int64_t stride, double *f, double *rho_out, double *u, double scaling,
                         double tau, int in_init)
{
    int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad) if (n > 50000)
    for (int64_t i = 0; i < n; ++i) {
        double F[9], k[9], s[9], h[9], feq[9], seq[9], heq[9], ds[9], dh[9];
        for (int q = 0; q < 9; q++) F[q] = f[q * stride + i];
        /* :113-116 */
        const double rho = F[0] + F[1] + F[2] + F[3] + F[4] + F[5] + F[6] + F[7] + F[8];
        rho_out[i] = rho;
        if (rho < 1e-10) { bad |= 1; continue; }
        if (!in_init) {   /* :126-134 */
            u[0 * n + i] = scaling / rho * (F[1] + F[5] + F[8] - F[3] - F[6] - F[7]);
            u[1 * n + i] = scaling / rho * (F[2] + F[5] + F[6] - F[4] - F[7] - F[8]);
        }
        const double ux = u[0 * n + i] / scaling;
        const double uy = u[1 * n + i] / scaling;
        /* moments :141-166 */
        double T = F[1] + F[2] + F[3] + F[4] + 2 * (F[5] + F[6] + F[7] + F[8]);
        T = T / rho;
        double N = F[1] - F[2] + F[3] - F[4];
        N = N / rho;
        double Pi_xy = F[5] - F[6] + F[7] - F[8];
        Pi_xy = Pi_xy / rho;
        /* k :171-179, s (KBC_C) :186-194 */
        k[0] = rho; k[1] = 0.5 * rho * ux; k[3] = 0.5 * rho * -ux; k[2] = 0.5 * rho * uy; k[4] = 0.5 * rho * -uy;
        k[5] = k[6] = k[7] = k[8] = 0;
        s[0] = -rho * T;
        s[1] = 0.5 * rho * 0.5 * (T + N); s[2] = 0.5 * rho * 0.5 * (T - N);
        s[3] = 0.5 * rho * 0.5 * (T + N); s[4] = 0.5 * rho * 0.5 * (T - N);
        s[5] = 0.25 * rho * Pi_xy; s[6] = 0.25 * rho * -Pi_xy; s[7] = 0.25 * rho * Pi_xy; s[8] = 0.25 * rho * -Pi_xy;
        for (int q = 0; q < 9; q++) h[q] = F[q] - k[q] - s[q];   /* :235-243 */
        /* product-form equilibrium :253-275 */
        const double u_x_i = ux * sqrt(3);
        const double u_y_i = uy * sqrt(3);
        const double sqrt_ux = sqrt(1 + u_x_i * u_x_i);
        const double sqrt_uy = sqrt(1 + u_y_i * u_y_i);
        const double prefactor = rho * (2 - sqrt_ux) * (2 - sqrt_uy);
        const double px = (2 * u_x_i / sqrt(3) + sqrt_ux) / (1 - u_x_i / sqrt(3));
        const double py = (2 * u_y_i / sqrt(3) + sqrt_uy) / (1 - u_y_i / sqrt(3));
        feq[0] = 4. / 9. * prefactor;
        feq[1] = 1. / 9. * prefactor * px;
        feq[2] = 1. / 9. * prefactor * py;
        feq[3] = 1. / 9. * prefactor / px;
        feq[4] = 1. / 9. * prefactor / py;
        feq[5] = 1. / 36. * prefactor * px * py;
        feq[6] = 1. / 36. * prefactor / px * py;
        feq[7] = 1. / 36. * prefactor / px / py;
        feq[8] = 1. / 36. * prefactor * px / py;
        /* moments of feq :279-300 */
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
        for (int q = 0; q < 9; q++) heq[q] = feq[q] - k[q] - seq[q];   /* :355-363 */
        double sum_s = 0, sum_h = 0;
        {   /* :376-421, left-to-right sums */
            double a[9], b[9];
            for (int q = 0; q < 9; q++) {
                ds[q] = s[q] - seq[q];
                dh[q] = h[q] - heq[q];
                a[q] = ds[q] * dh[q] / feq[q];
                b[q] = dh[q] * dh[q] / feq[q];
            }
            sum_s = a[0] + a[1] + a[2] + a[3] + a[4] + a[5] + a[6] + a[7] + a[8];
            sum_h = b[0] + b[1] + b[2] + b[3] + b[4] + b[5] + b[6] + b[7] + b[8];
        }
        const double beta = 1. / (tau + 0.5) / 2;            /* :424 */
        double gamma = 1. / beta - (2 - 1. / beta) * (sum_s / sum_h);   /* :427 */
        if (sum_h < 1e-16) gamma = 2;                        /* :430-433 */
        for (int q = 0; q < 9; q++) f[q * stride + i] = F[q] - beta * (2 * ds[q] + gamma * dh[q]);   /* :446-490 */
    }
    return bad;
}

int orc_collide_kbc_d3q15(int64_t n, int64_t stride, double *f, double *rho_out, double *u, double scaling,
                          double cs2_scaled, double tau, int in_init)
{
    int bad = 0;
    const double cs2 = cs2_scaled;              /* :507  QUIRK: the *scaled* cs2 ... */
    const double prefactor = scaling / cs2;     /* :508 */
#pragma omp parallel for schedule(static) reduction(| : bad) if (n > 50000)
    for (int64_t i = 0; i < n; ++i) {
        double F[15], k[15], s[15], h[15], feq[15], seq[15], heq[15], ds[15], dh[15];
        for (int q = 0; q < 15; q++) F[q] = f[q * stride + i];
        const double rho = F[0] + F[1] + F[2] + F[3] + F[4] + F[5] + F[6] + F[7] + F[8] + F[9] + F[10] + F[11] + F[12]
            + F[13] + F[14];                    /* :563-566 */
        rho_out[i] = rho;
        if (rho < 1e-10) { bad |= 1; continue; }
        if (!in_init) {                         /* :576-587 */
            u[0 * n + i] = scaling / rho * (F[1] - F[2] + F[7] - F[8] + F[9] - F[10] + F[11] - F[12] + F[13] - F[14]);
            u[1 * n + i] = scaling / rho * (F[3] - F[4] + F[7] - F[8] + F[9] - F[10] - F[11] + F[12] - F[13] + F[14]);
            u[2 * n + i] = scaling / rho * (F[5] - F[6] + F[7] - F[8] - F[9] + F[10] + F[11] - F[12] - F[13] + F[14]);
        }
        const double vx = u[0 * n + i], vy = u[1 * n + i], vz = u[2 * n + i];
        const double ux = vx / scaling, uy = vy / scaling, uz = vz / scaling;
        /* moments :593-643 (only those that enter k, s) */
        double T = F[1] + F[2] + F[3] + F[4] + F[5] + F[6] + 3 * F[7] + 3 * F[8] + 3 * F[9] + 3 * F[10] + 3 * F[11]
            + 3 * F[12] + 3 * F[13] + 3 * F[14];
        T /= rho;
        double N_xz = F[1] + F[2] - F[5] - F[6];
        N_xz /= rho;
        double N_yz = F[3] + F[4] - F[5] - F[6];
        N_yz /= rho;
        double Q_xyz = F[7] - F[8] - F[9] + F[10] - F[11] + F[12] + F[13] - F[14];
        Q_xyz /= rho;
        k[0] = rho;
        k[1] = rho / 6 * (3 * ux); k[2] = rho / 6 * (3 * -ux);
        k[3] = rho / 6 * (3 * uy); k[4] = rho / 6 * (3 * -uy);
        k[5] = rho / 6 * (3 * uz); k[6] = rho / 6 * (3 * -uz);
        for (int q = 7; q < 15; q++) k[q] = 0;
        s[0] = rho * -T;
        s[1] = 1. / 6. * rho * (2 * N_xz - N_yz + T); s[2] = s[1];
        s[3] = 1. / 6. * rho * (-N_xz + 2 * N_yz + T); s[4] = s[3];
        s[5] = 1. / 6. * rho * (-N_xz - N_yz + T); s[6] = s[5];
        s[7] = 1. / 8. * rho * Q_xyz;
        s[8] = -s[7]; s[9] = -s[7]; s[10] = s[7]; s[11] = -s[7]; s[12] = s[7]; s[13] = s[7]; s[14] = -s[7];
        for (int q = 0; q < 15; q++) h[q] = F[q] - k[q] - s[q];
        /* polynomial equilibrium :749-803.  QUIRK: u^2 term uses unscaled u with the scaled cs2. */
        const double scalar_product = ux * ux + uy * uy + uz * uz;
        const double uSquareTerm = -scalar_product / (2 * cs2);
        double weighting = 2. / 9. * rho, mixedTerm;
        feq[0] = weighting * (1 + uSquareTerm);
        weighting = 1. / 9. * rho;
        mixedTerm = prefactor * (vx);
        feq[1] = weighting * (1 + mixedTerm * (1 + 0.5 * mixedTerm) + uSquareTerm);
        feq[2] = weighting * (1 - mixedTerm * (1 - 0.5 * mixedTerm) + uSquareTerm);
        mixedTerm = prefactor * (vy);
        feq[3] = weighting * (1 + mixedTerm * (1 + 0.5 * mixedTerm) + uSquareTerm);
        feq[4] = weighting * (1 - mixedTerm * (1 - 0.5 * mixedTerm) + uSquareTerm);
        mixedTerm = prefactor * (vz);
        feq[5] = weighting * (1 + mixedTerm * (1 + 0.5 * mixedTerm) + uSquareTerm);
        feq[6] = weighting * (1 - mixedTerm * (1 - 0.5 * mixedTerm) + uSquareTerm);
        weighting = 1. / 72. * rho;
        mixedTerm = prefactor * (vx + vy + vz);
        feq[7] = weighting * (1 + mixedTerm * (1 + 0.5 * mixedTerm) + uSquareTerm);
        feq[8] = weighting * (1 - mixedTerm * (1 - 0.5 * mixedTerm) + uSquareTerm);
        mixedTerm = prefactor * (vx + vy - vz);
        feq[9] = weighting * (1 + mixedTerm * (1 + 0.5 * mixedTerm) + uSquareTerm);
        feq[10] = weighting * (1 - mixedTerm * (1 - 0.5 * mixedTerm) + uSquareTerm);
        mixedTerm = prefactor * (vx - vy + vz);
        feq[11] = weighting * (1 + mixedTerm * (1 + 0.5 * mixedTerm) + uSquareTerm);
        feq[12] = weighting * (1 - mixedTerm * (1 - 0.5 * mixedTerm) + uSquareTerm);
        mixedTerm = prefactor * (vx - vy - vz);
        feq[13] = weighting * (1 + mixedTerm * (1 + 0.5 * mixedTerm) + uSquareTerm);
        feq[14] = weighting * (1 - mixedTerm * (1 - 0.5 * mixedTerm) + uSquareTerm);
        /* moments of feq :807-859 */
        T = feq[1] + feq[2] + feq[3] + feq[4] + feq[5] + feq[6] + 3 * feq[7] + 3 * feq[8] + 3 * feq[9] + 3 * feq[10]
            + 3 * feq[11] + 3 * feq[12] + 3 * feq[13] + 3 * feq[14];
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
        for (int q = 0; q < 15; q++) heq[q] = feq[q] - k[q] - seq[q];
        double a[15], b[15];
        for (int q = 0; q < 15; q++) {
            ds[q] = s[q] - seq[q];
            dh[q] = h[q] - heq[q];
            a[q] = ds[q] * dh[q] / feq[q];
            b[q] = dh[q] * dh[q] / feq[q];
        }
        const double sum_s = a[0] + a[1] + a[2] + a[3] + a[4] + a[5] + a[6] + a[7] + a[8] + a[9] + a[10] + a[11] + a[12]
            + a[13] + a[14];
        const double sum_h = b[0] + b[1] + b[2] + b[3] + b[4] + b[5] + b[6] + b[7] + b[8] + b[9] + b[10] + b[11] + b[12]
            + b[13] + b[14];
        const double beta = 1. / (tau + 0.5) / 2;                    /* :989 */
        double gamma = 1. / beta - (2 - 1. / beta) * sum_s / sum_h;  /* :991 (note: (..)*sum_s/sum_h, not *(sum_s/sum_h)) */
        if (sum_h < 1e-16) gamma = 2;
        for (int q = 0; q < 15; q++) f[q * stride + i] = F[q] - beta * (2 * ds[q] + gamma * dh[q]);
    }
    return bad;
}

/* d'Humieres moment matrix as typed in L/collision/MRTEntropic.cpp:172-198 (applied to NATriuM's own D3Q19
 * direction order -- QUIRK kept) and its inverse :200-218. */
static const double orc_tm[19][19] = {
    {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1},
    {-30, -11, -11, -11, -11, -11, -11, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8},
    {12, -4, -4, -4, -4, -4, -4, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1},
    {0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0},
    {0, -4, 4, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0},
    {0, 0, 0, 1, -1, 0, 0, 1, 1, -1, -1, 0, 0, 0, 0, 1, -1, 1, -1},
    {0, 0, 0, -4, 4, 0, 0, 1, 1, -1, -1, 0, 0, 0, 0, 1, -1, 1, -1},
    {0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1},
    {0, 0, 0, 0, 0, -4, 4, 0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1},
    {0, 2, 2, -1, -1, -1, -1, 1, 1, 1, 1, 1, 1, 1, 1, -2, -2, -2, -2},
    {0, -4, -4, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, -2, -2, -2, -2},
    {0, 0, 0, 1, 1, -1, -1, 1, 1, 1, 1, -1, -1, -1, -1, 0, 0, 0, 0},
    {0, 0, 0, -2, -2, 2, 2, 1, 1, 1, 1, -1, -1, -1, -1, 0, 0, 0, 0},
    {0, 0, 0, 0, 0, 0, 0, 1, -1, -1, 1, 0, 0, 0, 0, 0, 0, 0, 0},
    {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, -1, -1, 1},
    {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, -1, -1, 1, 0, 0, 0, 0},
    {0, 0, 0, 0, 0, 0, 0, 1, -1, 1, -1, -1, 1, -1, 1, 0, 0, 0, 0},
    {0, 0, 0, 0, 0, 0, 0, -1, -1, 1, 1, 0, 0, 0, 0, 1, -1, 1, -1},
    {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, -1, -1, -1, -1, 1, 1}};

#define Z 0.
static const double orc_invm[19][19] = {
    {1. / 19., -5. / 399., 1. / 21., Z, Z, Z, Z, Z, Z, Z, Z, Z, Z, Z, Z, Z, Z, Z, Z},
    {1. / 19., -11. / 2394., -1. / 63., 1. / 10., -1. / 10., Z, Z, Z, Z, 1. / 18., -1. / 18., Z, Z, Z, Z, Z, Z, Z, Z},
    {1. / 19., -11. / 2394., -1. / 63., -1. / 10., 1. / 10., Z, Z, Z, Z, 1. / 18., -1. / 18., Z, Z, Z, Z, Z, Z, Z, Z},
    {1. / 19., -11. / 2394., -1. / 63., Z, Z, 1. / 10., -1. / 10., Z, Z, -1. / 36., 1. / 36., 1. / 12., -1. / 12., Z, Z, Z, Z, Z, Z},
    {1. / 19., -11. / 2394., -1. / 63., Z, Z, -1. / 10., 1. / 10., Z, Z, -1. / 36., 1. / 36., 1. / 12., -1. / 12., Z, Z, Z, Z, Z, Z},
    {1. / 19., -11. / 2394., -1. / 63., Z, Z, Z, Z, 1. / 10., -1. / 10., -1. / 36., 1. / 36., -1. / 12., 1. / 12., Z, Z, Z, Z, Z, Z},
    {1. / 19., -11. / 2394., -1. / 63., Z, Z, Z, Z, -1. / 10., 1. / 10., -1. / 36., 1. / 36., -1. / 12., 1. / 12., Z, Z, Z, Z, Z, Z},
    {1. / 19., 4. / 1197., 1. / 252., 1. / 10., 1. / 40., 1. / 10., 1. / 40., Z, Z, 1. / 36., 1. / 72., 1. / 12., 1. / 24., 1. / 4., Z, Z, 1. / 8., -1. / 8., Z},
    {1. / 19., 4. / 1197., 1. / 252., -1. / 10., -1. / 40., 1. / 10., 1. / 40., Z, Z, 1. / 36., 1. / 72., 1. / 12., 1. / 24., -1. / 4., Z, Z, -1. / 8., -1. / 8., Z},
    {1. / 19., 4. / 1197., 1. / 252., 1. / 10., 1. / 40., -1. / 10., -1. / 40., Z, Z, 1. / 36., 1. / 72., 1. / 12., 1. / 24., -1. / 4., Z, Z, 1. / 8., 1. / 8., Z},
    {1. / 19., 4. / 1197., 1. / 252., -1. / 10., -1. / 40., -1. / 10., -1. / 40., Z, Z, 1. / 36., 1. / 72., 1. / 12., 1. / 24., 1. / 4., Z, Z, -1. / 8., 1. / 8., Z},
    {1. / 19., 4. / 1197., 1. / 252., 1. / 10., 1. / 40., Z, Z, 1. / 10., 1. / 40., 1. / 36., 1. / 72., -1. / 12., -1. / 24., Z, Z, 1. / 4., -1. / 8., Z, 1. / 8.},
    {1. / 19., 4. / 1197., 1. / 252., -1. / 10., -1. / 40., Z, Z, 1. / 10., 1. / 40., 1. / 36., 1. / 72., -1. / 12., -1. / 24., Z, Z, -1. / 4., 1. / 8., Z, 1. / 8.},
    {1. / 19., 4. / 1197., 1. / 252., 1. / 10., 1. / 40., Z, Z, -1. / 10., -1. / 40., 1. / 36., 1. / 72., -1. / 12., -1. / 24., Z, Z, -1. / 4., -1. / 8., Z, -1. / 8.},
    {1. / 19., 4. / 1197., 1. / 252., -1. / 10., -1. / 40., Z, Z, -1. / 10., -1. / 40., 1. / 36., 1. / 72., -1. / 12., -1. / 24., Z, Z, 1. / 4., 1. / 8., Z, -1. / 8.},
    {1. / 19., 4. / 1197., 1. / 252., Z, Z, 1. / 10., 1. / 40., 1. / 10., 1. / 40., -1. / 18., -1. / 36., Z, Z, Z, 1. / 4., Z, Z, 1. / 8., -1. / 8.},
    {1. / 19., 4. / 1197., 1. / 252., Z, Z, -1. / 10., -1. / 40., 1. / 10., 1. / 40., -1. / 18., -1. / 36., Z, Z, Z, -1. / 4., Z, Z, -1. / 8., -1. / 8.},
    {1. / 19., 4. / 1197., 1. / 252., Z, Z, 1. / 10., 1. / 40., -1. / 10., -1. / 40., -1. / 18., -1. / 36., Z, Z, Z, -1. / 4., Z, Z, 1. / 8., 1. / 8.},
    {1. / 19., 4. / 1197., 1. / 252., Z, Z, -1. / 10., -1. / 40., -1. / 10., -1. / 40., -1. / 18., -1. / 36., Z, Z, Z, 1. / 4., Z, Z, -1. / 8., 1. / 8.}};
#undef Z

void orc_mrt_entropic_tables(double *tm, double *invm)
{
    for (int p = 0; p < 19; p++)
        for (int q = 0; q < 19; q++) {
            tm[p * 19 + q] = orc_tm[p][q];
            invm[p * 19 + q] = orc_invm[p][q];
        }
}

/* QUIRK (:229-232): the density guard looks at the density *stored from the previous call*, before the new
 * one is computed. */
int orc_collide_mrt_entropic_d3q19(int64_t n, int64_t stride, double *f, double *rho_io, double *u, double scaling,
                                   double tau, int in_init)
{
    int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad) if (n > 50000)
    for (int64_t i = 0; i < n; ++i) {
        double m[19], meq[19];
        if (rho_io[i] < 1e-10) { bad |= 1; continue; }
        for (int p = 0; p < 19; p++) {           /* :234-239 */
            m[p] = 0;
            for (int q = 0; q < 19; q++) m[p] += orc_tm[p][q] * f[q * stride + i];
        }
        const double rho = m[0], jx = m[3], jy = m[5], jz = m[7];
        rho_io[i] = rho;
        if (!in_init) {                          /* :248-253 */
            u[0 * n + i] = scaling / rho_io[i] * jx;
            u[1 * n + i] = scaling / rho_io[i] * jy;
            u[2 * n + i] = scaling / rho_io[i] * jz;
        }
        meq[1] = -11 * rho + 19. / rho * (jx * jx + jy * jy + jz * jz);   /* :255-260 */
        meq[9] = 1. / rho * (2 * jx * jx - (jy * jy + jz * jz));
        meq[11] = 1. / rho * (jy * jy - jz * jz);
        meq[13] = 1. / rho * jx * jy;
        meq[14] = 1. / rho * jy * jz;
        meq[15] = 1. / rho * jx * jz;
        m[1] = m[1] + -1. / (tau + 0.5) * (m[1] - meq[1]);               /* :262-279 */
        m[9] = m[9] + -1. / (tau + 0.5) * (m[9] - meq[9]);
        m[11] = m[11] + -1. / (tau + 0.5) * (m[11] - meq[11]);
        m[13] = m[13] + -1. / (tau + 0.5) * (m[13] - meq[13]);
        m[14] = m[14] + -1. / (tau + 0.5) * (m[14] - meq[14]);
        m[15] = m[15] + -1. / (tau + 0.5) * (m[15] - meq[15]);
        m[2] = -7. / 38 * rho - 11. / 38 * m[1];                         /* :281-289 */
        m[4] = -2. / 3. * jx;
        m[6] = -2. / 3. * jy;
        m[8] = -2. / 3. * jz;
        m[10] = -1. / 2. * m[9];
        m[12] = -1. / 2. * m[11];
        m[16] = 0; m[17] = 0; m[18] = 0;
        for (int p = 0; p < 19; p++) {           /* :291-296 */
            double acc = 0;
            for (int q = 0; q < 19; q++) acc += orc_invm[p][q] * m[q];
            f[p * stride + i] = acc;
        }
    }
    return bad;
}
