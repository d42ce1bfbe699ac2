/*
 * natrium_oracle.c -- CPU restatement of the NATriuM stream + collide hot path.
 *
 * TEST INFRASTRUCTURE ONLY: linked/loaded solely by tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs.  The product (natrium_b200/) never
 * touches it.  Build: see oracle/Makefile  (gcc -O2 -ffp-contract=off, no fast-math, so
 * that every operation is a separately rounded IEEE-754 double operation like the
 * reference's -O3 x86-64 build without FMA contraction).
 *
 * Each function cites the reference lines it follows; L = src/library/natrium.
 * Parity pinning: see oracle/__init__.py (collide pinned by the reference's own KATs,
 * SpMV arithmetic lives in un-vendored Trilinos -> property-pinned only).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MAXQ 48
#define ORC_MAXD 3

/* ------------------------------------------------------------------------------------------
 * stream: y = A x for one block, rows in Epetra local order, entries in stored (sorted) order.
 * Epetra_CrsMatrix::Multiply (Trilinos 13.0.1, not vendored; published algorithm =
 * sequential dot product per row) reached via dealii BlockSparseMatrix::vmult: block row r
 * is  dst_r = M_r0 x_0 ; dst_r += M_rc x_c (c>0), call sites L/solver/CFDSolver.cpp:672,
 * L/advection/SemiLagrangian.h:155.
 * ------------------------------------------------------------------------------------------ */
void orc_spmv_csr(int64_t n_rows, const int64_t *rowptr, const int32_t *col, const double *val,
                  const double *x, double *y, int add)
{
#pragma omp parallel for schedule(static) if (n_rows > 50000)
    for (int64_t i = 0; i < n_rows; ++i) {
        double s = 0.0;
        for (int64_t k = rowptr[i]; k < rowptr[i + 1]; ++k)
            s += val[k] * x[col[k]];
        if (add) y[i] += s; else y[i] = s;
    }
}

/* ------------------------------------------------------------------------------------------
 * helpers restating L/collision_advanced/AuxiliaryCollisionFunctions.h
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    int D, Q;
    double e[ORC_MAXQ][ORC_MAXD];   /* unscaled directions, GeneralCollisionData ctor :161-165 */
    double w[ORC_MAXQ];
    double cs2;                      /* unscaled, :172 */
    double scaling;
    double tau;                      /* :177 with the scaled cs2 */
    double H3[ORC_MAXQ][3][3][3];
    double H4[ORC_MAXQ][3][3][3][3];
} orc_params;

static void orc_calc_H3(orc_params *p)
{   /* calculateH3, AuxiliaryCollisionFunctions.h:519-535 */
    const int D = p->D;
    for (int i = 0; i < p->Q; i++)
        for (int a = 0; a < D; a++)
            for (int b = 0; b < D; b++)
                for (int c = 0; c < D; c++)
                    p->H3[i][a][b][c] = p->e[i][a] * p->e[i][b] * p->e[i][c]
                        - p->cs2 * (p->e[i][a] * (b == c) + p->e[i][b] * (a == c) + p->e[i][c] * (a == b));
}

static void orc_calc_H4(orc_params *p)
{   /* calculateH4, AuxiliaryCollisionFunctions.h:537-566 */
    const int D = p->D;
    const double cs2 = p->cs2;
    for (int i = 0; i < p->Q; i++) {
        const double *e = p->e[i];
        for (int a = 0; a < D; a++)
            for (int b = 0; b < D; b++)
                for (int c = 0; c < D; c++)
                    for (int d = 0; d < D; d++) {
                        const double power4 = e[a] * e[b] * e[c] * e[d];
                        const double power2 = e[a] * e[b] * (c == d) + e[a] * e[c] * (b == d)
                            + e[a] * e[d] * (b == c) + e[b] * e[c] * (a == d)
                            + e[b] * e[d] * (a == c) + e[c] * e[d] * (a == b);
                        const double power0 = (double)((a == b) * (c == d) + (a == c) * (b == d) + (a == d) * (b == c));
                        p->H4[i][a][b][c][d] = (power4 - cs2 * power2 + cs2 * cs2 * power0);
                    }
    }
}

static void orc_make_params(orc_params *p, int D, int Q, const double *e_scaled, const double *w,
                            double scaling, double cs2_scaled, double viscosity, double dt)
{
    memset(p, 0, sizeof(*p));
    p->D = D; p->Q = Q; p->scaling = scaling;
    for (int j = 0; j < Q; j++) {
        for (int i = 0; i < D; i++) p->e[j][i] = e_scaled[j * D + i] / scaling;
        p->w[j] = w[j];
    }
    p->cs2 = cs2_scaled / (scaling * scaling);
    p->tau = viscosity / (dt * cs2_scaled) + 0.5;     /* calculateTauFromNu :63-68 */
    orc_calc_H3(p);
    orc_calc_H4(p);
}

static double orc_density(const double *f, int Q)
{   /* calculateDensity :45-59 (the <1e-10 throw is reported through the return code) */
    double rho = 0.0;
    for (int p = 0; p < Q; ++p) rho += f[p];
    return rho;
}

static void orc_velocity(const orc_params *P, const double *f, double rho, double *u)
{   /* calculateVelocity: D2Q9 :258-268, D3Q19 :270-287, generic :231-242 */
    if (P->D == 2 && P->Q == 9) {
        u[0] = 1.0 / rho * (f[1] + f[5] + f[8] - f[3] - f[6] - f[7]);
        u[1] = 1.0 / rho * (f[2] + f[5] + f[6] - f[4] - f[7] - f[8]);
    } else if (P->D == 3 && P->Q == 19) {
        u[0] = 1.0 / rho * (f[1] - f[3] + f[7] - f[8] - f[9] + f[10] + f[11] + f[12] - f[13] - f[14]);
        u[1] = 1.0 / rho * (-f[5] + f[6] - f[11] + f[12] + f[13] - f[14] - f[15] + f[16] + f[17] - f[18]);
        u[2] = 1.0 / rho * (f[2] - f[4] + f[7] + f[8] - f[9] - f[10] + f[15] + f[16] - f[17] - f[18]);
    } else {
        for (int j = 0; j < P->D; j++) {
            u[j] = 0.0;
            for (int i = 0; i < P->Q; i++) u[j] += P->e[i][j] * f[i];
            u[j] = u[j] * 1.0 / rho;
        }
    }
}

static void orc_feq_bgk(const orc_params *P, double rho, const double *u, double *feq)
{   /* BGKEquilibrium::calc: D2Q9 specialisation Equilibria.h:33-63, generic :66-83 */
    if (P->D == 2 && P->Q == 9) {
        double prefactor = 1. / P->cs2;
        double scalar_product = u[0] * u[0] + u[1] * u[1];
        double uSquareTerm = -scalar_product / (2 * P->cs2);
        double weighting = 4. / 9. * rho;
        double mixedTerm;
        feq[0] = weighting * (1 + uSquareTerm);
        weighting = 1. / 9. * rho;
        mixedTerm = prefactor * (u[0]);
        feq[1] = weighting * (1 + mixedTerm * (1 + 0.5 * mixedTerm) + uSquareTerm);
        feq[3] = weighting * (1 - mixedTerm * (1 - 0.5 * mixedTerm) + uSquareTerm);
        mixedTerm = prefactor * (u[1]);
        feq[2] = weighting * (1 + mixedTerm * (1 + 0.5 * mixedTerm) + uSquareTerm);
        feq[4] = weighting * (1 - mixedTerm * (1 - 0.5 * mixedTerm) + uSquareTerm);
        weighting = 1. / 36. * rho;
        mixedTerm = prefactor * (u[0] + u[1]);
        feq[5] = weighting * (1 + mixedTerm * (1 + 0.5 * mixedTerm) + uSquareTerm);
        feq[7] = weighting * (1 - mixedTerm * (1 - 0.5 * mixedTerm) + uSquareTerm);
        mixedTerm = prefactor * (-u[0] + u[1]);
        feq[6] = weighting * (1 + mixedTerm * (1 + 0.5 * mixedTerm) + uSquareTerm);
        feq[8] = weighting * (1 - mixedTerm * (1 - 0.5 * mixedTerm) + uSquareTerm);
        return;
    }
    double uu_term = 0.0;
    for (int j = 0; j < P->D; j++) uu_term += -(u[j] * u[j]) / (2.0 * P->cs2);
    for (int i = 0; i < P->Q; i++) {
        double ue_term = 0.0;
        for (int j = 0; j < P->D; j++) ue_term += (u[j] * P->e[i][j]) / P->cs2;
        feq[i] = P->w[i] * rho * (1 + ue_term * (1 + 0.5 * (ue_term)) + uu_term);
    }
}

static void orc_feq_quartic(const orc_params *P, double rho, const double *uin, double T, double *feq)
{   /* QuarticEquilibrium::polynomial, Equilibria.h:119-267 */
    const int D = P->D, Q = P->Q;
    const double cs2 = P->cs2;
    double u[3] = {0, 0, 0};
    for (int j = 0; j < D; j++) u[j] = uin[j];
    double uu_term = 0.0;
    for (int j = 0; j < D; j++) uu_term += -(u[j] * u[j]) / (2.0 * cs2);
    const double T1 = cs2 * (T - 1);
    const double a_xxx = u[0] * u[0] * u[0] + T1 * (u[0] + u[0] + u[0]);
    const double a_xxy = u[0] * u[0] * u[1] + T1 * (u[1]);
    const double a_xyy = u[0] * u[1] * u[1] + T1 * (u[0]);
    const double a_yyy = u[1] * u[1] * u[1] + T1 * (u[1] + u[1] + u[1]);
    const double a_xxxx = u[0] * u[0] * u[0] * u[0] + T1 * u[0] * u[0] * 6.0 + T1 * T1 * 3.0;
    const double a_yyyy = u[1] * u[1] * u[1] * u[1] + T1 * u[1] * u[1] * 6.0 + T1 * T1 * 3.0;
    const double a_xxxy = u[0] * u[0] * u[0] * u[1] + T1 * (u[0] * u[1] * 3.0);
    const double a_xyyy = u[0] * u[1] * u[1] * u[1] + T1 * (u[0] * u[1] * 3.0);
    const double a_xxyy = u[0] * u[0] * u[1] * u[1] + T1 * (u[0] * u[0] + u[1] * u[1]) + T1 * T1;
    double a_zzz = 0.0, a_xxz = 0.0, a_xzz = 0.0, a_yzz = 0.0, a_yyz = 0.0, a_xyz = 0.0;
    double a_zzzz = 0.0, a_xzzz = 0.0, a_xxzz = 0.0, a_xxxz = 0.0, a_yzzz = 0.0, a_yyzz = 0.0,
           a_yyyz = 0.0, a_xxyz = 0.0, a_xyyz = 0.0, a_xyzz = 0.0;
    if (D == 3) {
        a_zzz = u[2] * u[2] * u[2] + T1 * (u[2] + u[2] + u[2]);
        a_xxz = u[0] * u[0] * u[2] + T1 * (u[2]);
        a_xzz = u[0] * u[2] * u[2] + T1 * (u[0]);
        a_yzz = u[1] * u[2] * u[2] + T1 * (u[1]);
        a_yyz = u[1] * u[1] * u[2] + T1 * (u[2]);
        a_xyz = u[0] * u[1] * u[2];
        a_zzzz = u[2] * u[2] * u[2] * u[2] + T1 * u[2] * u[2] * 6.0 + T1 * T1 * 3.0;
        a_xxxz = u[0] * u[0] * u[0] * u[2] + T1 * (u[0] * u[2] * 3.0);
        a_yyyz = u[1] * u[1] * u[1] * u[2] + T1 * (u[1] * u[2] * 3.0);
        a_xzzz = u[0] * u[2] * u[2] * u[2] + T1 * (u[0] * u[2] * 3.0);
        a_yzzz = u[1] * u[2] * u[2] * u[2] + T1 * (u[1] * u[2] * 3.0);
        a_xxzz = u[0] * u[0] * u[2] * u[2] + T1 * (u[0] * u[0] + u[2] * u[2]) + T1 * T1;
        a_yyzz = u[1] * u[1] * u[2] * u[2] + T1 * (u[1] * u[1] + u[2] * u[2]) + T1 * T1;
        a_xyzz = u[0] * u[1] * u[2] * u[2] + T1 * (u[0] * u[1]);
        a_xyyz = u[0] * u[1] * u[1] * u[2] + T1 * (u[0] * u[2]);
        a_xxyz = u[0] * u[0] * u[1] * u[2] + T1 * (u[1] * u[2]);
    }
    for (int i = 0; i < Q; i++) {
        const double *e = P->e[i];
        const double w = P->w[i];
        double ue_term = 0.0;
        for (int j = 0; j < D; j++) ue_term += (u[j] * e[j]) / cs2;
        feq[i] = w * rho * (1 + ue_term * (1 + 0.5 * (ue_term)) + uu_term);
        for (int alp = 0; alp < D; alp++)
            for (int bet = 0; bet < D; bet++) {
                const double eye = (alp == bet) ? 1.0 : 0.0;
                feq[i] += rho * w / (2.0 * cs2) * ((T - 1) * eye * e[alp] * e[bet] - cs2 * eye * (T - 1));
            }
        const double H_xxx = P->H3[i][0][0][0], H_xxy = P->H3[i][0][0][1];
        const double H_xyy = P->H3[i][0][1][1], H_yyy = P->H3[i][1][1][1];
        feq[i] += w * rho / (6. * cs2 * cs2 * cs2)
            * (a_xxx * H_xxx + 3 * (a_xxy * H_xxy + a_xyy * H_xyy) + a_yyy * H_yyy);
        if (D == 3) {
            const double H_zzz = P->H3[i][2][2][2], H_xxz = P->H3[i][0][0][2], H_xzz = P->H3[i][0][2][2];
            const double H_yzz = P->H3[i][1][2][2], H_yyz = P->H3[i][1][1][2], H_xyz = P->H3[i][0][1][2];
            feq[i] += w * rho / (6. * cs2 * cs2 * cs2)
                * (a_zzz * H_zzz + 3 * (a_xxz * H_xxz + a_xzz * H_xzz + a_yzz * H_yzz + a_yyz * H_yyz)
                   + 6.0 * a_xyz * H_xyz);
        }
        const double H_xxxx = P->H4[i][0][0][0][0], H_yyyy = P->H4[i][1][1][1][1];
        const double H_xxxy = P->H4[i][0][0][0][1], H_xyyy = P->H4[i][0][1][1][1];
        const double H_xxyy = P->H4[i][0][0][1][1];
        feq[i] += w * rho / (24. * cs2 * cs2 * cs2 * cs2)
            * (H_xxxx * a_xxxx + H_yyyy * a_yyyy + 6.0 * H_xxyy * a_xxyy + 4.0 * H_xyyy * a_xyyy
               + 4.0 * H_xxxy * a_xxxy);
        if (D == 3) {
            const double H_zzzz = P->H4[i][2][2][2][2], H_xzzz = P->H4[i][0][2][2][2];
            const double H_xxzz = P->H4[i][0][0][2][2], H_xxxz = P->H4[i][0][0][0][2];
            const double H_yzzz = P->H4[i][1][2][2][2], H_yyzz = P->H4[i][1][1][2][2];
            const double H_yyyz = P->H4[i][1][1][1][2], H_xxyz = P->H4[i][0][0][1][2];
            const double H_xyyz = P->H4[i][0][1][1][2], H_xyzz = P->H4[i][0][1][2][2];
            feq[i] += w * rho / (24. * cs2 * cs2 * cs2 * cs2)
                * (H_zzzz * a_zzzz
                   + 4.0 * (H_xzzz * a_xzzz + H_yzzz * a_yzzz + H_xxxz * a_xxxz + H_yyyz * a_yyyz)
                   + 6.0 * (H_xxzz * a_xxzz + H_yyzz * a_yyzz)
                   + 12.0 * (H_xxyz * a_xxyz + H_xyyz * a_xyyz + H_xyzz * a_xyzz));
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * collideAll, f only.  L/collision_advanced/CollisionOperator.h:26-111 with
 * BGKCollision::relax (CollisionSchemes.h:28-41).  Populations are Q arrays f[q*stride + i].
 * equilibrium: 0 = BGKEquilibrium, 1 = QuarticEquilibrium (temperature fixed to 1, :66).
 * Returns 0, or -1 if a density < 1e-10 was met (CollisionException, Aux...h:53-56).
 * No external force (problemDescription.hasExternalForce() == false).
 * ------------------------------------------------------------------------------------------ */
int orc_collide_bgk(int D, int Q, int64_t n, int64_t stride, double *f, double *rho_out, double *u_out,
                    const double *e_scaled, const double *w, double scaling, double cs2_scaled,
                    double viscosity, double dt, int equilibrium, int in_init)
{
    orc_params P;
    orc_make_params(&P, D, Q, e_scaled, w, scaling, cs2_scaled, viscosity, dt);
    int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad) if (n > 50000)
    for (int64_t ii = 0; ii < n; ii++) {
        double fl[ORC_MAXQ], feq[ORC_MAXQ], u[3];
        for (int p = 0; p < Q; ++p) fl[p] = f[p * stride + ii];
        double rho = orc_density(fl, Q);
        if (rho < 1e-10) bad |= 1;
        rho_out[ii] = rho;
        orc_velocity(&P, fl, rho, u);
        if (!in_init) {
            for (int j = 0; j < D; ++j) u_out[j * n + ii] = u[j] * P.scaling;
        } else {
            for (int j = 0; j < D; ++j) u[j] = u_out[j * n + ii] / P.scaling;
        }
        if (equilibrium == 0) orc_feq_bgk(&P, rho, u, feq);
        else orc_feq_quartic(&P, rho, u, 1.0, feq);
        for (int p = 0; p < Q; ++p) fl[p] -= 1. / P.tau * (fl[p] - feq[p]);
        for (int p = 0; p < Q; ++p) f[p * stride + ii] = fl[p];
    }
    return bad ? -1 : 0;
}

/* ------------------------------------------------------------------------------------------
 * collideAll, f + g.  CollisionOperator.h:113-224 with BGKCollision::relaxWithG
 * (CollisionSchemes.h:43-118), calculateTemperature (Aux...h:289-307), calculateGeqFromFeq
 * (:420-431), centred moments (:445-472), fStar/gStar (:484-515), Knudsen estimate (:474-481).
 * equilibrium: 0 = BGKEquilibrium (ignores T), 1 = QuarticEquilibrium.
 * ------------------------------------------------------------------------------------------ */
static int orc_collide_fg_core(int D, int Q, int64_t n, int64_t stride, double *f, double *g,
                       double *rho_out, double *u_out, double *T_out, double *mss_out,
                       const double *e_scaled, const double *w, double scaling, double cs2_scaled,
                       double viscosity, double dt, int equilibrium, double gamma,
                       int prandtl_set, double prandtl, int sutherland_set, int in_init,
                       int has_force, int force_type, const double *force)
{
    orc_params P;
    orc_make_params(&P, D, Q, e_scaled, w, scaling, cs2_scaled, viscosity, dt);
    if (has_force && force_type == 0) return -2;               /* Aux...h:335-339 */
    if (has_force && force_type != 1 && force_type != 2) return -3;
    int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad) if (n > 50000)
    for (int64_t ii = 0; ii < n; ii++) {
        double fl[ORC_MAXQ], gl[ORC_MAXQ], feq[ORC_MAXQ], geq[ORC_MAXQ];
        double fNeq[ORC_MAXQ], gNeq[ORC_MAXQ], fStar[ORC_MAXQ], gStar[ORC_MAXQ], u[3] = {0, 0, 0};
        for (int p = 0; p < Q; ++p) fl[p] = f[p * stride + ii];
        for (int p = 0; p < Q; ++p) gl[p] = g[p * stride + ii];
        double rho = orc_density(fl, Q);
        if (rho < 1e-10) bad |= 1;
        rho_out[ii] = rho;
        orc_velocity(&P, fl, rho, u);
        /* calculateTemperature */
        double T = 0.0;
        for (int i = 0; i < Q; i++) {
            double sum = 0.0;
            for (int a = 0; a < D; a++) sum += (P.e[i][a] - u[a]) * (P.e[i][a] - u[a]);
            T += sum * fl[i] / P.cs2 + gl[i];
        }
        const double C_v = 1. / (gamma - 1.0);
        T = T * 0.5 / (rho * C_v);
        T_out[ii] = T;
        if (!in_init) {
            for (int j = 0; j < D; ++j) u_out[j * n + ii] = u[j] * P.scaling;
            if (has_force && force_type == 1) {     /* applyMacroscopicForces + applyForces, SHIFTING_VELOCITY */
                for (int j = 0; j < D; j++) u_out[j * n + ii] = u_out[j * n + ii] + 0.5 * dt * force[j] / rho;
                for (int j = 0; j < D; j++) u[j] += P.tau * dt * force[j] / rho / P.scaling;
            }
        } else {
            for (int j = 0; j < D; ++j) u[j] = u_out[j * n + ii] / P.scaling;
        }
        /* relaxWithG */
        if (equilibrium == 0) orc_feq_bgk(&P, rho, u, feq);
        else orc_feq_quartic(&P, rho, u, T, feq);
        for (int i = 0; i < Q; i++) geq[i] = feq[i] * (T) * (2.0 * C_v - D);
        for (int p = 0; p < Q; ++p) {
            fStar[p] = 0.0; gStar[p] = 0.0;
            fNeq[p] = fl[p] - feq[p];
            gNeq[p] = gl[p] - geq[p];
        }
        if (prandtl_set) {
            double Qn[3][3][3];
            double qg[3] = {0, 0, 0};
            memset(Qn, 0, sizeof(Qn));
            for (int i = 0; i < Q; i++)
                for (int a = 0; a < D; a++)
                    for (int b = 0; b < D; b++)
                        for (int c = 0; c < D; c++)
                            Qn[a][b][c] += ((P.e[i][a] - u[a]) * (P.e[i][b] - u[b]) * (P.e[i][c] - u[c])) * fNeq[i];
            for (int i = 0; i < Q; i++)
                for (int a = 0; a < D; a++) qg[a] += (P.e[i][a] - u[a]) * gNeq[i];
            const double cs6 = 6.0 * P.cs2 * P.cs2 * P.cs2;
            for (int a = 0; a < D; a++)
                for (int b = 0; b < D; b++)
                    for (int c = 0; c < D; c++)
                        for (int i = 0; i < Q; i++)
                            fStar[i] += P.w[i] * (Qn[a][b][c] * (P.e[i][a] * P.e[i][b] * P.e[i][c]
                                                                 - 3 * P.cs2 * P.e[i][c] * (a == b))) / cs6;
            for (int a = 0; a < D; a++)
                for (int i = 0; i < Q; i++) gStar[i] += P.w[i] * (qg[a] * P.e[i][a]) / T;
        }
        double sutherland_factor = 1.0;
        if (sutherland_set) sutherland_factor = pow(T / 0.85, 0.7);
        const double visc_tau = (P.tau - 0.5) * sutherland_factor / (T * rho) + 0.5;
        double knudsen = 0.0;
        for (int i = 0; i < Q; i++) knudsen += fabs(fl[i] - feq[i]) / P.w[i];
        knudsen = knudsen / Q;
        mss_out[ii] = knudsen;
        const double prandtl_tau = (visc_tau - 0.5) / prandtl + 0.5;
        const double visc_omega = 1. / visc_tau;
        const double prandtl_omega = 1. / prandtl_tau;
        const double prandtl_diff = visc_omega - prandtl_omega;
        for (int p = 0; p < Q; ++p) {
            fl[p] -= visc_omega * fNeq[p] - prandtl_diff * fStar[p];
            gl[p] -= visc_omega * gNeq[p] - prandtl_diff * gStar[p];
        }
        if (has_force && force_type == 2) {     /* postCollisionApplyForces, EXACT_DIFFERENCE: f only (Aux...h:399-411) */
            double shifted[ORC_MAXQ];
            for (int j = 0; j < D; j++) {
                u[j] += dt * force[j] / rho / P.scaling;
                u_out[j * n + ii] = u_out[j * n + ii] + 0.5 * dt * force[j] / rho;
            }
            if (equilibrium == 0) orc_feq_bgk(&P, rho, u, shifted);
            else orc_feq_quartic(&P, rho, u, T, shifted);
            for (int i = 0; i < Q; i++) fl[i] += (shifted[i] - feq[i]);
        }
        for (int p = 0; p < Q; ++p) {
            f[p * stride + ii] = fl[p];
            g[p * stride + ii] = gl[p];
        }
    }
    return bad ? -1 : 0;
}

int orc_collide_bgk_fg(int D, int Q, int64_t n, int64_t stride, double *f, double *g,
                       double *rho_out, double *u_out, double *T_out, double *mss_out,
                       const double *e_scaled, const double *w, double scaling, double cs2_scaled,
                       double viscosity, double dt, int equilibrium, double gamma,
                       int prandtl_set, double prandtl, int sutherland_set, int in_init)
{
    return orc_collide_fg_core(D, Q, n, stride, f, g, rho_out, u_out, T_out, mss_out, e_scaled, w, scaling, cs2_scaled,
                               viscosity, dt, equilibrium, gamma, prandtl_set, prandtl, sutherland_set, in_init, 0, 0, NULL);
}

/* f + g with an external force (C5: EXACT_DIFFERENCE forcing of the channel flow, step-turbulent-channel.cpp) */
int orc_collide_bgk_fg_forced(int D, int Q, int64_t n, int64_t stride, double *f, double *g,
                              double *rho_out, double *u_out, double *T_out, double *mss_out,
                              const double *e_scaled, const double *w, double scaling, double cs2_scaled,
                              double viscosity, double dt, int equilibrium, double gamma,
                              int prandtl_set, double prandtl, int sutherland_set, int in_init,
                              int force_type, const double *force)
{
    return orc_collide_fg_core(D, Q, n, stride, f, g, rho_out, u_out, T_out, mss_out, e_scaled, w, scaling, cs2_scaled,
                               viscosity, dt, equilibrium, gamma, prandtl_set, prandtl, sutherland_set, in_init, 1, force_type, force);
}

/* equilibrium evaluation entry points (used by tests and by initial conditions) */
void orc_equilibrium(int D, int Q, const double *e_scaled, const double *w, double scaling, double cs2_scaled,
                     int equilibrium, double rho, const double *u_unscaled, double T, double *feq)
{
    orc_params P;
    orc_make_params(&P, D, Q, e_scaled, w, scaling, cs2_scaled, 1.0, 1.0);
    if (equilibrium == 0) orc_feq_bgk(&P, rho, u_unscaled, feq);
    else orc_feq_quartic(&P, rho, u_unscaled, T, feq);
}

/* ------------------------------------------------------------------------------------------
 * Legacy BGKStandard (the unit-tested path): getEquilibriumDistribution
 * L/collision/BGKStandard.cpp:23-41 (scaled directions, scaled cs2, physical u) and
 * BGK::collideSinglePoint L/collision/BGK.cpp:24-45 with prefactor -1/(tau_legacy+0.5),
 * tau_legacy = nu/(dt cs2) (CollisionModel.h:152-157).
 * ------------------------------------------------------------------------------------------ */
void orc_legacy_feq(int D, int Q, const double *e_scaled, const double *w, double cs2_scaled,
                    double rho, const double *u, double *feq)
{
    double uu = 0.0;
    for (int j = 0; j < D; j++) uu += u[j] * u[j];
    const double uSquareTerm = -uu / (2 * cs2_scaled);
    for (int i = 0; i < Q; i++) {
        const double prefactor = w[i] * rho;
        if (i == 0) { feq[i] = prefactor * (1 + uSquareTerm); continue; }
        double ue = 0.0;
        for (int j = 0; j < D; j++) ue += u[j] * e_scaled[i * D + j];
        const double mixedTerm = ue / cs2_scaled;
        feq[i] = prefactor * (1 + mixedTerm * (1 + 0.5 * (mixedTerm)) + uSquareTerm);
    }
}

void orc_legacy_collide_single_point(int D, int Q, const double *e_scaled, const double *w,
                                     double cs2_scaled, double tau_legacy, double *f)
{
    double rho = 0.0, u[3] = {0, 0, 0}, feq[ORC_MAXQ];
    for (int i = 0; i < Q; i++) rho += f[i];
    for (int i = 0; i < Q; i++)
        for (int j = 0; j < D; j++) u[j] += f[i] * e_scaled[i * D + j];
    for (int j = 0; j < D; j++) u[j] *= 1. / rho;
    orc_legacy_feq(D, Q, e_scaled, w, cs2_scaled, rho, u, feq);
    const double prefactor = -1. / (tau_legacy + 0.5);
    for (int i = 0; i < Q; i++) f[i] = f[i] + prefactor * (f[i] - feq[i]);
}

/* ------------------------------------------------------------------------------------------
 * One reference-ordered time step for the CPU baseline (CFDSolver::stream + collide,
 * L/solver/CFDSolver.cpp:659-754, 807-843): full copy f_tmp = f (:671), one CSR SpMV per
 * non-empty block (:672), then the separate pointwise collide pass.  Blocks are passed as
 * arrays of CSR triplets; populations are [Q][stride].  Periodic problems: no wall hits.
 * ------------------------------------------------------------------------------------------ */
int orc_step_f(int D, int Q, int64_t n, int64_t stride, double *f, double *f_tmp,
               int n_blocks, const int32_t *block_row, const int32_t *block_col,
               const int64_t *const *rowptr, const int32_t *const *col, const double *const *val,
               double *rho_out, double *u_out,
               const double *e_scaled, const double *w, double scaling, double cs2_scaled,
               double viscosity, double dt, int equilibrium)
{
    memcpy(f_tmp, f, (size_t)Q * stride * sizeof(double));                /* DistributionFunctions f_tmp(m_f) */
    char seen[ORC_MAXQ];
    memset(seen, 0, sizeof(seen));
    for (int b = 0; b < n_blocks; b++) {
        const int r = block_row[b], c = block_col[b];
        /* vmult(dst = m_f.FStream, src = f_tmp.FStream): reads the copy, writes f */
        orc_spmv_csr(n, rowptr[b], col[b], val[b], f_tmp + (int64_t)(c + 1) * stride,
                     f + (int64_t)(r + 1) * stride, seen[r]);
        seen[r] = 1;
    }
    return orc_collide_bgk(D, Q, n, stride, f, rho_out, u_out, e_scaled, w, scaling, cs2_scaled,
                           viscosity, dt, equilibrium, 0);
}

int orc_step_fg(int D, int Q, int64_t n, int64_t stride, double *f, double *g, double *tmp,
                int n_blocks, const int32_t *block_row, const int32_t *block_col,
                const int64_t *const *rowptr, const int32_t *const *col, const double *const *val,
                double *rho_out, double *u_out, double *T_out, double *mss_out,
                const double *e_scaled, const double *w, double scaling, double cs2_scaled,
                double viscosity, double dt, int equilibrium, double gamma,
                int prandtl_set, double prandtl, int sutherland_set)
{
    /* stream() then gStream() with the same matrix (CompressibleCFDSolver.h:181-314) */
    for (int which = 0; which < 2; which++) {
        double *x = which ? g : f;
        memcpy(tmp, x, (size_t)Q * stride * sizeof(double));
        char seen[ORC_MAXQ];
        memset(seen, 0, sizeof(seen));
        for (int b = 0; b < n_blocks; b++) {
            const int r = block_row[b], c = block_col[b];
            orc_spmv_csr(n, rowptr[b], col[b], val[b], tmp + (int64_t)(c + 1) * stride,
                         x + (int64_t)(r + 1) * stride, seen[r]);
            seen[r] = 1;
        }
    }
    return orc_collide_bgk_fg(D, Q, n, stride, f, g, rho_out, u_out, T_out, mss_out, e_scaled, w, scaling,
                              cs2_scaled, viscosity, dt, equilibrium, gamma, prandtl_set, prandtl,
                              sutherland_set, 0);
}

/* ------------------------------------------------------------------------------------------
 * collideAll, f only, remaining rows of selectCollision (CollisionSelection.h:85-88,179-196):
 *   scheme 0  BGKCollision::relax            CollisionSchemes.h:28-41
 *   scheme 1  Regularized::relax             CollisionSchemes.h:122-203 (Q tensor :136-148)
 *   scheme 2  MultipleRelaxationTime::relax  CollisionSchemes.h:206-265; M, T, omega are what
 *             AuxiliaryMRTFunctions::make_M / make_T / make_diag return (AuxiliaryMRTFunctions.cpp)
 * and the external-force hooks of collideAll (CollisionOperator.h:79-96):
 *   applyMacroscopicForces Aux...h:332-360, applyForces :362-386, postCollisionApplyForces :388-417.
 * force_type follows ForceType (ConfigNames.h:114-119): 0 NO_FORCING, 1 SHIFTING_VELOCITY,
 * 2 EXACT_DIFFERENCE, 3 GUO.  has_force mirrors problemDescription.hasExternalForce().
 * Returns 0, -1 (density < 1e-10), -2 (NATriuMException: forcing switched off), -3 (NotImplemented).
 * ------------------------------------------------------------------------------------------ */
int orc_collide_advanced_f(int D, int Q, int64_t n, int64_t stride, double *f, double *rho_out, double *u_out,
                           const double *e_scaled, const double *w, double scaling, double cs2_scaled,
                           double viscosity, double dt, int equilibrium, int scheme, int in_init,
                           int has_force, int force_type, const double *force,
                           const double *M, const double *T, const double *omega)
{
    orc_params P;
    orc_make_params(&P, D, Q, e_scaled, w, scaling, cs2_scaled, viscosity, dt);
    if (has_force && force_type == 0) return -2;
    if (has_force && force_type != 1 && force_type != 2) return -3;
    double Qt[ORC_MAXQ][3][3];                       /* Regularized::SpecificCollisionData::initializeQ */
    for (int a = 0; a < Q; a++)
        for (int b = 0; b < D; b++)
            for (int c = 0; c < D; c++) {
                Qt[a][b][c] = P.e[a][b] * P.e[a][c];
                if (b == c) Qt[a][b][c] -= P.cs2;
            }
    int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad) if (n > 50000)
    for (int64_t ii = 0; ii < n; ii++) {
        double fl[ORC_MAXQ], feq[ORC_MAXQ], u[3] = {0, 0, 0};
        for (int p = 0; p < Q; ++p) fl[p] = f[p * stride + ii];
        double rho = orc_density(fl, Q);
        if (rho < 1e-10) bad |= 1;
        rho_out[ii] = rho;
        orc_velocity(&P, fl, rho, u);
        if (!in_init) {
            for (int j = 0; j < D; ++j) u_out[j * n + ii] = u[j] * P.scaling;
            if (has_force) {
                if (force_type == 1) {      /* SHIFTING_VELOCITY */
                    for (int j = 0; j < D; j++) u_out[j * n + ii] = u_out[j * n + ii] + 0.5 * dt * force[j] / rho;
                    for (int j = 0; j < D; j++) u[j] += P.tau * dt * force[j] / rho / P.scaling;
                }
            }
        } else {
            for (int j = 0; j < D; ++j) u[j] = u_out[j * n + ii] / P.scaling;
        }
        if (equilibrium == 0) orc_feq_bgk(&P, rho, u, feq);
        else orc_feq_quartic(&P, rho, u, 1.0, feq);
        if (scheme == 0) {
            for (int p = 0; p < Q; ++p) fl[p] -= 1. / P.tau * (fl[p] - feq[p]);
        } else if (scheme == 1) {
            double pi[3][3], pieq[3][3], fi1[ORC_MAXQ];
            for (int m = 0; m < D; m++) for (int k = 0; k < D; k++) { pi[m][k] = 0.0; pieq[m][k] = 0.0; }
            for (int j = 0; j < Q; j++)
                for (int m = 0; m < D; m++)
                    for (int k = 0; k < D; k++) {
                        pi[m][k] += fl[j] * P.e[j][m] * P.e[j][k];
                        pieq[m][k] += feq[j] * P.e[j][m] * P.e[j][k];
                    }
            for (int m = 0; m < D; m++) for (int k = 0; k < D; k++) pi[m][k] -= pieq[m][k];
            for (int a = 0; a < Q; a++) {
                fi1[a] = 0.0;
                for (int b = 0; b < D; b++)
                    for (int c = 0; c < D; c++)
                        fi1[a] += P.w[a] / (2 * P.cs2 * P.cs2) * Qt[a][b][c] * pi[b][c];
            }
            for (int i = 0; i < Q; ++i) fl[i] = feq[i] + (1. - 1. / P.tau) * fi1[i];
        } else {
            double m[ORC_MAXQ], meq[ORC_MAXQ];
            for (int i = 0; i < Q; i++) {
                m[i] = 0.0; meq[i] = 0.0;
                for (int j = 0; j < Q; j++) m[i] += M[i * Q + j] * fl[j];
                for (int j = 0; j < Q; j++) meq[i] += M[i * Q + j] * feq[j];
            }
            for (int i = 0; i < Q; i++) m[i] = m[i] - omega[i] * (m[i] - meq[i]);
            for (int i = 0; i < Q; i++) {
                fl[i] = 0.0;
                for (int j = 0; j < Q; j++) fl[i] += T[i * Q + j] * m[j];
            }
        }
        if (has_force && force_type == 2) {     /* EXACT_DIFFERENCE, postCollisionApplyForces */
            double shifted[ORC_MAXQ];
            for (int j = 0; j < D; j++) {
                u[j] += dt * force[j] / rho / P.scaling;
                u_out[j * n + ii] = u_out[j * n + ii] + 0.5 * dt * force[j] / rho;
            }
            if (equilibrium == 0) orc_feq_bgk(&P, rho, u, shifted);
            else orc_feq_quartic(&P, rho, u, 1.0, shifted);
            for (int i = 0; i < Q; i++) fl[i] += (shifted[i] - feq[i]);
        }
        for (int p = 0; p < Q; ++p) f[p * stride + ii] = fl[p];
    }
    return bad ? -1 : 0;
}

/* ------------------------------------------------------------------------------------------
 * Wall hits after streaming: SemiLagrangianBoundaryHandler::apply / operate
 * (L/boundaries/SemiLagrangianBoundaryHandler.cpp:26-109) over a flattened hit list in HitList
 * iteration order.  kind 0: VelocityNeqBounceBack::calculateBoundaryValues
 * (L/boundaries/VelocityNeqBounceBack.cpp:137-195): f_new[dir](idx) += 2 w_dir rho (e_dir . u_wall) / cs2 with
 * rho = 1 -- the term is evaluated by the host (it owns the wall-velocity function) and arrives in value[].
 * kind 1: ThermalBounceBack::calculateBoundaryValues (L/boundaries/ThermalBounceBack.cpp:50-109): D3Q45 only,
 * gamma fixed to 1.4 there; re-equilibrates f and g of the destination DoF to the wall temperature value[].
 * f is the post-stream f, g the not-yet-streamed g (CompressibleCFDSolver.h:194-195: the BC runs before gStream).
 * ------------------------------------------------------------------------------------------ */
int orc_apply_wall_hits(int D, int Q, int64_t stride, double *f, double *g, int64_t n_hits,
                        const int32_t *dest_index, const int32_t *dest_dir, const int32_t *kind, const double *value,
                        const double *e_scaled, const double *w, double scaling, double cs2_scaled)
{
    orc_params P;
    orc_make_params(&P, D, Q, e_scaled, w, scaling, cs2_scaled, 1.0, 1.0);
    for (int64_t h = 0; h < n_hits; h++) {
        const int64_t idx = dest_index[h];
        if (kind[h] == 0) {
            f[(int64_t)dest_dir[h] * stride + idx] = f[(int64_t)dest_dir[h] * stride + idx] + value[h];
        } else if (kind[h] == 1) {
            if (Q != 45 || !g) return -3;
            const double gamma = 1.4, Tw = value[h];
            double fd[ORC_MAXQ], gd[ORC_MAXQ], feq[ORC_MAXQ], geq[ORC_MAXQ], u[3] = {0, 0, 0};
            for (int i = 0; i < Q; i++) { fd[i] = f[(int64_t)i * stride + idx]; gd[i] = g[(int64_t)i * stride + idx]; }
            const double rho = orc_density(fd, Q);
            for (int j = 0; j < D; j++) {          /* calculateVelocity(f, u, rho, e), Aux...h:244-255 */
                u[j] = 0.0;
                for (int i = 0; i < Q; i++) u[j] += P.e[i][j] * fd[i];
                u[j] = u[j] * 1.0 / rho;
            }
            double T = 0.0;                        /* calculateTemperature(f, g, u, rho, e, cs2, gamma), :309-325 */
            for (int i = 0; i < Q; i++) {
                double sum = 0.0;
                for (int a = 0; a < D; a++) sum += (P.e[i][a] - u[a]) * (P.e[i][a] - u[a]);
                T += sum * fd[i] / P.cs2 + gd[i];
            }
            const double C_v = 1. / (gamma - 1.0);
            T = T * 0.5 / (rho * C_v);
            if (fabs(T - Tw) > 0.00001) {
                orc_feq_quartic(&P, rho, u, T, feq);
                for (int i = 0; i < Q; i++) geq[i] = feq[i] * (T) * (2.0 * C_v - D);
                for (int i = 0; i < Q; i++) { fd[i] -= feq[i]; gd[i] -= geq[i]; }
                orc_feq_quartic(&P, rho, u, Tw, feq);
                for (int i = 0; i < Q; i++) geq[i] = feq[i] * (Tw) * (2.0 * C_v - D);
                for (int i = 0; i < Q; i++) {
                    f[(int64_t)i * stride + idx] = fd[i] + feq[i];
                    g[(int64_t)i * stride + idx] = geq[i];
                }
            }
        } else {
            return -3;
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * PseudoEntropicStabilizer::apply (L/dataprocessors/PseudoEntropicStabilizer.cpp:152-262): per owned DoF
 * f_new[i] = sum_j A[i][j] f[j], j ascending from 0; A = n/d (D2Q9, :27-40), nd_d2q9_with_e (:42-57) or nd_d3q19
 * (:59-150) -- passed in by the caller (tests use the literals extracted into tests/golden/stabilizer_tables.npz).
 * ------------------------------------------------------------------------------------------ */
void orc_apply_stabilizer(int Q, int64_t n, int64_t stride, double *f, const double *A)
{
#pragma omp parallel for schedule(static) if (n > 50000)
    for (int64_t ii = 0; ii < n; ii++) {
        double fi[ORC_MAXQ], fn[ORC_MAXQ];
        for (int j = 0; j < Q; j++) fi[j] = f[j * stride + ii];
        for (int i = 0; i < Q; i++) {
            fn[i] = 0;
            for (int j = 0; j < Q; j++) fn[i] += A[i * Q + j] * fi[j];
        }
        for (int j = 0; j < Q; j++) f[j * stride + ii] = fn[j];
    }
}

/* ExponentialFilter<dim>::applyFilter, smoothing/ExponentialFilter.cpp:139-199 (called once per population by
 * CFDSolver::filter, solver/CFDSolver.cpp:859-874).  Cells are visited one after the other in the given order
 * (cell->get_dof_indices of the active-cell loop); each reads its DoFs from the vector that earlier cells have
 * already written (a continuous FE shares face DoFs), projects to the Legendre modes (FullMatrix::vmult:
 * dst(i) = sum_j A(i,j) src(j), j ascending), scales the modes with degree >= Nc by sigma and projects back.
 * damped[i] != 0 marks the modes the reference multiplies (the others are left untouched, not multiplied by 1). */
void orc_exponential_filter(int64_t n_cells, int n, const int32_t *cell_dofs, const double *to_legendre,
                            const double *from_legendre, const double *sigma, const unsigned char *damped, double *v)
{
    double src[1024], leg[1024];
    for (int64_t c = 0; c < n_cells; c++) {
        const int32_t *idx = cell_dofs + c * n;
        for (int i = 0; i < n; i++) src[i] = v[idx[i]];
        for (int i = 0; i < n; i++) {
            double s = 0.0;
            for (int j = 0; j < n; j++) s += to_legendre[(size_t)i * n + j] * src[j];
            leg[i] = s;
        }
        for (int i = 0; i < n; i++) if (damped[i]) leg[i] = sigma[i] * leg[i];
        for (int i = 0; i < n; i++) {
            double s = 0.0;
            for (int j = 0; j < n; j++) s += from_legendre[(size_t)i * n + j] * leg[j];
            src[i] = s;
        }
        for (int i = 0; i < n; i++) v[idx[i]] = src[i];
    }
}

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_set_num_threads(int n)
{
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}
