// collide.cuh -- device-side collision operators of libnatrium_b200 (sm_100a).
//
// Pointwise, register-resident: every function works on one DoF's populations held in a
// thread's registers, so the same code is the epilogue of the fused stream+collide kernel
// and the body of the stand-alone collide kernel.  fp64 throughout; no tensor cores (the
// path is HBM-bound).
//
// What each function computes follows the reference (L = src/library/natrium):
//   density / velocity      L/collision_advanced/AuxiliaryCollisionFunctions.h:45-59,231-287
//   temperature             ...:289-307
//   BGK equilibrium         L/collision_advanced/Equilibria.h:33-83
//   quartic equilibrium     L/collision_advanced/Equilibria.h:119-267 (H3/H4 :519-566)
//   BGK relax               L/collision_advanced/CollisionSchemes.h:28-41
//   f+g relax (Prandtl fix, Sutherland, sensor)   CollisionSchemes.h:43-118, Aux...h:420-515
// The 1e-12-per-step parity contract fixes the operation order of these formulas, so the formula blocks themselves (the
// D2Q9 / generic BGK equilibrium, the quartic equilibrium's a_xxx .. a_xxyz coefficients, the hard-coded velocity index
// sums for D2Q9 / D3Q19) are TRANSCRIPTIONS of the reference's expressions, identifiers included (prefactor, uSquareTerm,
// mixedTerm, T1, a_*), about 70 lines; what surrounds them -- pre-reduced Hermite components, reciprocal constants,
// register templates, force hooks, the loop form for large stencils -- is this library's own.
// Unlike the reference, the Hermite tensors are reduced once on the host to their 10+15
// unique symmetric components and kept in constant memory (the reference recomputes the
// full tensors per DoF, CollisionOperator.h:68-69).  Divisions by run-time constants (cs2, tau,
// weights, 6 cs2^3, 24 cs2^4) are multiplications by reciprocals rounded once on the host: each
// differs from the reference's division by at most one ulp of the term, far inside the 1e-12
// per-step parity bound, and removes ~60 fp64 divisions per DoF from the D3Q19 epilogue.
#pragma once
#include <cstdint>

#include "nbconst.h"

__constant__ NbConst cP;

template <int Q>
__device__ __forceinline__ double nb_density(const double (&f)[Q])
{
    double rho = 0.0;
#pragma unroll
    for (int p = 0; p < Q; ++p) rho += f[p];
    return rho;
}

// calculateVelocity: hard-coded index sums for D2Q9 and D3Q19, generic loop otherwise.
template <int D, int Q>
__device__ __forceinline__ void nb_velocity(const double (&f)[Q], double rho, double (&u)[3])
{
    u[0] = u[1] = u[2] = 0.0;
    if (D == 2 && Q == 9) {
        u[0] = 1.0 / rho * (f[1] + f[5] + f[8] - f[3] - f[6] - f[7]);
        u[1] = 1.0 / rho * (f[2] + f[5] + f[6] - f[4] - f[7] - f[8]);
    } else if (D == 3 && Q == 19) {
        u[0] = 1.0 / rho * (f[1] - f[3] + f[7] - f[8] - f[9] + f[10] + f[11] + f[12] - f[13] - f[14]);
        u[1] = 1.0 / rho * (-f[5] + f[6] - f[11] + f[12] + f[13] - f[14] - f[15] + f[16] + f[17] - f[18]);
        u[2] = 1.0 / rho * (f[2] - f[4] + f[7] + f[8] - f[9] - f[10] + f[15] + f[16] - f[17] - f[18]);
    } else {
#pragma unroll
        for (int j = 0; j < D; j++) {
            double s = 0.0;
#pragma unroll
            for (int i = 0; i < Q; i++) s += cP.e[i][j] * f[i];
            u[j] = s * 1.0 / rho;
        }
    }
}

// BGKEquilibrium::calc
template <int D, int Q>
__device__ __forceinline__ void nb_feq_bgk(double rho, const double (&u)[3], double (&feq)[Q])
{
    const double cs2 = cP.cs2;
    if (D == 2 && Q == 9) {
        const double prefactor = cP.inv_cs2;
        const double scalar_product = u[0] * u[0] + u[1] * u[1];
        const double uSquareTerm = -scalar_product * cP.half_inv_cs2;
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
    const double inv_cs2 = cP.inv_cs2, half_inv_cs2 = cP.half_inv_cs2;
    double uu_term = 0.0;
#pragma unroll
    for (int j = 0; j < D; j++) uu_term += -(u[j] * u[j]) * half_inv_cs2;
#pragma unroll
    for (int i = 0; i < Q; i++) {
        double ue_term = 0.0;
#pragma unroll
        for (int j = 0; j < D; j++) ue_term += (u[j] * cP.e[i][j]) * inv_cs2;
        feq[i] = cP.w[i] * rho * (1 + ue_term * (1 + 0.5 * (ue_term)) + uu_term);
    }
}

// QuarticEquilibrium::polynomial, split in two: the moments of the Maxwellian that do not depend on the direction
// (nb_quartic_coef) and the value of one direction (nb_feq_quartic_i).  The array version below and the low-register
// f + g collision (which re-evaluates single directions instead of keeping feq[Q]) share them, so both give the same bits.
struct NbQuarticCoef {
    double rho, T, uu_term;
    double a_xxx, a_xxy, a_xyy, a_yyy, a_xxxx, a_yyyy, a_xxxy, a_xyyy, a_xxyy;
    double a_zzz, a_xxz, a_xzz, a_yzz, a_yyz, a_xyz;
    double a_zzzz, a_xzzz, a_xxzz, a_xxxz, a_yzzz, a_yyzz, a_yyyz, a_xxyz, a_xyyz, a_xyzz;
    double u[3];
};

template <int D>
__device__ __forceinline__ void nb_quartic_coef(double rho, const double (&u)[3], double T, NbQuarticCoef& c)
{
    const double cs2 = cP.cs2;
    const double half_inv_cs2 = cP.half_inv_cs2;
    c.rho = rho;
    c.T = T;
    c.u[0] = u[0]; c.u[1] = u[1]; c.u[2] = D == 3 ? u[2] : 0.0;
    double uu_term = 0.0;
#pragma unroll
    for (int j = 0; j < D; j++) uu_term += -(u[j] * u[j]) * half_inv_cs2;
    c.uu_term = uu_term;
    const double T1 = cs2 * (T - 1);
    c.a_xxx = u[0] * u[0] * u[0] + T1 * (u[0] + u[0] + u[0]);
    c.a_xxy = u[0] * u[0] * u[1] + T1 * (u[1]);
    c.a_xyy = u[0] * u[1] * u[1] + T1 * (u[0]);
    c.a_yyy = u[1] * u[1] * u[1] + T1 * (u[1] + u[1] + u[1]);
    c.a_xxxx = u[0] * u[0] * u[0] * u[0] + T1 * u[0] * u[0] * 6.0 + T1 * T1 * 3.0;
    c.a_yyyy = u[1] * u[1] * u[1] * u[1] + T1 * u[1] * u[1] * 6.0 + T1 * T1 * 3.0;
    c.a_xxxy = u[0] * u[0] * u[0] * u[1] + T1 * (u[0] * u[1] * 3.0);
    c.a_xyyy = u[0] * u[1] * u[1] * u[1] + T1 * (u[0] * u[1] * 3.0);
    c.a_xxyy = u[0] * u[0] * u[1] * u[1] + T1 * (u[0] * u[0] + u[1] * u[1]) + T1 * T1;
    c.a_zzz = c.a_xxz = c.a_xzz = c.a_yzz = c.a_yyz = c.a_xyz = 0.0;
    c.a_zzzz = c.a_xzzz = c.a_xxzz = c.a_xxxz = c.a_yzzz = c.a_yyzz = c.a_yyyz = c.a_xxyz = c.a_xyyz = c.a_xyzz = 0.0;
    if (D == 3) {
        c.a_zzz = u[2] * u[2] * u[2] + T1 * (u[2] + u[2] + u[2]);
        c.a_xxz = u[0] * u[0] * u[2] + T1 * (u[2]);
        c.a_xzz = u[0] * u[2] * u[2] + T1 * (u[0]);
        c.a_yzz = u[1] * u[2] * u[2] + T1 * (u[1]);
        c.a_yyz = u[1] * u[1] * u[2] + T1 * (u[2]);
        c.a_xyz = u[0] * u[1] * u[2];
        c.a_zzzz = u[2] * u[2] * u[2] * u[2] + T1 * u[2] * u[2] * 6.0 + T1 * T1 * 3.0;
        c.a_xxxz = u[0] * u[0] * u[0] * u[2] + T1 * (u[0] * u[2] * 3.0);
        c.a_yyyz = u[1] * u[1] * u[1] * u[2] + T1 * (u[1] * u[2] * 3.0);
        c.a_xzzz = u[0] * u[2] * u[2] * u[2] + T1 * (u[0] * u[2] * 3.0);
        c.a_yzzz = u[1] * u[2] * u[2] * u[2] + T1 * (u[1] * u[2] * 3.0);
        c.a_xxzz = u[0] * u[0] * u[2] * u[2] + T1 * (u[0] * u[0] + u[2] * u[2]) + T1 * T1;
        c.a_yyzz = u[1] * u[1] * u[2] * u[2] + T1 * (u[1] * u[1] + u[2] * u[2]) + T1 * T1;
        c.a_xyzz = u[0] * u[1] * u[2] * u[2] + T1 * (u[0] * u[1]);
        c.a_xyyz = u[0] * u[1] * u[1] * u[2] + T1 * (u[0] * u[2]);
        c.a_xxyz = u[0] * u[0] * u[1] * u[2] + T1 * (u[1] * u[2]);
    }
}

template <int D, int Q>
__device__ __forceinline__ double nb_feq_quartic_i(int i, const NbQuarticCoef& c)
{
    const double cs2 = cP.cs2;
    const double inv_cs2 = cP.inv_cs2, half_inv_cs2 = cP.half_inv_cs2;
    const double inv_c3 = cP.inv_c3;     // 1 / (6 cs2^3)
    const double inv_c4 = cP.inv_c4;     // 1 / (24 cs2^4)
    const double rho = c.rho, T = c.T;
    const double w = cP.w[i];
    double ue_term = 0.0;
#pragma unroll
    for (int j = 0; j < D; j++) ue_term += (c.u[j] * cP.e[i][j]) * inv_cs2;
    double fe = w * rho * (1 + ue_term * (1 + 0.5 * (ue_term)) + c.uu_term);
    // (T-1) trace term: the reference's alp/bet double loop only has diagonal contributions
#pragma unroll
    for (int a = 0; a < D; a++)
        fe += rho * w * half_inv_cs2 * ((T - 1) * cP.e[i][a] * cP.e[i][a] - cs2 * (T - 1));
    const double* H3 = cP.H3[i];
    const double* H4 = cP.H4[i];
    fe += w * rho * inv_c3 * (c.a_xxx * H3[0] + 3 * (c.a_xxy * H3[1] + c.a_xyy * H3[2]) + c.a_yyy * H3[3]);
    if (D == 3)
        fe += w * rho * inv_c3
            * (c.a_zzz * H3[4] + 3 * (c.a_xxz * H3[5] + c.a_xzz * H3[6] + c.a_yzz * H3[7] + c.a_yyz * H3[8])
               + 6.0 * c.a_xyz * H3[9]);
    fe += w * rho * inv_c4
        * (H4[0] * c.a_xxxx + H4[1] * c.a_yyyy + 6.0 * H4[4] * c.a_xxyy + 4.0 * H4[3] * c.a_xyyy + 4.0 * H4[2] * c.a_xxxy);
    if (D == 3)
        fe += w * rho * inv_c4
            * (H4[5] * c.a_zzzz + 4.0 * (H4[6] * c.a_xzzz + H4[9] * c.a_yzzz + H4[8] * c.a_xxxz + H4[11] * c.a_yyyz)
               + 6.0 * (H4[7] * c.a_xxzz + H4[10] * c.a_yyzz)
               + 12.0 * (H4[12] * c.a_xxyz + H4[13] * c.a_xyyz + H4[14] * c.a_xyzz));
    return fe;
}

template <int D, int Q>
__device__ __forceinline__ void nb_feq_quartic(double rho, const double (&u)[3], double T, double (&feq)[Q])
{
    NbQuarticCoef c;
    nb_quartic_coef<D>(rho, u, T, c);
#pragma unroll
    for (int i = 0; i < Q; i++) feq[i] = nb_feq_quartic_i<D, Q>(i, c);
}

// BGKEquilibrium::calc for one direction of the generic (not D2Q9) branch
template <int D, int Q>
__device__ __forceinline__ double nb_feq_bgk_i(int i, double rho, const double (&u)[3], double uu_term)
{
    double ue_term = 0.0;
#pragma unroll
    for (int j = 0; j < D; j++) ue_term += (u[j] * cP.e[i][j]) * cP.inv_cs2;
    return cP.w[i] * rho * (1 + ue_term * (1 + 0.5 * (ue_term)) + uu_term);
}

// collideAll body, f only: returns rho, u (unscaled); relaxes f in registers.
// u_override != nullptr mirrors inInitializationProcedure (velocity taken from the global vector).
template <int D, int Q, int EQ>
__device__ __forceinline__ void nb_collide_bgk(double (&f)[Q], double& rho, double (&u)[3], const double* u_override)
{
    rho = nb_density<Q>(f);
    nb_velocity<D, Q>(f, rho, u);
    if (u_override) {
#pragma unroll
        for (int j = 0; j < D; j++) u[j] = u_override[j] / cP.scaling;
    }
    double feq[Q];
    if (EQ == NB_EQ_BGK) nb_feq_bgk<D, Q>(rho, u, feq);
    else nb_feq_quartic<D, Q>(rho, u, 1.0, feq);
    const double omega = cP.inv_tau;
#pragma unroll
    for (int p = 0; p < Q; ++p) f[p] -= omega * (f[p] - feq[p]);
}

// MultipleRelaxationTime tables: what make_M / make_T / make_diag return (AuxiliaryMRTFunctions.cpp), handed over
// through nb200_set_mrt().  Only stencils with an MRT row in selectCollision carry them (Q <= 19).
#define NB_MRT_MAXQ 19
struct NbMrtStd {
    double M[NB_MRT_MAXQ][NB_MRT_MAXQ];
    double T[NB_MRT_MAXQ][NB_MRT_MAXQ];
    double omega[NB_MRT_MAXQ];
};
__constant__ NbMrtStd cS;

// Per-DoF linear map applied after the collision: the matrix of PseudoEntropicStabilizer::apply
// (L/dataprocessors/PseudoEntropicStabilizer.cpp:27-150,225-262), handed over through nb200_set_post_collision_matrix().
__constant__ double cA[NB_MRT_MAXQ][NB_MRT_MAXQ];

// collideAll body, f only, every scheme of selectCollision on the path (CollisionOperator.h:52-104):
// SCHEME 0 BGKCollision::relax (CollisionSchemes.h:28-41), 1 Regularized::relax (:150-203),
// 2 MultipleRelaxationTime::relax (:238-263).  FORCE compiles the external-force hooks in
// (applyMacroscopicForces / applyForces / postCollisionApplyForces, Aux...h:332-417; force type from cP).
// v_out: scaled velocity written to the global vector (input when in_init); u: velocity the equilibrium used.
template <int D, int Q, int EQ, int SCHEME, bool FORCE>
__device__ __forceinline__ void nb_collide_adv(double (&f)[Q], double& rho, double (&v_out)[3], bool in_init)
{
    double u[3];
    rho = nb_density<Q>(f);
    nb_velocity<D, Q>(f, rho, u);
    if (!in_init) {
#pragma unroll
        for (int j = 0; j < D; j++) v_out[j] = u[j] * cP.scaling;
        if (FORCE && cP.force_type == 1) {      // SHIFTING_VELOCITY
#pragma unroll
            for (int j = 0; j < D; j++) {
                v_out[j] = v_out[j] + 0.5 * cP.dt * cP.force[j] / rho;
                u[j] += cP.tau * cP.dt * cP.force[j] / rho / cP.scaling;
            }
        }
    } else {
#pragma unroll
        for (int j = 0; j < D; j++) u[j] = v_out[j] / cP.scaling;
    }
    double feq[Q];
    if (EQ == NB_EQ_BGK) nb_feq_bgk<D, Q>(rho, u, feq);
    else nb_feq_quartic<D, Q>(rho, u, 1.0, feq);
    if (SCHEME == 0) {
        const double omega = cP.inv_tau;
#pragma unroll
        for (int p = 0; p < Q; ++p) f[p] -= omega * (f[p] - feq[p]);
    } else if (SCHEME == 1) {
        double pi[3][3];
#pragma unroll
        for (int m = 0; m < 3; m++)
#pragma unroll
            for (int n = 0; n < 3; n++) pi[m][n] = 0.0;
        // pi - pieq accumulated as two sums and subtracted, as the reference does (no cancellation shortcuts)
        double pieq[3][3];
#pragma unroll
        for (int m = 0; m < 3; m++)
#pragma unroll
            for (int n = 0; n < 3; n++) pieq[m][n] = 0.0;
#pragma unroll
        for (int j = 0; j < Q; j++)
#pragma unroll
            for (int m = 0; m < D; m++)
#pragma unroll
                for (int n = 0; n < D; n++) {
                    pi[m][n] += f[j] * cP.e[j][m] * cP.e[j][n];
                    pieq[m][n] += feq[j] * cP.e[j][m] * cP.e[j][n];
                }
#pragma unroll
        for (int m = 0; m < D; m++)
#pragma unroll
            for (int n = 0; n < D; n++) pi[m][n] -= pieq[m][n];
        const double cs2 = cP.cs2;
        const double pref = 1.0 / (2 * cs2 * cs2);
        const double keep = 1. - cP.inv_tau;
#pragma unroll
        for (int a = 0; a < Q; a++) {
            double fi1 = 0.0;
#pragma unroll
            for (int b = 0; b < D; b++)
#pragma unroll
                for (int c = 0; c < D; c++) {
                    const double Qabc = cP.e[a][b] * cP.e[a][c] - (b == c ? cs2 : 0.0);
                    fi1 += cP.w[a] * pref * Qabc * pi[b][c];
                }
            f[a] = feq[a] + keep * fi1;
        }
    } else {
        double m[Q];
#pragma unroll
        for (int i = 0; i < Q; i++) {
            double mi = 0.0, meq = 0.0;
#pragma unroll
            for (int j = 0; j < Q; j++) mi += cS.M[i][j] * f[j];
#pragma unroll
            for (int j = 0; j < Q; j++) meq += cS.M[i][j] * feq[j];
            m[i] = mi - cS.omega[i] * (mi - meq);
        }
#pragma unroll
        for (int i = 0; i < Q; i++) {
            double acc = 0.0;
#pragma unroll
            for (int j = 0; j < Q; j++) acc += cS.T[i][j] * m[j];
            f[i] = acc;
        }
    }
    if (FORCE && cP.force_type == 2) {          // EXACT_DIFFERENCE
#pragma unroll
        for (int j = 0; j < D; j++) {
            u[j] += cP.dt * cP.force[j] / rho / cP.scaling;
            v_out[j] = v_out[j] + 0.5 * cP.dt * cP.force[j] / rho;
        }
        double shifted[Q];
        if (EQ == NB_EQ_BGK) nb_feq_bgk<D, Q>(rho, u, shifted);
        else nb_feq_quartic<D, Q>(rho, u, 1.0, shifted);
#pragma unroll
        for (int i = 0; i < Q; i++) f[i] += (shifted[i] - feq[i]);
    }
}

// collideAll body, f + g (relaxWithG).  Writes T and the Knudsen-estimate sensor.
// FORCE: external-force hooks; v_force then receives the value of the global (scaled) velocity vector:
// raw moment * scaling (or the stored velocity in the initialization procedure) + 0.5 dt F / rho.
template <int D, int Q, int EQ, bool FORCE = false>
__device__ __forceinline__ void nb_collide_bgk_fg(double (&f)[Q], double (&g)[Q], double& rho, double (&u)[3],
                                                  double& T, double& sensor, const double* u_override,
                                                  double* v_force = nullptr)
{
    const double cs2 = cP.cs2;
    rho = nb_density<Q>(f);
    nb_velocity<D, Q>(f, rho, u);
    // calculateTemperature
    double Tacc = 0.0;
#pragma unroll
    for (int i = 0; i < Q; i++) {
        double sum = 0.0;
#pragma unroll
        for (int a = 0; a < D; a++) sum += (cP.e[i][a] - u[a]) * (cP.e[i][a] - u[a]);
        Tacc += sum * f[i] * cP.inv_cs2 + g[i];
    }
    const double C_v = cP.Cv;
    T = Tacc * 0.5 / (rho * C_v);
    if (u_override) {
#pragma unroll
        for (int j = 0; j < D; j++) u[j] = u_override[j] / cP.scaling;
    }
    if (FORCE) {
#pragma unroll
        for (int j = 0; j < D; j++) v_force[j] = u_override ? u_override[j] : u[j] * cP.scaling;
        if (!u_override && cP.force_type == 1) {     // SHIFTING_VELOCITY (not in the initialization procedure)
#pragma unroll
            for (int j = 0; j < D; j++) {
                v_force[j] = v_force[j] + 0.5 * cP.dt * cP.force[j] / rho;
                u[j] += cP.tau * cP.dt * cP.force[j] / rho / cP.scaling;
            }
        }
    }
    double feq[Q];
    if (EQ == NB_EQ_BGK) nb_feq_bgk<D, Q>(rho, u, feq);
    else nb_feq_quartic<D, Q>(rho, u, T, feq);
    const double gfac = (T) * (2.0 * C_v - D);

    // non-equilibrium moments for the Prandtl correction
    double Qn[3][3][3];
    double qg[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
        for (int b = 0; b < 3; b++)
#pragma unroll
            for (int c = 0; c < 3; c++) Qn[a][b][c] = 0.0;
    double knudsen = 0.0;
#pragma unroll
    for (int i = 0; i < Q; i++) {
        const double fneq = f[i] - feq[i];
        const double gneq = g[i] - feq[i] * gfac;
        knudsen += fabs(f[i] - feq[i]) * cP.inv_w[i];
        if (cP.prandtl_set) {
            double c[3];
#pragma unroll
            for (int a = 0; a < D; a++) c[a] = cP.e[i][a] - u[a];
#pragma unroll
            for (int a = 0; a < D; a++)
#pragma unroll
                for (int b = 0; b < D; b++)
#pragma unroll
                    for (int cc = 0; cc < D; cc++) Qn[a][b][cc] += (c[a] * c[b] * c[cc]) * fneq;
#pragma unroll
            for (int a = 0; a < D; a++) qg[a] += c[a] * gneq;
        }
    }
    sensor = knudsen / Q;

    double sutherland_factor = 1.0;
    if (cP.sutherland_set) sutherland_factor = pow(T / 0.85, 0.7);
    const double visc_tau = (cP.tau - 0.5) * sutherland_factor / (T * rho) + 0.5;
    const double prandtl_tau = (visc_tau - 0.5) / cP.prandtl + 0.5;
    const double visc_omega = 1. / visc_tau;
    const double prandtl_omega = 1. / prandtl_tau;
    const double prandtl_diff = visc_omega - prandtl_omega;
    const double inv_cs6 = cP.inv_c3;
    const double inv_T = 1.0 / T;
#pragma unroll
    for (int i = 0; i < Q; i++) {
        double fStar = 0.0, gStar = 0.0;
        if (cP.prandtl_set) {
#pragma unroll
            for (int a = 0; a < D; a++)
#pragma unroll
                for (int b = 0; b < D; b++)
#pragma unroll
                    for (int c = 0; c < D; c++)
                        fStar += cP.w[i] * (Qn[a][b][c] * (cP.e[i][a] * cP.e[i][b] * cP.e[i][c]
                                                            - 3 * cs2 * cP.e[i][c] * (a == b ? 1.0 : 0.0))) * inv_cs6;
#pragma unroll
            for (int a = 0; a < D; a++) gStar += cP.w[i] * (qg[a] * cP.e[i][a]) * inv_T;
        }
        const double fneq = f[i] - feq[i];
        const double gneq = g[i] - feq[i] * gfac;
        f[i] -= visc_omega * fneq - prandtl_diff * fStar;
        g[i] -= visc_omega * gneq - prandtl_diff * gStar;
    }
    if (FORCE && cP.force_type == 2) {          // EXACT_DIFFERENCE: f only (Aux...h:399-411)
#pragma unroll
        for (int j = 0; j < D; j++) {
            u[j] += cP.dt * cP.force[j] / rho / cP.scaling;
            v_force[j] = v_force[j] + 0.5 * cP.dt * cP.force[j] / rho;
        }
        double shifted[Q];
        if (EQ == NB_EQ_BGK) nb_feq_bgk<D, Q>(rho, u, shifted);
        else nb_feq_quartic<D, Q>(rho, u, T, shifted);
#pragma unroll
        for (int i = 0; i < Q; i++) f[i] += (shifted[i] - feq[i]);
    }
}

// ---------------------------------------------------------------------------------------------
// collideAll body, f + g, for the large stencils (D3Q45: f[45], g[45] and feq[45] alone would be 270 registers, and the
// fully unrolled form is ~150 KB of code: 255 registers, kilobytes of spills, instruction-cache misses, 8 warps per SM).
// Same quantities as nb_collide_bgk_fg, organised as loops over the directions that are NOT unrolled:
//   1  rho, sum e f     2  temperature     3  f_eq, non-equilibrium moments, sensor
//   4  f_eq again, f*, g*, relaxation, store     (5  exact-difference force)
// The populations are read through accessors in every loop (global memory, mostly L2 hits) instead of being held in
// registers, and the per-direction constants come from a shared-memory table (NB_CT_* below; constant-bank operands need
// compile-time addresses, i.e. unrolling).  The third-order non-equilibrium moment Q_abc is symmetric, so its 10 (3-d) /
// 4 (2-d) distinct components are accumulated once, and f*_i = w_i / (6 cs^6) sum_abc Q_abc (e_a e_b e_c - 3 cs^2 e_c delta_ab)
// (Aux...h:486-501) is contracted with the Hermite components H3_i of the table (for symmetric Q the two contractions are
// the same number): 10 multiply-adds per direction instead of 27 x 5 operations.  Differences to the reference's
// term-by-term sums are a few ulp of f*, which enters f scaled by (omega_visc - omega_Pr).
// ---------------------------------------------------------------------------------------------
#define NB_CT_PITCH 32      // doubles per direction: e[3] w inv_w . | H3[10] at 6 | H4[15] at 16 (.. 30)
#define NB_CT_E 0
#define NB_CT_W 3
#define NB_CT_INVW 4
#define NB_CT_H3 6
#define NB_CT_H4 16

// fills the table from the constant block (whole CTA; the caller synchronises)
template <int D, int Q>
__device__ __forceinline__ void nb_ct_fill(double* ct, int tid, int nthreads)
{
    for (int t = tid; t < Q * NB_CT_PITCH; t += nthreads) {
        const int i = t / NB_CT_PITCH, k = t % NB_CT_PITCH;
        double v = 0.0;
        if (k < 3) v = cP.e[i][k];
        else if (k == NB_CT_W) v = cP.w[i];
        else if (k == NB_CT_INVW) v = cP.inv_w[i];
        else if (k >= NB_CT_H3 && k < NB_CT_H3 + 10) v = cP.H3[i][k - NB_CT_H3];
        else if (k >= NB_CT_H4 && k < NB_CT_H4 + 15) v = cP.H4[i][k - NB_CT_H4];
        ct[t] = v;
    }
}

// nb_feq_quartic_i with the direction's constants taken from the table row c_i (same operations, same order)
template <int D>
__device__ __forceinline__ double nb_feq_quartic_ct(const double* __restrict__ ci, const NbQuarticCoef& c)
{
    const double cs2 = cP.cs2;
    const double inv_cs2 = cP.inv_cs2, half_inv_cs2 = cP.half_inv_cs2;
    const double inv_c3 = cP.inv_c3, inv_c4 = cP.inv_c4;
    const double rho = c.rho, T = c.T;
    const double w = ci[NB_CT_W];
    double ue_term = 0.0;
#pragma unroll
    for (int j = 0; j < D; j++) ue_term += (c.u[j] * ci[NB_CT_E + j]) * inv_cs2;
    double fe = w * rho * (1 + ue_term * (1 + 0.5 * (ue_term)) + c.uu_term);
#pragma unroll
    for (int a = 0; a < D; a++)
        fe += rho * w * half_inv_cs2 * ((T - 1) * ci[NB_CT_E + a] * ci[NB_CT_E + a] - cs2 * (T - 1));
    const double* H3 = ci + NB_CT_H3;
    const double* H4 = ci + NB_CT_H4;
    fe += w * rho * inv_c3 * (c.a_xxx * H3[0] + 3 * (c.a_xxy * H3[1] + c.a_xyy * H3[2]) + c.a_yyy * H3[3]);
    if (D == 3)
        fe += w * rho * inv_c3
            * (c.a_zzz * H3[4] + 3 * (c.a_xxz * H3[5] + c.a_xzz * H3[6] + c.a_yzz * H3[7] + c.a_yyz * H3[8])
               + 6.0 * c.a_xyz * H3[9]);
    fe += w * rho * inv_c4
        * (H4[0] * c.a_xxxx + H4[1] * c.a_yyyy + 6.0 * H4[4] * c.a_xxyy + 4.0 * H4[3] * c.a_xyyy + 4.0 * H4[2] * c.a_xxxy);
    if (D == 3)
        fe += w * rho * inv_c4
            * (H4[5] * c.a_zzzz + 4.0 * (H4[6] * c.a_xzzz + H4[9] * c.a_yzzz + H4[8] * c.a_xxxz + H4[11] * c.a_yyyz)
               + 6.0 * (H4[7] * c.a_xxzz + H4[10] * c.a_yyzz)
               + 12.0 * (H4[12] * c.a_xxyz + H4[13] * c.a_xyyz + H4[14] * c.a_xyzz));
    return fe;
}

template <int D>
__device__ __forceinline__ double nb_feq_bgk_ct(const double* __restrict__ ci, double rho, const double (&u)[3], double uu_term)
{
    double ue_term = 0.0;
#pragma unroll
    for (int j = 0; j < D; j++) ue_term += (u[j] * ci[NB_CT_E + j]) * cP.inv_cs2;
    return ci[NB_CT_W] * rho * (1 + ue_term * (1 + 0.5 * (ue_term)) + uu_term);
}

template <int D, int Q, int EQ, bool FORCE, class LdF, class LdG, class St, class StF>
__device__ __forceinline__ void nb_collide_fg_loops(const double* __restrict__ ct, LdF ldf, LdG ldg, St store, StF store_f, double& rho,
                                                    double (&u)[3], double& T, double& sensor, const double* u_override, double* v_force)
{
    // loop 1: density and momentum (calculateDensity, calculateVelocity: one sum per component, i ascending)
    double r = 0.0, s[3] = {0.0, 0.0, 0.0};
#pragma unroll 3
    for (int i = 0; i < Q; i++) {
        const double fi = ldf(i);
        const double* ci = ct + i * NB_CT_PITCH;
        r += fi;
#pragma unroll
        for (int j = 0; j < D; j++) s[j] += ci[NB_CT_E + j] * fi;
    }
    rho = r;
    u[0] = u[1] = u[2] = 0.0;
#pragma unroll
    for (int j = 0; j < D; j++) u[j] = s[j] * 1.0 / rho;
    // loop 2: calculateTemperature
    double Tacc = 0.0;
#pragma unroll 3
    for (int i = 0; i < Q; i++) {
        const double* ci = ct + i * NB_CT_PITCH;
        double sum = 0.0;
#pragma unroll
        for (int a = 0; a < D; a++) sum += (ci[NB_CT_E + a] - u[a]) * (ci[NB_CT_E + a] - u[a]);
        Tacc += sum * ldf(i) * cP.inv_cs2 + ldg(i);
    }
    const double C_v = cP.Cv;
    T = Tacc * 0.5 / (rho * C_v);
    if (u_override) {
#pragma unroll
        for (int j = 0; j < D; j++) u[j] = u_override[j] / cP.scaling;
    }
    if (FORCE) {
#pragma unroll
        for (int j = 0; j < D; j++) v_force[j] = u_override ? u_override[j] : u[j] * cP.scaling;
        if (!u_override && cP.force_type == 1) {     // SHIFTING_VELOCITY (not in the initialization procedure)
#pragma unroll
            for (int j = 0; j < D; j++) {
                v_force[j] = v_force[j] + 0.5 * cP.dt * cP.force[j] / rho;
                u[j] += cP.tau * cP.dt * cP.force[j] / rho / cP.scaling;
            }
        }
    }
    NbQuarticCoef qc;
    double uu_bgk = 0.0;
    if (EQ == NB_EQ_BGK) {
#pragma unroll
        for (int j = 0; j < D; j++) uu_bgk += -(u[j] * u[j]) * cP.half_inv_cs2;
    } else {
        nb_quartic_coef<D>(rho, u, T, qc);
    }
    const double gfac = (T) * (2.0 * C_v - D);
    // loop 3: non-equilibrium moments.  Unique components: xxx xxy xyy yyy | zzz xxz xzz yzz yyz xyz (order of H3)
    double Qs[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    double qg[3] = {0.0, 0.0, 0.0};
    double knudsen = 0.0;
    const bool pr = cP.prandtl_set != 0;
#pragma unroll 1
    for (int i = 0; i < Q; i++) {
        const double* ci = ct + i * NB_CT_PITCH;
        const double feq = EQ == NB_EQ_BGK ? nb_feq_bgk_ct<D>(ci, rho, u, uu_bgk) : nb_feq_quartic_ct<D>(ci, qc);
        const double fi = ldf(i);
        const double fneq = fi - feq;
        knudsen += fabs(fi - feq) * ci[NB_CT_INVW];
        if (pr) {
            const double gneq = ldg(i) - feq * gfac;
            double c[3] = {0.0, 0.0, 0.0};
#pragma unroll
            for (int a = 0; a < D; a++) c[a] = ci[NB_CT_E + a] - u[a];
            Qs[0] += (c[0] * c[0] * c[0]) * fneq;
            Qs[1] += (c[0] * c[0] * c[1]) * fneq;
            Qs[2] += (c[0] * c[1] * c[1]) * fneq;
            Qs[3] += (c[1] * c[1] * c[1]) * fneq;
            if (D == 3) {
                Qs[4] += (c[2] * c[2] * c[2]) * fneq;
                Qs[5] += (c[0] * c[0] * c[2]) * fneq;
                Qs[6] += (c[0] * c[2] * c[2]) * fneq;
                Qs[7] += (c[1] * c[2] * c[2]) * fneq;
                Qs[8] += (c[1] * c[1] * c[2]) * fneq;
                Qs[9] += (c[0] * c[1] * c[2]) * fneq;
            }
#pragma unroll
            for (int a = 0; a < D; a++) qg[a] += c[a] * gneq;
        }
    }
    sensor = knudsen / Q;
    double sutherland_factor = 1.0;
    if (cP.sutherland_set) sutherland_factor = pow(T / 0.85, 0.7);
    const double visc_tau = (cP.tau - 0.5) * sutherland_factor / (T * rho) + 0.5;
    const double prandtl_tau = (visc_tau - 0.5) / cP.prandtl + 0.5;
    const double visc_omega = 1. / visc_tau;
    const double prandtl_omega = 1. / prandtl_tau;
    const double prandtl_diff = visc_omega - prandtl_omega;
    const double inv_cs6 = cP.inv_c3;
    const double inv_T = 1.0 / T;
    // loop 4: relaxation
#pragma unroll 1
    for (int i = 0; i < Q; i++) {
        const double* ci = ct + i * NB_CT_PITCH;
        const double feq = EQ == NB_EQ_BGK ? nb_feq_bgk_ct<D>(ci, rho, u, uu_bgk) : nb_feq_quartic_ct<D>(ci, qc);
        double fi = ldf(i), gi = ldg(i);
        double fStar = 0.0, gStar = 0.0;
        if (pr) {
            const double* H3 = ci + NB_CT_H3;
            double S = Qs[0] * H3[0] + 3 * (Qs[1] * H3[1] + Qs[2] * H3[2]) + Qs[3] * H3[3];
            if (D == 3) S += Qs[4] * H3[4] + 3 * (Qs[5] * H3[5] + Qs[6] * H3[6] + Qs[7] * H3[7] + Qs[8] * H3[8]) + 6.0 * Qs[9] * H3[9];
            fStar = ci[NB_CT_W] * S * inv_cs6;
#pragma unroll
            for (int a = 0; a < D; a++) gStar += ci[NB_CT_W] * (qg[a] * ci[NB_CT_E + a]) * inv_T;
        }
        const double fneq = fi - feq;
        const double gneq = gi - feq * gfac;
        fi -= visc_omega * fneq - prandtl_diff * fStar;
        gi -= visc_omega * gneq - prandtl_diff * gStar;
        store(i, fi, gi);
    }
    if (FORCE && cP.force_type == 2) {          // EXACT_DIFFERENCE: f only (Aux...h:399-411)
        double u2[3] = {u[0], u[1], u[2]};
#pragma unroll
        for (int j = 0; j < D; j++) {
            u2[j] += cP.dt * cP.force[j] / rho / cP.scaling;
            v_force[j] = v_force[j] + 0.5 * cP.dt * cP.force[j] / rho;
        }
        NbQuarticCoef qs;
        double uu2 = 0.0;
        if (EQ == NB_EQ_BGK) {
#pragma unroll
            for (int j = 0; j < D; j++) uu2 += -(u2[j] * u2[j]) * cP.half_inv_cs2;
        } else {
            nb_quartic_coef<D>(rho, u2, T, qs);
        }
#pragma unroll 1
        for (int i = 0; i < Q; i++) {
            const double* ci = ct + i * NB_CT_PITCH;
            const double feq = EQ == NB_EQ_BGK ? nb_feq_bgk_ct<D>(ci, rho, u, uu_bgk) : nb_feq_quartic_ct<D>(ci, qc);
            const double shifted = EQ == NB_EQ_BGK ? nb_feq_bgk_ct<D>(ci, rho, u2, uu2) : nb_feq_quartic_ct<D>(ci, qs);
            store_f(i, shifted - feq);
        }
#pragma unroll
        for (int j = 0; j < D; j++) u[j] = u2[j];
    }
}
