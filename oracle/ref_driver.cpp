// ref_driver.cpp -- C entry points over the REFERENCE's own collision / stencil code (test infrastructure, part of oracle/).
//
// Compiled by oracle/Makefile.ref together with the reference sources where they lie under /root/reference (a scratch
// tree of symbolic links + the stand-in headers of oracle/ref_stubs/, see its README) into oracle/_ref/libnatrium_ref.so.
// Nothing of the reference is restated here: this file only builds the arguments and calls
//   natrium::selectCollision<2|3>        L/collision_advanced/CollisionSelection.h:69-272 (both overloads)
//   natrium::KBCStandard / MRTEntropic / MRTStandard / BGKStandard ::collideAll      L/collision/*.cpp
//   natrium::D2Q9 / D3Q19 / D3Q15 / D2Q25H / D3Q45                                   L/stencils/*.cpp
//   AuxiliaryMRTFunctions::make_M / make_T / make_diag                               L/collision_advanced/AuxiliaryMRTFunctions.cpp
// Used by oracle/ref.py -> tests/test_oracle_vs_ref.py (oracle restatement == reference code) and, through the
// oracle, as the anchor of every GPU parity test.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>

#include "collision/CollisionModel.h"
#include "collision/BGKStandard.h"
#include "collision/KBCStandard.h"
#include "collision/MRTEntropic.h"
#include "collision/MRTStandard.h"
#include "collision_advanced/CollisionSelection.h"
#include "stencils/D2Q9.h"
#include "stencils/D3Q19.h"
#include "stencils/D3Q15.h"
#include "stencils/D2Q25H.h"
#include "stencils/D3Q45.h"

using namespace natrium;

namespace {

boost::shared_ptr<Stencil> make_stencil(const char* name, double scaling)
{
    const std::string s(name);
    if (s == "D2Q9") return boost::make_shared<D2Q9>(scaling);
    if (s == "D3Q19") return boost::make_shared<D3Q19>(scaling);
    if (s == "D3Q15") return boost::make_shared<D3Q15>(scaling);
    if (s == "D2Q25H") return boost::make_shared<D2Q25H>(scaling);
    if (s == "D3Q45") return boost::make_shared<D3Q45>(scaling);
    return boost::shared_ptr<Stencil>();
}

void put_err(char* err, int errlen, const char* what)
{
    if (err && errlen > 0) {
        strncpy(err, what, (size_t)errlen - 1);
        err[errlen - 1] = 0;
    }
}

template <size_t dim>
void fill_problem(ProblemDescription<dim>& pd, int has_force, const double* force)
{
    if (!has_force) return;
    dealii::Tensor<1, dim> F;
    for (size_t j = 0; j < dim; j++) F[j] = force[j];
    pd.setExternalForce(boost::make_shared<ConstantExternalForce<dim>>(F));
}

// exception -> status: what CFDSolver::collide() would report (CFDSolver.cpp:840-842)
template <class Fn>
int guarded(Fn&& fn, char* err, int errlen)
{
    try {
        fn();
    } catch (const DensityZeroException& e) {
        put_err(err, errlen, e.what());
        return -4;
    } catch (const CollisionException& e) {
        put_err(err, errlen, e.what());
        return -1;
    } catch (const NotImplementedException& e) {
        put_err(err, errlen, e.what());
        return -3;
    } catch (const NATriuMException& e) {
        put_err(err, errlen, e.what());
        return -2;
    } catch (const std::exception& e) {
        put_err(err, errlen, e.what());
        return -9;
    }
    return 0;
}

}  // namespace

extern "C" {

// Stencil tables of the reference classes.  e: [Q][D] scaled directions, w: [Q], opposite: [Q].
int ref_stencil(const char* name, double scaling, int* D, int* Q, double* e, double* w, double* cs2_scaled,
                double* max_speed, int* opposite)
{
    boost::shared_ptr<Stencil> st = make_stencil(name, scaling);
    if (!st) return -1;
    *D = (int)st->getD();
    *Q = (int)st->getQ();
    if (e)
        for (size_t i = 0; i < st->getQ(); i++)
            for (size_t j = 0; j < st->getD(); j++) e[i * st->getD() + j] = st->getDirection(i)(j);
    if (w)
        for (size_t i = 0; i < st->getQ(); i++) w[i] = st->getWeight(i);
    if (cs2_scaled) *cs2_scaled = st->getSpeedOfSoundSquare();
    if (max_speed) *max_speed = st->getMaxParticleVelocityMagnitude();
    if (opposite)
        for (size_t i = 0; i < st->getQ(); i++) opposite[i] = (int)st->getIndexOfOppositeDirection(i);
    return 0;
}

// selectCollision on caller-owned arrays (f, g: [Q][stride]; u: [D][n]).  Enums are passed as the reference's own
// integer values (ConfigNames.h).  with_g selects the compressible overload.
int ref_select_collision(const char* stencil_name, double scaling, int scheme, int equilibrium, int forcing,
                         int has_force, const double* force, int mrt_basis, int relax_mode, double viscosity, double dt,
                         int in_init, int with_g, double gamma, int prandtl_set, double prandtl, int sutherland,
                         int64_t n, int64_t stride, double* f, double* g, double* rho, double* u, double* T,
                         double* sensor, char* err, int errlen)
{
    boost::shared_ptr<Stencil> st = make_stencil(stencil_name, scaling);
    if (!st) { put_err(err, errlen, "unknown stencil"); return -8; }
    SolverConfiguration cfg;
    cfg.stencil = st->getStencilType();
    cfg.collision = (CollisionSchemeName)scheme;
    cfg.equilibrium = (EquilibriumSchemeName)equilibrium;
    cfg.forcing = (ForceType)forcing;
    cfg.mrt_basis = (MomentBasis)mrt_basis;
    cfg.mrt_relax = (RelaxMode)relax_mode;
    cfg.gamma = gamma;
    cfg.prandtl = prandtl;
    cfg.prandtl_set = prandtl_set != 0;
    cfg.sutherland_set = sutherland != 0;
    const size_t Q = st->getQ(), D = st->getD();
    DistributionFunctions F(f, Q, (size_t)n, (size_t)stride);
    distributed_vector densities(rho, (size_t)n);
    vector<distributed_vector> velocities;
    for (size_t j = 0; j < D; j++) velocities.emplace_back(u + j * (size_t)n, (size_t)n);
    const dealii::IndexSet owned((size_t)n);
    return guarded([&]() {
        if (D == 2) {
            ProblemDescription<2> pd;
            fill_problem<2>(pd, has_force, force);
            if (with_g) {
                DistributionFunctions G(g, Q, (size_t)n, (size_t)stride);
                distributed_vector temperature(T, (size_t)n), mss(sensor, (size_t)n);
                selectCollision<2>(cfg, pd, F, G, densities, velocities, temperature, mss, owned, viscosity, dt, *st, in_init != 0);
            } else {
                selectCollision<2>(cfg, pd, F, densities, velocities, owned, viscosity, dt, *st, in_init != 0);
            }
        } else {
            ProblemDescription<3> pd;
            fill_problem<3>(pd, has_force, force);
            if (with_g) {
                DistributionFunctions G(g, Q, (size_t)n, (size_t)stride);
                distributed_vector temperature(T, (size_t)n), mss(sensor, (size_t)n);
                selectCollision<3>(cfg, pd, F, G, densities, velocities, temperature, mss, owned, viscosity, dt, *st, in_init != 0);
            } else {
                selectCollision<3>(cfg, pd, F, densities, velocities, owned, viscosity, dt, *st, in_init != 0);
            }
        }
    }, err, errlen);
}

// Legacy CollisionModel family: model = "KBC_STANDARD" | "MRT_ENTROPIC" | "MRT_STANDARD" | "BGK_STANDARD".
// The relaxation parameter is CollisionModel::calculateRelaxationParameter(viscosity, dt, stencil)
// (CollisionModel.h:152-157), as CFDSolver's constructor computes it.
int ref_legacy_collide(const char* stencil_name, double scaling, const char* model, double viscosity, double dt,
                       int in_init, int64_t n, int64_t stride, double* f, double* rho, double* u, char* err, int errlen)
{
    boost::shared_ptr<Stencil> st = make_stencil(stencil_name, scaling);
    if (!st) { put_err(err, errlen, "unknown stencil"); return -8; }
    const size_t Q = st->getQ(), D = st->getD();
    DistributionFunctions F(f, Q, (size_t)n, (size_t)stride);
    distributed_vector densities(rho, (size_t)n);
    vector<distributed_vector> velocities;
    for (size_t j = 0; j < D; j++) velocities.emplace_back(u + j * (size_t)n, (size_t)n);
    const dealii::IndexSet owned((size_t)n);
    const double tau = CollisionModel::calculateRelaxationParameter(viscosity, dt, *st);
    const std::string m(model);
    return guarded([&]() {
        boost::shared_ptr<CollisionModel> cm;
        if (m == "KBC_STANDARD") cm = boost::make_shared<KBCStandard>(tau, dt, st);
        else if (m == "MRT_ENTROPIC") cm = boost::make_shared<MRTEntropic>(tau, dt, st);
        else if (m == "MRT_STANDARD") cm = boost::make_shared<MRTStandard>(tau, dt, st);
        else if (m == "BGK_STANDARD") cm = boost::make_shared<BGKStandard>(tau, dt, st);
        else throw NATriuMException("unknown legacy model");
        cm->collideAll(F, densities, velocities, owned, in_init != 0);
    }, err, errlen);
}

// BGKStandard::getEquilibriumDistribution (L/collision/BGKStandard.cpp:23-41): the formula CFDSolver's initialisation uses
int ref_legacy_feq(const char* stencil_name, double scaling, double rho, const double* u_scaled, double* feq)
{
    boost::shared_ptr<Stencil> st = make_stencil(stencil_name, scaling);
    if (!st) return -8;
    BGKStandard bgk(1.0, 1.0, st);
    numeric_vector u(st->getD());
    for (size_t j = 0; j < st->getD(); j++) u(j) = u_scaled[j];
    for (size_t i = 0; i < st->getQ(); i++) feq[i] = bgk.getEquilibriumDistribution(i, u, rho);
    return 0;
}

// make_M / make_T / make_diag (AuxiliaryMRTFunctions.cpp); M, T: [Q][Q], omega: [Q]
int ref_mrt_tables(int Q, int basis, int relax_mode, double tau, double* M, double* T, double* omega, char* err, int errlen)
{
    return guarded([&]() {
        if (Q == 9) {
            const auto m = AuxiliaryMRTFunctions::make_M<9>((MomentBasis)basis);
            const auto t = AuxiliaryMRTFunctions::make_T<9>((MomentBasis)basis);
            const auto d = AuxiliaryMRTFunctions::make_diag<9>(tau, (MomentBasis)basis, (RelaxMode)relax_mode);
            for (int i = 0; i < 9; i++) { omega[i] = d[i]; for (int j = 0; j < 9; j++) { M[i * 9 + j] = m[i][j]; T[i * 9 + j] = t[i][j]; } }
        } else if (Q == 19) {
            const auto m = AuxiliaryMRTFunctions::make_M<19>((MomentBasis)basis);
            const auto t = AuxiliaryMRTFunctions::make_T<19>((MomentBasis)basis);
            const auto d = AuxiliaryMRTFunctions::make_diag<19>(tau, (MomentBasis)basis, (RelaxMode)relax_mode);
            for (int i = 0; i < 19; i++) { omega[i] = d[i]; for (int j = 0; j < 19; j++) { M[i * 19 + j] = m[i][j]; T[i * 19 + j] = t[i][j]; } }
        } else {
            throw NATriuMException("no MRT tables for this Q");
        }
    }, err, errlen);
}

// ThermalBounceBack<3>::calculateBoundaryValues (L/boundaries/ThermalBounceBack.cpp:50-109) cannot be compiled here
// (the class sits on FEBoundaryValues / deal.II), but its body is a sequence of calls into the collision_advanced
// headers.  This entry point makes the same calls in the same order on one destination DoF (f, g: [45]), so the
// arithmetic -- density, velocity, temperature, QuarticEquilibrium::polynomial, calculateGeqFromFeq -- is the
// reference's own; only the eight-line call sequence is restated.  Returns 1 if the DoF was re-equilibrated.
int ref_thermal_wall_point(double scaling, double wall_temperature, double* f, double* g)
{
    boost::shared_ptr<Stencil> st = make_stencil("D3Q45", scaling);
    const Stencil& stencil = *st;
    const double cs2 = stencil.getSpeedOfSoundSquare() / (scaling * scaling);
    const double gamma = 1.4;
    std::array<double, 45> f_destination, g_destination, feq, geq, w;
    for (int i = 0; i < 45; i++) { f_destination[i] = f[i]; g_destination[i] = g[i]; w[i] = stencil.getWeight(i); }
    const double rho = calculateDensity<45>(f_destination);
    std::array<double, 3> u_local;
    std::array<std::array<double, 3>, 45> e = getParticleVelocitiesWithoutScaling<3, 45>(stencil);
    calculateVelocity<3, 45>(f_destination, u_local, rho, e);
    const double T_local = calculateTemperature<3, 45>(f_destination, g_destination, u_local, rho, e, cs2, gamma);
    if (std::abs(T_local - wall_temperature) > 0.00001) {
        QuarticEquilibrium<3, 45> eq(cs2, e);
        eq.polynomial(feq, rho, u_local, T_local, e, w, cs2);
        calculateGeqFromFeq<3, 45>(feq, geq, T_local, gamma);
        for (int i = 0; i < 45; i++) { f_destination[i] -= feq[i]; g_destination[i] -= geq[i]; }
        eq.polynomial(feq, rho, u_local, wall_temperature, e, w, cs2);
        calculateGeqFromFeq<3, 45>(feq, geq, wall_temperature, gamma);
        for (int i = 0; i < 45; i++) { f[i] = f_destination[i] + feq[i]; g[i] = geq[i]; }
        return 1;
    }
    return 0;
}

}  // extern "C"
