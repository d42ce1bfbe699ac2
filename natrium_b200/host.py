"""Host-side mirror of the reference interfaces the hot path sits behind.

Same names, argument meaning and error behaviour as the reference (L = src/library/natrium):

  DistributionFunctions   L/solver/DistributionFunctions.h:47-300
  SemiLagrangian          L/advection/AdvectionOperator.h:118-173,190 + SemiLagrangian.h:150-179
  selectCollision         L/collision_advanced/CollisionSelection.h:60-67,115-122
  CFDSolver               L/solver/CFDSolver.cpp:659-754 (stream), 807-843 (collide), 877-902 (run)
  CompressibleCFDSolver   L/solver/CompressibleCFDSolver.h:181-314,739-788,1006-1051

Storage and arithmetic live in libnatrium_b200 (CUDA); these classes only hold a context
handle and forward.  The mesh side (ProblemDescription) is the synthetic harness here; in a
NATriuM build it is deal.II (INTEGRATION.md shows the C++ shim against the real headers).
"""
import numpy as np

from . import _capi, harness, mrt
from ._capi import CollisionException  # noqa: F401  (re-exported like natrium::CollisionException)
from .stencils import Stencil

BGK_STANDARD, KBC_STANDARD, MRT_ENTROPIC = "BGK_STANDARD", "KBC_STANDARD", "MRT_ENTROPIC"
BGK_REGULARIZED, MRT_STANDARD = "BGK_REGULARIZED", "MRT_STANDARD"
BGK_EQUILIBRIUM, QUARTIC_EQUILIBRIUM = "BGK_EQUILIBRIUM", "QUARTIC_EQUILIBRIUM"
NO_FORCING, SHIFTING_VELOCITY, EXACT_DIFFERENCE, GUO = _capi.NO_FORCING, _capi.SHIFTING_VELOCITY, _capi.EXACT_DIFFERENCE, _capi.GUO
DELLAR_D2Q9, LALLEMAND_D2Q9, DHUMIERES_D3Q19 = mrt.DELLAR_D2Q9, mrt.LALLEMAND_D2Q9, mrt.DHUMIERES_D3Q19
RELAX_FULL, DELLAR_RELAX_ONLY_N, RELAX_DHUMIERES_PAPER = mrt.RELAX_FULL, mrt.DELLAR_RELAX_ONLY_N, mrt.RELAX_DHUMIERES_PAPER
_SCHEMES = {BGK_STANDARD: _capi.BGK_STANDARD, KBC_STANDARD: _capi.KBC_STANDARD, MRT_ENTROPIC: _capi.MRT_ENTROPIC,
            BGK_REGULARIZED: _capi.BGK_REGULARIZED, MRT_STANDARD: _capi.MRT_STANDARD}
_EQUILIBRIA = {BGK_EQUILIBRIUM: _capi.BGK_EQUILIBRIUM, QUARTIC_EQUILIBRIUM: _capi.QUARTIC_EQUILIBRIUM}


class SolverConfiguration:
    """The getters GeneralCollisionData / CFDSolver read on the path (defaults of
    L/solver/SolverConfiguration.cpp:17-293: CFL 0.4, D2Q9, scaling 1, BGK standard, BGK eq, gamma 1.4, Pr 1)."""

    def __init__(self):
        self._stencil, self._scaling, self._cfl, self._p = "Stencil_D2Q9", 1.0, 0.4, 1
        self._collision, self._equilibrium = BGK_STANDARD, BGK_EQUILIBRIUM
        self._gamma, self._prandtl, self._prandtl_set, self._sutherland = 1.4, 1.0, False, False
        self._n_steps = 0
        # "MRT basis" = Dellar D2Q9, "MRT relaxation times" = Full, forcing off (SolverConfiguration.cpp:133-141)
        self._mrt_basis, self._mrt_relax, self._forcing = DELLAR_D2Q9, RELAX_FULL, NO_FORCING

    def setMRTBasis(self, b): self._mrt_basis = b
    def getMRTBasis(self): return self._mrt_basis
    def setMRTRelaxationTimes(self, r): self._mrt_relax = r
    def getMRTRelaxationTimes(self): return self._mrt_relax
    def setForcingScheme(self, f): self._forcing = f
    def getForcingScheme(self): return self._forcing

    def setStencil(self, s): self._stencil = s if s.startswith("Stencil_") else "Stencil_" + s
    def getStencil(self): return self._stencil
    def setStencilScaling(self, s): self._scaling = float(s)
    def getStencilScaling(self): return self._scaling
    def setCFL(self, c): self._cfl = float(c)
    def getCFL(self): return self._cfl
    def setSedgOrderOfFiniteElement(self, p): self._p = int(p)
    def getSedgOrderOfFiniteElement(self): return self._p
    def setCollisionScheme(self, c): self._collision = c
    def getCollisionScheme(self): return self._collision
    def setEquilibriumScheme(self, e): self._equilibrium = e
    def getEquilibriumScheme(self): return self._equilibrium
    def setHeatCapacityRatioGamma(self, g): self._gamma = float(g)
    def getHeatCapacityRatioGamma(self): return self._gamma
    def setPrandtlNumber(self, pr): self._prandtl, self._prandtl_set = float(pr), True
    def getPrandtlNumber(self): return self._prandtl
    def isPrandtlNumberSet(self): return self._prandtl_set
    def setSutherlandLaw(self, on=True): self._sutherland = bool(on)
    def isSutherlandLawSet(self): return self._sutherland
    def setNumberOfTimeSteps(self, n): self._n_steps = int(n)
    def getNumberOfTimeSteps(self): return self._n_steps


class DistributionFunctions:
    """Device-resident container: Q direction-major arrays of n_owned (+ghost) doubles.
    ``which`` 0 = f, 1 = g of the owning context."""

    def __init__(self, ctx, which=0):
        self._ctx, self._which = ctx, which

    def getQ(self): return self._ctx.Q
    def size(self): return self._ctx.Q
    def n_owned(self): return self._ctx.n_owned

    def at(self, i):
        """Host copy of f.at(i) (owned entries) -- like ExtractView, but a download."""
        return self._ctx.download_population(self._which, i)

    def set(self, i, values):
        self._ctx.upload_population(self._which, i, values)

    def getFStream(self):
        return np.stack([self.at(i) for i in range(1, self.getQ())])

    def to_host(self, out=None):
        return self._ctx.download_populations(self._which, out)

    def from_host(self, arr):
        self._ctx.upload_populations(self._which, arr)

    def updateGhosted(self):
        self._ctx.update_ghosted()


class SemiLagrangian:
    """AdvectionOperator for semi-Lagrangian streaming, device backed."""

    def __init__(self, problem, orderOfFiniteElement, stencil, delta_t=0.0, ctx=None, rank=0, nranks=1):
        assert orderOfFiniteElement == problem.p
        self._problem, self._stencil, self._dt = problem, stencil, float(delta_t)
        self._ctx, self._rank, self._nranks = ctx, rank, nranks
        self._part = None
        self._nnz = 0

    # -- setup (host, once)
    def setupDoFs(self):
        pass   # DoFs of the synthetic problem are implicit in the Cartesian grid

    def setDeltaT(self, delta_t):
        """updateSparsityPattern() in the reference; here the partition/ghost plan depends on dt."""
        self._dt = float(delta_t)
        self._part = harness.SlabPartition(self._problem, self._stencil, self._dt, self._rank, self._nranks)

    def getDeltaT(self): return self._dt
    def getPartition(self): return self._part
    def getLocallyOwnedDofs(self): return self._part.owned_global_ids()
    def getNumberOfDoFs(self): return self._problem.N

    def reassemble(self):
        """fillSparseObject(false) + compress on the host, then the blocks go to the device."""
        ctx = self._ctx
        self._nnz = harness.upload_streaming_matrix(ctx, self._problem, self._part, self._stencil, self._dt)
        if self._nranks > 1:
            ctx.set_halo(*self._part.halo_plan())

    def getSystemMatrix(self):
        return self._ctx.matrix_info()

    # -- per step
    def stream(self, f_old, f, t):
        """f_old = f; f.FStream = M f_old.FStream; boundary handler (no hits for periodic); returns dt."""
        self._ctx.stream(f._which)
        return self._dt

    def applyBoundaryConditions(self, f_old, f, t):
        pass   # periodic problems: SemiLagrangianBoundaryHandler has no hits


def _apply_collision_setup(ctx, configuration, problemDescription, stencil, viscosity, delta_t, with_g, in_init=False):
    """What GeneralCollisionData + SpecificCollisionData pull out of the configuration, the problem and the stencil
    (AuxiliaryCollisionFunctions.h:151-200, CollisionSchemes.h:209-236), flattened for the C ABI."""
    scheme = configuration.getCollisionScheme()
    if scheme == MRT_STANDARD and stencil.getQ() in (9, 19) and not with_g:
        cs2 = stencil.getSpeedOfSoundSquare()
        tau = viscosity / (delta_t * cs2) + 0.5          # calculateTauFromNu
        basis = configuration.getMRTBasis()
        if (stencil.getQ() == 9) != (basis in (DELLAR_D2Q9, LALLEMAND_D2Q9)):
            raise CollisionException(_capi.NB200_ERR_UNSUPPORTED, f"MRT basis not defined for Q={stencil.getQ()}")
        ctx.set_mrt(mrt.make_M(basis), mrt.make_T(basis), mrt.make_diag(tau, basis, configuration.getMRTRelaxationTimes()))
    has_force = getattr(problemDescription, "hasExternalForce", lambda: False)()
    ctx.set_collision(viscosity, delta_t, scheme=_SCHEMES[scheme],
                      equilibrium=_EQUILIBRIA[configuration.getEquilibriumScheme()], with_g=with_g, in_init=in_init,
                      gamma=configuration.getHeatCapacityRatioGamma(),
                      prandtl=configuration.getPrandtlNumber() if configuration.isPrandtlNumberSet() else None,
                      sutherland=configuration.isSutherlandLawSet(),
                      force=problemDescription.getExternalForce().getForce() if has_force else None,
                      force_type=configuration.getForcingScheme())


def selectCollision(configuration, problemDescription, f, *args):
    """Both reference overloads:
       selectCollision(cfg, pd, f, densities, velocities, owned, viscosity, delta_t, stencil, inInit)
       selectCollision(cfg, pd, f, g, densities, velocities, temperature, maskShockSensor, owned, viscosity, delta_t, stencil, inInit)
    densities / velocities (/temperature/maskShockSensor) are numpy arrays that are overwritten
    (velocities is read when inInitializationProcedure is true)."""
    with_g = isinstance(args[0], DistributionFunctions)
    if with_g:
        g, densities, velocities, temperature, mask, owned, viscosity, delta_t, stencil, in_init = args
    else:
        densities, velocities, owned, viscosity, delta_t, stencil, in_init = args
    ctx = f._ctx
    if configuration.getStencil() != stencil.getStencilType():
        raise CollisionException(_capi.NB200_ERR_UNSUPPORTED, "Severe error: Collision model not implemented yet -- cf. CollisionSelection.h")
    _apply_collision_setup(ctx, configuration, problemDescription, stencil, viscosity, delta_t, with_g, in_init)
    if in_init:
        ctx.upload_velocity(np.asarray(velocities))
    ctx.collide()
    ctx.synchronize()     # surfaces the density exception here, like the reference's throw inside collideAll
    if with_g:
        rho, u, T, s = ctx.download_moments(want_T=True)
        temperature[...] = T
        mask[...] = s
    else:
        rho, u = ctx.download_moments()
    densities[...] = rho
    if not in_init:
        np.asarray(velocities)[...] = u


class PseudoEntropicStabilizer:
    """Mirror of natrium::PseudoEntropicStabilizer<dim> (L/dataprocessors/PseudoEntropicStabilizer.h:25-66): a
    DataProcessor appended to the solver; apply() multiplies the populations of every owned DoF by the stabilizer
    matrix.  On the device the matrix runs after the collision of every step."""

    def __init__(self, solver, with_e=False):
        self.m_solver, self.m_withE = solver, with_e
        name = solver.m_stencil.getStencilType()
        if name not in ("Stencil_D2Q9", "Stencil_D3Q19"):
            raise CollisionException(_capi.NB200_ERR_UNSUPPORTED, "PseudoEntropicStabilizer is only defined for D2Q9 and D3Q19")
        self.matrix = mrt.make_stabilizer(name, with_e and name == "Stencil_D2Q9")

    def apply(self):
        self.m_solver.ctx.set_post_collision_matrix(self.matrix)
        self.m_solver.ctx.apply_post_collision()


class ExponentialFilter:
    """Mirror of natrium::ExponentialFilter<dim> (L/smoothing/ExponentialFilter.h:25-98): the constructor builds, once, the
    projections between the nodal element basis and the tensor Legendre modes and the per-mode damping; applyFilter runs on
    the device (nb200_set_filter / nb200_apply_filter).  The reference integrates with the operator's own Gauss-Lobatto rule
    on the element's own nodes (L/advection/AdvectionOperator.cpp:34-39), for which quad_legendre^-1 quad_source
    (ExponentialFilter.cpp:46-66) is the inverse of the matrix Psi[node][mode] of mode values at the nodes: this mirror forms
    from_legendre = Psi and to_legendre = Psi^-1 directly.  Element-local numbering: lexicographic, x fastest."""

    def __init__(self, alpha, s, Nc, by_sum, p, dim):
        self.m_alpha, self.m_s, self.m_Nc, self.m_bySum, self.m_p, self.dim = float(alpha), float(s), int(Nc), bool(by_sum), int(p), int(dim)
        n1 = p + 1
        x = harness.gauss_lobatto_points(p)
        # dealii::Polynomials::Legendre: orthonormal on [0,1]; three-term recurrence in t = 2x - 1
        t = 2.0 * x - 1.0
        leg = np.zeros((n1, n1))
        leg[:, 0] = 1.0
        if p >= 1:
            leg[:, 1] = t
        for k in range(2, n1):
            leg[:, k] = ((2 * k - 1) * t * leg[:, k - 1] - (k - 1) * leg[:, k - 2]) / k
        leg *= np.sqrt(2.0 * np.arange(n1) + 1.0)[None, :]
        n = n1 ** dim
        psi = np.ones((n, n))
        for i in range(n):                    # node i: lexicographic, x fastest
            node = [(i // n1 ** d) % n1 for d in range(dim)]
            for m in range(n):                # mode m: x slowest (evaluateLegendreND, ExponentialFilter.cpp:70-94)
                mode = [(m // n1 ** (dim - 1 - d)) % n1 for d in range(dim)]
                v = 1.0
                for d in range(dim):
                    v *= leg[node[d], mode[d]]
                psi[i, m] = v
        self.m_projectFromLegendre = np.ascontiguousarray(psi)
        self.m_projectToLegendre = np.ascontiguousarray(np.linalg.inv(psi))
        # makeDegreeVectors (:96-137); in 3-D the reference's iy expression evaluates to i % (p+1) (operator precedence, :123)
        self.m_degreeMax, self.m_degreeSum = np.zeros(n, dtype=np.int64), np.zeros(n, dtype=np.int64)
        for i in range(n):
            if dim == 1:
                idx = (i,)
            elif dim == 2:
                idx = (i // n1, i % n1)
            else:
                idx = (i // (n1 * n1), i % n1, i % n1)
            self.m_degreeMax[i], self.m_degreeSum[i] = max(idx), sum(idx)
        deg = self.m_degreeSum if self.m_bySum else self.m_degreeMax
        max_degree = dim * p if self.m_bySum else p
        self.sigma = np.ones(n)
        hit = deg >= self.m_Nc
        self.sigma[hit] = np.exp(-self.m_alpha * ((deg[hit] + 1.0 - self.m_Nc) / (max_degree + 1.0 - self.m_Nc)) ** self.m_s)

    def getProjectToLegendre(self): return self.m_projectToLegendre
    def getProjectFromLegendre(self): return self.m_projectFromLegendre

    def attach(self, ctx, cell_dofs, interval=1):
        """Hands the tables to the device library (cell_dofs: cell->get_dof_indices per locally owned cell, in loop order)."""
        ctx.set_filter(cell_dofs, self.m_projectToLegendre, self.m_projectFromLegendre, self.sigma, interval)


class CFDSolver:
    """Time loop owner.  ``run()`` keeps everything on the device (one fused kernel per step);
    ``stream()`` / ``collide()`` are the reference-ordered single operators."""

    def __init__(self, configuration, problem, viscosity, device=0, rank=0, nranks=1, unique_id=None, with_g=False):
        self.m_configuration, self.m_problem, self.m_viscosity = configuration, problem, float(viscosity)
        self.m_stencil = Stencil(configuration.getStencil(), configuration.getStencilScaling())
        self.ctx = _capi.Context(device, rank, nranks, unique_id)
        st = self.m_stencil
        self.ctx.set_stencil(st.getDirections(), st.getWeights(), st.getScaling(), st.getSpeedOfSoundSquare())
        self.m_advectionOperator = SemiLagrangian(problem, configuration.getSedgOrderOfFiniteElement(), st, 0.0,
                                                  self.ctx, rank, nranks)
        self.m_advectionOperator.setupDoFs()
        dt = problem.timestep(st, configuration.getCFL())
        self.m_advectionOperator.setDeltaT(dt)
        part = self.m_advectionOperator.getPartition()
        self.ctx.set_layout(part.n_owned, part.n_ghost, with_g)
        self.ctx.set_dof_order(part.cell_blocked_order())     # cell-by-cell, the order fillSparseObject walks
        self.m_advectionOperator.reassemble()
        self.m_f = DistributionFunctions(self.ctx, 0)
        self.m_time, self.m_i = 0.0, 0
        n = part.n_owned
        self.m_density = np.ones(n)
        self.m_velocity = np.zeros((st.getD(), n))
        self._with_g = with_g

    def getTimeStepSize(self): return self.m_advectionOperator.getDeltaT()
    def getStencil(self): return self.m_stencil
    def getAdvectionOperator(self): return self.m_advectionOperator
    def getNumberOfDoFs(self): return self.m_problem.N
    def getF(self): return self.m_f
    def getDensity(self): return self.m_density
    def getVelocity(self): return self.m_velocity

    def setInitialFields(self, rho, u):
        """initializeDistributions(): f = f_eq(rho0, u0) on the owned DoFs."""
        self.m_density[...] = rho
        self.m_velocity[...] = u
        self.m_f.from_host(harness.equilibrium_distributions(self.m_stencil, rho, u))

    def _configure_collision(self):
        cfg = self.m_configuration
        _apply_collision_setup(self.ctx, cfg, self.m_problem, self.m_stencil, self.m_viscosity, self.getTimeStepSize(), self._with_g)

    def appendDataProcessor(self, proc):
        """CFDSolver::appendDataProcessor (CFDSolver.h): processors run after collide in every iteration of run()."""
        self.m_dataProcessors = getattr(self, "m_dataProcessors", []) + [proc]
        if isinstance(proc, PseudoEntropicStabilizer):
            self.ctx.set_post_collision_matrix(proc.matrix)      # device-resident run(): part of nb200_step

    def setFilter(self, flt, interval=1):
        """m_filter + getFilterInterval() (CFDSolver.cpp:448-476): the device applies it inside run() between stream and collide."""
        self.m_filter, self.m_filterInterval = flt, int(interval)
        part = self.m_advectionOperator.getPartition()
        flt.attach(self.ctx, part.cell_dofs(), interval)

    def filter(self):
        """CFDSolver::filter (CFDSolver.cpp:859-874) as a single operator: every population of f."""
        if getattr(self, "m_filter", None) is not None and self.m_i % self.m_filterInterval == 0:
            self.ctx.apply_filter(0)

    def stream(self):
        self.m_advectionOperator.stream(self.m_f, self.m_f, self.m_time)
        self.m_time += self.getTimeStepSize()

    def collide(self):
        selectCollision(self.m_configuration, self.m_problem, self.m_f, self.m_density, self.m_velocity,
                        None, self.m_viscosity, self.getTimeStepSize(), self.m_stencil, False)

    def run(self, n_steps=None):
        """collide(); then n x (stream; collide) -- fused on the device."""
        n = self.m_configuration.getNumberOfTimeSteps() if n_steps is None else n_steps
        self.collide()
        self._configure_collision()
        self.ctx.step(n)
        self.ctx.synchronize()
        self.m_i += n
        self.m_time += n * self.getTimeStepSize()
        self.m_density[...], self.m_velocity[...] = self.ctx.download_moments()[:2]


class CompressibleCFDSolver(CFDSolver):
    """f + g solver (second distribution carries the internal energy)."""

    def __init__(self, configuration, problem, viscosity, **kw):
        super().__init__(configuration, problem, viscosity, with_g=True, **kw)
        self.m_g = DistributionFunctions(self.ctx, 1)
        n = self.ctx.n_owned
        self.m_temperature = np.ones(n)
        self.m_maskShockSensor = np.zeros(n)

    def setInitialFields(self, rho, u, T):
        self.m_density[...], self.m_velocity[...], self.m_temperature[...] = rho, u, T
        f, g = harness.quartic_equilibrium_distributions(self.m_stencil, rho, u, T,
                                                         self.m_configuration.getHeatCapacityRatioGamma())
        self.m_f.from_host(f)
        self.m_g.from_host(g)

    def gStream(self):
        self.m_advectionOperator.stream(self.m_g, self.m_g, self.m_time)

    def collide(self):
        selectCollision(self.m_configuration, self.m_problem, self.m_f, self.m_g, self.m_density, self.m_velocity,
                        self.m_temperature, self.m_maskShockSensor, None, self.m_viscosity, self.getTimeStepSize(),
                        self.m_stencil, False)

    def run(self, n_steps=None):
        n = self.m_configuration.getNumberOfTimeSteps() if n_steps is None else n_steps
        self.collide()
        self._configure_collision()
        self.ctx.step(n)
        self.ctx.synchronize()
        self.m_i += n
        self.m_time += n * self.getTimeStepSize()
        rho, u, T, s = self.ctx.download_moments(want_T=True)
        self.m_density[...], self.m_velocity[...], self.m_temperature[...], self.m_maskShockSensor[...] = rho, u, T, s
