"""Shared problem builders for the parity tests: the same synthetic case is built twice,
once for the oracle (oracle/) and once for the product (natrium_b200/), from the same parameters."""
import math

import numpy as np

from natrium_b200 import harness
from natrium_b200.stencils import Stencil

CASES = {
    # name: dim, cells, p, stencil, Ma (-> scaling), viscosity, cfl, with_g
    "c1_tgv2d_d2q9": dict(dim=2, cells=8, p=4, stencil="D2Q9", scaling=math.sqrt(3) / 0.05, nu=1.0, cfl=0.4),
    "tgv2d_small": dict(dim=2, cells=4, p=3, stencil="D2Q9", scaling=math.sqrt(3) / 0.05, nu=1.0, cfl=0.4),
    "tgv3d_d3q19_small": dict(dim=3, cells=3, p=4, stencil="D3Q19", scaling=math.sqrt(3) / 0.05, nu=2 * math.pi, cfl=0.4),
    "tgv3d_d3q19_p2": dict(dim=3, cells=4, p=2, stencil="D3Q19", scaling=math.sqrt(3) / 0.05, nu=2 * math.pi, cfl=0.4),
    "tgv3d_d3q15": dict(dim=3, cells=3, p=2, stencil="D3Q15", scaling=math.sqrt(3) / 0.05, nu=2 * math.pi, cfl=0.4),
    "tgv2d_d2q25": dict(dim=2, cells=6, p=2, stencil="D2Q25H", scaling=1.0, nu=0.01, cfl=1.0, with_g=True),
    "tgv3d_d3q45": dict(dim=3, cells=2, p=3, stencil="D3Q45", scaling=1.0, nu=0.01, cfl=0.4, with_g=True),
}


def rel_err(a, b):
    """max elementwise relative error, relative to max(|b|) per array row to avoid 0/0 on exact zeros."""
    a, b = np.asarray(a), np.asarray(b)
    scale = np.maximum(np.abs(b), 1e-300)
    return float(np.max(np.abs(a - b) / scale))


def product_problem(case):
    c = CASES[case]
    st = Stencil(c["stencil"], c["scaling"])
    pb = harness.CartesianProblem(c["dim"], c["cells"], c["p"])
    dt = pb.timestep(st, c["cfl"])
    return c, st, pb, dt


def initial_fields(c, st, x):
    """(rho, u, T) of the Taylor-Green case at the given points (physical units)."""
    cs = st.getSpeedOfSound()
    if c["dim"] == 2:
        rho, u = harness.taylor_green_2d(x)
        if c.get("with_g"):
            rho = 1.0 + 0.05 * np.cos(x[:, 0]) * np.sin(x[:, 1])
            u = 0.1 * u
    else:
        if c.get("with_g"):
            rho, u = harness.taylor_green_3d(x, cs, compressible=True, density_numerator=0.1)
            u = 0.1 * u
        else:
            rho, u = harness.taylor_green_3d(x, cs)
    T = 1.0 + 0.02 * np.sin(x[:, 0]) * np.cos(x[:, -1]) if c.get("with_g") else np.ones(x.shape[0])
    return rho, u, T


def oracle_problem(case):
    """Builds mesh, matrix blocks and initial populations with the oracle only."""
    from oracle import assembly, cpu, fields
    c = CASES[case]
    st = cpu.Stencil(c["stencil"], c["scaling"])
    mesh = assembly.CartesianMesh.uniform(c["dim"], c["cells"])
    dt = assembly.calculate_timestep(mesh, c["p"], st.max_speed, c["cfl"])
    blocks, dofs = assembly.assemble_semilagrangian(mesh, c["p"], st.e, dt)
    x = dofs.support_points()
    pst = Stencil(c["stencil"], c["scaling"])
    rho, u, T = initial_fields(c, pst, x)
    if c.get("with_g"):
        f, g = fields.quartic_equilibrium_init(st.e, st.w, st.cs2, st.scaling, rho, u, T, 1.4)
    else:
        f, g = fields.equilibrium_init(st.e, st.w, st.cs2, rho, u), None
    return dict(c=c, st=st, mesh=mesh, dt=dt, blocks=blocks, dofs=dofs, x=x, f=f, g=g, rho=rho, u=u, T=T)


def poiseuille_problem(oracle_lib):
    """SemiLagrangianBoundaryHandler_PoiseuilleBB_test (test/boundaries/SemiLagrangianBoundaryHandler_test.cpp:175-229):
    PoiseuilleFlow2D (L/benchmarks/PoiseuilleFlow2D.cpp:21-95) on [0,2]x[0,1], refinement 2 (4x4 cells), FE order 2,
    D2Q9 with scaling sqrt(3)*1.5*u_bulk/Ma, CFL 1.5, Re 10, periodic in x, VelocityNeqBounceBack(zero velocity)
    walls in y, constant force Fx = 8 u_max nu / h^2 with SHIFTING_VELOCITY forcing, start at rest."""
    import math
    from oracle import assembly, fields
    u_bulk, height, length, Re, Ma, p, cfl = 0.0001 / 1.5, 1.0, 2.0, 10.0, 0.1, 2, 1.5
    nu = u_bulk * height / Re
    scaling = math.sqrt(3) * 1.5 * u_bulk / Ma
    st = oracle_lib.Stencil("D2Q9", scaling)
    mesh = assembly.CartesianMesh([np.linspace(0, length, 5), np.linspace(0, height, 5)], boundary=["periodic", "wall"])
    dt = assembly.calculate_timestep(mesh, p, st.max_speed, cfl)
    opp = np.array([int(np.argmin(np.abs(st.e + st.e[i]).sum(1))) for i in range(9)])
    blocks, dofs = assembly.assemble_semilagrangian(mesh, p, st.e, dt, opposite=opp)
    F = np.array([8 * (1.5 * u_bulk) * nu / (height * height), 0.0])
    hits = dofs.hits
    idx = np.array([h["index"] for h in hits], dtype=np.int32)
    dirs = np.array([h["direction"] for h in hits], dtype=np.int32)
    kinds = np.zeros(len(hits), dtype=np.int32)
    vals = np.zeros(len(hits))                                   # zero wall velocity: 2 w rho e.u_w / cs2 = 0
    f0 = fields.equilibrium_init(st.e, st.w, st.cs2, np.ones(dofs.N), np.zeros((2, dofs.N)))
    return dict(st=st, scaling=scaling, nu=nu, dt=dt, blocks=blocks, dofs=dofs, mesh=mesh, F=F, hits=(idx, dirs, kinds, vals), f0=f0, u_bulk=u_bulk)


def integral_mean(dofs, mesh, p, values):
    """PhysicalProperties<dim>::meanVelocityX (L/solver/PhysicalProperties.cpp:323-365): Gauss-Lobatto quadrature of a
    nodal field over all cells divided by the domain area (quadrature points = support points)."""
    from numpy.polynomial import legendre
    from oracle import assembly
    nodes = np.asarray(assembly.gll_nodes(p))
    Pp = legendre.legval(2 * nodes - 1, [0] * p + [1])
    w1 = 1.0 / (p * (p + 1) * Pp ** 2)                    # GLL weights on [0, 1]
    total, area = 0.0, 0.0
    for cell in mesh.cells():
        h = [mesh.verts[d][cell[d] + 1] - mesh.verts[d][cell[d]] for d in range(mesh.dim)]
        vol = float(np.prod(h))
        wq = w1
        for _ in range(mesh.dim - 1):
            wq = np.multiply.outer(w1, wq)
        # cell_dofs is lexicographic with x fastest; outer products above put the LAST axis fastest -> transpose order is
        # irrelevant because the weights are symmetric under axis permutation
        ids = dofs.cell_dofs(cell)
        total += vol * float(np.sum(wq.reshape(-1) * values[ids]))
        area += vol
    return total / area


def moving_walls_problem(oracle_lib):
    """WallTest fixture of the reference (test/boundaries/WallFixture.h:44-88, test/problemdescription/WallTestDomain2D.h):
    unit square, refinement 1 (2x2 cells), FE order 2, D2Q9 (scaling 1), CFL 1, viscosity 0.2, periodic in x, both
    y-walls VelocityNeqBounceBack moving with u_w = (0.01, 0); start at rest, 100 steps."""
    from oracle import assembly, fields
    st = oracle_lib.Stencil("D2Q9", 1.0)
    mesh = assembly.CartesianMesh([np.linspace(0, 1.0, 3), np.linspace(0, 1.0, 3)], boundary=["periodic", "wall"])
    p, nu = 2, 0.2
    dt = assembly.calculate_timestep(mesh, p, st.max_speed, 1.0)
    opp = np.array([int(np.argmin(np.abs(st.e + st.e[i]).sum(1))) for i in range(9)])
    blocks, dofs = assembly.assemble_semilagrangian(mesh, p, st.e, dt, opposite=opp)
    u_w = np.array([0.01, 0.0])
    hits = dofs.hits
    idx = np.array([h["index"] for h in hits], dtype=np.int32)
    dirs = np.array([h["direction"] for h in hits], dtype=np.int32)
    kinds = np.zeros(len(hits), dtype=np.int32)
    # VelocityNeqBounceBack::calculateBoundaryValues (VelocityNeqBounceBack.cpp:137-195): 2 w rho (e . u_w) / cs2, rho = 1
    vals = np.array([2 * st.w[h["direction"]] * 1.0 * float(st.e[h["direction"]] @ u_w) / st.cs2 for h in hits])
    f0 = fields.equilibrium_init(st.e, st.w, st.cs2, np.ones(dofs.N), np.zeros((2, dofs.N)))
    return dict(st=st, nu=nu, dt=dt, blocks=blocks, dofs=dofs, mesh=mesh, hits=(idx, dirs, kinds, vals), f0=f0)
