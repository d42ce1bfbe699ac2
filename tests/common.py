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
