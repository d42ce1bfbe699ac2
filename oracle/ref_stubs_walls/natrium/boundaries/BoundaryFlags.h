// Stand-in for L/boundaries/BoundaryFlags.h (test infrastructure): the flag type ThermalBounceBack::getUpdateFlags returns.
#pragma once
namespace natrium {
enum BoundaryFlags { only_distributions = 0, boundary_rho = 1, boundary_u = 2, boundary_drho_dt = 4, boundary_du_dt = 8, boundary_p = 16 };
}
