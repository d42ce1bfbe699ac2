// ref_walls_driver.cpp -- C entry point over the REFERENCE's own ThermalBounceBack (test infrastructure, part of oracle/).
//
// Compiled by oracle/Makefile.ref together with L/boundaries/ThermalBounceBack.cpp and L/stencils/{Stencil,D3Q45}.cpp where they
// lie under /root/reference (scratch tree of symbolic links; stand-ins of oracle/ref_stubs/ and oracle/ref_stubs_walls/) into
// oracle/_ref/libnatrium_ref_walls.so.  Nothing of the reference is restated here: the hits are replayed in list order as
// SemiLagrangianBoundaryHandler::apply does (L/boundaries/SemiLagrangianBoundaryHandler.cpp:43-109), each through
//   natrium::ThermalBounceBack<3>::calculateBoundaryValues      L/boundaries/ThermalBounceBack.cpp:50-109
// Used by oracle/ref.py -> tests/test_oracle_vs_ref.py (orc_apply_wall_hits, thermal kind == reference code).
#include <cstdint>

#include "boundaries/ThermalBounceBack.h"
#include "stencils/D3Q45.h"

using namespace natrium;

// f, g: [45][stride] (the just-streamed f and the not yet streamed g); dest_index / dest_direction: the hit list in order
extern "C" int ref_thermal_bounce_back(double scaling, int64_t n, int64_t stride, double* f, double* g, int64_t n_hits,
                                       const int32_t* dest_index, const int32_t* dest_direction, double wall_temperature)
{
    D3Q45 stencil(scaling);
    DistributionFunctions fnew(f, 45, (size_t)n, (size_t)stride), gg(g, 45, (size_t)n, (size_t)stride);
    GlobalBoundaryData data(fnew, fnew, gg, stencil, 0.0, 0.0);
    FEBoundaryValues<3> fe(data);
    dealii::Tensor<1, 3> u_wall;
    for (int i = 0; i < 3; i++) u_wall[i] = 0.0;
    ThermalBounceBack<3> wall(0, u_wall, wall_temperature);
    for (int64_t h = 0; h < n_hits; h++) {
        LagrangianPathDestination dest((size_t)dest_index[h], (size_t)dest_direction[h]);
        wall.calculateBoundaryValues(fe, 0, dest, 0.0, 0.0);
    }
    return 0;
}
