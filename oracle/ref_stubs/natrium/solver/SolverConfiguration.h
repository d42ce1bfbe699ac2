// Stand-in for L/solver/SolverConfiguration.h (test infrastructure, see ../../README.md): the getters the collision
// path reads (GeneralCollisionData ctor Aux...h:150-200, relaxWithG CollisionSchemes.h:63-103, MultipleRelaxationTime
// :224-234, selectCollision's dispatch macros CollisionSelection.h:26-55), backed by plain members.
#pragma once
#include "../utilities/BasicNames.h"
#include "../utilities/ConfigNames.h"
#include "../stencils/Stencil.h"
#include "../problemdescription/ProblemDescription.h"
namespace natrium {
class SolverConfiguration {
public:
    StencilType stencil = Stencil_D2Q9;
    CollisionSchemeName collision = BGK_STANDARD;
    EquilibriumSchemeName equilibrium = BGK_EQUILIBRIUM;
    ForceType forcing = NO_FORCING;
    MomentBasis mrt_basis = DELLAR_D2Q9;
    RelaxMode mrt_relax = RELAX_FULL;
    double gamma = 1.4, prandtl = 1.0, steady_gamma = 1.0;
    bool prandtl_set = false, sutherland_set = false;

    StencilType getStencil() const { return stencil; }
    CollisionSchemeName getCollisionScheme() const { return collision; }
    EquilibriumSchemeName getEquilibriumScheme() const { return equilibrium; }
    ForceType getForcingScheme() const { return forcing; }
    MomentBasis getMRTBasis() const { return mrt_basis; }
    RelaxMode getMRTRelaxationTimes() const { return mrt_relax; }
    double getHeatCapacityRatioGamma() const { return gamma; }
    double getPrandtlNumber() const { return prandtl; }
    bool isPrandtlNumberSet() const { return prandtl_set; }
    bool isSutherlandLawSet() const { return sutherland_set; }
    double getBGKSteadyStateGamma() const { return steady_gamma; }
};
}  // namespace natrium
