// Stand-in for L/problemdescription/ProblemDescription.h (test infrastructure, see ../../README.md): what the collision
// reads from it -- hasExternalForce() and getExternalForce()->getForce() (AuxiliaryCollisionFunctions.h:179-196).
#pragma once
#include "../utilities/BasicNames.h"
#include "ConstantExternalForce.h"
namespace natrium {
template <size_t dim>
class ProblemDescription {
    boost::shared_ptr<ConstantExternalForce<dim>> m_externalForce;
public:
    bool hasExternalForce() const { return (bool)m_externalForce; }
    const boost::shared_ptr<ConstantExternalForce<dim>>& getExternalForce() const { return m_externalForce; }
    void setExternalForce(boost::shared_ptr<ConstantExternalForce<dim>> f) { m_externalForce = f; }
};
}  // namespace natrium
