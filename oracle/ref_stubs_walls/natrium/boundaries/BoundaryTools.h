// Stand-in for L/boundaries/BoundaryTools.h as far as L/boundaries/ThermalBounceBack.{h,cpp} needs it (test infrastructure,
// see oracle/ref_stubs/README.md): BoundaryVelocity, the GlobalBoundaryData container (L/boundaries/BoundaryTools.h:235-290,
// same members and constructors), and the little of deal.II the two touch.
#pragma once
#include <memory>
#include <vector>

#include "../utilities/BasicNames.h"
#include "../utilities/NATriuMException.h"
#include "../solver/DistributionFunctions.h"
#include "../stencils/Stencil.h"
#include "../advection/SemiLagrangianTools.h"

namespace dealii {
template <int dim> class Point { double c_[dim]; public: Point() { for (int i = 0; i < dim; i++) c_[i] = 0; } double& operator()(unsigned i) { return c_[i]; } double operator()(unsigned i) const { return c_[i]; } };
template <typename number> class Vector { std::vector<number> v_; public: Vector() = default; explicit Vector(size_t n) : v_(n, 0) {} size_t size() const { return v_.size(); } number& operator()(size_t i) { return v_[i]; } const number& operator()(size_t i) const { return v_[i]; } };
template <int dim> class Function {
public:
    virtual ~Function() {}
    virtual void set_time(double) {}
    virtual void vector_value(const Point<dim>&, Vector<double>& values) const { for (size_t i = 0; i < values.size(); i++) values(i) = 0.0; }
};
}  // namespace dealii

namespace natrium {
namespace BoundaryTools {
template <size_t dim>
class BoundaryVelocity : public dealii::Function<dim> {
    dealii::Vector<double> m_Velocity;
public:
    BoundaryVelocity(const dealii::Vector<double>& velocity) : m_Velocity(velocity) {}
    BoundaryVelocity(const dealii::Tensor<1, dim>& velocity) : m_Velocity(dim) { for (size_t i = 0; i < dim; i++) m_Velocity(i) = velocity[i]; }
    virtual void vector_value(const dealii::Point<dim>&, dealii::Vector<double>& values) const { for (size_t i = 0; i < dim; i++) values(i) = m_Velocity(i); }
};
}  // namespace BoundaryTools

struct GlobalBoundaryData {
    const DistributionFunctions& m_fold;
    DistributionFunctions& m_fnew;
    DistributionFunctions& m_g;
    const Stencil& m_stencil;
    double m_viscosity;
    double m_dt;
    size_t m_Q;
    double m_cs2;
    GlobalBoundaryData(const DistributionFunctions& f_old, DistributionFunctions& f_new, DistributionFunctions& g, const Stencil& stencil,
                       double viscosity, double dt)
        : m_fold(f_old), m_fnew(f_new), m_g(g), m_stencil(stencil)
    {
        m_viscosity = viscosity;
        m_dt = dt;
        m_cs2 = stencil.getSpeedOfSoundSquare();
        m_Q = stencil.getQ();
    }
    virtual ~GlobalBoundaryData() {}
};
}  // namespace natrium
