// Stand-in for L/boundaries/Boundary.h (test infrastructure): the base class as far as ThermalBounceBack uses it --
// BoundaryName (L/boundaries/Boundary.h:29-43, same enumerators), PrescribedBoundaryValues with getVelocity(), the constructor
// ThermalBounceBack calls and FEBoundaryValues::getData() (L/boundaries/FEBoundaryValues.h).
#pragma once
#include "BoundaryTools.h"
#include "BoundaryFlags.h"

namespace natrium {

enum BoundaryName { BOUNDARY_NOT_SET, PERIODIC_BOUNDARY, ZERO_VELOCITY_NEQ_BB, CONSTANT_VELOCITY_NEQ_BB, NONCONSTANT_VELOCITY_NEQ_BB, FIRST_ORDER_BB,
                    PRESSURE_EQUILIBRIUM_BOUNDARY, VELOCITY_EQUILIBRIUM_BOUNDARY, DO_NOTHING_BC, THERMAL_BB };

template <size_t dim>
class PrescribedBoundaryValues {
    boost::shared_ptr<dealii::Function<dim> > m_velocity;
public:
    PrescribedBoundaryValues(boost::shared_ptr<dealii::Function<dim> > velocity) : m_velocity(velocity) {}
    boost::shared_ptr<dealii::Function<dim> > getVelocity() const { return m_velocity; }
};

template <size_t dim>
class FEBoundaryValues {
    GlobalBoundaryData& m_data;
public:
    explicit FEBoundaryValues(GlobalBoundaryData& data) : m_data(data) {}
    GlobalBoundaryData& getData() { return m_data; }
    const GlobalBoundaryData& getData() const { return m_data; }
    dealii::Point<dim> getPoint(size_t) const { return dealii::Point<dim>(); }
};

template <size_t dim>
class Boundary {
    size_t m_boundaryIndicator;
    BoundaryName m_boundaryName;
    PrescribedBoundaryValues<dim> m_boundaryValues;
public:
    Boundary(size_t boundaryIndicator, BoundaryName boundaryName, const PrescribedBoundaryValues<dim>& values)
        : m_boundaryIndicator(boundaryIndicator), m_boundaryName(boundaryName), m_boundaryValues(values) {}
    virtual ~Boundary() {}
    const PrescribedBoundaryValues<dim>& getBoundaryValues() const { return m_boundaryValues; }
    BoundaryName getBoundaryName() const { return m_boundaryName; }
    virtual bool isPeriodic() const { return false; }
    virtual bool isDGSupported() const { return false; }
    virtual bool isSLSupported() const { return false; }
    virtual void calculateBoundaryValues(FEBoundaryValues<dim>&, size_t, const LagrangianPathDestination&, double, double) = 0;
    virtual BoundaryFlags getUpdateFlags() const { return only_distributions; }
};
}  // namespace natrium
