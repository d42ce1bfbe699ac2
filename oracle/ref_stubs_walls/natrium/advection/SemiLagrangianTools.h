// Stand-in for L/advection/SemiLagrangianTools.h (test infrastructure): LagrangianPathDestination as defined there (:36-45).
#pragma once
#include <cstddef>
namespace natrium {
struct LagrangianPathDestination {
    size_t index;
    size_t direction;   // in case of a boundary: the outgoing direction
    LagrangianPathDestination(size_t i, size_t alpha) : index(i), direction(alpha) {}
    LagrangianPathDestination(const LagrangianPathDestination& other) : index(other.index), direction(other.direction) {}
};
}  // namespace natrium
