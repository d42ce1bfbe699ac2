// Stand-in for L/solver/DistributionFunctions.h:47-300 (test infrastructure, see ../../README.md): Q vectors, at(i).
#pragma once
#include "../utilities/BasicNames.h"
namespace natrium {
class DistributionFunctions {
    std::vector<distributed_vector> m_f;
public:
    DistributionFunctions() = default;
    DistributionFunctions(double* base, size_t Q, size_t n, size_t stride)
    {
        for (size_t q = 0; q < Q; q++) m_f.emplace_back(base + q * stride, n);
    }
    distributed_vector& at(size_t i) { return m_f.at(i); }
    const distributed_vector& at(size_t i) const { return m_f.at(i); }
    size_t size() const { return m_f.size(); }
};
}  // namespace natrium
