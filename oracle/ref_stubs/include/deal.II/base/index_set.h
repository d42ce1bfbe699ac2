// Stand-in for deal.II's IndexSet (test infrastructure, see ../../../README.md): a contiguous range [0, n).
#pragma once
#include <cstddef>
namespace dealii {
class IndexSet {
    std::size_t n_ = 0;
public:
    class ElementIterator {
        std::size_t i_;
    public:
        explicit ElementIterator(std::size_t i) : i_(i) {}
        std::size_t operator*() const { return i_; }
        ElementIterator& operator++() { ++i_; return *this; }
        ElementIterator operator++(int) { ElementIterator t(*this); ++i_; return t; }
        bool operator!=(const ElementIterator& o) const { return i_ != o.i_; }
        bool operator==(const ElementIterator& o) const { return i_ == o.i_; }
    };
    IndexSet() = default;
    explicit IndexSet(std::size_t n) : n_(n) {}
    ElementIterator begin() const { return ElementIterator(0); }
    ElementIterator end() const { return ElementIterator(n_); }
    std::size_t n_elements() const { return n_; }
    std::size_t size() const { return n_; }
    bool is_element(std::size_t i) const { return i < n_; }
};
}  // namespace dealii
