// Stand-in for boost/assign/std/vector.hpp (test infrastructure, see ../../../../README.md): `v += a, b, c;`
#pragma once
#include <vector>
namespace boost {
namespace assign {
template <class T>
class list_inserter_stub {
    std::vector<T>& v_;
public:
    explicit list_inserter_stub(std::vector<T>& v) : v_(v) {}
    template <class U>
    list_inserter_stub& operator,(const U& x) { v_.push_back(T(x)); return *this; }
};
template <class T, class U>
list_inserter_stub<T> operator+=(std::vector<T>& v, const U& x)
{
    v.push_back(T(x));
    return list_inserter_stub<T>(v);
}
}  // namespace assign
}  // namespace boost
