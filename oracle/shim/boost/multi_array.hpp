// Minimal stand-in for <boost/multi_array.hpp>, written for the oracle build only.
// Boost is not installed in this image; the reference's hot path uses exactly one
// boost type: `boost::multi_array<double, 2> __q_cache` (reference
// src/support/int_part.hh:13,23 and src/support/int_part.cc:28,34-51).  This header
// provides the few members those lines touch.  TEST INFRASTRUCTURE, not product code.
#ifndef ORACLE_SHIM_BOOST_MULTI_ARRAY_HPP
#define ORACLE_SHIM_BOOST_MULTI_ARRAY_HPP
#include <cstddef>
#include <vector>
#include <limits>
#include <tuple>
#include <cassert>
#include <algorithm>
#include <cmath>

namespace boost {

struct extent_gen2 { std::size_t a, b; };
struct extent_gen1 {
    std::size_t a;
    extent_gen2 operator[](std::size_t b) const { return extent_gen2{a, b}; }
};
struct extent_gen0 {
    extent_gen1 operator[](std::size_t a) const { return extent_gen1{a}; }
};
static const extent_gen0 extents = extent_gen0();

template <class T, std::size_t D>
class multi_array;

template <class T>
class multi_array<T, 2> {
    std::size_t shape_[2];
    std::vector<T> store_;
public:
    multi_array() { shape_[0] = 0; shape_[1] = 0; }
    const std::size_t* shape() const { return shape_; }
    void resize(const extent_gen2& e) {
        std::vector<T> fresh(e.a * e.b, T());
        std::size_t ra = std::min(e.a, shape_[0]), rb = std::min(e.b, shape_[1]);
        for (std::size_t i = 0; i < ra; ++i)
            for (std::size_t j = 0; j < rb; ++j)
                fresh[i * e.b + j] = store_[i * shape_[1] + j];
        store_.swap(fresh);
        shape_[0] = e.a; shape_[1] = e.b;
    }
    T* data() { return store_.data(); }
    const T* data() const { return store_.data(); }
    std::size_t num_elements() const { return store_.size(); }
    T* operator[](std::size_t i) { return store_.data() + i * shape_[1]; }
    const T* operator[](std::size_t i) const { return store_.data() + i * shape_[1]; }
};

}  // namespace boost
#endif
