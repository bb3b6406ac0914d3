#pragma once
#include <cstddef>
#include <iterator>
namespace boost {
template <typename R> auto begin(R const& r) -> decltype(r.begin()) { return r.begin(); }
template <typename R> auto end(R const& r) -> decltype(r.end()) { return r.end(); }
template <typename R> size_t size(R const& r) { return size_t(std::distance(r.begin(), r.end())); }
template <typename R> struct range_const_iterator { typedef typename R::const_iterator type; };
template <typename R> struct range_iterator { typedef typename R::const_iterator type; };
template <typename R> struct range_value { typedef typename R::value_type type; };
}
