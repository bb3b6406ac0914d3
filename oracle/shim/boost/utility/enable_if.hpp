#pragma once
namespace boost {
template <bool B, typename T> struct enable_if_c {};
template <typename T> struct enable_if_c<true, T> { typedef T type; };
template <typename C, typename T = void> struct enable_if : enable_if_c<C::value, T> {};
template <typename C, typename T = void> struct disable_if : enable_if_c<!C::value, T> {};
}
