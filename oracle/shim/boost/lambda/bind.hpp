#pragma once
#include <functional>
namespace boost { namespace lambda {
struct delete_array { template <typename T> void operator()(T* p) const { delete[] p; } };
struct delete_ptr { template <typename T> void operator()(T* p) const { delete p; } };
template <typename F, typename T> std::function<void()> bind(F f, T* p) { return [f, p]() { f(p); }; }
}}
