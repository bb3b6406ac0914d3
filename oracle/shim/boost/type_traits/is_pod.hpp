#pragma once
#include <type_traits>
namespace boost { template <typename T> struct is_pod : std::is_pod<T> {}; }
