#pragma once
// what succinct/CMakeLists.txt would generate from succinct_config.hpp.in
#define SUCCINCT_USE_LIBCXX 0
#define SUCCINCT_USE_INTRINSICS 1
#define SUCCINCT_USE_POPCNT 0
