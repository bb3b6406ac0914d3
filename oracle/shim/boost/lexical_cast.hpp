#pragma once
#include <sstream>
#include <string>
#include <cstring>
#include <stdexcept>
namespace boost {
template <typename T, typename S> T lexical_cast(S const& s) {
    std::stringstream ss; ss << s; T v; ss >> v;
    if (ss.fail()) throw std::runtime_error("bad lexical_cast");
    return v;
}
}
