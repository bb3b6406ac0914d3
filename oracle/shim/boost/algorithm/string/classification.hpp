#pragma once
#include <string>
namespace boost {
struct is_any_of_pred { std::string chars; bool operator()(char c) const { return chars.find(c) != std::string::npos; } };
inline is_any_of_pred is_any_of(std::string const& s) { is_any_of_pred p; p.chars = s; return p; }
namespace algorithm { using boost::is_any_of; }
}
