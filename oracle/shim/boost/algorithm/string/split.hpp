#pragma once
#include <string>
#include <boost/algorithm/string/classification.hpp>
namespace boost { namespace algorithm {
template <typename Seq, typename Pred>
Seq& split(Seq& out, std::string const& in, Pred pred) {
    out.clear(); std::string cur;
    for (size_t i = 0; i < in.size(); ++i) {
        if (pred(in[i])) { out.push_back(cur); cur.clear(); } else cur.push_back(in[i]);
    }
    out.push_back(cur); return out;
}
} using algorithm::split; }
