// Minimal stand-in for <boost/utility.hpp>, used ONLY to compile the reference
// (read-only /root/reference) as the parity oracle.  Test infrastructure.
#pragma once
#include <cassert>
#include <cstddef>
namespace boost {
class noncopyable {
protected:
    noncopyable() {}
    ~noncopyable() {}
private:
    noncopyable(noncopyable const&);
    noncopyable& operator=(noncopyable const&);
};
}
