#pragma once
#include <iterator>
#include <cstddef>
namespace boost {
struct forward_traversal_tag {};
class iterator_core_access {
public:
    template <typename D> static void increment(D& d) { d.increment(); }
    template <typename D> static bool equal(D const& a, D const& b) { return a.equal(b); }
    template <typename D> static auto dereference(D const& d) -> decltype(d.dereference()) { return d.dereference(); }
};
template <typename Derived, typename Value, typename Tag>
class iterator_facade {
public:
    typedef std::forward_iterator_tag iterator_category;
    typedef Value value_type; typedef ptrdiff_t difference_type; typedef Value* pointer; typedef Value& reference;
    Derived& operator++() { iterator_core_access::increment(self()); return self(); }
    Derived operator++(int) { Derived t(self()); ++*this; return t; }
    Value& operator*() const { return iterator_core_access::dereference(cself()); }
    Value* operator->() const { return &iterator_core_access::dereference(cself()); }
    friend bool operator==(Derived const& a, Derived const& b) { return iterator_core_access::equal(a, b); }
    friend bool operator!=(Derived const& a, Derived const& b) { return !iterator_core_access::equal(a, b); }
private:
    Derived& self() { return *static_cast<Derived*>(this); }
    Derived const& cself() const { return *static_cast<Derived const*>(this); }
};
}
