// Test stub (NOT Boost): boost::shared_ptr / make_shared mapped onto the standard library, enough for the PCL-interface stub.
#pragma once
#include <memory>
namespace boost {
template <class T> using shared_ptr = std::shared_ptr<T>;
using std::make_shared;
}  // namespace boost
