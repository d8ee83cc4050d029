// Stand-in for <boost/functional/hash.hpp> (TEST INFRASTRUCTURE): hash_combine
// only (example/cbs.cpp:38-44 etc.).  Hash values only order unordered
// containers, which never influences a result of the reference.
#pragma once
// standard headers the real Boost header drags in and the reference relies on
#include <algorithm>
#include <cassert>
#include <cstddef>
#include <functional>
#include <map>
#include <set>
#include <tuple>
#include <unordered_map>
#include <unordered_set>
#include <vector>
namespace boost {
template <class T>
inline void hash_combine(std::size_t& seed, const T& v) {
  seed ^= std::hash<T>()(v) + 0x9e3779b9 + (seed << 6) + (seed >> 2);
}
}  // namespace boost
