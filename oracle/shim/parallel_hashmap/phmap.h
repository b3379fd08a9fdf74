// TEST INFRASTRUCTURE ONLY (oracle/): parallel-hashmap is an un-vendored dependency of the
// reference (Makefile:110-120).  std::unordered_map with a pair-aware hasher provides the
// find/insert/operator[]/reserve/clear/swap/iteration surface pyci.h:157 relies on.
#pragma once
#include <cstddef>
#include <functional>
#include <unordered_map>
#include <utility>

namespace phmap {

template<class K>
struct shim_hash {
    std::size_t operator()(const K &k) const { return std::hash<K>()(k); }
};
template<class A, class B>
struct shim_hash<std::pair<A, B>> {
    std::size_t operator()(const std::pair<A, B> &k) const {
        std::size_t h = std::hash<A>()(k.first);
        return h ^ (std::hash<B>()(k.second) + 0x9e3779b97f4a7c15ULL + (h << 6) + (h >> 2));
    }
};

template<class K, class V>
using flat_hash_map = std::unordered_map<K, V, shim_hash<K>>;

} // namespace phmap
