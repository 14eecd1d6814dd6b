#pragma once
#include "hash_strmap.hpp"
namespace terark {
template <class K, class V, class H = std::hash<K>, class E = std::equal_to<K>>
class gold_hash_map : public std::unordered_map<K, V, H, E> {
  typedef std::unordered_map<K, V, H, E> base;
 public:
  using base::base;
  bool exists(const K& k) const { return base::count(k) != 0; }
  void enable_freelist() {}
  std::pair<typename base::iterator, bool> insert_i(const K& k, const V& v = V()) { return base::emplace(k, v); }
};
}
