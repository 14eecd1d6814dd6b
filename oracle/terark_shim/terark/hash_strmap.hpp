#pragma once
#include "fstring.hpp"
#include "valvec.hpp"
#include <unordered_map>
namespace terark {
template <class V>
using hash_strmap_base = std::unordered_map<std::string, V>;
// string-keyed hash map; std::unordered_map plus fstring lookups
template <class V = char>
class hash_strmap : public hash_strmap_base<V> {
  typedef hash_strmap_base<V> base;
 public:
  using base::base;
  using base::find;
  using base::operator[];
  using base::count;
  using base::erase;
  typename base::iterator find(fstring k) { return base::find(k.str()); }
  typename base::const_iterator find(fstring k) const { return base::find(k.str()); }
  V& operator[](fstring k) { return base::operator[](k.str()); }
  size_t count(fstring k) const { return base::count(k.str()); }
  size_t erase(fstring k) { return base::erase(k.str()); }
  bool exists(fstring k) const { return base::count(k.str()) != 0; }
  std::pair<typename base::iterator, bool> insert_i(fstring k, const V& v = V()) { return base::emplace(k.str(), v); }
  void enable_freelist() {}
  // "index" flavoured lookups, expressed with iterators
  typename base::const_iterator find_i(fstring k) const { return base::find(k.str()); }
  typename base::const_iterator end_i() const { return base::end(); }
  const V& val(typename base::const_iterator it) const { return it->second; }
};
}
