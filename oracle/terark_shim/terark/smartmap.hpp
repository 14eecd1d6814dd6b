#pragma once
#include "stdtypes.hpp"
#include <map>
namespace terark {
template <class K, class V, size_t InlineN = 1>
class SmartMap : public std::map<K, V> {
 public:
  using std::map<K, V>::map;
  template <class F> void for_each(F f) { for (auto& kv : *this) f(kv); }
  template <class F> void for_each(F f) const { for (auto& kv : *this) f(kv); }
};
}
