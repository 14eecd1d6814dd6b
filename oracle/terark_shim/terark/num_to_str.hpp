#pragma once
#include "fstring.hpp"
#include <sstream>
namespace terark {
// tmp | ReplaceSubStr<>{source, from, to}: appends source with every occurrence of `from` replaced by `to`
template <class Dummy = void>
struct ReplaceSubStr {
  fstring src, from, to;
};
// std::string that also accepts operator<< / operator|
template <class S = std::string>
struct string_appender : public S {
  using S::S;
  string_appender() = default;
  template <class Reserve> string_appender(Reserve, size_t cap) { this->reserve(cap); }  // (valvec_reserve(), n) of the original
  template <class N> void write(const char* p, N n) { this->append(p, (size_t)(n < 0 ? 0 : n)); }
  template <class D> string_appender& operator|(const ReplaceSubStr<D>& r) {
    const char* p = r.src.p;
    const char* end = p + r.src.n;
    while (p < end) {
      const char* hit = r.from.n > 0 ? (const char*)memmem(p, (size_t)(end - p), r.from.p, (size_t)r.from.n) : nullptr;
      if (!hit) break;
      this->append(p, (size_t)(hit - p));
      this->append(r.to.p, (size_t)r.to.n);
      p = hit + r.from.n;
    }
    this->append(p, (size_t)(end - p));
    return *this;
  }
  S& str() { return *this; }
  const S& str() const { return *this; }
  template <class T> string_appender& operator<<(const T& v) { std::ostringstream o; o << v; this->append(o.str()); return *this; }
  string_appender& operator<<(const char* s) { this->append(s); return *this; }
  string_appender& operator<<(const std::string& s) { this->append(s); return *this; }
  string_appender& operator<<(fstring s) { this->append(s.p, s.size()); return *this; }
  string_appender& operator<<(char c) { this->push_back(c); return *this; }
  template <class T> string_appender& operator|(const T& v) { return *this << v; }
  template <class T> string_appender& operator^(const T& v) { return *this << v; }
};
}
