#pragma once
#include "stdtypes.hpp"
#include <string>
#include <string_view>
namespace terark {
// non-owning (ptr,len) string view with the member names the reference uses
struct fstring {
  const char* p = "";
  ptrdiff_t n = 0;
  fstring() = default;
  fstring(const char* s) : p(s), n((ptrdiff_t)strlen(s)) {}
  fstring(const char* s, size_t l) : p(s), n((ptrdiff_t)l) {}
  fstring(const unsigned char* s, size_t l) : p((const char*)s), n((ptrdiff_t)l) {}
  fstring(const std::string& s) : p(s.data()), n((ptrdiff_t)s.size()) {}
  fstring(std::string_view s) : p(s.data()), n((ptrdiff_t)s.size()) {}
  template <class S, class = decltype(std::declval<const S&>().data_), class = decltype(std::declval<const S&>().size_)>
  fstring(const S& s) : p(s.data_), n((ptrdiff_t)s.size_) {}
  const char* data() const { return p; }
  const char* c_str() const { return p; }
  size_t size() const { return (size_t)n; }
  bool empty() const { return n == 0; }
  const char* begin() const { return p; }
  const char* end() const { return p + n; }
  std::string str() const { return std::string(p, (size_t)n); }
  operator std::string_view() const { return std::string_view(p, (size_t)n); }
  char operator[](ptrdiff_t i) const { return p[i]; }
  size_t commonPrefixLen(fstring y) const {
    size_t m = (size_t)(n < y.n ? n : y.n), i = 0;
    while (i < m && p[i] == y.p[i]) ++i;
    return i;
  }
  bool startsWith(fstring x) const { return n >= x.n && memcmp(p, x.p, (size_t)x.n) == 0; }
  bool starts_with(fstring x) const { return startsWith(x); }
  fstring substr(size_t pos) const { return pos < (size_t)n ? fstring(p + pos, (size_t)n - pos) : fstring(); }
  fstring substr(size_t pos, size_t len) const {
    if (pos >= (size_t)n) return fstring();
    const size_t m = (size_t)n - pos;
    return fstring(p + pos, len < m ? len : m);
  }
  int compare(fstring y) const {
    size_t m = (size_t)(n < y.n ? n : y.n);
    int r = memcmp(p, y.p, m);
    return r ? r : (n < y.n ? -1 : n > y.n);
  }
};
inline bool operator==(fstring a, fstring b) { return a.n == b.n && memcmp(a.p, b.p, (size_t)a.n) == 0; }
inline bool operator!=(fstring a, fstring b) { return !(a == b); }
inline bool operator<(fstring a, fstring b) { return a.compare(b) < 0; }
inline size_t commonPrefixLen(fstring a, fstring b) { return a.commonPrefixLen(b); }
struct fstring_func {
  struct hash { size_t operator()(fstring s) const { return std::hash<std::string_view>()(std::string_view(s)); } };
  struct equal { bool operator()(fstring a, fstring b) const { return a == b; } };
  typedef hash hash_align; typedef equal equal_align;
  struct less { bool operator()(fstring a, fstring b) const { return a < b; } };
};
inline bool getEnvBool(const char* name, bool dflt = false) {
  const char* v = getenv(name);
  if (!v) return dflt;
  return !(v[0] == '0' || v[0] == 'f' || v[0] == 'F' || v[0] == 'n' || v[0] == 'N' || v[0] == 0);
}
inline long getEnvLong(const char* name, long dflt = 0) {
  const char* v = getenv(name);
  return v ? strtol(v, nullptr, 0) : dflt;
}
inline double getEnvDouble(const char* name, double dflt = 0) {
  const char* v = getenv(name);
  return v ? strtod(v, nullptr) : dflt;
}
}
