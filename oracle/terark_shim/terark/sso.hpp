#pragma once
#include "fstring.hpp"
#include "valvec.hpp"
namespace terark {
// Byte string with char* iterators.  (The original is a small-string-optimised buffer;
// only its interface matters to the reference, so a valvec<char> backs it here.)
template <size_t N, bool = true>
class minimal_sso {
  valvec<char> b_;
 public:
  minimal_sso() = default;
  minimal_sso(const char* s, size_t n) { assign(s, n); }
  minimal_sso(const char* s) { assign(s, strlen(s)); }
  minimal_sso(fstring s) { assign(s.p, s.size()); }
  minimal_sso(const std::string& s) { assign(s.data(), s.size()); }
  minimal_sso(std::string_view s) { assign(s.data(), s.size()); }
  template <class S, class = decltype(std::declval<const S&>().data_), class = decltype(std::declval<const S&>().size_)>
  minimal_sso(const S& s) { assign(s.data_, s.size_); }
  void assign(const char* s, size_t n) { b_.assign(s, s + n); }
  void assign(fstring s) { assign(s.p, s.size()); }
  void assign(const std::string& s) { assign(s.data(), s.size()); }
  template <class S, class = decltype(std::declval<const S&>().data_), class = decltype(std::declval<const S&>().size_)>
  void assign(const S& s) { assign(s.data_, s.size_); }
  template <class F> auto assign(size_t n, F fill) -> decltype(fill((char*)nullptr, n), void()) { b_.resize(n); fill(b_.data(), n); }
  template <class F> auto risk_assign_local(size_t n, F fill) -> decltype(fill((char*)nullptr, n), void()) { b_.resize(n); fill(b_.data(), n); }
  void append(const char* s, size_t n) { b_.append(s, s + n); }
  void append(fstring s) { append(s.p, s.size()); }
  void push_back(char c) { b_.push_back(c); }
  void clear() { b_.erase_all(); }
  void destroy() { b_.clear(); }
  void swap(minimal_sso& y) { b_.swap(y.b_); }
  void reserve(size_t c) { b_.reserve(c); }
  void resize(size_t n) { b_.resize(n); }
  void resize_no_init(size_t n) { b_.resize(n); }
  void risk_set_size(size_t n) { b_.risk_set_size(n); }
  char* data() { return b_.data(); } const char* data() const { return b_.data(); }
  char* begin() { return b_.begin(); } const char* begin() const { return b_.begin(); }
  char* end() { return b_.end(); } const char* end() const { return b_.end(); }
  size_t size() const { return b_.size(); } bool empty() const { return b_.empty(); }
  size_t capacity() const { return b_.capacity(); }
  char& operator[](size_t i) { return b_[i]; } const char& operator[](size_t i) const { return b_[i]; }
  std::string str() const { return std::string(data(), size()); }
  operator fstring() const { return fstring(data(), size()); }
  template <class S> S to() const { return S(data(), size()); }
  template <class S> S notail(size_t tail) const { return S(data(), size() - tail); }
  int compare(const minimal_sso& y) const { return fstring(*this).compare(fstring(y)); }
  bool operator==(const minimal_sso& y) const { return fstring(*this) == fstring(y); }
  bool operator!=(const minimal_sso& y) const { return !(*this == y); }
  bool operator<(const minimal_sso& y) const { return compare(y) < 0; }
};
}
