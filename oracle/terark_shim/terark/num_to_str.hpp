#pragma once
#include "fstring.hpp"
#include <sstream>
namespace terark {
// std::string that also accepts operator<< / operator|
template <class S = std::string>
struct string_appender : public S {
  using S::S;
  template <class T> string_appender& operator<<(const T& v) { std::ostringstream o; o << v; this->append(o.str()); return *this; }
  string_appender& operator<<(const char* s) { this->append(s); return *this; }
  string_appender& operator<<(const std::string& s) { this->append(s); return *this; }
  string_appender& operator<<(fstring s) { this->append(s.p, s.size()); return *this; }
  string_appender& operator<<(char c) { this->push_back(c); return *this; }
  template <class T> string_appender& operator|(const T& v) { return *this << v; }
  template <class T> string_appender& operator^(const T& v) { return *this << v; }
};
}
