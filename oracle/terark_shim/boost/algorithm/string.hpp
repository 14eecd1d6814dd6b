#pragma once
// stand-in for the one boost string algorithm the reference calls
#include <string>
namespace boost {
inline std::string replace_all_copy(std::string s, const std::string& from, const std::string& to) {
  if (from.empty()) return s;
  for (size_t pos = 0; (pos = s.find(from, pos)) != std::string::npos; pos += to.size()) s.replace(pos, from.size(), to);
  return s;
}
}
