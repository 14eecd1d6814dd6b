#pragma once
#include "../stdtypes.hpp"
namespace terark {
// LEB128 (7 bits per byte, low group first)
inline unsigned char* save_var_uint64(unsigned char* p, uint64_t x) {
  while (x >= 0x80) { *p++ = (unsigned char)(x | 0x80); x >>= 7; }
  *p++ = (unsigned char)x;
  return p;
}
inline uint64_t load_var_uint64(const unsigned char* p, const unsigned char** end) {
  uint64_t x = 0; int sh = 0;
  for (;; sh += 7) { unsigned char b = *p++; x |= uint64_t(b & 0x7f) << sh; if (!(b & 0x80)) break; }
  *end = p;
  return x;
}
}
