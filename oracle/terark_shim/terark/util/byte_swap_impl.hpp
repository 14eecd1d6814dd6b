#pragma once
#include "../stdtypes.hpp"
namespace terark {
inline uint16_t byte_swap(uint16_t x) { return __builtin_bswap16(x); }
inline uint32_t byte_swap(uint32_t x) { return __builtin_bswap32(x); }
inline uint64_t byte_swap(uint64_t x) { return __builtin_bswap64(x); }
inline unsigned long long byte_swap(unsigned long long x) { return __builtin_bswap64(x); }
inline unsigned __int128 byte_swap(unsigned __int128 x) {
  return ((unsigned __int128)__builtin_bswap64((uint64_t)x) << 64) | __builtin_bswap64((uint64_t)(x >> 64));
}
}
// little-endian hosts only (see boost/predef stub)
#define NATIVE_OF_BIG_ENDIAN(x) ::terark::byte_swap(x)
#define BIG_ENDIAN_OF(x) ::terark::byte_swap(x)
#define NATIVE_OF_LITTLE_ENDIAN(x) (x)
#define LITTLE_ENDIAN_OF(x) (x)
