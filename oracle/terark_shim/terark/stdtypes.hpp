#pragma once
#include "config.hpp"
namespace terark {
typedef unsigned char byte_t;
template <class T> inline T pow2_align_up(T x, size_t a) { return T((x + (a - 1)) & ~T(a - 1)); }
template <class T> inline T pow2_align_down(T x, size_t a) { return T(x & ~T(a - 1)); }
}
namespace terark {
template <class T> inline T unaligned_load(const void* p) { T v; memcpy(&v, p, sizeof v); return v; }
template <class T> inline T unaligned_load(const void* p, size_t i) { T v; memcpy(&v, (const char*)p + i * sizeof(T), sizeof v); return v; }
template <class T> inline void unaligned_save(void* p, T v) { memcpy(p, &v, sizeof v); }
template <class T> inline T aligned_load(const void* p) { return *(const T*)p; }
}
using terark::unaligned_load;
