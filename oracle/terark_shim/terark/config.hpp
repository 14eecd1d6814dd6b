// oracle/terark_shim — std-backed stand-ins for the handful of topling-zip ("terark")
// helpers that ToplingDB's core includes.  TEST INFRASTRUCTURE ONLY: lets the
// unmodified reference sources under /root/reference compile into oracle/_ref.
// None of this is a port of topling-zip; each helper is the obvious std:: idiom.
#pragma once
#include <cassert>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
#define TERARK_DLL_EXPORT
#define terark_likely(x) __builtin_expect(!!(x), 1)
#define terark_unlikely(x) __builtin_expect(!!(x), 0)
#define terark_no_inline __attribute__((noinline))
#define terark_forceinline inline __attribute__((always_inline))
#define terark_flatten __attribute__((flatten))
#define terark_pure_func __attribute__((pure))
#define terark_no_return __attribute__((noreturn))
#define terark_warn_unused_result __attribute__((warn_unused_result))
#define terark_returns_nonnull __attribute__((returns_nonnull))
#define terark_nonnull __attribute__((nonnull))
#define TERARK_IF_DEBUG(d, r) r
#define TERARK_IF_MSVC(m, o) o
#define TERARK_UNUSED_VAR(x) (void)(x)
#define TERARK_DIE(fmt, ...) do { fprintf(stderr, "%s:%d: die: " fmt "\n", __FILE__, __LINE__, ##__VA_ARGS__); abort(); } while (0)
#define TERARK_VERIFY(expr) do { if (!(expr)) TERARK_DIE("verify(%s) failed", #expr); } while (0)
namespace terark { namespace shim_detail {
inline const char* to_cstr(const char* s) { return s; }
inline const char* to_cstr(const std::string& s) { return s.c_str(); }
template <class T> inline T to_cstr(T v) { return v; }
}}
#define TERARK_DIE_S(fmt, ...) TERARK_DIE(fmt, ##__VA_ARGS__)
#define TERARK_VERIFY_S(expr, fmt, ...) do { if (!(expr)) { fprintf(stderr, "%s:%d: verify(%s) failed\n", __FILE__, __LINE__, #expr); abort(); } } while (0)
#define TERARK_VERIFY_S_EQ(a, b) TERARK_VERIFY((a) == (b))
#define TERARK_VERIFY_EQ(a, b) TERARK_VERIFY((a) == (b))
#define TERARK_VERIFY_LT(a, b) TERARK_VERIFY((a) < (b))
#define TERARK_VERIFY_LE(a, b) TERARK_VERIFY((a) <= (b))
#define TERARK_ASSERT_EQ(a, b) assert((a) == (b))
#define TERARK_ASSERT_LT(a, b) assert((a) < (b))
#define TERARK_ASSERT_LE(a, b) assert((a) <= (b))
