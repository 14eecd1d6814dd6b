#pragma once
#include "../stdtypes.hpp"
#include <functional>
namespace terark {
using std::function;
template <class T> using MoveConsFunc = std::function<T>;
}
#define TERARK_CMP(field, op) [](const auto& x, const auto& y) { return x.field op y.field; }
#define TERARK_CMP_P(field, op) [](const auto* x, const auto* y) { return x->field op y->field; }
#define TERARK_GET(field) [](const auto& x) -> decltype(auto) { return (x field); }
#define TERARK_FIELD(field) [](const auto& x) -> decltype(auto) { return (x.field); }
#define TERARK_PP_CAT_(a, b) a##b
#define TERARK_PP_CAT(a, b) TERARK_PP_CAT_(a, b)
namespace terark { namespace shim_detail {
template <class F> struct scope_exit { F f; ~scope_exit() { f(); } };
struct scope_exit_tag {};
template <class F> scope_exit<F> operator+(scope_exit_tag, F f) { return scope_exit<F>{std::move(f)}; }
}}
#define TERARK_SCOPE_EXIT(...) auto TERARK_PP_CAT(scope_exit_, __LINE__) = ::terark::shim_detail::scope_exit_tag() + [&]() { __VA_ARGS__; }
#define TERARK_C_CALLBACK(lambda) \
  [](void* ctx, auto... a) { (*static_cast<std::remove_reference_t<decltype(lambda)>*>(ctx))(a...); }, (void*)&lambda
