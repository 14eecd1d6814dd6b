#pragma once
#include "stdtypes.hpp"
#include <algorithm>
#include <initializer_list>
#include <iterator>
#include <memory>
#include <new>
#include <type_traits>
namespace terark {
struct valvec_reserve {};
struct valvec_no_init {};
// Plain contiguous growable array with raw-pointer iterators (the reference relies on
// begin()/end() being T*).  Written for the oracle build; not a port of topling-zip.
template <class T, class SizeT = size_t>
class valvec_tpl {
  T* p_ = nullptr;
  SizeT n_ = 0, c_ = 0;
  void grow_to(size_t cap) {
    if (cap <= c_) return;
    T* q = static_cast<T*>(::operator new(sizeof(T) * cap, std::align_val_t(alignof(T) < 16 ? 16 : alignof(T))));
    for (size_t i = 0; i < n_; ++i) { new (q + i) T(std::move(p_[i])); p_[i].~T(); }
    release_mem();
    p_ = q; c_ = (SizeT)cap;
  }
  void release_mem() { if (p_) ::operator delete(p_, std::align_val_t(alignof(T) < 16 ? 16 : alignof(T))); }
  void grow_for(size_t extra) { if (n_ + extra > c_) grow_to(std::max<size_t>(n_ + extra, size_t(c_) * 2 + 4)); }
 public:
  typedef T value_type; typedef T* iterator; typedef const T* const_iterator;
  typedef T& reference; typedef const T& const_reference; typedef T* pointer; typedef const T* const_pointer;
  typedef size_t size_type; typedef ptrdiff_t difference_type;
  typedef std::reverse_iterator<T*> reverse_iterator; typedef std::reverse_iterator<const T*> const_reverse_iterator;
  valvec_tpl() = default;
  explicit valvec_tpl(size_t n) { resize(n); }
  valvec_tpl(size_t n, const T& v) { resize(n, v); }
  valvec_tpl(size_t cap, valvec_reserve) { reserve(cap); }
  valvec_tpl(size_t n, valvec_no_init) { resize(n); }
  valvec_tpl(const T* b, const T* e) { assign(b, e); }
  valvec_tpl(const T* b, size_t n) { assign(b, n); }
  valvec_tpl(std::initializer_list<T> il) { assign(il.begin(), il.end()); }
  valvec_tpl(const valvec_tpl& y) { assign(y.begin(), y.end()); }
  valvec_tpl(valvec_tpl&& y) noexcept : p_(y.p_), n_(y.n_), c_(y.c_) { y.p_ = nullptr; y.n_ = y.c_ = 0; }
  valvec_tpl& operator=(const valvec_tpl& y) { if (this != &y) assign(y.begin(), y.end()); return *this; }
  valvec_tpl& operator=(valvec_tpl&& y) noexcept { if (this != &y) { clear(); swap(y); } return *this; }
  ~valvec_tpl() { clear(); }
  void clear() { erase_all(); release_mem(); p_ = nullptr; c_ = 0; }
  void erase_all() { for (size_t i = 0; i < n_; ++i) p_[i].~T(); n_ = 0; }
  void swap(valvec_tpl& y) noexcept { std::swap(p_, y.p_); std::swap(n_, y.n_); std::swap(c_, y.c_); }
  T* data() { return p_; } const T* data() const { return p_; }
  T* begin() { return p_; } const T* begin() const { return p_; } const T* cbegin() const { return p_; }
  T* end() { return p_ + n_; } const T* end() const { return p_ + n_; } const T* cend() const { return p_ + n_; }
  reverse_iterator rbegin() { return reverse_iterator(end()); } reverse_iterator rend() { return reverse_iterator(begin()); }
  const_reverse_iterator rbegin() const { return const_reverse_iterator(end()); }
  const_reverse_iterator rend() const { return const_reverse_iterator(begin()); }
  size_t size() const { return n_; } size_t capacity() const { return c_; } bool empty() const { return n_ == 0; }
  size_t used_mem_size() const { return sizeof(T) * n_; } size_t full_mem_size() const { return sizeof(T) * c_; }
  T& operator[](size_t i) { return p_[i]; } const T& operator[](size_t i) const { return p_[i]; }
  T& at(size_t i) { return p_[i]; } const T& at(size_t i) const { return p_[i]; }
  T& front() { return p_[0]; } const T& front() const { return p_[0]; }
  T& back() { return p_[n_ - 1]; } const T& back() const { return p_[n_ - 1]; }
  T& ende(size_t i) { return p_[n_ - i]; } const T& ende(size_t i) const { return p_[n_ - i]; }
  void reserve(size_t cap) { grow_to(cap); }
  void reserve_aligned(size_t, size_t cap) { grow_to(cap); }
  void ensure_capacity(size_t cap) { if (cap > c_) grow_to(std::max<size_t>(cap, size_t(c_) * 2)); }
  void shrink_to_fit() {}
  void resize(size_t n) { if (n < n_) { for (size_t i = n; i < n_; ++i) p_[i].~T(); } else { grow_to(n); for (size_t i = n_; i < n; ++i) new (p_ + i) T(); } n_ = (SizeT)n; }
  void resize(size_t n, const T& v) { if (n < n_) { for (size_t i = n; i < n_; ++i) p_[i].~T(); } else { grow_to(n); for (size_t i = n_; i < n; ++i) new (p_ + i) T(v); } n_ = (SizeT)n; }
  void resize_no_init(size_t n) { resize(n); }
  void resize_fill(size_t n, const T& v = T()) { erase_all(); resize(n, v); }
  void risk_set_size(size_t n) { static_assert(std::is_trivially_destructible<T>::value, ""); grow_to(n); n_ = (SizeT)n; }
  void risk_set_data(T* p) { p_ = p; }
  void risk_set_capacity(size_t c) { c_ = (SizeT)c; }
  void risk_release_ownership() { p_ = nullptr; n_ = c_ = 0; }
  T* grow_no_init(size_t k) { grow_for(k); T* r = p_ + n_; for (size_t i = 0; i < k; ++i) new (r + i) T(); n_ += (SizeT)k; return r; }
  void push_back(const T& v) { if (n_ == c_) { T tmp(v); grow_for(1); new (p_ + n_) T(std::move(tmp)); } else new (p_ + n_) T(v); ++n_; }
  void push_back(T&& v) { if (n_ == c_) { T tmp(std::move(v)); grow_for(1); new (p_ + n_) T(std::move(tmp)); } else new (p_ + n_) T(std::move(v)); ++n_; }
  void unchecked_push_back(const T& v) { push_back(v); }
  template <class... A> T& emplace_back(A&&... a) { if (n_ == c_) { T tmp(std::forward<A>(a)...); grow_for(1); new (p_ + n_) T(std::move(tmp)); } else new (p_ + n_) T(std::forward<A>(a)...); return p_[n_++]; }
  void pop_back() { p_[--n_].~T(); }
  template <class It> void assign(It b, It e) { erase_all(); append(b, e); }
  template <class It, class = std::enable_if_t<std::is_pointer<It>::value>> void assign(It b, size_t n) { erase_all(); append(b, b + n); }
  void assign(size_t n, const T& v) { erase_all(); resize(n, v); }
  template <class It> void append(It b, It e) { size_t k = (size_t)std::distance(b, e); grow_for(k); for (; b != e; ++b) new (p_ + n_++) T(*b); }
  template <class U> void append(const U* b, size_t k) { static_assert(sizeof(U) == sizeof(T), ""); append((const T*)b, (const T*)b + k); }
  void append(const valvec_tpl& y) { append(y.begin(), y.end()); }
  T* insert(const T* pos, const T& v) { size_t i = size_t(pos - p_); push_back(v); std::rotate(p_ + i, p_ + n_ - 1, p_ + n_); return p_ + i; }
  template <class It> T* insert(const T* pos, It b, It e) { size_t i = size_t(pos - p_), o = n_; append(b, e); std::rotate(p_ + i, p_ + o, p_ + n_); return p_ + i; }
  T* erase(const T* pos) { return erase(pos, pos + 1); }
  T* erase(const T* b, const T* e) { size_t i = size_t(b - p_), j = size_t(e - p_); std::move(p_ + j, p_ + n_, p_ + i); size_t k = j - i; for (size_t t = n_ - k; t < n_; ++t) p_[t].~T(); n_ -= (SizeT)k; return p_ + i; }
  void erase_i(size_t i, size_t k = 1) { erase(p_ + i, p_ + i + k); }
  void fill(const T& v) { std::fill(p_, p_ + n_, v); }
  bool operator==(const valvec_tpl& y) const { return n_ == y.n_ && std::equal(p_, p_ + n_, y.p_); }
  bool operator!=(const valvec_tpl& y) const { return !(*this == y); }
};
template <class T> using valvec = valvec_tpl<T, size_t>;
template <class T> using valvec32 = valvec_tpl<T, uint32_t>;
template <class T, class S> inline void swap(valvec_tpl<T, S>& a, valvec_tpl<T, S>& b) noexcept { a.swap(b); }
// first index i in [0,n) with !(a[i] < key); upper_bound_0: first with key < a[i]
template <class It, class K>
inline size_t lower_bound_0(It a, size_t n, const K& key) { return size_t(std::lower_bound(a, a + n, key) - a); }
template <class It, class K>
inline size_t upper_bound_0(It a, size_t n, const K& key) { return size_t(std::upper_bound(a, a + n, key) - a); }
}
#include <alloca.h>
#define TERARK_FAST_ALLOC(Type, var, n) ::terark::valvec<Type> var##_holder_((size_t)(n)); Type* var = var##_holder_.data()
#define TERARK_FAST_ARRAY(Type, var, n) TERARK_FAST_ALLOC(Type, var, n)
#define TERARK_FAST_CLEAN(var, n, cap) ((void)0)
