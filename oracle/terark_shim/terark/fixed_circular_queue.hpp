#pragma once
#include "stdtypes.hpp"
#include <deque>
namespace terark {
template <class T, size_t Cap>
class fixed_circular_queue {
  std::deque<T> q_;
 public:
  bool full() const { return q_.size() >= Cap - 1; }
  bool empty() const { return q_.empty(); }
  size_t size() const { return q_.size(); }
  T& front() { return q_.front(); }
  T& back() { return q_.back(); }
  void pop_front() { q_.pop_front(); }
  void push_back(T&& v) { q_.push_back(std::move(v)); }
  template <class... A> void emplace_back(A&&... a) { q_.emplace_back(std::forward<A>(a)...); }
  void clear() { q_.clear(); }
  auto begin() { return q_.begin(); }
  auto end() { return q_.end(); }
};
}
