#pragma once
#include <terark/stdtypes.hpp>
// The reference only touches the fiber pool when ReadOptions::async_io is set (MultiGet).
// The oracle never does; tasks are simply run inline.
namespace boost { namespace fibers { struct context { static context** active_pp() { static context* c = nullptr; return &c; } }; } }
namespace terark {
class FiberPool {
 public:
  struct task { void (*fn)(void*, size_t); void* ctx; size_t arg; };
  explicit FiberPool(boost::fibers::context**) {}
  void update_fiber_count(int) {}
  void push(task t) { t.fn(t.ctx, t.arg); }
  void unchecked_yield() {}
  void yield() {}
};
}
