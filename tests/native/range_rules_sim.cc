// tests/native/range_rules_sim.cc — test infrastructure: toplingdb_b200/csrc/range_rules.h compiled for the host.
#include <stdint.h>
#include <string.h>

#include "range_rules.h"

using namespace b200c;

static RangeKey pack(const uint8_t* k, uint32_t n) {
  uint8_t b[16] = {0};
  memcpy(b, k, n);
  RangeKey r{0, 0, n};
  for (int i = 0; i < 8; i++) r.hi = (r.hi << 8) | b[i], r.lo = (r.lo << 8) | b[8 + i];
  return r;
}

// seps: n separator user keys in 16-byte slots; out[i] = 1 if block i has to be read for [start, end)
extern "C" void range_rules_select(const uint8_t* seps, const uint32_t* lens, uint32_t n, int has_start, const uint8_t* start, uint32_t start_len,
                                   int has_end, const uint8_t* end, uint32_t end_len, uint8_t* out) {
  const RangeKey s = pack(start, has_start ? start_len : 0), e = pack(end, has_end ? end_len : 0);
  for (uint32_t i = 0; i < n; i++) {
    const RangeKey cur = pack(seps + 16 * i, lens[i]);
    const RangeKey prev = i ? pack(seps + 16 * (i - 1), lens[i - 1]) : RangeKey{0, 0, 0};
    out[i] = block_may_touch_range(cur, i > 0, prev, has_start != 0, s, has_end != 0, e) ? 1 : 0;
  }
}
