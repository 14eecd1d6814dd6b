// tests/native/range_plan_sim.cc — toplingdb_b200/csrc/range_plan.h compiled for the host (test infrastructure): anchors of an index
// block, boundaries of equal-byte key ranges, byte cuts of a file per boundary.  tests/test_range_plan_host.py.
#include "range_plan.h"

using namespace b200c;

static InputTail tail_of(uint64_t ndb, uint32_t fv, int user_key) {
  InputTail t;
  t.num_data_blocks = ndb;
  t.format_version = fv;
  t.index_key_is_user_key = user_key ? 1 : 0;
  return t;
}

extern "C" {
// -> number of anchors (or -1); keys: 16 bytes per anchor
int plan_sim_anchors(const uint8_t* blk, uint64_t blk_len, uint64_t ndb, uint32_t fv, int user_key, uint32_t per_file, uint8_t* keys, uint32_t* klens,
                     uint64_t* bytes, uint32_t cap) {
  std::vector<Anchor> a;
  if (!index_anchors(blk, blk_len, tail_of(ndb, fv, user_key), per_file, &a).empty()) return -1;
  for (size_t i = 0; i < a.size() && i < cap; i++) {
    memcpy(keys + 16 * i, a[i].key, 16);
    klens[i] = a[i].klen;
    bytes[i] = a[i].bytes;
  }
  return (int)a.size();
}
// boundaries from anchors -> count
int plan_sim_boundaries(const uint8_t* keys, const uint32_t* klens, const uint64_t* bytes, uint32_t n, uint64_t total, uint32_t max_ranges,
                        uint64_t min_range_bytes, uint8_t* out_keys, uint32_t* out_lens) {
  std::vector<Anchor> a(n);
  for (uint32_t i = 0; i < n; i++) {
    memset(&a[i], 0, sizeof(Anchor));
    memcpy(a[i].key, keys + 16 * i, 16);
    a[i].klen = klens[i];
    a[i].bytes = bytes[i];
  }
  const std::vector<Anchor> b = plan_boundaries(a, total, max_ranges, min_range_bytes);
  for (size_t i = 0; i < b.size(); i++) {
    memcpy(out_keys + 16 * i, b[i].key, 16);
    out_lens[i] = b[i].klen;
  }
  return (int)b.size();
}
int plan_sim_cuts(const uint8_t* blk, uint64_t blk_len, uint64_t ndb, uint32_t fv, int user_key, uint64_t file_len, const uint8_t* bkeys,
                  const uint32_t* blens, uint32_t nb, uint64_t* cuts, uint64_t* data_end) {
  std::vector<Anchor> b(nb);
  for (uint32_t i = 0; i < nb; i++) {
    memset(&b[i], 0, sizeof(Anchor));
    memcpy(b[i].key, bkeys + 16 * i, blens[i]);
    b[i].klen = blens[i];
  }
  return index_range_cuts(blk, blk_len, tail_of(ndb, fv, user_key), file_len, b.data(), nb, cuts, data_end).empty() ? 0 : -1;
}
}
