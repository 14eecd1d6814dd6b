// toplingdb_b200/csrc/range_rules.h — which data blocks of an input file a sub-compaction's key range [start, end) can touch (host +
// device).  index_decode_kernel (decode.cu) marks the blocks outside the range as empty before anything of them is read; what is decoded
// is still clipped exactly (merge.cu clip_runs_kernel).  The reference gets the same effect from ClippingIterator's Seek
// (db/compaction/clipping_iterator.h:69-93).
//
// Index entry i of a BlockBasedTable holds a separator s_i with  last_key(block i) <= s_i < first_key(block i + 1)  in internal-key
// order (ShortenedIndexBuilder, table/block_based/index_builder.h:165-233); whether the 8-byte trailer is kept does not matter for a
// user-key bound.  So block i holds only user keys <= user(s_i) and, for i > 0, only user keys >= user(s_{i-1}):
//   it cannot touch [start, end)  iff  user(s_i) < start   or   (i > 0 and user(s_{i-1}) >= end).
// tests/test_range_rules_host.py checks on reference-written files that no block with a key in range is dropped and that at most one
// block on either side is read in vain.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define B200C_RR_HD __host__ __device__ __forceinline__
#else
#define B200C_RR_HD inline
#endif

namespace b200c {

struct RangeKey {  // a user key of at most 16 bytes as two big-endian words (zero padded) + its length, like the key columns
  uint64_t hi, lo;
  uint32_t len;
};
B200C_RR_HD int range_key_cmp(const RangeKey& a, const RangeKey& b) {
  if (a.hi != b.hi) return a.hi < b.hi ? -1 : 1;
  if (a.lo != b.lo) return a.lo < b.lo ? -1 : 1;
  return (int)a.len - (int)b.len;  // equal padded bytes: the shorter key is a prefix of the longer one and sorts first
}
// separator of this block, separator of the previous block (has_prev == false for the first block of the file)
B200C_RR_HD bool block_may_touch_range(const RangeKey& sep, bool has_prev, const RangeKey& prev_sep, bool has_start, const RangeKey& start,
                                       bool has_end, const RangeKey& end) {
  if (has_start && range_key_cmp(sep, start) < 0) return false;
  if (has_end && has_prev && range_key_cmp(prev_sep, end) >= 0) return false;
  return true;
}

}  // namespace b200c
