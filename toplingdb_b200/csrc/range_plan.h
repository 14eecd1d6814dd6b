// toplingdb_b200/csrc/range_plan.h — host-only planning of key ranges over a job's input files (plain C++, no CUDA): anchors read off
// an index block, boundaries of about equal input bytes, and the byte ranges of a file that a key range can touch.  It is the host
// half of b200c_job_plan_ranges / b200c_job_upload_by_ranges (api.cu) and stands in for TableReader::ApproximateKeyAnchors +
// CompactionJob::GenSubcompactionBoundaries (db/compaction/compaction_job.cc:465-640).  tests/test_range_plan_host.py compiles it for
// the host and checks it on reference-written tables.
#pragma once
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "sst_host.h"

namespace b200c {

constexpr int kPlanMaxUserKey = 16;  // = kMaxUserKey of the device columns (common.cuh); checked in api.cu

struct Anchor {
  uint8_t key[kPlanMaxUserKey];
  uint32_t klen;
  uint64_t bytes;  // data-block bytes of the anchor's file since its previous anchor
};
inline bool host_varint(const uint8_t*& p, const uint8_t* end, uint64_t* v) {
  uint64_t r = 0;
  for (int sft = 0; sft <= 63 && p < end; sft += 7) {
    const uint8_t c = *p++;
    r |= (uint64_t)(c & 127) << sft;
    if (c < 128) {
      *v = r;
      return true;
    }
  }
  return false;
}
// Entry i of an index block whose every entry is a restart point (index_block_restart_interval == 1, the default and the only value
// the device's index builder writes): the restart array addresses it directly, the key is stored whole and the handle in full.
struct IndexView {
  const uint8_t* blk;
  const uint8_t* end;      // end of the entries
  const uint8_t* restarts;
  uint32_t nr;
  bool value_delta, user_key;
  bool entry(uint32_t i, const uint8_t** key, size_t* ulen, uint64_t* off, uint64_t* size) const {
    const uint8_t* r = restarts + 4ull * i;
    const uint32_t ro = (uint32_t)r[0] | (uint32_t)r[1] << 8 | (uint32_t)r[2] << 16 | (uint32_t)r[3] << 24;
    const uint8_t* p = blk + ro;
    if (p >= end) return false;
    uint64_t shared, non_shared, vl = 0;
    if (!host_varint(p, end, &shared) || shared != 0 || !host_varint(p, end, &non_shared)) return false;
    if (!value_delta && !host_varint(p, end, &vl)) return false;
    if (non_shared > (uint64_t)(end - p)) return false;
    *key = p;
    size_t u = (size_t)non_shared;
    if (!user_key) {
      if (u < 8) return false;
      u -= 8;
    }
    *ulen = u;
    p += non_shared;
    return host_varint(p, end, off) && host_varint(p, end, size);
  }
};
inline bool index_view(const uint8_t* blk, uint64_t size, const InputTail& t, IndexView* v) {
  if (size < 8) return false;
  const uint32_t nr = ((uint32_t)blk[size - 4] | (uint32_t)blk[size - 3] << 8 | (uint32_t)blk[size - 2] << 16 | (uint32_t)blk[size - 1] << 24) & 0x7fffffffu;
  if (4ull * nr + 4 > size || nr != t.num_data_blocks || nr == 0) return false;  // (other restart intervals: the sequential walk)
  v->blk = blk;
  v->restarts = blk + size - 4 - 4ull * nr;
  v->end = v->restarts;
  v->nr = nr;
  v->value_delta = t.format_version >= 4;
  v->user_key = t.index_key_is_user_key != 0;
  return true;
}
// Walks one index block on the host (entry layout: table/block_based/block_builder.cc:21-32, values: table/format.cc:102-140) and
// appends about `per_file` anchors: the separator of every (nblocks / per_file)-th data block as a user key with the data bytes
// since the previous anchor -- what TableReader::ApproximateKeyAnchors gives GenSubcompactionBoundaries (compaction_job.cc:520-560).
inline std::string index_anchors(const uint8_t* blk, uint64_t size, const InputTail& t, uint32_t per_file, std::vector<Anchor>* out) {
  if (size < 8) return "index block too short";
  IndexView iv;
  if (index_view(blk, size, t, &iv)) {  // restart interval 1: read the ~per_file sampled entries directly
    const uint64_t stepf = std::max<uint64_t>(1, t.num_data_blocks / std::max<uint32_t>(per_file, 1));
    uint64_t last = 0;
    for (uint64_t n = stepf; n < t.num_data_blocks; n += stepf) {
      const uint8_t* kp;
      size_t ulen;
      uint64_t off, bsize;
      if (!iv.entry((uint32_t)(n - 1), &kp, &ulen, &off, &bsize)) return "malformed index entry";
      if (ulen > (size_t)kPlanMaxUserKey) continue;
      Anchor a;
      memset(&a, 0, sizeof a);
      memcpy(a.key, kp, ulen);
      a.klen = (uint32_t)ulen;
      a.bytes = off + bsize + 5 - last;
      last = off + bsize + 5;
      out->push_back(a);
    }
    return "";
  }
  const uint32_t nr = ((uint32_t)blk[size - 4] | (uint32_t)blk[size - 3] << 8 | (uint32_t)blk[size - 2] << 16 | (uint32_t)blk[size - 1] << 24) & 0x7fffffffu;
  if (4ull * nr + 4 > size) return "index block restart array out of range";
  const uint8_t* p = blk;
  const uint8_t* end = blk + size - 4 - 4ull * nr;
  const bool value_delta = t.format_version >= 4;
  const uint64_t step = std::max<uint64_t>(1, t.num_data_blocks / std::max<uint32_t>(per_file, 1));
  std::string key;
  uint64_t poff = 0, psize = 0, n = 0, last_end = 0;
  while (p < end) {
    uint64_t shared, non_shared, vl = 0, off, bsize;
    if (!host_varint(p, end, &shared) || !host_varint(p, end, &non_shared)) return "malformed index entry";
    if (!value_delta && !host_varint(p, end, &vl)) return "malformed index entry";
    if (shared > key.size() || non_shared > (uint64_t)(end - p)) return "malformed index entry";
    key.resize(shared);
    key.append(reinterpret_cast<const char*>(p), non_shared);
    p += non_shared;
    if (shared == 0 || !value_delta) {
      if (!host_varint(p, end, &off) || !host_varint(p, end, &bsize)) return "malformed index value";
    } else {
      uint64_t d;
      if (!host_varint(p, end, &d)) return "malformed index value";
      bsize = psize + (uint64_t)((int64_t)(d >> 1) ^ -(int64_t)(d & 1));
      off = poff + psize + 5;
    }
    poff = off;
    psize = bsize;
    n++;
    if (n % step == 0 && n < t.num_data_blocks) {  // (the last separator is the file's last key: nothing lies behind it)
      size_t ulen = key.size();
      if (!t.index_key_is_user_key) {
        if (ulen < 8) return "index separator shorter than a trailer";
        ulen -= 8;
      }
      if (ulen <= (size_t)kPlanMaxUserKey) {  // longer separators cannot bound a device range: their bytes go to the next anchor
        Anchor a;
        memset(&a, 0, sizeof a);
        memcpy(a.key, key.data(), ulen);
        a.klen = (uint32_t)ulen;
        a.bytes = off + bsize + 5 - last_end;
        last_end = off + bsize + 5;
        out->push_back(a);
      }
    }
  }
  return "";
}
inline int anchor_cmp(const Anchor& a, const Anchor& b) {
  const int c = memcmp(a.key, b.key, std::min(a.klen, b.klen));
  return c ? c : (int)a.klen - (int)b.klen;
}

// Boundaries of up to max_ranges key ranges of about equal input bytes, none smaller than min_range_bytes (the reference: at least one
// output file per range, compaction_job.cc:571-600).  `anchors` of all input files, in any order; total = data bytes of all files.
// A boundary is the first user key of the NEXT range: [.., key) | [key, ..).
inline std::vector<Anchor> plan_boundaries(std::vector<Anchor> anchors, uint64_t total, uint32_t max_ranges, uint64_t min_range_bytes) {
  std::vector<Anchor> out;
  if (max_ranges <= 1) return out;
  std::stable_sort(anchors.begin(), anchors.end(), [](const Anchor& a, const Anchor& b) { return anchor_cmp(a, b) < 0; });
  const uint64_t target = std::max<uint64_t>(std::max<uint64_t>(total / max_ranges, min_range_bytes), 1);
  uint64_t acc = 0;
  for (size_t i = 0; i < anchors.size() && out.size() + 1 < max_ranges; i++) {
    acc += anchors[i].bytes;
    if (acc < target) continue;
    if (anchors[i].klen == 0 || (!out.empty() && anchor_cmp(out.back(), anchors[i]) >= 0)) continue;
    out.push_back(anchors[i]);  // (any user key is a valid bound of [start, end) ranges; the separator's own key, if present, opens the next range)
    acc = 0;
  }
  return out;
}

// For one file: cuts[r] = end offset (in the file) of the data blocks the ranges 0..r can touch, i.e. the end of the first block
// whose index separator reaches boundary r (range_rules.h: that block is the last one [.., boundary r) can touch); *data_end = end
// of the last data block.  Returns "" or an error message.
inline std::string index_range_cuts(const uint8_t* blk, uint64_t blk_len, const InputTail& t, uint64_t file_len, const Anchor* bounds, uint32_t nb,
                                    uint64_t* cuts, uint64_t* data_end) {
  IndexView iv;
  if (index_view(blk, blk_len, t, &iv)) {  // restart interval 1: binary search per boundary
    const uint8_t* kp;
    size_t ul;
    uint64_t off, bsize;
    if (!iv.entry(iv.nr - 1, &kp, &ul, &off, &bsize) || off + bsize + 5 > file_len) return "malformed index entry";
    *data_end = off + bsize + 5;
    for (uint32_t r = 0; r < nb; r++) {
      uint32_t lo = 0, hi = iv.nr;  // first entry with separator >= boundary r
      while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if (!iv.entry(mid, &kp, &ul, &off, &bsize)) return "malformed index entry";
        const size_t m = std::min<size_t>(ul, bounds[r].klen);
        int c = memcmp(kp, bounds[r].key, m);
        if (c == 0) c = ul < bounds[r].klen ? -1 : (ul > bounds[r].klen ? 1 : 0);
        if (c < 0) lo = mid + 1;
        else hi = mid;
      }
      if (lo == iv.nr) {
        cuts[r] = *data_end;
      } else {
        if (!iv.entry(lo, &kp, &ul, &off, &bsize) || off + bsize + 5 > file_len) return "block handle out of range";
        cuts[r] = off + bsize + 5;
      }
    }
    return "";
  }
  if (blk_len < 8) return "index block too short";
  const uint32_t nr = ((uint32_t)blk[blk_len - 4] | (uint32_t)blk[blk_len - 3] << 8 | (uint32_t)blk[blk_len - 2] << 16 | (uint32_t)blk[blk_len - 1] << 24) & 0x7fffffffu;
  if (4ull * nr + 4 > blk_len) return "index block restart array out of range";
  const uint8_t* p = blk;
  const uint8_t* end = blk + blk_len - 4 - 4ull * nr;
  const bool value_delta = t.format_version >= 4;
  std::string key;
  uint64_t poff = 0, psize = 0, last_end = 0;
  uint32_t ri = 0;
  while (p < end) {
    uint64_t shared, non_shared, vl = 0, off, bsize;
    if (!host_varint(p, end, &shared) || !host_varint(p, end, &non_shared)) return "malformed index entry";
    if (!value_delta && !host_varint(p, end, &vl)) return "malformed index entry";
    if (shared > key.size() || non_shared > (uint64_t)(end - p)) return "malformed index entry";
    key.resize(shared);
    key.append(reinterpret_cast<const char*>(p), non_shared);
    p += non_shared;
    if (shared == 0 || !value_delta) {
      if (!host_varint(p, end, &off) || !host_varint(p, end, &bsize)) return "malformed index value";
    } else {
      uint64_t d;
      if (!host_varint(p, end, &d)) return "malformed index value";
      bsize = psize + (uint64_t)((int64_t)(d >> 1) ^ -(int64_t)(d & 1));
      off = poff + psize + 5;
    }
    poff = off;
    psize = bsize;
    if (off + bsize + 5 > file_len) return "block handle out of range";
    last_end = off + bsize + 5;
    size_t ulen = key.size();
    if (!t.index_key_is_user_key) {
      if (ulen < 8) return "index separator shorter than a trailer";
      ulen -= 8;
    }
    while (ri < nb) {  // separator >= boundary ri (user-key order; a longer separator compares by its bytes)
      const size_t m = std::min<size_t>(ulen, bounds[ri].klen);
      int c = memcmp(key.data(), bounds[ri].key, m);
      if (c == 0) c = ulen < bounds[ri].klen ? -1 : (ulen > bounds[ri].klen ? 1 : 0);
      if (c < 0) break;
      cuts[ri++] = last_end;
    }
  }
  *data_end = last_end;
  for (; ri < nb; ri++) cuts[ri] = last_end;  // every block of the file lies in front of these boundaries
  return "";
}

}  // namespace b200c
