// toplingdb_b200/csrc/group_rules.h — the CompactionIterator rules for ONE user key, as a serial walk over its versions (host + device).
//
// merge.cu evaluates the kTypeValue / kTypeDeletion rules of CompactionIterator::NextFromInput (db/compaction/compaction_iterator.cc:
// 475-1087) in parallel, entry by entry.  SingleDelete does not fit that shape: whether a SingleDelete and the Put below it cancel
// depends on what happened to the versions above them (:662-887), a chain through the whole key.  Versions of one key are few: keys
// that hold a kTypeSingleDeletion are walked serially by one thread with this function (merge.cu sd_walk_tile; the partition keeps a
// key's versions inside one tile then), all other keys take the parallel path.  tests/test_group_rules_host.py runs the same code on
// the CPU against the oracle's iterator on every scenario, SingleDelete or not.
//
// State that the reference keeps across keys and that matters here is per key: has_outputted_key_ and last_key_seq_zeroed_ are reset at
// every new user key (:578-580); clear_and_output_next_key_ never survives a key.  One wrinkle is global: has_outputted_key_ is set in
// Next() but not in SeekToFirst() (:223-226), so for the very first record of the job it is still false when the second is examined.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define B200C_GR_HD __host__ __device__ __forceinline__
#else
#define B200C_GR_HD inline
#endif

namespace b200c {

enum : uint8_t { kGrDeletion = 0, kGrValue = 1, kGrSingleDeletion = 7 };

struct GroupVersion {   // one version of the key, newest first
  uint64_t seq;
  uint8_t type;
};
struct GroupVerdict {
  uint8_t keep;         // the version is written out
  uint8_t out_type;     // its type on output (a filtered Put becomes a tombstone)
  uint8_t clear_value;  // bit 0: written without its value (:635-661 the Put behind a kept SingleDelete; :385-391 a filtered Put)
                        // bit 1 (kGrSkipped): consumed without being examined -- the iterator stepped over it inside another version's
                        // branch, so it is in none of the input statistics (total_input_raw_key_bytes, ..., :509-516)
  uint8_t zero_seq;     // sequence number zeroed (PrepareOutput :1296-1340)
};
struct GroupRules {
  const uint64_t* snapshots;  // ascending
  uint32_t num_snapshots;
  uint32_t bottommost;
  uint64_t earliest_write_conflict_snapshot;  // kMaxSequenceNumber unless a transaction DB holds one
  uint32_t key_not_exists_beyond_output_level;  // Compaction::KeyNotExistsBeyondOutputLevel for this key (worker: == bottommost)
  uint32_t filter_removes_newest;  // the compaction filter said kRemove for the newest version (only asked when it is a kTypeValue)
  uint32_t first_key_of_the_job;   // this key's first output would be the job's first record (see the wrinkle above)
};
struct GroupCounters {  // CompactionIterationStats
  uint32_t drop_hidden, drop_obsolete, optimized_del_drop_obsolete, drop_user;
};
constexpr uint64_t kGrMaxSeq = (1ull << 56) - 1;
constexpr uint8_t kGrSkipped = 2;

// findEarliestVisibleSnapshot (:1343-1396) without a snapshot checker
B200C_GR_HD uint64_t gr_earliest_visible_snapshot(const GroupRules& r, uint64_t seq, uint64_t* prev) {
  uint32_t lo = 0, hi = r.num_snapshots;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (r.snapshots[mid] < seq) lo = mid + 1;
    else hi = mid;
  }
  *prev = lo == 0 ? 0 : r.snapshots[lo - 1];
  return lo < r.num_snapshots ? r.snapshots[lo] : kGrMaxSeq;
}

// Walks the n versions of one user key (newest first).  Returns 0, or -1 when a SingleDelete meets a Delete of the same key in one
// snapshot stripe (enforce_single_del_contracts: the reference fails the job, :779-800).
B200C_GR_HD int group_walk(const GroupVersion* v, uint32_t n, const GroupRules& r, GroupVerdict* out, GroupCounters* cnt) {
  const uint64_t earliest_snapshot = r.num_snapshots ? r.snapshots[0] : kGrMaxSeq;
  const bool visible_at_tip = r.num_snapshots == 0;
  bool has_outputted_key = false, last_key_seq_zeroed = false, clear_and_output_next_key = false;
  bool first_output_pending = r.first_key_of_the_job != 0;  // Next() has not run yet: has_outputted_key stays false once more
  uint64_t cur_snap = 0;
  for (uint32_t i = 0; i < n; i++) out[i] = GroupVerdict{0, v[i].type, 0, 0};
  uint32_t i = 0;
  auto emit = [&](uint32_t at) {  // the version at `at` is valid output: Next() bookkeeping + PrepareOutput
    out[at].keep = 1;
    if (first_output_pending) first_output_pending = false;  // this record came out of SeekToFirst(): has_outputted_key_ is not set
    else has_outputted_key = true;
    if (r.bottommost && v[at].seq <= earliest_snapshot) {
      out[at].zero_seq = 1;
      last_key_seq_zeroed = true;
    }
  };
  while (i < n) {
    uint8_t type = v[i].type;
    const uint64_t seq = v[i].seq;
    bool filtered = false;
    if (i == 0 && r.filter_removes_newest && type == kGrValue) {  // InvokeFilterIfNeeded :385-391
      type = kGrDeletion;
      out[0].out_type = kGrDeletion;
      out[0].clear_value = 1;
      filtered = true;
      cnt->drop_user++;
    }
    (void)filtered;
    const uint64_t last_snapshot = cur_snap;
    uint64_t prev_snapshot = 0;
    cur_snap = visible_at_tip ? earliest_snapshot : gr_earliest_visible_snapshot(r, seq, &prev_snapshot);
    if (clear_and_output_next_key) {  // :635-661
      out[i].clear_value = 1;
      clear_and_output_next_key = false;
      emit(i);
      i++;
    } else if (type == kGrSingleDeletion) {  // :662-887
      if (i + 1 < n) {
        const uint64_t nseq = v[i + 1].seq;
        const uint8_t ntype = v[i + 1].type;
        if (last_key_seq_zeroed) {
          cnt->drop_hidden++;
          cnt->drop_obsolete++;
          out[i + 1].clear_value |= kGrSkipped;
          i += 2;
        } else if (prev_snapshot == 0 || nseq > prev_snapshot) {
          if (ntype == kGrSingleDeletion) {
            cnt->drop_obsolete++;
            i += 1;
          } else if (ntype == kGrDeletion) {
            cnt->drop_obsolete++;
            return -1;
          } else if (has_outputted_key || seq <= r.earliest_write_conflict_snapshot ||
                     (earliest_snapshot < r.earliest_write_conflict_snapshot && seq <= earliest_snapshot)) {
            cnt->drop_hidden++;
            cnt->drop_obsolete++;
            out[i + 1].clear_value |= kGrSkipped;
            i += 2;
          } else {
            clear_and_output_next_key = true;
            emit(i);
            i += 1;
          }
        } else {
          emit(i);
          i += 1;
        }
      } else {  // the oldest version of the key in this job
        if (seq <= earliest_snapshot && r.key_not_exists_beyond_output_level) {
          cnt->drop_obsolete++;
          if (!r.bottommost) cnt->optimized_del_drop_obsolete++;
        } else if (last_key_seq_zeroed) {
          cnt->drop_hidden++;
          cnt->drop_obsolete++;
        } else {
          emit(i);
        }
        i += 1;
      }
    } else if (last_snapshot == cur_snap || (last_snapshot > 0 && last_snapshot < cur_snap)) {  // :890-911 hidden by a newer version
      cnt->drop_hidden++;
      i++;
    } else if (type == kGrDeletion && seq <= earliest_snapshot && r.key_not_exists_beyond_output_level) {  // :912-946
      cnt->drop_obsolete++;
      if (!r.bottommost) cnt->optimized_del_drop_obsolete++;
      i++;
    } else if (type == kGrDeletion && r.bottommost) {  // :947-990 everything the tombstone covers in its stripe goes unseen; the
      const uint32_t d = i++;                          // tombstone itself stays only if an older snapshot still sees a version below it
      while (i < n && (prev_snapshot == 0 || v[i].seq > prev_snapshot)) out[i++].clear_value |= kGrSkipped;
      if (i < n) emit(d);
    } else {
      emit(i);
      i++;
    }
  }
  return 0;
}

}  // namespace b200c
