// tests/native/group_rules_sim.cc — test infrastructure: toplingdb_b200/csrc/group_rules.h (the per-user-key serial walk of the
// CompactionIterator rules, host + device) compiled for the host.  tests/test_group_rules_host.py groups a merged input stream by user
// key, walks every group and compares the result with the oracle's iterator.
#include <stdint.h>

#include "group_rules.h"

using namespace b200c;

extern "C" int group_rules_walk(const uint64_t* seqs, const uint8_t* types, uint32_t n, const uint64_t* snapshots, uint32_t num_snapshots,
                                uint32_t bottommost, uint64_t earliest_write_conflict_snapshot, uint32_t key_not_exists, uint32_t filter_removes_newest,
                                uint32_t first_key_of_the_job, uint8_t* verdicts /* n x 4 */, uint32_t* counters /* 4 */) {
  GroupVersion v[4096];
  GroupVerdict out[4096];
  if (n > 4096) return -2;
  for (uint32_t i = 0; i < n; i++) v[i] = GroupVersion{seqs[i], types[i]};
  GroupRules r{snapshots, num_snapshots, bottommost, earliest_write_conflict_snapshot, key_not_exists, filter_removes_newest, first_key_of_the_job};
  GroupCounters c{0, 0, 0, 0};
  const int rc = group_walk(v, n, r, out, &c);
  for (uint32_t i = 0; i < n; i++) {
    verdicts[4 * i] = out[i].keep;
    verdicts[4 * i + 1] = out[i].out_type;
    verdicts[4 * i + 2] = out[i].clear_value;
    verdicts[4 * i + 3] = out[i].zero_seq;
  }
  counters[0] += c.drop_hidden;
  counters[1] += c.drop_obsolete;
  counters[2] += c.optimized_del_drop_obsolete;
  counters[3] += c.drop_user;
  return rc;
}
