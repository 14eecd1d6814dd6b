// oracle/ref_link_stubs.cc — three symbols the reference core references from subtrees that the
// oracle build leaves out (utilities/transactions, utilities/write_batch_with_index, rockside).
// TEST INFRASTRUCTURE ONLY (oracle/_ref build).  None is reachable from a compaction.
#include <memory>

#include "db/snapshot_checker.h"
#include "rocksdb/cache.h"
#include "rocksdb/table.h"
#include "rocksdb/utilities/write_batch_with_index.h"

namespace ROCKSDB_NAMESPACE {

namespace {
struct OracleDisableGC : public DisableGCSnapshotChecker {
  OracleDisableGC() : DisableGCSnapshotChecker() {}
};
}  // namespace

DisableGCSnapshotChecker* DisableGCSnapshotChecker::Instance() {
  static OracleDisableGC* inst = new OracleDisableGC();
  return inst;
}

// options.cc installs this as the default WriteBatchWithIndex factory; the oracle never builds one.
std::shared_ptr<WBWIFactory> SingleSkipListWBWIFactory() { return nullptr; }

// internal_stats.cc asks the (absent) dispatcher table factory for its block cache.
Cache* GetBlockCacheFromAnyTableFactory(TableFactory*) { return nullptr; }

}  // namespace ROCKSDB_NAMESPACE
