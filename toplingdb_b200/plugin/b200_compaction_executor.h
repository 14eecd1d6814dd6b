// toplingdb_b200/plugin/b200_compaction_executor.h — ToplingDB plugin that hands whole compaction jobs to a B200.
//
// Mirrors the reference's executor surface one to one (db/compaction/compaction_executor.h:160-178):
//   class B200CompactionExecutorFactory : public CompactionExecutorFactory
//       ShouldRunLocal / AllowFallbackToLocal / NewExecutor / Name / JobUrl
//   class B200CompactionExecutor        : public CompactionExecutor
//       SetParams / Execute / RenameFile / CopyOneFile / CleanFiles
// and is selected exactly like DcompactEtcd is: ColumnFamilyOptions::compaction_executor_factory
// (include/rocksdb/options.h:335) or, with rockside, `CompactionExecutorFactory: {class: B200Compact}` in the JSON/YAML.
//
// It is compiled inside the ToplingDB tree (needs db/compaction/compaction_executor.h) and links libb200c.so; the only
// thing it calls for the data path is the C ABI of include/b200c.h.  Unlike topling-dcompact it runs in-process: the
// "remote worker" is the GPU.
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "db/compaction/compaction_executor.h"

namespace ROCKSDB_NAMESPACE {

struct B200CompactOptions {
  int device = 0;                       // CUDA ordinal; jobs are independent, run one factory per GPU to shard them
  bool allow_fallback_to_local = true;  // AllowFallbackToLocal(): NOT_SUPPORTED / device errors fall back to RunLocal()
  bool verify_input_checksums = true;   // ReadOptions::verify_checksums of the compaction read
  std::string scratch_dir;              // where output files are materialised before RenameFile(); default: <dbname>/b200c-tmp
  int io_threads = 8;                   // input files are read and output files written + synced by up to this many threads per job
  // Sub-compactions: a job the DB would split over threads (Compaction::ShouldFormSubcompactions, max_subcompactions > 1) is split
  // into that many key ranges of about equal input bytes (0: CompactionParams::max_subcompactions; 1: never split); the ranges run
  // concurrently and are dealt round-robin to `devices` (empty: {device}) -- one job over several GPUs.
  int max_subcompactions = 0;
  std::vector<int> devices;
};

class B200CompactionExecutorFactory : public CompactionExecutorFactory {
 public:
  explicit B200CompactionExecutorFactory(const B200CompactOptions& o = B200CompactOptions());
  ~B200CompactionExecutorFactory() override;
  // true when the job needs a rule outside the device rule set (merge operator, compaction filter, range tombstones,
  // compression, non-bytewise comparator, non-BlockBased output, > 64 runs) or no CUDA device is usable
  bool ShouldRunLocal(const Compaction*) const override;
  bool AllowFallbackToLocal() const override;
  CompactionExecutor* NewExecutor(const Compaction*) const override;
  const char* Name() const override;
  std::string JobUrl(const std::string& dbname, int job_id, int attempt) const override;

 private:
  B200CompactOptions opt_;
  bool have_device_;
};

std::shared_ptr<CompactionExecutorFactory> NewB200CompactionExecutorFactory(const B200CompactOptions& o = B200CompactOptions());

}  // namespace ROCKSDB_NAMESPACE
