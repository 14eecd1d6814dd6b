// toplingdb_b200/plugin/b200_table_factory.h — the TableFactory side of the drop-in boundary.
//
// Mirrors the reference's table-factory surface (include/rocksdb/table.h:844-934):
//   class B200TableFactory : public TableFactory
//       NewTableBuilder  -> B200TableBuilder: buffers the Add()ed records and, at Finish(), has the B200 encode the whole table
//                           (b200c_job_encode_kv, include/b200c.h) and appends the image to the file.  Bytes, FileSize() progression and
//                           table properties are those of BlockBasedTableBuilder (table/block_based/block_based_table_builder.cc:961-1133,
//                           1921-1977), so CompactionOutputs / FlushJob cut and install files exactly as with the stock factory.
//       NewTableReader   -> the stock BlockBasedTable reader: the files ARE BlockBasedTable files (same magic number), which is also
//                           why rockside's DispatcherTableFactory routes reads of them to "BlockBasedTable"
//                           (sideplugin/rockside/src/topling/builtin_table_factory.cc:490-560).
// A table the device encoder does not take (user keys > 16 bytes, merge operands, range tombstones, compression, collectors, ...) is
// written by the reference's own BlockBasedTableBuilder: the buffered records are replayed into it -- the reference's code, not a CPU
// copy of ours.  Selected like any table factory: ColumnFamilyOptions::table_factory, or per level through rockside's dispatcher
// (`"class": "B200BlockBasedTable"`).
#pragma once
#include <memory>

#include "rocksdb/table.h"

namespace ROCKSDB_NAMESPACE {

struct B200TableFactoryOptions {
  int device = 0;               // CUDA ordinal
  bool allow_fallback = true;   // false: a table the device cannot build fails the builder (Status::NotSupported) instead of replaying
  uint64_t min_device_bytes = 0;  // tables smaller than this go to the stock builder right away (launch overhead dominates tiny flushes)
};

class B200TableFactory : public TableFactory {
 public:
  explicit B200TableFactory(const BlockBasedTableOptions& table_options = BlockBasedTableOptions(),
                            const B200TableFactoryOptions& o = B200TableFactoryOptions());
  ~B200TableFactory() override;
  static const char* kClassName() { return "B200BlockBasedTable"; }
  const char* Name() const override { return kClassName(); }
  using TableFactory::NewTableReader;
  Status NewTableReader(const ReadOptions& ro, const TableReaderOptions& table_reader_options, std::unique_ptr<RandomAccessFileReader>&& file,
                        uint64_t file_size, std::unique_ptr<TableReader>* table_reader,
                        bool prefetch_index_and_filter_in_cache = true) const override;
  TableBuilder* NewTableBuilder(const TableBuilderOptions& table_builder_options, WritableFileWriter* file) const override;
  Status ValidateOptions(const DBOptions& db_opts, const ColumnFamilyOptions& cf_opts) const override;
  std::string GetPrintableOptions() const override;
  bool IsDeleteRangeSupported() const override { return true; }
  // the stock factory underneath (readers, fallback builders, option access)
  const std::shared_ptr<TableFactory>& inner() const { return inner_; }
  const BlockBasedTableOptions& table_options() const;
  // how many tables the device built / were replayed into the stock builder (tests, statistics)
  uint64_t device_tables() const;
  uint64_t fallback_tables() const;

 protected:
  const void* GetOptionsPtr(const std::string& name) const override;

 private:
  std::shared_ptr<TableFactory> inner_;
  B200TableFactoryOptions opt_;
  bool have_device_;
  struct Counters;
  std::unique_ptr<Counters> counters_;
};

std::shared_ptr<TableFactory> NewB200TableFactory(const BlockBasedTableOptions& table_options = BlockBasedTableOptions(),
                                                  const B200TableFactoryOptions& o = B200TableFactoryOptions());

}  // namespace ROCKSDB_NAMESPACE
