// toplingdb_b200/plugin/b200_table_factory.cc — see the header.  Host glue only: the table bytes come from libb200c.so
// (b200c_job_encode_kv, include/b200c.h) or, for tables outside the device rule set, from the reference's own BlockBasedTableBuilder.
#include "b200_table_factory.h"

#include <atomic>
#include <cstring>
#include <string>
#include <vector>

#include "b200c.h"
#include "db/dbformat.h"
#include "file/writable_file_writer.h"
#include "rocksdb/comparator.h"
#include "rocksdb/flush_block_policy.h"
#include "table/block_based/block_based_table_factory.h"
#include "table/block_based/filter_policy_internal.h"
#include "table/table_builder.h"
#include "util/coding.h"

namespace ROCKSDB_NAMESPACE {

struct B200TableFactory::Counters {
  std::atomic<uint64_t> device{0}, fallback{0};
};

namespace {

// Why a table has to be written by the stock builder (nullptr: the device takes it).  Decided from the OPTIONS; record-level reasons
// (a long key, a merge operand) surface in Add().
const char* WhyStock(const TableBuilderOptions& tbo, const BlockBasedTableOptions& t, bool have_device, int* bloom_millibits) {
  *bloom_millibits = 0;
  if (!have_device) return "no CUDA device";
  if (tbo.compression_type != kNoCompression) return "block compression";
  if (tbo.internal_comparator.user_comparator() != BytewiseComparator()) return "comparator other than the bytewise one";
  if (tbo.int_tbl_prop_collector_factories != nullptr && !tbo.int_tbl_prop_collector_factories->empty()) return "table-properties collectors";
  if (tbo.moptions.sample_for_compression > 0) return "sample_for_compression";
  if (t.index_type != BlockBasedTableOptions::kBinarySearch || t.data_block_index_type != BlockBasedTableOptions::kDataBlockBinarySearch ||
      t.index_block_restart_interval != 1 || t.block_align || !t.use_delta_encoding || t.format_version < 3 || t.format_version > 5 ||
      (t.checksum != kXXH3 && t.checksum != kCRC32c && t.checksum != kNoChecksum) ||
      t.prepopulate_block_cache != BlockBasedTableOptions::PrepopulateBlockCache::kDisable)
    return "BlockBasedTableOptions outside the device's format subset";
  if (t.flush_block_policy_factory && strcmp(t.flush_block_policy_factory->Name(), "FlushBlockBySizePolicyFactory") != 0)
    return "user flush-block policy";
  if (const FilterPolicy* fp = t.filter_policy.get()) {
    if (strcmp(fp->Name(), "bloomfilter") != 0 || t.partition_filters || !t.whole_key_filtering || t.optimize_filters_for_memory ||
        t.format_version < 5 || tbo.moptions.prefix_extractor != nullptr || tbo.skip_filters)
      return "filter policy other than a full Bloom filter over whole keys (format_version >= 5)";
    *bloom_millibits = static_cast<const BloomLikeFilterPolicy*>(fp)->GetMillibitsPerKey();
  }
  return nullptr;
}

// FileSize() of a table under construction, without building it: the size arithmetic of BlockBuilder (block_builder.cc:97-253) and
// FlushBlockBySizePolicy (flush_block_policy.cc:37-69).  CompactionOutputs cuts output files by what FileSize() reports after every
// Add() (compaction_outputs.cc:277,384-427), so the progression has to be the stock builder's: offset of the flushed data blocks.
class SizeModel {
 public:
  SizeModel(const BlockBasedTableOptions& t)
      : block_size_(t.block_size), limit_(((uint64_t)t.block_size * (100 - t.block_size_deviation) + 99) / 100),
        restart_interval_(t.block_restart_interval < 1 ? 1 : t.block_restart_interval) {}
  void Add(const Slice& key, const Slice& value) {
    if (entries_in_block_ > 0 && ShouldFlush(key, value)) Flush();
    size_t shared = 0;
    if (counter_ >= restart_interval_) {
      estimate_ += 4;  // a new restart point
      counter_ = 0;
    } else if (entries_in_block_ > 0) {
      const size_t n = std::min(last_key_.size(), key.size());
      while (shared < n && last_key_[shared] == key[shared]) shared++;
    }
    const size_t non_shared = key.size() - shared;
    estimate_ += VarintLength(shared) + VarintLength(non_shared) + VarintLength(value.size()) + non_shared + value.size();
    last_key_.assign(key.data(), key.size());
    counter_++;
    entries_in_block_++;
  }
  uint64_t offset() const { return offset_; }

 private:
  bool ShouldFlush(const Slice& key, const Slice& value) const {
    if (estimate_ >= block_size_) return true;
    if (limit_ == 0 || limit_ >= block_size_) {
      if (limit_ == 0) return false;  // block_size_deviation == 100
    }
    uint64_t after = estimate_ + key.size() + value.size();
    if (counter_ >= restart_interval_) after += 4;
    after += 4;  // varint for the shared-prefix length, as BlockBuilder::EstimateSizeAfterKV counts it
    after += VarintLength(key.size()) + VarintLength(value.size());
    return after > block_size_ && estimate_ > limit_;
  }
  void Flush() {
    offset_ += estimate_ + 5;  // block + trailer (type byte + checksum)
    estimate_ = 8;
    counter_ = 0;
    entries_in_block_ = 0;
  }
  const uint64_t block_size_, limit_;
  const uint32_t restart_interval_;
  uint64_t offset_ = 0, estimate_ = 8;  // one restart point + the restart count
  uint32_t counter_ = 0, entries_in_block_ = 0;
  std::string last_key_;
};

class B200TableBuilder : public TableBuilder {
 public:
  B200TableBuilder(const B200TableFactory* fac, const B200TableFactoryOptions& o, const BlockBasedTableOptions& t, const TableBuilderOptions& tbo,
                   WritableFileWriter* file, bool have_device, std::atomic<uint64_t>* n_device, std::atomic<uint64_t>* n_fallback)
      : fac_(fac), opt_(o), topt_(t), tbo_(tbo), file_(file), model_(t), n_device_(n_device), n_fallback_(n_fallback) {
    why_stock_ = WhyStock(tbo, t, have_device, &bloom_millibits_);
    props_.column_family_id = tbo.column_family_id;
    props_.column_family_name = tbo.column_family_name;
    props_.oldest_key_time = tbo.oldest_key_time;
    props_.file_creation_time = tbo.file_creation_time;
    props_.orig_file_number = tbo.cur_file_num;
    props_.db_id = tbo.db_id;
    props_.db_session_id = tbo.db_session_id;
    props_.db_host_id = tbo.ioptions.db_host_id;
    if (props_.db_host_id == kHostnameForDbHostId) {
      props_.db_host_id.clear();
      tbo.ioptions.env->GetHostNameString(&props_.db_host_id).PermitUncheckedError();
    }
    if (why_stock_ != nullptr) ToStock();
  }
  ~B200TableBuilder() override = default;

  void Add(const Slice& key, const Slice& value) override {
    if (inner_) return inner_->Add(key, value);
    if (!status_.ok()) return;
    // record-level rule set of the device encoder: user key <= 16 bytes, kTypeValue / kTypeDeletion
    bool ok = key.size() >= 8 && key.size() <= 16 + 8;
    if (ok) {
      const ValueType vt = ExtractValueType(key);
      ok = vt == kTypeValue || vt == kTypeDeletion;
    }
    if (!ok) {
      if (!opt_.allow_fallback) {
        status_ = Status::NotSupported("B200TableBuilder", "record outside the device rule set (user key > 16 bytes or a type other than Value / Deletion)");
        return;
      }
      ToStock();
      if (inner_) inner_->Add(key, value);
      return;
    }
    offs_.push_back(arena_.size());
    klens_.push_back((uint32_t)key.size());
    arena_.append(key.data(), key.size());
    arena_.append(value.data(), value.size());
    model_.Add(key, value);
    props_.num_entries++;
    props_.raw_key_size += key.size();
    props_.raw_value_size += value.size();
    if (ExtractValueType(key) == kTypeDeletion) props_.num_deletions++;
  }
  Status status() const override { return inner_ ? inner_->status() : status_; }
  IOStatus io_status() const override { return inner_ ? inner_->io_status() : io_status_; }

  Status Finish() override {
    if (!inner_ && (offs_.empty() || arena_.size() < opt_.min_device_bytes)) ToStock();  // (an empty table is the stock builder's too)
    if (inner_) return inner_->Finish();
    if (!status_.ok()) return status_;
    b200c_params bp;
    b200c_params_init(&bp);
    bp.device = opt_.device;
    bp.output_level = 0;  // one table, never cut: the caller decides where files end
    bp.bottommost_level = 0;
    bp.max_output_file_size = ~0ull;
    bp.block_size = (uint32_t)topt_.block_size;
    bp.block_size_deviation = (uint32_t)topt_.block_size_deviation;
    bp.block_restart_interval = (uint32_t)topt_.block_restart_interval;
    bp.index_block_restart_interval = (uint32_t)topt_.index_block_restart_interval;
    bp.format_version = topt_.format_version;
    bp.checksum = (uint32_t)topt_.checksum;
    bp.bloom_millibits_per_key = (uint32_t)bloom_millibits_;
    bp.column_family_id = props_.column_family_id;
    bp.column_family_name = props_.column_family_name.c_str();
    bp.db_id = props_.db_id.c_str();
    bp.db_session_id = props_.db_session_id.c_str();
    bp.db_host_id = props_.db_host_id.c_str();
    bp.creation_time = props_.creation_time;
    bp.oldest_key_time = props_.oldest_key_time;
    uint64_t fct = props_.file_creation_time;
    bp.file_creation_times = &fct;
    bp.num_file_creation_times = 1;
    bp.first_file_number = props_.orig_file_number;
    bp.output_mem = B200C_MEM_HOST;
    b200c_job* job = nullptr;
    int rc = b200c_job_create(&bp, &job);
    offs_.push_back(arena_.size());
    if (rc == B200C_OK) rc = b200c_job_encode_kv(job, klens_.size(), arena_.data(), offs_.data(), klens_.data());
    offs_.pop_back();
    const void* data = nullptr;
    uint64_t len = 0;
    b200c_file_meta m;
    if (rc == B200C_OK && b200c_job_output_count(job) != 1) rc = B200C_ERR_CUDA;
    if (rc == B200C_OK) rc = b200c_job_output_meta(job, 0, &m);
    if (rc == B200C_OK) rc = b200c_job_output_data(job, 0, &data, &len);
    if (rc != B200C_OK) {
      const std::string msg = b200c_last_error();
      if (job) b200c_job_destroy(job);
      if (!opt_.allow_fallback) {
        status_ = Status::NotSupported("B200TableBuilder", msg);
        return status_;
      }
      ToStock();
      return inner_->Finish();
    }
    io_status_ = file_->Append(Slice(static_cast<const char*>(data), (size_t)len));
    if (io_status_.ok()) io_status_ = file_->Flush();
    status_ = io_status_;
    file_size_ = len;
    finished_ = true;
    // table properties as the stock builder reports them after Finish() (the same values are inside the file)
    props_.data_size = m.data_size;
    props_.index_size = m.index_size + 5;  // with the block trailer (block_based_table_builder.cc:1560)
    props_.num_data_blocks = m.num_data_blocks;
    props_.format_version = topt_.format_version;
    props_.comparator_name = BytewiseComparator()->Name();
    props_.compression_name = "NoCompression";
    props_.index_key_is_user_key = 1;  // refined from the file below if any adjacent blocks share a user key
    props_.index_value_is_delta_encoded = topt_.format_version >= 4;
    props_.filter_policy_name = bloom_millibits_ ? topt_.filter_policy->Name() : "";
    props_.tail_start_offset = m.data_size;
    tail_size_ = len - m.data_size;
    n_device_->fetch_add(1, std::memory_order_relaxed);
    b200c_job_destroy(job);
    return status_;
  }
  void Abandon() override {
    if (inner_) inner_->Abandon();
    arena_.clear();
    offs_.clear();
    klens_.clear();
  }
  uint64_t NumEntries() const override { return inner_ ? inner_->NumEntries() : props_.num_entries; }
  bool IsEmpty() const override { return inner_ ? inner_->IsEmpty() : props_.num_entries == 0; }
  uint64_t FileSize() const override { return inner_ ? inner_->FileSize() : (finished_ ? file_size_ : model_.offset()); }
  uint64_t EstimatedFileSize() const override { return inner_ ? inner_->EstimatedFileSize() : FileSize(); }
  uint64_t GetTailSize() const override { return inner_ ? inner_->GetTailSize() : tail_size_; }
  bool NeedCompact() const override { return inner_ ? inner_->NeedCompact() : false; }
  TableProperties GetTableProperties() const override { return inner_ ? inner_->GetTableProperties() : props_; }
  std::string GetFileChecksum() const override { return inner_ ? inner_->GetFileChecksum() : (file_ ? file_->GetFileChecksum() : kUnknownFileChecksum); }
  const char* GetFileChecksumFuncName() const override {
    return inner_ ? inner_->GetFileChecksumFuncName() : (file_ ? file_->GetFileChecksumFuncName() : kUnknownFileChecksumFuncName);
  }
  void SetSeqnoTimeTableProperties(const std::string& encoded_seqno_to_time_mapping, uint64_t oldest_ancestor_time) override {
    if (!inner_ && !encoded_seqno_to_time_mapping.empty()) ToStock();  // the mapping is one more table property: stock builder
    if (inner_) return inner_->SetSeqnoTimeTableProperties(encoded_seqno_to_time_mapping, oldest_ancestor_time);
    props_.creation_time = oldest_ancestor_time;
    have_creation_time_ = true;
  }

 private:
  // hand the table to the reference's own builder: everything buffered so far is replayed, later calls are forwarded
  void ToStock() {
    if (inner_) return;
    inner_.reset(fac_->inner()->NewTableBuilder(tbo_, file_));
    if (have_creation_time_) inner_->SetSeqnoTimeTableProperties(std::string(), props_.creation_time);
    for (size_t i = 0; i < klens_.size(); i++) {
      const uint64_t o = offs_[i], e = i + 1 < offs_.size() ? offs_[i + 1] : arena_.size();
      inner_->Add(Slice(arena_.data() + o, klens_[i]), Slice(arena_.data() + o + klens_[i], (size_t)(e - o - klens_[i])));
    }
    std::string().swap(arena_);
    offs_.clear();
    klens_.clear();
    n_fallback_->fetch_add(1, std::memory_order_relaxed);
  }

  const B200TableFactory* fac_;
  const B200TableFactoryOptions opt_;
  const BlockBasedTableOptions topt_;
  const TableBuilderOptions tbo_;  // (holds references into the caller's options, as the stock builder's Rep does)
  WritableFileWriter* file_;
  std::unique_ptr<TableBuilder> inner_;
  std::string arena_;
  std::vector<uint64_t> offs_;
  std::vector<uint32_t> klens_;
  SizeModel model_;
  const char* why_stock_ = nullptr;
  int bloom_millibits_ = 0;
  Status status_;
  IOStatus io_status_;
  TableProperties props_;
  bool have_creation_time_ = false, finished_ = false;
  uint64_t file_size_ = 0, tail_size_ = 0;
  std::atomic<uint64_t>*n_device_, *n_fallback_;
};

}  // namespace

B200TableFactory::B200TableFactory(const BlockBasedTableOptions& table_options, const B200TableFactoryOptions& o)
    : inner_(NewBlockBasedTableFactory(table_options)), opt_(o), counters_(new Counters()) {
  have_device_ = b200c_device_count() > opt_.device;
}
B200TableFactory::~B200TableFactory() = default;
const BlockBasedTableOptions& B200TableFactory::table_options() const {
  return static_cast<const BlockBasedTableFactory*>(inner_.get())->table_options();
}
uint64_t B200TableFactory::device_tables() const { return counters_->device.load(); }
uint64_t B200TableFactory::fallback_tables() const { return counters_->fallback.load(); }
Status B200TableFactory::NewTableReader(const ReadOptions& ro, const TableReaderOptions& table_reader_options,
                                        std::unique_ptr<RandomAccessFileReader>&& file, uint64_t file_size,
                                        std::unique_ptr<TableReader>* table_reader, bool prefetch_index_and_filter_in_cache) const {
  return inner_->NewTableReader(ro, table_reader_options, std::move(file), file_size, table_reader, prefetch_index_and_filter_in_cache);
}
TableBuilder* B200TableFactory::NewTableBuilder(const TableBuilderOptions& tbo, WritableFileWriter* file) const {
  return new B200TableBuilder(this, opt_, table_options(), tbo, file, have_device_, &counters_->device, &counters_->fallback);
}
Status B200TableFactory::ValidateOptions(const DBOptions& db_opts, const ColumnFamilyOptions& cf_opts) const {
  return inner_->ValidateOptions(db_opts, cf_opts);
}
std::string B200TableFactory::GetPrintableOptions() const {
  return "  B200 table builder on cuda:" + std::to_string(opt_.device) + (have_device_ ? "" : " (no device: stock builder)") + "\n" +
         inner_->GetPrintableOptions();
}
// GetOptions<BlockBasedTableOptions>() of this factory answers with the stock factory's options (DB code and the B200 executor ask
// the configured table factory for them)
const void* B200TableFactory::GetOptionsPtr(const std::string& name) const {
  if (name == BlockBasedTableOptions::kName()) return &table_options();
  return TableFactory::GetOptionsPtr(name);
}

std::shared_ptr<TableFactory> NewB200TableFactory(const BlockBasedTableOptions& table_options, const B200TableFactoryOptions& o) {
  return std::make_shared<B200TableFactory>(table_options, o);
}

#ifdef B200C_WITH_SIDEPLUGIN
// rockside registration (sideplugin/rockside/src/topling/side_plugin_factory.h:290-293).  In JSON / YAML:
//   "TableFactory": { "b200_bbt": { "class": "B200BlockBasedTable", "params": { "device": 0, "block_size": 4096, "format_version": 5 } },
//                     "dispatch": { "class": "DispatcherTable", "params": { "default": "b200_bbt", "readers": { "BlockBasedTable": "bb" }, ... } } }
}  // namespace ROCKSDB_NAMESPACE
#include "topling/side_plugin_factory.h"
namespace ROCKSDB_NAMESPACE {
static std::shared_ptr<TableFactory> JS_NewB200TableFactory(const json& js, const SidePluginRepo&) {
  B200TableFactoryOptions o;
  BlockBasedTableOptions t;
  ROCKSDB_JSON_OPT_PROP_3(js, o.device, "device");
  ROCKSDB_JSON_OPT_PROP_3(js, o.allow_fallback, "allow_fallback");
  ROCKSDB_JSON_OPT_PROP_3(js, o.min_device_bytes, "min_device_bytes");
  ROCKSDB_JSON_OPT_PROP_3(js, t.block_size, "block_size");
  ROCKSDB_JSON_OPT_PROP_3(js, t.block_size_deviation, "block_size_deviation");
  ROCKSDB_JSON_OPT_PROP_3(js, t.block_restart_interval, "block_restart_interval");
  ROCKSDB_JSON_OPT_PROP_3(js, t.format_version, "format_version");
  return std::make_shared<B200TableFactory>(t, o);
}
ROCKSDB_FACTORY_REG("B200BlockBasedTable", JS_NewB200TableFactory);
#endif

}  // namespace ROCKSDB_NAMESPACE
