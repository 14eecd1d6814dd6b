// oracle/ref_compact.cc — drives the UNMODIFIED reference (oracle/_ref/libtoplingdb_ref.so) through
// its public DB API so that the reference's own FlushJob / CompactionJob::Run ->
// ProcessKeyValueCompaction (db/compaction/compaction_job.cc:642,1390) produce the input and output
// SSTs that this repo's CUDA path is compared against byte for byte.
//
// TEST INFRASTRUCTURE ONLY.  Nothing in the product (toplingdb_b200/) links or executes this.
//
// usage: ref_compact <ops.bin> <workdir> [key=value ...]
//   ops.bin  : write script (see oracle/ops_format.md): PUT/DEL/FLUSH/SNAPSHOT/COMPACT_ALL_TO records
//   workdir  : receives inputs/NNNNNN.sst (the L0 files fed to the job, newest first),
//              outputs/NNNNNN.sst, manifest.json (job parameters + CompactionJobStats)
//   options  : output_level=1 target_file_size=67108864 block_size=4096 restart_interval=16
//              checksum=xxh3|crc32c max_subcompactions=1 format_version=5 repeat=1 keep_db=0
#include <dirent.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "rocksdb/db.h"
#include "rocksdb/filter_policy.h"
#include "rocksdb/listener.h"
#include "rocksdb/options.h"
#include "rocksdb/table.h"
#include "env/composite_env_wrapper.h"
#include "rocksdb/compaction_filter.h"
#include "rocksdb/system_clock.h"
#include "rocksdb/utilities/db_ttl.h"
#include "rocksdb/write_batch.h"
#include "util/compression.h"
#include "utilities/compaction_filters/remove_emptyvalue_compactionfilter.h"
#ifdef WITH_B200_PLUGIN
#include "rocksdb/statistics.h"
#include "toplingdb_b200/plugin/b200_compaction_executor.h"
#include "toplingdb_b200/plugin/b200_table_factory.h"
#endif

using namespace ROCKSDB_NAMESPACE;

namespace {

struct Opts {
  int output_level = 1;
  uint64_t target_file_size = 64ull << 20;
  uint64_t block_size = 4096;
  int restart_interval = 16;
  std::string checksum = "xxh3";
  uint32_t max_subcompactions = 1;
  uint32_t format_version = 5;
  int keep_db = 0;
  int paranoid = 0;
  int blob = 0, ingest_behind = 0, ribbon = 0, partition_filters = 0;  // option shapes the B200 executor must leave to the CPU path
  double bloom_bits = 0;  // > 0: BlockBasedTableOptions::filter_policy = NewBloomFilterPolicy(bloom_bits) (full filter, whole keys)
  int copy = 1;  // 0: leave inputs/ and outputs/ empty (timing runs only need the manifest)
  uint64_t setup_file_size = UINT64_MAX;  // output file size limit of the set-up compactions (op 5): several files below the job
  std::string mode = "files";  // range: run the job through DB::CompactRange, the way the DB's own picker builds it -- the
                               // compaction then carries its grandparents (files at output_level + 1); CompactFiles never does
  int ttl = 0;  // > 0: open the DB as DBWithTTL (utilities/ttl): values carry a 4-byte timestamp, TtlCompactionFilter drops stale ones
  std::string filter = "none";  // remove_empty_value: the reference's RemoveEmptyValueCompactionFilter through a factory
  std::string barrier_dir;  // with barrier_n: wait until barrier_n processes have finished writing their inputs, so that
  int barrier_n = 0;        // concurrent timing runs compact at the same time
  int warm = 0;             // run script + job once through a throw-away DB first (timing runs: steady state of a long-lived process)
  std::string table_factory;  // "b200" / "b200+nofallback": flushes and (local) compactions write their tables through B200TableFactory
  std::string input_compression = "none";  // "zlib": the job's INPUT files (flushes, set-up compactions) are written with kZlibCompression
  int index_compression = 1;               // BlockBasedTableOptions::enable_index_compression
  int b200_subs = 0;          // B200CompactOptions::max_subcompactions
  std::string b200_devices;   // B200CompactOptions::devices, comma separated
  std::string executor;  // "b200": route the job through the B200 CompactionExecutor plugin (ref_compact_b200 build only)
};

void Die(const char* what, const Status& s) {
  fprintf(stderr, "ref_compact: %s: %s\n", what, s.ToString().c_str());
  exit(2);
}

std::string Hex(const std::string& s) {
  static const char* d = "0123456789abcdef";
  std::string r;
  for (unsigned char c : s) {
    r.push_back(d[c >> 4]);
    r.push_back(d[c & 15]);
  }
  return r;
}

void CopyFile(const std::string& from, const std::string& to) {
  std::ifstream in(from, std::ios::binary);
  std::ofstream out(to, std::ios::binary);
  out << in.rdbuf();
  if (!in.good() && !in.eof()) {
    fprintf(stderr, "ref_compact: copy %s failed\n", from.c_str());
    exit(2);
  }
}

class StatsListener : public EventListener {
 public:
  CompactionJobInfo last;
  int completed = 0;
  // range mode: DB::CompactRange goes on to push the result further down, so the L0 -> output_level job is captured here, right
  // after it has been installed (its files are still in place)
  int capture_level = -1;
  bool captured = false;
  bool copy = true;
  std::string dbdir, outdir;
  CompactionJobInfo job;
  std::vector<SstFileMetaData> outs;
  // per sub-compaction statistics (CompactionJob::ProcessKeyValueCompaction fills them, compaction_job.cc:1676-1700): which output
  // files belong to which key range is not reported anywhere else
  struct Sub {
    int job_id, sub_id;
    CompactionJobStats st;
  };
  std::vector<Sub> subs;
  std::mutex subs_mu;
  void OnSubcompactionCompleted(const SubcompactionJobInfo& si) override {
    std::lock_guard<std::mutex> g(subs_mu);
    subs.push_back(Sub{si.job_id, si.subcompaction_job_id, si.stats});
  }
  void OnCompactionCompleted(DB* db, const CompactionJobInfo& ci) override {
    last = ci;
    completed++;
    if (getenv("REF_COMPACT_DEBUG"))
      fprintf(stderr, "compaction #%d: L%d -> L%d, %zu inputs, %zu outputs, reason %d, in %llu out %llu records\n", completed,
              ci.base_input_level, ci.output_level, ci.input_files.size(), ci.output_files.size(), (int)ci.compaction_reason,
              (unsigned long long)ci.stats.num_input_records, (unsigned long long)ci.stats.num_output_records);
    if (capture_level >= 0 && !captured && ci.base_input_level == 0 && ci.output_level == capture_level) {
      captured = true;
      job = ci;
      ColumnFamilyMetaData cfm;
      db->GetColumnFamilyMetaData(&cfm);
      for (auto& lvl : cfm.levels)
        if (lvl.level == capture_level)
          for (auto& fm : lvl.files) {
            outs.push_back(fm);
            if (copy) CopyFile(dbdir + fm.name, outdir + fm.name);
          }
    }
  }
};

// RunRemote only accepts filters that come from a factory (compaction_job.cc:942-943)
class RemoveEmptyValueFactory : public CompactionFilterFactory {
 public:
  std::unique_ptr<CompactionFilter> CreateCompactionFilter(const CompactionFilter::Context&) override {
    return std::unique_ptr<CompactionFilter>(new RemoveEmptyValueCompactionFilter());
  }
  const char* Name() const override { return "RemoveEmptyValueCompactionFilterFactory"; }
};

// wall clock under the script's control (op 6): DBWithTTL stamps values with it and TtlCompactionFilter compares against it
class ScriptClock : public SystemClockWrapper {
 public:
  explicit ScriptClock(const std::shared_ptr<SystemClock>& base) : SystemClockWrapper(base) {}
  const char* Name() const override { return "ScriptClock"; }
  Status GetCurrentTime(int64_t* t) override {
    *t = now;
    return Status::OK();
  }
  int64_t now = 1700000000;
};

struct Reader {
  FILE* f;
  bool u8(uint8_t* v) { return fread(v, 1, 1, f) == 1; }
  uint32_t u32() {
    uint32_t v = 0;
    if (fread(&v, 4, 1, f) != 1) {
      fprintf(stderr, "ref_compact: truncated ops file\n");
      exit(2);
    }
    return v;
  }
  void bytes(std::string* s, size_t n) {
    s->resize(n);
    if (n && fread(&(*s)[0], 1, n, f) != n) {
      fprintf(stderr, "ref_compact: truncated ops file\n");
      exit(2);
    }
  }
};

}  // namespace

int main(int argc, char** argv) {
  if (argc < 3) {
    fprintf(stderr, "usage: ref_compact <ops.bin> <workdir> [key=value ...]\n");
    return 1;
  }
  Opts o;
  for (int i = 3; i < argc; i++) {
    std::string a = argv[i];
    size_t eq = a.find('=');
    if (eq == std::string::npos) continue;
    std::string k = a.substr(0, eq), v = a.substr(eq + 1);
    if (k == "output_level") o.output_level = atoi(v.c_str());
    else if (k == "target_file_size") o.target_file_size = strtoull(v.c_str(), nullptr, 0);
    else if (k == "block_size") o.block_size = strtoull(v.c_str(), nullptr, 0);
    else if (k == "restart_interval") o.restart_interval = atoi(v.c_str());
    else if (k == "checksum") o.checksum = v;
    else if (k == "max_subcompactions") o.max_subcompactions = (uint32_t)atoi(v.c_str());
    else if (k == "format_version") o.format_version = (uint32_t)atoi(v.c_str());
    else if (k == "keep_db") o.keep_db = atoi(v.c_str());
    else if (k == "paranoid") o.paranoid = atoi(v.c_str());
    else if (k == "bloom_bits") o.bloom_bits = atof(v.c_str());
    else if (k == "blob") o.blob = atoi(v.c_str());
    else if (k == "ingest_behind") o.ingest_behind = atoi(v.c_str());
    else if (k == "ribbon") o.ribbon = atoi(v.c_str());
    else if (k == "partition_filters") o.partition_filters = atoi(v.c_str());
    else if (k == "executor") o.executor = v;
    else if (k == "b200_subs") o.b200_subs = atoi(v.c_str());
    else if (k == "input_compression") o.input_compression = v;
    else if (k == "index_compression") o.index_compression = atoi(v.c_str());
    else if (k == "b200_devices") o.b200_devices = v;
    else if (k == "table_factory") o.table_factory = v;
    else if (k == "copy") o.copy = atoi(v.c_str());
    else if (k == "filter") o.filter = v;
    else if (k == "mode") o.mode = v;
    else if (k == "setup_file_size") o.setup_file_size = strtoull(v.c_str(), nullptr, 0);
    else if (k == "ttl") o.ttl = atoi(v.c_str());
    else if (k == "warm") o.warm = atoi(v.c_str());
    else if (k == "barrier_dir") o.barrier_dir = v;
    else if (k == "barrier_n") o.barrier_n = atoi(v.c_str());
    else {
      fprintf(stderr, "ref_compact: unknown option %s\n", k.c_str());
      return 1;
    }
  }
  const std::string work = argv[2];
  const std::string dbdir = work + "/db";
  mkdir(work.c_str(), 0755);
  mkdir((work + "/inputs").c_str(), 0755);
  mkdir((work + "/outputs").c_str(), 0755);

  Options opt;
  opt.create_if_missing = true;
  opt.disable_auto_compactions = true;
  const CompressionType in_comp = o.input_compression == "zlib" ? kZlibCompression : kNoCompression;
  if (o.input_compression != "zlib" && o.input_compression != "none") {
    fprintf(stderr, "ref_compact: unknown input_compression %s\n", o.input_compression.c_str());
    return 1;
  }
  if (in_comp != kNoCompression && !CompressionTypeSupported(in_comp)) {
    fprintf(stderr, "ref_compact: this build of the reference has no zlib (-DZLIB)\n");
    return 1;
  }
  opt.compression = in_comp;  // flushes write the job's L0 inputs; the measured job itself writes uncompressed (co.compression below)
  opt.bottommost_compression = kDisableCompressionOption;
  opt.compression_per_level.clear();
  opt.num_levels = 7;
  opt.write_buffer_size = size_t(8) << 30;  // flushes happen only where the script says FLUSH
  opt.max_write_buffer_number = 4;
  opt.level0_file_num_compaction_trigger = 1 << 20;
  opt.level0_slowdown_writes_trigger = 1 << 20;
  opt.level0_stop_writes_trigger = 1 << 20;
  opt.target_file_size_base = o.target_file_size;
  opt.max_subcompactions = o.max_subcompactions;
  opt.max_background_jobs = 2;
  opt.paranoid_file_checks = o.paranoid != 0;
  opt.enable_blob_files = o.blob != 0;
  opt.min_blob_size = 16;
  opt.allow_ingest_behind = o.ingest_behind != 0;
  if (o.mode == "range") opt.level_compaction_dynamic_level_bytes = false;  // CompactRange then goes L0 -> L1, not to a base level
  opt.info_log_level = WARN_LEVEL;
  opt.stats_dump_period_sec = 0;
  opt.stats_persist_period_sec = 0;
  BlockBasedTableOptions t;
  t.block_size = o.block_size;
  t.block_restart_interval = o.restart_interval;
  t.format_version = o.format_version;
  t.checksum = o.checksum == "crc32c" ? kCRC32c : kXXH3;
  t.no_block_cache = true;
  t.enable_index_compression = o.index_compression != 0;
  if (o.bloom_bits > 0) t.filter_policy.reset(NewBloomFilterPolicy(o.bloom_bits, false));
  if (o.ribbon) t.filter_policy.reset(NewRibbonFilterPolicy(10));
  if (o.partition_filters) {
    t.partition_filters = true;
    t.index_type = BlockBasedTableOptions::kTwoLevelIndexSearch;
  }
  opt.table_factory.reset(NewBlockBasedTableFactory(t));
  if (o.filter == "remove_empty_value") opt.compaction_filter_factory = std::make_shared<RemoveEmptyValueFactory>();
  else if (o.filter != "none") {
    fprintf(stderr, "ref_compact: unknown filter %s\n", o.filter.c_str());
    return 1;
  }
  auto listener = std::make_shared<StatsListener>();
  opt.listeners.push_back(listener);
  bool use_b200 = false;
#ifdef WITH_B200_PLUGIN
  std::shared_ptr<TableFactory> b200_tf;
  if (o.table_factory == "b200" || o.table_factory == "b200+nofallback") {
    B200TableFactoryOptions to;
    to.allow_fallback = o.table_factory == "b200";
    b200_tf = NewB200TableFactory(t, to);
    opt.table_factory = b200_tf;
  }
#else
  if (!o.table_factory.empty()) {
    fprintf(stderr, "ref_compact: table_factory=%s needs the ref_compact_b200 build\n", o.table_factory.c_str());
    return 1;
  }
#endif
#ifdef WITH_B200_PLUGIN
  if (o.executor == "b200" || o.executor == "b200+fallback") {
    opt.statistics = CreateDBStatistics();
    B200CompactOptions bo;
    // "b200": a job the device path rejects fails the compaction (the tests want to see the device path, not a silent
    // fallback); "b200+fallback": the reference re-runs such a job on its own CPU path (compaction_job.cc RunRemote -> RunLocal)
    bo.allow_fallback_to_local = o.executor == "b200+fallback";
    bo.max_subcompactions = o.b200_subs;  // 0: what the job asks for (max_subcompactions); 1: never split
    for (size_t a = 0; a < o.b200_devices.size();) {  // "0,1": the devices the ranges of one job are dealt to
      size_t b = o.b200_devices.find(',', a);
      if (b == std::string::npos) b = o.b200_devices.size();
      bo.devices.push_back(atoi(o.b200_devices.substr(a, b - a).c_str()));
      a = b + 1;
    }
    opt.compaction_executor_factory = NewB200CompactionExecutorFactory(bo);
    use_b200 = true;
  }
#endif
  if (!o.executor.empty() && !use_b200) {
    fprintf(stderr, "ref_compact: executor=%s needs the ref_compact_b200 build\n", o.executor.c_str());
    return 1;
  }

  std::shared_ptr<ScriptClock> clock;
  std::unique_ptr<Env> clock_env;
  if (o.ttl > 0) {
    clock = std::make_shared<ScriptClock>(SystemClock::Default());
    clock_env.reset(new CompositeEnvWrapper(Env::Default(), clock));
    opt.env = clock_env.get();
  }
  DestroyDB(dbdir, opt).PermitUncheckedError();
  DB* db = nullptr;
  Status s;
  if (o.ttl > 0) {
    DBWithTTL* tdb = nullptr;
    s = DBWithTTL::Open(opt, dbdir, &tdb, o.ttl);
    db = tdb;
  } else {
    s = DB::Open(opt, dbdir, &db);
  }
  if (!s.ok()) Die("open", s);

  auto load_ops = [&](DB* db, std::vector<const Snapshot*>& snaps) -> int {
  FILE* f = fopen(argv[1], "rb");
  if (!f) {
    perror(argv[1]);
    return 2;
  }
  char magic[8];
  if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "B2OPS\0\0\1", 8) != 0) {
    fprintf(stderr, "ref_compact: bad ops magic\n");
    return 2;
  }
  Reader rd{f};
  WriteOptions wo;
  wo.disableWAL = true;
  WriteBatch batch;
  size_t batch_n = 0;
  auto flush_batch = [&]() {
    if (batch_n) {
      Status ws = db->Write(wo, &batch);
      if (!ws.ok()) Die("write", ws);
      batch.Clear();
      batch_n = 0;
    }
  };
  std::string key, val;
  uint8_t op;
  while (rd.u8(&op) && op != 0) {
    switch (op) {
      case 1: {
        uint32_t kl = rd.u32(), vl = rd.u32();
        rd.bytes(&key, kl);
        rd.bytes(&val, vl);
        batch.Put(key, val);
        if (++batch_n >= 4096) flush_batch();
        break;
      }
      case 2: {
        uint32_t kl = rd.u32();
        rd.bytes(&key, kl);
        batch.Delete(key);
        if (++batch_n >= 4096) flush_batch();
        break;
      }
      case 7: {  // WriteBatch::SingleDelete
        uint32_t kl = rd.u32();
        rd.bytes(&key, kl);
        batch.SingleDelete(key);
        if (++batch_n >= 4096) flush_batch();
        break;
      }
      case 3: {
        flush_batch();
        FlushOptions fo;
        fo.wait = true;
        s = db->Flush(fo);
        if (!s.ok()) Die("flush", s);
        break;
      }
      case 4:
        flush_batch();
        snaps.push_back(db->GetSnapshot());
        break;
      case 5: {
        uint8_t lvl;
        rd.u8(&lvl);
        flush_batch();
        std::vector<LiveFileMetaData> live;
        db->GetLiveFilesMetaData(&live);
        std::vector<std::string> names;
        for (auto& m : live) names.push_back(m.name);
        if (!names.empty()) {
          CompactionOptions co;
          co.compression = in_comp;  // set-up compactions write inputs of the measured job
          co.output_file_size_limit = o.setup_file_size;
          s = db->CompactFiles(co, names, lvl);
          if (!s.ok()) Die("setup compaction", s);
        }
        break;
      }
      case 6: {  // SET_TIME: seconds since the epoch seen by the DB from now on (ttl mode)
        flush_batch();
        uint32_t t = rd.u32();
        if (clock) clock->now = t;
        break;
      }
      default:
        fprintf(stderr, "ref_compact: bad op %u\n", op);
        return 2;
    }
  }
  flush_batch();
  fclose(f);
  return 0;
  };

  if (o.warm) {
    // warm=1: the same script and job once through a throw-away DB first (same Options object, so the same executor / table factory):
    // the device context, the kernels' module and the library's buffer cache are those of a process that has compacted before,
    // which is what a DB sees from its second compaction on
    const std::string wdir = work + "/warmdb";
    DestroyDB(wdir, opt).PermitUncheckedError();
    DB* wdb = nullptr;
    Status ws = DB::Open(opt, wdir, &wdb);
    if (!ws.ok()) Die("open (warm)", ws);
    std::vector<const Snapshot*> wsnaps;
    if (int rc = load_ops(wdb, wsnaps)) return rc;
    std::vector<LiveFileMetaData> live;
    wdb->GetLiveFilesMetaData(&live);
    std::vector<std::string> names;
    for (auto& m : live)
      if (m.level == 0) names.push_back(m.name);
    if (!names.empty()) {
      CompactionOptions wco;
      wco.compression = kNoCompression;
      wco.output_file_size_limit = o.target_file_size;
      wco.max_subcompactions = o.max_subcompactions;
      ws = wdb->CompactFiles(wco, names, o.output_level);
      if (!ws.ok()) Die("warm compaction", ws);
    }
    for (auto* sn : wsnaps) wdb->ReleaseSnapshot(sn);
    delete wdb;
    DestroyDB(wdir, opt).PermitUncheckedError();
  }
  std::vector<const Snapshot*> snaps;
  if (int rc = load_ops(db, snaps)) return rc;

  // The job: every L0 file -> output_level.  Input order = newest L0 file first, which is the
  // order VersionSet::MakeInputIterator hands the children to the merging iterator
  // (db/version_set.cc:7298-7320).
  ColumnFamilyMetaData cfm;
  db->GetColumnFamilyMetaData(&cfm);
  std::vector<std::string> input_names;
  std::vector<SstFileMetaData> inputs;
  bool deeper_files = false;
  for (auto& lvl : cfm.levels) {
    if (lvl.level == 0) {
      for (auto& fm : lvl.files) {
        inputs.push_back(fm);
        input_names.push_back(fm.name);
      }
    } else if (lvl.level > o.output_level && !lvl.files.empty()) {
      deeper_files = true;
    } else if (lvl.level > 0 && lvl.level <= o.output_level && !lvl.files.empty()) {
      fprintf(stderr, "ref_compact: script left files at L%d (<= output level)\n", lvl.level);
      return 2;
    }
  }
  if (inputs.empty()) {
    fprintf(stderr, "ref_compact: no L0 files to compact\n");
    return 2;
  }
  if (o.copy)
    for (auto& fm : inputs) CopyFile(dbdir + fm.name, work + "/inputs" + fm.name);

  if (o.barrier_n > 1 && !o.barrier_dir.empty()) {
    const std::string mine = o.barrier_dir + "/ready." + std::to_string((long)getpid());
    fclose(fopen(mine.c_str(), "w"));
    for (int spins = 0; spins < 600000; spins++) {  // <= 10 minutes
      int n = 0;
      if (DIR* d = opendir(o.barrier_dir.c_str())) {
        while (dirent* e = readdir(d))
          if (strncmp(e->d_name, "ready.", 6) == 0) n++;
        closedir(d);
      }
      if (n >= o.barrier_n) break;
      usleep(1000);
    }
  }
  std::vector<SstFileMetaData> grandparents;  // what CompactionPicker::GetGrandparents would pick: files one level below the output
  for (auto& lvl : cfm.levels)
    if (lvl.level == o.output_level + 1)
      for (auto& fm : lvl.files) grandparents.push_back(fm);
  const bool range_mode = o.mode == "range";
  CompactionOptions co;
  co.compression = kNoCompression;
  co.output_file_size_limit = o.target_file_size;
  co.max_subcompactions = o.max_subcompactions;
  std::vector<std::string> out_names;
  CompactionJobInfo ji;
  auto t0 = std::chrono::steady_clock::now();
  if (range_mode) {
    listener->capture_level = o.output_level;
    listener->copy = o.copy != 0;
    listener->dbdir = dbdir;
    listener->outdir = work + "/outputs";
    CompactRangeOptions cro;
    cro.exclusive_manual_compaction = true;
    cro.bottommost_level_compaction = BottommostLevelCompaction::kSkip;
    s = db->CompactRange(cro, nullptr, nullptr);
    if (s.ok() && !listener->captured) {
      fprintf(stderr, "ref_compact: CompactRange ran no L0 -> L%d compaction (trivial move?)\n", o.output_level);
      return 2;
    }
    ji = listener->job;
  } else {
    grandparents.clear();  // DB::CompactFiles builds the compaction without grandparents (compaction_picker.cc CompactFiles)
    s = db->CompactFiles(co, input_names, o.output_level, -1, &out_names, &ji);
  }
  auto t1 = std::chrono::steady_clock::now();
  if (!s.ok()) Die("compaction", s);
  double wall_us = std::chrono::duration<double, std::micro>(t1 - t0).count();

  std::string db_id, session_id;
  db->GetDbIdentity(db_id).PermitUncheckedError();
  db->GetDbSessionId(session_id).PermitUncheckedError();

  FILE* m = fopen((work + "/manifest.json").c_str(), "w");
  fprintf(m, "{\n  \"reference\": \"toplingdb (rocksdb %d.%d.%d)\",\n", ROCKSDB_MAJOR, ROCKSDB_MINOR,
          ROCKSDB_PATCH);
  fprintf(m, "  \"output_level\": %d,\n  \"target_file_size\": %" PRIu64 ",\n", o.output_level,
          o.target_file_size);
  fprintf(m, "  \"block_size\": %" PRIu64 ",\n  \"restart_interval\": %d,\n  \"format_version\": %u,\n",
          o.block_size, o.restart_interval, o.format_version);
  fprintf(m, "  \"checksum\": \"%s\",\n  \"max_subcompactions\": %u,\n", o.checksum.c_str(),
          o.max_subcompactions);
  // BloomFilterPolicy keeps bits_per_key as millibits (filter_policy.cc: round(bits_per_key * 1000), at least 1000)
  fprintf(m, "  \"bloom_millibits_per_key\": %d,\n", o.bloom_bits > 0 ? std::max(1000, (int)(o.bloom_bits * 1000.0 + 0.500001)) : 0);
  fprintf(m, "  \"bottommost_level\": %s,\n", deeper_files ? "false" : "true");
  {
    // Compaction::max_output_file_size_ (compaction.cc:289-295): twice the target when the job has grandparents
    // range mode takes the size from target_file_size_base, which the reference clamps to >= 512 KiB (options/cf_options.cc:1030-1033)
    const uint64_t eff_target = range_mode ? std::max<uint64_t>(o.target_file_size, 512 << 10) : o.target_file_size;
    const bool doubled = !grandparents.empty() && deeper_files && opt.level_compaction_dynamic_file_size;
    fprintf(m, "  \"mode\": \"%s\",\n  \"max_output_file_size\": %" PRIu64 ",\n  \"max_compaction_bytes\": %" PRIu64 ",\n",
            o.mode.c_str(), doubled ? 2 * eff_target : eff_target,
            opt.max_compaction_bytes ? opt.max_compaction_bytes : eff_target * 25);
    fprintf(m, "  \"target_output_file_size\": %" PRIu64 ",\n", eff_target);
    fprintf(m, "  \"level_compaction_dynamic_file_size\": %s,\n  \"grandparents\": [", opt.level_compaction_dynamic_file_size ? "true" : "false");
    for (size_t i = 0; i < grandparents.size(); i++)
      fprintf(m, "%s\n    {\"smallestkey\": \"%s\", \"largestkey\": \"%s\", \"size\": %" PRIu64 "}", i ? "," : "",
              Hex(grandparents[i].smallestkey).c_str(), Hex(grandparents[i].largestkey).c_str(), grandparents[i].size);
    fprintf(m, "%s],\n", grandparents.empty() ? "" : "\n  ");
  }
  fprintf(m, "  \"compaction_filter\": \"%s\",\n", o.ttl > 0 ? "ttl" : o.filter.c_str());
  fprintf(m, "  \"ttl\": %d,\n  \"now\": %lld,\n", o.ttl, clock ? (long long)clock->now : 0ll);
  {
    uint64_t remote_read = 0;
#ifdef WITH_B200_PLUGIN
    if (opt.statistics) remote_read = opt.statistics->getTickerCount(REMOTE_COMPACT_READ_BYTES);
#endif
    fprintf(m, "  \"executor\": \"%s\",\n  \"remote_compact_read_bytes\": %" PRIu64 ",\n", use_b200 ? "B200Compact" : "local", remote_read);
#ifdef WITH_B200_PLUGIN
    if (b200_tf) {
      auto* tf = static_cast<B200TableFactory*>(b200_tf.get());
      fprintf(m, "  \"table_factory\": \"%s\",\n  \"b200_device_tables\": %" PRIu64 ",\n  \"b200_fallback_tables\": %" PRIu64 ",\n", tf->Name(),
              tf->device_tables(), tf->fallback_tables());
    }
#endif
  }
  {
    // Read the whole DB back through the reference's own table reader (block checksums verified): a digest of what a
    // user sees after the job, identical whichever executor produced the files.
    ReadOptions ro;
    ro.verify_checksums = true;
    ro.fill_cache = false;
    std::unique_ptr<Iterator> it(db->NewIterator(ro));
    uint64_t n = 0, h = 1469598103934665603ull;
    auto mix = [&h](const Slice& x) {
      for (size_t i = 0; i < x.size(); i++) h = (h ^ (unsigned char)x[i]) * 1099511628211ull;
      h = (h ^ x.size()) * 1099511628211ull;
    };
    // with a filter policy: every key the scan shows is also looked up with Get(), which consults the files' Bloom filters first -- a
    // filter with a missing bit turns into NotFound here (the reference counts useful / positive filter probes in its tickers)
    uint64_t get_found = 0, get_missing = 0;
    std::string got;
    for (it->SeekToFirst(); it->Valid(); it->Next()) {
      mix(it->key());
      mix(it->value());
      n++;
      if (o.bloom_bits > 0) {
        Status gs = db->Get(ro, it->key(), &got);
        if (gs.ok() && Slice(got) == it->value()) get_found++;
        else get_missing++;
      }
    }
    if (!it->status().ok()) Die("scan after compaction", it->status());
    fprintf(m, "  \"scan_count\": %" PRIu64 ",\n  \"scan_digest\": \"%016" PRIx64 "\",\n", n, h);
    fprintf(m, "  \"get_found\": %" PRIu64 ",\n  \"get_missing\": %" PRIu64 ",\n", get_found, get_missing);
  }
  fprintf(m, "  \"db_id\": \"%s\",\n  \"db_session_id\": \"%s\",\n", db_id.c_str(), session_id.c_str());
  fprintf(m, "  \"snapshots\": [");
  for (size_t i = 0; i < snaps.size(); i++)
    fprintf(m, "%s%" PRIu64, i ? ", " : "", snaps[i]->GetSequenceNumber());
  fprintf(m, "],\n  \"inputs\": [\n");
  for (size_t i = 0; i < inputs.size(); i++) {
    auto& fm = inputs[i];
    fprintf(m,
            "    {\"name\": \"%s\", \"level\": 0, \"size\": %" PRIu64 ", \"file_number\": %" PRIu64
            ", \"smallest_seqno\": %" PRIu64 ", \"largest_seqno\": %" PRIu64 ", \"num_entries\": %" PRIu64
            ", \"num_deletions\": %" PRIu64 "}%s\n",
            fm.name.c_str(), fm.size, fm.file_number, fm.smallest_seqno, fm.largest_seqno, fm.num_entries,
            fm.num_deletions, i + 1 < inputs.size() ? "," : "");
  }
  fprintf(m, "  ],\n  \"outputs\": [\n");
  db->GetColumnFamilyMetaData(&cfm);
  std::vector<SstFileMetaData> outs;
  if (range_mode) outs = listener->outs;  // captured (and copied) when the job completed
  else
    for (auto& lvl : cfm.levels)
      if (lvl.level == o.output_level)
        for (auto& fm : lvl.files) outs.push_back(fm);
  for (size_t i = 0; i < outs.size(); i++) {
    auto& fm = outs[i];
    if (o.copy && !range_mode) CopyFile(dbdir + fm.name, work + "/outputs" + fm.name);
    fprintf(m,
            "    {\"name\": \"%s\", \"size\": %" PRIu64 ", \"file_number\": %" PRIu64
            ", \"smallest_seqno\": %" PRIu64 ", \"largest_seqno\": %" PRIu64 ", \"num_entries\": %" PRIu64
            ", \"num_deletions\": %" PRIu64 ", \"smallestkey\": \"%s\", \"largestkey\": \"%s\"}%s\n",
            fm.name.c_str(), fm.size, fm.file_number, fm.smallest_seqno, fm.largest_seqno, fm.num_entries,
            fm.num_deletions, Hex(fm.smallestkey).c_str(), Hex(fm.largestkey).c_str(),
            i + 1 < outs.size() ? "," : "");
  }
  fprintf(m, "  ],\n  \"subcompactions\": [");
  {
    std::vector<StatsListener::Sub> subs;
    for (auto& sb : listener->subs)
      if (sb.job_id == ji.job_id && sb.sub_id >= 0) subs.push_back(sb);
    std::sort(subs.begin(), subs.end(), [](const StatsListener::Sub& a, const StatsListener::Sub& b) { return a.sub_id < b.sub_id; });
    for (size_t i = 0; i < subs.size(); i++) {
      const CompactionJobStats& ss = subs[i].st;
      fprintf(m,
              "%s\n    {\"id\": %d, \"num_input_records\": %" PRIu64 ", \"num_output_records\": %" PRIu64 ", \"num_output_files\": %" PRIu64
              ", \"num_input_deletion_records\": %" PRIu64 ", \"num_expired_deletion_records\": %" PRIu64
              ", \"num_records_replaced\": %" PRIu64 ", \"total_input_raw_key_bytes\": %" PRIu64
              ", \"total_input_raw_value_bytes\": %" PRIu64 "}",
              i ? "," : "", subs[i].sub_id, ss.num_input_records, ss.num_output_records, (uint64_t)ss.num_output_files,
              ss.num_input_deletion_records, ss.num_expired_deletion_records, ss.num_records_replaced, ss.total_input_raw_key_bytes,
              ss.total_input_raw_value_bytes);
    }
    fprintf(m, "%s],\n", subs.empty() ? "" : "\n  ");
  }
  const CompactionJobStats& st = ji.stats;
  fprintf(m, "  \"stats\": {\n");
  fprintf(m, "    \"wall_micros\": %.0f,\n    \"elapsed_micros\": %" PRIu64 ",\n    \"cpu_micros\": %" PRIu64 ",\n",
          wall_us, st.elapsed_micros, st.cpu_micros);
  fprintf(m, "    \"num_input_records\": %" PRIu64 ",\n    \"num_output_records\": %" PRIu64 ",\n",
          st.num_input_records, st.num_output_records);
  fprintf(m, "    \"total_input_bytes\": %" PRIu64 ",\n    \"total_output_bytes\": %" PRIu64 ",\n",
          st.total_input_bytes, st.total_output_bytes);
  fprintf(m, "    \"num_records_replaced\": %" PRIu64 ",\n    \"total_input_raw_key_bytes\": %" PRIu64 ",\n",
          st.num_records_replaced, st.total_input_raw_key_bytes);
  fprintf(m, "    \"total_input_raw_value_bytes\": %" PRIu64 ",\n    \"num_input_deletion_records\": %" PRIu64 ",\n",
          st.total_input_raw_value_bytes, st.num_input_deletion_records);
  fprintf(m, "    \"num_expired_deletion_records\": %" PRIu64 ",\n    \"num_corrupt_keys\": %" PRIu64 "\n  }\n}\n",
          st.num_expired_deletion_records, st.num_corrupt_keys);
  fclose(m);

  for (auto* sn : snaps) db->ReleaseSnapshot(sn);
  s = db->Close();
  delete db;
  if (!o.keep_db) DestroyDB(dbdir, opt).PermitUncheckedError();
  printf("ok inputs=%zu outputs=%zu elapsed_us=%" PRIu64 " cpu_us=%" PRIu64 " wall_us=%.0f\n", inputs.size(),
         outs.size(), st.elapsed_micros, st.cpu_micros, wall_us);
  return 0;
}
