// oracle/ref_sst_check.cc — test infrastructure.  What the reference's stock tools do with a table file, done through the reference's
// own classes (compiled into oracle/_ref/libtoplingdb_ref.so): SstFileDumper is what `sst_dump --command=verify|check|scan` drives
// (tools/sst_dump_tool.cc), SstFileReader is the public stand-alone reader (include/rocksdb/sst_file_reader.h).  Prints one JSON line:
// entries seen by both readers, a digest of the scan, and the table properties a tool would show.
//   ref_sst_check <file.sst>
#include <cinttypes>
#include <cstdio>
#include <memory>
#include <string>

#include "rocksdb/options.h"
#include "rocksdb/sst_file_reader.h"
#include "rocksdb/table_properties.h"
#include "table/sst_file_dumper.h"

using namespace ROCKSDB_NAMESPACE;

static int Fail(const char* what, const Status& s) {
  fprintf(stderr, "ref_sst_check: %s: %s\n", what, s.ToString().c_str());
  return 1;
}

int main(int argc, char** argv) {
  if (argc != 2) {
    fprintf(stderr, "usage: ref_sst_check <file.sst>\n");
    return 2;
  }
  const std::string file = argv[1];
  Options opt;
  // sst_dump: open, verify every block checksum, read the whole file sequentially
  SstFileDumper dumper(opt, file, Temperature::kUnknown, 2 << 20 /*readahead*/, true /*verify_checksum*/, false /*output_hex*/,
                       false /*decode_blob_index*/, EnvOptions(), true /*silent*/);
  Status s = dumper.getStatus();
  if (!s.ok()) return Fail("open (SstFileDumper)", s);
  s = dumper.VerifyChecksum();
  if (!s.ok()) return Fail("VerifyChecksum", s);
  s = dumper.ReadSequential(false, (uint64_t)-1, false, "", false, "");
  if (!s.ok()) return Fail("ReadSequential", s);
  const uint64_t dumped = dumper.GetReadNumber();
  // SstFileReader: the public reader (user keys and values as an application sees them)
  SstFileReader reader(opt);
  s = reader.Open(file);
  if (!s.ok()) return Fail("SstFileReader::Open", s);
  s = reader.VerifyChecksum();
  if (!s.ok()) return Fail("SstFileReader::VerifyChecksum", s);
  ReadOptions ro;
  ro.verify_checksums = true;
  std::unique_ptr<Iterator> it(reader.NewIterator(ro));
  uint64_t n = 0, h = 1469598103934665603ull;
  auto mix = [&h](const Slice& x) {
    for (size_t i = 0; i < x.size(); i++) h = (h ^ (unsigned char)x[i]) * 1099511628211ull;
    h = (h ^ x.size()) * 1099511628211ull;
  };
  for (it->SeekToFirst(); it->Valid(); it->Next()) {
    mix(it->key());
    mix(it->value());
    n++;
  }
  if (!it->status().ok()) return Fail("scan", it->status());
  std::shared_ptr<const TableProperties> tp = reader.GetTableProperties();
  printf("{\"dumper_entries\": %" PRIu64 ", \"reader_entries\": %" PRIu64 ", \"reader_digest\": \"%016" PRIx64
         "\", \"num_entries\": %" PRIu64 ", \"num_deletions\": %" PRIu64 ", \"num_data_blocks\": %" PRIu64 ", \"data_size\": %" PRIu64
         ", \"index_size\": %" PRIu64 ", \"filter_size\": %" PRIu64 ", \"num_filter_entries\": %" PRIu64
         ", \"filter_policy_name\": \"%s\", \"format_version\": %" PRIu64 "}\n",
         dumped, n, h, tp->num_entries, tp->num_deletions, tp->num_data_blocks, tp->data_size, tp->index_size, tp->filter_size,
         tp->num_filter_entries, tp->filter_policy_name.c_str(), tp->format_version);
  return 0;
}
