// tests/native/tail_sim.cc — test infrastructure: the product's host-side table code (toplingdb_b200/csrc/sst_host.cc: footer /
// metaindex / properties parsing of the inputs, properties + metaindex + footer construction of the outputs) compiled on its own for
// the CPU.  tests/test_sst_host.py feeds it the numbers of files the reference wrote and expects the reference's bytes back.
#include "../../toplingdb_b200/csrc/sst_host.cc"

using namespace b200c;

extern "C" {

// the tail (everything behind the index block trailer) of an output file with these properties; returns its length
uint64_t tail_sim_build(uint32_t cksum, uint32_t format_version, uint64_t data_size, uint64_t index_size, uint64_t filter_size,
                        uint64_t filter_entries, uint64_t num_entries, uint64_t num_deletions, uint64_t raw_key_size,
                        uint64_t raw_value_size, uint64_t num_data_blocks, int index_key_is_user_key, uint32_t cf_id, const char* cf_name,
                        const char* db_id, const char* db_session_id, const char* db_host_id, uint64_t creation_time,
                        uint64_t oldest_key_time, uint64_t file_creation_time, uint64_t orig_file_number, uint8_t* out, uint64_t cap) {
  OutputTailInput in;
  in.checksum_type = cksum;
  in.format_version = format_version;
  in.data_size = data_size;
  in.index_size = index_size;
  in.filter_size = filter_size;
  in.filter_entries = filter_entries;
  in.num_entries = num_entries;
  in.num_deletions = num_deletions;
  in.raw_key_size = raw_key_size;
  in.raw_value_size = raw_value_size;
  in.num_data_blocks = num_data_blocks;
  in.index_key_is_user_key = index_key_is_user_key != 0;
  in.column_family_id = cf_id;
  in.column_family_name = cf_name;
  in.db_id = db_id;
  in.db_session_id = db_session_id;
  in.db_host_id = db_host_id;
  in.creation_time = creation_time;
  in.oldest_key_time = oldest_key_time;
  in.file_creation_time = file_creation_time;
  in.orig_file_number = orig_file_number;
  std::vector<uint8_t> t = build_output_tail(in);
  if (t.size() <= cap) memcpy(out, t.data(), t.size());
  return t.size();
}

// parses the tail of a whole file image the way the job does for its inputs; fields[]: index_off, index_size, props_off, props_size,
// num_entries, num_data_blocks, raw_key_size, raw_value_size, data_size, checksum_type, format_version, has_filter.  0 = ok.
int tail_sim_parse(const uint8_t* file, uint64_t len, uint64_t* fields) {
  InputTail t;
  if (len < 53 || !parse_footer(file + len - 53, len, &t).empty()) return 1;
  std::map<std::string, std::pair<uint64_t, uint64_t>> meta;
  if (t.meta_off + t.meta_size > len || !parse_metaindex(file + t.meta_off, t.meta_size, &meta).empty()) return 2;
  auto it = meta.find("rocksdb.properties");
  if (it == meta.end()) return 3;
  t.props_off = it->second.first;
  t.props_size = it->second.second;
  if (t.props_off + t.props_size > len || !parse_properties(file + t.props_off, t.props_size, &t).empty()) return 4;
  const uint64_t v[12] = {t.index_off, t.index_size, t.props_off, t.props_size, t.num_entries, t.num_data_blocks, t.raw_key_size,
                          t.raw_value_size, t.data_size, t.checksum_type, t.format_version,
                          (uint64_t)(meta.count("fullfilter.rocksdb.BuiltinBloomFilter") != 0)};
  memcpy(fields, v, sizeof v);
  return 0;
}

}  // extern "C"
