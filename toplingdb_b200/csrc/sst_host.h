// toplingdb_b200/csrc/sst_host.h — host-side pieces of the BlockBasedTable format that are O(1) per file:
// footer / metaindex / properties parsing for inputs, and properties / metaindex / footer construction for
// outputs (the "tail" of a file, ~1 KB).  Everything per-entry or per-block runs on the device.
// Reference: table/format.cc:191-259 (footer), table/meta_blocks.cc:35-175 (metaindex, properties),
// table/block_based/block_based_table_builder.cc:1605-1764,1921-1977 (WritePropertiesBlock, WriteFooter, Finish).
#pragma once
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

namespace b200c {

struct InputTail {            // what the decoder needs to know about one input file
  uint32_t checksum_type = 0, format_version = 0;
  uint64_t index_off = 0, index_size = 0;
  uint64_t meta_off = 0, meta_size = 0;
  uint64_t props_off = 0, props_size = 0;
  bool has_range_del = false, has_filter = false, has_dict = false;
  uint64_t num_entries = 0, num_data_blocks = 0, raw_key_size = 0, raw_value_size = 0, num_range_deletions = 0,
           num_merge_operands = 0, data_size = 0;
  uint64_t index_key_is_user_key = 0;  // rocksdb.index.key.is.user.key: index separators carry no 8-byte trailer
  uint32_t index_type = 0;             // rocksdb.block.based.table.index.type (BlockBasedTableOptions::IndexType); 0 = kBinarySearch
  std::string compression_name, comparator_name;
};

// footer only (last 53 bytes).  Returns "" or an error message.
std::string parse_footer(const uint8_t* file_end_minus_53, uint64_t file_len, InputTail* t);
// a metaindex block payload (without trailer) -> name -> (offset,size)
std::string parse_metaindex(const uint8_t* blk, uint64_t size, std::map<std::string, std::pair<uint64_t, uint64_t>>* out);
// a properties block payload -> fills the numeric fields of t
std::string parse_properties(const uint8_t* blk, uint64_t size, InputTail* t);

uint32_t host_block_checksum(uint32_t type, const uint8_t* data, uint64_t n, uint8_t last_byte);
uint64_t host_xxh3_64(const uint8_t* data, uint64_t n);

struct OutputTailInput {      // per output file
  uint32_t checksum_type, format_version;
  uint64_t data_size, index_size /* payload, no trailer */;
  uint64_t filter_size = 0 /* filter block content (bits + metadata, no trailer); 0 = no filter block */, filter_entries = 0;
  uint64_t num_entries, num_deletions, raw_key_size, raw_value_size, num_data_blocks;
  bool index_key_is_user_key;
  uint32_t column_family_id;
  std::string column_family_name, db_id, db_session_id, db_host_id;
  uint64_t creation_time, oldest_key_time, file_creation_time, orig_file_number;
};
// bytes that follow the index block trailer: properties block + trailer, metaindex block + trailer, footer
std::vector<uint8_t> build_output_tail(const OutputTailInput& in);

}  // namespace b200c
