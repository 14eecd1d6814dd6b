/* oracle/compaction_oracle.h — CPU restatement of ToplingDB's compaction hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the checker the CUDA path is compared against; nothing under
 * toplingdb_b200/ may include, link or call it.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it.
 *
 * Parity status: PINNED.  The restatement is checked (tests/test_oracle_vs_reference.py and the
 * committed fixtures under tests/golden/) against (a) the reference's own known-answer vectors
 * (table/table_test.cc:2303-2389 block checksums, util/crc32c_test.cc:67, util/coding_test.cc,
 * db/compaction/compaction_iterator_test.cc drop-rule vectors) and (b) byte-for-byte against SST files
 * produced by the unmodified reference compiled into oracle/_ref (see oracle/Makefile, ref_compact.cc).
 *
 * Every function cites the reference file:line it restates (paths relative to /root/reference).
 */
#ifndef COMPACTION_ORACLE_H_
#define COMPACTION_ORACLE_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_TYPE_DELETION = 0, ORC_TYPE_VALUE = 1, ORC_TYPE_MERGE = 2, ORC_TYPE_SINGLE_DELETION = 7 };
enum { ORC_CKSUM_NONE = 0, ORC_CKSUM_CRC32C = 1, ORC_CKSUM_XXH3 = 4 };
#define ORC_MAX_SEQ ((1ull << 56) - 1) /* kMaxSequenceNumber, db/dbformat.h:96 */
#define ORC_VALUE_TYPE_FOR_SEEK 0x16ull  /* kValueTypeForSeek = kTypeWideColumnEntity, db/dbformat.cc:29 */

/* Job description: the subset of CompactionParams (db/compaction/compaction_executor.h:33-118),
 * BlockBasedTableOptions (include/rocksdb/table.h:237-564) and TableBuilderOptions
 * (db/compaction/compaction_job.cc:2323-2331) that shapes the output bytes. */
typedef struct orc_grandparent { /* one file of the level below the output: FileMetaData::{smallest,largest}.user_key(), fd.GetFileSize() */
  const uint8_t* smallest;
  uint32_t smallest_len;
  const uint8_t* largest;
  uint32_t largest_len;
  uint64_t file_size;
} orc_grandparent;

typedef struct orc_params {
  int32_t output_level;
  int32_t bottommost_level;        /* Compaction::bottommost_level() */
  uint64_t max_output_file_size;   /* compaction.cc:291-295 */
  uint32_t block_size;             /* 4096 */
  uint32_t block_size_deviation;   /* 10 */
  uint32_t block_restart_interval; /* 16 */
  uint32_t index_block_restart_interval; /* 1 */
  uint32_t format_version;         /* 5 */
  uint32_t checksum_type;          /* ORC_CKSUM_* */
  const uint64_t* snapshots;       /* ascending (existing_snapshots) */
  uint32_t num_snapshots;
  uint32_t column_family_id;
  const char* column_family_name;
  const char* db_id;
  const char* db_session_id;
  const char* db_host_id;
  uint64_t creation_time;          /* rocksdb.creation.time (oldest ancestor time) */
  uint64_t oldest_key_time;
  const uint64_t* file_creation_times; /* one per output file; last repeats */
  uint32_t num_file_creation_times;
  uint64_t first_file_number;      /* outputs are numbered consecutively from here */
  uint32_t compaction_filter;      /* ORC_FILTER_*: built-in CompactionFilter applied by the iterator (compaction_iterator.cc:231-473) */
  int32_t ttl;                     /* ORC_FILTER_TTL: seconds (<= 0: nothing is stale, db_ttl_impl.cc:445-461) */
  int64_t now;                     /* ORC_FILTER_TTL: the clock reading the filter compares against */
  /* grandparent-aware output cutting, CompactionOutputs::ShouldStopBefore (compaction_outputs.cc:231-354) */
  const orc_grandparent* grandparents; /* Compaction::grandparents(), sorted; NULL / 0 = none */
  uint32_t num_grandparents;
  uint32_t level_compaction_dynamic_file_size; /* ImmutableOptions (default true) */
  uint64_t max_compaction_bytes;       /* Compaction::max_compaction_bytes() */
  uint64_t target_output_file_size;    /* Compaction::target_output_file_size() (max_output_file_size is twice this with grandparents) */
  /* sub-compaction key range (SubcompactionState::start / end, compaction_job.cc:1433-1519): the merged input is clipped to
   * start <= user key < end by a ClippingIterator before CompactionIterator sees it; has_* == 0: unbounded on that side */
  const uint8_t* range_start;
  uint32_t range_start_len, has_range_start;
  const uint8_t* range_end;
  uint32_t range_end_len, has_range_end;
  /* BlockBasedTableOptions::filter_policy = NewBloomFilterPolicy(bits): BloomLikeFilterPolicy::millibits_per_key_
   * (table/block_based/filter_policy.cc:1327-1343); 0 = no filter block.  Full filter over whole user keys, format_version >= 5
   * (FastLocalBloom). */
  uint32_t bloom_millibits_per_key;
  /* CompactionIterator's earliest_write_conflict_snapshot (kMaxSequenceNumber unless a transaction DB holds write-conflict snapshots);
   * 0 = kMaxSequenceNumber.  Only the SingleDelete rule reads it (compaction_iterator.cc:803-808). */
  uint64_t earliest_write_conflict_snapshot;
  /* Compaction::KeyNotExistsBeyondOutputLevel (db/compaction/compaction.cc:548-586).  0 (default): a compaction worker's answer --
   * true only at the bottommost level (:553-556), which is what the executor plugin's jobs get.  1: the answer of a job the DB runs
   * itself: at a non-bottommost level the key may exist beyond the output level iff it falls into the user-key range of one of the
   * files of the deeper levels, listed in deeper_files (orc_grandparent: smallest / largest user key; file_size unused). */
  uint32_t key_not_exists_mode;
  const orc_grandparent* deeper_files;
  uint32_t num_deeper_files;
} orc_params;
#define ORC_FILTER_NONE 0
#define ORC_FILTER_REMOVE_EMPTY_VALUE 1 /* utilities/compaction_filters/remove_emptyvalue_compactionfilter.cc:15-22 */
#define ORC_FILTER_TTL 2                /* TtlCompactionFilter without a user filter, utilities/ttl/db_ttl_impl.cc:200-206,445-461 */

typedef struct orc_file_meta { /* FileMinMeta, compaction_executor.h:120-131 + table properties */
  uint64_t file_number, file_size;
  uint64_t smallest_seqno, largest_seqno;
  uint64_t num_entries, num_deletions, raw_key_size, raw_value_size, num_data_blocks;
  uint32_t smallest_len, largest_len;
  uint8_t smallest[256], largest[256]; /* internal keys (truncated copies if longer) */
} orc_file_meta;

typedef struct orc_stats { /* CompactionJobStats subset, include/rocksdb/compaction_job_stats.h */
  uint64_t num_input_records, num_output_records;
  uint64_t num_input_deletion_records;
  uint64_t num_records_replaced;          /* num_record_drop_hidden */
  uint64_t num_expired_deletion_records;  /* num_record_drop_obsolete */
  uint64_t total_input_raw_key_bytes, total_input_raw_value_bytes;
  uint64_t num_optimized_del_drop_obsolete;
  uint64_t num_record_drop_user;          /* entries the compaction filter turned into tombstones (:385-391) */
} orc_stats;

typedef struct orc_result orc_result;

/* leaf utilities */
uint32_t orc_crc32c_value(const void* data, size_t n);                    /* util/crc32c.h:25-33 */
uint32_t orc_crc32c_extend(uint32_t crc, const void* data, size_t n);
uint32_t orc_crc32c_mask(uint32_t crc);                                   /* util/crc32c.h:37-42 */
uint64_t orc_xxh3_64(const void* data, size_t n);                         /* util/xxhash.h XXH3_64bits */
uint32_t orc_block_checksum(uint32_t type, const void* data, size_t n, uint8_t last_byte); /* table/format.cc:468-509 */
uint32_t orc_checksum(uint32_t type, const void* data, size_t n);         /* table/format.cc:442-466 */
int orc_put_varint64(uint8_t* dst, uint64_t v);                           /* util/coding.h */
int orc_internal_key_less(const uint8_t* a, size_t an, const uint8_t* b, size_t bn); /* db/dbformat.h:1057-1097 */
size_t orc_shortest_separator(uint8_t* start, size_t start_len, const uint8_t* limit, size_t limit_len); /* index_builder.cc:77-94 */

/* "kv stream": repeated { u32 ikey_len, u32 value_len, ikey bytes, value bytes }, little endian. */

/* Decode every entry of a BlockBasedTable file (verifying block checksums) into a kv stream.
 * returns 0 or a negative error; *out is malloc'd. */
int orc_sst_to_kvstream(const uint8_t* file, size_t len, uint8_t** out, size_t* out_len, uint64_t* num_entries);

/* Build ONE BlockBasedTable file from a sorted kv stream (BlockBasedTableBuilder::Add/Finish). */
int orc_build_sst(const orc_params* p, const uint8_t* kv, size_t kv_len, uint8_t** out, size_t* out_len);

/* CompactionIterator over an already merged kv stream; emits the surviving kv stream. */
int orc_compaction_iterator(const orc_params* p, const uint8_t* kv, size_t kv_len, uint8_t** out, size_t* out_len,
                            orc_stats* stats);

/* Whole job: decode inputs (inputs[0] = newest L0 run first), k-way merge, drop rules, encode. */
int orc_file_cut_sim(const orc_params* p, int n, const uint8_t* const* ukeys, const uint32_t* ulens, uint64_t bytes_per_entry,
                     uint8_t* cut_before);
int orc_compact(const orc_params* p, int n_inputs, const uint8_t* const* inputs, const uint64_t* input_lens,
                orc_result** out);
int orc_result_num_files(const orc_result* r);
const uint8_t* orc_result_file(const orc_result* r, int i, uint64_t* len);
void orc_result_meta(const orc_result* r, int i, orc_file_meta* m);
void orc_result_stats(const orc_result* r, orc_stats* s);
void orc_result_free(orc_result* r);
void orc_free(void* p);
const char* orc_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
