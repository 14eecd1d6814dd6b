/* tests/native/mock_b200c.c -- test double of libb200c.so for the executor plugin (test infrastructure).  It claims one device, records
 * the b200c_params the plugin hands to b200c_job_create and the inputs it adds as one JSON object per job (file named by the environment
 * variable B200C_MOCK_DUMP, appended), and then refuses to run with B200C_ERR_NOT_SUPPORTED -- so a DB opened with
 * `executor=b200+fallback` goes on to compact locally.  tests/test_plugin_params.py compares what the plugin translated
 * (CompactionParams + BlockBasedTableOptions -> b200c_params) with what the reference itself reports about the same job.
 *
 * Second mode, B200C_MOCK_OUTPUTS=<dir>: b200c_job_run succeeds and the job's outputs are the files of <dir> as listed in <dir>/meta.txt
 * (written by tests/test_plugin_results.py from a run of the unmodified reference on the same data) -- a stand-in for a device that
 * produced exactly the reference's files, so that the plugin's result path (output directory, FileMinMeta, statistics) and RunRemote's
 * installation of the files can be exercised on the CPU.  No compaction logic of any kind lives here. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "b200c.h"

#define MOCK_MAX_OUT 256
struct b200c_job {
  FILE* f;
  int n_inputs, n_subs, sub_index; /* sub_index: -1 for a job of its own, else the position among its parent's sub-jobs */
  uint64_t first_file_number;
  int ran, n_out;
  b200c_file_meta meta[MOCK_MAX_OUT];
  unsigned char* data[MOCK_MAX_OUT];
  b200c_stats stats;
};
static const char* g_err = "";
static uint32_t unhex(const char* h, uint8_t* out);

static void hex(FILE* f, const void* p, uint32_t n) {
  const unsigned char* b = (const unsigned char*)p;
  fputc('"', f);
  for (uint32_t i = 0; i < n; i++) fprintf(f, "%02x", b[i]);
  fputc('"', f);
}
static void str(FILE* f, const char* s) {
  fputc('"', f);
  for (; s && *s; s++) {
    if (*s == '"' || *s == '\\') fputc('\\', f);
    fputc(*s, f);
  }
  fputc('"', f);
}

const char* b200c_last_error(void) { return g_err; }
int b200c_host_alloc(int device, uint64_t bytes, void** out) {
  (void)device;
  *out = malloc(bytes ? bytes : 1);
  return *out ? B200C_OK : B200C_ERR_OUT_OF_MEMORY;
}
void b200c_host_free(void* p) { free(p); }
int b200c_job_encode_kv(b200c_job* j, uint64_t n, const void* arena, const uint64_t* offs, const uint32_t* klens) {
  (void)j; (void)n; (void)arena; (void)offs; (void)klens;
  g_err = "test double: no encoder";
  return B200C_ERR_NOT_SUPPORTED;
}
uint32_t b200c_abi_version(void) { return B200C_ABI_VERSION; }
int b200c_device_count(void) { return 1; }
void b200c_params_init(b200c_params* p) { /* the defaults of the real library (api.cu) */
  memset(p, 0, sizeof *p);
  p->abi_version = B200C_ABI_VERSION;
  p->output_level = 1;
  p->max_output_file_size = 64ull << 20;
  p->block_size = 4096;
  p->block_size_deviation = 10;
  p->block_restart_interval = 16;
  p->index_block_restart_interval = 1;
  p->format_version = 5;
  p->checksum = B200C_CKSUM_XXH3;
  p->verify_input_checksums = 1;
  p->level_compaction_dynamic_file_size = 1;
  p->column_family_name = "default";
  p->output_mem = B200C_MEM_HOST;
}
int b200c_job_create(const b200c_params* p, b200c_job** out) {
  const char* path = getenv("B200C_MOCK_DUMP");
  b200c_job* j = (b200c_job*)calloc(1, sizeof *j);
  j->f = path ? fopen(path, "a") : NULL;
  if (j->f) {
    FILE* f = j->f;
    fprintf(f, "{\"abi_version\": %u, \"device\": %d, \"output_level\": %d, \"bottommost_level\": %d, \"max_output_file_size\": %llu, ",
            p->abi_version, p->device, p->output_level, p->bottommost_level, (unsigned long long)p->max_output_file_size);
    fprintf(f, "\"block_size\": %u, \"block_size_deviation\": %u, \"block_restart_interval\": %u, \"index_block_restart_interval\": %u, "
               "\"format_version\": %u, \"checksum\": %u, \"verify_input_checksums\": %u, ",
            p->block_size, p->block_size_deviation, p->block_restart_interval, p->index_block_restart_interval, p->format_version,
            p->checksum, p->verify_input_checksums);
    fprintf(f, "\"snapshots\": [");
    for (uint32_t i = 0; i < p->num_snapshots; i++) fprintf(f, "%s%llu", i ? ", " : "", (unsigned long long)p->snapshots[i]);
    fprintf(f, "], \"column_family_id\": %u, \"column_family_name\": ", p->column_family_id);
    str(f, p->column_family_name);
    fprintf(f, ", \"db_id\": ");
    str(f, p->db_id);
    fprintf(f, ", \"db_session_id\": ");
    str(f, p->db_session_id);
    fprintf(f, ", \"db_host_id\": ");
    str(f, p->db_host_id);
    fprintf(f, ", \"creation_time\": %llu, \"oldest_key_time\": %llu, \"file_creation_times\": [", (unsigned long long)p->creation_time,
            (unsigned long long)p->oldest_key_time);
    for (uint32_t i = 0; i < p->num_file_creation_times; i++)
      fprintf(f, "%s%llu", i ? ", " : "", (unsigned long long)p->file_creation_times[i]);
    fprintf(f, "], \"first_file_number\": %llu, \"output_mem\": %u, \"compaction_filter\": %u, \"ttl\": %d, \"ttl_now\": %lld, ",
            (unsigned long long)p->first_file_number, p->output_mem, p->compaction_filter, p->ttl, (long long)p->ttl_now);
    fprintf(f, "\"grandparents\": [");
    for (uint32_t i = 0; i < p->num_grandparents; i++) {
      fprintf(f, "%s{\"smallestkey\": ", i ? ", " : "");
      hex(f, p->grandparents[i].smallest_user_key, p->grandparents[i].smallest_len);
      fprintf(f, ", \"largestkey\": ");
      hex(f, p->grandparents[i].largest_user_key, p->grandparents[i].largest_len);
      fprintf(f, ", \"size\": %llu}", (unsigned long long)p->grandparents[i].file_size);
    }
    fprintf(f, "], \"level_compaction_dynamic_file_size\": %u, \"max_compaction_bytes\": %llu, \"target_output_file_size\": %llu, ",
            p->level_compaction_dynamic_file_size, (unsigned long long)p->max_compaction_bytes,
            (unsigned long long)p->target_output_file_size);
    fprintf(f, "\"has_range_start\": %u, \"has_range_end\": %u, \"range_start\": ", p->has_range_start, p->has_range_end);
    hex(f, p->range_start_user_key, p->has_range_start ? p->range_start_len : 0);
    fprintf(f, ", \"range_end\": ");
    hex(f, p->range_end_user_key, p->has_range_end ? p->range_end_len : 0);
    fprintf(f, ", \"paranoid_file_checks\": %u, \"bloom_millibits_per_key\": %u, \"inputs\": [", p->paranoid_file_checks, p->bloom_millibits_per_key);
  }
  j->first_file_number = p->first_file_number;
  j->sub_index = -1;
  *out = j;
  return B200C_OK;
}
/* B200C_MOCK_BOUNDARIES=<hex user key>,<hex user key>,...: the ranges the "planner" answers with (none: the job stays in one piece) */
int b200c_job_plan_ranges(b200c_job* j, uint32_t max_ranges, uint64_t min_range_bytes, uint8_t* keys, uint32_t* key_lens, uint32_t* n_boundaries) {
  (void)j; (void)min_range_bytes;
  *n_boundaries = 0;
  const char* b = getenv("B200C_MOCK_BOUNDARIES");
  if (!b || max_ranges < 2) return B200C_OK;
  char buf[1024];
  snprintf(buf, sizeof buf, "%s", b);
  for (char* tok = strtok(buf, ","); tok && *n_boundaries + 1 < max_ranges; tok = strtok(NULL, ",")) {
    uint8_t k[64];
    uint32_t n = unhex(tok, k);
    if (n > 16) n = 16;
    memcpy(keys + 16 * (size_t)*n_boundaries, k, n);
    key_lens[*n_boundaries] = n;
    (*n_boundaries)++;
  }
  return B200C_OK;
}
int b200c_job_upload_by_ranges(b200c_job* j, const uint8_t* keys, const uint32_t* key_lens, uint32_t n) {
  (void)j; (void)keys; (void)key_lens; (void)n;
  return B200C_OK;
}
int b200c_job_create_sub(b200c_job* parent, const b200c_params* p, b200c_job** out) {
  int rc = b200c_job_create(p, out);
  if (rc == B200C_OK) (*out)->sub_index = parent->n_subs++;
  return rc;
}
static uint32_t unhex(const char* h, uint8_t* out) {
  uint32_t n = 0;
  for (; h[0] && h[1]; h += 2) {
    unsigned v;
    sscanf(h, "%2x", &v);
    out[n++] = (uint8_t)v;
  }
  return n;
}
static int load_canned(b200c_job* j, const char* dir) {
  char path[4096], name[256], sk[256], lk[256];
  snprintf(path, sizeof path, "%s/meta.txt", dir);
  FILE* m = fopen(path, "r");
  if (!m) return 0;
  unsigned long long v[12];
  while (j->n_out < MOCK_MAX_OUT && fscanf(m, "%255s", name) == 1) {
    if (strcmp(name, "STATS") == 0) {
      if (fscanf(m, "%llu %llu %llu %llu %llu %llu %llu %llu %llu %llu", &v[0], &v[1], &v[2], &v[3], &v[4], &v[5], &v[6], &v[7], &v[8], &v[9]) != 10) break;
      j->stats.num_input_records = v[0];
      j->stats.num_output_records = v[1];
      j->stats.num_input_deletion_records = v[2];
      j->stats.num_records_replaced = v[3];
      j->stats.num_expired_deletion_records = v[4];
      j->stats.total_input_raw_key_bytes = v[5];
      j->stats.total_input_raw_value_bytes = v[6];
      j->stats.total_input_bytes = v[7];
      j->stats.total_output_bytes = v[8];
      j->stats.num_input_files = v[9];
      continue;
    }
    if (fscanf(m, "%255s %255s %llu %llu %llu %llu %llu %llu %llu %llu %llu", sk, lk, &v[0], &v[1], &v[2], &v[3], &v[4], &v[5], &v[6], &v[7], &v[8]) != 11) break;
    b200c_file_meta* fm = &j->meta[j->n_out];
    memset(fm, 0, sizeof *fm);
    snprintf(path, sizeof path, "%s/%s", dir, name);
    FILE* f = fopen(path, "rb");
    if (!f) break;
    fseek(f, 0, SEEK_END);
    long len = ftell(f);
    fseek(f, 0, SEEK_SET);
    j->data[j->n_out] = (unsigned char*)malloc((size_t)len);
    if (fread(j->data[j->n_out], 1, (size_t)len, f) != (size_t)len) {
      fclose(f);
      break;
    }
    fclose(f);
    fm->file_number = j->first_file_number + (uint64_t)j->n_out;
    fm->file_size = (uint64_t)len;
    fm->smallest_ikey_len = unhex(sk, fm->smallest_ikey);
    fm->largest_ikey_len = unhex(lk, fm->largest_ikey);
    fm->smallest_seqno = v[0];
    fm->largest_seqno = v[1];
    fm->num_entries = v[2];
    fm->num_deletions = v[3];
    fm->raw_key_size = v[4];
    fm->raw_value_size = v[5];
    fm->num_data_blocks = v[6];
    fm->data_size = v[7];
    fm->index_size = v[8];
    j->n_out++;
  }
  fclose(m);
  j->stats.num_output_files = (uint64_t)j->n_out;
  return 1;
}
int b200c_job_add_input(b200c_job* j, int level, uint64_t file_number, const void* data, uint64_t len, int mem_kind) {
  (void)data;
  if (j->f)
    fprintf(j->f, "%s{\"level\": %d, \"file_number\": %llu, \"len\": %llu, \"mem_kind\": %d}", j->n_inputs ? ", " : "", level,
            (unsigned long long)file_number, (unsigned long long)len, mem_kind);
  j->n_inputs++;
  return B200C_OK;
}
int b200c_job_run(b200c_job* j) {
  const char* dir = getenv("B200C_MOCK_OUTPUTS");
  const char* failcode = getenv("B200C_MOCK_FAIL");  /* third mode: fail the run with this enum b200c_status */
  if (failcode) {
    g_err = "mock library: injected failure";
    return atoi(failcode);
  }
  char sub[4096];
  if (dir && j->sub_index >= 0) { /* a sub-job "produces" the files of <dir>/sub<k> */
    snprintf(sub, sizeof sub, "%s/sub%d", dir, j->sub_index);
    dir = sub;
  }
  if (dir && load_canned(j, dir)) {
    j->ran = 1;
    return B200C_OK;
  }
  g_err = "mock library: records the job and leaves it to the reference";
  return B200C_ERR_NOT_SUPPORTED;
}
int b200c_job_output_count(const b200c_job* j) { return j->ran ? j->n_out : 0; }
int b200c_job_output_meta(const b200c_job* j, int i, b200c_file_meta* m) {
  if (!j->ran || i < 0 || i >= j->n_out) return B200C_ERR_STATE;
  *m = j->meta[i];
  return B200C_OK;
}
int b200c_job_output_data(b200c_job* j, int i, const void** d, uint64_t* l) {
  if (!j->ran || i < 0 || i >= j->n_out) return B200C_ERR_STATE;
  *d = j->data[i];
  *l = j->meta[i].file_size;
  return B200C_OK;
}
int b200c_job_output_read(b200c_job* j, int i, void* d, uint64_t c) {
  if (!j->ran || i < 0 || i >= j->n_out || c < j->meta[i].file_size) return B200C_ERR_STATE;
  memcpy(d, j->data[i], j->meta[i].file_size);
  return B200C_OK;
}
int b200c_job_get_stats(const b200c_job* j, b200c_stats* s) {
  *s = j->stats;
  return B200C_OK;
}
void b200c_job_destroy(b200c_job* j) {
  if (!j) return;
  if (j->f) {
    fprintf(j->f, "]}\n");
    fclose(j->f);
  }
  for (int i = 0; i < j->n_out; i++) free(j->data[i]);
  free(j);
}
int b200c_job_run_until(b200c_job* j, int stage) { (void)stage; return b200c_job_run(j); }
int b200c_job_debug_read(b200c_job* j, int w, int r, void* d, uint64_t c, uint64_t* l) { (void)j; (void)w; (void)r; (void)d; (void)c; (void)l; return B200C_ERR_STATE; }
int b200c_job_kernel_time_count(const b200c_job* j) { (void)j; return 0; }
int b200c_job_kernel_time(const b200c_job* j, int i, const char** n, double* us) { (void)j; (void)i; (void)n; (void)us; return B200C_ERR_STATE; }
int b200c_job_encode_columns(b200c_job* j, uint64_t n, const void* a, const void* b, const void* c, const void* d) { (void)j; (void)n; (void)a; (void)b; (void)c; (void)d; return B200C_ERR_NOT_SUPPORTED; }
int b200c_block_checksums(int dev, uint32_t t, const void* h, const uint64_t* o, uint32_t n, uint8_t lb, uint32_t* out) { (void)dev; (void)t; (void)h; (void)o; (void)n; (void)lb; (void)out; return B200C_ERR_NO_DEVICE; }
