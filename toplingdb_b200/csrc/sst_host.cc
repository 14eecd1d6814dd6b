// toplingdb_b200/csrc/sst_host.cc — see sst_host.h.  Host orchestration only; no per-entry work happens here.
#include "sst_host.h"

#include <string.h>

#include <algorithm>

namespace b200c {
namespace {

const uint64_t kMagic = 0x88e241b785f4cff7ull;  // kBlockBasedTableMagicNumber, block_based_table_builder.cc:202

bool get_varint(const uint8_t*& p, const uint8_t* end, uint64_t* v) {
  uint64_t r = 0;
  for (int s = 0; s <= 63 && p < end; s += 7) {
    uint8_t c = *p++;
    r |= (uint64_t)(c & 127) << s;
    if (c < 128) {
      *v = r;
      return true;
    }
  }
  return false;
}
void put_varint(std::vector<uint8_t>& b, uint64_t v) {
  while (v >= 128) {
    b.push_back((uint8_t)(v | 128));
    v >>= 7;
  }
  b.push_back((uint8_t)v);
}
void put_u32(std::vector<uint8_t>& b, uint32_t v) {
  for (int i = 0; i < 4; i++) b.push_back((uint8_t)(v >> (8 * i)));
}
uint32_t rd32(const uint8_t* p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }
uint64_t rd64(const uint8_t* p) { return (uint64_t)rd32(p) | (uint64_t)rd32(p + 4) << 32; }

// walk a (non value-delta) block: cb(key, value)
template <class F>
bool for_each_entry(const uint8_t* blk, uint64_t size, F cb) {
  if (size < 4) return false;
  uint32_t nr = rd32(blk + size - 4) & 0x7fffffffu;
  if ((uint64_t)nr * 4 + 4 > size) return false;
  const uint8_t *p = blk, *end = blk + size - 4 - 4ull * nr;
  std::string key;
  while (p < end) {
    uint64_t shared, non_shared, vlen;
    if (!get_varint(p, end, &shared) || !get_varint(p, end, &non_shared) || !get_varint(p, end, &vlen)) return false;
    if (shared > key.size() || (uint64_t)(end - p) < non_shared + vlen) return false;
    key.resize(shared);
    key.append((const char*)p, non_shared);
    p += non_shared;
    cb(key, p, vlen);
    p += vlen;
  }
  return true;
}

// ---- checksums for the two tiny tail blocks (XXH3 per util/xxhash.h, CRC32C per util/crc32c.cc)
const uint8_t kSecret[192] = {
    0xb8, 0xfe, 0x6c, 0x39, 0x23, 0xa4, 0x4b, 0xbe, 0x7c, 0x01, 0x81, 0x2c, 0xf7, 0x21, 0xad, 0x1c, 0xde, 0xd4, 0x6d, 0xe9,
    0x83, 0x90, 0x97, 0xdb, 0x72, 0x40, 0xa4, 0xa4, 0xb7, 0xb3, 0x67, 0x1f, 0xcb, 0x79, 0xe6, 0x4e, 0xcc, 0xc0, 0xe5, 0x78,
    0x82, 0x5a, 0xd0, 0x7d, 0xcc, 0xff, 0x72, 0x21, 0xb8, 0x08, 0x46, 0x74, 0xf7, 0x43, 0x24, 0x8e, 0xe0, 0x35, 0x90, 0xe6,
    0x81, 0x3a, 0x26, 0x4c, 0x3c, 0x28, 0x52, 0xbb, 0x91, 0xc3, 0x00, 0xcb, 0x88, 0xd0, 0x65, 0x8b, 0x1b, 0x53, 0x2e, 0xa3,
    0x71, 0x64, 0x48, 0x97, 0xa2, 0x0d, 0xf9, 0x4e, 0x38, 0x19, 0xef, 0x46, 0xa9, 0xde, 0xac, 0xd8, 0xa8, 0xfa, 0x76, 0x3f,
    0xe3, 0x9c, 0x34, 0x3f, 0xf9, 0xdc, 0xbb, 0xc7, 0xc7, 0x0b, 0x4f, 0x1d, 0x8a, 0x51, 0xe0, 0x4b, 0xcd, 0xb4, 0x59, 0x31,
    0xc8, 0x9f, 0x7e, 0xc9, 0xd9, 0x78, 0x73, 0x64, 0xea, 0xc5, 0xac, 0x83, 0x34, 0xd3, 0xeb, 0xc3, 0xc5, 0x81, 0xa0, 0xff,
    0xfa, 0x13, 0x63, 0xeb, 0x17, 0x0d, 0xdd, 0x51, 0xb7, 0xf0, 0xda, 0x49, 0xd3, 0x16, 0x55, 0x26, 0x29, 0xd4, 0x68, 0x9e,
    0x2b, 0x16, 0xbe, 0x58, 0x7d, 0x47, 0xa1, 0xfc, 0x8f, 0xf8, 0xb8, 0xd1, 0x7a, 0xd0, 0x31, 0xce, 0x45, 0xcb, 0x3a, 0x8f,
    0x95, 0x16, 0x04, 0x28, 0xaf, 0xd7, 0xfb, 0xca, 0xbb, 0x4b, 0x40, 0x7e};
const uint64_t P32_1 = 0x9E3779B1ull, P32_2 = 0x85EBCA77ull, P32_3 = 0xC2B2AE3Dull;
const uint64_t P64_1 = 0x9E3779B185EBCA87ull, P64_2 = 0xC2B2AE3D27D4EB4Full, P64_3 = 0x165667B19E3779F9ull,
               P64_4 = 0x85EBCA77C2B2AE63ull, P64_5 = 0x27D4EB2F165667C5ull;
uint64_t sec(int o) { return rd64(kSecret + o); }
uint64_t fold(uint64_t a, uint64_t b) {
  unsigned __int128 m = (unsigned __int128)a * b;
  return (uint64_t)m ^ (uint64_t)(m >> 64);
}
uint64_t aval(uint64_t h) {
  h ^= h >> 37;
  h *= 0x165667919E3779F9ull;
  return h ^ (h >> 32);
}
uint64_t aval64(uint64_t h) {
  h ^= h >> 33;
  h *= P64_2;
  h ^= h >> 29;
  h *= P64_3;
  return h ^ (h >> 32);
}
uint64_t mix16(const uint8_t* in, int so) { return fold(rd64(in) ^ sec(so), rd64(in + 8) ^ sec(so + 8)); }
void stripe(uint64_t acc[8], const uint8_t* in, int so) {
  for (int i = 0; i < 8; i++) {
    uint64_t dv = rd64(in + 8 * i), dk = dv ^ sec(so + 8 * i);
    acc[i ^ 1] += dv;
    acc[i] += (dk & 0xffffffffull) * (dk >> 32);
  }
}
uint32_t crc_tab[256];
bool crc_ready = false;
uint32_t crc_extend(uint32_t crc, const uint8_t* p, uint64_t n) {
  if (!crc_ready) {
    for (uint32_t i = 0; i < 256; i++) {
      uint32_t c = i;
      for (int k = 0; k < 8; k++) c = (c >> 1) ^ (0x82F63B78u & (0u - (c & 1)));
      crc_tab[i] = c;
    }
    crc_ready = true;
  }
  uint32_t c = crc ^ 0xffffffffu;
  for (uint64_t i = 0; i < n; i++) c = crc_tab[(c ^ p[i]) & 0xff] ^ (c >> 8);
  return c ^ 0xffffffffu;
}

// BlockBuilder with delta encoding and no value-delta (block_builder.cc:189-253), restart interval r
struct HostBlock {
  std::vector<uint8_t> b;
  std::vector<uint32_t> restarts{0};
  std::string last;
  int interval, counter = 0;
  explicit HostBlock(int r) : interval(r) {}
  void add(const std::string& k, const std::vector<uint8_t>& v) {
    size_t shared = 0;
    if (counter >= interval) {
      restarts.push_back((uint32_t)b.size());
      counter = 0;
    } else {
      size_t m = std::min(k.size(), last.size());
      while (shared < m && k[shared] == last[shared]) shared++;
    }
    put_varint(b, shared);
    put_varint(b, k.size() - shared);
    put_varint(b, v.size());
    b.insert(b.end(), k.begin() + shared, k.end());
    b.insert(b.end(), v.begin(), v.end());
    last = k;
    counter++;
  }
  void finish() {
    for (uint32_t r : restarts) put_u32(b, r);
    put_u32(b, (uint32_t)restarts.size());
  }
};
std::vector<uint8_t> vstr(const std::string& s) { return std::vector<uint8_t>(s.begin(), s.end()); }
std::vector<uint8_t> vu64(uint64_t v) {
  std::vector<uint8_t> b;
  put_varint(b, v);
  return b;
}

}  // namespace

uint64_t host_xxh3_64(const uint8_t* in, uint64_t len) {
  if (len <= 16) {
    if (len > 8) {
      uint64_t lo = rd64(in) ^ (sec(24) ^ sec(32)), hi = rd64(in + len - 8) ^ (sec(40) ^ sec(48));
      return aval(len + __builtin_bswap64(lo) + hi + fold(lo, hi));
    }
    if (len >= 4) {
      uint64_t i1 = rd32(in), i2 = rd32(in + len - 4), h = (i2 + (i1 << 32)) ^ (sec(8) ^ sec(16));
      h ^= ((h << 49) | (h >> 15)) ^ ((h << 24) | (h >> 40));
      h *= 0x9FB21C651E98DF25ull;
      h ^= (h >> 35) + len;
      h *= 0x9FB21C651E98DF25ull;
      return h ^ (h >> 28);
    }
    if (len) {
      uint32_t c = ((uint32_t)in[0] << 16) | ((uint32_t)in[len >> 1] << 24) | in[len - 1] | ((uint32_t)len << 8);
      return aval64((uint64_t)c ^ (uint64_t)(rd32(kSecret) ^ rd32(kSecret + 4)));
    }
    return aval64(sec(56) ^ sec(64));
  }
  if (len <= 128) {
    uint64_t acc = len * P64_1 + mix16(in, 0), e = mix16(in + len - 16, 16);
    if (len > 32) {
      acc += mix16(in + 16, 32);
      e += mix16(in + len - 32, 48);
      if (len > 64) {
        acc += mix16(in + 32, 64);
        e += mix16(in + len - 48, 80);
        if (len > 96) {
          acc += mix16(in + 48, 96);
          e += mix16(in + len - 64, 112);
        }
      }
    }
    return aval(acc + e);
  }
  if (len <= 240) {
    uint64_t acc = len * P64_1, e;
    for (int i = 0; i < 8; i++) acc += mix16(in + 16 * i, 16 * i);
    e = mix16(in + len - 16, 136 - 17);
    acc = aval(acc);
    for (unsigned i = 8; i < len / 16; i++) e += mix16(in + 16 * i, 16 * (i - 8) + 3);
    return aval(acc + e);
  }
  uint64_t acc[8] = {P32_3, P64_1, P64_2, P64_3, P64_4, P32_2, P64_5, P32_1};
  uint64_t nb = (len - 1) / 1024;
  for (uint64_t n = 0; n < nb; n++) {
    for (int s = 0; s < 16; s++) stripe(acc, in + n * 1024 + 64 * s, 8 * s);
    for (int i = 0; i < 8; i++) {
      uint64_t a = acc[i];
      a ^= a >> 47;
      a ^= sec(128 + 8 * i);
      acc[i] = a * P32_1;
    }
  }
  uint64_t ns = ((len - 1) - 1024 * nb) / 64;
  for (uint64_t s = 0; s < ns; s++) stripe(acc, in + nb * 1024 + 64 * s, (int)(8 * s));
  stripe(acc, in + len - 64, 192 - 64 - 7);
  uint64_t r = len * P64_1;
  for (int i = 0; i < 4; i++) r += fold(acc[2 * i] ^ sec(11 + 16 * i), acc[2 * i + 1] ^ sec(11 + 16 * i + 8));
  return aval(r);
}

uint32_t host_block_checksum(uint32_t type, const uint8_t* data, uint64_t n, uint8_t last_byte) {
  if (type == 1) {
    uint32_t c = crc_extend(crc_extend(0, data, n), &last_byte, 1);
    return ((c >> 15) | (c << 17)) + 0xa282ead8u;
  }
  if (type == 4) return (uint32_t)host_xxh3_64(data, n) ^ (uint32_t)last_byte * 0x6b9083d9u;
  return 0;
}

std::string parse_footer(const uint8_t* f, uint64_t file_len, InputTail* t) {
  if (file_len < 53) return "file shorter than a footer";
  if (rd64(f + 45) != kMagic) return "not a BlockBasedTable (bad magic number)";
  t->checksum_type = f[0];
  t->format_version = rd32(f + 41);
  const uint8_t *p = f + 1, *e = f + 41;
  if (!get_varint(p, e, &t->meta_off) || !get_varint(p, e, &t->meta_size) || !get_varint(p, e, &t->index_off) ||
      !get_varint(p, e, &t->index_size))
    return "bad footer handles";
  if (t->meta_off + t->meta_size + 5 > file_len || t->index_off + t->index_size + 5 > file_len) return "footer handle out of range";
  if (t->format_version < 2 || t->format_version > 5) return "unsupported format_version";
  if (t->checksum_type != 0 && t->checksum_type != 1 && t->checksum_type != 4) return "unsupported checksum type";
  return "";
}

std::string parse_metaindex(const uint8_t* blk, uint64_t size, std::map<std::string, std::pair<uint64_t, uint64_t>>* out) {
  bool ok = for_each_entry(blk, size, [&](const std::string& k, const uint8_t* v, uint64_t vl) {
    const uint8_t *p = v, *e = v + vl;
    uint64_t o = 0, s = 0;
    if (get_varint(p, e, &o) && get_varint(p, e, &s)) (*out)[k] = {o, s};
  });
  return ok ? "" : "corrupt metaindex block";
}

std::string parse_properties(const uint8_t* blk, uint64_t size, InputTail* t) {
  bool ok = for_each_entry(blk, size, [&](const std::string& k, const uint8_t* v, uint64_t vl) {
    auto num = [&](uint64_t* dst) {
      const uint8_t* p = v;
      uint64_t x;
      if (get_varint(p, v + vl, &x)) *dst = x;
    };
    if (k == "rocksdb.num.entries") num(&t->num_entries);
    else if (k == "rocksdb.num.data.blocks") num(&t->num_data_blocks);
    else if (k == "rocksdb.raw.key.size") num(&t->raw_key_size);
    else if (k == "rocksdb.raw.value.size") num(&t->raw_value_size);
    else if (k == "rocksdb.num.range-deletions") num(&t->num_range_deletions);
    else if (k == "rocksdb.merge.operands") num(&t->num_merge_operands);
    else if (k == "rocksdb.data.size") num(&t->data_size);
    else if (k == "rocksdb.index.key.is.user.key") num(&t->index_key_is_user_key);
    else if (k == "rocksdb.block.based.table.index.type" && vl >= 4)
      t->index_type = (uint32_t)v[0] | (uint32_t)v[1] << 8 | (uint32_t)v[2] << 16 | (uint32_t)v[3] << 24;
    else if (k == "rocksdb.compression") t->compression_name.assign((const char*)v, vl);
    else if (k == "rocksdb.comparator") t->comparator_name.assign((const char*)v, vl);
  });
  return ok ? "" : "corrupt properties block";
}

std::vector<uint8_t> build_output_tail(const OutputTailInput& in) {
  // PropertyBlockBuilder: std::map order, BlockBuilder(restart interval = INT32_MAX) (meta_blocks.cc:54-175)
  std::map<std::string, std::vector<uint8_t>> props;
  static const char kCompressionOpts[] =
      "window_bits=-14; level=32767; strategy=0; max_dict_bytes=0; zstd_max_train_bytes=0; enabled=0; "
      "max_dict_buffer_bytes=0; use_zstd_dict_trainer=1; ";
  props["rocksdb.block.based.table.index.type"] = {0, 0, 0, 0};
  props["rocksdb.block.based.table.prefix.filtering"] = vstr("0");
  props["rocksdb.block.based.table.whole.key.filtering"] = vstr("1");
  props["rocksdb.column.family.id"] = vu64(in.column_family_id);
  if (!in.column_family_name.empty()) props["rocksdb.column.family.name"] = vstr(in.column_family_name);
  props["rocksdb.comparator"] = vstr("leveldb.BytewiseComparator");
  props["rocksdb.compression"] = vstr("NoCompression");
  props["rocksdb.compression_options"] = vstr(kCompressionOpts);
  if (!in.db_id.empty()) props["rocksdb.creating.db.identity"] = vstr(in.db_id);
  if (!in.db_host_id.empty()) props["rocksdb.creating.host.identity"] = vstr(in.db_host_id);
  if (!in.db_session_id.empty()) props["rocksdb.creating.session.identity"] = vstr(in.db_session_id);
  props["rocksdb.creation.time"] = vu64(in.creation_time);
  props["rocksdb.data.size"] = vu64(in.data_size);
  props["rocksdb.deleted.keys"] = vu64(in.num_deletions);
  if (in.file_creation_time > 0) props["rocksdb.file.creation.time"] = vu64(in.file_creation_time);
  if (in.filter_size) props["rocksdb.filter.policy"] = vstr("bloomfilter");  // props.filter_policy_name (builder :1610)
  props["rocksdb.filter.size"] = vu64(in.filter_size);
  props["rocksdb.fixed.key.length"] = vu64(0);
  props["rocksdb.format.version"] = vu64(0);
  props["rocksdb.index.key.is.user.key"] = vu64(in.index_key_is_user_key ? 1 : 0);
  props["rocksdb.index.size"] = vu64(in.index_size + 5);
  props["rocksdb.index.value.is.delta.encoded"] = vu64(in.format_version >= 4 ? 1 : 0);
  props["rocksdb.merge.operands"] = vu64(0);
  props["rocksdb.merge.operator"] = vstr("nullptr");
  props["rocksdb.num.data.blocks"] = vu64(in.num_data_blocks);
  props["rocksdb.num.entries"] = vu64(in.num_entries);
  props["rocksdb.num.filter_entries"] = vu64(in.filter_entries);
  props["rocksdb.num.range-deletions"] = vu64(0);
  props["rocksdb.oldest.key.time"] = vu64(in.oldest_key_time);
  props["rocksdb.original.file.number"] = vu64(in.orig_file_number);
  props["rocksdb.prefix.extractor.name"] = vstr("nullptr");
  props["rocksdb.property.collectors"] = vstr("[]");
  props["rocksdb.raw.key.size"] = vu64(in.raw_key_size);
  props["rocksdb.raw.value.size"] = vu64(in.raw_value_size);
  props["rocksdb.tail.start.offset"] = vu64(in.data_size);
  HostBlock pb(0x7fffffff);
  for (auto& kv : props) pb.add(kv.first, kv.second);
  pb.finish();

  std::vector<uint8_t> out;
  auto append_block = [&](const std::vector<uint8_t>& blk) {
    out.insert(out.end(), blk.begin(), blk.end());
    out.push_back(0);
    put_u32(out, host_block_checksum(in.checksum_type, blk.data(), blk.size(), 0));
  };
  // file := data blocks | [filter block] | index block | properties | metaindex | footer (Finish :1948-1970)
  const uint64_t filter_off = in.data_size;
  const uint64_t index_off = in.data_size + (in.filter_size ? in.filter_size + 5 : 0);
  const uint64_t props_off = index_off + in.index_size + 5;
  append_block(pb.b);
  const uint64_t meta_off = props_off + pb.b.size() + 5;
  HostBlock mb(1);
  std::vector<uint8_t> h;
  if (in.filter_size) {  // "fullfilter." + FilterPolicy::CompatibilityName() (WriteFilterBlock :1532-1536); keys in sorted order
    put_varint(h, filter_off);
    put_varint(h, in.filter_size);
    mb.add("fullfilter.rocksdb.BuiltinBloomFilter", h);
    h.clear();
  }
  put_varint(h, props_off);
  put_varint(h, pb.b.size());
  mb.add("rocksdb.properties", h);
  mb.finish();
  append_block(mb.b);
  // footer (FooterBuilder::Build, table/format.cc:211-259), format_version >= 1
  std::vector<uint8_t> f;
  f.push_back((uint8_t)in.checksum_type);
  put_varint(f, meta_off);
  put_varint(f, mb.b.size());
  put_varint(f, index_off);
  put_varint(f, in.index_size);
  f.resize(41, 0);
  put_u32(f, in.format_version);
  for (int i = 0; i < 8; i++) f.push_back((uint8_t)(kMagic >> (8 * i)));
  out.insert(out.end(), f.begin(), f.end());
  return out;
}

}  // namespace b200c
