// toplingdb_b200/csrc/bloom_rules.h — the arithmetic of the full Bloom filter block, host + device.
//
// BlockBasedTableOptions::filter_policy = NewBloomFilterPolicy(bits) with format_version >= 5 builds a FastLocalBloom filter over
// Hash64 (= XXPH3_64bits, the frozen preview of XXH3 in util/xxph3.h; NOT the final XXH3 of the block checksums) of every user key
// (table/block_based/filter_policy.cc:60-127,304-506; util/bloom_impl.h:156-214; full_filter_block.cc).  User keys on the device path
// are at most 16 bytes and live as two big-endian words, so only the 0 / 1-3 / 4-8 / 9-16 byte paths of XXPH3 exist here
// (util/xxph3.h:1083-1138), fed straight from the key columns.  tests/native/bloom_rules_sim.cc runs this code on the CPU against the
// reference's own Hash64 known answers (util/hash_test.cc) and against filter blocks the reference wrote.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define B200C_BLOOM_HD __host__ __device__ __forceinline__
#else
#define B200C_BLOOM_HD inline
#endif

namespace b200c {

B200C_BLOOM_HD uint64_t bl_bswap64(uint64_t v) {
  v = ((v & 0x00ff00ff00ff00ffull) << 8) | ((v >> 8) & 0x00ff00ff00ff00ffull);
  v = ((v & 0x0000ffff0000ffffull) << 16) | ((v >> 16) & 0x0000ffff0000ffffull);
  return (v << 32) | (v >> 32);
}
B200C_BLOOM_HD uint32_t bl_bswap32(uint32_t v) {
  v = ((v & 0x00ff00ffu) << 8) | ((v >> 8) & 0x00ff00ffu);
  return (v << 16) | (v >> 16);
}
B200C_BLOOM_HD uint64_t bl_mul_fold(uint64_t a, uint64_t b) {  // XXPH3_mul128_fold64
#if defined(__CUDA_ARCH__)
  return (a * b) ^ __umul64hi(a, b);
#else
  const unsigned __int128 m = (unsigned __int128)a * b;
  return (uint64_t)m ^ (uint64_t)(m >> 64);
#endif
}
B200C_BLOOM_HD uint64_t bl_avalanche(uint64_t h) {  // XXPH3_avalanche
  h ^= h >> 37;
  h *= 0x165667B19E3779F9ull;
  h ^= h >> 32;
  return h;
}
// little-endian words of the default secret's first 16 bytes (b8 fe 6c 39 23 a4 4b be | 7c 01 81 2c f7 21 ad 1c)
constexpr uint64_t kBlSecret0 = 0xbe4ba423396cfeb8ull, kBlSecret8 = 0x1cad21f72c81017cull;

// Hash64 of the `len`-byte user key whose bytes are the first `len` bytes of the big-endian pair (hi, lo) (zero padded behind)
B200C_BLOOM_HD uint64_t xxph3_of_key(uint64_t hi, uint64_t lo, uint32_t len) {
  if (len > 8) {  // XXPH3_len_9to16_64b: the first 8 bytes and the LAST 8 bytes, read little-endian
    const uint32_t s = len - 8;  // byte offset of the last 8 bytes, 1..8
    const uint64_t tail_be = s == 8 ? lo : ((hi << (8 * s)) | (lo >> (64 - 8 * s)));
    const uint64_t in_lo = bl_bswap64(hi) ^ kBlSecret0;
    const uint64_t in_hi = bl_bswap64(tail_be) ^ kBlSecret8;
    return bl_avalanche((uint64_t)len + (in_lo + in_hi) + bl_mul_fold(in_lo, in_hi));
  }
  if (len >= 4) {  // XXPH3_len_4to8_64b: first 4 and last 4 bytes
    const uint32_t first = bl_bswap32((uint32_t)(hi >> 32));
    const uint32_t last = bl_bswap32((uint32_t)(hi >> (8 * (8 - len))));
    const uint64_t keyed = ((uint64_t)first | ((uint64_t)last << 32)) ^ kBlSecret0;
    const uint64_t mix = (uint64_t)len + ((keyed ^ (keyed >> 51)) * 0x9E3779B1ull);
    return bl_avalanche((mix ^ (mix >> 47)) * 0xC2B2AE3D27D4EB4Full);
  }
  if (len) {  // XXPH3_len_1to3_64b
    const uint32_t c1 = (uint32_t)(hi >> 56), c2 = (uint32_t)(hi >> (56 - 8 * (len >> 1))) & 0xff,
                   c3 = (uint32_t)(hi >> (56 - 8 * (len - 1))) & 0xff;
    const uint32_t comb = c1 | (c2 << 8) | (c3 << 16) | (len << 24);
    return bl_avalanche(((uint64_t)comb ^ (uint64_t)(uint32_t)kBlSecret0) * 0x9E3779B185EBCA87ull);
  }
  return bl_mul_fold(kBlSecret0, 0xC2B2AE3D27D4EB4Full);  // RocksDB's change to the preview: the empty key hashes the seed
}

// FastLocalBloomImpl::ChooseNumProbes (util/bloom_impl.h:156-198)
B200C_BLOOM_HD int bloom_num_probes(int millibits_per_key) {
  const int lim[12] = {2080, 3580, 5100, 6640, 8300, 10070, 11720, 14001, 16050, 18300, 22001, 25501};
  for (int i = 0; i < 12; i++)
    if (millibits_per_key <= lim[i]) return i + 1;
  if (millibits_per_key > 50000) return 24;
  return (millibits_per_key - 1) / 2000 - 1;
}
// FastLocalBloomBitsBuilder::CalculateSpace (filter_policy.cc:409-424) without the 5 metadata bytes: bytes of filter bits for n hashes
B200C_BLOOM_HD uint32_t bloom_bits_bytes(uint64_t n, uint32_t millibits_per_key) {
  uint64_t raw = (n * millibits_per_key + 7999) / 8000;
  if (raw >= 0xffffffc0ull) raw = 0xffffffc0ull;
  return (uint32_t)((raw + 63) & ~63ull);
}
constexpr uint32_t kBloomMetadataLen = 5;
// FastLocalBloomImpl::AddHash (util/bloom_impl.h:200-214): byte offset of the 64-byte line of hash h, and the k-th probed bit in it
B200C_BLOOM_HD uint32_t bloom_line_offset(uint64_t h, uint32_t bits_bytes) {
  return (uint32_t)(((uint64_t)(uint32_t)h * (bits_bytes >> 6)) >> 32) << 6;  // FastRange32(Lower32of64(h), lines) * 64
}
B200C_BLOOM_HD uint32_t bloom_first_probe(uint64_t h) { return (uint32_t)(h >> 32); }                  // Upper32of64
B200C_BLOOM_HD uint32_t bloom_next_probe(uint32_t p) { return p * 0x9e3779b9u; }
B200C_BLOOM_HD uint32_t bloom_probe_bit(uint32_t p) { return p >> (32 - 9); }                           // bit within the 512-bit line

}  // namespace b200c
