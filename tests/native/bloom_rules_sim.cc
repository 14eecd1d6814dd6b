// tests/native/bloom_rules_sim.cc — test infrastructure: the product's Bloom-filter arithmetic (toplingdb_b200/csrc/bloom_rules.h, the
// code the device kernels run) compiled for the host.  tests/test_bloom_rules_host.py checks it against the reference's Hash64 known
// answers and against filter blocks written by the reference itself.
#include <stdint.h>
#include <string.h>

#include "bloom_rules.h"

using namespace b200c;

static void pack(const uint8_t* key, uint32_t len, uint64_t* hi, uint64_t* lo) {
  uint8_t b[16] = {0};
  memcpy(b, key, len);
  *hi = *lo = 0;
  for (int i = 0; i < 8; i++) *hi = (*hi << 8) | b[i], *lo = (*lo << 8) | b[8 + i];
}

extern "C" {

uint64_t bloom_sim_hash(const uint8_t* key, uint32_t len) {
  uint64_t hi, lo;
  pack(key, len, &hi, &lo);
  return xxph3_of_key(hi, lo, len);
}

// keys: n user keys of at most 16 bytes, 16-byte slots; in table order (duplicates adjacent).  Writes the filter block content (bits +
// 5 metadata bytes) the way the device kernels do -- dedupe against the predecessor's hash, lines / probes of bloom_rules.h -- and
// returns its length (out must hold bloom_bits_bytes(n) + 5); *entries = number of hashes added.
uint64_t bloom_sim_build(const uint8_t* keys, const uint32_t* lens, uint64_t n, uint32_t millibits, uint8_t* out, uint64_t* entries) {
  uint64_t cnt = 0, prev = 0;
  for (uint64_t i = 0; i < n; i++) {
    const uint64_t h = bloom_sim_hash(keys + 16 * i, lens[i]);
    if (i == 0 || h != prev) cnt++;
    prev = h;
  }
  const uint32_t len = bloom_bits_bytes(cnt, millibits);
  memset(out, 0, len + kBloomMetadataLen);
  const int probes = bloom_num_probes((int)millibits);
  for (uint64_t i = 0; i < n; i++) {
    const uint64_t h = bloom_sim_hash(keys + 16 * i, lens[i]);
    if (i != 0 && h == bloom_sim_hash(keys + 16 * (i - 1), lens[i - 1])) continue;
    uint8_t* line = out + bloom_line_offset(h, len);
    uint32_t p = bloom_first_probe(h);
    for (int k = 0; k < probes; k++, p = bloom_next_probe(p)) {
      const uint32_t bit = bloom_probe_bit(p);
      line[bit >> 3] |= (uint8_t)(1u << (bit & 7));
    }
  }
  out[len] = 0xff;
  out[len + 1] = 0;
  out[len + 2] = (uint8_t)probes;
  *entries = cnt;
  return (uint64_t)len + kBloomMetadataLen;
}

}  // extern "C"
