// toplingdb_b200/csrc/decode.cu — BlockBasedTable input decode on the device.
//
// Replaces, for the compaction read path, BlockBasedTableIterator::{InitDataBlock,Next}
// (table/block_based/block_based_table_iterator.cc:224-411), BlockFetcher::ReadBlockContents + checksum check
// (table/block_fetcher.cc:32-40,211), DataBlockIter::ParseNextKey / DecodeEntry (table/block_based/block.cc:37-64,
// 617-665) and the IndexBlockIter walk that feeds them.
//
// Layout produced: every run decoded into one set of columns (KeyCols) indexed by a global entry number; run r
// owns [run_start[r], run_start[r+1]).  Values are NOT copied: vref[i] is the device address of the value bytes
// inside the resident file image, so each value byte crosses HBM once more only (file image -> output block).
//
// Kernels (all HBM-bound; algorithmic bytes = file bytes read + 36 B of columns written per entry):
//   index_decode_kernel   one thread per index restart point  -> data block handles
//   block_count_kernel    one warp per data block: stage block in shared memory with 16 B loads, verify the
//                         block checksum (warp-cooperative XXH3 / CRC32C), count entries per restart interval
//   block_decode_kernel   one THREAD per restart interval (the independent prefix-decode unit): aligned 8-byte loads
//                         straight from the image, key rebuilt in registers, emits (hi, lo, trailer, vref, meta)
#include "common.cuh"
#include "kernels.h"

namespace b200c {

__device__ __forceinline__ int file_of_block(const FileDesc* files, int nfiles, uint32_t gblk) {
  int f = 0;
  while (f + 1 < nfiles && gblk >= files[f + 1].gblk_first) f++;
  return f;
}

// ---------------------------------------------------------------------------------------------- index block
// IndexValue decode (table/format.cc:120-140): shared == 0 -> varint64 offset, varint64 size; else (fv >= 4)
// varsigned64 size delta and offset = prev_offset + prev_size + kBlockTrailerSize.
__device__ void index_decode_sequential(const FileDesc& fd, const uint8_t* blk, uint32_t nr, uint64_t* blk_off,
                                        uint32_t* blk_size, uint32_t* err) {
  const uint8_t* end = blk + fd.index_size - 4 - 4ull * nr;
  const uint8_t* p = blk;
  uint64_t poff = 0, psize = 0;
  uint32_t n = 0;
  while (p < end) {
    uint64_t shared, non_shared, vl = 0, off, size;
    int c;
    if (!(c = get_varint(p, end, &shared))) break;
    p += c;
    if (!(c = get_varint(p, end, &non_shared))) break;
    p += c;
    if (!fd.value_delta) {
      if (!(c = get_varint(p, end, &vl))) break;
      p += c;
    }
    p += non_shared;
    if (p > end) break;
    if (shared == 0 || !fd.value_delta) {
      if (!(c = get_varint(p, end, &off))) break;
      p += c;
      if (!(c = get_varint(p, end, &size))) break;
      p += c;
    } else {
      uint64_t t;
      if (!(c = get_varint(p, end, &t))) break;
      p += c;
      int64_t d = (int64_t)(t >> 1) ^ -(int64_t)(t & 1);
      size = psize + (uint64_t)d;
      off = poff + psize + 5;
    }
    if (n >= fd.nblocks) {
      n++;
      break;
    }
    blk_off[fd.gblk_first + n] = off;
    blk_size[fd.gblk_first + n] = (uint32_t)size;
    poff = off;
    psize = size;
    n++;
  }
  if (p != end || n != fd.nblocks) atomicOr(err, kErrCorruptBlock);
}

__global__ void index_decode_kernel(const FileDesc* __restrict__ files, int nfiles, uint64_t* __restrict__ blk_off,
                                    uint32_t* __restrict__ blk_size, uint32_t* __restrict__ err) {
  const int f = blockIdx.y;
  if (f >= nfiles) return;
  const FileDesc fd = files[f];
  if (fd.index_size < 8 || fd.index_off + fd.index_size + 5 > fd.len) {
    if (threadIdx.x == 0 && blockIdx.x == 0) atomicOr(err, kErrCorruptBlock);
    return;
  }
  const uint8_t* blk = fd.base + fd.index_off;
  const uint32_t nr = ld_u32(blk + fd.index_size - 4) & 0x7fffffffu;
  if (4ull * nr + 4 > fd.index_size) {
    if (threadIdx.x == 0 && blockIdx.x == 0) atomicOr(err, kErrCorruptBlock);
    return;
  }
  if (nr != fd.nblocks) {  // index_block_restart_interval != 1: rare, walk it with one thread
    if (threadIdx.x == 0 && blockIdx.x == 0) index_decode_sequential(fd, blk, nr, blk_off, blk_size, err);
    return;
  }
  const uint8_t* rs = blk + fd.index_size - 4 - 4ull * nr;
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < nr; j += gridDim.x * blockDim.x) {
    uint32_t ro = ld_u32(rs + 4ull * j);
    const uint8_t* p = blk + ro;
    uint64_t shared, non_shared, vl, off = 0, size = 0;
    int c, ok = ro < fd.index_size;
    ok = ok && (c = get_varint(p, rs, &shared)) && shared == 0;
    if (ok) p += c;
    ok = ok && (c = get_varint(p, rs, &non_shared));
    if (ok) p += c;
    if (ok && !fd.value_delta) {
      ok = (c = get_varint(p, rs, &vl)) != 0;
      if (ok) p += c;
    }
    if (ok) p += non_shared;
    ok = ok && p < rs && (c = get_varint(p, rs, &off));
    if (ok) p += c;
    ok = ok && (c = get_varint(p, rs, &size));
    if (!ok || off + size + 5 > fd.len || size < 4 || size > 0xffffffffull) {
      atomicOr(err, kErrCorruptBlock);
      off = 0;
      size = 4;
    }
    blk_off[fd.gblk_first + j] = off;
    blk_size[fd.gblk_first + j] = (uint32_t)size;
  }
}

// ---------------------------------------------------------------------------------------------- data blocks
struct Win {  // 40 bytes of the image starting at the 8-byte aligned address below p
  uint64_t w0, w1, w2, w3, w4;
};
__device__ __forceinline__ Win load_win(const uint8_t* p) {
  const uint64_t* a = reinterpret_cast<const uint64_t*>((uintptr_t)p & ~(uintptr_t)7);
  Win w;
  w.w0 = __ldg(a);
  w.w1 = __ldg(a + 1);
  w.w2 = __ldg(a + 2);
  w.w3 = __ldg(a + 3);
  w.w4 = __ldg(a + 4);
  return w;
}
// 8 bytes of the window starting at byte k (0 <= k <= 32)
__device__ __forceinline__ uint64_t win64(const Win& w, uint32_t k) {
  const uint32_t i = k >> 3, sh = (k & 7) * 8;
  uint64_t lo = i == 0 ? w.w0 : i == 1 ? w.w1 : i == 2 ? w.w2 : i == 3 ? w.w3 : w.w4;
  uint64_t hi = i == 0 ? w.w1 : i == 1 ? w.w2 : i == 2 ? w.w3 : i == 3 ? w.w4 : 0;
  return sh ? (lo >> sh) | (hi << (64 - sh)) : lo;
}
__device__ __forceinline__ uint64_t low_bytes_mask(uint32_t nbytes) {  // nbytes in 0..8
  return nbytes >= 8 ? ~0ull : ((1ull << (8 * nbytes)) - 1);
}


// entries in one restart interval [p, end), parsed with aligned word loads; 0xffffffff on malformed data
__device__ __forceinline__ uint32_t count_interval(const uint8_t* p, const uint8_t* end) {
  uint32_t n = 0;
  while (p < end) {
    const uint64_t h = ld_u64_funnel(p);
    uint32_t adv;
    if (((h | (h >> 8) | (h >> 16)) & 0x80) == 0) {  // DecodeEntry fast path (block.cc:44-50)
      adv = 3 + (uint32_t)((h >> 8) & 0xff) + (uint32_t)((h >> 16) & 0xff);
    } else {
      uint64_t shared, non_shared, vlen;
      const uint8_t* q = p;
      int c;
      if (!(c = get_varint(q, end, &shared))) return 0xffffffffu;
      q += c;
      if (!(c = get_varint(q, end, &non_shared))) return 0xffffffffu;
      q += c;
      if (!(c = get_varint(q, end, &vlen))) return 0xffffffffu;
      q += c;
      if (non_shared + vlen > (uint64_t)(end - q)) return 0xffffffffu;
      adv = (uint32_t)(q - p) + (uint32_t)(non_shared + vlen);
    }
    if (adv > (uint64_t)(end - p)) return 0xffffffffu;
    p += adv;
    n++;
  }
  return n;
}

// One warp per data block, software-pipelined: while the warp checksums / counts block i out of shared memory, the 16-byte
// vectors of block i+1 are already in flight into registers.  The staged copy keeps the file's 16-byte phase (no
// re-alignment work); the checksum and the parser read it with aligned 8-byte loads + funnel shifts.
constexpr int kCntWarps = 8;
constexpr int kCntSlice = 4608;            // bytes staged per warp (block + trailer + phase); larger blocks are read in place
constexpr int kCntVecs = kCntSlice / 16;   // 288
constexpr int kCntPerLane = kCntVecs / 32;  // 9 vectors per lane

__global__ void __launch_bounds__(kCntWarps * 32)
block_count_kernel(const FileDesc* __restrict__ files, int nfiles, const uint64_t* __restrict__ blk_off,
                   const uint32_t* __restrict__ blk_size, uint32_t nblk, uint32_t verify, uint32_t* __restrict__ blk_cnt,
                   uint32_t* __restrict__ blk_nr, uint32_t* __restrict__ blk_r, uint32_t* __restrict__ err) {
  extern __shared__ __align__(16) uint8_t smem[];
  const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint4* slice = reinterpret_cast<uint4*>(smem + (size_t)w * kCntSlice);
  const uint32_t stride = gridDim.x * kCntWarps;
  uint32_t b = blockIdx.x * kCntWarps + w;
  // prefetch state of the block about to be processed
  uint4 pre[kCntPerLane];
  const uint8_t* cur_src = nullptr;
  uint32_t cur_size = 0, cur_shift = 0, cur_cksum = 0;
  bool cur_staged = false;
  auto prefetch = [&](uint32_t bb) {
    const FileDesc& fd = files[file_of_block(files, nfiles, bb)];
    cur_src = fd.base + blk_off[bb];
    cur_size = blk_size[bb];
    cur_cksum = fd.cksum;
    const uintptr_t a0 = (uintptr_t)cur_src & ~(uintptr_t)15;
    cur_shift = (uint32_t)((uintptr_t)cur_src - a0);
    const uint32_t nvec = (cur_shift + cur_size + 5 + 15) >> 4;
    cur_staged = nvec <= (uint32_t)kCntVecs - 1;
    if (cur_staged) {
      const uint4* g = reinterpret_cast<const uint4*>(a0);
#pragma unroll
      for (int i = 0; i < kCntPerLane; i++) {
        const uint32_t v = lane + 32 * i;
        if (v < nvec) pre[i] = __ldg(g + v);
      }
    }
  };
  if (b < nblk) prefetch(b);
  for (; b < nblk; b += stride) {
    // commit the prefetched vectors of block b to shared memory
    const uint8_t* src = cur_src;
    const uint32_t size = cur_size, shift = cur_shift, cksum = cur_cksum;
    const bool staged = cur_staged;
    __syncwarp();
    if (staged) {
      const uint32_t nvec = (shift + size + 5 + 15) >> 4;
#pragma unroll
      for (int i = 0; i < kCntPerLane; i++) {
        const uint32_t v = lane + 32 * i;
        if (v < nvec) slice[v] = pre[i];
      }
    }
    __syncwarp();
    if (b + stride < nblk) prefetch(b + stride);  // next block's loads fly while this one is processed
    const uint8_t* p = staged ? reinterpret_cast<const uint8_t*>(slice) + shift : src;
    uint32_t cnt = 0, first = 0, nr = 0;
    bool ok = true;
    const uint8_t ctype = p[size];
    if (ctype != 0) {
      if (lane == 0) atomicOr(err, kErrCompressed);
      ok = false;
    }
    if (ok && verify && cksum != 0) {
      const uint32_t want = ld_u32(p + size + 1);
      const uint32_t got = block_checksum_warp(cksum, p, size, ctype);
      if (want != got) {
        if (lane == 0) atomicOr(err, kErrChecksum);
        ok = false;
      }
    }
    if (ok) {
      const uint32_t foot = ld_u32(p + size - 4);
      nr = foot & 0x7fffffffu;
      if ((foot >> 31) || nr == 0 || 4ull * nr + 4 > size) {  // data-block hash index: not produced by accepted configs
        if (lane == 0) atomicOr(err, kErrCorruptBlock);
        ok = false;
        nr = 0;
      }
    }
    if (ok) {
      const uint8_t* restarts = p + size - 4 - 4ull * nr;
      const uint32_t data_end = (uint32_t)(restarts - p);
      uint32_t irregular = 0;
      for (uint32_t j0 = 0; j0 < nr; j0 += 32) {
        const uint32_t j = j0 + lane;
        uint32_t c = 0;
        if (j < nr) {
          const uint32_t r0 = ld_u32(restarts + 4ull * j);
          const uint32_t r1 = j + 1 < nr ? ld_u32(restarts + 4ull * (j + 1)) : data_end;
          c = (r0 <= r1 && r1 <= data_end && (j != 0 || r0 == 0)) ? count_interval(p + r0, p + r1) : 0xffffffffu;
          if (c == 0xffffffffu) {
            atomicOr(err, kErrCorruptBlock);
            c = 0;
          }
        }
        if (j0 == 0) first = __shfl_sync(0xffffffffu, c, 0);
        // every restart interval but the last holds the same number of entries (what BlockBuilder writes): the decoder
        // derives an interval's output position from its index
        if (j + 1 < nr && c != first) irregular = 1;
        if (j + 1 == nr && c > first) irregular = 1;
        cnt += c;
      }
#pragma unroll
      for (int d = 16; d; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
      if (__any_sync(0xffffffffu, irregular)) {
        if (lane == 0) atomicOr(err, kErrIrregularRestarts);
      }
    }
    if (lane == 0) {
      blk_cnt[b] = ok ? cnt : 0;
      blk_nr[b] = ok ? nr : 0;
      blk_r[b] = first;
    }
  }
}

// ---- thread-per-restart-interval decode -----------------------------------------------------------------------
// A restart interval (16 entries by default) is the unit that can be prefix-decoded independently, so the decoder gives
// every interval its own thread: ~N/16 threads in flight instead of one warp per 4 KB block with 5 busy lanes.
// The thread streams through its ~1 KB of the file image with aligned 8-byte loads (L1 keeps the line for the next
// entry); the current internal key lives in three registers (K0..K2 = its 24 bytes as little-endian words) and is
// updated with mask / funnel-shift arithmetic instead of a byte buffer.

__global__ void __launch_bounds__(256)
block_decode_kernel(const FileDesc* __restrict__ files, int nfiles, const uint64_t* __restrict__ blk_off,
                    const uint32_t* __restrict__ blk_size, const uint64_t* __restrict__ blk_base, const uint32_t* __restrict__ blk_r,
                    const uint64_t* __restrict__ rbase, const uint64_t* __restrict__ total_intervals, uint32_t nblk, uint64_t n_total,
                    KeyColsMut out, uint32_t* __restrict__ err) {
  const unsigned lane = threadIdx.x & 31;
  const uint64_t total = *total_intervals;
  const uint64_t nwarps = (uint64_t)gridDim.x * (blockDim.x >> 5);
  for (uint64_t ii0 = ((uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32; ii0 < total; ii0 += nwarps * 32) {
    // block that holds interval ii0 (same search in every lane: uniform loads), then a short forward walk per lane
    uint32_t lo = 0, hi = nblk;
    while (hi - lo > 1) {
      uint32_t mid = (lo + hi) >> 1;
      if (rbase[mid] <= ii0) lo = mid;
      else hi = mid;
    }
    const uint64_t ii = ii0 + lane;
    if (ii >= total) continue;
    uint32_t b = lo;
    while (b + 1 < nblk && rbase[b + 1] <= ii) b++;
    const uint32_t j = (uint32_t)(ii - rbase[b]);
    const uint32_t nr = (uint32_t)((b + 1 < nblk ? rbase[b + 1] : total) - rbase[b]);
    const int f = file_of_block(files, nfiles, b);
    const uint8_t* blk = files[f].base + blk_off[b];
    const uint32_t size = blk_size[b];
    const uint8_t* restarts = blk + size - 4 - 4ull * nr;
    const uint32_t r0 = ld_u32(restarts + 4ull * j);
    const uint32_t r1 = j + 1 < nr ? ld_u32(restarts + 4ull * (j + 1)) : (uint32_t)(restarts - blk);
    uint64_t e = blk_base[b] + (uint64_t)j * blk_r[b];
    const uint8_t* p = blk + r0;
    const uint8_t* end = blk + r1;
    // the thread walks [p, end) front to back with dependent loads: pull its lines into L2 now so that every later miss
    // pays L2 latency instead of a DRAM round trip
    for (const uint8_t* q = p + 128; q < end; q += 128) prefetch_l2(q);
    uint64_t K0 = 0, K1 = 0, K2 = 0;
    uint32_t klen = 0;
    const uint64_t e_begin = e;
    const uint32_t want = j + 1 < nr ? blk_r[b] : 0xffffffffu;  // every interval but the last holds blk_r entries
    while (p < end) {
      const Win w = load_win(p);
      const uint32_t o = (uint32_t)((uintptr_t)p & 7);
      uint64_t h = win64(w, o);
      uint32_t shared, non_shared, vlen, hdr;
      if (((h | (h >> 8) | (h >> 16)) & 0x80) == 0) {  // DecodeEntry fast path: three one-byte lengths
        shared = (uint32_t)(h & 0xff);
        non_shared = (uint32_t)((h >> 8) & 0xff);
        vlen = (uint32_t)((h >> 16) & 0xff);
        hdr = 3;
      } else {  // varints from the 8 header bytes in the register (three varint32 of at most 8 bytes together)
        uint64_t vals[3];
        uint32_t k = 0;
        bool ok = true;
#pragma unroll
        for (int q = 0; q < 3; q++) {
          uint64_t v = 0;
          uint32_t sft = 0;
          for (;;) {
            if (k >= 8) {
              ok = false;
              break;
            }
            uint32_t c = (uint32_t)((h >> (8 * k)) & 0xff);
            k++;
            v |= (uint64_t)(c & 127) << sft;
            if (c < 128) break;
            sft += 7;
          }
          vals[q] = v;
        }
        if (!ok) {  // longer than 8 bytes: byte-wise from memory
          const uint8_t* q = p;
          int c1 = get_varint(q, end, &vals[0]);
          q += c1;
          int c2 = c1 ? get_varint(q, end, &vals[1]) : 0;
          q += c2;
          int c3 = c2 ? get_varint(q, end, &vals[2]) : 0;
          if (!c3) {
            atomicOr(err, kErrCorruptBlock);
            break;
          }
          k = (uint32_t)(c1 + c2 + c3);
        }
        if (vals[2] > kMetaVlenMask) {
          atomicOr(err, kErrValueTooLong);
          break;
        }
        if (vals[0] > 64 || vals[1] > 64 || k > 8) {  // keys longer than the device format / exotic header
          atomicOr(err, vals[0] + vals[1] > (uint64_t)(kMaxUserKey + 8) ? kErrKeyTooLong : kErrCorruptBlock);
          break;
        }
        shared = (uint32_t)vals[0];
        non_shared = (uint32_t)vals[1];
        vlen = (uint32_t)vals[2];
        hdr = k;
      }
      if (shared > klen || shared + non_shared < 8) {
        atomicOr(err, kErrCorruptBlock);
        break;
      }
      if (shared + non_shared > (uint32_t)(kMaxUserKey + 8)) {
        atomicOr(err, kErrKeyTooLong);
        break;
      }
      // key suffix: non_shared bytes that follow the header
      const uint32_t so = o + hdr;
      uint64_t S0 = win64(w, so), S1 = win64(w, so + 8), S2 = win64(w, so + 16);
      S0 &= low_bytes_mask(non_shared);
      S1 &= non_shared > 8 ? low_bytes_mask(non_shared - 8) : 0;
      S2 &= non_shared > 16 ? low_bytes_mask(non_shared - 16) : 0;
      // keep `shared` bytes of the previous key
      K0 &= low_bytes_mask(shared);
      K1 &= shared > 8 ? low_bytes_mask(shared - 8) : 0;
      K2 &= shared > 16 ? low_bytes_mask(shared - 16) : 0;
      // K |= S << (8 * shared)
      {
        const uint32_t ws = shared >> 3, bs = (shared & 7) * 8;
        const uint64_t c0 = bs ? S0 >> (64 - bs) : 0, c1 = bs ? S1 >> (64 - bs) : 0;
        const uint64_t T0 = S0 << bs, T1 = (S1 << bs) | c0, T2 = (S2 << bs) | c1;
        if (ws == 0) {
          K0 |= T0;
          K1 |= T1;
          K2 |= T2;
        } else if (ws == 1) {
          K1 |= T0;
          K2 |= T1;
        } else if (ws == 2) {
          K2 |= T0;
        }
      }
      klen = shared + non_shared;
      const uint32_t ulen = klen - 8;
      const uint64_t hi = bswap64(K0 & low_bytes_mask(ulen));
      const uint64_t lo = bswap64(ulen > 8 ? K1 & low_bytes_mask(ulen - 8) : 0);
      uint64_t tr;
      {
        const uint32_t ws = ulen >> 3, bs = (ulen & 7) * 8;
        const uint64_t a = ws == 0 ? K0 : ws == 1 ? K1 : K2, c = ws == 0 ? K1 : ws == 1 ? K2 : 0;
        tr = bs ? (a >> bs) | (c << (64 - bs)) : a;
      }
      if ((tr & 0xff) > 1) atomicOr(err, kErrBadType);
      const uint8_t* val = p + hdr + non_shared;
      if ((uint64_t)(end - val) < vlen) {
        atomicOr(err, kErrCorruptBlock);
        break;
      }
      if (val + vlen < end) prefetch_l1(val + vlen);  // next entry's header while this key is assembled
      if (e < n_total) {
        out.pfx[e] = make_ulonglong2(hi, lo);
        out.tr[e] = tr;
        out.vref[e] = (uint64_t)(uintptr_t)val;
        out.meta[e] = make_meta(ulen, vlen);
      } else {
        atomicOr(err, kErrCountMismatch);
      }
      e++;
      p = val + vlen;
    }
    if (want != 0xffffffffu ? (e - e_begin) != want : (e - e_begin) > blk_r[b]) atomicOr(err, kErrIrregularRestarts);
  }
}

// run_start[r] = number of entries before run r (blk_base at the run's first block); run_start[nfiles] = total
__global__ void run_starts_kernel(const FileDesc* __restrict__ files, int nfiles, const uint64_t* __restrict__ blk_base,
                                  const uint64_t* __restrict__ total, uint32_t nblk, uint64_t* __restrict__ run_start) {
  int r = threadIdx.x;
  if (r < nfiles) run_start[r] = files[r].gblk_first < nblk ? blk_base[files[r].gblk_first] : *total;
  if (r == nfiles) run_start[r] = *total;
}

// debug / test helper: gather value bytes through vref into a contiguous buffer (dst offsets from a scan of vlen)
__global__ void gather_values_kernel(KeyCols in, const uint64_t* __restrict__ dst_off, uint8_t* __restrict__ dst) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < in.n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint8_t* s = (const uint8_t*)(uintptr_t)in.vref[i];
    uint32_t n = meta_vlen(in.meta[i]);
    uint8_t* d = dst + dst_off[i];
    for (uint32_t t = 0; t < n; t++) d[t] = s[t];
  }
}
__global__ void meta_vlen_kernel(const uint32_t* __restrict__ meta, uint64_t n, uint32_t* __restrict__ vlen) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    vlen[i] = meta_vlen(meta[i]);
}

// ---------------------------------------------------------------------------------------------- host launchers
void launch_index_decode(const FileDesc* files_dev, int nfiles, uint32_t max_blocks_per_file, uint64_t* blk_off,
                         uint32_t* blk_size, uint32_t* err, cudaStream_t st) {
  dim3 grid((max_blocks_per_file + 255) / 256 ? (max_blocks_per_file + 255) / 256 : 1, nfiles);
  if (grid.x > 1024) grid.x = 1024;
  index_decode_kernel<<<grid, 256, 0, st>>>(files_dev, nfiles, blk_off, blk_size, err);
}
void launch_block_count(const FileDesc* files_dev, int nfiles, const uint64_t* blk_off, const uint32_t* blk_size,
                        uint32_t nblk, uint32_t verify, uint32_t* blk_cnt, uint32_t* blk_nr, uint32_t* blk_r, uint32_t* err, int sms,
                        cudaStream_t st) {
  unsigned want = (nblk + kCntWarps - 1) / kCntWarps, cap = (unsigned)sms * 5u;
  block_count_kernel<<<want < cap ? (want ? want : 1) : cap, kCntWarps * 32, kCntWarps * kCntSlice, st>>>(
      files_dev, nfiles, blk_off, blk_size, nblk, verify, blk_cnt, blk_nr, blk_r, err);
}
void launch_block_decode(const FileDesc* files_dev, int nfiles, const uint64_t* blk_off, const uint32_t* blk_size,
                         const uint64_t* blk_base, const uint32_t* blk_r, const uint64_t* rbase, const uint64_t* total_intervals,
                         uint32_t nblk, uint64_t n_total, KeyColsMut out, uint32_t* err, int sms, cudaStream_t st) {
  block_decode_kernel<<<(unsigned)sms * 8, 256, 0, st>>>(files_dev, nfiles, blk_off, blk_size, blk_base, blk_r, rbase,
                                                         total_intervals, nblk, n_total, out, err);
}
void launch_run_starts(const FileDesc* files_dev, int nfiles, const uint64_t* blk_base, const uint64_t* total, uint32_t nblk,
                       uint64_t* run_start, cudaStream_t st) {
  run_starts_kernel<<<1, 128, 0, st>>>(files_dev, nfiles, blk_base, total, nblk, run_start);
}
void launch_gather_values(KeyCols in, const uint64_t* dst_off, uint8_t* dst, cudaStream_t st) {
  if (in.n == 0) return;
  unsigned g = (unsigned)((in.n + 255) / 256);
  gather_values_kernel<<<g > 4096 ? 4096 : g, 256, 0, st>>>(in, dst_off, dst);
}
void launch_meta_vlen(const uint32_t* meta, uint64_t n, uint32_t* vlen, cudaStream_t st) {
  if (n == 0) return;
  unsigned g = (unsigned)((n + 255) / 256);
  meta_vlen_kernel<<<g > 4096 ? 4096 : g, 256, 0, st>>>(meta, n, vlen);
}

}  // namespace b200c
