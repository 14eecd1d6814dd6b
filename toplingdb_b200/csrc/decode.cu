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
//   block_decode_kernel   one warp per data block: lanes own restart intervals (independent prefix-decode units),
//                         rebuild keys, emit (hi, lo, trailer, vref, meta)
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
constexpr int kDecWarps = 8;
constexpr int kDecSlice = 6144;  // bytes of shared memory per warp for one staged block (+trailer +align slack)

// stage [src, src+total) into the warp's slice, RE-ALIGNED so that byte 0 of the block sits at slice[0]: the
// global side uses aligned 16 B loads, the funnel shift happens in registers, and everything that reads the staged
// block afterwards (checksum stripes, restart array) can use naturally aligned shared-memory loads.
__device__ __forceinline__ const uint8_t* stage_block(const uint8_t* src, uint32_t total, uint8_t* slice) {
  const unsigned lane = threadIdx.x & 31;
  if (total + 32 > (uint32_t)kDecSlice) return src;  // too big: parse straight from global memory
  uintptr_t a0 = (uintptr_t)src & ~(uintptr_t)15;
  const uint32_t shift = (uint32_t)((uintptr_t)src - a0);
  const uint32_t nvec = (total + 15) >> 4;
  const uint4* g = (const uint4*)a0;
  uint4* s = (uint4*)slice;
  if (shift == 0) {
    for (uint32_t i = lane; i < nvec; i += 32) s[i] = __ldg(g + i);
  } else {
    for (uint32_t i = lane; i < nvec; i += 32) s[i] = shift16(__ldg(g + i), __ldg(g + i + 1), shift);
  }
  __syncwarp();
  return slice;
}

// number of entries in [p, end) (one restart interval); 0xffffffff on malformed data
__device__ __forceinline__ uint32_t count_interval(const uint8_t* p, const uint8_t* end) {
  uint32_t n = 0;
  while (p < end) {
    uint64_t shared, non_shared, vlen;
    int c;
    if (p + 3 <= end && (p[0] | p[1] | p[2]) < 128) {  // DecodeEntry fast path (block.cc:44-50)
      non_shared = p[1];
      vlen = p[2];
      p += 3;
    } else {
      if (!(c = get_varint(p, end, &shared))) return 0xffffffffu;
      p += c;
      if (!(c = get_varint(p, end, &non_shared))) return 0xffffffffu;
      p += c;
      if (!(c = get_varint(p, end, &vlen))) return 0xffffffffu;
      p += c;
    }
    if (non_shared + vlen > (uint64_t)(end - p)) return 0xffffffffu;
    p += non_shared + vlen;
    n++;
  }
  return n;
}

struct BlockView {
  const uint8_t* p;      // staged (or global) payload
  const uint8_t* gsrc;   // payload in the file image (global)
  uint32_t size, nr;
  const uint8_t* restarts;
  bool ok;
};

__device__ __forceinline__ BlockView open_block(const FileDesc& fd, uint64_t off, uint32_t size, uint8_t* slice,
                                                uint32_t verify, uint32_t* err) {
  BlockView v;
  v.ok = false;
  v.gsrc = fd.base + off;
  v.size = size;
  v.p = stage_block(v.gsrc, size + 5, slice);
  const unsigned lane = threadIdx.x & 31;
  uint8_t ctype = v.p[size];
  if (ctype != 0) {
    if (lane == 0) atomicOr(err, kErrCompressed);
    return v;
  }
  if (verify && fd.cksum != 0) {
    uint32_t want = ld_u32(v.p + size + 1);
    uint32_t got = block_checksum_warp(fd.cksum, v.p, size, ctype);
    if (want != got) {
      if (lane == 0) atomicOr(err, kErrChecksum);
      return v;
    }
  }
  uint32_t foot = ld_u32(v.p + size - 4);
  v.nr = foot & 0x7fffffffu;
  if ((foot >> 31) || v.nr == 0 || 4ull * v.nr + 4 > size) {  // hash-index blocks are not produced by the configs we accept
    if (lane == 0) atomicOr(err, kErrCorruptBlock);
    return v;
  }
  v.restarts = v.p + size - 4 - 4ull * v.nr;
  v.ok = true;
  return v;
}

__global__ void __launch_bounds__(kDecWarps * 32)
block_count_kernel(const FileDesc* __restrict__ files, int nfiles, const uint64_t* __restrict__ blk_off,
                   const uint32_t* __restrict__ blk_size, uint32_t nblk, uint32_t verify, uint32_t* __restrict__ blk_cnt,
                   uint32_t* __restrict__ err) {
  extern __shared__ __align__(16) uint8_t smem[];
  const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint8_t* slice = smem + (size_t)w * kDecSlice;
  for (uint32_t b = blockIdx.x * kDecWarps + w; b < nblk; b += gridDim.x * kDecWarps) {
    const int f = file_of_block(files, nfiles, b);
    const FileDesc fd = files[f];
    __syncwarp();
    BlockView v = open_block(fd, blk_off[b], blk_size[b], slice, verify, err);
    uint32_t cnt = 0;
    if (v.ok) {
      for (uint32_t j = lane; j < v.nr; j += 32) {
        uint32_t r0 = ld_u32(v.restarts + 4ull * j);
        uint32_t r1 = j + 1 < v.nr ? ld_u32(v.restarts + 4ull * (j + 1)) : (uint32_t)(v.restarts - v.p);
        uint32_t c = (r0 <= r1 && v.p + r1 <= v.restarts) ? count_interval(v.p + r0, v.p + r1) : 0xffffffffu;
        if (c == 0xffffffffu) {
          atomicOr(err, kErrCorruptBlock);
          c = 0;
        }
        cnt += c;
      }
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
    if (lane == 0) blk_cnt[b] = cnt;
  }
}

__global__ void __launch_bounds__(kDecWarps * 32)
block_decode_kernel(const FileDesc* __restrict__ files, int nfiles, const uint64_t* __restrict__ blk_off,
                    const uint32_t* __restrict__ blk_size, const uint64_t* __restrict__ blk_base, uint32_t nblk,
                    uint64_t n_total, KeyColsMut out, uint32_t* __restrict__ err) {
  extern __shared__ __align__(16) uint8_t smem[];
  const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint8_t* slice = smem + (size_t)w * kDecSlice;
  for (uint32_t b = blockIdx.x * kDecWarps + w; b < nblk; b += gridDim.x * kDecWarps) {
    const int f = file_of_block(files, nfiles, b);
    const FileDesc fd = files[f];
    __syncwarp();
    BlockView v = open_block(fd, blk_off[b], blk_size[b], slice, 0, err);
    if (!v.ok) continue;
    uint64_t base = blk_base[b];
    for (uint32_t j0 = 0; j0 < v.nr; j0 += 32) {
      uint32_t j = j0 + lane;
      uint32_t r0 = 0, r1 = 0, c = 0;
      if (j < v.nr) {
        r0 = ld_u32(v.restarts + 4ull * j);
        r1 = j + 1 < v.nr ? ld_u32(v.restarts + 4ull * (j + 1)) : (uint32_t)(v.restarts - v.p);
        c = (r0 <= r1 && v.p + r1 <= v.restarts) ? count_interval(v.p + r0, v.p + r1) : 0;
        if (c == 0xffffffffu) c = 0;
      }
      uint32_t inc = warp_incl_scan(c);
      uint64_t e = base + (inc - c);
      base += __shfl_sync(0xffffffffu, inc, 31);
      if (c == 0) continue;
      // prefix-decode this restart interval
      uint8_t kb[kMaxUserKey + 8];
      uint32_t klen = 0;
      const uint8_t* p = v.p + r0;
      const uint8_t* end = v.p + r1;
      for (uint32_t i = 0; i < c; i++, e++) {
        uint64_t shared, non_shared, vlen;
        int cc;
        if ((p[0] | p[1] | p[2]) < 128) {
          shared = p[0];
          non_shared = p[1];
          vlen = p[2];
          p += 3;
        } else {
          cc = get_varint(p, end, &shared);
          p += cc;
          cc = get_varint(p, end, &non_shared);
          p += cc;
          cc = get_varint(p, end, &vlen);
          p += cc;
        }
        if (shared > klen || shared + non_shared < 8) {
          atomicOr(err, kErrCorruptBlock);
          break;
        }
        if (shared + non_shared > (uint64_t)(kMaxUserKey + 8)) {
          atomicOr(err, kErrKeyTooLong);
          break;
        }
        if (vlen > kMetaVlenMask) {
          atomicOr(err, kErrValueTooLong);
          break;
        }
        for (uint32_t t = 0; t < (uint32_t)non_shared; t++) kb[shared + t] = p[t];
        klen = (uint32_t)(shared + non_shared);
        p += non_shared;
        const uint32_t ulen = klen - 8;
        uint64_t hi = 0, lo = 0, tr = 0;
        for (uint32_t t = 0; t < 8; t++) hi = (hi << 8) | (t < ulen ? kb[t] : 0);
        for (uint32_t t = 8; t < 16; t++) lo = (lo << 8) | (t < ulen ? kb[t] : 0);
        for (int t = 7; t >= 0; t--) tr = (tr << 8) | kb[ulen + t];
        if ((tr & 0xff) > 1) atomicOr(err, kErrBadType);
        if (e < n_total) {
          out.pfx[e] = make_ulonglong2(hi, lo);
          out.tr[e] = tr;
          out.vref[e] = (uint64_t)(uintptr_t)(v.gsrc + (p - v.p));
          out.meta[e] = make_meta(ulen, (uint32_t)vlen);
        } else {
          atomicOr(err, kErrCountMismatch);
        }
        p += vlen;
      }
    }
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
static unsigned dec_grid(uint32_t nblk, int sms) {
  unsigned want = (nblk + kDecWarps - 1) / kDecWarps;
  unsigned cap = (unsigned)sms * 4u;  // 4 CTAs of 8 warps per SM: 48 KB of staging each
  return want < cap ? (want ? want : 1) : cap;
}
void launch_block_count(const FileDesc* files_dev, int nfiles, const uint64_t* blk_off, const uint32_t* blk_size,
                        uint32_t nblk, uint32_t verify, uint32_t* blk_cnt, uint32_t* err, int sms, cudaStream_t st) {
  block_count_kernel<<<dec_grid(nblk, sms), kDecWarps * 32, kDecWarps * kDecSlice, st>>>(files_dev, nfiles, blk_off, blk_size,
                                                                                       nblk, verify, blk_cnt, err);
}
void launch_block_decode(const FileDesc* files_dev, int nfiles, const uint64_t* blk_off, const uint32_t* blk_size,
                         const uint64_t* blk_base, uint32_t nblk, uint64_t n_total, KeyColsMut out, uint32_t* err, int sms,
                         cudaStream_t st) {
  block_decode_kernel<<<dec_grid(nblk, sms), kDecWarps * 32, kDecWarps * kDecSlice, st>>>(files_dev, nfiles, blk_off, blk_size,
                                                                                        blk_base, nblk, n_total, out, err);
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
