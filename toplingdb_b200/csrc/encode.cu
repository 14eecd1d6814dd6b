// toplingdb_b200/csrc/encode.cu — BlockBasedTable output encode on the device, bit-exact with the reference builder.
//
// Replaces BlockBasedTableBuilder::{Add,Flush,WriteBlock,WriteMaybeCompressedBlock} (table/block_based/
// block_based_table_builder.cc:961-1133,1277-1378), BlockBuilder (block_builder.cc:97-253), FlushBlockBySizePolicy
// (flush_block_policy.cc:37-69), ShortenedIndexBuilder (index_builder.h:165-233, index_builder.cc:77-94,
// util/comparator.cc:42-91), the block checksums (table/format.cc:436-509) and the output-file cut rule of
// CompactionOutputs::ShouldStopBefore (db/compaction/compaction_outputs.cc:231-354, max-file-size rule :277).
//
// The reference cuts blocks with a sequential greedy rule.  Here it is evaluated in parallel:
//   encode_sizes      per entry: shared-prefix length with the previous internal key and encoded size
//   encode_tables     per tile of kEncTile entries: next(a) = "where does a block that starts at entry a end" for every
//                     a (prefix sums + bisection in shared memory), then the tile's transfer function
//                     entry-point -> (exit point, bytes, #blocks) for every entry point a chain can arrive at
//   encode_stitch     one CTA walks the tile functions in order (rows prefetched in batches), applying the
//                     max_output_file_size rule exactly, and records the state at which the chain enters every tile
//   encode_blocklist  per tile: follow the real chain, emit one BlockRec per data block
//   encode_emit       one warp per data block: encode entries (varints, key suffix, value bytes fetched through vref)
//                     into a shared-memory image, restart array, checksum, coalesced store into the file image
//   encode_index_*    per block separator keys, per-file index block, its checksum
// HBM-bound; algorithmic bytes of encode_emit = 36 B of columns + value bytes read + block bytes written per entry.
#include "common.cuh"
#include "kernels.h"
#include "scan.cuh"

namespace b200c {

constexpr int kTT = kEncTile;
constexpr int kW = kEncTile + kEncHalo;
constexpr int kEncThreads = 256;

// ------------------------------------------------------------------------------------------------ entry sizes
__device__ __forceinline__ uint32_t ikey_byte(uint64_t hi, uint64_t lo, uint32_t ulen, uint64_t tr, uint32_t j) {
  if (j < ulen) return (uint32_t)(((j < 8) ? (hi >> (56 - 8 * j)) : (lo >> (56 - 8 * (j - 8)))) & 0xff);
  return (uint32_t)((tr >> (8 * (j - ulen))) & 0xff);
}
// bytes shared by two internal keys (Slice::difference_offset on the raw key bytes, block_builder.cc:214)
__device__ __forceinline__ uint32_t shared_prefix(uint64_t ahi, uint64_t alo, uint32_t alen, uint64_t atr, uint64_t bhi, uint64_t blo,
                                                  uint32_t blen, uint64_t btr) {
  if (alen == blen) {
    uint32_t cb;
    uint64_t x = ahi ^ bhi;
    if (x) cb = (uint32_t)__clzll((long long)x) >> 3;
    else {
      uint64_t y = alo ^ blo;
      cb = y ? 8 + ((uint32_t)__clzll((long long)y) >> 3) : 16;
    }
    if (cb < alen) return cb;
    uint64_t z = atr ^ btr;
    uint32_t tb = z ? ((uint32_t)(__ffsll((long long)z) - 1) >> 3) : 8;
    return alen + tb;
  }
  uint32_t n = (alen < blen ? alen : blen) + 8, j = 0;
  while (j < n && ikey_byte(ahi, alo, alen, atr, j) == ikey_byte(bhi, blo, blen, btr, j)) j++;
  return j;
}
__device__ __forceinline__ uint32_t entry_size(uint32_t shared, uint32_t ks, uint32_t vs) {
  return (uint32_t)varint_len(shared) + (uint32_t)varint_len(ks - shared) + (uint32_t)varint_len(vs) + (ks - shared) + vs;
}

__global__ void encode_sizes_kernel(KeyCols m, const unsigned long long* __restrict__ n_dev, uint32_t* __restrict__ esz,
                                    uint8_t* __restrict__ eshared, uint32_t* __restrict__ min_s1) {
  const uint64_t n = *n_dev;
  uint32_t mn = 0xffffffffu;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    ulonglong2 c = m.pfx[i];
    uint64_t ctr = m.tr[i];
    uint32_t cm = m.meta[i], ulen = meta_ulen(cm), vs = meta_vlen(cm), ks = ulen + 8, sh = 0;
    if (i > 0) {
      ulonglong2 p = m.pfx[i - 1];
      sh = shared_prefix(c.x, c.y, ulen, ctr, p.x, p.y, meta_ulen(m.meta[i - 1]), m.tr[i - 1]);
    }
    uint32_t s1 = entry_size(sh, ks, vs);
    esz[i] = s1;
    eshared[i] = (uint8_t)sh;
    mn = s1 < mn ? s1 : mn;
  }
#pragma unroll
  for (int d = 16; d; d >>= 1) {
    uint32_t o = __shfl_xor_sync(0xffffffffu, mn, d);
    mn = o < mn ? o : mn;
  }
  if ((threadIdx.x & 31) == 0 && mn != 0xffffffffu) atomicMin(min_s1, mn);
}

// ------------------------------------------------------------------------------------------------ block-cut window
struct Window {
  uint64_t P[kW + 1];   // P[j] = sum of s1 of window entries < j
  uint32_t Q[kW];       // Q[j] = D[j] + Q[j - R]: restart surcharge prefix per residue class (D = s0 - s1)
  uint64_t ws[33];
  uint32_t wlen;        // entries loaded
  uint32_t at_end;      // window reaches the end of the stream
};

// cooperative: load the window of tile `tile` and build P / Q
__device__ void build_window(Window& w, const KeyCols& m, const uint32_t* esz, const uint8_t* eshared, uint64_t n, uint64_t tile,
                             uint32_t R) {
  const uint64_t wstart = tile * (uint64_t)kTT;
  const uint32_t wlen = (uint32_t)((n - wstart) < (uint64_t)kW ? (n - wstart) : (uint64_t)kW);
  constexpr int kPer = (kW + kEncThreads - 1) / kEncThreads;  // 24 consecutive entries per thread
  const uint32_t j0 = threadIdx.x * kPer;
  uint64_t loc[kPer], sum = 0;
#pragma unroll
  for (int i = 0; i < kPer; i++) {
    uint32_t j = j0 + i;
    uint32_t s1 = 0;
    if (j < wlen) {
      s1 = esz[wstart + j];
      uint32_t sh = eshared[wstart + j], ks = meta_ulen(m.meta[wstart + j]) + 8;
      // D = s0 - s1 with s0 = encoded size when shared == 0
      w.Q[j] = 1u + (uint32_t)varint_len(ks) + sh - (uint32_t)varint_len(sh) - (uint32_t)varint_len(ks - sh);
    }
    loc[i] = s1;
    sum += s1;
  }
  uint64_t ex = block_excl_scan64(sum, nullptr, w.ws);
#pragma unroll
  for (int i = 0; i < kPer; i++) {
    uint32_t j = j0 + i;
    if (j <= wlen && j <= (uint32_t)kW) w.P[j] = ex;
    ex += loc[i];
  }
  if (threadIdx.x == kEncThreads - 1 && wlen == (uint32_t)kW) w.P[kW] = ex;
  if (threadIdx.x == 0) {
    w.wlen = wlen;
    w.at_end = (wstart + wlen == n);
  }
  __syncthreads();
  // strided inclusive scan of Q with stride R (Hillis-Steele doubling)
  for (uint32_t off = R; off < wlen; off <<= 1) {
    uint32_t add[kPer];
#pragma unroll
    for (int i = 0; i < kPer; i++) {
      uint32_t j = j0 + i;
      add[i] = (j < wlen && j >= off) ? w.Q[j - off] : 0;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kPer; i++) {
      uint32_t j = j0 + i;
      if (j < wlen) w.Q[j] += add[i];
    }
    __syncthreads();
  }
}

struct CutParams {
  uint32_t BS, LIM, R;
};
// payload bytes of a block holding window entries [a, b)
__device__ __forceinline__ uint64_t blk_payload(const Window& w, uint32_t a, uint32_t b, uint32_t R) {
  uint32_t nrm1 = (b - 1 - a) / R;  // restarts - 1
  uint32_t last = a + R * nrm1;
  uint64_t q = (uint64_t)w.Q[last] - (a >= R ? (uint64_t)w.Q[a - R] : 0);
  return w.P[b] - w.P[a] + q + 4ull * (nrm1 + 1) + 4;  // == BlockBuilder::CurrentSizeEstimate() (block_builder.cc:97,251)
}
// first b > a at which FlushBlockBySizePolicy::Update (flush_block_policy.cc:37-69) fires for a block started at a.
// returns wlen at the end of the stream, 0xffffffff if the block does not end inside the window.
__device__ __forceinline__ uint32_t next_block(const Window& w, uint32_t a, const CutParams& cp) {
  const uint32_t wlen = w.wlen;
  const uint64_t thr = cp.LIM ? cp.LIM : cp.BS - 1;
  uint32_t lo = a + 1, hi = wlen + 1;
  while (lo < hi) {  // first b with CurrentSizeEstimate > thr
    uint32_t mid = (lo + hi) >> 1;
    if (blk_payload(w, a, mid, cp.R) > thr) hi = mid;
    else lo = mid + 1;
  }
  for (uint32_t b = lo; b < wlen; b++) {
    uint64_t ec = blk_payload(w, a, b, cp.R);
    if (ec >= cp.BS) return b;
    if (cp.LIM) {  // BlockAlmostFull: EstimateSizeAfterKV (block_builder.cc:97-126) = ec + |k|+|v|+4+varints (+4 at a restart)
      uint64_t d = (uint64_t)w.Q[b] - (b >= cp.R ? (uint64_t)w.Q[b - cp.R] : 0);
      uint64_t s0 = (w.P[b + 1] - w.P[b]) + d;
      uint64_t after = ec + s0 + 3 + (((b - a) % cp.R) == 0 ? 4 : 0);
      if (after > cp.BS) return b;
    }
  }
  return w.at_end ? wlen : 0xffffffffu;
}

// per tile: nxt / disk for every block start inside the tile, then the transfer function for hc entry points
struct TablesSmem {
  Window w;
  uint16_t nxt[kTT];   // window-relative end of the block that starts at j (0xffff = does not fit the window)
  uint32_t disk[kTT];  // on-disk bytes of that block (payload + 5-byte trailer)
};
__global__ void __launch_bounds__(kEncThreads)
encode_tables_kernel(KeyCols m, EncodeParams ep, EncodeWork wk, uint64_t n, uint32_t hc, uint16_t* __restrict__ g_nxt,
                     uint32_t* __restrict__ g_disk, uint32_t* __restrict__ err) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  TablesSmem& s = *reinterpret_cast<TablesSmem*>(smem_raw);
  const uint64_t tile = blockIdx.x, wstart = tile * (uint64_t)kTT;
  const CutParams cp{ep.block_size, ep.block_size_limit, ep.restart_interval};
  build_window(s.w, m, wk.esz, wk.eshared, n, tile, cp.R);
  const uint32_t tl = s.w.wlen < (uint32_t)kTT ? s.w.wlen : (uint32_t)kTT;
  for (uint32_t j = threadIdx.x; j < tl; j += kEncThreads) {
    uint32_t b = next_block(s.w, j, cp);
    uint16_t nx = 0xffff;
    uint32_t dk = 0;
    if (b != 0xffffffffu) {
      nx = (uint16_t)b;
      uint64_t pay = blk_payload(s.w, j, b, cp.R) + 5;
      dk = pay > 0xffffffffull ? 0xffffffffu : (uint32_t)pay;
    }
    s.nxt[j] = nx;
    s.disk[j] = dk;
    g_nxt[wstart + j] = nx;
    g_disk[wstart + j] = dk;
  }
  __syncthreads();
  for (uint32_t c = threadIdx.x; c < hc; c += kEncThreads) {
    uint32_t x = c, nb = 0;
    uint64_t bytes = 0;
    bool bad = false;
    while (x < tl) {
      uint32_t y = s.nxt[x];
      if (y == 0xffff || y <= x) {
        bad = true;
        break;
      }
      bytes += s.disk[x];
      nb++;
      x = y;
    }
    TileRow r;
    r.exit = bad ? 0xffffffffu : x;  // window-relative; >= tile length unless the stream ended
    r.nblk = nb;
    r.bytes = bytes;
    wk.rows[tile * hc + c] = r;
  }
  (void)err;
}

// ------------------------------------------------------------------------------------------------ stitch
struct WalkState {
  uint64_t a;         // absolute entry index where the open block starts
  uint64_t blk;       // blocks completed so far
  uint64_t foff;      // bytes flushed to the current file
  uint64_t f_first_entry, f_first_blk;
  uint32_t f;         // current file index
};
__device__ __forceinline__ void close_file(FileRec* files, WalkState& st, uint64_t end_entry, uint32_t* err) {
  if (files == nullptr) return;  // block-list pass: the stitch kernel already wrote the file records
  if (st.f >= kMaxOutFiles) {
    atomicOr(err, kErrInternal);
    return;
  }
  FileRec& fr = files[st.f];
  fr.first_entry = st.f_first_entry;
  fr.n_entries = end_entry - st.f_first_entry;
  fr.first_block = st.f_first_blk;
  fr.n_blocks = st.blk - st.f_first_blk;
  fr.data_size = st.foff;
  fr.index_size = 0;
  fr.raw_key_size = fr.raw_value_size = fr.num_deletions = 0;
  fr.smallest_seq = ~0ull;
  fr.largest_seq = 0;
  fr.index_has_seq = 0;
  fr.index_cksum = 0;
}
// walk the real chain through one tile with the window in shared memory (serial; only used for tiles in which an
// output file ends).  emit != nullptr: also write BlockRecs.  Returns false on a block that leaves the window.
__device__ bool walk_tile_serial(const Window& w, const CutParams& cp, const EncodeParams& ep, uint64_t wstart, uint64_t tile_end,
                                 uint64_t n, WalkState& st, FileRec* files, BlockRec* emit, uint64_t emit_cap, uint32_t* err) {
  while (st.a < tile_end && st.a < n) {
    uint32_t ra = (uint32_t)(st.a - wstart);
    uint32_t rb = next_block(w, ra, cp);
    if (rb == 0xffffffffu) return false;
    uint64_t y = wstart + rb;
    uint64_t dsk = blk_payload(w, ra, rb, cp.R) + 5;
    if (emit && st.blk < emit_cap) emit[st.blk] = BlockRec{st.a, st.foff, st.f, (uint32_t)(y - st.a)};
    st.foff += dsk;
    st.blk++;
    if (y >= n) {  // Finish(): last block of the stream
      close_file(files, st, n, err);
      st.a = n;
      st.f++;
      return true;
    }
    if (ep.output_level != 0 && st.foff >= ep.max_output_file_size) {
      // ShouldStopBefore fires in front of entry y+1: entry y (whose Add flushed the block) ends the file alone
      if (rb + 1 > w.wlen) return false;
      uint64_t d1 = blk_payload(w, rb, rb + 1, cp.R) + 5;
      if (emit && st.blk < emit_cap) emit[st.blk] = BlockRec{y, st.foff, st.f, 1u};
      st.foff += d1;
      st.blk++;
      close_file(files, st, y + 1, err);
      st.f++;
      st.foff = 0;
      st.f_first_entry = y + 1;
      st.f_first_blk = st.blk;
      st.a = y + 1;
    } else {
      st.a = y;
    }
  }
  return true;
}

struct StitchSmem {
  Window w;
  WalkState st;
  uint64_t next_tile;
  uint64_t filled;   // tiles whose TileState has been written
  uint32_t need_window;
  uint32_t done;
};
__global__ void __launch_bounds__(kEncThreads)
encode_stitch_kernel(KeyCols m, EncodeParams ep, EncodeWork wk, uint64_t n, uint64_t ntiles, uint32_t hc, uint32_t batch,
                     uint32_t* __restrict__ err) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  StitchSmem& s = *reinterpret_cast<StitchSmem*>(smem_raw);
  TileRow* rows = reinterpret_cast<TileRow*>(smem_raw + ((sizeof(StitchSmem) + 15) & ~(size_t)15));
  const CutParams cp{ep.block_size, ep.block_size_limit, ep.restart_interval};
  if (threadIdx.x == 0) {
    s.st = WalkState{0, 0, 0, 0, 0, 0};
    s.done = (n == 0);
    s.filled = 0;
    s.need_window = 0;
    s.next_tile = 0;
  }
  __syncthreads();
  uint64_t t = 0;
  while (t < ntiles && !s.done) {
    const uint64_t tb = (ntiles - t) < batch ? (ntiles - t) : batch;
    {  // prefetch the transfer functions of tiles [t, t + tb)
      const uint4* src = reinterpret_cast<const uint4*>(wk.rows + t * hc);
      uint4* dst = reinterpret_cast<uint4*>(rows);
      for (uint64_t i = threadIdx.x; i < tb * hc; i += kEncThreads) dst[i] = src[i];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      WalkState st = s.st;
      uint64_t tt = t;
      uint32_t need = 0;
      for (; tt < t + tb; tt++) {
        const uint64_t tstart = tt * (uint64_t)kTT;
        const uint64_t tend = (tstart + kTT) < n ? (tstart + kTT) : n;
        if (st.a >= n) break;
        TileState ts;
        ts.entry = st.a;
        ts.blk = st.blk;
        ts.file_off = st.foff;
        ts.file_idx = st.f;
        ts.pad = 0;
        wk.tstate[tt] = ts;
        s.filled = tt + 1;
        if (st.a >= tend) continue;  // no block starts in this tile
        uint64_t c = st.a - tstart;
        if (c >= hc) {
          atomicOr(err, kErrBlockTooLong);
          s.done = 1;
          break;
        }
        TileRow r = rows[(tt - t) * hc + c];
        if (r.exit == 0xffffffffu) {
          atomicOr(err, kErrBlockTooLong);
          s.done = 1;
          break;
        }
        const uint64_t exit_abs = tstart + r.exit;
        const bool cut = ep.output_level != 0 && st.foff + r.bytes >= ep.max_output_file_size;
        if (cut || exit_abs >= n) {  // a file ends inside this tile: walk it block by block
          need = 1;
          break;
        }
        st.a = exit_abs;
        st.blk += r.nblk;
        st.foff += r.bytes;
      }
      s.st = st;
      s.next_tile = tt;
      s.need_window = need;
    }
    __syncthreads();
    t = s.next_tile;
    if (s.done) break;
    if (s.need_window) {
      build_window(s.w, m, wk.esz, wk.eshared, n, t, cp.R);
      __syncthreads();
      if (threadIdx.x == 0) {
        const uint64_t tstart = t * (uint64_t)kTT;
        const uint64_t tend = (tstart + kTT) < n ? (tstart + kTT) : n;
        WalkState st = s.st;
        if (!walk_tile_serial(s.w, cp, ep, tstart, tend, n, st, wk.files, nullptr, 0, err)) {
          atomicOr(err, kErrBlockTooLong);
          s.done = 1;
        }
        s.st = st;
        if (st.a >= n) s.done = 1;
      }
      __syncthreads();
      t = t + 1;
    }
    if (threadIdx.x == 0 && s.st.a >= n) s.done = 1;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    // tiles never reached (stream ended earlier) get an entry point past the end
    wk.totals[0] = s.st.blk;
    wk.totals[1] = s.st.f;
    for (uint64_t tt = s.filled; tt < ntiles; tt++) {
      TileState ts;
      ts.entry = n;
      ts.blk = s.st.blk;
      ts.file_off = 0;
      ts.file_idx = s.st.f;
      ts.pad = 0;
      wk.tstate[tt] = ts;
    }
  }
}

// ------------------------------------------------------------------------------------------------ block list
struct BlistSmem {
  Window w;              // only built on the serial path
  uint16_t nxt[kTT];
  uint32_t disk[kTT];
  uint8_t is_start[kTT];
  uint64_t ws[33];
  uint32_t serial;
  WalkState st;
};
__global__ void __launch_bounds__(kEncThreads)
encode_blocklist_kernel(KeyCols m, EncodeParams ep, EncodeWork wk, uint64_t n, const uint16_t* __restrict__ g_nxt,
                        const uint32_t* __restrict__ g_disk, uint64_t nblk_cap, uint32_t* __restrict__ err) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  BlistSmem& s = *reinterpret_cast<BlistSmem*>(smem_raw);
  const uint64_t tile = blockIdx.x, tstart = tile * (uint64_t)kTT;
  const uint64_t tend = (tstart + kTT) < n ? (tstart + kTT) : n;
  const uint32_t tl = (uint32_t)(tend - tstart);
  const TileState ts = wk.tstate[tile];
  const CutParams cp{ep.block_size, ep.block_size_limit, ep.restart_interval};
  if (ts.entry >= tend) return;  // no block starts here
  for (uint32_t j = threadIdx.x; j < tl; j += kEncThreads) {
    s.nxt[j] = g_nxt[tstart + j];
    s.disk[j] = g_disk[tstart + j];
    s.is_start[j] = 0;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // follow the chain; decide whether a file ends inside this tile
    uint32_t x = (uint32_t)(ts.entry - tstart);
    uint64_t bytes = 0;
    bool bad = false, ends = false;
    while (x < tl) {
      uint32_t y = s.nxt[x];
      if (y == 0xffff || y <= x) {
        bad = true;
        break;
      }
      s.is_start[x] = 1;
      bytes += s.disk[x];
      if (tstart + y >= n) ends = true;
      x = y;
    }
    if (bad) atomicOr(err, kErrBlockTooLong);
    s.serial = bad ? 2 : ((ends || (ep.output_level != 0 && ts.file_off + bytes >= ep.max_output_file_size)) ? 1 : 0);
  }
  __syncthreads();
  if (s.serial == 2) return;
  if (s.serial == 1) {  // an output file ends inside this tile: replay the exact serial rule
    build_window(s.w, m, wk.esz, wk.eshared, n, tile, cp.R);
    __syncthreads();
    if (threadIdx.x == 0) {
      WalkState st;
      st.a = ts.entry;
      st.blk = ts.blk;
      st.foff = ts.file_off;
      st.f = ts.file_idx;
      st.f_first_entry = 0;
      st.f_first_blk = 0;
      if (!walk_tile_serial(s.w, cp, ep, tstart, tend, n, st, nullptr, wk.blocks, nblk_cap, err)) atomicOr(err, kErrBlockTooLong);
    }
    return;
  }
  // parallel path: ranks and byte offsets of the block starts
  constexpr int kPer = kTT / kEncThreads;  // 16
  const uint32_t j0 = threadIdx.x * kPer;
  uint32_t cnt = 0;
  uint64_t bytes = 0;
#pragma unroll
  for (int i = 0; i < kPer; i++) {
    uint32_t j = j0 + i;
    if (j < tl && s.is_start[j]) {
      cnt++;
      bytes += s.disk[j];
    }
  }
  uint64_t packed = ((uint64_t)cnt << 48) | bytes;  // tile bytes < 2^48, count < 2^16
  uint64_t ex = block_excl_scan64(packed, nullptr, s.ws);
  uint64_t r = ts.blk + (ex >> 48), off = ts.file_off + (ex & ((1ull << 48) - 1));
#pragma unroll
  for (int i = 0; i < kPer; i++) {
    uint32_t j = j0 + i;
    if (j < tl && s.is_start[j]) {
      if (r < nblk_cap) wk.blocks[r] = BlockRec{tstart + j, off, ts.file_idx, (uint32_t)(s.nxt[j] - j)};
      r++;
      off += s.disk[j];
    }
  }
}

// ------------------------------------------------------------------------------------------------ per-file statistics
__global__ void encode_filestats_kernel(KeyCols m, EncodeWork wk, uint32_t nfiles) {
  // each CTA owns a contiguous chunk of entries and adds its part to every file it overlaps
  const uint64_t n = m.n;
  const uint64_t chunk = (n + gridDim.x - 1) / gridDim.x;
  const uint64_t c0 = (uint64_t)blockIdx.x * chunk, c1 = (c0 + chunk) < n ? (c0 + chunk) : n;
  __shared__ unsigned long long red[5];
  for (uint32_t f = 0; f < nfiles; f++) {
    const uint64_t f0 = wk.files[f].first_entry, f1 = f0 + wk.files[f].n_entries;
    const uint64_t lo = c0 > f0 ? c0 : f0, hi = c1 < f1 ? c1 : f1;
    if (lo >= hi) continue;  // uniform per CTA
    if (threadIdx.x < 5) red[threadIdx.x] = threadIdx.x == 3 ? ~0ull : 0ull;
    __syncthreads();
    unsigned long long kb = 0, vb = 0, nd = 0, smin = ~0ull, smax = 0;
    for (uint64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
      uint64_t tr = m.tr[i];
      uint32_t mt = m.meta[i];
      kb += meta_ulen(mt) + 8;
      vb += meta_vlen(mt);
      nd += (tr & 0xff) == kTypeDeletion;
      uint64_t sq = tr >> 8;
      smin = sq < smin ? sq : smin;
      smax = sq > smax ? sq : smax;
      if (i == f0 || i == f1 - 1) {
        ulonglong2 p = m.pfx[i];
        KeyRec kr{p.x, p.y, tr, meta_ulen(mt), 0};
        if (i == f0) wk.files[f].smallest = kr;
        if (i == f1 - 1) wk.files[f].largest = kr;
      }
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) {
      kb += __shfl_xor_sync(0xffffffffu, kb, d);
      vb += __shfl_xor_sync(0xffffffffu, vb, d);
      nd += __shfl_xor_sync(0xffffffffu, nd, d);
      unsigned long long a = __shfl_xor_sync(0xffffffffu, smin, d), b = __shfl_xor_sync(0xffffffffu, smax, d);
      smin = a < smin ? a : smin;
      smax = b > smax ? b : smax;
    }
    if ((threadIdx.x & 31) == 0) {
      atomicAdd(&red[0], kb);
      atomicAdd(&red[1], vb);
      atomicAdd(&red[2], nd);
      atomicMin(&red[3], smin);
      atomicMax(&red[4], smax);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      atomicAdd((unsigned long long*)&wk.files[f].raw_key_size, red[0]);
      atomicAdd((unsigned long long*)&wk.files[f].raw_value_size, red[1]);
      atomicAdd((unsigned long long*)&wk.files[f].num_deletions, red[2]);
      atomicMin((unsigned long long*)&wk.files[f].smallest_seq, red[3]);
      atomicMax((unsigned long long*)&wk.files[f].largest_seq, red[4]);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------ emit data blocks
constexpr int kEmitWarps = 8;
// copy n bytes to generic dst from global src, one lane
__device__ __forceinline__ void lane_copy(uint8_t* dst, const uint8_t* src, uint32_t n) {
  for (uint32_t i = 0; i < n; i++) dst[i] = src[i];
}
__global__ void __launch_bounds__(kEmitWarps * 32)
encode_emit_kernel(KeyCols m, EncodeParams ep, EncodeWork wk, uint64_t nblocks, uint8_t* const* __restrict__ out_base,
                   uint32_t slice_bytes, uint32_t* __restrict__ err) {
  extern __shared__ __align__(16) uint8_t smem[];
  const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint8_t* slice = smem + (size_t)w * slice_bytes;
  const uint32_t R = ep.restart_interval;
  for (uint64_t b = (uint64_t)blockIdx.x * kEmitWarps + w; b < nblocks; b += (uint64_t)gridDim.x * kEmitWarps) {
    const BlockRec br = wk.blocks[b];
    uint8_t* gdst = out_base[br.file_idx] + br.file_off;
    const uint32_t ne = br.n_entries, nrest = (ne + R - 1) / R;
    // pass 1: payload size (entries + restart array + footer)
    uint64_t body = 0;
    for (uint32_t j = lane; j < ne; j += 32) {
      uint64_t e = br.first_entry + j;
      uint32_t s1 = wk.esz[e];
      if (j % R == 0) {
        uint32_t mt = m.meta[e], ks = meta_ulen(mt) + 8;
        s1 = entry_size(0, ks, meta_vlen(mt));
      }
      body += s1;
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) body += __shfl_xor_sync(0xffffffffu, body, d);
    const uint64_t payload = body + 4ull * nrest + 4;
    const bool staged = payload + 5 + 16 <= slice_bytes;
    const uint32_t shift = (uint32_t)((uintptr_t)gdst & 15);
    uint8_t* img = staged ? slice + shift : gdst;
    __syncwarp();
    // pass 2: encode entries, 32 at a time
    uint64_t off_base = 0;
    for (uint32_t j0 = 0; j0 < ne; j0 += 32) {
      const uint32_t j = j0 + lane;
      uint32_t sz = 0, sh = 0, ks = 0, vs = 0, ulen = 0;
      uint64_t e = br.first_entry + j, hi = 0, lo = 0, tr = 0, vref = 0;
      if (j < ne) {
        uint32_t mt = m.meta[e];
        ulen = meta_ulen(mt);
        ks = ulen + 8;
        vs = meta_vlen(mt);
        sh = (j % R == 0) ? 0 : wk.eshared[e];
        sz = entry_size(sh, ks, vs);
        ulonglong2 p = m.pfx[e];
        hi = p.x;
        lo = p.y;
        tr = m.tr[e];
        vref = m.vref[e];
      }
      // exclusive offsets inside the block (64-bit: a single value may be large)
      uint64_t inc = warp_incl_scan64(sz);
      uint64_t off = off_base + inc - sz;
      off_base += __shfl_sync(0xffffffffu, inc, 31);
      if (j < ne) {
        uint8_t* p = img + off;
        p += put_varint(p, sh);
        p += put_varint(p, ks - sh);
        p += put_varint(p, vs);
        for (uint32_t t = sh; t < ks; t++) *p++ = (uint8_t)ikey_byte(hi, lo, ulen, tr, t);
        if (vs < 128) lane_copy(p, (const uint8_t*)(uintptr_t)vref, vs);
        if (j % R == 0) {  // restart array slot (block_builder.cc:207-210,128-133)
          uint8_t* rp = img + body + 4ull * (j / R);
          uint32_t o32 = (uint32_t)off;
          rp[0] = (uint8_t)o32;
          rp[1] = (uint8_t)(o32 >> 8);
          rp[2] = (uint8_t)(o32 >> 16);
          rp[3] = (uint8_t)(o32 >> 24);
        }
      }
      // large values: the whole warp copies each of them
      unsigned big = __ballot_sync(0xffffffffu, j < ne && vs >= 128);
      while (big) {
        int src_lane = __ffs(big) - 1;
        big &= big - 1;
        uint64_t voff = __shfl_sync(0xffffffffu, off + (sz - vs), src_lane);
        uint64_t vr = __shfl_sync(0xffffffffu, vref, src_lane);
        uint32_t vl = __shfl_sync(0xffffffffu, vs, src_lane);
        const uint8_t* sp = (const uint8_t*)(uintptr_t)vr;
        uint8_t* dp = img + voff;
        for (uint32_t t = lane; t < vl; t += 32) dp[t] = sp[t];
      }
    }
    if (lane == 0) {
      uint8_t* fp = img + body + 4ull * nrest;
      fp[0] = (uint8_t)nrest;
      fp[1] = (uint8_t)(nrest >> 8);
      fp[2] = (uint8_t)(nrest >> 16);
      fp[3] = (uint8_t)(nrest >> 24);
    }
    if (!staged) __threadfence();
    __syncwarp();
    // trailer: compression type 0 + checksum (WriteMaybeCompressedBlock :1305-1329)
    uint32_t ck = block_checksum_warp(ep.checksum, img, payload, 0);
    if (lane == 0) {
      uint8_t* tp = img + payload;
      tp[0] = 0;
      tp[1] = (uint8_t)ck;
      tp[2] = (uint8_t)(ck >> 8);
      tp[3] = (uint8_t)(ck >> 16);
      tp[4] = (uint8_t)(ck >> 24);
    }
    __syncwarp();
    if (staged) {  // coalesced store of the image: head bytes, 16-byte body, tail bytes
      const uint32_t total = (uint32_t)payload + 5;
      uint32_t head = shift ? 16 - shift : 0;
      if (head > total) head = total;
      if (lane < head) gdst[lane] = img[lane];
      const uint32_t nvec = (total - head) >> 4;
      const uint4* sv = reinterpret_cast<const uint4*>(img + head);
      uint4* gv = reinterpret_cast<uint4*>(gdst + head);
      for (uint32_t i = lane; i < nvec; i += 32) gv[i] = sv[i];
      const uint32_t done = head + (nvec << 4);
      if (done + lane < total) gdst[done + lane] = img[done + lane];
      __syncwarp();
    }
  }
  (void)err;
}

// ------------------------------------------------------------------------------------------------ index block
// separator between the last key of block b and the first key of block b+1 of the same file
__global__ void encode_index_sep_kernel(KeyCols m, EncodeWork wk, uint64_t nblocks) {
  for (uint64_t b = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; b < nblocks; b += (uint64_t)gridDim.x * blockDim.x) {
    const BlockRec br = wk.blocks[b];
    const uint64_t le = br.first_entry + br.n_entries - 1;
    ulonglong2 lp = m.pfx[le];
    uint32_t lul = meta_ulen(m.meta[le]);
    KeyRec sep{lp.x, lp.y, m.tr[le], lul, 0};
    const bool has_next = b + 1 < nblocks && wk.blocks[b + 1].file_idx == br.file_idx;
    if (has_next) {
      const uint64_t ne = le + 1;
      ulonglong2 np = m.pfx[ne];
      uint32_t nul = meta_ulen(m.meta[ne]);
      if (np.x == lp.x && np.y == lp.y && nul == lul) {
        atomicOr(&wk.files[br.file_idx].index_has_seq, 1u);  // index_builder.h:175-180
      } else {
        // BytewiseComparator::FindShortestSeparator (util/comparator.cc:42-91) on the user keys
        uint8_t s[kMaxUserKey], l[kMaxUserKey];
        for (uint32_t t = 0; t < lul; t++) s[t] = (uint8_t)ikey_byte(lp.x, lp.y, lul, 0, t);
        for (uint32_t t = 0; t < nul; t++) l[t] = (uint8_t)ikey_byte(np.x, np.y, nul, 0, t);
        uint32_t minl = lul < nul ? lul : nul, d = 0, tn = lul;
        while (d < minl && s[d] == l[d]) d++;
        bool changed = false;
        if (d < minl && s[d] < l[d]) {
          if (d < nul - 1 || (uint32_t)s[d] + 1 < (uint32_t)l[d]) {
            s[d]++;
            tn = d + 1;
            changed = true;
          } else {
            d++;
            while (d < tn) {
              if (s[d] < 0xff) {
                s[d]++;
                tn = d + 1;
                changed = true;
                break;
              }
              d++;
            }
          }
        }
        if (changed) {  // shorter physically, larger logically: append (kMaxSequenceNumber, kValueTypeForSeek)
          uint64_t hi = 0, lo = 0;
          for (uint32_t t = 0; t < 8; t++) hi = (hi << 8) | (t < tn ? s[t] : 0);
          for (uint32_t t = 8; t < 16; t++) lo = (lo << 8) | (t < tn ? s[t] : 0);
          sep = KeyRec{hi, lo, (kMaxSeq << 8) | 0x16, tn, 1};
        }
      }
    }
    wk.idx_sep[b] = sep;
  }
}
// block payload size of data block b = distance to the next block of the file (or data_size) minus the trailer
__device__ __forceinline__ uint64_t block_payload_size(const EncodeWork& wk, uint64_t b, uint64_t nblocks) {
  const BlockRec br = wk.blocks[b];
  const bool has_next = b + 1 < nblocks && wk.blocks[b + 1].file_idx == br.file_idx;
  uint64_t end = has_next ? wk.blocks[b + 1].file_off : wk.files[br.file_idx].data_size;
  return end - br.file_off - 5;
}
__global__ void encode_index_size_kernel(EncodeWork wk, uint64_t nblocks, uint32_t format_version) {
  for (uint64_t b = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; b < nblocks; b += (uint64_t)gridDim.x * blockDim.x) {
    const BlockRec br = wk.blocks[b];
    const KeyRec sep = wk.idx_sep[b];
    const uint32_t klen = sep.ulen + (wk.files[br.file_idx].index_has_seq ? 8 : 0);
    const uint32_t hlen = (uint32_t)varint_len(br.file_off) + (uint32_t)varint_len(block_payload_size(wk, b, nblocks));
    wk.idx_esz[b] = 1 + (uint32_t)varint_len(klen) + (format_version >= 4 ? 0 : (uint32_t)varint_len(hlen)) + klen + hlen;
  }
}
__global__ void encode_index_write_kernel(EncodeWork wk, uint64_t nblocks, uint32_t format_version,
                                          uint8_t* const* __restrict__ out_base) {
  for (uint64_t b = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; b < nblocks; b += (uint64_t)gridDim.x * blockDim.x) {
    const BlockRec br = wk.blocks[b];
    const FileRec& fr = wk.files[br.file_idx];
    const KeyRec sep = wk.idx_sep[b];
    const uint32_t klen = sep.ulen + (fr.index_has_seq ? 8 : 0);
    const uint64_t eoff = wk.idx_eoff[b] - wk.idx_eoff[fr.first_block];
    const uint64_t entries_bytes =
        wk.idx_eoff[fr.first_block + fr.n_blocks - 1] + wk.idx_esz[fr.first_block + fr.n_blocks - 1] - wk.idx_eoff[fr.first_block];
    uint8_t* ib = out_base[br.file_idx] + fr.data_size;  // index block follows the last data block
    uint8_t* p = ib + eoff;
    uint8_t h[20];
    uint32_t hn = (uint32_t)put_varint(h, br.file_off);
    hn += (uint32_t)put_varint(h + hn, block_payload_size(wk, b, nblocks));
    p += put_varint(p, 0);
    p += put_varint(p, klen);
    if (format_version < 4) p += put_varint(p, hn);
    for (uint32_t t = 0; t < klen; t++) *p++ = (uint8_t)ikey_byte(sep.hi, sep.lo, sep.ulen, sep.tr, t);
    for (uint32_t t = 0; t < hn; t++) *p++ = h[t];
    // restart array: one restart per entry (index_block_restart_interval == 1)
    const uint64_t bi = b - fr.first_block;
    uint8_t* rp = ib + entries_bytes + 4 * bi;
    uint32_t o32 = (uint32_t)eoff;
    rp[0] = (uint8_t)o32;
    rp[1] = (uint8_t)(o32 >> 8);
    rp[2] = (uint8_t)(o32 >> 16);
    rp[3] = (uint8_t)(o32 >> 24);
    if (bi == 0) {
      uint8_t* fp = ib + entries_bytes + 4 * fr.n_blocks;
      uint32_t nr = (uint32_t)fr.n_blocks;
      fp[0] = (uint8_t)nr;
      fp[1] = (uint8_t)(nr >> 8);
      fp[2] = (uint8_t)(nr >> 16);
      fp[3] = (uint8_t)(nr >> 24);
      wk.files[br.file_idx].index_size = entries_bytes + 4 * fr.n_blocks + 4;
    }
  }
}
// one warp per file: checksum of the index block, trailer written behind it
__global__ void encode_index_cksum_kernel(EncodeWork wk, uint32_t nfiles, uint32_t cksum, uint8_t* const* __restrict__ out_base) {
  const uint32_t f = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (f >= nfiles) return;
  const FileRec fr = wk.files[f];
  if (fr.n_blocks == 0) return;
  uint8_t* ib = out_base[f] + fr.data_size;
  uint32_t ck = block_checksum_warp(cksum, ib, fr.index_size, 0);
  if ((threadIdx.x & 31) == 0) {
    uint8_t* tp = ib + fr.index_size;
    tp[0] = 0;
    tp[1] = (uint8_t)ck;
    tp[2] = (uint8_t)(ck >> 8);
    tp[3] = (uint8_t)(ck >> 16);
    tp[4] = (uint8_t)(ck >> 24);
    wk.files[f].index_cksum = ck;
  }
}

// checksum test entry point: one warp per buffer
__global__ void block_checksums_kernel(uint32_t type, const uint8_t* __restrict__ data, const uint64_t* __restrict__ offsets, uint32_t n,
                                       uint8_t last_byte, uint32_t* __restrict__ out) {
  const uint32_t i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= n) return;
  uint32_t ck = block_checksum_warp(type, data + offsets[i], offsets[i + 1] - offsets[i], last_byte);
  if ((threadIdx.x & 31) == 0) out[i] = ck;
}

// ------------------------------------------------------------------------------------------------ launchers
void launch_encode_sizes(KeyCols m, const unsigned long long* n_dev, EncodeWork w, uint64_t n_cap, cudaStream_t st) {
  if (n_cap == 0) return;
  unsigned g = (unsigned)((n_cap + 255) / 256);
  encode_sizes_kernel<<<g > 148 * 16 ? 148 * 16 : g, 256, 0, st>>>(m, n_dev, w.esz, w.eshared, w.min_s1);
}
void launch_encode_tables(KeyCols m, EncodeParams ep, EncodeWork w, uint64_t ntiles, uint32_t hc, uint32_t* err, cudaStream_t st) {
  if (ntiles == 0) return;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(encode_tables_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TablesSmem));
    attr = true;
  }
  encode_tables_kernel<<<(unsigned)ntiles, kEncThreads, sizeof(TablesSmem), st>>>(m, ep, w, m.n, hc, w.nxt, w.disk, err);
}
void launch_encode_stitch(KeyCols m, EncodeParams ep, EncodeWork w, uint64_t ntiles, uint32_t hc, uint32_t* err, cudaStream_t st) {
  static bool attr = false;
  const size_t base = (sizeof(StitchSmem) + 15) & ~(size_t)15;
  size_t avail = 200 * 1024 - base;
  uint32_t batch = (uint32_t)(avail / ((size_t)hc * sizeof(TileRow)));
  if (batch > 512) batch = 512;
  if (batch < 1) batch = 1;
  size_t smem = base + (size_t)batch * hc * sizeof(TileRow);
  if (!attr) {
    cudaFuncSetAttribute(encode_stitch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr = true;
  }
  encode_stitch_kernel<<<1, kEncThreads, smem, st>>>(m, ep, w, m.n, ntiles, hc, batch, err);
}
void launch_encode_blocklist(KeyCols m, EncodeParams ep, EncodeWork w, uint64_t ntiles, uint64_t nblk_cap, uint32_t* err,
                             cudaStream_t st) {
  if (ntiles == 0) return;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(encode_blocklist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BlistSmem));
    attr = true;
  }
  encode_blocklist_kernel<<<(unsigned)ntiles, kEncThreads, sizeof(BlistSmem), st>>>(m, ep, w, m.n, w.nxt, w.disk, nblk_cap, err);
}
void launch_encode_filestats(KeyCols m, EncodeWork w, uint32_t nfiles, int sms, cudaStream_t st) {
  if (m.n == 0 || nfiles == 0) return;
  encode_filestats_kernel<<<sms * 4, 256, 0, st>>>(m, w, nfiles);
}
uint32_t encode_emit_slice(uint32_t block_size) {
  uint32_t s = block_size + block_size / 2 + 1024;
  s = (s + 1023) & ~1023u;
  if (s < 8192) s = 8192;
  if (s > 24 * 1024) s = 24 * 1024;
  return s;
}
void launch_encode_emit(KeyCols m, EncodeParams ep, EncodeWork w, uint64_t nblocks, uint8_t* const* out_base, uint32_t* err, int sms,
                        cudaStream_t st) {
  if (nblocks == 0) return;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(encode_emit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr = true;
  }
  uint32_t slice = encode_emit_slice(ep.block_size);
  size_t smem = (size_t)slice * kEmitWarps;
  unsigned per_sm = (unsigned)((220 * 1024) / smem);
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 4) per_sm = 4;
  uint64_t want = (nblocks + kEmitWarps - 1) / kEmitWarps;
  uint64_t cap = (uint64_t)sms * per_sm * 4;
  encode_emit_kernel<<<(unsigned)(want < cap ? want : cap), kEmitWarps * 32, smem, st>>>(m, ep, w, nblocks, out_base, slice, err);
}
void launch_encode_index(KeyCols m, EncodeParams ep, EncodeWork w, uint64_t nblocks, uint32_t nfiles, uint8_t* const* out_base,
                         uint32_t* err, cudaStream_t st, uint64_t* launches) {
  (void)err;
  if (nblocks == 0) return;
  unsigned g = (unsigned)((nblocks + 255) / 256);
  if (g > 148 * 8) g = 148 * 8;
  encode_index_sep_kernel<<<g, 256, 0, st>>>(m, w, nblocks);
  encode_index_size_kernel<<<g, 256, 0, st>>>(w, nblocks, ep.format_version);
  exclusive_scan<uint32_t>(w.idx_esz, w.idx_eoff, nblocks, w.scan_tmp, nullptr, st, launches);
  encode_index_write_kernel<<<g, 256, 0, st>>>(w, nblocks, ep.format_version, out_base);
  encode_index_cksum_kernel<<<(nfiles + 3) / 4, 128, 0, st>>>(w, nfiles, ep.checksum, out_base);
  if (launches) *launches += 4;
}
void launch_block_checksums(uint32_t type, const uint8_t* data, const uint64_t* offsets, uint32_t n, uint8_t last_byte, uint32_t* out,
                            cudaStream_t st) {
  if (n == 0) return;
  block_checksums_kernel<<<(n + 3) / 4, 128, 0, st>>>(type, data, offsets, n, last_byte, out);
}

}  // namespace b200c
