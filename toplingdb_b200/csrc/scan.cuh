// toplingdb_b200/csrc/scan.cuh — device-wide exclusive prefix sum (u32 counts -> u64 offsets), three small
// kernels (tile sums, single-CTA scan of the sums, downsweep).  Side structure only: never on the byte path.
#pragma once
#include "common.cuh"

namespace b200c {

constexpr int kScanThreads = 512;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;  // 4096

__device__ __forceinline__ uint64_t block_excl_scan64(uint64_t v, uint64_t* total, uint64_t* warp_sums /* >= 32 */) {
  const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  uint64_t inc = warp_incl_scan64(v);
  if (lane == 31) warp_sums[w] = inc;
  __syncthreads();
  if (w == 0) {
    uint64_t s = lane < nw ? warp_sums[lane] : 0;
    uint64_t si = warp_incl_scan64(s);
    warp_sums[lane] = si - s;
    if (lane == 31) warp_sums[32] = si;
  }
  __syncthreads();
  uint64_t r = warp_sums[w] + inc - v;
  if (total) *total = warp_sums[32];
  __syncthreads();
  return r;
}

template <typename T>
__global__ void scan_tile_sums(const T* __restrict__ in, uint64_t n, uint64_t* __restrict__ tile_sums) {
  __shared__ uint64_t ws[33];
  uint64_t base = (uint64_t)blockIdx.x * kScanTile;
  uint64_t s = 0;
  for (int i = 0; i < kScanItems; i++) {
    uint64_t idx = base + (uint64_t)i * kScanThreads + threadIdx.x;
    if (idx < n) s += in[idx];
  }
  uint64_t tot;
  block_excl_scan64(s, &tot, ws);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}
// single CTA: exclusive scan of tile_sums in place, grand total to *total
static __global__ void scan_of_sums(uint64_t* __restrict__ tile_sums, uint64_t ntiles, uint64_t* __restrict__ total) {
  __shared__ uint64_t ws[33];
  __shared__ uint64_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (uint64_t b = 0; b < ntiles; b += blockDim.x) {
    uint64_t i = b + threadIdx.x;
    uint64_t v = i < ntiles ? tile_sums[i] : 0, tot;
    uint64_t ex = block_excl_scan64(v, &tot, ws);
    if (i < ntiles) tile_sums[i] = carry + ex;
    __syncthreads();
    if (threadIdx.x == 0) carry += tot;
    __syncthreads();
  }
  if (threadIdx.x == 0 && total) *total = carry;
}
template <typename T>
__global__ void scan_downsweep(const T* __restrict__ in, uint64_t n, const uint64_t* __restrict__ tile_sums,
                               uint64_t* __restrict__ out) {
  __shared__ uint64_t ws[33];
  // blocked arrangement: thread t owns items [t*kScanItems, (t+1)*kScanItems) of the tile
  uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanItems;
  uint64_t v[kScanItems], s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; i++) {
    v[i] = base + i < n ? (uint64_t)in[base + i] : 0;
    s += v[i];
  }
  uint64_t ex = block_excl_scan64(s, nullptr, ws) + tile_sums[blockIdx.x];
#pragma unroll
  for (int i = 0; i < kScanItems; i++) {
    if (base + i < n) out[base + i] = ex;
    ex += v[i];
  }
}

// host launcher; tmp must hold ceil(n / kScanTile) u64; out may not alias in
template <typename T>
inline void exclusive_scan(const T* in, uint64_t* out, uint64_t n, uint64_t* tmp, uint64_t* total_dev, cudaStream_t st,
                           uint64_t* launches) {
  if (n == 0) {
    if (total_dev) cudaMemsetAsync(total_dev, 0, 8, st);
    return;
  }
  uint64_t ntiles = (n + kScanTile - 1) / kScanTile;
  scan_tile_sums<T><<<(unsigned)ntiles, kScanThreads, 0, st>>>(in, n, tmp);
  scan_of_sums<<<1, 1024, 0, st>>>(tmp, ntiles, total_dev);
  scan_downsweep<T><<<(unsigned)ntiles, kScanThreads, 0, st>>>(in, n, tmp, out);
  if (launches) *launches += 3;
}

}  // namespace b200c
