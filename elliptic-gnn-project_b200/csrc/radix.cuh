// Device-wide integer primitives shared by the graph build (graph_build.cu) and the epoch-tail metrics
// (metrics.cu): a three-kernel exclusive scan and one pass of a stable 8-bit LSD radix sort (in-tile ranking
// with __match_any_sync).  Grids are sized from a host-known capacity; the true length lives on the device.
#pragma once
#include "common.cuh"

namespace egnn {
namespace {

constexpr int kThreads = 256;
constexpr int kScanItems = 8;                       // per thread
constexpr int kScanTile = kThreads * kScanItems;    // 2048
constexpr int kSortItems = 8;
constexpr int kSortTile = kThreads * kSortItems;    // 2048

// ---- block-wide exclusive scan of one int per thread (256 threads) --------------------
__device__ __forceinline__ int block_excl_scan(int v, int* total, int* smem /*>=9 ints*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) smem[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int s = (lane < kThreads / 32) ? smem[lane] : 0;
    int si = s;
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, si, o);
      if (lane >= o) si += t;
    }
    if (lane < kThreads / 32) smem[lane] = si - s;  // exclusive warp offsets
    if (lane == kThreads / 32 - 1) smem[8] = si;    // block total
  }
  __syncthreads();
  int res = smem[warp] + inc - v;
  *total = smem[8];
  __syncthreads();  // smem reusable afterwards
  return res;
}

// ---- generic exclusive scan, 3 kernels -------------------------------------------------
__global__ void __launch_bounds__(kThreads) scan_tile_sums(const int* __restrict__ in, int64_t n,
                                                           int* __restrict__ tile_sums) {
  __shared__ int sm[9];
  int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  int s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i)
    if (base + i < n) s += in[base + i];
  int total;
  block_excl_scan(s, &total, sm);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// single block: in-place exclusive scan of tile_sums[0..m), total -> *total_out (may be null)
__global__ void __launch_bounds__(kThreads) scan_sums_inplace(int* __restrict__ sums, int64_t m,
                                                              int* __restrict__ total_out) {
  __shared__ int sm[9];
  int carry = 0;
  for (int64_t base = 0; base < m; base += kThreads) {
    int64_t i = base + threadIdx.x;
    int v = (i < m) ? sums[i] : 0;
    int total;
    int ex = block_excl_scan(v, &total, sm);
    if (i < m) sums[i] = carry + ex;
    carry += total;
  }
  if (threadIdx.x == 0 && total_out) *total_out = carry;
}

__global__ void __launch_bounds__(kThreads) scan_apply(const int* __restrict__ in, int64_t n,
                                                       const int* __restrict__ tile_offs,
                                                       int* __restrict__ out) {
  __shared__ int sm[9];
  int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  int v[kScanItems];
  int s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    v[i] = (base + i < n) ? in[base + i] : 0;
    s += v[i];
  }
  int total;
  int ex = block_excl_scan(s, &total, sm) + tile_offs[blockIdx.x];
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    if (base + i < n) out[base + i] = ex;
    ex += v[i];
  }
}

// exclusive scan of in[0..n) -> out[0..n) (in may alias out); total -> total_out (nullable)
int exclusive_scan(const int* in, int* out, int64_t n, int* tile_sums, int* total_out,
                   cudaStream_t st) {
  if (n <= 0) return 0;
  int64_t tiles = ceil_div(n, kScanTile);
  scan_tile_sums<<<(unsigned)tiles, kThreads, 0, st>>>(in, n, tile_sums);
  EGNN_LAUNCH_CHECK("scan_tile_sums");
  scan_sums_inplace<<<1, kThreads, 0, st>>>(tile_sums, tiles, total_out);
  EGNN_LAUNCH_CHECK("scan_sums_inplace");
  scan_apply<<<(unsigned)tiles, kThreads, 0, st>>>(in, n, tile_sums, out);
  EGNN_LAUNCH_CHECK("scan_apply");
  return 0;
}

// ---- LSD radix sort (stable) ------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) radix_hist(const int* __restrict__ keys,
                                                       const int* __restrict__ n_ptr, int shift,
                                                       int* __restrict__ table, int nblk) {
  __shared__ int h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const int n = *n_ptr;
  int64_t base = (int64_t)blockIdx.x * kSortTile;
#pragma unroll
  for (int it = 0; it < kSortItems; ++it) {
    int64_t i = base + it * kThreads + threadIdx.x;
    if (i < n) atomicAdd(&h[(keys[i] >> shift) & 255], 1);
  }
  __syncthreads();
  table[threadIdx.x * nblk + blockIdx.x] = h[threadIdx.x];
}

// vals_in == nullptr -> value = element index (first pass)
__global__ void __launch_bounds__(kThreads) radix_scatter(const int* __restrict__ keys_in,
                                                          const int* __restrict__ vals_in,
                                                          int* __restrict__ keys_out,
                                                          int* __restrict__ vals_out,
                                                          const int* __restrict__ n_ptr, int shift,
                                                          const int* __restrict__ table, int nblk) {
  __shared__ int warp_cnt[kThreads / 32][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (kThreads / 32) * 256; i += kThreads) (&warp_cnt[0][0])[i] = 0;
  __syncthreads();
  const int n = *n_ptr;
  const int64_t wbase = (int64_t)blockIdx.x * kSortTile + (int64_t)warp * (32 * kSortItems);
  int key[kSortItems], val[kSortItems], rank[kSortItems];
  const unsigned lt_mask = (1u << lane) - 1u;
#pragma unroll
  for (int it = 0; it < kSortItems; ++it) {
    int64_t i = wbase + it * 32 + lane;
    bool valid = i < n;
    key[it] = valid ? keys_in[i] : 0;
    val[it] = valid ? (vals_in ? vals_in[i] : (int)i) : 0;
    int digit = valid ? ((key[it] >> shift) & 255) : 256;
    unsigned peers = __match_any_sync(0xffffffffu, digit);
    int r = __popc(peers & lt_mask);
    int basecnt = 0;
    if (valid) basecnt = warp_cnt[warp][digit];
    __syncwarp();
    if (valid && r == 0) warp_cnt[warp][digit] = basecnt + __popc(peers);
    __syncwarp();
    rank[it] = basecnt + r;
  }
  __syncthreads();
  {  // thread d owns digit d: global base for (digit, this block) + prefix over warps
    int d = threadIdx.x;
    int run = table[d * nblk + blockIdx.x];
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) {
      int c = warp_cnt[w][d];
      warp_cnt[w][d] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int it = 0; it < kSortItems; ++it) {
    int64_t i = wbase + it * 32 + lane;
    if (i < n) {
      int pos = warp_cnt[warp][(key[it] >> shift) & 255] + rank[it];
      keys_out[pos] = key[it];
      vals_out[pos] = val[it];
    }
  }
}

}  // namespace
}  // namespace egnn
