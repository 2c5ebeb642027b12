// SURVEY.md 8(f) rank 3 -- the data formats in front of the path: the Elliptic CSV tables / graph.pt contents become
// the device-resident edge list the graph build consumes, without the reference's per-edge Python dictionary walk.
//
// Reference (src/data/dataset_elliptic.py):
//   :190-196  tx_to_idx = {int(tx): i for i, tx in enumerate(tx_ids)}     node order = CSV row order; a txId that
//                                                                         occurs twice maps to its LAST row
//   :221-232  keep = src.isin(tx_to_idx) & dst.isin(tx_to_idx);  src_idx / dst_idx = mapped endpoints, CSV order
//   :235-241  same_t = timestep[src_idx] == timestep[dst_idx];  only intra-timestep edges survive
//   :245      edge_index = int64 [2, E_kept]
//   :268-290  make_temporal_masks: labelled-node masks by timestep window
// Here: an open-addressing hash table over the N txIds (64-bit keys, linear probing, `atomicMax` on the row index so
// that the last duplicate wins like the dict), a probe per raw edge endpoint, the timestep test, and a stable
// compaction (flags -> exclusive scan -> scatter) so the surviving edges keep their CSV order.  Integer work only;
// bit-exact against the reference's own function (tests/golden/make_ingest_golden.py).
#include "radix.cuh"

namespace egnn {
namespace {

constexpr unsigned long long kEmpty = ~0ull;

__device__ __forceinline__ uint64_t mix64(uint64_t z) {   // splitmix64 finaliser
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

struct JoinWs {
  unsigned long long* keys;  // [cap]
  int* vals;                 // [cap + 1]; vals[cap] serves the one key that equals the empty marker
  int* keep;                 // [E]
  int* pos;                  // [E]
  int* s32;                  // [E]
  int* d32;                  // [E]
  int* tile_sums;            // [ceil(E / kScanTile) + 1]
  int64_t cap;
  size_t bytes;
};

JoinWs carve_join(char* base, int64_t n_nodes, int64_t E) {
  JoinWs w;
  int64_t cap = 1024;
  while (cap < 2 * n_nodes) cap <<= 1;
  w.cap = cap;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* p = base ? base + off : nullptr;
    off += (bytes + 255) & ~size_t(255);
    return p;
  };
  const int64_t Ec = E > 0 ? E : 1;
  w.keys = reinterpret_cast<unsigned long long*>(take(sizeof(unsigned long long) * (size_t)cap));
  w.vals = reinterpret_cast<int*>(take(sizeof(int) * (size_t)(cap + 1)));
  w.keep = reinterpret_cast<int*>(take(sizeof(int) * (size_t)Ec));
  w.pos = reinterpret_cast<int*>(take(sizeof(int) * (size_t)Ec));
  w.s32 = reinterpret_cast<int*>(take(sizeof(int) * (size_t)Ec));
  w.d32 = reinterpret_cast<int*>(take(sizeof(int) * (size_t)Ec));
  w.tile_sums = reinterpret_cast<int*>(take(sizeof(int) * (size_t)(ceil_div(Ec, kScanTile) + 1)));
  w.bytes = off;
  return w;
}

__global__ void __launch_bounds__(kThreads) join_insert(const int64_t* __restrict__ tx, int64_t n_nodes,
                                                        unsigned long long* __restrict__ keys, int* __restrict__ vals,
                                                        int64_t cap, int* __restrict__ n_dup) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n_nodes) return;
  const unsigned long long key = (unsigned long long)tx[i];
  if (key == kEmpty) {
    if (atomicMax(vals + cap, (int)i) >= 0) atomicAdd(n_dup, 1);
    return;
  }
  const uint64_t mask = (uint64_t)cap - 1;
  uint64_t h = mix64(key) & mask;
  for (int64_t probe = 0; probe < cap; ++probe, h = (h + 1) & mask) {
    const unsigned long long prev = atomicCAS(keys + h, kEmpty, key);
    if (prev == kEmpty || prev == key) {
      if (atomicMax(vals + h, (int)i) >= 0) atomicAdd(n_dup, 1);   // dict semantics: the last row of a txId wins
      return;
    }
  }
}

__device__ __forceinline__ int join_find(unsigned long long key, const unsigned long long* __restrict__ keys,
                                         const int* __restrict__ vals, int64_t cap) {
  if (key == kEmpty) return vals[cap];
  const uint64_t mask = (uint64_t)cap - 1;
  uint64_t h = mix64(key) & mask;
  for (int64_t probe = 0; probe < cap; ++probe, h = (h + 1) & mask) {
    const unsigned long long k = keys[h];
    if (k == key) return vals[h];
    if (k == kEmpty) return -1;
  }
  return -1;
}

__global__ void __launch_bounds__(kThreads) join_probe(const int64_t* __restrict__ e_src, const int64_t* __restrict__ e_dst,
                                                       int64_t E, const unsigned long long* __restrict__ keys,
                                                       const int* __restrict__ vals, int64_t cap,
                                                       const int64_t* __restrict__ timestep, int* __restrict__ keep,
                                                       int* __restrict__ s32, int* __restrict__ d32,
                                                       int* __restrict__ n_mapped) {
  const int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  bool mapped = false;
  if (e < E) {
    const int s = join_find((unsigned long long)e_src[e], keys, vals, cap);
    const int d = join_find((unsigned long long)e_dst[e], keys, vals, cap);
    mapped = s >= 0 && d >= 0;
    keep[e] = (mapped && timestep[s] == timestep[d]) ? 1 : 0;
    s32[e] = s;
    d32[e] = d;
  }
  const int cnt = __syncthreads_count(mapped);
  if (threadIdx.x == 0 && cnt) atomicAdd(n_mapped, cnt);
}

__global__ void __launch_bounds__(kThreads) join_compact(const int* __restrict__ keep, const int* __restrict__ pos,
                                                         const int* __restrict__ s32, const int* __restrict__ d32,
                                                         int64_t E, int64_t* __restrict__ out) {
  const int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (e >= E || !keep[e]) return;
  const int p = pos[e];
  out[p] = s32[e];
  out[E + p] = d32[e];
}

// make_temporal_masks (src/data/dataset_elliptic.py:268-290); window_k < 0 = None
__global__ void __launch_bounds__(kThreads) temporal_masks(const int64_t* __restrict__ y, const int64_t* __restrict__ t,
                                                           int64_t n, int64_t t_train_end, int64_t t_val_end,
                                                           int64_t window_k, uint8_t* __restrict__ train,
                                                           uint8_t* __restrict__ val, uint8_t* __restrict__ test) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  const bool labeled = y[i] >= 0;
  const int64_t ti = t[i];
  bool tr = ti <= t_train_end && labeled;
  if (window_k >= 0) {
    int64_t lo = t_train_end - window_k + 1;
    if (lo < 1) lo = 1;
    tr = ti >= lo && ti <= t_train_end && labeled;
  }
  train[i] = tr;
  val[i] = ti > t_train_end && ti <= t_val_end && labeled;
  test[i] = ti > t_val_end && labeled;
}

}  // namespace
}  // namespace egnn

using namespace egnn;

extern "C" size_t egnn_txid_join_workspace_bytes(int64_t n_nodes, int64_t n_edges_raw) {
  return carve_join(nullptr, n_nodes, n_edges_raw).bytes + 256;
}

extern "C" int egnn_txid_join(const int64_t* tx_ids, const int64_t* timestep, int64_t n_nodes, const int64_t* e_src_tx,
                              const int64_t* e_dst_tx, int64_t n_edges_raw, int64_t* edge_index_out, int32_t* info,
                              void* workspace, size_t workspace_bytes, void* stream) {
  const char* fn = "egnn_txid_join";
  EGNN_REQUIRE(n_nodes > 0 && n_edges_raw >= 0, fn, "bad sizes");
  EGNN_REQUIRE(tx_ids && timestep && info && workspace, fn, "null pointer");
  EGNN_REQUIRE(n_edges_raw == 0 || (e_src_tx && e_dst_tx && edge_index_out), fn, "null edge pointer");
  EGNN_REQUIRE(n_nodes < (int64_t)1073741824 && n_edges_raw < (int64_t)2147483647, fn, "too large for int32 indices");
  EGNN_REQUIRE(workspace_bytes >= egnn_txid_join_workspace_bytes(n_nodes, n_edges_raw), fn, "workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  char* base = reinterpret_cast<char*>(((uintptr_t)workspace + 255) & ~uintptr_t(255));
  JoinWs w = carve_join(base, n_nodes, n_edges_raw);
  // keys = all-ones (empty), vals = -1: one fill covers both (the arrays are adjacent up to alignment padding)
  cudaMemsetAsync(w.keys, 0xff, sizeof(unsigned long long) * (size_t)w.cap, st);
  cudaMemsetAsync(w.vals, 0xff, sizeof(int) * (size_t)(w.cap + 1), st);
  cudaMemsetAsync(info, 0, sizeof(int) * 3, st);
  join_insert<<<(unsigned)ceil_div(n_nodes, kThreads), kThreads, 0, st>>>(tx_ids, n_nodes, w.keys, w.vals, w.cap, info + 2);
  EGNN_LAUNCH_CHECK(fn);
  if (n_edges_raw == 0) return 0;
  const unsigned gE = (unsigned)ceil_div(n_edges_raw, kThreads);
  join_probe<<<gE, kThreads, 0, st>>>(e_src_tx, e_dst_tx, n_edges_raw, w.keys, w.vals, w.cap, timestep, w.keep, w.s32,
                                      w.d32, info + 1);
  EGNN_LAUNCH_CHECK(fn);
  int rc = exclusive_scan(w.keep, w.pos, n_edges_raw, w.tile_sums, info, st);   // info[0] = edges kept
  if (rc) return rc;
  join_compact<<<gE, kThreads, 0, st>>>(w.keep, w.pos, w.s32, w.d32, n_edges_raw, edge_index_out);
  EGNN_LAUNCH_CHECK(fn);
  return 0;
}

extern "C" int egnn_temporal_masks(const int64_t* y, const int64_t* timestep, int64_t n_nodes, int64_t t_train_end,
                                   int64_t t_val_end, int64_t train_window_k, uint8_t* train_mask, uint8_t* val_mask,
                                   uint8_t* test_mask, void* stream) {
  const char* fn = "egnn_temporal_masks";
  EGNN_REQUIRE(n_nodes >= 0, fn, "bad size");
  if (n_nodes == 0) return 0;
  EGNN_REQUIRE(y && timestep && train_mask && val_mask && test_mask, fn, "null pointer");
  temporal_masks<<<(unsigned)ceil_div(n_nodes, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
      y, timestep, n_nodes, t_train_end, t_val_end, train_window_k, train_mask, val_mask, test_mask);
  EGNN_LAUNCH_CHECK(fn);
  return 0;
}
