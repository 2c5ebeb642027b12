// K1 -- graph build on device: symmetrise, self-loop rewrite, stable destination/source
// sorted views (CSR / CSC), degree and gcn_norm.  Integer outputs are bit-exact with the
// reference's edge lists; float weights are bit-exact with torch CPU `deg.pow_(-0.5)`
// (SURVEY.md F10).  Replaces src/train_gnn.py:319-326 and PyG add_remaining_self_loops /
// gcn_norm, which the reference re-runs inside every GCNConv/GATConv forward.
//
// Algorithm: expand to an int32 edge list -> LSD radix sort (8-bit digits, stable
// in-tile ranking with __match_any_sync) of (key = dst | src, value = edge id) ->
// histogram + exclusive scan for the row pointers.  All kernels are HBM-bound integer
// streaming passes; grids are sized from the host-known capacity, the true edge count
// lives in info[0] on the device so no host synchronisation is needed.
#include "common.cuh"
#include "radix.cuh"

namespace egnn {
namespace {

// ---- edge-list expansion ---------------------------------------------------------------
// logical edge i of cat([ei, ei.flip(0)]) : i <  E -> (ei[0,i],   ei[1,i])
//                                           i >= E -> (ei[1,i-E], ei[0,i-E])
__device__ __forceinline__ void logical_edge(const int64_t* __restrict__ ei, int64_t E, int64_t i,
                                             int64_t& s, int64_t& d) {
  if (i < E) {
    s = ei[i];
    d = ei[E + i];
  } else {
    s = ei[E + (i - E)];
    d = ei[i - E];
  }
}

// no self-loop rewrite: straight int64 -> int32 conversion (+ range validation)
__global__ void __launch_bounds__(kThreads) expand_plain(const int64_t* __restrict__ ei, int64_t E,
                                                         int64_t E_log, int64_t n_nodes,
                                                         int* __restrict__ src32,
                                                         int* __restrict__ dst32,
                                                         int* __restrict__ info) {
  int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i == 0) info[0] = (int)E_log;
  if (i >= E_log) return;
  int64_t s, d;
  logical_edge(ei, E, i, s, d);
  bool bad = (s < 0) | (s >= n_nodes) | (d < 0) | (d >= n_nodes);
  if (bad) {
    atomicAdd(&info[1], 1);
    s = 0;
    d = 0;
  }
  src32[i] = (int)s;
  dst32[i] = (int)d;
}

// self-loop rewrite, pass 1: keep flags (src != dst)
__global__ void __launch_bounds__(kThreads) loop_flags(const int64_t* __restrict__ ei, int64_t E,
                                                       int64_t E_log, int64_t n_nodes,
                                                       int* __restrict__ keep,
                                                       int* __restrict__ info) {
  int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= E_log) return;
  int64_t s, d;
  logical_edge(ei, E, i, s, d);
  bool bad = (s < 0) | (s >= n_nodes) | (d < 0) | (d >= n_nodes);
  if (bad) atomicAdd(&info[1], 1);
  keep[i] = (!bad && s != d) ? 1 : 0;
}

// pass 2: stable compaction of the kept edges, then N self loops; info[0] = E_nl + N
__global__ void __launch_bounds__(kThreads) loop_compact(const int64_t* __restrict__ ei, int64_t E,
                                                         int64_t E_log, int64_t n_nodes,
                                                         const int* __restrict__ pos,
                                                         const int* __restrict__ n_kept,
                                                         int* __restrict__ src32,
                                                         int* __restrict__ dst32,
                                                         int* __restrict__ info) {
  int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  const int e_nl = *n_kept;
  if (i == 0) info[0] = e_nl + (int)n_nodes;
  if (i < E_log) {
    int64_t s, d;
    logical_edge(ei, E, i, s, d);
    bool ok = (s >= 0) & (s < n_nodes) & (d >= 0) & (d < n_nodes) & (s != d);
    if (ok) {
      int p = pos[i];
      src32[p] = (int)s;
      dst32[p] = (int)d;
    }
  }
  if (i < n_nodes) {
    src32[e_nl + i] = (int)i;
    dst32[e_nl + i] = (int)i;
  }
}

__global__ void __launch_bounds__(kThreads) count_keys(const int* __restrict__ keys,
                                                       const int* __restrict__ n_ptr,
                                                       int* __restrict__ counts, int64_t cap) {
  int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i < cap && i < *n_ptr) atomicAdd(&counts[keys[i]], 1);
}

__global__ void __launch_bounds__(kThreads) gather_csr(const int* __restrict__ perm,
                                                       const int* __restrict__ src32,
                                                       const int* __restrict__ n_ptr, int64_t cap,
                                                       int* __restrict__ csr_src,
                                                       int* __restrict__ csr_eid,
                                                       int* __restrict__ inv) {
  int64_t p = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (p >= cap) return;
  if (p < *n_ptr) {
    int e = perm[p];
    csr_eid[p] = e;
    csr_src[p] = src32[e];
    inv[e] = (int)p;
  } else {  // defined contents in the unused tail
    csr_eid[p] = -1;
    csr_src[p] = 0;
  }
}

__global__ void __launch_bounds__(kThreads) gather_csc(const int* __restrict__ perm,
                                                       const int* __restrict__ dst32,
                                                       const int* __restrict__ inv,
                                                       const int* __restrict__ n_ptr, int64_t cap,
                                                       int* __restrict__ csc_dst,
                                                       int* __restrict__ csc_pos) {
  int64_t q = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (q >= cap) return;
  if (q < *n_ptr) {
    int e = perm[q];
    csc_dst[q] = dst32[e];
    csc_pos[q] = inv[e];
  } else {
    csc_dst[q] = 0;
    csc_pos[q] = 0;
  }
}

__global__ void __launch_bounds__(kThreads) write_ei2(const int* __restrict__ src32,
                                                      const int* __restrict__ dst32,
                                                      const int* __restrict__ n_ptr, int64_t cap,
                                                      int64_t* __restrict__ ei2) {
  int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= cap) return;
  bool v = i < *n_ptr;
  ei2[i] = v ? src32[i] : -1;
  ei2[cap + i] = v ? dst32[i] : -1;
}

// deg^-1/2 exactly as torch CPU pow_(-0.5): rn(1 / rn(sqrt(deg))), inf -> 0
__global__ void __launch_bounds__(kThreads) deg_inv_sqrt(const int* __restrict__ csr_ptr,
                                                         int64_t n_nodes, float* __restrict__ dis) {
  int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n_nodes) return;
  int d = csr_ptr[i + 1] - csr_ptr[i];
  dis[i] = d > 0 ? __fdiv_rn(1.0f, __fsqrt_rn((float)d)) : 0.0f;
}

__global__ void __launch_bounds__(kThreads) norm_weights(const int* __restrict__ src32,
                                                         const int* __restrict__ dst32,
                                                         const int* __restrict__ inv,
                                                         const int* __restrict__ n_ptr, int64_t cap,
                                                         const float* __restrict__ dis,
                                                         float* __restrict__ w_edge,
                                                         float* __restrict__ w_csr) {
  int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (e >= cap) return;
  if (e < *n_ptr) {
    float w = __fmul_rn(dis[src32[e]], dis[dst32[e]]);  // dis[row]*1*dis[col]
    if (w_edge) w_edge[e] = w;
    if (w_csr) w_csr[inv[e]] = w;
  } else if (w_edge) {
    w_edge[e] = 0.f;
  }
}

__global__ void __launch_bounds__(kThreads) permute_weights(const float* __restrict__ w_csr,
                                                            const int* __restrict__ csc_pos,
                                                            const int* __restrict__ n_ptr,
                                                            int64_t cap, float* __restrict__ w_csc) {
  int64_t q = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (q >= cap) return;
  w_csc[q] = (q < *n_ptr) ? w_csr[csc_pos[q]] : 0.f;
}

// rows with more than kLongRow entries: appended (order irrelevant) to a list the SpMM serves
// with whole CTAs; count -> *n_long
__global__ void __launch_bounds__(kThreads) collect_long_rows(const int* __restrict__ ptr, int64_t n_nodes,
                                                              int threshold, int* __restrict__ list,
                                                              int* __restrict__ n_long) {
  int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n_nodes) return;
  if (ptr[i + 1] - ptr[i] > threshold) list[atomicAdd(n_long, 1)] = (int)i;
}

// key for the longest-rows-first schedule, applied inside tiles of 32768 consecutive rows so that a
// wave of CTAs still works on one timestep's feature rows (L2 locality) while every tile starts
// with its longest rows: key = tile << 7 | (64 - min(deg, 64)); stable sort => deterministic
__global__ void __launch_bounds__(kThreads) degree_keys(const int* __restrict__ ptr, int64_t n_nodes,
                                                        int* __restrict__ keys, int* __restrict__ n_dev) {
  int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i == 0) *n_dev = (int)n_nodes;
  if (i >= n_nodes) return;
  int d = ptr[i + 1] - ptr[i];
  keys[i] = (int)((i >> 15) << 7) | (64 - (d < 64 ? d : 64));
}
__global__ void __launch_bounds__(kThreads) copy_ints(const int* __restrict__ src, int64_t n, int* __restrict__ dst) {
  int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i < n) dst[i] = src[i];
}

inline size_t align_up(size_t v) { return (v + 255) & ~size_t(255); }

struct Workspace {
  int *src32, *dst32, *keysA, *keysB, *valsA, *valsB, *inv, *keep, *counts, *table, *tile_sums, *deg_keys, *n_dev;
  size_t bytes;
};

Workspace carve(char* base, int64_t n_nodes, int64_t E_log, int64_t cap) {
  Workspace w;
  size_t off = 0;
  auto take = [&](int64_t n_int) {
    int* p = reinterpret_cast<int*>(base + off);
    off += align_up((size_t)n_int * sizeof(int));
    return p;
  };
  int64_t sort_n = cap > n_nodes ? cap : n_nodes;  // the row-order sort re-uses the radix buffers
  int64_t nblk = ceil_div(sort_n, kSortTile);
  int64_t table_n = 256 * nblk;
  int64_t longest = cap;
  if (n_nodes + 1 > longest) longest = n_nodes + 1;
  if (table_n > longest) longest = table_n;
  w.src32 = take(cap);
  w.dst32 = take(cap);
  w.keysA = take(sort_n);
  w.keysB = take(sort_n);
  w.valsA = take(sort_n);
  w.valsB = take(sort_n);
  w.deg_keys = take(n_nodes);
  w.n_dev = take(4);
  w.inv = take(cap);
  w.keep = take(E_log + 1);
  w.counts = take(n_nodes + 2);
  w.table = take(table_n);
  w.tile_sums = take(ceil_div(longest, kScanTile) + 1);
  w.bytes = off;
  return w;
}

// stable sort of (keys, iota) by key over `bits` bits; result permutation in *perm_out
int radix_sort_perm(const int* keys, const int* n_ptr, int64_t cap, int bits, Workspace& w,
                    int** perm_out, cudaStream_t st) {
  int nblk = (int)ceil_div(cap, kSortTile);
  const int* kin = keys;
  const int* vin = nullptr;
  int* kout = w.keysA;
  int* vout = w.valsA;
  for (int shift = 0; shift < bits; shift += 8) {
    radix_hist<<<nblk, kThreads, 0, st>>>(kin, n_ptr, shift, w.table, nblk);
    EGNN_LAUNCH_CHECK("radix_hist");
    int rc = exclusive_scan(w.table, w.table, (int64_t)256 * nblk, w.tile_sums, nullptr, st);
    if (rc) return rc;
    radix_scatter<<<nblk, kThreads, 0, st>>>(kin, vin, kout, vout, n_ptr, shift, w.table, nblk);
    EGNN_LAUNCH_CHECK("radix_scatter");
    kin = kout;
    vin = vout;
    kout = (kout == w.keysA) ? w.keysB : w.keysA;
    vout = (vout == w.valsA) ? w.valsB : w.valsA;
  }
  *perm_out = const_cast<int*>(vin);
  return 0;
}


// ---- re-entrant graph variants ------------------------------------------------------------------
// Hub ablation (src/train_gnn.py:526-540, src/analysis/hub_ablation.py:56-71): deg = bincount(src) + bincount(dst);
// hubs = the num_hubs nodes of largest degree; keep the edges touching no hub, in their original order.  The reference
// round-trips the edge list through the CPU (`.cpu()`, torch.topk, boolean-mask indexing); here: degree histogram
// (integer atomics) -> stable radix sort of (max_deg - deg) -> the first num_hubs nodes -> edge flags -> exclusive scan
// -> stable compaction.  Ties at the k-th degree: LOWER node id first (torch.topk leaves ties unspecified -- its CPU
// kernel switches between partial_sort and nth_element with n and k -- so the hub set equals the reference's whenever
// the k-th and (k+1)-th largest degrees differ, and is the stable-sort choice otherwise).
__global__ void __launch_bounds__(kThreads) abl_degree(const int64_t* __restrict__ ei, int64_t E, int64_t n_nodes,
                                                       int* __restrict__ deg, int* __restrict__ n_bad) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= E) return;
  const int64_t s = ei[i], d = ei[E + i];
  if (s < 0 || s >= n_nodes || d < 0 || d >= n_nodes) {
    atomicAdd(n_bad, 1);
    return;
  }
  atomicAdd(deg + s, 1);
  atomicAdd(deg + d, 1);
}
__global__ void __launch_bounds__(kThreads) abl_keys(const int* __restrict__ deg, int64_t n_nodes, int max_key,
                                                     int* __restrict__ keys, int* __restrict__ n_dev) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i == 0) *n_dev = (int)n_nodes;
  if (i < n_nodes) keys[i] = max_key - deg[i];   // ascending key = descending degree; the sort is stable in node id
}
__global__ void __launch_bounds__(kThreads) abl_mark(const int* __restrict__ perm, int64_t num_hubs,
                                                     uint8_t* __restrict__ hub) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i < num_hubs) hub[perm[i]] = 1;
}
__global__ void __launch_bounds__(kThreads) abl_flags(const int64_t* __restrict__ ei, int64_t E, int64_t n_nodes,
                                                      const uint8_t* __restrict__ hub, int* __restrict__ keep) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= E) return;
  const int64_t s = ei[i], d = ei[E + i];
  const bool ok = s >= 0 && s < n_nodes && d >= 0 && d < n_nodes;
  keep[i] = (ok && !(hub[s] | hub[d])) ? 1 : 0;
}
__global__ void __launch_bounds__(kThreads) abl_compact(const int64_t* __restrict__ ei, int64_t E,
                                                        const int* __restrict__ keep, const int* __restrict__ pos,
                                                        int64_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= E || !keep[i]) return;
  const int p = pos[i];
  out[p] = ei[i];
  out[E + p] = ei[E + i];
}
// Random edge drop (src/analysis/robustness.py:65-82): out[:, j] = ei[:, idx[j]] with idx = perm[drop_count:]
__global__ void __launch_bounds__(kThreads) edge_gather(const int64_t* __restrict__ ei, int64_t E,
                                                        const int64_t* __restrict__ idx, int64_t n_idx,
                                                        int64_t* __restrict__ out, int* __restrict__ n_bad) {
  const int64_t j = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (j >= n_idx) return;
  const int64_t i = idx[j];
  if (i < 0 || i >= E) {
    atomicAdd(n_bad, 1);
    out[j] = 0;
    out[n_idx + j] = 0;
    return;
  }
  out[j] = ei[i];
  out[n_idx + j] = ei[E + i];
}

// byte-wise comparison of two device buffers (train.HostFeed: rebuild the sorted views only when the edge list a
// caller submits really differs from the one they were built from)
__global__ void __launch_bounds__(kThreads) buffers_differ_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b,
                                                                  int64_t n16, const uint8_t* __restrict__ ta,
                                                                  const uint8_t* __restrict__ tb, int tail,
                                                                  int* __restrict__ flag) {
  bool diff = false;
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n16; i += stride) {
    const uint4 u = a[i], v = b[i];
    diff |= (u.x != v.x) | (u.y != v.y) | (u.z != v.z) | (u.w != v.w);
  }
  if (blockIdx.x == 0 && (int)threadIdx.x < tail) diff |= ta[threadIdx.x] != tb[threadIdx.x];
  if (__syncthreads_or(diff) && threadIdx.x == 0) *flag = 1;
}

__global__ void clear_flag_kernel(int* flag) { *flag = 0; }

}  // namespace
}  // namespace egnn

using namespace egnn;

extern "C" int egnn_buffers_differ(const void* a, const void* b, int64_t n_bytes, int* flag, void* stream) {
  const char* fn = "egnn_buffers_differ";
  EGNN_REQUIRE(a && b && flag && n_bytes >= 0, fn, "bad arguments");
  EGNN_REQUIRE((uintptr_t)a % 16 == 0 && (uintptr_t)b % 16 == 0, fn, "buffers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  clear_flag_kernel<<<1, 1, 0, st>>>(flag);
  EGNN_LAUNCH_CHECK(fn);
  if (n_bytes == 0) return 0;
  const int64_t n16 = n_bytes / 16;
  const int tail = (int)(n_bytes - 16 * n16);
  int64_t blocks = ceil_div(n16 > 0 ? n16 : 1, (int64_t)kThreads * 4);
  if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
  buffers_differ_kernel<<<(unsigned)blocks, kThreads, 0, st>>>(
      reinterpret_cast<const uint4*>(a), reinterpret_cast<const uint4*>(b), n16,
      reinterpret_cast<const uint8_t*>(a) + 16 * n16, reinterpret_cast<const uint8_t*>(b) + 16 * n16, tail, flag);
  EGNN_LAUNCH_CHECK(fn);
  return 0;
}

extern "C" size_t egnn_hub_ablation_workspace_bytes(int64_t n_nodes, int64_t n_edges) {
  Workspace w = carve(nullptr, n_nodes, n_edges, n_edges > 0 ? n_edges : 1);
  return w.bytes + 256;
}

extern "C" int egnn_hub_ablation(const int64_t* ei, int64_t E, int64_t n_nodes, int64_t num_hubs, int64_t* ei_out,
                                 int32_t* info, uint8_t* hub_mask, int32_t* deg_out, void* workspace,
                                 size_t workspace_bytes, void* stream) {
  const char* fn = "egnn_hub_ablation";
  EGNN_REQUIRE(n_nodes > 0 && E >= 0 && num_hubs >= 0 && num_hubs <= n_nodes, fn, "bad sizes");
  EGNN_REQUIRE((E == 0 || (ei && ei_out)) && info && hub_mask && workspace, fn, "null pointer");
  EGNN_REQUIRE(2 * E < (int64_t)2147483647 && n_nodes < (int64_t)2147483647, fn, "graph too large for int32 indices");
  EGNN_REQUIRE(workspace_bytes >= egnn_hub_ablation_workspace_bytes(n_nodes, E), fn, "workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  char* base = reinterpret_cast<char*>(((uintptr_t)workspace + 255) & ~uintptr_t(255));
  Workspace w = carve(base, n_nodes, E, E > 0 ? E : 1);
  int* deg = deg_out ? deg_out : w.counts;
  cudaMemsetAsync(deg, 0, sizeof(int) * (size_t)n_nodes, st);
  cudaMemsetAsync(info, 0, sizeof(int) * 2, st);
  cudaMemsetAsync(hub_mask, 0, (size_t)n_nodes, st);
  const unsigned gE = (unsigned)ceil_div(E > 0 ? E : 1, kThreads), gN = (unsigned)ceil_div(n_nodes, kThreads);
  if (E > 0) {
    abl_degree<<<gE, kThreads, 0, st>>>(ei, E, n_nodes, deg, info + 1);
    EGNN_LAUNCH_CHECK(fn);
  }
  if (num_hubs > 0) {
    const int max_key = (int)(2 * E);
    int bits = 1;
    while (((int64_t)1 << bits) <= max_key) ++bits;
    bits = (bits + 7) / 8 * 8;
    abl_keys<<<gN, kThreads, 0, st>>>(deg, n_nodes, max_key, w.deg_keys, w.n_dev);
    EGNN_LAUNCH_CHECK(fn);
    int* perm = nullptr;
    int rc = radix_sort_perm(w.deg_keys, w.n_dev, n_nodes, bits, w, &perm, st);
    if (rc) return rc;
    abl_mark<<<(unsigned)ceil_div(num_hubs, kThreads), kThreads, 0, st>>>(perm, num_hubs, hub_mask);
    EGNN_LAUNCH_CHECK(fn);
  }
  if (E > 0) {
    abl_flags<<<gE, kThreads, 0, st>>>(ei, E, n_nodes, hub_mask, w.keep);
    EGNN_LAUNCH_CHECK(fn);
    int rc = exclusive_scan(w.keep, w.inv, E, w.tile_sums, info, st);   // info[0] = edges kept
    if (rc) return rc;
    abl_compact<<<gE, kThreads, 0, st>>>(ei, E, w.keep, w.inv, ei_out);
    EGNN_LAUNCH_CHECK(fn);
  }
  return 0;
}

extern "C" int egnn_edge_gather(const int64_t* ei, int64_t E, const int64_t* idx, int64_t n_idx, int64_t* out,
                                int32_t* n_bad, void* stream) {
  const char* fn = "egnn_edge_gather";
  EGNN_REQUIRE(n_idx >= 0 && E >= 0 && n_bad, fn, "bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(n_bad, 0, sizeof(int), st);
  if (n_idx == 0) return 0;
  EGNN_REQUIRE(ei && idx && out, fn, "null pointer");
  edge_gather<<<(unsigned)ceil_div(n_idx, kThreads), kThreads, 0, st>>>(ei, E, idx, n_idx, out, n_bad);
  EGNN_LAUNCH_CHECK(fn);
  return 0;
}

extern "C" size_t egnn_graph_workspace_bytes(int64_t n_nodes, int64_t n_edges_in, int flags) {
  int64_t E_log = (flags & EGNN_G_SYMMETRIZE) ? 2 * n_edges_in : n_edges_in;
  int64_t cap = E_log + ((flags & EGNN_G_SELF_LOOPS) ? n_nodes : 0);
  if (cap < 1) cap = 1;
  Workspace w = carve(nullptr, n_nodes, E_log, cap);
  return w.bytes + 256;
}

extern "C" int egnn_graph_build(const int64_t* ei, int64_t E, int64_t n_nodes, int flags,
                                int want_norm, int32_t* info, int32_t* csr_ptr, int32_t* csr_src,
                                int32_t* csr_eid, int32_t* csc_ptr, int32_t* csc_dst,
                                int32_t* csc_pos, int32_t* csr_long, int32_t* csc_long,
                                int32_t* csr_order, int32_t* csc_order, int64_t* ei2, float* dis, float* w_edge,
                                float* w_csr, float* w_csc, void* workspace,
                                size_t workspace_bytes, void* stream) {
  const char* fn = "egnn_graph_build";
  cudaStream_t st = (cudaStream_t)stream;
  EGNN_REQUIRE(n_nodes > 0 && E >= 0, fn, "n_nodes must be > 0 and n_edges >= 0");
  EGNN_REQUIRE(E == 0 || ei != nullptr, fn, "edge_index is null");
  EGNN_REQUIRE(info && csr_ptr && csr_src && csr_eid && csc_ptr && csc_dst && csc_pos, fn,
               "null output pointer");
  const bool sym = flags & EGNN_G_SYMMETRIZE, loops = flags & EGNN_G_SELF_LOOPS;
  const int64_t E_log = sym ? 2 * E : E;
  int64_t cap = E_log + (loops ? n_nodes : 0);
  EGNN_REQUIRE(cap < (int64_t)2147483647 && n_nodes < (int64_t)2147483647, fn,
               "graph too large for int32 indices");
  EGNN_REQUIRE(workspace_bytes >= egnn_graph_workspace_bytes(n_nodes, E, flags), fn,
               "workspace too small");
  const bool norm = want_norm || loops;
  if (norm) EGNN_REQUIRE(dis && w_csr && w_csc, fn, "norm outputs are null");
  const int64_t cap1 = cap < 1 ? 1 : cap;
  char* wbase = reinterpret_cast<char*>(((uintptr_t)workspace + 255) & ~uintptr_t(255));
  Workspace w = carve(wbase, n_nodes, E_log, cap1);
  int* n_ptr = info;  // info[0] = E2

  cudaMemsetAsync(info, 0, 4 * sizeof(int), st);
  const unsigned gE = (unsigned)ceil_div(E_log > 0 ? E_log : 1, kThreads);
  const unsigned gC = (unsigned)ceil_div(cap1, kThreads);
  if (!loops) {
    expand_plain<<<gE, kThreads, 0, st>>>(ei, E, E_log, n_nodes, w.src32, w.dst32, info);
    EGNN_LAUNCH_CHECK("expand_plain");
  } else {
    if (E_log > 0) {
      loop_flags<<<gE, kThreads, 0, st>>>(ei, E, E_log, n_nodes, w.keep, info);
      EGNN_LAUNCH_CHECK("loop_flags");
    }
    // exclusive scan over E_log+1 entries: entry E_log receives the kept count
    cudaMemsetAsync(w.keep + E_log, 0, sizeof(int), st);
    int rc = exclusive_scan(w.keep, w.keep, E_log + 1, w.tile_sums, nullptr, st);
    if (rc) return rc;
    int64_t m = E_log > n_nodes ? E_log : n_nodes;
    loop_compact<<<(unsigned)ceil_div(m, kThreads), kThreads, 0, st>>>(
        ei, E, E_log, n_nodes, w.keep, w.keep + E_log, w.src32, w.dst32, info);
    EGNN_LAUNCH_CHECK("loop_compact");
  }
  int bits = 1;
  while (((int64_t)1 << bits) < n_nodes) ++bits;

  for (int view = 0; view < 2; ++view) {  // 0: CSR by dst, 1: CSC by src
    const int* keys = view == 0 ? w.dst32 : w.src32;
    int* ptr = view == 0 ? csr_ptr : csc_ptr;
    cudaMemsetAsync(w.counts, 0, (size_t)(n_nodes + 2) * sizeof(int), st);
    count_keys<<<gC, kThreads, 0, st>>>(keys, n_ptr, w.counts, cap1);
    EGNN_LAUNCH_CHECK("count_keys");
    int rc = exclusive_scan(w.counts, ptr, n_nodes + 1, w.tile_sums, nullptr, st);
    if (rc) return rc;
    int* perm = nullptr;
    rc = radix_sort_perm(keys, n_ptr, cap1, bits, w, &perm, st);
    if (rc) return rc;
    int* long_list = view == 0 ? csr_long : csc_long;
    if (long_list) {
      collect_long_rows<<<(unsigned)ceil_div(n_nodes, kThreads), kThreads, 0, st>>>(ptr, n_nodes, 64, long_list,
                                                                                  info + 2 + view);
      EGNN_LAUNCH_CHECK("collect_long_rows");
    }
    int* order = view == 0 ? csr_order : csc_order;
    if (order) {
      const unsigned gN = (unsigned)ceil_div(n_nodes, kThreads);
      degree_keys<<<gN, kThreads, 0, st>>>(ptr, n_nodes, w.deg_keys, w.n_dev);
      EGNN_LAUNCH_CHECK("degree_keys");
      // consume the edge permutation before the radix buffers are re-used
      if (view == 0) {
        gather_csr<<<gC, kThreads, 0, st>>>(perm, w.src32, n_ptr, cap1, csr_src, csr_eid, w.inv);
        EGNN_LAUNCH_CHECK("gather_csr");
      } else {
        gather_csc<<<gC, kThreads, 0, st>>>(perm, w.dst32, w.inv, n_ptr, cap1, csc_dst, csc_pos);
        EGNN_LAUNCH_CHECK("gather_csc");
      }
      int* rperm = nullptr;
      int obits = 7;
      while (((int64_t)1 << (obits - 7)) < ceil_div(n_nodes, 32768)) ++obits;
      rc = radix_sort_perm(w.deg_keys, w.n_dev, n_nodes, obits, w, &rperm, st);
      if (rc) return rc;
      copy_ints<<<gN, kThreads, 0, st>>>(rperm, n_nodes, order);
      EGNN_LAUNCH_CHECK("copy_ints");
      continue;
    }
    if (view == 0) {
      gather_csr<<<gC, kThreads, 0, st>>>(perm, w.src32, n_ptr, cap1, csr_src, csr_eid, w.inv);
      EGNN_LAUNCH_CHECK("gather_csr");
    } else {
      gather_csc<<<gC, kThreads, 0, st>>>(perm, w.dst32, w.inv, n_ptr, cap1, csc_dst, csc_pos);
      EGNN_LAUNCH_CHECK("gather_csc");
    }
  }
  if (ei2) {
    write_ei2<<<gC, kThreads, 0, st>>>(w.src32, w.dst32, n_ptr, cap1, ei2);
    EGNN_LAUNCH_CHECK("write_ei2");
  }
  if (norm) {
    deg_inv_sqrt<<<(unsigned)ceil_div(n_nodes, kThreads), kThreads, 0, st>>>(csr_ptr, n_nodes, dis);
    EGNN_LAUNCH_CHECK("deg_inv_sqrt");
    norm_weights<<<gC, kThreads, 0, st>>>(w.src32, w.dst32, w.inv, n_ptr, cap1, dis, w_edge, w_csr);
    EGNN_LAUNCH_CHECK("norm_weights");
    permute_weights<<<gC, kThreads, 0, st>>>(w_csr, csc_pos, n_ptr, cap1, w_csc);
    EGNN_LAUNCH_CHECK("permute_weights");
  }
  return 0;
}
