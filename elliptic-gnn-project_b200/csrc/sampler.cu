// SURVEY.md 8(f) rank 4, last item -- the mini-batch path of the reference (src/train_gnn.py:329-348,212-245):
//   NeighborLoader(data, num_neighbors=fanout, batch_size=..., input_nodes=train_idx, shuffle=True)
// hands `train_epoch_minibatch` one sampled subgraph per batch of seed nodes: batch.x / y / timestep = rows of the
// sampled nodes (seeds first), batch.edge_index = the sampled edges in LOCAL ids, batch.batch_size = #seeds.
//
// What PyG's loader does per batch (torch_geometric.loader.NeighborLoader -> pyg-lib `neighbor_sample`, homogeneous graph,
// replace=False, directed, not disjoint; restated in oracle/neighbor_sample_np.py):
//   nodes = seeds;  for every hop h with fan-out k_h:  for every node v added in the previous hop (in order):
//     take all in-neighbours of v if indeg(v) <= k_h (or k_h < 0), else k_h distinct ones (Robert Floyd's algorithm over
//     the positions of v's CSC row); record the edge (u -> v); append u to `nodes` unless it is already there.
//   local id of a node = its position in `nodes`.
// PyG samples on the CPU with std::mt19937 (a stream that cannot be matched); here the whole batch is sampled on the
// device from the CSR-by-destination view the graph build already produced (stable: a row lists the in-edges in their
// original order = PyG's CSC), with Philox4x32-10 keyed on (seed, batch index, hop, local id of v, draw index), so a
// batch is a pure function of its inputs and the oracle reproduces it bit for bit.
//
// Per hop: count -> scan -> pick (+ atomicMin of the edge slot into first_pos[u] for nodes not yet in the batch) ->
// flag first occurrences -> scan -> assign local ids in slot order (= the sequential algorithm's order of first
// appearance).  All lengths live on the device (grids are sized from host-side upper bounds); nothing syncs.
// State: local_id[N] (-1 = not in the batch) and first_pos[N] (INT_MAX), reset for the touched nodes only at the end.
#include "philox.cuh"
#include "radix.cuh"

namespace egnn {
namespace {

constexpr int kMaxHops = 8;

struct SampleWs {
  int* cnt;        // [F_cap_max]     per-frontier-node sample counts, then their exclusive scan (in place)
  int* flag;       // [E_cap_max]     Floyd scratch positions, then first-occurrence flags
  int* rank;       // [E_cap_max]     exclusive scan of the flags
  int* src_g;      // [cap_edges]     global id of every sampled edge's source
  int* dst_l;      // [cap_edges]     local id of its destination
  int* tile_sums;  // scan scratch
  int* totals;     // [2]             this hop's edge count, this hop's new-node count
  size_t bytes;
};

struct Caps {
  int64_t F[kMaxHops], E[kMaxHops], nodes, edges, Fmax, Emax;
};

Caps make_caps(int64_t N, int64_t E, int64_t B, const int32_t* fanouts, int H) {
  Caps c{};
  int64_t F = B < N ? B : N, tot_e = 0;
  c.Fmax = 1; c.Emax = 1;
  for (int h = 0; h < H; ++h) {
    c.F[h] = F;
    int64_t e = fanouts[h] < 0 ? E : F * (int64_t)fanouts[h];
    if (e > E) e = E;
    c.E[h] = e;
    tot_e += e;
    if (F > c.Fmax) c.Fmax = F;
    if (e > c.Emax) c.Emax = e;
    F = e < N ? e : N;      // a hop adds at most one node per sampled edge
  }
  c.edges = tot_e > 0 ? tot_e : 1;
  c.nodes = B + tot_e < N ? B + tot_e : N;
  if (c.nodes < B) c.nodes = B;
  return c;
}

SampleWs carve_sample(char* base, const Caps& c) {
  SampleWs w;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* p = base ? base + off : nullptr;
    off += (bytes + 255) & ~size_t(255);
    return p;
  };
  const int64_t big = c.Fmax > c.Emax ? c.Fmax : c.Emax;
  w.cnt = reinterpret_cast<int*>(take(sizeof(int) * (size_t)c.Fmax));
  w.flag = reinterpret_cast<int*>(take(sizeof(int) * (size_t)c.Emax));
  w.rank = reinterpret_cast<int*>(take(sizeof(int) * (size_t)c.Emax));
  w.src_g = reinterpret_cast<int*>(take(sizeof(int) * (size_t)c.edges));
  w.dst_l = reinterpret_cast<int*>(take(sizeof(int) * (size_t)c.edges));
  w.tile_sums = reinterpret_cast<int*>(take(sizeof(int) * (size_t)(ceil_div(big, kScanTile) + 1)));
  w.totals = reinterpret_cast<int*>(take(sizeof(int) * 2));
  w.bytes = off;
  return w;
}

__global__ void __launch_bounds__(kThreads) ns_state_init(int* __restrict__ local_id, int* __restrict__ first_pos, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i < n) { local_id[i] = -1; first_pos[i] = 0x7fffffff; }
}

// counts: ne[0..H] = number of batch nodes after the seeds / after every hop, then ee[0..H] = number of edges likewise
__global__ void __launch_bounds__(kThreads) ns_seed(const int64_t* __restrict__ seeds, int64_t B, int* __restrict__ local_id,
                                                    int64_t* __restrict__ n_id, int* __restrict__ counts, int H) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i == 0) { counts[0] = (int)B; counts[H + 1] = 0; }
  if (i < B) {
    const int64_t s = seeds[i];
    local_id[s] = (int)i;
    n_id[i] = s;
  }
}

__global__ void __launch_bounds__(kThreads) ns_count(const int* __restrict__ ptr, const int64_t* __restrict__ n_id,
                                                     const int* __restrict__ counts, int h, int k, int64_t F_cap,
                                                     int* __restrict__ cnt) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= F_cap) return;
  const int fb = h == 0 ? 0 : counts[h - 1], fe = counts[h];
  int c = 0;
  if (i < fe - fb) {
    const int64_t v = n_id[fb + i];
    const int d = ptr[v + 1] - ptr[v];
    c = (k < 0 || d <= k) ? d : k;
  }
  cnt[i] = c;
}

__global__ void __launch_bounds__(kThreads) ns_pick(const int* __restrict__ ptr, const int* __restrict__ col,
                                                    const int* __restrict__ eid, const int64_t* __restrict__ n_id,
                                                    const int* __restrict__ counts, int H, int h, int k,
                                                    const int* __restrict__ off, int* __restrict__ tmp, uint64_t seed,
                                                    uint32_t batch_idx, const int* __restrict__ local_id,
                                                    int* __restrict__ first_pos, int* __restrict__ src_g,
                                                    int* __restrict__ dst_l, int64_t* __restrict__ e_id_out, int64_t F_cap) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= F_cap) return;
  const int fb = h == 0 ? 0 : counts[h - 1], fe = counts[h];
  if (i >= fe - fb) return;
  const int vl = fb + (int)i;                 // local id of the frontier node
  const int64_t v = n_id[vl];
  const int row = ptr[v], d = ptr[v + 1] - row;
  const int c = (k < 0 || d <= k) ? d : k;
  const int t0 = off[i];                      // slot of this node's first edge inside the hop
  const int e0 = counts[H + 1 + h] + t0;      // ... and inside the batch
  if (c < d) {
    // Robert Floyd: for jj = d-c .. d-1: t = uniform{0..jj}; take t unless already taken, else jj (insertion order kept)
    Philox4 w{};
    for (int j = 0; j < c; ++j) {
      if ((j & 3) == 0)
        w = philox4x32_10((uint32_t)vl, (uint32_t)(j >> 2), (uint32_t)h, batch_idx, (uint32_t)(seed & 0xffffffffu),
                          (uint32_t)(seed >> 32));
      const int jj = d - c + j;
      int t = (int)__umulhi(w.v[j & 3], (uint32_t)(jj + 1));
      bool dup = false;
      for (int q = 0; q < j; ++q) dup |= tmp[t0 + q] == t;
      tmp[t0 + j] = dup ? jj : t;
    }
  }
  for (int j = 0; j < c; ++j) {
    const int p = row + (c < d ? tmp[t0 + j] : j);
    const int u = col[p];
    src_g[e0 + j] = u;
    dst_l[e0 + j] = vl;
    if (e_id_out) e_id_out[e0 + j] = eid[p];
    if (local_id[u] < 0) atomicMin(&first_pos[u], e0 + j);
  }
}

__global__ void __launch_bounds__(kThreads) ns_flag(const int* __restrict__ counts, int H, int h, const int* __restrict__ totals,
                                                    const int* __restrict__ src_g, const int* __restrict__ local_id,
                                                    const int* __restrict__ first_pos, int* __restrict__ flag, int64_t E_cap) {
  const int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (t >= E_cap) return;
  int f = 0;
  if (t < totals[0]) {
    const int e = counts[H + 1 + h] + (int)t;
    const int u = src_g[e];
    f = (local_id[u] < 0 && first_pos[u] == e) ? 1 : 0;
  }
  flag[t] = f;
}

__global__ void __launch_bounds__(kThreads) ns_assign(int* __restrict__ counts, int H, int h, const int* __restrict__ totals,
                                                      const int* __restrict__ src_g, const int* __restrict__ flag,
                                                      const int* __restrict__ rank, int* __restrict__ local_id,
                                                      int64_t* __restrict__ n_id, int64_t E_cap) {
  const int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (t == 0) {
    counts[h + 1] = counts[h] + totals[1];
    counts[H + 2 + h] = counts[H + 1 + h] + totals[0];
  }
  if (t >= E_cap || t >= totals[0] || !flag[t]) return;
  const int u = src_g[counts[H + 1 + h] + (int)t];
  const int id = counts[h] + rank[t];
  local_id[u] = id;
  n_id[id] = u;
}

__global__ void __launch_bounds__(kThreads) ns_finish(const int* __restrict__ counts, int H, const int* __restrict__ src_g,
                                                      const int* __restrict__ dst_l, const int* __restrict__ local_id,
                                                      int64_t* __restrict__ ei, int64_t cap_edges) {
  const int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (e >= cap_edges || e >= counts[2 * H + 1]) return;
  ei[e] = local_id[src_g[e]];
  ei[cap_edges + e] = dst_l[e];
}

__global__ void __launch_bounds__(kThreads) ns_reset(const int* __restrict__ counts, int H, const int64_t* __restrict__ n_id,
                                                     int* __restrict__ local_id, int* __restrict__ first_pos,
                                                     int32_t* __restrict__ info, int64_t cap_nodes) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i == 0) { info[0] = counts[H]; info[1] = counts[2 * H + 1]; }
  if (i >= cap_nodes || i >= counts[H]) return;
  const int64_t u = n_id[i];
  local_id[u] = -1;
  first_pos[u] = 0x7fffffff;
}

// rows of `row_bytes` bytes: out[i] = in[idx[i]]; U = the widest unit every address is a multiple of
template <typename U>
__global__ void __launch_bounds__(kThreads) gather_rows_kernel(const char* __restrict__ in, int64_t ld_in_bytes,
                                                               const int64_t* __restrict__ idx, int64_t n,
                                                               int64_t row_bytes, char* __restrict__ out,
                                                               int64_t ld_out_bytes) {
  const int units = (int)(row_bytes / sizeof(U));
  const int lanes = units < 32 ? (units <= 1 ? 1 : (units <= 2 ? 2 : (units <= 4 ? 4 : (units <= 8 ? 8 : (units <= 16 ? 16 : 32))))) : 32;
  const int rows_per_block = kThreads / lanes;
  const int lane = threadIdx.x % lanes;
  const int64_t r = (int64_t)blockIdx.x * rows_per_block + threadIdx.x / lanes;
  if (r >= n) return;
  const U* src = reinterpret_cast<const U*>(in + idx[r] * ld_in_bytes);
  U* dst = reinterpret_cast<U*>(out + r * ld_out_bytes);
  for (int u = lane; u < units; u += lanes) dst[u] = __ldg(src + u);
}

}  // namespace
}  // namespace egnn

using namespace egnn;

extern "C" int egnn_neighbor_sample_caps(int64_t n_nodes, int64_t n_edges, int64_t batch, const int32_t* fanouts, int n_hops,
                                         int64_t* cap_nodes, int64_t* cap_edges) {
  const char* fn = "egnn_neighbor_sample_caps";
  EGNN_REQUIRE(n_nodes > 0 && n_edges >= 0 && batch > 0 && batch <= n_nodes, fn, "bad sizes");
  EGNN_REQUIRE(fanouts && n_hops >= 1 && n_hops <= kMaxHops && cap_nodes && cap_edges, fn, "bad fan-out list");
  const Caps c = make_caps(n_nodes, n_edges, batch, fanouts, n_hops);
  *cap_nodes = c.nodes;
  *cap_edges = c.edges;
  return 0;
}

extern "C" size_t egnn_neighbor_sample_workspace_bytes(int64_t n_nodes, int64_t n_edges, int64_t batch,
                                                       const int32_t* fanouts, int n_hops) {
  if (!fanouts || n_hops < 1 || n_hops > kMaxHops || batch < 1 || n_nodes < 1) return 0;
  return carve_sample(nullptr, make_caps(n_nodes, n_edges, batch, fanouts, n_hops)).bytes + 256;
}

extern "C" int egnn_neighbor_sample_state_init(int32_t* state, int64_t n_nodes, void* stream) {
  const char* fn = "egnn_neighbor_sample_state_init";
  EGNN_REQUIRE(state && n_nodes > 0, fn, "bad arguments");
  ns_state_init<<<(unsigned)ceil_div(n_nodes, kThreads), kThreads, 0, (cudaStream_t)stream>>>(state, state + n_nodes, n_nodes);
  EGNN_LAUNCH_CHECK(fn);
  return 0;
}

extern "C" int egnn_neighbor_sample(const int32_t* csr_ptr, const int32_t* csr_src, const int32_t* csr_eid, int64_t n_nodes,
                                    int64_t n_edges, const int64_t* seeds, int64_t batch, const int32_t* fanouts, int n_hops,
                                    uint64_t seed, uint64_t batch_idx, int32_t* state, int64_t* n_id, int64_t cap_nodes,
                                    int64_t* edge_index, int64_t* e_id, int64_t cap_edges, int32_t* counts, int32_t* info,
                                    void* workspace, size_t workspace_bytes, void* stream) {
  const char* fn = "egnn_neighbor_sample";
  EGNN_REQUIRE(n_nodes > 0 && n_edges >= 0 && batch > 0 && batch <= n_nodes, fn, "bad sizes");
  EGNN_REQUIRE(n_nodes < (int64_t)1073741824 && n_edges < (int64_t)2147483647, fn, "too large for int32 indices");
  EGNN_REQUIRE(fanouts && n_hops >= 1 && n_hops <= kMaxHops, fn, "1..8 hops");
  EGNN_REQUIRE(csr_ptr && seeds && state && n_id && edge_index && counts && info && workspace, fn, "null pointer");
  EGNN_REQUIRE(n_edges == 0 || (csr_src && csr_eid), fn, "null graph pointer");
  const Caps c = make_caps(n_nodes, n_edges, batch, fanouts, n_hops);
  EGNN_REQUIRE(cap_nodes >= c.nodes && cap_edges >= c.edges, fn, "output capacity below egnn_neighbor_sample_caps");
  EGNN_REQUIRE(workspace_bytes >= egnn_neighbor_sample_workspace_bytes(n_nodes, n_edges, batch, fanouts, n_hops), fn,
               "workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  char* base = reinterpret_cast<char*>(((uintptr_t)workspace + 255) & ~uintptr_t(255));
  SampleWs w = carve_sample(base, c);
  int* local_id = state;
  int* first_pos = state + n_nodes;
  const int H = n_hops;
  ns_seed<<<(unsigned)ceil_div(batch, kThreads), kThreads, 0, st>>>(seeds, batch, local_id, n_id, counts, H);
  EGNN_LAUNCH_CHECK(fn);
  for (int h = 0; h < H; ++h) {
    const int k = fanouts[h];
    if (c.F[h] == 0 || c.E[h] == 0) {   // an edgeless graph: nothing to pick, the counters just carry over
      cudaMemsetAsync(w.totals, 0, 2 * sizeof(int), st);
      ns_assign<<<1, kThreads, 0, st>>>(counts, H, h, w.totals, w.src_g, w.flag, w.rank, local_id, n_id, 0);
      EGNN_LAUNCH_CHECK(fn);
      continue;
    }
    const unsigned gF = (unsigned)ceil_div(c.F[h], kThreads);
    ns_count<<<gF, kThreads, 0, st>>>(csr_ptr, n_id, counts, h, k, c.F[h], w.cnt);
    EGNN_LAUNCH_CHECK(fn);
    int rc = exclusive_scan(w.cnt, w.cnt, c.F[h], w.tile_sums, w.totals, st);
    if (rc) return rc;
    ns_pick<<<gF, kThreads, 0, st>>>(csr_ptr, csr_src, csr_eid, n_id, counts, H, h, k, w.cnt, w.flag, seed,
                                     (uint32_t)batch_idx, local_id, first_pos, w.src_g, w.dst_l, e_id, c.F[h]);
    EGNN_LAUNCH_CHECK(fn);
    const unsigned gE = (unsigned)ceil_div(c.E[h], kThreads);
    ns_flag<<<gE, kThreads, 0, st>>>(counts, H, h, w.totals, w.src_g, local_id, first_pos, w.flag, c.E[h]);
    EGNN_LAUNCH_CHECK(fn);
    rc = exclusive_scan(w.flag, w.rank, c.E[h], w.tile_sums, w.totals + 1, st);
    if (rc) return rc;
    ns_assign<<<gE, kThreads, 0, st>>>(counts, H, h, w.totals, w.src_g, w.flag, w.rank, local_id, n_id, c.E[h]);
    EGNN_LAUNCH_CHECK(fn);
  }
  ns_finish<<<(unsigned)ceil_div(cap_edges, kThreads), kThreads, 0, st>>>(counts, H, w.src_g, w.dst_l, local_id, edge_index,
                                                                         cap_edges);
  EGNN_LAUNCH_CHECK(fn);
  ns_reset<<<(unsigned)ceil_div(cap_nodes, kThreads), kThreads, 0, st>>>(counts, H, n_id, local_id, first_pos, info, cap_nodes);
  EGNN_LAUNCH_CHECK(fn);
  return 0;
}

extern "C" int egnn_gather_rows(const void* in, int64_t ld_in_bytes, const int64_t* idx, int64_t n, int64_t row_bytes,
                                void* out, int64_t ld_out_bytes, void* stream) {
  const char* fn = "egnn_gather_rows";
  EGNN_REQUIRE(n >= 0 && row_bytes > 0 && ld_in_bytes >= row_bytes && ld_out_bytes >= row_bytes, fn, "bad sizes");
  if (n == 0) return 0;
  EGNN_REQUIRE(in && idx && out, fn, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const uintptr_t a = (uintptr_t)in | (uintptr_t)out | (uintptr_t)ld_in_bytes | (uintptr_t)ld_out_bytes | (uintptr_t)row_bytes;
  auto launch = [&](auto unit) {
    using U = decltype(unit);
    const int64_t units = row_bytes / (int64_t)sizeof(U);
    const int lanes = units < 32 ? (units <= 1 ? 1 : (units <= 2 ? 2 : (units <= 4 ? 4 : (units <= 8 ? 8 : (units <= 16 ? 16 : 32))))) : 32;
    const int rows_per_block = kThreads / lanes;
    gather_rows_kernel<U><<<(unsigned)ceil_div(n, rows_per_block), kThreads, 0, st>>>(
        (const char*)in, ld_in_bytes, idx, n, row_bytes, (char*)out, ld_out_bytes);
  };
  if (a % 16 == 0) launch(uint4{});
  else if (a % 8 == 0) launch(uint2{});
  else if (a % 4 == 0) launch((uint32_t)0);
  else launch((uint8_t)0);
  EGNN_LAUNCH_CHECK(fn);
  return 0;
}
