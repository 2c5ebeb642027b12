// K2/K3, production kernel -- lean sub-warp gather-reduce (sum / mean / weighted).
//
// What the profiles said (profiles/r01_spmm.md): on the Elliptic degree profile (mean in-degree
// 2.3, 45 % of rows have one edge) neither HBM nor L2 limits the aggregation -- the first three
// kernels (sub-warp with a generic epilogue, register-tiled, TMA/LDGSTS shared-memory staged;
// see profiles/experiments/) all sat at 45-68 % issue-slot utilisation with DRAM at 7-21 %:
// ~110 warp instructions per row, of which ~7 were the useful loads and adds.  This kernel is
// built to minimise instructions per row and to run at full occupancy so that thread-level
// parallelism hides the rowptr -> col -> feature-row latency chain:
//   * no shared memory, no barriers on the main path, <= 40 registers -> 6-8 CTAs per SM;
//   * a lane group of G lanes owns a row, 16-byte loads (8 bf16 / 4 fp32 features per lane);
//   * edges are consumed in predicated batches (all column indices, then all feature loads,
//     then the ordered fp32 adds): three dependent latencies per batch whatever the degree;
//   * exact mean: x/deg for deg = 2^k is x * 2^-k bit for bit (80 % of rows); only the other
//     rows pay the IEEE division; lean store when there is no bias / activation / accumulate.
// Summation order per row = stored edge order, sequential fp32 adds (no FMA contraction with the
// edge weight), so fp32 output stays bitwise equal to the CPU scatter_add_ oracle.
// Rows longer than kLongRow go to the whole-CTA path of spmm.cuh (first kLongCtas blocks).
#include <stdlib.h>

#include "spmm.cuh"

namespace egnn {
namespace {
using namespace spmm_detail;

template <int VEC>
struct Acc {
  float v[VEC];
};
template <typename TI, int VEC>
__device__ __forceinline__ Acc<VEC> ldg_vec(const TI* p);
template <>
__device__ __forceinline__ Acc<4> ldg_vec<float, 4>(const float* p) {
  float4 t = __ldg(reinterpret_cast<const float4*>(p));
  return Acc<4>{{t.x, t.y, t.z, t.w}};
}
template <>
__device__ __forceinline__ Acc<4> ldg_vec<__nv_bfloat16, 4>(const __nv_bfloat16* p) {
  uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
  return Acc<4>{{__uint_as_float(v.x << 16), __uint_as_float(v.x & 0xffff0000u), __uint_as_float(v.y << 16),
                 __uint_as_float(v.y & 0xffff0000u)}};
}
template <>
__device__ __forceinline__ Acc<8> ldg_vec<__nv_bfloat16, 8>(const __nv_bfloat16* p) {
  uint4 q = __ldg(reinterpret_cast<const uint4*>(p));
  return Acc<8>{{__uint_as_float(q.x << 16), __uint_as_float(q.x & 0xffff0000u), __uint_as_float(q.y << 16),
                 __uint_as_float(q.y & 0xffff0000u), __uint_as_float(q.z << 16), __uint_as_float(q.z & 0xffff0000u),
                 __uint_as_float(q.w << 16), __uint_as_float(q.w & 0xffff0000u)}};
}
template <int VEC>
__device__ __forceinline__ void stv(float* p, const Acc<VEC>& a) {
#pragma unroll
  for (int h = 0; h < VEC / 4; ++h)
    *reinterpret_cast<float4*>(p + 4 * h) = make_float4(a.v[4 * h], a.v[4 * h + 1], a.v[4 * h + 2], a.v[4 * h + 3]);
}
__device__ __forceinline__ void stv(__nv_bfloat16* p, const Acc<4>& a) {
  *reinterpret_cast<uint2*>(p) = make_uint2(pack_bf16x2(a.v[0], a.v[1]), pack_bf16x2(a.v[2], a.v[3]));
}
__device__ __forceinline__ void stv(__nv_bfloat16* p, const Acc<8>& a) {
  *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16x2(a.v[0], a.v[1]), pack_bf16x2(a.v[2], a.v[3]),
                                            pack_bf16x2(a.v[4], a.v[5]), pack_bf16x2(a.v[6], a.v[7]));
}

template <bool EXACT, int VEC>
__device__ __forceinline__ void mean_scale(Acc<VEC>& a, int deg) {
  if (deg <= 1) return;
  const float c = (float)deg;
  if (!EXACT || (deg & (deg - 1)) == 0) {
    // exact for deg = 2^k; for a bf16 result the 0.5-ulp(fp32) error of x * rn(1/deg) vanishes in
    // the final rounding to 8 mantissa bits, so the division is only paid for fp32 output
    const float inv = __fdiv_rn(1.0f, c);  // exact power of two
#pragma unroll
    for (int i = 0; i < VEC; ++i) a.v[i] = __fmul_rn(a.v[i], inv);
  } else {
#pragma unroll
    for (int i = 0; i < VEC; ++i) a.v[i] = __fdiv_rn(a.v[i], c);
  }
}

template <typename TO, int VEC>
__device__ __forceinline__ void generic_epilogue(const Params& P, TO* o, const TO* add, int f, Acc<VEC>& a) {
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    float v = a.v[i];
    if (P.bias) v = __fadd_rn(v, __ldg(P.bias + f + i));
    v = apply_act(v, P.act);
    if (P.accumulate) v = __fadd_rn(to_f32(add[i]), v);
    a.v[i] = v;
  }
  stv(o, a);
}

// minimum resident CTAs per SM: the kernel is bound by rows in flight (ncu: issue active 31 %, DRAM 28 %,
// 52 registers -> 4 CTAs), so the register budget is capped to buy occupancy
constexpr int lean_min_blocks(int vpl, int batch, int vec) { return (vpl >= 3 && batch == 1 && vec == 4) ? 6 : 1; }

template <typename TI, typename TO, int MODE, int VEC, int G, int VPL, int BATCH, bool LEAN>
__global__ void __launch_bounds__(kThreads, lean_min_blocks(VPL, BATCH, VEC)) spmm_lean(Params P) {
  constexpr int kRows = kThreads / G;
  __shared__ __align__(16) float s_stage[kStageEdges][kSliceFeat];  // used by the long-row CTAs only
  __shared__ float s_scale[kStageEdges];
  __shared__ int s_col[kStageEdges];
  const bool has_long = P.long_rows != nullptr;
  int bx = blockIdx.x;
  if (has_long) {
    if (bx < kLongCtas) {
      long_row_path<TI, TO, MODE>(P, bx, s_stage, s_scale, s_col);
      return;
    }
    bx -= kLongCtas;
  }
  const int g = threadIdx.x / G, lane = threadIdx.x % G;
  const int64_t slot = (int64_t)bx * kRows + g;
  if (slot >= P.n_rows) return;
  // longest-rows-first schedule: blocks are dispatched in index order, so the medium-degree rows
  // start at t = 0 and overlap with the bulk of 1-2 edge rows instead of forming the tail; rows of
  // one warp also have similar degrees (no intra-warp imbalance)
  const int64_t row = P.row_order ? (int64_t)__ldg(P.row_order + slot) : slot;
  const int p0 = __ldg(P.ptr + row), p1 = __ldg(P.ptr + row + 1);
  const int deg = p1 - p0;
  if (has_long && deg > kLongRow) return;

  const TI* __restrict__ in = reinterpret_cast<const TI*>(P.in) + VEC * lane;
  bool on[VPL];
#pragma unroll
  for (int k = 0; k < VPL; ++k) on[k] = VEC * (lane + k * G) < P.n_feat;
  Acc<VEC> acc[VPL];
#pragma unroll
  for (int k = 0; k < VPL; ++k)
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[k].v[i] = 0.f;

  for (int p = p0; p < p1; p += BATCH) {
    int c[BATCH];
    float w[BATCH];
#pragma unroll
    for (int b = 0; b < BATCH; ++b) {
      const bool ok = p + b < p1;
      c[b] = ok ? __ldg(P.col + p + b) : -1;
      if (MODE == M_WEIGHTED) w[b] = ok ? __ldg(P.w + p + b) : 0.f;
    }
    Acc<VEC> x[BATCH][VPL];
#pragma unroll
    for (int b = 0; b < BATCH; ++b) {
      if (c[b] >= 0) {
        const TI* src = in + (int64_t)c[b] * P.ld_in;
#pragma unroll
        for (int k = 0; k < VPL; ++k)
          if (k < VPL - 1 || on[k]) x[b][k] = ldg_vec<TI, VEC>(src + k * G * VEC);
      }
    }
#pragma unroll
    for (int b = 0; b < BATCH; ++b) {
      if (c[b] >= 0) {
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
          if (k < VPL - 1 || on[k]) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
              float t = x[b][k].v[i];
              if (MODE == M_WEIGHTED) t = __fmul_rn(w[b], t);
              acc[k].v[i] = __fadd_rn(acc[k].v[i], t);
            }
          }
        }
      }
    }
  }
  TO* o = reinterpret_cast<TO*>(P.out) + row * P.ld_out + VEC * lane;
  const TO* add = reinterpret_cast<const TO*>(P.add_in) + row * P.ld_add + VEC * lane;
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    if (k < VPL - 1 || on[k]) {
      if (P.mean) mean_scale<sizeof(TO) == 4, VEC>(acc[k], deg);
      if (LEAN) stv(o + k * G * VEC, acc[k]);
      else generic_epilogue<TO, VEC>(P, o + k * G * VEC, add + k * G * VEC, VEC * (lane + k * G), acc[k]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Pipelined persistent variant (the production forward / transposed kernel).
//
// The lean kernel above still pays three dependent global latencies per row (row pointers -> column
// indices -> feature rows) inside every lane group, and with mean degree 2.3 only the last of them
// carries feature bytes.  Here a CTA walks tiles of kTileRows consecutive rows and keeps a three-deep
// software pipeline in shared memory with cp.async (LDGSTS): while the lane groups gather tile i, the
// row pointers of tile i+2 and the column indices (and edge weights) of tile i+1 are already in
// flight.  The gather phase therefore starts from shared-memory indices and every resident warp
// spends its time with feature loads outstanding.  Tiles are handed out by a global counter (two
// tiles of look-ahead), rows inside a tile by a shared-memory counter, so degree skew does not idle
// lane groups.  Per-row arithmetic is unchanged: sequential fp32 adds in stored edge order.
constexpr int kTileRows = 32;
constexpr int kColCap = 768;      // staged column indices per tile; longer tiles read the rest from global
constexpr int kCounterSlots = 16;

__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
               "l"(gmem_src)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

struct PipeSmem {
  int ptr[3][kTileRows + 1];
  int col[2][kColCap];
  int tile[4];
  int next_row[2];
};

template <typename TI, typename TO, int MODE, int VEC, int G, int VPL, int BATCH, bool LEAN>
__global__ void __launch_bounds__(kThreads) spmm_pipe(Params P, int* __restrict__ counters, int n_tiles) {
  __shared__ PipeSmem S;
  __shared__ float s_w[MODE == M_WEIGHTED ? 2 : 1][MODE == M_WEIGHTED ? kColCap : 1];
  __shared__ __align__(16) float s_stage[kStageEdges][kSliceFeat];  // long-row path only
  __shared__ float s_scale[kStageEdges];
  __shared__ int s_lcol[kStageEdges];
  const int tid = threadIdx.x;
  const bool has_long = P.long_rows != nullptr;
  if (has_long && blockIdx.x < kLongCtas) {
    long_row_path<TI, TO, MODE>(P, blockIdx.x, s_stage, s_scale, s_lcol);
    __syncthreads();
  }
  int* work = counters;        // next tile to hand out
  int* done = counters + 1;    // CTAs that have finished (the last one resets the slot)

  auto issue_ptr = [&](int t, int buf) {
    const int64_t r0 = (int64_t)t * kTileRows;
    for (int i = tid; i <= kTileRows; i += kThreads) {
      int64_t r = r0 + i;
      cp_async4(&S.ptr[buf][i], P.ptr + (r <= P.n_rows ? r : P.n_rows));
    }
  };
  auto issue_col = [&](int pbuf, int cbuf) {
    const int base = S.ptr[pbuf][0];
    int n = S.ptr[pbuf][kTileRows] - base;
    n = n < kColCap ? n : kColCap;
    for (int i = tid; i < n; i += kThreads) {
      cp_async4(&S.col[cbuf][i], P.col + base + i);
      if (MODE == M_WEIGHTED) cp_async4(&s_w[MODE == M_WEIGHTED ? cbuf : 0][i], P.w + base + i);
    }
  };

  // ---- prologue: three tile ids, row pointers of tiles 0 and 1, column indices of tile 0
  if (tid == 0) {
    S.tile[0] = atomicAdd(work, 1);
    S.tile[1] = atomicAdd(work, 1);
    S.tile[2] = atomicAdd(work, 1);
    S.next_row[0] = 0;
    S.next_row[1] = 0;
  }
  __syncthreads();
  if (S.tile[0] < n_tiles) issue_ptr(S.tile[0], 0);
  if (S.tile[1] < n_tiles) issue_ptr(S.tile[1], 1);
  cp_async_commit_wait_all();
  __syncthreads();
  if (S.tile[0] < n_tiles) issue_col(0, 0);
  cp_async_commit_wait_all();
  __syncthreads();

  const int g = tid / G, lane = tid % G;
  const unsigned gmask = G == 32 ? 0xffffffffu : (((1u << G) - 1u) << (((tid & 31) / G) * G));
  const TI* __restrict__ in = reinterpret_cast<const TI*>(P.in) + VEC * lane;
  bool on[VPL];
#pragma unroll
  for (int k = 0; k < VPL; ++k) on[k] = VEC * (lane + k * G) < P.n_feat;

  for (int it = 0;; ++it) {
    const int t = S.tile[it & 3];
    if (t >= n_tiles) break;  // block-uniform
    const int pb = it % 3, cb = it & 1;
    const int t1 = S.tile[(it + 1) & 3], t2 = S.tile[(it + 2) & 3];
    if (t2 < n_tiles) issue_ptr(t2, (it + 2) % 3);
    if (t1 < n_tiles) issue_col((it + 1) % 3, cb ^ 1);
    asm volatile("cp.async.commit_group;" ::: "memory");
    if (tid == 0) {
      S.tile[(it + 3) & 3] = atomicAdd(work, 1);
      S.next_row[cb ^ 1] = 0;
    }
    const int64_t r0 = (int64_t)t * kTileRows;
    const int rows_here = (int)min((int64_t)kTileRows, P.n_rows - r0);
    const int base = S.ptr[pb][0];
    for (;;) {
      int rr = 0;
      if (lane == 0) rr = atomicAdd(&S.next_row[cb], 1);
      rr = __shfl_sync(gmask, rr, 0, G);
      if (rr >= rows_here) break;
      const int p0 = S.ptr[pb][rr], p1 = S.ptr[pb][rr + 1];
      const int deg = p1 - p0;
      if (has_long && deg > kLongRow) continue;
      const int64_t row = r0 + rr;
      Acc<VEC> acc[VPL];
#pragma unroll
      for (int k = 0; k < VPL; ++k)
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[k].v[i] = 0.f;
      for (int p = p0; p < p1; p += BATCH) {
        int c[BATCH];
        float w[BATCH];
#pragma unroll
        for (int b = 0; b < BATCH; ++b) {
          const int q = p + b, off = q - base;
          const bool ok = q < p1;
          c[b] = ok ? (off < kColCap ? S.col[cb][off] : __ldg(P.col + q)) : -1;
          if (MODE == M_WEIGHTED)
            w[b] = ok ? (off < kColCap ? s_w[MODE == M_WEIGHTED ? cb : 0][off] : __ldg(P.w + q)) : 0.f;
        }
        Acc<VEC> x[BATCH][VPL];
#pragma unroll
        for (int b = 0; b < BATCH; ++b) {
          if (c[b] >= 0) {
            const TI* src = in + (int64_t)c[b] * P.ld_in;
#pragma unroll
            for (int k = 0; k < VPL; ++k)
              if (k < VPL - 1 || on[k]) x[b][k] = ldg_vec<TI, VEC>(src + k * G * VEC);
          }
        }
#pragma unroll
        for (int b = 0; b < BATCH; ++b) {
          if (c[b] >= 0) {
#pragma unroll
            for (int k = 0; k < VPL; ++k) {
              if (k < VPL - 1 || on[k]) {
#pragma unroll
                for (int i = 0; i < VEC; ++i) {
                  float tv = x[b][k].v[i];
                  if (MODE == M_WEIGHTED) tv = __fmul_rn(w[b], tv);
                  acc[k].v[i] = __fadd_rn(acc[k].v[i], tv);
                }
              }
            }
          }
        }
      }
      TO* o = reinterpret_cast<TO*>(P.out) + row * P.ld_out + VEC * lane;
      const TO* add = reinterpret_cast<const TO*>(P.add_in) + row * P.ld_add + VEC * lane;
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        if (k < VPL - 1 || on[k]) {
          if (P.mean) mean_scale<sizeof(TO) == 4, VEC>(acc[k], deg);
          if (LEAN) stv(o + k * G * VEC, acc[k]);
          else generic_epilogue<TO, VEC>(P, o + k * G * VEC, add + k * G * VEC, VEC * (lane + k * G), acc[k]);
        }
      }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
  }
  // the last CTA to leave resets the counter slot for the next launch that uses it
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(done, 1) == (int)gridDim.x - 1) {
      *work = 0;
      *done = 0;
      __threadfence();
    }
  }
}

int* pipe_counters() {  // kCounterSlots x {work, done}, zero-initialised once per device
  static int* buf[64] = {nullptr};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return nullptr;
  if (!buf[dev]) {
    if (cudaMalloc(&buf[dev], kCounterSlots * 2 * sizeof(int)) != cudaSuccess) return nullptr;
    cudaMemset(buf[dev], 0, kCounterSlots * 2 * sizeof(int));
  }
  return buf[dev];
}

template <typename TI, typename TO, int MODE, int VEC, int G, int VPL, int BATCH>
int launch_pipe(const Params& P, cudaStream_t st) {
  static std::atomic<unsigned> slot{0};
  int* ctr = pipe_counters();
  if (!ctr) return -2;
  ctr += 2 * (slot.fetch_add(1) % kCounterSlots);
  const int n_tiles = (int)ceil_div(P.n_rows, kTileRows);
  const bool lean = !P.bias && P.act == EGNN_ACT_NONE && !P.accumulate;
  int per_sm = 0;
  if (lean) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, spmm_pipe<TI, TO, MODE, VEC, G, VPL, BATCH, true>, kThreads, 0);
  else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, spmm_pipe<TI, TO, MODE, VEC, G, VPL, BATCH, false>, kThreads, 0);
  if (per_sm < 1) per_sm = 1;
  int grid = kNumSMs * per_sm;
  const int min_grid = P.long_rows ? kLongCtas : 1;
  if (grid > n_tiles) grid = n_tiles > min_grid ? n_tiles : min_grid;
  if (lean) spmm_pipe<TI, TO, MODE, VEC, G, VPL, BATCH, true><<<grid, kThreads, 0, st>>>(P, ctr, n_tiles);
  else spmm_pipe<TI, TO, MODE, VEC, G, VPL, BATCH, false><<<grid, kThreads, 0, st>>>(P, ctr, n_tiles);
  EGNN_LAUNCH_CHECK("egnn_spmm(pipe)");
  return 0;
}

template <typename TI, typename TO, int MODE, int VEC, int G, int VPL, int BATCH>
int launch_exp(const Params& P, cudaStream_t st) {
  constexpr int rows = kThreads / G;
  dim3 grid((unsigned)(ceil_div(P.n_rows, rows) + (P.long_rows ? kLongCtas : 0)), 1);
  spmm_lean<TI, TO, MODE, VEC, G, VPL, BATCH, true><<<grid, kThreads, 0, st>>>(P);
  EGNN_LAUNCH_CHECK("egnn_spmm(lean exp)");
  return 0;
}

template <typename TI, typename TO, int MODE, int VEC, int G, int VPL>
int launch_cfg(const Params& P, cudaStream_t st) {
  constexpr int BATCH = VPL >= 3 ? 1 : 2;  // wide rows: one edge per batch keeps the kernel at 6 CTAs per SM
  static const bool use_pipe = [] { const char* e = getenv("EGNN_SPMM_IMPL"); return e && e[0] == 'p'; }();
  if (use_pipe) {
    int rc = launch_pipe<TI, TO, MODE, VEC, G, VPL, BATCH>(P, st);
    if (rc != -2) return rc;
  }
  constexpr int rows = kThreads / G;
  const int long_ctas = P.long_rows ? kLongCtas : 0;
  dim3 grid((unsigned)(ceil_div(P.n_rows, rows) + long_ctas), 1);
  const bool lean = !P.bias && P.act == EGNN_ACT_NONE && !P.accumulate;
  if (lean) spmm_lean<TI, TO, MODE, VEC, G, VPL, BATCH, true><<<grid, kThreads, 0, st>>>(P);
  else spmm_lean<TI, TO, MODE, VEC, G, VPL, BATCH, false><<<grid, kThreads, 0, st>>>(P);
  EGNN_LAUNCH_CHECK("egnn_spmm(lean)");
  return 0;
}

template <typename TI, typename TO, int MODE, int VEC>
int launch(const Params& P, cudaStream_t st) {
  const int nvec = P.n_feat / VEC;
#ifdef EGNN_SPMM_EXPERIMENT
  // tuning hook (profiles/spmm_probe.py): EGNN_SPMM_CFG = "G,VPL,BATCH"
  if (const char* e = getenv("EGNN_SPMM_CFG")) {
    int G = 0, V = 0, B = 0;
    sscanf(e, "%d,%d,%d", &G, &V, &B);
    if (MODE == M_PLAIN && !P.bias && !P.act && !P.accumulate && G * V >= nvec) {
#define EXP(g, v, b) if (G == g && V == v && B == b) return launch_exp<TI, TO, MODE, VEC, g, v, b>(P, st);
      EXP(2, 4, 1) EXP(2, 4, 2) EXP(4, 2, 1) EXP(4, 2, 2) EXP(4, 2, 4) EXP(8, 1, 1) EXP(8, 1, 2) EXP(8, 1, 4)
      EXP(8, 6, 1) EXP(8, 6, 2) EXP(16, 3, 1) EXP(16, 3, 2) EXP(16, 3, 4) EXP(32, 2, 1) EXP(32, 2, 2) EXP(4, 11, 1)
#undef EXP
    }
  }
#endif
  if (nvec <= 4) return launch_cfg<TI, TO, MODE, VEC, 4, 1>(P, st);
  if (nvec <= 8) return launch_cfg<TI, TO, MODE, VEC, 8, 1>(P, st);
  if (nvec <= 16) return launch_cfg<TI, TO, MODE, VEC, 16, 1>(P, st);
  if (nvec <= 24) return launch_cfg<TI, TO, MODE, VEC, 8, 3>(P, st);
  if (nvec <= 32) return launch_cfg<TI, TO, MODE, VEC, 32, 1>(P, st);
  if (nvec <= 48) return launch_cfg<TI, TO, MODE, VEC, 16, 3>(P, st);
  if (nvec <= 64) return launch_cfg<TI, TO, MODE, VEC, 32, 2>(P, st);
  return -2;  // wider rows: the chunked sub-warp kernel of spmm.cu
}

template <int MODE>
int by_dtype(const Params& P, int in_dt, int out_dt, cudaStream_t st) {
  const bool v8 = in_dt == EGNN_BF16 && P.n_feat % 8 == 0 && P.ld_in % 8 == 0 && P.ld_out % 8 == 0 &&
                  ((uintptr_t)P.in % 16 == 0) && ((uintptr_t)P.out % 16 == 0) &&
                  (!P.accumulate || (P.ld_add % 8 == 0 && (uintptr_t)P.add_in % 16 == 0));
  if (in_dt == EGNN_F32 && out_dt == EGNN_F32) return launch<float, float, MODE, 4>(P, st);
  if (in_dt == EGNN_F32 && out_dt == EGNN_BF16) return launch<float, __nv_bfloat16, MODE, 4>(P, st);
  if (in_dt == EGNN_BF16 && out_dt == EGNN_BF16)
    return v8 ? launch<__nv_bfloat16, __nv_bfloat16, MODE, 8>(P, st)
              : launch<__nv_bfloat16, __nv_bfloat16, MODE, 4>(P, st);
  if (in_dt == EGNN_BF16 && out_dt == EGNN_F32)
    return v8 ? launch<__nv_bfloat16, float, MODE, 8>(P, st) : launch<__nv_bfloat16, float, MODE, 4>(P, st);
  return -2;
}

}  // namespace

int spmm_tile_launch(const spmm_detail::Params& P, int in_dt, int out_dt, bool weighted, cudaStream_t st) {
  // production: the streaming lane-group kernel (spmm_stream.cu); EGNN_SPMM_IMPL=lean / pipe select the
  // per-row kernels of this file (kept for 4-lane groups and as the measured baseline)
#ifdef EGNN_SPMM_EXPERIMENT
  const char* impl_env = getenv("EGNN_SPMM_IMPL");  // re-read on every launch so one probe process can sweep it
  const bool use_stream = !impl_env || impl_env[0] == 's';
#else
  static const bool use_stream = [] { const char* e = getenv("EGNN_SPMM_IMPL"); return !e || e[0] == 's'; }();
#endif
  if (use_stream) {
    int rc = spmm_stream_launch(P, in_dt, out_dt, weighted, st);
    if (rc != -2) return rc;
  }
  return weighted ? by_dtype<M_WEIGHTED>(P, in_dt, out_dt, st) : by_dtype<M_PLAIN>(P, in_dt, out_dt, st);
}

}  // namespace egnn
