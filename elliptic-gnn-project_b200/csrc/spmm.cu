// K2/K3 -- deterministic segmented gather-reduce (CSR SpMM forward, transposed-CSR backward).
//
// Replaces PyG's un-fused propagate (index_select [E,F] -> scatter_add_ -> count ->
// clamp -> div) and its autograd backward.  HBM-bound: every source
// row is read with 128-bit loads, nothing of size [E,F] is ever materialised.
//
// Mapping: one sub-warp group of G lanes per destination row, each lane owning VPL
// 4-feature vectors; a row's edges are visited in stored (= original edge) order and added
// one after another in fp32 (`__fadd_rn`, never FMA-contracted with the edge weight), so
// fp32 output is bitwise what CPU `scatter_add_` produces.  Row pointers of
// the CTA's row tile are staged in shared memory.  Rows longer than kLongRow are skipped by
// the sub-warp path and handled by whole CTAs (the first kLongCtas blocks of the grid): the
// CTA stages 128 source rows x 32 features at a time in shared memory with all threads
// loading (memory-level parallelism), then each thread adds its feature over the staged rows
// sequentially -- same summation order, no float atomics anywhere.
#include "spmm.cuh"

namespace egnn {
namespace {

using namespace spmm_detail;

// ---------------------------------------------------------------------------------------
// Vectorised kernel.  blockIdx.y selects a feature chunk of G*VPL*4 features.
// ---------------------------------------------------------------------------------------
template <typename TI, typename TO, int MODE, int G, int VPL>
__global__ void __launch_bounds__(kThreads) spmm_vec(Params P) {
  constexpr int kRows = kThreads / G;
  constexpr int kChunkFeat = G * VPL * 4;
  static_assert(kChunkFeat <= kStageFeat, "chunk too wide for the staging buffer");
  __shared__ int s_ptr[kRows + 1];
  __shared__ __align__(16) float s_stage[kStageEdges][kSliceFeat];
  __shared__ float s_scale[kStageEdges];
  __shared__ int s_col[kStageEdges];

  const TI* __restrict__ in = reinterpret_cast<const TI*>(P.in);
  const int f_chunk0 = blockIdx.y * kChunkFeat;
  const bool has_long = P.long_rows != nullptr;
  int bx = blockIdx.x;

  if (has_long) {
    if (bx < kLongCtas) {
      long_row_path<TI, TO, MODE>(P, bx, s_stage, s_scale, s_col);
      return;
    }
    bx -= kLongCtas;
  }

  // ---------------- sub-warp path ----------------
  const int64_t row0 = (int64_t)bx * kRows;
  for (int i = threadIdx.x; i <= kRows; i += kThreads) {
    int64_t r = row0 + i;
    s_ptr[i] = P.ptr[r <= P.n_rows ? r : P.n_rows];
  }
  __syncthreads();
  const int g = threadIdx.x / G, lane = threadIdx.x % G;
  const int64_t row = row0 + g;
  if (row >= P.n_rows) return;
  const int p0 = s_ptr[g], p1 = s_ptr[g + 1];
  const int deg = p1 - p0;
  if (has_long && deg > kLongRow) return;

  bool act_v[VPL];
  int f_v[VPL];
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    f_v[k] = f_chunk0 + 4 * (lane + k * G);
    act_v[k] = f_v[k] < P.n_feat;
  }
  F4 acc[VPL];
#pragma unroll
  for (int k = 0; k < VPL; ++k) acc[k] = F4{0.f, 0.f, 0.f, 0.f};

  int p = p0;
  for (; p + 1 < p1; p += 2) {  // two edges in flight, added in order
    const int c0 = __ldg(P.col + p), c1 = __ldg(P.col + p + 1);
    const float s0 = edge_scale<MODE>(P, p, c0), s1 = edge_scale<MODE>(P, p + 1, c1);
    F4 v0[VPL], v1[VPL];
#pragma unroll
    for (int k = 0; k < VPL; ++k)
      if (act_v[k]) {
        v0[k] = ld4(in + (int64_t)c0 * P.ld_in + f_v[k]);
        v1[k] = ld4(in + (int64_t)c1 * P.ld_in + f_v[k]);
      }
#pragma unroll
    for (int k = 0; k < VPL; ++k)
      if (act_v[k]) {
        acc_add(acc[k], edge_term<MODE>(v0[k], s0));
        acc_add(acc[k], edge_term<MODE>(v1[k], s1));
      }
  }
  if (p < p1) {
    const int c0 = __ldg(P.col + p);
    const float s0 = edge_scale<MODE>(P, p, c0);
#pragma unroll
    for (int k = 0; k < VPL; ++k)
      if (act_v[k]) acc_add(acc[k], edge_term<MODE>(ld4(in + (int64_t)c0 * P.ld_in + f_v[k]), s0));
  }
#pragma unroll
  for (int k = 0; k < VPL; ++k)
    if (act_v[k]) epilogue_store<TO>(P, row, f_v[k], acc[k], deg);
}

// ---------------------------------------------------------------------------------------
// Scalar fallback: any n_feat / leading dimension / alignment (e.g. F = 167, F = 2).
// G lanes per row, lane owns features lane, lane+G, ...
// ---------------------------------------------------------------------------------------
template <typename TI, typename TO, int MODE, int G>
__global__ void __launch_bounds__(kThreads) spmm_scalar(Params P) {
  constexpr int kRows = kThreads / G;
  __shared__ int s_ptr[kRows + 1];
  const TI* __restrict__ in = reinterpret_cast<const TI*>(P.in);
  TO* __restrict__ out = reinterpret_cast<TO*>(P.out);
  const int64_t row0 = (int64_t)blockIdx.x * kRows;
  for (int i = threadIdx.x; i <= kRows; i += kThreads) {
    int64_t r = row0 + i;
    s_ptr[i] = P.ptr[r <= P.n_rows ? r : P.n_rows];
  }
  __syncthreads();
  const int g = threadIdx.x / G, lane = threadIdx.x % G;
  const int64_t row = row0 + g;
  if (row >= P.n_rows) return;
  const int p0 = s_ptr[g], p1 = s_ptr[g + 1];
  const int deg = p1 - p0;
  for (int f0 = lane; f0 < P.n_feat; f0 += 4 * G) {  // 4 features per lane per sweep
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int p = p0; p < p1; ++p) {
      const int c = __ldg(P.col + p);
      const float s = edge_scale<MODE>(P, p, c);
      const TI* src = in + (int64_t)c * P.ld_in;
      float v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        int f = f0 + k * G;
        v[k] = f < P.n_feat ? to_f32(__ldg(src + f)) : 0.f;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float t = v[k];
        if (MODE == M_WEIGHTED) t = __fmul_rn(s, t);
        if (MODE == M_DIV_NBR) t = __fdiv_rn(t, s);
        acc[k] = __fadd_rn(acc[k], t);
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      int f = f0 + k * G;
      if (f >= P.n_feat) continue;
      float a = acc[k];
      if (P.mean) a = __fdiv_rn(a, (float)(deg > 1 ? deg : 1));
      if (P.bias) a = __fadd_rn(a, __ldg(P.bias + f));
      a = apply_act(a, P.act);
      TO* o = out + row * P.ld_out + f;
      if (P.accumulate) a = __fadd_rn(to_f32(reinterpret_cast<const TO*>(P.add_in)[row * P.ld_add + f]), a);
      *o = from_f32<TO>(a);
    }
  }
}

template <typename TI, typename TO, int MODE>
int launch_vec(const Params& P, cudaStream_t st) {
  const int nvec = P.n_feat / 4;
  const int long_ctas = P.long_rows ? kLongCtas : 0;
#define EGNN_SPMM_CFG(G, VPL)                                                           \
  {                                                                                      \
    constexpr int rows = kThreads / (G);                                                 \
    dim3 grid((unsigned)(ceil_div(P.n_rows, rows) + long_ctas),                          \
              (unsigned)ceil_div(nvec, (G) * (VPL)));                                    \
    spmm_vec<TI, TO, MODE, G, VPL><<<grid, kThreads, 0, st>>>(P);                        \
  }
  if (nvec <= 4) EGNN_SPMM_CFG(4, 1)
  else if (nvec <= 8) EGNN_SPMM_CFG(8, 1)
  else if (nvec <= 16) EGNN_SPMM_CFG(16, 1)
  else if (nvec <= 32) EGNN_SPMM_CFG(32, 1)
  else if (nvec <= 48) EGNN_SPMM_CFG(16, 3)
  else EGNN_SPMM_CFG(32, 2)
#undef EGNN_SPMM_CFG
  EGNN_LAUNCH_CHECK("egnn_spmm(vec)");
  return 0;
}

template <typename TI, typename TO, int MODE>
int launch_scalar(const Params& P, cudaStream_t st) {
#define EGNN_SPMM_SC(G)                                                      \
  {                                                                          \
    constexpr int rows = kThreads / (G);                                     \
    spmm_scalar<TI, TO, MODE, G><<<(unsigned)ceil_div(P.n_rows, rows), kThreads, 0, st>>>(P); \
  }
  if (P.n_feat <= 2) EGNN_SPMM_SC(2)
  else if (P.n_feat <= 8) EGNN_SPMM_SC(8)
  else EGNN_SPMM_SC(32)
#undef EGNN_SPMM_SC
  EGNN_LAUNCH_CHECK("egnn_spmm(scalar)");
  return 0;
}

template <typename TI, typename TO, int MODE>
int dispatch_layout(const Params& P, bool vec_ok, cudaStream_t st) {
  return vec_ok ? launch_vec<TI, TO, MODE>(P, st) : launch_scalar<TI, TO, MODE>(P, st);
}

template <int MODE>
int dispatch_dtype(const Params& P, int in_dt, int out_dt, bool vec_ok, cudaStream_t st) {
  if (in_dt == EGNN_F32 && out_dt == EGNN_F32) return dispatch_layout<float, float, MODE>(P, vec_ok, st);
  if (in_dt == EGNN_F32 && out_dt == EGNN_BF16)
    return dispatch_layout<float, __nv_bfloat16, MODE>(P, vec_ok, st);
  if (in_dt == EGNN_BF16 && out_dt == EGNN_BF16)
    return dispatch_layout<__nv_bfloat16, __nv_bfloat16, MODE>(P, vec_ok, st);
  if (in_dt == EGNN_BF16 && out_dt == EGNN_F32)
    return dispatch_layout<__nv_bfloat16, float, MODE>(P, vec_ok, st);
  return fail("egnn_spmm", "unsupported dtype combination");
}

}  // namespace
}  // namespace egnn

using namespace egnn;

extern "C" int egnn_spmm(int mode, const int32_t* ptr, const int32_t* col, const float* w,
                         const int32_t* nbr_ptr, const int32_t* long_rows, const int32_t* n_long,
                         const int32_t* row_order, const int32_t* part, int64_t n_tasks, const void* in, int in_dtype,
                         int64_t ld_in, void* out, int out_dtype,
                         int64_t ld_out, int64_t n_rows, int64_t n_feat, const float* bias, int act,
                         int accumulate, const void* addend, int64_t ld_addend, void* stream) {
  const char* fn = "egnn_spmm";
  EGNN_REQUIRE(ptr && col && in && out, fn, "null pointer");
  EGNN_REQUIRE(n_rows >= 0 && n_feat > 0 && n_feat <= (1 << 20), fn, "bad shape");
  EGNN_REQUIRE(ld_in >= n_feat && ld_out >= n_feat, fn, "leading dimension < n_feat");
  EGNN_REQUIRE(mode >= EGNN_SPMM_SUM && mode <= EGNN_SPMM_WEIGHTED, fn, "bad mode");
  EGNN_REQUIRE(mode != EGNN_SPMM_WEIGHTED || w, fn, "WEIGHTED needs w");
  EGNN_REQUIRE(mode != EGNN_SPMM_DIV_NBR || nbr_ptr, fn, "DIV_NBR needs nbr_ptr");
  EGNN_REQUIRE((long_rows == nullptr) == (n_long == nullptr), fn, "long_rows/n_long mismatch");
  if (n_rows == 0) return 0;
  Params P;
  P.ptr = ptr; P.col = col; P.w = w; P.nbr_ptr = nbr_ptr;
  P.long_rows = long_rows; P.n_long = n_long; P.row_order = row_order;
  P.part = part; P.n_tasks = part ? n_tasks : 0;
  P.in = in; P.out = out; P.bias = bias;
  P.ld_in = ld_in; P.ld_out = ld_out; P.n_rows = n_rows; P.n_feat = (int)n_feat;
  P.mean = (mode == EGNN_SPMM_MEAN); P.act = act;
  P.accumulate = (accumulate || addend) ? 1 : 0;
  P.add_in = addend ? addend : out;
  P.ld_add = addend ? ld_addend : ld_out;
  EGNN_REQUIRE(!addend || ld_addend >= n_feat, fn, "ld_addend < n_feat");
  const size_t in_es = in_dtype == EGNN_F32 ? 4 : 2, out_es = out_dtype == EGNN_F32 ? 4 : 2;
  // 4-feature vector accesses need 4-element-aligned rows and 16 B (fp32) / 8 B (bf16) bases
  bool vec_ok = (n_feat % 4 == 0) && (ld_in % 4 == 0) && (ld_out % 4 == 0) &&
                ((uintptr_t)in % (4 * in_es) == 0) && ((uintptr_t)out % (4 * out_es) == 0) &&
                (!bias || (uintptr_t)bias % 16 == 0) &&
                (!addend || (ld_addend % 4 == 0 && (uintptr_t)addend % (4 * out_es) == 0));
  if (!vec_ok) { P.long_rows = nullptr; P.n_long = nullptr; }
  cudaStream_t st = (cudaStream_t)stream;
  if (vec_ok && mode != EGNN_SPMM_DIV_NBR) {
    int rc = spmm_tile_launch(P, in_dtype, out_dtype, mode == EGNN_SPMM_WEIGHTED, st);
    if (rc != -2) return rc;
  }
  switch (mode) {
    case EGNN_SPMM_SUM:
    case EGNN_SPMM_MEAN: return dispatch_dtype<M_PLAIN>(P, in_dtype, out_dtype, vec_ok, st);
    case EGNN_SPMM_DIV_NBR: return dispatch_dtype<M_DIV_NBR>(P, in_dtype, out_dtype, vec_ok, st);
    default: return dispatch_dtype<M_WEIGHTED>(P, in_dtype, out_dtype, vec_ok, st);
  }
}
