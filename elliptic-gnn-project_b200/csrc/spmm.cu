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
#include "common.cuh"

namespace egnn {
namespace {

constexpr int kThreads = 256;
constexpr int kLongRow = 64;    // rows with more edges go to the CTA path (if a list is given)
constexpr int kLongCtas = 32;   // leading blocks that serve the long-row list
constexpr int kStageEdges = 128; // source rows staged per round in the CTA path
constexpr int kSliceFeat = 32;   // features per CTA work item in the CTA path
constexpr int kStageFeat = 256;  // max features per feature-chunk (G*VPL*4 <= 256)

enum { M_PLAIN = 0, M_DIV_NBR = 1, M_WEIGHTED = 2 };

struct Params {
  const int32_t* ptr;
  const int32_t* col;
  const float* w;
  const int32_t* nbr_ptr;
  const int32_t* long_rows;
  const int32_t* n_long;
  const void* in;
  void* out;
  const float* bias;
  int64_t ld_in, ld_out, n_rows;
  int n_feat;      // total features
  int mean;        // divide by max(rowlen,1) after the reduction
  int act;
  int accumulate;
};

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == EGNN_ACT_RELU) return v > 0.f ? v : 0.f;
  if (act == EGNN_ACT_ELU) return v > 0.f ? v : expm1f(v);
  return v;
}

template <int MODE>
__device__ __forceinline__ F4 edge_term(F4 v, float s) {
  if (MODE == M_WEIGHTED) {
    v.x = __fmul_rn(s, v.x); v.y = __fmul_rn(s, v.y); v.z = __fmul_rn(s, v.z); v.w = __fmul_rn(s, v.w);
  } else if (MODE == M_DIV_NBR) {
    v.x = __fdiv_rn(v.x, s); v.y = __fdiv_rn(v.y, s); v.z = __fdiv_rn(v.z, s); v.w = __fdiv_rn(v.w, s);
  }
  return v;
}
__device__ __forceinline__ void acc_add(F4& a, const F4& v) {
  a.x = __fadd_rn(a.x, v.x); a.y = __fadd_rn(a.y, v.y); a.z = __fadd_rn(a.z, v.z); a.w = __fadd_rn(a.w, v.w);
}

template <int MODE>
__device__ __forceinline__ float edge_scale(const Params& P, int p, int c) {
  if (MODE == M_WEIGHTED) return __ldg(P.w + p);
  if (MODE == M_DIV_NBR) {
    int d = __ldg(P.nbr_ptr + c + 1) - __ldg(P.nbr_ptr + c);
    return (float)(d > 1 ? d : 1);
  }
  return 1.f;
}

template <typename TO>
__device__ __forceinline__ void epilogue_store(const Params& P, int64_t row, int f, F4 a, int deg) {
  if (P.mean) {
    float c = (float)(deg > 1 ? deg : 1);
    a.x = __fdiv_rn(a.x, c); a.y = __fdiv_rn(a.y, c); a.z = __fdiv_rn(a.z, c); a.w = __fdiv_rn(a.w, c);
  }
  if (P.bias) {
    float4 b = __ldg(reinterpret_cast<const float4*>(P.bias + f));
    a.x = __fadd_rn(a.x, b.x); a.y = __fadd_rn(a.y, b.y); a.z = __fadd_rn(a.z, b.z); a.w = __fadd_rn(a.w, b.w);
  }
  a.x = apply_act(a.x, P.act); a.y = apply_act(a.y, P.act);
  a.z = apply_act(a.z, P.act); a.w = apply_act(a.w, P.act);
  TO* o = reinterpret_cast<TO*>(P.out) + row * P.ld_out + f;
  if (P.accumulate) {
    F4 old = ld4(o);
    a.x = __fadd_rn(old.x, a.x); a.y = __fadd_rn(old.y, a.y);
    a.z = __fadd_rn(old.z, a.z); a.w = __fadd_rn(old.w, a.w);
  }
  st4(o, a);
}

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
      // ---------------- CTA path over the long-row list ----------------
      // work item = (long row, 32-feature slice); 128 source rows staged per round
      if (blockIdx.y != 0) return;
      const int n_long = *P.n_long;
      const int n_slices = (P.n_feat + kSliceFeat - 1) / kSliceFeat;
      for (int item = bx; item < n_long * n_slices; item += kLongCtas) {
        const int li = item / n_slices, f0 = (item - li * n_slices) * kSliceFeat;
        const int nf = min(kSliceFeat, P.n_feat - f0);  // multiple of 4
        const int nv = nf >> 2;
        const int row = P.long_rows[li];
        const int p0 = P.ptr[row], p1 = P.ptr[row + 1];
        float acc = 0.f;  // thread t < nf owns feature f0 + t
        for (int pb = p0; pb < p1; pb += kStageEdges) {
          const int ne = min(kStageEdges, p1 - pb);
          __syncthreads();  // previous round fully consumed
          if (threadIdx.x < ne) {
            int c = __ldg(P.col + pb + threadIdx.x);
            s_col[threadIdx.x] = c;
            s_scale[threadIdx.x] = edge_scale<MODE>(P, pb + threadIdx.x, c);
          }
          __syncthreads();
          for (int i = threadIdx.x; i < ne * nv; i += kThreads) {
            int e = i / nv, v = i - e * nv;
            F4 x = ld4(in + (int64_t)s_col[e] * P.ld_in + f0 + 4 * v);
            *reinterpret_cast<float4*>(&s_stage[e][4 * v]) = make_float4(x.x, x.y, x.z, x.w);
          }
          __syncthreads();
          if (threadIdx.x < nf) {
            for (int e = 0; e < ne; ++e) {
              float t = s_stage[e][threadIdx.x];
              if (MODE == M_WEIGHTED) t = __fmul_rn(s_scale[e], t);
              if (MODE == M_DIV_NBR) t = __fdiv_rn(t, s_scale[e]);
              acc = __fadd_rn(acc, t);
            }
          }
        }
        // epilogue: regroup 4 features per thread through shared memory
        __syncthreads();
        if (threadIdx.x < nf) s_stage[0][threadIdx.x] = acc;
        __syncthreads();
        if (threadIdx.x < nv) {
          float4 a4 = *reinterpret_cast<float4*>(&s_stage[0][4 * threadIdx.x]);
          epilogue_store<TO>(P, row, f0 + 4 * threadIdx.x, F4{a4.x, a4.y, a4.z, a4.w}, p1 - p0);
        }
      }
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
      if (P.accumulate) a = __fadd_rn(to_f32(*o), a);
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
                         const void* in, int in_dtype, int64_t ld_in, void* out, int out_dtype,
                         int64_t ld_out, int64_t n_rows, int64_t n_feat, const float* bias, int act,
                         int accumulate, void* stream) {
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
  P.long_rows = long_rows; P.n_long = n_long;
  P.in = in; P.out = out; P.bias = bias;
  P.ld_in = ld_in; P.ld_out = ld_out; P.n_rows = n_rows; P.n_feat = (int)n_feat;
  P.mean = (mode == EGNN_SPMM_MEAN); P.act = act; P.accumulate = accumulate;
  const size_t in_es = in_dtype == EGNN_F32 ? 4 : 2, out_es = out_dtype == EGNN_F32 ? 4 : 2;
  // 4-feature vector accesses need 4-element-aligned rows and 16 B (fp32) / 8 B (bf16) bases
  bool vec_ok = (n_feat % 4 == 0) && (ld_in % 4 == 0) && (ld_out % 4 == 0) &&
                ((uintptr_t)in % (4 * in_es) == 0) && ((uintptr_t)out % (4 * out_es) == 0) &&
                (!bias || (uintptr_t)bias % 16 == 0);
  if (!vec_ok) { P.long_rows = nullptr; P.n_long = nullptr; }
  cudaStream_t st = (cudaStream_t)stream;
  switch (mode) {
    case EGNN_SPMM_SUM:
    case EGNN_SPMM_MEAN: return dispatch_dtype<M_PLAIN>(P, in_dtype, out_dtype, vec_ok, st);
    case EGNN_SPMM_DIV_NBR: return dispatch_dtype<M_DIV_NBR>(P, in_dtype, out_dtype, vec_ok, st);
    default: return dispatch_dtype<M_WEIGHTED>(P, in_dtype, out_dtype, vec_ok, st);
  }
}
