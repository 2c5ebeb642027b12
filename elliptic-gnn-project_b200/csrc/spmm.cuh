// Shared declarations of the SpMM kernels (spmm.cu: sub-warp + long-row CTA path and scalar
// fallback; spmm_tile.cu: tile-staged kernel).
#pragma once
#include "common.cuh"

namespace egnn {
namespace spmm_detail {

constexpr int kThreads = 256;
constexpr int kLongRow = 64;    // rows with more edges go to the CTA path (if a list is given)
// Leading blocks that serve the long-row list, one (row, 32-feature slice) item at a time.  4 per SM: with 32
// (the first value) these CTAs were the critical path of EVERY aggregation kernel of round 1 -- 68 long rows x 6
// slices on the base graph at ~5 us per item = 65 us of an ~80 us launch, whatever the main path did; blocks
// with no item left return at once.
constexpr int kLongCtas = 4 * 148;
constexpr int kStageEdges = 128; // source rows staged per round in the CTA path
constexpr int kSliceFeat = 32;   // features per CTA work item in the CTA path
constexpr int kStageFeat = 256;  // max features per feature-chunk (G*VPL*4 <= 256)
constexpr int kTaskCost = 32;    // cost units (rows + 2 * entries) per task of the streaming kernel's row partition

enum { M_PLAIN = 0, M_DIV_NBR = 1, M_WEIGHTED = 2 };

struct Params {
  const int32_t* ptr;
  const int32_t* col;
  const float* w;
  const int32_t* nbr_ptr;
  const int32_t* long_rows;
  const int32_t* n_long;
  const int32_t* row_order;  // optional: rows in descending-degree order (lean kernel schedule)
  const int32_t* part;       // optional: cost-balanced row partition (egnn_spmm_partition), n_tasks + 1 row ids
  int64_t n_tasks;
  const void* in;
  void* out;
  const float* bias;
  int64_t ld_in, ld_out, n_rows;
  int n_feat;      // total features
  int mean;        // divide by max(rowlen,1) after the reduction
  int act;
  int accumulate;       // out = add_in + result (add_in == out for an in-place accumulate)
  const void* add_in;
  int64_t ld_add;
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
    F4 old = ld4(reinterpret_cast<const TO*>(P.add_in) + row * P.ld_add + f);
    a.x = __fadd_rn(old.x, a.x); a.y = __fadd_rn(old.y, a.y);
    a.z = __fadd_rn(old.z, a.z); a.w = __fadd_rn(old.w, a.w);
  }
  st4(o, a);
}

// Whole-CTA path over the long-row list (rows with more than kLongRow entries).
// work item = (long row, 32-feature slice); 128 source rows staged per round in shared memory by
// all threads (memory-level parallelism), then thread t adds feature t over the staged rows
// sequentially -- the same summation order as the sub-warp path.
template <typename TI, typename TO, int MODE, int NT = kThreads>
__device__ __forceinline__ void long_row_path(const Params& P, int bx, float (*s_stage)[kSliceFeat],
                                              float* s_scale, int* s_col) {
  const TI* __restrict__ in = reinterpret_cast<const TI*>(P.in);
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
          for (int i = threadIdx.x; i < ne * nv; i += NT) {
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
}

}  // namespace spmm_detail

// spmm_lean.cu (lane groups, production): returns -2 when the configuration is not covered and the caller
// falls through to the chunked kernels of spmm.cu
int spmm_tile_launch(const spmm_detail::Params& P, int in_dt, int out_dt, bool weighted, cudaStream_t st);
// spmm_stream.cu (lane groups streaming the edges of R consecutive rows); -2 when not covered
int spmm_stream_launch(const spmm_detail::Params& P, int in_dt, int out_dt, bool weighted, cudaStream_t st);
}  // namespace egnn
