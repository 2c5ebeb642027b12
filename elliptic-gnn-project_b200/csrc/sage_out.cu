// Narrow-output SAGEConv (the `hidden -> 2` logits layer of every SAGE net, src/models/gnn.py:44,128).
//
// PyG computes lin_l(mean_j h_j) + lin_r(h_i): a mean aggregation at the layer's INPUT width
// (64 / 128 columns) followed by two [N,K]x[K,2] products.  Aggregation and projection are both
// linear, so for C <= 4 output channels the layer is evaluated project-first:
//     p      = h . [W_l ; W_r]^T                      [N, 2C]   one pass over h (skinny_project)
//     out_i  = (mean_{j->i} p_j[0:C] + b) + p_i[C:2C]           gather at width C (sage_out_fwd)
// and the backward mirrors it:
//     dp_j   = [ sum_{j->i} dout_i / deg_i  |  dout_j ]         transposed gather at width C
//     dW     = dp^T h,  db = sum_i dout_i                        one pass over h (skinny_wgrad)
//     dh     = dp . [W_l ; W_r]                                  one pass writing dh (skinny_dgrad)
// HBM traffic drops from ~10 passes over an [N,K] matrix to 3.  All sums are sequential in stored
// edge order (gathers) or fixed-shape trees (reductions): deterministic, no float atomics.  The
// projection keeps fp32 weights and fp32 accumulation whatever the activation dtype.
#include "common.cuh"

namespace egnn {
namespace {

constexpr int kThreads = 256;
constexpr int kMaxP = 8;      // 2C
constexpr int kMaxK = 1024;   // weight panel staged in shared memory: P * K floats <= 32 KB

template <typename T>
__device__ __forceinline__ void load8(const T* p, float (&v)[8]);
template <>
__device__ __forceinline__ void load8<float>(const float* p, float (&v)[8]) {
  float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <>
__device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
  F8 r = ld8(p);
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = r.v[i];
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *(reinterpret_cast<float4*>(p) + 1) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&v)[8]) {
  F8 r;
#pragma unroll
  for (int i = 0; i < 8; ++i) r.v[i] = v[i];
  st8(p, r);
}

// ---- p[r, 0:P] = sum_k a[r,k] * W[p,k] ------------------------------------------------------
// G lanes per row, each lane owns 8 consecutive k per sweep; xor-tree over the G lanes.
template <typename T, int P, bool VEC>
__global__ void __launch_bounds__(kThreads) skinny_project_kernel(const T* __restrict__ a, int64_t ld,
                                                                  int64_t n_rows, int K,
                                                                  const float* __restrict__ W,
                                                                  float* __restrict__ out, int G) {
  extern __shared__ float sW[];  // [P][K]
  for (int i = threadIdx.x; i < P * K; i += kThreads) sW[i] = W[i];
  __syncthreads();
  const int rows_per_block = kThreads / G;
  const int g = threadIdx.x / G, lane = threadIdx.x % G;
  // block-uniform trip count: every lane reaches the shuffles below
  for (int64_t base = (int64_t)blockIdx.x * rows_per_block; base < n_rows;
       base += (int64_t)gridDim.x * rows_per_block) {
    const int64_t r = base + g;
    const bool valid = r < n_rows;
    float s[P];
#pragma unroll
    for (int p = 0; p < P; ++p) s[p] = 0.f;
    const T* row = a + (valid ? r : 0) * ld;
    if (VEC) {
      for (int k0 = lane * 8; valid && k0 < K; k0 += G * 8) {
        float v[8];
        load8<T>(row + k0, v);
#pragma unroll
        for (int p = 0; p < P; ++p) {
          const float4 w0 = *reinterpret_cast<const float4*>(sW + p * K + k0);
          const float4 w1 = *reinterpret_cast<const float4*>(sW + p * K + k0 + 4);
          s[p] = fmaf(v[0], w0.x, s[p]); s[p] = fmaf(v[1], w0.y, s[p]);
          s[p] = fmaf(v[2], w0.z, s[p]); s[p] = fmaf(v[3], w0.w, s[p]);
          s[p] = fmaf(v[4], w1.x, s[p]); s[p] = fmaf(v[5], w1.y, s[p]);
          s[p] = fmaf(v[6], w1.z, s[p]); s[p] = fmaf(v[7], w1.w, s[p]);
        }
      }
    } else {
      for (int k = lane; valid && k < K; k += G) {
        const float v = to_f32(row[k]);
#pragma unroll
        for (int p = 0; p < P; ++p) s[p] = fmaf(v, sW[p * K + k], s[p]);
      }
    }
#pragma unroll
    for (int p = 0; p < P; ++p)
      for (int o = G >> 1; o > 0; o >>= 1) s[p] += __shfl_xor_sync(0xffffffffu, s[p], o, G);
    if (lane == 0 && valid) {
#pragma unroll
      for (int p = 0; p < P; ++p) out[r * P + p] = s[p];
    }
  }
}

// Width-C gathers in two uniform passes (no per-row dependent chains, whatever the degree skew: the
// synthetic generator's preferential attachment puts a timestep's hubs into the same warp):
//   pass 1, one thread per ENTRY of the sorted view: term[e] = value gathered for entry e;
//   pass 2, one thread per ROW: sequential fp32 sum of its contiguous terms in stored order.
// `w` (GCN: gcn_norm weight of the entry, rounded product like PyG's w * x_j) or NULL (SAGE); p has row stride ldp
template <int C>
__global__ void __launch_bounds__(kThreads) sage_out_fwd_terms(const int32_t* __restrict__ ptr,
                                                               const int32_t* __restrict__ col,
                                                               const float* __restrict__ w,
                                                               const float* __restrict__ p, int ldp,
                                                               float* __restrict__ term, int64_t n_rows) {
  const int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (e >= __ldg(ptr + n_rows)) return;
  const float* src = p + (int64_t)__ldg(col + e) * ldp;
  const float we = w ? __ldg(w + e) : 1.f;
#pragma unroll
  for (int c = 0; c < C; ++c) term[e * C + c] = w ? __fmul_rn(we, __ldg(src + c)) : __ldg(src + c);
}

template <typename TD, int C>
__global__ void __launch_bounds__(kThreads) sage_out_bwd_terms(const int32_t* __restrict__ csc_ptr,
                                                               const int32_t* __restrict__ csc_dst,
                                                               const int32_t* __restrict__ csr_ptr,
                                                               const float* __restrict__ w,
                                                               const TD* __restrict__ dout,
                                                               float* __restrict__ term, int64_t n_rows) {
  const int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (e >= __ldg(csc_ptr + n_rows)) return;
  const int i = __ldg(csc_dst + e);
  if (w) {  // GCN: dp_j += w_e * dout_i
    const float we = __ldg(w + e);
#pragma unroll
    for (int c = 0; c < C; ++c) term[e * C + c] = __fmul_rn(we, to_f32(dout[(int64_t)i * C + c]));
    return;
  }
  const int d = __ldg(csr_ptr + i + 1) - __ldg(csr_ptr + i);
  const float cnt = (float)(d > 1 ? d : 1);
#pragma unroll
  for (int c = 0; c < C; ++c) term[e * C + c] = __fdiv_rn(to_f32(dout[(int64_t)i * C + c]), cnt);
}

template <int C>
__device__ __forceinline__ void sum_terms(const float* __restrict__ term, int p0, int p1, float (&acc)[C]) {
#pragma unroll
  for (int c = 0; c < C; ++c) acc[c] = 0.f;
  int e = p0;
  for (; e + 4 <= p1; e += 4) {  // four independent loads in flight, added in stored order
    float v[4][C];
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int c = 0; c < C; ++c) v[b][c] = __ldg(term + (int64_t)(e + b) * C + c);
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int c = 0; c < C; ++c) acc[c] = __fadd_rn(acc[c], v[b][c]);
  }
  for (; e < p1; ++e)
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = __fadd_rn(acc[c], __ldg(term + (int64_t)e * C + c));
}

// One thread per row sums its contiguous terms in stored order; a row with more than kWarpRow terms (a hub: several
// hundred) would be the tail of the whole launch, so the WARP takes those rows one after the other -- lane l adds
// terms l, l+32, ... and a fixed butterfly combines the lanes (deterministic; the logits layer is evaluated
// project-first, so its summation order differs from PyG's anyway).  Every lane of the warp must call this.
constexpr int kWarpRow = 32;
template <int C>
__device__ __forceinline__ void row_sum(const float* __restrict__ term, int p0, int p1, bool in_range, float (&acc)[C]) {
  const int lane = threadIdx.x & 31;
  const bool is_long = in_range && (p1 - p0) > kWarpRow;
  if (in_range && !is_long) {
    sum_terms<C>(term, p0, p1, acc);
  } else {
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = 0.f;
  }
  unsigned m = __ballot_sync(0xffffffffu, is_long);
  while (m) {
    const int src = __ffs(m) - 1;
    m &= m - 1;
    const int q0 = __shfl_sync(0xffffffffu, p0, src), q1 = __shfl_sync(0xffffffffu, p1, src);
    float t[C];
#pragma unroll
    for (int c = 0; c < C; ++c) t[c] = 0.f;
    for (int e = q0 + lane; e < q1; e += 32) {
#pragma unroll
      for (int c = 0; c < C; ++c) t[c] = __fadd_rn(t[c], __ldg(term + (int64_t)e * C + c));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int c = 0; c < C; ++c) t[c] = __fadd_rn(t[c], __shfl_xor_sync(0xffffffffu, t[c], o));
    }
    if (lane == src) {
#pragma unroll
      for (int c = 0; c < C; ++c) acc[c] = t[c];
    }
  }
}

// ---- out[i, c] = (mean_{j->i} p[j, c] + b[c]) + p[i, C + c] ----------------------------------
template <int C>
__global__ void __launch_bounds__(kThreads) sage_out_fwd_kernel(const int32_t* __restrict__ ptr,
                                                                const float* __restrict__ term,
                                                                const float* __restrict__ p,
                                                                const float* __restrict__ bias,
                                                                float* __restrict__ out, int64_t n_rows, int gcn) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  const bool in_range = i < n_rows;
  const int p0 = in_range ? __ldg(ptr + i) : 0, p1 = in_range ? __ldg(ptr + i + 1) : 0;
  float acc[C];
  row_sum<C>(term, p0, p1, in_range, acc);
  if (!in_range) return;
  if (gcn) {  // GCNConv: weighted sum + bias (no mean, no root term)
#pragma unroll
    for (int c = 0; c < C; ++c) out[i * C + c] = bias ? __fadd_rn(acc[c], __ldg(bias + c)) : acc[c];
    return;
  }
  const int deg = p1 - p0;
  const float cnt = (float)(deg > 1 ? deg : 1);
#pragma unroll
  for (int c = 0; c < C; ++c) {
    float m = __fdiv_rn(acc[c], cnt);
    if (bias) m = __fadd_rn(m, __ldg(bias + c));
    out[i * C + c] = __fadd_rn(m, __ldg(p + i * (2 * C) + C + c));
  }
}

// ---- dp[j, 0:C] = sum_{j->i} dout[i, :] / max(deg_in(i), 1);  dp[j, C:2C] = dout[j, :] ---------
template <typename TD, int C>
__global__ void __launch_bounds__(kThreads) sage_out_bwd_kernel(const int32_t* __restrict__ csc_ptr,
                                                                const float* __restrict__ term,
                                                                const TD* __restrict__ dout,
                                                                float* __restrict__ dp, int64_t n_rows) {
  const int64_t j = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  const bool in_range = j < n_rows;
  const int p0 = in_range ? __ldg(csc_ptr + j) : 0, p1 = in_range ? __ldg(csc_ptr + j + 1) : 0;
  float acc[C];
  row_sum<C>(term, p0, p1, in_range, acc);
  if (!in_range) return;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    dp[j * (2 * C) + c] = acc[c];
    dp[j * (2 * C) + C + c] = to_f32(dout[j * C + c]);
  }
}

// ---- dW[p, k] = sum_r dp[r,p] * a[r,k];  dsum[p] = sum_r dp[r,p] -----------------------------
// block = row chunk; thread (row lane rl, k group kg) accumulates P x 8 partials over its rows;
// row lanes are combined in shared memory in a fixed order; partial[blk][P][Kp + 8].
template <typename T, int P>
__global__ void __launch_bounds__(kThreads) skinny_wgrad_kernel(const T* __restrict__ a, int64_t ld,
                                                                const float* __restrict__ dp, int64_t n_rows,
                                                                int K, int KG, int64_t rows_per_block,
                                                                float* __restrict__ partial) {
  extern __shared__ float sred[];  // [RL][P*8] per k group, reused
  const int RL = kThreads / KG;
  const int kg = threadIdx.x % KG, rl = threadIdx.x / KG;
  const int k0 = kg * 8;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block, r1 = min(n_rows, r0 + rows_per_block);
  float acc[P][8];
  float ds[P];
#pragma unroll
  for (int p = 0; p < P; ++p) {
    ds[p] = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[p][i] = 0.f;
  }
  if (k0 < K) {
    for (int64_t r = r0 + rl; r < r1; r += RL) {
      float v[8];
      load8<T>(a + r * ld + k0, v);
      float d[P];
#pragma unroll
      for (int p = 0; p < P; ++p) d[p] = __ldg(dp + r * P + p);
#pragma unroll
      for (int p = 0; p < P; ++p) {
        ds[p] += d[p];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[p][i] = fmaf(d[p], v[i], acc[p][i]);
      }
    }
  }
  const int Kp = KG * 8;
  float* out = partial + (int64_t)blockIdx.x * P * (Kp + 8);
  // combine the RL row lanes: each (kg, p, i) column summed over rl in order 0..RL-1
#pragma unroll
  for (int p = 0; p < P; ++p) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) sred[(rl * KG + kg) * 9 + i] = acc[p][i];
    sred[(rl * KG + kg) * 9 + 8] = ds[p];
    __syncthreads();
    if (rl == 0) {
      float s[9];
#pragma unroll
      for (int i = 0; i < 9; ++i) s[i] = 0.f;
      for (int l = 0; l < RL; ++l)
#pragma unroll
        for (int i = 0; i < 9; ++i) s[i] += sred[(l * KG + kg) * 9 + i];
#pragma unroll
      for (int i = 0; i < 8; ++i) out[p * (Kp + 8) + k0 + i] = s[i];
      if (kg == 0) out[p * (Kp + 8) + Kp] = s[8];
    }
  }
}

// final: dW[p,k] = sum_blk partial (double combine, fixed order); dsum[p] likewise.
// block = 32 output columns x 8 block lanes; lane bl sums blocks bl, bl+8, ... then a fixed tree.
// block = 8 output columns x 32 block lanes; lane bl sums blocks bl, bl+32, ... (four independent
// accumulators keep the loads in flight), then a fixed-order combine over the 32 lanes.
__global__ void __launch_bounds__(kThreads) skinny_wgrad_final(const float* __restrict__ partial, int nblk, int P,
                                                               int K, int Kp, float* __restrict__ dW,
                                                               float* __restrict__ dsum, int split = 1 << 30,
                                                               float* __restrict__ dW_hi = nullptr,
                                                               float* __restrict__ dsum_hi = nullptr) {
  __shared__ double sm[32][9];
  const int kl = threadIdx.x & 7, bl = threadIdx.x >> 3;
  const int k = blockIdx.x * 8 + kl, p = blockIdx.y;
  const bool ok = k <= K;  // k == K is the dsum column
  const int src = k < K ? k : Kp;
  const int64_t stride = (int64_t)P * (Kp + 8);
  const float* base = partial + (int64_t)p * (Kp + 8) + src;
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  if (ok) {
    int b = bl;
    for (; b + 96 < nblk; b += 128) {
      s0 += (double)base[(int64_t)b * stride];
      s1 += (double)base[(int64_t)(b + 32) * stride];
      s2 += (double)base[(int64_t)(b + 64) * stride];
      s3 += (double)base[(int64_t)(b + 96) * stride];
    }
    for (; b < nblk; b += 32) s0 += (double)base[(int64_t)b * stride];
  }
  sm[bl][kl] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (bl == 0 && ok) {
    double t = 0;
#pragma unroll
    for (int l = 0; l < 32; ++l) t += sm[l][kl];
    // rows >= split go to their own destinations (d lin_r.weight / the bias gradient of the project-first logits layer)
    if (p >= split) {
      if (k < K) dW_hi[(p - split) * K + k] = (float)t;
      else if (dsum_hi) dsum_hi[p - split] = (float)t;
    } else if (k < K) dW[p * K + k] = (float)t;
    else if (dsum) dsum[p] = (float)t;
  }
}

// ---- dh[r, k] = sum_p dp[r,p] * W[p,k] --------------------------------------------------------
template <typename T, int P>
__global__ void __launch_bounds__(kThreads) skinny_dgrad_kernel(const float* __restrict__ dp,
                                                                const float* __restrict__ W, T* __restrict__ dh,
                                                                int64_t ld, int64_t n_rows, int K) {
  extern __shared__ float sW[];  // [P][K]
  for (int i = threadIdx.x; i < P * K; i += kThreads) sW[i] = W[i];
  __syncthreads();
  const int KG = K / 8;
  const int64_t total = n_rows * KG;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kThreads) {
    const int64_t r = i / KG;
    const int k0 = (int)(i - r * KG) * 8;
    float d[P];
#pragma unroll
    for (int p = 0; p < P; ++p) d[p] = __ldg(dp + r * P + p);
    float v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = 0.f;
#pragma unroll
    for (int p = 0; p < P; ++p)
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = fmaf(d[p], sW[p * K + k0 + q], v[q]);
    store8(dh + r * ld + k0, v);
  }
}

// the same product when K / 8 is a power of two <= 256 (hidden widths 64 / 128 / 256): a thread keeps ONE column
// group for its whole life, so its P x 8 slice of W lives in registers (no shared-memory reads, no 64-bit division
// per element) and the loop is: 16 bytes of dp in, 8 results out.  (The staged version above ran at 1.2 TB/s of
// output on the 8x graph: 32 LDS wavefronts per 512 bytes stored.)
template <typename T, int P>
__global__ void __launch_bounds__(kThreads) skinny_dgrad_reg_kernel(const float* __restrict__ dp,
                                                                    const float* __restrict__ W, T* __restrict__ dh,
                                                                    int64_t ld, int64_t n_rows, int K, int kg_shift) {
  const int KG = 1 << kg_shift;
  const int kg = threadIdx.x & (KG - 1);
  const int k0 = kg * 8;
  float w[P][8];
#pragma unroll
  for (int p = 0; p < P; ++p)
#pragma unroll
    for (int q = 0; q < 8; ++q) w[p][q] = __ldg(W + p * K + k0 + q);
  const int64_t rows_per_pass = ((int64_t)gridDim.x * kThreads) >> kg_shift;
  for (int64_t r = ((int64_t)blockIdx.x * kThreads + threadIdx.x) >> kg_shift; r < n_rows; r += rows_per_pass) {
    float d[P];
    if constexpr (P % 4 == 0) {
#pragma unroll
      for (int p = 0; p < P; p += 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(dp + r * P + p));
        d[p] = t.x; d[p + 1] = t.y; d[p + 2] = t.z; d[p + 3] = t.w;
      }
    } else {
      const float2 t = __ldg(reinterpret_cast<const float2*>(dp + r * P));
      d[0] = t.x; d[1] = t.y;
    }
    float v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = 0.f;
#pragma unroll
    for (int p = 0; p < P; ++p)
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = fmaf(d[p], w[p][q], v[q]);
    store8(dh + r * ld + k0, v);
  }
}

inline int pick_groups(int K) {  // lanes per row for the projection: 8 elements per lane
  int g = 1;
  while (g < 32 && g * 8 < K) g <<= 1;
  return g;
}

}  // namespace
}  // namespace egnn

using namespace egnn;

#define BY_P(P, CALL)        \
  switch (P) {               \
    case 2: { constexpr int kP = 2; CALL; } break; \
    case 4: { constexpr int kP = 4; CALL; } break; \
    case 8: { constexpr int kP = 8; CALL; } break; \
    default: return fail(fn, "P must be 2, 4 or 8"); \
  }

extern "C" int egnn_skinny_project(const void* a, int dtype, int64_t ld, int64_t n_rows, int64_t K,
                                   const float* W, int P, float* out, void* stream) {
  const char* fn = "egnn_skinny_project";
  EGNN_REQUIRE(a && W && out, fn, "null pointer");
  EGNN_REQUIRE(K > 0 && K <= kMaxK && ld >= K, fn, "bad K / ld");
  if (n_rows == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t es = dtype == EGNN_F32 ? 4 : 2;
  const bool vec = K % 8 == 0 && ld % 8 == 0 && (uintptr_t)a % (8 * es) == 0;
  const int G = pick_groups((int)K);
  int64_t blocks = ceil_div(n_rows, kThreads / G);
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  const size_t smem = (size_t)P * K * sizeof(float);
#define LAUNCH(T, V) skinny_project_kernel<T, kP, V><<<(unsigned)blocks, kThreads, smem, st>>>( \
      reinterpret_cast<const T*>(a), ld, n_rows, (int)K, W, out, G)
  if (dtype == EGNN_F32) { if (vec) { BY_P(P, LAUNCH(float, true)) } else { BY_P(P, LAUNCH(float, false)) } }
  else if (dtype == EGNN_BF16) {
    if (vec) { BY_P(P, LAUNCH(__nv_bfloat16, true)) } else { BY_P(P, LAUNCH(__nv_bfloat16, false)) }
  } else return fail(fn, "bad dtype");
#undef LAUNCH
  EGNN_LAUNCH_CHECK(fn);
  return 0;
}

static int narrow_out_fwd(const char* fn, const int32_t* csr_ptr, const int32_t* csr_src, const float* w,
                          const float* p, const float* bias, int C, float* out, int64_t n_rows, float* edge_tmp,
                          int64_t edge_cap, void* stream);

extern "C" int egnn_sage_out_fwd(const int32_t* csr_ptr, const int32_t* csr_src, const float* p,
                                 const float* bias, int C, float* out, int64_t n_rows, float* edge_tmp,
                                 int64_t edge_cap, void* stream) {
  return narrow_out_fwd("egnn_sage_out_fwd", csr_ptr, csr_src, nullptr, p, bias, C, out, n_rows, edge_tmp, edge_cap,
                        stream);
}

extern "C" int egnn_gcn_out_fwd(const int32_t* csr_ptr, const int32_t* csr_src, const float* w_csr, const float* p,
                                const float* bias, int C, float* out, int64_t n_rows, float* edge_tmp,
                                int64_t edge_cap, void* stream) {
  if (!w_csr) return fail("egnn_gcn_out_fwd", "null weights");
  return narrow_out_fwd("egnn_gcn_out_fwd", csr_ptr, csr_src, w_csr, p, bias, C, out, n_rows, edge_tmp, edge_cap,
                        stream);
}

static int narrow_out_fwd(const char* fn, const int32_t* csr_ptr, const int32_t* csr_src, const float* w,
                          const float* p, const float* bias, int C, float* out, int64_t n_rows, float* edge_tmp,
                          int64_t edge_cap, void* stream) {
  EGNN_REQUIRE(csr_ptr && csr_src && p && out && edge_tmp, fn, "null pointer");
  if (n_rows == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)ceil_div(n_rows, kThreads);
  const unsigned egrid = (unsigned)ceil_div(edge_cap > 0 ? edge_cap : 1, kThreads);
#define LAUNCH(CC)                                                                                   \
  sage_out_fwd_terms<CC><<<egrid, kThreads, 0, st>>>(csr_ptr, csr_src, w, p, w ? CC : 2 * CC, edge_tmp, n_rows); \
  EGNN_LAUNCH_CHECK(fn);                                                                             \
  sage_out_fwd_kernel<CC><<<grid, kThreads, 0, st>>>(csr_ptr, edge_tmp, p, bias, out, n_rows, w ? 1 : 0)
  switch (C) {
    case 1: LAUNCH(1); break;
    case 2: LAUNCH(2); break;
    case 4: LAUNCH(4); break;
    default: return fail(fn, "C must be 1, 2 or 4");
  }
#undef LAUNCH
  EGNN_LAUNCH_CHECK(fn);
  return 0;
}

static int narrow_out_bwd(const char* fn, const int32_t* csc_ptr, const int32_t* csc_dst, const int32_t* csr_ptr,
                          const float* w, const void* dout, int dtype, int C, float* dp, int64_t n_rows,
                          float* edge_tmp, int64_t edge_cap, void* stream);

extern "C" int egnn_sage_out_bwd(const int32_t* csc_ptr, const int32_t* csc_dst, const int32_t* csr_ptr,
                                 const void* dout, int dtype, int C, float* dp, int64_t n_rows, float* edge_tmp,
                                 int64_t edge_cap, void* stream) {
  if (!csr_ptr) return fail("egnn_sage_out_bwd", "null pointer");
  return narrow_out_bwd("egnn_sage_out_bwd", csc_ptr, csc_dst, csr_ptr, nullptr, dout, dtype, C, dp, n_rows, edge_tmp,
                        edge_cap, stream);
}

extern "C" int egnn_gcn_out_bwd(const int32_t* csc_ptr, const int32_t* csc_dst, const float* w_csc, const void* dout,
                                int dtype, int C, float* dp, int64_t n_rows, float* edge_tmp, int64_t edge_cap,
                                void* stream) {
  if (!w_csc) return fail("egnn_gcn_out_bwd", "null weights");
  return narrow_out_bwd("egnn_gcn_out_bwd", csc_ptr, csc_dst, csc_ptr, w_csc, dout, dtype, C, dp, n_rows, edge_tmp,
                        edge_cap, stream);
}

static int narrow_out_bwd(const char* fn, const int32_t* csc_ptr, const int32_t* csc_dst, const int32_t* csr_ptr,
                          const float* w, const void* dout, int dtype, int C, float* dp, int64_t n_rows,
                          float* edge_tmp, int64_t edge_cap, void* stream) {
  EGNN_REQUIRE(csc_ptr && csc_dst && csr_ptr && dout && dp && edge_tmp, fn, "null pointer");
  if (n_rows == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)ceil_div(n_rows, kThreads);
  const unsigned egrid = (unsigned)ceil_div(edge_cap > 0 ? edge_cap : 1, kThreads);
#define LAUNCH(T, CC)                                                                                          \
  sage_out_bwd_terms<T, CC><<<egrid, kThreads, 0, st>>>(csc_ptr, csc_dst, csr_ptr, w,                          \
                                                        reinterpret_cast<const T*>(dout), edge_tmp, n_rows);   \
  EGNN_LAUNCH_CHECK(fn);                                                                                       \
  sage_out_bwd_kernel<T, CC><<<grid, kThreads, 0, st>>>(csc_ptr, edge_tmp, reinterpret_cast<const T*>(dout), dp, \
                                                        n_rows)
  if (dtype == EGNN_F32) {
    switch (C) { case 1: LAUNCH(float, 1); break; case 2: LAUNCH(float, 2); break; case 4: LAUNCH(float, 4); break;
      default: return fail(fn, "C must be 1, 2 or 4"); }
  } else if (dtype == EGNN_BF16) {
    switch (C) { case 1: LAUNCH(__nv_bfloat16, 1); break; case 2: LAUNCH(__nv_bfloat16, 2); break;
      case 4: LAUNCH(__nv_bfloat16, 4); break; default: return fail(fn, "C must be 1, 2 or 4"); }
  } else return fail(fn, "bad dtype");
#undef LAUNCH
  EGNN_LAUNCH_CHECK(fn);
  return 0;
}

namespace {
struct WgradPlan { int KG, RL; int64_t rpb; int nblk; int Kp; };
WgradPlan wgrad_plan(int64_t n_rows, int64_t K) {
  WgradPlan w;
  int kg = 1;
  while (kg * 8 < K) kg <<= 1;  // power of two so that it divides 256
  if (kg > kThreads) kg = kThreads;
  w.KG = kg; w.RL = kThreads / kg; w.Kp = kg * 8;
  int64_t target = kNumSMs * 4;
  int64_t rpb = ceil_div(ceil_div(n_rows > 0 ? n_rows : 1, target), w.RL) * w.RL;
  if (rpb < w.RL * 8) rpb = w.RL * 8;
  w.rpb = rpb;
  w.nblk = (int)ceil_div(n_rows > 0 ? n_rows : 1, rpb);
  return w;
}
}  // namespace

extern "C" size_t egnn_skinny_wgrad_workspace_floats(int64_t n_rows, int64_t K, int P) {
  WgradPlan w = wgrad_plan(n_rows, K);
  return (size_t)w.nblk * P * (w.Kp + 8);
}

static int skinny_wgrad_impl(const void* a, int dtype, int64_t ld, const float* dp, int P, int64_t n_rows, int64_t K,
                             float* dW, float* dsum, float* workspace, int split, float* dW_hi, float* dsum_hi,
                             void* stream);

extern "C" int egnn_skinny_wgrad(const void* a, int dtype, int64_t ld, const float* dp, int P, int64_t n_rows,
                                 int64_t K, float* dW, float* dsum, float* workspace, void* stream) {
  return skinny_wgrad_impl(a, dtype, ld, dp, P, n_rows, K, dW, dsum, workspace, 1 << 30, nullptr, nullptr, stream);
}

extern "C" int egnn_skinny_wgrad_split(const void* a, int dtype, int64_t ld, const float* dp, int P, int64_t n_rows,
                                       int64_t K, float* dW_lo, float* dW_hi, float* dsum_hi, float* workspace,
                                       void* stream) {
  if (!dW_hi || P % 2) return fail("egnn_skinny_wgrad_split", "dW_hi missing or odd P");
  return skinny_wgrad_impl(a, dtype, ld, dp, P, n_rows, K, dW_lo, nullptr, workspace, P / 2, dW_hi, dsum_hi, stream);
}

static int skinny_wgrad_impl(const void* a, int dtype, int64_t ld, const float* dp, int P, int64_t n_rows, int64_t K,
                             float* dW, float* dsum, float* workspace, int split, float* dW_hi, float* dsum_hi,
                             void* stream) {
  const char* fn = "egnn_skinny_wgrad";
  EGNN_REQUIRE(a && dp && dW && workspace, fn, "null pointer");
  EGNN_REQUIRE(K > 0 && K <= kMaxK && K % 8 == 0 && ld % 8 == 0, fn, "K and ld must be multiples of 8");
  const size_t es = dtype == EGNN_F32 ? 4 : 2;
  EGNN_REQUIRE((uintptr_t)a % (8 * es) == 0, fn, "misaligned input");
  cudaStream_t st = (cudaStream_t)stream;
  WgradPlan w = wgrad_plan(n_rows, K);
  const size_t smem = (size_t)kThreads * 9 * sizeof(float);
#define LAUNCH(T) skinny_wgrad_kernel<T, kP><<<(unsigned)w.nblk, kThreads, smem, st>>>( \
      reinterpret_cast<const T*>(a), ld, dp, n_rows, (int)K, w.KG, w.rpb, workspace)
  if (dtype == EGNN_F32) { BY_P(P, LAUNCH(float)) }
  else if (dtype == EGNN_BF16) { BY_P(P, LAUNCH(__nv_bfloat16)) }
  else return fail(fn, "bad dtype");
#undef LAUNCH
  EGNN_LAUNCH_CHECK(fn);
  skinny_wgrad_final<<<dim3((unsigned)ceil_div(K + 1, 8), (unsigned)P), kThreads, 0, st>>>(
      workspace, w.nblk, P, (int)K, w.Kp, dW, dsum, split, dW_hi, dsum_hi);
  EGNN_LAUNCH_CHECK(fn);
  return 0;
}

extern "C" int egnn_skinny_dgrad(const float* dp, const float* W, int P, void* dh, int dtype, int64_t ld,
                                 int64_t n_rows, int64_t K, void* stream) {
  const char* fn = "egnn_skinny_dgrad";
  EGNN_REQUIRE(dp && W && dh, fn, "null pointer");
  EGNN_REQUIRE(K > 0 && K <= kMaxK && K % 8 == 0 && ld % 8 == 0, fn, "K and ld must be multiples of 8");
  const size_t es = dtype == EGNN_F32 ? 4 : 2;
  EGNN_REQUIRE((uintptr_t)dh % (8 * es) == 0, fn, "misaligned output");
  if (n_rows == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  int64_t blocks = ceil_div(n_rows * (K / 8), kThreads);
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  const size_t smem = (size_t)P * K * sizeof(float);
  const int KG = (int)(K / 8);
  if ((KG & (KG - 1)) == 0 && KG <= kThreads && (uintptr_t)dp % 16 == 0) {   // register-resident weight slice
    int sh = 0;
    while ((1 << sh) < KG) ++sh;
#define LAUNCH(T) skinny_dgrad_reg_kernel<T, kP><<<(unsigned)blocks, kThreads, 0, st>>>( \
      dp, W, reinterpret_cast<T*>(dh), ld, n_rows, (int)K, sh)
    if (dtype == EGNN_F32) { BY_P(P, LAUNCH(float)) }
    else if (dtype == EGNN_BF16) { BY_P(P, LAUNCH(__nv_bfloat16)) }
    else return fail(fn, "bad dtype");
#undef LAUNCH
    EGNN_LAUNCH_CHECK(fn);
    return 0;
  }
#define LAUNCH(T) skinny_dgrad_kernel<T, kP><<<(unsigned)blocks, kThreads, smem, st>>>( \
      dp, W, reinterpret_cast<T*>(dh), ld, n_rows, (int)K)
  if (dtype == EGNN_F32) { BY_P(P, LAUNCH(float)) }
  else if (dtype == EGNN_BF16) { BY_P(P, LAUNCH(__nv_bfloat16)) }
  else return fail(fn, "bad dtype");
#undef LAUNCH
  EGNN_LAUNCH_CHECK(fn);
  return 0;
}
