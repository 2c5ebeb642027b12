// K7 and the step tail: cast / time-feature injection, deterministic column reductions,
// BatchNorm(batch stats) + ReLU/ELU + dropout + residual forward/backward, masked weighted
// cross-entropy, global-norm clip + Adam.  All HBM-bound streaming kernels: 4 features per
// thread (128-bit fp32 / 64-bit bf16 accesses), fp32 math, fp64 only for cross-thread
// combination of statistics (torch CPU BatchNorm accumulates in double too).
#include "common.cuh"
#include "philox.cuh"

#include <initializer_list>

namespace egnn {
namespace {

constexpr int kThreads = 256;
constexpr int kMaxPartBlocks = 1024;

template <typename T>
__device__ __forceinline__ F4 load4g(const T* base, int64_t ld, int64_t r, int c, int F, bool vec) {
  const T* p = base + r * ld + c;
  if (vec) return ld4(p);
  F4 o{0.f, 0.f, 0.f, 0.f};
  if (c < F) o.x = to_f32(p[0]);
  if (c + 1 < F) o.y = to_f32(p[1]);
  if (c + 2 < F) o.z = to_f32(p[2]);
  if (c + 3 < F) o.w = to_f32(p[3]);
  return o;
}
template <typename T>
__device__ __forceinline__ void store4g(T* base, int64_t ld, int64_t r, int c, int F, bool vec, F4 v) {
  T* p = base + r * ld + c;
  if (vec) {
    st4(p, v);
    return;
  }
  if (c < F) p[0] = from_f32<T>(v.x);
  if (c + 1 < F) p[1] = from_f32<T>(v.y);
  if (c + 2 < F) p[2] = from_f32<T>(v.z);
  if (c + 3 < F) p[3] = from_f32<T>(v.w);
}
inline bool vec_ok(const void* p, int dtype, int64_t ld, int64_t F) {
  size_t es = dtype == EGNN_F32 ? 4 : 2;
  return p == nullptr || ((F % 4 == 0) && (ld % 4 == 0) && ((uintptr_t)p % (4 * es) == 0));
}

// ---- cast / pad -------------------------------------------------------------------------
template <typename TI, typename TO>
__global__ void __launch_bounds__(kThreads) cast_kernel(const TI* __restrict__ in, int64_t ld_in,
                                                        TO* __restrict__ out, int64_t ld_out,
                                                        int64_t n_rows, int F, int Fo4, bool vin,
                                                        bool vout) {
  int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n_rows * Fo4) return;
  int64_t r = i / Fo4;
  int c = (int)(i - r * Fo4) * 4;
  F4 v = load4g(in, ld_in, r, c, F, vin && c + 3 < F);
  store4g(out, ld_out, r, c, (int)ld_out, vout, v);
}

// ---- time-feature injection ----------------------------------------------------------------
constexpr int kInjectRows = 48;  // rows per block
// 4 output columns per thread; a thread walks the block's (row, column group) items with an incremental
// index (no per-item division).  x rows are only 8-byte aligned (ld_x = 166), so they are read as float2.
__global__ void __launch_bounds__(kThreads) inject_time_kernel(
    const float* __restrict__ x, int64_t ld_x, const int64_t* __restrict__ t,
    const float* __restrict__ table, int64_t T, int D, float* __restrict__ o32,
    __nv_bfloat16* __restrict__ o16, int64_t ld_out, int64_t ld_out16, int64_t n_rows, int F) {
  const int W4 = (int)(ld_out / 4);  // ld_out is a multiple of 4
  const int64_t row0 = (int64_t)blockIdx.x * kInjectRows;
  const int nr = (int)min((int64_t)kInjectRows, n_rows - row0);
  const bool x2 = (ld_x % 2 == 0) && ((uintptr_t)x % 8 == 0);
  const int dq = kThreads / W4, dm = kThreads - dq * W4;  // advance of (row, group) per 256 items
  int rr = threadIdx.x / W4, cg = threadIdx.x - rr * W4;
  for (; rr < nr; rr += dq, cg += dm) {
    if (cg >= W4) { cg -= W4; ++rr; if (rr >= nr) break; }
    const int c = cg * 4;
    const int64_t r = row0 + rr;
    float v[4];
    if (x2 && c + 3 < F) {
      const float2 p0 = __ldg(reinterpret_cast<const float2*>(x + r * ld_x + c));
      const float2 p1 = __ldg(reinterpret_cast<const float2*>(x + r * ld_x + c + 2));
      v[0] = p0.x; v[1] = p0.y; v[2] = p1.x; v[3] = p1.y;
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int cc = c + k;
        if (cc < F) {
          v[k] = __ldg(x + r * ld_x + cc);
        } else if (cc < F + D) {
          int64_t ti = t[r] - 1;
          ti = ti < 0 ? 0 : (ti > T - 1 ? T - 1 : ti);
          v[k] = __ldg(table + ti * D + (cc - F));
        } else {
          v[k] = 0.f;
        }
      }
    }
    if (o32) *reinterpret_cast<float4*>(o32 + r * ld_out + c) = make_float4(v[0], v[1], v[2], v[3]);
    if (o16)
      *reinterpret_cast<uint2*>(o16 + r * ld_out16 + c) = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
  }
}

// Gradient of the learned time-embedding table (nn.Embedding backward, src/models/gnn.py:152,172-176):
// dtab[r, d] = sum over nodes n with clamp(t[n]-1, 0, T-1) == r of dout[n, col0 + d].  torch's index_add_ uses float
// atomics on CUDA; here block (r, chunk) scans its node chunk for row r with a fixed thread -> node assignment and a
// fixed shared-memory tree, and embed_grad_final adds the chunk partials in chunk order: same bits on every run.
constexpr int kEmbChunk = 8192;
constexpr int kEmbCols = 8;

template <typename T>
__global__ void __launch_bounds__(kThreads) embed_grad_partial(const T* __restrict__ dout, int64_t ld, int col0, int D,
                                                               const int64_t* __restrict__ t, int64_t Tn, int64_t n_rows,
                                                               double* __restrict__ partial) {
  __shared__ double s[kThreads];
  const int r = blockIdx.x, chunk = blockIdx.y, nchunks = gridDim.y;
  const int64_t lo = (int64_t)chunk * kEmbChunk, hi = min(lo + (int64_t)kEmbChunk, n_rows);
  for (int d0 = 0; d0 < D; d0 += kEmbCols) {
    double acc[kEmbCols];
#pragma unroll
    for (int k = 0; k < kEmbCols; ++k) acc[k] = 0.0;
    for (int64_t n = lo + threadIdx.x; n < hi; n += kThreads) {
      int64_t ti = t[n] - 1;
      ti = ti < 0 ? 0 : (ti > Tn - 1 ? Tn - 1 : ti);
      if (ti != r) continue;
#pragma unroll
      for (int k = 0; k < kEmbCols; ++k)
        if (d0 + k < D) acc[k] += (double)to_f32(dout[n * ld + col0 + d0 + k]);
    }
    for (int k = 0; k < kEmbCols && d0 + k < D; ++k) {
      s[threadIdx.x] = acc[k];
      __syncthreads();
      for (int w = kThreads / 2; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w) s[threadIdx.x] += s[threadIdx.x + w];
        __syncthreads();
      }
      if (threadIdx.x == 0) partial[((int64_t)r * nchunks + chunk) * D + d0 + k] = s[0];
      __syncthreads();
    }
  }
}

__global__ void embed_grad_final(const double* __restrict__ partial, int nchunks, int D, int64_t Tn,
                                 float* __restrict__ dtab) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Tn * D) return;
  const int64_t r = i / D;
  const int d = (int)(i - r * D);
  double a = 0.0;
  for (int c = 0; c < nchunks; ++c) a += partial[(r * nchunks + c) * D + d];
  dtab[i] = (float)a;
}

// ---- fused BN / activation / dropout element math ------------------------------------------
struct ActCtx {
  const float *mean, *rstd, *gamma, *beta;
  int act;
  int drop;  // dropout active
  float scale;
  uint32_t thr;
  uint64_t seed;
  const int64_t* seed_off;  // optional device-side offset added to seed (CUDA-graph replays)
  uint32_t layer;
  int64_t row0;
};

// for 4 columns starting at c (c % 4 == 0): y = dropout(act(bn(z))), dfac = d y / d u, xhat
__device__ __forceinline__ void act_eval(const ActCtx& C, int64_t r, int c, int F, F4 z, F4& y,
                                         F4& dfac, F4& xhat) {
  float zz[4] = {z.x, z.y, z.z, z.w}, yy[4], dd[4], xh[4];
  uint32_t keep = 0xfu;
  if (C.drop) {
    const uint64_t seed = C.seed + (C.seed_off ? (uint64_t)*C.seed_off : 0ull);
    keep = (dropout_keep8(seed, C.layer, C.row0 + r, (uint32_t)(c >> 3), C.thr) >> (c & 4)) & 0xfu;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    int cc = c + k;
    float u = zz[k];
    xh[k] = 0.f;
    if (C.mean && cc < F) {
      xh[k] = (zz[k] - C.mean[cc]) * C.rstd[cc];
      u = xh[k] * C.gamma[cc] + C.beta[cc];
    }
    float a = u, da = 1.f;
    if (C.act == EGNN_ACT_RELU) {
      a = u > 0.f ? u : 0.f;
      da = u > 0.f ? 1.f : 0.f;
    } else if (C.act == EGNN_ACT_ELU) {
      float e = expm1f(u);
      a = u > 0.f ? u : e;
      da = u > 0.f ? 1.f : e + 1.f;
    }
    float ks = 1.f;
    if (C.drop) ks = (keep >> k) & 1u ? C.scale : 0.f;
    yy[k] = a * ks;
    dd[k] = da * ks;
  }
  y = F4{yy[0], yy[1], yy[2], yy[3]};
  dfac = F4{dd[0], dd[1], dd[2], dd[3]};
  xhat = F4{xh[0], xh[1], xh[2], xh[3]};
}

template <typename T>
__global__ void __launch_bounds__(kThreads) bn_act_fwd_kernel(const T* __restrict__ z,
                                                              const T* __restrict__ res,
                                                              T* __restrict__ yout, int64_t ld,
                                                              int64_t n_rows, int F, bool vec,
                                                              ActCtx C) {
  const int F4n = (F + 3) >> 2;
  int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n_rows * F4n) return;
  int64_t r = i / F4n;
  int c = (int)(i - r * F4n) * 4;
  F4 zv = load4g(z, ld, r, c, F, vec), y, d, xh;
  act_eval(C, r, c, F, zv, y, d, xh);
  if (res) {
    F4 rv = load4g(res, ld, r, c, F, vec);
    y.x += rv.x; y.y += rv.y; y.z += rv.z; y.w += rv.w;
  }
  store4g(yout, ld, r, c, F, vec, y);
}

template <typename T>
__global__ void __launch_bounds__(kThreads) bn_act_bwd_apply_kernel(
    const T* __restrict__ dy, const T* __restrict__ z, T* __restrict__ dz, int64_t ld,
    int64_t n_rows, int F, bool vec, ActCtx C, const double* __restrict__ sum_g,
    const double* __restrict__ sum_gx, double inv_n) {
  const int F4n = (F + 3) >> 2;
  int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n_rows * F4n) return;
  int64_t r = i / F4n;
  int c = (int)(i - r * F4n) * 4;
  F4 zv = load4g(z, ld, r, c, F, vec), y, d, xh;
  act_eval(C, r, c, F, zv, y, d, xh);
  F4 g = load4g(dy, ld, r, c, F, vec);
  float gg[4] = {g.x * d.x, g.y * d.y, g.z * d.z, g.w * d.w};
  float xx[4] = {xh.x, xh.y, xh.z, xh.w};
  if (C.mean) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      int cc = c + k;
      if (cc < F) {
        float mg = (float)(sum_g[cc] * inv_n), mgx = (float)(sum_gx[cc] * inv_n);
        gg[k] = C.gamma[cc] * C.rstd[cc] * (gg[k] - mg - xx[k] * mgx);
      }
    }
  }
  store4g(dz, ld, r, c, F, vec, F4{gg[0], gg[1], gg[2], gg[3]});
}

// ---- deterministic column reductions ----------------------------------------------------------
// Producer::eval(r, c, v0, v1): two 4-column values per (row, column group).
template <typename T>
struct PlainProd {
  const T* a;
  int64_t ld;
  int F;
  bool vec;
  __device__ __forceinline__ void eval(int64_t r, int c, F4& v0, F4& v1) const {
    v0 = load4g(a, ld, r, c, F, vec);
    v1 = F4{v0.x * v0.x, v0.y * v0.y, v0.z * v0.z, v0.w * v0.w};
  }
};
template <typename T>
struct BnBwdProd {
  const T* dy;
  const T* z;
  int64_t ld;
  int F;
  bool vec;
  ActCtx C;
  __device__ __forceinline__ void eval(int64_t r, int c, F4& v0, F4& v1) const {
    F4 zv = load4g(z, ld, r, c, F, vec), y, d, xh;
    act_eval(C, r, c, F, zv, y, d, xh);
    F4 g = load4g(dy, ld, r, c, F, vec);
    v0 = F4{g.x * d.x, g.y * d.y, g.z * d.z, g.w * d.w};
    v1 = F4{v0.x * xh.x, v0.y * xh.y, v0.z * xh.z, v0.w * xh.w};
  }
};

// GAT: datt_src[h,c] = sum_n da_s[n,h]*xs[n,h,c], datt_dst likewise (SURVEY.md A.3 backward)
struct GatAttProd {
  const float* xs;
  const float* da_s;
  const float* da_d;
  int H, C, F;
  bool vec;
  __device__ __forceinline__ void eval(int64_t r, int c, F4& v0, F4& v1) const {
    F4 x = load4g(xs, (int64_t)F, r, c, F, vec);
    float xv[4] = {x.x, x.y, x.z, x.w}, o0[4], o1[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      int cc = c + k;
      int h = cc < F ? cc / C : 0;
      o0[k] = da_s[r * H + h] * xv[k];
      o1[k] = da_d[r * H + h] * xv[k];
    }
    v0 = F4{o0[0], o0[1], o0[2], o0[3]};
    v1 = F4{o1[0], o1[1], o1[2], o1[3]};
  }
};

// block = (CW4 column groups) x (RL row lanes); partial[(blk*2+which)*Fp + col], Fp = 4*ceil(F/4)
template <typename Prod>
__global__ void __launch_bounds__(kThreads) colreduce_partial(Prod prod, int64_t n_rows, int F,
                                                              int CW4, int64_t rows_per_block,
                                                              double* __restrict__ partial) {
  __shared__ double sm[kThreads][8];
  const int RL = kThreads / CW4;
  const int cgl = threadIdx.x % CW4, rl = threadIdx.x / CW4;
  const int cg = blockIdx.y * CW4 + cgl;
  const int c = cg * 4;
  const int Fp = ((F + 3) >> 2) << 2;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r1 = min(n_rows, r0 + rows_per_block);
  double d[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (c < F) {
    float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    int cnt = 0;
    // four rows in flight per thread (one row at a time left every thread with a single load group outstanding:
    // 33 us for the 33 MB of the GAT attention gradient); rows past the end contribute exact zeros
    for (int64_t r = r0 + rl; r < r1; r += 4 * (int64_t)RL) {
      F4 v0[4], v1[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t rr = r + (int64_t)u * RL;
        if (rr < r1) prod.eval(rr, c, v0[u], v1[u]);
        else { v0[u] = F4{0.f, 0.f, 0.f, 0.f}; v1[u] = F4{0.f, 0.f, 0.f, 0.f}; }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        a[0] += v0[u].x; a[1] += v0[u].y; a[2] += v0[u].z; a[3] += v0[u].w;
        a[4] += v1[u].x; a[5] += v1[u].y; a[6] += v1[u].z; a[7] += v1[u].w;
      }
      cnt += 4;
      if (cnt >= 32) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { d[k] += (double)a[k]; a[k] = 0.f; }
        cnt = 0;
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) d[k] += (double)a[k];
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) sm[threadIdx.x][k] = d[k];
  __syncthreads();
  if (rl == 0 && c < F) {
    double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int l = 0; l < RL; ++l)
#pragma unroll
      for (int k = 0; k < 8; ++k) s[k] += sm[l * CW4 + cgl][k];
    double* p0 = partial + ((int64_t)blockIdx.x * 2 + 0) * Fp + c;
    double* p1 = partial + ((int64_t)blockIdx.x * 2 + 1) * Fp + c;
#pragma unroll
    for (int k = 0; k < 4; ++k) { p0[k] = s[k]; p1[k] = s[4 + k]; }
  }
}

// one block per (which, column): fixed-order strided sum + fixed tree
__global__ void __launch_bounds__(kThreads) colreduce_final(const double* __restrict__ partial,
                                                            int nblk, int F, double* __restrict__ out0,
                                                            double* __restrict__ out1, float* __restrict__ f0 = nullptr,
                                                            float* __restrict__ f1 = nullptr) {
  __shared__ double sm[kThreads];
  const int Fp = ((F + 3) >> 2) << 2;
  const int col = blockIdx.x, which = blockIdx.y;
  double* out = which == 0 ? out0 : out1;
  float* fout = which == 0 ? f0 : f1;   // optional fp32 copy (a parameter gradient: d beta / d gamma of BatchNorm)
  if (!out) return;
  double s = 0;
  for (int b = threadIdx.x; b < nblk; b += kThreads) s += partial[((int64_t)b * 2 + which) * Fp + col];
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int o = kThreads / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out[col] = sm[0];
    if (fout) fout[col] = (float)sm[0];
  }
}

template <typename Prod>
int run_colreduce(const Prod& prod, int64_t n_rows, int F, double* out0, double* out1, void* ws,
                  cudaStream_t st, const char* fn) {
  int ncg = (F + 3) / 4;
  int CW4 = 1;
  while (CW4 < ncg && CW4 < kThreads) CW4 <<= 1;
  int RL = kThreads / CW4;
  // rows per thread: a multiple of 4 (the rows in flight), 32 for big inputs, fewer when that would leave the grid
  // below ~4 blocks per SM
  int64_t rpt = ceil_div(n_rows > 0 ? n_rows : 1, (int64_t)RL * kNumSMs * 4);
  rpt = rpt < 4 ? 4 : rpt > 32 ? 32 : (rpt + 3) / 4 * 4;
  int64_t rpb = (int64_t)RL * rpt;
  if (ceil_div(n_rows, rpb) > kMaxPartBlocks) rpb = ceil_div(ceil_div(n_rows, kMaxPartBlocks), RL) * RL;
  int nblk = (int)ceil_div(n_rows > 0 ? n_rows : 1, rpb);
  dim3 grid(nblk, (unsigned)ceil_div(ncg, CW4));
  double* partial = reinterpret_cast<double*>(ws);
  colreduce_partial<Prod><<<grid, kThreads, 0, st>>>(prod, n_rows, F, CW4, rpb, partial);
  EGNN_LAUNCH_CHECK(fn);
  colreduce_final<<<dim3(F, 2), kThreads, 0, st>>>(partial, nblk, F, out0, out1);
  EGNN_LAUNCH_CHECK(fn);
  return 0;
}

// ---- fast path: 8 columns per thread, column constants in registers, rows looped ---------------
// Used when F % 8 == 0, F/8 is a power of two <= 256 and every pointer is 16-byte aligned with
// ld % 8 == 0 (the hidden widths 32 / 64 / 128 of the reference configs).  A thread owns 8 fixed
// columns and walks rows; per-column BatchNorm constants are loaded once.  Same element math as
// act_eval (the generic kernels remain the fallback for ragged shapes).
template <typename T>
__device__ __forceinline__ void ld8f(const T* p, float (&v)[8]);
template <>
__device__ __forceinline__ void ld8f<float>(const float* p, float (&v)[8]) {
  float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <>
__device__ __forceinline__ void ld8f<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
  F8 r = ld8(p);
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = r.v[i];
}
__device__ __forceinline__ void st8f(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *(reinterpret_cast<float4*>(p) + 1) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void st8f(__nv_bfloat16* p, const float (&v)[8]) {
  F8 r;
#pragma unroll
  for (int i = 0; i < 8; ++i) r.v[i] = v[i];
  st8(p, r);
}

// raw 8-element loads (held packed while the next iteration's loads are in flight)
template <typename T>
struct Raw8;
template <>
struct Raw8<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) {
    a = __ldg(reinterpret_cast<const float4*>(p));
    b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  }
  __device__ __forceinline__ void unpack(float (&v)[8]) const {
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
};
template <>
struct Raw8<__nv_bfloat16> {
  uint4 q;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) { q = __ldg(reinterpret_cast<const uint4*>(p)); }
  __device__ __forceinline__ void unpack(float (&v)[8]) const {
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
};

// ---- lean streaming kernels: 4 columns per thread, column constants in REGISTERS, 4 rows in flight ---
// (the first 8-column version kept its constants in shared memory and was bound by LDS wavefronts:
//  l1tex 68 %, 43 us for 78 MB, profiles/r01/ncu_kernels_after_fusion.txt).  Templated on BatchNorm
// presence and the activation so the element math is ~12 instructions.  The dropout keep bits of a
// (row, 4-column group) are one byte of `keep_bits` [n_rows, F/4] (low nibble): written by the forward,
// read by the backward kernels instead of re-running Philox.
template <typename T>
struct Raw4;
template <>
struct Raw4<float> {
  float4 a;
  __device__ __forceinline__ void load(const float* p) { a = __ldg(reinterpret_cast<const float4*>(p)); }
  __device__ __forceinline__ void unpack(float (&v)[4]) const { v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; }
};
template <>
struct Raw4<__nv_bfloat16> {
  uint2 q;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) { q = __ldg(reinterpret_cast<const uint2*>(p)); }
  __device__ __forceinline__ void unpack(float (&v)[4]) const {
    v[0] = __uint_as_float(q.x << 16); v[1] = __uint_as_float(q.x & 0xffff0000u);
    v[2] = __uint_as_float(q.y << 16); v[3] = __uint_as_float(q.y & 0xffff0000u);
  }
};
__device__ __forceinline__ void st4f(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void st4f(__nv_bfloat16* p, const float (&v)[4]) {
  *reinterpret_cast<uint2*>(p) = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
}

constexpr int kRowsInFlight = 4;

struct Col4 {
  float mean[4], rstd[4], gamma[4], beta[4];
};
template <bool BN>
__device__ __forceinline__ void load_col4(const ActCtx& C, int c, Col4& k) {
  if (BN) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      k.mean[i] = C.mean[c + i]; k.rstd[i] = C.rstd[c + i]; k.gamma[i] = C.gamma[c + i]; k.beta[i] = C.beta[c + i];
    }
  }
}
// 4 keep bits of (row r, columns c..c+3) from Philox (same words as egnn_dropout_mask)
__device__ __forceinline__ uint32_t keep_bits4(const ActCtx& C, uint64_t seed, int64_t r, int c) {
  return (dropout_keep8(seed, C.layer, C.row0 + r, (uint32_t)(c >> 3), C.thr) >> (c & 4)) & 0xfu;
}
// u = bn(z); a = act(u); returns y = a*ks, dfac = act'(u)*ks, xhat.  bits: low nibble = dropout keep bits; with
// STORED_GATE (backward, ReLU, bits saved by the forward) the high nibble holds the ReLU gates u > 0, so the
// backward needs neither gamma / beta nor the affine map; `gate_out` receives the gates computed here.
template <bool BN, int ACT, bool STORED_GATE = false>
__device__ __forceinline__ void act4(const Col4& k, float scale, uint32_t bits, const float (&z)[4], float (&y)[4],
                                     float (&dfac)[4], float (&xhat)[4], uint32_t* gate_out = nullptr) {
  uint32_t gates = 0u;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float u = z[i];
    xhat[i] = 0.f;
    if (BN) {
      xhat[i] = (z[i] - k.mean[i]) * k.rstd[i];
      if (!STORED_GATE) u = xhat[i] * k.gamma[i] + k.beta[i];
    }
    float a = u, da = 1.f;
    if (ACT == EGNN_ACT_RELU) {
      const bool on = STORED_GATE ? ((bits >> (4 + i)) & 1u) != 0u : u > 0.f;
      gates |= (on ? 1u : 0u) << i;
      a = on ? u : 0.f;
      da = on ? 1.f : 0.f;
    } else if (ACT == EGNN_ACT_ELU) {
      const float e = expm1f(u);
      a = u > 0.f ? u : e;
      da = u > 0.f ? 1.f : e + 1.f;
    }
    const float ks = (bits >> i) & 1u ? scale : 0.f;
    y[i] = a * ks;
    dfac[i] = da * ks;
  }
  if (gate_out) *gate_out = gates;
}

template <typename T, bool BN, int ACT>
__global__ void __launch_bounds__(kThreads, 3) bn_act_fwd_lean(const T* __restrict__ z, const T* __restrict__ res,
                                                            T* __restrict__ yout, int64_t ld, int64_t ld_res,
                                                            int64_t ld_y, int64_t n_rows, int F, int cg_shift,
                                                            int64_t rows_per_block, ActCtx C,
                                                            uint8_t* __restrict__ keep_bits) {
  const int CG = 1 << cg_shift, RL = kThreads >> cg_shift;
  const int cgi = threadIdx.x & (CG - 1), c = cgi * 4, rl = threadIdx.x >> cg_shift;
  Col4 k;
  load_col4<BN>(C, c, k);
  const uint64_t seed = C.seed + (C.seed_off ? (uint64_t)*C.seed_off : 0ull);
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block, r1 = min(n_rows, r0 + rows_per_block);
  for (int64_t rb = r0 + rl; rb < r1; rb += (int64_t)RL * kRowsInFlight) {
    Raw4<T> zr[kRowsInFlight], rr[kRowsInFlight];
#pragma unroll
    for (int u = 0; u < kRowsInFlight; ++u) {
      const int64_t r = rb + (int64_t)u * RL;
      if (r < r1) {
        zr[u].load(z + r * ld + c);
        if (res) rr[u].load(res + r * ld_res + c);
      }
    }
#pragma unroll
    for (int u = 0; u < kRowsInFlight; ++u) {
      const int64_t r = rb + (int64_t)u * RL;
      if (r < r1) {
        float zv[4], y[4], d[4], xh[4];
        zr[u].unpack(zv);
        const uint32_t bits = C.drop ? keep_bits4(C, seed, r, c) : 0xfu;
        uint32_t gates;
        act4<BN, ACT>(k, C.scale, bits, zv, y, d, xh, &gates);
        if (keep_bits) keep_bits[r * CG + cgi] = (uint8_t)(bits | (gates << 4));  // keep bits | ReLU gates
        if (res) {
          float rv[4];
          rr[u].unpack(rv);
#pragma unroll
          for (int i = 0; i < 4; ++i) y[i] += rv[i];
        }
        st4f(yout + r * ld_y + c, y);
      }
    }
  }
}

// ---- 8 columns per thread: one Philox draw (all eight 16-bit lanes) per thread-row, 16-byte bf16 accesses, and --
// optionally (PROJ) the logits-layer projection p[r, 0:4] = y[r, :] . Wp[0:4, :]^T of the NEXT SAGEConv
// (`SAGEConv(hidden, 2)` evaluated project-first: [W_l ; W_r], src/models/gnn.py:128,193) as a by-product, so that
// layer does not read the activation again.  The projection uses the activation AS STORED (rounded to T).
// store 8 values, returning them as stored (rounded to the element type)
__device__ __forceinline__ void st8r(float* p, float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *(reinterpret_cast<float4*>(p) + 1) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void st8r(__nv_bfloat16* p, float (&v)[8]) {
  uint4 q;
  q.x = pack_bf16x2(v[0], v[1]); q.y = pack_bf16x2(v[2], v[3]);
  q.z = pack_bf16x2(v[4], v[5]); q.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = q;
  v[0] = __uint_as_float(q.x << 16); v[1] = __uint_as_float(q.x & 0xffff0000u);
  v[2] = __uint_as_float(q.y << 16); v[3] = __uint_as_float(q.y & 0xffff0000u);
  v[4] = __uint_as_float(q.z << 16); v[5] = __uint_as_float(q.z & 0xffff0000u);
  v[6] = __uint_as_float(q.w << 16); v[7] = __uint_as_float(q.w & 0xffff0000u);
}

constexpr int kRows8 = 4;   // rows in flight per thread
constexpr int kProjP = 4;   // projected outputs (2 classes x [W_l ; W_r])

template <typename T, bool BN, int ACT, bool PROJ>
__global__ void __launch_bounds__(kThreads, (BN || PROJ || sizeof(T) == 4) ? 2 : 3) bn_act_fwd8(const T* __restrict__ z, const T* __restrict__ res,
                                                          T* __restrict__ yout, int64_t ld, int64_t ld_res,
                                                          int64_t ld_y, int64_t n_rows, int F, int cg_shift,
                                                          int64_t rows_per_block, ActCtx C,
                                                          uint8_t* __restrict__ keep_bits,
                                                          const float* __restrict__ Wp, float* __restrict__ pout) {
  const int CG = 1 << cg_shift, RL = kThreads >> cg_shift;   // CG = F / 8 lanes per row (<= 32: one warp)
  const int cgi = threadIdx.x & (CG - 1), c = cgi * 8, rl = threadIdx.x >> cg_shift;
  float mean[8], rstd[8], gamma[8], beta[8];
  if (BN) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      mean[i] = C.mean[c + i]; rstd[i] = C.rstd[c + i]; gamma[i] = C.gamma[c + i]; beta[i] = C.beta[c + i];
    }
  }
  float wp[PROJ ? kProjP : 1][8];
  if (PROJ) {
#pragma unroll
    for (int p = 0; p < kProjP; ++p)
#pragma unroll
      for (int i = 0; i < 8; ++i) wp[PROJ ? p : 0][i] = Wp[p * F + c + i];
  }
  const uint64_t seed = C.seed + (C.seed_off ? (uint64_t)*C.seed_off : 0ull);
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block, r1 = min(n_rows, r0 + rows_per_block);
  // uniform trip count over the block (rows_per_block is a multiple of RL * kRows8): every lane reaches the shuffles
  // fp32 rows with the projection folded in: 2 rows in flight (32-byte raw rows of z and res x 4 rows + the 32
  // projection weights + the 32 BatchNorm constants do not fit 128 registers: 0.9 KB of spill traffic per thread-row)
  constexpr int kR = (sizeof(T) == 4 && PROJ && BN) ? 2 : kRows8;
  for (int64_t rb = r0 + rl; rb < r0 + rows_per_block; rb += (int64_t)RL * kR) {
    Raw8<T> zr[kR], rr[kR];
#pragma unroll
    for (int u = 0; u < kR; ++u) {
      const int64_t r = rb + (int64_t)u * RL;
      if (r < r1) {
        zr[u].load(z + r * ld + c);
        if (res) rr[u].load(res + r * ld_res + c);
      }
    }
#pragma unroll
    for (int u = 0; u < kR; ++u) {
      const int64_t r = rb + (int64_t)u * RL;
      const bool valid = r < r1;
      float y[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) y[i] = 0.f;
      if (valid) {
        float zv[8];
        zr[u].unpack(zv);
        const uint32_t bits = C.drop ? dropout_keep8(seed, C.layer, C.row0 + r, (uint32_t)cgi, C.thr) : 0xffu;
        uint32_t gates = 0u;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float uu = zv[i];
          if (BN) uu = ((zv[i] - mean[i]) * rstd[i]) * gamma[i] + beta[i];
          float a = uu;
          if (ACT == EGNN_ACT_RELU) {
            const bool on = uu > 0.f;
            gates |= (on ? 1u : 0u) << i;
            a = on ? uu : 0.f;
          } else if (ACT == EGNN_ACT_ELU) {
            a = uu > 0.f ? uu : expm1f(uu);
          }
          y[i] = a * ((bits >> i) & 1u ? C.scale : 0.f);
        }
        if (keep_bits)   // one byte per 4 columns: low nibble = keep bits, high nibble = ReLU gates
          *reinterpret_cast<uint16_t*>(keep_bits + r * (2 * CG) + 2 * cgi) =
              (uint16_t)(((bits & 0xfu) | ((gates & 0xfu) << 4)) | ((((bits >> 4) & 0xfu) | (gates & 0xf0u)) << 8));
        if (res) {
          float rv[8];
          rr[u].unpack(rv);
#pragma unroll
          for (int i = 0; i < 8; ++i) y[i] += rv[i];
        }
        st8r(yout + r * ld_y + c, y);
      }
      if (PROJ) {
        float sp[kProjP];
#pragma unroll
        for (int p = 0; p < kProjP; ++p) {
          float a = 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i) a = fmaf(y[i], wp[PROJ ? p : 0][i], a);
          sp[p] = a;
        }
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {   // fixed butterfly over the CG lanes of the row
          if (o < CG) {
#pragma unroll
            for (int p = 0; p < kProjP; ++p) sp[p] += __shfl_xor_sync(0xffffffffu, sp[p], o);
          }
        }
        if (valid && cgi == 0) *reinterpret_cast<float4*>(pout + r * kProjP) = make_float4(sp[0], sp[1], sp[2], sp[3]);
      }
    }
  }
}

// partial[(blk*2 + which)*F + col] (double), which: 0 = sum g, 1 = sum g*xhat
template <typename T, bool BN, int ACT, bool KB>
__global__ void __launch_bounds__(kThreads, 4) bn_act_bwd_reduce_lean(const T* __restrict__ dy, int64_t ld_dy,
                                                                   const T* __restrict__ z, int64_t ld,
                                                                   int64_t n_rows, int F, int cg_shift,
                                                                   int64_t rows_per_block, ActCtx C,
                                                                   double* __restrict__ partial,
                                                                   const uint8_t* __restrict__ keep_bits) {
  __shared__ float sm[kThreads][9];
  const int CG = 1 << cg_shift, RL = kThreads >> cg_shift;
  const int cgi = threadIdx.x & (CG - 1), c = cgi * 4, rl = threadIdx.x >> cg_shift;
  Col4 k;
  load_col4<BN>(C, c, k);
  const uint64_t seed = C.seed + (C.seed_off ? (uint64_t)*C.seed_off : 0ull);
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block, r1 = min(n_rows, r0 + rows_per_block);
  float a0[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f};
  for (int64_t rb = r0 + rl; rb < r1; rb += (int64_t)RL * kRowsInFlight) {
    Raw4<T> zr[kRowsInFlight], gr[kRowsInFlight];
    uint32_t kb[kRowsInFlight];
#pragma unroll
    for (int u = 0; u < kRowsInFlight; ++u) {
      const int64_t r = rb + (int64_t)u * RL;
      if (r < r1) {
        zr[u].load(z + r * ld + c);
        gr[u].load(dy + r * ld_dy + c);
        kb[u] = KB ? keep_bits[r * CG + cgi] : 0xfu;
      }
    }
#pragma unroll
    for (int u = 0; u < kRowsInFlight; ++u) {
      const int64_t r = rb + (int64_t)u * RL;
      if (r < r1) {
        float zv[4], g[4], y[4], d[4], xh[4];
        zr[u].unpack(zv);
        gr[u].unpack(g);
        const uint32_t bits = (!KB && C.drop) ? keep_bits4(C, seed, r, c) : kb[u];
        act4<BN, ACT, KB && ACT == EGNN_ACT_RELU>(k, C.scale, bits, zv, y, d, xh);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float gg = g[i] * d[i];
          a0[i] += gg;
          a1[i] = fmaf(gg, xh[i], a1[i]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) { sm[threadIdx.x][i] = a0[i]; sm[threadIdx.x][4 + i] = a1[i]; }
  __syncthreads();
  for (int item = threadIdx.x; item < CG * 8; item += kThreads) {
    const int g_ = item >> 3, i = item & 7;
    double s = 0.0;
    for (int l = 0; l < RL; ++l) s += (double)sm[l * CG + g_][i];
    partial[((int64_t)blockIdx.x * 2 + (i >> 2)) * F + g_ * 4 + (i & 3)] = s;
  }
}

// dz = gamma*rstd*(g - mean(g) - xhat*mean(g*xhat)) (or g without BN); optionally the per-block
// column sums of the dz values written (the conv-bias gradient) -> dzsum_partial[blk*F + col]
template <typename T, bool BN, int ACT, bool KB>
__global__ void __launch_bounds__(kThreads, 4) bn_act_bwd_apply_lean(const T* __restrict__ dy, int64_t ld_dy,
                                                                  const T* __restrict__ z, int64_t ld_z,
                                                                  T* __restrict__ dz, int64_t ld, int64_t n_rows,
                                                                  int F, int cg_shift, int64_t rows_per_block,
                                                                  ActCtx C, const double* __restrict__ sum_g,
                                                                  const double* __restrict__ sum_gx, double inv_n,
                                                                  double* __restrict__ dzsum_partial,
                                                                  const uint8_t* __restrict__ keep_bits) {
  __shared__ float sm[kThreads][5];
  const int CG = 1 << cg_shift, RL = kThreads >> cg_shift;
  const int cgi = threadIdx.x & (CG - 1), c = cgi * 4, rl = threadIdx.x >> cg_shift;
  Col4 k;
  load_col4<BN>(C, c, k);
  float c1[4], c2[4], c3[4];  // dz = c1*gg - c2 - xhat*c3
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (BN) {
      const float gr = k.gamma[i] * k.rstd[i];
      c1[i] = gr;
      c2[i] = gr * (float)(sum_g[c + i] * inv_n);
      c3[i] = gr * (float)(sum_gx[c + i] * inv_n);
    } else {
      c1[i] = 1.f; c2[i] = 0.f; c3[i] = 0.f;
    }
  }
  const uint64_t seed = C.seed + (C.seed_off ? (uint64_t)*C.seed_off : 0ull);
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block, r1 = min(n_rows, r0 + rows_per_block);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int64_t rb = r0 + rl; rb < r1; rb += (int64_t)RL * kRowsInFlight) {
    Raw4<T> zr[kRowsInFlight], gr[kRowsInFlight];
    uint32_t kb[kRowsInFlight];
#pragma unroll
    for (int u = 0; u < kRowsInFlight; ++u) {
      const int64_t r = rb + (int64_t)u * RL;
      if (r < r1) {
        zr[u].load(z + r * ld_z + c);
        gr[u].load(dy + r * ld_dy + c);
        kb[u] = KB ? keep_bits[r * CG + cgi] : 0xfu;
      }
    }
#pragma unroll
    for (int u = 0; u < kRowsInFlight; ++u) {
      const int64_t r = rb + (int64_t)u * RL;
      if (r < r1) {
        float zv[4], g[4], y[4], d[4], xh[4], o[4];
        zr[u].unpack(zv);
        gr[u].unpack(g);
        const uint32_t bits = (!KB && C.drop) ? keep_bits4(C, seed, r, c) : kb[u];
        act4<BN, ACT, KB && ACT == EGNN_ACT_RELU>(k, C.scale, bits, zv, y, d, xh);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float gg = g[i] * d[i];
          o[i] = BN ? (c1[i] * gg - c2[i]) - xh[i] * c3[i] : gg;
        }
        st4f(dz + r * ld + c, o);
        if (dzsum_partial) {  // sum what the consumer will read (the value after rounding to T)
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[i] += to_f32(from_f32<T>(o[i]));
        }
      }
    }
  }
  if (dzsum_partial) {
#pragma unroll
    for (int i = 0; i < 4; ++i) sm[threadIdx.x][i] = acc[i];
    __syncthreads();
    for (int item = threadIdx.x; item < CG * 4; item += kThreads) {
      const int g_ = item >> 2, i = item & 3;
      double s = 0.0;
      for (int l = 0; l < RL; ++l) s += (double)sm[l * CG + g_][i];
      dzsum_partial[(int64_t)blockIdx.x * F + g_ * 4 + i] = s;
    }
  }
}

// ---- 8-column backward of BatchNorm + ReLU + dropout with the forward's saved keep / gate bits (the SAGE-ResBN hot
// path): 16-byte accesses.  DYP: the incoming gradient is not read but COMPUTED as dy[r, :] = dp[r, 0:4] . Wp -- the
// input gradient of the project-first logits layer (`SAGEConv(hidden, 2)`, src/models/gnn.py:128,193) -- rounded to T
// and written to dy_out for the later consumers (apply pass, residual path), which replaces a separate dgrad pass.
template <typename T, bool DYP>
__global__ void __launch_bounds__(kThreads, 2) bn_relu_bwd_reduce8(const T* __restrict__ dy, T* __restrict__ dy_out,
                                                                  const float* __restrict__ dp,
                                                                  const float* __restrict__ Wp,
                                                                  const T* __restrict__ z, int64_t ld_dy, int64_t ld_z,
                                                                  int64_t n_rows, int F, int cg_shift,
                                                                  int64_t rows_per_block, ActCtx C,
                                                                  double* __restrict__ partial,
                                                                  const uint8_t* __restrict__ keep_bits) {
  __shared__ float sm[kThreads][17];
  const int CG = 1 << cg_shift, RL = kThreads >> cg_shift;
  const int cgi = threadIdx.x & (CG - 1), c = cgi * 8, rl = threadIdx.x >> cg_shift;
  float mean[8], rstd[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { mean[i] = C.mean[c + i]; rstd[i] = C.rstd[c + i]; }
  float wp[DYP ? kProjP : 1][8];
  if (DYP) {
#pragma unroll
    for (int p = 0; p < kProjP; ++p)
#pragma unroll
      for (int i = 0; i < 8; ++i) wp[DYP ? p : 0][i] = Wp[p * F + c + i];
  }
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block, r1 = min(n_rows, r0 + rows_per_block);
  float a0[8], a1[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a0[i] = 0.f; a1[i] = 0.f; }
  for (int64_t rb = r0 + rl; rb < r1; rb += (int64_t)RL * kRows8) {
    Raw8<T> zr[kRows8], gr[kRows8];
    float4 dpr[DYP ? kRows8 : 1];
    uint32_t kb[kRows8];
#pragma unroll
    for (int u = 0; u < kRows8; ++u) {
      const int64_t r = rb + (int64_t)u * RL;
      if (r < r1) {
        zr[u].load(z + r * ld_z + c);
        if (DYP) dpr[DYP ? u : 0] = __ldg(reinterpret_cast<const float4*>(dp + r * kProjP));
        else gr[u].load(dy + r * ld_dy + c);
        kb[u] = *reinterpret_cast<const uint16_t*>(keep_bits + r * (2 * CG) + 2 * cgi);
      }
    }
#pragma unroll
    for (int u = 0; u < kRows8; ++u) {
      const int64_t r = rb + (int64_t)u * RL;
      if (r < r1) {
        float zv[8], g[8];
        zr[u].unpack(zv);
        if (DYP) {
          const float4 d = dpr[DYP ? u : 0];
#pragma unroll
          for (int i = 0; i < 8; ++i)
            g[i] = fmaf(d.w, wp[DYP ? 3 : 0][i], fmaf(d.z, wp[DYP ? 2 : 0][i], fmaf(d.y, wp[DYP ? 1 : 0][i], d.x * wp[0][i])));
          st8r(dy_out + r * ld_dy + c, g);   // g now holds the values as stored
        } else {
          gr[u].unpack(g);
        }
        const uint32_t keep = (kb[u] & 0xfu) | ((kb[u] >> 4) & 0xf0u), gate = ((kb[u] >> 4) & 0xfu) | ((kb[u] >> 8) & 0xf0u);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float xh = (zv[i] - mean[i]) * rstd[i];
          const float d = ((keep & gate) >> i) & 1u ? C.scale : 0.f;
          const float gg = g[i] * d;
          a0[i] += gg;
          a1[i] = fmaf(gg, xh, a1[i]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) { sm[threadIdx.x][i] = a0[i]; sm[threadIdx.x][8 + i] = a1[i]; }
  __syncthreads();
  for (int item = threadIdx.x; item < CG * 16; item += kThreads) {
    const int g_ = item >> 4, i = item & 15;
    double s = 0.0;
    for (int l = 0; l < RL; ++l) s += (double)sm[l * CG + g_][i];
    partial[((int64_t)blockIdx.x * 2 + (i >> 3)) * F + g_ * 8 + (i & 7)] = s;
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads, 2) bn_relu_bwd_apply8(const T* __restrict__ dy, int64_t ld_dy,
                                                                 const T* __restrict__ z, int64_t ld_z,
                                                                 T* __restrict__ dz, int64_t ld, int64_t n_rows, int F,
                                                                 int cg_shift, int64_t rows_per_block, ActCtx C,
                                                                 const double* __restrict__ sum_g,
                                                                 const double* __restrict__ sum_gx, double inv_n,
                                                                 double* __restrict__ dzsum_partial,
                                                                 const uint8_t* __restrict__ keep_bits) {
  __shared__ float sm[kThreads][9];
  const int CG = 1 << cg_shift, RL = kThreads >> cg_shift;
  const int cgi = threadIdx.x & (CG - 1), c = cgi * 8, rl = threadIdx.x >> cg_shift;
  float mean[8], rstd[8], c1[8], c2[8], c3[8];   // dz = c1*gg - c2 - xhat*c3
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    mean[i] = C.mean[c + i]; rstd[i] = C.rstd[c + i];
    const float gr = C.gamma[c + i] * rstd[i];
    c1[i] = gr;
    c2[i] = gr * (float)(sum_g[c + i] * inv_n);
    c3[i] = gr * (float)(sum_gx[c + i] * inv_n);
  }
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block, r1 = min(n_rows, r0 + rows_per_block);
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  for (int64_t rb = r0 + rl; rb < r1; rb += (int64_t)RL * kRows8) {
    Raw8<T> zr[kRows8], gr[kRows8];
    uint32_t kb[kRows8];
#pragma unroll
    for (int u = 0; u < kRows8; ++u) {
      const int64_t r = rb + (int64_t)u * RL;
      if (r < r1) {
        zr[u].load(z + r * ld_z + c);
        gr[u].load(dy + r * ld_dy + c);
        kb[u] = *reinterpret_cast<const uint16_t*>(keep_bits + r * (2 * CG) + 2 * cgi);
      }
    }
#pragma unroll
    for (int u = 0; u < kRows8; ++u) {
      const int64_t r = rb + (int64_t)u * RL;
      if (r < r1) {
        float zv[8], g[8], o[8];
        zr[u].unpack(zv);
        gr[u].unpack(g);
        const uint32_t keep = (kb[u] & 0xfu) | ((kb[u] >> 4) & 0xf0u), gate = ((kb[u] >> 4) & 0xfu) | ((kb[u] >> 8) & 0xf0u);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float xh = (zv[i] - mean[i]) * rstd[i];
          const float d = ((keep & gate) >> i) & 1u ? C.scale : 0.f;
          const float gg = g[i] * d;
          o[i] = (c1[i] * gg - c2[i]) - xh * c3[i];
        }
        st8r(dz + r * ld + c, o);   // o now holds the values as stored
        if (dzsum_partial) {
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] += o[i];
        }
      }
    }
  }
  if (dzsum_partial) {
#pragma unroll
    for (int i = 0; i < 8; ++i) sm[threadIdx.x][i] = acc[i];
    __syncthreads();
    for (int item = threadIdx.x; item < CG * 8; item += kThreads) {
      const int g_ = item >> 3, i = item & 7;
      double s = 0.0;
      for (int l = 0; l < RL; ++l) s += (double)sm[l * CG + g_][i];
      dzsum_partial[(int64_t)blockIdx.x * F + g_ * 8 + i] = s;
    }
  }
}

// host-side dispatch over (dtype, BatchNorm, activation); the backward kernels are also specialised on whether
// the forward's keep bits are supplied (the Philox recomputation is then compiled out: fewer registers)
#define EGNN_LEAN_KB_T(KERNEL, TT_, HAVE_KB, GRID, ...)                                                    \
  {                                                                                                        \
    using TT = TT_;                                                                                        \
    const int sel_ = (C.mean != nullptr ? 0 : 6) +                                                         \
                     (C.act == EGNN_ACT_RELU ? 0 : C.act == EGNN_ACT_NONE ? 2 : 4) + ((HAVE_KB) ? 0 : 1);  \
    switch (sel_) {                                                                                        \
      case 0: KERNEL<TT, true, EGNN_ACT_RELU, true><<<GRID, kThreads, 0, st>>>(__VA_ARGS__); break;        \
      case 1: KERNEL<TT, true, EGNN_ACT_RELU, false><<<GRID, kThreads, 0, st>>>(__VA_ARGS__); break;       \
      case 2: KERNEL<TT, true, EGNN_ACT_NONE, true><<<GRID, kThreads, 0, st>>>(__VA_ARGS__); break;        \
      case 3: KERNEL<TT, true, EGNN_ACT_NONE, false><<<GRID, kThreads, 0, st>>>(__VA_ARGS__); break;       \
      case 4: KERNEL<TT, true, EGNN_ACT_ELU, true><<<GRID, kThreads, 0, st>>>(__VA_ARGS__); break;         \
      case 5: KERNEL<TT, true, EGNN_ACT_ELU, false><<<GRID, kThreads, 0, st>>>(__VA_ARGS__); break;        \
      case 6: KERNEL<TT, false, EGNN_ACT_RELU, true><<<GRID, kThreads, 0, st>>>(__VA_ARGS__); break;       \
      case 7: KERNEL<TT, false, EGNN_ACT_RELU, false><<<GRID, kThreads, 0, st>>>(__VA_ARGS__); break;      \
      case 8: KERNEL<TT, false, EGNN_ACT_NONE, true><<<GRID, kThreads, 0, st>>>(__VA_ARGS__); break;       \
      case 9: KERNEL<TT, false, EGNN_ACT_NONE, false><<<GRID, kThreads, 0, st>>>(__VA_ARGS__); break;      \
      case 10: KERNEL<TT, false, EGNN_ACT_ELU, true><<<GRID, kThreads, 0, st>>>(__VA_ARGS__); break;       \
      default: KERNEL<TT, false, EGNN_ACT_ELU, false><<<GRID, kThreads, 0, st>>>(__VA_ARGS__); break;      \
    }                                                                                                      \
  }
#define EGNN_LEAN_DISPATCH_KB(KERNEL, HAVE_KB, GRID, ...)                                  \
  do {                                                                                     \
    if (dtype == EGNN_F32) EGNN_LEAN_KB_T(KERNEL, float, HAVE_KB, GRID, __VA_ARGS__)       \
    else EGNN_LEAN_KB_T(KERNEL, __nv_bfloat16, HAVE_KB, GRID, __VA_ARGS__)                 \
  } while (0)

// host-side dispatch over (dtype, BatchNorm, activation)
#define EGNN_LEAN_DISPATCH(KERNEL, GRID, ...)                                                            \
  do {                                                                                                   \
    const bool bn_ = C.mean != nullptr;                                                                  \
    if (dtype == EGNN_F32) {                                                                             \
      using TT = float;                                                                                  \
      if (bn_ && C.act == EGNN_ACT_RELU) KERNEL<TT, true, EGNN_ACT_RELU><<<GRID, kThreads, 0, st>>>(__VA_ARGS__);   \
      else if (bn_ && C.act == EGNN_ACT_NONE) KERNEL<TT, true, EGNN_ACT_NONE><<<GRID, kThreads, 0, st>>>(__VA_ARGS__); \
      else if (bn_) KERNEL<TT, true, EGNN_ACT_ELU><<<GRID, kThreads, 0, st>>>(__VA_ARGS__);             \
      else if (C.act == EGNN_ACT_RELU) KERNEL<TT, false, EGNN_ACT_RELU><<<GRID, kThreads, 0, st>>>(__VA_ARGS__);  \
      else if (C.act == EGNN_ACT_NONE) KERNEL<TT, false, EGNN_ACT_NONE><<<GRID, kThreads, 0, st>>>(__VA_ARGS__);  \
      else KERNEL<TT, false, EGNN_ACT_ELU><<<GRID, kThreads, 0, st>>>(__VA_ARGS__);                     \
    } else {                                                                                             \
      using TT = __nv_bfloat16;                                                                          \
      if (bn_ && C.act == EGNN_ACT_RELU) KERNEL<TT, true, EGNN_ACT_RELU><<<GRID, kThreads, 0, st>>>(__VA_ARGS__);   \
      else if (bn_ && C.act == EGNN_ACT_NONE) KERNEL<TT, true, EGNN_ACT_NONE><<<GRID, kThreads, 0, st>>>(__VA_ARGS__); \
      else if (bn_) KERNEL<TT, true, EGNN_ACT_ELU><<<GRID, kThreads, 0, st>>>(__VA_ARGS__);             \
      else if (C.act == EGNN_ACT_RELU) KERNEL<TT, false, EGNN_ACT_RELU><<<GRID, kThreads, 0, st>>>(__VA_ARGS__);  \
      else if (C.act == EGNN_ACT_NONE) KERNEL<TT, false, EGNN_ACT_NONE><<<GRID, kThreads, 0, st>>>(__VA_ARGS__);  \
      else KERNEL<TT, false, EGNN_ACT_ELU><<<GRID, kThreads, 0, st>>>(__VA_ARGS__);                     \
    }                                                                                                    \
  } while (0)

// column statistics of a matrix (BatchNorm forward): partial[(blk*2+which)*F + col]
template <typename T>
__global__ void __launch_bounds__(kThreads) colstats_fast(const T* __restrict__ a, int64_t ld, int64_t n_rows,
                                                          int F, int cg_shift, int64_t rows_per_block,
                                                          double* __restrict__ partial) {
  __shared__ float sm[kThreads][17];
  const int CG = 1 << cg_shift, RL = kThreads >> cg_shift;
  const int cgi = threadIdx.x & (CG - 1), c = cgi * 8, rl = threadIdx.x >> cg_shift;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block, r1 = min(n_rows, r0 + rows_per_block);
  float a0[8], a1[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a0[i] = a1[i] = 0.f;
#pragma unroll 4
  for (int64_t r = r0 + rl; r < r1; r += RL) {
    float v[8];
    ld8f<T>(a + r * ld + c, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) { a0[i] += v[i]; a1[i] = fmaf(v[i], v[i], a1[i]); }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) { sm[threadIdx.x][i] = a0[i]; sm[threadIdx.x][8 + i] = a1[i]; }
  __syncthreads();
  for (int item = threadIdx.x; item < CG * 16; item += kThreads) {
    const int g_ = item >> 4, i = item & 15;
    double s = 0.0;
    for (int l = 0; l < RL; ++l) s += (double)sm[l * CG + g_][i];
    partial[((int64_t)blockIdx.x * 2 + (i >> 3)) * F + g_ * 8 + (i & 7)] = s;
  }
}

constexpr int kFastMaxF = 2048;
struct FastPlan {
  bool ok;
  int cg_shift;
  int64_t rpb;
  int nblk;
};
// rows per block: a multiple of the row lanes, <= kMaxPartBlocks blocks, ~8 rows per thread
inline FastPlan fast_plan(int64_t n_rows, int64_t F, int64_t ld, int dtype, std::initializer_list<const void*> ptrs) {
  FastPlan p{false, 0, 0, 0};
  if (F % 8 != 0 || ld % 8 != 0 || F > kFastMaxF) return p;
  int cg = (int)(F / 8), sh = 0;
  while ((1 << sh) < cg) ++sh;
  if ((1 << sh) != cg) return p;
  const size_t es = dtype == EGNN_F32 ? 4 : 2;
  for (const void* q : ptrs)
    if (q && (uintptr_t)q % (8 * es) != 0) return p;
  const int RL = kThreads >> sh;
  int64_t rpb = (int64_t)RL * 8;
  if (ceil_div(n_rows, rpb) > kMaxPartBlocks) rpb = ceil_div(ceil_div(n_rows, kMaxPartBlocks), RL) * RL;
  p.ok = true; p.cg_shift = sh; p.rpb = rpb; p.nblk = (int)ceil_div(n_rows > 0 ? n_rows : 1, rpb);
  return p;
}

// plan of the 4-column lean kernels: F/4 a power of two <= 256, 4-element aligned pointers / strides
inline FastPlan lean_plan(int64_t n_rows, int64_t F, int dtype, std::initializer_list<int64_t> lds,
                          std::initializer_list<const void*> ptrs) {
  FastPlan p{false, 0, 0, 0};
  if (F % 4 != 0 || F > 4 * kThreads) return p;
  int cg = (int)(F / 4), sh = 0;
  while ((1 << sh) < cg) ++sh;
  if ((1 << sh) != cg) return p;
  const size_t es = dtype == EGNN_F32 ? 4 : 2;
  for (int64_t l : lds)
    if (l % 4 != 0) return p;
  for (const void* q : ptrs)
    if (q && (uintptr_t)q % (4 * es) != 0) return p;
  const int RL = kThreads >> sh;
  // at most one wave of blocks at 3 resident blocks per SM (no partially filled second wave)
  const int64_t unit = (int64_t)RL * kRowsInFlight;
  int64_t rpb = ceil_div(ceil_div(n_rows > 0 ? n_rows : 1, (int64_t)kNumSMs * 3), unit) * unit;
  if (rpb < unit * 2) rpb = unit * 2;
  p.ok = true; p.cg_shift = sh; p.rpb = rpb; p.nblk = (int)ceil_div(n_rows > 0 ? n_rows : 1, rpb);
  return p;
}

inline FastPlan plan8(int64_t n_rows, int64_t F, int dtype, std::initializer_list<int64_t> lds,
                      std::initializer_list<const void*> ptrs, int blocks_per_sm = 2) {
  FastPlan p{false, 0, 0, 0};
  if (F % 8 != 0 || F > 256) return p;
  int cg = (int)(F / 8), sh = 0;
  while ((1 << sh) < cg) ++sh;
  if ((1 << sh) != cg) return p;
  for (int64_t l : lds)
    if (l % 8 != 0) return p;
  (void)dtype;
  for (const void* q : ptrs)
    if (q && (uintptr_t)q % 16 != 0) return p;
  const int RL = kThreads >> sh;
  const int64_t unit = (int64_t)RL * kRows8;
  int64_t rpb = ceil_div(ceil_div(n_rows > 0 ? n_rows : 1, (int64_t)kNumSMs * blocks_per_sm), unit) * unit;   // one wave
  if (rpb < unit) rpb = unit;
  p.ok = true; p.cg_shift = sh; p.rpb = rpb; p.nblk = (int)ceil_div(n_rows > 0 ? n_rows : 1, rpb);
  return p;
}

// one block per column: fixed-order sum of nblk doubles
__global__ void __launch_bounds__(kThreads) colsum_final(const double* __restrict__ partial, int nblk, int F,
                                                         float* __restrict__ out) {
  __shared__ double sm[kThreads];
  const int col = blockIdx.x;
  double s = 0;
  for (int b = threadIdx.x; b < nblk; b += kThreads) s += partial[(int64_t)b * F + col];
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int o = kThreads / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[col] = (float)sm[0];
}

__global__ void bn_finalize_kernel(const double* __restrict__ sums, const double* __restrict__ sumsq,
                                   double count, int F, float eps, float momentum,
                                   float* __restrict__ mean, float* __restrict__ rstd,
                                   float* __restrict__ rmean, float* __restrict__ rvar) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= F) return;
  double m = sums[c] / count;
  double var = sumsq[c] / count - m * m;
  if (var < 0) var = 0;
  mean[c] = (float)m;
  rstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (rmean) rmean[c] = (1.f - momentum) * rmean[c] + momentum * (float)m;
  if (rvar) {
    double unb = count > 1 ? var * count / (count - 1) : var;
    rvar[c] = (1.f - momentum) * rvar[c] + momentum * (float)unb;
  }
}

// BatchNorm statistics from the per-CTA partial rows the layer GEMM's epilogue wrote
// (egnn_linear_tc colstats): fixed-order float64 sum over the parts; FINALIZE also turns them into mean / rstd and
// updates the running buffers (single-GPU: one launch between the GEMM and the BatchNorm apply).
constexpr int kPartCols = 64;     // columns per block
constexpr int kPartGroups = 16;   // part groups per block (threads = 64 x 16)
template <bool FINALIZE>
__global__ void __launch_bounds__(kPartCols * kPartGroups)
colstats_parts_kernel(const float* __restrict__ parts, int n_parts, int F, double* __restrict__ sums, double count,
                      float eps, float momentum, float* __restrict__ mean, float* __restrict__ rstd,
                      float* __restrict__ rmean, float* __restrict__ rvar, int64_t* __restrict__ num_batches) {
  __shared__ double sm[2][kPartGroups][kPartCols];
  const int cl = threadIdx.x & (kPartCols - 1), gq = threadIdx.x / kPartCols;
  const int c = blockIdx.x * kPartCols + cl;
  if (FINALIZE && num_batches && blockIdx.x == 0 && threadIdx.x == 0) *num_batches += 1;   // bn.num_batches_tracked
  double s = 0.0, q = 0.0;
  if (c < F) {
    // group gq sums parts gq, gq + 16, ... in that order: every (group, column) chain is fixed, and so is the
    // order in which the 16 group totals are combined below
    int p = gq;
    for (; p + 3 * kPartGroups < n_parts; p += 4 * kPartGroups) {   // four parts in flight, same summation order
      float a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        a[u] = parts[(size_t)(p + u * kPartGroups) * 2 * F + c];
        b[u] = parts[(size_t)(p + u * kPartGroups) * 2 * F + F + c];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) { s += (double)a[u]; q += (double)b[u]; }
    }
    for (; p < n_parts; p += kPartGroups) {
      s += (double)parts[(size_t)p * 2 * F + c];
      q += (double)parts[(size_t)p * 2 * F + F + c];
    }
  }
  sm[0][gq][cl] = s;
  sm[1][gq][cl] = q;
  __syncthreads();
  if (gq != 0 || c >= F) return;
  s = 0.0;
  q = 0.0;
#pragma unroll
  for (int k = 0; k < kPartGroups; ++k) {
    s += sm[0][k][cl];
    q += sm[1][k][cl];
  }
  if (sums) {
    sums[c] = s;
    sums[F + c] = q;
  }
  if (FINALIZE) {
    const double m = s / count;
    double var = q / count - m * m;
    if (var < 0) var = 0;
    mean[c] = (float)m;
    rstd[c] = (float)(1.0 / sqrt(var + (double)eps));
    if (rmean) rmean[c] = (1.f - momentum) * rmean[c] + momentum * (float)m;
    if (rvar) {
      const double unb = count > 1 ? var * count / (count - 1) : var;
      rvar[c] = (1.f - momentum) * rvar[c] + momentum * (float)unb;
    }
  }
}

__global__ void __launch_bounds__(kThreads) dropout_mask_kernel(uint8_t* __restrict__ mask,
                                                                int64_t n_rows, int F, uint32_t thr,
                                                                uint64_t seed, const int64_t* seed_off,
                                                                uint32_t layer, int64_t row0) {
  const int F8n = (F + 7) >> 3;
  int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n_rows * F8n) return;
  int64_t r = i / F8n;
  int cb = (int)(i - r * F8n);
  const uint32_t keep = dropout_keep8(seed + (seed_off ? (uint64_t)*seed_off : 0ull), layer, row0 + r, (uint32_t)cb, thr);
#pragma unroll
  for (int k = 0; k < 8; ++k)
    if (cb * 8 + k < F) mask[r * F + cb * 8 + k] = (keep >> k) & 1u;
}

// ---- masked loss over the train rows (2 classes): `_make_loss_fn` (src/train_gnn.py:136-183) ----------------------
//   plain   : w[y] * CE                                  (F.cross_entropy(weight=cw, reduction='none'), :159-162)
//   focal   : (1 - p_y)^gamma * CE, NO class weights     (:153-158)
//   time    : * clamp(f((t - t_min) / max(t_max - t_min, 1)), 1e-3), f = identity | sqrt(clamp(., 0))   (:165-174)
//   then .mean() over the train rows (inv_n = 1 / global train-row count).  d loss / d logits is produced in the same
//   pass; for the focal term  d/dz_k = [gamma (1-p)^(gamma-1) p log p - (1-p)^gamma] (delta_ky - p_k).
template <typename T>
__global__ void __launch_bounds__(kThreads) masked_ce_kernel(const T* __restrict__ logits,
                                                             const int64_t* __restrict__ y,
                                                             const int64_t* __restrict__ idx,
                                                             int64_t n_idx, const float* __restrict__ cw,
                                                             float inv_n, float focal_gamma,
                                                             const int64_t* __restrict__ timestep, float t_min,
                                                             float t_denom, int time_scheme,
                                                             T* __restrict__ dlogits,
                                                             float* __restrict__ partial) {
  __shared__ float sm[kThreads];
  int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  float li = 0.f;
  if (i < n_idx) {
    int64_t r = idx ? idx[i] : i;
    float l0 = to_f32(logits[2 * r]), l1 = to_f32(logits[2 * r + 1]);
    int64_t yi = y[r];
    float m = fmaxf(l0, l1);
    float e0 = expf(l0 - m), e1 = expf(l1 - m);
    float s = e0 + e1;
    float lse = m + logf(s);
    float ce = lse - (yi == 0 ? l0 : l1);
    float p0 = e0 / s, p1 = e1 / s;
    float c;      // d li / d z_k = c * (p_k - delta_ky)
    if (focal_gamma >= 0.f) {
      float pt = yi == 0 ? p0 : p1;
      float om = 1.f - pt;
      float mod = powf(om, focal_gamma);
      li = mod * ce;
      // (1-p)^g - g (1-p)^(g-1) p log p, with log p = -ce; (1-p)^(g-1) p taken as 0 at p = 1
      float dm = focal_gamma > 0.f ? focal_gamma * (om > 0.f ? powf(om, focal_gamma - 1.f) : (focal_gamma == 1.f ? 1.f : 0.f)) : 0.f;
      c = mod + dm * pt * ce;
    } else {
      float w = cw[yi];
      li = w * ce;
      c = w;
    }
    if (time_scheme != 0) {
      float wt = __fdiv_rn((float)timestep[r] - t_min, t_denom);
      if (time_scheme == 2) wt = sqrtf(fmaxf(wt, 0.f));
      wt = fmaxf(wt, 1e-3f);
      li *= wt;
      c *= wt;
    }
    c *= inv_n;
    dlogits[2 * r] = from_f32<T>(c * (p0 - (yi == 0 ? 1.f : 0.f)));
    dlogits[2 * r + 1] = from_f32<T>(c * (p1 - (yi == 1 ? 1.f : 0.f)));
  }
  sm[threadIdx.x] = li;
  __syncthreads();
  for (int o = kThreads / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sm[0];
}

// loss += lambda * mean(w^2); grad += 2 lambda / n * w   (optional L2 on the learned time table, :178-180)
__global__ void __launch_bounds__(kThreads) l2_mean_penalty_kernel(const float* __restrict__ w, int64_t n, float lambda,
                                                                   float* __restrict__ loss, float* __restrict__ grad) {
  __shared__ double sm[kThreads];
  double s = 0;
  const float gs = 2.f * lambda / (float)n;
  for (int64_t i = threadIdx.x; i < n; i += kThreads) {
    const float v = w[i];
    s += (double)v * (double)v;
    if (grad) grad[i] += gs * v;
  }
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int o = kThreads / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss += lambda * (float)(sm[0] / (double)n);
}

__global__ void __launch_bounds__(kThreads) ce_final_kernel(const float* __restrict__ partial, int nblk,
                                                            double inv_n, float* __restrict__ loss) {
  __shared__ double sm[kThreads];
  double s = 0;
  for (int b = threadIdx.x; b < nblk; b += kThreads) s += (double)partial[b];
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int o = kThreads / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss = (float)(sm[0] * inv_n);
}

// ---- global-norm clip + Adam over a flat buffer ---------------------------------------------------
__global__ void __launch_bounds__(kThreads) sqnorm_partial_kernel(const float* __restrict__ g, int64_t n,
                                                                  float* __restrict__ partial) {
  __shared__ float sm[kThreads];
  float s = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads)
    s += g[i] * g[i];
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int o = kThreads / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sm[0];
}

// ws layout: [0]=clip coef, [1]=bias_correction1, [2]=sqrt(bias_correction2), [3]=grad norm
__global__ void __launch_bounds__(kThreads) adam_prepare_kernel(const float* __restrict__ partial, int nblk,
                                                                float max_norm, float beta1, float beta2,
                                                                int64_t* __restrict__ step,
                                                                float* __restrict__ coefs,
                                                                float* __restrict__ norm_out) {
  __shared__ double sm[kThreads];
  double s = 0;
  for (int b = threadIdx.x; b < nblk; b += kThreads) s += (double)partial[b];
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int o = kThreads / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    float norm = (float)sqrt(sm[0]);
    float coef = 1.f;
    if (max_norm > 0.f) {
      coef = max_norm / (norm + 1e-6f);
      coef = coef > 1.f ? 1.f : coef;
    }
    int64_t t = *step + 1;
    *step = t;
    coefs[0] = coef;
    coefs[1] = (float)(1.0 - pow((double)beta1, (double)t));
    coefs[2] = (float)sqrt(1.0 - pow((double)beta2, (double)t));
    coefs[3] = norm;
    if (norm_out) *norm_out = norm;
  }
}

__global__ void __launch_bounds__(kThreads) adam_apply_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                              float* __restrict__ m, float* __restrict__ v,
                                                              int64_t n, float lr, float beta1, float beta2,
                                                              float eps, float wd,
                                                              const float* __restrict__ coefs) {
  int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  const float coef = coefs[0], bc1 = coefs[1], bc2s = coefs[2];
  float gi = g[i] * coef;
  float pi = p[i];
  gi = gi + wd * pi;                       // coupled L2 (torch.optim.Adam, not AdamW)
  float mi = m[i] + (gi - m[i]) * (1.f - beta1);  // lerp, as torch does
  float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
  float denom = sqrtf(vi) / bc2s + eps;
  p[i] = pi - (lr / bc1) * (mi / denom);
  m[i] = mi;
  v[i] = vi;
}

}  // namespace
}  // namespace egnn

using namespace egnn;

extern "C" int egnn_cast(const void* in, int in_dtype, int64_t ld_in, void* out, int out_dtype,
                         int64_t ld_out, int64_t n_rows, int64_t n_feat, void* stream) {
  const char* fn = "egnn_cast";
  EGNN_REQUIRE(in && out, fn, "null pointer");
  EGNN_REQUIRE(ld_in >= n_feat && ld_out >= n_feat && n_feat > 0, fn, "bad shape");
  if (n_rows == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int Fo4 = (int)ceil_div(ld_out, 4);
  // the output row is written over its whole leading dimension (zero padding); vector stores
  // need ld_out % 4 == 0, else fall back to element stores bounded by ld_out
  bool vin = vec_ok(in, in_dtype, ld_in, 4) && (ld_in % 4 == 0);
  bool vout = vec_ok(out, out_dtype, ld_out, 4);
  unsigned grid = (unsigned)ceil_div(n_rows * Fo4, kThreads);
#define EGNN_CAST(TI, TO)                                                                          \
  cast_kernel<TI, TO><<<grid, kThreads, 0, st>>>((const TI*)in, ld_in, (TO*)out, ld_out, n_rows, \
                                                 (int)n_feat, Fo4, vin, vout)
  if (in_dtype == EGNN_F32 && out_dtype == EGNN_F32) EGNN_CAST(float, float);
  else if (in_dtype == EGNN_F32 && out_dtype == EGNN_BF16) EGNN_CAST(float, __nv_bfloat16);
  else if (in_dtype == EGNN_BF16 && out_dtype == EGNN_F32) EGNN_CAST(__nv_bfloat16, float);
  else if (in_dtype == EGNN_BF16 && out_dtype == EGNN_BF16) EGNN_CAST(__nv_bfloat16, __nv_bfloat16);
  else return fail(fn, "unsupported dtype");
#undef EGNN_CAST
  EGNN_LAUNCH_CHECK(fn);
  return 0;
}

// out[r, :] (bf16, [No + Nr, 2*Kp]) = r < No ? [w_l[r] | 0 | w_r[r] | 0] : [0 | w_res[r - No] | 0];  bias_out = [b_l | 0]
template <typename TO>
__global__ void __launch_bounds__(kThreads) pack_sage_weights_kernel(
    const float* __restrict__ w_l, const float* __restrict__ w_r, const float* __restrict__ w_res,
    const float* __restrict__ b_l, int No, int Nr, int K, int Kp, TO* __restrict__ out,
    float* __restrict__ bias_out, TO* __restrict__ out_t) {
  const int i = blockIdx.x * kThreads + threadIdx.x;
  const int W = 2 * Kp;
  if (i < No + Nr && bias_out) bias_out[i] = (i < No && b_l) ? b_l[i] : 0.f;
  if (i >= (No + Nr) * W) return;
  const int r = i / W, c = i - r * W;
  const int half = c >= Kp, k = c - half * Kp;
  float v = 0.f;
  if (k < K) {
    if (r < No) v = half ? w_r[r * K + k] : w_l[r * K + k];
    else if (half) v = w_res[(r - No) * K + k];
  }
  out[i] = from_f32<TO>(v);
  // [W_l | W_r]^T  ([2*Kp, No]): the contraction-contiguous B operand of the concatenated dgrad GEMM
  if (out_t && r < No) out_t[(size_t)c * No + r] = from_f32<TO>(v);
}

extern "C" int egnn_pack_sage_weights(const float* w_l, const float* w_r, const float* w_res, const float* b_l,
                                      int64_t n_out, int64_t n_res, int64_t K, int64_t K_padded, void* out_bf16,
                                      float* bias_out, void* out_t_bf16, int out_dtype, void* stream) {
  const char* fn = "egnn_pack_sage_weights";
  EGNN_REQUIRE(w_l && w_r && out_bf16 && (n_res == 0 || w_res), fn, "null pointer");
  EGNN_REQUIRE(out_dtype == EGNN_BF16 || out_dtype == EGNN_F32, fn, "bad output dtype");
  EGNN_REQUIRE(n_out > 0 && K > 0 && K_padded >= K && (n_out + n_res) * 2 * K_padded < (1 << 30), fn, "bad shape");
  const int64_t total = (n_out + n_res) * 2 * K_padded;
  if (out_dtype == EGNN_BF16)
    pack_sage_weights_kernel<__nv_bfloat16><<<(unsigned)ceil_div(total, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        w_l, w_r, w_res, b_l, (int)n_out, (int)n_res, (int)K, (int)K_padded, (__nv_bfloat16*)out_bf16, bias_out,
        (__nv_bfloat16*)out_t_bf16);
  else
    pack_sage_weights_kernel<float><<<(unsigned)ceil_div(total, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        w_l, w_r, w_res, b_l, (int)n_out, (int)n_res, (int)K, (int)K_padded, (float*)out_bf16, bias_out,
        (float*)out_t_bf16);
  EGNN_LAUNCH_CHECK(fn);
  return 0;
}

extern "C" int egnn_inject_time(const float* x, int64_t ld_x, const int64_t* t, const float* table,
                                int64_t T, int64_t D, float* out_f32, void* out_bf16, int64_t ld_out,
                                int64_t ld_out_bf16, int64_t n_rows, int64_t n_feat, void* stream) {
  const char* fn = "egnn_inject_time";
  if (ld_out_bf16 <= 0) ld_out_bf16 = ld_out;
  EGNN_REQUIRE(ld_out_bf16 % 4 == 0 && ld_out_bf16 >= ld_out, fn, "bad ld_out_bf16");
  EGNN_REQUIRE((!out_f32 || (uintptr_t)out_f32 % 16 == 0) && (!out_bf16 || (uintptr_t)out_bf16 % 8 == 0) &&
                   ld_out / 4 <= kThreads,
               fn, "outputs must be 16-byte (fp32) / 8-byte (bf16) aligned, ld_out <= 1024");
  EGNN_REQUIRE(x && (out_f32 || out_bf16), fn, "null pointer");
  EGNN_REQUIRE(D == 0 || (t && table && T > 0), fn, "time table / indices missing");
  EGNN_REQUIRE(ld_out % 4 == 0 && ld_out >= n_feat + D, fn, "ld_out must be a multiple of 4 and >= F+D");
  if (n_rows == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  unsigned grid = (unsigned)ceil_div(n_rows, kInjectRows);
  inject_time_kernel<<<grid, kThreads, 0, st>>>(x, ld_x, t, table, T, (int)D, out_f32,
                                                (__nv_bfloat16*)out_bf16, ld_out, ld_out_bf16, n_rows,
                                                (int)n_feat);
  EGNN_LAUNCH_CHECK(fn);
  return 0;
}

extern "C" size_t egnn_embed_grad_workspace_bytes(int64_t n_rows, int64_t T, int64_t D) {
  return (size_t)T * (size_t)ceil_div(n_rows > 0 ? n_rows : 1, kEmbChunk) * (size_t)D * sizeof(double);
}

extern "C" int egnn_embed_grad(const void* dout, int dtype, int64_t ld, int64_t col0, int64_t D, const int64_t* t,
                               int64_t T, int64_t n_rows, float* dtab, void* workspace, void* stream) {
  const char* fn = "egnn_embed_grad";
  EGNN_REQUIRE(dout && t && dtab && workspace, fn, "null pointer");
  EGNN_REQUIRE(T > 0 && T <= 65535 && D > 0 && col0 >= 0 && ld >= col0 + D && n_rows >= 0, fn, "bad shape");
  EGNN_REQUIRE(dtype == EGNN_F32 || dtype == EGNN_BF16, fn, "dtype must be EGNN_F32 or EGNN_BF16");
  cudaStream_t st = (cudaStream_t)stream;
  const int nchunks = (int)ceil_div(n_rows > 0 ? n_rows : 1, kEmbChunk);
  double* partial = reinterpret_cast<double*>(workspace);
  dim3 grid((unsigned)T, (unsigned)nchunks);
  if (dtype == EGNN_F32)
    embed_grad_partial<float><<<grid, kThreads, 0, st>>>((const float*)dout, ld, (int)col0, (int)D, t, T, n_rows, partial);
  else
    embed_grad_partial<__nv_bfloat16><<<grid, kThreads, 0, st>>>((const __nv_bfloat16*)dout, ld, (int)col0, (int)D, t, T,
                                                               n_rows, partial);
  EGNN_LAUNCH_CHECK(fn);
  embed_grad_final<<<(unsigned)ceil_div(T * D, 128), 128, 0, st>>>(partial, nchunks, (int)D, T, dtab);
  EGNN_LAUNCH_CHECK(fn);
  return 0;
}

extern "C" size_t egnn_colreduce_workspace_bytes(int64_t n_feat) {
  return (size_t)(kMaxPartBlocks + 8) * 2 * (size_t)(((n_feat + 3) / 4) * 4) * sizeof(double);
}

extern "C" int egnn_colreduce(const void* a, int dtype, int64_t ld, int64_t n_rows, int64_t n_feat,
                              double* sums, double* sumsq, void* workspace, void* stream) {
  const char* fn = "egnn_colreduce";
  EGNN_REQUIRE(a && sums && workspace, fn, "null pointer");
  EGNN_REQUIRE(n_feat > 0 && ld >= n_feat, fn, "bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  FastPlan fp = fast_plan(n_rows, n_feat, ld, dtype, {a});
  if (fp.ok) {
    double* partial = reinterpret_cast<double*>(workspace);
    if (dtype == EGNN_F32)
      colstats_fast<float><<<fp.nblk, kThreads, 0, st>>>((const float*)a, ld, n_rows, (int)n_feat, fp.cg_shift,
                                                         fp.rpb, partial);
    else
      colstats_fast<__nv_bfloat16><<<fp.nblk, kThreads, 0, st>>>((const __nv_bfloat16*)a, ld, n_rows, (int)n_feat,
                                                                 fp.cg_shift, fp.rpb, partial);
    EGNN_LAUNCH_CHECK(fn);
    colreduce_final<<<dim3((unsigned)n_feat, 2), kThreads, 0, st>>>(partial, fp.nblk, (int)n_feat, sums, sumsq);
    EGNN_LAUNCH_CHECK(fn);
    return 0;
  }
  bool v = vec_ok(a, dtype, ld, n_feat);
  if (dtype == EGNN_F32)
    return run_colreduce(PlainProd<float>{(const float*)a, ld, (int)n_feat, v}, n_rows, (int)n_feat, sums,
                         sumsq, workspace, st, fn);
  return run_colreduce(PlainProd<__nv_bfloat16>{(const __nv_bfloat16*)a, ld, (int)n_feat, v}, n_rows,
                       (int)n_feat, sums, sumsq, workspace, st, fn);
}

extern "C" int egnn_gat_att_grad(const float* xs, const float* da_s, const float* da_d, int64_t n_rows, int H,
                                 int C, double* datt_src, double* datt_dst, void* workspace, void* stream) {
  const char* fn = "egnn_gat_att_grad";
  EGNN_REQUIRE(xs && da_s && da_d && datt_src && datt_dst && workspace && H > 0 && C > 0, fn, "bad arguments");
  int F = H * C;
  bool v = vec_ok(xs, EGNN_F32, F, F);
  return run_colreduce(GatAttProd{xs, da_s, da_d, H, C, F, v}, n_rows, F, datt_src, datt_dst, workspace,
                       (cudaStream_t)stream, fn);
}

extern "C" int egnn_bn_finalize(const double* sums, const double* sumsq, double count, int64_t n_feat,
                                float eps, float momentum, float* mean, float* rstd, float* running_mean,
                                float* running_var, void* stream) {
  const char* fn = "egnn_bn_finalize";
  EGNN_REQUIRE(sums && sumsq && mean && rstd && count > 0, fn, "bad arguments");
  bn_finalize_kernel<<<(unsigned)ceil_div(n_feat, 128), 128, 0, (cudaStream_t)stream>>>(
      sums, sumsq, count, (int)n_feat, eps, momentum, mean, rstd, running_mean, running_var);
  EGNN_LAUNCH_CHECK(fn);
  return 0;
}

__global__ void concat2_f32_kernel(const float* __restrict__ a, int64_t na, const float* __restrict__ b, int64_t nb,
                                   float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < na) out[i] = a[i];
  else if (i < na + nb) out[i] = b[i - na];
}
extern "C" int egnn_concat2_f32(const float* a, int64_t na, const float* b, int64_t nb, float* out, void* stream) {
  EGNN_REQUIRE(a && b && out && na >= 0 && nb >= 0, "egnn_concat2_f32", "bad arguments");
  if (na + nb == 0) return 0;
  concat2_f32_kernel<<<(unsigned)ceil_div(na + nb, 256), 256, 0, (cudaStream_t)stream>>>(a, na, b, nb, out);
  EGNN_LAUNCH_CHECK("egnn_concat2_f32");
  return 0;
}

__global__ void f64_to_f32_kernel(const double* __restrict__ in, float* __restrict__ out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (float)in[i];
}
extern "C" int egnn_f64_to_f32(const double* in, float* out, int64_t n, void* stream) {
  EGNN_REQUIRE(in && out && n >= 0, "egnn_f64_to_f32", "bad arguments");
  if (n == 0) return 0;
  f64_to_f32_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(in, out, n);
  EGNN_LAUNCH_CHECK("egnn_f64_to_f32");
  return 0;
}

extern "C" int egnn_colstats_reduce(const float* parts, int64_t n_parts, int64_t n_feat, double* sums, void* stream) {
  const char* fn = "egnn_colstats_reduce";
  EGNN_REQUIRE(parts && sums && n_parts > 0 && n_feat > 0, fn, "bad arguments");
  colstats_parts_kernel<false><<<(unsigned)ceil_div(n_feat, kPartCols), kPartCols * kPartGroups, 0, (cudaStream_t)stream>>>(
      parts, (int)n_parts, (int)n_feat, sums, 1.0, 0.f, 0.f, nullptr, nullptr, nullptr, nullptr, nullptr);
  EGNN_LAUNCH_CHECK(fn);
  return 0;
}

extern "C" int egnn_bn_finalize_parts(const float* parts, int64_t n_parts, int64_t n_feat, double count, float eps,
                                      float momentum, float* mean, float* rstd, float* running_mean,
                                      float* running_var, int64_t* num_batches_tracked, void* stream) {
  const char* fn = "egnn_bn_finalize_parts";
  EGNN_REQUIRE(parts && mean && rstd && n_parts > 0 && n_feat > 0 && count > 0, fn, "bad arguments");
  colstats_parts_kernel<true><<<(unsigned)ceil_div(n_feat, kPartCols), kPartCols * kPartGroups, 0, (cudaStream_t)stream>>>(
      parts, (int)n_parts, (int)n_feat, nullptr, count, eps, momentum, mean, rstd, running_mean, running_var,
      num_batches_tracked);
  EGNN_LAUNCH_CHECK(fn);
  return 0;
}

static ActCtx make_ctx(const float* mean, const float* rstd, const float* gamma, const float* beta, int act,
                       float p, uint64_t seed, const int64_t* seed_off, uint32_t layer, int64_t row0) {
  ActCtx C;
  C.mean = mean; C.rstd = rstd; C.gamma = gamma; C.beta = beta;
  C.act = act;
  C.drop = p > 0.f;
  C.scale = p > 0.f ? (float)(1.0 / (1.0 - (double)p)) : 1.f;
  C.thr = dropout_threshold16(p);
  C.seed = seed; C.seed_off = seed_off; C.layer = layer; C.row0 = row0;
  return C;
}

extern "C" int egnn_bn_act_dropout_res_fwd(const void* z, const void* res, void* y, int dtype, int64_t ld,
                                           int64_t n_rows, int64_t n_feat, const float* mean,
                                           const float* rstd, const float* gamma, const float* beta,
                                           int act, float p, uint64_t seed, const int64_t* seed_off, uint32_t layer,
                                           int64_t row0, int64_t ld_res, int64_t ld_y, uint8_t* keep_bits,
                                           const float* proj_w, float* proj_out, void* stream) {
  const char* fn = "egnn_bn_act_dropout_res_fwd";
  if (ld_y <= 0) ld_y = ld;
  if (ld_res <= 0) ld_res = ld;
  EGNN_REQUIRE(ld_y >= n_feat && ld_res >= n_feat, fn, "ld_y / ld_res < n_feat");
  EGNN_REQUIRE(z && y, fn, "null pointer");
  EGNN_REQUIRE(!mean || (rstd && gamma && beta), fn, "incomplete BatchNorm arguments");
  EGNN_REQUIRE(p >= 0.f && p < 1.f, fn, "dropout p must be in [0,1)");
  EGNN_REQUIRE((proj_w == nullptr) == (proj_out == nullptr), fn, "proj_w / proj_out must be given together");
  if (n_rows == 0) return 0;
  ActCtx C = make_ctx(mean, rstd, gamma, beta, act, p, seed, seed_off, layer, row0);
  cudaStream_t st = (cudaStream_t)stream;
  // the plain relu / dropout variant (no BatchNorm constants, no projection, bf16) fits 3 CTAs per SM: 50 % more rows in
  // flight where the pass is bound by bytes in flight (SAGENet on the replicated graphs)
  const bool three = dtype == EGNN_BF16 && !(mean && rstd) && !(proj_w && proj_out);
  FastPlan f8 = plan8(n_rows, n_feat, dtype, {ld, res ? ld_res : 8, ld_y}, {z, res, y}, three ? 3 : 2);
  EGNN_REQUIRE(!proj_w || (f8.ok && act == EGNN_ACT_RELU && (uintptr_t)proj_out % 16 == 0), fn,
               "the fused projection needs the 8-column path (F/8 a power of two <= 32, 16-byte rows) and ReLU");
  if (f8.ok) {
#define EGNN_FWD8(TT, BNF, ACTF, PJ)                                                                              \
  bn_act_fwd8<TT, BNF, ACTF, PJ><<<f8.nblk, kThreads, 0, st>>>((const TT*)z, (const TT*)res, (TT*)y, ld, ld_res, ld_y, \
                                                               n_rows, (int)n_feat, f8.cg_shift, f8.rpb, C, keep_bits, \
                                                               proj_w, proj_out)
#define EGNN_FWD8_ACT(TT, BNF)                                               \
  do {                                                                       \
    if (proj_w) EGNN_FWD8(TT, BNF, EGNN_ACT_RELU, true);                     \
    else if (act == EGNN_ACT_RELU) EGNN_FWD8(TT, BNF, EGNN_ACT_RELU, false); \
    else if (act == EGNN_ACT_ELU) EGNN_FWD8(TT, BNF, EGNN_ACT_ELU, false);   \
    else EGNN_FWD8(TT, BNF, EGNN_ACT_NONE, false);                           \
  } while (0)
    if (dtype == EGNN_F32) {
      if (mean) EGNN_FWD8_ACT(float, true);
      else EGNN_FWD8_ACT(float, false);
    } else {
      if (mean) EGNN_FWD8_ACT(__nv_bfloat16, true);
      else EGNN_FWD8_ACT(__nv_bfloat16, false);
    }
#undef EGNN_FWD8_ACT
#undef EGNN_FWD8
    EGNN_LAUNCH_CHECK(fn);
    return 0;
  }
  FastPlan fp = lean_plan(n_rows, n_feat, dtype, {ld, res ? ld_res : 4, ld_y}, {z, res, y});
  if (fp.ok) {
    EGNN_LEAN_DISPATCH(bn_act_fwd_lean, fp.nblk, (const TT*)z, (const TT*)res, (TT*)y, ld, ld_res, ld_y, n_rows,
                       (int)n_feat, fp.cg_shift, fp.rpb, C, keep_bits);
    EGNN_LAUNCH_CHECK(fn);
    return 0;
  }
  EGNN_REQUIRE(ld_y == ld && ld_res == ld && !keep_bits, fn,
               "separate leading dimensions / keep_bits need the 4-column streaming path");
  bool v = vec_ok(z, dtype, ld, n_feat) && vec_ok(res, dtype, ld, n_feat) && vec_ok(y, dtype, ld, n_feat);
  unsigned grid = (unsigned)ceil_div(n_rows * ceil_div(n_feat, 4), kThreads);
  if (dtype == EGNN_F32)
    bn_act_fwd_kernel<float><<<grid, kThreads, 0, st>>>((const float*)z, (const float*)res, (float*)y, ld,
                                                        n_rows, (int)n_feat, v, C);
  else
    bn_act_fwd_kernel<__nv_bfloat16><<<grid, kThreads, 0, st>>>(
        (const __nv_bfloat16*)z, (const __nv_bfloat16*)res, (__nv_bfloat16*)y, ld, n_rows, (int)n_feat, v, C);
  EGNN_LAUNCH_CHECK(fn);
  return 0;
}

extern "C" int64_t egnn_bn_bwd_reduce_parts(int64_t n_rows, int64_t n_feat) {
  FastPlan f8 = plan8(n_rows, n_feat, EGNN_BF16, {}, {});
  return f8.ok ? f8.nblk : 0;
}

extern "C" int egnn_bn_act_dropout_bwd_reduce(const void* dy, const void* z, int dtype, int64_t ld,
                                              int64_t n_rows, int64_t n_feat, const float* mean,
                                              const float* rstd, const float* gamma, const float* beta,
                                              int act, float p, uint64_t seed, const int64_t* seed_off,
                                              uint32_t layer, int64_t row0, double* sum_g, double* sum_gx, void* workspace,
                                              int64_t ld_z, const uint8_t* keep_bits, const float* dp,
                                              const float* dp_w, float* sum_g_f32, float* sum_gx_f32, void* stream) {
  const char* fn = "egnn_bn_act_dropout_bwd_reduce";
  if (ld_z <= 0) ld_z = ld;
  EGNN_REQUIRE(dy && z && workspace && mean && rstd && gamma && beta, fn, "null pointer");
  EGNN_REQUIRE((sum_g == nullptr) == (sum_gx == nullptr), fn, "sum_g / sum_gx must be given together");
  EGNN_REQUIRE((dp == nullptr) == (dp_w == nullptr), fn, "dp / dp_w must be given together");
  ActCtx C = make_ctx(mean, rstd, gamma, beta, act, p, seed, seed_off, layer, row0);
  cudaStream_t st = (cudaStream_t)stream;
  FastPlan f8 = plan8(n_rows, n_feat, dtype, {ld, ld_z}, {z, dy});
  const bool use8 = f8.ok && act == EGNN_ACT_RELU && keep_bits && n_rows > 0 && f8.nblk <= kMaxPartBlocks;
  EGNN_REQUIRE(!dp || (use8 && (uintptr_t)dp % 16 == 0), fn,
               "dy-from-dp needs the 8-column path (ReLU, saved keep bits, F/8 a power of two <= 32, 16-byte rows)");
  if (use8) {
    double* partial = reinterpret_cast<double*>(workspace);
    if (dtype == EGNN_F32) {
      if (dp) bn_relu_bwd_reduce8<float, true><<<f8.nblk, kThreads, 0, st>>>(nullptr, (float*)const_cast<void*>(dy), dp, dp_w, (const float*)z, ld, ld_z, n_rows, (int)n_feat, f8.cg_shift, f8.rpb, C, partial, keep_bits);
      else bn_relu_bwd_reduce8<float, false><<<f8.nblk, kThreads, 0, st>>>((const float*)dy, nullptr, nullptr, nullptr, (const float*)z, ld, ld_z, n_rows, (int)n_feat, f8.cg_shift, f8.rpb, C, partial, keep_bits);
    } else {
      using B = __nv_bfloat16;
      if (dp) bn_relu_bwd_reduce8<B, true><<<f8.nblk, kThreads, 0, st>>>(nullptr, (B*)const_cast<void*>(dy), dp, dp_w, (const B*)z, ld, ld_z, n_rows, (int)n_feat, f8.cg_shift, f8.rpb, C, partial, keep_bits);
      else bn_relu_bwd_reduce8<B, false><<<f8.nblk, kThreads, 0, st>>>((const B*)dy, nullptr, nullptr, nullptr, (const B*)z, ld, ld_z, n_rows, (int)n_feat, f8.cg_shift, f8.rpb, C, partial, keep_bits);
    }
    EGNN_LAUNCH_CHECK(fn);
    if (sum_g) {   // NULL: the caller reduces the egnn_bn_bwd_reduce_parts(...) partial rows itself (fused exchange)
      colreduce_final<<<dim3((unsigned)n_feat, 2), kThreads, 0, st>>>(partial, f8.nblk, (int)n_feat, sum_g, sum_gx,
                                                                      sum_g_f32, sum_gx_f32);
      EGNN_LAUNCH_CHECK(fn);
    }
    return 0;
  }
  EGNN_REQUIRE(sum_g, fn, "partials-only mode needs the 8-column path");
  EGNN_REQUIRE(!sum_g_f32 && !sum_gx_f32, fn, "fp32 copies of the sums need the 8-column path");
  FastPlan fp = lean_plan(n_rows, n_feat, dtype, {ld, ld_z}, {z, dy});
  if (fp.ok) {
    double* partial = reinterpret_cast<double*>(workspace);
    EGNN_LEAN_DISPATCH_KB(bn_act_bwd_reduce_lean, keep_bits != nullptr, fp.nblk, (const TT*)dy, ld, (const TT*)z, ld_z,
                          n_rows, (int)n_feat, fp.cg_shift, fp.rpb, C, partial, keep_bits);
    EGNN_LAUNCH_CHECK(fn);
    colreduce_final<<<dim3((unsigned)n_feat, 2), kThreads, 0, st>>>(partial, fp.nblk, (int)n_feat, sum_g, sum_gx);
    EGNN_LAUNCH_CHECK(fn);
    return 0;
  }
  if (ld_z != ld || keep_bits) return fail(fn, "separate ld_z / keep_bits need the 4-column streaming path");
  bool v = vec_ok(z, dtype, ld, n_feat) && vec_ok(dy, dtype, ld, n_feat);
  if (dtype == EGNN_F32)
    return run_colreduce(BnBwdProd<float>{(const float*)dy, (const float*)z, ld, (int)n_feat, v, C}, n_rows,
                         (int)n_feat, sum_g, sum_gx, workspace, st, fn);
  return run_colreduce(
      BnBwdProd<__nv_bfloat16>{(const __nv_bfloat16*)dy, (const __nv_bfloat16*)z, ld, (int)n_feat, v, C},
      n_rows, (int)n_feat, sum_g, sum_gx, workspace, st, fn);
}

extern "C" int egnn_bn_act_dropout_bwd_apply(const void* dy, const void* z, void* dz, int dtype, int64_t ld,
                                             int64_t n_rows, int64_t n_feat, const float* mean,
                                             const float* rstd, const float* gamma, const float* beta,
                                             int act, float p, uint64_t seed, const int64_t* seed_off, uint32_t layer,
                                             int64_t row0, const double* sum_g, const double* sum_gx, double n_total,
                                             float* dz_colsum, void* workspace, int64_t ld_z,
                                             const uint8_t* keep_bits, void* stream) {
  const char* fn = "egnn_bn_act_dropout_bwd_apply";
  if (ld_z <= 0) ld_z = ld;
  EGNN_REQUIRE(!dz_colsum || workspace, fn, "dz_colsum needs a workspace (egnn_colreduce_workspace_bytes)");
  EGNN_REQUIRE(dy && z && dz, fn, "null pointer");
  EGNN_REQUIRE(!mean || (rstd && gamma && beta && sum_g && sum_gx && n_total > 0), fn,
               "incomplete BatchNorm arguments");
  if (n_rows == 0) return 0;
  ActCtx C = make_ctx(mean, rstd, gamma, beta, act, p, seed, seed_off, layer, row0);
  cudaStream_t st = (cudaStream_t)stream;
  double inv_n = mean ? 1.0 / n_total : 0.0;
  FastPlan f8 = plan8(n_rows, n_feat, dtype, {ld, ld_z}, {z, dy, dz});
  if (f8.ok && mean && act == EGNN_ACT_RELU && keep_bits && f8.nblk <= kMaxPartBlocks) {
    double* partial = dz_colsum ? reinterpret_cast<double*>(workspace) : nullptr;
    if (dtype == EGNN_F32)
      bn_relu_bwd_apply8<float><<<f8.nblk, kThreads, 0, st>>>((const float*)dy, ld, (const float*)z, ld_z, (float*)dz, ld, n_rows, (int)n_feat, f8.cg_shift, f8.rpb, C, sum_g, sum_gx, inv_n, partial, keep_bits);
    else
      bn_relu_bwd_apply8<__nv_bfloat16><<<f8.nblk, kThreads, 0, st>>>((const __nv_bfloat16*)dy, ld, (const __nv_bfloat16*)z, ld_z, (__nv_bfloat16*)dz, ld, n_rows, (int)n_feat, f8.cg_shift, f8.rpb, C, sum_g, sum_gx, inv_n, partial, keep_bits);
    EGNN_LAUNCH_CHECK(fn);
    if (dz_colsum) {
      colsum_final<<<(unsigned)n_feat, kThreads, 0, st>>>(partial, f8.nblk, (int)n_feat, dz_colsum);
      EGNN_LAUNCH_CHECK(fn);
    }
    return 0;
  }
  FastPlan fp = lean_plan(n_rows, n_feat, dtype, {ld, ld_z}, {z, dy, dz});
  if (fp.ok) {
    double* partial = dz_colsum ? reinterpret_cast<double*>(workspace) : nullptr;
    EGNN_LEAN_DISPATCH_KB(bn_act_bwd_apply_lean, keep_bits != nullptr, fp.nblk, (const TT*)dy, ld, (const TT*)z, ld_z,
                          (TT*)dz, ld, n_rows, (int)n_feat, fp.cg_shift, fp.rpb, C, sum_g, sum_gx, inv_n, partial,
                          keep_bits);
    EGNN_LAUNCH_CHECK(fn);
    if (dz_colsum) {
      colsum_final<<<(unsigned)n_feat, kThreads, 0, st>>>(partial, fp.nblk, (int)n_feat, dz_colsum);
      EGNN_LAUNCH_CHECK(fn);
    }
    return 0;
  }
  if (ld_z != ld || keep_bits) return fail(fn, "separate ld_z / keep_bits need the 4-column streaming path");
  bool v = vec_ok(z, dtype, ld, n_feat) && vec_ok(dy, dtype, ld, n_feat) && vec_ok(dz, dtype, ld, n_feat);
  unsigned grid = (unsigned)ceil_div(n_rows * ceil_div(n_feat, 4), kThreads);
  if (dtype == EGNN_F32)
    bn_act_bwd_apply_kernel<float><<<grid, kThreads, 0, st>>>((const float*)dy, (const float*)z, (float*)dz,
                                                              ld, n_rows, (int)n_feat, v, C, sum_g, sum_gx,
                                                              inv_n);
  else
    bn_act_bwd_apply_kernel<__nv_bfloat16><<<grid, kThreads, 0, st>>>(
        (const __nv_bfloat16*)dy, (const __nv_bfloat16*)z, (__nv_bfloat16*)dz, ld, n_rows, (int)n_feat, v, C,
        sum_g, sum_gx, inv_n);
  EGNN_LAUNCH_CHECK(fn);
  if (dz_colsum) {  // generic shapes: a separate deterministic column reduction of dz
    double* sums = reinterpret_cast<double*>(workspace);
    char* ws2 = reinterpret_cast<char*>(workspace) + 8 * sizeof(double) * (size_t)(((n_feat + 3) / 4) * 4);
    int rc = egnn_colreduce(dz, dtype, ld, n_rows, n_feat, sums, nullptr, ws2, stream);
    if (rc) return rc;
    colsum_final<<<(unsigned)n_feat, kThreads, 0, st>>>(sums, 1, (int)n_feat, dz_colsum);
    EGNN_LAUNCH_CHECK(fn);
  }
  return 0;
}

extern "C" int egnn_dropout_mask(uint8_t* mask, int64_t n_rows, int64_t n_feat, float p, uint64_t seed,
                                 const int64_t* seed_off, uint32_t layer, int64_t row0, void* stream) {
  const char* fn = "egnn_dropout_mask";
  EGNN_REQUIRE(mask && n_feat > 0, fn, "bad arguments");
  if (n_rows == 0) return 0;
  unsigned grid = (unsigned)ceil_div(n_rows * ceil_div(n_feat, 8), kThreads);
  dropout_mask_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(mask, n_rows, (int)n_feat,
                                                                   dropout_threshold16(p), seed, seed_off, layer, row0);
  EGNN_LAUNCH_CHECK(fn);
  return 0;
}

__global__ void counter_add_kernel(int64_t* c, int64_t inc) { *c += inc; }
extern "C" int egnn_counter_add(int64_t* counter, int64_t inc, void* stream) {
  EGNN_REQUIRE(counter, "egnn_counter_add", "null pointer");
  counter_add_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(counter, inc);
  EGNN_LAUNCH_CHECK("egnn_counter_add");
  return 0;
}

extern "C" size_t egnn_ce_workspace_floats(int64_t n_idx) { return (size_t)ceil_div(n_idx, kThreads) + 8; }

extern "C" int egnn_masked_loss(const void* logits, int dtype, int64_t n_rows, const int64_t* y,
                                const int64_t* idx, int64_t n_idx, const float* cw, double n_total,
                                double focal_gamma, const int64_t* timestep, double t_min, double t_max,
                                int time_scheme, float* loss, void* dlogits, float* workspace, void* stream) {
  const char* fn = "egnn_masked_loss";
  EGNN_REQUIRE(logits && y && loss && dlogits && workspace, fn, "null pointer");
  EGNN_REQUIRE(focal_gamma >= 0 || cw, fn, "class weights missing");
  EGNN_REQUIRE(n_total > 0 && n_idx >= 0 && (idx || n_idx <= n_rows), fn, "bad arguments");
  EGNN_REQUIRE(time_scheme >= 0 && time_scheme <= 2 && (time_scheme == 0 || timestep), fn, "bad time weighting");
  cudaStream_t st = (cudaStream_t)stream;
  size_t es = dtype == EGNN_F32 ? 4 : 2;
  cudaMemsetAsync(dlogits, 0, (size_t)n_rows * 2 * es, st);
  int nblk = (int)ceil_div(n_idx, kThreads);
  const float denom = (float)(t_max - t_min > 1.0 ? t_max - t_min : 1.0);   // _norm_train_time (:131-133)
  if (nblk > 0) {
    if (dtype == EGNN_F32)
      masked_ce_kernel<float><<<nblk, kThreads, 0, st>>>((const float*)logits, y, idx, n_idx, cw, (float)(1.0 / n_total),
                                                         (float)focal_gamma, timestep, (float)t_min, denom, time_scheme,
                                                         (float*)dlogits, workspace);
    else
      masked_ce_kernel<__nv_bfloat16><<<nblk, kThreads, 0, st>>>((const __nv_bfloat16*)logits, y, idx, n_idx, cw,
                                                                 (float)(1.0 / n_total), (float)focal_gamma, timestep,
                                                                 (float)t_min, denom, time_scheme,
                                                                 (__nv_bfloat16*)dlogits, workspace);
    EGNN_LAUNCH_CHECK(fn);
  }
  ce_final_kernel<<<1, kThreads, 0, st>>>(workspace, nblk, 1.0 / n_total, loss);
  EGNN_LAUNCH_CHECK(fn);
  return 0;
}

extern "C" int egnn_masked_ce(const void* logits, int dtype, int64_t n_rows, const int64_t* y,
                              const int64_t* idx, int64_t n_idx, const float* cw, double n_total,
                              float* loss, void* dlogits, float* workspace, void* stream) {
  EGNN_REQUIRE(cw && (n_idx == 0 || idx), "egnn_masked_ce", "bad arguments");
  return egnn_masked_loss(logits, dtype, n_rows, y, idx, n_idx, cw, n_total, -1.0, nullptr, 0.0, 1.0, 0, loss, dlogits,
                          workspace, stream);
}

extern "C" int egnn_l2_mean_penalty(const float* w, int64_t n, double lambda, float* loss, float* grad, void* stream) {
  const char* fn = "egnn_l2_mean_penalty";
  EGNN_REQUIRE(w && loss && n > 0, fn, "bad arguments");
  l2_mean_penalty_kernel<<<1, kThreads, 0, (cudaStream_t)stream>>>(w, n, (float)lambda, loss, grad);
  EGNN_LAUNCH_CHECK(fn);
  return 0;
}

extern "C" size_t egnn_adam_workspace_floats(int64_t n) {
  (void)n;
  return (size_t)kNumSMs * 4 + 16;
}

extern "C" int egnn_clip_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                                   int64_t n, float lr, float beta1, float beta2, float eps,
                                   float weight_decay, float max_norm, int64_t* step_count,
                                   float* grad_norm_out, float* workspace, void* stream) {
  const char* fn = "egnn_clip_adam_step";
  EGNN_REQUIRE(param && grad && exp_avg && exp_avg_sq && step_count && workspace, fn, "null pointer");
  if (n <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  int nblk = (int)ceil_div(n, kThreads);
  if (nblk > kNumSMs * 4) nblk = kNumSMs * 4;
  float* coefs = workspace;
  float* partial = workspace + 16;
  sqnorm_partial_kernel<<<nblk, kThreads, 0, st>>>(grad, n, partial);
  EGNN_LAUNCH_CHECK(fn);
  adam_prepare_kernel<<<1, kThreads, 0, st>>>(partial, nblk, max_norm, beta1, beta2, step_count, coefs,
                                              grad_norm_out);
  EGNN_LAUNCH_CHECK(fn);
  adam_apply_kernel<<<(unsigned)ceil_div(n, kThreads), kThreads, 0, st>>>(
      param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, coefs);
  EGNN_LAUNCH_CHECK(fn);
  return 0;
}
