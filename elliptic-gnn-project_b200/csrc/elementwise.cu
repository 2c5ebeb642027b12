// K7 and the step tail: cast / time-feature injection, deterministic column reductions,
// BatchNorm(batch stats) + ReLU/ELU + dropout + residual forward/backward, masked weighted
// cross-entropy, global-norm clip + Adam.  All HBM-bound streaming kernels: 4 features per
// thread (128-bit fp32 / 64-bit bf16 accesses), fp32 math, fp64 only for cross-thread
// combination of statistics (torch CPU BatchNorm accumulates in double too).
#include "common.cuh"
#include "philox.cuh"

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
__global__ void __launch_bounds__(kThreads) inject_time_kernel(
    const float* __restrict__ x, int64_t ld_x, const int64_t* __restrict__ t,
    const float* __restrict__ table, int64_t T, int D, float* __restrict__ o32,
    __nv_bfloat16* __restrict__ o16, int64_t ld_out, int64_t n_rows, int F) {
  const int W2 = (int)(ld_out / 2);  // ld_out is even (multiple of 4)
  int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n_rows * W2) return;
  int64_t r = i / W2;
  int c = (int)(i - r * W2) * 2;
  float v[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    int cc = c + k;
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
  if (o32) *reinterpret_cast<float2*>(o32 + r * ld_out + c) = make_float2(v[0], v[1]);
  if (o16) *reinterpret_cast<uint32_t*>(o16 + r * ld_out + c) = pack_bf16x2(v[0], v[1]);
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
  uint32_t words[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
  if (C.drop) {
    const uint64_t seed = C.seed + (C.seed_off ? (uint64_t)*C.seed_off : 0ull);
    Philox4 w = dropout_words(seed, C.layer, C.row0 + r, (uint32_t)(c >> 2));
#pragma unroll
    for (int k = 0; k < 4; ++k) words[k] = w.v[k];
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
    if (C.drop) ks = words[k] >= C.thr ? C.scale : 0.f;
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
    for (int64_t r = r0 + rl; r < r1; r += RL) {
      F4 v0, v1;
      prod.eval(r, c, v0, v1);
      a[0] += v0.x; a[1] += v0.y; a[2] += v0.z; a[3] += v0.w;
      a[4] += v1.x; a[5] += v1.y; a[6] += v1.z; a[7] += v1.w;
      if (++cnt == 32) {
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
                                                            double* __restrict__ out1) {
  __shared__ double sm[kThreads];
  const int Fp = ((F + 3) >> 2) << 2;
  const int col = blockIdx.x, which = blockIdx.y;
  double* out = which == 0 ? out0 : out1;
  if (!out) return;
  double s = 0;
  for (int b = threadIdx.x; b < nblk; b += kThreads) s += partial[((int64_t)b * 2 + which) * Fp + col];
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int o = kThreads / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[col] = sm[0];
}

template <typename Prod>
int run_colreduce(const Prod& prod, int64_t n_rows, int F, double* out0, double* out1, void* ws,
                  cudaStream_t st, const char* fn) {
  int ncg = (F + 3) / 4;
  int CW4 = 1;
  while (CW4 < ncg && CW4 < kThreads) CW4 <<= 1;
  int RL = kThreads / CW4;
  int64_t rpb = (int64_t)RL * 32;
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

__global__ void __launch_bounds__(kThreads) dropout_mask_kernel(uint8_t* __restrict__ mask,
                                                                int64_t n_rows, int F, uint32_t thr,
                                                                uint64_t seed, const int64_t* seed_off,
                                                                uint32_t layer, int64_t row0) {
  const int F4n = (F + 3) >> 2;
  int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n_rows * F4n) return;
  int64_t r = i / F4n;
  int cb = (int)(i - r * F4n);
  Philox4 w = dropout_words(seed + (seed_off ? (uint64_t)*seed_off : 0ull), layer, row0 + r, (uint32_t)cb);
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (cb * 4 + k < F) mask[r * F + cb * 4 + k] = w.v[k] >= thr ? 1 : 0;
}

// ---- masked weighted cross-entropy (2 classes) --------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads) masked_ce_kernel(const T* __restrict__ logits,
                                                             const int64_t* __restrict__ y,
                                                             const int64_t* __restrict__ idx,
                                                             int64_t n_idx, const float* __restrict__ cw,
                                                             float inv_n, T* __restrict__ dlogits,
                                                             float* __restrict__ partial) {
  __shared__ float sm[kThreads];
  int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  float li = 0.f;
  if (i < n_idx) {
    int64_t r = idx[i];
    float l0 = to_f32(logits[2 * r]), l1 = to_f32(logits[2 * r + 1]);
    int64_t yi = y[r];
    float m = fmaxf(l0, l1);
    float e0 = expf(l0 - m), e1 = expf(l1 - m);
    float s = e0 + e1;
    float lse = m + logf(s);
    float w = cw[yi];
    li = w * (lse - (yi == 0 ? l0 : l1));
    float p0 = e0 / s, p1 = e1 / s;
    float g0 = w * (p0 - (yi == 0 ? 1.f : 0.f)) * inv_n;
    float g1 = w * (p1 - (yi == 1 ? 1.f : 0.f)) * inv_n;
    dlogits[2 * r] = from_f32<T>(g0);
    dlogits[2 * r + 1] = from_f32<T>(g1);
  }
  sm[threadIdx.x] = li;
  __syncthreads();
  for (int o = kThreads / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sm[0];
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

extern "C" int egnn_inject_time(const float* x, int64_t ld_x, const int64_t* t, const float* table,
                                int64_t T, int64_t D, float* out_f32, void* out_bf16, int64_t ld_out,
                                int64_t n_rows, int64_t n_feat, void* stream) {
  const char* fn = "egnn_inject_time";
  EGNN_REQUIRE(x && (out_f32 || out_bf16), fn, "null pointer");
  EGNN_REQUIRE(D == 0 || (t && table && T > 0), fn, "time table / indices missing");
  EGNN_REQUIRE(ld_out % 4 == 0 && ld_out >= n_feat + D, fn, "ld_out must be a multiple of 4 and >= F+D");
  if (n_rows == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  unsigned grid = (unsigned)ceil_div(n_rows * (ld_out / 2), kThreads);
  inject_time_kernel<<<grid, kThreads, 0, st>>>(x, ld_x, t, table, T, (int)D, out_f32,
                                                (__nv_bfloat16*)out_bf16, ld_out, n_rows, (int)n_feat);
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

static ActCtx make_ctx(const float* mean, const float* rstd, const float* gamma, const float* beta, int act,
                       float p, uint64_t seed, const int64_t* seed_off, uint32_t layer, int64_t row0) {
  ActCtx C;
  C.mean = mean; C.rstd = rstd; C.gamma = gamma; C.beta = beta;
  C.act = act;
  C.drop = p > 0.f;
  C.scale = p > 0.f ? (float)(1.0 / (1.0 - (double)p)) : 1.f;
  C.thr = dropout_threshold(p);
  C.seed = seed; C.seed_off = seed_off; C.layer = layer; C.row0 = row0;
  return C;
}

extern "C" int egnn_bn_act_dropout_res_fwd(const void* z, const void* res, void* y, int dtype, int64_t ld,
                                           int64_t n_rows, int64_t n_feat, const float* mean,
                                           const float* rstd, const float* gamma, const float* beta,
                                           int act, float p, uint64_t seed, const int64_t* seed_off, uint32_t layer,
                                           int64_t row0, void* stream) {
  const char* fn = "egnn_bn_act_dropout_res_fwd";
  EGNN_REQUIRE(z && y, fn, "null pointer");
  EGNN_REQUIRE(!mean || (rstd && gamma && beta), fn, "incomplete BatchNorm arguments");
  EGNN_REQUIRE(p >= 0.f && p < 1.f, fn, "dropout p must be in [0,1)");
  if (n_rows == 0) return 0;
  ActCtx C = make_ctx(mean, rstd, gamma, beta, act, p, seed, seed_off, layer, row0);
  bool v = vec_ok(z, dtype, ld, n_feat) && vec_ok(res, dtype, ld, n_feat) && vec_ok(y, dtype, ld, n_feat);
  unsigned grid = (unsigned)ceil_div(n_rows * ceil_div(n_feat, 4), kThreads);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == EGNN_F32)
    bn_act_fwd_kernel<float><<<grid, kThreads, 0, st>>>((const float*)z, (const float*)res, (float*)y, ld,
                                                        n_rows, (int)n_feat, v, C);
  else
    bn_act_fwd_kernel<__nv_bfloat16><<<grid, kThreads, 0, st>>>(
        (const __nv_bfloat16*)z, (const __nv_bfloat16*)res, (__nv_bfloat16*)y, ld, n_rows, (int)n_feat, v, C);
  EGNN_LAUNCH_CHECK(fn);
  return 0;
}

extern "C" int egnn_bn_act_dropout_bwd_reduce(const void* dy, const void* z, int dtype, int64_t ld,
                                              int64_t n_rows, int64_t n_feat, const float* mean,
                                              const float* rstd, const float* gamma, const float* beta,
                                              int act, float p, uint64_t seed, const int64_t* seed_off,
                                              uint32_t layer, int64_t row0, double* sum_g, double* sum_gx, void* workspace,
                                              void* stream) {
  const char* fn = "egnn_bn_act_dropout_bwd_reduce";
  EGNN_REQUIRE(dy && z && sum_g && sum_gx && workspace && mean && rstd && gamma && beta, fn, "null pointer");
  ActCtx C = make_ctx(mean, rstd, gamma, beta, act, p, seed, seed_off, layer, row0);
  bool v = vec_ok(z, dtype, ld, n_feat) && vec_ok(dy, dtype, ld, n_feat);
  cudaStream_t st = (cudaStream_t)stream;
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
                                             void* stream) {
  const char* fn = "egnn_bn_act_dropout_bwd_apply";
  EGNN_REQUIRE(dy && z && dz, fn, "null pointer");
  EGNN_REQUIRE(!mean || (rstd && gamma && beta && sum_g && sum_gx && n_total > 0), fn,
               "incomplete BatchNorm arguments");
  if (n_rows == 0) return 0;
  ActCtx C = make_ctx(mean, rstd, gamma, beta, act, p, seed, seed_off, layer, row0);
  bool v = vec_ok(z, dtype, ld, n_feat) && vec_ok(dy, dtype, ld, n_feat) && vec_ok(dz, dtype, ld, n_feat);
  unsigned grid = (unsigned)ceil_div(n_rows * ceil_div(n_feat, 4), kThreads);
  cudaStream_t st = (cudaStream_t)stream;
  double inv_n = mean ? 1.0 / n_total : 0.0;
  if (dtype == EGNN_F32)
    bn_act_bwd_apply_kernel<float><<<grid, kThreads, 0, st>>>((const float*)dy, (const float*)z, (float*)dz,
                                                              ld, n_rows, (int)n_feat, v, C, sum_g, sum_gx,
                                                              inv_n);
  else
    bn_act_bwd_apply_kernel<__nv_bfloat16><<<grid, kThreads, 0, st>>>(
        (const __nv_bfloat16*)dy, (const __nv_bfloat16*)z, (__nv_bfloat16*)dz, ld, n_rows, (int)n_feat, v, C,
        sum_g, sum_gx, inv_n);
  EGNN_LAUNCH_CHECK(fn);
  return 0;
}

extern "C" int egnn_dropout_mask(uint8_t* mask, int64_t n_rows, int64_t n_feat, float p, uint64_t seed,
                                 const int64_t* seed_off, uint32_t layer, int64_t row0, void* stream) {
  const char* fn = "egnn_dropout_mask";
  EGNN_REQUIRE(mask && n_feat > 0, fn, "bad arguments");
  if (n_rows == 0) return 0;
  unsigned grid = (unsigned)ceil_div(n_rows * ceil_div(n_feat, 4), kThreads);
  dropout_mask_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(mask, n_rows, (int)n_feat,
                                                                   dropout_threshold(p), seed, seed_off, layer, row0);
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

extern "C" int egnn_masked_ce(const void* logits, int dtype, int64_t n_rows, const int64_t* y,
                              const int64_t* idx, int64_t n_idx, const float* cw, double n_total,
                              float* loss, void* dlogits, float* workspace, void* stream) {
  const char* fn = "egnn_masked_ce";
  EGNN_REQUIRE(logits && y && cw && loss && dlogits && workspace, fn, "null pointer");
  EGNN_REQUIRE(n_total > 0 && (n_idx == 0 || idx), fn, "bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  size_t es = dtype == EGNN_F32 ? 4 : 2;
  cudaMemsetAsync(dlogits, 0, (size_t)n_rows * 2 * es, st);
  int nblk = (int)ceil_div(n_idx, kThreads);
  if (nblk > 0) {
    if (dtype == EGNN_F32)
      masked_ce_kernel<float><<<nblk, kThreads, 0, st>>>((const float*)logits, y, idx, n_idx, cw,
                                                         (float)(1.0 / n_total), (float*)dlogits, workspace);
    else
      masked_ce_kernel<__nv_bfloat16><<<nblk, kThreads, 0, st>>>((const __nv_bfloat16*)logits, y, idx, n_idx,
                                                                 cw, (float)(1.0 / n_total),
                                                                 (__nv_bfloat16*)dlogits, workspace);
    EGNN_LAUNCH_CHECK(fn);
  }
  ce_final_kernel<<<1, kThreads, 0, st>>>(workspace, nblk, 1.0 / n_total, loss);
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
