// K6 (SIMT path) -- dense projections with exact-fp32 FFMA accumulation.
//
// This is the fp32-parity path (plain TF32 cannot meet rel 1e-5) and the
// reference the tcgen05 bf16 path (gemm_tcgen05.cu) is checked against.  It covers the
// three products of a Linear layer through one strided interface:
//   forward  C = X W^T      (A k-contiguous, B k-contiguous)
//   dgrad    D = G W        (A k-contiguous, B n-contiguous)
//   wgrad    dW = G^T X     (A m-contiguous, B n-contiguous, reduction over the node axis,
//                            split-K with a fixed-order second stage => deterministic)
// plus two skinny specialisations for the 2-class output layer (N<=8 forward, M<=8 wgrad),
// which are pure streaming reductions.
#include <algorithm>

#include <type_traits>

#include "common.cuh"

namespace egnn {
namespace {

constexpr int BM = 128, BN = 64, BK = 16, kThreads = 256;

struct GemmParams {
  const void* A;
  const void* B;
  void* C;
  const float* bias;
  const int32_t* row_div_ptr;
  int64_t row_div_cols;  // columns [0, row_div_cols) are divided (<= 0: all)
  float* ws;
  int64_t a_sm, a_sk, b_sk, b_sn, ld_c;
  int64_t M, N, K;
  int64_t k_per_split;  // multiple of BK
  int c_dtype, accumulate, split_k;
};

__device__ __forceinline__ void store_c(const GemmParams& P, int64_t m, int64_t n, float v) {
  if (P.bias) v += P.bias[n];
  if (P.row_div_ptr && (P.row_div_cols <= 0 || n < P.row_div_cols)) {
    int d = P.row_div_ptr[m + 1] - P.row_div_ptr[m];
    v = __fdiv_rn(v, (float)(d > 1 ? d : 1));
  }
  if (P.c_dtype == EGNN_F32) {
    float* c = reinterpret_cast<float*>(P.C) + m * P.ld_c + n;
    if (P.accumulate) v += *c;
    *c = v;
  } else {
    __nv_bfloat16* c = reinterpret_cast<__nv_bfloat16*>(P.C) + m * P.ld_c + n;
    if (P.accumulate) v += __bfloat162float(*c);
    *c = __float2bfloat16_rn(v);
  }
}

template <typename TA, typename TB>
__global__ void __launch_bounds__(kThreads) gemm_simt(GemmParams P) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const TA* __restrict__ A = reinterpret_cast<const TA*>(P.A);
  const TB* __restrict__ B = reinterpret_cast<const TB*>(P.B);
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.x * BM, n0 = (int64_t)blockIdx.y * BN;
  const int64_t kb = (int64_t)blockIdx.z * P.k_per_split;
  const int64_t ke = min(P.K, kb + P.k_per_split);
  const bool a_mfast = P.a_sm <= P.a_sk, b_nfast = P.b_sn <= P.b_sk;

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int64_t k0 = kb; k0 < ke; k0 += BK) {
#pragma unroll
    for (int i = 0; i < (BM * BK) / kThreads; ++i) {
      int idx = tid + i * kThreads;
      int mm = a_mfast ? (idx % BM) : (idx / BK);
      int kk = a_mfast ? (idx / BM) : (idx % BK);
      int64_t m = m0 + mm, k = k0 + kk;
      As[kk][mm] = (m < P.M && k < ke) ? to_f32(A[m * P.a_sm + k * P.a_sk]) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < (BN * BK) / kThreads; ++i) {
      int idx = tid + i * kThreads;
      int nn = b_nfast ? (idx % BN) : (idx / BK);
      int kk = b_nfast ? (idx / BN) : (idx % BK);
      int64_t n = n0 + nn, k = k0 + kk;
      Bs[kk][nn] = (n < P.N && k < ke) ? to_f32(B[k * P.b_sk + n * P.b_sn]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
      float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int64_t m = m0 + ty * 8 + i;
    if (m >= P.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int64_t n = n0 + tx * 4 + j;
      if (n >= P.N) continue;
      if (P.split_k > 1)
        P.ws[((int64_t)blockIdx.z * P.M + m) * P.N + n] = acc[i][j];
      else
        store_c(P, m, n, acc[i][j]);
    }
  }
}

// fp32 fast path (the exact-fp32 parity / eval mode): A [M,K] and W [N,K] both contiguous along K (Linear forward,
// and dgrad through a pre-transposed weight), 16-byte aligned rows.  Same 128x64x16 tiling and 8x4 micro-tile as the
// generic kernel, but 128-bit global loads, register prefetch of the next K slab while the current one is
// multiplied, and double-buffered shared memory (one barrier per slab).  FFMA only: TF32 would break rel 1e-5.
__global__ void __launch_bounds__(kThreads) gemm_simt_f32_kmajor(GemmParams P) {
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];
  const float* __restrict__ A = reinterpret_cast<const float*>(P.A);
  const float* __restrict__ B = reinterpret_cast<const float*>(P.B);
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.x * BM, n0 = (int64_t)blockIdx.y * BN;
  const int K = (int)P.K;
  // this thread's slab loads: two float4 of A (rows ar0, ar0 + 64), one float4 of W
  const int ar0 = tid >> 2, akq = (tid & 3) * 4;
  const int64_t am0 = m0 + ar0, am1 = m0 + ar0 + 64, bn = n0 + ar0;
  const float* pa0 = A + (am0 < P.M ? am0 : 0) * P.a_sm + akq;
  const float* pa1 = A + (am1 < P.M ? am1 : 0) * P.a_sm + akq;
  const float* pb = B + (bn < P.N ? bn : 0) * P.b_sn + akq;
  const bool va0 = am0 < P.M, va1 = am1 < P.M, vb = bn < P.N;
  float4 ra0, ra1, rb;
  auto gload = [&](int k0) {
    const bool kin = k0 + akq < K;  // K % 4 == 0: a float4 is entirely inside or outside
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    ra0 = (va0 && kin) ? __ldg(reinterpret_cast<const float4*>(pa0 + k0)) : z;
    ra1 = (va1 && kin) ? __ldg(reinterpret_cast<const float4*>(pa1 + k0)) : z;
    rb = (vb && kin && ar0 < BN) ? __ldg(reinterpret_cast<const float4*>(pb + k0)) : z;
  };
  auto sstore = [&](int buf) {
    As[buf][akq + 0][ar0] = ra0.x; As[buf][akq + 1][ar0] = ra0.y; As[buf][akq + 2][ar0] = ra0.z; As[buf][akq + 3][ar0] = ra0.w;
    As[buf][akq + 0][ar0 + 64] = ra1.x; As[buf][akq + 1][ar0 + 64] = ra1.y;
    As[buf][akq + 2][ar0 + 64] = ra1.z; As[buf][akq + 3][ar0 + 64] = ra1.w;
    if (ar0 < BN) {
      Bs[buf][akq + 0][ar0] = rb.x; Bs[buf][akq + 1][ar0] = rb.y; Bs[buf][akq + 2][ar0] = rb.z; Bs[buf][akq + 3][ar0] = rb.w;
    }
  };
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  gload(0);
  sstore(0);
  __syncthreads();
  int buf = 0;
  for (int k0 = 0; k0 < K; k0 += BK) {
    const bool more = k0 + BK < K;
    if (more) gload(k0 + BK);  // in flight while this slab is multiplied
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8 + 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (more) {
      sstore(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + ty * 8 + i;
    if (m >= P.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t n = n0 + tx * 4 + j;
      if (n < P.N) store_c(P, m, n, acc[i][j]);
    }
  }
}

// fp32 weight-gradient shape: C[M,N] = sum_k A(m,k) B(k,n) with BOTH operands contiguous along the non-contracted
// axis (A = G^T: element (m,k) at A[k*a_sk + m];  B = X: element (k,n) at B[k*b_sk + n]) and a long contraction
// (the node axis), split over blockIdx.z.  128-bit loads straight into the [k][m] / [k][n] shared tiles, register
// prefetch and double buffering as above; partial sums go to the split-K workspace (fixed-order reduce).
__global__ void __launch_bounds__(kThreads) gemm_simt_f32_mnmajor(GemmParams P) {
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];
  const float* __restrict__ A = reinterpret_cast<const float*>(P.A);
  const float* __restrict__ B = reinterpret_cast<const float*>(P.B);
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.x * BM, n0 = (int64_t)blockIdx.y * BN;
  const int64_t kb = (int64_t)blockIdx.z * P.k_per_split;
  const int64_t ke = min(P.K, kb + P.k_per_split);
  // slab loads: A 16 x 128 floats = 512 float4 (two per thread), B 16 x 64 floats = 256 float4 (one per thread)
  const int ak0 = tid >> 5, am4 = (tid & 31) * 4;     // k rows ak0 and ak0 + 8
  const int bk = tid >> 4, bn4 = (tid & 15) * 4;
  const bool vam = m0 + am4 < P.M, vbn = n0 + bn4 < P.N;  // M % 4 == 0, N % 4 == 0
  float4 ra0, ra1, rb;
  auto gload = [&](int64_t k0) {
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    const int64_t k_a0 = k0 + ak0, k_a1 = k0 + ak0 + 8, k_b = k0 + bk;
    ra0 = (vam && k_a0 < ke) ? __ldg(reinterpret_cast<const float4*>(A + k_a0 * P.a_sk + m0 + am4)) : z;
    ra1 = (vam && k_a1 < ke) ? __ldg(reinterpret_cast<const float4*>(A + k_a1 * P.a_sk + m0 + am4)) : z;
    rb = (vbn && k_b < ke) ? __ldg(reinterpret_cast<const float4*>(B + k_b * P.b_sk + n0 + bn4)) : z;
  };
  auto sstore = [&](int buf) {
    *reinterpret_cast<float4*>(&As[buf][ak0][am4]) = ra0;
    *reinterpret_cast<float4*>(&As[buf][ak0 + 8][am4]) = ra1;
    *reinterpret_cast<float4*>(&Bs[buf][bk][bn4]) = rb;
  };
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  gload(kb);
  sstore(0);
  __syncthreads();
  int buf = 0;
  for (int64_t k0 = kb; k0 < ke; k0 += BK) {
    const bool more = k0 + BK < ke;
    if (more) gload(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8 + 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (more) {
      sstore(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + ty * 8 + i;
    if (m >= P.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t n = n0 + tx * 4 + j;
      if (n >= P.N) continue;
      if (P.split_k > 1) P.ws[((int64_t)blockIdx.z * P.M + m) * P.N + n] = acc[i][j];
      else store_c(P, m, n, acc[i][j]);
    }
  }
}

// fixed-order second stage of split-K
__global__ void __launch_bounds__(kThreads) splitk_reduce(GemmParams P) {
  int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= P.M * P.N) return;
  float s = 0.f;
  for (int z = 0; z < P.split_k; ++z) s += P.ws[(int64_t)z * P.M * P.N + i];
  store_c(P, i / P.N, i % P.N, s);
}

// ---- skinny forward: C[M, N<=8] = A[M,K] . W[N,K]^T, one warp per row, lanes split K ----
template <typename TA, typename TB, int NOUT>
__global__ void __launch_bounds__(kThreads) gemm_skinny_fwd(GemmParams P) {
  const TA* __restrict__ A = reinterpret_cast<const TA*>(P.A);
  const TB* __restrict__ W = reinterpret_cast<const TB*>(P.B);
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * kThreads + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * kThreads) >> 5;
  for (int64_t m = warp; m < P.M; m += n_warps) {
    float s[NOUT];
#pragma unroll
    for (int j = 0; j < NOUT; ++j) s[j] = 0.f;
    for (int64_t k = lane; k < P.K; k += 32) {
      float a = to_f32(A[m * P.a_sm + k]);
#pragma unroll
      for (int j = 0; j < NOUT; ++j)
        if (j < P.N) s[j] = fmaf(a, to_f32(W[j * P.b_sn + k]), s[j]);
    }
#pragma unroll
    for (int j = 0; j < NOUT; ++j)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s[j] += __shfl_xor_sync(0xffffffffu, s[j], o);
    if (lane == 0)
      for (int j = 0; j < NOUT && j < P.N; ++j) store_c(P, m, j, s[j]);
  }
}

// ---- skinny wgrad: dW[Mo<=8, N] = G^T . X ; G [R, Mo] (a_sk = ldg, a_sm = 1), X [R, N] ----
// CTA = row chunk, thread = output column; partials -> ws[chunk][Mo][N], then splitk_reduce.
template <typename TA, typename TB, int MOUT>
__global__ void __launch_bounds__(kThreads) gemm_skinny_wgrad(GemmParams P) {
  const TA* __restrict__ G = reinterpret_cast<const TA*>(P.A);
  const TB* __restrict__ X = reinterpret_cast<const TB*>(P.B);
  const int64_t r0 = (int64_t)blockIdx.x * P.k_per_split;
  const int64_t r1 = min(P.K, r0 + P.k_per_split);
  for (int64_t n = threadIdx.x; n < P.N; n += kThreads) {
    float s[MOUT];
#pragma unroll
    for (int j = 0; j < MOUT; ++j) s[j] = 0.f;
    for (int64_t r = r0; r < r1; ++r) {
      float x = to_f32(X[r * P.b_sk + n]);
#pragma unroll
      for (int j = 0; j < MOUT; ++j)
        if (j < P.M) s[j] = fmaf(to_f32(G[r * P.a_sk + j]), x, s[j]);
    }
    for (int j = 0; j < MOUT && j < P.M; ++j)
      P.ws[((int64_t)blockIdx.x * P.M + j) * P.N + n] = s[j];
  }
}

template <typename TA, typename TB>
int run(GemmParams& P, cudaStream_t st) {
  const char* fn = "egnn_gemm";
  // skinny forward (2-class logits): N <= 8, both operands k-contiguous
  if (P.N <= 8 && P.a_sk == 1 && P.b_sk == 1 && P.split_k <= 1 && P.K <= 4096) {
    int64_t blocks = ceil_div(P.M, kThreads / 32);
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    if (P.N <= 2) gemm_skinny_fwd<TA, TB, 2><<<(unsigned)blocks, kThreads, 0, st>>>(P);
    else gemm_skinny_fwd<TA, TB, 8><<<(unsigned)blocks, kThreads, 0, st>>>(P);
    EGNN_LAUNCH_CHECK(fn);
    return 0;
  }
  // skinny wgrad: M <= 8, A m-contiguous, B n-contiguous, long reduction
  if (P.M <= 8 && P.a_sm == 1 && P.b_sn == 1 && P.K >= 4096 && P.ws) {
    int64_t chunks = kNumSMs * 4;
    int64_t per = ceil_div(P.K, chunks);
    chunks = ceil_div(P.K, per);
    P.k_per_split = per;
    P.split_k = (int)chunks;
    if (P.M <= 2) gemm_skinny_wgrad<TA, TB, 2><<<(unsigned)chunks, kThreads, 0, st>>>(P);
    else gemm_skinny_wgrad<TA, TB, 8><<<(unsigned)chunks, kThreads, 0, st>>>(P);
    EGNN_LAUNCH_CHECK(fn);
    splitk_reduce<<<(unsigned)ceil_div(P.M * P.N, kThreads), kThreads, 0, st>>>(P);
    EGNN_LAUNCH_CHECK(fn);
    return 0;
  }
  if (std::is_same<TA, float>::value && std::is_same<TB, float>::value && P.a_sk == 1 && P.b_sk == 1 &&
      P.split_k <= 1 && P.K % 4 == 0 && P.a_sm % 4 == 0 && P.b_sn % 4 == 0 &&
      (((uintptr_t)P.A | (uintptr_t)P.B) & 15) == 0 && P.K < (1 << 30)) {
    dim3 grid((unsigned)ceil_div(P.M, BM), (unsigned)ceil_div(P.N, BN), 1);
    gemm_simt_f32_kmajor<<<grid, kThreads, 0, st>>>(P);
    EGNN_LAUNCH_CHECK(fn);
    return 0;
  }
  int split = P.split_k < 1 ? 1 : P.split_k;
  int64_t per = ceil_div(ceil_div(P.K, split), BK) * BK;
  split = (int)ceil_div(P.K, per);
  P.k_per_split = per;
  P.split_k = split;
  dim3 grid((unsigned)ceil_div(P.M, BM), (unsigned)ceil_div(P.N, BN), (unsigned)split);
  if (std::is_same<TA, float>::value && std::is_same<TB, float>::value && P.a_sm == 1 && P.b_sn == 1 &&
      P.M % 4 == 0 && P.N % 4 == 0 && P.a_sk % 4 == 0 && P.b_sk % 4 == 0 &&
      (((uintptr_t)P.A | (uintptr_t)P.B) & 15) == 0 && (split == 1 || P.ws))
    gemm_simt_f32_mnmajor<<<grid, kThreads, 0, st>>>(P);
  else
    gemm_simt<TA, TB><<<grid, kThreads, 0, st>>>(P);
  EGNN_LAUNCH_CHECK(fn);
  if (split > 1) {
    splitk_reduce<<<(unsigned)ceil_div(P.M * P.N, kThreads), kThreads, 0, st>>>(P);
    EGNN_LAUNCH_CHECK(fn);
  }
  return 0;
}

}  // namespace

int gemm_tcgen05_dispatch(const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int c_dtype,
                          int64_t ld_c, int64_t M, int64_t N, int64_t K, const float* bias, int accumulate,
                          const int32_t* row_div_ptr, int64_t row_div_cols, cudaStream_t st, const void* addend = nullptr,
                          int64_t ld_add = 0, int64_t add_col0 = 0, float* stats = nullptr,
                          int64_t stats_cols = 0);  // gemm_tcgen05.cu
bool gemm_tf32x3_supported(int64_t lda, int64_t ldw, int64_t M, int64_t N, int64_t K, const void* A, const void* W);
size_t gemm_tf32x3_workspace_floats(int64_t N, int64_t K);
int gemm_tf32x3_dispatch(const void* A, int64_t lda, const void* W, int64_t ldw, void* C, int c_dtype, int64_t ld_c,
                         int64_t M, int64_t N, int64_t K, const float* bias, int accumulate,
                         const int32_t* row_div_ptr, int64_t row_div_cols, float* workspace, cudaStream_t st,
                         const void* addend = nullptr, int64_t ld_add = 0, int64_t add_col0 = 0, float* stats = nullptr,
                         int64_t stats_cols = 0);
bool wgrad_tf32x3_supported(const void* G, int64_t ldg, const void* X, int64_t ldx, int64_t M_rows, int64_t N_out,
                            int64_t K_in);
int wgrad_tf32x3_dispatch(const void* G, int64_t ldg, const void* X, int64_t ldx, float* dW, int64_t M_rows,
                          int64_t N_out, int64_t K_in, int accumulate, float* workspace, cudaStream_t st,
                          float* dst1 = nullptr, int64_t split = 0, int64_t valid = 0);
bool gemm_tcgen05_epilogue_supported(bool bias, bool row_div, bool accumulate, bool addend, bool stats, int64_t stats_cols,
                                     int64_t add_col0);
int64_t gemm_tcgen05_stats_parts(int64_t M);
bool gemm_tcgen05_supported(int64_t lda, int64_t ldb, int64_t ld_c, int64_t M, int64_t N, int64_t K,
                            const void* A, const void* B, const void* C);
size_t wgrad_tcgen05_workspace_floats(int64_t N_out, int64_t K_in);
bool wgrad_tcgen05_supported(const void* G, int64_t ldg, const void* X, int64_t ldx, int64_t M_rows, int64_t N_out,
                             int64_t K_in);
int wgrad_tcgen05_dispatch(const void* G, int64_t ldg, const void* X, int64_t ldx, float* dW, int64_t M_rows,
                           int64_t N_out, int64_t K_in, int accumulate, float* workspace, cudaStream_t st,
                           float* dst1 = nullptr, int64_t split = 0, int64_t valid = 0, const void* G2 = nullptr,
                           int64_t ldg2 = 0, int64_t N2 = 0, float* dst2 = nullptr);
}  // namespace egnn

using namespace egnn;

// C[M,N] = A[M,K] . W[N,K]^T on the tcgen05 kernel with the fused layer epilogues (bf16 operands, both K-major).
extern "C" int64_t egnn_linear_stats_parts(int64_t M) { return gemm_tcgen05_stats_parts(M); }

extern "C" size_t egnn_linear_tc_workspace_floats(int ab_dtype, int64_t N, int64_t K) {
  return ab_dtype == EGNN_F32 ? gemm_tf32x3_workspace_floats(N, K) : 0;
}

extern "C" int egnn_linear_tc(const void* A, int64_t lda, const void* W, int64_t ldw, void* C, int c_dtype, int64_t ld_c,
                              int64_t M, int64_t N, int64_t K, const float* bias, const int32_t* row_div_ptr,
                              int64_t row_div_cols, int accumulate, const void* addend, int64_t ld_addend,
                              int64_t addend_col0, float* colstats, int64_t colstats_cols, int ab_dtype,
                              float* workspace, void* stream) {
  const char* fn = "egnn_linear_tc";
  EGNN_REQUIRE(A && W && C, fn, "null pointer");
  EGNN_REQUIRE(ab_dtype == EGNN_BF16 || ab_dtype == EGNN_F32, fn, "bad operand dtype");
  if (ab_dtype == EGNN_F32) {   // fp32 operands: 3xTF32, same epilogues
    EGNN_REQUIRE(M >= 0 && N > 0 && K > 0 && ld_c >= N && workspace, fn, "bad shape / missing workspace");
    EGNN_REQUIRE(!addend || (ld_addend >= N - addend_col0 && addend_col0 >= 0 && addend_col0 < N), fn, "bad addend range");
    if (M == 0) return 0;
    if (!gemm_tf32x3_supported(lda, ldw, M, N, K, A, W))
      return fail(fn, "shape / alignment outside the 3xTF32 kernel (8 <= N <= 256, K % 4 == 0, 16-byte rows)");
    if (!gemm_tcgen05_epilogue_supported(bias != nullptr, row_div_ptr != nullptr, accumulate != 0, addend != nullptr,
                                         colstats != nullptr, colstats_cols, addend_col0))
      return fail(fn, "epilogue combination not compiled in");
    return gemm_tf32x3_dispatch(A, lda, W, ldw, C, c_dtype, ld_c, M, N, K, bias, accumulate, row_div_ptr, row_div_cols,
                                workspace, (cudaStream_t)stream, addend, ld_addend, addend_col0, colstats, colstats_cols);
  }
  EGNN_REQUIRE(M >= 0 && N > 0 && K > 0 && ld_c >= N, fn, "bad shape");
  EGNN_REQUIRE(c_dtype == EGNN_F32 || c_dtype == EGNN_BF16, fn, "bad output dtype");
  EGNN_REQUIRE(!addend || (ld_addend >= N - addend_col0 && addend_col0 >= 0 && addend_col0 < N), fn, "bad addend range");
  EGNN_REQUIRE(!colstats || (colstats_cols > 0 && colstats_cols <= N), fn, "bad colstats_cols");
  if (M == 0) return 0;
  if (!gemm_tcgen05_supported(lda, ldw, ld_c, M, N, K, A, W, C))
    return fail(fn, "shape / alignment outside the tcgen05 kernel (K <= 512, 8 <= N <= 256, 16-byte rows)");
  if (!gemm_tcgen05_epilogue_supported(bias != nullptr, row_div_ptr != nullptr, accumulate != 0, addend != nullptr,
                                       colstats != nullptr, colstats_cols, addend_col0))
    return fail(fn, "epilogue combination not compiled in");
  return gemm_tcgen05_dispatch(A, lda, W, ldw, C, c_dtype, ld_c, M, N, K, bias, accumulate, row_div_ptr, row_div_cols,
                               (cudaStream_t)stream, addend, ld_addend, addend_col0, colstats, colstats_cols);
}

// dW[N_out, K_in] = G[M, N_out]^T . X[M, K_in] on the tcgen05 wgrad kernel, the result written as two dense parameter
// gradients (see wgrad_reduce_split_kernel)
extern "C" size_t egnn_wgrad_tc_workspace_floats(int64_t N_out, int64_t K_in) {
  return wgrad_tcgen05_workspace_floats(N_out, K_in);
}
extern "C" int egnn_wgrad_tc(const void* G, int64_t ldg, const void* X, int64_t ldx, int64_t M, int64_t N_out,
                             int64_t K_in, float* dst0, float* dst1, int64_t split_col, int64_t valid_cols,
                             const void* G2, int64_t ldg2, int64_t N2, float* dst2, int dtype, float* workspace,
                             void* stream) {
  const char* fn = "egnn_wgrad_tc";
  EGNN_REQUIRE(G && X && dst0 && workspace, fn, "null pointer");
  EGNN_REQUIRE(M > 0 && split_col > 0 && split_col <= K_in && valid_cols > 0 && valid_cols <= split_col, fn, "bad split");
  EGNN_REQUIRE(dtype == EGNN_BF16 || dtype == EGNN_F32, fn, "bad operand dtype");
  if (dtype == EGNN_F32) {   // fp32 operands: 3xTF32 (one gradient operand per call)
    EGNN_REQUIRE(!G2, fn, "the fp32 kernel takes one gradient operand");
    if (!wgrad_tf32x3_supported(G, ldg, X, ldx, M, N_out, K_in))
      return fail(fn, "shape / alignment outside the 3xTF32 wgrad kernel (N_out <= 256, K_in <= 384, 16-byte rows)");
    return wgrad_tf32x3_dispatch(G, ldg, X, ldx, dst0, M, N_out, K_in, 0, workspace, (cudaStream_t)stream, dst1, split_col,
                                 valid_cols);
  }
  EGNN_REQUIRE(!G2 || (dst2 && N2 > 0 && N_out % 64 == 0 && ldg2 % 8 == 0 && (uintptr_t)G2 % 16 == 0 &&
                       2 * split_col == K_in),
               fn, "second gradient operand: N_out must be a multiple of 64, K_in = 2 * split_col, 16-byte rows");
  if (!wgrad_tcgen05_supported(G, ldg, X, ldx, M, N_out + (G2 ? N2 : 0), K_in))
    return fail(fn, "shape / alignment outside the tcgen05 wgrad kernel (N_out <= 256, K_in <= 384, 16-byte rows)");
  return wgrad_tcgen05_dispatch(G, ldg, X, ldx, dst0, M, N_out, K_in, 0, workspace, (cudaStream_t)stream, dst1, split_col,
                                valid_cols, G2, ldg2, N2, dst2);
}

constexpr int64_t kTf32MinRows = 1024;   // fp32 products with fewer rows stay on the FFMA kernel

extern "C" size_t egnn_gemm_workspace_floats(int64_t M, int64_t N, int64_t K, int split_k) {
  (void)K;
  size_t need = split_k > 1 ? (size_t)split_k * (size_t)M * (size_t)N : 0;
  if (M <= 8) need = std::max(need, (size_t)(kNumSMs * 4 + 1) * (size_t)M * (size_t)N);  // skinny wgrad chunks
  if (M <= 256 && N <= 384) need = std::max(need, wgrad_tcgen05_workspace_floats(M, N));     // tcgen05 wgrad
  if (M >= kTf32MinRows && N <= 256) need = std::max(need, gemm_tf32x3_workspace_floats(N, K));  // 3xTF32 weight split
  return need;
}

extern "C" int egnn_gemm(const void* A, int a_dtype, int64_t a_sm, int64_t a_sk, const void* B,
                         int b_dtype, int64_t b_sk, int64_t b_sn, void* C, int c_dtype, int64_t ld_c,
                         int64_t M, int64_t N, int64_t K, const float* bias, const int32_t* row_div_ptr,
                         int64_t row_div_cols, int accumulate, int split_k, float* workspace, int impl,
                         void* stream) {
  const char* fn = "egnn_gemm";
  EGNN_REQUIRE(A && B && C, fn, "null pointer");
  EGNN_REQUIRE(M >= 0 && N > 0 && K > 0, fn, "bad shape");
  EGNN_REQUIRE(ld_c >= N, fn, "ld_c < N");
  EGNN_REQUIRE(split_k <= 1 || workspace, fn, "split_k > 1 needs a workspace");
  EGNN_REQUIRE(impl >= 0 && impl <= 2, fn, "bad impl");
  if (M == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const bool both_bf16 = a_dtype == EGNN_BF16 && b_dtype == EGNN_BF16;
  // tensor-core forward / dgrad: both operands contiguous along the contraction
  const bool tc_tn = both_bf16 && a_sk == 1 && b_sk == 1 && N >= 8 &&
                     gemm_tcgen05_supported(a_sm, b_sn, ld_c, M, N, K, A, B, C);
  // tensor-core wgrad: both operands contiguous along the non-contracted dimension, long reduction
  const bool tc_wg = both_bf16 && a_sm == 1 && b_sn == 1 && c_dtype == EGNN_F32 && ld_c == N && !bias &&
                     !row_div_ptr && workspace && M >= 8 && K >= 1024 &&
                     wgrad_tcgen05_supported(A, a_sk, B, b_sk, K, M, N);
  // fp32 operands, both contiguous along the contraction: 3xTF32 on the tensor cores (fp32-exact to ~1e-6)
  const bool both_f32 = a_dtype == EGNN_F32 && b_dtype == EGNN_F32;
  const bool tc_f32 = both_f32 && a_sk == 1 && b_sk == 1 && workspace && split_k <= 1 && M >= kTf32MinRows &&
                      gemm_tf32x3_supported(a_sm, b_sn, M, N, K, A, B);
  if (impl != 1 && tc_f32)
    return gemm_tf32x3_dispatch(A, a_sm, B, b_sn, C, c_dtype, ld_c, M, N, K, bias, accumulate, row_div_ptr, row_div_cols,
                                workspace, st);
  // fp32 weight gradient (both operands contiguous along the non-contracted dimension, long reduction): 3xTF32
  const bool tc_wg32 = both_f32 && a_sm == 1 && b_sn == 1 && c_dtype == EGNN_F32 && ld_c == N && !bias && !row_div_ptr &&
                       workspace && M >= 8 && K >= kTf32MinRows && wgrad_tf32x3_supported(A, a_sk, B, b_sk, K, M, N);
  if (impl != 1 && tc_wg32)
    return wgrad_tf32x3_dispatch(A, a_sk, B, b_sk, (float*)C, K, M, N, accumulate, workspace, st);
  if (impl == 2 && !(tc_tn || tc_wg)) return fail(fn, "shape/dtype/layout not supported by the tcgen05 path");
  if (impl != 1 && tc_tn)
    return gemm_tcgen05_dispatch(A, a_sm, B, b_sn, C, c_dtype, ld_c, M, N, K, bias, accumulate, row_div_ptr, row_div_cols, st);
  if (impl != 1 && tc_wg)
    return wgrad_tcgen05_dispatch(A, a_sk, B, b_sk, (float*)C, K, M, N, accumulate, workspace, st);
  if (both_bf16 && (M >= 65536 || K >= 65536)) {  // a large bf16 product that misses the tensor-core path: say so once
    static bool warned = false;
    if (!warned) {
      warned = true;
      fprintf(stderr, "[egnn_b200] note: bf16 GEMM M=%lld N=%lld K=%lld (a_sm=%lld a_sk=%lld) runs on the SIMT kernel "
                      "(shape/layout outside the tcgen05 kernels)\n",
              (long long)M, (long long)N, (long long)K, (long long)a_sm, (long long)a_sk);
    }
  }
  GemmParams P;
  P.A = A; P.B = B; P.C = C; P.bias = bias; P.row_div_ptr = row_div_ptr; P.row_div_cols = row_div_cols; P.ws = workspace;
  P.a_sm = a_sm; P.a_sk = a_sk; P.b_sk = b_sk; P.b_sn = b_sn; P.ld_c = ld_c;
  P.M = M; P.N = N; P.K = K; P.k_per_split = K;
  P.c_dtype = c_dtype; P.accumulate = accumulate; P.split_k = split_k;
  if (a_dtype == EGNN_F32 && b_dtype == EGNN_F32) return run<float, float>(P, st);
  if (a_dtype == EGNN_BF16 && b_dtype == EGNN_BF16) return run<__nv_bfloat16, __nv_bfloat16>(P, st);
  if (a_dtype == EGNN_BF16 && b_dtype == EGNN_F32) return run<__nv_bfloat16, float>(P, st);
  if (a_dtype == EGNN_F32 && b_dtype == EGNN_BF16) return run<float, __nv_bfloat16>(P, st);
  return fail(fn, "unsupported dtype");
}
