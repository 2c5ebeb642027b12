// K4/K5 -- fused GAT attention (PyG GATConv) over the self-loop CSR/CSC.
//
// Forward: one warp per destination row.  Lanes < H own a head for the segment softmax
// (edge score -> running max -> sum of exp -> alpha, alpha stored in CSR order for the
// backward); then all lanes own output features and accumulate alpha * xs[src] over the
// row's edges in stored order (rounded product, sequential add: the oracle's scatter order).
// Backward: a destination pass (per (row, head): g, sum alpha*g, dpre, da_d) and a source
// pass over the CSC view (da_s, dxs), addressing the per-edge arrays through csc_pos.
// PyG launches ~20 ATen kernels, 3 scatters and 5 gathers of [E2,H(,C)] for the same work.
#include "common.cuh"

namespace egnn {
namespace {

constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads) gat_scores_kernel(const float* __restrict__ xs, int64_t n_rows,
                                                              int H, int C, const float* __restrict__ att_src,
                                                              const float* __restrict__ att_dst,
                                                              float* __restrict__ a_s, float* __restrict__ a_d) {
  int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (t >= n_rows * H) return;
  int h = (int)(t % H);
  const float* x = xs + t * C;  // (n*H + h)*C
  float s = 0.f, d = 0.f;
  for (int c = 0; c < C; ++c) {
    float v = x[c];
    s = __fadd_rn(s, __fmul_rn(v, att_src[h * C + c]));
    d = __fadd_rn(d, __fmul_rn(v, att_dst[h * C + c]));
  }
  a_s[t] = s;
  a_d[t] = d;
}

__device__ __forceinline__ float leaky(float v, float slope) { return v > 0.f ? v : v * slope; }

__global__ void __launch_bounds__(kThreads) gat_fwd_kernel(const int* __restrict__ ptr, const int* __restrict__ src,
                                                           const float* __restrict__ xs,
                                                           const float* __restrict__ a_s,
                                                           const float* __restrict__ a_d, float slope, int H,
                                                           int C, int concat, const float* __restrict__ bias,
                                                           float* alpha, float* __restrict__ out,
                                                           int64_t n_rows) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * kThreads + threadIdx.x) >> 5;
  if (row >= n_rows) return;
  const int p0 = ptr[row], p1 = ptr[row + 1];
  const int F = H * C;
  for (int h = lane; h < H; h += 32) {
    const float ad = a_d[row * H + h];
    float mx = -INFINITY;
    for (int p = p0; p < p1; ++p) mx = fmaxf(mx, leaky(a_s[(int64_t)src[p] * H + h] + ad, slope));
    float den = 0.f;
    for (int p = p0; p < p1; ++p) {
      float e = expf(leaky(a_s[(int64_t)src[p] * H + h] + ad, slope) - mx);
      alpha[(int64_t)p * H + h] = e;
      den = __fadd_rn(den, e);
    }
    den = den + 1e-16f;
    for (int p = p0; p < p1; ++p) alpha[(int64_t)p * H + h] = __fdiv_rn(alpha[(int64_t)p * H + h], den);
  }
  __syncwarp();
  if (concat) {
    for (int f = lane; f < F; f += 32) {
      const int h = f / C;
      float acc = 0.f;
      for (int p = p0; p < p1; ++p)
        acc = __fadd_rn(acc, __fmul_rn(alpha[(int64_t)p * H + h], xs[(int64_t)src[p] * F + f]));
      out[row * F + f] = acc + (bias ? bias[f] : 0.f);
    }
  } else {
    for (int c = lane; c < C; c += 32) {
      float tot = 0.f;
      for (int h = 0; h < H; ++h) {
        float acc = 0.f;
        for (int p = p0; p < p1; ++p)
          acc = __fadd_rn(acc, __fmul_rn(alpha[(int64_t)p * H + h], xs[(int64_t)src[p] * F + h * C + c]));
        tot = __fadd_rn(tot, acc);
      }
      out[row * C + c] = __fdiv_rn(tot, (float)H) + (bias ? bias[c] : 0.f);
    }
  }
}

// destination pass: thread per (row, head)
__global__ void __launch_bounds__(kThreads) gat_bwd_dst_kernel(
    const int* __restrict__ ptr, const int* __restrict__ src, const float* __restrict__ xs,
    const float* __restrict__ a_s, const float* __restrict__ a_d, const float* __restrict__ alpha,
    const float* __restrict__ dout, float slope, int H, int C, int concat, float* dpre,
    float* __restrict__ da_d, int64_t n_rows) {
  int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (t >= n_rows * H) return;
  const int64_t row = t / H;
  const int h = (int)(t - row * H);
  const int F = H * C;
  const int p0 = ptr[row], p1 = ptr[row + 1];
  const float* dorow = concat ? dout + row * F + h * C : dout + row * C;
  const float dscale = concat ? 1.f : 1.f / (float)H;
  float s = 0.f;
  for (int p = p0; p < p1; ++p) {
    const float* xj = xs + (int64_t)src[p] * F + h * C;
    float g = 0.f;
    for (int c = 0; c < C; ++c) g = fmaf(dorow[c] * dscale, xj[c], g);
    dpre[(int64_t)p * H + h] = g;
    s = fmaf(alpha[(int64_t)p * H + h], g, s);
  }
  const float ad = a_d[t];
  float acc = 0.f;
  for (int p = p0; p < p1; ++p) {
    float g = dpre[(int64_t)p * H + h];
    float de = alpha[(int64_t)p * H + h] * (g - s);
    float pre = a_s[(int64_t)src[p] * H + h] + ad;
    float dp = pre > 0.f ? de : de * slope;
    dpre[(int64_t)p * H + h] = dp;
    acc += dp;
  }
  da_d[t] = acc;
}

// source pass: warp per source row over the CSC view
__global__ void __launch_bounds__(kThreads) gat_bwd_src_kernel(
    const int* __restrict__ ptr, const int* __restrict__ dst, const int* __restrict__ pos,
    const float* __restrict__ alpha, const float* __restrict__ dpre, const float* __restrict__ dout,
    const float* __restrict__ da_d, const float* __restrict__ att_src, const float* __restrict__ att_dst, int H,
    int C, int concat, float* __restrict__ dxs, float* da_s, int64_t n_rows) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * kThreads + threadIdx.x) >> 5;
  if (row >= n_rows) return;
  const int q0 = ptr[row], q1 = ptr[row + 1];
  const int F = H * C;
  for (int h = lane; h < H; h += 32) {
    float acc = 0.f;
    for (int q = q0; q < q1; ++q) acc += dpre[(int64_t)pos[q] * H + h];
    da_s[row * H + h] = acc;
  }
  __syncwarp();
  const float dscale = concat ? 1.f : 1.f / (float)H;
  for (int f = lane; f < F; f += 32) {
    const int h = f / C, c = f - h * C;
    float acc = 0.f;
    for (int q = q0; q < q1; ++q) {
      const float d = concat ? dout[(int64_t)dst[q] * F + f] : dout[(int64_t)dst[q] * C + c] * dscale;
      acc = fmaf(alpha[(int64_t)pos[q] * H + h], d, acc);
    }
    acc = fmaf(da_s[row * H + h], att_src[f], acc);
    acc = fmaf(da_d[row * H + h], att_dst[f], acc);
    dxs[row * F + f] = acc;
  }
}

}  // namespace
}  // namespace egnn

using namespace egnn;

extern "C" int egnn_gat_scores(const float* xs, int64_t n_rows, int H, int C, const float* att_src,
                               const float* att_dst, float* a_s, float* a_d, void* stream) {
  const char* fn = "egnn_gat_scores";
  EGNN_REQUIRE(xs && att_src && att_dst && a_s && a_d && H > 0 && C > 0, fn, "bad arguments");
  if (n_rows == 0) return 0;
  gat_scores_kernel<<<(unsigned)ceil_div(n_rows * H, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
      xs, n_rows, H, C, att_src, att_dst, a_s, a_d);
  EGNN_LAUNCH_CHECK(fn);
  return 0;
}

extern "C" int egnn_gat_fwd(const int32_t* csr_ptr, const int32_t* csr_src, const float* xs, const float* a_s,
                            const float* a_d, float negative_slope, int H, int C, int concat,
                            const float* bias, float* alpha, float* out, int64_t n_rows, void* stream) {
  const char* fn = "egnn_gat_fwd";
  EGNN_REQUIRE(csr_ptr && csr_src && xs && a_s && a_d && alpha && out && H > 0 && C > 0, fn, "bad arguments");
  if (n_rows == 0) return 0;
  gat_fwd_kernel<<<(unsigned)ceil_div(n_rows * 32, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
      csr_ptr, csr_src, xs, a_s, a_d, negative_slope, H, C, concat, bias, alpha, out, n_rows);
  EGNN_LAUNCH_CHECK(fn);
  return 0;
}

extern "C" int egnn_gat_bwd_dst(const int32_t* csr_ptr, const int32_t* csr_src, const float* xs,
                                const float* a_s, const float* a_d, const float* alpha, const float* dout,
                                float negative_slope, int H, int C, int concat, float* dpre, float* da_d,
                                int64_t n_rows, void* stream) {
  const char* fn = "egnn_gat_bwd_dst";
  EGNN_REQUIRE(csr_ptr && csr_src && xs && a_s && a_d && alpha && dout && dpre && da_d, fn, "null pointer");
  if (n_rows == 0) return 0;
  gat_bwd_dst_kernel<<<(unsigned)ceil_div(n_rows * H, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
      csr_ptr, csr_src, xs, a_s, a_d, alpha, dout, negative_slope, H, C, concat, dpre, da_d, n_rows);
  EGNN_LAUNCH_CHECK(fn);
  return 0;
}

extern "C" int egnn_gat_bwd_src(const int32_t* csc_ptr, const int32_t* csc_dst, const int32_t* csc_pos,
                                const float* alpha, const float* dpre, const float* dout, const float* da_d,
                                const float* att_src, const float* att_dst, int H, int C, int concat,
                                float* dxs, float* da_s, int64_t n_rows, void* stream) {
  const char* fn = "egnn_gat_bwd_src";
  EGNN_REQUIRE(csc_ptr && csc_dst && csc_pos && alpha && dpre && dout && da_d && att_src && att_dst && dxs &&
                   da_s,
               fn, "null pointer");
  if (n_rows == 0) return 0;
  gat_bwd_src_kernel<<<(unsigned)ceil_div(n_rows * 32, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
      csc_ptr, csc_dst, csc_pos, alpha, dpre, dout, da_d, att_src, att_dst, H, C, concat, dxs, da_s, n_rows);
  EGNN_LAUNCH_CHECK(fn);
  return 0;
}
