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
constexpr int kGatLanes = 8;  // lanes per row in the forward and the source pass (4 rows per warp)

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

constexpr int kGatLongRow = 64;  // rows with more entries are processed by their whole warp (entry-parallel)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One hub row, all 32 lanes of the warp (H <= 8).  Softmax sweeps take 32 entries at a time; the aggregation splits
// the ENTRIES over the lanes (4 entry slots x 8 feature lanes with 16-byte gathers for the 4 x 8 layer, 32 entry
// slots per feature otherwise) and combines the partial sums with shuffles, so a 481-entry row costs ~15 dependent
// round trips instead of ~180.  The summation order differs from the short-row path; results agree to rounding.
template <int HT>
__device__ void gat_fwd_long_row(int64_t row, int p0, int p1, const int* __restrict__ src, const float* __restrict__ xs,
                                 const float* __restrict__ a_s, const float* __restrict__ a_d, float slope, int H, int C,
                                 int concat, const float* __restrict__ bias, float* alpha, float* __restrict__ out) {
  const int Hc = HT > 0 ? HT : H;   // compile-time head count (HT > 0): the per-head loops below fold
  constexpr int kHMax = HT > 0 ? HT : 8;
  const int wl = threadIdx.x & 31, F = Hc * C;
  float ad[kHMax], mx[kHMax], den[kHMax];
#pragma unroll
  for (int h = 0; h < kHMax; ++h) {
    ad[h] = h < Hc ? __ldg(a_d + row * Hc + h) : 0.f;
    mx[h] = -INFINITY;
    den[h] = 0.f;
  }
#pragma unroll 4
  for (int base = p0; base < p1; base += 32) {
    const int p = base + wl;
    const bool valid = p < p1;
    const int64_t sj = valid ? __ldg(src + p) : 0;
#pragma unroll
    for (int h = 0; h < kHMax; ++h)
      if (h < Hc && valid) mx[h] = fmaxf(mx[h], leaky(__ldg(a_s + sj * Hc + h) + ad[h], slope));
  }
#pragma unroll
  for (int h = 0; h < kHMax; ++h)
    if (h < Hc)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], o));
#pragma unroll 4
  for (int base = p0; base < p1; base += 32) {
    const int p = base + wl;
    const bool valid = p < p1;
    const int64_t sj = valid ? __ldg(src + p) : 0;
#pragma unroll
    for (int h = 0; h < kHMax; ++h)
      if (h < Hc && valid) {
        const float ex = expf(leaky(__ldg(a_s + sj * Hc + h) + ad[h], slope) - mx[h]);
        den[h] += ex;
        alpha[(int64_t)p * Hc + h] = ex;
      }
  }
#pragma unroll
  for (int h = 0; h < kHMax; ++h)
    if (h < Hc) den[h] = warp_sum(den[h]) + 1e-16f;
#pragma unroll 4
  for (int base = p0; base < p1; base += 32) {
    const int p = base + wl;
    if (p < p1) {
#pragma unroll
      for (int h = 0; h < kHMax; ++h)
        if (h < Hc) alpha[(int64_t)p * Hc + h] = __fdiv_rn(alpha[(int64_t)p * Hc + h], den[h]);
    }
  }
  __syncwarp();
  constexpr int kB = 8;
  if (concat && F == 32 && (C & 3) == 0) {
    const int fl = wl & 7, es = wl >> 3, f0 = 4 * fl, h = f0 / C;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int p = p0 + es; p < p1; p += 4 * kB) {
      float a[kB];
      float4 x[kB];
#pragma unroll
      for (int b = 0; b < kB; ++b) {
        const int pp = p + 4 * b;
        const bool ok = pp < p1;
        a[b] = ok ? alpha[(int64_t)pp * Hc + h] : 0.f;
        x[b] = ok ? __ldg(reinterpret_cast<const float4*>(xs + (int64_t)__ldg(src + pp) * F + f0))
                  : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int b = 0; b < kB; ++b) {
        acc.x = __fadd_rn(acc.x, __fmul_rn(a[b], x[b].x));
        acc.y = __fadd_rn(acc.y, __fmul_rn(a[b], x[b].y));
        acc.z = __fadd_rn(acc.z, __fmul_rn(a[b], x[b].z));
        acc.w = __fadd_rn(acc.w, __fmul_rn(a[b], x[b].w));
      }
    }
#pragma unroll
    for (int o = 8; o <= 16; o <<= 1) {
      acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
      acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
      acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o);
      acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
    }
    if (es == 0) {
      float* o = out + row * F + f0;
      o[0] = acc.x + (bias ? bias[f0] : 0.f);
      o[1] = acc.y + (bias ? bias[f0 + 1] : 0.f);
      o[2] = acc.z + (bias ? bias[f0 + 2] : 0.f);
      o[3] = acc.w + (bias ? bias[f0 + 3] : 0.f);
    }
  } else {
    const int n_out = concat ? F : C;
    for (int oc = 0; oc < n_out; ++oc) {
      const int h_lo = concat ? oc / C : 0, h_hi = concat ? h_lo + 1 : Hc;
      float tot = 0.f;
      for (int h = h_lo; h < h_hi; ++h) {
        const int f = concat ? oc : h * C + oc;
        float part = 0.f;
#pragma unroll 4
        for (int p = p0 + wl; p < p1; p += 32)
          part = __fadd_rn(part, __fmul_rn(alpha[(int64_t)p * Hc + h], __ldg(xs + (int64_t)__ldg(src + p) * F + f)));
        tot = __fadd_rn(tot, warp_sum(part));
      }
      if (wl == 0) out[row * n_out + oc] = (concat ? tot : __fdiv_rn(tot, (float)Hc)) + (bias ? bias[oc] : 0.f);
    }
  }
}

template <int HT>
__device__ void gat_bwd_src_long_row(int64_t row, int q0, int q1, const int* __restrict__ dst, const int* __restrict__ pos,
                                     const float* __restrict__ alpha, const float* __restrict__ dpre,
                                     const float* __restrict__ dout, const float* __restrict__ da_d,
                                     const float* __restrict__ att_src, const float* __restrict__ att_dst, int H, int C,
                                     int concat, float* __restrict__ dxs, float* da_s) {
  const int Hc = HT > 0 ? HT : H;   // compile-time head count (HT > 0): the per-head loops below fold
  constexpr int kHMax = HT > 0 ? HT : 8;
  const int wl = threadIdx.x & 31, F = Hc * C;
  float das[kHMax];
#pragma unroll
  for (int h = 0; h < kHMax; ++h) das[h] = 0.f;
#pragma unroll 4
  for (int base = q0; base < q1; base += 32) {
    const int q = base + wl;
    if (q < q1) {
      const int64_t pq = __ldg(pos + q);
#pragma unroll
      for (int h = 0; h < kHMax; ++h)
        if (h < Hc) das[h] += __ldg(dpre + pq * Hc + h);
    }
  }
#pragma unroll
  for (int h = 0; h < kHMax; ++h)
    if (h < Hc) {
      das[h] = warp_sum(das[h]);
      if (wl == h) da_s[row * Hc + h] = das[h];
    }
  const float dscale = concat ? 1.f : 1.f / (float)Hc;
  constexpr int kB = 8;
  if (concat && F == 32 && (C & 3) == 0) {
    const int fl = wl & 7, es = wl >> 3, f0 = 4 * fl, h = f0 / C;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int q = q0 + es; q < q1; q += 4 * kB) {
      float a[kB];
      float4 d[kB];
#pragma unroll
      for (int b = 0; b < kB; ++b) {
        const int qq = q + 4 * b;
        const bool ok = qq < q1;
        a[b] = ok ? __ldg(alpha + (int64_t)__ldg(pos + qq) * Hc + h) : 0.f;
        d[b] = ok ? __ldg(reinterpret_cast<const float4*>(dout + (int64_t)__ldg(dst + qq) * F + f0))
                  : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int b = 0; b < kB; ++b) {
        acc.x = fmaf(a[b], d[b].x, acc.x);
        acc.y = fmaf(a[b], d[b].y, acc.y);
        acc.z = fmaf(a[b], d[b].z, acc.z);
        acc.w = fmaf(a[b], d[b].w, acc.w);
      }
    }
#pragma unroll
    for (int o = 8; o <= 16; o <<= 1) {
      acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
      acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
      acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o);
      acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
    }
    if (es == 0) {
      float das_h = 0.f;
#pragma unroll
      for (int k = 0; k < kHMax; ++k) das_h = k == h ? das[k] : das_h;
      const float dad = da_d[row * Hc + h];
      float r[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        r[i] = fmaf(das_h, att_src[f0 + i], r[i]);
        r[i] = fmaf(dad, att_dst[f0 + i], r[i]);
        dxs[row * F + f0 + i] = r[i];
      }
    }
  } else {
    for (int f = 0; f < F; ++f) {
      const int h = f / C, c = f - h * C;
      float part = 0.f;
#pragma unroll 4
      for (int q = q0 + wl; q < q1; q += 32) {
        const int64_t dq = __ldg(dst + q);
        const float d = concat ? __ldg(dout + dq * F + f) : __ldg(dout + dq * C + c) * dscale;
        part = fmaf(__ldg(alpha + (int64_t)__ldg(pos + q) * Hc + h), d, part);
      }
      part = warp_sum(part);
      if (wl == 0) {
        float das_h = 0.f;
#pragma unroll
        for (int k = 0; k < kHMax; ++k) das_h = k == h ? das[k] : das_h;
        float acc = fmaf(das_h, att_src[f], part);
        acc = fmaf(da_d[row * Hc + h], att_dst[f], acc);
        dxs[row * F + f] = acc;
      }
    }
  }
}

template <int G, int HT>
__global__ void __launch_bounds__(kThreads) gat_fwd_kernel(const int* __restrict__ ptr, const int* __restrict__ src,
                                                           const float* __restrict__ xs,
                                                           const float* __restrict__ a_s,
                                                           const float* __restrict__ a_d, float slope, int H,
                                                           int C, int concat, const float* __restrict__ bias,
                                                           float* alpha, float* __restrict__ out,
                                                           int64_t n_rows) {
  const int Hc = HT > 0 ? HT : H;   // compile-time head count (HT > 0): the per-head loops below fold
  // G lanes per destination row (32 / G rows per warp): the mean row of the self-loop graph has 2-3 entries, so a
  // whole warp per row left 90 % of the lanes idle and 203 769 warps queued behind eight dependent round trips
  // each.  Rows past the end keep their lanes in the warp-wide shuffles with an empty range.
  const int lane = threadIdx.x % G;
  int64_t row = ((int64_t)blockIdx.x * kThreads + threadIdx.x) / G;
  const bool in_range = row < n_rows;
  if (!in_range) row = 0;
  int p0 = in_range ? ptr[row] : 0, p1 = in_range ? ptr[row + 1] : 0;
  // hub rows: the whole warp takes them one after the other, then the groups go on with their short rows
  const bool is_long = G < 32 && Hc <= 8 && in_range && p1 - p0 > kGatLongRow;
  for (unsigned lm = __ballot_sync(0xffffffffu, is_long && lane == 0); lm; lm &= lm - 1) {
    const int sl = __ffs(lm) - 1;
    gat_fwd_long_row<HT>(__shfl_sync(0xffffffffu, row, sl), __shfl_sync(0xffffffffu, p0, sl),
                     __shfl_sync(0xffffffffu, p1, sl), src, xs, a_s, a_d, slope, Hc, C, concat, bias, alpha, out);
  }
  const bool live = in_range && !is_long;
  if (!live) p0 = p1 = 0;
  const int F = Hc * C;
  // segment softmax with ONE LANE PER ENTRY (32 entries per sweep; the usual row is one sweep): all score
  // gathers of a row are issued together, max / sum are warp reductions, nothing walks the row serially
  constexpr int kHMax = HT > 0 ? HT : 8;
  if (Hc <= kHMax) {
    float ad[kHMax], mx[kHMax], den[kHMax], e0[kHMax];
#pragma unroll
    for (int h = 0; h < kHMax; ++h) {
      ad[h] = (h < Hc && live) ? __ldg(a_d + row * Hc + h) : 0.f;
      mx[h] = -INFINITY;
      den[h] = 0.f;
      e0[h] = 0.f;
    }
    const bool single = p1 - p0 <= G;
#pragma unroll 4
    for (int base = p0; base < p1; base += G) {  // pass A: scores and per-head max
      const int p = base + lane;
      const bool valid = p < p1;
      const int64_t sj = valid ? __ldg(src + p) : 0;
#pragma unroll
      for (int h = 0; h < kHMax; ++h)
        if (h < Hc) {
          const float e = leaky(__ldg(a_s + sj * Hc + h) + ad[h], slope);
          if (base == p0) e0[h] = e;
          if (valid) mx[h] = fmaxf(mx[h], e);
        }
    }
#pragma unroll
    for (int h = 0; h < kHMax; ++h)
      if (h < Hc)
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], o));
#pragma unroll 4
    for (int base = p0; base < p1; base += G) {  // pass B: exp and per-head sum (unnormalised alpha parked)
      const int p = base + lane;
      const bool valid = p < p1;
      const int64_t sj = (valid && !single) ? __ldg(src + p) : 0;
#pragma unroll
      for (int h = 0; h < kHMax; ++h)
        if (h < Hc) {
          const float e = single ? e0[h] : leaky(__ldg(a_s + sj * Hc + h) + ad[h], slope);
          const float ex = valid ? expf(e - mx[h]) : 0.f;
          den[h] += ex;
          if (valid) alpha[(int64_t)p * Hc + h] = ex;
          if (single) e0[h] = ex;
        }
    }
#pragma unroll
    for (int h = 0; h < kHMax; ++h)
      if (h < Hc) {
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) den[h] += __shfl_xor_sync(0xffffffffu, den[h], o);
        den[h] += 1e-16f;
      }
#pragma unroll 4
    for (int base = p0; base < p1; base += G) {  // pass C: normalise
      const int p = base + lane;
      if (p < p1) {
#pragma unroll
        for (int h = 0; h < kHMax; ++h)
          if (h < Hc) alpha[(int64_t)p * Hc + h] = __fdiv_rn(single ? e0[h] : alpha[(int64_t)p * Hc + h], den[h]);
      }
    }
  } else {
  for (int h = lane; live && h < Hc; h += G) {
    const float ad = a_d[row * Hc + h];
    float mx = -INFINITY;
    int p = p0;
    for (; p + 4 <= p1; p += 4) {  // four (src -> a_s) chains in flight: a hub row is hundreds of entries long
      float v[4];
#pragma unroll
      for (int b = 0; b < 4; ++b) v[b] = __ldg(a_s + (int64_t)__ldg(src + p + b) * Hc + h);
#pragma unroll
      for (int b = 0; b < 4; ++b) mx = fmaxf(mx, leaky(v[b] + ad, slope));
    }
    for (; p < p1; ++p) mx = fmaxf(mx, leaky(__ldg(a_s + (int64_t)__ldg(src + p) * Hc + h) + ad, slope));
    float den = 0.f;
    p = p0;
    for (; p + 4 <= p1; p += 4) {
      float v[4];
#pragma unroll
      for (int b = 0; b < 4; ++b) v[b] = __ldg(a_s + (int64_t)__ldg(src + p + b) * Hc + h);
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const float e = expf(leaky(v[b] + ad, slope) - mx);
        alpha[(int64_t)(p + b) * Hc + h] = e;
        den = __fadd_rn(den, e);
      }
    }
    for (; p < p1; ++p) {
      const float e = expf(leaky(__ldg(a_s + (int64_t)__ldg(src + p) * Hc + h) + ad, slope) - mx);
      alpha[(int64_t)p * Hc + h] = e;
      den = __fadd_rn(den, e);
    }
    den = den + 1e-16f;
    for (p = p0; p < p1; ++p) alpha[(int64_t)p * Hc + h] = __fdiv_rn(alpha[(int64_t)p * Hc + h], den);
  }
  }
  __syncwarp();
  // Aggregation.  A hub row is hundreds of entries long and its lanes walk it alone, so the row's time is
  // (entries / batch) dependent round trips: eight entries in flight per batch, and -- when every lane can own four
  // contiguous features of one head (F = 4 G, C % 4 == 0: the 4 x 8 hidden layer) -- one 16-byte gather and one
  // alpha per entry instead of four scalar walks over the row.  Products are rounded, adds sequential in stored
  // order (the oracle's scatter order) in every variant.
  constexpr int kB = 4;   // entries in flight per batch: the usual row has 2-3 entries, 8 only cost registers
  if (concat && F == 4 * G && (C & 3) == 0) {
    const int f0 = 4 * lane, h = f0 / C;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int p = p0; p < p1; p += kB) {
      float a[kB];
      float4 x[kB];
#pragma unroll
      for (int b = 0; b < kB; ++b) {
        const bool ok = p + b < p1;
        a[b] = ok ? alpha[(int64_t)(p + b) * Hc + h] : 0.f;
        x[b] = ok ? __ldg(reinterpret_cast<const float4*>(xs + (int64_t)__ldg(src + p + b) * F + f0))
                  : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int b = 0; b < kB; ++b) {
        if (p + b < p1) {
          acc.x = __fadd_rn(acc.x, __fmul_rn(a[b], x[b].x));
          acc.y = __fadd_rn(acc.y, __fmul_rn(a[b], x[b].y));
          acc.z = __fadd_rn(acc.z, __fmul_rn(a[b], x[b].z));
          acc.w = __fadd_rn(acc.w, __fmul_rn(a[b], x[b].w));
        }
      }
    }
    if (live) {
      float* o = out + row * F + f0;
      o[0] = acc.x + (bias ? bias[f0] : 0.f);
      o[1] = acc.y + (bias ? bias[f0 + 1] : 0.f);
      o[2] = acc.z + (bias ? bias[f0 + 2] : 0.f);
      o[3] = acc.w + (bias ? bias[f0 + 3] : 0.f);
    }
  } else if (concat) {
    for (int f = lane; live && f < F; f += G) {
      const int h = f / C;
      float acc = 0.f;
      for (int p = p0; p < p1; p += kB) {
        float a[kB], x[kB];
#pragma unroll
        for (int b = 0; b < kB; ++b) {
          const bool ok = p + b < p1;
          a[b] = ok ? alpha[(int64_t)(p + b) * Hc + h] : 0.f;
          x[b] = ok ? __ldg(xs + (int64_t)__ldg(src + p + b) * F + f) : 0.f;
        }
#pragma unroll
        for (int b = 0; b < kB; ++b)
          if (p + b < p1) acc = __fadd_rn(acc, __fmul_rn(a[b], x[b]));
      }
      out[row * F + f] = acc + (bias ? bias[f] : 0.f);
    }
  } else {
    for (int c = lane; live && c < C; c += G) {
      float tot = 0.f;
      for (int h = 0; h < Hc; ++h) {
        float acc = 0.f;
        for (int p = p0; p < p1; p += kB) {
          float a[kB], x[kB];
#pragma unroll
          for (int b = 0; b < kB; ++b) {
            const bool ok = p + b < p1;
            a[b] = ok ? alpha[(int64_t)(p + b) * Hc + h] : 0.f;
            x[b] = ok ? __ldg(xs + (int64_t)__ldg(src + p + b) * F + h * C + c) : 0.f;
          }
#pragma unroll
          for (int b = 0; b < kB; ++b)
            if (p + b < p1) acc = __fadd_rn(acc, __fmul_rn(a[b], x[b]));
        }
        tot = __fadd_rn(tot, acc);
      }
      out[row * C + c] = __fdiv_rn(tot, (float)Hc) + (bias ? bias[c] : 0.f);
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
  const bool valid = t < n_rows * H;
  if (!valid) t = 0;
  const int64_t row = t / H;
  const int h = (int)(t - row * H);
  const int F = H * C;
  const int p0 = ptr[row], p1 = ptr[row + 1];
  const float dscale = concat ? 1.f : 1.f / (float)H;
  // Hub rows (> 64 entries; up to ~500 in the Elliptic-shaped graph): a thread walking one alone is the critical path of
  // the whole launch (two passes of ~120 dependent batches).  The warp takes its hub (row, head) items one after the
  // other with ONE LANE PER ENTRY and warp-reduced sums; the summation order differs from the sequential walk, results
  // agree to rounding (same convention as the hub paths of the forward and the source pass).
  // (measured: 96 -> 84 us for the two calls of a gat.yaml step; a threshold of 16 instead of 64 entries is slower, 97 us)
  const bool is_long = valid && (p1 - p0 > kGatLongRow);
  for (unsigned lm = __ballot_sync(0xffffffffu, is_long); lm; lm &= lm - 1) {
    const int sl = __ffs(lm) - 1, wl = threadIdx.x & 31;
    const int64_t lrow = __shfl_sync(0xffffffffu, row, sl);
    const int lh = __shfl_sync(0xffffffffu, h, sl), q0 = __shfl_sync(0xffffffffu, p0, sl), q1 = __shfl_sync(0xffffffffu, p1, sl);
    const float* drow = concat ? dout + lrow * F + lh * C : dout + lrow * C;
    float sp = 0.f;
    for (int p = q0 + wl; p < q1; p += 32) {
      const float* xj = xs + (int64_t)__ldg(src + p) * F + lh * C;
      float g = 0.f;
      for (int c = 0; c < C; ++c) g = fmaf(drow[c] * dscale, __ldg(xj + c), g);
      dpre[(int64_t)p * H + lh] = g;
      sp = fmaf(alpha[(int64_t)p * H + lh], g, sp);
    }
    const float ls = warp_sum(sp);
    const float lad = a_d[lrow * H + lh];
    float ap = 0.f;
    for (int p = q0 + wl; p < q1; p += 32) {       // each lane re-reads the entries it wrote itself
      const float g = dpre[(int64_t)p * H + lh];
      const float de = alpha[(int64_t)p * H + lh] * (g - ls);
      const float pre = __ldg(a_s + (int64_t)__ldg(src + p) * H + lh) + lad;
      const float dp = pre > 0.f ? de : de * slope;
      dpre[(int64_t)p * H + lh] = dp;
      ap += dp;
    }
    ap = warp_sum(ap);
    if (wl == 0) da_d[lrow * H + lh] = ap;
  }
  if (!valid || is_long) return;
  const float* dorow = concat ? dout + row * F + h * C : dout + row * C;
  float s = 0.f;
  const bool vec8 = C == 8 && concat && ((uintptr_t)xs & 15) == 0 && ((uintptr_t)dout & 15) == 0;
  if (p1 - p0 <= 4) {
    // The usual row (mean degree 2.2 with the self loop): everything for its <= 4 entries is requested at once and
    // stays in registers -- one pass, three dependent round trips (row pointer -> sources -> gathers) instead of the
    // two passes with a store / reload of dpre in between.  Same operations in the same order as the general path.
    const int deg = p1 - p0;
    int sj[4];
    float al[4], as4[4], g[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) sj[k] = k < deg ? __ldg(src + p0 + k) : 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      al[k] = k < deg ? alpha[(int64_t)(p0 + k) * H + h] : 0.f;
      as4[k] = __ldg(a_s + (int64_t)sj[k] * H + h);
      g[k] = 0.f;
    }
    if (vec8) {
      const float4 d0 = *reinterpret_cast<const float4*>(dorow), d1 = *reinterpret_cast<const float4*>(dorow + 4);
      float4 a[4], b[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float* xj = xs + (int64_t)sj[k] * F + h * 8;
        a[k] = __ldg(reinterpret_cast<const float4*>(xj));
        b[k] = __ldg(reinterpret_cast<const float4*>(xj) + 1);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float q = 0.f;
        q = fmaf(d0.x, a[k].x, q); q = fmaf(d0.y, a[k].y, q); q = fmaf(d0.z, a[k].z, q); q = fmaf(d0.w, a[k].w, q);
        q = fmaf(d1.x, b[k].x, q); q = fmaf(d1.y, b[k].y, q); q = fmaf(d1.z, b[k].z, q); q = fmaf(d1.w, b[k].w, q);
        g[k] = q;
      }
    } else {
      for (int c = 0; c < C; ++c) {
        const float d = dorow[c] * dscale;
#pragma unroll
        for (int k = 0; k < 4; ++k) g[k] = fmaf(d, __ldg(xs + (int64_t)sj[k] * F + h * C + c), g[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (k < deg) s = fmaf(al[k], g[k], s);
    const float ad0 = a_d[t];
    float acc0 = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (k < deg) {
        const float de = al[k] * (g[k] - s);
        const float dp = (as4[k] + ad0) > 0.f ? de : de * slope;
        dpre[(int64_t)(p0 + k) * H + h] = dp;
        acc0 += dp;
      }
    }
    da_d[t] = acc0;
    return;
  }
  if (vec8) {
    // 8 channels per head (the 4 x 8 hidden layer): the head's slice of a row is 32 aligned bytes -> two 16-byte
    // loads per gathered row instead of eight scalar ones, the output-gradient slice kept in registers
    const float4 d0 = *reinterpret_cast<const float4*>(dorow), d1 = *reinterpret_cast<const float4*>(dorow + 4);
    int p = p0;
    for (; p + 4 <= p1; p += 4) {  // four gathers in flight
      float4 a[4], b[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float* xj = xs + (int64_t)__ldg(src + p + k) * F + h * 8;
        a[k] = __ldg(reinterpret_cast<const float4*>(xj));
        b[k] = __ldg(reinterpret_cast<const float4*>(xj) + 1);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float g = 0.f;   // same order of fused multiply-adds as the scalar loop below: c = 0 .. 7
        g = fmaf(d0.x, a[k].x, g); g = fmaf(d0.y, a[k].y, g); g = fmaf(d0.z, a[k].z, g); g = fmaf(d0.w, a[k].w, g);
        g = fmaf(d1.x, b[k].x, g); g = fmaf(d1.y, b[k].y, g); g = fmaf(d1.z, b[k].z, g); g = fmaf(d1.w, b[k].w, g);
        dpre[(int64_t)(p + k) * H + h] = g;
        s = fmaf(alpha[(int64_t)(p + k) * H + h], g, s);
      }
    }
    for (; p < p1; ++p) {
      const float* xj = xs + (int64_t)__ldg(src + p) * F + h * 8;
      const float4 a = __ldg(reinterpret_cast<const float4*>(xj)), b = __ldg(reinterpret_cast<const float4*>(xj) + 1);
      float g = 0.f;
      g = fmaf(d0.x, a.x, g); g = fmaf(d0.y, a.y, g); g = fmaf(d0.z, a.z, g); g = fmaf(d0.w, a.w, g);
      g = fmaf(d1.x, b.x, g); g = fmaf(d1.y, b.y, g); g = fmaf(d1.z, b.z, g); g = fmaf(d1.w, b.w, g);
      dpre[(int64_t)p * H + h] = g;
      s = fmaf(alpha[(int64_t)p * H + h], g, s);
    }
  } else {
    int p = p0;
    for (; p + 4 <= p1; p += 4) {  // four gathers in flight
      const float* xj[4];
      float g[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int b = 0; b < 4; ++b) xj[b] = xs + (int64_t)__ldg(src + p + b) * F + h * C;
      for (int c = 0; c < C; ++c) {
        const float d = dorow[c] * dscale;
#pragma unroll
        for (int b = 0; b < 4; ++b) g[b] = fmaf(d, __ldg(xj[b] + c), g[b]);
      }
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        dpre[(int64_t)(p + b) * H + h] = g[b];
        s = fmaf(alpha[(int64_t)(p + b) * H + h], g[b], s);
      }
    }
    for (; p < p1; ++p) {
      const float* xj = xs + (int64_t)__ldg(src + p) * F + h * C;
      float g = 0.f;
      for (int c = 0; c < C; ++c) g = fmaf(dorow[c] * dscale, __ldg(xj + c), g);
      dpre[(int64_t)p * H + h] = g;
      s = fmaf(alpha[(int64_t)p * H + h], g, s);
    }
  }
  const float ad = a_d[t];
  float acc = 0.f;
  {
    int p = p0;
    for (; p + 4 <= p1; p += 4) {
      float as4[4];
#pragma unroll
      for (int b = 0; b < 4; ++b) as4[b] = __ldg(a_s + (int64_t)__ldg(src + p + b) * H + h);
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const float g = dpre[(int64_t)(p + b) * H + h];
        const float de = alpha[(int64_t)(p + b) * H + h] * (g - s);
        const float dp = (as4[b] + ad) > 0.f ? de : de * slope;
        dpre[(int64_t)(p + b) * H + h] = dp;
        acc += dp;
      }
    }
    for (; p < p1; ++p) {
      float g = dpre[(int64_t)p * H + h];
      float de = alpha[(int64_t)p * H + h] * (g - s);
      float pre = __ldg(a_s + (int64_t)__ldg(src + p) * H + h) + ad;
      float dp = pre > 0.f ? de : de * slope;
      dpre[(int64_t)p * H + h] = dp;
      acc += dp;
    }
  }
  da_d[t] = acc;
}

// source pass: G lanes per source row over the CSC view
template <int G, int HT>
__global__ void __launch_bounds__(kThreads) gat_bwd_src_kernel(
    const int* __restrict__ ptr, const int* __restrict__ dst, const int* __restrict__ pos,
    const float* __restrict__ alpha, const float* __restrict__ dpre, const float* __restrict__ dout,
    const float* __restrict__ da_d, const float* __restrict__ att_src, const float* __restrict__ att_dst, int H,
    int C, int concat, float* __restrict__ dxs, float* da_s, int64_t n_rows) {
  const int Hc = HT > 0 ? HT : H;   // compile-time head count (HT > 0): the per-head loops below fold
  const int lane = threadIdx.x % G;
  int64_t row = ((int64_t)blockIdx.x * kThreads + threadIdx.x) / G;
  const bool in_range = row < n_rows;
  if (!in_range) row = 0;
  int q0 = in_range ? ptr[row] : 0, q1 = in_range ? ptr[row + 1] : 0;
  const bool is_long = G < 32 && Hc <= 8 && in_range && q1 - q0 > kGatLongRow;
  for (unsigned lm = __ballot_sync(0xffffffffu, is_long && lane == 0); lm; lm &= lm - 1) {
    const int sl = __ffs(lm) - 1;
    gat_bwd_src_long_row<HT>(__shfl_sync(0xffffffffu, row, sl), __shfl_sync(0xffffffffu, q0, sl),
                         __shfl_sync(0xffffffffu, q1, sl), dst, pos, alpha, dpre, dout, da_d, att_src, att_dst, Hc, C,
                         concat, dxs, da_s);
  }
  const bool live = in_range && !is_long;
  if (!live) q0 = q1 = 0;
  const int F = Hc * C;
  constexpr int kHMax = HT > 0 ? HT : 8;
  float das[kHMax];
#pragma unroll
  for (int h = 0; h < kHMax; ++h) das[h] = 0.f;
  if (Hc <= kHMax) {  // one lane per entry, warp-reduced: no serial walk over the row
#pragma unroll 4
    for (int base = q0; base < q1; base += G) {
      const int q = base + lane;
      if (q < q1) {
        const int64_t pq = __ldg(pos + q);
#pragma unroll
        for (int h = 0; h < kHMax; ++h)
          if (h < Hc) das[h] += __ldg(dpre + pq * Hc + h);
      }
    }
#pragma unroll
    for (int h = 0; h < kHMax; ++h)
      if (h < Hc) {
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) das[h] += __shfl_xor_sync(0xffffffffu, das[h], o);
        if (lane == h % G && live) da_s[row * Hc + h] = das[h];
      }
  } else {
    for (int h = lane; live && h < Hc; h += G) {
      float acc = 0.f;
      for (int q = q0; q < q1; ++q) acc += __ldg(dpre + (int64_t)__ldg(pos + q) * Hc + h);
      da_s[row * Hc + h] = acc;
    }
  }
  __syncwarp();
  const float dscale = concat ? 1.f : 1.f / (float)Hc;
  constexpr int kB = 4;  // entries in flight per batch (a hub row is walked by its lanes alone)
  if (concat && F == 4 * G && (C & 3) == 0) {  // four contiguous features of one head per lane: 16-byte gathers
    const int f0 = 4 * lane, h = f0 / C;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int q = q0; q < q1; q += kB) {
      float a[kB];
      float4 d[kB];
#pragma unroll
      for (int b = 0; b < kB; ++b) {
        const bool ok = q + b < q1;
        a[b] = ok ? __ldg(alpha + (int64_t)__ldg(pos + q + b) * Hc + h) : 0.f;
        d[b] = ok ? __ldg(reinterpret_cast<const float4*>(dout + (int64_t)__ldg(dst + q + b) * F + f0))
                  : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int b = 0; b < kB; ++b) {
        if (q + b < q1) {
          acc.x = fmaf(a[b], d[b].x, acc.x);
          acc.y = fmaf(a[b], d[b].y, acc.y);
          acc.z = fmaf(a[b], d[b].z, acc.z);
          acc.w = fmaf(a[b], d[b].w, acc.w);
        }
      }
    }
    if (live) {
      float das_h = 0.f;
      if (Hc <= kHMax) {
#pragma unroll
        for (int k = 0; k < kHMax; ++k) das_h = k == h ? das[k] : das_h;
      } else {
        das_h = da_s[row * Hc + h];
      }
      const float dad = da_d[row * Hc + h];
      float r[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        r[i] = fmaf(das_h, att_src[f0 + i], r[i]);
        r[i] = fmaf(dad, att_dst[f0 + i], r[i]);
        dxs[row * F + f0 + i] = r[i];
      }
    }
    return;
  }
  for (int f = lane; live && f < F; f += G) {
    const int h = f / C, c = f - h * C;
    float acc = 0.f;
    for (int q = q0; q < q1; q += kB) {
      float d[kB], a[kB];
#pragma unroll
      for (int b = 0; b < kB; ++b) {
        const bool ok = q + b < q1;
        const int64_t dq = ok ? __ldg(dst + q + b) : 0;
        d[b] = !ok ? 0.f : concat ? __ldg(dout + dq * F + f) : __ldg(dout + dq * C + c) * dscale;
        a[b] = ok ? __ldg(alpha + (int64_t)__ldg(pos + q + b) * Hc + h) : 0.f;
      }
#pragma unroll
      for (int b = 0; b < kB; ++b)
        if (q + b < q1) acc = fmaf(a[b], d[b], acc);
    }
    float das_h = 0.f;
    if (Hc <= kHMax) {
#pragma unroll
      for (int k = 0; k < kHMax; ++k) das_h = k == h ? das[k] : das_h;
    } else {
      das_h = da_s[row * Hc + h];
    }
    acc = fmaf(das_h, att_src[f], acc);
    acc = fmaf(da_d[row * Hc + h], att_dst[f], acc);
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
  // the head count is a template parameter for the usual values: with a runtime H the per-head loops are unrolled
  // to 8 and predicated (ncu, round 2: 38.5 M warp instructions for the 1-head logits layer, 71 us)
#define EGNN_GAT_FWD(HT_)                                                                                         \
  gat_fwd_kernel<kGatLanes, HT_><<<(unsigned)ceil_div(n_rows * kGatLanes, kThreads), kThreads, 0, (cudaStream_t)stream>>>( \
      csr_ptr, csr_src, xs, a_s, a_d, negative_slope, H, C, concat, bias, alpha, out, n_rows)
  if (H == 1 && C <= 4) {   // the logits layer (1 head x 2): 4 lanes per row -- half the warps, half the waves
    gat_fwd_kernel<4, 1><<<(unsigned)ceil_div(n_rows * 4, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        csr_ptr, csr_src, xs, a_s, a_d, negative_slope, H, C, concat, bias, alpha, out, n_rows);
    EGNN_LAUNCH_CHECK(fn);
    return 0;
  }
  switch (H) {
    case 1: EGNN_GAT_FWD(1); break;
    case 2: EGNN_GAT_FWD(2); break;
    case 4: EGNN_GAT_FWD(4); break;
    case 8: EGNN_GAT_FWD(8); break;
    default: EGNN_GAT_FWD(0); break;
  }
#undef EGNN_GAT_FWD
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
#define EGNN_GAT_BSRC(HT_)                                                                                        \
  gat_bwd_src_kernel<kGatLanes, HT_><<<(unsigned)ceil_div(n_rows * kGatLanes, kThreads), kThreads, 0, (cudaStream_t)stream>>>( \
      csc_ptr, csc_dst, csc_pos, alpha, dpre, dout, da_d, att_src, att_dst, H, C, concat, dxs, da_s, n_rows)
  if (H == 1 && C <= 4) {
    gat_bwd_src_kernel<4, 1><<<(unsigned)ceil_div(n_rows * 4, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        csc_ptr, csc_dst, csc_pos, alpha, dpre, dout, da_d, att_src, att_dst, H, C, concat, dxs, da_s, n_rows);
    EGNN_LAUNCH_CHECK(fn);
    return 0;
  }
  switch (H) {
    case 1: EGNN_GAT_BSRC(1); break;
    case 2: EGNN_GAT_BSRC(2); break;
    case 4: EGNN_GAT_BSRC(4); break;
    case 8: EGNN_GAT_BSRC(8); break;
    default: EGNN_GAT_BSRC(0); break;
  }
#undef EGNN_GAT_BSRC
  EGNN_LAUNCH_CHECK(fn);
  return 0;
}
