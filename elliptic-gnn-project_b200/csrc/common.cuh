// Shared helpers for the egnn_b200 C-ABI library (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/egnn_b200.h"

namespace egnn {

constexpr int kNumSMs = 148;  // B200

// ---- error plumbing ------------------------------------------------------------------
char* err_buf();                 // thread-local message buffer (api.cu)
extern std::atomic<uint64_t> g_launches;

inline int fail(const char* fn, const char* msg) {
  snprintf(err_buf(), 512, "%s: %s", fn, msg);
  return -1;
}
inline int check_launch(const char* fn) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(err_buf(), 512, "%s: CUDA error %d (%s)", fn, (int)e, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}
#define EGNN_REQUIRE(cond, fn, msg) \
  do {                              \
    if (!(cond)) return egnn::fail(fn, msg); \
  } while (0)
#define EGNN_LAUNCH_CHECK(fn)             \
  do {                                    \
    int _e = egnn::check_launch(fn);      \
    if (_e) return _e;                    \
  } while (0)

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- dtype helpers -------------------------------------------------------------------
template <typename T>
struct Vec4;  // 4 consecutive elements

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) {
  return __float2bfloat16_rn(v);
}

// 4 features held in fp32 registers, loaded/stored as one 16-byte (fp32) or 8-byte (bf16) access
struct F4 {
  float x, y, z, w;
};
__device__ __forceinline__ F4 ld4(const float* p) {
  float4 v = __ldg(reinterpret_cast<const float4*>(p));
  return {v.x, v.y, v.z, v.w};
}
__device__ __forceinline__ F4 ld4(const __nv_bfloat16* p) {
  uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
  F4 r;
  r.x = __uint_as_float(v.x << 16);
  r.y = __uint_as_float(v.x & 0xffff0000u);
  r.z = __uint_as_float(v.y << 16);
  r.w = __uint_as_float(v.y & 0xffff0000u);
  return r;
}
__device__ __forceinline__ void st4(float* p, F4 v) {
  *reinterpret_cast<float4*>(p) = make_float4(v.x, v.y, v.z, v.w);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ void st4(__nv_bfloat16* p, F4 v) {
  uint2 o;
  o.x = pack_bf16x2(v.x, v.y);
  o.y = pack_bf16x2(v.z, v.w);
  *reinterpret_cast<uint2*>(p) = o;
}

// 8 features (one 16-byte access of bf16)
struct F8 {
  float v[8];
};
__device__ __forceinline__ F8 ld8(const __nv_bfloat16* p) {
  uint4 q = __ldg(reinterpret_cast<const uint4*>(p));
  F8 r;
  uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    r.v[2 * i] = __uint_as_float(w[i] << 16);
    r.v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
  return r;
}
__device__ __forceinline__ void st8(__nv_bfloat16* p, const F8& r) {
  uint4 q;
  q.x = pack_bf16x2(r.v[0], r.v[1]);
  q.y = pack_bf16x2(r.v[2], r.v[3]);
  q.z = pack_bf16x2(r.v[4], r.v[5]);
  q.w = pack_bf16x2(r.v[6], r.v[7]);
  *reinterpret_cast<uint4*>(p) = q;
}

}  // namespace egnn
