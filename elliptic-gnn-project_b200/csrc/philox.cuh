// Philox4x32-10 counter-based RNG for the dropout keep-mask.  The mask for element
// (global row r, column c) of dropout layer `layer` is a pure function of
// (seed, layer, r, c): counter = (r_lo, r_hi, c/8, layer), key = (seed_lo, seed_hi); the four
// 32-bit output words are used as EIGHT 16-bit lanes (all 128 bits of a draw): column c takes
// lane c%8 = the low (even c) or high (odd c) half of word (c%8)/2, and is kept iff
// lane >= floor(p * 2^16)  (p = 0.2: keep probability 0.800003; torch's own dropout compares a
// 24-bit uniform).  Keyed on the GLOBAL row id so timestep-sharded runs draw the same mask as the
// single-GPU run.  oracle/graph_build_np.py holds the NumPy twin used by the tests.
#pragma once
#include <stdint.h>

namespace egnn {

struct Philox4 {
  uint32_t v[4];
};

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2,
                                                          uint32_t c3, uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += W0; k1 += W1;
  }
  return Philox4{{c0, c1, c2, c3}};
}

__host__ __device__ __forceinline__ uint32_t dropout_threshold(float p) {
  double t = (double)p * 4294967296.0;
  if (t >= 4294967295.0) return 4294967295u;
  if (t <= 0.0) return 0u;
  return (uint32_t)t;  // floor
}

__host__ __device__ __forceinline__ uint32_t dropout_threshold16(float p) {
  double t = (double)p * 65536.0;
  if (t >= 65535.0) return 65535u;
  if (t <= 0.0) return 0u;
  return (uint32_t)t;  // floor
}

// 8 keep flags (bit i = column 8*cb8 + i) of global row `row`
__device__ __forceinline__ uint32_t dropout_keep8(uint64_t seed, uint32_t layer, int64_t row, uint32_t cb8,
                                                  uint32_t thr16) {
  const Philox4 w = philox4x32_10((uint32_t)((uint64_t)row & 0xffffffffu), (uint32_t)((uint64_t)row >> 32), cb8, layer,
                                  (uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32));
  uint32_t bits = 0u;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    bits |= ((w.v[i] & 0xffffu) >= thr16 ? 1u : 0u) << (2 * i);
    bits |= ((w.v[i] >> 16) >= thr16 ? 1u : 0u) << (2 * i + 1);
  }
  return bits;
}

}  // namespace egnn
