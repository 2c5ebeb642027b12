// Philox4x32-10 counter-based RNG for the dropout keep-mask.  The mask for element
// (global row r, column c) of dropout layer `layer` is a pure function of
// (seed, layer, r, c): counter = (r_lo, r_hi, c/4, layer), key = (seed_lo, seed_hi), output
// word c%4; keep iff word >= floor(p * 2^32).  Keyed on the GLOBAL row id so timestep-sharded
// runs draw the same mask as the single-GPU run, and recomputed in the backward
// instead of being stored.  oracle/graph_build_np.py holds the NumPy twin used by the tests.
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

// 4 keep flags for columns 4*cb .. 4*cb+3 of global row `row`
__device__ __forceinline__ Philox4 dropout_words(uint64_t seed, uint32_t layer, int64_t row, uint32_t cb) {
  return philox4x32_10((uint32_t)((uint64_t)row & 0xffffffffu), (uint32_t)((uint64_t)row >> 32), cb, layer,
                       (uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32));
}

}  // namespace egnn
