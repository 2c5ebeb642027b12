// K6 (tensor-core path) -- bf16 dense projections on tcgen05 / TMEM / TMA (sm_100a).
//
// Two persistent, warp-specialised kernels (1 CTA per SM, 256 threads):
//   gemm_tn_kernel    C[M,N] = A[M,K] . W[N,K]^T (+bias, /rowcount, +=C)   Linear forward, and
//                     dgrad through a pre-transposed weight.  Both operands K-major.
//                     W stays resident in shared memory for the whole kernel; 128-row A tiles
//                     stream through a TMA ring; two TMEM accumulators overlap the epilogue of
//                     tile i with the MMAs of tile i+1.
//   gemm_wgrad_kernel dW[N,K] = G[M,N]^T . X[M,K]   reduction over the node axis M.  Both
//                     operands are MN-major (contiguous along the non-contracted dimension);
//                     every CTA accumulates its 64-node slabs into TMEM and writes one fp32
//                     partial, a second kernel sums the partials in a fixed order
//                     (deterministic, no float atomics).
// warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane), warp 2 = TMEM allocator,
// warps 4..7 = epilogue (tcgen05.ld -> registers -> global).
//
// Shared-memory operand layouts are the canonical UMMA SWIZZLE_128B layouts, written by TMA
// with CU_TENSOR_MAP_SWIZZLE_128B boxes whose inner extent is 64 bf16 (128 bytes):
//   K-major : row r of the operand = 128 B at r*128; 8-row groups are 1024 B apart (SBO);
//             one MMA (K=16) advances the start address by 32 B inside the swizzled row.
//   MN-major: row k (contraction index) = 128 B holding 64 consecutive MN elements; 8-row
//             groups 1024 B apart (SBO); the next 64 MN elements are a separate TMA box LBO
//             bytes further; one MMA (K=16) advances the start address by 2048 B.
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"

namespace egnn {
namespace {

constexpr int kThreads = 256;
constexpr int kThreadsTN = 384;   // gemm_tn_kernel: warps 0-3 control, 4-11 epilogue (two per TMEM sub-partition)
constexpr int kEpiWarps = 8;
constexpr int BM = 128;          // UMMA M (cta_group::1)
constexpr int BK = 64;           // bf16 elements per 128-byte swizzled row
constexpr int kStageBytesA = BM * BK * 2;  // 16 KB

// ------------------------------------------------------------------ PTX wrappers ---------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tcgen05_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// two 16-column slices, one wait: the second load's latency hides behind the first
__device__ __forceinline__ void tmem_ld16x2(uint32_t taddr0, uint32_t taddr1, float* v0, float* v1) {
  uint32_t r[16], u[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr0));
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
        "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
      : "r"(taddr1));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) { v0[i] = __uint_as_float(r[i]); v1[i] = __uint_as_float(u[i]); }
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred P;\n"
      "elect.sync _|P, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, P;\n"
      "}"
      : "=r"(pred));
  return pred != 0;
}

// UMMA shared-memory descriptor, SWIZZLE_128B, version 1 (Blackwell)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // version
  d |= (uint64_t)2 << 61;  // LayoutType::SWIZZLE_128B
  return d;
}
// the same with LayoutType::SWIZZLE_128B_BASE32B (1): the only shared-memory layout the tensor core accepts for MN-major
// 32-bit (tf32) operands -- 32-byte chunks of a 128-byte row are XOR-ed with the row index inside 4-row groups (TMA:
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B); SBO = distance between 4-row groups, LBO = distance between 128-byte column blocks
__device__ __forceinline__ uint64_t make_desc_b32(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // version
  d |= (uint64_t)1 << 61;  // LayoutType::SWIZZLE_128B_BASE32B
  return d;
}
// instruction descriptor: D=f32, A=B=bf16, M=128, N=n
__host__ __device__ constexpr uint32_t make_idesc(int n, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

struct Epilogue {
  void* C;
  const float* bias;
  const int32_t* row_div_ptr;  // optional: divide row m by max(ptr[m+1]-ptr[m], 1)
  int64_t ld_c;
  int c_dtype, accumulate;
  int row_div_cols;  // only columns [0, row_div_cols) are divided (<= 0: all)
  // optional second addend: C[m, add_col0 + j] += addend[m, j] (same dtype as C; add_col0 a multiple of 16) -- the
  // identity-residual gradient joining the root half of the concatenated SAGE dgrad
  const void* addend;
  int64_t ld_add;
  int add_col0;
  // optional BatchNorm statistics of the STORED values of columns [0, stats_cols), stats_cols <= 64:
  // stats[(cta*2 + {0: sum, 1: sum of squares})*stats_cols + c]: one partial row per CTA
  float* stats;
  int stats_cols;
};

// Epilogue of gemm_tn_kernel for one warp: FLAGS bit 0 = +bias, bit 1 = /row count, bit 2 = += C, bit 3 = second
// addend, bit 4 = column statistics.
//
// kExact (fp32 operands, Npad <= 128): the tensor core adds into its fp32 accumulator with TRUNCATION -- a bias of
// half an ulp per MMA that does not average out over a 40-MMA chain and is coherent across outputs.  In this mode
// every hi.hi product of one 8-wide k-step leaves the tensor core in a FRESH TMEM buffer (accumulate = 0) and is
// added here, on the CUDA cores, with round-to-nearest into 64 accumulator registers per thread (the scheme of
// Ootomo & Yokota for recovering fp32 accuracy).  TMEM is a ring of two "pairs" [2 k-steps][Npad columns]; t_full /
// t_empty are that ring's barriers; K_exact = the contraction length (the pair sequence is derived from it exactly as
// the MMA warp derives it).  The small hi.lo / lo.hi terms of a pair ride in its first buffer.
template <typename TC, int FLAGS, bool kExact = false>
__device__ __forceinline__ void epilogue_loop(const Epilogue& ep, const float* s_bias, uint64_t* t_full,
                                              uint64_t* t_empty, uint32_t tmem_base, int Npad, int M, int N,
                                              int n_tiles, int q, int half, int lane, int K_exact = 0) {
  constexpr bool kBias = FLAGS & 1, kDiv = (FLAGS & 2) != 0, kAcc = (FLAGS & 4) != 0;
  constexpr bool kAdd2 = (FLAGS & 8) != 0, kStats = (FLAGS & 16) != 0;
  float st_sum[kStats ? 2 : 1][16], st_sq[kStats ? 2 : 1][16];   // this warp's <= 2 chunks of the statistics columns
  if (kStats) {
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int j = 0; j < 16; ++j) { st_sum[kStats ? h : 0][j] = 0.f; st_sq[kStats ? h : 0][j] = 0.f; }
  }
  const TC* __restrict__ Abase = reinterpret_cast<const TC*>(ep.addend);
  // the second addend of a tile is requested BEFORE the wait for its accumulator (its address depends on the tile
  // index only), so the loads overlap the MMAs instead of sitting on the epilogue's critical path: 2 chunks per warp
  constexpr int kPreVec = sizeof(TC) == 2 ? 2 : 4;   // 16-byte vectors per 16-column chunk
  uint4 pre[kAdd2 ? 2 : 1][kPreVec];
  const bool pre_ok = kAdd2 && (N - ep.add_col0 <= 64) && (ep.add_col0 % 32 == 0) && (N % 16 == 0) &&
                      (ep.ld_add % 8 == 0) && ((uintptr_t)ep.addend % 16 == 0);
  const int div_cols = ep.row_div_cols > 0 ? ep.row_div_cols : N;
  TC* __restrict__ Cbase = reinterpret_cast<TC*>(ep.C);
  const bool vec_ok = (ep.ld_c % 8 == 0) && ((uintptr_t)ep.C % 16 == 0);
  int it = 0;
  uint32_t pc = 0;   // kExact: pairs consumed so far (ring position of the next one)
  for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
    const int buf = it & 1;
    const uint32_t use = (uint32_t)(it >> 1);
    const int64_t row = (int64_t)t * BM + q * 32 + lane;
    const bool row_ok = row < M;
    if (kAdd2 && pre_ok && row_ok) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int cc = half * 16 + 32 * i;   // column inside the addend
        if (ep.add_col0 + cc < N) {
          const uint4* src = reinterpret_cast<const uint4*>(Abase + row * ep.ld_add + cc);
#pragma unroll
          for (int u = 0; u < kPreVec; ++u) pre[kAdd2 ? i : 0][u] = __ldg(src + u);
        }
      }
    }
    float acc[kExact ? 4 : 1][16];
    if constexpr (kExact) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[i][j] = 0.f;
      const int n_steps = (K_exact + 7) >> 3;            // 8-wide k-steps of this tile
      for (int k0 = 0; k0 < n_steps; k0 += 2, ++pc) {    // (a 32-wide chunk holds 4 steps: pairs never straddle chunks)
        const uint32_t b = pc & 1;
        mbar_wait(&t_full[b], (pc >> 1) & 1);
        tcgen05_fence_after();
        // a pair has its second k-step unless it is the last step of an odd-length tail
        const bool two = k0 + 1 < n_steps;
        const uint32_t p_addr = tmem_base + ((uint32_t)(q * 32) << 16) + b * 2u * (uint32_t)Npad;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c0 = half * 16 + 32 * i;
          if (c0 < Npad) {
            float v0[16], v1[16];
            if (two) {
              tmem_ld16x2(p_addr + c0, p_addr + Npad + c0, v0, v1);
#pragma unroll
              for (int j = 0; j < 16; ++j) acc[kExact ? i : 0][j] = (acc[kExact ? i : 0][j] + v0[j]) + v1[j];
            } else {
              tmem_ld16(p_addr + c0, v0);
#pragma unroll
              for (int j = 0; j < 16; ++j) acc[kExact ? i : 0][j] += v0[j];
            }
          }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&t_empty[b]);
      }
    } else {
      mbar_wait(&t_full[buf], use & 1);
      tcgen05_fence_after();
    }
    float rdiv = 1.f, rinv = 1.f;
    if (kDiv && row_ok) {
      int d = ep.row_div_ptr[row + 1] - ep.row_div_ptr[row];
      rdiv = (float)(d > 1 ? d : 1);
      rinv = __frcp_rn(rdiv);
    }
    const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * Npad);
    constexpr int kSl = kExact ? 4 : 1;   // kExact: all (<= 4) slices of this warp, statically indexed registers
    for (int cb = half * 16; cb < N; cb += 32 * kSl) {
#pragma unroll
    for (int sl = 0; sl < kSl; ++sl) {
      const int c0 = cb + 32 * sl;
      if (c0 >= N) break;
      float v[16];
      if constexpr (kExact) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = acc[kExact ? sl : 0][j];
      } else {
        tmem_ld16(t_addr + c0, v);
      }
      if (!row_ok) continue;
      const int nv = min(16, N - c0);
      TC* c = Cbase + row * ep.ld_c + c0;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if (kBias) v[j] += s_bias[c0 + j];
        if (kDiv) {
          // bf16 result: x * rn(1/deg) differs from the IEEE quotient by < 1 ulp(fp32), invisible after
          // rounding to 8 mantissa bits; fp32 result keeps the exact division (parity with CPU torch)
          if (c0 + j < div_cols) v[j] = sizeof(TC) == 2 ? v[j] * rinv : __fdiv_rn(v[j], rdiv);
        }
      }
      if (kAdd2 && c0 >= ep.add_col0 && pre_ok) {
        const bool second = (c0 - ep.add_col0) >= 32;
#pragma unroll
        for (int u = 0; u < kPreVec; ++u) {
          const uint4 w = second ? pre[kAdd2 ? 1 : 0][u] : pre[0][u];
          if (sizeof(TC) == 4) {
            v[4 * u] += __uint_as_float(w.x); v[4 * u + 1] += __uint_as_float(w.y);
            v[4 * u + 2] += __uint_as_float(w.z); v[4 * u + 3] += __uint_as_float(w.w);
          } else {
            v[8 * u] += __uint_as_float(w.x << 16); v[8 * u + 1] += __uint_as_float(w.x & 0xffff0000u);
            v[8 * u + 2] += __uint_as_float(w.y << 16); v[8 * u + 3] += __uint_as_float(w.y & 0xffff0000u);
            v[8 * u + 4] += __uint_as_float(w.z << 16); v[8 * u + 5] += __uint_as_float(w.z & 0xffff0000u);
            v[8 * u + 6] += __uint_as_float(w.w << 16); v[8 * u + 7] += __uint_as_float(w.w & 0xffff0000u);
          }
        }
      } else if (kAdd2 && c0 >= ep.add_col0) {
        const TC* a = Abase + row * ep.ld_add + (c0 - ep.add_col0);
        if (nv == 16 && ep.ld_add % 8 == 0 && ((uintptr_t)ep.addend % 16 == 0)) {
          if (sizeof(TC) == 4) {
#pragma unroll
            for (int h = 0; h < 4; ++h) {
              const float4 o = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(a) + 4 * h);
              v[4 * h] += o.x; v[4 * h + 1] += o.y; v[4 * h + 2] += o.z; v[4 * h + 3] += o.w;
            }
          } else {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const F8 o = ld8(reinterpret_cast<const __nv_bfloat16*>(a) + 8 * h);
#pragma unroll
              for (int j = 0; j < 8; ++j) v[8 * h + j] += o.v[j];
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (j < nv) v[j] += to_f32(a[j]);
        }
      }
      if (kStats && c0 < ep.stats_cols) {   // statistics of the values as stored (rounded to the output dtype)
        const int h = c0 >> 5;               // chunks half*16 + 32*h of this warp
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float r = (kAcc || c0 + j >= ep.stats_cols) ? 0.f : to_f32(from_f32<TC>(v[j]));
          if (h == 0) { st_sum[0][j] += r; st_sq[0][j] = fmaf(r, r, st_sq[0][j]); }
          else { st_sum[kStats ? 1 : 0][j] += r; st_sq[kStats ? 1 : 0][j] = fmaf(r, r, st_sq[kStats ? 1 : 0][j]); }
        }
      }
      if (nv == 16 && vec_ok) {
        if (sizeof(TC) == 4) {
          float* cf = reinterpret_cast<float*>(c);
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            float4 o = make_float4(v[4 * h], v[4 * h + 1], v[4 * h + 2], v[4 * h + 3]);
            if (kAcc) {
              const float4 old = *reinterpret_cast<const float4*>(cf + 4 * h);
              o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
            }
            *reinterpret_cast<float4*>(cf + 4 * h) = o;
          }
        } else {
          __nv_bfloat16* cb = reinterpret_cast<__nv_bfloat16*>(c);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            F8 r;
            if (kAcc) {
              const F8 old = ld8(cb + 8 * h);
#pragma unroll
              for (int j = 0; j < 8; ++j) r.v[j] = v[8 * h + j] + old.v[j];
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) r.v[j] = v[8 * h + j];
            }
            st8(cb + 8 * h, r);
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (j < nv) {
            float o = v[j];
            if (kAcc) o += to_f32(c[j]);
            c[j] = from_f32<TC>(o);
          }
        }
      }
    }
    }
    if constexpr (!kExact) {
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_empty[buf]);
    }
  }
  if (kStats) {
    // fixed butterfly over the 32 rows of this warp, the four sub-partitions combined in order through shared
    // memory, then ONE partial row per CTA: the assignment of tiles to CTAs is static, so the partials (and their
    // fixed-order sum in colstats_parts_kernel) are reproducible
    float (*s_st)[2][64] = reinterpret_cast<float (*)[2][64]>(const_cast<float*>(s_bias) + 256 + 16);   // 2 KB behind the bias
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c0 = half * 16 + 32 * h;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float a = st_sum[kStats ? h : 0][j], b = st_sq[kStats ? h : 0][j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          a += __shfl_xor_sync(0xffffffffu, a, o);
          b += __shfl_xor_sync(0xffffffffu, b, o);
        }
        if (lane == 0) {
          s_st[q][0][c0 + j] = a;
          s_st[q][1][c0 + j] = b;
        }
      }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");   // the 8 epilogue warps only
    const int tid = (half * 4 + q) * 32 + lane;      // 0..255
    if (tid < 128) {
      const int which = tid >> 6, c = tid & 63;
      if (c < ep.stats_cols)
        ep.stats[((size_t)blockIdx.x * 2 + which) * ep.stats_cols + c] =
            ((s_st[0][which][c] + s_st[1][which][c]) + s_st[2][which][c]) + s_st[3][which][c];
    }
  }
}

// ------------------------------------------------------------------ forward / dgrad --------
// smem map (dynamic, 1024-aligned): [W chunks: KC * Npad*128][A ring: S * 16 KB][barriers]
template <int kStages>
__global__ void __launch_bounds__(kThreadsTN, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, int M, int N,
               int Npad, int K, Epilogue ep) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~uintptr_t(1023));
  const int KC = (K + BK - 1) / BK;
  const int w_chunk_bytes = Npad * 128;
  uint8_t* sW = smem;
  uint8_t* sA = smem + (size_t)KC * w_chunk_bytes;  // Npad*128 is a multiple of 1024 (Npad % 8 == 0)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sA + (size_t)kStages * kStageBytesA);
  uint64_t* full = bars;                    // [kStages]
  uint64_t* empty = bars + kStages;         // [kStages]
  uint64_t* w_full = bars + 2 * kStages;    // [1]
  uint64_t* t_full = w_full + 1;            // [2]
  uint64_t* t_empty = t_full + 2;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (M + BM - 1) / BM;
  float* s_bias = reinterpret_cast<float*>(tmem_slot + 4);  // [256 + 16] floats behind the barriers
  for (int i = threadIdx.x; i < 256 + 16; i += kThreadsTN) s_bias[i] = (ep.bias && i < N) ? ep.bias[i] : 0.f;
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < 2 * Npad) tmem_cols <<= 1;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(w_full, 1);
    for (int b = 0; b < 2; ++b) { mbar_init(&t_full[b], 1); mbar_init(&t_empty[b], kEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      // weights: all K chunks, resident for the whole kernel
      mbar_expect_tx(w_full, (uint32_t)(KC * w_chunk_bytes));
      for (int kc = 0; kc < KC; ++kc) tma_load_2d(sW + (size_t)kc * w_chunk_bytes, &tmW, w_full, kc * BK, 0);
      int s = 0;
      uint32_t ph = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        for (int kc = 0; kc < KC; ++kc) {
          mbar_wait(&empty[s], ph ^ 1);
          mbar_expect_tx(&full[s], kStageBytesA);
          tma_load_2d(sA + (size_t)s * kStageBytesA, &tmA, &full[s], kc * BK, t * BM);
          if (++s == kStages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc(Npad, 0, 0);
    mbar_wait(w_full, 0);
    int s = 0;
    uint32_t ph = 0;
    int it = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t use = (uint32_t)(it >> 1);  // how many times this buffer was used before
      mbar_wait(&t_empty[buf], (use & 1) ^ 1);
      tcgen05_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(buf * Npad);
      for (int kc = 0; kc < KC; ++kc) {
        mbar_wait(&full[s], ph);
        tcgen05_fence_after();
        if (elect_one()) {
          const int k_left = K - kc * BK;
          const int n_k = k_left >= BK ? BK / 16 : (k_left + 15) / 16;
          const uint32_t a_addr = smem_u32(sA + (size_t)s * kStageBytesA);
          const uint32_t b_addr = smem_u32(sW + (size_t)kc * w_chunk_bytes);
          for (int k4 = 0; k4 < n_k; ++k4) {
            uint64_t ad = make_desc(a_addr + k4 * 32, 16, 1024);
            uint64_t bd = make_desc(b_addr + k4 * 32, 16, 1024);
            tcgen05_mma_bf16(d_tmem, ad, bd, idesc, (kc | k4) != 0);
          }
          tcgen05_commit(&empty[s]);                       // frees the smem stage when the MMAs retire
          if (kc == KC - 1) tcgen05_commit(&t_full[buf]);  // accumulator ready for the epilogue
        }
        __syncwarp();
        if (++s == kStages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // 8 epilogue warps: warp w reads TMEM lanes 32*(w&3).. (its sub-partition) and every second
    // 16-column chunk, so each scheduler has two epilogue warps to interleave
    const int q = warp & 3, half = (warp - 4) >> 2;
    // the combinations the layers use are compiled in: plain 0-7, +second addend (with / without row division), and
    // +statistics with bias (the SAGE layer GEMM feeding BatchNorm)
    const int flags = (ep.bias ? 1 : 0) | (ep.row_div_ptr ? 2 : 0) | (ep.accumulate ? 4 : 0) | (ep.addend ? 8 : 0) |
                      (ep.stats ? 16 : 0);
#define EGNN_EPI_CASE(F)                                                                                      \
  case F:                                                                                                     \
    if (ep.c_dtype == EGNN_F32)                                                                               \
      epilogue_loop<float, F>(ep, s_bias, t_full, t_empty, tmem_base, Npad, M, N, n_tiles, q, half, lane);    \
    else                                                                                                      \
      epilogue_loop<__nv_bfloat16, F>(ep, s_bias, t_full, t_empty, tmem_base, Npad, M, N, n_tiles, q, half,   \
                                      lane);                                                                  \
    break;
    switch (flags) {
      EGNN_EPI_CASE(0) EGNN_EPI_CASE(1) EGNN_EPI_CASE(2) EGNN_EPI_CASE(3)
      EGNN_EPI_CASE(4) EGNN_EPI_CASE(5) EGNN_EPI_CASE(6) EGNN_EPI_CASE(7)
      EGNN_EPI_CASE(8) EGNN_EPI_CASE(10) EGNN_EPI_CASE(16) EGNN_EPI_CASE(17)
    }
#undef EGNN_EPI_CASE
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols));
  }
}

// ------------------------------------------------------------------ fp32 operands: 3xTF32 ---
// The exact-fp32 path (every `eval_split` forward, src/train_gnn.py:248-257, and the fp32 configs) needs rel 1e-5:
// plain TF32 (10-bit mantissa) is not enough, FFMA leaves the tensor cores idle.  3xTF32: a = a_hi + a_lo with
// a_hi = rna_tf32(a), and  a.w ~= a_lo.w_hi + a_hi.w_lo + a_hi.w_hi  accumulated in fp32 in TMEM (the dropped
// a_lo.w_lo term and the truncation of the lo parts are ~2^-22 relative).  W_hi / W_lo are split once per call by a
// small kernel; the A tile arrives by TMA as fp32 and two converter warps split it IN SHARED MEMORY (hi written back
// in place, lo into a second tile at the same swizzled offsets), then hand the stage to the MMA warp.
constexpr int BK32 = 32;   // fp32 elements per 128-byte swizzled row

__host__ __device__ constexpr uint32_t make_idesc_tf32(int n, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
__device__ __forceinline__ void tcgen05_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ float rna_tf32(float a) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(a));
  return __uint_as_float(r);
}
// hi / lo split of `n16` 16-byte vectors by `nthreads` threads: hi in place, lo at the same offset of `lo`
// (batches of 8 independent loads: the in-place stores would otherwise serialise every iteration on its own load)
__device__ __forceinline__ void split_tile_tf32(uint8_t* hi, uint8_t* lo, int n16, int tid, int nthreads) {
  constexpr int kBatch = 8;
  for (int i0 = tid; i0 < n16; i0 += nthreads * kBatch) {
    float4 v[kBatch];
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      const int i = i0 + u * nthreads;
      if (i < n16) v[u] = *reinterpret_cast<const float4*>(hi + (size_t)i * 16);
    }
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      const int i = i0 + u * nthreads;
      if (i < n16) {
        const float4 h = make_float4(rna_tf32(v[u].x), rna_tf32(v[u].y), rna_tf32(v[u].z), rna_tf32(v[u].w));
        *reinterpret_cast<float4*>(hi + (size_t)i * 16) = h;
        // lo rounded to tf32 here (the tensor core would TRUNCATE it: a biased 2^-21 error instead of a 2^-22 rounding)
        *reinterpret_cast<float4*>(lo + (size_t)i * 16) =
            make_float4(rna_tf32(v[u].x - h.x), rna_tf32(v[u].y - h.y), rna_tf32(v[u].z - h.z), rna_tf32(v[u].w - h.w));
      }
    }
  }
}

__global__ void split_tf32_kernel(const float* __restrict__ w, int64_t n, float* __restrict__ hi, float* __restrict__ lo) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = w[i], h = rna_tf32(v);
  hi[i] = h;
  lo[i] = rna_tf32(v - h);
}

// smem map (dynamic, 1024-aligned): stages of [A 16 KB][A_lo 16 KB][W_hi chunk Npad*128][W_lo chunk Npad*128], then
// barriers / TMEM slot / bias / statistics scratch as in gemm_tn_kernel
// kConvWarps: 2 (warps 2-3; 384 threads, every epilogue compiled in) or 6 (warps 2-3 and 12-15; 512 threads, i.e. 128
// registers per thread: the statistics epilogues, which need more, are left out of that variant)
// kExact: see epilogue_loop -- every hi.hi k-step product into a fresh TMEM buffer, summed with IEEE adds by the epilogue
// warps (Npad <= 128: ring of 2 pairs x 2 buffers x Npad columns = the whole 512-column TMEM at Npad = 128).
template <int kStages, int kConvWarps, bool kExact>
__global__ void __launch_bounds__(kThreadsTN + (kConvWarps - 2) * 32, 1)
gemm_tn_tf32x3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmWhi,
                      const __grid_constant__ CUtensorMap tmWlo, int M, int N, int Npad, int K, Epilogue ep, int dbg) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~uintptr_t(1023));
  const int KC = (K + BK32 - 1) / BK32;
  const int w_chunk_bytes = Npad * 128;
  const int stage_bytes = 2 * kStageBytesA + 2 * w_chunk_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)kStages * stage_bytes);
  uint64_t* full = bars;                     // [kStages]  TMA landed
  uint64_t* conv = bars + kStages;           // [kStages]  A split into hi / lo
  uint64_t* empty = bars + 2 * kStages;      // [kStages]  MMAs retired
  uint64_t* t_full = bars + 3 * kStages;     // [2]
  uint64_t* t_empty = t_full + 2;            // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (M + BM - 1) / BM;
  float* s_bias = reinterpret_cast<float*>(tmem_slot + 4);
  for (int i = threadIdx.x; i < 256 + 16; i += blockDim.x) s_bias[i] = (ep.bias && i < N) ? ep.bias[i] : 0.f;
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < (kExact ? 4 : 2) * Npad) tmem_cols <<= 1;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&conv[s], kConvWarps); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&t_full[b], 1); mbar_init(&t_empty[b], kEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        for (int kc = 0; kc < KC; ++kc) {
          mbar_wait(&empty[s], ph ^ 1);
          uint8_t* st = smem + (size_t)s * stage_bytes;
          mbar_expect_tx(&full[s], (uint32_t)(kStageBytesA + 2 * w_chunk_bytes));
          tma_load_2d(st, &tmA, &full[s], kc * BK32, t * BM);
          tma_load_2d(st + 2 * kStageBytesA, &tmWhi, &full[s], kc * BK32, 0);
          tma_load_2d(st + 2 * kStageBytesA + w_chunk_bytes, &tmWlo, &full[s], kc * BK32, 0);
          if (++s == kStages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 2 || warp == 3 || warp >= 12) {
    // converters: kConvWarps warps split the A tile of every stage
    const int tid = (warp >= 12 ? warp - 10 : warp - 2) * 32 + lane;
    int s = 0;
    uint32_t ph = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      for (int kc = 0; kc < KC; ++kc) {
        mbar_wait(&full[s], ph);
        uint8_t* st = smem + (size_t)s * stage_bytes;
        if (!(dbg & 1)) split_tile_tf32(st, st + kStageBytesA, kStageBytesA / 16, tid, kConvWarps * 32);
        fence_proxy_async_smem();   // generic-proxy stores -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(&conv[s]);
        if (++s == kStages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_tf32(Npad, 0, 0);
    int s = 0;
    uint32_t ph = 0;
    int it = 0;
    uint32_t pc = 0;   // kExact: pairs issued so far
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t use = (uint32_t)(it >> 1);
      if constexpr (!kExact) {
        mbar_wait(&t_empty[buf], (use & 1) ^ 1);
        tcgen05_fence_after();
      }
      const uint32_t d_tmem = tmem_base + (uint32_t)(buf * Npad);
      for (int kc = 0; kc < KC; ++kc) {
        mbar_wait(&conv[s], ph);
        tcgen05_fence_after();
        if (elect_one()) {
          const int k_left = K - kc * BK32;
          const int n_k = k_left >= BK32 ? BK32 / 8 : (k_left + 7) / 8;
          const uint32_t a_hi = smem_u32(smem + (size_t)s * stage_bytes);
          const uint32_t a_lo = a_hi + kStageBytesA;
          const uint32_t w_hi = a_hi + 2 * kStageBytesA;
          const uint32_t w_lo = w_hi + w_chunk_bytes;
          if constexpr (kExact) {
            for (int k4 = 0; k4 < n_k; k4 += 2, ++pc) {
              const uint32_t b = pc & 1;
              mbar_wait(&t_empty[b], ((pc >> 1) & 1) ^ 1);
              tcgen05_fence_after();
              const uint32_t d0 = tmem_base + b * 2u * (uint32_t)Npad, d1 = d0 + (uint32_t)Npad;
              const bool two = k4 + 1 < n_k;
              const uint64_t ah0 = make_desc(a_hi + k4 * 32, 16, 1024), al0 = make_desc(a_lo + k4 * 32, 16, 1024);
              const uint64_t wh0 = make_desc(w_hi + k4 * 32, 16, 1024), wl0 = make_desc(w_lo + k4 * 32, 16, 1024);
              const uint64_t ah1 = make_desc(a_hi + k4 * 32 + 32, 16, 1024), al1 = make_desc(a_lo + k4 * 32 + 32, 16, 1024);
              const uint64_t wh1 = make_desc(w_hi + k4 * 32 + 32, 16, 1024), wl1 = make_desc(w_lo + k4 * 32 + 32, 16, 1024);
              // the small terms of both k-steps first (their truncation is 2^-11 of the main term's), then ONE hi.hi
              // product on top of them and the other hi.hi product alone in the pair's second buffer
              tcgen05_mma_tf32(d0, al0, wh0, idesc, 0u);
              tcgen05_mma_tf32(d0, ah0, wl0, idesc, 1u);
              if (two) {
                tcgen05_mma_tf32(d0, al1, wh1, idesc, 1u);
                tcgen05_mma_tf32(d0, ah1, wl1, idesc, 1u);
              }
              tcgen05_mma_tf32(d0, ah0, wh0, idesc, 1u);
              if (two) tcgen05_mma_tf32(d1, ah1, wh1, idesc, 0u);
              tcgen05_commit(&t_full[b]);
            }
            tcgen05_commit(&empty[s]);
          } else {
          for (int k4 = 0; k4 < n_k; ++k4) {
            const uint64_t ah = make_desc(a_hi + k4 * 32, 16, 1024), al = make_desc(a_lo + k4 * 32, 16, 1024);
            const uint64_t wh = make_desc(w_hi + k4 * 32, 16, 1024), wl = make_desc(w_lo + k4 * 32, 16, 1024);
            if (dbg & 2) {
              tcgen05_mma_tf32(d_tmem, ah, wh, idesc, (kc | k4) != 0);
              continue;
            }
            tcgen05_mma_tf32(d_tmem, al, wh, idesc, (kc | k4) != 0);   // small terms first
            tcgen05_mma_tf32(d_tmem, ah, wl, idesc, 1u);
            tcgen05_mma_tf32(d_tmem, ah, wh, idesc, 1u);
          }
          tcgen05_commit(&empty[s]);
          if (kc == KC - 1) tcgen05_commit(&t_full[buf]);
          }
        }
        __syncwarp();
        if (++s == kStages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp >= 4 && warp < 12) {
    const int q = warp & 3, half = (warp - 4) >> 2;
    const int flags = (ep.bias ? 1 : 0) | (ep.row_div_ptr ? 2 : 0) | (ep.accumulate ? 4 : 0) | (ep.addend ? 8 : 0) |
                      (ep.stats ? 16 : 0);
#define EGNN_EPI_CASE(F)                                                                                      \
  case F:                                                                                                     \
    if (ep.c_dtype == EGNN_F32)                                                                               \
      epilogue_loop<float, F, kExact>(ep, s_bias, t_full, t_empty, tmem_base, Npad, M, N, n_tiles, q, half,   \
                                      lane, K);                                                               \
    else                                                                                                      \
      epilogue_loop<__nv_bfloat16, F, kExact>(ep, s_bias, t_full, t_empty, tmem_base, Npad, M, N, n_tiles, q, \
                                              half, lane, K);                                                 \
    break;
    switch (flags) {
      EGNN_EPI_CASE(0) EGNN_EPI_CASE(1) EGNN_EPI_CASE(2) EGNN_EPI_CASE(3)
      EGNN_EPI_CASE(4) EGNN_EPI_CASE(5) EGNN_EPI_CASE(6) EGNN_EPI_CASE(7)
      EGNN_EPI_CASE(8) EGNN_EPI_CASE(10)
      default:
        if constexpr (kConvWarps == 2) {
          switch (flags) { EGNN_EPI_CASE(16) EGNN_EPI_CASE(17) }
        }
    }
#undef EGNN_EPI_CASE
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols));
  }
}

// ------------------------------------------------------------------ wgrad ------------------
// dW^T tile: UMMA M axis = K_in (up to three 128-wide M tiles), N axis = N_out.
// smem stage = [X blocks: MB * 8 KB][G blocks: NB * 8 KB], each block = 64 nodes x 64 columns.
template <int kStages>
__global__ void __launch_bounds__(kThreads, 1)
gemm_wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmG,
                  const __grid_constant__ CUtensorMap tmG2, int NB1, int skip_rows_from, int skip_cols_below, int M_rows,
                  int N_out, int Nopad, int K_in, float* __restrict__ partial) {
  // the gradient operand may be TWO matrices side by side (tmG: its first NB1 64-column blocks, tmG2: the rest) --
  // dz and the residual gradient of a SAGE-ResBN layer share one pass over the layer input; output elements of rows
  // >= skip_rows_from and columns < skip_cols_below (residual gradient x aggregation half) are not needed, not written
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~uintptr_t(1023));
  const int MT = (K_in + 127) / 128;   // 128-wide M tiles over K_in (<= 3)
  const int MB = MT * 2;               // 64-column X blocks per stage
  const int NB = (Nopad + 63) / 64;    // 64-column G blocks per stage
  const int stage_bytes = (MB + NB) * 8192;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)kStages * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kStages;
  uint64_t* done = bars + 2 * kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_slabs = (M_rows + 63) / 64;
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < MT * Nopad) tmem_cols <<= 1;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const bool has_work = (int)blockIdx.x < n_slabs;

  if (warp == 0) {
    if (elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      for (int sl = blockIdx.x; sl < n_slabs; sl += gridDim.x) {
        mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx(&full[s], (uint32_t)stage_bytes);
        uint8_t* st = smem + (size_t)s * stage_bytes;
        for (int b = 0; b < MB; ++b) tma_load_2d(st + b * 8192, &tmX, &full[s], b * 64, sl * 64);
        for (int b = 0; b < NB; ++b) {
          if (b < NB1) tma_load_2d(st + (MB + b) * 8192, &tmG, &full[s], b * 64, sl * 64);
          else tma_load_2d(st + (MB + b) * 8192, &tmG2, &full[s], (b - NB1) * 64, sl * 64);
        }
        if (++s == kStages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc(Nopad, 1, 1);
    int s = 0;
    uint32_t ph = 0;
    bool first = true;
    for (int sl = blockIdx.x; sl < n_slabs; sl += gridDim.x) {
      mbar_wait(&full[s], ph);
      tcgen05_fence_after();
      if (elect_one()) {
        const uint32_t st = smem_u32(smem + (size_t)s * stage_bytes);
        for (int mt = 0; mt < MT; ++mt) {
          for (int k4 = 0; k4 < 4; ++k4) {
            uint64_t ad = make_desc(st + mt * 16384 + k4 * 2048, 8192, 1024);
            uint64_t bd = make_desc(st + MB * 8192 + k4 * 2048, 8192, 1024);
            tcgen05_mma_bf16(tmem_base + (uint32_t)(mt * Nopad), ad, bd, idesc, (!first || k4 != 0) ? 1u : 0u);
          }
        }
        tcgen05_commit(&empty[s]);
      }
      __syncwarp();
      first = false;
      if (++s == kStages) { s = 0; ph ^= 1; }
    }
    if (has_work && elect_one()) tcgen05_commit(done);
    __syncwarp();
  } else if (warp >= 4) {
    const int q = warp & 3;
    float* out = partial + (size_t)blockIdx.x * N_out * K_in;
    if (has_work) {
      mbar_wait(done, 0);
      tcgen05_fence_after();
    }
    for (int mt = 0; mt < MT; ++mt) {
      const int k = mt * 128 + q * 32 + lane;  // K_in index owned by this thread (TMEM lane)
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * Nopad);
      for (int c0 = 0; c0 < N_out; c0 += 8) {
        float v[8];
        if (has_work) {
          tmem_ld8(t_addr + c0, v);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = 0.f;
        }
        if (k < K_in && !(c0 >= skip_rows_from && k < skip_cols_below)) {
          const int nv = min(8, N_out - c0);
          for (int j = 0; j < nv; ++j) out[(size_t)(c0 + j) * K_in + k] = v[j];  // coalesced over lanes
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols));
  }
}

// ---- fp32 weight gradient: 3xTF32 -------------------------------------------------------------------------------
// dW[N_out, K_in] = G[M, N_out]^T . X[M, K_in] with fp32 operands.  Both operands are MN-major (32 fp32 = 128 bytes of one
// node's row per swizzled shared-memory row); a stage holds a slab of 16 nodes: X blocks (K_in / 32 x 2 KB), G blocks
// (N_out / 32 x 2 KB) and, behind them, the lo parts written by the converter warps.  Per slab: 2 K-steps (8 nodes) x
// (X_lo G_hi + X_hi G_lo + X_hi G_hi) per 128-wide M tile over K_in.  Per-CTA partial in TMEM, fixed-order reduction.
constexpr int kSlab32 = 16;              // nodes per stage
constexpr int kBlk32 = kSlab32 * 128;    // bytes of one 32-column block of a slab
constexpr int kThreadsWG32 = 384;        // warp 0 TMA, 1 MMA, 2-3 + 8-11 converters, 4-7 epilogue

template <int kStages>
__global__ void __launch_bounds__(kThreadsWG32, 1)
gemm_wgrad_tf32x3_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmG, int M_rows,
                         int N_out, int Nopad, int K_in, int P, float* __restrict__ partial) {
  // P phases per CTA, one partial each: the tensor core adds into its fp32 accumulator with truncation, so the error of
  // a long accumulation chain grows linearly (1e-5 after ~500 steps); every phase keeps the chain short (~130 steps)
  // and the phases are then combined with IEEE adds by wgrad_reduce_kernel.
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~uintptr_t(1023));
  const int MT = (K_in + 127) / 128;     // 128-wide M tiles over K_in (<= 3)
  const int XB = MT * 4;                 // 32-column X blocks per stage
  const int GB = (Nopad + 31) / 32;      // 32-column G blocks per stage
  const int half_bytes = (XB + GB) * kBlk32;   // hi part of a stage; the lo part follows
  const int stage_bytes = 2 * half_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)kStages * stage_bytes);
  uint64_t* full = bars;
  uint64_t* conv = bars + kStages;
  uint64_t* empty = bars + 2 * kStages;
  uint64_t* done = bars + 3 * kStages;      // MMA -> epilogue: a phase's accumulator is complete
  uint64_t* drained = done + 1;             // epilogue -> MMA: the accumulator has been read
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(drained + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_slabs = (M_rows + kSlab32 - 1) / kSlab32;
  const int n_my = (int)blockIdx.x < n_slabs ? (n_slabs - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < MT * Nopad) tmem_cols <<= 1;
  constexpr int kConvWarps = 6;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&conv[s], kConvWarps); mbar_init(&empty[s], 1); }
    mbar_init(done, 1);
    mbar_init(drained, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      for (int it = 0; it < n_my; ++it) {
        const int sl = blockIdx.x + it * gridDim.x;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx(&full[s], (uint32_t)half_bytes);
        uint8_t* st = smem + (size_t)s * stage_bytes;
        for (int b = 0; b < XB; ++b) tma_load_2d(st + b * kBlk32, &tmX, &full[s], b * 32, sl * kSlab32);
        for (int b = 0; b < GB; ++b) tma_load_2d(st + (XB + b) * kBlk32, &tmG, &full[s], b * 32, sl * kSlab32);
        if (++s == kStages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 2 || warp == 3 || warp >= 8) {
    const int tid = (warp >= 8 ? warp - 6 : warp - 2) * 32 + lane;
    int s = 0;
    uint32_t ph = 0;
    for (int it = 0; it < n_my; ++it) {
      mbar_wait(&full[s], ph);
      uint8_t* st = smem + (size_t)s * stage_bytes;
      split_tile_tf32(st, st + half_bytes, half_bytes / 16, tid, kConvWarps * 32);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&conv[s]);
      if (++s == kStages) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_tf32(Nopad, 1, 1);
    int s = 0;
    uint32_t ph = 0;
    for (int p = 0; p < P; ++p) {
      const int it0 = (int)((int64_t)n_my * p / P), it1 = (int)((int64_t)n_my * (p + 1) / P);
      if (p > 0) {
        mbar_wait(drained, (uint32_t)((p - 1) & 1));
        tcgen05_fence_after();
      }
      bool first = true;
      for (int it = it0; it < it1; ++it) {
        mbar_wait(&conv[s], ph);
        tcgen05_fence_after();
        if (elect_one()) {
          const uint32_t hi = smem_u32(smem + (size_t)s * stage_bytes), lo = hi + half_bytes;
          for (int mt = 0; mt < MT; ++mt) {
            for (int k2 = 0; k2 < kSlab32 / 8; ++k2) {
              // operand descriptors: 8 nodes = two 4-row (512-byte) swizzle groups; the next 32 columns are kBlk32 bytes further
              const uint32_t xo = mt * 4 * kBlk32 + k2 * 1024, go = XB * kBlk32 + k2 * 1024;
              const uint64_t xh = make_desc_b32(hi + xo, kBlk32, 512), xl = make_desc_b32(lo + xo, kBlk32, 512);
              const uint64_t gh = make_desc_b32(hi + go, kBlk32, 512), gl = make_desc_b32(lo + go, kBlk32, 512);
              const uint32_t d = tmem_base + (uint32_t)(mt * Nopad);
              tcgen05_mma_tf32(d, xl, gh, idesc, (!first || k2 != 0) ? 1u : 0u);
              tcgen05_mma_tf32(d, xh, gl, idesc, 1u);
              tcgen05_mma_tf32(d, xh, gh, idesc, 1u);
            }
          }
          tcgen05_commit(&empty[s]);
        }
        __syncwarp();
        first = false;
        if (++s == kStages) { s = 0; ph ^= 1; }
      }
      if (elect_one()) tcgen05_commit(done);
      __syncwarp();
    }
  } else if (warp >= 4 && warp < 8) {
    const int q = warp & 3;
    for (int p = 0; p < P; ++p) {
      const int it0 = (int)((int64_t)n_my * p / P), it1 = (int)((int64_t)n_my * (p + 1) / P);
      const bool has_work = it1 > it0;
      float* out = partial + ((size_t)blockIdx.x * P + p) * N_out * K_in;
      mbar_wait(done, (uint32_t)(p & 1));
      tcgen05_fence_after();
      for (int mt = 0; mt < MT; ++mt) {
        const int k = mt * 128 + q * 32 + lane;
        const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * Nopad);
        for (int c0 = 0; c0 < N_out; c0 += 8) {
          float v[8];
          if (has_work) {
            tmem_ld8(t_addr + c0, v);
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = 0.f;
          }
          if (k < K_in) {
            const int nv = min(8, N_out - c0);
            for (int j = 0; j < nv; ++j) out[(size_t)(c0 + j) * K_in + k] = v[j];
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(drained);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols));
  }
}

// block = 64 output elements x 4 partial lanes; lane q sums partials q, q+4, ... with four independent
// accumulators (loads in flight), then the four lanes are combined in a fixed order: deterministic
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ partial, int n_part,
                                                           int64_t n_elem, float* __restrict__ out,
                                                           int accumulate) {
  __shared__ float sm[4][64];
  const int el = threadIdx.x & 63, q = threadIdx.x >> 6;
  const int64_t i = (int64_t)blockIdx.x * 64 + el;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (i < n_elem) {
    int p = q;
    for (; p + 12 < n_part; p += 16) {
      s0 += partial[(size_t)p * n_elem + i];
      s1 += partial[(size_t)(p + 4) * n_elem + i];
      s2 += partial[(size_t)(p + 8) * n_elem + i];
      s3 += partial[(size_t)(p + 12) * n_elem + i];
    }
    for (; p < n_part; p += 4) s0 += partial[(size_t)p * n_elem + i];
  }
  sm[q][el] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (q == 0 && i < n_elem) {
    const float s = (sm[0][el] + sm[1][el]) + (sm[2][el] + sm[3][el]);
    out[i] = accumulate ? out[i] + s : s;
  }
}

// the same reduction writing the [N_out, K_in] result as two parameter gradients: columns [0, split) -> dst0, columns
// [split, K_in) -> dst1 (may be null), each [N_out, valid] dense with the zero-padding columns >= valid dropped --
// dW_l / dW_r of a SAGE layer straight into the flat gradient buffer, no slicing copies
__global__ void __launch_bounds__(256) wgrad_reduce_split_kernel(const float* __restrict__ partial, int n_part,
                                                                 int N_out, int K_in, float* __restrict__ dst0,
                                                                 float* __restrict__ dst1, int split, int valid,
                                                                 int N1, float* __restrict__ dst2) {
  __shared__ float sm[4][64];
  const int el = threadIdx.x & 63, q = threadIdx.x >> 6;
  const int64_t n_elem = (int64_t)N_out * K_in;
  const int64_t i = (int64_t)blockIdx.x * 64 + el;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  const int n = (int)(i / K_in), k = (int)(i - (int64_t)n * K_in);
  const bool wanted = i < n_elem && !(n >= N1 && k < split);   // (second operand) x (first half) was never written
  if (wanted) {
    int p = q;
    for (; p + 12 < n_part; p += 16) {
      s0 += partial[(size_t)p * n_elem + i];
      s1 += partial[(size_t)(p + 4) * n_elem + i];
      s2 += partial[(size_t)(p + 8) * n_elem + i];
      s3 += partial[(size_t)(p + 12) * n_elem + i];
    }
    for (; p < n_part; p += 4) s0 += partial[(size_t)p * n_elem + i];
  }
  sm[q][el] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (q == 0 && wanted) {
    const float s = (sm[0][el] + sm[1][el]) + (sm[2][el] + sm[3][el]);
    if (n >= N1) {
      if (k - split < valid) dst2[(size_t)(n - N1) * valid + (k - split)] = s;
    } else if (k < split) {
      if (k < valid) dst0[(size_t)n * valid + k] = s;
    } else if (dst1 && k - split < valid) {
      dst1[(size_t)n * valid + (k - split)] = s;
    }
  }
}

// ------------------------------------------------------------------ host side ---------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  // cuTensorMapEncodeTiled is a DRIVER call: it needs a current context on the calling thread and returns
  // CUDA_ERROR_INVALID_CONTEXT (201) otherwise.  A thread whose first piece of CUDA work is a tensor-core GEMM -- the
  // autograd engine's worker when a backward starts with a weight gradient -- has none yet (the runtime binds the
  // primary context lazily, on its own first call): bind it once per thread.
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {
    cudaFree(nullptr);
    ctx_bound = true;
  }
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// bf16 [rows, cols] row-major with leading dimension ld (elements); box = 64 cols x box_rows
bool make_map(CUtensorMap* m, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

thread_local char g_map_err[256] = "cuTensorMapEncodeTiled failed";
// fp32 [rows, cols] row-major, box = 32 cols (128 bytes) x box_rows
bool make_map_f32(CUtensorMap* m, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                  CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    snprintf(g_map_err, sizeof(g_map_err), "cuTensorMapEncodeTiled(f32) -> %d: base %p rows %lld cols %lld ld %lld box_rows %d swizzle %d",
             (int)r, base, (long long)rows, (long long)cols, (long long)ld, box_rows, (int)swz);
  return r == CUDA_SUCCESS;
}

constexpr int kStagesTN = 6;
constexpr size_t kMaxDynSmemTN = 227 * 1024;
constexpr size_t kStatsSmem = 2048;   // [4 sub-partitions][sum, sumsq][64] floats behind the bias (statistics epilogue)
constexpr int kStagesWG = 4;

}  // namespace

bool gemm_tcgen05_supported(int64_t lda, int64_t ldb, int64_t ld_c, int64_t M, int64_t N, int64_t K,
                            const void* A, const void* B, const void* C) {
  (void)ld_c; (void)C;
  if (M < 1 || N < 8 || N > 256 || K < 8 || K > 512) return false;
  if (lda % 8 || ldb % 8) return false;                          // TMA: 16-byte global strides
  if (((uintptr_t)A | (uintptr_t)B) & 15) return false;
  if (M >= (int64_t)1 << 31) return false;
  const int Npad = (int)((N + 15) / 16 * 16);
  const int KC = (int)((K + BK - 1) / BK);
  size_t smem = (size_t)KC * Npad * 128 + (size_t)kStagesTN * kStageBytesA + 256 + 1088 + 1024;
  return smem <= kMaxDynSmemTN && 2 * Npad <= 512;
}

// which (bias, row_div, accumulate, addend, stats) combinations gemm_tn_kernel has an epilogue for
bool gemm_tcgen05_epilogue_supported(bool bias, bool row_div, bool accumulate, bool addend, bool stats, int64_t stats_cols,
                                     int64_t add_col0) {
  const int flags = (bias ? 1 : 0) | (row_div ? 2 : 0) | (accumulate ? 4 : 0) | (addend ? 8 : 0) | (stats ? 16 : 0);
  if (flags < 8) return true;
  if (addend && add_col0 % 16 != 0) return false;
  if (stats && (stats_cols < 1 || stats_cols > 64)) return false;
  return flags == 8 || flags == 10 || flags == 16 || flags == 17;
}

int64_t gemm_tcgen05_stats_parts(int64_t M) {
  const int64_t n_tiles = (M + BM - 1) / BM;
  return n_tiles < kNumSMs ? n_tiles : kNumSMs;
}

int gemm_tcgen05_dispatch(const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int c_dtype,
                          int64_t ld_c, int64_t M, int64_t N, int64_t K, const float* bias, int accumulate,
                          const int32_t* row_div_ptr, int64_t row_div_cols, cudaStream_t st, const void* addend,
                          int64_t ld_add, int64_t add_col0, float* stats, int64_t stats_cols) {
  const char* fn = "egnn_gemm(tcgen05)";
  const int Npad = (int)((N + 15) / 16 * 16);
  const int KC = (int)((K + BK - 1) / BK);
  CUtensorMap tmA, tmW;
  if (!make_map(&tmA, A, M, K, lda, BM) || !make_map(&tmW, B, N, K, ldb, Npad))
    return fail(fn, "cuTensorMapEncodeTiled failed");
  size_t smem = (size_t)KC * Npad * 128 + (size_t)kStagesTN * kStageBytesA + 256 + 1088 + 1024 + (stats ? kStatsSmem : 0);
  if (smem > kMaxDynSmemTN) return fail(fn, "shared memory: statistics epilogue does not fit this shape");
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(gemm_tn_kernel<kStagesTN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmemTN);
    attr_set = true;
  }
  int n_tiles = (int)((M + BM - 1) / BM);
  int grid = n_tiles < kNumSMs ? n_tiles : kNumSMs;
  Epilogue ep{C, bias, row_div_ptr, ld_c, c_dtype, accumulate, (int)row_div_cols, addend, ld_add, (int)add_col0,
              stats, (int)stats_cols};
  gemm_tn_kernel<kStagesTN><<<grid, kThreadsTN, smem, st>>>(tmA, tmW, (int)M, (int)N, Npad, (int)K, ep);
  EGNN_LAUNCH_CHECK(fn);
  return 0;
}

// ---- fp32 operands (3xTF32) ----------------------------------------------------------------------------------------
std::atomic<int> g_f32_tc_exact{1};
static int tf32_stages(int Npad, bool stats) {
  const size_t stage = 2 * (size_t)kStageBytesA + 2 * (size_t)Npad * 128;
  const size_t fixed = 256 + 1088 + 1024 + (stats ? kStatsSmem : 0) + 64;
  int st = 4;
  while (st > 1 && st * stage + fixed > kMaxDynSmemTN) --st;
  return st;
}

bool gemm_tf32x3_supported(int64_t lda, int64_t ldw, int64_t M, int64_t N, int64_t K, const void* A, const void* W) {
  if (M < 1 || N < 8 || N > 256 || K < 8 || K > 4096) return false;
  if (lda % 4 || ldw % 4 || K % 4) return false;                  // TMA: 16-byte global strides
  if (((uintptr_t)A | (uintptr_t)W) & 15) return false;
  if (M >= (int64_t)1 << 31) return false;
  const int Npad = (int)((N + 15) / 16 * 16);
  return tf32_stages(Npad, true) >= 2 && 2 * Npad <= 512;
}

size_t gemm_tf32x3_workspace_floats(int64_t N, int64_t K) { return 2 * (size_t)N * (size_t)K + 64; }

// C[M,N] = A[M,K] . W[N,K]^T, fp32 operands, fp32 accumulate; workspace >= gemm_tf32x3_workspace_floats(N, K) floats
int gemm_tf32x3_dispatch(const void* A, int64_t lda, const void* W, int64_t ldw, void* C, int c_dtype, int64_t ld_c,
                         int64_t M, int64_t N, int64_t K, const float* bias, int accumulate,
                         const int32_t* row_div_ptr, int64_t row_div_cols, float* workspace, cudaStream_t st,
                         const void* addend, int64_t ld_add, int64_t add_col0, float* stats, int64_t stats_cols) {
  const char* fn = "egnn_gemm(tcgen05 3xTF32)";
  const int Npad = (int)((N + 15) / 16 * 16);
  if (Npad > 128 && g_f32_tc_exact.load(std::memory_order_relaxed) != 0) {
    // exact accumulation holds one 128-column tile in TMEM: a wider product (the [dm | dx_root] dgrad of a 128-wide
    // fp32 SAGE layer) runs as column blocks of 128 -- A is read once per block, every block exact
    const size_t c_es = c_dtype == EGNN_F32 ? 4 : 2;
    for (int64_t c0 = 0; c0 < N; c0 += 128) {
      const int64_t n = N - c0 < 128 ? N - c0 : 128;
      const int32_t* rd = row_div_ptr;
      int64_t rdc = row_div_cols;
      if (rd && row_div_cols > 0) {
        rdc = row_div_cols - c0;
        if (rdc <= 0) rd = nullptr;
        else if (rdc > n) rdc = n;
      }
      const void* ad = nullptr;
      int64_t ac0 = 0;
      if (addend && add_col0 < c0 + n) {
        ac0 = add_col0 > c0 ? add_col0 - c0 : 0;
        ad = reinterpret_cast<const char*>(addend) + (size_t)(c0 > add_col0 ? c0 - add_col0 : 0) * c_es;
      }
      int rc = gemm_tf32x3_dispatch(A, lda, reinterpret_cast<const float*>(W) + c0 * ldw, ldw,
                                    reinterpret_cast<char*>(C) + (size_t)c0 * c_es, c_dtype, ld_c, M, n, K,
                                    bias ? bias + c0 : nullptr, accumulate, rd, rdc, workspace, st, ad, ld_add, ac0,
                                    c0 == 0 ? stats : nullptr, c0 == 0 ? stats_cols : 0);
      if (rc) return rc;
    }
    return 0;
  }
  // dense [N, K] hi / lo copies of the weight (16-byte aligned halves of the workspace)
  float* whi = reinterpret_cast<float*>(((uintptr_t)workspace + 15) & ~uintptr_t(15));
  float* wlo = whi + ((N * K + 3) / 4) * 4;
  if (ldw == K) {
    split_tf32_kernel<<<(unsigned)ceil_div(N * K, 256), 256, 0, st>>>((const float*)W, N * K, whi, wlo);
    EGNN_LAUNCH_CHECK(fn);
  } else {
    for (int64_t r = 0; r < N; ++r) {   // strided weight: row by row (rare; the layers pass dense weights)
      split_tf32_kernel<<<(unsigned)ceil_div(K, 256), 256, 0, st>>>((const float*)W + r * ldw, K, whi + r * K, wlo + r * K);
      EGNN_LAUNCH_CHECK(fn);
    }
  }
  CUtensorMap tmA, tmWhi, tmWlo;
  if (!make_map_f32(&tmA, A, M, K, lda, BM) || !make_map_f32(&tmWhi, whi, N, K, K, Npad) ||
      !make_map_f32(&tmWlo, wlo, N, K, K, Npad))
    return fail(fn, "cuTensorMapEncodeTiled failed");
  const int stages = tf32_stages(Npad, stats != nullptr);
  if (stages < 2) return fail(fn, "shared memory: shape does not fit");
  const size_t stage = 2 * (size_t)kStageBytesA + 2 * (size_t)Npad * 128;
  const size_t smem = stages * stage + 256 + 1088 + 1024 + (stats ? kStatsSmem : 0) + 64;
  static bool attr_set = false;
  if (!attr_set) {
#define EGNN_TF32_ATTR(S, C, X) \
  cudaFuncSetAttribute(gemm_tn_tf32x3_kernel<S, C, X>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmemTN)
    EGNN_TF32_ATTR(4, 2, false); EGNN_TF32_ATTR(3, 2, false); EGNN_TF32_ATTR(2, 2, false);
    EGNN_TF32_ATTR(4, 6, false); EGNN_TF32_ATTR(3, 6, false); EGNN_TF32_ATTR(2, 6, false);
    EGNN_TF32_ATTR(4, 2, true); EGNN_TF32_ATTR(3, 2, true); EGNN_TF32_ATTR(2, 2, true);
    EGNN_TF32_ATTR(4, 6, true); EGNN_TF32_ATTR(3, 6, true); EGNN_TF32_ATTR(2, 6, true);
#undef EGNN_TF32_ATTR
    attr_set = true;
  }
  const int n_tiles = (int)((M + BM - 1) / BM);
  const int grid = n_tiles < kNumSMs ? n_tiles : kNumSMs;
  Epilogue ep{C, bias, row_div_ptr, ld_c, c_dtype, accumulate, (int)row_div_cols, addend, ld_add, (int)add_col0,
              stats, (int)stats_cols};
  static int dbg = -1;
  if (dbg < 0) { const char* e = getenv("EGNN_TF32_DEBUG"); dbg = e ? atoi(e) : 0; }
#define EGNN_TF32_LAUNCH(S, C, X)                                                                                  \
  gemm_tn_tf32x3_kernel<S, C, X><<<grid, kThreadsTN + (C - 2) * 32, smem, st>>>(tmA, tmWhi, tmWlo, (int)M, (int)N, Npad, \
                                                                                (int)K, ep, dbg)
  // exact accumulation (IEEE adds of the k-step products outside the tensor core) wherever the TMEM ring fits,
  // unless the caller asked for the in-TMEM accumulate (egnn_set_f32_tc_exact(0): no-grad forwards)
  const bool exact = Npad <= 128 && !(dbg & 4) && g_f32_tc_exact.load(std::memory_order_relaxed) != 0;
  if (exact) {
    if (stats) {   // statistics epilogue: the 384-thread variant
      if (stages == 4) EGNN_TF32_LAUNCH(4, 2, true); else if (stages == 3) EGNN_TF32_LAUNCH(3, 2, true); else EGNN_TF32_LAUNCH(2, 2, true);
    } else {
      if (stages == 4) EGNN_TF32_LAUNCH(4, 6, true); else if (stages == 3) EGNN_TF32_LAUNCH(3, 6, true); else EGNN_TF32_LAUNCH(2, 6, true);
    }
  } else if (stats) {
    if (stages == 4) EGNN_TF32_LAUNCH(4, 2, false); else if (stages == 3) EGNN_TF32_LAUNCH(3, 2, false); else EGNN_TF32_LAUNCH(2, 2, false);
  } else {
    if (stages == 4) EGNN_TF32_LAUNCH(4, 6, false); else if (stages == 3) EGNN_TF32_LAUNCH(3, 6, false); else EGNN_TF32_LAUNCH(2, 6, false);
  }
#undef EGNN_TF32_LAUNCH
  EGNN_LAUNCH_CHECK(fn);
  return 0;
}

// phases (= partials) per CTA of the fp32 wgrad: accumulation chains of at most ~352 rows (132 tensor-core adds)
constexpr int kMaxPhases32 = 8;
static int wgrad32_phases(int64_t M_rows) {
  const int64_t rows_per_cta = (M_rows + kNumSMs - 1) / kNumSMs;
  int64_t P = (rows_per_cta + 351) / 352;
  return (int)(P < 1 ? 1 : (P > kMaxPhases32 ? kMaxPhases32 : P));
}

static int wgrad32_stages(int XB, int GB) {
  const size_t stage = 2 * (size_t)(XB + GB) * kBlk32;
  int st = 6;
  while (st > 1 && st * stage + 256 + 1024 > 227 * 1024) --st;
  return st;
}

bool wgrad_tf32x3_supported(const void* G, int64_t ldg, const void* X, int64_t ldx, int64_t M_rows, int64_t N_out,
                            int64_t K_in) {
  if (N_out < 8 || N_out > 256 || K_in < 8 || K_in > 384 || M_rows < 1 || M_rows >= (int64_t)1 << 31) return false;
  if (ldg % 4 || ldx % 4 || N_out % 4 || K_in % 4) return false;
  if (((uintptr_t)G | (uintptr_t)X) & 15) return false;
  const int Nopad = (int)((N_out + 15) / 16 * 16);
  const int MT = (int)((K_in + 127) / 128);
  return wgrad32_stages(MT * 4, (Nopad + 31) / 32) >= 2 && MT * Nopad <= 512;
}

// dW[N_out, K_in] (fp32 dense) = G[M, N_out]^T X[M, K_in], fp32 operands; workspace >= wgrad_tcgen05_workspace_floats
int wgrad_tf32x3_dispatch(const void* G, int64_t ldg, const void* X, int64_t ldx, float* dW, int64_t M_rows,
                          int64_t N_out, int64_t K_in, int accumulate, float* workspace, cudaStream_t st, float* dst1,
                          int64_t split, int64_t valid) {
  const char* fn = "egnn_gemm(tcgen05 3xTF32 wgrad)";
  const int Nopad = (int)((N_out + 15) / 16 * 16);
  const int MT = (int)((K_in + 127) / 128);
  const int XB = MT * 4, GB = (Nopad + 31) / 32;
  CUtensorMap tmX, tmG;
  if (!make_map_f32(&tmX, X, M_rows, K_in, ldx, kSlab32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) ||
      !make_map_f32(&tmG, G, M_rows, N_out, ldg, kSlab32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))
    return fail(fn, g_map_err);
  const int stages = wgrad32_stages(XB, GB);
  const size_t smem = (size_t)stages * 2 * (XB + GB) * kBlk32 + 256 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(gemm_wgrad_tf32x3_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(gemm_wgrad_tf32x3_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(gemm_wgrad_tf32x3_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(gemm_wgrad_tf32x3_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr_set = true;
  }
  const int n_slabs = (int)((M_rows + kSlab32 - 1) / kSlab32);
  const int grid = n_slabs < kNumSMs ? n_slabs : kNumSMs;
  const int P = wgrad32_phases(M_rows);
  if (stages >= 6)
    gemm_wgrad_tf32x3_kernel<6><<<grid, kThreadsWG32, smem, st>>>(tmX, tmG, (int)M_rows, (int)N_out, Nopad, (int)K_in, P, workspace);
  else if (stages >= 4)
    gemm_wgrad_tf32x3_kernel<4><<<grid, kThreadsWG32, (size_t)4 * 2 * (XB + GB) * kBlk32 + 256 + 1024, st>>>(
        tmX, tmG, (int)M_rows, (int)N_out, Nopad, (int)K_in, P, workspace);
  else if (stages == 3)
    gemm_wgrad_tf32x3_kernel<3><<<grid, kThreadsWG32, smem, st>>>(tmX, tmG, (int)M_rows, (int)N_out, Nopad, (int)K_in, P, workspace);
  else
    gemm_wgrad_tf32x3_kernel<2><<<grid, kThreadsWG32, smem, st>>>(tmX, tmG, (int)M_rows, (int)N_out, Nopad, (int)K_in, P, workspace);
  EGNN_LAUNCH_CHECK(fn);
  const int64_t n_elem = N_out * K_in;
  if (valid > 0)
    wgrad_reduce_split_kernel<<<(unsigned)ceil_div(n_elem, 64), 256, 0, st>>>(workspace, grid * P, (int)N_out, (int)K_in,
                                                                             dW, dst1, (int)split, (int)valid, 1 << 30,
                                                                             nullptr);
  else
    wgrad_reduce_kernel<<<(unsigned)ceil_div(n_elem, 64), 256, 0, st>>>(workspace, grid * P, n_elem, dW, accumulate);
  EGNN_LAUNCH_CHECK(fn);
  return 0;
}

size_t wgrad_tcgen05_workspace_floats(int64_t N_out, int64_t K_in) {
  return (size_t)kNumSMs * kMaxPhases32 * N_out * K_in;   // (the fp32 kernel writes up to kMaxPhases32 partials per CTA)
}

bool wgrad_tcgen05_supported(const void* G, int64_t ldg, const void* X, int64_t ldx, int64_t M_rows, int64_t N_out,
                             int64_t K_in) {
  if (N_out < 8 || N_out > 256 || K_in < 8 || K_in > 384 || M_rows < 1 || M_rows >= (int64_t)1 << 31) return false;
  if (ldg % 8 || ldx % 8) return false;
  if (((uintptr_t)G | (uintptr_t)X) & 15) return false;
  const int Nopad = (int)((N_out + 15) / 16 * 16);
  const int MT = (int)((K_in + 127) / 128);
  const int stage = (MT * 2 + (Nopad + 63) / 64) * 8192;
  return (size_t)2 * stage + 256 + 1024 <= 227 * 1024 && MT * Nopad <= 512;  // at least a 2-deep ring
}

// dW[N_out, K_in] (fp32, ld = K_in) (+)= G[M,N_out]^T X[M,K_in]; workspace >= wgrad_tcgen05_workspace_floats
int wgrad_tcgen05_dispatch(const void* G, int64_t ldg, const void* X, int64_t ldx, float* dW, int64_t M_rows,
                           int64_t N_out, int64_t K_in, int accumulate, float* workspace, cudaStream_t st,
                           float* dst1, int64_t split, int64_t valid, const void* G2, int64_t ldg2, int64_t N2,
                           float* dst2) {
  const char* fn = "egnn_gemm(tcgen05 wgrad)";
  const int64_t N1 = N_out;          // columns of the first gradient operand
  if (G2) N_out = N1 + N2;           // [G | G2]: N1 must be a multiple of 64 (whole TMA boxes)
  const int Nopad = (int)((N_out + 15) / 16 * 16);
  const int MT = (int)((K_in + 127) / 128);
  const int stage = (MT * 2 + (Nopad + 63) / 64) * 8192;
  CUtensorMap tmX, tmG, tmG2;
  if (!make_map(&tmX, X, M_rows, K_in, ldx, 64) || !make_map(&tmG, G, M_rows, N1, ldg, 64) ||
      !make_map(&tmG2, G2 ? G2 : G, M_rows, G2 ? N2 : N1, G2 ? ldg2 : ldg, 64))
    return fail(fn, "cuTensorMapEncodeTiled failed");
  const int NB1 = G2 ? (int)(N1 / 64) : 1 << 20;
  const int skip_rows = G2 ? (int)N1 : 1 << 30, skip_cols = G2 ? (int)split : 0;
  // deepest TMA ring that fits: 4 stages for the 64-wide layers, 3 / 2 for the wide concatenated operands
  int stages = kStagesWG;
  while (stages > 2 && (size_t)stages * stage + 256 + 1024 > 227 * 1024) --stages;
  size_t smem = (size_t)stages * stage + 256 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(gemm_wgrad_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(gemm_wgrad_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(gemm_wgrad_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr_set = true;
  }
  int n_slabs = (int)((M_rows + 63) / 64);
  int grid = n_slabs < kNumSMs ? n_slabs : kNumSMs;
  if (stages == 4)
    gemm_wgrad_kernel<4><<<grid, kThreads, smem, st>>>(tmX, tmG, tmG2, NB1, skip_rows, skip_cols, (int)M_rows, (int)N_out,
                                                       Nopad, (int)K_in, workspace);
  else if (stages == 3)
    gemm_wgrad_kernel<3><<<grid, kThreads, smem, st>>>(tmX, tmG, tmG2, NB1, skip_rows, skip_cols, (int)M_rows, (int)N_out,
                                                       Nopad, (int)K_in, workspace);
  else
    gemm_wgrad_kernel<2><<<grid, kThreads, smem, st>>>(tmX, tmG, tmG2, NB1, skip_rows, skip_cols, (int)M_rows, (int)N_out,
                                                       Nopad, (int)K_in, workspace);
  EGNN_LAUNCH_CHECK(fn);
  int64_t n_elem = N_out * K_in;
  if (valid > 0)
    wgrad_reduce_split_kernel<<<(unsigned)ceil_div(n_elem, 64), 256, 0, st>>>(workspace, grid, (int)N_out, (int)K_in, dW,
                                                                             dst1, (int)split, (int)valid,
                                                                             G2 ? (int)N1 : 1 << 30, dst2);
  else
    wgrad_reduce_kernel<<<(unsigned)ceil_div(n_elem, 64), 256, 0, st>>>(workspace, grid, n_elem, dW, accumulate);
  EGNN_LAUNCH_CHECK(fn);
  return 0;
}

}  // namespace egnn

extern "C" int egnn_set_f32_tc_exact(int exact) { return egnn::g_f32_tc_exact.exchange(exact ? 1 : 0); }
