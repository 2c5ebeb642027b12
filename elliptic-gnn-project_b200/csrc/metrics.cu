// Epoch tail on the device: validation PR-AUC and the early-stopping bookkeeping.
//
// The reference ends every epoch on the host (src/train_gnn.py:248-257,387-402): softmax -> .cpu().numpy() ->
// boolean-mask indexing -> sklearn `average_precision_score` (src/utils/metrics.py:11-13) -> `if pr_val > best_val`
// -> a CPU clone of the whole state dict.  Each of these is a device synchronisation that costs more than the
// ~0.7 ms train step.  Here the same quantities are produced by kernels and stay on the device:
//   egnn_average_precision : AP (and ROC-AUC, out[4]) of the positive class (label == 1) over the selected rows, sklearn's definition
//       AP = sum over distinct thresholds k (descending score) of (R_k - R_{k-1}) * P_k,  R = tp / P_total,
//       P_k = tp_k / (tp_k + fp_k)   (precision_recall_curve + step integral; ties share one threshold),
//     computed as: descending order-preserving 32-bit keys -> stable LSD radix sort (radix.cuh, four 8-bit passes)
//     -> one CTA walks the sorted run in 1024 contiguous chunks: integer tp counts (exact), tie-group ends, the
//     per-threshold float64 terms, fixed-order float64 reduction (deterministic).
//   egnn_early_stop_update : best/bad/epoch bookkeeping plus the best-parameter snapshot, conditional on the
//     device-side comparison (no host round trip; the host polls `state` when it wants to).
// HBM-bound integer / streaming work: sort passes move 16 bytes per row per pass.
#include "radix.cuh"

namespace egnn {
namespace {

constexpr int kFinThreads = 1024;

// float -> unsigned key whose ASCENDING order is DESCENDING float order (total order, -0 < +0, NaNs at the ends)
__device__ __forceinline__ unsigned desc_key(float f) {
  unsigned b = __float_as_uint(f);
  unsigned u = (b & 0x80000000u) ? ~b : (b | 0x80000000u);  // ascending order-preserving
  return ~u;
}

__device__ __forceinline__ float prob_pos(const float* __restrict__ logits, int64_t ld, int64_t i) {
  // softmax(logits[i, 0:2])[1] the way torch.softmax evaluates it in fp32: exp(l - max) / sum
  const float a = logits[i * ld], b = logits[i * ld + 1];
  const float m = fmaxf(a, b);
  const float ea = expf(a - m), eb = expf(b - m);
  return eb / (ea + eb);
}

// keys for every row (unselected rows sort to the very end), selected / positive counts, n -> device
__global__ void __launch_bounds__(kThreads) ap_keys(const float* __restrict__ logits, int64_t ld,
                                                    const float* __restrict__ scores, const uint8_t* __restrict__ mask,
                                                    const int64_t* __restrict__ y, int64_t n, int* __restrict__ keys,
                                                    float* __restrict__ score_out, int* __restrict__ counters) {
  __shared__ int s_cnt[2];
  if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i < n) {
    const bool sel = mask ? mask[i] != 0 : true;
    const float s = scores ? scores[i] : prob_pos(logits, ld, i);
    if (score_out) score_out[i] = s;
    keys[i] = sel ? (int)desc_key(s) : (int)0xffffffffu;
    if (sel) {
      atomicAdd(&s_cnt[0], 1);
      if (y[i] == 1) atomicAdd(&s_cnt[1], 1);
    }
  }
  __syncthreads();
  if (threadIdx.x < 2 && s_cnt[threadIdx.x]) atomicAdd(counters + threadIdx.x, s_cnt[threadIdx.x]);
  if (i == 0) counters[2] = (int)n;
}

__global__ void ap_clear(int* counters) {
  if (threadIdx.x < 4) counters[threadIdx.x] = 0;
}

// one CTA: walk the first n_sel sorted entries in kFinThreads contiguous chunks
__global__ void __launch_bounds__(kFinThreads) ap_finish(const int* __restrict__ keys, const int* __restrict__ perm,
                                                         const int64_t* __restrict__ y, const int* __restrict__ counters,
                                                         double* __restrict__ out) {
  __shared__ int s_tp[kFinThreads];       // positives in the chunk -> exclusive prefix
  __shared__ int s_end_tp[kFinThreads];   // cumulative positives at the chunk's last threshold (-1: chunk has none)
  __shared__ int s_groups[kFinThreads];
  __shared__ int s_end_cnt[kFinThreads];  // entries up to and including the chunk's last threshold (-1: none)
  __shared__ double s_sum[kFinThreads];
  const int n_sel = counters[0], n_pos = counters[1];
  const int t = threadIdx.x;
  const int chunk = (n_sel + kFinThreads - 1) / kFinThreads;
  const int lo = min(t * chunk, n_sel), hi = min(lo + chunk, n_sel);
  // pass 1: positives per chunk, positives up to the chunk's last threshold
  int tp = 0, tp_at_end = -1, cnt_at_end = -1, groups = 0;
  for (int i = lo; i < hi; ++i) {
    tp += y[perm[i]] == 1;
    if (i == n_sel - 1 || keys[i] != keys[i + 1]) {
      tp_at_end = tp;
      cnt_at_end = i + 1;
      ++groups;
    }
  }
  s_tp[t] = tp;
  s_end_tp[t] = tp_at_end;
  s_end_cnt[t] = cnt_at_end;
  s_groups[t] = groups;
  __syncthreads();
  if (t == 0) {  // 1024 sequential steps: exclusive prefix of the counts, cumulative counts at the previous threshold
    int run = 0, prev_end = 0, prev_cnt = 0, g = 0;
    for (int q = 0; q < kFinThreads; ++q) {
      const int c = s_tp[q], e = s_end_tp[q], ec = s_end_cnt[q];
      s_tp[q] = run;
      s_end_tp[q] = prev_end;
      s_end_cnt[q] = prev_cnt;
      if (e >= 0) {
        prev_end = run + e;
        prev_cnt = ec;
      }
      run += c;
      g += s_groups[q];
    }
    s_groups[0] = g;
  }
  __syncthreads();
  // pass 2: the per-threshold terms (R_k - R_{k-1}) * P_k in float64, like numpy's float64 arrays
  // and, over the same thresholds, the ROC trapezoids (fpr_k - fpr_{k-1}) * (tpr_k + tpr_{k-1}) / 2
  // (`roc_auc_illicit` = sklearn roc_auc_score, src/utils/metrics.py:15-16; collinear points do not change the area)
  double sum = 0.0, roc = 0.0;
  int run = s_tp[t], prev = s_end_tp[t], prev_cnt = s_end_cnt[t];
  const double P = (double)n_pos, Nn = (double)(n_sel - n_pos);
  for (int i = lo; i < hi; ++i) {
    run += y[perm[i]] == 1;
    if (i == n_sel - 1 || keys[i] != keys[i + 1]) {
      if (n_pos > 0) {
        const double recall = (double)run / P, recall_prev = (double)prev / P;
        const double precision = (double)run / (double)(i + 1);
        sum += (recall - recall_prev) * precision;
        if (Nn > 0.0) {
          const double fpr = (double)(i + 1 - run) / Nn, fpr_prev = (double)(prev_cnt - prev) / Nn;
          roc += (fpr - fpr_prev) * (recall + recall_prev) * 0.5;
        }
      }
      prev = run;
      prev_cnt = i + 1;
    }
  }
  s_sum[t] = sum;
  __syncthreads();
  for (int w = kFinThreads / 2; w > 0; w >>= 1) {  // fixed tree: same result on every run
    if (t < w) s_sum[t] += s_sum[t + w];
    __syncthreads();
  }
  const double ap_total = s_sum[0];
  __syncthreads();
  s_sum[t] = roc;
  __syncthreads();
  for (int w = kFinThreads / 2; w > 0; w >>= 1) {
    if (t < w) s_sum[t] += s_sum[t + w];
    __syncthreads();
  }
  if (t == 0) {
    out[0] = n_pos > 0 ? fmax(0.0, ap_total) : 0.0;  // sklearn: no positive class -> 0.0 (recall defined as 1)
    out[1] = (double)n_sel;
    out[2] = (double)n_pos;
    out[3] = (double)s_groups[0];
    // ROC-AUC is undefined with a single class (sklearn raises): NaN
    out[4] = (n_pos > 0 && n_sel > n_pos) ? s_sum[0] : __longlong_as_double(0x7ff8000000000000LL);
    out[5] = out[6] = out[7] = 0.0;
  }
}

struct ApWorkspace {
  int *keysA, *keysB, *valsA, *valsB, *table, *tile_sums, *counters;
  size_t bytes;
};

ApWorkspace ap_carve(char* base, int64_t n) {
  ApWorkspace w;
  size_t off = 0;
  auto take = [&](int64_t n_int) {
    int* p = reinterpret_cast<int*>(base + off);
    off += ((size_t)n_int * sizeof(int) + 255) & ~size_t(255);
    return p;
  };
  const int64_t nblk = ceil_div(n > 0 ? n : 1, kSortTile), table_n = 256 * nblk;
  w.keysA = take(n + 1);
  w.keysB = take(n + 1);
  w.valsA = take(n + 1);
  w.valsB = take(n + 1);
  w.table = take(table_n);
  w.tile_sums = take(ceil_div(table_n, kScanTile) + 1);
  w.counters = take(4);
  w.bytes = off;
  return w;
}

__global__ void early_stop_kernel(const double* __restrict__ ap, double* __restrict__ state) {
  // state: [0] best value, [1] epochs since the best, [2] epoch of the best (1-based), [3] epochs seen,
  //        [4] 1.0 when this update improved the best (the snapshot kernel reads it)
  const double v = ap[0];
  const double epoch = state[3] + 1.0;
  state[3] = epoch;
  if (v > state[0]) {  // `if pr_val > best_val` (src/train_gnn.py:394)
    state[0] = v;
    state[1] = 0.0;
    state[2] = epoch;
    state[4] = 1.0;
  } else {
    state[1] += 1.0;
    state[4] = 0.0;
  }
}

__global__ void __launch_bounds__(kThreads) snapshot_kernel(const float4* __restrict__ src, float4* __restrict__ dst,
                                                            int64_t n4, const float* __restrict__ src_tail,
                                                            float* __restrict__ dst_tail, int tail,
                                                            const double* __restrict__ state) {
  if (state[4] == 0.0) return;
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n4; i += stride) dst[i] = src[i];
  if (blockIdx.x == 0 && threadIdx.x < tail) dst_tail[threadIdx.x] = src_tail[threadIdx.x];
}

__global__ void __launch_bounds__(kThreads) snapshot_words_kernel(const uint32_t* __restrict__ src,
                                                                  uint32_t* __restrict__ dst, int64_t n_words,
                                                                  const double* __restrict__ state) {
  if (state[4] == 0.0) return;
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n_words; i += stride) dst[i] = src[i];
}

}  // namespace
}  // namespace egnn

using namespace egnn;

extern "C" size_t egnn_ap_workspace_bytes(int64_t n_rows) {
  return ap_carve(nullptr, n_rows > 0 ? n_rows : 1).bytes;
}

extern "C" int egnn_average_precision(const float* logits, int64_t ld_logits, const float* scores, const int64_t* y,
                                      const uint8_t* mask, int64_t n_rows, float* scores_out, double* out,
                                      void* workspace, size_t workspace_bytes, void* stream) {
  const char* fn = "egnn_average_precision";
  EGNN_REQUIRE((logits != nullptr) != (scores != nullptr), fn, "exactly one of logits / scores");
  EGNN_REQUIRE(y && out && workspace, fn, "null pointer");
  EGNN_REQUIRE(n_rows >= 0 && n_rows < ((int64_t)1 << 31) - kSortTile, fn, "bad row count");
  EGNN_REQUIRE(!logits || ld_logits >= 2, fn, "logits need two columns");
  EGNN_REQUIRE(workspace_bytes >= egnn_ap_workspace_bytes(n_rows), fn, "workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  ApWorkspace w = ap_carve(reinterpret_cast<char*>(workspace), n_rows > 0 ? n_rows : 1);
  ap_clear<<<1, 32, 0, st>>>(w.counters);
  EGNN_LAUNCH_CHECK("ap_clear");
  if (n_rows > 0) {
    ap_keys<<<(unsigned)ceil_div(n_rows, kThreads), kThreads, 0, st>>>(logits, ld_logits, scores, mask, y, n_rows,
                                                                      w.keysA, scores_out, w.counters);
    EGNN_LAUNCH_CHECK("ap_keys");
    const int nblk = (int)ceil_div(n_rows, kSortTile);
    const int* kin = w.keysA;
    const int* vin = nullptr;
    int *kout = w.keysB, *vout = w.valsB;
    for (int shift = 0; shift < 32; shift += 8) {
      radix_hist<<<nblk, kThreads, 0, st>>>(kin, w.counters + 2, shift, w.table, nblk);
      EGNN_LAUNCH_CHECK("radix_hist");
      int rc = exclusive_scan(w.table, w.table, (int64_t)256 * nblk, w.tile_sums, nullptr, st);
      if (rc) return rc;
      radix_scatter<<<nblk, kThreads, 0, st>>>(kin, vin, kout, vout, w.counters + 2, shift, w.table, nblk);
      EGNN_LAUNCH_CHECK("radix_scatter");
      kin = kout;
      vin = vout;
      kout = (kout == w.keysA) ? w.keysB : w.keysA;
      vout = (vout == w.valsA) ? w.valsB : w.valsA;
    }
    ap_finish<<<1, kFinThreads, 0, st>>>(kin, vin, y, w.counters, out);
  } else {
    ap_finish<<<1, kFinThreads, 0, st>>>(w.keysA, w.valsA, y, w.counters, out);
  }
  EGNN_LAUNCH_CHECK("ap_finish");
  return 0;
}

extern "C" int egnn_early_stop_update(const double* ap, double* state, const float* params, float* best_params,
                                      int64_t n_params, void* stream) {
  const char* fn = "egnn_early_stop_update";
  EGNN_REQUIRE(ap && state, fn, "null pointer");
  EGNN_REQUIRE((params == nullptr) == (best_params == nullptr) && n_params >= 0, fn, "params / best_params mismatch");
  cudaStream_t st = (cudaStream_t)stream;
  early_stop_kernel<<<1, 1, 0, st>>>(ap, state);
  EGNN_LAUNCH_CHECK("early_stop_kernel");
  if (params && n_params > 0) {
    EGNN_REQUIRE((uintptr_t)params % 16 == 0 && (uintptr_t)best_params % 16 == 0, fn, "buffers must be 16-byte aligned");
    const int64_t n4 = n_params / 4;
    const int tail = (int)(n_params - 4 * n4);
    int64_t blocks = ceil_div(n4 > 0 ? n4 : 1, kThreads);
    if (blocks > 4 * kNumSMs) blocks = 4 * kNumSMs;
    snapshot_kernel<<<(unsigned)blocks, kThreads, 0, st>>>(reinterpret_cast<const float4*>(params),
                                                           reinterpret_cast<float4*>(best_params), n4, params + 4 * n4,
                                                           best_params + 4 * n4, tail, state);
    EGNN_LAUNCH_CHECK("snapshot_kernel");
  }
  return 0;
}

extern "C" int egnn_snapshot_if_improved(const double* state, const void* src, void* dst, int64_t n_bytes, void* stream) {
  const char* fn = "egnn_snapshot_if_improved";
  EGNN_REQUIRE(state && src && dst && n_bytes >= 0, fn, "bad arguments");
  EGNN_REQUIRE(n_bytes % 4 == 0 && (uintptr_t)src % 4 == 0 && (uintptr_t)dst % 4 == 0, fn, "4-byte aligned buffers only");
  if (n_bytes == 0) return 0;
  const int64_t n_words = n_bytes / 4;
  int64_t blocks = ceil_div(n_words, kThreads);
  if (blocks > 4 * kNumSMs) blocks = 4 * kNumSMs;
  snapshot_words_kernel<<<(unsigned)blocks, kThreads, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint32_t*>(src), reinterpret_cast<uint32_t*>(dst), n_words, state);
  EGNN_LAUNCH_CHECK(fn);
  return 0;
}
