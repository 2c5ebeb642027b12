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

__device__ __forceinline__ float key_score(int key) {   // inverse of desc_key
  const unsigned u = ~(unsigned)key;
  const unsigned b = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
  return __uint_as_float(b);
}

struct F1Best {
  double f1;
  int idx;   // position of the threshold in the descending run (-1: none)
};
__device__ __forceinline__ F1Best better(F1Best a, F1Best b) {
  // np.nanargmax over ASCENDING thresholds returns the first maximum = the LOWEST threshold among ties
  // = the largest position in the descending run
  if (b.idx < 0) return a;
  if (a.idx < 0) return b;
  if (b.f1 > a.f1 || (b.f1 == a.f1 && b.idx > a.idx)) return b;
  return a;
}

constexpr int kMaxEceBins = 32;

// Final metrics of src/train_gnn.py:449-470 from the same sorted run (src/utils/metrics.py:18-66):
//   out[8]  max F1 over the precision-recall curve, out[9] its threshold          (pick_threshold_max_f1)
//   out[10] precision among the top-k scores, out[15] = min(k, n)                 (precision_at_k)
//   out[11] max recall with precision >= target                                   (recall_at_precision)
//   out[12] lowest threshold with precision >= target (1.0 when there is none)    (pick_threshold_for_precision)
//   out[13] F1 of `score >= thr` (thr = *thr_dev, or out[9] when thr_dev is NULL) (f1_at_threshold)
//   out[14] expected calibration error over `bins` equal-width bins               (expected_calibration_error)
__global__ void __launch_bounds__(kFinThreads) rank_finish(const int* __restrict__ keys, const int* __restrict__ perm,
                                                           const int64_t* __restrict__ y,
                                                           const int* __restrict__ counters, int64_t top_k,
                                                           double target_p, const double* __restrict__ thr_dev,
                                                           int bins, double* __restrict__ out) {
  __shared__ int s_tp[kFinThreads];
  __shared__ double s_d[kFinThreads];
  __shared__ int s_i[kFinThreads];
  __shared__ int s_j[kFinThreads];
  __shared__ double s_thr;
  const int n_sel = counters[0], n_pos = counters[1];
  const int t = threadIdx.x;
  const int chunk = (n_sel + kFinThreads - 1) / kFinThreads;
  const int lo = min(t * chunk, n_sel), hi = min(lo + chunk, n_sel);
  int tp = 0;
  for (int i = lo; i < hi; ++i) tp += y[perm[i]] == 1;
  s_tp[t] = tp;
  __syncthreads();
  if (t == 0) {
    int run = 0;
    for (int q = 0; q < kFinThreads; ++q) {
      const int c = s_tp[q];
      s_tp[q] = run;
      run += c;
    }
  }
  __syncthreads();
  const double P = (double)n_pos;
  // ---- pass A: thresholds (tie-group ends): max F1, recall@precision, lowest threshold reaching the precision
  F1Best best{0.0, -1};
  double best_rec = 0.0;
  int thr_p_idx = -1;
  int run = s_tp[t];
  for (int i = lo; i < hi; ++i) {
    run += y[perm[i]] == 1;
    if (i == n_sel - 1 || keys[i] != keys[i + 1]) {
      const double precision = (double)run / (double)(i + 1);
      const double recall = n_pos > 0 ? (double)run / P : 1.0;   // sklearn: no positives -> recall = 1 everywhere
      const double f1 = 2.0 * precision * recall / (precision + recall + 1e-12);
      best = better(best, F1Best{f1, i});
      if (precision >= target_p) {
        best_rec = fmax(best_rec, recall);
        thr_p_idx = max(thr_p_idx, i);
      }
    }
  }
  s_d[t] = best.f1;
  s_i[t] = best.idx;
  __syncthreads();
  for (int w = kFinThreads / 2; w > 0; w >>= 1) {
    if (t < w) {
      const F1Best r = better(F1Best{s_d[t], s_i[t]}, F1Best{s_d[t + w], s_i[t + w]});
      s_d[t] = r.f1;
      s_i[t] = r.idx;
    }
    __syncthreads();
  }
  const double max_f1 = s_d[0];
  const int max_f1_idx = s_i[0];
  __syncthreads();
  s_d[t] = best_rec;
  s_i[t] = thr_p_idx;
  __syncthreads();
  for (int w = kFinThreads / 2; w > 0; w >>= 1) {
    if (t < w) {
      s_d[t] = fmax(s_d[t], s_d[t + w]);
      s_i[t] = max(s_i[t], s_i[t + w]);
    }
    __syncthreads();
  }
  const double rec_at_p = s_d[0];
  const int thr_idx = s_i[0];
  __syncthreads();
  // the appended (precision 1, recall 0, threshold 1.0) point of precision_recall_curve has F1 = 0: it wins only when
  // every real threshold has F1 == 0 exactly (then nanargmax returns index 0 = the LOWEST real threshold anyway)
  const double thr_f1 = max_f1_idx >= 0 ? (double)key_score(keys[max_f1_idx]) : 1.0;
  if (t == 0) s_thr = thr_dev ? *thr_dev : thr_f1;
  __syncthreads();
  const double thr = s_thr;
  // ---- pass B: prefix counts (top-k, score >= thr) and the calibration bins
  const int k_used = (int)min((int64_t)n_sel, top_k > 0 ? top_k : (int64_t)0);
  int tp_k = 0, cnt_thr = 0, tp_thr = 0;
  int b_cnt[kMaxEceBins], b_pos[kMaxEceBins];
  double b_sum[kMaxEceBins];
#pragma unroll
  for (int b = 0; b < kMaxEceBins; ++b) {
    b_cnt[b] = 0;
    b_pos[b] = 0;
    b_sum[b] = 0.0;
  }
  const double step = 1.0 / (double)bins;   // np.linspace(0, 1, bins + 1): i * step, last edge exactly 1.0
  for (int i = lo; i < hi; ++i) {
    const int pos = y[perm[i]] == 1;
    const double sc = (double)key_score(keys[i]);
    if (i < k_used) tp_k += pos;
    if (sc >= thr) {
      ++cnt_thr;
      tp_thr += pos;
    }
#pragma unroll
    for (int b = 0; b < kMaxEceBins; ++b) {
      if (b < bins) {
        const double e_lo = (double)b * step, e_hi = (b == bins - 1) ? 1.0 : (double)(b + 1) * step;
        const bool in = (sc >= e_lo) && (b < bins - 1 ? sc < e_hi : sc <= e_hi);
        if (in) {
          ++b_cnt[b];
          b_pos[b] += pos;
          b_sum[b] += sc;
        }
      }
    }
  }
  auto block_sum_int = [&](int v) {
    s_i[t] = v;
    __syncthreads();
    for (int w = kFinThreads / 2; w > 0; w >>= 1) {
      if (t < w) s_i[t] += s_i[t + w];
      __syncthreads();
    }
    const int r = s_i[0];
    __syncthreads();
    return r;
  };
  auto block_sum_dbl = [&](double v) {
    s_d[t] = v;
    __syncthreads();
    for (int w = kFinThreads / 2; w > 0; w >>= 1) {
      if (t < w) s_d[t] += s_d[t + w];
      __syncthreads();
    }
    const double r = s_d[0];
    __syncthreads();
    return r;
  };
  const int tot_tp_k = block_sum_int(tp_k);
  const int tot_cnt_thr = block_sum_int(cnt_thr);
  const int tot_tp_thr = block_sum_int(tp_thr);
  double ece = 0.0;
  for (int b = 0; b < bins; ++b) {
    const int c = block_sum_int(b_cnt[b]);
    const int ps = block_sum_int(b_pos[b]);
    const double sm = block_sum_dbl(b_sum[b]);
    if (c > 0) {
      const double conf = sm / (double)c, acc = (double)ps / (double)c;
      ece += ((double)c / (double)n_sel) * fabs(acc - conf);
    }
  }
  (void)s_j;
  if (t == 0) {
    out[8] = max_f1_idx >= 0 ? max_f1 : 0.0;
    out[9] = thr_f1;
    out[10] = k_used > 0 ? (double)tot_tp_k / (double)k_used : __longlong_as_double(0x7ff8000000000000LL);
    out[11] = rec_at_p;
    out[12] = thr_idx >= 0 ? (double)key_score(keys[thr_idx]) : 1.0;
    // sklearn f1_score: 2 tp / (2 tp + fp + fn) = 2 tp / (predicted positives + actual positives); 0 when undefined
    const double den = (double)tot_cnt_thr + P;
    out[13] = den > 0.0 ? 2.0 * (double)tot_tp_thr / den : 0.0;
    out[14] = ece;
    out[15] = (double)k_used;
  }
}

// Temperature scaling (src/utils/calibrate.py:8-30: minimise CrossEntropyLoss(logits / T, y) over one scalar T with
// LBFGS from T = 1).  Two classes: loss_i = softplus(beta * d_i), beta = 1 / T, d_i = l_other - l_true, which is convex
// in beta; one CTA runs a damped Newton iteration on beta with float64 fixed-order sums over the selected rows.
// out: [T, mean NLL at T = 1, mean NLL at T, Newton iterations, selected rows]
__global__ void __launch_bounds__(kFinThreads) temperature_fit_kernel(const float* __restrict__ logits, int64_t ld,
                                                                      const int64_t* __restrict__ y,
                                                                      const uint8_t* __restrict__ mask, int64_t n,
                                                                      int max_iter, double* __restrict__ out) {
  __shared__ double s_a[kFinThreads], s_b[kFinThreads], s_c[kFinThreads];
  __shared__ double s_beta, s_cnt;
  __shared__ int s_done;
  const int t = threadIdx.x;
  auto reduce3 = [&](double a, double b, double c) {
    s_a[t] = a;
    s_b[t] = b;
    s_c[t] = c;
    __syncthreads();
    for (int w = kFinThreads / 2; w > 0; w >>= 1) {
      if (t < w) {
        s_a[t] += s_a[t + w];
        s_b[t] += s_b[t + w];
        s_c[t] += s_c[t + w];
      }
      __syncthreads();
    }
  };
  auto eval = [&](double beta, double& f, double& g, double& h) {
    double a = 0.0, b = 0.0, c = 0.0;
    for (int64_t i = t; i < n; i += kFinThreads) {
      if (mask && !mask[i]) continue;
      const int64_t yi = y[i];
      if (yi != 0 && yi != 1) continue;
      const double l0 = (double)logits[i * ld], l1 = (double)logits[i * ld + 1];
      const double d = yi == 1 ? l0 - l1 : l1 - l0;
      const double z = beta * d;
      a += z > 0.0 ? z + log1p(exp(-z)) : log1p(exp(z));   // softplus
      const double sg = 1.0 / (1.0 + exp(-z));
      b += d * sg;
      c += d * d * sg * (1.0 - sg);
    }
    reduce3(a, b, c);
    f = s_a[0];
    g = s_b[0];
    h = s_c[0];
    __syncthreads();
  };
  {   // number of selected rows
    double c = 0.0;
    for (int64_t i = t; i < n; i += kFinThreads)
      if ((!mask || mask[i]) && (y[i] == 0 || y[i] == 1)) c += 1.0;
    reduce3(c, 0.0, 0.0);
    if (t == 0) s_cnt = s_a[0];
    __syncthreads();
  }
  const double cnt = s_cnt;
  if (cnt == 0.0) {
    if (t == 0) {
      out[0] = 1.0;
      out[1] = out[2] = out[3] = out[4] = 0.0;
    }
    return;
  }
  double beta = 1.0, f, g, h;
  eval(beta, f, g, h);
  const double f_start = f;
  int it = 0;
  for (; it < max_iter; ++it) {
    if (fabs(g) <= 1e-13 * cnt || !(h > 0.0)) break;
    double step = g / h, nb, nf, ng, nh;
    int tries = 0;
    for (;;) {   // damped Newton: halve the step until the objective does not increase and beta stays positive
      nb = beta - step;
      if (nb > 0.0) {
        eval(nb, nf, ng, nh);
        if (nf <= f + 1e-15 * fabs(f)) break;
      }
      step *= 0.5;
      if (++tries > 40) {
        nb = beta;
        nf = f;
        ng = g;
        nh = h;
        break;
      }
    }
    const bool converged = fabs(nb - beta) <= 1e-14 * fabs(beta);
    beta = nb;
    f = nf;
    g = ng;
    h = nh;
    if (converged) {
      ++it;
      break;
    }
  }
  (void)s_beta;
  (void)s_done;
  if (t == 0) {
    out[0] = 1.0 / beta;
    out[1] = f_start / cnt;
    out[2] = f / cnt;
    out[3] = (double)it;
    out[4] = cnt;
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

__global__ void early_stop_kernel(const double* __restrict__ ap, double* __restrict__ state, double patience) {
  // state: [0] best value, [1] epochs since the best, [2] epoch of the best (1-based), [3] epochs seen,
  //        [4] 1.0 when this update improved the best (the snapshot kernel reads it)
  // `if bad >= patience: break` (src/train_gnn.py:411): once the reference would have left its loop the state is
  // frozen -- epochs the host runs before it polls (fit(poll_every=...)) can neither move the best value nor
  // overwrite the best-parameter snapshot.
  if (patience > 0.0 && state[1] >= patience) {
    state[4] = 0.0;
    return;
  }
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

// keys -> stable descending sort; on return *keys_sorted / *perm_sorted point at the sorted run inside the workspace
static int ap_sort(const float* logits, int64_t ld_logits, const float* scores, const int64_t* y, const uint8_t* mask,
                   int64_t n_rows, float* scores_out, ApWorkspace& w, const int** keys_sorted, const int** perm_sorted,
                   cudaStream_t st) {
  ap_clear<<<1, 32, 0, st>>>(w.counters);
  EGNN_LAUNCH_CHECK("ap_clear");
  *keys_sorted = w.keysA;
  *perm_sorted = w.valsA;
  if (n_rows == 0) return 0;
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
  *keys_sorted = kin;
  *perm_sorted = vin;
  return 0;
}

static int ap_check_args(const char* fn, const float* logits, int64_t ld_logits, const float* scores, const int64_t* y,
                         double* out, void* workspace, size_t workspace_bytes, int64_t n_rows) {
  EGNN_REQUIRE((logits != nullptr) != (scores != nullptr), fn, "exactly one of logits / scores");
  EGNN_REQUIRE(y && out && workspace, fn, "null pointer");
  EGNN_REQUIRE(n_rows >= 0 && n_rows < ((int64_t)1 << 31) - kSortTile, fn, "bad row count");
  EGNN_REQUIRE(!logits || ld_logits >= 2, fn, "logits need two columns");
  EGNN_REQUIRE(workspace_bytes >= egnn_ap_workspace_bytes(n_rows), fn, "workspace too small");
  return 0;
}

extern "C" int egnn_average_precision(const float* logits, int64_t ld_logits, const float* scores, const int64_t* y,
                                      const uint8_t* mask, int64_t n_rows, float* scores_out, double* out,
                                      void* workspace, size_t workspace_bytes, void* stream) {
  const char* fn = "egnn_average_precision";
  int rc = ap_check_args(fn, logits, ld_logits, scores, y, out, workspace, workspace_bytes, n_rows);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  ApWorkspace w = ap_carve(reinterpret_cast<char*>(workspace), n_rows > 0 ? n_rows : 1);
  const int *ks, *ps;
  rc = ap_sort(logits, ld_logits, scores, y, mask, n_rows, scores_out, w, &ks, &ps, st);
  if (rc) return rc;
  ap_finish<<<1, kFinThreads, 0, st>>>(ks, ps, y, w.counters, out);
  EGNN_LAUNCH_CHECK("ap_finish");
  return 0;
}

extern "C" int egnn_ranking_metrics(const float* logits, int64_t ld_logits, const float* scores, const int64_t* y,
                                    const uint8_t* mask, int64_t n_rows, int64_t top_k, double target_precision,
                                    const double* threshold_dev, int ece_bins, float* scores_out, double* out,
                                    void* workspace, size_t workspace_bytes, void* stream) {
  const char* fn = "egnn_ranking_metrics";
  int rc = ap_check_args(fn, logits, ld_logits, scores, y, out, workspace, workspace_bytes, n_rows);
  if (rc) return rc;
  EGNN_REQUIRE(ece_bins >= 1 && ece_bins <= kMaxEceBins, fn, "ece_bins must be in [1, 32]");
  EGNN_REQUIRE(top_k >= 0, fn, "top_k must be >= 0");
  cudaStream_t st = (cudaStream_t)stream;
  ApWorkspace w = ap_carve(reinterpret_cast<char*>(workspace), n_rows > 0 ? n_rows : 1);
  const int *ks, *ps;
  rc = ap_sort(logits, ld_logits, scores, y, mask, n_rows, scores_out, w, &ks, &ps, st);
  if (rc) return rc;
  ap_finish<<<1, kFinThreads, 0, st>>>(ks, ps, y, w.counters, out);
  EGNN_LAUNCH_CHECK("ap_finish");
  rank_finish<<<1, kFinThreads, 0, st>>>(ks, ps, y, w.counters, top_k, target_precision, threshold_dev, ece_bins, out);
  EGNN_LAUNCH_CHECK("rank_finish");
  return 0;
}

extern "C" int egnn_temperature_fit(const float* logits, int64_t ld_logits, const int64_t* y, const uint8_t* mask,
                                    int64_t n_rows, int max_iter, double* out, void* stream) {
  const char* fn = "egnn_temperature_fit";
  EGNN_REQUIRE(logits && y && out, fn, "null pointer");
  EGNN_REQUIRE(ld_logits >= 2 && n_rows >= 0 && max_iter > 0, fn, "bad arguments");
  temperature_fit_kernel<<<1, kFinThreads, 0, (cudaStream_t)stream>>>(logits, ld_logits, y, mask, n_rows, max_iter, out);
  EGNN_LAUNCH_CHECK(fn);
  return 0;
}

extern "C" int egnn_early_stop_update(const double* ap, double* state, const float* params, float* best_params,
                                      int64_t n_params, int64_t patience, void* stream) {
  const char* fn = "egnn_early_stop_update";
  EGNN_REQUIRE(ap && state, fn, "null pointer");
  EGNN_REQUIRE((params == nullptr) == (best_params == nullptr) && n_params >= 0, fn, "params / best_params mismatch");
  cudaStream_t st = (cudaStream_t)stream;
  early_stop_kernel<<<1, 1, 0, st>>>(ap, state, (double)patience);
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
