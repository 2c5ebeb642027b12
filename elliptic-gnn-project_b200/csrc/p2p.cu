// One-shot all-reduce of small vectors over NVLink peer memory: the BatchNorm statistics of the timestep-sharded
// step (2 x hidden doubles per layer and direction) and the flat weight-gradient
// buffer (41 090 floats for rec_k8).
//
// A NCCL all-reduce of 1 KB costs ~30 us of launch + protocol latency, and SAGE-ResBN needs four of them on
// the critical path of every step.  Here each rank PUSHES its vector straight into a slot of every peer's
// symmetric buffer with plain stores over NVLink/NVSwitch (the buffers are mapped into each process by
// torch.distributed._symmetric_memory; this file only sees an array of peer base pointers), publishes a flag
// per peer with a system-scope release store, spins on the flags in its OWN memory and sums the world_size
// slots in rank order -- every rank computes bit-identical sums.  Slots and flags are double-buffered on the
// parity of a device-side epoch counter, so the kernel is CUDA-graph replayable and a peer can never overwrite
// a slot that is still being read (it cannot enter epoch e+2 before this rank has published epoch e+1).
#include "common.cuh"

namespace egnn {
namespace {

constexpr int kP2PThreads = 256;
constexpr int kP2PChunk = 2048;     // elements per CTA
constexpr int kP2PMaxChunks = 64;   // -> vectors of up to 131 072 elements
constexpr int kP2PMaxWorld = 16;

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// buffer layout of every rank: T data[2][world][n_max]; u64 flags[2][kP2PMaxChunks][kP2PMaxWorld]
__host__ __device__ inline size_t flag_offset(int world, int64_t n_max, size_t es) {
  return (es * 2 * (size_t)world * (size_t)n_max + 15) & ~size_t(15);
}

// CTA c owns elements [c*kP2PChunk, ...) and its own flag row, so chunks complete independently
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

template <typename T>
__device__ __forceinline__ T poison();
template <>
__device__ __forceinline__ float poison<float>() { return __int_as_float(0x7fc00000); }
template <>
__device__ __forceinline__ double poison<double>() { return __longlong_as_double(0x7ff8000000000000LL); }

// `in` may alias `out` (in-place call): no __restrict__ on either
template <typename T>
__global__ void __launch_bounds__(kP2PThreads) p2p_allreduce_kernel(const T* in, T* out, int64_t n,
                                                                    int64_t n_max, void* const* __restrict__ peer_bufs,
                                                                    int rank, int world,
                                                                    unsigned long long* __restrict__ epoch,
                                                                    int* __restrict__ error_flag,
                                                                    unsigned long long timeout_ns) {
  __shared__ int s_fail;
  if (threadIdx.x == 0) s_fail = 0;
  // epoch[0] = calls completed, epoch[1] = ticket counter of the running call: the LAST block to finish advances the
  // epoch (no block can still be reading it then) and clears the ticket -- no separate launch
  const unsigned long long e = *epoch + 1ull;
  const int par = (int)(e & 1ull);
  const int64_t i0 = (int64_t)blockIdx.x * kP2PChunk;
  const int cnt = (int)min((int64_t)kP2PChunk, n - i0);
  __syncthreads();
  // 1. push my chunk into slot [par][rank] of every rank's buffer (remote stores over NVLink)
  for (int p = 0; p < world; ++p) {
    T* dst = reinterpret_cast<T*>(peer_bufs[p]) + ((size_t)par * world + rank) * n_max + i0;
    for (int i = threadIdx.x; i < cnt; i += kP2PThreads) dst[i] = in[i0 + i];
  }
  __threadfence_system();
  __syncthreads();
  // 2. publish: flag [par][chunk][rank] = e on every rank; 3. wait for every rank's flag in my own buffer
  const size_t foff = flag_offset(world, n_max, sizeof(T));
  if (threadIdx.x < world) {
    unsigned long long* pf =
        reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(peer_bufs[threadIdx.x]) + foff);
    st_release_sys(pf + ((size_t)par * kP2PMaxChunks + blockIdx.x) * kP2PMaxWorld + rank, e);
    const unsigned long long* mine =
        reinterpret_cast<const unsigned long long*>(reinterpret_cast<const char*>(peer_bufs[rank]) + foff);
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(mine + ((size_t)par * kP2PMaxChunks + blockIdx.x) * kP2PMaxWorld + threadIdx.x) < e) {
      if (global_ns() - t0 > timeout_ns) {  // a peer never arrived; fail loudly instead of hanging the GPU
        s_fail = 1;
        break;
      }
    }
  }
  __syncthreads();
  if (s_fail) {
    // sticky and loud: the error flag stays set (the host checks it at every synchronisation point) and the result
    // is NaN, so a loss / parameter computed from an un-reduced vector can never look valid
    if (threadIdx.x == 0 && error_flag) *error_flag = 1;
    for (int i = threadIdx.x; i < cnt; i += kP2PThreads) out[i0 + i] = poison<T>();
  } else {
    // 4. sum the slots in rank order (identical on every rank)
    const T* data = reinterpret_cast<const T*>(peer_bufs[rank]) + (size_t)par * world * n_max + i0;
    for (int i = threadIdx.x; i < cnt; i += kP2PThreads) {
      T s = 0;
      for (int r = 0; r < world; ++r) s += data[(size_t)r * n_max + i];
      out[i0 + i] = s;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned long long done = atomicAdd(epoch + 1, 1ull);
    if (done == gridDim.x - 1) {
      epoch[1] = 0ull;
      __threadfence();
      epoch[0] = e;
    }
  }
}

// ---- single-CTA exchange fused with its producer and consumer --------------------------------------------------------
// The BatchNorm statistics of the timestep-sharded step are 2 x hidden doubles per layer and direction; as separate
// launches (reduce partials -> all-reduce -> finalise) each exchange costs three kernels plus their launch gaps on the
// critical path.  Here ONE block reduces the partial rows, pushes the 2F sums to every peer, waits, adds the world's
// slots in rank order and finalises -- the same buffer, flag and epoch protocol as p2p_allreduce_kernel (chunk 0).
constexpr int kXThreads = 1024;

// vals[0..n) (shared memory, n <= n_max): on return the sum over ranks (rank order); false on timeout
__device__ bool cta_exchange_f64(double* vals, int n, int64_t n_max, void* const* __restrict__ peer_bufs, int rank,
                                 int world, unsigned long long e, unsigned long long timeout_ns, int* s_fail) {
  const int par = (int)(e & 1ull);
  for (int p = 0; p < world; ++p) {
    double* dst = reinterpret_cast<double*>(peer_bufs[p]) + ((size_t)par * world + rank) * n_max;
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = vals[i];
  }
  __threadfence_system();
  __syncthreads();
  const size_t foff = flag_offset(world, n_max, sizeof(double));
  if ((int)threadIdx.x < world) {
    unsigned long long* pf =
        reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(peer_bufs[threadIdx.x]) + foff);
    st_release_sys(pf + ((size_t)par * kP2PMaxChunks) * kP2PMaxWorld + rank, e);
    const unsigned long long* mine =
        reinterpret_cast<const unsigned long long*>(reinterpret_cast<const char*>(peer_bufs[rank]) + foff);
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(mine + ((size_t)par * kP2PMaxChunks) * kP2PMaxWorld + threadIdx.x) < e) {
      if (global_ns() - t0 > timeout_ns) {
        *s_fail = 1;
        break;
      }
    }
  }
  __syncthreads();
  if (*s_fail) return false;
  const double* data = reinterpret_cast<const double*>(peer_bufs[rank]) + (size_t)par * world * n_max;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double s = 0.0;
    for (int r = 0; r < world; ++r) s += data[(size_t)r * n_max + i];
    vals[i] = s;
  }
  __syncthreads();
  return true;
}

// local fixed-order reduction of `n_parts` rows of 2F values into vals[0..2F): item j = which*F + c read at
// base[(p*2 + which)*F + c]; G = blockDim / 2F groups take parts g, g+G, ... and are combined in group order
template <typename TP>
__device__ void cta_reduce_parts(const TP* __restrict__ parts, int n_parts, int F, double* vals, double* red) {
  const int n = 2 * F, G = blockDim.x / n;
  const int j = threadIdx.x % n, g = threadIdx.x / n;
  if (g < G) {
    // four loads in flight per thread, added in the same fixed order (one dependent L2 round trip per part made this
    // reduction the longest phase of an exchange: ~19 parts per thread at hidden 64)
    double s = 0.0;
    int p = g;
    for (; p + 3 * G < n_parts; p += 4 * G) {
      const TP v0 = parts[(size_t)p * n + j], v1 = parts[(size_t)(p + G) * n + j];
      const TP v2 = parts[(size_t)(p + 2 * G) * n + j], v3 = parts[(size_t)(p + 3 * G) * n + j];
      s += (double)v0; s += (double)v1; s += (double)v2; s += (double)v3;
    }
    for (; p < n_parts; p += G) s += (double)parts[(size_t)p * n + j];
    red[g * n + j] = s;
  }
  __syncthreads();
  if ((int)threadIdx.x < n) {
    double s = 0.0;
    for (int k = 0; k < G; ++k) s += red[k * n + threadIdx.x];
    vals[threadIdx.x] = s;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kXThreads) bn_fwd_exchange_kernel(
    const float* __restrict__ parts, int n_parts, int F, double count, float eps, float momentum, float* __restrict__ mean,
    float* __restrict__ rstd, float* __restrict__ rmean, float* __restrict__ rvar, int64_t* __restrict__ num_batches,
    int64_t n_max, void* const* __restrict__ peer_bufs, int rank, int world, unsigned long long* __restrict__ epoch,
    int* __restrict__ error_flag, unsigned long long timeout_ns) {
  __shared__ double vals[kXThreads], red[kXThreads];
  __shared__ int s_fail;
  if (threadIdx.x == 0) s_fail = 0;
  const unsigned long long e = *epoch + 1ull;
  __syncthreads();
  cta_reduce_parts<float>(parts, n_parts, F, vals, red);
  const bool ok = cta_exchange_f64(vals, 2 * F, n_max, peer_bufs, rank, world, e, timeout_ns, &s_fail);
  const int c = threadIdx.x;
  if (c < F) {
    if (!ok) {   // loud: NaN statistics poison everything downstream, the sticky flag tells the host why
      mean[c] = poison<float>();
      rstd[c] = poison<float>();
    } else {
      const double m = vals[c] / count;
      double var = vals[F + c] / count - m * m;
      if (var < 0) var = 0;
      mean[c] = (float)m;
      rstd[c] = (float)(1.0 / sqrt(var + (double)eps));
      if (rmean) rmean[c] = (1.f - momentum) * rmean[c] + momentum * (float)m;
      if (rvar) {
        const double unb = count > 1 ? var * count / (count - 1) : var;
        rvar[c] = (1.f - momentum) * rvar[c] + momentum * (float)unb;
      }
    }
  }
  if (threadIdx.x == 0) {
    if (!ok && error_flag) *error_flag = 1;
    if (num_batches) *num_batches += 1;
    epoch[0] = e;
  }
}

__global__ void __launch_bounds__(kXThreads) bn_bwd_exchange_kernel(
    const double* __restrict__ partial, int n_parts, int F, double* __restrict__ sums, float* __restrict__ f0,
    float* __restrict__ f1, int64_t n_max, void* const* __restrict__ peer_bufs, int rank, int world,
    unsigned long long* __restrict__ epoch, int* __restrict__ error_flag, unsigned long long timeout_ns) {
  __shared__ double vals[kXThreads], red[kXThreads];
  __shared__ int s_fail;
  if (threadIdx.x == 0) s_fail = 0;
  const unsigned long long e = *epoch + 1ull;
  __syncthreads();
  cta_reduce_parts<double>(partial, n_parts, F, vals, red);
  const int j = threadIdx.x;
  if (j < 2 * F) {
    // d beta / d gamma: THIS rank's share -- the flat weight-gradient all-reduce adds the ranks' shares later, like
    // every other parameter gradient (handing out the world-wide sums here would count them `world` times)
    float* f = j < F ? f0 : f1;
    if (f) f[j < F ? j : j - F] = (float)vals[j];
  }
  __syncthreads();
  const bool ok = cta_exchange_f64(vals, 2 * F, n_max, peer_bufs, rank, world, e, timeout_ns, &s_fail);
  if (j < 2 * F) sums[j] = ok ? vals[j] : poison<double>();   // [sum g | sum g*xhat] over all ranks
  if (threadIdx.x == 0) {
    if (!ok && error_flag) *error_flag = 1;
    epoch[0] = e;
  }
}

}  // namespace
}  // namespace egnn

using namespace egnn;

extern "C" size_t egnn_p2p_allreduce_buffer_bytes(int world, int64_t n_max, int dtype) {
  const size_t es = dtype == EGNN_F64 ? 8 : 4;
  return flag_offset(world, n_max, es) + sizeof(unsigned long long) * 2 * kP2PMaxChunks * kP2PMaxWorld;
}

extern "C" int egnn_p2p_allreduce(const void* in, void* out, int64_t n, int dtype, int64_t n_max,
                                  void* const* peer_bufs_dev, int rank, int world, int64_t* epoch, int* error_flag,
                                  int64_t timeout_ms, void* stream) {
  const char* fn = "egnn_p2p_allreduce";
  EGNN_REQUIRE(in && out && peer_bufs_dev && epoch, fn, "null pointer");
  EGNN_REQUIRE(dtype == EGNN_F32 || dtype == EGNN_F64, fn, "dtype must be EGNN_F32 or EGNN_F64");
  EGNN_REQUIRE(n > 0 && n <= n_max && n_max <= (int64_t)kP2PChunk * kP2PMaxChunks, fn, "n out of range");
  EGNN_REQUIRE(world >= 1 && world <= kP2PMaxWorld && rank >= 0 && rank < world, fn, "bad rank / world");
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned long long tmo = (unsigned long long)(timeout_ms > 0 ? timeout_ms : 2000) * 1000000ull;
  const unsigned grid = (unsigned)ceil_div(n, kP2PChunk);
  unsigned long long* ep = reinterpret_cast<unsigned long long*>(epoch);
  if (dtype == EGNN_F64)
    p2p_allreduce_kernel<double><<<grid, kP2PThreads, 0, st>>>((const double*)in, (double*)out, n, n_max,
                                                               peer_bufs_dev, rank, world, ep, error_flag, tmo);
  else
    p2p_allreduce_kernel<float><<<grid, kP2PThreads, 0, st>>>((const float*)in, (float*)out, n, n_max, peer_bufs_dev,
                                                              rank, world, ep, error_flag, tmo);
  EGNN_LAUNCH_CHECK(fn);
  return 0;
}

static int check_exchange_args(const char* fn, int64_t F, int64_t n_parts, int64_t n_max, void* const* peers, int rank,
                               int world, int64_t* epoch) {
  EGNN_REQUIRE(peers && epoch, fn, "null pointer");
  EGNN_REQUIRE(F > 0 && 2 * F <= kXThreads && 2 * F <= n_max && n_parts > 0, fn, "2 * n_feat must fit 1024 and n_max");
  EGNN_REQUIRE(world >= 1 && world <= kP2PMaxWorld && rank >= 0 && rank < world, fn, "bad rank / world");
  return 0;
}

extern "C" int egnn_bn_stats_exchange(const float* parts, int64_t n_parts, int64_t n_feat, double count, float eps,
                                      float momentum, float* mean, float* rstd, float* running_mean,
                                      float* running_var, int64_t* num_batches_tracked, int64_t n_max,
                                      void* const* peer_bufs_dev, int rank, int world, int64_t* epoch, int* error_flag,
                                      int64_t timeout_ms, void* stream) {
  const char* fn = "egnn_bn_stats_exchange";
  EGNN_REQUIRE(parts && mean && rstd && count > 0, fn, "bad arguments");
  int rc = check_exchange_args(fn, n_feat, n_parts, n_max, peer_bufs_dev, rank, world, epoch);
  if (rc) return rc;
  const unsigned long long tmo = (unsigned long long)(timeout_ms > 0 ? timeout_ms : 2000) * 1000000ull;
  bn_fwd_exchange_kernel<<<1, kXThreads, 0, (cudaStream_t)stream>>>(
      parts, (int)n_parts, (int)n_feat, count, eps, momentum, mean, rstd, running_mean, running_var, num_batches_tracked,
      n_max, peer_bufs_dev, rank, world, reinterpret_cast<unsigned long long*>(epoch), error_flag, tmo);
  EGNN_LAUNCH_CHECK(fn);
  return 0;
}

extern "C" int egnn_bn_bwd_sums_exchange(const double* partial, int64_t n_parts, int64_t n_feat, double* sums,
                                         float* sum_g_f32, float* sum_gx_f32, int64_t n_max, void* const* peer_bufs_dev,
                                         int rank, int world, int64_t* epoch, int* error_flag, int64_t timeout_ms,
                                         void* stream) {
  const char* fn = "egnn_bn_bwd_sums_exchange";
  EGNN_REQUIRE(partial && sums, fn, "bad arguments");
  int rc = check_exchange_args(fn, n_feat, n_parts, n_max, peer_bufs_dev, rank, world, epoch);
  if (rc) return rc;
  const unsigned long long tmo = (unsigned long long)(timeout_ms > 0 ? timeout_ms : 2000) * 1000000ull;
  bn_bwd_exchange_kernel<<<1, kXThreads, 0, (cudaStream_t)stream>>>(
      partial, (int)n_parts, (int)n_feat, sums, sum_g_f32, sum_gx_f32, n_max, peer_bufs_dev, rank, world,
      reinterpret_cast<unsigned long long*>(epoch), error_flag, tmo);
  EGNN_LAUNCH_CHECK(fn);
  return 0;
}
