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
                                                                    const unsigned long long* __restrict__ epoch,
                                                                    int* __restrict__ error_flag,
                                                                    unsigned long long timeout_ns) {
  __shared__ int s_fail;
  if (threadIdx.x == 0) s_fail = 0;
  const unsigned long long e = *epoch + 1ull;   // advanced by p2p_epoch_advance after the whole grid
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
    return;
  }
  // 4. sum the slots in rank order (identical on every rank)
  const T* data = reinterpret_cast<const T*>(peer_bufs[rank]) + (size_t)par * world * n_max + i0;
  for (int i = threadIdx.x; i < cnt; i += kP2PThreads) {
    T s = 0;
    for (int r = 0; r < world; ++r) s += data[(size_t)r * n_max + i];
    out[i0 + i] = s;
  }
}

__global__ void p2p_epoch_advance(unsigned long long* epoch) { *epoch += 1ull; }

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
  p2p_epoch_advance<<<1, 1, 0, st>>>(ep);
  EGNN_LAUNCH_CHECK(fn);
  return 0;
}
