// Data-parallel exchange step for SMALL buffers: a one-shot all-reduce over NVLink peer memory.
// At the published configuration a training step is ~0.2 ms and its two exchanges are tiny -- the (L, N) column sums
// before the loss (4 KB) and the flat gradient buffer after the backward (~200 KB) -- so the cost of a collective is
// its latency, not its bandwidth.  Every rank owns a staging buffer and a signal pad in symmetric (peer-mapped)
// memory.  One launch per rank, no host involvement, CUDA-graph capturable:
//   1. block b copies slice b of the local input into this rank's staging buffer (parity = epoch & 1);
//   2. it stores `epoch` into slot (b, rank) of every peer's signal pad (st.release.sys over NVLink) and spins until
//      its own pad shows >= epoch from every peer (ld.acquire.sys);
//   3. it reads slice b of EVERY rank's staging buffer over NVLink, adds them in rank order (every rank gets bit-identical
//      sums, so replicated parameters stay identical), scales, and writes the local output.
// The staging buffer is double-buffered on the epoch parity: a rank can be at most one call ahead of its slowest
// peer (it needs that peer's signal to finish a call), so the buffer of call e is never overwritten before every
// rank has finished reading it.  The epoch lives on the device and is advanced by the last block to finish.
// Large buffers (the 268 MB output-layer gradient at T = 2^19) stay with NCCL: one-shot moves (N-1) x the buffer.
#include <algorithm>

#include "common.cuh"

namespace gngf {

constexpr int AR_THREADS = 256;
// A dead peer must not hang the GPU: a rank gives up waiting after this many clock cycles (default ~30 s; every rank
// must reach the collective within the bound -- gngf_peer_allreduce_set_timeout_ms).  Giving up is FATAL for the
// result: the error flag state[2] is raised (sticky) and the whole output of the call is NaN, so that neither stale
// nor partial sums can pass for reduced gradients; the host side raises on either signal (dp.PeerAllReduce.check).
static long long g_timeout_cycles = 60000000000ll;

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer_v4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ float ld_peer(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}

// state: [0] epoch, [1] ticket, [2] error flag (1 = a peer did not arrive within the timeout)
__global__ void __launch_bounds__(AR_THREADS)
    peer_allreduce_kernel(float* const* __restrict__ stage_ptrs, uint32_t* const* __restrict__ signal_ptrs, int rank,
                          int world, const float* __restrict__ in, float* __restrict__ out, int64_t n, int64_t cap,
                          float scale, uint32_t* __restrict__ state, long long timeout_cycles) {
  __shared__ int timed_out_s;
  if (threadIdx.x == 0) timed_out_s = 0;
  const uint32_t epoch = state[0] + 1;
  const int64_t par_off = static_cast<int64_t>(epoch & 1u) * cap;
  float* mine = stage_ptrs[rank] + par_off;
  const int tid = threadIdx.x;
  // slice of this block, in units of 4 floats (the buffers are 16-byte aligned; a ragged tail goes element-wise)
  const int64_t n4 = n / 4;
  const int64_t per = (n4 + gridDim.x - 1) / gridDim.x;
  const int64_t lo = blockIdx.x * per, hi = min(n4, lo + per);
  const bool tail = blockIdx.x == gridDim.x - 1 && (n & 3);

  for (int64_t i = lo + tid; i < hi; i += AR_THREADS)
    reinterpret_cast<float4*>(mine)[i] = reinterpret_cast<const float4*>(in)[i];
  if (tail && tid < (n & 3)) mine[n4 * 4 + tid] = in[n4 * 4 + tid];
  __syncthreads();

  if (tid < world) {
    __threadfence_system();
    st_release_sys(signal_ptrs[tid] + static_cast<int64_t>(blockIdx.x) * world + rank, epoch);
    const uint32_t* slot = signal_ptrs[rank] + static_cast<int64_t>(blockIdx.x) * world + tid;
    const long long t0 = clock64();
    while (static_cast<int32_t>(ld_acquire_sys(slot) - epoch) < 0) {
      if (clock64() - t0 > timeout_cycles) {
        state[2] = 1u;
        timed_out_s = 1;
        break;
      }
    }
  }
  __syncthreads();
  // a peer that never arrived, now or in an earlier call (the flag is sticky): poison the output
  const bool dead = timed_out_s || *reinterpret_cast<volatile uint32_t*>(state + 2) != 0u;
  if (dead) scale = __int_as_float(0x7fc00000);

  for (int64_t i = lo + tid; i < hi; i += AR_THREADS) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < world; ++r) {
      const float4 v = ld_peer_v4(stage_ptrs[r] + par_off + i * 4);
      acc.x += v.x;
      acc.y += v.y;
      acc.z += v.z;
      acc.w += v.w;
    }
    reinterpret_cast<float4*>(out)[i] = make_float4(acc.x * scale, acc.y * scale, acc.z * scale, acc.w * scale);
  }
  if (tail && tid < (n & 3)) {
    float acc = 0.0f;
    for (int r = 0; r < world; ++r) acc += ld_peer(stage_ptrs[r] + par_off + n4 * 4 + tid);
    out[n4 * 4 + tid] = acc * scale;
  }

  __syncthreads();
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(state + 1, 1u) == gridDim.x - 1) {
      state[0] = epoch;
      state[1] = 0u;
    }
  }
}

}  // namespace gngf

extern "C" {

int gngf_peer_allreduce(const void* stage_ptrs_dev, const void* signal_ptrs_dev, int32_t rank, int32_t world,
                        const float* in, float* out, int64_t n, int64_t cap_floats, int32_t max_blocks, float scale,
                        uint32_t* state, void* stream) {
  if (!stage_ptrs_dev || !signal_ptrs_dev || !in || !out || !state || world < 1 || world > 32 || rank < 0 ||
      rank >= world || n < 0 || n > cap_floats || max_blocks < 1)
    return GNGF_ERR_INVALID_ARGUMENT;
  if ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) return GNGF_ERR_INVALID_ARGUMENT;
  if (n == 0) return GNGF_OK;
  const int64_t want = gngf::ceil_div(gngf::ceil_div(n, 4), gngf::AR_THREADS);
  const int grid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(want, max_blocks)));
  gngf::peer_allreduce_kernel<<<grid, gngf::AR_THREADS, 0, gngf::as_stream(stream)>>>(
      reinterpret_cast<float* const*>(stage_ptrs_dev), reinterpret_cast<uint32_t* const*>(signal_ptrs_dev), rank, world, in,
      out, n, cap_floats, scale, state, gngf::g_timeout_cycles);
  gngf::note_launch();
  return gngf::check_launch();
}


int gngf_peer_allreduce_set_timeout_ms(int64_t ms) {
  if (ms < 1) return GNGF_ERR_INVALID_ARGUMENT;
  gngf::g_timeout_cycles = ms * 2000000ll;   // clock64 ticks at <= 2 GHz: the bound is at least `ms`
  return GNGF_OK;
}

}  // extern "C"
