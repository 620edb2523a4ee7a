// f-3: the optimizer step that follows every backward (functions.py:96-127, 281: torch.optim.Adam with per-group
// learning rate / weight decay, betas (0.9, 0.99), eps 1e-15) as ONE launch over all parameter tensors.
// torch's fused Adam issues one multi-tensor launch per parameter group plus a step-counter launch per group (6
// launches, ~33 us at the published configuration where the whole step is ~0.25 ms); here the tensor descriptors
// travel in kernel-parameter space, a block finds its tensor with a search over the chunk prefix sums, and the step
// counters (one per tensor, as in torch: a parameter whose gradient is missing in some step does not advance) are
// device scalars advanced by the last block to finish (CUDA-graph friendly: no host state).
//   g' = g + wd p ;  m += (1 - b1)(g' - m) ;  v = b2 v + (1 - b2) g'^2 ;
//   p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)            (torch.optim.Adam, amsgrad = False)
#include "common.cuh"

namespace gngf {

constexpr int ADAM_THREADS = 256;
constexpr int ADAM_CHUNK = ADAM_THREADS * 4 * 4;   // elements per block: 4 float4 per thread

struct AdamArgs {
  gngf_adam_tensor t[GNGF_ADAM_MAX_TENSORS];
  int64_t chunk_end[GNGF_ADAM_MAX_TENSORS];   // prefix sums of the tensors' chunk counts
  int count;
};

__global__ void __launch_bounds__(ADAM_THREADS)
    adam_kernel(const __grid_constant__ AdamArgs a, float beta1, float beta2, float eps,
                unsigned int* __restrict__ ticket) {
  // which tensor does this block's chunk belong to?
  const int64_t b = blockIdx.x;
  int lo = 0, hi = a.count - 1;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (b < a.chunk_end[mid]) hi = mid; else lo = mid + 1;
  }
  const gngf_adam_tensor& T = a.t[lo];
  const int64_t chunk = b - (lo ? a.chunk_end[lo - 1] : 0);
  const int64_t base = chunk * ADAM_CHUNK;
  const bool vec = ((reinterpret_cast<uintptr_t>(T.p) | reinterpret_cast<uintptr_t>(T.g) | reinterpret_cast<uintptr_t>(T.m) |
                     reinterpret_cast<uintptr_t>(T.v)) & 15) == 0;
  // The kernel is a latency chain (step counter -> bias corrections -> loads -> stores -> ticket -> counters) on a few
  // dozen blocks, so the links overlap: every thread first requests its 4 x 4 float4 operands, then thread 0 reads the
  // step counter and evaluates the bias corrections while they are in flight.
  float4 p4[4], g4[4], m4[4], v4[4];
  bool on[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int64_t i = base + (static_cast<int64_t>(r) * ADAM_THREADS + threadIdx.x) * 4;
    on[r] = vec && i + 4 <= T.n;
    if (on[r]) {
      p4[r] = *reinterpret_cast<const float4*>(T.p + i);
      g4[r] = *reinterpret_cast<const float4*>(T.g + i);
      m4[r] = *reinterpret_cast<const float4*>(T.m + i);
      v4[r] = *reinterpret_cast<const float4*>(T.v + i);
    }
  }
  // bias corrections: one double-precision evaluation per block (a per-thread pow() dominated this small kernel)
  __shared__ float bc_s[2];
  __shared__ int last_s;
  if (threadIdx.x == 0) {
    const int t_now = static_cast<int>(*T.step) + 1;   // float32 scalar holding an integer (torch's state-dict dtype)
    bc_s[0] = 1.0f - static_cast<float>(pow(static_cast<double>(beta1), t_now));
    bc_s[1] = 1.0f - static_cast<float>(pow(static_cast<double>(beta2), t_now));
  }
  __syncthreads();
  const float bc1 = bc_s[0], bc2 = bc_s[1];
  const float step_size = T.lr / bc1;
  const float inv_sqrt_bc2 = rsqrtf(bc2);
  const float wd = T.weight_decay, omb1 = 1.0f - beta1, omb2 = 1.0f - beta2;
  auto upd = [&](float& p, float g, float& m, float& v) {
    g = fmaf(wd, p, g);
    m = fmaf(omb1, g - m, m);
    v = fmaf(beta2, v, omb2 * g * g);
    p -= step_size * (m / (sqrtf(v) * inv_sqrt_bc2 + eps));
  };
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int64_t i = base + (static_cast<int64_t>(r) * ADAM_THREADS + threadIdx.x) * 4;
    if (on[r]) {
      upd(p4[r].x, g4[r].x, m4[r].x, v4[r].x);
      upd(p4[r].y, g4[r].y, m4[r].y, v4[r].y);
      upd(p4[r].z, g4[r].z, m4[r].z, v4[r].z);
      upd(p4[r].w, g4[r].w, m4[r].w, v4[r].w);
      *reinterpret_cast<float4*>(T.p + i) = p4[r];
      *reinterpret_cast<float4*>(T.m + i) = m4[r];
      *reinterpret_cast<float4*>(T.v + i) = v4[r];
    } else if (i < T.n) {
      for (int64_t j = i; j < min(i + 4, T.n); ++j) {
        float p = T.p[j], m = T.m[j], v = T.v[j];
        upd(p, T.g[j], m, v);
        T.p[j] = p;
        T.m[j] = m;
        T.v[j] = v;
      }
    }
  }
  // the last block to finish advances the step counters (every block has read its own by then), one thread per counter
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    last_s = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last_s) {
    for (int i = threadIdx.x; i < a.count; i += ADAM_THREADS) *a.t[i].step += 1.0f;
    if (threadIdx.x == 0) *ticket = 0u;
  }
}

}  // namespace gngf

extern "C" {

int gngf_adam_step(const gngf_adam_tensor* tensors, int32_t count, float beta1, float beta2, float eps, uint32_t* ticket,
                   void* stream) {
  if (count < 0 || count > GNGF_ADAM_MAX_TENSORS || !ticket) return GNGF_ERR_INVALID_ARGUMENT;
  gngf::AdamArgs a;
  int64_t chunks = 0;
  int used = 0;
  for (int i = 0; i < count; ++i) {
    if (tensors[i].n < 0) return GNGF_ERR_INVALID_ARGUMENT;
    if (tensors[i].n == 0) continue;
    if (!tensors[i].p || !tensors[i].g || !tensors[i].m || !tensors[i].v || !tensors[i].step)
      return GNGF_ERR_INVALID_ARGUMENT;
    a.t[used] = tensors[i];
    chunks += gngf::ceil_div(tensors[i].n, gngf::ADAM_CHUNK);
    a.chunk_end[used] = chunks;
    ++used;
  }
  a.count = used;
  if (used == 0) return GNGF_OK;
  if (chunks >= (1ll << 31)) return GNGF_ERR_UNSUPPORTED;
  gngf::adam_kernel<<<static_cast<unsigned>(chunks), gngf::ADAM_THREADS, 0, gngf::as_stream(stream)>>>(a, beta1, beta2, eps,
                                                                                                      ticket);
  gngf::note_launch();
  return gngf::check_launch();
}

}  // extern "C"
