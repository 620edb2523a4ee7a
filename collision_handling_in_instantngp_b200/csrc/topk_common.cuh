// Order-preserving keys and the warp-level K-round selection shared by the softmax/top-k kernels.
// Selection order is (value descending, index ascending); see k3_topk.cu.
#pragma once
#include "common.cuh"

namespace gngf {

__device__ __forceinline__ uint32_t ordered_bits(float v) {
  v = v + 0.0f;  // -0 -> +0
  const uint32_t b = __float_as_uint(v);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float from_ordered_bits(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
__device__ __forceinline__ uint64_t make_key(float v, uint32_t idx) {
  return (static_cast<uint64_t>(ordered_bits(v)) << 32) | static_cast<uint64_t>(~idx);
}
__device__ __forceinline__ uint64_t warp_max_u64(uint64_t k) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const uint64_t other = __shfl_xor_sync(0xffffffffu, k, o);
    k = other > k ? other : k;
  }
  return k;
}

// K rounds of selection over `vals` (T entries, lane-strided); lane 0 writes the winners.
template <typename IdxT>
__device__ __forceinline__ void select_topk(const float* vals, int64_t T, int K, int lane, float* topv, IdxT* topi) {
  uint64_t prev = ~0ull;
  for (int k = 0; k < K; ++k) {
    uint64_t best = 0ull;
    for (int64_t t = lane; t < T; t += 32) {
      const uint64_t key = make_key(vals[t], static_cast<uint32_t>(t));
      if (key < prev && key > best) best = key;
    }
    best = warp_max_u64(best);
    if (lane == 0) {
      topi[k] = static_cast<IdxT>(~static_cast<uint32_t>(best));
      topv[k] = from_ordered_bits(static_cast<uint32_t>(best >> 32));
    }
    prev = best;
  }
}

}  // namespace gngf
