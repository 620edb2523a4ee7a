// The reference's loss assembly (utils.py:78-174, functions.py:243-245) as one kernel that also emits its own
// adjoints:  total = l_mse * MSE(rgb, target) + sum_l (l_js_kl * level_l + coll_l),
//   pbar_l = colsum_l / rows,  q = 1/N,  m = (pbar + q) / 2
//   kl_l   = sum q (ln q - ln pbar) / N                      (KLDivLoss 'batchmean' on a 1-D vector: / N)
//   js_l   = [sum m (ln m - ln pbar) + sum m (ln m - ln q)] / (2N)   (gradient also through the target m)
//   level_l = -(gamma + epsilon) js_l + epsilon kl_l
// `Loss` itself stays the reference's module when the reference's train loop drives the model; this kernel is
// what bench.py and the fused training step use: ~45 tiny ATen launches (forward + autograd) become one.
#include <algorithm>

#include "common.cuh"

namespace gngf {

__device__ __forceinline__ float block_sum_256(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x % 32, w = threadIdx.x / 32;
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += red[i];
  return s;
}

// blocks [0, L): one level each; blocks >= L: grid-stride MSE.  out: [0] total, [1] mse, [2 + l] level_l
__global__ void __launch_bounds__(256)
    loss_kernel(const float* __restrict__ rgb, const float* __restrict__ target, int64_t n_rgb,
                const float* __restrict__ colsum, int L, int64_t N, float rows, float gamma, float epsilon, float l_mse,
                float l_js_kl, const float* __restrict__ coll, float* __restrict__ out, float* __restrict__ d_rgb,
                float* __restrict__ d_colsum) {
  __shared__ float red[8];
  if (static_cast<int>(blockIdx.x) < L) {
    const int l = blockIdx.x;
    const float q = 1.0f / static_cast<float>(N), lq = logf(q), invN = 1.0f / static_cast<float>(N);
    const float ge = gamma + epsilon;
    float kl = 0.0f, js = 0.0f;
    for (int64_t t = threadIdx.x; t < N; t += 256) {
      const float p = colsum[l * N + t] / rows;
      const float lp = logf(p);
      const float m = 0.5f * (p + q), lm = logf(m);
      kl += q * (lq - lp);
      js += m * (lm - lp) + m * (lm - lq);
      const float dkl = -q / p * invN;
      const float djs = (0.5f * (lm - lp) + 0.5f - m / p + 0.5f * (lm - lq) + 0.5f) * (0.5f * invN);
      d_colsum[l * N + t] = l_js_kl * (-ge * djs + epsilon * dkl) / rows;
    }
    kl = block_sum_256(kl, red);
    js = block_sum_256(js, red);
    if (threadIdx.x == 0) {
      const float level = -ge * (js * invN * 0.5f) + epsilon * (kl * invN);
      out[2 + l] = level;
      atomicAdd(out, l_js_kl * level + (coll ? coll[l] : 1.0f));
    }
    return;
  }
  const float inv_n = 1.0f / static_cast<float>(n_rgb);
  const float gscale = l_mse * 2.0f * inv_n;
  float acc = 0.0f;
  const int64_t stride = static_cast<int64_t>(gridDim.x - L) * 256;
  for (int64_t i = static_cast<int64_t>(blockIdx.x - L) * 256 + threadIdx.x; i < n_rgb; i += stride) {
    const float d = rgb[i] - target[i];
    acc = fmaf(d, d, acc);
    d_rgb[i] = gscale * d;
  }
  acc = block_sum_256(acc, red);
  if (threadIdx.x == 0) {
    atomicAdd(out + 1, acc * inv_n);
    atomicAdd(out, l_mse * acc * inv_n);
  }
}

}  // namespace gngf

extern "C" {

int gngf_loss_fwd_bwd(const float* rgb, const float* target, int64_t n_rgb, const float* colsum, int32_t L, int64_t N,
                      float rows, float gamma, float epsilon, float l_mse, float l_js_kl, const float* coll_term,
                      float* out, float* d_rgb, float* d_colsum, void* stream) {
  if (n_rgb <= 0 || L < 0 || L > GNGF_MAX_LEVELS || (L > 0 && N <= 0) || rows <= 0) return GNGF_ERR_INVALID_ARGUMENT;
  cudaStream_t st = gngf::as_stream(stream);
  if (cudaMemsetAsync(out, 0, sizeof(float) * (2 + L), st) != cudaSuccess) return gngf::check_launch();
  const int mse_blocks = static_cast<int>(std::min<int64_t>(gngf::ceil_div(n_rgb, 256 * 4), 2 * gngf::sm_count()));
  gngf::loss_kernel<<<L + std::max(1, mse_blocks), 256, 0, st>>>(rgb, target, n_rgb, colsum, L, N, rows, gamma, epsilon,
                                                                 l_mse, l_js_kl, coll_term, out, d_rgb, d_colsum);
  gngf::note_launch();
  return gngf::check_launch();
}

// The two halves separately, for a caller that runs them on different streams (trainer.GraphedTrainer: the MSE half
// seeds the decoder backward on the main stream while the divergence half waits for the column sums -- and under data
// parallelism for their all-reduce -- on the side stream).  parts: 1 = MSE (out[0] += l_mse * mse, out[1] = mse, d_rgb),
// 2 = levels (out[0] += sum_l ..., out[2 + l], d_colsum), 3 = both.  `out` (2 + L floats) must be ZERO on entry of the
// first part: no memset here.
int gngf_loss_parts(const float* rgb, const float* target, int64_t n_rgb, const float* colsum, int32_t L, int64_t N,
                    float rows, float gamma, float epsilon, float l_mse, float l_js_kl, const float* coll_term,
                    float* out, float* d_rgb, float* d_colsum, int32_t parts, void* stream) {
  if ((parts & 3) == 0 || !out) return GNGF_ERR_INVALID_ARGUMENT;
  const bool mse = parts & 1, lev = parts & 2;
  if (mse && (n_rgb <= 0 || !rgb || !target || !d_rgb)) return GNGF_ERR_INVALID_ARGUMENT;
  if (lev && (L <= 0 || L > GNGF_MAX_LEVELS || N <= 0 || rows <= 0 || !colsum || !d_colsum))
    return GNGF_ERR_INVALID_ARGUMENT;
  const int Lb = lev ? L : 0;
  const int mse_blocks =
      mse ? std::max(1, static_cast<int>(std::min<int64_t>(gngf::ceil_div(n_rgb, 256 * 4), 2 * gngf::sm_count()))) : 0;
  gngf::loss_kernel<<<Lb + mse_blocks, 256, 0, gngf::as_stream(stream)>>>(rgb, target, mse ? n_rgb : 0, colsum, Lb, N, rows,
                                                                         gamma, epsilon, l_mse, l_js_kl, coll_term, out,
                                                                         d_rgb, d_colsum);
  gngf::note_launch();
  return gngf::check_launch();
}

}  // extern "C"
