// K3: softmax + nan_to_num + top-k, one warp per row (models.py:85, 111, DifferentiableTopk 5-42), and
// K5b: the fused softmax / top-k / column-sum backward that turns per-node adjoints into dlogits.
//
// Selection order is (value descending, index ascending) -- torch.topk leaves ties unspecified; the oracle
// defines them the same way.  Every candidate is a 64-bit key (ordered value bits << 32 | ~index), so one
// unsigned max-reduction picks the winner and "strictly below the previous winner" removes chosen entries
// without any marking: K rounds, each a strided scan + 5 shuffle steps.
#include "topk_common.cuh"

namespace gngf {

constexpr int TOPK_WARPS = 4;
constexpr int SMEM_ROW_MAX = 8192;  // floats cached per warp

// logits (R,T) -> probs (R,T) [optional, may alias], topv/topi (R,K), row_max/row_sum (R) [optional]
__global__ void __launch_bounds__(TOPK_WARPS * 32)
    softmax_topk_kernel(const float* logits, int64_t R, int64_t T, int K, float* probs, float* __restrict__ topv,
                        int32_t* __restrict__ topi, float* __restrict__ row_max, float* __restrict__ row_sum,
                        int cache_row) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int64_t r = static_cast<int64_t>(blockIdx.x) * TOPK_WARPS + warp;
  if (r >= R) return;
  const float* z = logits + r * T;
  float* cache = cache_row ? smem + static_cast<int64_t>(warp) * T : nullptr;
  float* pout = probs ? probs + r * T : nullptr;

  float m = -INFINITY;
  for (int64_t t = lane; t < T; t += 32) {
    const float v = z[t];
    if (cache) cache[t] = v;
    m = fmaxf(m, v);
  }
  m = warp_max(m);
  // Kahan-compensated per-lane sum: rows are up to 2^22 entries long and the normaliser feeds every output
  float s = 0.0f, comp = 0.0f;
  for (int64_t t = lane; t < T; t += 32) {
    const float e = expf((cache ? cache[t] : z[t]) - m);
    if (cache) cache[t] = e;
    const float yk = __fsub_rn(e, comp);
    const float tk = __fadd_rn(s, yk);
    comp = __fsub_rn(__fsub_rn(tk, s), yk);
    s = tk;
  }
  s = warp_sum(s);
  // p = e / sum, NaN -> 0 (nan_to_num; +-inf cannot occur in a softmax output)
  if (cache || pout) {
    for (int64_t t = lane; t < T; t += 32) {
      float p = (cache ? cache[t] : expf(z[t] - m)) / s;
      if (p != p) p = 0.0f;
      if (cache) cache[t] = p;
      if (pout) pout[t] = p;
    }
  }
  __syncwarp();
  if (lane == 0) {
    if (row_max) row_max[r] = m;
    if (row_sum) row_sum[r] = s;
  }
  if (cache) {
    select_topk<int32_t>(cache, T, K, lane, topv + r * K, topi + r * K);
  } else if (pout) {
    select_topk<int32_t>(pout, T, K, lane, topv + r * K, topi + r * K);
  } else {
    // large row, nothing materialised: recompute p on every scan
    uint64_t prev = ~0ull;
    for (int k = 0; k < K; ++k) {
      uint64_t best = 0ull;
      for (int64_t t = lane; t < T; t += 32) {
        float p = expf(z[t] - m) / s;
        if (p != p) p = 0.0f;
        const uint64_t key = make_key(p, static_cast<uint32_t>(t));
        if (key < prev && key > best) best = key;
      }
      best = warp_max_u64(best);
      if (lane == 0) {
        topi[r * K + k] = static_cast<int32_t>(~static_cast<uint32_t>(best));
        topv[r * K + k] = from_ordered_bits(static_cast<uint32_t>(best >> 32));
      }
      prev = best;
    }
  }
}

__global__ void __launch_bounds__(TOPK_WARPS * 32)
    topk_kernel(const float* __restrict__ values, int64_t R, int64_t T, int K, float* __restrict__ topv,
                int64_t* __restrict__ topi) {
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int64_t r = static_cast<int64_t>(blockIdx.x) * TOPK_WARPS + warp;
  if (r >= R) return;
  select_topk<int64_t>(values + r * T, T, K, lane, topv + r * K, topi + r * K);
}

__global__ void __launch_bounds__(256) topk_scatter_kernel(const float* __restrict__ gv,
                                                           const int64_t* __restrict__ topi, int64_t R, int64_t T,
                                                           int K, float* __restrict__ gin) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= R * K) return;
  const int64_t r = i / K;
  const int64_t t = topi[i];
  if (t >= 0 && t < T) gin[r * T + t] = gv[i];
}

// K5b.  One warp per lattice node u (rows [u0, u0 + n_rows) of the node array; `in`/`out` are chunk-local):
//   G[t]      = sum_l c_l * gcol[l,t] + gdense[u,t] + sum_k g_k [t == utopi[u,k]],   c_l = cnt[s(l,u)],
//   g_k       = dtv[u,k] + sum_l c_l * gcol_k[l,k]
//   dlogit[t] = p[t] * (G[t] - <G,p>)
// `in` holds the probabilities, or -- when row_max/row_sum are given -- the logits, and p = exp(z - max)/sum is
// recomputed on the fly (streaming path: the forward kept only the softmax statistics).  out may alias in.
// TPR threads cooperate on one row: a warp (4 rows per CTA) for short rows, a whole 256-thread CTA for long ones.
template <int TPR>
__global__ void __launch_bounds__(TPR == 32 ? TOPK_WARPS * 32 : TPR)
    hpd_dlogits_kernel(const __grid_constant__ gngf_lattice lat, const float* in, int64_t T, int K,
                       const int32_t* __restrict__ utopi, const float* __restrict__ dtv,
                       const int32_t* __restrict__ cnt, const float* __restrict__ gcol,
                       const float* __restrict__ gcol_k, const float* __restrict__ gdense,
                       const float* __restrict__ row_max, const float* __restrict__ row_sum, int64_t u0,
                       int64_t n_rows, float* out_base) {
  constexpr int ROWS = TPR == 32 ? TOPK_WARPS : 1;
  const int sub = threadIdx.x / TPR, lane = threadIdx.x % TPR;
  const int64_t r = static_cast<int64_t>(blockIdx.x) * ROWS + sub;
  if (r >= n_rows) return;
  const int64_t u = u0 + r;
  const int L = lat.num_levels;
  const int cx = lat.ox + static_cast<int>(u / lat.wy), cy = lat.oy + static_cast<int>(u % lat.wy);
  // cl[l] = multiplicity of this node on level l (0 when the node is outside that level's box)
  __shared__ float cl_s[ROWS][GNGF_MAX_LEVELS];
  __shared__ float red_s[TPR / 32];
  float* cl = cl_s[sub];
  if (lane < L) {
    float c = 0.0f;
    if (cnt) {
      const int i = cx - lat.lox[lane], j = cy - lat.loy[lane];
      if (i >= 0 && i < lat.lwx[lane] && j >= 0 && j < lat.lwy[lane])
        c = static_cast<float>(cnt[lat.loff[lane] + static_cast<int64_t>(i) * lat.lwy[lane] + j]);
    }
    cl[lane] = c;
  }
  if (TPR == 32) __syncwarp(); else __syncthreads();
  const float* p = in + r * T;
  float* out = out_base + r * T;
  const bool from_logits = row_max != nullptr;
  const float mx = from_logits ? row_max[u] : 0.0f;
  const float inv = from_logits ? 1.0f / row_sum[u] : 1.0f;
  auto prob_of = [&](float v) -> float {
    if (from_logits) {
      v = expf(v - mx) * inv;
      if (v != v) v = 0.0f;
    }
    return v;
  };

  // sparse (top-k) part: threads stride K; keep (t_k, g_k * p_k) so that `out` may alias `in`
  constexpr int SP = (GNGF_MAX_TOPK + TPR - 1) / TPR;
  float dot = 0.0f;
  float sp_add[SP];
  int sp_t[SP];
#pragma unroll
  for (int it = 0; it < SP; ++it) {
    const int k = lane + TPR * it;
    sp_add[it] = 0.0f;
    sp_t[it] = -1;
    if (k < K) {
      float g = dtv[u * K + k];
      if (gcol_k)
        for (int l = 0; l < L; ++l) g = fmaf(cl[l], gcol_k[l * K + k], g);
      const int t = utopi[u * K + k];
      const float pk = prob_of(p[t]);
      sp_t[it] = t;
      sp_add[it] = pk * g;
      dot += pk * g;
    }
  }
  // ... plus the dense column-sum part
  const float* gd = gdense ? gdense + u * T : nullptr;
  if (gcol || gd) {
    for (int64_t t = lane; t < T; t += TPR) {
      float g = gd ? gd[t] : 0.0f;
      if (gcol)
        for (int l = 0; l < L; ++l) g = fmaf(cl[l], gcol[l * T + t], g);
      dot = fmaf(g, prob_of(p[t]), dot);
    }
  }
  dot = warp_sum(dot);
  if (TPR > 32) {
    if ((lane & 31) == 0) red_s[lane >> 5] = dot;
    __syncthreads();  // also orders every read of p[t_k] above before the writes below when out aliases in
    dot = 0.0f;
#pragma unroll
    for (int w = 0; w < TPR / 32; ++w) dot += red_s[w];
  }
  if (!gcol && !gd && (T & 3) == 0 && ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
    // streaming / top-k-only case: dlogit = -<G,p> * p, vectorised
    const float4* p4 = reinterpret_cast<const float4*>(p);
    float4* o4 = reinterpret_cast<float4*>(out);
    for (int64_t t = lane; t < T / 4; t += TPR) {
      const float4 v = p4[t];
      float4 o;
      o.x = -dot * prob_of(v.x);
      o.y = -dot * prob_of(v.y);
      o.z = -dot * prob_of(v.z);
      o.w = -dot * prob_of(v.w);
      o4[t] = o;
    }
  } else {
    for (int64_t t = lane; t < T; t += TPR) {
      float g = gd ? gd[t] : 0.0f;
      if (gcol)
        for (int l = 0; l < L; ++l) g = fmaf(cl[l], gcol[l * T + t], g);
      out[t] = prob_of(p[t]) * (g - dot);
    }
  }
  if (TPR == 32) __syncwarp(); else __syncthreads();
#pragma unroll
  for (int it = 0; it < SP; ++it)
    if (sp_t[it] >= 0) out[sp_t[it]] += sp_add[it];
}

}  // namespace gngf

extern "C" {

int gngf_softmax_topk_fwd(const float* logits, int64_t R, int64_t T, int32_t K, float* probs, float* topv,
                          int32_t* topi, float* row_max, float* row_sum, void* stream) {
  if (R < 0 || T <= 0 || K <= 0 || K > T || K > GNGF_MAX_TOPK || T >= (1ll << 31)) return GNGF_ERR_INVALID_ARGUMENT;
  if (R == 0) return GNGF_OK;
  const int cache = T <= gngf::SMEM_ROW_MAX;
  const size_t smem = cache ? sizeof(float) * T * gngf::TOPK_WARPS : 0;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(gngf::softmax_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return gngf::check_launch();
  }
  gngf::softmax_topk_kernel<<<static_cast<unsigned>(gngf::ceil_div(R, gngf::TOPK_WARPS)), gngf::TOPK_WARPS * 32, smem,
                              gngf::as_stream(stream)>>>(logits, R, T, K, probs, topv, topi, row_max, row_sum, cache);
  gngf::note_launch();
  return gngf::check_launch();
}

int gngf_topk_fwd(const float* values, int64_t R, int64_t T, int32_t K, float* topv, int64_t* topi, void* stream) {
  if (R < 0 || T <= 0 || K <= 0 || K > T || T >= (1ll << 31)) return GNGF_ERR_INVALID_ARGUMENT;
  if (R == 0) return GNGF_OK;
  gngf::topk_kernel<<<static_cast<unsigned>(gngf::ceil_div(R, gngf::TOPK_WARPS)), gngf::TOPK_WARPS * 32, 0,
                      gngf::as_stream(stream)>>>(values, R, T, K, topv, topi);
  gngf::note_launch();
  return gngf::check_launch();
}

int gngf_topk_bwd(const float* grad_values, const int64_t* topi, int64_t R, int64_t T, int32_t K, float* grad_in,
                  void* stream) {
  if (R < 0 || T <= 0 || K <= 0) return GNGF_ERR_INVALID_ARGUMENT;
  if (R == 0) return GNGF_OK;
  cudaStream_t st = gngf::as_stream(stream);
  if (cudaMemsetAsync(grad_in, 0, sizeof(float) * R * T, st) != cudaSuccess) return gngf::check_launch();
  gngf::topk_scatter_kernel<<<static_cast<unsigned>(gngf::ceil_div(R * K, 256)), 256, 0, st>>>(grad_values, topi, R, T,
                                                                                               K, grad_in);
  gngf::note_launch();
  return gngf::check_launch();
}

int gngf_hpd_dlogits(gngf_lattice lat, const float* uprobs, int64_t T, int32_t K, const int32_t* utopi,
                     const float* dtv, const int32_t* cnt, const float* gcol, const float* gcol_k,
                     const float* gdense, const float* row_max, const float* row_sum, int64_t u0, int64_t n_rows,
                     float* dlogits, void* stream) {
  const int64_t U = static_cast<int64_t>(lat.wx) * lat.wy;
  if (U <= 0 || T <= 0 || K <= 0 || K > T || K > GNGF_MAX_TOPK || u0 < 0 || n_rows < 0 || u0 + n_rows > U)
    return GNGF_ERR_INVALID_ARGUMENT;
  if ((gcol || gcol_k) && !cnt) return GNGF_ERR_INVALID_ARGUMENT;
  if ((row_max == nullptr) != (row_sum == nullptr)) return GNGF_ERR_INVALID_ARGUMENT;
  if (n_rows == 0) return GNGF_OK;
  if (T >= 2048) {
    if (n_rows >= (1ll << 31)) return GNGF_ERR_UNSUPPORTED;
    gngf::hpd_dlogits_kernel<256><<<static_cast<unsigned>(n_rows), 256, 0, gngf::as_stream(stream)>>>(
        lat, uprobs, T, K, utopi, dtv, cnt, gcol, gcol_k, gdense, row_max, row_sum, u0, n_rows, dlogits);
  } else {
    gngf::hpd_dlogits_kernel<32><<<static_cast<unsigned>(gngf::ceil_div(n_rows, gngf::TOPK_WARPS)),
                                   gngf::TOPK_WARPS * 32, 0, gngf::as_stream(stream)>>>(
        lat, uprobs, T, K, utopi, dtv, cnt, gcol, gcol_k, gdense, row_max, row_sum, u0, n_rows, dlogits);
  }
  gngf::note_launch();
  return gngf::check_launch();
}

}  // extern "C"
