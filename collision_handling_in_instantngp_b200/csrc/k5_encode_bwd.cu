// K5a: backward of the feature lookup, again split along the lattice:
//
//   point pass  dnf[s(p,l,v), f] += denc[p, l*F+f] * w_bil[p,l,v]      one vector reduction per corner
//   node pass   table_grad_l[utopi[u,k], f] += dnf[s,f] * w_k           K vector reductions per level node
//               dtv[u,k] += d mix / d topv (closed form of the autograd of models.py:212-217)
//
// The reference scatter-adds K*F values per (point, level, corner) row into the tables
// (embedding_dense_backward, 16 launches) and materialises (rows, T) zero tensors for the top-k adjoint
// (models.py:27-35).  Because mix and gather are linear in the incoming gradient and identical for every row
// that shares a node, summing the F-vector per level node first cuts the atomics by 2K and leaves the
// data-dependent scatter to a pass over S level nodes instead of 4*L*P rows.
#include <algorithm>

#include "common.cuh"

namespace gngf {

// Persistent grid, one thread per point with a loop over levels (see encode_fwd_kernel); the first
// `private_nodes` level nodes (coarsest levels) accumulate in shared memory and are flushed once per CTA.
template <int F>
__global__ void __launch_bounds__(256)
    encode_bwd_kernel(const float2* __restrict__ x, int64_t P, const __grid_constant__ gngf_lattice lat,
                      const float* __restrict__ denc, float* __restrict__ dnf, int private_nodes) {
  extern __shared__ float dnf_s[];
  const int L = lat.num_levels;
  if (private_nodes > 0) {
    for (int i = threadIdx.x; i < private_nodes * F; i += blockDim.x) dnf_s[i] = 0.0f;
    __syncthreads();
  }
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t p = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; p < P; p += stride) {
    const float2 xy = x[p];
    const float* din = denc + p * (L * F);
#pragma unroll 4
    for (int l = 0; l < L; ++l) {
      const Cell c = cell_of(xy.x, xy.y, lat.n[l]);
      bool outside = false;
      float d[F];
      if constexpr (F == 2) {
        const float2 t = reinterpret_cast<const float2*>(din)[l];
        d[0] = t.x;
        d[1] = t.y;
      } else {
#pragma unroll
        for (int f = 0; f < F; ++f) d[f] = din[l * F + f];
      }
      if constexpr (F == 2) {
        // corners (cx, cy) / (cx, cy+1) are adjacent level nodes: one 16-byte reduction when the pair is aligned
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          const int64_t s0 = level_node(lat, l, c.cx + dx, c.cy, outside);
          const int64_t s1 = level_node(lat, l, c.cx + dx, c.cy + 1, outside);
          const float w0 = c.w[dx], w1 = c.w[dx + 2];
          if (s0 < private_nodes && s1 < private_nodes) {
            if (w0 != 0.0f) {
              atomicAdd(dnf_s + s0 * 2, d[0] * w0);
              atomicAdd(dnf_s + s0 * 2 + 1, d[1] * w0);
            }
            if (w1 != 0.0f) {
              atomicAdd(dnf_s + s1 * 2, d[0] * w1);
              atomicAdd(dnf_s + s1 * 2 + 1, d[1] * w1);
            }
          } else if (s1 == s0 + 1 && (s0 & 1) == 0 && s0 >= private_nodes) {
            red_add_v4(dnf + s0 * 2, d[0] * w0, d[1] * w0, d[0] * w1, d[1] * w1);
          } else {
            if (w0 != 0.0f) {
              if (s0 < private_nodes) {
                atomicAdd(dnf_s + s0 * 2, d[0] * w0);
                atomicAdd(dnf_s + s0 * 2 + 1, d[1] * w0);
              } else {
                red_add_v2(dnf + s0 * 2, d[0] * w0, d[1] * w0);
              }
            }
            if (w1 != 0.0f) {
              if (s1 < private_nodes) {
                atomicAdd(dnf_s + s1 * 2, d[0] * w1);
                atomicAdd(dnf_s + s1 * 2 + 1, d[1] * w1);
              } else {
                red_add_v2(dnf + s1 * 2, d[0] * w1, d[1] * w1);
              }
            }
          }
        }
      } else {
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const int64_t s = level_node(lat, l, c.cx + (v & 1), c.cy + (v >> 1), outside);
          const float w = c.w[v];
          if (w == 0.0f) continue;  // x == 1.0 rows: the far corners carry exactly zero weight
          if (s < private_nodes) {
#pragma unroll
            for (int f = 0; f < F; ++f) atomicAdd(dnf_s + s * F + f, d[f] * w);
          } else if constexpr (F % 2 == 0) {
#pragma unroll
            for (int f = 0; f < F; f += 2) red_add_v2(dnf + s * F + f, d[f] * w, d[f + 1] * w);
          } else {
#pragma unroll
            for (int f = 0; f < F; ++f) atomicAdd(dnf + s * F + f, d[f] * w);
          }
        }
      }
    }
  }
  if (private_nodes > 0) {
    __syncthreads();
    for (int i = threadIdx.x; i < private_nodes * F; i += blockDim.x) {
      const float v = dnf_s[i];
      if (v != 0.0f) atomicAdd(dnf + i, v);
    }
  }
}

// One level node: table scatter-add (into `tgrad`: the global gradient table, or a CTA-private copy of it in shared
// memory) and the adjoint of the selected probabilities.
template <bool SHARED>
__device__ __forceinline__ void node_bwd(const gngf_lattice& lat, int l, int64_t i, const float* __restrict__ table,
                                         float* tgrad, int F, int K, int mode, const float* __restrict__ utopv,
                                         const int32_t* __restrict__ utopi, const float* __restrict__ dnf,
                                         float* __restrict__ dtv) {
  const int wy = lat.lwy[l];
  float d[GNGF_MAX_FEATURES];
  bool any = false;
#pragma unroll
  for (int f = 0; f < GNGF_MAX_FEATURES; ++f) {
    d[f] = f < F ? dnf[(lat.loff[l] + i) * F + f] : 0.0f;
    any |= d[f] != 0.0f;
  }
  if (!any) return;  // untouched node (or exactly zero gradient): contributes nothing
  const int cx = lat.lox[l] + static_cast<int>(i / wy), cy = lat.loy[l] + static_cast<int>(i % wy);
  const int64_t u = global_node(lat, cx, cy);
  const float* tv = utopv + u * K;
  const int32_t* ti = utopi + u * K;

  // mix weights (same arithmetic as the forward node pass)
  float mx = 0.0f, norm = 1.0f;
  if (mode == GNGF_MIX_SOFTMAX) {
    mx = tv[0];
    for (int k = 1; k < K; ++k) mx = fmaxf(mx, tv[k]);
    norm = 0.0f;
    for (int k = 0; k < K; ++k) norm += expf(tv[k] - mx);
  } else if (mode == GNGF_MIX_WEIGHTED_AVG) {
    norm = 0.0f;
    for (int k = 0; k < K; ++k) norm += tv[k];
  }
  // pass 1: table gradient + <dw, w>
  float dot = 0.0f;
  for (int k = 0; k < K; ++k) {
    const float w = mode == GNGF_MIX_SOFTMAX ? expf(tv[k] - mx) / norm
                                             : (mode == GNGF_MIX_WEIGHTED_AVG ? tv[k] / norm : tv[k]);
    const int64_t row = static_cast<int64_t>(ti[k]) * F;
    float dw = 0.0f;
#pragma unroll
    for (int f = 0; f < GNGF_MAX_FEATURES; ++f)
      if (f < F) dw = fmaf(d[f], table[row + f], dw);
    dot = fmaf(dw, w, dot);
    if (!SHARED && F == 2) {
      red_add_v2(tgrad + row, d[0] * w, d[1] * w);
    } else {
#pragma unroll
      for (int f = 0; f < GNGF_MAX_FEATURES; ++f)
        if (f < F) atomicAdd(tgrad + row + f, d[f] * w);
    }
  }
  if (!dtv) return;
  // pass 2: adjoint of the top-k probabilities
  for (int k = 0; k < K; ++k) {
    const int64_t row = static_cast<int64_t>(ti[k]) * F;
    float dw = 0.0f;
#pragma unroll
    for (int f = 0; f < GNGF_MAX_FEATURES; ++f)
      if (f < F) dw = fmaf(d[f], table[row + f], dw);
    float g;
    if (mode == GNGF_MIX_SOFTMAX) g = (expf(tv[k] - mx) / norm) * (dw - dot);
    else if (mode == GNGF_MIX_WEIGHTED_AVG) g = (dw - dot) / norm;
    else g = dw;
    atomicAdd(dtv + u * K + k, g);
  }
}

// grid (ceil(max level box / 256), L); thread per level node
__global__ void __launch_bounds__(256)
    node_features_bwd_kernel(const __grid_constant__ gngf_lattice lat, const __grid_constant__ gngf_tables tables,
                             const __grid_constant__ gngf_tables tgrads, int64_t T, int F, int K, int mode,
                             const float* __restrict__ utopv, const int32_t* __restrict__ utopi,
                             const float* __restrict__ dnf, float* __restrict__ dtv) {
  const int l = blockIdx.y;
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= static_cast<int64_t>(lat.lwx[l]) * lat.lwy[l]) return;
  node_bwd<false>(lat, l, i, tables.ptr[l], tgrads.ptr[l], F, K, mode, utopv, utopi, dnf, dtv);
}

// Small tables (T * F * 4 bytes fit shared memory: T = 2^14, F = 2 is 128 KB): millions of level nodes scatter-add into
// a few thousand rows, and same-address reductions in the L2 serialise (64 ms at BASELINE.json configs[3] with T = 2^14:
// 40 M touched level nodes into 16 x 16 384 rows).  A CTA keeps a private copy of its level's gradient table in shared
// memory, walks its share of the level's nodes with shared-memory atomics, and adds the non-zero entries to the global
// table once.  grid (ctas, L): CTA b of level l takes nodes b, b + ctas_l, ... in 512-node strides (ctas_l grows with the
// level: coarse levels have a few hundred nodes).
__global__ void __launch_bounds__(512)
    node_features_bwd_private_kernel(const __grid_constant__ gngf_lattice lat, const __grid_constant__ gngf_tables tables,
                                     const __grid_constant__ gngf_tables tgrads, int64_t T, int F, int K, int mode,
                                     const float* __restrict__ utopv, const int32_t* __restrict__ utopi,
                                     const float* __restrict__ dnf, float* __restrict__ dtv, int nodes_per_cta) {
  extern __shared__ float tg_s[];
  const int l = blockIdx.y;
  const int64_t n = static_cast<int64_t>(lat.lwx[l]) * lat.lwy[l];
  const int64_t ctas = min(static_cast<int64_t>(gridDim.x), (n + nodes_per_cta - 1) / nodes_per_cta);
  if (blockIdx.x >= ctas) return;
  const int64_t TF = T * F;
  for (int64_t e = threadIdx.x; e < TF; e += blockDim.x) tg_s[e] = 0.0f;
  __syncthreads();
  const float* table = tables.ptr[l];
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += ctas * blockDim.x)
    node_bwd<true>(lat, l, i, table, tg_s, F, K, mode, utopv, utopi, dnf, dtv);
  __syncthreads();
  float* tgrad = tgrads.ptr[l];
  for (int64_t e = threadIdx.x; e < TF; e += blockDim.x) {
    const float v = tg_s[e];
    if (v != 0.0f) atomicAdd(tgrad + e, v);
  }
}

template <int F>
__global__ void __launch_bounds__(256)
    encode_hash_bwd_kernel(const float2* __restrict__ x, int64_t P, const __grid_constant__ gngf_lattice lat,
                           const __grid_constant__ gngf_tables tgrads, int64_t T, const float* __restrict__ denc) {
  const int L = lat.num_levels;
  const bool pow2 = (T & (T - 1)) == 0;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t p = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; p < P; p += stride) {
    const float2 xy = x[p];
#pragma unroll 4
    for (int l = 0; l < L; ++l) {
      const Cell c = cell_of(xy.x, xy.y, lat.n[l]);
      float* tgrad = tgrads.ptr[l];
      float d[F];
#pragma unroll
      for (int f = 0; f < F; ++f) d[f] = denc[(p * L + l) * F + f];
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const uint32_t gx = static_cast<uint32_t>(c.cx + (v & 1)), gy = static_cast<uint32_t>(c.cy + (v >> 1));
        const uint32_t h32 = gx ^ (gy * 2654435761u);
        int64_t h;
        if (pow2) {
          h = static_cast<int64_t>(h32 & static_cast<uint32_t>(T - 1));
        } else {
          h = static_cast<int64_t>(static_cast<int32_t>(h32)) % T;
          if (h < 0) h += T;
        }
        const float w = c.w[v];
        if constexpr (F % 2 == 0) {
#pragma unroll
          for (int f = 0; f < F; f += 2) red_add_v2(tgrad + h * F + f, d[f] * w, d[f + 1] * w);
        } else {
#pragma unroll
          for (int f = 0; f < F; ++f) atomicAdd(tgrad + h * F + f, d[f] * w);
        }
      }
    }
  }
}

}  // namespace gngf

extern "C" {

int gngf_encode_bwd(const float* x, int64_t P, gngf_lattice lat, int32_t F, const float* denc, float* dnf,
                    void* stream) {
  if (lat.num_levels <= 0 || lat.num_levels > GNGF_MAX_LEVELS || P < 0) return GNGF_ERR_INVALID_ARGUMENT;
  if (P == 0) return GNGF_OK;
  cudaStream_t st = gngf::as_stream(stream);
  const float2* x2 = reinterpret_cast<const float2*>(x);
  const int priv = gngf::private_node_count(lat, (40 * 1024) / (4 * F));   // <= 40 KB of shared accumulators
  const size_t smem = sizeof(float) * priv * F;
  const unsigned blocks = static_cast<unsigned>(std::min<int64_t>(gngf::ceil_div(P, 256), 8 * gngf::sm_count()));
  switch (F) {
    case 1: gngf::encode_bwd_kernel<1><<<blocks, 256, smem, st>>>(x2, P, lat, denc, dnf, priv); break;
    case 2: gngf::encode_bwd_kernel<2><<<blocks, 256, smem, st>>>(x2, P, lat, denc, dnf, priv); break;
    case 4: gngf::encode_bwd_kernel<4><<<blocks, 256, smem, st>>>(x2, P, lat, denc, dnf, priv); break;
    case 8: gngf::encode_bwd_kernel<8><<<blocks, 256, smem, st>>>(x2, P, lat, denc, dnf, priv); break;
    default: return GNGF_ERR_UNSUPPORTED;
  }
  gngf::note_launch();
  return gngf::check_launch();
}

int gngf_node_features_bwd(gngf_lattice lat, gngf_tables tables, gngf_tables table_grads, int64_t T, int32_t F,
                           int32_t K, int32_t mix_mode, const float* utopv, const int32_t* utopi, const float* dnf,
                           float* dtv, void* stream) {
  if (lat.num_levels <= 0 || lat.num_levels > GNGF_MAX_LEVELS || F <= 0 || F > GNGF_MAX_FEATURES || K <= 0 ||
      K > GNGF_MAX_TOPK)
    return GNGF_ERR_INVALID_ARGUMENT;
  int64_t box = 0;
  for (int l = 0; l < lat.num_levels; ++l) box = std::max<int64_t>(box, static_cast<int64_t>(lat.lwx[l]) * lat.lwy[l]);
  const size_t private_bytes = sizeof(float) * static_cast<size_t>(T) * F;
  if (private_bytes <= 160 * 1024 && box >= (1 << 18)) {
    // small table, many nodes: CTA-private gradient tables in shared memory (one CTA per SM and level)
    if (cudaFuncSetAttribute(gngf::node_features_bwd_private_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             static_cast<int>(private_bytes)) != cudaSuccess)
      return gngf::check_launch();
    const int nodes_per_cta = 8192;
    dim3 grid(static_cast<unsigned>(std::min<int64_t>(gngf::sm_count(), gngf::ceil_div(box, nodes_per_cta))),
              lat.num_levels);
    gngf::node_features_bwd_private_kernel<<<grid, 512, private_bytes, gngf::as_stream(stream)>>>(
        lat, tables, table_grads, T, F, K, mix_mode, utopv, utopi, dnf, dtv, nodes_per_cta);
    gngf::note_launch();
    return gngf::check_launch();
  }
  dim3 grid(static_cast<unsigned>(gngf::ceil_div(box, 256)), lat.num_levels);
  gngf::node_features_bwd_kernel<<<grid, 256, 0, gngf::as_stream(stream)>>>(lat, tables, table_grads, T, F, K,
                                                                            mix_mode, utopv, utopi, dnf, dtv);
  gngf::note_launch();
  return gngf::check_launch();
}

int gngf_encode_hash_bwd(const float* x, int64_t P, gngf_lattice lat, gngf_tables table_grads, int64_t T, int32_t F,
                         const float* denc, void* stream) {
  if (lat.num_levels <= 0 || lat.num_levels > GNGF_MAX_LEVELS || P < 0 || T <= 0) return GNGF_ERR_INVALID_ARGUMENT;
  if (P == 0) return GNGF_OK;
  const unsigned blocks = static_cast<unsigned>(std::min<int64_t>(gngf::ceil_div(P, 256), 8 * gngf::sm_count()));
  cudaStream_t st = gngf::as_stream(stream);
  const float2* x2 = reinterpret_cast<const float2*>(x);
  switch (F) {
    case 1: gngf::encode_hash_bwd_kernel<1><<<blocks, 256, 0, st>>>(x2, P, lat, table_grads, T, denc); break;
    case 2: gngf::encode_hash_bwd_kernel<2><<<blocks, 256, 0, st>>>(x2, P, lat, table_grads, T, denc); break;
    case 4: gngf::encode_hash_bwd_kernel<4><<<blocks, 256, 0, st>>>(x2, P, lat, table_grads, T, denc); break;
    case 8: gngf::encode_hash_bwd_kernel<8><<<blocks, 256, 0, st>>>(x2, P, lat, table_grads, T, denc); break;
    default: return GNGF_ERR_UNSUPPORTED;
  }
  gngf::note_launch();
  return gngf::check_launch();
}

}  // extern "C"
