// K2/K3/K5b/K5c for SMALL lattices, fused: when the HPD sees only a few hundred to a few thousand nodes (782 at the
// published configuration) its whole forward is ~70 MFLOP -- as a chain of separate GEMM / softmax / top-k launches
// it is pure launch latency (18 launches, ~120 us of a 320 us step).  Here one CTA owns NB = 8 nodes and walks
//   forward : layer 0 from the node coordinates, hidden layers, output layer, softmax + nan_to_num, top-k
//   backward: dlogits (softmax / top-k / column-sum adjoint, see k3_topk.cu), then dX through the layers with the
//             ReLU masks, bias gradients and the first layer's gradients by atomics
// with activations in shared memory and the weights streamed from L2 (coalesced: a warp reads one weight row).
// The weight gradients dW_i = g_i^T h_{i-1} reduce over all nodes and stay with the generic split-K layer kernel.
// Limits: <= 6 layers, hidden widths <= 256, T <= 1024, K <= 128; anything else takes the general path.
#include <algorithm>

#include "topk_common.cuh"

namespace gngf {

constexpr int NB = 8;          // nodes per CTA
constexpr int SM_MAXW = 256;   // widest hidden layer
constexpr int SM_MAXT = 1024;  // widest output layer
constexpr int SM_MAXL = 6;
constexpr int SM_THREADS = 256;

struct HpdNet {
  int n_layers;                 // linear layers, >= 2
  int width[SM_MAXL + 1];       // width[0] = 2, ..., width[n_layers] = T
  const float* w[SM_MAXL];      // (width[i+1], width[i]) row-major
  const float* b[SM_MAXL];
  float* act[SM_MAXL];          // act[i] (U, width[i+1]) outputs of hidden layers i < n_layers-1 (global, saved)
  float* gact[SM_MAXL];         // backward: g[i] (U, width[i+1]) adjoint of layer i's pre-activation (global)
  float* dbias[SM_MAXL];        // backward: bias gradients (+=)
  float* dw0;                   // backward: first layer weight gradient (+=)
};

// The per-level-node passes of the encoding (k4_encode.cu node_features_fwd, k5_encode_bwd.cu node_features_bwd) folded
// into the HPD kernels: a level node belongs to exactly one lattice node, and the warp that owns the node already holds
// its selection.  Two launches leave the critical path of the step.
struct SmallEnc {
  gngf_tables tables;           // forward + backward: level tables (T, F)
  gngf_tables tgrads;           // backward: their gradients (+=)
  int F, mode;
  float* nfeat;                 // forward: (S, F) mixed features per level node; NULL = not fused
  const float* dnf;             // backward: (S, F) adjoint of nfeat; NULL = not fused
};

// 8 partial sums (one per node) x 32 lanes -> each 4-lane group ends with one node's total
__device__ __forceinline__ float reduce_scatter8(const float (&acc)[NB], int lane, int& node_of_lane) {
  const bool hi16 = lane & 16, hi8 = lane & 8, hi4 = lane & 4;
  float a4[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float send = hi16 ? acc[i] : acc[i + 4];
    const float keep = hi16 ? acc[i + 4] : acc[i];
    a4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
  float a2[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float send = hi8 ? a4[i] : a4[i + 2];
    const float keep = hi8 ? a4[i + 2] : a4[i];
    a2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  float a1;
  {
    const float send = hi4 ? a2[0] : a2[1];
    const float keep = hi4 ? a2[1] : a2[0];
    a1 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  a1 += __shfl_xor_sync(0xffffffffu, a1, 2);
  a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
  node_of_lane = (hi16 ? 4 : 0) + (hi8 ? 2 : 0) + (hi4 ? 1 : 0);
  return a1;
}

// One layer for the CTA's NB nodes: out[j][n] = act(sum_k in[k][n] W[j][k] + b[j]).
// Thread = output unit j (a pass covers SM_THREADS of them); the weight rows of the pass stream through shared memory
// in tiles of TK input units (cp.async, double-buffered, rows padded to TK + 4 floats so that the float4 reads of a
// quarter warp hit distinct banks); activations live unit-major ([unit][NB]) so that the NB values of one input unit
// are two broadcast 16-byte loads.  No shuffles: the previous version (lanes stride k, tree-reduce 8 partial sums per
// output) spent its time in the shuffle chains (ncu: short-scoreboard + issue stalls, 24 us for 782 nodes).
constexpr int TK = 32;
constexpr int WT_LD = TK + 4;
constexpr int WT_FLOATS = SM_THREADS * WT_LD;   // one tile buffer

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst))),
               "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// rows [j0, j0 + SM_THREADS) x columns [k0, k0 + TK) of W (N x K) -> tile (zero-filled outside the matrix)
__device__ __forceinline__ void stage_w_tile(const float* __restrict__ W, int N, int K, int j0, int k0, float* tile,
                                             bool vec_ok) {
  if (vec_ok && k0 + TK <= K) {
    for (int e = threadIdx.x; e < SM_THREADS * (TK / 4); e += SM_THREADS) {
      const int r = e / (TK / 4), c = e % (TK / 4);
      float* dst = tile + r * WT_LD + c * 4;
      if (j0 + r < N) cp_async16(dst, W + static_cast<int64_t>(j0 + r) * K + k0 + c * 4);
      else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  } else {
    for (int e = threadIdx.x; e < SM_THREADS * TK; e += SM_THREADS) {
      const int r = e / TK, c = e % TK;
      tile[r * WT_LD + c] = (j0 + r < N && k0 + c < K) ? __ldg(W + static_cast<int64_t>(j0 + r) * K + k0 + c) : 0.0f;
    }
  }
}

// in: [K][NB] unit-major.  hidden layers: out_t [N][NB] unit-major (+ gsave (U,N) global); last layer: logits [NB][ldl]
__device__ __forceinline__ void small_layer(const float* __restrict__ W, const float* __restrict__ bias, int K, int N,
                                            const float* in, float* out_t, float* logits, int ldl, bool relu,
                                            float* gsave, int64_t u0, int64_t U, float* wt) {
  const bool vec_ok = (K & 3) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0;
  const int ktiles = (K + TK - 1) / TK;
  for (int j0 = 0; j0 < N; j0 += SM_THREADS) {
    const int j = j0 + threadIdx.x;
    float acc[NB];
#pragma unroll
    for (int n = 0; n < NB; ++n) acc[n] = 0.0f;
    stage_w_tile(W, N, K, j0, 0, wt, vec_ok);
    cp_async_commit();
    for (int t = 0; t < ktiles; ++t) {
      if (t + 1 < ktiles) {
        stage_w_tile(W, N, K, j0, (t + 1) * TK, wt + ((t + 1) & 1) * WT_FLOATS, vec_ok);
        cp_async_commit();
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncthreads();
      const float* wrow = wt + (t & 1) * WT_FLOATS + threadIdx.x * WT_LD;
      const float* xin = in + t * TK * NB;
      const int kmax = min(TK, K - t * TK);
#pragma unroll 2
      for (int kk = 0; kk < TK; kk += 4) {
        if (kk >= kmax) break;
        const float4 w4 = *reinterpret_cast<const float4*>(wrow + kk);
        const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (kk + i < kmax) {
            const float4 xa = *reinterpret_cast<const float4*>(xin + (kk + i) * NB);
            const float4 xb = *reinterpret_cast<const float4*>(xin + (kk + i) * NB + 4);
            acc[0] = fmaf(wv[i], xa.x, acc[0]); acc[1] = fmaf(wv[i], xa.y, acc[1]);
            acc[2] = fmaf(wv[i], xa.z, acc[2]); acc[3] = fmaf(wv[i], xa.w, acc[3]);
            acc[4] = fmaf(wv[i], xb.x, acc[4]); acc[5] = fmaf(wv[i], xb.y, acc[5]);
            acc[6] = fmaf(wv[i], xb.z, acc[6]); acc[7] = fmaf(wv[i], xb.w, acc[7]);
          }
        }
      }
      __syncthreads();   // the buffer just read is refilled two iterations later
    }
    if (j < N) {
      const float bj = __ldg(bias + j);
#pragma unroll
      for (int n = 0; n < NB; ++n) {
        float v = acc[n] + bj;
        if (relu) v = fmaxf(v, 0.0f);
        acc[n] = v;
      }
      if (logits) {
#pragma unroll
        for (int n = 0; n < NB; ++n) logits[n * ldl + j] = acc[n];
      } else {
        *reinterpret_cast<float4*>(out_t + j * NB) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        *reinterpret_cast<float4*>(out_t + j * NB + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
      }
      if (gsave) {
#pragma unroll
        for (int n = 0; n < NB; ++n)
          if (u0 + n < U) gsave[(u0 + n) * N + j] = acc[n];
      }
    }
  }
}

// Resident-weights variant: when every layer's weight matrix fits in shared memory next to the activations (the
// published HPD: 2-32-64-128-256 = 179 KB padded), ALL of them are requested with cp.async at kernel start -- one
// commit group per layer -- and a layer only waits for its own group.  The weight stream (L2 -> SM) then overlaps
// layer 0 and the earlier layers instead of being paid tile by tile inside every layer.
// Layout: layer i at wres + woff[i], N_i rows of K_i + 4 floats.
__device__ __forceinline__ void stage_w_resident(const float* __restrict__ W, int N, int K, float* dst) {
  const int ld = K + 4, cpr = K / 4;
  for (int e = threadIdx.x; e < N * cpr; e += SM_THREADS) {
    const int r = e / cpr, c = e % cpr;
    cp_async16(dst + r * ld + c * 4, W + static_cast<int64_t>(r) * K + c * 4);
  }
}

__device__ __forceinline__ void small_layer_resident(const float* wsm, const float* __restrict__ bias, int K, int N,
                                                     const float* in, float* out_t, float* logits, int ldl, bool relu,
                                                     float* gsave, int64_t u0, int64_t U) {
  const int j = threadIdx.x;   // N <= SM_THREADS
  if (j >= N) return;
  float acc[NB];
#pragma unroll
  for (int n = 0; n < NB; ++n) acc[n] = 0.0f;
  const float* wrow = wsm + j * (K + 4);
#pragma unroll 2
  for (int k = 0; k < K; k += 4) {
    const float4 w4 = *reinterpret_cast<const float4*>(wrow + k);
    const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 xa = *reinterpret_cast<const float4*>(in + (k + i) * NB);
      const float4 xb = *reinterpret_cast<const float4*>(in + (k + i) * NB + 4);
      acc[0] = fmaf(wv[i], xa.x, acc[0]); acc[1] = fmaf(wv[i], xa.y, acc[1]);
      acc[2] = fmaf(wv[i], xa.z, acc[2]); acc[3] = fmaf(wv[i], xa.w, acc[3]);
      acc[4] = fmaf(wv[i], xb.x, acc[4]); acc[5] = fmaf(wv[i], xb.y, acc[5]);
      acc[6] = fmaf(wv[i], xb.z, acc[6]); acc[7] = fmaf(wv[i], xb.w, acc[7]);
    }
  }
  const float bj = __ldg(bias + j);
#pragma unroll
  for (int n = 0; n < NB; ++n) {
    float v = acc[n] + bj;
    if (relu) v = fmaxf(v, 0.0f);
    acc[n] = v;
  }
  if (logits) {
#pragma unroll
    for (int n = 0; n < NB; ++n) logits[n * ldl + j] = acc[n];
  } else {
    *reinterpret_cast<float4*>(out_t + j * NB) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    *reinterpret_cast<float4*>(out_t + j * NB + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
  if (gsave) {
#pragma unroll
    for (int n = 0; n < NB; ++n)
      if (u0 + n < U) gsave[(u0 + n) * N + j] = acc[n];
  }
}

__global__ void __launch_bounds__(SM_THREADS)
    hpd_small_fwd_kernel(const __grid_constant__ gngf_lattice lat, const __grid_constant__ HpdNet net,
                         const __grid_constant__ SmallEnc enc, int K, float* __restrict__ uprobs, float* utopv,
                         int32_t* utopi, int resident) {
  extern __shared__ float sm[];
  float* bufA = sm;                       // [SM_MAXW][NB] unit-major activations
  float* bufB = sm + NB * SM_MAXW;        // [SM_MAXW][NB]
  float* logits = sm + 2 * NB * SM_MAXW;  // [NB][T]
  const int64_t U = static_cast<int64_t>(lat.wx) * lat.wy;
  const int64_t u0 = static_cast<int64_t>(blockIdx.x) * NB;
  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
  const int nl = net.n_layers, T = net.width[nl];
  float* wt = logits + NB * T;            // [2][SM_THREADS][WT_LD] weight tiles, or all layers' weights (resident)
  if (resident) {   // request every layer's weights now; group i-1 <-> layer i
    float* dst = wt;
    for (int i = 1; i < nl; ++i) {
      stage_w_resident(net.w[i], net.width[i + 1], net.width[i], dst);
      cp_async_commit();
      dst += net.width[i + 1] * (net.width[i] + 4);
    }
  }

  // layer 0 from the node coordinates (models.py:416: the HPD input is the integer corner)
  {
    const int N0 = net.width[1];
    const float2* w0 = reinterpret_cast<const float2*>(net.w[0]);
    for (int e = tid; e < NB * N0; e += SM_THREADS) {
      const int j = e / NB, n = e % NB;
      const int64_t u = min(u0 + n, U - 1);
      const float cx = static_cast<float>(lat.ox + static_cast<int>(u / lat.wy));
      const float cy = static_cast<float>(lat.oy + static_cast<int>(u % lat.wy));
      const float2 w = w0[j];
      float v = fmaf(cy, w.y, fmaf(cx, w.x, net.b[0][j]));
      v = fmaxf(v, 0.0f);
      bufA[j * NB + n] = v;
      if (u0 + n < U) net.act[0][(u0 + n) * N0 + j] = v;
    }
  }
  __syncthreads();
  float* in = bufA;
  float* out = bufB;
  const float* wres = wt;
  for (int i = 1; i < nl; ++i) {
    const int Kd = net.width[i], N = net.width[i + 1];
    const bool last = i == nl - 1;
    if (resident) {
      // groups are committed in layer order: layer i may proceed once at most (nl - 1 - i) newer groups are pending
      switch (nl - 1 - i) {
        case 0: cp_async_wait<0>(); break;
        case 1: cp_async_wait<1>(); break;
        case 2: cp_async_wait<2>(); break;
        case 3: cp_async_wait<3>(); break;
        default: cp_async_wait<4>(); break;
      }
      __syncthreads();
      small_layer_resident(wres, net.b[i], Kd, N, in, out, last ? logits : nullptr, T, !last,
                           last ? nullptr : net.act[i], u0, U);
      wres += N * (Kd + 4);
    } else {
      small_layer(net.w[i], net.b[i], Kd, N, in, out, last ? logits : nullptr, T, !last, last ? nullptr : net.act[i],
                  u0, U, wt);
    }
    __syncthreads();
    if (!last) {
      float* t = in;
      in = out;
      out = t;
    }
  }
  // softmax + nan_to_num + top-k: warp n owns node n (NB == number of warps)
  {
    const int n = warp;
    const int64_t u = u0 + n;
    if (u < U) {
      float* z = logits + n * T;
      float m = -INFINITY;
      for (int t = lane; t < T; t += 32) m = fmaxf(m, z[t]);
      m = warp_max(m);
      float ssum = 0.0f;
      for (int t = lane; t < T; t += 32) {
        const float e = expf(z[t] - m);
        z[t] = e;
        ssum += e;
      }
      ssum = warp_sum(ssum);
      for (int t = lane; t < T; t += 32) {
        float p = z[t] / ssum;
        if (p != p) p = 0.0f;
        z[t] = p;
        uprobs[u * T + t] = p;
      }
      __syncwarp();
      select_topk<int32_t>(z, T, K, lane, utopv + u * K, utopi + u * K);
      if (enc.nfeat) {
        // node pass of the encoding for this node's level nodes (same arithmetic as node_features_fwd_kernel); lane = level
        __syncwarp();
        const int cx = lat.ox + static_cast<int>(u / lat.wy), cy = lat.oy + static_cast<int>(u % lat.wy);
        const float* tv = utopv + u * K;
        const int32_t* ti = utopi + u * K;
        const int F = enc.F, mode = enc.mode;
        for (int l = lane; l < lat.num_levels; l += 32) {
          const int i = cx - lat.lox[l], j = cy - lat.loy[l];
          if (i < 0 || i >= lat.lwx[l] || j < 0 || j >= lat.lwy[l]) continue;
          const float* table = enc.tables.ptr[l];
          for (int k = 0; k < K; ++k) prefetch_l1(table + static_cast<int64_t>(ti[k]) * F);   // K misses in flight, not in turn
          float mx;
          const float norm = mix_weight_norm(tv, K, mode, mx);
          float acc[GNGF_MAX_FEATURES];
#pragma unroll
          for (int f = 0; f < GNGF_MAX_FEATURES; ++f) acc[f] = 0.0f;
          for (int k = 0; k < K; ++k) {
            const float w = (mode == GNGF_MIX_WEIGHTED_AVG) ? tv[k] : mix_weight(tv[k], mode, mx, norm);
            const float* row = table + static_cast<int64_t>(ti[k]) * F;
#pragma unroll
            for (int f = 0; f < GNGF_MAX_FEATURES; ++f)
              if (f < F) acc[f] = fmaf(__ldg(row + f), w, acc[f]);
          }
          float* out = enc.nfeat + (lat.loff[l] + static_cast<int64_t>(i) * lat.lwy[l] + j) * F;
#pragma unroll
          for (int f = 0; f < GNGF_MAX_FEATURES; ++f)
            if (f < F) out[f] = (mode == GNGF_MIX_WEIGHTED_AVG) ? acc[f] / norm : acc[f];
        }
      }
    }
  }
}

// backward over NB nodes per CTA.  gT buffers are stored transposed ([width][NB]) so that a thread reads the NB
// adjoints of one unit with two 16-byte loads.
__global__ void __launch_bounds__(SM_THREADS)
    hpd_small_bwd_kernel(const __grid_constant__ gngf_lattice lat, const __grid_constant__ HpdNet net,
                         const __grid_constant__ SmallEnc enc, int K, const float* __restrict__ uprobs, const int32_t* __restrict__ utopi,
                         const float* __restrict__ dtv, const int32_t* __restrict__ cnt,
                         const float* __restrict__ gcol, const float* __restrict__ gcol_k,
                         const float* __restrict__ gdense, int gt_rows, int resident) {
  extern __shared__ float sm[];
  const int nl = net.n_layers, T = net.width[nl];
  float* gT = sm;                         // [gt_rows = max width][NB] adjoint of the current layer's pre-activation
  float* gN = sm + gt_rows * NB;          // [SM_MAXW][NB] next (lower) layer's adjoint
  float* part = gN + SM_MAXW * NB;        // [2][SM_MAXW][NB] partial sums of the two j-halves
  float* wtile = part + 2 * SM_MAXW * NB; // [2][32][SM_MAXW] weight-row tiles of the dX products, or -- resident -- every
                                          // layer's (N_i, K_i) matrix, top layer first, requested at kernel start
  __shared__ float cl_s[NB][GNGF_MAX_LEVELS];
  __shared__ float dtv_s[NB][GNGF_MAX_TOPK];
  const int64_t U = static_cast<int64_t>(lat.wx) * lat.wy;
  const int64_t u0 = static_cast<int64_t>(blockIdx.x) * NB;
  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
  const int L = lat.num_levels;
  if (resident) {   // one cp.async group per layer, in the order the backward walks them: they land during the dlogits phase
    float* dst = wtile;
    for (int i = nl - 1; i >= 1; --i) {
      const int N = net.width[i + 1], Kd = net.width[i], cpr = Kd / 4;
      for (int e = tid; e < N * cpr; e += SM_THREADS) {
        const int r = e / cpr, c = e % cpr;
        cp_async16(dst + r * Kd + c * 4, net.w[i] + static_cast<int64_t>(r) * Kd + c * 4);
      }
      cp_async_commit();
      dst += N * Kd;
    }
  }
  const float* wres = wtile;

  // ---- adjoint of the selected probabilities: the incoming one, plus -- fused -- the encoding's node pass ----
  for (int e = tid; e < NB * K; e += SM_THREADS) {
    const int n = e / K, k = e % K;
    dtv_s[n][k] = (dtv && u0 + n < U) ? dtv[(u0 + n) * K + k] : 0.0f;
  }
  __syncthreads();
  if (enc.dnf) {
    // thread = (node, level) = one level node: same arithmetic as node_features_bwd_kernel (k5_encode_bwd.cu)
    const int F = enc.F, mode = enc.mode;
    for (int e = tid; e < NB * L; e += SM_THREADS) {
      const int n = e / L, l = e % L;
      const int64_t u = u0 + n;
      if (u >= U) continue;
      const int cx = lat.ox + static_cast<int>(u / lat.wy), cy = lat.oy + static_cast<int>(u % lat.wy);
      const int i = cx - lat.lox[l], j = cy - lat.loy[l];
      if (i < 0 || i >= lat.lwx[l] || j < 0 || j >= lat.lwy[l]) continue;
      const float* dn = enc.dnf + (lat.loff[l] + static_cast<int64_t>(i) * lat.lwy[l] + j) * F;
      float d[GNGF_MAX_FEATURES];
      bool any = false;
#pragma unroll
      for (int f = 0; f < GNGF_MAX_FEATURES; ++f) {
        d[f] = f < F ? dn[f] : 0.0f;
        any |= d[f] != 0.0f;
      }
      if (!any) continue;   // untouched level node
      const float* p = uprobs + u * T;
      const int32_t* ti = utopi + u * K;
      const float* table = enc.tables.ptr[l];
      float* tgrad = enc.tgrads.ptr[l];
      float mx = 0.0f, norm = 1.0f;
      if (mode == GNGF_MIX_SOFTMAX) {
        mx = p[ti[0]];
        for (int k = 1; k < K; ++k) mx = fmaxf(mx, p[ti[k]]);
        norm = 0.0f;
        for (int k = 0; k < K; ++k) norm += expf(p[ti[k]] - mx);
      } else if (mode == GNGF_MIX_WEIGHTED_AVG) {
        norm = 0.0f;
        for (int k = 0; k < K; ++k) norm += p[ti[k]];
      }
      float dotw = 0.0f;
      for (int k = 0; k < K; ++k) {
        const float tvk = p[ti[k]];
        const float w = mode == GNGF_MIX_SOFTMAX ? expf(tvk - mx) / norm
                                                 : (mode == GNGF_MIX_WEIGHTED_AVG ? tvk / norm : tvk);
        const int64_t row = static_cast<int64_t>(ti[k]) * F;
        float dw = 0.0f;
#pragma unroll
        for (int f = 0; f < GNGF_MAX_FEATURES; ++f)
          if (f < F) dw = fmaf(d[f], __ldg(table + row + f), dw);
        dotw = fmaf(dw, w, dotw);
        if (F == 2) {
          red_add_v2(tgrad + row, d[0] * w, d[1] * w);
        } else {
#pragma unroll
          for (int f = 0; f < GNGF_MAX_FEATURES; ++f)
            if (f < F) atomicAdd(tgrad + row + f, d[f] * w);
        }
      }
      for (int k = 0; k < K; ++k) {
        const float tvk = p[ti[k]];
        const int64_t row = static_cast<int64_t>(ti[k]) * F;
        float dw = 0.0f;
#pragma unroll
        for (int f = 0; f < GNGF_MAX_FEATURES; ++f)
          if (f < F) dw = fmaf(d[f], __ldg(table + row + f), dw);
        float g;
        if (mode == GNGF_MIX_SOFTMAX) g = (expf(tvk - mx) / norm) * (dw - dotw);
        else if (mode == GNGF_MIX_WEIGHTED_AVG) g = (dw - dotw) / norm;
        else g = dw;
        atomicAdd(&dtv_s[n][k], g);
      }
    }
    __syncthreads();
  }

  // ---- dlogits of node (u0 + warp): same arithmetic as hpd_dlogits_kernel ----
  {
    const int n = warp;
    const int64_t u = u0 + n;
    if (u < U) {
      const int cx = lat.ox + static_cast<int>(u / lat.wy), cy = lat.oy + static_cast<int>(u % lat.wy);
      float* cl = cl_s[n];
      if (lane < L) {
        float c = 0.0f;
        if (cnt) {
          const int i = cx - lat.lox[lane], j = cy - lat.loy[lane];
          if (i >= 0 && i < lat.lwx[lane] && j >= 0 && j < lat.lwy[lane])
            c = static_cast<float>(cnt[lat.loff[lane] + static_cast<int64_t>(i) * lat.lwy[lane] + j]);
        }
        cl[lane] = c;
      }
      __syncwarp();
      const float* p = uprobs + u * T;
      const float* gd = gdense ? gdense + u * T : nullptr;
      // adjoint of the K selected probabilities (dtv_s: the incoming one plus the encoding's node pass above); lane
      // owns k = lane + 32 m
      float dtv_r[(GNGF_MAX_TOPK + 31) / 32];
#pragma unroll
      for (int m = 0; m < (GNGF_MAX_TOPK + 31) / 32; ++m) dtv_r[m] = lane + 32 * m < K ? dtv_s[n][lane + 32 * m] : 0.0f;
      float dot = 0.0f;
#pragma unroll
      for (int m = 0; m < (GNGF_MAX_TOPK + 31) / 32; ++m) {
        const int k = lane + 32 * m;
        if (k < K) {
          float g = dtv_r[m];
          if (gcol_k)
            for (int l = 0; l < L; ++l) g = fmaf(cl[l], gcol_k[l * K + k], g);
          dtv_r[m] = g;   // complete adjoint of slot k
          dot = fmaf(g, p[utopi[u * K + k]], dot);
        }
      }
      if (gcol || gd) {
        for (int t = lane; t < T; t += 32) {
          float g = gd ? gd[t] : 0.0f;
          if (gcol)
            for (int l = 0; l < L; ++l) g = fmaf(cl[l], gcol[l * T + t], g);
          dot = fmaf(g, p[t], dot);
        }
      }
      dot = warp_sum(dot);
      for (int t = lane; t < T; t += 32) {
        float g = gd ? gd[t] : 0.0f;
        if (gcol)
          for (int l = 0; l < L; ++l) g = fmaf(cl[l], gcol[l * T + t], g);
        gT[t * NB + n] = p[t] * (g - dot);
      }
      __syncwarp();
#pragma unroll
      for (int m = 0; m < (GNGF_MAX_TOPK + 31) / 32; ++m) {
        const int k = lane + 32 * m;
        if (k < K) {
          const int t = utopi[u * K + k];
          gT[t * NB + n] += p[t] * dtv_r[m];
        }
      }
    } else {
      for (int t = lane; t < T; t += 32) gT[t * NB + n] = 0.0f;
    }
  }
  __syncthreads();

  // ---- walk down the layers: store g_i, bias gradient, dX with the ReLU mask ----
  for (int i = nl - 1; i >= 0; --i) {
    const int N = net.width[i + 1], Kd = net.width[i];
    // g_i to global (the weight-gradient GEMM reads it) + bias gradient
    for (int e = tid; e < N * NB; e += SM_THREADS) {
      const int j = e / NB, n = e % NB;
      if (u0 + n < U) net.gact[i][(u0 + n) * N + j] = gT[e];
    }
    for (int j = tid; j < N; j += SM_THREADS) {
      const float4 a = *reinterpret_cast<const float4*>(gT + j * NB);
      const float4 b = *reinterpret_cast<const float4*>(gT + j * NB + 4);
      atomicAdd(net.dbias[i] + j, ((a.x + a.y) + (a.z + a.w)) + ((b.x + b.y) + (b.z + b.w)));
    }
    if (i == 0) {
      // first layer: dw0[j] += sum_n g0[n][j] * (cx_n, cy_n)
      for (int j = tid; j < N; j += SM_THREADS) {
        float sx = 0.0f, sy = 0.0f;
#pragma unroll
        for (int n = 0; n < NB; ++n) {
          const int64_t u = min(u0 + n, U - 1);
          const float g = gT[j * NB + n];   // zero for padded nodes
          sx = fmaf(g, static_cast<float>(lat.ox + static_cast<int>(u / lat.wy)), sx);
          sy = fmaf(g, static_cast<float>(lat.oy + static_cast<int>(u % lat.wy)), sy);
        }
        atomicAdd(net.dw0 + j * 2 + 0, sx);
        atomicAdd(net.dw0 + j * 2 + 1, sy);
      }
      break;
    }
    // dX[n][k] = sum_j g_i[n][j] * W_i[j][k]; thread = (k, half of the j range); coalesced weight rows
    {
      const int k = tid % SM_MAXW >= Kd ? -1 : tid % SM_MAXW;   // Kd <= 256 == SM_THREADS: one k per thread ...
      const int halves = Kd <= SM_THREADS / 2 ? 2 : 1;          // ... and two j-halves when Kd <= 128
      const int kk = halves == 2 ? tid % (SM_THREADS / 2) : tid;
      const int hf = halves == 2 ? tid / (SM_THREADS / 2) : 0;
      (void)k;
      float acc[NB];
#pragma unroll
      for (int n = 0; n < NB; ++n) acc[n] = 0.0f;
      if (resident) {
        // layer i's matrix is group (nl - 1 - i): wait until at most (i - 1) newer groups are pending
        switch (i - 1) {
          case 0: cp_async_wait<0>(); break;
          case 1: cp_async_wait<1>(); break;
          case 2: cp_async_wait<2>(); break;
          case 3: cp_async_wait<3>(); break;
          default: cp_async_wait<4>(); break;
        }
        __syncthreads();
        if (kk < Kd) {
#pragma unroll 4
          for (int j = hf; j < N; j += halves) {
            const float w = wres[j * Kd + kk];
            const float4 a = *reinterpret_cast<const float4*>(gT + j * NB);
            const float4 b = *reinterpret_cast<const float4*>(gT + j * NB + 4);
            acc[0] = fmaf(w, a.x, acc[0]); acc[1] = fmaf(w, a.y, acc[1]);
            acc[2] = fmaf(w, a.z, acc[2]); acc[3] = fmaf(w, a.w, acc[3]);
            acc[4] = fmaf(w, b.x, acc[4]); acc[5] = fmaf(w, b.y, acc[5]);
            acc[6] = fmaf(w, b.z, acc[6]); acc[7] = fmaf(w, b.w, acc[7]);
          }
#pragma unroll
          for (int n = 0; n < NB; ++n) part[(hf * SM_MAXW + kk) * NB + n] = acc[n];
        }
        wres += N * Kd;
      } else
      {
        // W_i rows stream through shared memory in tiles of TJ output units (cp.async, double-buffered): the loop was
        // bound by the L2 latency of its weight loads.  Thread (kk, hf) takes rows hf, hf + halves, ... of every tile.
        constexpr int TJ = 32;
        const float* Wg = net.w[i];
        const bool vec_ok = (Kd & 3) == 0 && (reinterpret_cast<uintptr_t>(Wg) & 15) == 0;
        const int tiles = (N + TJ - 1) / TJ;
        auto stage = [&](int t, float* buf) {
          const int jb = t * TJ;
          if (vec_ok) {
            const int cpr = Kd / 4;   // 16-byte chunks per row
            for (int e = tid; e < TJ * cpr; e += SM_THREADS) {
              const int r = e / cpr, c = e % cpr;
              if (jb + r < N) cp_async16(buf + r * SM_MAXW + c * 4, Wg + static_cast<int64_t>(jb + r) * Kd + c * 4);
            }
          } else {
            for (int e = tid; e < TJ * Kd; e += SM_THREADS) {
              const int r = e / Kd, c = e % Kd;
              if (jb + r < N) buf[r * SM_MAXW + c] = __ldg(Wg + static_cast<int64_t>(jb + r) * Kd + c);
            }
          }
        };
        stage(0, wtile);
        cp_async_commit();
        for (int t = 0; t < tiles; ++t) {
          if (t + 1 < tiles) {
            stage(t + 1, wtile + ((t + 1) & 1) * TJ * SM_MAXW);
            cp_async_commit();
            cp_async_wait<1>();
          } else {
            cp_async_wait<0>();
          }
          __syncthreads();
          if (kk < Kd) {
            const float* buf = wtile + (t & 1) * TJ * SM_MAXW;
            const int rows = min(TJ, N - t * TJ);
#pragma unroll 4
            for (int r = hf; r < rows; r += halves) {
              const float w = buf[r * SM_MAXW + kk];
              const int j = t * TJ + r;
              const float4 a = *reinterpret_cast<const float4*>(gT + j * NB);
              const float4 b = *reinterpret_cast<const float4*>(gT + j * NB + 4);
              acc[0] = fmaf(w, a.x, acc[0]); acc[1] = fmaf(w, a.y, acc[1]);
              acc[2] = fmaf(w, a.z, acc[2]); acc[3] = fmaf(w, a.w, acc[3]);
              acc[4] = fmaf(w, b.x, acc[4]); acc[5] = fmaf(w, b.y, acc[5]);
              acc[6] = fmaf(w, b.z, acc[6]); acc[7] = fmaf(w, b.w, acc[7]);
            }
          }
          __syncthreads();
        }
        if (kk < Kd) {
#pragma unroll
          for (int n = 0; n < NB; ++n) part[(hf * SM_MAXW + kk) * NB + n] = acc[n];
        }
      }
      __syncthreads();
      // combine the halves, apply the ReLU mask of h_{i-1} (saved forward activation), -> gN
      for (int e = tid; e < Kd * NB; e += SM_THREADS) {
        const int k2 = e / NB, n = e % NB;
        float v = part[e];
        if (halves == 2) v += part[SM_MAXW * NB + e];
        const int64_t u = u0 + n;
        const float h = (u < U) ? net.act[i - 1][u * Kd + k2] : 0.0f;
        gN[e] = h > 0.0f ? v : 0.0f;
      }
      __syncthreads();
      for (int e = tid; e < Kd * NB; e += SM_THREADS) gT[e] = gN[e];
      __syncthreads();
    }
  }
}

static size_t small_fwd_smem(int T) { return sizeof(float) * (2 * NB * SM_MAXW + NB * T + 2 * WT_FLOATS); }
// resident-weights forward: every hidden/output width <= SM_THREADS, K % 4 == 0, and everything fits in 227 KB
static size_t small_fwd_resident_smem(int n_layers, const int32_t* widths, const float* const* w) {
  size_t wfl = 0;
  for (int i = 1; i < n_layers; ++i) {
    if (widths[i + 1] > SM_THREADS || (widths[i] & 3) || (reinterpret_cast<uintptr_t>(w[i]) & 15)) return 0;
    wfl += static_cast<size_t>(widths[i + 1]) * (widths[i] + 4);
  }
  const size_t bytes = sizeof(float) * (2 * NB * SM_MAXW + NB * widths[n_layers] + wfl);
  return (n_layers <= 6 && bytes <= 227 * 1024 - 1024) ? bytes : 0;
}
static int small_bwd_gt_rows(int n_layers, const int32_t* widths) {
  int m = SM_MAXW;
  for (int i = 1; i <= n_layers; ++i) m = std::max(m, widths[i]);
  return m;
}
static size_t small_bwd_smem(int gt_rows) {
  return sizeof(float) * (static_cast<size_t>(gt_rows) * NB + SM_MAXW * NB + 2 * SM_MAXW * NB + 2 * 32 * SM_MAXW);
}
// resident-weights backward: all (N_i, K_i) matrices of layers >= 1 in shared memory (K_i % 4 == 0)
static size_t small_bwd_resident_smem(int n_layers, const int32_t* widths, const float* const* w, int gt_rows) {
  size_t wfl = 0;
  for (int i = 1; i < n_layers; ++i) {
    if ((widths[i] & 3) || (reinterpret_cast<uintptr_t>(w[i]) & 15)) return 0;
    wfl += static_cast<size_t>(widths[i + 1]) * widths[i];
  }
  const size_t bytes = sizeof(float) * (static_cast<size_t>(gt_rows) * NB + SM_MAXW * NB + 2 * SM_MAXW * NB + wfl);
  return (n_layers <= 6 && bytes <= 227 * 1024 - 6144) ? bytes : 0;   // (5 KB of static shared memory in the kernel)
}

}  // namespace gngf

extern "C" {

int gngf_hpd_small_supported(int32_t n_layers, const int32_t* widths, int32_t topk) {
  if (n_layers < 2 || n_layers > gngf::SM_MAXL || widths[0] != 2 || topk < 1 || topk > GNGF_MAX_TOPK) return 0;
  for (int i = 1; i < n_layers; ++i)
    if (widths[i] < 1 || widths[i] > gngf::SM_MAXW) return 0;
  return widths[n_layers] >= topk && widths[n_layers] <= gngf::SM_MAXT;
}

// w / b / act: arrays of n_layers device pointers (act: n_layers - 1 hidden outputs, (U, width))
int gngf_hpd_small_fwd_enc(gngf_lattice lat, int32_t n_layers, const int32_t* widths, const float* const* w,
                           const float* const* b, float* const* act, int32_t topk, float* uprobs, float* utopv,
                           int32_t* utopi, gngf_tables tables, int32_t F, int32_t mix_mode, float* nfeat, void* stream) {
  if (!gngf_hpd_small_supported(n_layers, widths, topk)) return GNGF_ERR_UNSUPPORTED;
  const int64_t U = static_cast<int64_t>(lat.wx) * lat.wy;
  if (U <= 0) return GNGF_ERR_INVALID_ARGUMENT;
  if (nfeat && (F <= 0 || F > GNGF_MAX_FEATURES || lat.num_levels <= 0 || lat.num_levels > GNGF_MAX_LEVELS))
    return GNGF_ERR_INVALID_ARGUMENT;
  gngf::SmallEnc enc{};
  enc.tables = tables;
  enc.F = F;
  enc.mode = mix_mode;
  enc.nfeat = nfeat;
  gngf::HpdNet net{};
  net.n_layers = n_layers;
  for (int i = 0; i <= n_layers; ++i) net.width[i] = widths[i];
  for (int i = 0; i < n_layers; ++i) {
    net.w[i] = w[i];
    net.b[i] = b[i];
    net.act[i] = i < n_layers - 1 ? act[i] : nullptr;
  }
  const size_t res = gngf::small_fwd_resident_smem(n_layers, widths, w);
  const size_t smem = res ? res : gngf::small_fwd_smem(widths[n_layers]);
  if (cudaFuncSetAttribute(gngf::hpd_small_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           static_cast<int>(smem)) != cudaSuccess)
    return gngf::check_launch();
  gngf::hpd_small_fwd_kernel<<<static_cast<unsigned>(gngf::ceil_div(U, gngf::NB)), gngf::SM_THREADS, smem,
                               gngf::as_stream(stream)>>>(lat, net, enc, topk, uprobs, utopv, utopi, res ? 1 : 0);
  gngf::note_launch();
  return gngf::check_launch();
}

int gngf_hpd_small_fwd(gngf_lattice lat, int32_t n_layers, const int32_t* widths, const float* const* w,
                       const float* const* b, float* const* act, int32_t topk, float* uprobs, float* utopv,
                       int32_t* utopi, void* stream) {
  return gngf_hpd_small_fwd_enc(lat, n_layers, widths, w, b, act, topk, uprobs, utopv, utopi, gngf_tables{}, 0, 0, nullptr,
                                stream);
}

// gact: n_layers outputs (U, width[i+1]) receiving the pre-activation adjoints; dbias: n_layers bias gradients (+=);
// dw0 (width[1], 2) (+=).  The weight gradients of layers >= 1 are dW_i += gact[i]^T act[i-1] (gngf_linear_bwd).
int gngf_hpd_small_bwd_enc(gngf_lattice lat, int32_t n_layers, const int32_t* widths, const float* const* w,
                           float* const* act, float* const* gact, float* const* dbias, float* dw0, int32_t topk,
                           const float* uprobs, const int32_t* utopi, const float* dtv, const int32_t* cnt,
                           const float* gcol, const float* gcol_k, const float* gdense, gngf_tables tables,
                           gngf_tables table_grads, int32_t F, int32_t mix_mode, const float* dnf, void* stream) {
  if (!gngf_hpd_small_supported(n_layers, widths, topk)) return GNGF_ERR_UNSUPPORTED;
  const int64_t U = static_cast<int64_t>(lat.wx) * lat.wy;
  if (U <= 0) return GNGF_ERR_INVALID_ARGUMENT;
  if ((gcol || gcol_k) && !cnt) return GNGF_ERR_INVALID_ARGUMENT;
  if (!dtv && !dnf) return GNGF_ERR_INVALID_ARGUMENT;
  if (dnf && (F <= 0 || F > GNGF_MAX_FEATURES || lat.num_levels <= 0 || lat.num_levels > GNGF_MAX_LEVELS))
    return GNGF_ERR_INVALID_ARGUMENT;
  gngf::SmallEnc enc{};
  enc.tables = tables;
  enc.tgrads = table_grads;
  enc.F = F;
  enc.mode = mix_mode;
  enc.dnf = dnf;
  gngf::HpdNet net{};
  net.n_layers = n_layers;
  for (int i = 0; i <= n_layers; ++i) net.width[i] = widths[i];
  for (int i = 0; i < n_layers; ++i) {
    net.w[i] = w[i];
    net.act[i] = i < n_layers - 1 ? act[i] : nullptr;
    net.gact[i] = gact[i];
    net.dbias[i] = dbias[i];
  }
  net.dw0 = dw0;
  const int gt_rows = gngf::small_bwd_gt_rows(n_layers, widths);
  const size_t res = gngf::small_bwd_resident_smem(n_layers, widths, w, gt_rows);
  const size_t smem = res ? res : gngf::small_bwd_smem(gt_rows);
  if (cudaFuncSetAttribute(gngf::hpd_small_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           static_cast<int>(smem)) != cudaSuccess)
    return gngf::check_launch();
  gngf::hpd_small_bwd_kernel<<<static_cast<unsigned>(gngf::ceil_div(U, gngf::NB)), gngf::SM_THREADS, smem,
                               gngf::as_stream(stream)>>>(lat, net, enc, topk, uprobs, utopi, dtv, cnt, gcol, gcol_k,
                                                          gdense, gt_rows, res ? 1 : 0);
  gngf::note_launch();
  return gngf::check_launch();
}

int gngf_hpd_small_bwd(gngf_lattice lat, int32_t n_layers, const int32_t* widths, const float* const* w,
                       float* const* act, float* const* gact, float* const* dbias, float* dw0, int32_t topk,
                       const float* uprobs, const int32_t* utopi, const float* dtv, const int32_t* cnt,
                       const float* gcol, const float* gcol_k, const float* gdense, void* stream) {
  if (!dtv) return GNGF_ERR_INVALID_ARGUMENT;
  return gngf_hpd_small_bwd_enc(lat, n_layers, widths, w, act, gact, dbias, dw0, topk, uprobs, utopi, dtv, cnt, gcol,
                                gcol_k, gdense, gngf_tables{}, gngf_tables{}, 0, 0, nullptr, stream);
}

}  // extern "C"
