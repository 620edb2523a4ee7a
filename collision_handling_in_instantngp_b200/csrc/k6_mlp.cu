// K6: the decoder MLP  enc (P, IN) -> 64 -> 64 -> OUT (3 or 1), ReLU/LeakyReLU hidden, sigmoid output
// (models.py:382-392, 468-470) fused into one forward kernel and one backward kernel for the reference's
// shape (two hidden layers of 64).  Activations never leave the SM: a CTA of 128 threads takes tiles of 128
// points, keeps them k-major in shared memory ([feature][point], row stride 132 floats so that float4 accesses of
// a quarter warp hit distinct banks) and runs each layer as a register-tiled fp32 GEMM (true fp32 FMA = the
// reference's cuBLAS SGEMM arithmetic).  The thread tile is 8 points x 8 outputs: 4 shared-memory vector loads feed
// 64 FMAs, which keeps the kernel FMA-bound rather than shared-memory-bound (a 4x4 tile needs 2 loads per 16
// FMAs and was limited by the one-wavefront-per-cycle shared-memory pipe).  The backward recomputes the hidden
// activations instead of reading them back; weight gradients are accumulated in registers across all tiles of a
// (persistent) CTA, written once per CTA to a partials buffer and summed by a second tiny kernel -- no atomics.
// Other decoder shapes go through the generic layers of k2_linear.cu.
#include <algorithm>

#include "common.cuh"

namespace gngf {

constexpr int H = 64;        // hidden width
constexpr int TP = 128;      // points per tile
constexpr int LDP = TP + 4;  // row stride of the k-major activation buffers
constexpr int MLP_THREADS = 128;

__device__ __forceinline__ float hidden_act(float v, int leaky) { return v > 0.0f ? v : (leaky ? v * 0.01f : 0.0f); }

// thread (tp = tid % 16, tj = tid / 16) owns points {4tp..4tp+3, 64+4tp..} x outputs {4tj..4tj+3, 32+4tj..}
// acc[pi][ji] += sum_k A[k][p] * B[k][j]
__device__ __forceinline__ void gemm_tile(const float* __restrict__ A, const float* __restrict__ B, int bstride, int K,
                                          int tp, int tj, float acc[8][8]) {
#pragma unroll 2
  for (int k = 0; k < K; ++k) {
    const float4 a0 = *reinterpret_cast<const float4*>(A + k * LDP + 4 * tp);
    const float4 a1 = *reinterpret_cast<const float4*>(A + k * LDP + 64 + 4 * tp);
    const float4 b0 = *reinterpret_cast<const float4*>(B + k * bstride + 4 * tj);
    const float4 b1 = *reinterpret_cast<const float4*>(B + k * bstride + 32 + 4 * tj);
    const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
  }
}
__device__ __forceinline__ int tile_point(int tp, int i) { return (i < 4 ? 0 : 64) + 4 * tp + (i & 3); }
__device__ __forceinline__ int tile_out(int tj, int j) { return (j < 4 ? 0 : 32) + 4 * tj + (j & 3); }

struct MlpSmem {
  float* X;    // [INP][LDP]   enc, k-major
  float* A1;   // [H][LDP]
  float* A2;   // [H][LDP]
  float* W0t;  // [INP][H]     W0^T  (forward B operand)
  float* W1t;  // [H][H]       W1^T
  float* W0n;  // [H][INP]     W0 as stored (backward B operand)      -- backward only
  float* W1n;  // [H][H]
  float* dz2;  // [4][LDP]                                            -- backward only
  // the output layer's weights and all biases (~650 floats) are read through the read-only path instead:
  // that keeps the backward at 110.8 KB of shared memory, i.e. two CTAs per SM
  const float* w2;  // (OUT, H) global
  const float* b0;
  const float* b1;
  const float* b2;
  int OUT;
};

__host__ __device__ inline int mlp_smem_floats(int INP, bool bwd) {
  int n = INP * LDP + 2 * H * LDP + INP * H + H * H;
  if (bwd) n += H * INP + H * H + 4 * LDP;
  return n;
}

__device__ __forceinline__ MlpSmem carve(float* base, int INP, bool bwd) {
  MlpSmem s;
  float* p = base;
  s.X = p; p += INP * LDP;
  s.A1 = p; p += H * LDP;
  s.A2 = p; p += H * LDP;
  s.W0t = p; p += INP * H;
  s.W1t = p; p += H * H;
  if (bwd) {
    s.W0n = p; p += H * INP;
    s.W1n = p; p += H * H;
    s.dz2 = p; p += 4 * LDP;
  } else {
    s.W0n = s.W1n = s.dz2 = nullptr;
  }
  s.w2 = s.b0 = s.b1 = s.b2 = nullptr;
  s.OUT = 0;
  return s;
}

__device__ __forceinline__ void load_weights(MlpSmem& s, int IN, int INP, int OUT, bool bwd,
                                             const float* __restrict__ w0, const float* __restrict__ b0,
                                             const float* __restrict__ w1, const float* __restrict__ b1,
                                             const float* __restrict__ w2, const float* __restrict__ b2) {
  const int tid = threadIdx.x;
  // (transposed copies are written with the shared-memory index fastest: conflict-free stores, the strided
  //  global reads come from L2)
  for (int e = tid; e < H * INP; e += MLP_THREADS) {  // w0 is (H, IN)
    const int k = e / H, j = e % H;
    s.W0t[e] = k < IN ? w0[j * IN + k] : 0.0f;
  }
  for (int e = tid; e < H * H; e += MLP_THREADS) {  // w1 is (H, H)
    const int k = e / H, j = e % H;
    s.W1t[e] = w1[j * H + k];
  }
  if (bwd) {
    for (int e = tid; e < H * INP; e += MLP_THREADS) {
      const int j = e / INP, k = e % INP;
      s.W0n[e] = k < IN ? w0[j * IN + k] : 0.0f;
    }
    for (int e = tid; e < H * H; e += MLP_THREADS) s.W1n[e] = w1[e];
  }
  s.w2 = w2;
  s.b0 = b0;
  s.b1 = b1;
  s.b2 = b2;
  s.OUT = OUT;
}

// enc tile -> X (k-major), zero-padded rows/points
__device__ __forceinline__ void load_enc_tile(const MlpSmem& s, const float* __restrict__ enc, int64_t p0, int64_t P,
                                              int IN, int INP) {
  for (int e = threadIdx.x; e < TP * INP; e += MLP_THREADS) {
    int p, k;
    if (IN == INP) {
      p = e / IN;
      k = e % IN;
    } else {
      p = e / INP;
      k = e % INP;
    }
    float v = 0.0f;
    if (p0 + p < P && k < IN) v = enc[(p0 + p) * IN + k];
    s.X[k * LDP + p] = v;
  }
}

// dst[j][p] = act(sum_k src[k][p] * Wt[k][j] + b[j])
__device__ __forceinline__ void hidden_layer(const float* src, const float* Wt, const float* b, int K, float* dst,
                                             int leaky) {
  const int tp = threadIdx.x % 16, tj = threadIdx.x / 16;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
  gemm_tile(src, Wt, H, K, tp, tj, acc);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int jj = tile_out(tj, j);
    const float bj = __ldg(b + jj);
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      float4 o;
      o.x = hidden_act(acc[4 * g + 0][j] + bj, leaky);
      o.y = hidden_act(acc[4 * g + 1][j] + bj, leaky);
      o.z = hidden_act(acc[4 * g + 2][j] + bj, leaky);
      o.w = hidden_act(acc[4 * g + 3][j] + bj, leaky);
      *reinterpret_cast<float4*>(dst + jj * LDP + 64 * g + 4 * tp) = o;
    }
  }
}

// thread p < TP: out[c] = sigmoid(sum_k A2[k][p] * w2[c][k] + b2[c]), c < OUT (others 0.5, never used)
__device__ __forceinline__ void output_layer(const MlpSmem& s, int p, float out[4]) {
  float acc[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) acc[c] = c < s.OUT ? __ldg(s.b2 + c) : 0.0f;
#pragma unroll 8
  for (int k = 0; k < H; ++k) {
    const float a = s.A2[k * LDP + p];
#pragma unroll
    for (int c = 0; c < 4; ++c)
      if (c < s.OUT) acc[c] = fmaf(a, __ldg(s.w2 + c * H + k), acc[c]);
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) out[c] = 1.0f / (1.0f + expf(-acc[c]));
}

__global__ void __launch_bounds__(MLP_THREADS)
    mlp3_fwd_kernel(const float* __restrict__ enc, int64_t P, int IN, int INP, int OUT, int leaky,
                    const float* __restrict__ w0, const float* __restrict__ b0, const float* __restrict__ w1,
                    const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
                    float* __restrict__ rgb) {
  extern __shared__ __align__(16) float smem_f[];
  MlpSmem s = carve(smem_f, INP, false);
  load_weights(s, IN, INP, OUT, false, w0, b0, w1, b1, w2, b2);
  const int64_t tiles = (P + TP - 1) / TP;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t p0 = tile * TP;
    __syncthreads();  // weights visible / previous tile done with X and A2
    load_enc_tile(s, enc, p0, P, IN, INP);
    __syncthreads();
    hidden_layer(s.X, s.W0t, s.b0, INP, s.A1, leaky);
    __syncthreads();
    hidden_layer(s.A1, s.W1t, s.b1, H, s.A2, leaky);
    __syncthreads();
    const int p = threadIdx.x;
    if (p < TP && p0 + p < P) {
      float out[4];
      output_layer(s, p, out);
      for (int c = 0; c < OUT; ++c) rgb[(p0 + p) * OUT + c] = out[c];
    }
  }
}

// partial-gradient layout per CTA (floats): dw0 [H*IN] | db0 [H] | dw1 [H*H] | db1 [H] | dw2 [OUT*H] | db2 [OUT]
__host__ __device__ inline int mlp_param_floats(int IN, int OUT) { return H * IN + H + H * H + H + OUT * H + OUT; }

template <int NM>  // NM = number of dw0 columns per thread = ceil(INP / 2) rounded to {4, 16, 32}
__global__ void __launch_bounds__(MLP_THREADS)
    mlp3_bwd_kernel(const float* __restrict__ enc, const float* __restrict__ drgb, int64_t P, int IN, int INP, int OUT,
                    int leaky, const float* __restrict__ w0, const float* __restrict__ b0,
                    const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ w2,
                    const float* __restrict__ b2, float* __restrict__ denc, float* __restrict__ partials) {
  extern __shared__ __align__(16) float smem_f[];
  MlpSmem s = carve(smem_f, INP, true);
  load_weights(s, IN, INP, OUT, true, w0, b0, w1, b1, w2, b2);
  const int tid = threadIdx.x;
  const int tp = tid % 16, tj = tid / 16;
  const float slope = leaky ? 0.01f : 0.0f;

  // gradient accumulators that live in registers across all tiles of this CTA
  float g_w1[8][4];   // dw1[j = tid % 8 + 8 a][k = tid / 8 + 16 b]
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) g_w1[a][b] = 0.0f;
  float g_w0[NM];     // dw0[j = tid % 64][i = tid / 64 + 2 m], m < INP / 2
#pragma unroll
  for (int m = 0; m < NM; ++m) g_w0[m] = 0.0f;
  float g_w2[2] = {0.0f, 0.0f};  // dw2[c = tid / 64 + 2 q][k = tid % 64], q < 2
  float g_b1 = 0.0f, g_b0 = 0.0f, g_b2 = 0.0f;  // tid < 64: db1[tid], db0[tid]; tid < 4: db2[tid]

  const int64_t tiles = (P + TP - 1) / TP;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t p0 = tile * TP;
    __syncthreads();
    load_enc_tile(s, enc, p0, P, IN, INP);
    __syncthreads();
    hidden_layer(s.X, s.W0t, s.b0, INP, s.A1, leaky);
    __syncthreads();
    hidden_layer(s.A1, s.W1t, s.b1, H, s.A2, leaky);
    __syncthreads();
    // dz2 = drgb * y * (1 - y), stored [c][p]; padded points / channels contribute zero  (thread = point)
    {
      float out[4];
      output_layer(s, tid, out);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float g = 0.0f;
        if (c < OUT && p0 + tid < P) g = drgb[(p0 + tid) * OUT + c] * out[c] * (1.0f - out[c]);
        s.dz2[c * LDP + tid] = g;
      }
    }
    __syncthreads();
    // dw2[c][k] += sum_p dz2[c][p] * a2[k][p] ; db2[c] += sum_p dz2[c][p]
    {
      const int k = tid % H;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int c = tid / H + 2 * q;
        float acc = 0.0f;
        for (int p = 0; p < TP; p += 4) {
          const float4 d = *reinterpret_cast<const float4*>(s.dz2 + c * LDP + p);
          const float4 a = *reinterpret_cast<const float4*>(s.A2 + k * LDP + p);
          acc = fmaf(d.x, a.x, acc); acc = fmaf(d.y, a.y, acc); acc = fmaf(d.z, a.z, acc); acc = fmaf(d.w, a.w, acc);
        }
        g_w2[q] += acc;
      }
      if (tid < 4) {
        float sb = 0.0f;
        for (int p = 0; p < TP; p += 4) {
          const float4 d = *reinterpret_cast<const float4*>(s.dz2 + tid * LDP + p);
          sb += (d.x + d.y) + (d.z + d.w);
        }
        g_b2 += sb;
      }
    }
    __syncthreads();
    // dz1[j][p] = (sum_c dz2[c][p] * w2[c][j]) * act'(a2[j][p]), in place over A2 (each thread its own 8x8 tile)
    {
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const int pp = g * 64 + 4 * tp;
        const float4 d0 = *reinterpret_cast<const float4*>(s.dz2 + 0 * LDP + pp);
        const float4 d1 = *reinterpret_cast<const float4*>(s.dz2 + 1 * LDP + pp);
        const float4 d2 = *reinterpret_cast<const float4*>(s.dz2 + 2 * LDP + pp);
        const float4 d3 = *reinterpret_cast<const float4*>(s.dz2 + 3 * LDP + pp);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int jj = tile_out(tj, j);
          const float w20 = __ldg(s.w2 + jj), w21 = OUT > 1 ? __ldg(s.w2 + H + jj) : 0.0f;
          const float w22 = OUT > 2 ? __ldg(s.w2 + 2 * H + jj) : 0.0f, w23 = OUT > 3 ? __ldg(s.w2 + 3 * H + jj) : 0.0f;
          const float4 a = *reinterpret_cast<const float4*>(s.A2 + jj * LDP + pp);
          float4 r;
          r.x = fmaf(d3.x, w23, fmaf(d2.x, w22, fmaf(d1.x, w21, d0.x * w20))) * (a.x > 0.0f ? 1.0f : slope);
          r.y = fmaf(d3.y, w23, fmaf(d2.y, w22, fmaf(d1.y, w21, d0.y * w20))) * (a.y > 0.0f ? 1.0f : slope);
          r.z = fmaf(d3.z, w23, fmaf(d2.z, w22, fmaf(d1.z, w21, d0.z * w20))) * (a.z > 0.0f ? 1.0f : slope);
          r.w = fmaf(d3.w, w23, fmaf(d2.w, w22, fmaf(d1.w, w21, d0.w * w20))) * (a.w > 0.0f ? 1.0f : slope);
          *reinterpret_cast<float4*>(s.A2 + jj * LDP + pp) = r;
        }
      }
    }
    __syncthreads();
    // dw1[j][k] += sum_p dz1[j][p] * a1[k][p] ; db1[j] += sum_p dz1[j][p]
    {
      const int rj = tid % 8, rk = tid / 8;
      float acc[8][4];
#pragma unroll
      for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0f;
      for (int p = 0; p < TP; p += 4) {
        float4 dz[8], av[4];
#pragma unroll
        for (int a = 0; a < 8; ++a) dz[a] = *reinterpret_cast<const float4*>(s.A2 + (rj + 8 * a) * LDP + p);
#pragma unroll
        for (int b = 0; b < 4; ++b) av[b] = *reinterpret_cast<const float4*>(s.A1 + (rk + 16 * b) * LDP + p);
#pragma unroll
        for (int a = 0; a < 8; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            acc[a][b] = fmaf(dz[a].x, av[b].x, acc[a][b]);
            acc[a][b] = fmaf(dz[a].y, av[b].y, acc[a][b]);
            acc[a][b] = fmaf(dz[a].z, av[b].z, acc[a][b]);
            acc[a][b] = fmaf(dz[a].w, av[b].w, acc[a][b]);
          }
      }
#pragma unroll
      for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) g_w1[a][b] += acc[a][b];
      if (tid < H) {
        float sb = 0.0f;
        for (int p = 0; p < TP; p += 4) {
          const float4 d = *reinterpret_cast<const float4*>(s.A2 + tid * LDP + p);
          sb += (d.x + d.y) + (d.z + d.w);
        }
        g_b1 += sb;
      }
    }
    // dz0[k][p] = (sum_j dz1[j][p] * w1[j][k]) * act'(a1[k][p]) -> registers, then in place over A1
    {
      float acc[8][8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
      gemm_tile(s.A2, s.W1n, H, H, tp, tj, acc);
      __syncthreads();  // every thread is done reading A1 (dw1) before it is overwritten
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int kk = tile_out(tj, j);
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          float* ptr = s.A1 + kk * LDP + g * 64 + 4 * tp;
          const float4 a = *reinterpret_cast<const float4*>(ptr);
          float4 r;
          r.x = acc[g * 4 + 0][j] * (a.x > 0.0f ? 1.0f : slope);
          r.y = acc[g * 4 + 1][j] * (a.y > 0.0f ? 1.0f : slope);
          r.z = acc[g * 4 + 2][j] * (a.z > 0.0f ? 1.0f : slope);
          r.w = acc[g * 4 + 3][j] * (a.w > 0.0f ? 1.0f : slope);
          *reinterpret_cast<float4*>(ptr) = r;
        }
      }
    }
    __syncthreads();
    // dw0[j][i] += sum_p dz0[j][p] * x[i][p] ; db0[j] += sum_p dz0[j][p]
    {
      const int j = tid % H, i0 = tid / H;  // i = i0 + 2 m
      const int nm = INP / 2;
      for (int p = 0; p < TP; p += 4) {
        const float4 d = *reinterpret_cast<const float4*>(s.A1 + j * LDP + p);
#pragma unroll
        for (int m = 0; m < NM; ++m) {
          if (m < nm) {
            const float4 xv = *reinterpret_cast<const float4*>(s.X + (i0 + 2 * m) * LDP + p);
            g_w0[m] = fmaf(d.x, xv.x, fmaf(d.y, xv.y, fmaf(d.z, xv.z, fmaf(d.w, xv.w, g_w0[m]))));
          }
        }
      }
      if (tid < H) {
        float sb = 0.0f;
        for (int p = 0; p < TP; p += 4) {
          const float4 d = *reinterpret_cast<const float4*>(s.A1 + tid * LDP + p);
          sb += (d.x + d.y) + (d.z + d.w);
        }
        g_b0 += sb;
      }
    }
    // denc[p][i] = sum_k dz0[k][p] * w0[k][i]   (INP / 4 column groups of 4; thread -> 4 points x 4 columns)
    {
      const int groups = INP / 4;                       // column groups
      for (int e = tid; e < (TP / 4) * groups; e += MLP_THREADS) {
        const int pg = e % (TP / 4), cg = e / (TP / 4);
        float acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) acc[a][b] = 0.0f;
#pragma unroll 4
        for (int k = 0; k < H; ++k) {
          const float4 d = *reinterpret_cast<const float4*>(s.A1 + k * LDP + 4 * pg);
          const float4 w = *reinterpret_cast<const float4*>(s.W0n + k * INP + 4 * cg);
          const float dv[4] = {d.x, d.y, d.z, d.w}, wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(dv[a], wv[b], acc[a][b]);
        }
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const int64_t p = p0 + 4 * pg + a;
          if (p < P) {
#pragma unroll
            for (int b = 0; b < 4; ++b)
              if (4 * cg + b < IN) denc[p * IN + 4 * cg + b] = acc[a][b];
          }
        }
      }
    }
  }

  // one partial-gradient record per CTA
  float* out = partials + static_cast<int64_t>(blockIdx.x) * mlp_param_floats(IN, OUT);
  float* o_dw0 = out;
  float* o_db0 = o_dw0 + H * IN;
  float* o_dw1 = o_db0 + H;
  float* o_db1 = o_dw1 + H * H;
  float* o_dw2 = o_db1 + H;
  float* o_db2 = o_dw2 + OUT * H;
  {
    const int j = tid % H, i0 = tid / H;
#pragma unroll
    for (int m = 0; m < NM; ++m) {
      const int i = i0 + 2 * m;
      if (m < INP / 2 && i < IN) o_dw0[j * IN + i] = g_w0[m];
    }
  }
  {
    const int rj = tid % 8, rk = tid / 8;
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) o_dw1[(rj + 8 * a) * H + (rk + 16 * b)] = g_w1[a][b];
  }
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int c = tid / H + 2 * q;
    if (c < OUT) o_dw2[c * H + tid % H] = g_w2[q];
  }
  if (tid < H) {
    o_db1[tid] = g_b1;
    o_db0[tid] = g_b0;
  }
  if (tid < OUT) o_db2[tid] = g_b2;
}

// grads[i] += sum over CTAs of partials[cta][i], routed to the six parameter-gradient buffers
__global__ void __launch_bounds__(256)
    mlp3_reduce_kernel(const float* __restrict__ partials, int n_cta, int IN, int OUT, float* __restrict__ dw0,
                       float* __restrict__ db0, float* __restrict__ dw1, float* __restrict__ db1,
                       float* __restrict__ dw2, float* __restrict__ db2) {
  const int n = mlp_param_floats(IN, OUT);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.0f;
  for (int c = 0; c < n_cta; ++c) s += partials[static_cast<int64_t>(c) * n + i];
  int o = i;
  if (o < H * IN) { dw0[o] += s; return; }
  o -= H * IN;
  if (o < H) { db0[o] += s; return; }
  o -= H;
  if (o < H * H) { dw1[o] += s; return; }
  o -= H * H;
  if (o < H) { db1[o] += s; return; }
  o -= H;
  if (o < OUT * H) { dw2[o] += s; return; }
  o -= OUT * H;
  db2[o] += s;
}

static int mlp_grid(int64_t P, size_t smem_bytes) {
  const int per_sm = static_cast<int>(std::max<size_t>(1, std::min<size_t>(4, (227 * 1024) / (smem_bytes + 1024))));
  return static_cast<int>(std::min<int64_t>(ceil_div(P, TP), static_cast<int64_t>(per_sm) * sm_count()));
}

}  // namespace gngf

extern "C" {

int gngf_mlp3_supported(int32_t in_dim, int32_t h1, int32_t h2, int32_t out_dim) {
  return in_dim >= 1 && in_dim <= 64 && h1 == gngf::H && h2 == gngf::H && out_dim >= 1 && out_dim <= 4;
}

int gngf_mlp3_fwd(const float* enc, int64_t P, int32_t in_dim, int32_t out_dim, int32_t leaky, const float* w0,
                  const float* b0, const float* w1, const float* b1, const float* w2, const float* b2, float* rgb,
                  void* stream) {
  if (!gngf_mlp3_supported(in_dim, gngf::H, gngf::H, out_dim) || P < 0) return GNGF_ERR_INVALID_ARGUMENT;
  if (P == 0) return GNGF_OK;
  const int INP = (in_dim + 3) & ~3;
  const size_t smem = sizeof(float) * gngf::mlp_smem_floats(INP, false);
  if (cudaFuncSetAttribute(gngf::mlp3_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           static_cast<int>(smem)) != cudaSuccess)
    return gngf::check_launch();
  gngf::mlp3_fwd_kernel<<<gngf::mlp_grid(P, smem), gngf::MLP_THREADS, smem, gngf::as_stream(stream)>>>(
      enc, P, in_dim, INP, out_dim, leaky, w0, b0, w1, b1, w2, b2, rgb);
  gngf::note_launch();
  return gngf::check_launch();
}

int64_t gngf_mlp3_bwd_workspace_floats(int32_t in_dim, int32_t out_dim) {
  return static_cast<int64_t>(4 * gngf::sm_count()) * gngf::mlp_param_floats(in_dim, out_dim);
}

int gngf_mlp3_bwd(const float* enc, const float* drgb, int64_t P, int32_t in_dim, int32_t out_dim, int32_t leaky,
                  const float* w0, const float* b0, const float* w1, const float* b1, const float* w2, const float* b2,
                  float* denc, float* dw0, float* db0, float* dw1, float* db1, float* dw2, float* db2, float* workspace,
                  void* stream) {
  if (!gngf_mlp3_supported(in_dim, gngf::H, gngf::H, out_dim) || P < 0) return GNGF_ERR_INVALID_ARGUMENT;
  if (P == 0) return GNGF_OK;
  const int INP = (in_dim + 3) & ~3;
  const size_t smem = sizeof(float) * gngf::mlp_smem_floats(INP, true);
  cudaStream_t st = gngf::as_stream(stream);
  const int grid = gngf::mlp_grid(P, smem);
  auto launch = [&](auto kernel) -> int {
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess)
      return gngf::check_launch();
    kernel<<<grid, gngf::MLP_THREADS, smem, st>>>(enc, drgb, P, in_dim, INP, out_dim, leaky, w0, b0, w1, b1, w2, b2,
                                                   denc, workspace);
    return GNGF_OK;
  };
  int lrc = INP <= 8 ? launch(gngf::mlp3_bwd_kernel<4>)
                     : (INP <= 32 ? launch(gngf::mlp3_bwd_kernel<16>) : launch(gngf::mlp3_bwd_kernel<32>));
  if (lrc) return lrc;
  gngf::note_launch();
  int rc = gngf::check_launch();
  if (rc) return rc;
  const int n = gngf::mlp_param_floats(in_dim, out_dim);
  gngf::mlp3_reduce_kernel<<<(n + 255) / 256, 256, 0, st>>>(workspace, grid, in_dim, out_dim, dw0, db0, dw1, db1, dw2,
                                                            db2);
  gngf::note_launch();
  return gngf::check_launch();
}

}  // extern "C"
