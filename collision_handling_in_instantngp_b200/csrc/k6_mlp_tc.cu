// K6 on the tensor cores: the decoder MLP  enc (P, IN) -> 64 -> 64 -> OUT  (models.py:382-392, 468-470), forward and
// backward, as chains of tcgen05 products whose operands never leave the SM.
//
// A CTA of 256 threads takes tiles of 128 points.  Thread = (point, 32-column half of a layer's output): it reads its
// accumulator row from TMEM, applies bias + activation, splits the fp32 result into bf16 planes and writes them into
// shared memory in the canonical K-major / 128-byte-swizzle layout -- which is exactly the A operand of the next
// layer's product.  Weights are staged once per CTA as bf16 planes in the same layout.
//
//   forward : three planes (hi, mid, lo), six products per k-step -> ~1.5e-6 relative (bar: 1e-5)
//   backward: two planes (hi, mid), three products                -> ~1e-5 relative   (bar: 1e-4)
//
// The backward needs every tile in two orientations (dX = dZ W contracts over features, dW = dZ^T A over points).
// Nothing is transposed: the same shared-memory tile is read K-major by one product and MN-major by the other
// (UMMA descriptor major bits; see k2_hpd_tc_bwd.cu), and weight tiles stored (out, in) serve as the K-major B operand
// of the forward layer and the MN-major B operand of the dX product.  Weight gradients accumulate in TMEM (M = 64
// accumulators, 16 lanes per subpartition) across all tiles of a persistent CTA and are reduced into the gradient
// buffers once per CTA; bias gradients are column sums taken with a 31-shuffle transpose-reduce per warp.
#include <algorithm>

#include "tc_common.cuh"

namespace gngf {
namespace tc {
namespace dec {

constexpr int H = 64;                       // hidden width
constexpr int TP = 128;                     // points per tile (= M of the forward products)
constexpr int THREADS = 256;
constexpr uint32_t ACT_PLANE = 128 * 128;   // 128 rows x 64 bf16
constexpr uint32_t W_PLANE = 64 * 128;      // 64 rows x 64 bf16
constexpr uint32_t W2_PLANE = 16 * 128;     // 16 rows x 64 bf16

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ float lo_f(uint32_t p) { return __uint_as_float(p << 16); }
__device__ __forceinline__ float hi_f(uint32_t p) { return __uint_as_float(p & 0xffff0000u); }

// byte offset of 16-byte chunk `ch` (8 bf16 columns) of row r inside a K-major, 128-byte-swizzled plane
__device__ __forceinline__ uint32_t chunk_off(int r, int ch) {
  return static_cast<uint32_t>((r >> 3) * 1024 + (r & 7) * 128 + ((ch ^ (r & 7)) << 4));
}

// 8 fp32 values -> NPL bf16 planes, one 16-byte chunk each
template <int NPL>
__device__ __forceinline__ void store_chunk(uint8_t* planes, uint32_t plane_bytes, uint32_t off, const float (&x)[8]) {
  uint32_t p0[4], p1[4], p2[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float a = x[2 * i], b = x[2 * i + 1];
    p0[i] = pack2(a, b);
    const float ra = a - lo_f(p0[i]), rb = b - hi_f(p0[i]);
    p1[i] = pack2(ra, rb);
    if (NPL > 2) p2[i] = pack2(ra - lo_f(p1[i]), rb - hi_f(p1[i]));
  }
  *reinterpret_cast<uint4*>(planes + off) = make_uint4(p0[0], p0[1], p0[2], p0[3]);
  *reinterpret_cast<uint4*>(planes + plane_bytes + off) = make_uint4(p1[0], p1[1], p1[2], p1[3]);
  if (NPL > 2) *reinterpret_cast<uint4*>(planes + 2 * plane_bytes + off) = make_uint4(p2[0], p2[1], p2[2], p2[3]);
}

// w (rows, cols) fp32 row-major -> planes of a (rows_pad x 64) tile, zero padded
template <int NPL>
__device__ __forceinline__ void stage_weight(const float* __restrict__ w, int rows, int cols, int rows_pad,
                                             uint8_t* planes, uint32_t plane_bytes, int nthreads = THREADS) {
  for (int e = threadIdx.x; e < rows_pad * 8; e += nthreads) {
    const int r = e >> 3, ch = e & 7;
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = ch * 8 + i;
      x[i] = (r < rows && c < cols) ? __ldg(w + r * cols + c) : 0.0f;
    }
    store_chunk<NPL>(planes, plane_bytes, chunk_off(r, ch), x);
  }
}

// enc rows of the tile -> planes (columns >= IN and rows >= P are zero); thread = (row, half) takes chunks half, half+2, ..
template <int NPL>
__device__ __forceinline__ void stage_x(const float* __restrict__ enc, int64_t p0, int64_t P, int IN, int nchunks, int row,
                                        int half, uint8_t* planes) {
  const int64_t p = p0 + row;
  for (int ch = half; ch < nchunks; ch += 2) {
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = 0.0f;
    if (p < P) {
      const float* src = enc + p * IN + ch * 8;
      if ((IN & 3) == 0 && ch * 8 + 8 <= IN) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(src));
        const float4 b = __ldg(reinterpret_cast<const float4*>(src) + 1);
        x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (ch * 8 + i < IN) x[i] = __ldg(src + i);
      }
    }
    store_chunk<NPL>(planes, ACT_PLANE, chunk_off(row, ch), x);
  }
}

// one product  D (+)= sum over plane pairs and k-steps of A B^T  issued by an elected lane of a converged warp.
// a_step / b_step: descriptor-lo increment per k-step (K-major: 32 bytes >> 4 = 2; MN-major: 2048 >> 4 = 128)
template <int NPL>
__device__ __forceinline__ void issue_product(uint32_t d, uint32_t a_lo, uint32_t a_plane16, uint32_t a_step, uint32_t b_lo,
                                              uint32_t b_plane16, uint32_t b_step, int ksteps, uint32_t idesc,
                                              bool accumulate) {
  constexpr int NPROD = NPL == 3 ? 6 : 3;
  // partial products of order <= NPL - 1
  constexpr int pa[6] = {0, 0, 1, 0, 2, 1};
  constexpr int pb[6] = {0, 1, 0, 2, 0, 1};
#pragma unroll
  for (int pr = 0; pr < NPROD; ++pr) {
    for (int k = 0; k < ksteps; ++k) {
      const uint64_t ad = umma_desc_pack(a_lo + pa[pr] * a_plane16 + k * a_step);
      const uint64_t bd = umma_desc_pack(b_lo + pb[pr] * b_plane16 + k * b_step);
      umma_bf16_lead(d, ad, bd, idesc, accumulate || (pr | k) != 0);
    }
  }
}

__device__ __forceinline__ float hidden_act(float v, int leaky) { return v > 0.0f ? v : (leaky ? v * 0.01f : 0.0f); }

__host__ __device__ constexpr uint32_t idesc_of(int M, int N, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (a_mn ? UMMA_A_MN_MAJOR : 0u) | (b_mn ? UMMA_B_MN_MAJOR : 0u) |
         (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}
__device__ __forceinline__ uint32_t idesc_rt(int M, int N, bool a_mn, bool b_mn) { return idesc_of(M, N, a_mn, b_mn); }

// ------------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------------
constexpr uint32_t FWD_SMEM = 3 * ACT_PLANE + 2 * 3 * W_PLANE + 3 * W2_PLANE + 1024 /*align*/ + 1024 /*bias, barrier*/;
constexpr uint32_t FWD_TMEM_COLS = 128;   // hidden accumulator [0,64), output accumulator [64,80)

__global__ void __launch_bounds__(THREADS, 2)
    mlp3_tc_fwd_kernel(const float* __restrict__ enc, int64_t P, int IN, int OUT, int leaky, const float* __restrict__ w0,
                       const float* __restrict__ b0, const float* __restrict__ w1, const float* __restrict__ b1,
                       const float* __restrict__ w2, const float* __restrict__ b2, float* __restrict__ rgb,
                       uint32_t* __restrict__ masks) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* act = smem;
  uint8_t* w0p = act + 3 * ACT_PLANE;
  uint8_t* w1p = w0p + 3 * W_PLANE;
  uint8_t* w2p = w1p + 3 * W_PLANE;
  float* sb = reinterpret_cast<float*>(w2p + 3 * W2_PLANE);   // b0 [64] | b1 [64] | b2 [16]
  uint64_t* bar = reinterpret_cast<uint64_t*>(sb + 160);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int q = warp & 3, half = warp >> 2;
  const int row = q * 32 + lane;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(FWD_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  stage_weight<3>(w0, H, IN, H, w0p, W_PLANE);
  stage_weight<3>(w1, H, H, H, w1p, W_PLANE);
  stage_weight<3>(w2, OUT, H, 16, w2p, W2_PLANE);
  for (int i = threadIdx.x; i < 144; i += THREADS)
    sb[i] = i < 64 ? __ldg(b0 + i) : (i < 128 ? __ldg(b1 + i - 64) : (i - 128 < OUT ? __ldg(b2 + i - 128) : 0.0f));
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
  const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;

  const int ksteps0 = (IN + 15) / 16, nchunks = ksteps0 * 2;
  const uint32_t act_lo = umma_desc_lo(smem_u32(act)), w0_lo = umma_desc_lo(smem_u32(w0p)),
                 w1_lo = umma_desc_lo(smem_u32(w1p)), w2_lo = umma_desc_lo(smem_u32(w2p));
  uint32_t phase = 0;
  const int64_t tiles = (P + TP - 1) / TP;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t p0 = tile * TP;
    stage_x<3>(enc, p0, P, IN, nchunks, row, half, act);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
      tc_fence_after();
      issue_product<3>(tmem_u, act_lo, ACT_PLANE >> 4, 2, w0_lo, W_PLANE >> 4, 2, ksteps0, idesc_of(128, 64, false, false),
                       false);
      umma_commit_lead(bar);
    }
#pragma unroll 1
    for (int layer = 0; layer < 2; ++layer) {
      mbar_wait(bar, phase);
      phase ^= 1;
      tc_fence_after();
      uint32_t v[32];
      tmem_ld32(tmem_base + lane_off + half * 32, v);
      const float* bias = sb + layer * 64 + half * 32;
      uint32_t m = 0;   // bit j: pre-activation of column half*32 + j is positive (what the backward gates with)
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        float x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float z = __uint_as_float(v[ch * 8 + i]) + bias[ch * 8 + i];
          m |= (z > 0.0f ? 1u : 0u) << (ch * 8 + i);
          x[i] = hidden_act(z, leaky);
        }
        store_chunk<3>(act, ACT_PLANE, chunk_off(row, half * 4 + ch), x);
      }
      if (masks && p0 + row < P) masks[(p0 + row) * 4 + layer * 2 + half] = m;
      fence_proxy_async();
      tc_fence_before();
      __syncthreads();
      if (warp == 0) {
        tc_fence_after();
        if (layer == 0)
          issue_product<3>(tmem_u, act_lo, ACT_PLANE >> 4, 2, w1_lo, W_PLANE >> 4, 2, 4, idesc_of(128, 64, false, false),
                           false);
        else
          issue_product<3>(tmem_u + 64, act_lo, ACT_PLANE >> 4, 2, w2_lo, W2_PLANE >> 4, 2, 4,
                           idesc_of(128, 16, false, false), false);
        umma_commit_lead(bar);
      }
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();
    if (half == 0) {
      uint32_t v[16];
      tmem_ld16(tmem_base + lane_off + 64, v);
      if (p0 + row < P) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (c < OUT) rgb[(p0 + row) * OUT + c] = 1.0f / (1.0f + expf(-(__uint_as_float(v[c]) + sb[128 + c])));
      }
    }
    tc_fence_before();   // the next tile's first product overwrites the accumulators
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(FWD_TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------------------------
// shared memory: XB | A1B | A2B | G | G2 (2 planes x 16 KB each) | w0p | w1p (2 x 8 KB) | w2p (2 x 2 KB) | biases | barriers
constexpr uint32_t BWD_BUF = 2 * ACT_PLANE;
constexpr uint32_t BWD_SMEM = 5 * BWD_BUF + 2 * 2 * W_PLANE + 2 * W2_PLANE + 1024 /*align*/ + 1024;
constexpr uint32_t BWD_TMEM_COLS = 512;
constexpr int BWD_THREADS = THREADS + 32;   // 8 epilogue warps + the MMA issuer
// TMEM columns: main accumulator (Z0, Z1, dA2, dA1) | dX | dW1 | dW0 | dW2
constexpr uint32_t C_MAIN = 0, C_DX = 64, C_DW1 = 128, C_DW0 = 192, C_DW2 = 256;

// column sums over the 32 rows held by a warp: d[j] (thread = row, j = column) -> lane j returns the sum of column j
__device__ __forceinline__ float warp_colsum32(float (&d)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float send = up ? d[i] : d[i + s];
      const float keep = up ? d[i + s] : d[i];
      d[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return d[0];
}

// 288 threads: warps 0..7 = epilogue (thread = (point, 32-column half)), warp 8 = MMA issuer.  The issuer waits on `ready`
// (one arrival per epilogue warp, after its TMEM reads and shared-memory writes) instead of a __syncthreads, issues the
// product the epilogue warps are waiting for, commits it, and only then issues the weight-gradient product of the same
// stage -- which therefore costs the epilogue nothing (clock64 instrumentation of the first version: ~117 MMAs per tile
// at ~35 clocks of issue each = a quarter of the tile time, all on the critical path through warp 0).
__global__ void __launch_bounds__(BWD_THREADS, 1)
    mlp3_tc_bwd_kernel(const float* __restrict__ enc, const float* __restrict__ rgb, const float* __restrict__ drgb,
                       int64_t P, int IN, int OUT, int leaky, const float* __restrict__ w0, const float* __restrict__ b0,
                       const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ w2,
                       const uint32_t* __restrict__ masks, float* __restrict__ denc, float* __restrict__ dw0, float* __restrict__ db0,
                       float* __restrict__ dw1, float* __restrict__ db1, float* __restrict__ dw2,
                       float* __restrict__ db2) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* XB = smem;
  uint8_t* A1B = XB + BWD_BUF;
  uint8_t* A2B = A1B + BWD_BUF;
  uint8_t* G = A2B + BWD_BUF;
  uint8_t* G2 = G + BWD_BUF;
  uint8_t* w0p = G2 + BWD_BUF;
  uint8_t* w1p = w0p + 2 * W_PLANE;
  uint8_t* w2p = w1p + 2 * W_PLANE;
  float* sb = reinterpret_cast<float*>(w2p + 2 * W2_PLANE);   // b0 [64] | b1 [64]
  uint64_t* bar_main = reinterpret_cast<uint64_t*>(sb + 128);
  uint64_t* bar_tile = bar_main + 1;
  uint64_t* bar_ready = bar_tile + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_ready + 1);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int q = warp & 3, half = warp >> 2;
  const int row = q * 32 + lane;
  if (threadIdx.x == 0) {
    mbar_init(bar_main, 1);
    mbar_init(bar_tile, 1);
    mbar_init(bar_ready, 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(BWD_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  stage_weight<2>(w0, H, IN, H, w0p, W_PLANE, BWD_THREADS);
  stage_weight<2>(w1, H, H, H, w1p, W_PLANE, BWD_THREADS);
  stage_weight<2>(w2, OUT, H, 16, w2p, W2_PLANE, BWD_THREADS);
  for (int i = threadIdx.x; i < 128; i += BWD_THREADS) sb[i] = i < 64 ? __ldg(b0 + i) : __ldg(b1 + i - 64);
  // G holds dz2 in its first 16 columns (two chunks); chunk 1 (channels 8..15) stays zero for the whole kernel
  if (warp < 8 && half == 1) {
    const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    store_chunk<2>(G, ACT_PLANE, chunk_off(row, 1), z);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
  const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;

  const int ksteps0 = (IN + 15) / 16, nchunks = ksteps0 * 2, INP = ksteps0 * 16;
  const uint32_t xb_lo = umma_desc_lo(smem_u32(XB)), a1_lo = umma_desc_lo(smem_u32(A1B)),
                 a2_lo = umma_desc_lo(smem_u32(A2B)), g_lo = umma_desc_lo(smem_u32(G)), g2_lo = umma_desc_lo(smem_u32(G2)),
                 w0_lo = umma_desc_lo(smem_u32(w0p)), w1_lo = umma_desc_lo(smem_u32(w1p)),
                 w2_lo = umma_desc_lo(smem_u32(w2p));
  constexpr uint32_t AP = ACT_PLANE >> 4, WP = W_PLANE >> 4, W2P = W2_PLANE >> 4;
  constexpr uint32_t KS = 2, MS = 128;   // descriptor step per k-step: K-major (32 B) / MN-major (16 rows = 2048 B)
  const float slope = leaky ? 0.01f : 0.0f;

  uint32_t ph_main = 0, ph_tile = 0;
  float db1_acc = 0.0f, db0_acc = 0.0f;   // lane j of warp (q, half): column half*32 + j, rows of this warp, all tiles
  float db2_acc[4] = {0.f, 0.f, 0.f, 0.f};
  bool first = true;
  const int64_t tiles = (P + TP - 1) / TP;
  if (warp == 8) {
    // ---- MMA issuer ----
    uint32_t ph_ready = 0;
    bool first_i = true;
    auto wait_ready = [&]() {
      mbar_wait(bar_ready, ph_ready);
      ph_ready ^= 1;
      tc_fence_after();
    };
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      wait_ready();   // X, dz2 staged: Z0 = X W0^T
      issue_product<2>(tmem_u + C_MAIN, xb_lo, AP, KS, w0_lo, WP, KS, ksteps0, idesc_of(128, 64, false, false), false);
      umma_commit_lead(bar_main);
      wait_ready();   // A1 staged: Z1 = A1 W1^T
      issue_product<2>(tmem_u + C_MAIN, a1_lo, AP, KS, w1_lo, WP, KS, 4, idesc_of(128, 64, false, false), false);
      umma_commit_lead(bar_main);
      wait_ready();   // A2 staged: dA2 (128 x 64) = dz2 (K = 16 channels) W2, B = W2 planes (16 x 64) read MN-major
      issue_product<2>(tmem_u + C_MAIN, g_lo, AP, KS, w2_lo, W2P, MS, 1, idesc_of(128, 64, false, true), false);
      umma_commit_lead(bar_main);
      //                dW2^T (64 x 16) += A2^T dz2 : both operands MN-major, K = 128 points
      issue_product<2>(tmem_u + C_DW2, a2_lo, AP, MS, g_lo, AP, MS, 8, idesc_of(64, 16, true, true), !first_i);
      wait_ready();   // dZ1 staged: dA1 = dZ1 W1 (B = W1 planes (out, in) read MN-major); dW1 (64 x 64) += dZ1^T A1
      issue_product<2>(tmem_u + C_MAIN, g2_lo, AP, KS, w1_lo, WP, MS, 4, idesc_of(128, 64, false, true), false);
      umma_commit_lead(bar_main);
      issue_product<2>(tmem_u + C_DW1, g2_lo, AP, MS, a1_lo, AP, MS, 8, idesc_of(64, 64, true, true), !first_i);
      wait_ready();   // dZ0 staged: dX (128 x INP) = dZ0 W0 (B = W0 planes read MN-major); dW0 (64 x INP) += dZ0^T X
      issue_product<2>(tmem_u + C_DX, a2_lo, AP, KS, w0_lo, WP, MS, 4, idesc_rt(128, INP, false, true), false);
      umma_commit_lead(bar_main);
      issue_product<2>(tmem_u + C_DW0, a2_lo, AP, MS, xb_lo, AP, MS, 8, idesc_rt(64, INP, true, true), !first_i);
      umma_commit_lead(bar_tile);
      first_i = false;
    }
  }
  if (warp < 8) {
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t p0 = tile * TP;
    const bool live = p0 + row < P;
    // dz2 = drgb y (1 - y) for this thread's point (half 0 threads own G's chunk 0)
    float g[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (half == 0 && live) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (c < OUT) {
          const float y = __ldg(rgb + (p0 + row) * OUT + c);
          g[c] = __ldg(drgb + (p0 + row) * OUT + c) * y * (1.0f - y);
        }
      }
    }
    if (!first) {   // the previous tile's weight-gradient products have read XB, A1B, A2B, G, G2
      mbar_wait(bar_tile, ph_tile);
      ph_tile ^= 1;
    }
    stage_x<2>(enc, p0, P, IN, nchunks, row, half, XB);
    if (half == 0) {
      store_chunk<2>(G, ACT_PLANE, chunk_off(row, 0), g);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float s = warp_sum(g[c]);
        db2_acc[c] += s;
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_ready);
    // bit j: the FORWARD's pre-activation of column half*32 + j was positive.  (The two-plane recompute below is good
    // to ~1e-5; deriving the gates from it would flip ~1e-5 of them against the forward that produced the loss.)
    uint32_t mask1 = 0, mask2 = 0;
    if (live) {
      mask1 = __ldg(masks + (p0 + row) * 4 + half);
      mask2 = __ldg(masks + (p0 + row) * 4 + 2 + half);
    }
    // ---- recompute the hidden activations: A1 -> A1B, A2 -> A2B
#pragma unroll 1
    for (int layer = 0; layer < 2; ++layer) {
      mbar_wait(bar_main, ph_main);
      ph_main ^= 1;
      tc_fence_after();
      uint32_t v[32];
      tmem_ld32(tmem_base + lane_off + C_MAIN + half * 32, v);
      const float* bias = sb + layer * 64 + half * 32;
      uint8_t* dst = layer == 0 ? A1B : A2B;
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        float x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = hidden_act(__uint_as_float(v[ch * 8 + i]) + bias[ch * 8 + i], leaky);
        store_chunk<2>(dst, ACT_PLANE, chunk_off(row, half * 4 + ch), x);
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_ready);
    }
    // ---- dZ1 = dA2 .* act'(Z1) -> G2 ; db1
    {
      mbar_wait(bar_main, ph_main);
      ph_main ^= 1;
      tc_fence_after();
      uint32_t v[32];
      tmem_ld32(tmem_base + lane_off + C_MAIN + half * 32, v);
      float d[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) d[j] = __uint_as_float(v[j]) * ((mask2 >> j) & 1u ? 1.0f : slope);
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        float x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = d[ch * 8 + i];
        store_chunk<2>(G2, ACT_PLANE, chunk_off(row, half * 4 + ch), x);
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_ready);
      db1_acc += warp_colsum32(d, lane);
    }
    // ---- dZ0 = dA1 .* act'(Z0) -> A2B ; db0
    {
      mbar_wait(bar_main, ph_main);   // (also: the dW2 product, issued earlier, has finished reading A2B)
      ph_main ^= 1;
      tc_fence_after();
      uint32_t v[32];
      tmem_ld32(tmem_base + lane_off + C_MAIN + half * 32, v);
      float d[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) d[j] = __uint_as_float(v[j]) * ((mask1 >> j) & 1u ? 1.0f : slope);
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        float x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = d[ch * 8 + i];
        store_chunk<2>(A2B, ACT_PLANE, chunk_off(row, half * 4 + ch), x);
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_ready);
      db0_acc += warp_colsum32(d, lane);
    }
    // ---- dX -> global
    {
      mbar_wait(bar_main, ph_main);
      ph_main ^= 1;
      tc_fence_after();
      if (half * 32 < INP) {   // warp-uniform
        uint32_t v[32];
        tmem_ld32(tmem_base + lane_off + C_DX + half * 32, v);
        if (live) {
          float* o = denc + (p0 + row) * IN + half * 32;
          if ((IN & 3) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              if (half * 32 + j < IN)
                *reinterpret_cast<float4*>(o + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (half * 32 + j < IN) o[j] = __uint_as_float(v[j]);
          }
        }
      }
      tc_fence_before();
    }
    first = false;
  }
  }
  // ---- weight gradients: M = 64 accumulators live in lanes 0..15 of every 32-lane subpartition (row = 16 q + lane)
  if (!first && warp < 8) {
    mbar_wait(bar_tile, ph_tile);
    tc_fence_after();
    const int m = q * 16 + lane;
    {
      uint32_t v[32];
      tmem_ld32(tmem_base + lane_off + C_DW1 + half * 32, v);
      if (lane < 16) {
        float* o = dw1 + m * H + half * 32;
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          red_add_v4(o + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                     __uint_as_float(v[j + 3]));
      }
    }
    if (half * 32 < INP) {
      uint32_t v[32];
      tmem_ld32(tmem_base + lane_off + C_DW0 + half * 32, v);
      if (lane < 16) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (half * 32 + j < IN) atomicAdd(dw0 + m * IN + half * 32 + j, __uint_as_float(v[j]));
      }
    }
    if (half == 0) {
      uint32_t v[16];
      tmem_ld16(tmem_base + lane_off + C_DW2, v);
      if (lane < 16) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (c < OUT) atomicAdd(dw2 + c * H + m, __uint_as_float(v[c]));
      }
      if (lane == 0) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (c < OUT) atomicAdd(db2 + c, db2_acc[c]);
      }
    }
    atomicAdd(db1 + half * 32 + lane, db1_acc);
    atomicAdd(db0 + half * 32 + lane, db0_acc);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(BWD_TMEM_COLS) : "memory");
  }
}

}  // namespace dec
}  // namespace tc
}  // namespace gngf

extern "C" {

int gngf_mlp3_tc_supported(int32_t in_dim, int32_t h1, int32_t h2, int32_t out_dim) {
  return in_dim >= 1 && in_dim <= 64 && h1 == gngf::tc::dec::H && h2 == gngf::tc::dec::H && out_dim >= 1 && out_dim <= 4;
}

int gngf_mlp3_tc_fwd(const float* enc, int64_t P, int32_t in_dim, int32_t out_dim, int32_t leaky, const float* w0,
                     const float* b0, const float* w1, const float* b1, const float* w2, const float* b2, float* rgb,
                     uint32_t* masks, void* stream) {
  using namespace gngf::tc::dec;
  if (!gngf_mlp3_tc_supported(in_dim, H, H, out_dim) || P < 0) return GNGF_ERR_INVALID_ARGUMENT;
  if (P == 0) return GNGF_OK;
  if (cudaFuncSetAttribute(mlp3_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(FWD_SMEM)) !=
      cudaSuccess)
    return gngf::check_launch();
  const int grid = static_cast<int>(std::min<int64_t>(gngf::ceil_div(P, TP), 2 * static_cast<int64_t>(gngf::sm_count())));
  mlp3_tc_fwd_kernel<<<grid, THREADS, FWD_SMEM, gngf::as_stream(stream)>>>(enc, P, in_dim, out_dim, leaky, w0, b0, w1, b1,
                                                                          w2, b2, rgb, masks);
  gngf::note_launch();
  return gngf::check_launch();
}

int gngf_mlp3_tc_bwd(const float* enc, const float* rgb, const float* drgb, int64_t P, int32_t in_dim, int32_t out_dim,
                     int32_t leaky, const float* w0, const float* b0, const float* w1, const float* b1, const float* w2,
                     const uint32_t* masks, float* denc, float* dw0, float* db0, float* dw1, float* db1, float* dw2,
                     float* db2, void* stream) {
  using namespace gngf::tc::dec;
  if (!gngf_mlp3_tc_supported(in_dim, H, H, out_dim) || P < 0) return GNGF_ERR_INVALID_ARGUMENT;
  if (P == 0) return GNGF_OK;
  if ((reinterpret_cast<uintptr_t>(dw1) & 15) || !masks) return GNGF_ERR_INVALID_ARGUMENT;
  if (cudaFuncSetAttribute(mlp3_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(BWD_SMEM)) !=
      cudaSuccess)
    return gngf::check_launch();
  const int grid = static_cast<int>(std::min<int64_t>(gngf::ceil_div(P, TP), gngf::sm_count()));
  mlp3_tc_bwd_kernel<<<grid, BWD_THREADS, BWD_SMEM, gngf::as_stream(stream)>>>(enc, rgb, drgb, P, in_dim, out_dim, leaky, w0, b0,
                                                                          w1, b1, w2, masks, denc, dw0, db0, dw1, db1, dw2,
                                                                          db2);
  gngf::note_launch();
  return gngf::check_launch();
}

}  // extern "C"
