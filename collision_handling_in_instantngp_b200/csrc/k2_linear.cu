// Generic fp32 linear layers on CUDA cores (true fp32 FMA, the reference's arithmetic: cuBLAS SGEMM with
// allow_tf32=False).  One strided 64x64x16 tile kernel serves
//   forward   y  = act(x w^T + b)               (models.py:105-106, 469-470)
//   backward  dx = (dz w) .* act'(x)            autograd of the same lines
//             dw += dz^T x  (split over rows, vector atomics), db += colsum(dz)
// The HPD's first layer (in_features = 2) never materialises its input: it is evaluated directly from the
// lattice node coordinates.  The tensor-core path for the wide HPD output layer lives in k2_hpd_tc.cu.
#include <algorithm>

#include "common.cuh"

namespace gngf {

constexpr int BK = 16;  // tiles: 64x64 (4x4 per thread) for large problems, 32x32 (2x2) to spread small ones

struct Epilogue {
  const float* bias;   // (N) or null
  int act;             // GNGF_ACT_*
  const float* mask;   // same shape as C, or null: C *= act'(mask)
  int mask_act;
  int atomic;          // split-K accumulation into a zeroed C
};

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case GNGF_ACT_RELU: return fmaxf(v, 0.0f);
    case GNGF_ACT_LEAKY_RELU: return v > 0.0f ? v : v * 0.01f;
    case GNGF_ACT_SIGMOID: return 1.0f / (1.0f + expf(-v));
    default: return v;
  }
}

// C[m,n] = sum_k A[m*sam + k*sak] * B[k*sbk + n*sbn]
template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(256) sgemm_kernel(const float* __restrict__ A, int64_t sam, int64_t sak,
                                                    const float* __restrict__ B, int64_t sbk, int64_t sbn,
                                                    float* __restrict__ C, int64_t ldc, int64_t M, int64_t N,
                                                    int64_t K, int64_t k_chunk, Epilogue ep) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  // (row tiles on grid.x: up to 2^31 - 1 of them -- 67 M lattice nodes at the 8192^2 configuration)
  const int64_t m0 = static_cast<int64_t>(blockIdx.x) * BM, n0 = static_cast<int64_t>(blockIdx.y) * BN;
  const int64_t k_begin = static_cast<int64_t>(blockIdx.z) * k_chunk;
  const int64_t k_end = min(K, k_begin + k_chunk);
  const bool a_kfast = (sak == 1), b_kfast = (sbk == 1);

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;

  for (int64_t kt = k_begin; kt < k_end; kt += BK) {
#pragma unroll
    for (int r = 0; r < (BM * BK) / 256; ++r) {
      const int e = tid + r * 256;
      const int kk = a_kfast ? (e % BK) : (e / BM);
      const int mm = a_kfast ? (e / BK) : (e % BM);
      const int64_t m = m0 + mm, k = kt + kk;
      As[kk][mm] = (m < M && k < k_end) ? A[m * sam + k * sak] : 0.0f;
    }
#pragma unroll
    for (int r = 0; r < (BN * BK) / 256; ++r) {
      const int e = tid + r * 256;
      const int kk = b_kfast ? (e % BK) : (e / BN);
      const int nn = b_kfast ? (e / BK) : (e % BN);
      const int64_t n = n0 + nn, k = kt + kk;
      Bs[kk][nn] = (n < N && k < k_end) ? B[k * sbk + n * sbn] : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float av[TM], bv[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) av[i] = As[kk][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) bv[j] = Bs[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int64_t m = m0 + ty * TM + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int64_t n = n0 + tx * TN + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (ep.atomic) {
        atomicAdd(&C[m * ldc + n], v);
        continue;
      }
      if (ep.bias) v += ep.bias[n];
      v = apply_act(v, ep.act);
      if (ep.mask) {
        const float x = ep.mask[m * ldc + n];
        if (ep.mask_act == GNGF_ACT_RELU) v = x > 0.0f ? v : 0.0f;
        else if (ep.mask_act == GNGF_ACT_LEAKY_RELU) v = x > 0.0f ? v : v * 0.01f;
      }
      C[m * ldc + n] = v;
    }
  }
}

static int launch_sgemm(const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbk, int64_t sbn, float* C,
                        int64_t ldc, int64_t M, int64_t N, int64_t K, int split, Epilogue ep, cudaStream_t st) {
  if (M == 0 || N == 0) return GNGF_OK;
  const bool small = ceil_div(M, 64) * ceil_div(N, 64) < 2 * sm_count();
  const int bm = small ? 32 : 64;
  if (ep.atomic) {  // reduction-heavy (dW): enough CTAs to fill the chip
    const int64_t tiles = ceil_div(M, bm) * ceil_div(N, bm);
    split = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(ceil_div(K, 64), (4 * sm_count()) / tiles)));
  }
  split = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(split, ceil_div(K, BK))));
  int64_t k_chunk = ceil_div(ceil_div(K, split), BK) * BK;
  split = static_cast<int>(ceil_div(K, k_chunk));
  if (split > 1 && !ep.atomic) return GNGF_ERR_INVALID_ARGUMENT;
  const int64_t gy = ceil_div(M, bm), gx = ceil_div(N, bm);
  if (gx > 65535 || gy >= (1ll << 31)) return GNGF_ERR_UNSUPPORTED;
  dim3 grid(static_cast<unsigned>(gy), static_cast<unsigned>(gx), static_cast<unsigned>(split));
  if (small)
    sgemm_kernel<32, 32, 2, 2><<<grid, 256, 0, st>>>(A, sam, sak, B, sbk, sbn, C, ldc, M, N, K, k_chunk, ep);
  else
    sgemm_kernel<64, 64, 4, 4><<<grid, 256, 0, st>>>(A, sam, sak, B, sbk, sbn, C, ldc, M, N, K, k_chunk, ep);
  note_launch();
  return check_launch();
}

// ---- skinny layers: M huge, N and K <= 128 (the hidden HPD layers on 30 M lattice nodes) ----------------------------
// The generic tile kernel above spends its time per BLOCK there, not per FLOP: 468 k blocks of two k-steps each, scalar
// strided loads, four scalar stores per 16 bytes -- 16.8 ms per launch at BASELINE.json configs[3] whatever the layer
// (2 -> 8 x the HBM time).  Here a persistent CTA keeps the whole weight matrix in shared memory (k-major: B[k][n]),
// walks 128-row tiles (rows read as float4 along k, parked k-major so that the 8 rows of a thread are two LDS.128), and
// every thread owns 8 rows x TN = N / 16 columns: for N = 128 that is 64 FMAs per 4 LDS.128 and k.  Same fp32 FMA
// arithmetic as the generic kernel.
//   TRANS_B = true : B[k][n] = w[n * K + k]   (forward: y = x w^T, w is (N, K))
//   TRANS_B = false: B[k][n] = w[k * N + n]   (backward: dx = dz w, w is (K = out features, N = in features))
constexpr int SK_BM = 128;
template <int NOUT, bool TRANS_B>
__global__ void __launch_bounds__(256)
    skinny_linear_kernel(const float* __restrict__ X, const float* __restrict__ W, float* __restrict__ Y, int64_t M, int N,
                         int K, Epilogue ep) {
  // NOUT = N in {32, 64, 128}: TN columns x TM rows per thread, (N / TN) column groups x (256 TN / N) row groups = 128 rows
  constexpr int TN = NOUT == 128 ? 8 : 4, CG = NOUT / TN, TM = SK_BM / (256 / CG);
  extern __shared__ float sk_smem[];
  const int ldb = N + 4, lda = SK_BM + 4;
  float* Bs = sk_smem;                 // [K][N + 4]
  float* As = sk_smem + K * ldb;       // [K][128 + 4]
  const int tid = threadIdx.x, tx = tid % CG, ty = tid / CG;
  for (int e = tid; e < N * K; e += 256) {
    const int k = TRANS_B ? (e % K) : (e / N), n = TRANS_B ? (e / K) : (e % N);   // (coalesced read of w either way)
    Bs[k * ldb + n] = W[e];
  }
  const int64_t tiles = (M + SK_BM - 1) / SK_BM;
  const int k4 = K / 4;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t m0 = tile * SK_BM;
    __syncthreads();   // (previous tile's reads of As are over; first pass: Bs is complete)
    for (int e = tid; e < SK_BM * k4; e += 256) {
      const int r = e / k4, c = (e - r * k4) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m0 + r < M) v = *reinterpret_cast<const float4*>(X + (m0 + r) * K + c);
      As[(c + 0) * lda + r] = v.x;
      As[(c + 1) * lda + r] = v.y;
      As[(c + 2) * lda + r] = v.z;
      As[(c + 3) * lda + r] = v.w;
    }
    __syncthreads();
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;
#pragma unroll 4
    for (int kk = 0; kk < K; ++kk) {
      float av[TM];
#pragma unroll
      for (int i = 0; i < TM; i += 4) {
        const float4 a = *reinterpret_cast<const float4*>(As + kk * lda + ty * TM + i);
        av[i] = a.x; av[i + 1] = a.y; av[i + 2] = a.z; av[i + 3] = a.w;
      }
      float bv[TN];
#pragma unroll
      for (int j = 0; j < TN; j += 4) {
        const float4 b = *reinterpret_cast<const float4*>(Bs + kk * ldb + tx * TN + j);
        bv[j] = b.x; bv[j + 1] = b.y; bv[j + 2] = b.z; bv[j + 3] = b.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int64_t m = m0 + ty * TM + i;
      if (m >= M) continue;
#pragma unroll
      for (int j = 0; j < TN; j += 4) {
        const int n = tx * TN + j;
        float v[4] = {acc[i][j], acc[i][j + 1], acc[i][j + 2], acc[i][j + 3]};
        if (ep.bias) {
          const float4 b = *reinterpret_cast<const float4*>(ep.bias + n);
          v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = apply_act(v[q], ep.act);
        if (ep.mask) {
          const float4 x = *reinterpret_cast<const float4*>(ep.mask + m * N + n);
          const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (ep.mask_act == GNGF_ACT_RELU) v[q] = xs[q] > 0.0f ? v[q] : 0.0f;
            else if (ep.mask_act == GNGF_ACT_LEAKY_RELU) v[q] = xs[q] > 0.0f ? v[q] : v[q] * 0.01f;
          }
        }
        *reinterpret_cast<float4*>(Y + m * N + n) = make_float4(v[0], v[1], v[2], v[3]);
      }
    }
  }
}

// dw[n, k] += sum_m dz[m, n] x[m, k],  db[n] += sum_m dz[m, n]   (N, K <= 128; N % 16 == 0, K % 16 == 0, N / 16 * K / 16 <= 32)
// Persistent CTAs own a contiguous run of rows; thread (tn, tk) keeps an (N / 16) x (K / 16) block of dw in registers over
// the whole run, rows travel through shared memory in chunks of 32 (float4 loads), one atomic per CTA and element at the end.
template <int TN, int TK>
__global__ void __launch_bounds__(256)
    skinny_dw_kernel(const float* __restrict__ dz, const float* __restrict__ x, int64_t M, int N, int K, int64_t rows_per_cta,
                     float* __restrict__ dw, float* __restrict__ db) {
  extern __shared__ float sk_smem[];
  constexpr int CH = 32;
  float* Ds = sk_smem;              // [CH][N]
  float* Xs = sk_smem + CH * N;     // [CH][K]
  const int tid = threadIdx.x, tn = tid / 16, tk = tid % 16;
  const int64_t r_begin = static_cast<int64_t>(blockIdx.x) * rows_per_cta, r_end = min(M, r_begin + rows_per_cta);
  float acc[TN][TK], bsum[TN];
#pragma unroll
  for (int i = 0; i < TN; ++i) {
    bsum[i] = 0.0f;
#pragma unroll
    for (int j = 0; j < TK; ++j) acc[i][j] = 0.0f;
  }
  const int n4 = N / 4, k4 = K / 4;
  for (int64_t r0 = r_begin; r0 < r_end; r0 += CH) {
    __syncthreads();
    for (int e = tid; e < CH * n4; e += 256) {
      const int r = e / n4, c = (e - r * n4) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r0 + r < r_end) v = *reinterpret_cast<const float4*>(dz + (r0 + r) * N + c);
      *reinterpret_cast<float4*>(Ds + r * N + c) = v;
    }
    for (int e = tid; e < CH * k4; e += 256) {
      const int r = e / k4, c = (e - r * k4) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r0 + r < r_end) v = *reinterpret_cast<const float4*>(x + (r0 + r) * K + c);
      *reinterpret_cast<float4*>(Xs + r * K + c) = v;
    }
    __syncthreads();
#pragma unroll 4
    for (int r = 0; r < CH; ++r) {
      float a[TN], b[TK];
#pragma unroll
      for (int i = 0; i < TN; i += 4) {
        const float4 v = *reinterpret_cast<const float4*>(Ds + r * N + tn * TN + i);
        a[i] = v.x; a[i + 1] = v.y; a[i + 2] = v.z; a[i + 3] = v.w;
      }
      if constexpr (TK == 2) {
        const float2 v = *reinterpret_cast<const float2*>(Xs + r * K + tk * TK);
        b[0] = v.x; b[1] = v.y;
      } else {
#pragma unroll
        for (int j = 0; j < TK; j += 4) {
          const float4 v = *reinterpret_cast<const float4*>(Xs + r * K + tk * TK + j);
          b[j] = v.x; b[j + 1] = v.y; b[j + 2] = v.z; b[j + 3] = v.w;
        }
      }
#pragma unroll
      for (int i = 0; i < TN; ++i) {
        if (tk == 0) bsum[i] += a[i];
#pragma unroll
        for (int j = 0; j < TK; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < TN; ++i) {
    const int n = tn * TN + i;
#pragma unroll
    for (int j = 0; j < TK; ++j) atomicAdd(dw + static_cast<int64_t>(n) * K + tk * TK + j, acc[i][j]);
    if (tk == 0 && db) atomicAdd(db + n, bsum[i]);
  }
}

static bool skinny_ok(int64_t M, int N, int K) {
  return M >= 16384 && (N == 32 || N == 64 || N == 128) && K <= 128 && (K % 4) == 0;
}

template <bool TRANS_B>
static int launch_skinny_linear(const float* X, const float* W, float* Y, int64_t M, int N, int K, Epilogue ep,
                                cudaStream_t st) {
  const size_t smem = sizeof(float) * (static_cast<size_t>(K) * (N + 4) + static_cast<size_t>(K) * (SK_BM + 4));
  const unsigned grid = static_cast<unsigned>(std::min<int64_t>(ceil_div(M, SK_BM), 2 * sm_count()));
  auto go = [&](auto kernel) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    kernel<<<grid, 256, smem, st>>>(X, W, Y, M, N, K, ep);
  };
  if (N == 128) go(skinny_linear_kernel<128, TRANS_B>);
  else if (N == 64) go(skinny_linear_kernel<64, TRANS_B>);
  else go(skinny_linear_kernel<32, TRANS_B>);
  note_launch();
  return check_launch();
}

// db[n] += sum_m dz[m, n]
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ dz, int64_t M, int64_t N,
                                                     int64_t rows_per_block, float* __restrict__ db) {
  __shared__ float part[8][33];
  const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;
  const int64_t n = static_cast<int64_t>(blockIdx.y) * 32 + tx;
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * rows_per_block;
  const int64_t r1 = min(M, r0 + rows_per_block);
  float s = 0.0f;
  if (n < N)
    for (int64_t r = r0 + ty; r < r1; r += 8) s += dz[r * N + n];
  part[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && n < N) {
#pragma unroll
    for (int i = 1; i < 8; ++i) s += part[i][tx];
    atomicAdd(&db[n], s);
  }
}

// h[r, j] = act(cx * w0[j,0] + cy * w0[j,1] + b0[j]);  row r <-> node u = node_ids ? node_ids[r] : r,
// (cx, cy) = (ox + u / wy, oy + u % wy)
__global__ void __launch_bounds__(256) first_layer_fwd_kernel(const __grid_constant__ gngf_lattice lat,
                                                              const int* __restrict__ node_ids, int64_t rows,
                                                              const float2* __restrict__ w0,
                                                              const float* __restrict__ b0, int n_out, int act,
                                                              float* __restrict__ h) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= rows * n_out) return;
  const int64_t r = i / n_out;
  const int j = static_cast<int>(i - r * n_out);
  const int64_t u = node_ids ? node_ids[r] : r;
  const float cx = static_cast<float>(lat.ox + static_cast<int>(u / lat.wy));
  const float cy = static_cast<float>(lat.oy + static_cast<int>(u % lat.wy));
  const float2 w = w0[j];
  h[i] = apply_act(fmaf(cy, w.y, fmaf(cx, w.x, b0[j])), act);
}

// dw0[j, :] += sum_r dz[r, j] * (cx, cy);  db0[j] += sum_r dz[r, j]
// A block takes a contiguous run of rows; warp w walks rows r0 + w, r0 + w + 8, ... with lane = output unit (coalesced
// 128-byte row reads, the node coordinate computed once per row and warp), the eight warps' partials meet in shared
// memory and leave as one atomic per (block, unit, term).  (The first version walked the rows serially in n_out
// threads: 18 ms for 30 M rows; this form streams dz at HBM speed.)
__global__ void __launch_bounds__(256) first_layer_bwd_kernel(const __grid_constant__ gngf_lattice lat,
                                                              const int* __restrict__ node_ids, int64_t rows,
                                                              const float* __restrict__ dz, int n_out,
                                                              int64_t rows_per_block, float* __restrict__ dw0,
                                                              float* __restrict__ db0) {
  __shared__ float part[8][3][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * rows_per_block;
  const int64_t r1 = min(rows, r0 + rows_per_block);
  for (int j0 = 0; j0 < n_out; j0 += 32) {
    const int j = j0 + lane;
    float sx = 0.0f, sy = 0.0f, sb = 0.0f;
    for (int64_t r = r0 + warp; r < r1; r += 8) {
      const int64_t u = node_ids ? node_ids[r] : r;
      const int qx = static_cast<int>(u / lat.wy);
      const float cx = static_cast<float>(lat.ox + qx), cy = static_cast<float>(lat.oy + static_cast<int>(u - int64_t(qx) * lat.wy));
      const float g = j < n_out ? dz[r * n_out + j] : 0.0f;
      sx = fmaf(g, cx, sx);
      sy = fmaf(g, cy, sy);
      sb += g;
    }
    part[warp][0][lane] = sx;
    part[warp][1][lane] = sy;
    part[warp][2][lane] = sb;
    __syncthreads();
    if (warp < 3 && j < n_out) {
      float t = 0.0f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += part[w][warp][lane];
      if (warp == 2) atomicAdd(&db0[j], t);
      else atomicAdd(&dw0[j * 2 + warp], t);
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) sigmoid_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                                          int64_t n, float* __restrict__ dz) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i < n) {
    const float v = y[i];
    dz[i] = dy[i] * v * (1.0f - v);
  }
}

}  // namespace gngf

extern "C" {

int gngf_hpd_first_layer_fwd_nodes(gngf_lattice lat, const int32_t* node_ids, int64_t n_nodes, const float* w0,
                                   const float* b0, int32_t n_out, int32_t act, float* h, void* stream) {
  const int64_t U = static_cast<int64_t>(lat.wx) * lat.wy;
  const int64_t rows = node_ids ? n_nodes : U;
  if (U <= 0 || n_out <= 0 || rows < 0 || rows > U) return GNGF_ERR_INVALID_ARGUMENT;
  if (rows == 0) return GNGF_OK;
  gngf::first_layer_fwd_kernel<<<static_cast<unsigned>(gngf::ceil_div(rows * n_out, 256)), 256, 0,
                                 gngf::as_stream(stream)>>>(lat, node_ids, rows, reinterpret_cast<const float2*>(w0), b0,
                                                            n_out, act, h);
  gngf::note_launch();
  return gngf::check_launch();
}

int gngf_hpd_first_layer_fwd(gngf_lattice lat, const float* w0, const float* b0, int32_t n_out, int32_t act, float* h,
                             void* stream) {
  return gngf_hpd_first_layer_fwd_nodes(lat, nullptr, 0, w0, b0, n_out, act, h, stream);
}

int gngf_hpd_first_layer_bwd_nodes(gngf_lattice lat, const int32_t* node_ids, int64_t n_nodes, const float* dz,
                                   int32_t n_out, float* dw0, float* db0, void* stream) {
  const int64_t U = static_cast<int64_t>(lat.wx) * lat.wy;
  const int64_t rows = node_ids ? n_nodes : U;
  if (U <= 0 || n_out <= 0 || rows < 0 || rows > U) return GNGF_ERR_INVALID_ARGUMENT;
  if (rows == 0) return GNGF_OK;
  // enough blocks to cover the chip even for a few hundred nodes
  const int64_t rows_per_block = std::max<int64_t>(8, gngf::ceil_div(rows, 8 * gngf::sm_count()));
  gngf::first_layer_bwd_kernel<<<static_cast<unsigned>(gngf::ceil_div(rows, rows_per_block)), 256, 0,
                                 gngf::as_stream(stream)>>>(lat, node_ids, rows, dz, n_out, rows_per_block, dw0, db0);
  gngf::note_launch();
  return gngf::check_launch();
}

int gngf_hpd_first_layer_bwd(gngf_lattice lat, const float* dz, int32_t n_out, float* dw0, float* db0, void* stream) {
  return gngf_hpd_first_layer_bwd_nodes(lat, nullptr, 0, dz, n_out, dw0, db0, stream);
}

int gngf_linear_fwd(const float* x, const float* w, const float* b, int64_t M, int32_t N, int32_t K, int32_t act,
                    float* y, void* stream) {
  if (M < 0 || N <= 0 || K <= 0) return GNGF_ERR_INVALID_ARGUMENT;
  gngf::Epilogue ep{b, act, nullptr, 0, 0};
  if (gngf::skinny_ok(M, N, K) && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) |
                                    reinterpret_cast<uintptr_t>(b)) & 15) == 0)
    return gngf::launch_skinny_linear<true>(x, w, y, M, N, K, ep, gngf::as_stream(stream));
  // y[m,n] = sum_k x[m*K + k] * w[n*K + k]
  return gngf::launch_sgemm(x, K, 1, w, 1, K, y, N, M, N, K, 1, ep, gngf::as_stream(stream));
}

int gngf_linear_bwd(const float* dz, const float* x, const float* w, int64_t M, int32_t N, int32_t K, int32_t act_prev,
                    float* dx, float* dw, float* db, void* stream) {
  if (M < 0 || N <= 0 || K <= 0) return GNGF_ERR_INVALID_ARGUMENT;
  if (M == 0) return GNGF_OK;
  cudaStream_t st = gngf::as_stream(stream);
  int rc = GNGF_OK;
  const bool aligned = ((reinterpret_cast<uintptr_t>(dz) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0;
  if (dw && aligned && M >= 16384 && N <= 128 && K <= 128 && (N % 64) == 0 && (K % 32) == 0 && (N / 16) * (K / 16) <= 32) {
    // skinny layer: dw and db from one persistent pass over the rows
    const int ctas = 2 * gngf::sm_count();
    const int64_t rows_per_cta = gngf::ceil_div(gngf::ceil_div(M, ctas), 32) * 32;
    const unsigned grid = static_cast<unsigned>(gngf::ceil_div(M, rows_per_cta));
    const size_t smem = sizeof(float) * 32 * (static_cast<size_t>(N) + K);
    if (N == 128 && K == 64) gngf::skinny_dw_kernel<8, 4><<<grid, 256, smem, st>>>(dz, x, M, N, K, rows_per_cta, dw, db);
    else if (N == 64 && K == 32) gngf::skinny_dw_kernel<4, 2><<<grid, 256, smem, st>>>(dz, x, M, N, K, rows_per_cta, dw, db);
    else if (N == 128 && K == 32) gngf::skinny_dw_kernel<8, 2><<<grid, 256, smem, st>>>(dz, x, M, N, K, rows_per_cta, dw, db);
    else if (N == 64 && K == 64) gngf::skinny_dw_kernel<4, 4><<<grid, 256, smem, st>>>(dz, x, M, N, K, rows_per_cta, dw, db);
    else if (N == 128 && K == 128) gngf::skinny_dw_kernel<8, 8><<<grid, 256, smem, st>>>(dz, x, M, N, K, rows_per_cta, dw, db);
    else if (N == 64 && K == 128) gngf::skinny_dw_kernel<4, 8><<<grid, 256, smem, st>>>(dz, x, M, N, K, rows_per_cta, dw, db);
    else goto generic_dw;
    gngf::note_launch();
    if ((rc = gngf::check_launch())) return rc;
    dw = nullptr;
    db = nullptr;
  }
generic_dw:
  if (dw) {
    // dw[n,k] = sum_m dz[m*N + n] * x[m*K + k]; rows are the reduction dimension -> split + atomics
    gngf::Epilogue ep{nullptr, 0, nullptr, 0, 1};
    rc = gngf::launch_sgemm(dz, 1, N, x, K, 1, dw, K, N, K, M, 0, ep, st);
    if (rc) return rc;
  }
  if (db) {
    const int64_t col_blocks = gngf::ceil_div(N, 32);
    const int64_t rows_per_block =
        std::min<int64_t>(256, std::max<int64_t>(32, gngf::ceil_div(M * col_blocks, 4 * gngf::sm_count())));
    if (col_blocks > 65535) return GNGF_ERR_UNSUPPORTED;
    dim3 grid(static_cast<unsigned>(gngf::ceil_div(M, rows_per_block)), static_cast<unsigned>(col_blocks));
    gngf::colsum_kernel<<<grid, 256, 0, st>>>(dz, M, N, rows_per_block, db);
    gngf::note_launch();
    rc = gngf::check_launch();
    if (rc) return rc;
  }
  if (dx) {
    // dx[m,k] = sum_n dz[m*N + n] * w[n*K + k], masked by act'(x)
    gngf::Epilogue ep{nullptr, 0, act_prev == GNGF_ACT_NONE ? nullptr : x, act_prev, 0};
    // output width K, contraction over N: B[n][k] = w[n * K + k].  (Measured at 30 M rows: 128 -> 64 is 21 ms here against
    // 16.8 ms in the generic kernel -- 100 KB of shared memory per CTA --, 64 -> 32 is 8.0 against 16.8.)
    if (aligned && N <= 64 && gngf::skinny_ok(M, K, N))
      return gngf::launch_skinny_linear<false>(dz, w, dx, M, K, N, ep, st);
    rc = gngf::launch_sgemm(dz, N, 1, w, K, 1, dx, K, M, K, N, 1, ep, st);
  }
  return rc;
}

int gngf_sigmoid_bwd(const float* dy, const float* y, int64_t n, float* dz, void* stream) {
  if (n < 0) return GNGF_ERR_INVALID_ARGUMENT;
  if (n == 0) return GNGF_OK;
  gngf::sigmoid_bwd_kernel<<<static_cast<unsigned>(gngf::ceil_div(n, 256)), 256, 0, gngf::as_stream(stream)>>>(dy, y, n,
                                                                                                              dz);
  gngf::note_launch();
  return gngf::check_launch();
}

}  // extern "C"
