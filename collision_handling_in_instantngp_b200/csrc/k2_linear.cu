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
  // y[m,n] = sum_k x[m*K + k] * w[n*K + k]
  return gngf::launch_sgemm(x, K, 1, w, 1, K, y, N, M, N, K, 1, ep, gngf::as_stream(stream));
}

int gngf_linear_bwd(const float* dz, const float* x, const float* w, int64_t M, int32_t N, int32_t K, int32_t act_prev,
                    float* dx, float* dw, float* db, void* stream) {
  if (M < 0 || N <= 0 || K <= 0) return GNGF_ERR_INVALID_ARGUMENT;
  if (M == 0) return GNGF_OK;
  cudaStream_t st = gngf::as_stream(stream);
  int rc = GNGF_OK;
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
