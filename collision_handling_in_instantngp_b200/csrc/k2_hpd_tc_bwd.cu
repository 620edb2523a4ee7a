// K5c (tensor-core path): backward of the streaming HPD output layer -- softmax + top-k + Linear(Kd, T) with T up to
// 2^22 slots (models.py:80-88, 105-123 and DifferentiableTopk.backward, models.py:21-42) -- without ever writing the
// (U, T) logits, probabilities or dlogits.
//
// In top-k-only mode the adjoint of the probabilities is non-zero at the K selected slots only, so per lattice node r
//     dlogit[r, t] = -<G,p>_r * p[r, t]  +  sum_k [t == t_k] p_k g_k ,      p[r, t] = exp(z[r, t] - max_r) / sum_r .
// The K-sparse part is a gather / scatter of K rows per node (hpd_stream_bwd_sparse_kernel).  The dense part is an
// attention-shaped pair of products with E[r, t] = a_r exp(z[r, t] - max_r), a_r = -<G,p>_r / sum_r:
//     dh  (U, Kd) = E   W3           dW3 (T, Kd) += E^T h           db3 (T) += E^T 1
// Both are instances of ONE kernel (hpd_stream_bwd_kernel<DW>):
//     X (128 x Kd) resident, Y tiles (64 x Kd) streamed by TMA;   S = X Y^T  (tcgen05, TMEM, double-buffered)
//     E = rowscale_i colscale_j exp2(S log2e + rowoff_i + coloff_j)   (8 epilogue warps; E -> bf16 planes -> smem)
//     O (128 x Kd) += E Y   (second tcgen05 product: A = E K-major from shared memory, B = the SAME Y tile read
//                            MN-major -- no transposed copy of anything), accumulated in TMEM over all Y tiles
//   DW = false:  X = h tile,  Y = W3 tiles:  O = dh tile          (rows carry (a_r, -max_r log2e), columns the bias)
//   DW = true :  X = W3 tile, Y = h tiles :  O = dW3 tile, db3    (rows carry the bias, columns (a_r, -max_r log2e))
// Operands are two bf16 planes (hi, mid) of the fp32 values and each product is hi.hi + hi.mid + mid.hi (relative
// error ~1e-5: the gradient tolerance is 1e-4; the forward, which must reproduce top-k selections exactly, uses three
// planes and six products).  MMA issue order S(j+1), O(j) keeps the tensor pipe busy while the epilogue turns S(j)
// into E(j).
#include <algorithm>

#include "tc_common.cuh"

namespace gngf {
namespace tc {
namespace sb {

constexpr int SBN = 64;                                   // rows of a streamed Y tile (= columns of S and E)
constexpr int NP = 2;                                     // planes used: hi, mid
constexpr int YSTAGES = 3;
constexpr int EPI_WARPS = 8;                              // (TMEM lane quarter) x (32-column half of the S tile)
constexpr int THREADS = 64 + 32 * EPI_WARPS;
constexpr uint32_t XP_BYTES = 128 * 128;                  // one (k-block, plane) of X: 128 rows x 64 bf16
constexpr uint32_t X_BYTES = 2 * NP * XP_BYTES;           // 64 KB
constexpr uint32_t YP_BYTES = SBN * 128;                  // one (k-block, plane) of a Y tile: 64 rows x 64 bf16
constexpr uint32_t Y_STAGE_BYTES = 2 * NP * YP_BYTES;     // 32 KB
constexpr uint32_t EP_BYTES = 128 * 128;                  // one plane of E: 128 rows x 64 bf16
constexpr uint32_t E_BUF_BYTES = NP * EP_BYTES;           // 32 KB
constexpr uint32_t TMEM_COLS = 512;                       // S: 2 x 64 columns, O: 128, resident X planes: 2 x 64
constexpr uint32_t O_COL = 128;
constexpr uint32_t X_COL = 256;                           // plane p of X: columns [X_COL + 64 p, +64) (bf16 pairs)
constexpr size_t SMEM_BYTES = X_BYTES + YSTAGES * Y_STAGE_BYTES + 2 * E_BUF_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&v);
}

// x_rows / y_rows: number of valid rows of X / Y (U or T); Kdim <= 128.
// row_off / row_scale: per X row; col_off / col_scale: per Y row.  DW = false: row_off = -max log2e, row_scale = a,
// col vectors come from `bias`; DW = true: the other way round.
template <bool DW>
__global__ void __launch_bounds__(THREADS, 1)
    hpd_stream_bwd_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_y,
                          int x_rows, int y_rows, int Kdim, int n_split, const float* __restrict__ bias,
                          const float* __restrict__ m2neg, const float* __restrict__ ascale, float* __restrict__ out,
                          float* __restrict__ dbias) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* x_buf = smem;
  uint8_t* y_ring = smem + X_BYTES;
  uint8_t* e_bufs = y_ring + YSTAGES * Y_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(e_bufs + 2 * E_BUF_BYTES);
  uint64_t* x_full = bars;
  uint64_t* x_empty = bars + 1;
  uint64_t* y_full = bars + 2;
  uint64_t* y_empty = y_full + YSTAGES;
  uint64_t* s_full = y_empty + YSTAGES;
  uint64_t* s_empty = s_full + 2;
  uint64_t* e_full = s_empty + 2;
  uint64_t* e_empty = e_full + 2;
  uint64_t* o_full = e_empty + 2;
  uint64_t* o_empty = o_full + 1;
  uint64_t* xt_full = o_empty + 1;          // the epilogue warps have copied the X tile into tensor memory
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xt_full + 1);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_x)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_y)) : "memory");
    mbar_init(x_full, 1);
    mbar_init(x_empty, 1);
    for (int s = 0; s < YSTAGES; ++s) {
      mbar_init(y_full + s, 1);
      mbar_init(y_empty + s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(s_full + s, 1);
      mbar_init(s_empty + s, EPI_WARPS);
      mbar_init(e_full + s, EPI_WARPS);
      mbar_init(e_empty + s, 1);
    }
    mbar_init(o_full, 1);
    mbar_init(o_empty, EPI_WARPS);
    mbar_init(xt_full, EPI_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int x_tiles = (x_rows + BM - 1) / BM, y_tiles = (y_rows + SBN - 1) / SBN;
  const int tiles_per_split = (y_tiles + n_split - 1) / n_split;
  const int items = x_tiles * n_split;
  const int kblocks = (Kdim + BK - 1) / BK;   // 1 or 2
  const int n2 = kblocks * BK;                // N of the second product (columns of O)

  if (warp == 0) {
    if (lane == 0) {  // ---- TMA producer ----
      int stage = 0;
      uint32_t phase = 0, x_phase = 0;
      for (int w = blockIdx.x; w < items; w += gridDim.x) {
        const int m0 = (w / n_split) * BM, sp = w % n_split;
        const int t0 = sp * tiles_per_split, t1 = min(y_tiles, t0 + tiles_per_split);
        if (t0 >= t1) continue;   // empty split (every role skips it)
        mbar_wait(x_empty, x_phase ^ 1);
        x_phase ^= 1;
        mbar_expect_tx(x_full, kblocks * NP * XP_BYTES);
        for (int kb = 0; kb < kblocks; ++kb)
          for (int pl = 0; pl < NP; ++pl) tma_load_3d(x_buf + (kb * NP + pl) * XP_BYTES, &map_x, kb * BK, m0, pl, x_full);
        for (int t = t0; t < t1; ++t) {
          mbar_wait(y_empty + stage, phase ^ 1);
          mbar_expect_tx(y_full + stage, kblocks * NP * YP_BYTES);
          uint8_t* yb = y_ring + stage * Y_STAGE_BYTES;
          for (int kb = 0; kb < kblocks; ++kb)
            for (int pl = 0; pl < NP; ++pl)
              tma_load_3d(yb + (kb * NP + pl) * YP_BYTES, &map_y, kb * BK, t * SBN, pl, y_full + stage);
          if (++stage == YSTAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    {  // ---- MMA issuer: the whole warp runs the loop, lane 0 issues (see tc_common.cuh) ----
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);   // provably warp-uniform copy
      constexpr uint32_t idesc1 = umma_idesc(BM, SBN);
      const uint32_t idesc2 = umma_idesc(BM, n2) | UMMA_B_MN_MAJOR;
      const uint32_t y_lo = umma_desc_lo(smem_u32(y_ring));                      // K-major view (first product)
      const uint32_t y_lo_mn = umma_desc_lo(smem_u32(y_ring), NP * YP_BYTES);    // MN-major view (second product)
      const uint32_t e_lo = umma_desc_lo(smem_u32(e_bufs));
      int s1_stage = 0, s2_stage = 0;          // Y ring positions of the first / second product
      uint32_t s1_phase = 0;
      uint32_t it1 = 0, it2 = 0;               // running tile counters (S / E buffer = counter & 1)
      uint32_t x_phase = 0, o_phase = 0;
      for (int w = blockIdx.x; w < items; w += gridDim.x) {
        const int sp = w % n_split;
        const int t0 = sp * tiles_per_split, t1 = min(y_tiles, t0 + tiles_per_split);
        const int nt = t1 - t0;
        if (nt <= 0) continue;
        mbar_wait(xt_full, x_phase);   // X tile resident in tensor memory (A operand of the first product)
        x_phase ^= 1;
        tc_fence_after();
        for (int j = 0; j <= nt; ++j) {
          if (j < nt) {  // S(j) = X Y_j^T : products hi.hi, hi.mid, mid.hi
            const uint32_t buf = it1 & 1, ph = (it1 >> 1) & 1;
            mbar_wait(y_full + s1_stage, s1_phase);
            mbar_wait(s_empty + buf, ph ^ 1);
            tc_fence_after();
            const uint32_t yb = y_lo + s1_stage * (Y_STAGE_BYTES >> 4);
            const uint32_t d = tmem_u + buf * SBN;
#pragma unroll
            for (int pr = 0; pr < 3; ++pr) {
              const int pa = pr >> 1, pb = pr & 1;   // (0,0) (0,1) (1,0)
#pragma unroll
              for (int kb = 0; kb < 2; ++kb) {
                if (kb < kblocks) {
#pragma unroll
                  for (int k = 0; k < BK / UMMA_K; ++k) {
                    // A from tensor memory (no shared-memory traffic: with N = 64 the product was bound by the 6 KB of
                    // operand reads per MMA); 8 columns per K = 16 step, 32 per k-block, 64 per plane
                    const uint32_t at = tmem_u + X_COL + pa * 64 + kb * 32 + k * 8;
                    const uint64_t bd = umma_desc_pack(yb + (((kb * NP + pb) * YP_BYTES + k * UMMA_K * 2) >> 4));
                    umma_bf16_ts_lead(d, at, bd, idesc1, (pr | kb | k) != 0);
                  }
                }
              }
            }
            umma_commit_lead(s_full + buf);
            ++it1;
            if (++s1_stage == YSTAGES) {
              s1_stage = 0;
              s1_phase ^= 1;
            }
          }
          if (j >= 1) {  // O += E(j-1) Y_{j-1}
            const uint32_t buf = it2 & 1, ph = (it2 >> 1) & 1;
            mbar_wait(e_full + buf, ph);
            if (j == 1) {
              mbar_wait(o_empty, o_phase ^ 1);   // the previous item's O has been read out
              o_phase ^= 1;
            }
            tc_fence_after();
            const uint32_t eb = e_lo + buf * (E_BUF_BYTES >> 4);
            const uint32_t yb = y_lo_mn + s2_stage * (Y_STAGE_BYTES >> 4);
            const uint32_t d = tmem_u + O_COL;
#pragma unroll
            for (int pr = 0; pr < 3; ++pr) {
              const int pa = pr >> 1, pb = pr & 1;
#pragma unroll
              for (int k = 0; k < SBN / UMMA_K; ++k) {
                // A: E plane, K-major (K = streamed index, 16 columns = 32 bytes per step)
                const uint64_t ad = umma_desc_pack(eb + ((pa * EP_BYTES + k * UMMA_K * 2) >> 4));
                // B: Y plane read MN-major: N = feature index (64 contiguous per k-block, k-blocks NP*YP_BYTES apart),
                //    K = streamed index (rows of 128 bytes, 16 rows = 2048 bytes per step)
                const uint64_t bd = umma_desc_pack(yb + ((pb * YP_BYTES + k * UMMA_K * 128) >> 4));
                umma_bf16_lead(d, ad, bd, idesc2, (j != 1) || (pr | k) != 0);
              }
            }
            umma_commit_lead(e_empty + buf);
            umma_commit_lead(y_empty + s2_stage);
            ++it2;
            if (++s2_stage == YSTAGES) s2_stage = 0;
          }
        }
        umma_commit_lead(o_full);
        umma_commit_lead(x_empty);
      }
    }
  } else {  // ---- epilogue warps 2..9 ----
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;       // which 32 of the 64 S columns
    const int row_l = q * 32 + lane;        // row of the X tile / of E / of O
    uint32_t it = 0, o_phase = 0, x_par = 0;
    for (int w = blockIdx.x; w < items; w += gridDim.x) {
      const int m0 = (w / n_split) * BM, sp = w % n_split;
      const int t0 = sp * tiles_per_split, t1 = min(y_tiles, t0 + tiles_per_split);
      if (t0 >= t1) continue;
      const int row = m0 + row_l;
      const bool row_ok = row < x_rows;
      float r_off, r_scale;
      if (DW) {
        r_off = row_ok ? __ldg(bias + row) * LOG2E : 0.0f;
        r_scale = row_ok ? 1.0f : 0.0f;
      } else {
        r_off = row_ok ? __ldg(m2neg + row) : 0.0f;
        r_scale = row_ok ? __ldg(ascale + row) : 0.0f;
      }
      {  // X tile: shared memory (TMA, 128-byte swizzle) -> tensor memory; warp (q, half) copies plane `half` of rows 32q..
        mbar_wait(x_full, x_par);
        x_par ^= 1;
        const uint8_t* xrow = x_buf + (row_l >> 3) * 1024 + (row_l & 7) * 128;
#pragma unroll 1
        for (int kb = 0; kb < 2; ++kb) {
          uint32_t v[32];
          if (kb < kblocks) {
            const uint8_t* src = xrow + (kb * NP + half) * XP_BYTES;
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) {
              const uint4 c = *reinterpret_cast<const uint4*>(src + ((ch ^ (row_l & 7)) << 4));
              v[ch * 4 + 0] = c.x; v[ch * 4 + 1] = c.y; v[ch * 4 + 2] = c.z; v[ch * 4 + 3] = c.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0u;
          }
          tmem_st32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + X_COL + half * 64 + kb * 32, v);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(xt_full);
      }
      float rsum = 0.0f;   // DW: db3[row] = sum of E over all streamed nodes
      for (int t = t0; t < t1; ++t, ++it) {
        const uint32_t buf = it & 1, ph = (it >> 1) & 1;
        const int c0 = t * SBN + half * 32;
        mbar_wait(s_full + buf, ph);
        tc_fence_after();
        uint32_t v[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * SBN + half * 32, v);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_empty + buf);     // S(buf) may be overwritten by the next-but-one product

        float e[32];
        if (c0 + 32 <= y_rows) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 off, sc;
            if (DW) {
              off = __ldg(reinterpret_cast<const float4*>(m2neg + c0 + j));
              sc = __ldg(reinterpret_cast<const float4*>(ascale + c0 + j));
            } else {
              const float4 b = __ldg(reinterpret_cast<const float4*>(bias + c0 + j));
              off = make_float4(b.x * LOG2E, b.y * LOG2E, b.z * LOG2E, b.w * LOG2E);
              sc = make_float4(1.0f, 1.0f, 1.0f, 1.0f);
            }
            e[j + 0] = exp2f(fmaf(__uint_as_float(v[j + 0]), LOG2E, r_off + off.x)) * (r_scale * sc.x);
            e[j + 1] = exp2f(fmaf(__uint_as_float(v[j + 1]), LOG2E, r_off + off.y)) * (r_scale * sc.y);
            e[j + 2] = exp2f(fmaf(__uint_as_float(v[j + 2]), LOG2E, r_off + off.z)) * (r_scale * sc.z);
            e[j + 3] = exp2f(fmaf(__uint_as_float(v[j + 3]), LOG2E, r_off + off.w)) * (r_scale * sc.w);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int c = c0 + j;
            float off = 0.0f, sc = 0.0f;
            if (c < y_rows) {
              if (DW) {
                off = __ldg(m2neg + c);
                sc = __ldg(ascale + c);
              } else {
                off = __ldg(bias + c) * LOG2E;
                sc = 1.0f;
              }
            }
            const float ev = exp2f(fmaf(__uint_as_float(v[j]), LOG2E, r_off + off)) * (r_scale * sc);
            e[j] = c < y_rows ? ev : 0.0f;   // (zero-filled Y rows give S = 0, not a logit: exp2 may overflow there)
          }
        }
        if (!row_ok) {
#pragma unroll
          for (int j = 0; j < 32; ++j) e[j] = 0.0f;
        }
        if (DW) {
#pragma unroll
          for (int j = 0; j < 32; ++j) rsum += e[j];
        }

        // E -> two bf16 planes, K-major with the 128-byte swizzle: 16-byte chunk ch of row r lives at chunk ch ^ (r & 7)
        mbar_wait(e_empty + buf, ph ^ 1);
        uint8_t* eb = e_bufs + buf * E_BUF_BYTES + (row_l >> 3) * 1024 + (row_l & 7) * 128;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          uint32_t hi[4], mid[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float a = e[ch * 8 + 2 * i], b = e[ch * 8 + 2 * i + 1];
            hi[i] = pack_bf16x2(a, b);
            const float ah = __uint_as_float(hi[i] << 16), bh = __uint_as_float(hi[i] & 0xffff0000u);
            mid[i] = pack_bf16x2(a - ah, b - bh);
          }
          const int pch = ((half * 4 + ch) ^ (row_l & 7)) * 16;
          *reinterpret_cast<uint4*>(eb + pch) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(eb + EP_BYTES + pch) = make_uint4(mid[0], mid[1], mid[2], mid[3]);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(e_full + buf);
      }

      // ---- O tile: accumulated in TMEM over the item's Y tiles -> global (reductions: other splits / the sparse part
      //      / earlier batches add into the same buffer)
      mbar_wait(o_full, o_phase);
      o_phase ^= 1;
      tc_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < 64; cc += 32) {
        const int col0 = half * 64 + cc;
        if (col0 >= n2) continue;   // warp-uniform
        uint32_t v[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + O_COL + col0, v);
        if (row_ok) {
          float* o = out + static_cast<int64_t>(row) * Kdim + col0;
          if ((Kdim & 3) == 0 && col0 + 32 <= Kdim) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              red_add_v4(o + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                         __uint_as_float(v[j + 3]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < Kdim) atomicAdd(o + j, __uint_as_float(v[j]));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty);
      if (DW && row_ok && dbias) atomicAdd(dbias + row, rsum);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// per node u: g_k = dtv[u,k] + sum_l cnt[s(l,u)] gcol_k[l,k];  spk[u,k] = p_k g_k;  a_u = -sum_k spk / row_sum;
// m2neg_u = -row_max log2e.  (The dense-adjoint inputs of gngf_hpd_dlogits do not exist in top-k-only mode.)
__global__ void __launch_bounds__(256)
    hpd_stream_bwd_prep_kernel(gngf_lattice lat, int64_t U, int K, const float* __restrict__ utopv,
                               const float* __restrict__ dtv, const int* __restrict__ cnt,
                               const float* __restrict__ gcol_k, const float* __restrict__ row_max,
                               const float* __restrict__ row_sum, float* __restrict__ ascale,
                               float* __restrict__ m2neg, float* __restrict__ spk) {
  const int64_t u = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (u >= U) return;
  const int L = lat.num_levels;
  const int cx = lat.ox + static_cast<int>(u / lat.wy), cy = lat.oy + static_cast<int>(u % lat.wy);
  float dot = 0.0f;
  for (int k = 0; k < K; ++k) {
    float g = dtv[u * K + k];
    if (gcol_k) {
      for (int l = 0; l < L; ++l) {
        const int i = cx - lat.lox[l], j = cy - lat.loy[l];
        if (i >= 0 && i < lat.lwx[l] && j >= 0 && j < lat.lwy[l]) {
          const float c = static_cast<float>(cnt[lat.loff[l] + static_cast<int64_t>(i) * lat.lwy[l] + j]);
          g = fmaf(c, gcol_k[l * K + k], g);
        }
      }
    }
    const float pg = utopv[u * K + k] * g;
    spk[u * K + k] = pg;
    dot += pg;
  }
  float a = -dot / row_sum[u];
  float m = -row_max[u] * LOG2E;
  if (!(fabsf(a) <= 3.0e38f)) a = 0.0f;   // nan_to_num of a degenerate row (models.py:111)
  if (!(fabsf(m) <= 3.0e38f)) m = 0.0f;
  ascale[u] = a;
  m2neg[u] = m;
}

// warp per node: the K selected slots.  dh[u,:] = (dh[u,:] + sum_k spk W3[t_k,:]) .* act'(h[u,:]);
// dW3[t_k,:] += spk h[u,:];  db3[t_k] += spk.  Runs after the dense pass, as the last writer of dh.
__global__ void __launch_bounds__(256)
    hpd_stream_bwd_sparse_kernel(int64_t U, int K, int Kdim, const int* __restrict__ utopi, const float* __restrict__ spk,
                                 const float* __restrict__ h, const float* __restrict__ w, int act_prev,
                                 float* __restrict__ dh, float* __restrict__ dw, float* __restrict__ db) {
  const int64_t u = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) / 32;
  const int lane = threadIdx.x % 32;
  if (u >= U) return;
  const int c = lane * 4;
  const bool on = c < Kdim;   // Kdim % 4 == 0
  float4 hv = make_float4(0.f, 0.f, 0.f, 0.f), acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (on) {
    hv = *reinterpret_cast<const float4*>(h + u * Kdim + c);
    acc = *reinterpret_cast<const float4*>(dh + u * Kdim + c);
  }
  for (int k = 0; k < K; ++k) {
    const int t = utopi[u * K + k];
    const float s = spk[u * K + k];
    if (on) {
      const float4 wv = __ldg(reinterpret_cast<const float4*>(w + static_cast<int64_t>(t) * Kdim + c));
      acc.x = fmaf(s, wv.x, acc.x);
      acc.y = fmaf(s, wv.y, acc.y);
      acc.z = fmaf(s, wv.z, acc.z);
      acc.w = fmaf(s, wv.w, acc.w);
      red_add_v4(dw + static_cast<int64_t>(t) * Kdim + c, s * hv.x, s * hv.y, s * hv.z, s * hv.w);
    }
    if (lane == 0 && db) atomicAdd(db + t, s);
  }
  if (on) {
    if (act_prev == GNGF_ACT_RELU) {
      acc.x = hv.x > 0.0f ? acc.x : 0.0f;
      acc.y = hv.y > 0.0f ? acc.y : 0.0f;
      acc.z = hv.z > 0.0f ? acc.z : 0.0f;
      acc.w = hv.w > 0.0f ? acc.w : 0.0f;
    } else if (act_prev == GNGF_ACT_LEAKY_RELU) {
      acc.x = hv.x > 0.0f ? acc.x : 0.01f * acc.x;
      acc.y = hv.y > 0.0f ? acc.y : 0.01f * acc.y;
      acc.z = hv.z > 0.0f ? acc.z : 0.01f * acc.z;
      acc.w = hv.w > 0.0f ? acc.w : 0.01f * acc.w;
    }
    *reinterpret_cast<float4*>(dh + u * Kdim + c) = acc;
  }
}

static int split_count(int64_t x_tiles, int64_t y_tiles) {
  return static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(y_tiles, (2 * gngf::sm_count()) / x_tiles)));
}

}  // namespace sb
}  // namespace tc
}  // namespace gngf

extern "C" {

int64_t gngf_hpd_stream_bwd_workspace_floats(int64_t U, int32_t topk) { return U * (2 + static_cast<int64_t>(topk)) + 4; }

int gngf_hpd_stream_bwd(gngf_lattice lat, const uint16_t* h_planes, const uint16_t* w_planes, const float* h,
                        const float* w, const float* bias, int64_t U, int64_t T, int64_t Kdim, int32_t topk,
                        const float* utopv, const int32_t* utopi, const float* dtv, const int32_t* cnt,
                        const float* gcol_k, const float* row_max, const float* row_sum, int32_t act_prev, float* dh,
                        float* dw, float* db, float* workspace, void* stream) {
  using namespace gngf::tc;
  using namespace gngf::tc::sb;
  if (U <= 0 || T <= 0 || Kdim <= 0 || (Kdim % 8) != 0 || Kdim > 2 * BK || topk <= 0 || topk > GNGF_MAX_TOPK ||
      U >= (1ll << 31) || T >= (1ll << 31) || U != static_cast<int64_t>(lat.wx) * lat.wy)
    return GNGF_ERR_UNSUPPORTED;
  if (gcol_k && !cnt) return GNGF_ERR_INVALID_ARGUMENT;
  if (!h_planes || !w_planes || !h || !w || !bias || !utopv || !utopi || !dtv || !row_max || !row_sum || !dh || !dw ||
      !workspace)
    return GNGF_ERR_INVALID_ARGUMENT;
  if ((reinterpret_cast<uintptr_t>(bias) | reinterpret_cast<uintptr_t>(workspace) | reinterpret_cast<uintptr_t>(dh) |
       reinterpret_cast<uintptr_t>(dw) | reinterpret_cast<uintptr_t>(h) | reinterpret_cast<uintptr_t>(w)) & 15)
    return GNGF_ERR_INVALID_ARGUMENT;
  cudaStream_t st = gngf::as_stream(stream);
  // workspace: ascale (U, rounded up to 4) | m2neg (U, rounded up to 4) | spk (U, topk)
  const int64_t U4 = (U + 3) & ~int64_t(3);
  float* ascale = workspace;
  float* m2neg = workspace + U4;
  float* spk = workspace + 2 * U4;
  hpd_stream_bwd_prep_kernel<<<static_cast<unsigned>(gngf::ceil_div(U, 256)), 256, 0, st>>>(
      lat, U, topk, utopv, dtv, cnt, gcol_k, row_max, row_sum, ascale, m2neg, spk);
  gngf::note_launch();
  int rc = gngf::check_launch();
  if (rc) return rc;

  CUtensorMap map_h128, map_h64, map_w128, map_w64;
  if ((rc = make_plane_map(&map_h128, h_planes, U, Kdim, BM))) return rc;
  if ((rc = make_plane_map(&map_h64, h_planes, U, Kdim, SBN))) return rc;
  if ((rc = make_plane_map(&map_w128, w_planes, T, Kdim, BM))) return rc;
  if ((rc = make_plane_map(&map_w64, w_planes, T, Kdim, SBN))) return rc;
  if (cudaFuncSetAttribute(hpd_stream_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           static_cast<int>(SMEM_BYTES)) != cudaSuccess ||
      cudaFuncSetAttribute(hpd_stream_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           static_cast<int>(SMEM_BYTES)) != cudaSuccess)
    return gngf::check_launch();
  const int sms = gngf::sm_count();
  {  // dh (U, Kdim) += E W3 : X = h tiles, Y = W3 tiles
    const int64_t xt = gngf::ceil_div(U, BM), yt = gngf::ceil_div(T, SBN);
    const int ns = split_count(xt, yt);
    const int grid = static_cast<int>(std::min<int64_t>(xt * ns, sms));
    hpd_stream_bwd_kernel<false><<<grid, THREADS, SMEM_BYTES, st>>>(map_h128, map_w64, static_cast<int>(U),
                                                                   static_cast<int>(T), static_cast<int>(Kdim), ns, bias,
                                                                   m2neg, ascale, dh, nullptr);
    gngf::note_launch();
    if ((rc = gngf::check_launch())) return rc;
  }
  {  // dW3 (T, Kdim) += E^T h, db3 += E^T 1 : X = W3 tiles, Y = h tiles
    const int64_t xt = gngf::ceil_div(T, BM), yt = gngf::ceil_div(U, SBN);
    const int ns = split_count(xt, yt);
    const int grid = static_cast<int>(std::min<int64_t>(xt * ns, sms));
    hpd_stream_bwd_kernel<true><<<grid, THREADS, SMEM_BYTES, st>>>(map_w128, map_h64, static_cast<int>(T),
                                                                  static_cast<int>(U), static_cast<int>(Kdim), ns, bias,
                                                                  m2neg, ascale, dw, db);
    gngf::note_launch();
    if ((rc = gngf::check_launch())) return rc;
  }
  hpd_stream_bwd_sparse_kernel<<<static_cast<unsigned>(gngf::ceil_div(U * 32, 256)), 256, 0, st>>>(
      U, topk, static_cast<int>(Kdim), utopi, spk, h, w, act_prev, dh, dw, db);
  gngf::note_launch();
  return gngf::check_launch();
}

}  // extern "C"
