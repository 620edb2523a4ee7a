// K2 (tensor-core path): the wide HPD output layer  logits (U, T) = h3 (U, 128) W3^T + b3  -- the one real dense
// contraction of the path (models.py:80-88,105-106; 79 GFLOP per batch at the published configuration when
// evaluated per row, 134 MFLOP per row at T = 2^19) -- on the 5th-generation tensor cores:
//
//   * fp32 accuracy from bf16 tensor cores: every fp32 operand is split into three bf16 planes
//     x = hi + mid + lo (8 mantissa bits each) and the product is the six partial products of order <= 2
//     (hi.hi, hi.mid, mid.hi, hi.lo, lo.hi, mid.mid) accumulated in fp32 in TMEM.  Plain TF32 / BF16 operands
//     flip 1-10 % of the top-k selections; the 3-way split flips none.
//   * operands are staged by TMA (cp.async.bulk.tensor, 128-byte swizzle) into a 2-stage shared-memory ring,
//     tcgen05.mma (cta_group::1, kind::f16, M=128, N=128, K=16) is issued by one thread, accumulators live in
//     TMEM (2 x 128 columns, double-buffered so that the epilogue of tile i overlaps the MMAs of tile i+1),
//     the epilogue warps read them back with tcgen05.ld and add the bias.
//   * persistent: one CTA per SM walks the output tiles; warp 0 = TMA producer, warp 1 = MMA issuer / TMEM
//     owner, warps 2-5 = epilogue (warp w owns TMEM lanes 32*(w%4)...).
//
// gngf_split_bf16x3 produces the planes (one elementwise pass; the weight planes are reusable until the next
// optimizer step).
#include <algorithm>

#include <cuda_fp16.h>

#include "tc_common.cuh"

namespace gngf {
namespace tc {

constexpr int STAGES = 2;
constexpr uint32_t STAGE_BYTES = 2 * PLANES * PLANE_BYTES;            // 96 KB: A planes then B planes
constexpr uint32_t TMEM_COLS = 2 * BN;                                // two accumulators
constexpr int GEMM_EPI_WARPS = 8;                                     // two per TMEM lane quarter / scheduler
constexpr int THREADS = 64 + 32 * GEMM_EPI_WARPS;
constexpr uint32_t EPI_LD = 33;                                        // padded row of the epilogue staging tile
constexpr uint32_t EPI_BYTES = GEMM_EPI_WARPS * 32 * EPI_LD * 4;       // one 32x32 fp32 tile per epilogue warp
constexpr size_t SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024 /*align*/ + 256 /*barriers*/;


__device__ __forceinline__ float act_apply(float v, int act) {
  switch (act) {
    case GNGF_ACT_RELU: return fmaxf(v, 0.0f);
    case GNGF_ACT_LEAKY_RELU: return v > 0.0f ? v : v * 0.01f;
    case GNGF_ACT_SIGMOID: return 1.0f / (1.0f + expf(-v));
    default: return v;
  }
}

// C (M,N) = act(sum over plane pairs of A_i (M,K) B_j (N,K)^T + bias)
__global__ void __launch_bounds__(THREADS, 1)
    gemm_bf16x3_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                       const float* __restrict__ bias, float* __restrict__ C, int M, int N, int K, int act,
                       int accumulate, int k_splits, uint32_t idesc_arg) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* epi = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + EPI_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_b)) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull + s, 1);
      mbar_init(tempty + s, GEMM_EPI_WARPS);
    }
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

  // work item = (output tile, K split); with k_splits > 1 the partial tiles are summed with atomics into a
  // zero-initialised C (used when M x N alone cannot fill the chip, e.g. dX = dlogits W3 with K = T)
  const int num_m = (M + BM - 1) / BM, num_n = (N + BN - 1) / BN;
  const int kblocks_all = (K + BK - 1) / BK;
  const int kb_per = (kblocks_all + k_splits - 1) / k_splits;
  const int tiles = num_m * num_n * k_splits;

  if (warp == 0) {
    if (lane == 0) {  // ---- TMA producer ----
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int ks = t % k_splits, tt = t / k_splits;
        const int m0 = (tt / num_n) * BM, n0 = (tt % num_n) * BN;
        const int kb0 = ks * kb_per, kb1 = min(kblocks_all, kb0 + kb_per);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty + stage, phase ^ 1);
          mbar_expect_tx(full + stage, STAGE_BYTES);
          uint8_t* base = smem + stage * STAGE_BYTES;
          for (int pl = 0; pl < PLANES; ++pl) {
            tma_load_3d(base + pl * PLANE_BYTES, &map_a, kb * BK, m0, pl, full + stage);
            tma_load_3d(base + (PLANES + pl) * PLANE_BYTES, &map_b, kb * BK, n0, pl, full + stage);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    {  // ---- MMA issuer: the whole warp runs the loop, an elected lane issues (see tc_common.cuh) ----
      const uint32_t idesc = idesc_arg;   // (operand formats are a launch parameter: gngf_tc_gemm_set_formats)
      // partial products of order <= 2 of (hi + mid + lo)(hi + mid + lo)
      constexpr int pa[6] = {0, 0, 1, 0, 2, 1};
      constexpr int pb[6] = {0, 1, 0, 2, 0, 1};
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t ring_lo = umma_desc_lo(smem_u32(smem));
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int ks = t % k_splits;
        const int kb0 = ks * kb_per, kb1 = min(kblocks_all, kb0 + kb_per);
        mbar_wait(tempty + acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d = tmem_u + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full + stage, phase);
          tc_fence_after();
          const uint32_t a_lo = ring_lo + stage * (STAGE_BYTES >> 4);
          const uint32_t b_lo = a_lo + ((PLANES * PLANE_BYTES) >> 4);
#pragma unroll
          for (int pr = 0; pr < 6; ++pr) {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint64_t ad = umma_desc_pack(a_lo + ((pa[pr] * PLANE_BYTES + k * UMMA_K * 2) >> 4));
              const uint64_t bd = umma_desc_pack(b_lo + ((pb[pr] * PLANE_BYTES + k * UMMA_K * 2) >> 4));
              umma_bf16_lead(d, ad, bd, idesc, ((kb - kb0) | pr | k) != 0);
            }
          }
          umma_commit_lead(empty + stage);  // the stage may be refilled once these MMAs have read it
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_lead(tfull + acc);  // accumulator complete -> epilogue
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {  // ---- epilogue warps 2..9: (TMEM lane quarter, column half of the tile) ----
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
      const int ks = t % k_splits, tt = t / k_splits;
      const int m0 = (tt / num_n) * BM, n0 = (tt % num_n) * BN;
      const bool empty_split = ks * kb_per >= kblocks_all;
      mbar_wait(tfull + acc, acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;
      float* stage_tile = epi + (warp - 2) * 32 * EPI_LD;
#pragma unroll 1
      for (int c0 = half * (BN / 2); c0 < (half + 1) * (BN / 2); c0 += 32) {
        // thread = accumulator row (TMEM lane); transpose through shared memory so that a warp stores
        // 32 consecutive floats (one 128-byte line) of one output row per instruction
        uint32_t v[32];
        tmem_ld32(taddr + c0, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) stage_tile[lane * EPI_LD + j] = __uint_as_float(v[j]);
        __syncwarp();
        const int col = n0 + c0 + lane;
        if (col < N && !empty_split) {
          const float bcol = (bias && ks == 0) ? bias[col] : 0.0f;
          const int r_end = min(32, M - (m0 + q * 32));
          float* out = C + static_cast<int64_t>(m0 + q * 32) * N + col;
          if (k_splits > 1) {
            for (int r = 0; r < r_end; ++r) atomicAdd(out + static_cast<int64_t>(r) * N, stage_tile[r * EPI_LD + lane] + bcol);
          } else if (accumulate) {
            for (int r = 0; r < r_end; ++r) out[static_cast<int64_t>(r) * N] += stage_tile[r * EPI_LD + lane] + bcol;
          } else {
            for (int r = 0; r < r_end; ++r)
              out[static_cast<int64_t>(r) * N] = act_apply(stage_tile[r * EPI_LD + lane] + bcol, act);
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty + acc);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------
// K2 + K3 fused: streaming HPD output layer.  logits are never written: the epilogue keeps, per lattice node
// (= accumulator row = one thread), the online-softmax statistics (running max, compensated running sum of
// exp) and a sorted running top-K of the logits, while the MMA warp streams W3 tile by tile past a resident
// 128-row tile of h3.  Work items are (row tile, column split); splits exist so that a small lattice with a
// huge table still fills the chip, and are merged by hpd_stream_merge_kernel.
// Selection runs on the logits (monotone in the probabilities); ties keep the lower index.
// ---------------------------------------------------------------------------------------------------------
constexpr int KTOP = 8;                                  // largest K handled by the fused epilogue
constexpr int A_KBLOCKS = 2;                             // resident A tile: K <= 128
constexpr uint32_t A_BYTES = A_KBLOCKS * PLANES * PLANE_BYTES;   // 96 KB
constexpr uint32_t B_STAGE_BYTES = PLANES * PLANE_BYTES;         // 48 KB: one k-block of one column tile
constexpr size_t STREAM_SMEM_BYTES = A_BYTES + STAGES * B_STAGE_BYTES + 1024 + 256;
// Shared-memory plan of the streaming forward by number of split products: three planes (NPROD = 6) leave room for a
// 2-deep ring of 48 KB column-tile stages next to the 96 KB resident row tile; the two-plane pass (NPROD = 3) packs its
// operands (64 KB resident tile, 32 KB stages) and spends what it saves on a 4-deep ring -- 128 KB of W3 in flight per SM
// instead of 64 KB (every (row tile, column tile) pair pulls 64 KB through the L2 in ~1.1 us of MMA time: with two stages
// the stream was bound by TMA latency, not bandwidth).
template <int NPROD>
struct StreamPlan {
  static constexpr int NPL = NPROD == 3 ? 2 : PLANES;
  static constexpr int NSTAGES = NPROD == 3 ? 4 : 2;
  static constexpr uint32_t A_BYTES_ = A_KBLOCKS * NPL * PLANE_BYTES;
  static constexpr uint32_t B_STAGE_ = NPL * PLANE_BYTES;
  static constexpr size_t SMEM = A_BYTES_ + NSTAGES * B_STAGE_ + 1024 + 256;
};
constexpr int STREAM_EPI_WARPS = 16;                     // four per TMEM lane quarter: each takes 32 of a tile's 128 columns
constexpr int STREAM_COL_PARTS = STREAM_EPI_WARPS / 4;   // column parts per tile
constexpr int STREAM_THREADS = 64 + 32 * STREAM_EPI_WARPS;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float DEAD_CHUNK_LOG2 = -130.0f;               // 2^x flushes to zero (ex2.approx.ftz) with 4 binades to spare

struct RowTop {
  float v[KTOP];
  int i[KTOP];
};

// insert (z, n) into the list sorted by (value desc, index asc); callers feed candidates in increasing index order
__device__ __forceinline__ void top_insert(RowTop& t, int K, float z, int n) {
  float cz = z;
  int ci = n;
  bool carried = false;
#pragma unroll
  for (int k = 0; k < KTOP; ++k) {
    if (k < K) {
      const bool sw = carried ? (cz >= t.v[k]) : (cz > t.v[k]);
      const float tz = t.v[k];
      const int tn = t.i[k];
      t.v[k] = sw ? cz : tz;
      t.i[k] = sw ? ci : tn;
      cz = sw ? tz : cz;
      ci = sw ? tn : ci;
      carried |= sw;
    }
  }
}

// NPROD = 6: three planes, all products of order <= 2 (1.5e-6: selections exact on their own);
// NPROD = 3: two planes, hi.hi + hi.mid + mid.hi (1e-5) -- half the tensor-core work; the caller then keeps KTOP = 8
// candidates per row and re-evaluates them in fp32 (hpd_stream_refine_kernel), which restores exact selections and
// 1e-6 probabilities as long as the true top-K lie within the approximate top-8.
template <int NPROD>
__global__ void __launch_bounds__(STREAM_THREADS, 1)
    hpd_stream_fwd_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                          const float* __restrict__ bias, const float* __restrict__ a_scale,
                          const float* __restrict__ b_scale, int U, int T, int Kdim, int topk, int n_split, int no_skip,
                          float* __restrict__ part_max, float* __restrict__ part_sum, float* __restrict__ part_topv,
                          int* __restrict__ part_topi) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  using Plan = StreamPlan<NPROD>;
  constexpr int NPL = Plan::NPL;                 // planes loaded = planes laid out
  constexpr int STAGES = Plan::NSTAGES;          // (shadow the namespace-level constants of the plain GEMM)
  constexpr uint32_t A_BYTES = Plan::A_BYTES_, B_STAGE_BYTES = Plan::B_STAGE_;
  uint8_t* a_buf = smem;
  uint8_t* b_ring = smem + A_BYTES;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + A_BYTES + STAGES * B_STAGE_BYTES);
  uint64_t* a_empty = a_full + 1;
  uint64_t* b_full = a_empty + 1;
  uint64_t* b_empty = b_full + STAGES;
  // Accumulator ring.  With two buffers a tile's MMAs may only start when the epilogue of the tile before last has
  // drained its buffer, so the pair runs at (MMA + epilogue + hand-offs) / 2 per tile whenever epilogue + hand-offs
  // exceed the MMA time (measured: 2 407 cycles per tile against 1 578 of tensor work, both sides 27 % idle waiting for
  // each other).  The two-plane pass keeps its row tile in tensor memory, which leaves exactly one more 128-column
  // accumulator: 3 x 128 + 128 = 512 columns.
  constexpr int NACC = NPROD == 3 ? 3 : 2;
  uint64_t* tfull = b_empty + STAGES;
  uint64_t* tempty = tfull + NACC;
  uint64_t* at_full = tempty + NACC;             // A_IN_TMEM: the row tile has been copied into tensor memory
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(at_full + 1);
  // Two-plane pass: the resident row tile is the A operand from TENSOR MEMORY (columns 256..383: plane p at +64 p,
  // 8 columns per K = 16 step), copied there once per item by eight epilogue warps -- with both operands in shared
  // memory every 128x128x16 MMA reads 8 KB in 64 cycles, which is all the 128 B/clk an SM's shared memory delivers,
  // before the 64 KB per tile that TMA writes into the ring: 256 KB per tile = 2 048 cycles against 1 536 of tensor
  // work.  (Measured once before and rejected, 338 vs 343 ms, when the epilogue bounded the kernel; it no longer does.)
  constexpr bool A_IN_TMEM = NPROD == 3;
  constexpr uint32_t A_COL = NACC * BN;
  constexpr uint32_t TMEM_ALLOC = A_IN_TMEM ? 512u : TMEM_COLS;

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_b)) : "memory");
    mbar_init(a_full, 1);
    mbar_init(a_empty, 1);
    mbar_init(at_full, 8);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(b_full + s, 1);
      mbar_init(b_empty + s, 1);
    }
    for (int s = 0; s < NACC; ++s) {
      mbar_init(tfull + s, 1);
      mbar_init(tempty + s, STREAM_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TMEM_ALLOC)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int row_tiles = (U + BM - 1) / BM, col_tiles = (T + BN - 1) / BN;
  const int tiles_per_split = (col_tiles + n_split - 1) / n_split;
  const int items = row_tiles * n_split;
  const int kblocks = (Kdim + BK - 1) / BK;   // <= A_KBLOCKS

  if (warp == 0) {
    if (lane == 0) {  // ---- TMA producer ----
      int stage = 0;
      uint32_t phase = 0, a_phase = 0;
      for (int w = blockIdx.x; w < items; w += gridDim.x) {
        const int m0 = (w / n_split) * BM, sp = w % n_split;
        const int nt0 = sp * tiles_per_split, nt1 = min(col_tiles, nt0 + tiles_per_split);
        mbar_wait(a_empty, a_phase ^ 1);
        mbar_expect_tx(a_full, kblocks * NPL * PLANE_BYTES);
        for (int kb = 0; kb < kblocks; ++kb)
          for (int pl = 0; pl < NPL; ++pl)
            tma_load_3d(a_buf + (kb * NPL + pl) * PLANE_BYTES, &map_a, kb * BK, m0, pl, a_full);
        a_phase ^= 1;
        for (int nt = nt0; nt < nt1; ++nt) {
          for (int kb = 0; kb < kblocks; ++kb) {
            mbar_wait(b_empty + stage, phase ^ 1);
            mbar_expect_tx(b_full + stage, NPL * PLANE_BYTES);
            for (int pl = 0; pl < NPL; ++pl)
              tma_load_3d(b_ring + stage * B_STAGE_BYTES + pl * PLANE_BYTES, &map_b, kb * BK, nt * BN, pl, b_full + stage);
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    {  // ---- MMA issuer: the whole warp runs the loop, an elected lane issues (see tc_common.cuh) ----
      // NPROD = 3: two fp16 planes (11-bit mantissas, operands pre-scaled by powers of two: gngf_split_f16x2);
      // NPROD = 6: three bf16 planes
      constexpr uint32_t idesc = umma_idesc_fmt(BM, BN, NPROD == 3 ? 0 : 1, NPROD == 3 ? 0 : 1);
      constexpr int pa[6] = {0, 0, 1, 0, 2, 1};
      constexpr int pb[6] = {0, 1, 0, 2, 0, 1};
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t a_lo0 = umma_desc_lo(smem_u32(a_buf));
      const uint32_t b_lo0 = umma_desc_lo(smem_u32(b_ring));
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0, a_phase = 0;
      for (int w = blockIdx.x; w < items; w += gridDim.x) {
        const int sp = w % n_split;
        const int nt0 = sp * tiles_per_split, nt1 = min(col_tiles, nt0 + tiles_per_split);
        if (A_IN_TMEM) {
          mbar_wait(at_full, a_phase);
          a_phase ^= 1;
          tc_fence_after();
          if (elect_one()) mbar_arrive(a_empty);   // the shared-memory copy is free for the next item's row tile
        } else {
          mbar_wait(a_full, a_phase);
          a_phase ^= 1;
          tc_fence_after();
        }
        for (int nt = nt0; nt < nt1; ++nt) {
          mbar_wait(tempty + acc, acc_phase ^ 1);
          tc_fence_after();
          const uint32_t d = tmem_u + acc * BN;
          for (int kb = 0; kb < kblocks; ++kb) {
            mbar_wait(b_full + stage, phase);
            tc_fence_after();
            const uint32_t a_lo = a_lo0 + kb * ((NPL * PLANE_BYTES) >> 4);
            const uint32_t b_lo = b_lo0 + stage * (B_STAGE_BYTES >> 4);
#pragma unroll
            for (int pr = 0; pr < NPROD; ++pr) {
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; ++k) {
                const uint64_t bd = umma_desc_pack(b_lo + ((pb[pr] * PLANE_BYTES + k * UMMA_K * 2) >> 4));
                if (A_IN_TMEM) {
                  umma_bf16_ts_lead(d, tmem_u + A_COL + pa[pr] * 64 + kb * 32 + k * 8, bd, idesc, (kb | pr | k) != 0);
                } else {
                  const uint64_t ad = umma_desc_pack(a_lo + ((pa[pr] * PLANE_BYTES + k * UMMA_K * 2) >> 4));
                  umma_bf16_lead(d, ad, bd, idesc, (kb | pr | k) != 0);
                }
              }
            }
            umma_commit_lead(b_empty + stage);
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
          umma_commit_lead(tfull + acc);
          if (++acc == NACC) {
            acc = 0;
            acc_phase ^= 1;
          }
        }
        if (!A_IN_TMEM) umma_commit_lead(a_empty);  // every MMA that reads the resident A tile has completed
      }
    }
  } else {  // ---- epilogue warps 2..17: thread = (lattice node, 32-column part of every tile) ----
    // Sixteen warps, four per scheduler, so that the dependent chains of the online softmax overlap: with a single
    // epilogue warp per scheduler the epilogue took 7.5 k cycles per tile against 3.1 k for the MMAs (tensor pipe
    // 41 % active); with two per scheduler 5.4 k (57 %).
    const int q = warp & 3;                  // TMEM lane quarter this warp may access
    const int part = (warp - 2) >> 2;        // which 32 of the tile's 128 columns
    const int n_parts = STREAM_COL_PARTS * n_split;
    // the accumulator holds (h 2^sa)(W 2^sb)^T: undo the operands' power-of-two scales in the bias FMA
    const float inv = (a_scale ? __ldg(a_scale) : 1.0f) * (b_scale ? __ldg(b_scale) : 1.0f);
    int acc = 0;
    uint32_t acc_phase = 0, at_phase = 0;
    for (int w = blockIdx.x; w < items; w += gridDim.x) {
      const int m0 = (w / n_split) * BM, sp = w % n_split;
      const int nt0 = sp * tiles_per_split, nt1 = min(col_tiles, nt0 + tiles_per_split);
      if (A_IN_TMEM && part < 2) {
        // (every MMA of the previous item has completed: this warp has seen the tfull of its last tile.)  Row tile:
        // shared memory (TMA, 128-byte swizzle, [k-block][plane] x 16 KB) -> tensor memory; warp (q, part) copies
        // plane `part` of rows 32 q .. 32 q + 31
        mbar_wait(a_full, at_phase);
        at_phase ^= 1;
        const int row_l = q * 32 + lane;
        const uint8_t* xrow = a_buf + (row_l >> 3) * 1024 + (row_l & 7) * 128;
#pragma unroll 1
        for (int kb = 0; kb < A_KBLOCKS; ++kb) {
          uint32_t v[32];
          if (kb < kblocks) {
            const uint8_t* src = xrow + (kb * NPL + part) * PLANE_BYTES;
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) {
              const uint4 c = *reinterpret_cast<const uint4*>(src + ((ch ^ (row_l & 7)) << 4));
              v[ch * 4 + 0] = c.x; v[ch * 4 + 1] = c.y; v[ch * 4 + 2] = c.z; v[ch * 4 + 3] = c.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0u;
          }
          tmem_st32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + A_COL + part * 64 + kb * 32, v);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(at_full);
      }
      // running statistics in the base-2 domain: m2 = max(z) * log2(e), ssum = sum 2^(z*log2e - m2)
      float m2 = -INFINITY, ssum = 0.0f, comp = 0.0f;
      RowTop top;
#pragma unroll
      for (int k = 0; k < KTOP; ++k) {
        top.v[k] = -INFINITY;
        top.i[k] = 0x7fffffff;
      }
      for (int nt = nt0; nt < nt1; ++nt) {
        const int n0 = nt * BN + part * 32;
        // this thread's 32 bias values are requested BEFORE the wait for the accumulator: their L1/L2 latency was the
        // largest single stall of the epilogue (ncu source page: 14 % of all samples on the first add after the load)
        float4 bb[8];
        const bool full32 = n0 + 32 <= T;
        if (full32) {
          const float4* b4 = reinterpret_cast<const float4*>(bias + n0);   // n0 is a multiple of 32
#pragma unroll
          for (int j = 0; j < 8; ++j) bb[j] = __ldg(b4 + j);
        }
        mbar_wait(tfull + acc, acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + part * 32;
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          const int c0 = ch * 16;
          uint32_t v[16];
          tmem_ld16(taddr + c0, v);
          const int nb = n0 + c0;
          if (nb >= T) continue;
          float z[16];
          float cmax = -INFINITY;
          if (full32) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const float4 b = bb[ch * 4 + j / 4];
              z[j + 0] = fmaf(__uint_as_float(v[j + 0]), inv, b.x);
              z[j + 1] = fmaf(__uint_as_float(v[j + 1]), inv, b.y);
              z[j + 2] = fmaf(__uint_as_float(v[j + 2]), inv, b.z);
              z[j + 3] = fmaf(__uint_as_float(v[j + 3]), inv, b.w);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              z[j] = (nb + j < T) ? fmaf(__uint_as_float(v[j]), inv, __ldg(bias + nb + j)) : -INFINITY;
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) cmax = fmaxf(cmax, z[j]);
          const float cmax2 = cmax * LOG2E;
          if (cmax2 > m2) {  // rescale the running sum to the new maximum
            const float sc = exp2f(m2 - cmax2);
            ssum *= sc;
            comp *= sc;
            m2 = cmax2;
          }
          // A chunk whose largest logit sits more than 130 binades under the running maximum adds exactly nothing:
          // ex2.approx.ftz returns 0 below 2^-126.  That is the common case once the HPD is fed integer lattice
          // coordinates (models.py:416-418: logits O(1e2)..O(1e4), a one-hot softmax), and it takes the 16 MUFU.EX2 and
          // the summation chain out of an epilogue that otherwise bounds this kernel (tensor pipe 58 % active).
          // (!(x < y) rather than x >= y: a NaN logit keeps the full path and reaches the row sum as before.)
          if (no_skip || !(cmax2 - m2 < DEAD_CHUNK_LOG2)) {
            float cs[4] = {0.0f, 0.0f, 0.0f, 0.0f};   // four chains instead of one 16-deep FADD chain
#pragma unroll
            for (int j = 0; j < 16; ++j) cs[j & 3] += fast_exp2(fmaf(z[j], LOG2E, -m2));   // one FFMA + MUFU.EX2 per element
            const float csum = (cs[0] + cs[1]) + (cs[2] + cs[3]);
            // compensated accumulation of the chunk sums (rows are up to 2^22 columns long)
            const float y = csum - comp;
            const float tsum = ssum + y;
            comp = (tsum - ssum) - y;
            ssum = tsum;
          }
          if (cmax > top.v[KTOP - 1]) {  // (a sorted top-KTOP is kept whatever K is; its head is the top-K)
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (z[j] > top.v[KTOP - 1]) top_insert(top, KTOP, z[j], nb + j);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty + acc);
        if (++acc == NACC) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
      const int row = m0 + q * 32 + lane;
      if (row < U) {
        const int64_t o = static_cast<int64_t>(row) * n_parts + sp * STREAM_COL_PARTS + part;
        part_max[o] = m2;          // base-2 domain (see merge kernel)
        part_sum[o] = ssum;
#pragma unroll
        for (int k = 0; k < KTOP; ++k) {
          if (k < topk) {
            part_topv[o * topk + k] = top.v[k];
            part_topi[o * topk + k] = top.i[k];
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_ALLOC) : "memory");
  }
}

// insert (z, n) keeping (value desc, index asc) order for candidates that arrive in any index order
__device__ __forceinline__ void top_insert_any(RowTop& t, float z, int n) {
  float cz = z;
  int ci = n;
#pragma unroll
  for (int k = 0; k < KTOP; ++k) {
    const bool sw = (cz > t.v[k]) || (cz == t.v[k] && ci < t.i[k]);
    const float tz = t.v[k];
    const int tn = t.i[k];
    t.v[k] = sw ? cz : tz;
    t.i[k] = sw ? ci : tn;
    cz = sw ? tz : cz;
    ci = sw ? tn : ci;
  }
}

// merges the partial results of one row (n_parts = column splits x 4 column parts of every tile): global max / normaliser,
// top-K of the logits, p = exp(z - max) / sum.  The partial statistics are in the base-2 domain
// (m2 = max * log2 e, sum of 2^(z log2 e - m2)); the winners' probabilities use the accurate expf.
__global__ void __launch_bounds__(128)
    hpd_stream_merge_kernel(const float* __restrict__ part_max, const float* __restrict__ part_sum,
                            const float* __restrict__ part_topv, const int* __restrict__ part_topi, int U, int n_parts,
                            int topk, float* __restrict__ row_max, float* __restrict__ row_sum,
                            float* __restrict__ utopv, int* __restrict__ utopi, float* __restrict__ cand_v,
                            int* __restrict__ cand_i) {
  // cand_v / cand_i (U, KTOP), optional: the merged top-KTOP (approximate logits, indices) for hpd_stream_refine_kernel;
  // utopv / utopi are then not written (topk = number of entries stored per part)
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= U) return;
  float M2 = -INFINITY;
  for (int s = 0; s < n_parts; ++s) M2 = fmaxf(M2, part_max[static_cast<int64_t>(row) * n_parts + s]);
  float S = 0.0f;
  RowTop top;
#pragma unroll
  for (int k = 0; k < KTOP; ++k) {
    top.v[k] = -INFINITY;
    top.i[k] = 0x7fffffff;
  }
  for (int s = 0; s < n_parts; ++s) {
    const int64_t o = static_cast<int64_t>(row) * n_parts + s;
    const float pm = part_max[o];
    if (pm > -INFINITY) S += part_sum[o] * exp2f(pm - M2);
    for (int k = 0; k < topk; ++k) {
      const float z = part_topv[o * topk + k];
      if (z >= top.v[KTOP - 1]) top_insert_any(top, z, part_topi[o * topk + k]);
    }
  }
  const float M = top.v[0];                      // the row maximum itself (M2 == M * log2 e up to rounding)
  // normaliser relative to M: S is relative to 2^M2, so rescale by 2^(M2 - M log2 e) (== 1 up to rounding)
  S *= exp2f(M2 - M * LOG2E);
  if (row_max) row_max[row] = M;
  if (row_sum) row_sum[row] = S;
  if (cand_v) {
#pragma unroll
    for (int k = 0; k < KTOP; ++k) {
      cand_v[static_cast<int64_t>(row) * KTOP + k] = top.v[k];
      cand_i[static_cast<int64_t>(row) * KTOP + k] = top.i[k];
    }
    return;
  }
#pragma unroll
  for (int k = 0; k < KTOP; ++k) {
    if (k < topk) {
      float p = expf(top.v[k] - M) / S;
      if (p != p) p = 0.0f;
      utopv[static_cast<int64_t>(row) * topk + k] = p;
      utopi[static_cast<int64_t>(row) * topk + k] = top.i[k];
    }
  }
}

// Refinement of the two-plane (1e-5) streaming pass: warp per node.  The KTOP = 8 candidates' logits are re-evaluated
// in fp32 (z = <h[u,:], W[t,:]> + b[t]), re-ranked (value desc, index asc), and the softmax statistics are corrected for
// the exact maximum and the exact candidate terms: everything the caller sees about the selected slots is then as
// accurate as an fp32 evaluation of those logits; only the sum of the remaining (individually tiny) terms carries the
// 1e-5 of the approximate pass.  row_max / row_sum hold the approximate statistics on entry, the corrected ones on exit.
__global__ void __launch_bounds__(256)
    hpd_stream_refine_kernel(const float* __restrict__ h, const float* __restrict__ w, const float* __restrict__ bias, int U,
                             int T, int Kdim, int topk, const float* __restrict__ cand_v, const int* __restrict__ cand_i,
                             float* __restrict__ row_max, float* __restrict__ row_sum, float* __restrict__ utopv,
                             int* __restrict__ utopi) {
  const int64_t u = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) / 32;
  const int lane = threadIdx.x % 32;
  if (u >= U) return;
  const int c0 = lane * 4;   // Kdim <= 128, Kdim % 4 == 0
  float4 hv = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c0 < Kdim) hv = *reinterpret_cast<const float4*>(h + u * Kdim + c0);
  float z[KTOP], za[KTOP];
  int idx[KTOP];
#pragma unroll
  for (int k = 0; k < KTOP; ++k) {
    idx[k] = cand_i[u * KTOP + k];
    za[k] = cand_v[u * KTOP + k];
    float acc = 0.0f;
    if (idx[k] >= 0 && idx[k] < T && c0 < Kdim) {
      const float4 wv = __ldg(reinterpret_cast<const float4*>(w + static_cast<int64_t>(idx[k]) * Kdim + c0));
      acc = fmaf(hv.w, wv.w, fmaf(hv.z, wv.z, fmaf(hv.y, wv.y, hv.x * wv.x)));
    }
    acc = warp_sum(acc);
    z[k] = (idx[k] >= 0 && idx[k] < T) ? acc + __ldg(bias + idx[k]) : -INFINITY;   // (fewer than KTOP slots: T < 8)
  }
  // re-rank (insertion sort of 8, every lane redundantly)
  float zs[KTOP];
  int is[KTOP];
#pragma unroll
  for (int k = 0; k < KTOP; ++k) {
    zs[k] = -INFINITY;
    is[k] = 0x7fffffff;
  }
#pragma unroll
  for (int k = 0; k < KTOP; ++k) {
    float cz = z[k];
    int ci = idx[k];
#pragma unroll
    for (int j = 0; j < KTOP; ++j) {
      const bool sw = (cz > zs[j]) || (cz == zs[j] && ci < is[j]);
      const float tz = zs[j];
      const int tn = is[j];
      zs[j] = sw ? cz : tz;
      is[j] = sw ? ci : tn;
      cz = sw ? tz : cz;
      ci = sw ? tn : ci;
    }
  }
  const float Ma = row_max[u], M = zs[0];
  float S = row_sum[u] * expf(Ma - M);
#pragma unroll
  for (int k = 0; k < KTOP; ++k)
    if (z[k] > -INFINITY) S += expf(z[k] - M) - expf(za[k] - M);
  if (lane == 0) {
    row_max[u] = M;
    row_sum[u] = S;
  }
  if (lane < topk) {
    float zk = zs[0];
    int ik = is[0];
#pragma unroll
    for (int k = 1; k < KTOP; ++k)
      if (lane == k) {
        zk = zs[k];
        ik = is[k];
      }
    float p = expf(zk - M) / S;
    if (p != p) p = 0.0f;
    utopv[u * topk + lane] = p;
    utopi[u * topk + lane] = ik;
  }
}

// ---- two fp16 planes with a power-of-two scale ---------------------------------------------------------------------
// x 2^s = hi + mid (fp16 each, round-to-nearest), s chosen so that max |x| 2^s lies in [2^13, 2^14): 22 mantissa bits
// per element (two bf16 planes: 16), at the same tensor-core cost.  With two bf16 planes the recomputed logits of the
// streaming backward carry an absolute error of ~1.5e-5 sum|h w|, which exp() turns into a RELATIVE error of that
// size in every probability: measured at BASELINE.json configs[2] (T = 2^19, random-init HPD, |logit| up to 73)
// 1.8e-3 in dW3 against the 1e-4 bar, where a plain fp32 evaluation is at 1e-6.  fp16 has no exponent range to spare,
// hence the scale: exact (a power of two), per tensor, computed on the device (no host round trip), undone in the
// consumers' epilogues.  Elements more than 2^27 below the tensor's maximum lose relative (not absolute) precision.
__global__ void __launch_bounds__(256) absmax_kernel(const float* __restrict__ src, int64_t n, unsigned* __restrict__ out) {
  float m = 0.0f;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x * 4;
  for (int64_t i = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 4 <= n) {
      const float4 v = *reinterpret_cast<const float4*>(src + i);
      m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    } else {
      for (int64_t j = i; j < n; ++j) m = fmaxf(m, fabsf(src[j]));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.0f && m <= 3.0e38f) atomicMax(out, __float_as_uint(m));   // (non-negative floats order like their bits)
}

__device__ __forceinline__ int f16_shift_of(float absmax) {
  if (!(absmax > 0.0f) || !(absmax <= 3.0e38f)) return 0;
  return 13 - ilogbf(absmax);   // max |x| 2^s in [2^13, 2^14)
}

__global__ void __launch_bounds__(256) split_f16x2_kernel(const float* __restrict__ src, int64_t n,
                                                         float* __restrict__ scale, __half* __restrict__ planes) {
  const int s = f16_shift_of(__uint_as_float(*reinterpret_cast<const unsigned*>(scale + 1)));
  const float f = ldexpf(1.0f, s);
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i == 0) scale[0] = ldexpf(1.0f, -s);   // what a consumer multiplies its accumulator with
  if (i >= n) return;
  const float x = src[i] * f;
  const __half hi = __float2half_rn(x);
  const __half mid = __float2half_rn(x - __half2float(hi));
  planes[i] = hi;
  planes[n + i] = mid;
}

// x = hi + mid + lo, each bf16 (round-to-nearest): planes[0][i], planes[1][i], planes[2][i]
__global__ void __launch_bounds__(256) split_bf16x3_kernel(const float* __restrict__ src, int64_t n,
                                                          __nv_bfloat16* __restrict__ planes) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const float x = src[i];
  const __nv_bfloat16 hi = __float2bfloat16_rn(x);
  const float r1 = x - __bfloat162float(hi);
  const __nv_bfloat16 mid = __float2bfloat16_rn(r1);
  const float r2 = r1 - __bfloat162float(mid);
  const __nv_bfloat16 lo = __float2bfloat16_rn(r2);
  planes[i] = hi;
  planes[n + i] = mid;
  planes[2 * n + i] = lo;
}

// planes[pl][c][r] = split(src[r][c]) for src (rows, cols); the plane matrices are (cols, ld) with ld >= rows,
// columns r >= rows are zero-filled (ld keeps the row pitch a multiple of 16 bytes for TMA)
__global__ void __launch_bounds__(256) split_bf16x3_t_kernel(const float* __restrict__ src, int64_t rows, int64_t cols,
                                                            int64_t ld, __nv_bfloat16* __restrict__ planes) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;  // 32 x 8
  const int64_t r0 = static_cast<int64_t>(blockIdx.y) * 32, c0 = static_cast<int64_t>(blockIdx.x) * 32;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t r = r0 + ty + 8 * i, c = c0 + tx;
    tile[ty + 8 * i][tx] = (r < rows && c < cols) ? src[r * cols + c] : 0.0f;
  }
  __syncthreads();
  const int64_t plane = cols * ld;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t c = c0 + ty + 8 * i, r = r0 + tx;
    if (c < cols && r < ld) {
      const float x = tile[tx][ty + 8 * i];
      const __nv_bfloat16 hi = __float2bfloat16_rn(x);
      const float r1 = x - __bfloat162float(hi);
      const __nv_bfloat16 mid = __float2bfloat16_rn(r1);
      const __nv_bfloat16 lo = __float2bfloat16_rn(r1 - __bfloat162float(mid));
      planes[c * ld + r] = hi;
      planes[plane + c * ld + r] = mid;
      planes[2 * plane + c * ld + r] = lo;
    }
  }
}


}  // namespace tc
}  // namespace gngf

extern "C" {

int gngf_split_bf16x3(const float* src, int64_t n, uint16_t* planes, void* stream) {
  if (n < 0) return GNGF_ERR_INVALID_ARGUMENT;
  if (n == 0) return GNGF_OK;
  gngf::tc::split_bf16x3_kernel<<<static_cast<unsigned>(gngf::ceil_div(n, 256)), 256, 0, gngf::as_stream(stream)>>>(
      src, n, reinterpret_cast<__nv_bfloat16*>(planes));
  gngf::note_launch();
  return gngf::check_launch();
}

int gngf_split_f16x2(const float* src, int64_t n, uint16_t* planes, float* scale, void* stream) {
  if (n < 0 || !scale) return GNGF_ERR_INVALID_ARGUMENT;
  if (reinterpret_cast<uintptr_t>(src) & 15) return GNGF_ERR_INVALID_ARGUMENT;
  cudaStream_t st = gngf::as_stream(stream);
  if (cudaMemsetAsync(scale, 0, 2 * sizeof(float), st) != cudaSuccess) return gngf::check_launch();
  if (n == 0) return GNGF_OK;
  const unsigned blocks = static_cast<unsigned>(std::min<int64_t>(gngf::ceil_div(n, 1024), 8 * gngf::sm_count()));
  gngf::tc::absmax_kernel<<<blocks, 256, 0, st>>>(src, n, reinterpret_cast<unsigned*>(scale + 1));
  gngf::note_launch();
  int rc = gngf::check_launch();
  if (rc) return rc;
  gngf::tc::split_f16x2_kernel<<<static_cast<unsigned>(gngf::ceil_div(n, 256)), 256, 0, st>>>(
      src, n, scale, reinterpret_cast<__half*>(planes));
  gngf::note_launch();
  return gngf::check_launch();
}

int gngf_split_bf16x3_t(const float* src, int64_t rows, int64_t cols, int64_t ld, uint16_t* planes, void* stream) {
  if (rows <= 0 || cols <= 0 || ld < rows) return GNGF_ERR_INVALID_ARGUMENT;
  dim3 grid(static_cast<unsigned>(gngf::ceil_div(cols, 32)), static_cast<unsigned>(gngf::ceil_div(ld, 32)));
  if (grid.y > 65535) return GNGF_ERR_UNSUPPORTED;
  gngf::tc::split_bf16x3_t_kernel<<<grid, 256, 0, gngf::as_stream(stream)>>>(src, rows, cols, ld,
                                                                            reinterpret_cast<__nv_bfloat16*>(planes));
  gngf::note_launch();
  return gngf::check_launch();
}

// 16-bit operand formats of gngf_tc_gemm_bf16x3's planes (tcgen05 instruction descriptor: 0 = fp16, 1 = bf16; default
// bf16 x bf16).  The A and B fields are independent: bf16 x fp16 products are what the streaming backward uses for
// E (bf16, wide range) times fp16 operand planes (11-bit mantissas); tests/test_kernels_gpu.py checks the combination.
static int g_gemm_a_fmt = 1, g_gemm_b_fmt = 1;
int gngf_tc_gemm_set_formats(int32_t a_fmt, int32_t b_fmt) {
  if ((a_fmt != 0 && a_fmt != 1) || (b_fmt != 0 && b_fmt != 1)) return GNGF_ERR_INVALID_ARGUMENT;
  g_gemm_a_fmt = a_fmt;
  g_gemm_b_fmt = b_fmt;
  return GNGF_OK;
}

int gngf_tc_gemm_bf16x3(const uint16_t* a_planes, const uint16_t* b_planes, const float* bias, int64_t M, int64_t N,
                        int64_t K, int32_t act, int32_t accumulate, int32_t k_splits, float* C, void* stream) {
  using namespace gngf::tc;
  if (M <= 0 || N <= 0 || K <= 0 || (K % 8) != 0 || M >= (1ll << 31) || N >= (1ll << 31)) return GNGF_ERR_INVALID_ARGUMENT;
  if ((reinterpret_cast<uintptr_t>(a_planes) | reinterpret_cast<uintptr_t>(b_planes)) & 15) return GNGF_ERR_INVALID_ARGUMENT;
  CUtensorMap map_a, map_b;
  int rc = make_plane_map(&map_a, a_planes, M, K);
  if (rc) return rc;
  rc = make_plane_map(&map_b, b_planes, N, K);
  if (rc) return rc;
  if (cudaFuncSetAttribute(gemm_bf16x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           static_cast<int>(SMEM_BYTES)) != cudaSuccess)
    return gngf::check_launch();
  const int64_t out_tiles = gngf::ceil_div(M, BM) * gngf::ceil_div(N, BN);
  const int64_t kblocks = gngf::ceil_div(K, BK);
  if (k_splits <= 0)  // auto: split K only when the output tiles cannot fill the chip (C must then be zeroed)
    k_splits = 1;
  k_splits = static_cast<int32_t>(std::min<int64_t>(k_splits, kblocks));
  if (k_splits > 1 && (act != GNGF_ACT_NONE)) return GNGF_ERR_INVALID_ARGUMENT;
  const int64_t tiles = out_tiles * k_splits;
  const int grid = static_cast<int>(std::min<int64_t>(tiles, gngf::sm_count()));
  gemm_bf16x3_kernel<<<grid, THREADS, SMEM_BYTES, gngf::as_stream(stream)>>>(
      map_a, map_b, bias, C, static_cast<int>(M), static_cast<int>(N), static_cast<int>(K), act, accumulate, k_splits,
      umma_idesc_fmt(BM, BN, g_gemm_a_fmt, g_gemm_b_fmt));
  gngf::note_launch();
  return gngf::check_launch();
}

int64_t gngf_hpd_stream_workspace_floats(int64_t U, int64_t T, int32_t topk) {
  using namespace gngf::tc;
  const int64_t row_tiles = gngf::ceil_div(U, BM), col_tiles = gngf::ceil_div(T, BN);
  int64_t n_split = std::max<int64_t>(1, std::min<int64_t>(col_tiles, (2 * gngf::sm_count()) / row_tiles));
  return U * (STREAM_COL_PARTS * n_split) * (2 + 2 * static_cast<int64_t>(topk)) + 1;
}

int gngf_hpd_stream_fwd(const uint16_t* a_planes, const uint16_t* b_planes, const float* bias, int64_t U, int64_t T,
                        int64_t Kdim, int32_t topk, float* utopv, int32_t* utopi, float* row_max, float* row_sum,
                        float* workspace, void* stream) {
  using namespace gngf::tc;
  if (U <= 0 || T <= 0 || Kdim <= 0 || (Kdim % 8) != 0 || Kdim > A_KBLOCKS * BK || topk <= 0 || topk > KTOP ||
      topk > T || U >= (1ll << 31) || T >= (1ll << 31))
    return GNGF_ERR_UNSUPPORTED;
  CUtensorMap map_a, map_b;
  int rc = make_plane_map(&map_a, a_planes, U, Kdim);
  if (rc) return rc;
  rc = make_plane_map(&map_b, b_planes, T, Kdim);
  if (rc) return rc;
  const int64_t row_tiles = gngf::ceil_div(U, BM), col_tiles = gngf::ceil_div(T, BN);
  const int n_split =
      static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(col_tiles, (2 * gngf::sm_count()) / row_tiles)));
  const int64_t n_parts = static_cast<int64_t>(STREAM_COL_PARTS) * n_split;
  float* part_max = workspace;
  float* part_sum = part_max + U * n_parts;
  float* part_topv = part_sum + U * n_parts;
  int* part_topi = reinterpret_cast<int*>(part_topv + U * n_parts * topk);
  cudaStream_t st = gngf::as_stream(stream);
  if (cudaFuncSetAttribute(hpd_stream_fwd_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           static_cast<int>(StreamPlan<6>::SMEM)) != cudaSuccess)
    return gngf::check_launch();
  const int grid = static_cast<int>(std::min<int64_t>(row_tiles * n_split, gngf::sm_count()));
  hpd_stream_fwd_kernel<6><<<grid, STREAM_THREADS, StreamPlan<6>::SMEM, st>>>(map_a, map_b, bias, nullptr, nullptr, static_cast<int>(U),
                                                                     static_cast<int>(T), static_cast<int>(Kdim), topk,
                                                                     n_split, gngf::debug_no_skip(), part_max, part_sum, part_topv,
                                                                     part_topi);
  gngf::note_launch();
  rc = gngf::check_launch();
  if (rc) return rc;
  hpd_stream_merge_kernel<<<static_cast<unsigned>(gngf::ceil_div(U, 128)), 128, 0, st>>>(
      part_max, part_sum, part_topv, part_topi, static_cast<int>(U), static_cast<int>(n_parts), topk, row_max, row_sum,
      utopv, utopi, nullptr, nullptr);
  gngf::note_launch();
  return gngf::check_launch();
}

int64_t gngf_hpd_stream_refined_workspace_floats(int64_t U, int64_t T) {
  using namespace gngf::tc;
  // partial records with KTOP candidates each + the merged candidates (U, KTOP) x (value, index)
  return gngf_hpd_stream_workspace_floats(U, T, KTOP) + 2 * U * KTOP + 4;
}

int gngf_hpd_stream_fwd_refined(const uint16_t* a_planes, const float* a_scale, const uint16_t* b_planes,
                                const float* b_scale, const float* h, const float* w, const float* bias, int64_t U,
                                int64_t T, int64_t Kdim, int32_t topk, float* utopv, int32_t* utopi, float* row_max,
                                float* row_sum, float* workspace, void* stream) {
  using namespace gngf::tc;
  if (U <= 0 || T <= 0 || Kdim <= 0 || (Kdim % 8) != 0 || Kdim > A_KBLOCKS * BK || topk <= 0 || 2 * topk > KTOP ||
      topk > T || U >= (1ll << 31) || T >= (1ll << 31) || !h || !w || !row_max || !row_sum)
    return GNGF_ERR_UNSUPPORTED;
  if (!a_planes || !b_planes || !a_scale || !b_scale) return GNGF_ERR_INVALID_ARGUMENT;
  if ((reinterpret_cast<uintptr_t>(h) | reinterpret_cast<uintptr_t>(w)) & 15) return GNGF_ERR_INVALID_ARGUMENT;
  CUtensorMap map_a, map_b;   // two fp16 planes each (gngf_split_f16x2)
  int rc = make_plane_map(&map_a, a_planes, U, Kdim, BM, 2);
  if (rc) return rc;
  rc = make_plane_map(&map_b, b_planes, T, Kdim, BM, 2);
  if (rc) return rc;
  const int64_t row_tiles = gngf::ceil_div(U, BM), col_tiles = gngf::ceil_div(T, BN);
  const int n_split =
      static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(col_tiles, (2 * gngf::sm_count()) / row_tiles)));
  const int64_t n_parts = static_cast<int64_t>(STREAM_COL_PARTS) * n_split;
  float* part_max = workspace;
  float* part_sum = part_max + U * n_parts;
  float* part_topv = part_sum + U * n_parts;
  int* part_topi = reinterpret_cast<int*>(part_topv + U * n_parts * KTOP);
  float* cand_v = reinterpret_cast<float*>(part_topi + U * n_parts * KTOP);
  int* cand_i = reinterpret_cast<int*>(cand_v + U * KTOP);
  cudaStream_t st = gngf::as_stream(stream);
  if (cudaFuncSetAttribute(hpd_stream_fwd_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           static_cast<int>(StreamPlan<3>::SMEM)) != cudaSuccess)
    return gngf::check_launch();
  const int grid = static_cast<int>(std::min<int64_t>(row_tiles * n_split, gngf::sm_count()));
  hpd_stream_fwd_kernel<3><<<grid, STREAM_THREADS, StreamPlan<3>::SMEM, st>>>(map_a, map_b, bias, a_scale, b_scale, static_cast<int>(U),
                                                                     static_cast<int>(T), static_cast<int>(Kdim), KTOP,
                                                                     n_split, gngf::debug_no_skip(), part_max, part_sum, part_topv,
                                                                     part_topi);
  gngf::note_launch();
  if ((rc = gngf::check_launch())) return rc;
  hpd_stream_merge_kernel<<<static_cast<unsigned>(gngf::ceil_div(U, 128)), 128, 0, st>>>(
      part_max, part_sum, part_topv, part_topi, static_cast<int>(U), static_cast<int>(n_parts), KTOP, row_max, row_sum,
      nullptr, nullptr, cand_v, cand_i);
  gngf::note_launch();
  if ((rc = gngf::check_launch())) return rc;
  hpd_stream_refine_kernel<<<static_cast<unsigned>(gngf::ceil_div(U * 32, 256)), 256, 0, st>>>(
      h, w, bias, static_cast<int>(U), static_cast<int>(T), static_cast<int>(Kdim), topk, cand_v, cand_i, row_max, row_sum,
      utopv, utopi);
  gngf::note_launch();
  return gngf::check_launch();
}

}  // extern "C"
