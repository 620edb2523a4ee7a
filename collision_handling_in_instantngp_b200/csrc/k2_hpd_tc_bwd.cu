// K5c (tensor-core path): backward of the streaming HPD output layer -- softmax + top-k + Linear(Kd, T) with T up to
// 2^22 slots (models.py:80-88, 105-123 and DifferentiableTopk.backward, models.py:21-42) -- without ever writing the
// (U, T) logits, probabilities or dlogits.
//
// In top-k-only mode the adjoint of the probabilities is non-zero at the K selected slots only, so per lattice node r
//     dlogit[r, t] = p[r, t] (G[r, t] - <G,p>_r) = -<G,p>_r p[r, t]           for t not selected,
//                                                 p_k (g_k - <G,p>_r)          for t = t_k,     p = exp(z - max_r) / sum_r .
// The selected slots are a gather / scatter of K rows per node in fp32 (hpd_stream_bwd_sparse_kernel) and are MASKED OUT
// of the dense part.  (The first version let the dense part cover every t and added p_k g_k on top: with a peaked softmax
// -- p_k -> 1, the regime of a random-init HPD fed integer lattice coordinates -- the two terms cancel to a result
// 1/(1 - p_k) times smaller, and the dense term's independent rounding, ~3e-6 from the exp2 argument alone, was measured
// as 1.6e-3 in dW3 at the size of BASELINE.json configs[2]; computing p_k (g_k - <G,p>) as ONE product is what a plain
// fp32 evaluation does.)  The dense part is an attention-shaped pair of products with E[r, t] = a_r exp(z[r, t] - max_r)
// for the unselected t, a_r = -<G,p>_r / sum_r:
//     dh  (U, Kd) = E   W3           dW3 (T, Kd) += E^T h           db3 (T) += E^T 1
// Both are instances of ONE kernel (hpd_stream_bwd_kernel<DW>):
//     X (128 x Kd) resident, Y tiles (128 x Kd) streamed by TMA through a 3-slot ring (X travels through it once per item)
//     S = X Y^T   : A = the X tile, copied once per item into TENSOR MEMORY (bf16 pairs per column), B = Y from shared
//                   memory (K-major); accumulator in TMEM, double-buffered
//     E = rowscale_i colscale_j exp2(S log2e + rowoff_i + coloff_j)  : 8 epilogue warps read S with tcgen05.ld and write
//                   the two bf16 planes of E back IN PLACE over S with tcgen05.st
//     O (128 x Kd) += E Y : A = E from tensor memory, B = the SAME Y tile read MN-major (no transposed copy of anything),
//                   accumulated in TMEM over all Y tiles
//   DW = false:  X = h tile,  Y = W3 tiles:  O = dh tile          (rows carry (a_r, -max_r log2e), columns the bias)
//   DW = true :  X = W3 tile, Y = h tiles :  O = dW3 tile, db3    (rows carry the bias, columns (a_r, -max_r log2e))
// Neither product reads its A operand from shared memory: the only shared-memory operand traffic is the Y tile (4 KB
// per 128x128x16 MMA = 64 B/clk).  History: E through shared memory with 64-row Y tiles reached 67 % of the measured
// bf16 peak (the N = 64 first product was bound by 6 KB of operand reads per MMA), X as a TMEM operand 71 %, E in TMEM
// with 128-row tiles 85 %.  TMEM: S/E 2 x 128 columns, O 128, X 128 = all 512.
// Operands are two fp16 planes (hi, mid) of the power-of-two-scaled fp32 values (gngf_split_f16x2: 22 mantissa bits)
// and each product is hi.hi + hi.mid + mid.hi.  (Two bf16 planes -- 16 bits -- were measured at the size of BASELINE.json
// configs[2] to put 1.8e-3 into dW3: the recomputed logits' absolute error becomes a relative error of every E.)
// tcgen05.mma kind::f16 wants A and B in the SAME 16-bit format (bf16 x fp16 traps as an illegal instruction), so E is
// fp16 as well and its range is managed by exponent offsets inside the exp2 argument:
//   DW = false:  E' = 2^12 exp(z - max_r)            the row factor a_r is applied to the O tile when it is flushed;
//   DW = true :  E''= sgn(a_r) 2^(log2|a_r| + sa) exp(z - max_r),   sa = 13 - ceil(max_r log2|a_r|)  (device scalar)
// and the flush multiplies O by the inverse powers of two (and the operand scale of Y).  MMA issue order S(j+1), O(j)
// keeps the tensor pipe busy while the epilogue turns S(j) into E(j).
//
// Exact zeros are not computed.  The HPD sees integer lattice coordinates, its logits are O(1e2)..O(1e4) and the softmax is
// one-hot: exp2 of an argument below -126 IS zero (ex2.approx.ftz), and so are both fp16 planes of anything below 2^-25.
//   * an epilogue thread bounds the 32 arguments of a chunk by max(v) k1 + (r_off + max(off)) (monotone in fp32) and writes
//     zero planes when that is under -130; the exact per-element maximum decides the chunks the bound leaves open;
//   * a tile whose E has no non-zero fp16 entry (warp votes -> a stamp in shared memory, released by the e_full arrive) does
//     not issue O += E Y; its buffers are released by plain arrives.  98-99 % of the tiles at BASELINE.json configs[3].
//   * SCREENING: the first of the three split products of S, X_hi Y_hi^T, is within errv[u] = 2^-9 |h_u| max|W3 row| of
//     their sum (hpd_stream_bwd_errv_kernel).  A work item in screening mode issues only that product per tile; the epilogue
//     threads bound their arguments from it (+ errv), vote (v_full, a named barrier among themselves), and a tile nobody
//     objects to is dropped there and then -- nothing written, nothing more issued.  The other tiles get the two remaining
//     products into the same accumulator (s2_full) and the exact epilogue.  The mode is adaptive per item: see the issuer.
// GNGF_DEBUG_NO_SKIP=1 turns all of it off (tests: the results do not change -- dh bit for bit); gngf_hpd_stream_bwd_stats
// counts tiles / tiles with all three logit products / tiles that issued their second product (bench.py: executed FLOPs).
#include <algorithm>

#include <cuda_fp16.h>

#include "tc_common.cuh"

namespace gngf {
namespace tc {
namespace sb {

constexpr int SBN = 128;                                  // rows of a streamed Y tile (= columns of S and E)
constexpr int NP = 2;                                     // planes used: hi, mid
constexpr int RING = 3;                                   // shared-memory ring of (X | Y) tiles
constexpr int EPI_WARPS = 8;                              // (TMEM lane quarter) x (64-column half of the S tile)
constexpr int THREADS = 64 + 32 * EPI_WARPS;
constexpr uint32_t TP_BYTES = 128 * 128;                  // one (k-block, plane) of a tile: 128 rows x 64 bf16
constexpr uint32_t TILE_BYTES = 2 * NP * TP_BYTES;        // 64 KB: [k-block][plane]
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t SE_COL = 0;                            // two S / E buffers of 128 columns
constexpr uint32_t O_COL = 256;                           // O accumulator, 128 columns
constexpr uint32_t X_COL = 384;                           // resident X tile: plane p at [X_COL + 64 p, +64) (bf16 pairs)
constexpr uint32_t CTRL_BYTES = 512;                      // barriers, flags, per-chunk maxima
constexpr size_t SMEM_BYTES = RING * TILE_BYTES + 1024 /*align*/ + CTRL_BYTES + 2 * 128 * 16 /*selected-slot masks*/ +
                              2 * 256 * 4 /*per-column offset / scale of two Y tiles*/;
constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  const __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_f16x2(uint32_t p) {
  return __half22float2(*reinterpret_cast<const __half2*>(&p));
}

// device scalars shared by the prep kernel and the two dense passes (consts[] in the workspace)
constexpr int C_K1 = 0;        // log2(e) / (scale_h scale_w): turns the scaled accumulator into a base-2 logit
constexpr int C_OUT_DH = 1;    // 2^-12 / scale_w
constexpr int C_OUT_DW = 2;    // 2^-sa / scale_h
constexpr int C_SA = 3;        // sa (as a float): exponent offset of the DW = true pass
constexpr int C_MAXLA = 4;     // scratch: max over rows of log2|a_r| as ordered-int bits
constexpr int C_MAXW = 5;      // largest row norm of W3 (bits of a non-negative float: integer atomicMax)
constexpr float SCREEN_ERR = 1.0f / 512.0f;   // |S - S_hihi| <= 2^-10 (1 + 2^-11) |x||y| for fp16-rounded planes; x2 margin
constexpr float E_SHIFT = 12.0f;
constexpr int SB_MAX_K = 8;    // selected slots per node on the streaming path (the forward's KTOP)
constexpr float DEAD_CHUNK_LOG2 = -130.0f;   // exp2 arguments below this give exactly 0 (ex2.approx.ftz flushes under 2^-126)

// x_rows / y_rows: number of valid rows of X / Y (U or T); Kdim <= 128.
// DW = false: X = h tile, rows carry (a, -max log2e), columns the bias;  DW = true: X = W3 tile, the other way round.
//
// Data flow per (X tile, Y tile):  S = X Y^T (A = X from tensor memory, B = Y from shared memory, K-major)
//   -> epilogue: E = scale * exp2(S log2e + offsets), two bf16 planes written IN PLACE over S in tensor memory
//   -> O += E Y (A = E from tensor memory, B = the same Y tile read MN-major).
// Neither product reads its A operand from shared memory, so the only shared-memory operand traffic is the Y tile
// (4 KB per 128x128x16 MMA = 64 B/clk), and the ring holds 128-row tiles (X travels through it once per item).
// [0] tiles, [1] tiles that needed the two remaining logit products after the screening product, [2] tiles whose second
// product O += E Y was issued -- DW = false; [3..5] the same for DW = true.  Accumulated over launches until
// gngf_hpd_stream_bwd_stats() reads and clears them: what bench.py needs to state the EXECUTED tensor work of passes
// whose products are data dependent.
__device__ unsigned long long g_stream_stats[6];

template <bool DW>
__global__ void __launch_bounds__(THREADS, 1)
    hpd_stream_bwd_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_y,
                          int x_rows, int y_rows, int Kdim, int n_split, const float* __restrict__ bias,
                          const float* __restrict__ m2neg, const float* __restrict__ ascale,
                          const float* __restrict__ errv, const float* __restrict__ consts,
                          const int* __restrict__ utopi, int topk, int no_skip,
                          float* __restrict__ out, float* __restrict__ dbias) {
  // no_skip (GNGF_DEBUG_NO_SKIP=1, tests only): every chunk takes the full path and every tile issues its second product
  // -- the reference point for "skipping changes nothing"
  // errv (nodes): bound on |S - S_hihi| of a node against ANY slot, in accumulator units (hpd_stream_bwd_errv_kernel)
  // utopi (nodes, topk): the selected slots, masked out of E (DW = false: nodes are the X rows; DW = true: the Y rows)
  // m2neg / ascale: DW = false: per ROW  -max_r log2e          / a_r (applied at the flush);
  //                 DW = true : per COLUMN -max_r log2e + log2|a_r| / sgn(a_r)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ring = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + RING * TILE_BYTES);
  uint64_t* r_full = bars;                 // [RING] TMA landed
  uint64_t* r_empty = r_full + RING;       // [RING] slot may be refilled
  uint64_t* s_full = r_empty + RING;       // [2] S complete in tensor memory
  uint64_t* e_full = s_full + 2;           // [2] E written over S
  uint64_t* e_empty = e_full + 2;          // [2] second product done: the S / E buffer is free
  uint64_t* o_full = e_empty + 2;
  uint64_t* o_empty = o_full + 1;
  uint64_t* xt_full = o_empty + 1;         // X tile copied into tensor memory
  uint64_t* v_full = xt_full + 1;          // [2] the epilogue warps have voted on the screening product
  uint64_t* s2_full = v_full + 2;          // [2] the two remaining logit products have been added (live tiles only)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s2_full + 2);
  volatile uint32_t* o_valid_s = tmem_slot + 1;   // the item's O accumulator holds something (else: nothing to flush)
  volatile uint32_t* mode_s = tmem_slot + 2;      // this item screens its tiles with one product (1) or issues all three (0)
  // live_s[buf] = 1 + (running number of the last tile whose E in S / E buffer `buf` has a non-zero entry): written by
  // the epilogue warps before they arrive on e_full, read by the MMA issuer after it (never cleared: a stale stamp
  // cannot equal the current tile's)
  volatile uint32_t* live_s = reinterpret_cast<volatile uint32_t*>(bars) + 48;   // byte 192 of the 256-byte barrier block
  volatile uint32_t* live1_s = live_s + 2;   // the same stamp for "the screening product could not rule this tile out"
  float* cmax_s = reinterpret_cast<float*>(bars) + 52;   // [2 buffers][4 chunks of 32 columns]: largest column offset
  float* emax_s = reinterpret_cast<float*>(bars) + 64;   // [2 buffers][4 chunks]: largest errv of the chunk's columns (DW)
  unsigned* kill_s = reinterpret_cast<unsigned*>(ring + RING * TILE_BYTES + CTRL_BYTES);   // [2 buffers][128 slots][4 words]
  float* col_s = reinterpret_cast<float*>(kill_s + 2 * 128 * 4);                    // [2 buffers][offset 128 | scale 128]

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_x)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_y)) : "memory");
    for (int s = 0; s < RING; ++s) {
      mbar_init(r_full + s, 1);
      mbar_init(r_empty + s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(s_full + s, 1);
      mbar_init(e_full + s, EPI_WARPS);
      mbar_init(e_empty + s, 1);
    }
    mbar_init(o_full, 1);
    mbar_init(o_empty, EPI_WARPS);
    mbar_init(xt_full, EPI_WARPS);
    for (int s = 0; s < 2; ++s) {
      mbar_init(v_full + s, EPI_WARPS);
      mbar_init(s2_full + s, 1);
    }
    live_s[0] = 0u;
    live_s[1] = 0u;
    live1_s[0] = 0u;
    live1_s[1] = 0u;
    *o_valid_s = 0u;
    *mode_s = 1u;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp >= 2)
    for (int i = threadIdx.x - 64; i < 2 * 128 * 4; i += 32 * EPI_WARPS) kill_s[i] = 0u;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int x_tiles = (x_rows + BM - 1) / BM, y_tiles = (y_rows + SBN - 1) / SBN;
  const int tiles_per_split = (y_tiles + n_split - 1) / n_split;
  const int items = x_tiles * n_split;
  const int kblocks = (Kdim + BK - 1) / BK;   // 1 or 2
  const int n2 = kblocks * BK;                // N of the second product (columns of O)
  // Item w -> (X tile, split of the Y stream).  DW = false: the Y stream is W3 (L2-resident), consecutive items walk the
  // splits of one X tile.  DW = true: the Y stream is the h planes of every node (15 GB at BASELINE.json configs[3]) and
  // is streamed once PER X TILE, so consecutive items -- the CTAs that run at the same time -- take the SAME split of
  // different X tiles: one HBM read of an h tile serves ~128 CTAs out of the L2 instead of ~4.
  auto item_x = [&](int w) { return DW ? w % x_tiles : w / n_split; };
  auto item_split = [&](int w) { return DW ? w / x_tiles : w % n_split; };

  if (warp == 0) {
    if (lane == 0) {  // ---- TMA producer: per item the X tile, then its Y tiles, all through one ring ----
      uint32_t n = 0;   // ring entries produced
      auto load_tile = [&](const CUtensorMap* map, int row0) {
        const uint32_t slot = n % RING, ph = (n / RING) & 1;
        mbar_wait(r_empty + slot, ph ^ 1);
        mbar_expect_tx(r_full + slot, kblocks * NP * TP_BYTES);
        uint8_t* dst = ring + slot * TILE_BYTES;
        for (int kb = 0; kb < kblocks; ++kb)
          for (int pl = 0; pl < NP; ++pl) tma_load_3d(dst + (kb * NP + pl) * TP_BYTES, map, kb * BK, row0, pl, r_full + slot);
        ++n;
      };
      for (int w = blockIdx.x; w < items; w += gridDim.x) {
        const int m0 = item_x(w) * BM, sp = item_split(w);
        const int t0 = sp * tiles_per_split, t1 = min(y_tiles, t0 + tiles_per_split);
        if (t0 >= t1) continue;   // empty split (every role skips it)
        load_tile(&map_x, m0);
        for (int t = t0; t < t1; ++t) load_tile(&map_y, t * SBN);
      }
    }
  } else if (warp == 1) {
    {  // ---- MMA issuer: the whole warp runs the loop, an elected lane issues (see tc_common.cuh) ----
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);   // provably warp-uniform copy
      constexpr uint32_t idesc1 = umma_idesc_fmt(BM, SBN, 0, 0);                    // fp16 x fp16
      const uint32_t idesc2 = umma_idesc_fmt(BM, n2, 0, 0) | UMMA_B_MN_MAJOR;
      const uint32_t y_lo = umma_desc_lo(smem_u32(ring));                      // K-major view (first product)
      const uint32_t y_lo_mn = umma_desc_lo(smem_u32(ring), NP * TP_BYTES);    // MN-major view (second product)
      // Ring position and S / E buffer of the next S product and of the next O product as wrap-around counters: they
      // stay in the uniform datapath, and so do the 24 descriptors derived from them.  (`entry % RING` is an IMAD.HI in
      // the vector datapath: every descriptor then took an R2UR.BROADCAST to reach the UTCHMMA -- 424 instructions per
      // tile in this warp, ~1 400 cycles against 1 536 of tensor work, serial with the wait for the epilogue.)
      uint32_t rs = 0, rph = 0;                // ring slot / phase of the next entry to consume
      uint32_t os = 0;                         // ring slot of the next O product's Y tile
      uint32_t sbuf = 0, sph = 0;              // S / E buffer and phase of the next S product
      uint32_t obuf = 0, oph = 0;              // ... of the next O product
      uint32_t it2 = 0;                        // running tile counter of the O side (the epilogue's liveness stamps)
      uint32_t p2a = 0, p2b = 0;               // phases of e_full of buffer 0 / 1 (not every tile gets that far)
      uint32_t pva = 0, pvb = 0;               // phases of v_full of buffer 0 / 1 (screening items only)
      uint32_t n_s2 = 0, n_live = 0;           // tiles that needed all three logit products / issued their second product
      // Screening pays where it rules tiles out (a one-hot softmax: BASELINE.json configs[3]) and costs a round trip
      // through the epilogue per tile where it does not (logits within +-73 at configs[2]: nothing is 130 binades down).
      // Each item decides from the items before: screening goes on while it rules out more than half of an item's tiles;
      // after TWO failures in a row (one alone means little: at configs[3] a CTA's consecutive items are row tiles 148
      // apart, and only those near the lattice origin have small logits) `backoff` plain items follow (8, doubling up
      // to 64 with every failed probe) before screening is tried again.
      bool screen = no_skip == 0;
      uint32_t c_tiles = 0, c_s2 = 0, c_live = 0;   // the previous item's tiles / not ruled out / second product issued
      uint32_t backoff = 8, plain_left = 0, fails = 0;
      uint32_t x_phase = 0, o_phase = 0;
      // one logit product (planes pa x pb of X and Y) into S / E buffer d: 8 MMAs at Kdim = 128
      auto logit_product = [&](uint32_t d, uint32_t yb, int pa, int pb, bool first) {
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          if (kb < kblocks) {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              // A: resident X tile in tensor memory, 8 columns per K = 16 step, 32 per k-block, 64 per plane
              const uint32_t at = tmem_u + X_COL + pa * 64 + kb * 32 + k * 8;
              const uint64_t bd = umma_desc_pack(yb + (((kb * NP + pb) * TP_BYTES + k * UMMA_K * 2) >> 4));
              umma_bf16_ts_lead(d, at, bd, idesc1, !first || (kb | k) != 0);
            }
          }
        }
      };
      for (int w = blockIdx.x; w < items; w += gridDim.x) {
        const int sp = item_split(w);
        const int t0 = sp * tiles_per_split, t1 = min(y_tiles, t0 + tiles_per_split);
        const int nt = t1 - t0;
        if (nt <= 0) continue;
        // the epilogue warps copy the X tile (ring entry nent) into tensor memory; its slot is free afterwards
        mbar_wait(xt_full, x_phase);
        x_phase ^= 1;
        tc_fence_after();
        if (elect_one()) mbar_arrive(r_empty + rs);
        if (++rs == RING) {
          rs = 0;
          rph ^= 1;
        }
        os = rs;                               // ring slot of this item's first Y tile
        bool o_init = false;                   // the item's O accumulator has been written
        if (c_tiles > 0 && no_skip == 0) {
          if (screen) {
            if (2 * c_s2 < c_tiles) {
              backoff = 8;
              fails = 0;
            } else if (++fails >= 2) {
              screen = false;
              plain_left = backoff;
              backoff = min(64u, backoff * 2);
            }
          } else if (--plain_left == 0) {
            screen = true;   // (fails stays >= 2: one more failure sends the next items back to plain at once)
          }
        }
        c_tiles = c_s2 = c_live = 0;
        if (elect_one()) {                     // (read by the epilogue warps after the item's first s_full)
          *mode_s = screen ? 1u : 0u;
          __threadfence_block();
        }
        for (int j = 0; j <= nt; ++j) {
          if (j < nt) {
            // SCREENING product S1(j) = X_hi Y_hi^T -- one of the three split products.  It differs from the full S by at
            // most errv (hpd_stream_bwd_errv_kernel), and at a one-hot softmax that is enough to rule out almost every
            // tile: S1 + errv so far below the row maximum that exp2 underflows.  The other two products are issued
            // only for the tiles the epilogue could not rule out.
            const uint32_t buf = sbuf, ph = sph;
            const uint32_t slot = rs;
            mbar_wait(r_full + slot, rph);
            mbar_wait(e_empty + buf, ph ^ 1);    // the tile before last has finished with this buffer
            tc_fence_after();
            logit_product(tmem_u + SE_COL + buf * SBN, y_lo + slot * (TILE_BYTES >> 4), 0, 0, true);
            if (!screen) {                       // a plain item: all three products at once
              logit_product(tmem_u + SE_COL + buf * SBN, y_lo + slot * (TILE_BYTES >> 4), 0, 1, false);
              logit_product(tmem_u + SE_COL + buf * SBN, y_lo + slot * (TILE_BYTES >> 4), 1, 0, false);
            }
            umma_commit_lead(s_full + buf);
            if (++rs == RING) {
              rs = 0;
              rph ^= 1;
            }
            sph ^= sbuf;                         // (the phase flips when the buffer index wraps 1 -> 0)
            sbuf ^= 1;
          }
          if (j >= 1) {  // resolve tile j - 1
            const uint32_t buf = obuf, ph = oph;
            const uint32_t slot = os;
            (void)ph;
            bool live1 = true;                   // a plain item: every tile has all three products already
            if (screen) {
              mbar_wait(v_full + buf, buf ? pvb : pva);   // the eight epilogue warps have looked at S1
              if (buf) pvb ^= 1; else pva ^= 1;
              // (votes: provably warp-uniform, which keeps every descriptor below in the uniform datapath)
              live1 = __any_sync(0xffffffffu, live1_s[buf] == it2 + 1u);
            }
            if (j == 1) {
              mbar_wait(o_empty, o_phase ^ 1);   // the previous item's O has been read out
              o_phase ^= 1;
            }
            bool issued_o = false;
            ++c_tiles;
            if (live1) {
              ++n_s2;
              ++c_s2;
              if (screen) {
                tc_fence_after();
                const uint32_t d1 = tmem_u + SE_COL + buf * SBN;
                const uint32_t yk = y_lo + slot * (TILE_BYTES >> 4);
                logit_product(d1, yk, 0, 1, false);   // + X_hi Y_mid^T
                logit_product(d1, yk, 1, 0, false);   // + X_mid Y_hi^T
                umma_commit_lead(s2_full + buf);
              }
              const uint32_t p2 = buf ? p2b : p2a;
              mbar_wait(e_full + buf, p2);          // E written over S by the epilogue
              if (buf) p2b ^= 1; else p2a ^= 1;
              tc_fence_after();
              // An E tile without a single non-zero fp16 entry adds nothing to O: its three products are not issued.
              const bool live2 = no_skip != 0 || __any_sync(0xffffffffu, live_s[buf] == it2 + 1u);
              if (live2) {
                const uint32_t eb = tmem_u + SE_COL + buf * SBN;
                const uint32_t yb = y_lo_mn + slot * (TILE_BYTES >> 4);
                const uint32_t d = tmem_u + O_COL;
#pragma unroll
                for (int pr = 0; pr < 3; ++pr) {
                  const int pa = pr >> 1, pb = pr & 1;
#pragma unroll
                  for (int k = 0; k < SBN / UMMA_K; ++k) {
                    // A: E plane pa in tensor memory: S columns 16k.. live at packed columns 64 (k / 4) + 32 pa + 8 (k % 4)
                    const uint32_t at = eb + 64 * (k >> 2) + 32 * pa + 8 * (k & 3);
                    // B: Y plane read MN-major: N = feature index (64 contiguous per k-block, k-blocks NP*TP_BYTES apart),
                    //    K = streamed index (rows of 128 bytes, 16 rows = 2048 bytes per step)
                    const uint64_t bd = umma_desc_pack(yb + ((pb * TP_BYTES + k * UMMA_K * 128) >> 4));
                    umma_bf16_ts_lead(d, at, bd, idesc2, o_init || (pr | k) != 0);
                  }
                }
                o_init = true;
                issued_o = true;
                ++n_live;
                ++c_live;
              }
            }
            if (live1) {
              // (commits: behind the products that read this S / E buffer and this Y tile)
              umma_commit_lead(e_empty + buf);
              umma_commit_lead(r_empty + slot);
            } else if (elect_one()) {
              // nothing reads S1(j-1) or Y_{j-1} any more (the epilogue has read S1; its product completed before that):
              // both are released at once instead of behind the S1(j) product a commit would wait for
              mbar_arrive(e_empty + buf);
              mbar_arrive(r_empty + slot);
            }
            (void)issued_o;
            ++it2;
            if (++os == RING) os = 0;
            oph ^= obuf;
            obuf ^= 1;
          }
        }
        // o_full: behind the item's last O product if there was one; else there is nothing to flush and the epilogue is
        // told so (the flag is written, fenced, and only then the barrier is signalled)
        if (elect_one()) {
          *o_valid_s = o_init ? 1u : 0u;
          __threadfence_block();
          if (o_init) umma_commit(o_full);
          else mbar_arrive(o_full);
        }
      }
      if (elect_one()) {
        atomicAdd(g_stream_stats + (DW ? 3 : 0), static_cast<unsigned long long>(it2));
        atomicAdd(g_stream_stats + (DW ? 4 : 1), static_cast<unsigned long long>(n_s2));
        atomicAdd(g_stream_stats + (DW ? 5 : 2), static_cast<unsigned long long>(n_live));
      }
    }
  } else {  // ---- epilogue warps 2..9 ----
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;       // which 64 of the 128 S columns
    const int row_l = q * 32 + lane;        // row of the X tile / of E / of O
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    uint32_t it = 0, o_phase = 0, nent = 0;
    uint32_t p2a = 0, p2b = 0;   // phases of s2_full of buffer 0 / 1 (tiles of screening items that were not ruled out)
    bool screen = true;          // this item's mode (mode_s, read after the item's first s_full)
    for (int w = blockIdx.x; w < items; w += gridDim.x) {
      const int m0 = item_x(w) * BM, sp = item_split(w);
      const int t0 = sp * tiles_per_split, t1 = min(y_tiles, t0 + tiles_per_split);
      if (t0 >= t1) continue;
      const int row = m0 + row_l;
      const bool row_ok = row < x_rows;
      const float k1 = __ldg(consts + C_K1);
      const float out_scale = __ldg(consts + (DW ? C_OUT_DW : C_OUT_DH));
      float r_off, r_scale;
      if (DW) {
        r_off = row_ok ? fmaf(__ldg(bias + row), LOG2E, __ldg(consts + C_SA)) : 0.0f;
        r_scale = row_ok ? 1.0f : 0.0f;
      } else {
        r_off = row_ok ? __ldg(m2neg + row) + E_SHIFT : 0.0f;
        r_scale = row_ok ? 1.0f : 0.0f;
      }
      const float flush_scale = DW ? out_scale : (row_ok ? out_scale * __ldg(ascale + row) : 0.0f);
      const float err_row = (!DW && row_ok) ? __ldg(errv + row) : 0.0f;   // DW: per column, staged as chunk maxima
      int tk[SB_MAX_K];   // DW = false: this row's selected slots
#pragma unroll
      for (int k = 0; k < SB_MAX_K; ++k)
        tk[k] = (!DW && row_ok && k < topk) ? __ldg(utopi + static_cast<int64_t>(row) * topk + k) : -1;
      {  // X tile: ring slot (TMA, 128-byte swizzle) -> tensor memory; warp (q, half) copies plane `half` of rows 32q..
        const uint32_t slot = nent % RING;
        mbar_wait(r_full + slot, (nent / RING) & 1);
        const uint8_t* xrow = ring + slot * TILE_BYTES + (row_l >> 3) * 1024 + (row_l & 7) * 128;
#pragma unroll 1
        for (int kb = 0; kb < 2; ++kb) {
          uint32_t v[32];
          if (kb < kblocks) {
            const uint8_t* src = xrow + (kb * NP + half) * TP_BYTES;
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) {
              const uint4 c = *reinterpret_cast<const uint4*>(src + ((ch ^ (row_l & 7)) << 4));
              v[ch * 4 + 0] = c.x; v[ch * 4 + 1] = c.y; v[ch * 4 + 2] = c.z; v[ch * 4 + 3] = c.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0u;
          }
          tmem_st32(tmem_base + lane_off + X_COL + half * 64 + kb * 32, v);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(xt_full);
        nent += 1 + (t1 - t0);
      }
      float rsum = 0.0f;   // DW: db3[row] = sum of E over all streamed nodes
      const int etid = threadIdx.x - 64;   // 0..255 among the epilogue threads
      int nxt[SB_MAX_K / 2];   // DW: selected slots of entries etid, etid + 256, .. of the next Y tile (-1: none)
      auto load_sel = [&](int tile) {
#pragma unroll
        for (int i = 0; i < SB_MAX_K / 2; ++i) {
          const int e_i = i * 256 + etid;          // k-major: entry = k * 128 + node
          const int k = e_i >> 7, node = tile * SBN + (e_i & 127);
          // (the RAW loaded value: any arithmetic on it here would wait for the load and undo the prefetch)
          nxt[i] = -1;
          if (k < topk && node < y_rows) nxt[i] = __ldg(utopi + static_cast<int64_t>(node) * topk + k);
        }
      };
      // Per-column offset / scale of a Y tile (DW: -max log2e + log2|a| and sgn(a) of its nodes; else bias log2e and 1):
      // every epilogue thread fetches ONE value one tile ahead and parks it in shared memory before the tile's barrier,
      // so the inner loop reads its 64 columns with broadcast LDS instead of waiting on 16 global loads per 32 columns
      // (22 % of this kernel's stall samples in the first version).  Columns past the last row get offset -inf: E = 0.
      float pf_col = 0.0f, pf_err = 0.0f;
      auto load_col = [&](int tile) {
        const int c = tile * SBN + (etid & 127);
        const bool second = etid >= 128;
        pf_col = second ? 0.0f : -INFINITY;
        pf_err = 0.0f;
        if (c < y_rows) {
          pf_col = second ? (DW ? __ldg(ascale + c) : 1.0f) : (DW ? __ldg(m2neg + c) : __ldg(bias + c));
          if (DW && !second) pf_err = __ldg(errv + c);
        }
      };
      load_col(t0);
      if (DW) load_sel(t0);
      for (int t = t0; t < t1; ++t, ++it) {
        const uint32_t buf = it & 1, ph = (it >> 1) & 1;
        const uint32_t se = tmem_base + lane_off + SE_COL + buf * SBN + half * 64;
        // DW = true: which of this thread's 64 columns (nodes) selected this thread's slot?  The 256 epilogue threads share
        // the tile's 128 x topk selections (one coalesced pass, requested ONE TILE AHEAD; thread e takes entries e, e + 256,
        // ... in k-major order, so the 32 lanes of a warp hold the k-th choice of 32 neighbouring nodes -- mostly the SAME
        // slot): an entry that falls into the CTA's 128 slots sets the node's bit in that slot's 128-bit mask in shared
        // memory -- one atomic per warp when the lanes agree, one per lane otherwise -- and after a barrier among the
        // epilogue warps every thread reads (and clears) the 64 bits of its own slot and half.  Two mask buffers
        // alternate, so the clears of tile j are ordered before the atomics of tile j + 2 by the barrier of tile j + 1.
        float* cbuf = col_s + (it & 1) * 256;
        {
          const float cval = (!DW && etid < 128) ? pf_col * LOG2E : pf_col;
          cbuf[etid] = cval;
          if (etid < 128) {   // warp-uniform: the four warps that hold the offsets; warp w = columns 32w.. = chunk w
            float m = cval, me = pf_err;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
              if (DW) me = fmaxf(me, __shfl_xor_sync(0xffffffffu, me, o));
            }
            if (lane == 0) {
              cmax_s[(it & 1) * 4 + (etid >> 5)] = m;
              emax_s[(it & 1) * 4 + (etid >> 5)] = me;
            }
          }
        }
        if (t + 1 < t1) load_col(t + 1);
        uint64_t kill = 0;
        if (DW) {
          unsigned* kb = kill_s + (it & 1) * (128 * 4);
          int cur[SB_MAX_K / 2];
#pragma unroll
          for (int i = 0; i < SB_MAX_K / 2; ++i) cur[i] = nxt[i];
          if (t + 1 < t1) load_sel(t + 1);
#pragma unroll
          for (int i = 0; i < SB_MAX_K / 2; ++i) {
            if (i * 256 < 128 * topk) {   // uniform
              const int d = cur[i] - m0;      // slot - first slot of the X tile (cur = -1: no entry)
              const bool hit = cur[i] >= 0 && d >= 0 && d < 128;
              const unsigned hits = __ballot_sync(0xffffffffu, hit);
              if (hits) {                 // warp-uniform
                const int d0 = __shfl_sync(0xffffffffu, d, __ffs(hits) - 1);
                const int word = (etid & 127) >> 5;   // the warp's 32 nodes share one 32-bit word of every mask
                if (__all_sync(0xffffffffu, !hit || d == d0)) {
                  if (lane == 0) atomicOr(kb + d0 * 4 + word, hits);
                } else if (hit) {
                  atomicOr(kb + d * 4 + word, 1u << lane);
                }
              }
            }
          }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
        if (DW) {
          unsigned* kb = kill_s + (it & 1) * (128 * 4);
          const uint2 mine = *reinterpret_cast<const uint2*>(kb + row_l * 4 + half * 2);
          if (mine.x | mine.y) {
            kill = (static_cast<uint64_t>(mine.y) << 32) | mine.x;
            *reinterpret_cast<uint2*>(kb + row_l * 4 + half * 2) = make_uint2(0u, 0u);
          }
        }
        mbar_wait(s_full + buf, ph);
        tc_fence_after();
        // both 32-column chunks are requested before the wait: what follows the accumulator is a latency chain (load ->
        // maximum -> bound -> vote -> arrive) that the MMA issuer waits for, not a throughput problem
        uint32_t va[32], vb[32];
        tmem_ld32_nowait(se, va);
        tmem_ld32_nowait(se + 32, vb);
        tmem_wait_ld();
        if (t == t0) screen = *mode_s != 0u;
        if (screen) {  // ---- verdict on the SCREENING product (X_hi Y_hi^T): can any of this thread's 64 entries be non-zero?
           // arg_j <= (max(v1) + err) k1 + (r_off + max(off)) with err >= |S - S1| (k1 > 0; fp32 fma / add are monotone)
          auto tree_max = [](const uint32_t(&v)[32]) {
            float m8[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
              m8[j] = fmaxf(fmaxf(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1])),
                            fmaxf(__uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])));
            return fmaxf(fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3])), fmaxf(fmaxf(m8[4], m8[5]), fmaxf(m8[6], m8[7])));
          };
          const int ci = (it & 1) * 4 + half * 2;
          const float e0 = DW ? emax_s[ci] : err_row, e1 = DW ? emax_s[ci + 1] : err_row;
          // per chunk: the cheap bound (largest accumulator + largest offset), and where that is not enough the bound
          // element by element (the DW pass's column offsets -- minus the row maximum of 32 different nodes -- vary by
          // hundreds within a chunk: largest accumulator and largest offset rarely belong to the same column)
          auto chunk_dead = [&](const uint32_t(&v)[32], float err, int cb) {
            if (fmaf(tree_max(v) + err, k1, r_off + cmax_s[ci + cb]) < DEAD_CHUNK_LOG2) return true;
            const float* offp = cbuf + half * 64 + cb * 32;
            float am[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 off = *reinterpret_cast<const float4*>(offp + j);
              am[0] = fmaxf(am[0], fmaf(__uint_as_float(v[j + 0]) + err, k1, r_off + off.x));
              am[1] = fmaxf(am[1], fmaf(__uint_as_float(v[j + 1]) + err, k1, r_off + off.y));
              am[2] = fmaxf(am[2], fmaf(__uint_as_float(v[j + 2]) + err, k1, r_off + off.z));
              am[3] = fmaxf(am[3], fmaf(__uint_as_float(v[j + 3]) + err, k1, r_off + off.w));
            }
            return fmaxf(fmaxf(am[0], am[1]), fmaxf(am[2], am[3])) < DEAD_CHUNK_LOG2;
          };
          const bool dead = !row_ok || (chunk_dead(va, e0, 0) && chunk_dead(vb, e1, 1));
          const bool warp_dead = __all_sync(0xffffffffu, dead);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (!warp_dead) live1_s[buf] = it + 1u;   // every warp that objects stores the same stamp
            mbar_arrive(v_full + buf);                // (release: the issuer reads the stamp after its wait)
          }
          asm volatile("bar.sync 2, %0;" ::"n"(32 * EPI_WARPS) : "memory");   // ... and so do the other epilogue warps
          if (live1_s[buf] != it + 1u) continue;   // ruled out: E = 0, nothing is written, nothing is issued
          // ---- the tile stays: the issuer adds the two remaining logit products to the same buffer
          mbar_wait(s2_full + buf, buf ? p2b : p2a);
          if (buf) p2b ^= 1; else p2a ^= 1;
          tc_fence_after();
          tmem_ld32_nowait(se, va);
          tmem_ld32_nowait(se + 32, vb);
          tmem_wait_ld();
        }
        uint32_t hi[32], mid[32];   // this thread's 64 columns of E as fp16 pairs: hi plane, mid plane
        uint32_t nz = 0u;           // OR of every fp16 pair this thread writes
#pragma unroll
        for (int cb = 0; cb < 2; ++cb) {
          const int c0 = t * SBN + half * 64 + cb * 32;
          uint32_t(&v)[32] = cb ? vb : va;
          {  // cheap bound first: arg_j = v_j k1 + (r_off + off_j) <= max(v) k1 + (r_off + max(off)) (k1 > 0, and fp32
             // fma / add are monotone, so the bound holds in floating point too): a tree of 16 FMNMX3 decides most chunks
            float m8[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
              m8[j] = fmaxf(fmaxf(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1])),
                            fmaxf(__uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])));
            const float vm = fmaxf(fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3])),
                                   fmaxf(fmaxf(m8[4], m8[5]), fmaxf(m8[6], m8[7])));
            if (!no_skip && fmaf(vm, k1, r_off + cmax_s[(it & 1) * 4 + half * 2 + cb]) < DEAD_CHUNK_LOG2) {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                hi[cb * 16 + i] = 0u;
                mid[cb * 16 + i] = 0u;
              }
              continue;
            }
          }
          float e[32];
          float amax = -INFINITY;   // largest exp2 argument of the chunk
          {
            const float* offp = cbuf + half * 64 + cb * 32;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 off = *reinterpret_cast<const float4*>(offp + j);
              e[j + 0] = fmaf(__uint_as_float(v[j + 0]), k1, r_off + off.x);
              e[j + 1] = fmaf(__uint_as_float(v[j + 1]), k1, r_off + off.y);
              e[j + 2] = fmaf(__uint_as_float(v[j + 2]), k1, r_off + off.z);
              e[j + 3] = fmaf(__uint_as_float(v[j + 3]), k1, r_off + off.w);
              amax = fmaxf(fmaxf(amax, fmaxf(e[j + 0], e[j + 1])), fmaxf(e[j + 2], e[j + 3]));
            }
          }
          // Dead chunk: every argument more than 130 binades down -- ex2.approx.ftz returns exactly 0 for each of them
          // (it flushes below 2^-126), so E, its two planes and the chunk's share of db3 are exactly zero and the 32
          // MUFU.EX2, the masks and the plane split are skipped.
          if (!no_skip && amax < DEAD_CHUNK_LOG2) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              hi[cb * 16 + i] = 0u;
              mid[cb * 16 + i] = 0u;
            }
            continue;
          }
          {
            const float* offp = cbuf + half * 64 + cb * 32 + 128;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 sc = *reinterpret_cast<const float4*>(offp + j);
              e[j + 0] = fast_exp2(e[j + 0]) * sc.x;
              e[j + 1] = fast_exp2(e[j + 1]) * sc.y;
              e[j + 2] = fast_exp2(e[j + 2]) * sc.z;
              e[j + 3] = fast_exp2(e[j + 3]) * sc.w;
            }
          }
          {  // the selected slots belong to the sparse part
            uint32_t k32 = static_cast<uint32_t>(kill >> (cb * 32));
            if (!DW) {
#pragma unroll
              for (int k = 0; k < SB_MAX_K; ++k) {
                const unsigned d = static_cast<unsigned>(tk[k] - c0);
                if (tk[k] >= 0 && d < 32u) k32 |= 1u << d;
              }
            }
            if (k32) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if ((k32 >> j) & 1u) e[j] = 0.0f;
            }
          }
          if (!row_ok) {
#pragma unroll
            for (int j = 0; j < 32; ++j) e[j] = 0.0f;
          }
          if (DW) {   // two-level sum: a chunk's 32 terms first (pairwise), then into the running total
            float cs[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) cs[j] = (e[j] + e[j + 8]) + (e[j + 16] + e[j + 24]);
            rsum += ((cs[0] + cs[1]) + (cs[2] + cs[3])) + ((cs[4] + cs[5]) + (cs[6] + cs[7]));
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            // (|E| <= 2^13 (1 + 1e-4) by construction of the exponent offsets: fp16 cannot overflow here)
            const float a = e[2 * i], b = e[2 * i + 1];
            const uint32_t h2 = pack_f16x2(a, b);
            const float2 hf = unpack_f16x2(h2);
            hi[cb * 16 + i] = h2;
            mid[cb * 16 + i] = pack_f16x2(a - hf.x, b - hf.y);
            nz |= h2;             // (mid = 0 wherever hi = 0: |a| < 2^-25 rounds to zero in both)
          }
        }
        // E over S, in place: this thread's 64 fp32 columns become 32 columns of hi pairs + 32 columns of mid pairs
        tmem_st32_nowait(se, hi);
        tmem_st32_nowait(se + 32, mid);
        tmem_wait_st();
        tc_fence_before();
        const bool warp_live = __any_sync(0xffffffffu, (nz & 0x7fff7fffu) != 0u);   // (-0 is zero)
        __syncwarp();
        if (lane == 0) {
          if (warp_live) live_s[buf] = it + 1u;   // every live warp stores the same stamp; released by the arrive below
          mbar_arrive(e_full + buf);
        }
      }

      // ---- O tile: accumulated in TMEM over the item's Y tiles -> global (reductions: other splits / the sparse part
      //      / earlier batches add into the same buffer)
      mbar_wait(o_full, o_phase);
      o_phase ^= 1;
      tc_fence_after();
      const bool o_valid = *o_valid_s != 0u;   // (no tile of the item issued its second product: O was never written)
#pragma unroll 1
      for (int cc = 0; cc < 64; cc += 32) {
        const int col0 = half * 64 + cc;
        if (col0 >= n2 || !o_valid) continue;   // warp-uniform
        uint32_t v[32];
        tmem_ld32(tmem_base + lane_off + O_COL + col0, v);
        if (row_ok) {
          float* o = out + static_cast<int64_t>(row) * Kdim + col0;
          if ((Kdim & 3) == 0 && col0 + 32 <= Kdim) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              red_add_v4(o + j, __uint_as_float(v[j]) * flush_scale, __uint_as_float(v[j + 1]) * flush_scale,
                         __uint_as_float(v[j + 2]) * flush_scale, __uint_as_float(v[j + 3]) * flush_scale);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < Kdim) atomicAdd(o + j, __uint_as_float(v[j]) * flush_scale);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty);
      if (DW && row_ok && dbias) atomicAdd(dbias + row, rsum * exp2f(-__ldg(consts + C_SA)));   // (E'' carries 2^sa)
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// per node u: g_k = dtv[u,k] + sum_l cnt[s(l,u)] gcol_k[l,k];  spk[u,k] = p_k (g_k - <g,p>) (the dlogits of the selected
// slots);  a_u = -<g,p> / row_sum;
// m2neg_u = -row_max log2e;  coff_u = m2neg_u + log2|a_u| (-inf when a_u = 0), sgn_u = sign(a_u): the column offset / sign
// of the DW = true pass, whose E carries a_u inside the exp2 argument; the largest log2|a_u| goes to consts[C_MAXLA].
// (The dense-adjoint inputs of gngf_hpd_dlogits do not exist in top-k-only mode.)
// With node_ids the rows are the active nodes (k11_active_nodes.cu): row r <-> lattice node node_ids[r]; dtv and cnt are
// indexed by the lattice node, everything else by the row.
__device__ __forceinline__ int ordered_key(float x) {   // monotone float -> int map (for atomicMax on signed values)
  const int b = __float_as_int(x);
  return b >= 0 ? b : b ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_value(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff); }
constexpr int MAXLA_INIT = static_cast<int>(0x80808080u);   // what cudaMemsetAsync(.., 0x80, ..) leaves: below every key

__global__ void __launch_bounds__(256)
    hpd_stream_bwd_prep_kernel(gngf_lattice lat, const int* __restrict__ node_ids, int64_t U, int K,
                               const float* __restrict__ utopv,
                               const float* __restrict__ dtv, const int* __restrict__ cnt,
                               const float* __restrict__ gcol_k, const float* __restrict__ row_max,
                               const float* __restrict__ row_sum, float* __restrict__ ascale,
                               float* __restrict__ m2neg, float* __restrict__ coff, float* __restrict__ sgn,
                               float* __restrict__ spk, float* __restrict__ consts) {
  const int64_t u = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  float la = -INFINITY;
  if (u < U) {
    const int L = lat.num_levels;
    const int64_t un = node_ids ? node_ids[u] : u;
    const int cx = lat.ox + static_cast<int>(un / lat.wy), cy = lat.oy + static_cast<int>(un % lat.wy);
    float dot = 0.0f;
    float gk[SB_MAX_K];
#pragma unroll
    for (int k = 0; k < SB_MAX_K; ++k) {
      if (k >= K) break;
      float g = dtv[un * K + k];
      if (gcol_k) {
        for (int l = 0; l < L; ++l) {
          const int i = cx - lat.lox[l], j = cy - lat.loy[l];
          if (i >= 0 && i < lat.lwx[l] && j >= 0 && j < lat.lwy[l]) {
            const float c = static_cast<float>(cnt[lat.loff[l] + static_cast<int64_t>(i) * lat.lwy[l] + j]);
            g = fmaf(c, gcol_k[l * K + k], g);
          }
        }
      }
      gk[k] = g;
      dot = fmaf(utopv[u * K + k], g, dot);
    }
    // dlogit at the selected slots, as ONE product (see the header): p_k (g_k - <g,p>)
#pragma unroll
    for (int k = 0; k < SB_MAX_K; ++k) {
      if (k >= K) break;
      float v = utopv[u * K + k] * (gk[k] - dot);
      if (!(fabsf(v) <= 3.0e38f)) v = 0.0f;
      spk[u * K + k] = v;
    }
    float a = -dot / row_sum[u];
    float m = -row_max[u] * LOG2E;
    if (!(fabsf(a) <= 3.0e38f)) a = 0.0f;   // nan_to_num of a degenerate row (models.py:111)
    if (!(fabsf(m) <= 3.0e38f)) m = 0.0f;
    ascale[u] = a;
    m2neg[u] = m;
    if (a != 0.0f) la = log2f(fabsf(a));
    coff[u] = m + la;
    sgn[u] = a > 0.0f ? 1.0f : (a < 0.0f ? -1.0f : 0.0f);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) la = fmaxf(la, __shfl_xor_sync(0xffffffffu, la, o));
  if ((threadIdx.x & 31) == 0 && la > -INFINITY)
    atomicMax(reinterpret_cast<int*>(consts + C_MAXLA), ordered_key(la));
}

// one thread: the device scalars of the dense passes from the operand scales (gngf_split_f16x2) and the largest log2|a|
__global__ void hpd_stream_bwd_consts_kernel(const float* __restrict__ scale_h, const float* __restrict__ scale_w,
                                             float* __restrict__ consts) {
  const float ih = scale_h[0], iw = scale_w[0];   // inverse operand scales (powers of two)
  const int key = *reinterpret_cast<const int*>(consts + C_MAXLA);
  float sa = 0.0f;
  if (key != MAXLA_INIT) sa = 13.0f - ceilf(ordered_value(key));
  sa = fminf(fmaxf(sa, -100.0f), 100.0f);
  consts[C_K1] = LOG2E * ih * iw;
  consts[C_OUT_DH] = exp2f(-E_SHIFT) * iw;
  consts[C_OUT_DW] = exp2f(-sa) * ih;
  consts[C_SA] = sa;
}

// The K selected slots of every node:  dh[u,:] = (dh[u,:] + sum_k spk W3[t_k,:]) .* act'(h[u,:]);
// dW3[t_k,:] += spk h[u,:];  db3[t_k] += spk.  Runs after the dense pass, as the last writer of dh.
//
// A warp walks a RUN of consecutive rows (lane = 4 of the row's <= 128 features).  Consecutive rows are neighbouring
// lattice nodes and the HPD is a smooth function of the node coordinate, so neighbours mostly select the SAME slots:
// the warp keeps the last SP_CACHE distinct slots it met -- their W3 row (read once instead of once per node) and the
// running sums  sum_u spk h[u,:]  and  sum_u spk  in registers -- and only issues the vector reduction into dW3 / db3
// when a slot leaves the cache or the run ends.  With one reduction per (node, slot) (the first form of this kernel,
// a warp per node) 30 M nodes x 4 slots x 128 floats went into the 16 384 rows of BASELINE.json configs[3]'s T = 2^14
// table as same-address L2 reductions: 132 ms per step; cache misses only occur where the selection changes.
constexpr int SP_CACHE = 8;
constexpr int SP_RUN = 64;

__global__ void __launch_bounds__(256)
    hpd_stream_bwd_sparse_kernel(int64_t U, int K, int Kdim, const int* __restrict__ utopi, const float* __restrict__ spk,
                                 const float* __restrict__ h, const float* __restrict__ w, int act_prev,
                                 float* __restrict__ dh, float* __restrict__ dw, float* __restrict__ db) {
  const int64_t run = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) / 32;
  const int lane = threadIdx.x % 32;
  const int64_t u0 = run * SP_RUN;
  if (u0 >= U) return;
  const int64_t u1 = min(U, u0 + SP_RUN);
  const int c = lane * 4;
  const bool on = c < Kdim;   // Kdim % 4 == 0
  int cs[SP_CACHE];           // cached slot ids (-1: empty); identical in every lane
  float4 cw[SP_CACHE], ca[SP_CACHE];
  float cb[SP_CACHE];
#pragma unroll
  for (int e = 0; e < SP_CACHE; ++e) cs[e] = -1;
  int victim = 0;
  auto flush = [&](int t, const float4& a, float bsum) {
    if (on) red_add_v4(dw + static_cast<int64_t>(t) * Kdim + c, a.x, a.y, a.z, a.w);
    if (lane == 0 && db) atomicAdd(db + t, bsum);
  };
  for (int64_t u = u0; u < u1; ++u) {
    float4 hv = make_float4(0.f, 0.f, 0.f, 0.f), acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (on) {
      hv = *reinterpret_cast<const float4*>(h + u * Kdim + c);
      acc = *reinterpret_cast<const float4*>(dh + u * Kdim + c);
    }
    for (int k = 0; k < K; ++k) {
      const int t = utopi[u * K + k];
      const float sv = spk[u * K + k];
      bool hit = false;
#pragma unroll
      for (int e = 0; e < SP_CACHE; ++e) {
        if (cs[e] == t) {   // (warp-uniform: t and cs[] are the same in every lane)
          hit = true;
          acc.x = fmaf(sv, cw[e].x, acc.x);
          acc.y = fmaf(sv, cw[e].y, acc.y);
          acc.z = fmaf(sv, cw[e].z, acc.z);
          acc.w = fmaf(sv, cw[e].w, acc.w);
          ca[e].x = fmaf(sv, hv.x, ca[e].x);
          ca[e].y = fmaf(sv, hv.y, ca[e].y);
          ca[e].z = fmaf(sv, hv.z, ca[e].z);
          ca[e].w = fmaf(sv, hv.w, ca[e].w);
          cb[e] += sv;
        }
      }
      if (!hit) {
        float4 wv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (on) wv = __ldg(reinterpret_cast<const float4*>(w + static_cast<int64_t>(t) * Kdim + c));
        acc.x = fmaf(sv, wv.x, acc.x);
        acc.y = fmaf(sv, wv.y, acc.y);
        acc.z = fmaf(sv, wv.z, acc.z);
        acc.w = fmaf(sv, wv.w, acc.w);
#pragma unroll
        for (int e = 0; e < SP_CACHE; ++e) {
          if (e == victim) {
            if (cs[e] >= 0) flush(cs[e], ca[e], cb[e]);
            cs[e] = t;
            cw[e] = wv;
            ca[e] = make_float4(sv * hv.x, sv * hv.y, sv * hv.z, sv * hv.w);
            cb[e] = sv;
          }
        }
        victim = (victim + 1) % SP_CACHE;
      }
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
#pragma unroll
  for (int e = 0; e < SP_CACHE; ++e)
    if (cs[e] >= 0) flush(cs[e], ca[e], cb[e]);
}

// Column splits per X tile.  Work items are (X tile, split) pairs handed out round-robin to one CTA per SM.  Plenty of X
// tiles (the dh pass of a large lattice): no split.  Few (the dW3 pass: T / 128 tiles): split so that the number of
// items is a MULTIPLE of the SM count -- 128 tiles x 2 splits on 148 SMs left 40 SMs with half the work of the others
// (14 % of that pass) -- with at least 64 Y tiles per item to amortise the X-tile load and the accumulator flush.
// largest row norm of W3 -> consts[C_MAXW] (bits of a non-negative float order like integers)
__global__ void __launch_bounds__(256)
    hpd_stream_bwd_wnorm_kernel(const float* __restrict__ w, int64_t T, int Kdim, float* __restrict__ consts) {
  const int64_t row = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) / 32;
  const int lane = threadIdx.x % 32;
  float ss = 0.0f;
  if (row < T)
    for (int c = lane; c < Kdim; c += 32) {
      const float v = __ldg(w + row * Kdim + c);
      ss = fmaf(v, v, ss);
    }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if (lane == 0 && row < T) atomicMax(reinterpret_cast<int*>(consts) + C_MAXW, __float_as_int(sqrtf(ss) * 1.0001f));
}

// errv[u] >= |S[u, t] - S_hihi[u, t]| for every slot t, in accumulator units (both operands carry their plane scale):
// with fp16-rounded planes x = x_hi + x_mid + r, |x - x_hi| <= 2^-11 |x|, so the products the screening pass leaves out
// are bounded by 2^-10 (1 + 2^-11) sum |x||y| <= 2^-10 (1 + 2^-11) |x||y| (Cauchy-Schwarz); SCREEN_ERR doubles that, which
// also covers the fp32 accumulation order and the sub-normal tail of the planes.
__global__ void __launch_bounds__(256)
    hpd_stream_bwd_errv_kernel(const float* __restrict__ h, int64_t U, int Kdim, const float* __restrict__ h_scale,
                               const float* __restrict__ w_scale, const float* __restrict__ consts,
                               float* __restrict__ errv) {
  const int64_t row = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) / 32;
  const int lane = threadIdx.x % 32;
  if (row >= U) return;
  float ss = 0.0f;
  for (int c = lane; c < Kdim; c += 32) {
    const float v = __ldg(h + row * Kdim + c);
    ss = fmaf(v, v, ss);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if (lane == 0)
    // (h_scale / w_scale hold the INVERSE plane scales, see hpd_stream_bwd_consts_kernel)
    errv[row] = SCREEN_ERR * (sqrtf(ss) * 1.0001f) * __ldg(consts + C_MAXW) / (__ldg(h_scale) * __ldg(w_scale)) + 1.0f;
}

static int split_count(int64_t x_tiles, int64_t y_tiles) {
  const int64_t sms = gngf::sm_count();
  if (x_tiles >= 8 * sms) return 1;
  int64_t g = x_tiles, b = sms;
  while (b) { const int64_t t = g % b; g = b; b = t; }      // gcd
  const int64_t unit = sms / g;                             // smallest split count with x_tiles * ns % sms == 0
  const int64_t most = std::max<int64_t>(1, y_tiles / 64);  // keep items at >= 64 Y tiles
  int64_t ns = std::max<int64_t>(1, (2 * sms) / x_tiles);   // the old rule: at least two items per SM
  if (unit <= most) ns = unit * std::max<int64_t>(1, std::min<int64_t>(most / unit, (ns + unit - 1) / unit));
  return static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(ns, y_tiles)));
}

}  // namespace sb
}  // namespace tc
}  // namespace gngf

extern "C" {

int64_t gngf_hpd_stream_bwd_workspace_floats(int64_t U, int32_t topk) {
  return 5 * ((U + 3) & ~int64_t(3)) + U * static_cast<int64_t>(topk) + 16;
}

int gngf_hpd_stream_bwd_nodes(gngf_lattice lat, const int32_t* node_ids, const uint16_t* h_planes, const float* h_scale,
                              const uint16_t* w_planes, const float* w_scale, const float* h, const float* w,
                              const float* bias, int64_t U,
                              int64_t T, int64_t Kdim, int32_t topk, const float* utopv, const int32_t* utopi,
                              const float* dtv, const int32_t* cnt, const float* gcol_k, const float* row_max,
                              const float* row_sum, int32_t act_prev, float* dh, float* dw, float* db, float* workspace,
                              void* stream) {
  using namespace gngf::tc;
  using namespace gngf::tc::sb;
  const int64_t box = static_cast<int64_t>(lat.wx) * lat.wy;
  if (U <= 0 || T <= 0 || Kdim <= 0 || (Kdim % 8) != 0 || Kdim > 2 * BK || topk <= 0 || topk > SB_MAX_K ||
      U >= (1ll << 31) || T >= (1ll << 31) || (node_ids ? U > box : (gcol_k != nullptr && U != box)))
    return GNGF_ERR_UNSUPPORTED;   // (no node list and no column-sum adjoint: plain rows, dtv indexed by the row)
  if (gcol_k && !cnt) return GNGF_ERR_INVALID_ARGUMENT;
  if (!h_planes || !w_planes || !h_scale || !w_scale || !h || !w || !bias || !utopv || !utopi || !dtv || !row_max ||
      !row_sum || !dh || !dw || !workspace)
    return GNGF_ERR_INVALID_ARGUMENT;
  if ((reinterpret_cast<uintptr_t>(bias) | reinterpret_cast<uintptr_t>(workspace) | reinterpret_cast<uintptr_t>(dh) |
       reinterpret_cast<uintptr_t>(dw) | reinterpret_cast<uintptr_t>(h) | reinterpret_cast<uintptr_t>(w)) & 15)
    return GNGF_ERR_INVALID_ARGUMENT;
  cudaStream_t st = gngf::as_stream(stream);
  // workspace: ascale | m2neg | coff | sgn | errv (U each, rounded up to 4) | spk (U, topk) | consts (8, 16-byte aligned)
  const int64_t U4 = (U + 3) & ~int64_t(3);
  float* ascale = workspace;
  float* m2neg = workspace + U4;
  float* coff = workspace + 2 * U4;
  float* sgn = workspace + 3 * U4;
  float* errv = workspace + 4 * U4;
  float* spk = workspace + 5 * U4;
  float* consts = spk + ((U * topk + 3) & ~int64_t(3));
  if (cudaMemsetAsync(consts, 0x80, 8 * sizeof(float), st) != cudaSuccess) return gngf::check_launch();
  hpd_stream_bwd_prep_kernel<<<static_cast<unsigned>(gngf::ceil_div(U, 256)), 256, 0, st>>>(
      lat, node_ids, U, topk, utopv, dtv, cnt, gcol_k, row_max, row_sum, ascale, m2neg, coff, sgn, spk, consts);
  gngf::note_launch();
  int rc = gngf::check_launch();
  if (rc) return rc;
  hpd_stream_bwd_consts_kernel<<<1, 1, 0, st>>>(h_scale, w_scale, consts);
  gngf::note_launch();
  if ((rc = gngf::check_launch())) return rc;
  // the screening pass's error bound per node: |W3 row|_max, then SCREEN_ERR |h_u| |W3 row|_max in accumulator units
  hpd_stream_bwd_wnorm_kernel<<<static_cast<unsigned>(gngf::ceil_div(T * 32, 256)), 256, 0, st>>>(w, T, static_cast<int>(Kdim),
                                                                                                   consts);
  gngf::note_launch();
  hpd_stream_bwd_errv_kernel<<<static_cast<unsigned>(gngf::ceil_div(U * 32, 256)), 256, 0, st>>>(
      h, U, static_cast<int>(Kdim), h_scale, w_scale, consts, errv);
  gngf::note_launch();
  if ((rc = gngf::check_launch())) return rc;

  CUtensorMap map_h, map_w;   // 128-row boxes serve both roles (resident X tile, streamed Y tiles); two fp16 planes
  if ((rc = make_plane_map(&map_h, h_planes, U, Kdim, BM, NP))) return rc;
  if ((rc = make_plane_map(&map_w, w_planes, T, Kdim, BM, NP))) return rc;
  if (cudaFuncSetAttribute(hpd_stream_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           static_cast<int>(SMEM_BYTES)) != cudaSuccess ||
      cudaFuncSetAttribute(hpd_stream_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           static_cast<int>(SMEM_BYTES)) != cudaSuccess)
    return gngf::check_launch();
  const int sms = gngf::sm_count();
  const int no_skip = gngf::debug_no_skip();
  {  // dh (U, Kdim) += E W3 : X = h tiles, Y = W3 tiles
    const int64_t xt = gngf::ceil_div(U, BM), yt = gngf::ceil_div(T, SBN);
    const int ns = split_count(xt, yt);
    const int grid = static_cast<int>(std::min<int64_t>(xt * ns, sms));
    hpd_stream_bwd_kernel<false><<<grid, THREADS, SMEM_BYTES, st>>>(map_h, map_w, static_cast<int>(U),
                                                                   static_cast<int>(T), static_cast<int>(Kdim), ns, bias,
                                                                   m2neg, ascale, errv, consts, utopi, topk, no_skip, dh, nullptr);
    gngf::note_launch();
    if ((rc = gngf::check_launch())) return rc;
  }
  {  // dW3 (T, Kdim) += E^T h, db3 += E^T 1 : X = W3 tiles, Y = h tiles
    const int64_t xt = gngf::ceil_div(T, BM), yt = gngf::ceil_div(U, SBN);
    const int ns = split_count(xt, yt);
    const int grid = static_cast<int>(std::min<int64_t>(xt * ns, sms));
    hpd_stream_bwd_kernel<true><<<grid, THREADS, SMEM_BYTES, st>>>(map_w, map_h, static_cast<int>(T),
                                                                  static_cast<int>(U), static_cast<int>(Kdim), ns, bias,
                                                                  coff, sgn, errv, consts, utopi, topk, no_skip, dw, db);
    gngf::note_launch();
    if ((rc = gngf::check_launch())) return rc;
  }
  hpd_stream_bwd_sparse_kernel<<<static_cast<unsigned>(gngf::ceil_div(gngf::ceil_div(U, SP_RUN) * 32, 256)), 256, 0, st>>>(
      U, topk, static_cast<int>(Kdim), utopi, spk, h, w, act_prev, dh, dw, db);
  gngf::note_launch();
  return gngf::check_launch();
}

int gngf_hpd_stream_bwd(gngf_lattice lat, const uint16_t* h_planes, const float* h_scale, const uint16_t* w_planes,
                        const float* w_scale, const float* h, const float* w, const float* bias, int64_t U, int64_t T,
                        int64_t Kdim, int32_t topk,
                        const float* utopv, const int32_t* utopi, const float* dtv, const int32_t* cnt,
                        const float* gcol_k, const float* row_max, const float* row_sum, int32_t act_prev, float* dh,
                        float* dw, float* db, float* workspace, void* stream) {
  return gngf_hpd_stream_bwd_nodes(lat, nullptr, h_planes, h_scale, w_planes, w_scale, h, w, bias, U, T, Kdim, topk, utopv,
                                   utopi, dtv, cnt,
                                   gcol_k, row_max, row_sum, act_prev, dh, dw, db, workspace, stream);
}

int gngf_hpd_stream_bwd_stats(uint64_t* out6, int32_t reset) {
  if (!out6) return GNGF_ERR_INVALID_ARGUMENT;
  unsigned long long h[6] = {0, 0, 0, 0, 0, 0};
  if (cudaDeviceSynchronize() != cudaSuccess ||
      cudaMemcpyFromSymbol(h, gngf::tc::sb::g_stream_stats, sizeof(h)) != cudaSuccess)
    return gngf::check_launch();
  for (int i = 0; i < 6; ++i) out6[i] = h[i];
  if (reset) {
    const unsigned long long z[6] = {0, 0, 0, 0, 0, 0};
    if (cudaMemcpyToSymbol(gngf::tc::sb::g_stream_stats, z, sizeof(z)) != cudaSuccess) return gngf::check_launch();
  }
  return GNGF_OK;
}

}  // extern "C"
