/*
 * gngf.h -- C ABI of libgngf_sm100.so: the B200 (sm_100a) kernels behind the GNGF training hot path of
 * FedeMont/collision_handling_in_instantNGP (forward + backward of GeneralNeuralGaugeFields).
 *
 * The reference is pure PyTorch and has no FFI of its own; each entry point below names the reference
 * function (file:line, relative to the reference root) whose arithmetic it replaces.  The host side
 * (collision_handling_in_instantngp_b200/models.py) keeps the reference's class signatures and calls these
 * through ctypes with raw device pointers.  INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to contiguous row-major memory owned by the caller (PyTorch's
 *     caching allocator); nothing is allocated or freed by the library;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), thread-safe for distinct
 *     streams, and returns 0 or a negative gngf_status; gngf_strerror() names it;
 *   - fp32 everywhere; indices int32 on the lattice, int64 in API outputs (torch.topk's dtype);
 *   - corner order v in {0,1,2,3} <-> (dx,dy) = (0,0),(1,0),(0,1),(1,1)            (models.py:322-331);
 *   - "lattice": the HPD only ever sees integer corner coordinates and is shared by all levels
 *     (models.py:353-359,416-418), so it is evaluated once per lattice *node* u = (cx-ox)*wy + (cy-oy) of
 *     the bounding box of all levels, and per-level quantities live on "level nodes"
 *     s = loff[l] + (cx-lox[l])*lwy[l] + (cy-loy[l]).
 */
#ifndef GNGF_H
#define GNGF_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GNGF_MAX_LEVELS 32
#define GNGF_MAX_FEATURES 8
#define GNGF_MAX_TOPK 128

typedef enum {
  GNGF_OK = 0,
  GNGF_ERR_INVALID_ARGUMENT = -1,
  GNGF_ERR_UNSUPPORTED = -2,
  GNGF_ERR_CUDA = -3,
  GNGF_ERR_NO_DEVICE = -4
} gngf_status;

/* top-k mix modes == params.should_softmax_topk_features (models.py:212-217) */
#define GNGF_MIX_SOFTMAX 1      /* True : sum_k g * softmax_K(topk probs)           */
#define GNGF_MIX_WEIGHTED_AVG 0 /* False: sum_k g * p / sum_k p                     */
#define GNGF_MIX_RAW 2          /* None : sum_k g * p                               */

/* activations of gngf_linear_fwd */
#define GNGF_ACT_NONE 0
#define GNGF_ACT_RELU 1
#define GNGF_ACT_LEAKY_RELU 2 /* slope 0.01 (nn.LeakyReLU default, models.py:388) */
#define GNGF_ACT_SIGMOID 3

/* Geometry of one forward call; filled by the host from the level table (models.py:305-317) and the
 * coordinate bounds of the batch.  Passed by value. */
typedef struct {
  int32_t num_levels;
  int32_t n[GNGF_MAX_LEVELS];   /* n_l */
  int32_t ox, oy, wx, wy;       /* global node box: origin and extent; U = wx*wy                      */
  int32_t lox[GNGF_MAX_LEVELS]; /* per-level node boxes                                              */
  int32_t loy[GNGF_MAX_LEVELS];
  int32_t lwx[GNGF_MAX_LEVELS];
  int32_t lwy[GNGF_MAX_LEVELS];
  int64_t loff[GNGF_MAX_LEVELS + 1]; /* prefix offsets of the level boxes; S = loff[num_levels]       */
} gngf_lattice;

/* L table pointers, each (T, F) fp32 (encoding._hash_tables.{l}.weight, models.py:159-164) */
typedef struct {
  float* ptr[GNGF_MAX_LEVELS];
} gngf_tables;

const char* gngf_strerror(int status);
/* version of this ABI; bumped on any signature change */
int gngf_abi_version(void);
/* number of SMs / compute capability of the current device (negative status when no device) */
int gngf_device_info(int* sm_count, int* cc_major, int* cc_minor);
/* cumulative number of kernel launches issued by this library in this process (bench.py: gpu_launches) */
int64_t gngf_launch_count(void);

/* ---- K1: grid corners (models.py:486-502, _scale_to_grid) ------------------------------------------
 * scaled (P,2,L,1) = x * n_l ; grid (P,2,L,4) = floor(scaled) + corner offsets; fp32, integer-valued. */
int gngf_corners_fwd(const float* x, int64_t P, gngf_lattice lat, float* scaled, float* grid, void* stream);

/* spatial hash baseline (models.py:504-528, _fast_hash): idx (P,L,4) int64 in [0,T) */
int gngf_fast_hash_fwd(const float* x, int64_t P, gngf_lattice lat, int64_t table_size, int64_t* idx, void* stream);

/* ---- K2: HPD MLP on lattice nodes (models.py:80-88,105-106) ----------------------------------------
 * first layer (in_features = 2) evaluated directly from the node coordinates:
 *   h (U,N) = relu(c(u) * W0^T + b0), c(u) = (ox + u / wy, oy + u % wy) as fp32                        */
int gngf_hpd_first_layer_fwd(gngf_lattice lat, const float* w0, const float* b0, int32_t n_out, int32_t act,
                             float* h, void* stream);
/* generic linear layer: y (M,N) = act(x (M,K) * w(N,K)^T + b(N)); fp32 CUDA-core tiles                  */
int gngf_linear_fwd(const float* x, const float* w, const float* b, int64_t M, int32_t N, int32_t K, int32_t act,
                    float* y, void* stream);
/* backward of a linear layer whose INPUT was x (M,K) (an activation of the previous layer):
 *   dw (N,K) += dz^T x ; db (N) += colsum(dz) ; dx (M,K) = (dz * w) .* act'(x)   (dx may be NULL)
 * act_prev describes how x was produced (GNGF_ACT_NONE: no mask).  dw/db must be zeroed by the caller.   */
int gngf_linear_bwd(const float* dz, const float* x, const float* w, int64_t M, int32_t N, int32_t K,
                    int32_t act_prev, float* dx, float* dw, float* db, void* stream);
/* dz = dy .* y .* (1 - y)  (sigmoid backward of the decoder's last layer, models.py:389)               */
int gngf_sigmoid_bwd(const float* dy, const float* y, int64_t n, float* dz, void* stream);
/* backward of the first HPD layer: dw0 (N,2) += dz^T c(u), db0 (N) += colsum(dz)                        */
int gngf_hpd_first_layer_bwd(gngf_lattice lat, const float* dz, int32_t n_out, float* dw0, float* db0, void* stream);
/* the same two on a list of lattice nodes (gngf_compact_nodes): row r of h / dz <-> node node_ids[r];
 * node_ids == NULL: all U nodes of the box (n_nodes ignored)                                            */
int gngf_hpd_first_layer_fwd_nodes(gngf_lattice lat, const int32_t* node_ids, int64_t n_nodes, const float* w0,
                                   const float* b0, int32_t n_out, int32_t act, float* h, void* stream);
int gngf_hpd_first_layer_bwd_nodes(gngf_lattice lat, const int32_t* node_ids, int64_t n_nodes, const float* dz,
                                   int32_t n_out, float* dw0, float* db0, void* stream);

/* ---- active nodes: the lattice nodes a batch touches ------------------------------------------------
 * models.py:416-418 evaluates the HPD on one row per (point, level, corner); only the distinct corner
 * coordinates matter, and on large lattices (BASELINE.json configs[3]: 8192^2) a batch touches a fraction
 * of the bounding box.  gngf_lattice_mark_nodes sets bit u of `bitmap` (gngf_active_nodes_bitmap_words(U)
 * ZERO-INITIALISED 32-bit words, 16-byte aligned) for every corner node u of every (point, level);
 * gngf_compact_nodes writes the set bits as an ascending list node_ids (at most `capacity` entries are
 * written) and their number to count[0]; chunk_offsets: gngf_active_nodes_chunks(U) int32 of scratch.
 * gngf_scatter_node_rows: dst[node_ids[r], :] = src[r, :] for rows of `row_words` 32-bit elements.       */
int64_t gngf_active_nodes_bitmap_words(int64_t U);
int64_t gngf_active_nodes_chunks(int64_t U);
int gngf_lattice_mark_nodes(const float* x, int64_t P, gngf_lattice lat, uint32_t* bitmap, void* stream);
int gngf_compact_nodes(const uint32_t* bitmap, int64_t U, int32_t* chunk_offsets, int32_t* node_ids, int64_t capacity,
                       int32_t* count, void* stream);
/* Node-parallel HPD across data-parallel ranks (last sentence: "HPD work could alternatively be sharded
 * over unique rows"): gngf_bitmap_or: out[w] = OR over the n_maps all-gathered per-rank bitmaps maps[r * words + w]
 * (words % 4 == 0, 16-byte aligned).  gngf_gather_node_adjoints: out[r,k] = dtv[node_ids[r],k] + sum_l cnt[s(l,node)]
 * gcol_k[l,k] -- this rank's share of the adjoint of the selected probabilities per row of the agreed node list
 * (node_ids NULL: row r is node r), ready for a sum-reduce-scatter to the rows' owners; gcol_k / cnt may be NULL.  The
 * owner then calls gngf_hpd_stream_bwd with node list NULL, cnt = gcol_k = NULL and the reduced rows as dtv.       */
int gngf_bitmap_or(const uint32_t* maps, int32_t n_maps, int64_t words, uint32_t* out, void* stream);
int gngf_gather_node_adjoints(gngf_lattice lat, const int32_t* node_ids, int64_t n_nodes, int32_t K, const float* dtv,
                              const int32_t* cnt, const float* gcol_k, float* out, void* stream);
int gngf_scatter_node_rows(const int32_t* node_ids, int64_t n_nodes, const void* src, int64_t row_words, void* dst,
                           void* stream);

/* ---- K2 (tensor cores): split-precision GEMM on tcgen05 ------------------------------------------------------
 * gngf_split_bf16x3: x = hi + mid + lo in bf16; planes (3, n) row-major after src's own layout.
 * gngf_tc_gemm_bf16x3: C (M,N) fp32 = act(sum of the six partial products of order <= 2 of
 *   (A_hi+A_mid+A_lo)(M,K) (B_hi+B_mid+B_lo)(N,K)^T + bias(N)), fp32 accumulation in TMEM; TMA-staged operands;
 *   a_planes (3,M,K), b_planes (3,N,K) bf16, 16-byte aligned, K % 8 == 0.  This is the HPD output layer
 *   (models.py:80-88: Linear(128, T)) when T is large.                                                       */
int gngf_split_bf16x3(const float* src, int64_t n, uint16_t* planes, void* stream);
/* Two fp16 planes with a power-of-two scale (the operands of the STREAMING kernels): x 2^s = hi + mid, s chosen on the
 * device so that max|x| 2^s lies in [2^13, 2^14) -- 22 mantissa bits per element against 16 for two bf16 planes, at the
 * same tensor-core cost (measured at BASELINE.json configs[2]: dW3 error 1.8e-3 -> fp32 level).  planes (2, n) fp16;
 * scale: 2 device floats, scale[0] receives 2^-s (what a consumer multiplies its accumulator with), scale[1] is
 * scratch.  src 16-byte aligned.  No host synchronisation (CUDA-graph capturable).                               */
int gngf_split_f16x2(const float* src, int64_t n, uint16_t* planes, float* scale, void* stream);
/* transposing variant: planes (3, cols, ld) of src (rows, cols)^T, ld >= rows, columns >= rows zero-filled      */
int gngf_split_bf16x3_t(const float* src, int64_t rows, int64_t cols, int64_t ld, uint16_t* planes, void* stream);
/* accumulate != 0: C += product + bias (no activation) -- used for weight gradients summed over row chunks.
 * k_splits > 1: the K range is split over CTAs and summed with atomics into a ZERO-INITIALISED C (for products
 * whose M x N alone cannot fill the chip, e.g. dX = dlogits W3 with K = T).                                    */
/* operand formats of the planes given to gngf_tc_gemm_bf16x3 (0 = fp16, 1 = bf16; default 1, 1); process-wide */
int gngf_tc_gemm_set_formats(int32_t a_fmt, int32_t b_fmt);
int gngf_tc_gemm_bf16x3(const uint16_t* a_planes, const uint16_t* b_planes, const float* bias, int64_t M, int64_t N,
                        int64_t K, int32_t act, int32_t accumulate, int32_t k_splits, float* C, void* stream);

/* K2+K3 fused, streaming: top-k probabilities/indices and softmax statistics of
 *   softmax(h (U,Kdim) W (T,Kdim)^T + bias) per row, without materialising the (U,T) logits (Kdim <= 128,
 *   topk <= 8).  Ties -> lower index; selection on the logits.  utopv (U,topk) = exp(z - max)/sum of the
 *   winners; row_max/row_sum (U) are what a recomputing backward needs.  workspace:
 *   gngf_hpd_stream_workspace_floats(U, T, topk) floats.                                                     */
int64_t gngf_hpd_stream_workspace_floats(int64_t U, int64_t T, int32_t topk);
int gngf_hpd_stream_fwd(const uint16_t* a_planes, const uint16_t* b_planes, const float* bias, int64_t U, int64_t T,
                        int64_t Kdim, int32_t topk, float* utopv, int32_t* utopi, float* row_max, float* row_sum,
                        float* workspace, void* stream);

/* The same result from HALF the tensor-core work (topk <= 4): a two-plane / three-product pass (1e-5) keeps the 8
 *   best candidates per row, then the candidates' logits are re-evaluated in fp32 from h (U,Kdim) and w (T,Kdim),
 *   re-ranked, and row_max / row_sum corrected for the exact maximum and the exact candidate terms.  Selections are
 *   those of an fp32 evaluation as long as the true top-k lie within the approximate top-8 (a violation needs five
 *   logits within 2e-5 of the k-th largest).  workspace: gngf_hpd_stream_refined_workspace_floats(U, T) floats.   */
int64_t gngf_hpd_stream_refined_workspace_floats(int64_t U, int64_t T);
/*   a_planes (2,U,Kdim) / b_planes (2,T,Kdim) with their scales: gngf_split_f16x2 of h / w.                          */
int gngf_hpd_stream_fwd_refined(const uint16_t* a_planes, const float* a_scale, const uint16_t* b_planes,
                                const float* b_scale, const float* h, const float* w, const float* bias, int64_t U,
                                int64_t T, int64_t Kdim, int32_t topk, float* utopv, int32_t* utopi, float* row_max,
                                float* row_sum, float* workspace, void* stream);

/* K5c fused, streaming (top-k-only mode): backward of gngf_hpd_stream_fwd's layer + softmax + top-k
 *   (models.py:80-88, 105-123; DifferentiableTopk.backward, models.py:21-42; column-sum adjoint of utils.py:138)
 *   without materialising logits, probabilities or dlogits.  Per node u
 *     g_k = dtv[u,k] + sum_l cnt[s(l,u)] gcol_k[l,k],   dlogit[u,:] = -<g,p_top> p[u,:] + scatter_k(p_k g_k),
 *   and   dh (U,Kdim) = (dlogit W) .* act_prev'(h),   dw (T,Kdim) += dlogit^T h,   db (T) += colsum(dlogit).
 *   The dense products run on tcgen05 (two fp16 planes of the power-of-two-scaled operands, hi.hi + hi.mid + mid.hi),
 *   recomputing the logits tile by tile from the planes and the forward's row_max / row_sum.
 *   h_planes (2,U,Kdim) / w_planes (2,T,Kdim) + h_scale / w_scale: gngf_split_f16x2; h (U,Kdim), w (T,Kdim): the fp32 originals
 *   (used by the K-sparse part); utopv / utopi (U,topk): the forward's outputs; dh must be ZERO-INITIALISED, dw / db
 *   are accumulated into; workspace: gngf_hpd_stream_bwd_workspace_floats(U, topk) floats, 16-byte aligned.      */
int64_t gngf_hpd_stream_bwd_workspace_floats(int64_t U, int32_t topk);
/* Measurement aid.  The dense passes screen every 128 x 128 tile with ONE of the three split products of the logits and
 * add the other two -- and the second product (E W3 / E^T h) after them -- only where the tile cannot be ruled out (with
 * a one-hot softmax: almost nowhere), so the tensor work they execute depends on the data.
 * out6 = {dh-pass tiles, of which needed all three logit products, of which issued the second product,
 *         dW3-pass tiles, ..., ...}, accumulated over the calls on the current device since the last reset;
 * synchronises the device.                                                                                            */
int gngf_hpd_stream_bwd_stats(uint64_t* out6, int32_t reset);
int gngf_hpd_stream_bwd(gngf_lattice lat, const uint16_t* h_planes, const float* h_scale, const uint16_t* w_planes,
                        const float* w_scale, const float* h, const float* w, const float* bias, int64_t U, int64_t T,
                        int64_t Kdim, int32_t topk, const float* utopv, const int32_t* utopi, const float* dtv, const int32_t* cnt,
                        const float* gcol_k, const float* row_max, const float* row_sum, int32_t act_prev, float* dh,
                        float* dw, float* db, float* workspace, void* stream);
/* the same on the active nodes: rows of h_planes / h / utopv / utopi / row_max / row_sum / dh are the nodes
 * node_ids[0..U), while dtv (box, topk) and cnt stay indexed by the lattice node; node_ids == NULL: U = box     */
int gngf_hpd_stream_bwd_nodes(gngf_lattice lat, const int32_t* node_ids, const uint16_t* h_planes, const float* h_scale,
                              const uint16_t* w_planes, const float* w_scale, const float* h, const float* w,
                              const float* bias, int64_t U, int64_t T, int64_t Kdim, int32_t topk, const float* utopv, const int32_t* utopi,
                              const float* dtv, const int32_t* cnt, const float* gcol_k, const float* row_max,
                              const float* row_sum, int32_t act_prev, float* dh, float* dw, float* db, float* workspace,
                              void* stream);

/* ---- K2/K3/K5 fused for small lattices (a few hundred to a few thousand nodes, T <= 1024, hidden widths <= 256) ----
 * One CTA per 8 nodes walks the whole HPD (models.py:80-123): layer 0 from the node coordinates, hidden layers,
 * output layer, softmax + nan_to_num, top-k; the backward computes dlogits (as gngf_hpd_dlogits), the bias
 * gradients, the first layer's gradients and the pre-activation adjoints gact[i] (U, width[i+1]) of every layer;
 * the weight gradients of layers i >= 1 are dW_i += gact[i]^T act[i-1] (gngf_linear_bwd with dx = db = NULL).
 * widths: n_layers + 1 entries (2, hidden..., T); w / b / act / gact / dbias: host arrays of device pointers.     */
int gngf_hpd_small_supported(int32_t n_layers, const int32_t* widths, int32_t topk);
int gngf_hpd_small_fwd(gngf_lattice lat, int32_t n_layers, const int32_t* widths, const float* const* w,
                       const float* const* b, float* const* act, int32_t topk, float* uprobs, float* utopv,
                       int32_t* utopi, void* stream);
int gngf_hpd_small_bwd(gngf_lattice lat, int32_t n_layers, const int32_t* widths, const float* const* w,
                       float* const* act, float* const* gact, float* const* dbias, float* dw0, int32_t topk,
                       const float* uprobs, const int32_t* utopi, const float* dtv, const int32_t* cnt,
                       const float* gcol, const float* gcol_k, const float* gdense, void* stream);
/* The same two with the encoding's per-level-node passes folded in (a level node belongs to exactly one lattice
 * node, whose selection the owning warp already holds): the forward also writes nfeat (S,F) as
 * gngf_node_features_fwd would, the backward consumes dnf (S,F) as gngf_node_features_bwd would -- table
 * gradients += and the adjoint of the K selected probabilities kept in registers (dtv may then be NULL, or hold
 * an additional adjoint of utopv).  nfeat / dnf == NULL: identical to the plain entry points.               */
int gngf_hpd_small_fwd_enc(gngf_lattice lat, int32_t n_layers, const int32_t* widths, const float* const* w,
                           const float* const* b, float* const* act, int32_t topk, float* uprobs, float* utopv,
                           int32_t* utopi, gngf_tables tables, int32_t F, int32_t mix_mode, float* nfeat, void* stream);
int gngf_hpd_small_bwd_enc(gngf_lattice lat, int32_t n_layers, const int32_t* widths, const float* const* w,
                           float* const* act, float* const* gact, float* const* dbias, float* dw0, int32_t topk,
                           const float* uprobs, const int32_t* utopi, const float* dtv, const int32_t* cnt,
                           const float* gcol, const float* gcol_k, const float* gdense, gngf_tables tables,
                           gngf_tables table_grads, int32_t F, int32_t mix_mode, const float* dnf, void* stream);

/* ---- K6: fused decoder MLP (models.py:382-392, 468-470) for the reference's shape IN -> 64 -> 64 -> OUT -----
 * rgb (P,OUT) = sigmoid(W2 act(W1 act(W0 enc + b0) + b1) + b2), act = ReLU or LeakyReLU(0.01); activations stay
 * in shared memory.  The backward recomputes them, writes denc (P,IN) and ADDS the parameter gradients into
 * dw0 (64,IN), db0, dw1 (64,64), db1, dw2 (OUT,64), db2; workspace: gngf_mlp3_bwd_workspace_floats() floats.
 * gngf_mlp3_supported() says whether a decoder shape is covered (otherwise use gngf_linear_fwd/bwd).        */
int gngf_mlp3_supported(int32_t in_dim, int32_t h1, int32_t h2, int32_t out_dim);
int gngf_mlp3_fwd(const float* enc, int64_t P, int32_t in_dim, int32_t out_dim, int32_t leaky, const float* w0,
                  const float* b0, const float* w1, const float* b1, const float* w2, const float* b2, float* rgb,
                  void* stream);
int64_t gngf_mlp3_bwd_workspace_floats(int32_t in_dim, int32_t out_dim);
int gngf_mlp3_bwd(const float* enc, const float* drgb, int64_t P, int32_t in_dim, int32_t out_dim, int32_t leaky,
                  const float* w0, const float* b0, const float* w1, const float* b1, const float* w2, const float* b2,
                  float* denc, float* dw0, float* db0, float* dw1, float* db1, float* dw2, float* db2, float* workspace,
                  void* stream);

/* K6 on the tensor cores (k6_mlp_tc.cu): the same decoder (models.py:382-392, 468-470) as chains of tcgen05 products,
 * 128-point tiles, operands as bf16 planes of the fp32 values kept in shared memory between layers.
 *   forward : three planes / six partial products per k-step (~1.5e-6 relative);
 *   backward: two planes / three partial products (~1e-5 relative; gradient bar 1e-4).  It re-reads the forward's
 *   output `rgb` (P,out_dim) instead of recomputing the last layer and gates with the forward's own ReLU pattern
 *   `masks` (P,4) uint32 -- words [layer*2 + half], bit j = pre-activation of hidden unit half*32+j was positive;
 *   the forward writes them when `masks` is non-NULL -- recomputes the hidden activation VALUES, reads every
 *   shared-memory tile K-major for the dX products and MN-major for the dW products, accumulates dW in TMEM over all
 *   tiles of a persistent CTA and adds (red.global) into dw0/db0/dw1/db1/dw2/db2 -- which must be valid accumulators
 *   (zero-initialised or holding earlier contributions); dw1 16-byte aligned.                                     */
int gngf_mlp3_tc_supported(int32_t in_dim, int32_t h1, int32_t h2, int32_t out_dim);
int gngf_mlp3_tc_fwd(const float* enc, int64_t P, int32_t in_dim, int32_t out_dim, int32_t leaky, const float* w0,
                     const float* b0, const float* w1, const float* b1, const float* w2, const float* b2, float* rgb,
                     uint32_t* masks, void* stream);
int gngf_mlp3_tc_bwd(const float* enc, const float* rgb, const float* drgb, int64_t P, int32_t in_dim, int32_t out_dim,
                     int32_t leaky, const float* w0, const float* b0, const float* w1, const float* b1, const float* w2,
                     const uint32_t* masks, float* denc, float* dw0, float* db0, float* dw1, float* db1, float* dw2,
                     float* db2, void* stream);

/* ---- loss assembly (utils.py:78-174 Loss + functions.py:243-245) with its adjoints, one kernel ---------------
 * total = l_mse * MSE(rgb, target) + sum_l (l_js_kl * level_l + coll_l), level_l = -(gamma+epsilon) JS_l + epsilon KL_l
 * of pbar_l = colsum_l / rows against the uniform distribution.  out (2+L): [0] total, [1] mse, [2+l] level_l;
 * d_rgb (n_rgb) = d total / d rgb; d_colsum (L,N) = d total / d colsum.  coll_term (L) = l_collisions *
 * collisions / (min_possible + delta) or NULL (epoch 0: the scalar 1 per level, functions.py:245).           */
int gngf_loss_fwd_bwd(const float* rgb, const float* target, int64_t n_rgb, const float* colsum, int32_t L, int64_t N,
                      float rows, float gamma, float epsilon, float l_mse, float l_js_kl, const float* coll_term,
                      float* out, float* d_rgb, float* d_colsum, void* stream);
/* the two halves separately (parts: 1 = MSE half -> out[0] +=, out[1], d_rgb; 2 = divergence half -> out[0] +=,
 * out[2 + l], d_colsum; 3 = both), for callers that run them on different streams; `out` must be zero on entry of
 * the first part (no memset inside)                                                                          */
int gngf_loss_parts(const float* rgb, const float* target, int64_t n_rgb, const float* colsum, int32_t L, int64_t N,
                    float rows, float gamma, float epsilon, float l_mse, float l_js_kl, const float* coll_term,
                    float* out, float* d_rgb, float* d_colsum, int32_t parts, void* stream);

/* ---- K3: softmax + nan_to_num + top-k (models.py:85,111 and DifferentiableTopk.forward 7-19) --------
 * logits (R,T) -> probs (R,T) (may alias logits, may be NULL), topv (R,K) sorted descending,
 * topi (R,K) int32, ties broken towards the lower index; row_max / row_sum (R) optional.               */
int gngf_softmax_topk_fwd(const float* logits, int64_t R, int64_t T, int32_t K, float* probs, float* topv,
                          int32_t* topi, float* row_max, float* row_sum, void* stream);
/* top-k only (DifferentiableTopk.forward on given values): int64 indices, API dtype                     */
int gngf_topk_fwd(const float* values, int64_t R, int64_t T, int32_t K, float* topv, int64_t* topi, void* stream);
/* DifferentiableTopk.backward (models.py:21-42): grad_in (R,T) = scatter(zeros, idx, grad_values)       */
int gngf_topk_bwd(const float* grad_values, const int64_t* topi, int64_t R, int64_t T, int32_t K, float* grad_in,
                  void* stream);

/* ---- K4: fused per-level feature lookup + top-k mix + bilinear interpolation -----------------------
 * node features (models.py:194-222 evaluated once per level node):
 *   nfeat[s, f] = mix_k(table_l[utopi[u,k], f], utopv[u,:])                                             */
int gngf_node_features_fwd(gngf_lattice lat, gngf_tables tables, int64_t T, int32_t F, int32_t K, int32_t mix_mode,
                           const float* utopv, const int32_t* utopi, float* nfeat, void* stream);
/* per point: corners + bilinear weights (models.py:621-655) + gather of 4 node features per level:
 *   enc (P, L*F); optional cnt (S) int32 += multiplicity of every level node; err_flag (int32, optional)
 *   is set to 1 when a coordinate falls outside the lattice box (the access is clamped).
 *   cell_cnt (S int32, zero-initialised, optional, needs cnt != NULL): count ONE update per (point, level) -- the
 *   cell, indexed by its floor-corner node -- into cell_cnt instead of four into cnt; the caller then derives
 *   cnt = gngf_cell_to_node_counts(cell_cnt) (a node's multiplicity is the sum of its four surrounding cells).  */
int gngf_encode_fwd(const float* x, int64_t P, gngf_lattice lat, int32_t F, const float* nfeat, float* enc,
                    int32_t* cnt, int32_t* cell_cnt, int32_t* err_flag, void* stream);
int gngf_cell_to_node_counts(gngf_lattice lat, const int32_t* cell_cnt, int32_t* cnt, void* stream);
/* hash-function mode (models.py:181-190 + 504-528): enc straight from table_l[hash(corner)]             */
int gngf_encode_hash_fwd(const float* x, int64_t P, gngf_lattice lat, gngf_tables tables, int64_t T, int32_t F,
                         float* enc, int64_t* idx_out, void* stream);
/* column sums of the (virtual) probability tensor: out (L,N) = sum_s cnt[s] * uvals[u(s), :]
 * (the numerator of p-bar in utils.py:138; uvals = uprobs (U,T) or utopv (U,K)); out must be zeroed      */
int gngf_lattice_colsum(gngf_lattice lat, const int32_t* cnt, const float* uvals, int64_t N, float* out, void* stream);
/* materialise rows of a per-node array: out (P,L,4,N) = uvals[u(p,l,v), :]
 * (f32: top-k probs / full probs -- the module's `probs` output; i32->i64: the `idx_topk` output)        */
int gngf_lattice_gather_rows(const float* x, int64_t P, gngf_lattice lat, const float* uvals, int64_t N, float* out,
                             void* stream);
int gngf_lattice_gather_rows_i64(const float* x, int64_t P, gngf_lattice lat, const int32_t* uvals, int64_t N,
                                 int64_t* out, void* stream);
/* and its adjoint: dvals (U,N) += sum over rows of dout (P,L,4,N)                                       */
int gngf_lattice_scatter_rows(const float* x, int64_t P, gngf_lattice lat, const float* dout, int64_t N, float* dvals,
                              void* stream);

/* ---- K5: backward ----------------------------------------------------------------------------------
 * K5a stage 1, per point: dnf[s, f] += denc[p, l*F+f] * w_bil[p,l,v]   (vector red.global.add)          */
int gngf_encode_bwd(const float* x, int64_t P, gngf_lattice lat, int32_t F, const float* denc, float* dnf,
                    void* stream);
/* K5a stage 2, per level node: table_grad[l][utopi[u,k], f] += dnf[s,f] * w_k ;
 *   dtv[u,k] += mix-backward(dnf . table rows)   (closed form of models.py:212-217's autograd)          */
int gngf_node_features_bwd(gngf_lattice lat, gngf_tables tables, gngf_tables table_grads, int64_t T, int32_t F,
                           int32_t K, int32_t mix_mode, const float* utopv, const int32_t* utopi, const float* dnf,
                           float* dtv, void* stream);
int gngf_encode_hash_bwd(const float* x, int64_t P, gngf_lattice lat, gngf_tables table_grads, int64_t T, int32_t F,
                         const float* denc, void* stream);
/* K5b, per node u in [u0, u0 + n_rows): G = sum_l cnt[s(l,u)] * gcol[l,:] + gdense[u,:] + scatter(g_k at utopi[u,:]),
 *   g_k = dtv[u,k] + sum_l cnt[s(l,u)] * gcol_k[l,k];   dlogit[r,:] = p .* (G - <G,p>)        (r = u - u0)
 * (softmax backward, models.py:85, with DifferentiableTopk.backward, models.py:21-42, and the adjoint of
 * the loss's column sums, utils.py:138, folded in).  gcol (L,T): adjoint of the full-probability column
 * sums; gcol_k (L,K): adjoint of the top-k column sums; gdense (U,T): dense adjoint of uprobs; each may be
 * NULL.  uprobs (n_rows,T) holds the probabilities of the chunk, or -- when row_max/row_sum (U) are given --
 * the LOGITS of the chunk, and p = exp(z - max)/sum is recomputed.  dlogits (n_rows,T) may alias uprobs.   */
int gngf_hpd_dlogits(gngf_lattice lat, const float* uprobs, int64_t T, int32_t K, const int32_t* utopi,
                     const float* dtv, const int32_t* cnt, const float* gcol, const float* gcol_k,
                     const float* gdense, const float* row_max, const float* row_sum, int64_t u0, int64_t n_rows,
                     float* dlogits, void* stream);

/* ---- f-1: calc_hash_collisions (models.py:568-619) ------------------------------------------------------------
 * indices (P,L,V,C) as float32 (train_step's buffer, functions.py:179,216) or int64; uniq (C,L) int32 = number of
 * distinct integer values in [0,range) per (column, level); *outliers is set to 1 when any value is not such an
 * integer (the caller then counts exactly by other means).  bitmap: gngf_count_distinct_workspace_words() words. */
int64_t gngf_count_distinct_workspace_words(int32_t L, int32_t C, int64_t range);
int gngf_count_distinct_f32(const float* indices, int64_t P, int32_t L, int32_t V, int32_t C, int64_t range,
                            uint32_t* bitmap, int32_t* uniq, int32_t* outliers, void* stream);
int gngf_count_distinct_i64(const int64_t* indices, int64_t P, int32_t L, int32_t V, int32_t C, int64_t range,
                            uint32_t* bitmap, int32_t* uniq, int32_t* outliers, void* stream);

/* ---- f-4: _calc_counts_per_level (models.py:530-566) -----------------------------------------------------------
 * grid (P,2,L,4) fp32 grid corners (gngf_corners_fwd); hashed: int64 slot of (point p, level l, corner v) at
 * hashed[((p*L + l)*4 + v) * hashed_stride] (stride K selects the best top-k column of a (P,L,4,K) index tensor, 1 a
 * (P,L,4) hash tensor).  hist (L,T) int32 = per level, for every DISTINCT grid cell, one count for the slot
 * hashed_flat[l][j], j = index of the first point (batch order) in that cell -- the reference's np.unique(axis=0,
 * return_index=True) + flattened "(p v)" indexing, as is.  first: lat.loff[L] int32 of scratch; *outliers = 1 when a
 * corner lies outside the lattice box or a slot outside [0,T) (the caller then counts on the host).               */
int gngf_counts_per_level(const float* grid, int64_t P, gngf_lattice lat, const int64_t* hashed, int64_t hashed_stride,
                          int64_t T, int32_t* first, int32_t* hist, int32_t* outliers, void* stream);

/* ---- f-3: fused Adam over all parameter tensors (functions.py:96-127, 281) ------------------------------------------
 * One launch: for every tensor  g' = g + weight_decay p;  m += (1-beta1)(g' - m);  v = beta2 v + (1-beta2) g'^2;
 *   p -= lr / (1 - beta1^t) * m / (sqrt(v) / sqrt(1 - beta2^t) + eps)   with t = *step + 1  (torch.optim.Adam semantics,
 * amsgrad off).  `tensors` is a HOST array (copied into kernel-parameter space); p/g/m/v/step are device pointers;
 * every tensor has its own `step` (device float32 scalar holding an integer value, as in torch.optim.Adam's
 * state dict; distinct addresses), advanced by the kernel; `ticket` (device
 * uint32, zero-initialised) is scratch.                                                                             */
#define GNGF_ADAM_MAX_TENSORS 64
typedef struct gngf_adam_tensor {
  float* p;
  const float* g;
  float* m;
  float* v;
  float* step;
  int64_t n;
  float lr;
  float weight_decay;
} gngf_adam_tensor;
int gngf_adam_step(const gngf_adam_tensor* tensors, int32_t count, float beta1, float beta2, float eps, uint32_t* ticket,
                   void* stream);

/* ---- 8e: one-shot all-reduce over NVLink peer memory for small buffers (k10_allreduce.cu) -------------------------
 * out (n) = scale * sum over ranks of in (n), identical bits on every rank (fixed summation order).
 * stage_ptrs_dev / signal_ptrs_dev: DEVICE arrays of `world` peer-mapped pointers (symmetric memory): staging buffers
 *   of 2 * cap_floats floats each, and zero-initialised signal pads of at least max_blocks * world uint32 each.
 * state: 3 device uint32 (epoch, ticket, error), zero-initialised, owned by this communicator.  Every rank must call
 *   with the same n and max_blocks, in the same order, and reach the call within the timeout (default ~30 s,
 *   gngf_peer_allreduce_set_timeout_ms).  A peer that does not arrive is FATAL for the result: state[2] becomes 1 and
 *   stays 1, and the output of that call and of every later call is NaN -- never stale or partial sums.
 *   in / out 16-byte aligned; out may alias in.  Asynchronous on `stream`, CUDA-graph capturable.                  */
int gngf_peer_allreduce(const void* stage_ptrs_dev, const void* signal_ptrs_dev, int32_t rank, int32_t world,
                        const float* in, float* out, int64_t n, int64_t cap_floats, int32_t max_blocks, float scale,
                        uint32_t* state, void* stream);
/* Host-side setting, applies to launches made after it (process-wide). */
int gngf_peer_allreduce_set_timeout_ms(int64_t ms);

#ifdef __cplusplus
}
#endif
#endif /* GNGF_H */
