// K11: the set of lattice nodes a batch touches ("active nodes").
//
// The HPD is evaluated per lattice node (models.py:416-418 feeds it one row per (point, level, corner), but the rows are
// integer corner coordinates and the network is shared by the levels).  On a small lattice every node of the bounding box
// is touched; on the 8192^2 lattice of BASELINE.json configs[3] a batch of 2^22 points touches ~40 % of the 67 M nodes --
// and a data-parallel rank with 1/N of the batch a fraction of that.  So the path marks the touched nodes in a bitmap
// (8 MB for 67 M nodes: L2-resident), compacts the bitmap into an ascending list of node ids, evaluates the HPD chain
// on that list only, and scatters the per-node results back into the full arrays the gather / scatter kernels index.
//
//   gngf_lattice_mark_nodes   thread per point, warp-uniform level loop; 4 corners -> atomicOr (skipped when the bit is set)
//   gngf_compact_nodes        popcount per 4096-node chunk -> one-block exclusive scan of the chunk sums -> ordered write
//   gngf_scatter_node_rows    dst[node_ids[r], :] = src[r, :]  (32-bit elements)
#include <algorithm>

#include "common.cuh"

namespace gngf {

constexpr int CHUNK_THREADS = 256;
constexpr int WORDS_PER_THREAD = 4;
constexpr int CHUNK_WORDS = CHUNK_THREADS * WORDS_PER_THREAD;   // 1024 words = 32768 nodes per block

__global__ void __launch_bounds__(256) mark_nodes_kernel(const float2* __restrict__ x, int64_t P,
                                                         const __grid_constant__ gngf_lattice lat,
                                                         unsigned* __restrict__ bitmap) {
  const int64_t p = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (p >= P) return;
  const float2 xy = __ldg(x + p);
  const int L = lat.num_levels;
  for (int l = 0; l < L; ++l) {
    const Cell c = cell_of(xy.x, xy.y, lat.n[l]);
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int64_t u = global_node(lat, c.cx + (v & 1), c.cy + (v >> 1));
      unsigned* word = bitmap + (u >> 5);
      const unsigned bit = 1u << (u & 31);
      if (!(__ldcg(word) & bit)) atomicOr(word, bit);
    }
  }
}

__device__ __forceinline__ uint4 load_words(const unsigned* __restrict__ bitmap, int64_t w0, int64_t W) {
  uint4 v = make_uint4(0u, 0u, 0u, 0u);
  if (w0 + 3 < W) {
    v = *reinterpret_cast<const uint4*>(bitmap + w0);
  } else {
    if (w0 < W) v.x = bitmap[w0];
    if (w0 + 1 < W) v.y = bitmap[w0 + 1];
    if (w0 + 2 < W) v.z = bitmap[w0 + 2];
  }
  return v;
}

// block-wide exclusive scan of one int per thread (256 threads); returns the exclusive prefix, *total = block sum
__device__ __forceinline__ int block_exclusive_scan(int v, int* smem_warp, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) smem_warp[warp] = inc;
  __syncthreads();
  int base = 0, sum = 0;
#pragma unroll
  for (int i = 0; i < CHUNK_THREADS / 32; ++i) {
    const int s = smem_warp[i];
    if (i < warp) base += s;
    sum += s;
  }
  *total = sum;
  __syncthreads();
  return base + inc - v;
}

__global__ void __launch_bounds__(CHUNK_THREADS) chunk_popcount_kernel(const unsigned* __restrict__ bitmap, int64_t W,
                                                                       int* __restrict__ chunk_sums) {
  __shared__ int sw[CHUNK_THREADS / 32];
  const int64_t w0 = (static_cast<int64_t>(blockIdx.x) * CHUNK_THREADS + threadIdx.x) * WORDS_PER_THREAD;
  const uint4 v = load_words(bitmap, w0, W);
  int c = __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) sw[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
#pragma unroll
    for (int i = 0; i < CHUNK_THREADS / 32; ++i) s += sw[i];
    chunk_sums[blockIdx.x] = s;
  }
}

// one block: chunk_sums (n) -> exclusive offsets in place; count[0] = total
__global__ void __launch_bounds__(CHUNK_THREADS) chunk_scan_kernel(int* __restrict__ chunk_sums, int n,
                                                                   int* __restrict__ count) {
  __shared__ int sw[CHUNK_THREADS / 32];
  int carry = 0;
  for (int base = 0; base < n; base += CHUNK_THREADS) {
    const int i = base + threadIdx.x;
    const int v = i < n ? chunk_sums[i] : 0;
    int total;
    const int ex = block_exclusive_scan(v, sw, &total);
    if (i < n) chunk_sums[i] = carry + ex;
    carry += total;
  }
  if (threadIdx.x == 0) count[0] = carry;
}

__global__ void __launch_bounds__(CHUNK_THREADS) chunk_write_kernel(const unsigned* __restrict__ bitmap, int64_t W,
                                                                    const int* __restrict__ chunk_offsets,
                                                                    int64_t capacity, int* __restrict__ node_ids) {
  __shared__ int sw[CHUNK_THREADS / 32];
  const int64_t w0 = (static_cast<int64_t>(blockIdx.x) * CHUNK_THREADS + threadIdx.x) * WORDS_PER_THREAD;
  const uint4 v = load_words(bitmap, w0, W);
  const unsigned words[4] = {v.x, v.y, v.z, v.w};
  const int c = __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);
  int total;
  int64_t o = chunk_offsets[blockIdx.x] + static_cast<int64_t>(block_exclusive_scan(c, sw, &total));
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    unsigned m = words[i];
    const int base = static_cast<int>((w0 + i) << 5);
    while (m) {
      const int b = __ffs(m) - 1;
      m &= m - 1;
      if (o < capacity) node_ids[o] = base + b;
      ++o;
    }
  }
}

__global__ void __launch_bounds__(256) scatter_node_rows_kernel(const int* __restrict__ node_ids, int64_t n, int N,
                                                                const unsigned* __restrict__ src,
                                                                unsigned* __restrict__ dst) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n * N) return;
  const int64_t r = i / N;
  const int c = static_cast<int>(i - r * N);
  dst[static_cast<int64_t>(node_ids[r]) * N + c] = src[i];
}

// Node-parallel HPD (dp.NodeSharding): every rank marks the nodes ITS points touch; the union over ranks is the list all
// ranks agree on.  out[w] = OR_r maps[r * W + w]   (maps: the all-gathered per-rank bitmaps, 16 bytes per thread).
__global__ void __launch_bounds__(256) bitmap_or_kernel(const uint4* __restrict__ maps, int n_maps, int64_t W4,
                                                        uint4* __restrict__ out) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= W4) return;
  uint4 acc = maps[i];
  for (int r = 1; r < n_maps; ++r) {
    const uint4 v = maps[static_cast<int64_t>(r) * W4 + i];
    acc.x |= v.x;
    acc.y |= v.y;
    acc.z |= v.z;
    acc.w |= v.w;
  }
  out[i] = acc;
}

// Node-parallel HPD, backward: this rank's share of the adjoint of the selected probabilities, per row of the agreed
// node list:  out[r,k] = dtv[node_ids[r],k] + sum_l cnt[s(l, node)] gcol_k[l,k]   (the column-sum adjoint folded in:
// it is linear in this rank's multiplicities).  The rows are then summed over ranks and scattered to their owners
// (reduce-scatter); the owner's streaming backward takes them row-indexed.
__global__ void __launch_bounds__(256)
    gather_node_adjoints_kernel(const __grid_constant__ gngf_lattice lat, const int* __restrict__ node_ids, int64_t n, int K,
                                const float* __restrict__ dtv, const int* __restrict__ cnt,
                                const float* __restrict__ gcol_k, float* __restrict__ out) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n * K) return;
  const int64_t r = i / K;
  const int k = static_cast<int>(i - r * K);
  const int64_t un = node_ids ? node_ids[r] : r;
  float g = dtv[un * K + k];
  if (gcol_k) {
    const int cx = lat.ox + static_cast<int>(un / lat.wy), cy = lat.oy + static_cast<int>(un % lat.wy);
    for (int l = 0; l < lat.num_levels; ++l) {
      const int a = cx - lat.lox[l], b = cy - lat.loy[l];
      if (a >= 0 && a < lat.lwx[l] && b >= 0 && b < lat.lwy[l]) {
        const float c = static_cast<float>(cnt[lat.loff[l] + static_cast<int64_t>(a) * lat.lwy[l] + b]);
        g = fmaf(c, gcol_k[l * K + k], g);
      }
    }
  }
  out[i] = g;
}

}  // namespace gngf

extern "C" {

int gngf_bitmap_or(const uint32_t* maps, int32_t n_maps, int64_t words, uint32_t* out, void* stream) {
  if (n_maps < 1 || words < 0 || (words & 3) || !maps || !out) return GNGF_ERR_INVALID_ARGUMENT;
  if ((reinterpret_cast<uintptr_t>(maps) | reinterpret_cast<uintptr_t>(out)) & 15) return GNGF_ERR_INVALID_ARGUMENT;
  if (words == 0) return GNGF_OK;
  const int64_t W4 = words / 4;
  gngf::bitmap_or_kernel<<<static_cast<unsigned>(gngf::ceil_div(W4, 256)), 256, 0, gngf::as_stream(stream)>>>(
      reinterpret_cast<const uint4*>(maps), n_maps, W4, reinterpret_cast<uint4*>(out));
  gngf::note_launch();
  return gngf::check_launch();
}

int gngf_gather_node_adjoints(gngf_lattice lat, const int32_t* node_ids, int64_t n_nodes, int32_t K, const float* dtv,
                              const int32_t* cnt, const float* gcol_k, float* out, void* stream) {
  if (n_nodes < 0 || K <= 0 || K > GNGF_MAX_TOPK || lat.num_levels <= 0 || lat.num_levels > GNGF_MAX_LEVELS)
    return GNGF_ERR_INVALID_ARGUMENT;
  if (n_nodes == 0) return GNGF_OK;
  if (!dtv || !out || (gcol_k && !cnt)) return GNGF_ERR_INVALID_ARGUMENT;
  gngf::gather_node_adjoints_kernel<<<static_cast<unsigned>(gngf::ceil_div(n_nodes * K, 256)), 256, 0,
                                      gngf::as_stream(stream)>>>(lat, node_ids, n_nodes, K, dtv, cnt, gcol_k, out);
  gngf::note_launch();
  return gngf::check_launch();
}

int64_t gngf_active_nodes_bitmap_words(int64_t U) { return U <= 0 ? 0 : (U + 31) / 32; }
int64_t gngf_active_nodes_chunks(int64_t U) {
  return U <= 0 ? 0 : gngf::ceil_div(gngf_active_nodes_bitmap_words(U), gngf::CHUNK_WORDS);
}

int gngf_lattice_mark_nodes(const float* x, int64_t P, gngf_lattice lat, uint32_t* bitmap, void* stream) {
  if (P < 0 || !bitmap || (P > 0 && !x) || lat.num_levels <= 0 || lat.num_levels > GNGF_MAX_LEVELS)
    return GNGF_ERR_INVALID_ARGUMENT;
  if (P == 0) return GNGF_OK;
  gngf::mark_nodes_kernel<<<static_cast<unsigned>(gngf::ceil_div(P, 256)), 256, 0, gngf::as_stream(stream)>>>(
      reinterpret_cast<const float2*>(x), P, lat, bitmap);
  gngf::note_launch();
  return gngf::check_launch();
}

int gngf_compact_nodes(const uint32_t* bitmap, int64_t U, int32_t* chunk_offsets, int32_t* node_ids, int64_t capacity,
                       int32_t* count, void* stream) {
  if (U <= 0 || U >= (1ll << 31) || capacity < 0 || !bitmap || !chunk_offsets || !node_ids || !count)
    return GNGF_ERR_INVALID_ARGUMENT;
  if (reinterpret_cast<uintptr_t>(bitmap) & 15) return GNGF_ERR_INVALID_ARGUMENT;
  cudaStream_t st = gngf::as_stream(stream);
  const int64_t W = gngf_active_nodes_bitmap_words(U);
  const int64_t chunks = gngf_active_nodes_chunks(U);
  gngf::chunk_popcount_kernel<<<static_cast<unsigned>(chunks), gngf::CHUNK_THREADS, 0, st>>>(bitmap, W, chunk_offsets);
  gngf::note_launch();
  int rc = gngf::check_launch();
  if (rc) return rc;
  gngf::chunk_scan_kernel<<<1, gngf::CHUNK_THREADS, 0, st>>>(chunk_offsets, static_cast<int>(chunks), count);
  gngf::note_launch();
  if ((rc = gngf::check_launch())) return rc;
  gngf::chunk_write_kernel<<<static_cast<unsigned>(chunks), gngf::CHUNK_THREADS, 0, st>>>(bitmap, W, chunk_offsets, capacity,
                                                                                        node_ids);
  gngf::note_launch();
  return gngf::check_launch();
}

int gngf_scatter_node_rows(const int32_t* node_ids, int64_t n_nodes, const void* src, int64_t row_words, void* dst,
                           void* stream) {
  if (n_nodes < 0 || row_words <= 0 || row_words >= (1 << 20)) return GNGF_ERR_INVALID_ARGUMENT;
  if (n_nodes == 0) return GNGF_OK;
  if (!node_ids || !src || !dst) return GNGF_ERR_INVALID_ARGUMENT;
  gngf::scatter_node_rows_kernel<<<static_cast<unsigned>(gngf::ceil_div(n_nodes * row_words, 256)), 256, 0,
                                   gngf::as_stream(stream)>>>(node_ids, n_nodes, static_cast<int>(row_words),
                                                              static_cast<const unsigned*>(src),
                                                              static_cast<unsigned*>(dst));
  gngf::note_launch();
  return gngf::check_launch();
}

}  // extern "C"
