// K4: the feature lookup of MultiResHashEncoding.forward (models.py:194-222) and _bilinear_interpolate
// (models.py:621-655), restructured around the lattice:
//
//   node pass   nfeat[s,:] = mix_k(table_l[utopi[u,k],:], utopv[u,:])     once per level node s = (l, cx, cy)
//   point pass  enc[p, l*F+f] = sum_v w_bil[p,l,v] * nfeat[s(p,l,v), f]   4 gathers per (point, level)
//
// Every (point, level, corner) row of the reference that lands on the same node has the same top-k slots and
// probabilities, so mixing K table rows per *row* (the reference) and per *node* (here) is the same
// arithmetic done once.  The point pass is an HBM/L2-bound gather: one thread per (point, level), the
// (P, L*F) output row written as contiguous float2.
#include <algorithm>

#include "common.cuh"

namespace gngf {

// grid (ceil(max level box / 256), L); thread per level node
__global__ void __launch_bounds__(256)
    node_features_fwd_kernel(const __grid_constant__ gngf_lattice lat, const __grid_constant__ gngf_tables tables,
                             int64_t T, int F, int K, int mode, const float* __restrict__ utopv,
                             const int32_t* __restrict__ utopi, float* __restrict__ nfeat) {
  const int l = blockIdx.y;
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int wy = lat.lwy[l];
  if (i >= static_cast<int64_t>(lat.lwx[l]) * wy) return;
  const int cx = lat.lox[l] + static_cast<int>(i / wy), cy = lat.loy[l] + static_cast<int>(i % wy);
  const int64_t u = global_node(lat, cx, cy);
  const float* tv = utopv + u * K;
  const int32_t* ti = utopi + u * K;
  const float* table = tables.ptr[l];
  float mx;
  const float norm = mix_weight_norm(tv, K, mode, mx);
  float acc[GNGF_MAX_FEATURES];
#pragma unroll
  for (int f = 0; f < GNGF_MAX_FEATURES; ++f) acc[f] = 0.0f;
  for (int k = 0; k < K; ++k) {
    // weighted-average mode follows the reference's op order: (sum_k g*p) / (sum_k p)
    const float w = (mode == GNGF_MIX_WEIGHTED_AVG) ? tv[k] : mix_weight(tv[k], mode, mx, norm);
    const float* row = table + static_cast<int64_t>(ti[k]) * F;
#pragma unroll
    for (int f = 0; f < GNGF_MAX_FEATURES; ++f)
      if (f < F) acc[f] = fmaf(row[f], w, acc[f]);
  }
  float* out = nfeat + (lat.loff[l] + i) * F;
#pragma unroll
  for (int f = 0; f < GNGF_MAX_FEATURES; ++f)
    if (f < F) out[f] = (mode == GNGF_MIX_WEIGHTED_AVG) ? acc[f] / norm : acc[f];
}

// Persistent grid, one thread per point, levels walked by an (unrolled) loop: the level index is uniform across
// the warp, so the per-level geometry comes from the constant bank as a broadcast (a lane-per-level mapping
// serialises those reads 16-way), the 4*L gathers of a point are independent loads in flight together, and on
// the coarse levels the 32 points of a warp fall into a handful of cache lines.
// The node multiplicities `cnt` are a histogram with heavy same-address traffic on the coarse levels (level 0 of
// the published configuration: 3 280 increments per node and batch; an L2 atomic unit retires roughly one
// same-address update per 50-80 cycles), so the first `private_nodes` level nodes -- the coarsest levels, which
// are stored first -- are counted in shared memory and flushed once per CTA; finer levels go straight to global
// atomics where contention is naturally low.
template <int F>
__global__ void __launch_bounds__(256)
    encode_fwd_kernel(const float2* __restrict__ x, int64_t P, const __grid_constant__ gngf_lattice lat,
                      const float* __restrict__ nfeat, float* __restrict__ enc, int32_t* __restrict__ cnt,
                      int32_t* __restrict__ err_flag, int private_nodes, int cells) {
  // cells != 0: `cnt` receives CELL counts (one update per (point, level), indexed by the cell's floor corner) and
  // cell_to_node_counts_kernel turns them into node multiplicities: a node's count is the sum of its four
  // surrounding cells.  At 2^22 points x 16 levels the four per-corner atomics doubled the pass (2.07 vs 1.00 ms).
  extern __shared__ int32_t cnt_s[];
  const int L = lat.num_levels;
  const bool priv = cnt != nullptr && private_nodes > 0;
  if (priv) {
    for (int i = threadIdx.x; i < private_nodes; i += blockDim.x) cnt_s[i] = 0;
    __syncthreads();
  }
  bool outside = false;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t p = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; p < P; p += stride) {
    const float2 xy = __ldcs(x + p);     // streamed once: keep the (L2-resident) node features in the cache instead
    float* out = enc + p * (L * F);
#pragma unroll 4
    for (int l = 0; l < L; ++l) {
      const Cell c = cell_of(xy.x, xy.y, lat.n[l]);
      int64_t s[4];
#pragma unroll
      for (int v = 0; v < 4; ++v) s[v] = level_node(lat, l, c.cx + (v & 1), c.cy + (v >> 1), outside);
      if constexpr (F == 2) {
        float2 nf[4];
#pragma unroll
        for (int v = 0; v < 4; ++v) nf[v] = __ldg(reinterpret_cast<const float2*>(nfeat) + s[v]);
        float a0 = 0.0f, a1 = 0.0f;
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          a0 = fmaf(nf[v].x, c.w[v], a0);
          a1 = fmaf(nf[v].y, c.w[v], a1);
        }
        __stcs(reinterpret_cast<float2*>(out) + l, make_float2(a0, a1));
      } else {
        float acc[F];
#pragma unroll
        for (int f = 0; f < F; ++f) acc[f] = 0.0f;
#pragma unroll
        for (int v = 0; v < 4; ++v)
#pragma unroll
          for (int f = 0; f < F; ++f) acc[f] = fmaf(__ldg(nfeat + s[v] * F + f), c.w[v], acc[f]);
#pragma unroll
        for (int f = 0; f < F; ++f) out[l * F + f] = acc[f];
      }
      if (cnt) {
        if (cells) {
          if (s[0] < private_nodes) atomicAdd(cnt_s + s[0], 1);
          else atomicAdd(cnt + s[0], 1);
        } else {
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            if (s[v] < private_nodes) atomicAdd(cnt_s + s[v], 1);
            else atomicAdd(cnt + s[v], 1);
          }
        }
      }
    }
  }
  if (outside && err_flag) *err_flag = 1;
  if (priv) {
    __syncthreads();
    for (int i = threadIdx.x; i < private_nodes; i += blockDim.x) {
      const int c = cnt_s[i];
      if (c) atomicAdd(cnt + i, c);
    }
  }
}

// node multiplicities from cell counts: cnt[node (i, j) of level l] = cells (i, j) + (i-1, j) + (i, j-1) + (i-1, j-1),
// a cell being indexed by its floor-corner node.  grid (ceil(max box / 256), L)
__global__ void __launch_bounds__(256)
    cell_to_node_counts_kernel(const __grid_constant__ gngf_lattice lat, const int32_t* __restrict__ cell_cnt,
                               int32_t* __restrict__ cnt) {
  const int l = blockIdx.y;
  const int wx = lat.lwx[l], wy = lat.lwy[l];
  const int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (e >= static_cast<int64_t>(wx) * wy) return;
  const int i = static_cast<int>(e / wy), j = static_cast<int>(e % wy);
  const int32_t* cc = cell_cnt + lat.loff[l];
  int c = cc[e];
  if (i > 0) c += cc[e - wy];
  if (j > 0) c += cc[e - 1];
  if (i > 0 && j > 0) c += cc[e - wy - 1];
  cnt[lat.loff[l] + e] = c;
}

// hash-function mode: enc straight from table_l[hash(corner)], optional idx output (P,L,4) int64;
// thread per point, loop over levels (see encode_fwd_kernel)
template <int F>
__global__ void __launch_bounds__(256)
    encode_hash_fwd_kernel(const float2* __restrict__ x, int64_t P, const __grid_constant__ gngf_lattice lat,
                           const __grid_constant__ gngf_tables tables, int64_t T, float* __restrict__ enc,
                           int64_t* __restrict__ idx_out) {
  const int L = lat.num_levels;
  const bool pow2 = (T & (T - 1)) == 0;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t p = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; p < P; p += stride) {
    const float2 xy = x[p];
#pragma unroll 4
    for (int l = 0; l < L; ++l) {
      const Cell c = cell_of(xy.x, xy.y, lat.n[l]);
      const float* table = tables.ptr[l];
      float acc[F];
#pragma unroll
      for (int f = 0; f < F; ++f) acc[f] = 0.0f;
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const uint32_t gx = static_cast<uint32_t>(c.cx + (v & 1)), gy = static_cast<uint32_t>(c.cy + (v >> 1));
        const uint32_t h32 = gx ^ (gy * 2654435761u);
        int64_t h;
        if (pow2) {
          h = static_cast<int64_t>(h32 & static_cast<uint32_t>(T - 1));   // == non-negative remainder of the int32
        } else {
          h = static_cast<int64_t>(static_cast<int32_t>(h32)) % T;
          if (h < 0) h += T;
        }
        if (idx_out) idx_out[(p * L + l) * 4 + v] = h;
        if constexpr (F == 2) {
          const float2 t = __ldg(reinterpret_cast<const float2*>(table) + h);
          acc[0] = fmaf(t.x, c.w[v], acc[0]);
          acc[1] = fmaf(t.y, c.w[v], acc[1]);
        } else {
#pragma unroll
          for (int f = 0; f < F; ++f) acc[f] = fmaf(__ldg(table + h * F + f), c.w[v], acc[f]);
        }
      }
#pragma unroll
      for (int f = 0; f < F; ++f) enc[(p * L + l) * F + f] = acc[f];
    }
  }
}

// out[l, n] += sum over level nodes s of level l: cnt[s] * uvals[u(s), n]
// grid (ceil(N/128), node chunks, L), block 128: thread = one column n.  The nodes of a chunk are staged 32 at a time
// (multiplicity + lattice node in shared memory, untouched nodes dropped), so that the column loop is a run of
// independent, coalesced row reads: the first version chained two dependent L2 reads per node (cnt -> branch -> row),
// 32 nodes deep -- 16 us of pure latency at the published configuration, on the branch the loss waits for.
__global__ void __launch_bounds__(128)
    lattice_colsum_kernel(const __grid_constant__ gngf_lattice lat, const int32_t* __restrict__ cnt,
                          const float* __restrict__ uvals, int64_t N, int64_t nodes_per_block,
                          float* __restrict__ out) {
  __shared__ float c_s[32];
  __shared__ int64_t u_s[32];
  __shared__ int n_s;
  const int l = blockIdx.z;
  const int64_t n = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int wy = lat.lwy[l];
  const int64_t box = static_cast<int64_t>(lat.lwx[l]) * wy;
  const int64_t i0 = static_cast<int64_t>(blockIdx.y) * nodes_per_block;
  const int64_t i1 = min(box, i0 + nodes_per_block);
  if (i0 >= box) return;   // block-uniform
  float acc = 0.0f;
  for (int64_t base = i0; base < i1; base += 32) {
    if (threadIdx.x < 32) {   // warp 0: compact the touched nodes of this group
      const int64_t i = base + threadIdx.x;
      const int c = i < i1 ? cnt[lat.loff[l] + i] : 0;
      const unsigned live = __ballot_sync(0xffffffffu, c != 0);
      if (c != 0) {
        const int slot = __popc(live & ((1u << threadIdx.x) - 1u));
        c_s[slot] = static_cast<float>(c);
        u_s[slot] = global_node(lat, lat.lox[l] + static_cast<int>(i / wy), lat.loy[l] + static_cast<int>(i % wy));
      }
      if (threadIdx.x == 0) n_s = __popc(live);
    }
    __syncthreads();
    const int m = n_s;
    if (n < N) {
#pragma unroll 8
      for (int k = 0; k < m; ++k) acc = fmaf(c_s[k], __ldg(uvals + u_s[k] * N + n), acc);
    }
    __syncthreads();
  }
  if (n < N) atomicAdd(out + l * N + n, acc);
}

// the same for a narrow value row (top-k-only mode: N = K <= 8): thread per level node, N register accumulators,
// block reduction, one atomic per (block, column).  The column-per-thread kernel above leaves all but N threads idle
// (315 ms at the 8192^2 lattice, 119 M level nodes).  grid (node blocks, L)
template <int NMAX>
__global__ void __launch_bounds__(256)
    lattice_colsum_narrow_kernel(const __grid_constant__ gngf_lattice lat, const int32_t* __restrict__ cnt,
                                 const float* __restrict__ uvals, int N, float* __restrict__ out) {
  const int l = blockIdx.y;
  const int wy = lat.lwy[l];
  const int64_t box = static_cast<int64_t>(lat.lwx[l]) * wy;
  float acc[NMAX];
#pragma unroll
  for (int n = 0; n < NMAX; ++n) acc[n] = 0.0f;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < box;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = cnt[lat.loff[l] + i];
    if (c == 0) continue;
    const int64_t u = global_node(lat, lat.lox[l] + static_cast<int>(i / wy), lat.loy[l] + static_cast<int>(i % wy));
    const float cf = static_cast<float>(c);
#pragma unroll
    for (int n = 0; n < NMAX; ++n)
      if (n < N) acc[n] = fmaf(cf, uvals[u * N + n], acc[n]);
  }
  __shared__ float red[8][NMAX];
  const int lane = threadIdx.x % 32, w = threadIdx.x / 32;
#pragma unroll
  for (int n = 0; n < NMAX; ++n) {
    const float s = warp_sum(acc[n]);
    if (lane == 0) red[w][n] = s;
  }
  __syncthreads();
  if (threadIdx.x < N) {
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[i][threadIdx.x];
    if (s != 0.0f) atomicAdd(out + l * N + threadIdx.x, s);
  }
}

// out (P,L,4,N) = uvals[u(p,l,v), :]; thread per output element
template <typename Tin, typename Tout>
__global__ void __launch_bounds__(256)
    gather_rows_kernel(const float2* __restrict__ x, int64_t P, const __grid_constant__ gngf_lattice lat,
                       const Tin* __restrict__ uvals, int64_t N, Tout* __restrict__ out) {
  const int L = lat.num_levels;
  const int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (e >= P * L * 4 * N) return;
  const int64_t row = e / N;
  const int64_t k = e - row * N;
  const int v = static_cast<int>(row & 3);
  const int64_t pl = row >> 2;
  const int64_t p = pl / L;
  const int l = static_cast<int>(pl - p * L);
  const float2 xy = x[p];
  const Cell c = cell_of(xy.x, xy.y, lat.n[l]);
  const int64_t u = global_node(lat, c.cx + (v & 1), c.cy + (v >> 1));
  out[e] = static_cast<Tout>(uvals[u * N + k]);
}

// adjoint of gather_rows (f32): dvals[u(p,l,v), k] += dout[p,l,v,k]
__global__ void __launch_bounds__(256)
    scatter_rows_kernel(const float2* __restrict__ x, int64_t P, const __grid_constant__ gngf_lattice lat,
                        const float* __restrict__ dout, int64_t N, float* __restrict__ dvals) {
  const int L = lat.num_levels;
  const int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (e >= P * L * 4 * N) return;
  const int64_t row = e / N;
  const int64_t k = e - row * N;
  const int v = static_cast<int>(row & 3);
  const int64_t pl = row >> 2;
  const int64_t p = pl / L;
  const int l = static_cast<int>(pl - p * L);
  const float2 xy = x[p];
  const Cell c = cell_of(xy.x, xy.y, lat.n[l]);
  const int64_t u = global_node(lat, c.cx + (v & 1), c.cy + (v >> 1));
  const float g = dout[e];
  if (g != 0.0f) atomicAdd(dvals + u * N + k, g);
}

static int valid_lat(const gngf_lattice& lat) {
  return lat.num_levels > 0 && lat.num_levels <= GNGF_MAX_LEVELS && lat.wx > 0 && lat.wy > 0;
}
static int64_t max_level_box(const gngf_lattice& lat) {
  int64_t m = 0;
  for (int l = 0; l < lat.num_levels; ++l) m = std::max<int64_t>(m, static_cast<int64_t>(lat.lwx[l]) * lat.lwy[l]);
  return m;
}

}  // namespace gngf

extern "C" {

int gngf_node_features_fwd(gngf_lattice lat, gngf_tables tables, int64_t T, int32_t F, int32_t K, int32_t mix_mode,
                           const float* utopv, const int32_t* utopi, float* nfeat, void* stream) {
  if (!gngf::valid_lat(lat) || F <= 0 || F > GNGF_MAX_FEATURES || K <= 0 || K > GNGF_MAX_TOPK || T <= 0)
    return GNGF_ERR_INVALID_ARGUMENT;
  dim3 grid(static_cast<unsigned>(gngf::ceil_div(gngf::max_level_box(lat), 256)), lat.num_levels);
  gngf::node_features_fwd_kernel<<<grid, 256, 0, gngf::as_stream(stream)>>>(lat, tables, T, F, K, mix_mode, utopv,
                                                                            utopi, nfeat);
  gngf::note_launch();
  return gngf::check_launch();
}

int gngf_encode_fwd(const float* x, int64_t P, gngf_lattice lat, int32_t F, const float* nfeat, float* enc,
                    int32_t* cnt, int32_t* cell_cnt, int32_t* err_flag, void* stream) {
  if (!gngf::valid_lat(lat) || P < 0 || F <= 0 || F > GNGF_MAX_FEATURES) return GNGF_ERR_INVALID_ARGUMENT;
  if (cell_cnt && !cnt) return GNGF_ERR_INVALID_ARGUMENT;
  if (P == 0) return GNGF_OK;
  cudaStream_t st = gngf::as_stream(stream);
  const float2* x2 = reinterpret_cast<const float2*>(x);
  const int priv = cnt ? gngf::private_node_count(lat, 10240) : 0;      // <= 40 KB of shared counters
  const size_t smem = sizeof(int32_t) * priv;
  const unsigned blocks = static_cast<unsigned>(std::min<int64_t>(gngf::ceil_div(P, 256), 8 * gngf::sm_count()));
  int32_t* counters = cell_cnt ? cell_cnt : cnt;
  const int cells = cell_cnt != nullptr;
  switch (F) {
    case 1: gngf::encode_fwd_kernel<1><<<blocks, 256, smem, st>>>(x2, P, lat, nfeat, enc, counters, err_flag, priv, cells); break;
    case 2: gngf::encode_fwd_kernel<2><<<blocks, 256, smem, st>>>(x2, P, lat, nfeat, enc, counters, err_flag, priv, cells); break;
    case 4: gngf::encode_fwd_kernel<4><<<blocks, 256, smem, st>>>(x2, P, lat, nfeat, enc, counters, err_flag, priv, cells); break;
    case 8: gngf::encode_fwd_kernel<8><<<blocks, 256, smem, st>>>(x2, P, lat, nfeat, enc, counters, err_flag, priv, cells); break;
    default: return GNGF_ERR_UNSUPPORTED;
  }
  gngf::note_launch();
  return gngf::check_launch();
}

int gngf_cell_to_node_counts(gngf_lattice lat, const int32_t* cell_cnt, int32_t* cnt, void* stream) {
  if (!gngf::valid_lat(lat) || !cell_cnt || !cnt) return GNGF_ERR_INVALID_ARGUMENT;
  dim3 grid(static_cast<unsigned>(gngf::ceil_div(gngf::max_level_box(lat), 256)), lat.num_levels);
  gngf::cell_to_node_counts_kernel<<<grid, 256, 0, gngf::as_stream(stream)>>>(lat, cell_cnt, cnt);
  gngf::note_launch();
  return gngf::check_launch();
}

int gngf_encode_hash_fwd(const float* x, int64_t P, gngf_lattice lat, gngf_tables tables, int64_t T, int32_t F,
                         float* enc, int64_t* idx_out, void* stream) {
  if (!gngf::valid_lat(lat) || P < 0 || T <= 0) return GNGF_ERR_INVALID_ARGUMENT;
  if (P == 0) return GNGF_OK;
  const unsigned blocks = static_cast<unsigned>(std::min<int64_t>(gngf::ceil_div(P, 256), 8 * gngf::sm_count()));
  cudaStream_t st = gngf::as_stream(stream);
  const float2* x2 = reinterpret_cast<const float2*>(x);
  switch (F) {
    case 1: gngf::encode_hash_fwd_kernel<1><<<blocks, 256, 0, st>>>(x2, P, lat, tables, T, enc, idx_out); break;
    case 2: gngf::encode_hash_fwd_kernel<2><<<blocks, 256, 0, st>>>(x2, P, lat, tables, T, enc, idx_out); break;
    case 4: gngf::encode_hash_fwd_kernel<4><<<blocks, 256, 0, st>>>(x2, P, lat, tables, T, enc, idx_out); break;
    case 8: gngf::encode_hash_fwd_kernel<8><<<blocks, 256, 0, st>>>(x2, P, lat, tables, T, enc, idx_out); break;
    default: return GNGF_ERR_UNSUPPORTED;
  }
  gngf::note_launch();
  return gngf::check_launch();
}

int gngf_lattice_colsum(gngf_lattice lat, const int32_t* cnt, const float* uvals, int64_t N, float* out, void* stream) {
  if (!gngf::valid_lat(lat) || N <= 0) return GNGF_ERR_INVALID_ARGUMENT;
  const int64_t box = gngf::max_level_box(lat);
  if (N <= 8) {
    dim3 grid(static_cast<unsigned>(std::min<int64_t>(gngf::ceil_div(box, 256), 8 * gngf::sm_count())), lat.num_levels);
    gngf::lattice_colsum_narrow_kernel<8><<<grid, 256, 0, gngf::as_stream(stream)>>>(lat, cnt, uvals, static_cast<int>(N),
                                                                                    out);
    gngf::note_launch();
    return gngf::check_launch();
  }
  const int64_t nodes_per_block = std::max<int64_t>(32, gngf::ceil_div(box, 64));
  dim3 grid(static_cast<unsigned>(gngf::ceil_div(N, 128)), static_cast<unsigned>(gngf::ceil_div(box, nodes_per_block)),
            lat.num_levels);
  gngf::lattice_colsum_kernel<<<grid, 128, 0, gngf::as_stream(stream)>>>(lat, cnt, uvals, N, nodes_per_block, out);
  gngf::note_launch();
  return gngf::check_launch();
}

int gngf_lattice_gather_rows(const float* x, int64_t P, gngf_lattice lat, const float* uvals, int64_t N, float* out,
                             void* stream) {
  if (!gngf::valid_lat(lat) || P < 0 || N <= 0) return GNGF_ERR_INVALID_ARGUMENT;
  if (P == 0) return GNGF_OK;
  const int64_t n = P * lat.num_levels * 4 * N;
  gngf::gather_rows_kernel<float, float><<<static_cast<unsigned>(gngf::ceil_div(n, 256)), 256, 0,
                                           gngf::as_stream(stream)>>>(reinterpret_cast<const float2*>(x), P, lat, uvals,
                                                                      N, out);
  gngf::note_launch();
  return gngf::check_launch();
}

int gngf_lattice_gather_rows_i64(const float* x, int64_t P, gngf_lattice lat, const int32_t* uvals, int64_t N,
                                 int64_t* out, void* stream) {
  if (!gngf::valid_lat(lat) || P < 0 || N <= 0) return GNGF_ERR_INVALID_ARGUMENT;
  if (P == 0) return GNGF_OK;
  const int64_t n = P * lat.num_levels * 4 * N;
  gngf::gather_rows_kernel<int32_t, int64_t><<<static_cast<unsigned>(gngf::ceil_div(n, 256)), 256, 0,
                                               gngf::as_stream(stream)>>>(reinterpret_cast<const float2*>(x), P, lat,
                                                                          uvals, N, out);
  gngf::note_launch();
  return gngf::check_launch();
}

int gngf_lattice_scatter_rows(const float* x, int64_t P, gngf_lattice lat, const float* dout, int64_t N, float* dvals,
                              void* stream) {
  if (!gngf::valid_lat(lat) || P < 0 || N <= 0) return GNGF_ERR_INVALID_ARGUMENT;
  if (P == 0) return GNGF_OK;
  const int64_t n = P * lat.num_levels * 4 * N;
  gngf::scatter_rows_kernel<<<static_cast<unsigned>(gngf::ceil_div(n, 256)), 256, 0, gngf::as_stream(stream)>>>(
      reinterpret_cast<const float2*>(x), P, lat, dout, N, dvals);
  gngf::note_launch();
  return gngf::check_launch();
}

}  // extern "C"
