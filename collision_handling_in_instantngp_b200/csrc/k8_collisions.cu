// f-1: the distinct-slot count behind calc_hash_collisions (models.py:568-619).  The reference calls
// torch.unique(dim=0) once per (top-k column, level) -- 48 sort-based calls per epoch, 87 % of its CPU epoch time.
// Here: one pass sets one bit per (column, level, slot) in a bitmap (test-before-atomicOr, so the contended words
// are written ~once per bit), a second tiny pass popcounts.  Values that are not integers in [0, range) (train_step
// hands over a float32 tensor allocated with torch.empty, functions.py:179) are reported through `outliers` so the
// caller can take the exact path for them.
#include <algorithm>

#include "common.cuh"

namespace gngf {

template <typename T>
__global__ void __launch_bounds__(256)
    mark_slots_kernel(const T* __restrict__ idx, int64_t n, int L, int V, int C, int64_t range, int64_t words,
                      uint32_t* __restrict__ bitmap, int32_t* __restrict__ outliers) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  bool bad = false;
  for (int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; e < n; e += stride) {
    const int c = static_cast<int>(e % C);
    const int l = static_cast<int>((e / (static_cast<int64_t>(C) * V)) % L);
    const T raw = idx[e];
    const int64_t iv = static_cast<int64_t>(raw);
    if (static_cast<T>(iv) != raw || iv < 0 || iv >= range) {
      bad = true;
      continue;
    }
    uint32_t* w = bitmap + (static_cast<int64_t>(c) * L + l) * words + (iv >> 5);
    const uint32_t bit = 1u << (iv & 31);
    if (!(__ldcg(w) & bit)) atomicOr(w, bit);
  }
  if (bad) *outliers = 1;
}

// one block per (column, level)
__global__ void __launch_bounds__(256)
    count_slots_kernel(const uint32_t* __restrict__ bitmap, int64_t words, int32_t* __restrict__ uniq) {
  __shared__ int red[8];
  const uint32_t* w = bitmap + static_cast<int64_t>(blockIdx.x) * words;
  int s = 0;
  for (int64_t i = threadIdx.x; i < words; i += 256) s += __popc(w[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int i = 0; i < 8; ++i) t += red[i];
    uniq[blockIdx.x] = t;
  }
}

template <typename T>
static int count_distinct(const T* idx, int64_t P, int L, int V, int C, int64_t range, uint32_t* bitmap, int32_t* uniq,
                          int32_t* outliers, cudaStream_t st) {
  if (P < 0 || L <= 0 || V <= 0 || C <= 0 || range <= 0) return GNGF_ERR_INVALID_ARGUMENT;
  const int64_t words = ceil_div(range, 32);
  if (cudaMemsetAsync(bitmap, 0, sizeof(uint32_t) * words * C * L, st) != cudaSuccess) return check_launch();
  if (cudaMemsetAsync(outliers, 0, sizeof(int32_t), st) != cudaSuccess) return check_launch();
  const int64_t n = P * L * V * C;
  if (n > 0) {
    const int blocks = static_cast<int>(std::min<int64_t>(ceil_div(n, 256), 8 * sm_count()));
    mark_slots_kernel<T><<<blocks, 256, 0, st>>>(idx, n, L, V, C, range, words, bitmap, outliers);
    note_launch();
    int rc = check_launch();
    if (rc) return rc;
  }
  count_slots_kernel<<<C * L, 256, 0, st>>>(bitmap, words, uniq);
  note_launch();
  return check_launch();
}

// ---- f-4: _calc_counts_per_level (models.py:530-566) -----------------------------------------------------------------
// The reference de-duplicates, per level, the rows "p (v xy)" -- the 8 corner coordinates of a point's grid cell -- with
// np.unique(axis=0, return_index=True) on the host, and then indexes the FLATTENED "(p v)" slot vector with the
// returned first-occurrence POINT indices (models.py:556-558): for every distinct cell the slot counted is that of
// corner (j % 4) of point (j / 4), j = first point (batch order) in the cell.  Two passes here: (1) per (point, level)
// an atomicMin of the point index into the cell's entry (cells are indexed like the level nodes of their floor
// corner), (2) per cell with an entry, one histogram increment.  Exact, order-free.
__global__ void __launch_bounds__(256)
    cell_first_point_kernel(const float* __restrict__ grid, int64_t P, const __grid_constant__ gngf_lattice lat,
                            int32_t* __restrict__ first, int32_t* __restrict__ outliers) {
  const int L = lat.num_levels;
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= P * L) return;
  const int64_t p = i / L;
  const int l = static_cast<int>(i - p * L);
  // grid (P,2,L,4): the floor corner is v = 0
  const float fx = grid[((p * 2 + 0) * L + l) * 4], fy = grid[((p * 2 + 1) * L + l) * 4];
  const int a = static_cast<int>(fx) - lat.lox[l], b = static_cast<int>(fy) - lat.loy[l];
  if (!(fx == floorf(fx) && fy == floorf(fy)) || a < 0 || a >= lat.lwx[l] || b < 0 || b >= lat.lwy[l]) {
    *outliers = 1;
    return;
  }
  atomicMin(first + lat.loff[l] + static_cast<int64_t>(a) * lat.lwy[l] + b, static_cast<int32_t>(p));
}

__global__ void __launch_bounds__(256)
    cell_slot_histogram_kernel(const int32_t* __restrict__ first, const __grid_constant__ gngf_lattice lat,
                               const int64_t* __restrict__ hashed, int64_t stride, int64_t P, int64_t T,
                               int32_t* __restrict__ hist, int32_t* __restrict__ outliers) {
  const int l = blockIdx.y;
  const int L = lat.num_levels;
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= static_cast<int64_t>(lat.lwx[l]) * lat.lwy[l]) return;
  const int32_t j = first[lat.loff[l] + i];
  if (j == 0x7f7f7f7f) return;   // (the memset pattern: no point in this cell)
  // element j of "p l v -> l (p v)": point j / 4, corner j % 4
  const int64_t slot = hashed[(((j >> 2) * static_cast<int64_t>(L) + l) * 4 + (j & 3)) * stride];
  if (slot < 0 || slot >= T) {
    *outliers = 1;
    return;
  }
  atomicAdd(hist + static_cast<int64_t>(l) * T + slot, 1);
}

}  // namespace gngf

extern "C" {

int gngf_counts_per_level(const float* grid, int64_t P, gngf_lattice lat, const int64_t* hashed, int64_t hashed_stride,
                          int64_t T, int32_t* first, int32_t* hist, int32_t* outliers, void* stream) {
  if (P < 0 || P >= 0x7f7f7f7fll || T <= 0 || hashed_stride <= 0 || lat.num_levels <= 0 ||
      lat.num_levels > GNGF_MAX_LEVELS || !first || !hist || !outliers || (P > 0 && (!grid || !hashed)))
    return GNGF_ERR_INVALID_ARGUMENT;
  cudaStream_t st = gngf::as_stream(stream);
  const int L = lat.num_levels;
  const int64_t S = lat.loff[L];
  if (cudaMemsetAsync(first, 0x7f, sizeof(int32_t) * S, st) != cudaSuccess) return gngf::check_launch();   // 0x7f7f7f7f
  if (cudaMemsetAsync(hist, 0, sizeof(int32_t) * L * T, st) != cudaSuccess) return gngf::check_launch();
  if (cudaMemsetAsync(outliers, 0, sizeof(int32_t), st) != cudaSuccess) return gngf::check_launch();
  if (P == 0) return GNGF_OK;
  gngf::cell_first_point_kernel<<<static_cast<unsigned>(gngf::ceil_div(P * L, 256)), 256, 0, st>>>(grid, P, lat, first,
                                                                                                 outliers);
  gngf::note_launch();
  int rc = gngf::check_launch();
  if (rc) return rc;
  int64_t box = 0;
  for (int l = 0; l < L; ++l) box = std::max<int64_t>(box, static_cast<int64_t>(lat.lwx[l]) * lat.lwy[l]);
  dim3 g(static_cast<unsigned>(gngf::ceil_div(box, 256)), L);
  gngf::cell_slot_histogram_kernel<<<g, 256, 0, st>>>(first, lat, hashed, hashed_stride, P, T, hist, outliers);
  gngf::note_launch();
  return gngf::check_launch();
}

int64_t gngf_count_distinct_workspace_words(int32_t L, int32_t C, int64_t range) {
  return gngf::ceil_div(range, 32) * C * L;
}

int gngf_count_distinct_f32(const float* indices, int64_t P, int32_t L, int32_t V, int32_t C, int64_t range,
                            uint32_t* bitmap, int32_t* uniq, int32_t* outliers, void* stream) {
  return gngf::count_distinct<float>(indices, P, L, V, C, range, bitmap, uniq, outliers, gngf::as_stream(stream));
}

int gngf_count_distinct_i64(const int64_t* indices, int64_t P, int32_t L, int32_t V, int32_t C, int64_t range,
                            uint32_t* bitmap, int32_t* uniq, int32_t* outliers, void* stream) {
  return gngf::count_distinct<int64_t>(indices, P, L, V, C, range, bitmap, uniq, outliers, gngf::as_stream(stream));
}

}  // extern "C"
