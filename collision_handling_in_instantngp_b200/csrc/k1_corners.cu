// K1: grid corners of every (point, level) -- the materialised form of _scale_to_grid (models.py:486-502)
// -- and the Instant-NGP spatial hash used as the baseline index source (models.py:504-528).
// The fused kernels (k4/k5) recompute the same cell_of() in registers; these two exist for API parity
// (`_scale_to_grid`, `_fast_hash`) and for the bit-exactness tests.
#include "common.cuh"

namespace gngf {

// thread per (point, level); scaled (P,2,L,1), grid (P,2,L,4)
__global__ void __launch_bounds__(256) corners_kernel(const float2* __restrict__ x, int64_t P,
                                                      const __grid_constant__ gngf_lattice lat,
                                                      float* __restrict__ scaled, float4* __restrict__ grid) {
  const int L = lat.num_levels;
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= P * L) return;
  const int64_t p = i / L;
  const int l = static_cast<int>(i - p * L);
  const float2 xy = x[p];
  const Cell c = cell_of(xy.x, xy.y, lat.n[l]);
  const float fx = static_cast<float>(c.cx), fy = static_cast<float>(c.cy);
  scaled[(p * 2 + 0) * L + l] = c.sx;
  scaled[(p * 2 + 1) * L + l] = c.sy;
  grid[(p * 2 + 0) * L + l] = make_float4(fx, fx + 1.0f, fx, fx + 1.0f);
  grid[(p * 2 + 1) * L + l] = make_float4(fy, fy, fy + 1.0f, fy + 1.0f);
}

// _fast_hash: per dimension (int32)(g_i * prime_i) wraps, xor is done on sign-extended int64, then the
// non-negative remainder.  primes = (1, 2654435761).
__global__ void __launch_bounds__(256) fast_hash_kernel(const float2* __restrict__ x, int64_t P,
                                                        const __grid_constant__ gngf_lattice lat, int64_t T,
                                                        int64_t* __restrict__ idx) {
  const int L = lat.num_levels;
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= P * L) return;
  const int64_t p = i / L;
  const int l = static_cast<int>(i - p * L);
  const float2 xy = x[p];
  const Cell c = cell_of(xy.x, xy.y, lat.n[l]);
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    const uint32_t gx = static_cast<uint32_t>(c.cx + (v & 1));
    const uint32_t gy = static_cast<uint32_t>(c.cy + (v >> 1));
    const int32_t h32 = static_cast<int32_t>(gx ^ (gy * 2654435761u));
    int64_t h = static_cast<int64_t>(h32) % T;
    if (h < 0) h += T;
    idx[i * 4 + v] = h;
  }
}

}  // namespace gngf

extern "C" {

int gngf_corners_fwd(const float* x, int64_t P, gngf_lattice lat, float* scaled, float* grid, void* stream) {
  if (P < 0 || lat.num_levels <= 0 || lat.num_levels > GNGF_MAX_LEVELS) return GNGF_ERR_INVALID_ARGUMENT;
  if (P == 0) return GNGF_OK;
  const int64_t n = P * lat.num_levels;
  gngf::corners_kernel<<<static_cast<unsigned>(gngf::ceil_div(n, 256)), 256, 0, gngf::as_stream(stream)>>>(
      reinterpret_cast<const float2*>(x), P, lat, scaled, reinterpret_cast<float4*>(grid));
  gngf::note_launch();
  return gngf::check_launch();
}

int gngf_fast_hash_fwd(const float* x, int64_t P, gngf_lattice lat, int64_t table_size, int64_t* idx, void* stream) {
  if (P < 0 || table_size <= 0 || lat.num_levels <= 0 || lat.num_levels > GNGF_MAX_LEVELS)
    return GNGF_ERR_INVALID_ARGUMENT;
  if (P == 0) return GNGF_OK;
  const int64_t n = P * lat.num_levels;
  gngf::fast_hash_kernel<<<static_cast<unsigned>(gngf::ceil_div(n, 256)), 256, 0, gngf::as_stream(stream)>>>(
      reinterpret_cast<const float2*>(x), P, lat, table_size, idx);
  gngf::note_launch();
  return gngf::check_launch();
}

}  // extern "C"
