// Shared helpers of libgngf_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "gngf.h"

namespace gngf {

// bookkeeping shared by all translation units (defined in abi.cu)
void note_launch(int n = 1);
int check_launch();  // cudaGetLastError -> gngf_status, remembers the message for gngf_strerror

int sm_count();       // multiprocessors of the current device (cached)
int debug_no_skip();  // GNGF_DEBUG_NO_SKIP=1 (read on every call; tests only): the streaming kernels skip nothing

// number of leading level nodes (the coarsest levels are stored first) that fit a shared-memory budget
inline int private_node_count(const gngf_lattice& lat, int64_t max_nodes) {
  int64_t n = 0;
  for (int l = 0; l < lat.num_levels; ++l) {
    if (lat.loff[l + 1] > max_nodes) break;
    n = lat.loff[l + 1];
  }
  return static_cast<int>(n);
}

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ----------------------------------------------------------------------------------------------------
// One grid cell of one level: floor corner and the four bilinear weights.
// Reference arithmetic (models.py:492-500, 626-637), every rounding kept separate (no FMA contraction):
//   s = x * n_l ; a = floor(s) ; d = a + 1
//   w = [(xd-x)(yd-y), (x-xa)(yd-y), (xd-x)(y-ya), (x-xa)(y-ya)]   corner order (0,0),(1,0),(0,1),(1,1)
// ----------------------------------------------------------------------------------------------------
struct Cell {
  int cx, cy;     // floor corner (integer lattice coordinates)
  float sx, sy;   // scaled coordinates
  float w[4];
};

__device__ __forceinline__ Cell cell_of(float x, float y, int n) {
  Cell c;
  const float fn = static_cast<float>(n);
  c.sx = __fmul_rn(x, fn);
  c.sy = __fmul_rn(y, fn);
  const float fx = floorf(c.sx), fy = floorf(c.sy);
  const float xd = __fadd_rn(fx, 1.0f), yd = __fadd_rn(fy, 1.0f);
  const float ax = __fsub_rn(xd, c.sx), bx = __fsub_rn(c.sx, fx);
  const float ay = __fsub_rn(yd, c.sy), by = __fsub_rn(c.sy, fy);
  c.w[0] = __fmul_rn(ax, ay);
  c.w[1] = __fmul_rn(bx, ay);
  c.w[2] = __fmul_rn(ax, by);
  c.w[3] = __fmul_rn(bx, by);
  c.cx = static_cast<int>(fx);
  c.cy = static_cast<int>(fy);
  return c;
}

// index of corner (cx,cy) inside level l's node box; clamps and reports when outside
__device__ __forceinline__ int64_t level_node(const gngf_lattice& lat, int l, int cx, int cy, bool& outside) {
  int i = cx - lat.lox[l], j = cy - lat.loy[l];
  const int wx = lat.lwx[l], wy = lat.lwy[l];
  if (i < 0 || i >= wx || j < 0 || j >= wy) {
    outside = true;
    i = min(max(i, 0), wx - 1);
    j = min(max(j, 0), wy - 1);
  }
  return lat.loff[l] + static_cast<int64_t>(i) * wy + j;
}

// index of corner (cx,cy) in the global node box
__device__ __forceinline__ int64_t global_node(const gngf_lattice& lat, int cx, int cy) {
  int i = min(max(cx - lat.ox, 0), lat.wx - 1);
  int j = min(max(cy - lat.oy, 0), lat.wy - 1);
  return static_cast<int64_t>(i) * lat.wy + j;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// vector reduction into global memory (sm_90+): one L2 atomic for two adjacent floats
__device__ __forceinline__ void red_add_v2(float* addr, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}

// four adjacent floats (16-byte aligned) in one L2 reduction
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// top-k mix weights (models.py:212-217): the normaliser of a node's K probabilities, and one weight
__device__ __forceinline__ float mix_weight_norm(const float* tv, int K, int mode, float& mx) {
  // returns the normaliser; mx = max (softmax mode)
  mx = 0.0f;
  if (mode == GNGF_MIX_SOFTMAX) {
    mx = tv[0];
    for (int k = 1; k < K; ++k) mx = fmaxf(mx, tv[k]);
    float s = 0.0f;
    for (int k = 0; k < K; ++k) s += expf(tv[k] - mx);
    return s;
  }
  if (mode == GNGF_MIX_WEIGHTED_AVG) {
    float s = 0.0f;
    for (int k = 0; k < K; ++k) s += tv[k];
    return s;
  }
  return 1.0f;
}
__device__ __forceinline__ float mix_weight(float tv, int mode, float mx, float norm) {
  if (mode == GNGF_MIX_SOFTMAX) return expf(tv - mx) / norm;
  if (mode == GNGF_MIX_WEIGHTED_AVG) return tv / norm;
  return tv;
}

}  // namespace gngf
