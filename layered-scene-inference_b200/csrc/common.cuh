// Shared device helpers for the B200 LDI renderer kernels (sm_100a).
//
// The arithmetic below restates, per source pixel and in registers, what the reference expresses as a chain of
// TF ops (reference paths relative to the upstream tree):
//   helpers.transform_pts  lsi/nnutils/helpers.py:116-137      helpers.divide_safe    helpers.py:82-85
//   helpers.zbuffer_weights helpers.py:180-193                 sampling.splat corners lsi/geometry/sampling.py:183-222
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lsi {

constexpr float kEpsDiv = 1e-8f;     // helpers.py:83
constexpr float kWtThresh = 1e-3f;   // sampling.py:219-222

// divide_safe's denominator: den + 1e-8*[den == 0]  (helpers.py:82-85)
__device__ __forceinline__ float safe_den(float d) { return d == 0.f ? kEpsDiv : d; }

// Streaming (evict-first) loads/stores for data touched exactly once.
__device__ __forceinline__ float ld_stream(const float* p) { return __ldcs(p); }

struct Mat34 {  // rows 0..2 of the 4x4 src->trg matrix; row 3 is (0,0,0,1) by construction (projection.py:27-68)
  float m[12];
};

__device__ __forceinline__ Mat34 load_mat(const float* __restrict__ mats, int b) {
  Mat34 M;
  const float4* p = reinterpret_cast<const float4*>(mats + (size_t)b * 16);
  float4 r0 = __ldg(p), r1 = __ldg(p + 1), r2 = __ldg(p + 2);
  M.m[0] = r0.x; M.m[1] = r0.y; M.m[2] = r0.z; M.m[3] = r0.w;
  M.m[4] = r1.x; M.m[5] = r1.y; M.m[6] = r1.z; M.m[7] = r1.w;
  M.m[8] = r2.x; M.m[9] = r2.y; M.m[10] = r2.z; M.m[11] = r2.w;
  return M;
}

// Per-source-pixel geometry: projection, z-buffer weight and the four bilinear corners.
struct PixGeom {
  float up, vp, np_, dp;      // (u', v', n, d') = M (x, y, one, d)
  float nh;                   // safe normaliser
  float x, y;                 // target coords after the -0.5 centre shift (sampling.py:183)
  float dt;                   // target-frame disparity (+ focal)
  float r;                    // dt / max_disp
  float zb;                   // exp((clip(r,0,1)-0.5)*scale) * [r > 0]
  float wx0, wx1, wy0, wy1;   // 1-D weights with validity folded in (sampling.py:208-211)
  float vx0, vx1, vy0, vy1;   // validity as 0/1 (needed by the backward)
  int ix0, ix1, iy0, iy1;     // clipped integer corner coordinates (meaningful where valid)
  float w[4];                 // thresholded corner weights tl, tr, bl, br (sampling.py:213-222)
  bool keep[4];
};

struct GeomParams {
  int w_t, h_t;
  float ds, inv_max_disp, scale;
};

__device__ __forceinline__ void project(const Mat34& M, float xs, float ys, float one, float d_in, float focal,
                                        const GeomParams& gp, PixGeom& g) {
  g.up = fmaf(M.m[3], d_in, fmaf(M.m[2], one, fmaf(M.m[1], ys, M.m[0] * xs)));
  g.vp = fmaf(M.m[7], d_in, fmaf(M.m[6], one, fmaf(M.m[5], ys, M.m[4] * xs)));
  g.np_ = fmaf(M.m[11], d_in, fmaf(M.m[10], one, fmaf(M.m[9], ys, M.m[8] * xs)));
  g.dp = d_in;  // row 3 of M is (0,0,0,1)
  g.nh = safe_den(g.np_);
  // ldi.py:138-140: uv = divide_safe(uv, n) * ds ; d_t = divide_safe(d', n); then sampling.py:183: -= 0.5
  g.x = (g.up / g.nh) * gp.ds - 0.5f;
  g.y = (g.vp / g.nh) * gp.ds - 0.5f;
  g.dt = g.dp / g.nh + focal;
  g.r = g.dt * gp.inv_max_disp;
  float c = fminf(fmaxf(g.r, 0.f), 1.f);
  g.zb = g.r > 0.f ? expf((c - 0.5f) * gp.scale) : 0.f;
}

__device__ __forceinline__ void corners(float x, float y, int w_t, int h_t, PixGeom& g) {
  float x0 = floorf(x), y0 = floorf(y);
  float x1 = x0 + 1.f, y1 = y0 + 1.f;
  float xmax = (float)(w_t - 1), ymax = (float)(h_t - 1);
  // validity == "clip_by_value leaves it unchanged" (sampling.py:202-211); NaN/inf compare false -> dropped
  bool bx0 = (x0 >= 0.f) && (x0 <= xmax), bx1 = (x1 >= 0.f) && (x1 <= xmax);
  bool by0 = (y0 >= 0.f) && (y0 <= ymax), by1 = (y1 >= 0.f) && (y1 <= ymax);
  g.vx0 = bx0 ? 1.f : 0.f; g.vx1 = bx1 ? 1.f : 0.f; g.vy0 = by0 ? 1.f : 0.f; g.vy1 = by1 ? 1.f : 0.f;
  g.wx0 = bx0 ? (x1 - x) : 0.f;
  g.wx1 = bx1 ? (x - x0) : 0.f;
  g.wy0 = by0 ? (y1 - y) : 0.f;
  g.wy1 = by1 ? (y - y0) : 0.f;
  g.ix0 = bx0 ? (int)x0 : 0; g.ix1 = bx1 ? (int)x1 : 0;
  g.iy0 = by0 ? (int)y0 : 0; g.iy1 = by1 ? (int)y1 : 0;
  float wtl = g.wx0 * g.wy0, wtr = g.wx1 * g.wy0, wbl = g.wx0 * g.wy1, wbr = g.wx1 * g.wy1;
  g.keep[0] = wtl > kWtThresh; g.keep[1] = wtr > kWtThresh; g.keep[2] = wbl > kWtThresh; g.keep[3] = wbr > kWtThresh;
  g.w[0] = g.keep[0] ? wtl : 0.f; g.w[1] = g.keep[1] ? wtr : 0.f;
  g.w[2] = g.keep[2] ? wbl : 0.f; g.w[3] = g.keep[3] ? wbr : 0.f;
}

__device__ __forceinline__ int corner_index(const PixGeom& g, int c, int w_t) {
  int ix = (c & 1) ? g.ix1 : g.ix0;
  int iy = (c & 2) ? g.iy1 : g.iy0;
  return iy * w_t + ix;
}

// Block-wide sum (blockDim.x multiple of 32, <= 1024).  Result valid in thread 0.
__device__ __forceinline__ float block_sum(float v) {
  __shared__ float red[32];
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) red[wid] = v;
  __syncthreads();
  int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
  if (wid == 0) for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  return v;
}

}  // namespace lsi
