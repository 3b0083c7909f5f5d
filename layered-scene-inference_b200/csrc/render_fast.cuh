// Fast forward-splat kernel (the training/inference hot path): one source pixel position x all L layers per
// thread, pose-specialised arithmetic, vector reductions into an L2-resident accumulator.
//
// Why this shape (measured on B200, profiles/r1_*): the one-thread-per-pixel kernel is instruction-issue bound
// (244 SASS instructions per source pixel, 75 % issue-slot utilisation, 26 % of HBM peak).  At B200's
// bytes-per-instruction ratio the HBM roofline allows ~91 issue slots per source pixel, so the kernel
//   * hoists everything that does not depend on the layer (pixel grid, M*(x,y,1)) out of the layer loop,
//   * removes the three IEEE divisions when the pose has n == 1 (pure translation in the image plane: rectified
//     stereo, the KITTI configs) -- dividing by exactly 1 is the identity, so results are unchanged,
//   * hoists the vertical corner weights when y does not depend on x or d (same case),
//   * uses F2I.FLOOR/I2F instead of floorf + casts and ex2.approx instead of expf.
// (A 4-pixels-per-thread variant with 128-bit loads was measured SLOWER, 0.81 vs 0.52 ms/step: its reductions hit
// every fourth accumulator cell per lane, doubling the L2 sector operations; lanes must map to consecutive pixels.)
// Semantics are those of render.cu's reference kernel (same corner/threshold/validity rules).
#pragma once
#include "common.cuh"

namespace lsi {

struct FastParams {
  const float* tex; const float* disp; const float* mask; const float* mats;
  float4* acc4;             // [nl_acc][bc][Nt]
  int L, B, H, W, b0, bc;
  int acc_per_layer;
  int w_t, h_t;
  float ds, inv_max_disp;
  float k2, k2h;            // zb = ex2(clip(r)*k2 - k2h), k2 = scale*log2(e), k2h = 0.5*k2
  int ablate;               // measurement only: 1 = no reductions issued, 2 = reductions to the pixel's own cell
  const int* skip;          // per batch element: 1 = already rendered by the row-owner / row-gather kernel (nullable)
  const int* n_flagged;     // number of such images in the whole batch (with skip)
};

struct AxisW {   // one axis of the bilinear footprint: integer base, the two weights (validity folded in)
  int i0; float w0, w1;
};

// floor / weights / validity along one axis (sampling.py:193-211); extent = target size along the axis
__device__ __forceinline__ AxisW axis_weights(float x, int extent) {
  AxisW a;
  const float xc = fminf(fmaxf(x, -2.f), (float)extent + 1.f);   // keeps the int conversion defined; NaN -> -2 -> invalid
  a.i0 = __float2int_rd(xc);
  const float x0 = (float)a.i0;
  const float w1 = x - x0, w0 = (x0 + 1.f) - x;                  // wt_x1 = x - x0 ; wt_x0 = x1 - x
  a.w0 = ((unsigned)a.i0 < (unsigned)extent) ? w0 : 0.f;
  a.w1 = ((unsigned)(a.i0 + 1) < (unsigned)extent) ? w1 : 0.f;
  return a;
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// z-buffer weight (helpers.py:180-193) with exp(t) = 2^(t*log2 e)
__device__ __forceinline__ float zb_weight(float dt, const FastParams& p) {
  const float r = dt * p.inv_max_disp;
  const float z = ex2_approx(fmaf(__saturatef(r), p.k2, -p.k2h));
  return r > 0.f ? z : 0.f;
}

__device__ __forceinline__ void red4(float4* acc, int q, const float4& v) {
  atomicAdd(acc + q, v);   // REDG.E.ADD.F32x4
}

// One bilinear row (two horizontally adjacent cells q, q+1) of a pixel's footprint, with the warp-level
// pre-reduction: lane i's right cell is lane i+1's left cell whenever consecutive source pixels land on
// consecutive target cells (the common case), so lane i+1 folds its left neighbour's right-cell value into its own
// left-cell value and lane i skips that reduction: ~1 L2 reduction per pixel-row instead of 2.
// All 32 lanes must call this (shuffles); lanes with nothing to add pass zero weights.
__device__ __forceinline__ void splat_row(float4* acc, int q, float ol, float orr, float vr, float vg, float vb,
                                          float w, int q_prev, int q_next, float orr_prev, float vr_p, float vg_p,
                                          float vb_p, float w_p, int lane, int ablate) {
  float4 left = make_float4(vr * ol, vg * ol, vb * ol, w * ol);
  if (lane > 0 && q_prev + 1 == q) {
    left.x = fmaf(vr_p, orr_prev, left.x); left.y = fmaf(vg_p, orr_prev, left.y);
    left.z = fmaf(vb_p, orr_prev, left.z); left.w = fmaf(w_p, orr_prev, left.w);
  }
  if (left.w != 0.f && ablate != 1) red4(acc, q, left);
  const bool taken = (lane < 31) && (q_next == q + 1);
  if (!taken && orr > 0.f && w != 0.f && ablate != 1) red4(acc, q + 1, make_float4(vr * orr, vg * orr, vb * orr, w * orr));
}

__device__ __forceinline__ float thresh(float o) { return o > kWtThresh ? o : 0.f; }   // sampling.py:219-222

// kMode 0: general pose; 1: n == 1 (no division); 2: n == 1 and y independent of (x, d) (vertical weights hoisted)
// Warp-convergent: every lane of the warp calls this for every layer; `w` is zero for lanes past the row end.
template <int kMode>
__device__ __forceinline__ void splat_pixel(const FastParams& p, const Mat34& M, float bu, float bv, float bn, float d,
                                            float tr, float tg, float tb, float mk, const AxisW& ay_row, float4* acc,
                                            int lane) {
  float x, y, dt;
  const float up = fmaf(M.m[3], d, bu);
  if (kMode == 0) {
    const float vp = fmaf(M.m[7], d, bv);
    const float nh = safe_den(fmaf(M.m[11], d, bn));
    x = (up / nh) * p.ds - 0.5f; y = (vp / nh) * p.ds - 0.5f; dt = d / nh;
  } else {
    x = fmaf(up, p.ds, -0.5f); dt = d;
    y = (kMode == 1) ? fmaf(fmaf(M.m[7], d, bv), p.ds, -0.5f) : 0.f;
  }
  const float w = zb_weight(dt, p) * mk;
  const AxisW ax = axis_weights(x, p.w_t);
  const AxisW ay = (kMode == 2) ? ay_row : axis_weights(y, p.h_t);
  const float vr = tr * w, vg = tg * w, vb = tb * w;
  const int q = ay.i0 * p.w_t + ax.i0;
  // neighbour exchange (left neighbour's cell index, right weight and weighted values; right neighbour's cell index)
  const unsigned full = 0xffffffffu;
  const int q_prev = __shfl_up_sync(full, q, 1), q_next = __shfl_down_sync(full, q, 1);
  const float w1x_p = __shfl_up_sync(full, ax.w1, 1);
  const float vr_p = __shfl_up_sync(full, vr, 1), vg_p = __shfl_up_sync(full, vg, 1);
  const float vb_p = __shfl_up_sync(full, vb, 1), w_p = __shfl_up_sync(full, w, 1);
  float w0y_p = ay.w0, w1y_p = ay.w1;
  if (kMode != 2) { w0y_p = __shfl_up_sync(full, ay.w0, 1); w1y_p = __shfl_up_sync(full, ay.w1, 1); }
  if (kMode != 2 || ay.w0 > 0.f)
    splat_row(acc, q, thresh(ax.w0 * ay.w0), thresh(ax.w1 * ay.w0), vr, vg, vb, w, q_prev, q_next,
              thresh(w1x_p * w0y_p), vr_p, vg_p, vb_p, w_p, lane, p.ablate);
  if (kMode != 2 || ay.w1 > 0.f)
    splat_row(acc, q + p.w_t, thresh(ax.w0 * ay.w1), thresh(ax.w1 * ay.w1), vr, vg, vb, w, q_prev + p.w_t,
              q_next + p.w_t, thresh(w1x_p * w1y_p), vr_p, vg_p, vb_p, w_p, lane, p.ablate);
}

template <int kMode, bool kHasMask, bool kPacked>
__device__ __forceinline__ void splat_column_layers(const FastParams& p, const Mat34& M, int b, int bl, int i, int j,
                                                    bool active) {
  const int n_src = p.H * p.W;
  const int n_trg = p.h_t * p.w_t;
  const int lane = threadIdx.x & 31;
  const float xs = (float)j + 0.5f, ys = (float)i + 0.5f;            // helpers.py:88-113
  const float bu = fmaf(M.m[1], ys, M.m[0] * xs) + M.m[2];
  const float bv = fmaf(M.m[5], ys, M.m[4] * xs) + M.m[6];
  const float bn = fmaf(M.m[9], ys, M.m[8] * xs) + M.m[10];
  AxisW ay_row;
  ay_row.i0 = 0; ay_row.w0 = 0.f; ay_row.w1 = 0.f;
  if (kMode == 2) ay_row = axis_weights(fmaf(bv, p.ds, -0.5f), p.h_t);   // uniform over the row
  const size_t pix = (size_t)i * p.W + (active ? j : 0);
  const size_t lstride = (size_t)p.B * n_src;
  size_t img = (size_t)b * n_src + pix;
  float4* acc = p.acc4 + (size_t)bl * n_trg;
  const size_t acc_lstride = p.acc_per_layer ? (size_t)p.bc * n_trg : 0;
  // layers in groups of 4: issue every load of the group first (16-20 independent requests per thread in flight --
  // the kernel is latency/bytes-in-flight bound otherwise), then do the arithmetic and reductions
  for (int l0 = 0; l0 < p.L; l0 += 4) {
    const int n = min(4, p.L - l0);
    float tr[4], tg[4], tb[4], d[4], mk[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (u < n) {
        const size_t im = img + (size_t)u * lstride;
        if (kPacked) {   // [.,H,W,4] = (r,g,b,disp) per pixel: the head output layout (nets.py:204)
          const float4 v = __ldcs(reinterpret_cast<const float4*>(p.tex) + im);
          tr[u] = v.x; tg[u] = v.y; tb[u] = v.z; d[u] = v.w;
        } else {
          const float* t = p.tex + im * 3;
          tr[u] = __ldcs(t); tg[u] = __ldcs(t + 1); tb[u] = __ldcs(t + 2);
          d[u] = __ldcs(p.disp + im);
        }
        mk[u] = active ? 1.f : 0.f;
        if (kHasMask) mk[u] *= __ldcs(p.mask + im);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (u < n) {
        if (p.ablate == 2) {   // measurement only: loads + one perfectly coalesced reduction per pixel, no geometry
          if (active && i < p.h_t && j < p.w_t) red4(acc + (size_t)u * acc_lstride, i * p.w_t + j, make_float4(tr[u], tg[u], tb[u], d[u]));
        } else {
          splat_pixel<kMode>(p, M, bu, bv, bn, d[u], tr[u], tg[u], tb[u], mk[u], ay_row, acc + (size_t)u * acc_lstride, lane);
        }
      }
    }
    img += 4 * lstride; acc += 4 * acc_lstride;
  }
}

// One thread per source pixel position (all L layers); consecutive lanes = consecutive pixels of one row, so that the
// vector reductions of a warp land on consecutive 16-byte accumulator cells and neighbours can pre-reduce by shuffle.
// blockDim.x is a multiple of 32 and every warp stays converged (lanes past the row end carry zero weight).
template <bool kHasMask, bool kPacked>
__global__ void __launch_bounds__(64, 16) splat_fwd_fast_kernel(const FastParams p) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = j < p.W;
  const int i = blockIdx.y;
  const int bl = blockIdx.z, b = p.b0 + bl;
  if (p.skip && p.skip[b]) return;   // uniform per block
  const Mat34 M = load_mat(p.mats, b);
  // pose class (uniform per block): n == 1 for every (x, y, d)?  y independent of x and d?
  const bool affine = (M.m[8] == 0.f) && (M.m[9] == 0.f) && (M.m[10] == 1.f) && (M.m[11] == 0.f);
  const bool yconst = affine && (M.m[4] == 0.f) && (M.m[7] == 0.f);
  if (yconst) splat_column_layers<2, kHasMask, kPacked>(p, M, b, bl, i, j, active);
  else if (affine) splat_column_layers<1, kHasMask, kPacked>(p, M, b, bl, i, j, active);
  else splat_column_layers<0, kHasMask, kPacked>(p, M, b, bl, i, j, active);
}

// normalise/compose, 4 target pixels per thread (n_trg % 4 == 0): 4 x LDG.128 in, 3 x STG.128 + 1 x STG.128 out
struct NormFastParams {
  const float4* acc4; float* img; float* wts;
  int B, b0, bc, n_trg;
  float nb;   // bg_wt * (number of canvases summed into one accumulator)
  const int* skip;
};

__global__ void __launch_bounds__(256) normalize_fast_kernel(const NormFastParams p) {
  const int q4 = blockIdx.x * blockDim.x + threadIdx.x;
  if (q4 * 4 >= p.n_trg) return;
  const int bl = blockIdx.y, b = p.b0 + bl, lo = blockIdx.z;
  if (p.skip && p.skip[b]) return;
  const float4* a = p.acc4 + ((size_t)lo * p.bc + bl) * p.n_trg + (size_t)q4 * 4;
  const size_t o = ((size_t)lo * p.B + b) * p.n_trg + (size_t)q4 * 4;
  float v[12], w[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float4 s = a[k];
    // one correctly rounded reciprocal + three multiplies instead of three IEEE divisions (<= 1 ulp apart): the
    // division sequences made this pass instruction-bound at a third of the HBM write rate
    const float W = s.w + p.nb, Wi = __frcp_rn(safe_den(W));
    v[3 * k] = (s.x + p.nb) * Wi; v[3 * k + 1] = (s.y + p.nb) * Wi; v[3 * k + 2] = (s.z + p.nb) * Wi;
    w[k] = W;
  }
  float4* ip = reinterpret_cast<float4*>(p.img + o * 3);
  __stcs(ip, make_float4(v[0], v[1], v[2], v[3]));
  __stcs(ip + 1, make_float4(v[4], v[5], v[6], v[7]));
  __stcs(ip + 2, make_float4(v[8], v[9], v[10], v[11]));
  __stcs(reinterpret_cast<float4*>(p.wts + o), make_float4(w[0], w[1], w[2], w[3]));
}

}  // namespace lsi
