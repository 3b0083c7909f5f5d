// Row-owner forward splat for the rectified pose class (n == 1, y' independent of x and d: stereo pairs, the KITTI
// configurations): the north star's shared-memory-tile design, made conflict-free by OWNERSHIP instead of atomics.
//
// In this pose class a source row maps onto one or two fixed target rows with row-constant vertical weights, so a
// warp can own a band of R consecutive TARGET rows outright: it visits every source row that touches the band (all
// layers), accumulates (sum w*omega*rgb, sum w*omega) for its rows in a warp-private shared-memory tile with plain
// LDS.128 / FADD / STS.128 (shared-memory float atomics are CAS loops on sm_100a, hence ownership), then normalises
// and writes each target pixel exactly once with no global reduction, no accumulator memset and no separate
// normalise pass.  Deterministic.  HBM traffic = the algorithmic bytes (sources that straddle two bands are re-read
// through L2).
//
// Within a warp step (32 consecutive source pixels of one layer) lanes pre-reduce with their left neighbour by
// shuffle, then the 32 read-modify-writes are issued together when the target cells are provably distinct (strictly
// increasing cell index: the common, smooth case) or in match_any-ranked rounds when they are not (fold-overs at
// occlusion boundaries, noisy disparities).
//
// Arithmetic is bit-identical in structure to render_fast.cuh (same projection/weight/threshold formulas); only the
// summation order differs.
#pragma once
#include "render_fast.cuh"

namespace lsi {

struct RowOwnerParams {
  const float* tex; const float* disp; const float* mask; const float* mats; const int* flags;
  float* img; float* wts;
  int L, B, H, W, h_t, w_t, R, bands;
  int compose;
  float ds, inv_max_disp, k2, k2h, bg_wt;
};

// acc[cell] += v for the lanes with `valid`; conflict-free rounds when two lanes of the warp share a cell
__device__ __forceinline__ void tile_rmw(float4* row, int cell, const float4& v, bool valid, bool distinct, int lane) {
  const unsigned full = 0xffffffffu;
  if (distinct) {
    if (valid) { float4 a = row[cell]; a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w; row[cell] = a; }
    __syncwarp();
    return;
  }
  const int key = valid ? cell : (-1 - lane);
  const unsigned grp = __match_any_sync(full, key);
  const int rank = __popc(grp & ((1u << lane) - 1u));
  const int rounds = __reduce_max_sync(full, __popc(grp));
  for (int r = 0; r < rounds; ++r) {
    if (valid && rank == r) { float4 a = row[cell]; a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w; row[cell] = a; }
    __syncwarp();
  }
}

// one target row's share of a warp step: left cells (pre-reduced with the left neighbour), then the right cells that
// no neighbour took over
__device__ __forceinline__ void rowowner_row(float4* row, int w_t, int q, float ol, float orr, float vr, float vg, float vb,
                                             float w, int q_prev, int q_next, float orr_prev, float vr_p, float vg_p,
                                             float vb_p, float w_p, int lane, bool increasing) {
  float4 left = make_float4(vr * ol, vg * ol, vb * ol, w * ol);
  if (lane > 0 && q_prev + 1 == q) {
    left.x = fmaf(vr_p, orr_prev, left.x); left.y = fmaf(vg_p, orr_prev, left.y);
    left.z = fmaf(vb_p, orr_prev, left.z); left.w = fmaf(w_p, orr_prev, left.w);
  }
  const bool lvalid = left.w != 0.f && (unsigned)q < (unsigned)w_t;
  tile_rmw(row, q, left, lvalid, increasing, lane);
  const bool taken = (lane < 31) && (q_next == q + 1);
  const bool rvalid = !taken && orr > 0.f && w != 0.f && (unsigned)(q + 1) < (unsigned)w_t;
  const unsigned rmask = __ballot_sync(0xffffffffu, rvalid);
  if (rmask)   // usually just lane 31 (its right neighbour lives in the next warp step): a single writer needs no ranking
    tile_rmw(row, q + 1, make_float4(vr * orr, vg * orr, vb * orr, w * orr), rvalid, (rmask & (rmask - 1)) == 0, lane);
}

template <bool kHasMask, bool kPacked>
__global__ void __launch_bounds__(128) splat_rowowner_kernel(const RowOwnerParams p) {
  extern __shared__ __align__(16) float4 tile_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int task = blockIdx.x * (blockDim.x >> 5) + warp;
  if (task >= p.B * p.bands) return;
  const int b = task / p.bands, band = task - b * p.bands;
  if (!p.flags[b]) return;                                  // not the rectified class: the reduction kernels handle it
  float4* tile = tile_smem + (size_t)warp * p.R * p.w_t;   // [R][w_t], warp private
  const int r0 = band * p.R;
  const int rows = min(p.R, p.h_t - r0);
  const Mat34 M = load_mat(p.mats, b);
  const unsigned full = 0xffffffffu;
  const int n_src = p.H * p.W;
  const int n_trg = p.h_t * p.w_t;
  FastParams fp;                                            // for zb_weight()
  fp.inv_max_disp = p.inv_max_disp; fp.k2 = p.k2; fp.k2h = p.k2h;

  // source rows that can touch [r0, r0+rows): y(i) = (M5*(i+0.5)+M6)*ds - 0.5 is increasing (flag requires M5 > 0)
  const float a = M.m[5] * p.ds, c = M.m[6] * p.ds - 0.5f;
  int i_lo = (int)floorf(((float)(r0 - 1) - c) / a - 0.5f) - 1, i_hi = (int)ceilf(((float)(r0 + rows) - c) / a - 0.5f) + 1;
  i_lo = max(i_lo, 0); i_hi = min(i_hi, p.H - 1);
  const int chunks = (p.W + 31) >> 5;
  const int l_outer = p.compose ? 1 : p.L;                 // per-layer outputs: one pass per layer

  for (int lo = 0; lo < l_outer; ++lo) {
    for (int k = lane; k < rows * p.w_t; k += 32) tile[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();
    const int l_begin = p.compose ? 0 : lo, l_end = p.compose ? p.L : lo + 1;
    for (int i = i_lo; i <= i_hi; ++i) {
      const float ys = (float)i + 0.5f;
      const float bv = fmaf(M.m[5], ys, 0.f) + M.m[6];     // == render_fast's bv with M4 == 0
      const AxisW ay = axis_weights(fmaf(bv, p.ds, -0.5f), p.h_t);
      const int rt = ay.i0 - r0, rb = rt + 1;
      const bool top = ay.w0 > 0.f && rt >= 0 && rt < rows;
      const bool bot = ay.w1 > 0.f && rb >= 0 && rb < rows;
      if (!top && !bot) continue;                           // warp uniform
      // layers in groups of 4; within a source row the loads of chunk ch+1 (all layers of the group) are issued before
      // the arithmetic of chunk ch, so every warp keeps 4..16 independent requests in flight (the kernel runs at ~16
      // warps per SM: latency has to be hidden inside the warp)
      for (int lg = l_begin; lg < l_end; lg += 4) {
        const int nl = min(4, l_end - lg);
        const size_t lstride = (size_t)p.B * n_src;
        const size_t row_base = (size_t)(lg * p.B + b) * n_src + (size_t)i * p.W;
        float ctr[4], ctg[4], ctb[4], cd[4], cmk[4], ntr[4], ntg[4], ntb[4], nd[4], nmk[4];
        auto load = [&](int ch, float* tr, float* tg, float* tb, float* d, float* mk) {
          const int j = (ch << 5) + lane;
          const bool active = j < p.W;
          const size_t im0 = row_base + (active ? j : 0);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (u < nl) {
              const size_t im = im0 + (size_t)u * lstride;
              if (kPacked) {
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
        };
        load(0, ctr, ctg, ctb, cd, cmk);
        for (int ch = 0; ch < chunks; ++ch) {
          if (ch + 1 < chunks) load(ch + 1, ntr, ntg, ntb, nd, nmk);
          const int j = (ch << 5) + lane;
          const float xs = (float)j + 0.5f;
          const float bu = fmaf(M.m[1], ys, M.m[0] * xs) + M.m[2];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (u < nl) {
              const float d = cd[u];
              const float x = fmaf(fmaf(M.m[3], d, bu), p.ds, -0.5f);
              const float w = zb_weight(d, fp) * cmk[u];
              const AxisW ax = axis_weights(x, p.w_t);
              const float vr = ctr[u] * w, vg = ctg[u] * w, vb = ctb[u] * w;
              const int q = ax.i0;
              const int q_prev = __shfl_up_sync(full, q, 1), q_next = __shfl_down_sync(full, q, 1);
              const float w1x_p = __shfl_up_sync(full, ax.w1, 1);
              const float vr_p = __shfl_up_sync(full, vr, 1), vg_p = __shfl_up_sync(full, vg, 1);
              const float vb_p = __shfl_up_sync(full, vb, 1), w_p = __shfl_up_sync(full, w, 1);
              const bool increasing = __all_sync(full, lane == 0 || q > q_prev);
              if (top)
                rowowner_row(tile + (size_t)rt * p.w_t, p.w_t, q, thresh(ax.w0 * ay.w0), thresh(ax.w1 * ay.w0), vr, vg, vb, w,
                             q_prev, q_next, thresh(w1x_p * ay.w0), vr_p, vg_p, vb_p, w_p, lane, increasing);
              if (bot)
                rowowner_row(tile + (size_t)rb * p.w_t, p.w_t, q, thresh(ax.w0 * ay.w1), thresh(ax.w1 * ay.w1), vr, vg, vb, w,
                             q_prev, q_next, thresh(w1x_p * ay.w1), vr_p, vg_p, vb_p, w_p, lane, increasing);
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) { ctr[u] = ntr[u]; ctg[u] = ntg[u]; ctb[u] = ntb[u]; cd[u] = nd[u]; cmk[u] = nmk[u]; }
        }
      }
    }
    __syncwarp();
    // normalise + write the owned rows once (ldi.py:165-173; bg canvases folded in as in normalize_kernel)
    const float nb = p.compose ? (float)p.L * p.bg_wt : p.bg_wt;
    const size_t out_base = ((size_t)lo * p.B + b) * n_trg + (size_t)r0 * p.w_t;
    for (int k = lane; k < rows * p.w_t; k += 32) {
      const float4 s = tile[k];
      const float Wsum = s.w + nb, Wh = safe_den(Wsum);
      float* ip = p.img + (out_base + k) * 3;
      __stcs(ip, (s.x + nb) / Wh); __stcs(ip + 1, (s.y + nb) / Wh); __stcs(ip + 2, (s.z + nb) / Wh);
      __stcs(p.wts + out_base + k, Wsum);
    }
    __syncwarp();
  }
}

}  // namespace lsi
