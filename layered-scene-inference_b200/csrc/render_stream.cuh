// Streaming forward-splat kernel: the default forward path.
//
// Same arithmetic as render_fast.cuh (one source pixel position x up to 4 layers per lane, consecutive lanes =
// consecutive pixels, neighbour pre-reduction, vector reductions into the L2-resident accumulator) but with the
// input stream taken out of the warps: the measured limit of the block-per-64-pixels kernel was not HBM but latency --
// loads, ~130 dependent instructions and the reductions of a thread ran back to back, one short-lived CTA per 64
// pixels (profiles/r1_splat_ablation.txt: 0.33 ms/step for loads + one coalesced reduction and no geometry at all).
// Here every warp is persistent, owns a contiguous range of (image, row, 64-pixel segment) units and a private ring
// of shared-memory stages that its elected lane fills with bulk asynchronous copies (cp.async.bulk -> UBLKCP, completion
// on an mbarrier): 3 stages x 4 KB per warp are in flight while the warp does the geometry of the current one, so
// the bytes in flight per SM no longer depend on occupancy x registers.  Row constants (matrix, pose class, vertical
// weights) are recomputed only when the warp's range crosses into a new row.
#pragma once
#include "render_fast.cuh"

namespace lsi {

// mbarrier.try_wait suspend-time hint: a waiting thread sleeps until the phase completes (or this many ns pass) instead
// of re-polling -- in the halo kernel 27 % of all issued instructions were YIELD/TRYWAIT/BRA of waiting warps
#ifndef LSI_SUSPEND_HINT_DEFINED
#define LSI_SUSPEND_HINT_DEFINED
constexpr unsigned kSuspendHintNs = 0x989680u;
#endif


constexpr int kSegPx = 64;          // source pixels of one row per unit (two 32-lane sub-steps)
constexpr int kStreamWarps = 4;     // warps per CTA, each with its own ring
#ifndef LSI_STREAM_QUAD
#define LSI_STREAM_QUAD 1
#endif
constexpr bool kStreamQuad = LSI_STREAM_QUAD != 0;   // interleave 4 layers per lane (else 2)
constexpr int kStreamCtasPerSm = kStreamQuad ? 4 : 6;

struct StreamParams {
  FastParams f;
  int segs;                 // segments per row = ceil(W / kSegPx)
  int groups;               // layer groups of <= 4 per unit = ceil(L / 4)
  int stages;               // ring depth per warp
  int stage_bytes;          // bytes of one stage (4 layers x kSegPx pixels x bytes per pixel)
  long long units;          // bc * H * segs
};

__device__ __forceinline__ uint32_t st_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void st_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(st_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void st_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(st_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void st_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(st_smem_u32(bar)), "r"(parity), "r"(kSuspendHintNs) : "memory");
}
// 1-D bulk copy global -> shared, evict-first in L2 (the LDI streams through once; the accumulator must stay resident)
__device__ __forceinline__ void st_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          st_smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(st_smem_u32(bar)), "l"(policy)
      : "memory");
}

struct StreamCursor {   // (image in chunk, row, segment, layer group) of an item, advanced incrementally
  int bl, i, s, g;
};

__device__ __forceinline__ void cursor_advance(StreamCursor& c, const StreamParams& p) {
  if (++c.g < p.groups) return;
  c.g = 0;
  if (++c.s < p.segs) return;
  c.s = 0;
  if (++c.i < p.f.H) return;
  c.i = 0; ++c.bl;
}

// stage layout: [tex: 4 x kSegPx x (16 | 12) B][disp: 4 x kSegPx x 4 B (planar only)][mask: 4 x kSegPx x 4 B (if any)]
template <bool kHasMask, bool kPacked>
__device__ __forceinline__ void stream_issue(const StreamParams& p, const StreamCursor& c, unsigned char* stage,
                                             uint64_t* bar, uint64_t policy) {
  const FastParams& f = p.f;
  const int j0 = c.s * kSegPx;
  const int npx = min(kSegPx, f.W - j0);
  const int nl = min(4, f.L - c.g * 4);
  constexpr int kTexB = kPacked ? 16 : 12;
  const uint32_t per_layer = (uint32_t)npx * (kTexB + (kPacked ? 0 : 4) + (kHasMask ? 4 : 0));
  st_mbar_expect_tx(bar, per_layer * nl);
  const size_t n_src = (size_t)f.H * f.W;
  size_t img = ((size_t)(c.g * 4) * f.B + (f.b0 + c.bl)) * n_src + (size_t)c.i * f.W + j0;
  const size_t lstride = (size_t)f.B * n_src;
  for (int u = 0; u < nl; ++u, img += lstride) {
    st_bulk_g2s(stage + u * kSegPx * kTexB, reinterpret_cast<const unsigned char*>(f.tex) + img * kTexB, npx * kTexB, bar,
                policy);
    if (!kPacked)
      st_bulk_g2s(stage + 4 * kSegPx * kTexB + u * kSegPx * 4, f.disp + img, npx * 4, bar, policy);
    if (kHasMask)
      st_bulk_g2s(stage + 4 * kSegPx * (kPacked ? 16 : 16) + u * kSegPx * 4, f.mask + img, npx * 4, bar, policy);
  }
}

struct RowConst {   // per (image, row): recomputed when a warp's range crosses a row boundary
  Mat34 M;
  AxisW ay_row;
  int mode;
  bool two_rows;   // mode 2: both target rows of this source row carry weight
};

__device__ __forceinline__ void row_setup(const FastParams& f, int b, int i, RowConst& r) {
  r.M = load_mat(f.mats, b);
  const Mat34& M = r.M;
  const bool affine = (M.m[8] == 0.f) && (M.m[9] == 0.f) && (M.m[10] == 1.f) && (M.m[11] == 0.f);
  const bool yconst = affine && (M.m[4] == 0.f) && (M.m[7] == 0.f);
  r.mode = yconst ? 2 : (affine ? 1 : 0);
  r.ay_row.i0 = 0; r.ay_row.w0 = 0.f; r.ay_row.w1 = 0.f;
  if (yconst) {
    const float ys = (float)i + 0.5f;
    const float bv = fmaf(M.m[5], ys, M.m[4] * 0.5f) + M.m[6];   // M[4] == 0: independent of x, same value as the per-pixel form
    r.ay_row = axis_weights(fmaf(bv, f.ds, -0.5f), f.h_t);
  }
  r.two_rows = r.ay_row.w0 > 0.f && r.ay_row.w1 > 0.f;
}

__device__ __forceinline__ bool st_elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
  return pred != 0;
}

// base + q cells, as one IMAD.WIDE (the compiler otherwise rebuilds the 64-bit element index: 4-5 ALU-pipe instructions)
__device__ __forceinline__ float4* cell_ptr(float4* base, int q) {
  float4* r;
  asm("mad.wide.s32 %0, %1, 16, %2;" : "=l"(r) : "r"(q), "l"(base));
  return r;
}

__device__ __forceinline__ void red4(float4* p, float x, float y, float z, float w) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}

// One 32-pixel sub-step of a stage: the arithmetic of splat_pixel (render_fast.cuh) restated as straight-line code over
// the kNL layers of the group, so that their dependent chains (LDS -> FMA -> MUFU -> F2I -> SHFL -> RED) interleave
// instead of running back to back behind branches.  What the ncu source view of the first version asked for
// (profiles/r1_splat_stream_*): the half-rate ALU pipe (compares, selects, logic, 64-bit adds) was the busiest unit, so
//   * the left-cell reduction is unconditional: a lane with nothing to add adds exact zeros -- to its own cell or, when the
//     cell is outside the image, to a neighbouring image's cell or the guard band around the chunk accumulator
//     (plan_forward reserves 2 rows + 4 cells on both sides; coordinates are clamped to [-2, extent + 1]);
//   * the sender thresholds its right-cell weights once and ships them (6 shuffles for one target row, 7 for two);
//   * cell addresses are one IMAD.WIDE off a per-item base pointer;
//   * only the (rare: lane 31, disparity jumps) unmerged right cells take a branch.
template <int kMode, bool kTwoRows, int kNL, bool kHasMask>
__device__ __forceinline__ void splat_group(const FastParams& f, const RowConst& rc, int i, int j, const float4 (&v)[4],
                                            const float (&mk)[4], float4* const (&base)[4], int lane) {
  const Mat34& M = rc.M;
  const float xs = (float)j + 0.5f, ys = (float)i + 0.5f;            // helpers.py:88-113
  const float bu = fmaf(M.m[1], ys, M.m[0] * xs) + M.m[2];
  float bv = 0.f, bn = 1.f;
  if (kMode != 2) bv = fmaf(M.m[5], ys, M.m[4] * xs) + M.m[6];
  if (kMode == 0) bn = fmaf(M.m[9], ys, M.m[8] * xs) + M.m[10];
  const unsigned full = 0xffffffffu;
  const unsigned next_bit = 2u << lane;   // lane 31: 0
  float vr[kNL], vg[kNL], vb[kNL], w[kNL], ol0[kNL], or0[kNL], ol1[kNL], or1[kNL];   // ol1/or1 dead unless kTwoRows
  int q[kNL];
#pragma unroll
  for (int u = 0; u < kNL; ++u) {
    const float d = v[u].w;
    float x, y = 0.f, dt;
    const float up = fmaf(M.m[3], d, bu);
    if (kMode == 0) {
      const float vp = fmaf(M.m[7], d, bv);
      const float nh = safe_den(fmaf(M.m[11], d, bn));
      x = (up / nh) * f.ds - 0.5f; y = (vp / nh) * f.ds - 0.5f; dt = d / nh;
    } else {
      x = fmaf(up, f.ds, -0.5f); dt = d;
      if (kMode == 1) y = fmaf(fmaf(M.m[7], d, bv), f.ds, -0.5f);
    }
    w[u] = zb_weight(dt, f);
    if (kHasMask) w[u] *= mk[u];
    const AxisW ax = axis_weights(x, f.w_t);
    const AxisW ay = (kMode == 2) ? rc.ay_row : axis_weights(y, f.h_t);
    vr[u] = v[u].x * w[u]; vg[u] = v[u].y * w[u]; vb[u] = v[u].z * w[u];
    if (kMode == 2 && !kTwoRows) {   // the one row with weight (row 1 when row 0 has none)
      const bool first_is_1 = !(rc.ay_row.w0 > 0.f);
      const float wy0 = first_is_1 ? ay.w1 : ay.w0;
      q[u] = (ay.i0 + (first_is_1 ? 1 : 0)) * f.w_t + ax.i0;
      ol0[u] = thresh(ax.w0 * wy0); or0[u] = thresh(ax.w1 * wy0);
      ol1[u] = 0.f; or1[u] = 0.f;
    } else {
      q[u] = ay.i0 * f.w_t + ax.i0;
      ol0[u] = thresh(ax.w0 * ay.w0); or0[u] = thresh(ax.w1 * ay.w0);
      ol1[u] = thresh(ax.w0 * ay.w1); or1[u] = thresh(ax.w1 * ay.w1);
    }
  }
  // neighbour exchange: lane k folds lane k-1's right cell into its left cell when the two coincide
  float jf[kNL], vr_p[kNL], vg_p[kNL], vb_p[kNL], w_p[kNL], or0_p[kNL], or1_p[kNL];
  unsigned joins[kNL];
#pragma unroll
  for (int u = 0; u < kNL; ++u) {
    const int q_prev = __shfl_up_sync(full, q[u], 1);
    or0_p[u] = __shfl_up_sync(full, or0[u], 1);
    vr_p[u] = __shfl_up_sync(full, vr[u], 1); vg_p[u] = __shfl_up_sync(full, vg[u], 1);
    vb_p[u] = __shfl_up_sync(full, vb[u], 1); w_p[u] = __shfl_up_sync(full, w[u], 1);
    or1_p[u] = 0.f;
    if (kTwoRows) or1_p[u] = __shfl_up_sync(full, or1[u], 1);
    const bool joined = (q_prev + 1 == q[u]) && lane > 0;
    joins[u] = __ballot_sync(full, joined);
    jf[u] = joined ? 1.f : 0.f;
  }
#pragma unroll
  for (int u = 0; u < kNL; ++u) {
    float4* cell = cell_ptr(base[u], q[u]);
    {
      const float op = or0_p[u] * jf[u];
      red4(cell, fmaf(vr_p[u], op, vr[u] * ol0[u]), fmaf(vg_p[u], op, vg[u] * ol0[u]), fmaf(vb_p[u], op, vb[u] * ol0[u]),
           fmaf(w_p[u], op, w[u] * ol0[u]));
      const float rw = w[u] * or0[u];
      if (rw != 0.f && !(joins[u] & next_bit)) red4(cell + 1, vr[u] * or0[u], vg[u] * or0[u], vb[u] * or0[u], rw);
    }
    if (kTwoRows) {
      cell = cell_ptr(cell, f.w_t);
      const float op = or1_p[u] * jf[u];
      red4(cell, fmaf(vr_p[u], op, vr[u] * ol1[u]), fmaf(vg_p[u], op, vg[u] * ol1[u]), fmaf(vb_p[u], op, vb[u] * ol1[u]),
           fmaf(w_p[u], op, w[u] * ol1[u]));
      const float rw = w[u] * or1[u];
      if (rw != 0.f && !(joins[u] & next_bit)) red4(cell + 1, vr[u] * or1[u], vg[u] * or1[u], vb[u] * or1[u], rw);
    }
  }
}

// kN (1, 2 or 4) layers of a stage x the stage's 32-lane sub-steps, as a rolled loop whose body is one splat_group.
// Shipped configuration (kStreamQuad): all four layers of a group interleaved in one body (~350 instructions, 128
// registers, 16 resident warps per SM); the two-layer variant (80 registers, 24 warps) and a fully unrolled stage
// (~1100 instructions: instruction-fetch stalls) measured slower (profiles/r1_splat_stream_versions.txt).
// `u0` = first layer of the group within the stage; `last` = these are the stage's last shared-memory reads.
template <int kMode, bool kTwoRows, int kN, bool kHasMask, bool kPacked>
__device__ __forceinline__ void pair_run(const StreamParams& p, const RowConst& rc, const unsigned char* stage, int u0,
                                         int i, int j0, int n_sub, float4* pb0, size_t acc_lstride, int lane, bool last,
                                         bool refill, const StreamCursor& pc, uint64_t* bar, uint64_t policy) {
  const FastParams& f = p.f;
  float4* pb[4] = {pb0, pb0 + acc_lstride, pb0 + 2 * acc_lstride, pb0 + 3 * acc_lstride};
  const unsigned char* tex_ptr = stage + (u0 * kSegPx + lane) * (kPacked ? 16 : 12);
  const unsigned char* aux_ptr = stage + 4 * kSegPx * 12 + (u0 * kSegPx + lane) * 4;    // planar disparity
  const unsigned char* msk_ptr = stage + 4 * kSegPx * 16 + (u0 * kSegPx + lane) * 4;
#pragma unroll 1
  for (int h = 0; h < n_sub; ++h) {
    float4 v[4];
    float mk[4];
#pragma unroll
    for (int k = 0; k < kN; ++k) {
      if (kPacked) {
        v[k] = *reinterpret_cast<const float4*>(tex_ptr + k * kSegPx * 16);
      } else {
        const float* t = reinterpret_cast<const float*>(tex_ptr + k * kSegPx * 12);
        v[k].x = t[0]; v[k].y = t[1]; v[k].z = t[2];
        v[k].w = *reinterpret_cast<const float*>(aux_ptr + k * kSegPx * 4);
      }
      mk[k] = 1.f;
      if (kHasMask) mk[k] = *reinterpret_cast<const float*>(msk_ptr + k * kSegPx * 4);
    }
    if (last && h == n_sub - 1) {   // hand the slot back to the copy engine
      __syncwarp();
      if (refill && st_elect_one()) stream_issue<kHasMask, kPacked>(p, pc, const_cast<unsigned char*>(stage), bar, policy);
    }
    splat_group<kMode, kTwoRows, kN, kHasMask>(f, rc, i, j0 + h * 32 + lane, v, mk, pb, lane);
    tex_ptr += 32 * (kPacked ? 16 : 12); aux_ptr += 32 * 4; msk_ptr += 32 * 4;
  }
}

// One stage = nl (<= 4) layers x kSegPx pixels of one source row, layer pair by layer pair.
template <int kMode, bool kTwoRows, bool kHasMask, bool kPacked>
__device__ __forceinline__ void stage_run(const StreamParams& p, const RowConst& rc, unsigned char* stage, int i, int j0,
                                       float4* base0, size_t acc_lstride, int nl, int lane, bool refill,
                                       const StreamCursor& pc, uint64_t* bar, uint64_t policy) {
  const FastParams& f = p.f;
  const int npx = min(kSegPx, f.W - j0);
  const int n_sub = (npx + 31) >> 5;
  if (npx & 31) {
    // ragged row end: the lanes past it would read stale shared memory -> write exact zeros there (zero disparity =
    // zero weight, added to a valid cell).  Generic-proxy writes that the next bulk copy into this slot overwrites:
    // ordered by the proxy fence below.
    for (int u = 0; u < nl; ++u)
      for (int px = npx + lane; px < n_sub * 32; px += 32) {
        if (kPacked) {
          *reinterpret_cast<float4*>(stage + (u * kSegPx + px) * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
          float* t = reinterpret_cast<float*>(stage + (u * kSegPx + px) * 12);
          t[0] = 0.f; t[1] = 0.f; t[2] = 0.f;
          *reinterpret_cast<float*>(stage + 4 * kSegPx * 12 + (u * kSegPx + px) * 4) = 0.f;
        }
        if (kHasMask) *reinterpret_cast<float*>(stage + 4 * kSegPx * 16 + (u * kSegPx + px) * 4) = 0.f;
      }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
  }
  if (kStreamQuad && nl == 4) {   // all four layers interleaved (128 registers, 16 warps per SM)
    pair_run<kMode, kTwoRows, 4, kHasMask, kPacked>(p, rc, stage, 0, i, j0, n_sub, base0, acc_lstride, lane, true, refill, pc, bar, policy);
    return;
  }
  const int pairs = (nl + 1) >> 1;
#pragma unroll 1
  for (int pr = 0; pr < pairs; ++pr) {
    const int u0 = 2 * pr;
    float4* pb0 = base0 + (size_t)u0 * acc_lstride;
    const bool last = pr == pairs - 1;
    if (u0 + 1 < nl)
      pair_run<kMode, kTwoRows, 2, kHasMask, kPacked>(p, rc, stage, u0, i, j0, n_sub, pb0, acc_lstride, lane, last, refill, pc, bar, policy);
    else
      pair_run<kMode, kTwoRows, 1, kHasMask, kPacked>(p, rc, stage, u0, i, j0, n_sub, pb0, acc_lstride, lane, last, refill, pc, bar, policy);
  }
}

template <bool kHasMask, bool kPacked>
__global__ void __launch_bounds__(kStreamWarps * 32, kStreamCtasPerSm) splat_fwd_stream_kernel(const StreamParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const FastParams& f = p.f;
  if (f.skip && *f.n_flagged == f.B) return;   // every image was rendered by the row-gather kernel (render_rowgather.cuh)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // barriers first (8 B each), then the rings
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem) + warp * p.stages;
  unsigned char* ring = smem + 128 + (size_t)warp * p.stages * p.stage_bytes;
  if (lane == 0) {
    for (int s = 0; s < p.stages; ++s) st_mbar_init(bars + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  uint64_t policy;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));

  // contiguous unit range of this warp
  const long long wg = (long long)blockIdx.x * kStreamWarps + warp, wt = (long long)gridDim.x * kStreamWarps;
  const long long u0 = p.units * wg / wt, u1 = p.units * (wg + 1) / wt;
  if (u0 >= u1) return;
  StreamCursor cc;   // consumer cursor
  {
    const long long per_img = (long long)f.H * p.segs;
    cc.bl = (int)(u0 / per_img);
    const int rem = (int)(u0 - (long long)cc.bl * per_img);
    cc.i = rem / p.segs; cc.s = rem - cc.i * p.segs; cc.g = 0;
  }
  StreamCursor pc = cc;   // producer cursor (kept by every lane, advanced uniformly)
  const int items = (int)(u1 - u0) * p.groups;   // host keeps units * groups per warp below 2^31
  int issued = 0;
  for (; issued < p.stages && issued < items; ++issued) {
    if (st_elect_one())
      stream_issue<kHasMask, kPacked>(p, pc, ring + (size_t)issued * p.stage_bytes, bars + issued, policy);
    cursor_advance(pc, p);
  }

  const int n_trg = f.h_t * f.w_t;
  const size_t acc_lstride = f.acc_per_layer ? (size_t)f.bc * n_trg : 0;
  RowConst rc;
  int row_b = -1, row_i = -1;
  bool row_skip = false;
  int slot = 0;
  uint32_t parity = 0;
  for (int it = 0; it < items; ++it) {
    if (cc.bl != row_b || cc.i != row_i) {
      row_b = cc.bl; row_i = cc.i;
      row_setup(f, f.b0 + cc.bl, cc.i, rc);
      row_skip = f.skip && f.skip[f.b0 + cc.bl];   // rendered elsewhere: its stages are streamed (mixed batches are rare) but not processed
    }
    unsigned char* stage = ring + (size_t)slot * p.stage_bytes;
    st_mbar_wait(bars + slot, parity);
    const int nl = min(4, f.L - cc.g * 4);
    const int j0 = cc.s * kSegPx;
    float4* base0 = f.acc4 + (size_t)cc.bl * n_trg + (size_t)(cc.g * 4) * acc_lstride;
    const bool refill = issued < items;
    if (row_skip) {
      __syncwarp();
      if (refill && st_elect_one()) stream_issue<kHasMask, kPacked>(p, pc, stage, bars + slot, policy);
    } else if (rc.mode == 2) {
      if (rc.two_rows) stage_run<2, true, kHasMask, kPacked>(p, rc, stage, cc.i, j0, base0, acc_lstride, nl, lane, refill, pc, bars + slot, policy);
      else stage_run<2, false, kHasMask, kPacked>(p, rc, stage, cc.i, j0, base0, acc_lstride, nl, lane, refill, pc, bars + slot, policy);
    } else if (rc.mode == 1) {
      stage_run<1, true, kHasMask, kPacked>(p, rc, stage, cc.i, j0, base0, acc_lstride, nl, lane, refill, pc, bars + slot, policy);
    } else {
      stage_run<0, true, kHasMask, kPacked>(p, rc, stage, cc.i, j0, base0, acc_lstride, nl, lane, refill, pc, bars + slot, policy);
    }
    if (refill) { cursor_advance(pc, p); ++issued; }
    cursor_advance(cc, p);
    if (++slot == p.stages) { slot = 0; parity ^= 1; }
  }
}

}  // namespace lsi
