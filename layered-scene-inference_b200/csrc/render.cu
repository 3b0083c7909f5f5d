// LDI renderer for B200 (sm_100a): fused project -> z-weight -> bilinear forward splat -> normalise/compose, and
// its gather-form backward.  Replaces the TF graph built by lsi/geometry/ldi.py:71-182 (reference tree paths).
//
// Data flow (forward), per batch chunk that keeps the accumulators L2-resident:
//   proj_matrix_kernel      k_s,k_t,rot,t -> M[B,4,4]                         (projection.py:71-86)
//   splat_fwd_*_kernel      one thread per (layer, source pixel): registers only, then vector reductions
//                           (red.global.add.v4.f32) into acc4[b][q] = (sum w*omega*rgb, sum w*omega)
//   normalize_kernel        acc4 (+ per-layer (sum w, sum w*d)) -> trg_img, trg_wts, trg_disp   (ldi.py:165-173)
#include "capi_common.h"
#include "common.cuh"
#include "render_fast.cuh"
#include "render_stream.cuh"
#include "render_rowowner.cuh"
#include "render_rowgather.cuh"

namespace lsi {

// ---------------------------------------------------------------------------------------------------------
// projection.py:27-106.  One thread per batch element; fp64 inside so the fp32 result is the correctly
// rounded matrix (the reference rounds every intermediate to fp32; both are within fp32 noise of this).
// ---------------------------------------------------------------------------------------------------------
__device__ void inv3x3(const double* a, double* o) {
  double c00 = a[4] * a[8] - a[5] * a[7], c01 = a[5] * a[6] - a[3] * a[8], c02 = a[3] * a[7] - a[4] * a[6];
  double det = a[0] * c00 + a[1] * c01 + a[2] * c02;
  double id = 1.0 / det;
  o[0] = c00 * id; o[1] = (a[2] * a[7] - a[1] * a[8]) * id; o[2] = (a[1] * a[5] - a[2] * a[4]) * id;
  o[3] = c01 * id; o[4] = (a[0] * a[8] - a[2] * a[6]) * id; o[5] = (a[2] * a[3] - a[0] * a[5]) * id;
  o[6] = c02 * id; o[7] = (a[1] * a[6] - a[0] * a[7]) * id; o[8] = (a[0] * a[4] - a[1] * a[3]) * id;
}

__global__ void proj_matrix_kernel(const float* __restrict__ k_s, const float* __restrict__ k_t,
                                   const float* __restrict__ rot, const float* __restrict__ t, int batch, int inverse,
                                   float* __restrict__ out, int* __restrict__ rect_flags) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  double ks[9], kt[9], r[9], tr[3], kinv[9], rt[9], tt[3];
  for (int i = 0; i < 9; ++i) { ks[i] = k_s[b * 9 + i]; kt[i] = k_t[b * 9 + i]; r[i] = rot[b * 9 + i]; }
  for (int i = 0; i < 3; ++i) tr[i] = t[b * 3 + i];
  const double* kout;
  if (!inverse) {  // pad(K_t) . [R t; 0 1] . pad(K_s^-1)
    inv3x3(ks, kinv);
    for (int i = 0; i < 9; ++i) rt[i] = r[i];
    for (int i = 0; i < 3; ++i) tt[i] = tr[i];
    kout = kt;
  } else {         // pad(K_s) . [R^T  -R^T t; 0 1] . pad(K_t^-1)
    inv3x3(kt, kinv);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) rt[i * 3 + j] = r[j * 3 + i];
    for (int i = 0; i < 3; ++i) tt[i] = -(rt[i * 3] * tr[0] + rt[i * 3 + 1] * tr[1] + rt[i * 3 + 2] * tr[2]);
    kout = ks;
  }
  double rk[9], m3[9], col[3];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j)
    rk[i * 3 + j] = rt[i * 3] * kinv[j] + rt[i * 3 + 1] * kinv[3 + j] + rt[i * 3 + 2] * kinv[6 + j];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j)
    m3[i * 3 + j] = kout[i * 3] * rk[j] + kout[i * 3 + 1] * rk[3 + j] + kout[i * 3 + 2] * rk[6 + j];
  for (int i = 0; i < 3; ++i) col[i] = kout[i * 3] * tt[0] + kout[i * 3 + 1] * tt[1] + kout[i * 3 + 2] * tt[2];
  float* o = out + (size_t)b * 16;
  for (int i = 0; i < 3; ++i) {
    o[i * 4 + 0] = (float)m3[i * 3 + 0]; o[i * 4 + 1] = (float)m3[i * 3 + 1]; o[i * 4 + 2] = (float)m3[i * 3 + 2];
    o[i * 4 + 3] = (float)col[i];
  }
  o[12] = 0.f; o[13] = 0.f; o[14] = 0.f; o[15] = 1.f;
  // rectified pose class (row-owner kernel): n == 1, y' independent of x and d, rows keep their order
  if (rect_flags) {      // rect_flags[batch] counts the flagged images (zeroed by the caller)
    const int f = (o[8] == 0.f && o[9] == 0.f && o[10] == 1.f && o[11] == 0.f && o[4] == 0.f && o[7] == 0.f && o[5] > 0.f) ? 1 : 0;
    rect_flags[b] = f;
    if (f) atomicAdd(rect_flags + batch, 1);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Forward splat, global-reduction variant.
// ---------------------------------------------------------------------------------------------------------
struct FwdParams {
  const float* tex; const float* mask; const float* disp; const float* pc; const float* focal; const float* mats;
  float4* acc4;   // [nl_acc][bc][Nt]  (chunk-local)
  float2* accd;   // [L][accd_b][Nt]   (sum w*omega, sum w*omega*d_t) per layer, or nullptr
  int L, B, H, W, b0, bc;
  int tex_s, disp_s, mask_s;
  int acc_per_layer;       // 0: all layers reduce into one accumulator (compose), 1: per layer
  int accd_b, accd_b0;
  GeomParams gp;
};

template <bool kHasMask, bool kHasPc, bool kHasDisp>
__global__ void __launch_bounds__(256) splat_fwd_atomic_kernel(const FwdParams p) {
  const int n_src = p.H * p.W;
  const int n_trg = p.gp.h_t * p.gp.w_t;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= n_src) return;
  const int l = blockIdx.y, bl = blockIdx.z, b = p.b0 + bl;
  const size_t img = (size_t)(l * p.B + b) * n_src + pix;
  const Mat34 M = load_mat(p.mats, b);
  const int i = pix / p.W, j = pix - i * p.W;
  float xs, ys, one;
  if (kHasPc) {
    const float* c = p.pc + ((size_t)b * n_src + pix) * 3;
    xs = c[0]; ys = c[1]; one = c[2];
  } else {
    xs = (float)j + 0.5f; ys = (float)i + 0.5f; one = 1.f;   // helpers.py:88-113
  }
  const float focal = p.focal ? __ldg(p.focal + b) : 0.f;
  const float d = ld_stream(p.disp + img * p.disp_s) - focal;   // ldi.py:131-132
  PixGeom g;
  project(M, xs, ys, one, d, focal, p.gp, g);
  float w = g.zb;
  if (kHasMask) w *= ld_stream(p.mask + img * p.mask_s);       // ldi.py:145-146
  corners(g.x, g.y, p.gp.w_t, p.gp.h_t, g);
  const float* tp = p.tex + img * p.tex_s;
  const float vr = ld_stream(tp) * w, vg = ld_stream(tp + 1) * w, vb = ld_stream(tp + 2) * w;   // ldi.py:148
  const float vd = g.dt * w;                                                                    // ldi.py:149
  float4* acc = p.acc4 + ((size_t)(p.acc_per_layer ? l : 0) * p.bc + bl) * n_trg;
  float2* accd = kHasDisp ? p.accd + ((size_t)l * p.accd_b + (b - p.accd_b0)) * n_trg : nullptr;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    if (g.keep[c]) {
      const int q = corner_index(g, c, p.gp.w_t);
      const float o = g.w[c];
      atomicAdd(acc + q, make_float4(vr * o, vg * o, vb * o, w * o));   // REDG.E.ADD.F32x4
      if (kHasDisp) atomicAdd(accd + q, make_float2(w * o, vd * o));
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// ldi.py:165-173: per-layer disparity normalisation, compose (sum, sum, max), image normalisation.
// bg canvases (ldi.py:115-125) are folded in here: every layer canvas starts at bg_wt for img and wts.
// ---------------------------------------------------------------------------------------------------------
// zero the chunk accumulators of the images the reduction kernels will render (skips row-owner images)
__global__ void __launch_bounds__(256) zero_acc_kernel(float4* __restrict__ acc4, int nl_acc, int bc, int n_trg, int b0,
                                                       const int* __restrict__ skip) {
  const int bl = blockIdx.y;
  if (skip && skip[b0 + bl]) return;
  for (int l = 0; l < nl_acc; ++l) {
    float4* a = acc4 + ((size_t)l * bc + bl) * n_trg;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_trg; i += gridDim.x * blockDim.x) a[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

struct NormParams {
  const float4* acc4; const float2* accd;
  float* img; float* wts; float* disp;
  int L, B, b0, bc, n_trg, compose, accd_b, accd_b0;
  float bg_wt;
  const int* skip;
};

__global__ void __launch_bounds__(256) normalize_kernel(const NormParams p) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= p.n_trg) return;
  const int bl = blockIdx.y, b = p.b0 + bl;
  if (p.skip && p.skip[b]) return;
  const int lo = blockIdx.z;   // output layer (0 when composing)
  const float4 a = p.acc4[((size_t)lo * p.bc + bl) * p.n_trg + q];
  const float nb = p.compose ? (float)p.L * p.bg_wt : p.bg_wt;
  const float W = a.w + nb;
  const float Wh = safe_den(W);
  const size_t o = ((size_t)lo * p.B + b) * p.n_trg + q;
  float* ip = p.img + o * 3;
  __stcs(ip, (a.x + nb) / Wh); __stcs(ip + 1, (a.y + nb) / Wh); __stcs(ip + 2, (a.z + nb) / Wh);
  __stcs(p.wts + o, W);
  if (p.disp) {
    float dmax;
    if (p.compose) {
      dmax = -INFINITY;
      for (int l = 0; l < p.L; ++l) {
        const float2 s = p.accd[((size_t)l * p.accd_b + (b - p.accd_b0)) * p.n_trg + q];
        dmax = fmaxf(dmax, s.y / safe_den(s.x + p.bg_wt));
      }
    } else {
      const float2 s = p.accd[((size_t)lo * p.accd_b + (b - p.accd_b0)) * p.n_trg + q];
      dmax = s.y / safe_den(s.x + p.bg_wt);
    }
    __stcs(p.disp + o, dmax);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Backward.  Stage A (per target pixel): gradients w.r.t. the accumulators.  Stage B (per source pixel): gather
// them at the four corners and push through weights / z-weight / projection in registers.  No atomics.
// ---------------------------------------------------------------------------------------------------------
struct BwdAParams {
  const float* img; const float* wts; const float2* accd;   // saved forward outputs (+ per-layer accumulators)
  const float* g_img; const float* g_wts; const float* g_disp;
  float4* g4;   // [nl_g][bc][Nt]: (dL/dA_img rgb, dL/dA_w)
  float* gd;    // [L][bc][Nt]: dL/dA_d, or nullptr
  int L, B, b0, bc, n_trg, compose, nl_g;
  float bg_wt;
};

__global__ void __launch_bounds__(256) splat_bwd_target_kernel(const BwdAParams p) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= p.n_trg) return;
  const int bl = blockIdx.y, b = p.b0 + bl;
  const int nl_out = p.compose ? 1 : p.L;
  for (int lo = 0; lo < nl_out; ++lo) {
    const size_t o = ((size_t)lo * p.B + b) * p.n_trg + q;
    const float W = p.wts[o], Wh = safe_den(W);
    float gr = 0.f, gg = 0.f, gb = 0.f, gw = 0.f;
    if (p.g_img) {
      const float* gi = p.g_img + o * 3; const float* im = p.img + o * 3;
      const float a = gi[0], bb = gi[1], c = gi[2];
      gr = a / Wh; gg = bb / Wh; gb = c / Wh;
      // I = A/Wh with Wh = W + eps*[W==0] (the indicator carries no gradient) => dI/dW = -A/Wh^2 = -I/Wh
      gw = -(a * im[0] + bb * im[1] + c * im[2]) / Wh;
    }
    if (p.g_wts) gw += p.g_wts[o];
    if (!p.gd) {
      p.g4[((size_t)lo * p.bc + bl) * p.n_trg + q] = make_float4(gr, gg, gb, gw);
      continue;
    }
    // trg_disp gradient: D_l = A_d[l] / safe(A_w[l]); compose -> max over layers (gradient split over ties)
    const float gD = p.g_disp ? p.g_disp[o] : 0.f;
    if (p.compose) {
      float dmax = -INFINITY; int nsel = 0;
      for (int l = 0; l < p.L; ++l) {
        const float2 s = p.accd[((size_t)l * p.B + b) * p.n_trg + q];
        const float dl = s.y / safe_den(s.x + p.bg_wt);
        if (dl > dmax) { dmax = dl; nsel = 1; } else if (dl == dmax) { ++nsel; }
      }
      for (int l = 0; l < p.L; ++l) {
        const float2 s = p.accd[((size_t)l * p.B + b) * p.n_trg + q];
        const float awh = safe_den(s.x + p.bg_wt);
        const float dl = s.y / awh;
        float gdl = 0.f, gwl = 0.f;
        if (dl == dmax) { gdl = gD / ((float)nsel * awh); gwl = -gD * dl / ((float)nsel * awh); }
        p.g4[((size_t)l * p.bc + bl) * p.n_trg + q] = make_float4(gr, gg, gb, gw + gwl);
        p.gd[((size_t)l * p.bc + bl) * p.n_trg + q] = gdl;
      }
    } else {
      const float2 s = p.accd[((size_t)lo * p.B + b) * p.n_trg + q];
      const float dl = s.y / Wh;
      const float gwl = -gD * dl / Wh;
      p.g4[((size_t)lo * p.bc + bl) * p.n_trg + q] = make_float4(gr, gg, gb, gw + gwl);
      p.gd[((size_t)lo * p.bc + bl) * p.n_trg + q] = gD / Wh;
    }
  }
}

struct BwdBParams {
  const float* tex; const float* mask; const float* disp; const float* pc; const float* focal; const float* mats;
  const float4* g4; const float* gd;
  float* d_tex; float* d_mask; float* d_disp;
  int L, B, H, W, b0, bc;
  int tex_s, disp_s, mask_s;
  int g_per_layer;
  GeomParams gp;
};

template <bool kHasMask, bool kHasPc, bool kHasGd>
__global__ void __launch_bounds__(256) splat_bwd_source_kernel(const BwdBParams p) {
  const int n_src = p.H * p.W;
  const int n_trg = p.gp.h_t * p.gp.w_t;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= n_src) return;
  const int l = blockIdx.y, bl = blockIdx.z, b = p.b0 + bl;
  const size_t img = (size_t)(l * p.B + b) * n_src + pix;
  const Mat34 M = load_mat(p.mats, b);
  const int i = pix / p.W, j = pix - i * p.W;
  float xs, ys, one;
  if (kHasPc) {
    const float* c = p.pc + ((size_t)b * n_src + pix) * 3;
    xs = c[0]; ys = c[1]; one = c[2];
  } else {
    xs = (float)j + 0.5f; ys = (float)i + 0.5f; one = 1.f;
  }
  const float focal = p.focal ? __ldg(p.focal + b) : 0.f;
  const float d = ld_stream(p.disp + img * p.disp_s) - focal;
  PixGeom g;
  project(M, xs, ys, one, d, focal, p.gp, g);
  const float mk = kHasMask ? ld_stream(p.mask + img * p.mask_s) : 1.f;
  const float w = g.zb * mk;
  corners(g.x, g.y, p.gp.w_t, p.gp.h_t, g);
  const float* tp = p.tex + img * p.tex_s;
  const float tr = ld_stream(tp), tg = ld_stream(tp + 1), tb = ld_stream(tp + 2);
  const float4* g4 = p.g4 + ((size_t)(p.g_per_layer ? l : 0) * p.bc + bl) * n_trg;
  const float* gdp = kHasGd ? p.gd + ((size_t)l * p.bc + bl) * n_trg : nullptr;
  float Gr = 0.f, Gg = 0.f, Gb = 0.f, Gw = 0.f, Gd = 0.f, gx = 0.f, gy = 0.f;
  // d omega_c / dx and / dy (sampling.py:208-216); zero for dropped corners (sampling.py:219-222)
  const float dwx[4] = {-g.vx0 * g.wy0, g.vx1 * g.wy0, -g.vx0 * g.wy1, g.vx1 * g.wy1};
  const float dwy[4] = {-g.wx0 * g.vy0, -g.wx1 * g.vy0, g.wx0 * g.vy1, g.wx1 * g.vy1};
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    if (g.keep[c]) {
      const int q = corner_index(g, c, p.gp.w_t);
      const float4 a = __ldg(g4 + q);
      const float ad = kHasGd ? __ldg(gdp + q) : 0.f;
      const float o = g.w[c];
      Gr = fmaf(o, a.x, Gr); Gg = fmaf(o, a.y, Gg); Gb = fmaf(o, a.z, Gb); Gw = fmaf(o, a.w, Gw);
      if (kHasGd) Gd = fmaf(o, ad, Gd);
      const float s = w * (tr * a.x + tg * a.y + tb * a.z + a.w + g.dt * ad);   // dL/d omega_c
      gx = fmaf(s, dwx[c], gx); gy = fmaf(s, dwy[c], gy);
    }
  }
  float* dt_ = p.d_tex + img * 3;
  __stcs(dt_, w * Gr); __stcs(dt_ + 1, w * Gg); __stcs(dt_ + 2, w * Gb);
  const float g_w = tr * Gr + tg * Gg + tb * Gb + Gw + g.dt * Gd;
  if (p.d_mask) __stcs(p.d_mask + img, g.zb * g_w);
  // zbuffer_weights: clip passes gradient on the closed interval; [r>0] is already inside zb (helpers.py:189-192)
  const float clipg = (g.r >= 0.f && g.r <= 1.f) ? 1.f : 0.f;
  const float g_dt = g_w * mk * g.zb * (p.gp.scale * p.gp.inv_max_disp) * clipg + w * Gd;
  const float inh = 1.f / g.nh;
  const float gxs = gx * p.gp.ds, gys = gy * p.gp.ds;
  const float g_u = gxs * inh, g_v = gys * inh, g_dp = g_dt * inh;
  const float g_n = -(gxs * g.up + gys * g.vp + g_dt * g.dp) * inh * inh;
  __stcs(p.d_disp + img, M.m[3] * g_u + M.m[7] * g_v + M.m[11] * g_n + g_dp);
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
static bool rowowner_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("LSI_B200_ROWOWNER");   // measured slower than the reduction kernels: opt-in
    on = (e && atoi(e) != 0) ? 1 : 0;
  }
  return on == 1;
}

static int chunk_budget_mb() {
  static int mb = -1;
  if (mb < 0) {
    const char* e = getenv("LSI_B200_CHUNK_MB");
    mb = e ? atoi(e) : 64;
    if (mb < 1) mb = 1;
  }
  return mb;
}

struct FwdPlan { int nl_acc, bc; size_t off_mats, off_acc4, off_accd, total; };

static FwdPlan plan_forward(const lsi_b200_splat_desc* d) {
  FwdPlan pl;
  const size_t n_trg = (size_t)d->h_t * d->w_t;
  pl.nl_acc = d->compose_layers ? 1 : d->n_layers;
  const size_t per_img = n_trg * (16 * pl.nl_acc + (d->compute_trg_disp ? 8 * (size_t)d->n_layers : 0));
  size_t bc = ((size_t)chunk_budget_mb() << 20) / (per_img ? per_img : 1);
  if (bc < 1) bc = 1;
  if (bc > (size_t)d->batch) bc = d->batch;
  if (bc > 65535) bc = 65535;
  const size_t n_chunks = ((size_t)d->batch + bc - 1) / bc;   // equal chunks: no short last launch
  bc = ((size_t)d->batch + n_chunks - 1) / n_chunks;
  pl.bc = (int)bc;
  pl.off_mats = 0;
  // guard band on both sides of the chunk accumulator: the streaming kernel adds exact zeros to the (clamped) cell of a
  // pixel that projects outside the image instead of branching around the reduction (render_stream.cuh)
  const size_t guard = align_up(((size_t)4 * d->w_t + 8) * 16, 256);
  pl.off_acc4 = align_up(((size_t)d->batch * 17 + 1) * sizeof(float), 256) + guard;   // matrices + rectified-class flags + their count
  pl.off_accd = pl.off_acc4 + align_up(n_trg * 16 * pl.nl_acc * bc, 256) + guard;
  pl.total = pl.off_accd + (d->compute_trg_disp ? align_up(n_trg * 8 * d->n_layers * bc, 256) : 0);
  return pl;
}

struct BwdPlan { int nl_g, bc; size_t off_mats, off_g4, off_gd, total; };

static BwdPlan plan_backward(const lsi_b200_splat_desc* d) {
  BwdPlan pl;
  const size_t n_trg = (size_t)d->h_t * d->w_t;
  // g4 is per layer when not composing or when the trg_disp gradient makes dL/dA_w layer dependent
  pl.nl_g = (d->compose_layers && !d->compute_trg_disp) ? 1 : d->n_layers;
  const size_t per_img = n_trg * (16 * pl.nl_g + (d->compute_trg_disp ? 4 * (size_t)d->n_layers : 0));
  size_t bc = ((size_t)chunk_budget_mb() << 20) / (per_img ? per_img : 1);
  if (bc < 1) bc = 1;
  if (bc > (size_t)d->batch) bc = d->batch;
  if (bc > 65535) bc = 65535;
  pl.bc = (int)bc;
  pl.off_mats = 0;
  pl.off_g4 = align_up((size_t)d->batch * 16 * sizeof(float), 256);
  pl.off_gd = pl.off_g4 + align_up(n_trg * 16 * pl.nl_g * bc, 256);
  pl.total = pl.off_gd + (d->compute_trg_disp ? align_up(n_trg * 4 * d->n_layers * bc, 256) : 0);
  return pl;
}

static int check_desc(const lsi_b200_splat_desc* d) {
  LSI_REQUIRE(d != nullptr, "descriptor is NULL");
  LSI_REQUIRE(d->n_layers >= 1 && d->n_layers <= 65535, "n_layers=%d out of range", d->n_layers);
  LSI_REQUIRE(d->batch >= 1, "batch=%d must be >= 1", d->batch);
  LSI_REQUIRE(d->h_s >= 1 && d->w_s >= 1 && d->h_t >= 1 && d->w_t >= 1, "image sizes must be >= 1");
  LSI_REQUIRE((long long)d->h_s * d->w_s < (1ll << 30) && (long long)d->h_t * d->w_t < (1ll << 30),
              "image too large for 32-bit pixel indices");
  LSI_REQUIRE(d->max_disp != 0.f, "max_disp must be non-zero");
  LSI_REQUIRE(d->tex_px_stride >= 3 && d->disp_px_stride >= 1 && d->mask_px_stride >= 1, "bad pixel strides");
  return LSI_B200_OK;
}

static GeomParams geom_of(const lsi_b200_splat_desc* d) {
  GeomParams gp;
  gp.w_t = d->w_t; gp.h_t = d->h_t; gp.ds = d->trg_downsampling;
  gp.inv_max_disp = 1.f / d->max_disp; gp.scale = d->zbuf_scale;
  return gp;
}

// zbuffer_weights on the scalar bg_layer_disp / max_disp (ldi.py:115), evaluated in fp32 like the reference
static float bg_weight(const lsi_b200_splat_desc* d) {
  float r = d->bg_layer_disp / d->max_disp;
  float c = fminf(fmaxf(r, 0.f), 1.f);
  return r > 0.f ? expf((c - 0.5f) * d->zbuf_scale) : 0.f;
}

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

static bool stream_enabled() {
  static int on = -1;
  if (on < 0) on = env_int("LSI_B200_SPLAT_STREAM", 1) != 0 ? 1 : 0;
  return on == 1;
}

// Persistent streaming splat (render_stream.cuh): grid = SMs x resident CTAs, each warp a contiguous unit range.
static int launch_stream(const FastParams& f, bool has_mask, bool packed, cudaStream_t st) {
  static int sms = 0, stages = 0, ctas_env = -1;
  if (!sms) {
    int dev = 0;
    LSI_CUDA(cudaGetDevice(&dev));
    LSI_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    stages = env_int("LSI_B200_STREAM_STAGES", kStreamQuad ? 3 : 2);
    if (stages < 2) stages = 2;
    if (stages > 4) stages = 4;
    ctas_env = env_int("LSI_B200_STREAM_CTAS", 0);
  }
  StreamParams p;
  p.f = f;
  p.segs = (f.W + kSegPx - 1) / kSegPx;
  p.groups = (f.L + 3) / 4;
  p.stages = stages;
  p.stage_bytes = 4 * kSegPx * 16 + (has_mask ? 4 * kSegPx * 4 : 0);
  p.units = (long long)f.bc * f.H * p.segs;
  LSI_REQUIRE(p.units * p.groups < (1ll << 31), "LDI chunk too large for the streaming splat kernel");
  const size_t smem = 128 + (size_t)kStreamWarps * stages * p.stage_bytes;
  int per_sm = (int)(232448 / (smem + 1024));
  if (per_sm > 8) per_sm = 8;
  if (ctas_env > 0 && ctas_env < per_sm) per_sm = ctas_env;
  long long grid = (long long)sms * per_sm;
  const long long need = (p.units + kStreamWarps - 1) / kStreamWarps;
  if (grid > need) grid = need;
  const int ki = (has_mask ? 2 : 0) | (packed ? 1 : 0);
  auto kern = ki == 0 ? splat_fwd_stream_kernel<false, false> : ki == 1 ? splat_fwd_stream_kernel<false, true>
            : ki == 2 ? splat_fwd_stream_kernel<true, false> : splat_fwd_stream_kernel<true, true>;
  static size_t smem_set[4] = {0, 0, 0, 0};
  if (smem > smem_set[ki]) {
    LSI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set[ki] = smem;
  }
  {
    ScopedTiming tm(kSplatFwd, st);
    kern<<<(unsigned)grid, kStreamWarps * 32, smem, st>>>(p);
  }
  LSI_LAUNCH_CHECK();
  return LSI_B200_OK;
}


static bool rowgather_enabled() {
  static int on = -1;
  if (on < 0) on = env_int("LSI_B200_ROWGATHER", 1) != 0 ? 1 : 0;
  return on == 1;
}

// Row-gather splat (render_rowgather.cuh): persistent CTAs, each a contiguous range of (image, output layer, target row)
// tasks.  Returns LSI_B200_OK with *launched = false when the shape does not fit the kernel (the caller falls back).
static int launch_rowgather(const lsi_b200_splat_desc* d, const float* tex, const float* disp, const float* mask,
                            const float* mats, const int* flags, float* img, float* wts, bool packed, bool all_flagged,
                            cudaStream_t st, bool* launched) {
  *launched = false;
  const int w_lists = d->w_s > d->w_t + 1 ? d->w_s : d->w_t + 1;                       // pixels per row / lists per row (w_t + 1)
  const int threads = ((w_lists + kRgPerThread - 1) / kRgPerThread + 31) / 32 * 32;   // consumer threads
  if (threads > kRgMaxThreads || d->w_s > 65534) return LSI_B200_OK;
  const int l_outer = d->compose_layers ? 1 : d->n_layers;
  if ((long long)l_outer * d->batch * d->h_t * d->w_t >= (1ll << 32)) return LSI_B200_OK;
  static int sms = 0, stages_env = 0;
  if (!sms) {
    int dev = 0;
    LSI_CUDA(cudaGetDevice(&dev));
    LSI_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    stages_env = env_int("LSI_B200_ROWGATHER_STAGES", 0);
  }
  RowGatherParams p;
  p.tex = tex; p.disp = disp; p.mask = mask; p.mats = mats; p.flags = flags; p.img = img; p.wts = wts;
  p.L = d->n_layers; p.B = d->batch; p.H = d->h_s; p.W = d->w_s; p.h_t = d->h_t; p.w_t = d->w_t;
  p.l_outer = l_outer;
  p.all_flagged = all_flagged ? 1 : 0;
  p.ds = d->trg_downsampling; p.inv_max_disp = 1.f / d->max_disp;
  p.k2 = d->zbuf_scale * 1.4426950408889634f; p.k2h = 0.5f * p.k2;
  p.nb = d->compose_layers ? (float)d->n_layers * bg_weight(d) : bg_weight(d);
  p.tasks = (long long)d->batch * p.l_outer * d->h_t;
  p.threads = threads;
  const bool small = threads + 32 <= 256;
  size_t off = 128;                                         // full / empty mbarriers (<= 8 stages)
  p.off_desc = (int)off; off += 8 * 32;                     // item descriptors
  const size_t wpad = (size_t)kRgPerThread * (threads + (packed ? 0 : 2));       // every per-node array: 4 entries per consumer thread (node stride threads + 2 for planar rows)
  p.off_head = (int)off; off = align_up(off + ((size_t)d->w_t + 34) * 4, 16);
  p.off_next = (int)off; off = align_up(off + wpad * 2, 16);
  p.off_wl = (int)off; off = align_up(off + wpad * 4, 16);
  p.off_wr = (int)off; off = align_up(off + wpad * 4, 16);
  p.off_bnd = (int)off; off += ((size_t)kRgPerThread * (threads / 32) + 1) * 16;
  p.off_val4 = (int)off; if (!packed) off += wpad * 16;
  off = align_up(off, 128);
  p.off_ring = (int)off;
  p.stage_bytes = (int)align_up(wpad * (16 + (mask ? 4 : 0)), 128);
  const int want_ctas = small ? LSI_RG_MIN_CTAS : 1;
  int stages = stages_env > 0 ? stages_env : 3;
  if (stages > 8) stages = 8;
  while (stages > 2 && (off + (size_t)stages * p.stage_bytes + 1024) * want_ctas > 232448) --stages;
  const size_t smem = off + (size_t)stages * p.stage_bytes;
  if (smem > 232448 - 1024) return LSI_B200_OK;            // row too wide for one CTA's shared memory
  p.stages = stages;
  int per_sm = (int)(232448 / (smem + 1024));
  if (per_sm > want_ctas) per_sm = want_ctas;
  long long grid = (long long)sms * per_sm;
  if (grid > p.tasks) grid = p.tasks;
  // instantiations: consumer threads as a compile-time constant for the two BASELINE widths (832 -> 224, 1664 -> 448), run-time otherwise
  const int tsel = threads == 224 ? 0 : threads == 448 ? 2 : small ? 1 : 3;
  const int ki = tsel * 4 + ((mask ? 2 : 0) | (packed ? 1 : 0));
  void (*kern)(const RowGatherParams) = nullptr;
#define LSI_RG_CASE(i, cta, kt, m, pk) case i: kern = splat_fwd_rowgather_kernel<cta, kt, m, pk>; break;
  switch (ki) {
    LSI_RG_CASE(0, 256, 224, false, false) LSI_RG_CASE(1, 256, 224, false, true) LSI_RG_CASE(2, 256, 224, true, false) LSI_RG_CASE(3, 256, 224, true, true)
    LSI_RG_CASE(4, 256, 0, false, false) LSI_RG_CASE(5, 256, 0, false, true) LSI_RG_CASE(6, 256, 0, true, false) LSI_RG_CASE(7, 256, 0, true, true)
    LSI_RG_CASE(8, 544, 448, false, false) LSI_RG_CASE(9, 544, 448, false, true) LSI_RG_CASE(10, 544, 448, true, false) LSI_RG_CASE(11, 544, 448, true, true)
    LSI_RG_CASE(12, 544, 0, false, false) LSI_RG_CASE(13, 544, 0, false, true) LSI_RG_CASE(14, 544, 0, true, false)
    default: kern = splat_fwd_rowgather_kernel<544, 0, true, true>; break;
  }
#undef LSI_RG_CASE
  static size_t smem_set[16] = {0};
  if (smem > smem_set[ki]) {
    LSI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set[ki] = smem;
  }
  {
    ScopedTiming tm(kSplatFwd, st);
    kern<<<(unsigned)grid, threads + 32, smem, st>>>(p);
  }
  LSI_LAUNCH_CHECK();
  *launched = true;
  return LSI_B200_OK;
}

template <typename K, typename P>
static void launch3(K kern, dim3 grid, dim3 block, cudaStream_t st, const P& p) { kern<<<grid, block, 0, st>>>(p); }

}  // namespace lsi

using namespace lsi;

extern "C" int lsi_b200_projection_matrix(const float* k_s, const float* k_t, const float* rot, const float* t,
                                          int batch, int inverse, float* out, void* stream) {
  LSI_REQUIRE(k_s && k_t && rot && t && out, "NULL pointer argument");
  LSI_REQUIRE(batch >= 1, "batch=%d must be >= 1", batch);
  proj_matrix_kernel<<<(batch + 63) / 64, 64, 0, as_stream(stream)>>>(k_s, k_t, rot, t, batch, inverse ? 1 : 0, out, nullptr);
  LSI_LAUNCH_CHECK();
  return LSI_B200_OK;
}

extern "C" size_t lsi_b200_forward_splat_workspace_bytes(const lsi_b200_splat_desc* d) {
  if (check_desc(d) != LSI_B200_OK) return 0;
  return plan_forward(d).total;
}

extern "C" size_t lsi_b200_forward_splat_backward_workspace_bytes(const lsi_b200_splat_desc* d) {
  if (check_desc(d) != LSI_B200_OK) return 0;
  return plan_backward(d).total;
}

extern "C" int lsi_b200_forward_splat(const lsi_b200_splat_desc* d, const float* tex, const float* mask,
                                      const float* disp, const float* pixel_coords, const float* k_s,
                                      const float* k_t, const float* rot, const float* t, const float* focal_disps,
                                      float* trg_img, float* trg_wts, float* trg_disp, float* layer_acc,
                                      void* workspace, size_t workspace_bytes, void* stream) {
  if (int rc = check_desc(d)) return rc;
  LSI_REQUIRE(tex && disp && k_s && k_t && rot && t && trg_img && trg_wts, "NULL pointer argument");
  LSI_REQUIRE(!d->compute_trg_disp || trg_disp, "compute_trg_disp needs trg_disp");
  const FwdPlan pl = plan_forward(d);
  LSI_REQUIRE(workspace && workspace_bytes >= pl.total, "workspace too small: %zu < %zu", workspace_bytes, pl.total);
  LSI_REQUIRE(((uintptr_t)workspace & 255) == 0, "workspace must be 256-byte aligned");
  cudaStream_t st = as_stream(stream);
  char* ws = static_cast<char*>(workspace);
  float* mats = reinterpret_cast<float*>(ws + pl.off_mats);
  int* rect_flags = reinterpret_cast<int*>(mats + (size_t)d->batch * 16);
  LSI_CUDA(cudaMemsetAsync(rect_flags + d->batch, 0, sizeof(int), st));
  proj_matrix_kernel<<<(d->batch + 63) / 64, 64, 0, st>>>(k_s, k_t, rot, t, d->batch, 0, mats, rect_flags);
  LSI_LAUNCH_CHECK();

  const int n_src = d->h_s * d->w_s, n_trg = d->h_t * d->w_t;
  const bool has_disp = d->compute_trg_disp != 0;
  const bool own_accd = has_disp && layer_acc == nullptr;
  if (has_disp && layer_acc)
    LSI_CUDA(cudaMemsetAsync(layer_acc, 0, (size_t)d->n_layers * d->batch * n_trg * 8, st));

  const bool bc_fits_grid = pl.bc <= 65535;
  // fast path: standard grid, no focal shift, no trg_disp, 16-byte aligned rows, planar (3/1/1) or packed (4/4) layout
  const bool packed = d->tex_px_stride == 4 && d->disp_px_stride == 4 && disp == tex + 3;
  const bool planar = d->tex_px_stride == 3 && d->disp_px_stride == 1;
  const bool fast = (d->variant == 0 || d->variant == 2 || d->variant == 3 || d->variant == 5 || d->variant == 6 || d->variant >= 100) && !pixel_coords && !focal_disps && !has_disp && (packed || planar) &&
                    (!mask || d->mask_px_stride == 1) && (!packed || ((uintptr_t)tex & 15) == 0) && d->h_s <= 65535 &&
                    bc_fits_grid;
  // rectified-stereo pose class: warp-owned target rows in shared memory, written once (render_rowowner.cuh); images of
  // any other class fall through to the reduction kernels below, which skip the flagged ones
  const int* skip = nullptr;
  const bool fast_eligible = fast;
  // streaming kernel (render_stream.cuh): bulk copies need 16-byte aligned row segments
  const bool use_stream_shape = fast && ((uintptr_t)tex & 15) == 0 &&
                      (packed || (d->w_s % 4 == 0 && ((uintptr_t)disp & 15) == 0)) &&
                      (!mask || (d->w_s % 4 == 0 && ((uintptr_t)mask & 15) == 0));
  const bool use_stream = use_stream_shape && (d->variant == 0 || d->variant == 5 || d->variant == 6) && !rowowner_enabled() && stream_enabled();
  // rectified pose class, default: row-gather kernel (render_rowgather.cuh) -- target rows owned by CTAs, the scatter inverted
  // in shared memory, normalisation fused.  variant 5 = the caller knows (from these flags, read back earlier for the same
  // camera tensors) that every image is in the class: the reduction kernels below are not launched at all (a hint: shapes the
  // kernel does not take run the default path); variant 6 = the caller knows that NO image is in the class: no row-gather launch.
  const bool gather_ok = use_stream_shape && (d->variant == 0 || d->variant == 5) && rowgather_enabled() && !rowowner_enabled();
  if (gather_ok) {
    bool launched = false;
    if (int rc = launch_rowgather(d, tex, disp, mask, mats, rect_flags, trg_img, trg_wts, packed, d->variant == 5, st, &launched))
      return rc;
    if (launched) {
      if (d->variant == 5) return LSI_B200_OK;
      skip = rect_flags;
    }
  }
  if (!skip && fast_eligible && (d->variant == 2 || rowowner_enabled())) {
    int R = (int)(14336 / ((size_t)d->w_t * 16));
    if (R > 4) R = 4;
    if (R > d->h_t) R = d->h_t;
    if (R >= 1) {
      RowOwnerParams r;
      r.tex = tex; r.disp = disp; r.mask = mask; r.mats = mats; r.flags = rect_flags; r.img = trg_img; r.wts = trg_wts;
      r.L = d->n_layers; r.B = d->batch; r.H = d->h_s; r.W = d->w_s; r.h_t = d->h_t; r.w_t = d->w_t; r.R = R;
      r.bands = (d->h_t + R - 1) / R; r.compose = d->compose_layers ? 1 : 0;
      r.ds = d->trg_downsampling; r.inv_max_disp = 1.f / d->max_disp;
      r.k2 = d->zbuf_scale * 1.4426950408889634f; r.k2h = 0.5f * r.k2; r.bg_wt = bg_weight(d);
      const size_t smem = (size_t)4 * R * d->w_t * 16;
      static size_t smem_set[4] = {0, 0, 0, 0};
      const int ki = (mask ? 2 : 0) | (packed ? 1 : 0);
      auto kern = ki == 0 ? splat_rowowner_kernel<false, false> : ki == 1 ? splat_rowowner_kernel<false, true>
                : ki == 2 ? splat_rowowner_kernel<true, false> : splat_rowowner_kernel<true, true>;
      if (smem > smem_set[ki]) {
        LSI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set[ki] = smem;
      }
      const long long tasks = (long long)d->batch * r.bands;
      {
        ScopedTiming tm(kSplatFwd, st);
        kern<<<(unsigned)((tasks + 3) / 4), 128, smem, st>>>(r);
      }
      LSI_LAUNCH_CHECK();
      skip = rect_flags;
    }
  }
  for (int b0 = 0; b0 < d->batch; b0 += pl.bc) {
    const int bc = (d->batch - b0 < pl.bc) ? d->batch - b0 : pl.bc;
    FwdParams p;
    p.tex = tex; p.mask = mask; p.disp = disp; p.pc = pixel_coords; p.focal = focal_disps; p.mats = mats;
    p.acc4 = reinterpret_cast<float4*>(ws + pl.off_acc4);
    p.accd = has_disp ? (own_accd ? reinterpret_cast<float2*>(ws + pl.off_accd) : reinterpret_cast<float2*>(layer_acc))
                      : nullptr;
    p.L = d->n_layers; p.B = d->batch; p.H = d->h_s; p.W = d->w_s; p.b0 = b0; p.bc = bc;
    p.tex_s = d->tex_px_stride; p.disp_s = d->disp_px_stride; p.mask_s = d->mask_px_stride;
    p.acc_per_layer = d->compose_layers ? 0 : 1;
    p.accd_b = own_accd ? bc : d->batch; p.accd_b0 = own_accd ? b0 : 0;
    p.gp = geom_of(d);
    if (skip) {
      zero_acc_kernel<<<dim3(64, bc), 256, 0, st>>>(p.acc4, pl.nl_acc, bc, n_trg, b0, skip);
      LSI_LAUNCH_CHECK();
    } else {
      LSI_CUDA(cudaMemsetAsync(p.acc4, 0, (size_t)pl.nl_acc * bc * n_trg * 16, st));
    }
    if (own_accd) LSI_CUDA(cudaMemsetAsync(p.accd, 0, (size_t)d->n_layers * bc * n_trg * 8, st));
    if (fast) {
      FastParams f;
      f.tex = tex; f.disp = disp; f.mask = mask; f.mats = mats; f.acc4 = p.acc4;
      f.L = d->n_layers; f.B = d->batch; f.H = d->h_s; f.W = d->w_s; f.b0 = b0; f.bc = bc;
      f.acc_per_layer = p.acc_per_layer; f.w_t = d->w_t; f.h_t = d->h_t;
      f.ds = d->trg_downsampling; f.inv_max_disp = 1.f / d->max_disp;
      f.k2 = d->zbuf_scale * 1.4426950408889634f; f.k2h = 0.5f * f.k2;
      f.ablate = (d->variant >= 100) ? d->variant - 100 : 0;
      f.skip = skip; f.n_flagged = skip ? rect_flags + d->batch : nullptr;
      dim3 fgrid((d->w_s + 63) / 64, d->h_s, bc), fblock(64);
      if (use_stream) {
        if (int rc = launch_stream(f, mask != nullptr, packed, st)) return rc;
      } else {
        ScopedTiming tm(kSplatFwd, st);
        if (packed) {
          if (mask) launch3(splat_fwd_fast_kernel<true, true>, fgrid, fblock, st, f);
          else launch3(splat_fwd_fast_kernel<false, true>, fgrid, fblock, st, f);
        } else {
          if (mask) launch3(splat_fwd_fast_kernel<true, false>, fgrid, fblock, st, f);
          else launch3(splat_fwd_fast_kernel<false, false>, fgrid, fblock, st, f);
        }
      }
      LSI_LAUNCH_CHECK();
      if (n_trg % 4 == 0 && ((uintptr_t)trg_img & 15) == 0 && ((uintptr_t)trg_wts & 15) == 0) {
        NormFastParams nf;
        nf.acc4 = p.acc4; nf.img = trg_img; nf.wts = trg_wts; nf.B = d->batch; nf.b0 = b0; nf.bc = bc; nf.n_trg = n_trg;
        nf.nb = d->compose_layers ? (float)d->n_layers * bg_weight(d) : bg_weight(d);
        nf.skip = skip;
        {
          ScopedTiming tm(kNormalize, st);
          normalize_fast_kernel<<<dim3((n_trg / 4 + 255) / 256, bc, pl.nl_acc), 256, 0, st>>>(nf);
        }
        LSI_LAUNCH_CHECK();
        continue;
      }
    } else {
    dim3 grid((n_src + 255) / 256, d->n_layers, bc), block(256);
    const int sel = (mask ? 4 : 0) | (pixel_coords ? 2 : 0) | (has_disp ? 1 : 0);
    {
    ScopedTiming tm(kSplatFwd, st);
    switch (sel) {
      case 0: launch3(splat_fwd_atomic_kernel<false, false, false>, grid, block, st, p); break;
      case 1: launch3(splat_fwd_atomic_kernel<false, false, true>, grid, block, st, p); break;
      case 2: launch3(splat_fwd_atomic_kernel<false, true, false>, grid, block, st, p); break;
      case 3: launch3(splat_fwd_atomic_kernel<false, true, true>, grid, block, st, p); break;
      case 4: launch3(splat_fwd_atomic_kernel<true, false, false>, grid, block, st, p); break;
      case 5: launch3(splat_fwd_atomic_kernel<true, false, true>, grid, block, st, p); break;
      case 6: launch3(splat_fwd_atomic_kernel<true, true, false>, grid, block, st, p); break;
      default: launch3(splat_fwd_atomic_kernel<true, true, true>, grid, block, st, p); break;
    }
    }
    LSI_LAUNCH_CHECK();
    }
    NormParams np;
    np.acc4 = p.acc4; np.accd = p.accd; np.img = trg_img; np.wts = trg_wts; np.disp = has_disp ? trg_disp : nullptr;
    np.L = d->n_layers; np.B = d->batch; np.b0 = b0; np.bc = bc; np.n_trg = n_trg; np.compose = d->compose_layers ? 1 : 0;
    np.accd_b = p.accd_b; np.accd_b0 = p.accd_b0; np.bg_wt = bg_weight(d); np.skip = skip;
    {
      ScopedTiming tm(kNormalize, st);
      normalize_kernel<<<dim3((n_trg + 255) / 256, bc, pl.nl_acc), 256, 0, st>>>(np);
    }
    LSI_LAUNCH_CHECK();
  }
  return LSI_B200_OK;
}

extern "C" int lsi_b200_forward_splat_backward(const lsi_b200_splat_desc* d, const float* tex, const float* mask,
                                               const float* disp, const float* pixel_coords, const float* k_s,
                                               const float* k_t, const float* rot, const float* t,
                                               const float* focal_disps, const float* trg_img, const float* trg_wts,
                                               const float* layer_acc, const float* g_img, const float* g_wts,
                                               const float* g_disp, float* d_tex, float* d_mask, float* d_disp,
                                               void* workspace, size_t workspace_bytes, void* stream) {
  if (int rc = check_desc(d)) return rc;
  LSI_REQUIRE(tex && disp && k_s && k_t && rot && t && trg_img && trg_wts && d_tex && d_disp, "NULL pointer argument");
  LSI_REQUIRE(d->tex_px_stride == 3 && d->disp_px_stride == 1 && d->mask_px_stride == 1,
              "backward expects the reference's dense tex/disp/mask layouts");
  const bool has_gd = d->compute_trg_disp != 0;
  LSI_REQUIRE(!has_gd || layer_acc, "the gradient of trg_disp needs layer_acc from the forward");
  const BwdPlan pl = plan_backward(d);
  LSI_REQUIRE(workspace && workspace_bytes >= pl.total, "workspace too small: %zu < %zu", workspace_bytes, pl.total);
  LSI_REQUIRE(((uintptr_t)workspace & 255) == 0, "workspace must be 256-byte aligned");
  cudaStream_t st = as_stream(stream);
  char* ws = static_cast<char*>(workspace);
  float* mats = reinterpret_cast<float*>(ws + pl.off_mats);
  proj_matrix_kernel<<<(d->batch + 63) / 64, 64, 0, st>>>(k_s, k_t, rot, t, d->batch, 0, mats, nullptr);
  LSI_LAUNCH_CHECK();
  const int n_src = d->h_s * d->w_s, n_trg = d->h_t * d->w_t;
  for (int b0 = 0; b0 < d->batch; b0 += pl.bc) {
    const int bc = (d->batch - b0 < pl.bc) ? d->batch - b0 : pl.bc;
    BwdAParams a;
    a.img = trg_img; a.wts = trg_wts; a.accd = reinterpret_cast<const float2*>(layer_acc);
    a.g_img = g_img; a.g_wts = g_wts; a.g_disp = g_disp;
    a.g4 = reinterpret_cast<float4*>(ws + pl.off_g4);
    a.gd = has_gd ? reinterpret_cast<float*>(ws + pl.off_gd) : nullptr;
    a.L = d->n_layers; a.B = d->batch; a.b0 = b0; a.bc = bc; a.n_trg = n_trg; a.compose = d->compose_layers ? 1 : 0;
    a.nl_g = pl.nl_g; a.bg_wt = bg_weight(d);
    {
      ScopedTiming tm(kBwdTarget, st);
      splat_bwd_target_kernel<<<dim3((n_trg + 255) / 256, bc), 256, 0, st>>>(a);
    }
    LSI_LAUNCH_CHECK();
    BwdBParams p;
    p.tex = tex; p.mask = mask; p.disp = disp; p.pc = pixel_coords; p.focal = focal_disps; p.mats = mats;
    p.g4 = a.g4; p.gd = a.gd; p.d_tex = d_tex; p.d_mask = d_mask; p.d_disp = d_disp;
    p.L = d->n_layers; p.B = d->batch; p.H = d->h_s; p.W = d->w_s; p.b0 = b0; p.bc = bc;
    p.tex_s = 3; p.disp_s = 1; p.mask_s = 1;
    p.g_per_layer = pl.nl_g > 1 ? 1 : 0;
    p.gp = geom_of(d);
    dim3 grid((n_src + 255) / 256, d->n_layers, bc), block(256);
    const int sel = (mask ? 4 : 0) | (pixel_coords ? 2 : 0) | (has_gd ? 1 : 0);
    {
    ScopedTiming tm(kBwdSource, st);
    switch (sel) {
      case 0: launch3(splat_bwd_source_kernel<false, false, false>, grid, block, st, p); break;
      case 1: launch3(splat_bwd_source_kernel<false, false, true>, grid, block, st, p); break;
      case 2: launch3(splat_bwd_source_kernel<false, true, false>, grid, block, st, p); break;
      case 3: launch3(splat_bwd_source_kernel<false, true, true>, grid, block, st, p); break;
      case 4: launch3(splat_bwd_source_kernel<true, false, false>, grid, block, st, p); break;
      case 5: launch3(splat_bwd_source_kernel<true, false, true>, grid, block, st, p); break;
      case 6: launch3(splat_bwd_source_kernel<true, true, false>, grid, block, st, p); break;
      default: launch3(splat_bwd_source_kernel<true, true, true>, grid, block, st, p); break;
    }
    }
    LSI_LAUNCH_CHECK();
  }
  return LSI_B200_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Host-buffer entry point (reference-facing end-to-end path).  Device staging is cached across calls.
// ---------------------------------------------------------------------------------------------------------
namespace {
struct HostCtx {
  void* buf = nullptr; size_t cap = 0; cudaStream_t st = nullptr;
};
HostCtx g_host;
}  // namespace

extern "C" int lsi_b200_forward_splat_host(const lsi_b200_splat_desc* d, const float* tex_host,
                                           const float* mask_host, const float* disp_host, const float* k_s_host,
                                           const float* k_t_host, const float* rot_host, const float* t_host,
                                           float* trg_img_host, float* trg_wts_host, float* trg_disp_host) {
  if (int rc = check_desc(d)) return rc;
  LSI_REQUIRE(tex_host && disp_host && k_s_host && k_t_host && rot_host && t_host && trg_img_host && trg_wts_host,
              "NULL pointer argument");
  LSI_REQUIRE(d->tex_px_stride == 3 && d->disp_px_stride == 1 && d->mask_px_stride == 1, "host path expects dense layouts");
  LSI_REQUIRE(!d->compute_trg_disp || trg_disp_host, "compute_trg_disp needs trg_disp_host");
  const size_t n_src = (size_t)d->h_s * d->w_s, n_trg = (size_t)d->h_t * d->w_t;
  const size_t LB = (size_t)d->n_layers * d->batch;
  const size_t nl_out = d->compose_layers ? 1 : d->n_layers;
  const size_t b_tex = align_up(LB * n_src * 12, 256), b_one = align_up(LB * n_src * 4, 256);
  const size_t b_cam = align_up((size_t)d->batch * 9 * 4, 256);
  const size_t b_img = align_up(nl_out * d->batch * n_trg * 12, 256), b_w = align_up(nl_out * d->batch * n_trg * 4, 256);
  const size_t ws_bytes = plan_forward(d).total;
  const size_t total = b_tex + 2 * b_one + 4 * b_cam + b_img + 2 * b_w + ws_bytes;
  if (!g_host.st) LSI_CUDA(cudaStreamCreateWithFlags(&g_host.st, cudaStreamNonBlocking));
  if (g_host.cap < total) {
    if (g_host.buf) LSI_CUDA(cudaFree(g_host.buf));
    g_host.buf = nullptr; g_host.cap = 0;
    LSI_CUDA(cudaMalloc(&g_host.buf, total));
    g_host.cap = total;
  }
  char* p = static_cast<char*>(g_host.buf);
  float* tex = (float*)p; p += b_tex;
  float* disp = (float*)p; p += b_one;
  float* mask = (float*)p; p += b_one;
  float* ks = (float*)p; p += b_cam;
  float* kt = (float*)p; p += b_cam;
  float* rot = (float*)p; p += b_cam;
  float* tt = (float*)p; p += b_cam;
  float* img = (float*)p; p += b_img;
  float* wts = (float*)p; p += b_w;
  float* dsp = (float*)p; p += b_w;
  void* ws = p;
  cudaStream_t st = g_host.st;
  LSI_CUDA(cudaMemcpyAsync(tex, tex_host, LB * n_src * 12, cudaMemcpyHostToDevice, st));
  LSI_CUDA(cudaMemcpyAsync(disp, disp_host, LB * n_src * 4, cudaMemcpyHostToDevice, st));
  if (mask_host) LSI_CUDA(cudaMemcpyAsync(mask, mask_host, LB * n_src * 4, cudaMemcpyHostToDevice, st));
  LSI_CUDA(cudaMemcpyAsync(ks, k_s_host, (size_t)d->batch * 36, cudaMemcpyHostToDevice, st));
  LSI_CUDA(cudaMemcpyAsync(kt, k_t_host, (size_t)d->batch * 36, cudaMemcpyHostToDevice, st));
  LSI_CUDA(cudaMemcpyAsync(rot, rot_host, (size_t)d->batch * 36, cudaMemcpyHostToDevice, st));
  LSI_CUDA(cudaMemcpyAsync(tt, t_host, (size_t)d->batch * 12, cudaMemcpyHostToDevice, st));
  if (int rc = lsi_b200_forward_splat(d, tex, mask_host ? mask : nullptr, disp, nullptr, ks, kt, rot, tt, nullptr, img,
                                      wts, d->compute_trg_disp ? dsp : nullptr, nullptr, ws, ws_bytes, st))
    return rc;
  LSI_CUDA(cudaMemcpyAsync(trg_img_host, img, nl_out * d->batch * n_trg * 12, cudaMemcpyDeviceToHost, st));
  LSI_CUDA(cudaMemcpyAsync(trg_wts_host, wts, nl_out * d->batch * n_trg * 4, cudaMemcpyDeviceToHost, st));
  if (d->compute_trg_disp)
    LSI_CUDA(cudaMemcpyAsync(trg_disp_host, dsp, nl_out * d->batch * n_trg * 4, cudaMemcpyDeviceToHost, st));
  LSI_CUDA(cudaStreamSynchronize(st));
  return LSI_B200_OK;
}
