// GPU-native synthetic planar-room generator: the reference renders a box world of textured planes with a
// TF graph per view -- per plane an inverse-homography warp of its texture and of its mask (homography.transform_plane_imgs,
// lsi/geometry/homography.py:95-118), a per-pixel disparity map (trg_disp_maps, :140-156), then hard soft-z selection over the
// planes plus a white background layer (layers.compose / compose_depth, lsi/geometry/layers.py:29-118) -- in a second TF session,
// one sample at a time (lsi/data/syntheticPlanes/data.py:372-420, 598-619).  Here one fused kernel renders a whole batch of
// scenes: a thread per target pixel walks the planes (homography, bilinear texture + mask sample, plane disparity, log
// layer probability), keeps the running arg-max for the foreground and for the background selection, and writes the image
// and both disparity maps; the warped layers never exist in memory.  Textures are procedural (the PASCAL / SUN images the
// reference pastes on the planes are not available): sums of sinusoids per channel, objects with a super-ellipse alpha mask.
#include "capi_common.h"
#include "common.cuh"

namespace lsi {

constexpr int kMaxPlanes = 16;

struct PlanesParams {
  const float* imgs; const float* masks;   // [B][n][h][w][3], [B][n][h][w][1]
  const float* hom; const float* dmat;     // [B][n][9] target pixel -> texture pixel, [B][n][3] target pixel -> disparity
  const float* gmax;                       // [B] max over the scene of relu(disparity) and min_disp (layers.py:103)
  float* out_img; float* out_fg; float* out_bg;
  int B, n, h, w, H, W;
  float min_disp, inv_temp;
};

__device__ __forceinline__ float plane_disp(const float* m, float u, float v) {
  return fmaxf((m[0] * u + m[1] * v) + m[2], 0.f);                      // trg_disp_maps + relu (layers.py:50)
}

// log layer probability of helpers.soft_z_buffering (helpers.py:140-160) before its (arg-max preserving) normalisation
__device__ __forceinline__ float layer_logp(float mask, float disp, float inv_temp) {
  const float depth = 1.f / safe_den(disp);
  return logf(mask + 1e-8f) - depth * inv_temp;
}

__global__ void __launch_bounds__(64) planes_gmax_kernel(const float* __restrict__ dmat, int B, int n, int H, int W, float min_disp,
                                                         float* __restrict__ gmax) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float m = min_disp;
  const float us[2] = {0.5f, (float)W - 0.5f}, vs[2] = {0.5f, (float)H - 0.5f};
  for (int l = 0; l < n; ++l)            // a plane's disparity is affine in (u, v): its maximum over the pixel grid sits at a corner
    for (int c = 0; c < 4; ++c) m = fmaxf(m, plane_disp(dmat + ((size_t)b * n + l) * 3, us[c & 1], vs[c >> 1]));
  gmax[b] = m;
}

__global__ void __launch_bounds__(128) render_planes_kernel(const PlanesParams p) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per = (long long)p.H * p.W;
  if (q >= per * p.B) return;
  const int b = (int)(q / per);
  const int r = (int)(q - (long long)b * per);
  const float u = (float)(r % p.W) + 0.5f, v = (float)(r / p.W) + 0.5f;
  const float gmax = p.gmax[b];
  // background layer (layers.py:52-58): white, opaque, disparity min_disp; listed last, so a tie goes to a plane
  float best_fg = -INFINITY, best_bg = -INFINITY;
  float fg_r = 1.f, fg_g = 1.f, fg_b = 1.f, fg_d = p.min_disp, bg_d = p.min_disp;
  const size_t tex_px = (size_t)p.h * p.w;
  for (int l = 0; l < p.n; ++l) {
    const float* Hm = p.hom + ((size_t)b * p.n + l) * 9;
    const float xs = (Hm[0] * u + Hm[1] * v) + Hm[2], ys = (Hm[3] * u + Hm[4] * v) + Hm[5], ns = (Hm[6] * u + Hm[7] * v) + Hm[8];
    const float nd = safe_den(ns);
    PixGeom g;
    corners(xs / nd - 0.5f, ys / nd - 0.5f, p.w, p.h, g);      // sampling.bilinear (sampling.py:41-132): weights not thresholded
    const float wt[4] = {g.wx0 * g.wy0, g.wx1 * g.wy0, g.wx0 * g.wy1, g.wx1 * g.wy1};
    const float* im = p.imgs + ((size_t)b * p.n + l) * tex_px * 3;
    const float* mk = p.masks + ((size_t)b * p.n + l) * tex_px;
    float cr = 0.f, cg = 0.f, cb = 0.f, cm = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (wt[c] != 0.f) {
        const size_t ci = (size_t)corner_index(g, c, p.w);
        cr = fmaf(wt[c], __ldg(im + ci * 3), cr); cg = fmaf(wt[c], __ldg(im + ci * 3 + 1), cg); cb = fmaf(wt[c], __ldg(im + ci * 3 + 2), cb);
        cm = fmaf(wt[c], __ldg(mk + ci), cm);
      }
    }
    const float d = plane_disp(p.dmat + ((size_t)b * p.n + l) * 3, u, v);
    const float lp = layer_logp(cm, d, p.inv_temp);
    if (lp > best_fg) { best_fg = lp; fg_r = cr; fg_g = cg; fg_b = cb; fg_d = d; }
    if (p.out_bg) {
      const float lpb = layer_logp(cm, fmaxf(gmax - d, 0.f), p.inv_temp);      // layers.py:103-107; soft_z_buffering relu's again
      if (lpb > best_bg) { best_bg = lpb; bg_d = d; }
    }
  }
  const float lp_bgl = layer_logp(1.f, p.min_disp, p.inv_temp);
  if (lp_bgl > best_fg) { fg_r = 1.f; fg_g = 1.f; fg_b = 1.f; fg_d = p.min_disp; }
  float* o = p.out_img + (size_t)q * 3;
  o[0] = fg_r; o[1] = fg_g; o[2] = fg_b;
  if (p.out_fg) p.out_fg[q] = fg_d;
  if (p.out_bg) p.out_bg[q] = (lp_bgl > best_bg) ? p.min_disp : bg_d;
}

// procedural textures: img[n][h][w][3] = 0.5 + sum_k amp_k sin(fx_k x + fy_k y + ph_k) per channel (clamped to [0,1]);
// mask = 1 (walls) or a super-ellipse alpha with a soft edge (objects: kind != 0).  params [n][3][K][4] = (amp, fx, fy, phase).
__global__ void __launch_bounds__(256) procedural_texture_kernel(const float* __restrict__ params, const int* __restrict__ kind, int n, int K,
                                                                 int h, int w, float* __restrict__ img, float* __restrict__ mask) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per = (long long)h * w;
  if (q >= per * n) return;
  const int l = (int)(q / per);
  const int r = (int)(q - (long long)l * per);
  const float x = (float)(r % w) + 0.5f, y = (float)(r / w) + 0.5f;
  for (int c = 0; c < 3; ++c) {
    float acc = 0.5f;
    const float* pr = params + ((size_t)(l * 3 + c) * K) * 4;
    for (int k = 0; k < K; ++k) acc = fmaf(pr[4 * k], sinf(pr[4 * k + 1] * x + pr[4 * k + 2] * y + pr[4 * k + 3]), acc);
    img[(size_t)q * 3 + c] = fminf(fmaxf(acc, 0.f), 1.f);
  }
  float m = 1.f;
  if (kind[l] != 0) {
    const float ex = (2.f * x / (float)w - 1.f) / 0.9f, ey = (2.f * y / (float)h - 1.f) / 0.9f;
    const float rr = ex * ex * ex * ex + ey * ey * ey * ey;                 // super-ellipse |x|^4 + |y|^4 <= 1
    m = fminf(fmaxf((1.f - rr) * 8.f, 0.f), 1.f);
  }
  mask[q] = m;
}

// tf.image.resize_images(AREA) by an integer factor (data.py:364-368,...): box mean, float NHWC
__global__ void __launch_bounds__(256) box_downsample_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int H, int W, int C,
                                                             int f) {
  const int Ho = H / f, Wo = W / f;
  const long long total = (long long)B * Ho * Wo * C;
  const float inv = 1.f / (float)(f * f);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long t = i / C;
    const int x = (int)(t % Wo); t /= Wo;
    const int y = (int)(t % Ho);
    const int b = (int)(t / Ho);
    float acc = 0.f;
    for (int dy = 0; dy < f; ++dy)
      for (int dx = 0; dx < f; ++dx) acc += in[(((size_t)b * H + y * f + dy) * W + x * f + dx) * C + c];
    out[i] = acc * inv;
  }
}

}  // namespace lsi

using namespace lsi;

extern "C" int lsi_b200_render_planes(const float* imgs_w, const float* masks_w, const float* hom_t2w, const float* dmat_t, int batch,
                                      int n_planes, int h_tex, int w_tex, int h_out, int w_out, float min_disp, float depth_softmax_temp,
                                      float* out_img, float* out_disp_fg, float* out_disp_bg, float* scratch_gmax, void* stream) {
  LSI_REQUIRE(imgs_w && masks_w && hom_t2w && dmat_t && out_img && scratch_gmax, "NULL pointer argument");
  LSI_REQUIRE(batch >= 1 && n_planes >= 1 && n_planes <= kMaxPlanes, "n_planes=%d out of range (1..%d)", n_planes, kMaxPlanes);
  LSI_REQUIRE(h_tex >= 1 && w_tex >= 1 && h_out >= 1 && w_out >= 1 && depth_softmax_temp > 0.f, "bad sizes");
  LSI_REQUIRE((long long)h_tex * w_tex < (1ll << 30), "texture too large");
  cudaStream_t st = as_stream(stream);
  planes_gmax_kernel<<<(batch + 63) / 64, 64, 0, st>>>(dmat_t, batch, n_planes, h_out, w_out, min_disp, scratch_gmax);
  LSI_LAUNCH_CHECK();
  PlanesParams p{imgs_w, masks_w, hom_t2w, dmat_t, scratch_gmax, out_img, out_disp_fg, out_disp_bg, batch, n_planes, h_tex, w_tex,
                 h_out, w_out, min_disp, 1.f / depth_softmax_temp};
  const long long total = (long long)batch * h_out * w_out;
  render_planes_kernel<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(p);
  LSI_LAUNCH_CHECK();
  return LSI_B200_OK;
}

extern "C" int lsi_b200_procedural_texture(const float* params, const int* kind, int n_textures, int n_waves, int h, int w, float* img,
                                           float* mask, void* stream) {
  LSI_REQUIRE(params && kind && img && mask, "NULL pointer argument");
  LSI_REQUIRE(n_textures >= 1 && n_waves >= 1 && h >= 1 && w >= 1, "bad sizes");
  const long long total = (long long)n_textures * h * w;
  procedural_texture_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(params, kind, n_textures, n_waves, h, w, img, mask);
  LSI_LAUNCH_CHECK();
  return LSI_B200_OK;
}

extern "C" int lsi_b200_box_downsample(const float* in, float* out, int batch, int h, int w, int channels, int factor, void* stream) {
  LSI_REQUIRE(in && out, "NULL pointer argument");
  LSI_REQUIRE(batch >= 1 && channels >= 1 && factor >= 1 && h >= factor && w >= factor && h % factor == 0 && w % factor == 0,
              "sizes must be multiples of the factor");
  const long long total = (long long)batch * (h / factor) * (w / factor) * channels;
  long long grid = (total + 255) / 256;
  if (grid > 148 * 16) grid = 148 * 16;
  box_downsample_kernel<<<(unsigned)grid, 256, 0, as_stream(stream)>>>(in, out, batch, h, w, channels, factor);
  LSI_LAUNCH_CHECK();
  return LSI_B200_OK;
}
