// View-synthesis loss kernels (forward + backward), fused per pixel with deterministic two-stage reductions.
//   zbuffer_composition_loss   lsi/loss/loss.py:66-115
//   photometric splat loss     ldi_enc_dec.py:337-357 (AREA downsample, L1, channel mean, layer min, border crop, mean)
//   disp_smoothness_loss       lsi/geometry/ldi.py:33-68
//   decreasing_disp_loss       lsi/loss/loss.py:48-63
// Scalars live on the device; the upstream gradient of a loss is a device scalar (no host sync anywhere).
#include "capi_common.h"
#include "common.cuh"

namespace lsi {

constexpr int kMaxPartials = 2048;

__global__ void finalize_sum_kernel(const float* __restrict__ partials, int n, double scale, float* __restrict__ out) {
  __shared__ double sh[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += (double)partials[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = (float)(sh[0] * scale);
}

__device__ __forceinline__ float zbw(float r, float scale) {   // helpers.py:180-193
  float c = fminf(fmaxf(r, 0.f), 1.f);
  return r > 0.f ? expf((c - 0.5f) * scale) : 0.f;
}

// ---------------------------------------------------------------------------------------------------------
// zbuffer_composition_loss
// ---------------------------------------------------------------------------------------------------------
struct ZclParams {
  const float* tex; const float* mask; const float* disp; const float* trg; const float* g_loss;
  float* partials; float* d_tex; float* d_mask; float* d_disp;
  int L; long long N;
  float bg_disp, inv_max_disp, scale;
};

template <bool kBackward>
__global__ void __launch_bounds__(256) zcl_kernel(const ZclParams p) {
  float local = 0.f;
  const float gscale = kBackward ? (*p.g_loss) * 0.5f / (float)((double)p.N * 3.0) : 0.f;
  for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < p.N; n += (long long)gridDim.x * blockDim.x) {
    const float t0 = p.trg[n * 3], t1 = p.trg[n * 3 + 1], t2 = p.trg[n * 3 + 2];
    // white background layer: img 1, mask 1, disp bg_layer_disp (loss.py:97-103)
    const float pbg = zbw(p.bg_disp * p.inv_max_disp, p.scale);
    const float ebg = (1.f - t0) * (1.f - t0) + (1.f - t1) * (1.f - t1) + (1.f - t2) * (1.f - t2);
    float P = pbg, E = pbg * ebg;
    for (int l = 0; l < p.L; ++l) {
      const long long i = (long long)l * p.N + n;
      const float m = p.mask ? p.mask[i] : 1.f;
      const float pl = zbw(p.disp[i] * p.inv_max_disp, p.scale) * m;
      const float a = p.tex[i * 3] - t0, b = p.tex[i * 3 + 1] - t1, c = p.tex[i * 3 + 2] - t2;
      P += pl; E = fmaf(pl, a * a + b * b + c * c, E);
    }
    const float Ph = safe_den(P);
    const float cost = E / Ph;   // sum_l (p_l/Ph) e_l
    if (!kBackward) { local += cost; continue; }
    for (int l = 0; l < p.L; ++l) {
      const long long i = (long long)l * p.N + n;
      const float m = p.mask ? p.mask[i] : 1.f;
      const float r = p.disp[i] * p.inv_max_disp;
      const float z = zbw(r, p.scale);
      const float q = z * m / Ph;
      const float a = p.tex[i * 3] - t0, b = p.tex[i * 3 + 1] - t1, c = p.tex[i * 3 + 2] - t2;
      p.d_tex[i * 3] = gscale * 2.f * q * a; p.d_tex[i * 3 + 1] = gscale * 2.f * q * b; p.d_tex[i * 3 + 2] = gscale * 2.f * q * c;
      const float dp = gscale * ((a * a + b * b + c * c) - cost) / Ph;   // dL/dp_l
      if (p.d_mask) p.d_mask[i] = dp * z;
      const float clipg = (r >= 0.f && r <= 1.f) ? 1.f : 0.f;
      p.d_disp[i] = dp * m * z * p.scale * p.inv_max_disp * clipg;
    }
  }
  if (!kBackward) {
    local = block_sum(local);
    if (threadIdx.x == 0) p.partials[blockIdx.x] = local;
  }
}

// ---------------------------------------------------------------------------------------------------------
// photometric loss on the splatted render
// ---------------------------------------------------------------------------------------------------------
struct PhotoParams {
  const float* render; const float* gt; const float* g_loss;
  float* partials; float* d_render;
  int nl, B, H, W, Ht, Wt, fh, fw, x_min, x_max, y_min, y_max;
};

__device__ __forceinline__ void area_mean(const PhotoParams& p, int b, int yt, int xt, float* o) {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f;   // tf.image.resize_images(AREA) with an integer factor == box mean
  for (int dy = 0; dy < p.fh; ++dy)
    for (int dx = 0; dx < p.fw; ++dx) {
      const float* g = p.gt + (((size_t)b * p.H + (yt * p.fh + dy)) * p.W + (xt * p.fw + dx)) * 3;
      s0 += g[0]; s1 += g[1]; s2 += g[2];
    }
  const float inv = 1.f / (float)(p.fh * p.fw);
  o[0] = s0 * inv; o[1] = s1 * inv; o[2] = s2 * inv;
}

template <bool kBackward>
__global__ void __launch_bounds__(256) photo_kernel(const PhotoParams p) {
  const int cw = p.x_max - p.x_min, ch = p.y_max - p.y_min;
  const long long total = (long long)p.B * ch * cw;
  const size_t n_trg = (size_t)p.Ht * p.Wt;
  float local = 0.f;
  const float gscale = kBackward ? (*p.g_loss) / (float)total : 0.f;
  for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < total; n += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(n / ((long long)ch * cw));
    const int rem = (int)(n - (long long)b * ch * cw);
    const int yt = p.y_min + rem / cw, xt = p.x_min + rem % cw;
    float gt[3];
    area_mean(p, b, yt, xt, gt);
    float best = INFINITY; int nbest = 0;
    for (int l = 0; l < p.nl; ++l) {
      const float* r = p.render + (((size_t)l * p.B + b) * n_trg + (size_t)yt * p.Wt + xt) * 3;
      const float e = (fabsf(gt[0] - r[0]) + fabsf(gt[1] - r[1]) + fabsf(gt[2] - r[2])) * (1.f / 3.f);
      if (e < best) { best = e; nbest = 1; } else if (e == best) { ++nbest; }
    }
    if (!kBackward) { local += best; continue; }
    for (int l = 0; l < p.nl; ++l) {   // reduce_min routes the gradient to the arg-min (split over exact ties)
      const size_t o = (((size_t)l * p.B + b) * n_trg + (size_t)yt * p.Wt + xt) * 3;
      const float* r = p.render + o;
      const float e = (fabsf(gt[0] - r[0]) + fabsf(gt[1] - r[1]) + fabsf(gt[2] - r[2])) * (1.f / 3.f);
      if (e != best) continue;
      const float gg = gscale / (3.f * (float)nbest);
      for (int c = 0; c < 3; ++c) {
        const float df = r[c] - gt[c];
        p.d_render[o + c] = df > 0.f ? gg : (df < 0.f ? -gg : 0.f);
      }
    }
  }
  if (!kBackward) {
    local = block_sum(local);
    if (threadIdx.x == 0) p.partials[blockIdx.x] = local;
  }
}

// ---------------------------------------------------------------------------------------------------------
// disparity smoothness: mean|dxx| + mean|dxy| + mean|dyx| + mean|dyy| of second differences (ldi.py:33-68)
//   dx[y][x] = d[y][x+1]-d[y][x]  (H x W-1)    dy[y][x] = d[y+1][x]-d[y][x]  (H-1 x W)
//   gradient(dx) -> (dx2 = d/dx dx : H x W-2, dxdy = d/dy dx : H-1 x W-1)
//   gradient(dy) -> (dydx = d/dx dy : H-1 x W-1, dy2 = d/dy dy : H-2 x W)
// ---------------------------------------------------------------------------------------------------------
struct SmoothParams {
  const float* disp; const float* g_loss; float* partials; float* d_disp;
  int n_img, H, W;
  float inv_xx, inv_xy, inv_yy;   // 1/count of each term (0 when the term is empty)
};

__device__ __forceinline__ float sgn(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); }

template <bool kBackward>
__global__ void __launch_bounds__(256) smooth_kernel(const SmoothParams p) {
  const long long total = (long long)p.n_img * p.H * p.W;
  const float gl = kBackward ? *p.g_loss : 0.f;
  float local = 0.f;
  for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < total; n += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(n % p.W);
    const int y = (int)((n / p.W) % p.H);
    const float* d = p.disp + (n - (long long)y * p.W - x);   // image base
    auto at = [&](int yy, int xx) { return d[(size_t)yy * p.W + xx]; };
    if (!kBackward) {
      // each term is owned by its top-left element
      float v = 0.f;
      if (x + 2 < p.W) v += fabsf(at(y, x + 2) - 2.f * at(y, x + 1) + at(y, x)) * p.inv_xx;
      if (y + 2 < p.H) v += fabsf(at(y + 2, x) - 2.f * at(y + 1, x) + at(y, x)) * p.inv_yy;
      if (x + 1 < p.W && y + 1 < p.H)   // dxdy and dydx are the same mixed difference, counted twice
        v += 2.f * fabsf(at(y + 1, x + 1) - at(y + 1, x) - at(y, x + 1) + at(y, x)) * p.inv_xy;
      local += v;
    } else {
      float gsum = 0.f;
      // xx terms containing (y,x): owners x, x-1, x-2 with coefficients 1, -2, 1
      for (int k = 0; k < 3; ++k) {
        const int xo = x - k;
        if (xo >= 0 && xo + 2 < p.W) {
          const float s = sgn(at(y, xo + 2) - 2.f * at(y, xo + 1) + at(y, xo));
          gsum += s * (k == 1 ? -2.f : 1.f) * p.inv_xx;
        }
        const int yo = y - k;
        if (yo >= 0 && yo + 2 < p.H) {
          const float s = sgn(at(yo + 2, x) - 2.f * at(yo + 1, x) + at(yo, x));
          gsum += s * (k == 1 ? -2.f : 1.f) * p.inv_yy;
        }
      }
      for (int ky = 0; ky < 2; ++ky)
        for (int kx = 0; kx < 2; ++kx) {
          const int yo = y - ky, xo = x - kx;
          if (yo >= 0 && xo >= 0 && yo + 1 < p.H && xo + 1 < p.W) {
            const float s = sgn(at(yo + 1, xo + 1) - at(yo + 1, xo) - at(yo, xo + 1) + at(yo, xo));
            gsum += 2.f * s * ((ky ^ kx) ? -1.f : 1.f) * p.inv_xy;
          }
        }
      p.d_disp[n] = gl * gsum;
    }
  }
  if (!kBackward) {
    local = block_sum(local);
    if (threadIdx.x == 0) p.partials[blockIdx.x] = local;
  }
}

// ---------------------------------------------------------------------------------------------------------
// decreasing_disp_loss: mean relu(d[l+1] - stopgrad(d[l]))   (loss.py:48-63)
// ---------------------------------------------------------------------------------------------------------
struct DecrParams {
  const float* disp; const float* g_loss; float* partials; float* d_disp;
  int L; long long N;
};

template <bool kBackward>
__global__ void __launch_bounds__(256) decr_kernel(const DecrParams p) {
  const long long total = (long long)(p.L - 1) * p.N;
  const float gs = kBackward ? (*p.g_loss) / (float)total : 0.f;
  float local = 0.f;
  for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < total; n += (long long)gridDim.x * blockDim.x) {
    const float v = p.disp[n + p.N] - p.disp[n];
    if (!kBackward) local += fmaxf(v, 0.f);
    else p.d_disp[n + p.N] = v > 0.f ? gs : 0.f;   // layer 0 gets no gradient (stop_gradient on d[l])
  }
  if (!kBackward) {
    local = block_sum(local);
    if (threadIdx.x == 0) p.partials[blockIdx.x] = local;
  }
}

static int grid_for(long long n) {
  long long g = (n + 255) / 256;
  if (g > kMaxPartials) g = kMaxPartials;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace lsi

using namespace lsi;

extern "C" size_t lsi_b200_loss_partials_count(void) { return kMaxPartials; }

extern "C" int lsi_b200_zbuf_composition_loss(const float* tex, const float* mask, const float* disp, const float* trg,
                                              int n_layers, long long n_pixels, float bg_layer_disp, float max_disp,
                                              float zbuf_scale, float* loss, float* partials, void* stream) {
  LSI_REQUIRE(tex && disp && trg && loss && partials, "NULL pointer argument");
  LSI_REQUIRE(n_layers >= 1 && n_pixels >= 1 && max_disp != 0.f, "bad sizes");
  ZclParams p{tex, mask, disp, trg, nullptr, partials, nullptr, nullptr, nullptr, n_layers, n_pixels,
              bg_layer_disp, 1.f / max_disp, zbuf_scale};
  const int g = grid_for(n_pixels);
  zcl_kernel<false><<<g, 256, 0, as_stream(stream)>>>(p);
  LSI_LAUNCH_CHECK();
  finalize_sum_kernel<<<1, 256, 0, as_stream(stream)>>>(partials, g, 0.5 / ((double)n_pixels * 3.0), loss);
  LSI_LAUNCH_CHECK();
  return LSI_B200_OK;
}

extern "C" int lsi_b200_zbuf_composition_loss_backward(const float* tex, const float* mask, const float* disp,
                                                       const float* trg, int n_layers, long long n_pixels,
                                                       float bg_layer_disp, float max_disp, float zbuf_scale,
                                                       const float* g_loss, float* d_tex, float* d_mask,
                                                       float* d_disp, void* stream) {
  LSI_REQUIRE(tex && disp && trg && g_loss && d_tex && d_disp, "NULL pointer argument");
  LSI_REQUIRE(n_layers >= 1 && n_pixels >= 1 && max_disp != 0.f, "bad sizes");
  ZclParams p{tex, mask, disp, trg, g_loss, nullptr, d_tex, d_mask, d_disp, n_layers, n_pixels,
              bg_layer_disp, 1.f / max_disp, zbuf_scale};
  zcl_kernel<true><<<grid_for(n_pixels), 256, 0, as_stream(stream)>>>(p);
  LSI_LAUNCH_CHECK();
  return LSI_B200_OK;
}

static int photo_setup(PhotoParams& p, int n_layers, int batch, int h, int w, int h_t, int w_t, float bdry) {
  LSI_REQUIRE(n_layers >= 1 && batch >= 1 && h >= 1 && w >= 1 && h_t >= 1 && w_t >= 1, "bad sizes");
  LSI_REQUIRE(h % h_t == 0 && w % w_t == 0, "AREA resize supports integer factors only (%dx%d -> %dx%d)", h, w, h_t, w_t);
  p.nl = n_layers; p.B = batch; p.H = h; p.W = w; p.Ht = h_t; p.Wt = w_t; p.fh = h / h_t; p.fw = w / w_t;
  // python: int(round(loss_w * bdry)) (ldi_enc_dec.py:348-351); Python-2 round() rounds half away from zero
  p.x_min = (int)floor((double)w_t * (double)bdry + 0.5); p.x_max = w_t - p.x_min;
  p.y_min = (int)floor((double)h_t * (double)bdry + 0.5); p.y_max = h_t - p.y_min;
  LSI_REQUIRE(p.x_max > p.x_min && p.y_max > p.y_min, "splat_bdry_ignore leaves no pixels");
  return LSI_B200_OK;
}

extern "C" int lsi_b200_photo_loss(const float* render, const float* gt, int n_layers, int batch, int h, int w, int h_t,
                                   int w_t, float bdry_ignore, float* loss, float* partials, void* stream) {
  LSI_REQUIRE(render && gt && loss && partials, "NULL pointer argument");
  PhotoParams p{};
  if (int rc = photo_setup(p, n_layers, batch, h, w, h_t, w_t, bdry_ignore)) return rc;
  p.render = render; p.gt = gt; p.partials = partials;
  const long long total = (long long)batch * (p.y_max - p.y_min) * (p.x_max - p.x_min);
  const int g = grid_for(total);
  photo_kernel<false><<<g, 256, 0, as_stream(stream)>>>(p);
  LSI_LAUNCH_CHECK();
  finalize_sum_kernel<<<1, 256, 0, as_stream(stream)>>>(partials, g, 1.0 / (double)total, loss);
  LSI_LAUNCH_CHECK();
  return LSI_B200_OK;
}

extern "C" int lsi_b200_photo_loss_backward(const float* render, const float* gt, int n_layers, int batch, int h, int w,
                                            int h_t, int w_t, float bdry_ignore, const float* g_loss, float* d_render,
                                            void* stream) {
  LSI_REQUIRE(render && gt && g_loss && d_render, "NULL pointer argument");
  PhotoParams p{};
  if (int rc = photo_setup(p, n_layers, batch, h, w, h_t, w_t, bdry_ignore)) return rc;
  p.render = render; p.gt = gt; p.g_loss = g_loss; p.d_render = d_render;
  LSI_CUDA(cudaMemsetAsync(d_render, 0, (size_t)n_layers * batch * h_t * w_t * 3 * 4, as_stream(stream)));
  const long long total = (long long)batch * (p.y_max - p.y_min) * (p.x_max - p.x_min);
  photo_kernel<true><<<grid_for(total), 256, 0, as_stream(stream)>>>(p);
  LSI_LAUNCH_CHECK();
  return LSI_B200_OK;
}

static void smooth_setup(SmoothParams& p, int n_img, int h, int w) {
  p.n_img = n_img; p.H = h; p.W = w;
  const double nxx = (double)n_img * h * (w > 2 ? w - 2 : 0), nyy = (double)n_img * (h > 2 ? h - 2 : 0) * w;
  const double nxy = (double)n_img * (h > 1 ? h - 1 : 0) * (w > 1 ? w - 1 : 0);
  p.inv_xx = nxx > 0 ? (float)(1.0 / nxx) : 0.f; p.inv_yy = nyy > 0 ? (float)(1.0 / nyy) : 0.f;
  p.inv_xy = nxy > 0 ? (float)(1.0 / nxy) : 0.f;
}

extern "C" int lsi_b200_disp_smoothness_loss(const float* disp, int n_images, int h, int w, float* loss,
                                             float* partials, void* stream) {
  LSI_REQUIRE(disp && loss && partials, "NULL pointer argument");
  LSI_REQUIRE(n_images >= 1 && h >= 3 && w >= 3, "disp_smoothness_loss needs h,w >= 3");
  SmoothParams p{}; smooth_setup(p, n_images, h, w);
  p.disp = disp; p.partials = partials;
  const int g = grid_for((long long)n_images * h * w);
  smooth_kernel<false><<<g, 256, 0, as_stream(stream)>>>(p);
  LSI_LAUNCH_CHECK();
  finalize_sum_kernel<<<1, 256, 0, as_stream(stream)>>>(partials, g, 1.0, loss);
  LSI_LAUNCH_CHECK();
  return LSI_B200_OK;
}

extern "C" int lsi_b200_disp_smoothness_loss_backward(const float* disp, int n_images, int h, int w,
                                                      const float* g_loss, float* d_disp, void* stream) {
  LSI_REQUIRE(disp && g_loss && d_disp, "NULL pointer argument");
  LSI_REQUIRE(n_images >= 1 && h >= 3 && w >= 3, "disp_smoothness_loss needs h,w >= 3");
  SmoothParams p{}; smooth_setup(p, n_images, h, w);
  p.disp = disp; p.g_loss = g_loss; p.d_disp = d_disp;
  smooth_kernel<true><<<grid_for((long long)n_images * h * w), 256, 0, as_stream(stream)>>>(p);
  LSI_LAUNCH_CHECK();
  return LSI_B200_OK;
}

extern "C" int lsi_b200_decreasing_disp_loss(const float* disp, int n_layers, long long n_pixels, float* loss,
                                             float* partials, void* stream) {
  LSI_REQUIRE(disp && loss && partials, "NULL pointer argument");
  LSI_REQUIRE(n_layers >= 1 && n_pixels >= 1, "bad sizes");
  if (n_layers == 1) {   // loss.py:57-58 returns 0 for a single layer
    LSI_CUDA(cudaMemsetAsync(loss, 0, 4, as_stream(stream)));
    return LSI_B200_OK;
  }
  DecrParams p{disp, nullptr, partials, nullptr, n_layers, n_pixels};
  const long long total = (long long)(n_layers - 1) * n_pixels;
  const int g = grid_for(total);
  decr_kernel<false><<<g, 256, 0, as_stream(stream)>>>(p);
  LSI_LAUNCH_CHECK();
  finalize_sum_kernel<<<1, 256, 0, as_stream(stream)>>>(partials, g, 1.0 / (double)total, loss);
  LSI_LAUNCH_CHECK();
  return LSI_B200_OK;
}

extern "C" int lsi_b200_decreasing_disp_loss_backward(const float* disp, int n_layers, long long n_pixels,
                                                      const float* g_loss, float* d_disp, void* stream) {
  LSI_REQUIRE(disp && g_loss && d_disp, "NULL pointer argument");
  LSI_REQUIRE(n_layers >= 1 && n_pixels >= 1, "bad sizes");
  LSI_CUDA(cudaMemsetAsync(d_disp, 0, (size_t)n_layers * n_pixels * 4, as_stream(stream)));
  if (n_layers == 1) return LSI_B200_OK;
  DecrParams p{disp, g_loss, nullptr, d_disp, n_layers, n_pixels};
  decr_kernel<true><<<grid_for((long long)(n_layers - 1) * n_pixels), 256, 0, as_stream(stream)>>>(p);
  LSI_LAUNCH_CHECK();
  return LSI_B200_OK;
}
