// Generic sampling primitives with the reference's public signatures: sampling.splat (lsi/geometry/sampling.py:171-254)
// and sampling.bilinear (sampling.py:41-132), forward and backward.  These are the API-complete (any channel
// count, caller-supplied coordinates) versions; the training hot path uses the fused kernels in render.cu.
#include "capi_common.h"
#include "common.cuh"

namespace lsi {

struct SampParams {
  const float* a;        // splat: src [B,Hs,Ws,C]      bilinear: imgs [B,Hs,Ws,C]
  const float* coords;   // splat: [B,Hs,Ws,2]          bilinear: [B,Ht,Wt,2]
  const float* g;        // upstream gradient (backward only)
  float* out;            // forward output / d_src or d_imgs
  float* d_coords;
  int B, Hs, Ws, Ht, Wt, C;
};

// out was initialised with init; one thread per source pixel, scalar reductions per channel.
__global__ void __launch_bounds__(256) splat_generic_kernel(const SampParams p) {
  const long long n_src = (long long)p.Hs * p.Ws;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_src * p.B) return;
  const int b = (int)(idx / n_src);
  PixGeom g;
  const float x = p.coords[idx * 2] - 0.5f, y = p.coords[idx * 2 + 1] - 0.5f;   // sampling.py:183
  corners(x, y, p.Wt, p.Ht, g);
  const float* s = p.a + idx * p.C;
  float* o = p.out + (size_t)b * p.Ht * p.Wt * p.C;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    if (!g.keep[c]) continue;
    float* oc = o + (size_t)corner_index(g, c, p.Wt) * p.C;
    for (int ch = 0; ch < p.C; ++ch) atomicAdd(oc + ch, s[ch] * g.w[c]);
  }
}

__global__ void __launch_bounds__(256) splat_generic_bwd_kernel(const SampParams p) {
  const long long n_src = (long long)p.Hs * p.Ws;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_src * p.B) return;
  const int b = (int)(idx / n_src);
  PixGeom g;
  const float x = p.coords[idx * 2] - 0.5f, y = p.coords[idx * 2 + 1] - 0.5f;
  corners(x, y, p.Wt, p.Ht, g);
  const float* s = p.a + idx * p.C;
  const float* gg = p.g + (size_t)b * p.Ht * p.Wt * p.C;
  const float dwx[4] = {-g.vx0 * g.wy0, g.vx1 * g.wy0, -g.vx0 * g.wy1, g.vx1 * g.wy1};
  const float dwy[4] = {-g.wx0 * g.vy0, -g.wx1 * g.vy0, g.wx0 * g.vy1, g.wx1 * g.vy1};
  float gx = 0.f, gy = 0.f;
  float* ds = p.out + idx * p.C;
  for (int ch = 0; ch < p.C; ++ch) {
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (!g.keep[c]) continue;
      const float gv = __ldg(gg + (size_t)corner_index(g, c, p.Wt) * p.C + ch);
      acc = fmaf(g.w[c], gv, acc);
      const float sv = s[ch] * gv;
      gx = fmaf(sv, dwx[c], gx); gy = fmaf(sv, dwy[c], gy);
    }
    ds[ch] = acc;
  }
  p.d_coords[idx * 2] = gx; p.d_coords[idx * 2 + 1] = gy;
}

// sampling.py:41-132: gather at the four (clipped) corners, weights NOT thresholded, validity zeroes a corner.
__global__ void __launch_bounds__(256) bilinear_kernel(const SampParams p) {
  const long long n_trg = (long long)p.Ht * p.Wt;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_trg * p.B) return;
  const int b = (int)(idx / n_trg);
  const float x = p.coords[idx * 2] - 0.5f, y = p.coords[idx * 2 + 1] - 0.5f;   // sampling.py:54
  PixGeom g;
  corners(x, y, p.Ws, p.Hs, g);
  const float wt[4] = {g.wx0 * g.wy0, g.wx1 * g.wy0, g.wx0 * g.wy1, g.wx1 * g.wy1};
  const float* im = p.a + (size_t)b * p.Hs * p.Ws * p.C;
  float* o = p.out + idx * p.C;
  for (int ch = 0; ch < p.C; ++ch) {
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c)
      if (wt[c] != 0.f) acc = fmaf(wt[c], __ldg(im + (size_t)corner_index(g, c, p.Ws) * p.C + ch), acc);
    o[ch] = acc;
  }
}

__global__ void __launch_bounds__(256) bilinear_bwd_kernel(const SampParams p) {
  const long long n_trg = (long long)p.Ht * p.Wt;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_trg * p.B) return;
  const int b = (int)(idx / n_trg);
  const float x = p.coords[idx * 2] - 0.5f, y = p.coords[idx * 2 + 1] - 0.5f;
  PixGeom g;
  corners(x, y, p.Ws, p.Hs, g);
  const float wt[4] = {g.wx0 * g.wy0, g.wx1 * g.wy0, g.wx0 * g.wy1, g.wx1 * g.wy1};
  const float dwx[4] = {-g.vx0 * g.wy0, g.vx1 * g.wy0, -g.vx0 * g.wy1, g.vx1 * g.wy1};
  const float dwy[4] = {-g.wx0 * g.vy0, -g.wx1 * g.vy0, g.wx0 * g.vy1, g.wx1 * g.vy1};
  const bool valid[4] = {g.vx0 * g.vy0 != 0.f, g.vx1 * g.vy0 != 0.f, g.vx0 * g.vy1 != 0.f, g.vx1 * g.vy1 != 0.f};
  const float* im = p.a + (size_t)b * p.Hs * p.Ws * p.C;
  float* di = p.out + (size_t)b * p.Hs * p.Ws * p.C;
  const float* gg = p.g + idx * p.C;
  float gx = 0.f, gy = 0.f;
  for (int ch = 0; ch < p.C; ++ch) {
    const float gv = gg[ch];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (!valid[c]) continue;
      const size_t q = (size_t)corner_index(g, c, p.Ws) * p.C + ch;
      if (wt[c] != 0.f) atomicAdd(di + q, wt[c] * gv);
      const float sv = __ldg(im + q) * gv;
      gx = fmaf(sv, dwx[c], gx); gy = fmaf(sv, dwy[c], gy);
    }
  }
  p.d_coords[idx * 2] = gx; p.d_coords[idx * 2 + 1] = gy;
}

// sampling.py:117-131, compose=False: the four corner samples and their weights separately (the reference's data generator
// composites layers from them).  out_ims[i] = [corner valid] * imgs[corner]; out_wts[i] = the raw bilinear weight (NOT masked by
// validity, exactly as sampling.py:83-86,123-126); i runs over (x0,y0), (x0,y1), (x1,y0), (x1,y1) -- the reference's list order.
__global__ void __launch_bounds__(256) bilinear_corners_kernel(const SampParams p, float* __restrict__ out_wts) {
  const long long n_trg = (long long)p.Ht * p.Wt;
  const long long total = n_trg * p.B;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int b = (int)(idx / n_trg);
  const float x = p.coords[idx * 2] - 0.5f, y = p.coords[idx * 2 + 1] - 0.5f;
  PixGeom g;
  corners(x, y, p.Ws, p.Hs, g);
  const float x0 = floorf(x), y0 = floorf(y);
  const float rx0 = (x0 + 1.f) - x, rx1 = x - x0, ry0 = (y0 + 1.f) - y, ry1 = y - y0;
  const int order[4] = {0, 2, 1, 3};                       // corners(): bit 0 = x1, bit 1 = y1
  const float raw[4] = {rx0 * ry0, rx0 * ry1, rx1 * ry0, rx1 * ry1};
  const float valid[4] = {g.vx0 * g.vy0, g.vx0 * g.vy1, g.vx1 * g.vy0, g.vx1 * g.vy1};
  const float* im = p.a + (size_t)b * p.Hs * p.Ws * p.C;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    out_wts[(size_t)i * total + idx] = raw[i];
    float* o = p.out + ((size_t)i * total + idx) * p.C;
    const float* src = im + (size_t)corner_index(g, order[i], p.Ws) * p.C;
    for (int ch = 0; ch < p.C; ++ch) o[ch] = valid[i] != 0.f ? __ldg(src + ch) : 0.f;
  }
}

static int check_samp(const void* a, const void* c, const void* o, int B, int Hs, int Ws, int Ht, int Wt, int C) {
  LSI_REQUIRE(a && c && o, "NULL pointer argument");
  LSI_REQUIRE(B >= 1 && Hs >= 1 && Ws >= 1 && Ht >= 1 && Wt >= 1 && C >= 1, "sizes must be >= 1");
  LSI_REQUIRE((long long)Hs * Ws < (1ll << 30) && (long long)Ht * Wt < (1ll << 30), "image too large");
  return LSI_B200_OK;
}

static unsigned blocks_for(long long n) { return (unsigned)((n + 255) / 256); }

}  // namespace lsi

using namespace lsi;

extern "C" int lsi_b200_splat(const float* src, const float* coords, const float* init, float* out, int batch, int h_s,
                              int w_s, int h_t, int w_t, int channels, void* stream) {
  if (int rc = check_samp(src, coords, out, batch, h_s, w_s, h_t, w_t, channels)) return rc;
  LSI_REQUIRE(init != nullptr, "NULL pointer argument");
  cudaStream_t st = as_stream(stream);
  if (init != out)   // functional: init is never modified (sampling.py:283)
    LSI_CUDA(cudaMemcpyAsync(out, init, (size_t)batch * h_t * w_t * channels * 4, cudaMemcpyDeviceToDevice, st));
  SampParams p{src, coords, nullptr, out, nullptr, batch, h_s, w_s, h_t, w_t, channels};
  splat_generic_kernel<<<blocks_for((long long)batch * h_s * w_s), 256, 0, st>>>(p);
  LSI_LAUNCH_CHECK();
  return LSI_B200_OK;
}

extern "C" int lsi_b200_splat_backward(const float* src, const float* coords, const float* g, float* d_src,
                                       float* d_coords, int batch, int h_s, int w_s, int h_t, int w_t, int channels,
                                       void* stream) {
  if (int rc = check_samp(src, coords, d_src, batch, h_s, w_s, h_t, w_t, channels)) return rc;
  LSI_REQUIRE(g && d_coords, "NULL pointer argument");
  SampParams p{src, coords, g, d_src, d_coords, batch, h_s, w_s, h_t, w_t, channels};
  splat_generic_bwd_kernel<<<blocks_for((long long)batch * h_s * w_s), 256, 0, as_stream(stream)>>>(p);
  LSI_LAUNCH_CHECK();
  return LSI_B200_OK;
}

extern "C" int lsi_b200_bilinear(const float* imgs, const float* coords, float* out, int batch, int h_s, int w_s,
                                 int h_t, int w_t, int channels, void* stream) {
  if (int rc = check_samp(imgs, coords, out, batch, h_s, w_s, h_t, w_t, channels)) return rc;
  SampParams p{imgs, coords, nullptr, out, nullptr, batch, h_s, w_s, h_t, w_t, channels};
  bilinear_kernel<<<blocks_for((long long)batch * h_t * w_t), 256, 0, as_stream(stream)>>>(p);
  LSI_LAUNCH_CHECK();
  return LSI_B200_OK;
}

extern "C" int lsi_b200_bilinear_backward(const float* imgs, const float* coords, const float* g, float* d_imgs,
                                          float* d_coords, int batch, int h_s, int w_s, int h_t, int w_t,
                                          int channels, void* stream) {
  if (int rc = check_samp(imgs, coords, d_imgs, batch, h_s, w_s, h_t, w_t, channels)) return rc;
  LSI_REQUIRE(g && d_coords, "NULL pointer argument");
  SampParams p{imgs, coords, g, d_imgs, d_coords, batch, h_s, w_s, h_t, w_t, channels};
  bilinear_bwd_kernel<<<blocks_for((long long)batch * h_t * w_t), 256, 0, as_stream(stream)>>>(p);
  LSI_LAUNCH_CHECK();
  return LSI_B200_OK;
}

// bilinear(imgs, coords, compose=False) (sampling.py:117-131): out_ims [4,B,Ht,Wt,C], out_wts [4,B,Ht,Wt,1].
extern "C" int lsi_b200_bilinear_corners(const float* imgs, const float* coords, float* out_ims, float* out_wts, int batch, int h_s,
                                         int w_s, int h_t, int w_t, int channels, void* stream) {
  if (int rc = check_samp(imgs, coords, out_ims, batch, h_s, w_s, h_t, w_t, channels)) return rc;
  LSI_REQUIRE(out_wts != nullptr, "NULL pointer argument");
  SampParams p{imgs, coords, nullptr, out_ims, nullptr, batch, h_s, w_s, h_t, w_t, channels};
  bilinear_corners_kernel<<<blocks_for((long long)batch * h_t * w_t), 256, 0, as_stream(stream)>>>(p, out_wts);
  LSI_LAUNCH_CHECK();
  return LSI_B200_OK;
}
