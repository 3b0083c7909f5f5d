// Data-path kernel of the KITTI loader (lsi/data/kitti/data.py:247-266 of the reference): the decoded 8-bit image is scaled
// to [0, 1] and resized to the network resolution with tf.image.resize_images(method=AREA).  [TF1.4] ResizeArea semantics,
// restated: scale = in / out per axis; output pixel y covers the input interval [y * scale, (y + 1) * scale); every input
// row i it touches contributes with the covered fraction of the row (1 for fully covered rows), indices clamped to the
// image; the sum is divided by scale_y * scale_x.  For integer factors this is the box mean used elsewhere.
// (TF evaluates the interval bounds in fp32; here they are exact, which is what the op is defined to compute.)
// One thread per output pixel (all channels): the footprint at KITTI's 1242x375 -> 832x256 is at most 3 x 3 input pixels.
#include "capi_common.h"
#include "common.cuh"

namespace lsi {

struct AreaParams {
  const unsigned char* in; float* out;
  int h_in, w_in, c_in, h_out, w_out, nc;
  float sy, sx, norm;     // in/out scales, 1 / (255 * sy * sx)
};

__device__ __forceinline__ float area_weight(int i, double lo, double hi) {
  // fraction of input cell [i, i + 1) covered by [lo, hi)
  const double fi = (double)i;
  if (fi < lo) return (float)((fi + 1.0 > hi) ? hi - lo : fi + 1.0 - lo);
  return (float)((fi + 1.0 > hi) ? hi - fi : 1.0);
}

__global__ void __launch_bounds__(256) area_resize_u8_kernel(const AreaParams p) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= p.w_out) return;
  // interval bounds as exact rationals y * h_in / h_out in fp64 (a few per thread): in fp32 the last column's upper bound
  // 832 * (1242 / 832) rounds to 1242.0001 and picks up a spurious clamped pixel (7e-5 of full scale)
  const double y0 = (double)((long long)y * p.h_in) / p.h_out, y1 = (double)((long long)(y + 1) * p.h_in) / p.h_out;
  const double x0 = (double)((long long)x * p.w_in) / p.w_out, x1 = (double)((long long)(x + 1) * p.w_in) / p.w_out;
  const int iy0 = (int)floor(y0), iy1 = (int)ceil(y1), ix0 = (int)floor(x0), ix1 = (int)ceil(x1);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int i = iy0; i < iy1; ++i) {
    const float wy = area_weight(i, y0, y1);
    const int yi = min(max(i, 0), p.h_in - 1);
    for (int j = ix0; j < ix1; ++j) {
      const float w = wy * area_weight(j, x0, x1);
      const int xj = min(max(j, 0), p.w_in - 1);
      const unsigned char* px = p.in + ((size_t)yi * p.w_in + xj) * p.c_in;
      for (int c = 0; c < p.nc; ++c) acc[c] = fmaf((float)__ldg(px + c), w, acc[c]);
    }
  }
  float* o = p.out + ((size_t)y * p.w_out + x) * p.nc;
  for (int c = 0; c < p.nc; ++c) o[c] = acc[c] * p.norm;
}

}  // namespace lsi

using namespace lsi;

extern "C" int lsi_b200_area_resize_u8(const unsigned char* in, int h_in, int w_in, int c_in, float* out, int h_out, int w_out,
                                       int nc, void* stream) {
  LSI_REQUIRE(in && out, "NULL pointer argument");
  LSI_REQUIRE(h_in >= 1 && w_in >= 1 && h_out >= 1 && w_out >= 1 && h_out <= 65535, "bad image sizes");
  LSI_REQUIRE(nc >= 1 && nc <= 4 && nc <= c_in, "nc=%d channels requested from a %d-channel image (1..4 supported)", nc, c_in);
  AreaParams p;
  p.in = in; p.out = out; p.h_in = h_in; p.w_in = w_in; p.c_in = c_in; p.h_out = h_out; p.w_out = w_out; p.nc = nc;
  p.sy = (float)h_in / (float)h_out; p.sx = (float)w_in / (float)w_out;
  p.norm = 1.f / (255.f * p.sy * p.sx);
  area_resize_u8_kernel<<<dim3((w_out + 255) / 256, h_out), 256, 0, as_stream(stream)>>>(p);
  LSI_LAUNCH_CHECK();
  return LSI_B200_OK;
}
