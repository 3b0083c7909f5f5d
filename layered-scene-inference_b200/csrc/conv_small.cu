// Two CUDA-core kernels for the layers of the CNN whose contraction is too thin for the tensor-core paths (TMA needs 16-byte
// pixels and 32-channel chunks) and too lopsided for the generic fp32 implicit GEMM of conv.cu, which the profile of the training
// step showed at 6.3 ms (stem weight gradient) and 0.8 ms (data gradient of the prediction conv) per call
// (profiles/r2_train_calls_v1.txt):
//
//   thin gather      out[n,oy,ox,0..Co) = sum_{taps, ci < Ci <= 8} in[n, iy, ix, ci] * W(tap, ci, co), unit stride, Co <= 32:
//                    the data gradient of pixelwise_predictor's 3x3 conv (nets.py:139-155: 4 -> 32 channels).  One thread
//                    per output pixel, the whole filter bank in shared memory (broadcast reads), 32 accumulators in registers.
//   stem wgrad       dW[ky,kx,ci < 4,co < 32] = sum_pixels x[n, oy*s - pad + ky, ox*s - pad + kx, ci] * dz[n,oy,ox,co]:
//                    the weight gradient of cnv1 (nets.py:273: 7x7 stride 2, 3 -> 32).  Persistent CTAs; a warp owns a
//                    subset of the (tap, ci) rows, a lane one output channel; input patch and dz tile staged in shared
//                    memory; partial sums live in registers across all tiles of the CTA and are added to dW once.
#include "capi_common.h"
#include "common.cuh"

namespace lsi {

struct ThinParams {
  const float* in; const float* w; float* out;
  int N, Hi, Wi, Ci, Ho, Wo, Co, kh, kw, pad_t, pad_l, mode, w_tap, w_ci, w_co, in_cs, out_cs, accumulate;
};

constexpr int kThinMaxCi = 8, kThinCo = 32, kThinMaxTaps = 9;

__global__ void __launch_bounds__(128) conv_thin_kernel(const ThinParams p) {
  __shared__ float ws[kThinMaxTaps * kThinMaxCi * kThinCo];           // [tap][ci][co], zero for co >= Co
  const int taps = p.kh * p.kw;
  for (int i = threadIdx.x; i < taps * p.Ci * kThinCo; i += blockDim.x) {
    const int co = i % kThinCo, ci = (i / kThinCo) % p.Ci, tap = i / (kThinCo * p.Ci);
    ws[i] = co < p.Co ? p.w[(size_t)tap * p.w_tap + (size_t)ci * p.w_ci + (size_t)co * p.w_co] : 0.f;
  }
  __syncthreads();
  const long long total = (long long)p.N * p.Ho * p.Wo;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(q % p.Wo);
    const long long r = q / p.Wo;
    const int oy = (int)(r % p.Ho), n = (int)(r / p.Ho);
    float acc[kThinCo];
#pragma unroll
    for (int c = 0; c < kThinCo; ++c) acc[c] = 0.f;
    for (int ky = 0; ky < p.kh; ++ky) {
      const int iy = p.mode == 0 ? oy - p.pad_t + ky : oy + p.pad_t - ky;
      if ((unsigned)iy >= (unsigned)p.Hi) continue;
      for (int kx = 0; kx < p.kw; ++kx) {
        const int ix = p.mode == 0 ? ox - p.pad_l + kx : ox + p.pad_l - kx;
        if ((unsigned)ix >= (unsigned)p.Wi) continue;
        const float* src = p.in + ((size_t)(n * p.Hi + iy) * p.Wi + ix) * p.in_cs;
        const float* wt = ws + (ky * p.kw + kx) * p.Ci * kThinCo;
        for (int ci = 0; ci < p.Ci; ++ci) {
          const float v = __ldg(src + ci);
          const float4* w4 = reinterpret_cast<const float4*>(wt + ci * kThinCo);
#pragma unroll
          for (int c = 0; c < kThinCo / 4; ++c) {
            const float4 wv = w4[c];
            acc[4 * c] = fmaf(v, wv.x, acc[4 * c]); acc[4 * c + 1] = fmaf(v, wv.y, acc[4 * c + 1]);
            acc[4 * c + 2] = fmaf(v, wv.z, acc[4 * c + 2]); acc[4 * c + 3] = fmaf(v, wv.w, acc[4 * c + 3]);
          }
        }
      }
    }
    float* dst = p.out + (size_t)q * p.out_cs;
    if (p.Co == kThinCo && (p.out_cs & 3) == 0 && ((uintptr_t)p.out & 15) == 0) {
#pragma unroll
      for (int c = 0; c < kThinCo; c += 4) {
        float4 o = make_float4(acc[c], acc[c + 1], acc[c + 2], acc[c + 3]);
        if (p.accumulate) { const float4 a = *reinterpret_cast<const float4*>(dst + c); o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w; }
        *reinterpret_cast<float4*>(dst + c) = o;
      }
    } else {
#pragma unroll
      for (int c = 0; c < kThinCo; ++c)
        if (c < p.Co) dst[c] = p.accumulate ? dst[c] + acc[c] : acc[c];
    }
  }
}

// The shape the training step actually sends here (data gradient of the 3x3 prediction conv: 4 -> 32 channels at full resolution,
// 1.9 ms per step in the one-pixel-per-thread kernel above: one broadcast LDS.128 of weights per four FMAs).  Four lanes share four
// horizontally adjacent output pixels, each lane 8 of the 32 output channels: a weight vector is loaded once for four pixels
// (72 instead of 288 shared-memory loads per thread, 1152 FMAs either way), the six input columns of a filter row come in as six
// 128-bit loads, and the stores of a warp cover 4 KB contiguously.
__global__ void __launch_bounds__(128) conv_thin4_kernel(const ThinParams p) {
  __shared__ __align__(16) float ws[9 * 4 * kThinCo];                 // [tap][ci][co]
  for (int i = threadIdx.x; i < 9 * 4 * kThinCo; i += blockDim.x) {
    const int co = i % kThinCo, ci = (i / kThinCo) % 4, tap = i / (kThinCo * 4);
    ws[i] = p.w[(size_t)tap * p.w_tap + (size_t)ci * p.w_ci + (size_t)co * p.w_co];
  }
  __syncthreads();
  const int q = threadIdx.x & 3;                                       // channel octet
  const int gx = (p.Wo + 3) >> 2;                                      // pixel quads per output row
  const long long groups = (long long)p.N * p.Ho * gx;
  const float4* in4 = reinterpret_cast<const float4*>(p.in);
  for (long long g = (long long)blockIdx.x * 32 + (threadIdx.x >> 2); g < groups; g += (long long)gridDim.x * 32) {
    const int ox0 = (int)(g % gx) * 4;
    const long long r = g / gx;
    const int oy = (int)(r % p.Ho), n = (int)(r / p.Ho);
    float acc[4][8];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[j][c] = 0.f;
    // input columns of one filter row: mode 0 (gather): ix = ox - pad_l + kx; mode 1 (transposed): ix = ox + pad_l - kx.  Either way the
    // four pixels and three taps touch six consecutive columns starting at c0; pixel j with tap kx reads column c0 + j + (mode ? 2 - kx : kx)
    const int c0 = p.mode == 0 ? ox0 - p.pad_l : ox0 + p.pad_l - 2;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = p.mode == 0 ? oy - p.pad_t + ky : oy + p.pad_t - ky;
      if ((unsigned)iy >= (unsigned)p.Hi) continue;
      const float4* row = in4 + (size_t)(n * p.Hi + iy) * p.Wi;
      float4 v[6];
#pragma unroll
      for (int c = 0; c < 6; ++c) v[c] = (unsigned)(c0 + c) < (unsigned)p.Wi ? __ldg(row + c0 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const float* wt = ws + ((ky * 3 + kx) * 4) * kThinCo + q * 8;
        const int sh = p.mode == 0 ? kx : 2 - kx;
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
          const float4 wa = *reinterpret_cast<const float4*>(wt + ci * kThinCo), wb = *reinterpret_cast<const float4*>(wt + ci * kThinCo + 4);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 vv = sh == 0 ? v[j] : (sh == 1 ? v[j + 1] : v[j + 2]);
            const float x = ci == 0 ? vv.x : (ci == 1 ? vv.y : (ci == 2 ? vv.z : vv.w));
            acc[j][0] = fmaf(x, wa.x, acc[j][0]); acc[j][1] = fmaf(x, wa.y, acc[j][1]); acc[j][2] = fmaf(x, wa.z, acc[j][2]); acc[j][3] = fmaf(x, wa.w, acc[j][3]);
            acc[j][4] = fmaf(x, wb.x, acc[j][4]); acc[j][5] = fmaf(x, wb.y, acc[j][5]); acc[j][6] = fmaf(x, wb.z, acc[j][6]); acc[j][7] = fmaf(x, wb.w, acc[j][7]);
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (ox0 + j < p.Wo) {
        float4* dst = reinterpret_cast<float4*>(p.out + ((size_t)(n * p.Ho + oy) * p.Wo + ox0 + j) * p.out_cs + q * 8);
        float4 o0 = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]), o1 = make_float4(acc[j][4], acc[j][5], acc[j][6], acc[j][7]);
        if (p.accumulate) { const float4 a = dst[0], b = dst[1]; o0.x += a.x; o0.y += a.y; o0.z += a.z; o0.w += a.w; o1.x += b.x; o1.y += b.y; o1.z += b.z; o1.w += b.w; }
        dst[0] = o0; dst[1] = o1;
      }
    }
  }
}

// ---- stem weight gradient -------------------------------------------------------------------------------------------
constexpr int kSwTile = 32;          // output pixels of one row per tile
constexpr int kSwWarps = 8;
constexpr int kSwMaxRows = 7 * 7 * 4;   // (tap, ci) rows of dW

struct StemWgradParams {
  const float* x; const float* dz; float* dw;
  int N, Hi, Wi, Ci, Ho, Wo, kh, kw, stride, pad_t, pad_l, x_cs, dz_cs;
  int tiles_x; long long tiles;
};

__global__ void __launch_bounds__(kSwWarps * 32, 2) stem_wgrad_kernel(const StemWgradParams p) {
  extern __shared__ float sw_smem[];
  const int patch_w = (kSwTile - 1) * p.stride + p.kw;                 // input columns one tile touches
  float* xs = sw_smem;                                                 // [kh][patch_w][Ci]
  float* dzs = xs + p.kh * patch_w * p.Ci;                             // [kSwTile][32]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rows = p.kh * p.kw * p.Ci;                                 // (tap, ci) rows; row r -> warp r % 8
  constexpr int kMaxPerWarp = (kSwMaxRows + kSwWarps - 1) / kSwWarps;  // 25
  float acc[kMaxPerWarp];
  int xoff[kMaxPerWarp];                                               // offset of row r's first input value inside the patch
#pragma unroll
  for (int k = 0; k < kMaxPerWarp; ++k) {
    acc[k] = 0.f;
    const int r = warp + k * kSwWarps;
    const int ci = r % p.Ci, tap = r / p.Ci, ky = tap / p.kw, kx = tap - ky * p.kw;
    xoff[k] = r < rows ? (ky * patch_w + kx) * p.Ci + ci : 0;
  }
  const int step = p.stride * p.Ci;
  for (long long t = blockIdx.x; t < p.tiles; t += gridDim.x) {
    const int tx = (int)(t % p.tiles_x);
    const long long rr = t / p.tiles_x;
    const int oy = (int)(rr % p.Ho), n = (int)(rr / p.Ho);
    const int ox0 = tx * kSwTile;
    const int ix0 = ox0 * p.stride - p.pad_l, iy0 = oy * p.stride - p.pad_t;
    __syncthreads();                                                   // previous tile fully consumed
    for (int i = threadIdx.x; i < p.kh * patch_w * p.Ci; i += blockDim.x) {
      const int ci = i % p.Ci, px = (i / p.Ci) % patch_w, ky = i / (p.Ci * patch_w);
      const int iy = iy0 + ky, ix = ix0 + px;
      xs[i] = ((unsigned)iy < (unsigned)p.Hi && (unsigned)ix < (unsigned)p.Wi)
                  ? __ldg(p.x + ((size_t)(n * p.Hi + iy) * p.Wi + ix) * p.x_cs + ci) : 0.f;
    }
    for (int i = threadIdx.x; i < kSwTile * 32; i += blockDim.x) {
      const int co = i & 31, px = i >> 5;
      dzs[i] = (ox0 + px < p.Wo) ? __ldg(p.dz + ((size_t)(n * p.Ho + oy) * p.Wo + ox0 + px) * p.dz_cs + co) : 0.f;
    }
    __syncthreads();
#pragma unroll 4
    for (int px = 0; px < kSwTile; ++px) {
      const float g = dzs[px * 32 + lane];
      const float* xp = xs + px * step;
#pragma unroll
      for (int k = 0; k < kMaxPerWarp; ++k) acc[k] = fmaf(xp[xoff[k]], g, acc[k]);   // broadcast read: one address per warp
    }
  }
#pragma unroll
  for (int k = 0; k < kMaxPerWarp; ++k) {
    const int r = warp + k * kSwWarps;
    if (r < rows) atomicAdd(p.dw + (size_t)r * 32 + lane, acc[k]);
  }
}

}  // namespace lsi

using namespace lsi;

extern "C" int lsi_b200_conv2d_thin_supported(const lsi_b200_conv_desc* d) {
  if (!d) return 0;
  return (d->stride == 1 && (d->mode == 0 || d->mode == 1) && d->c_in >= 1 && d->c_in <= kThinMaxCi && d->c_out >= 1 &&
          d->c_out <= kThinCo && d->kh * d->kw <= kThinMaxTaps && d->epilogue == 0) ? 1 : 0;
}

// lsi_b200_conv2d for thin contractions (c_in <= 8, c_out <= 32, unit stride, <= 9 taps, no epilogue): same descriptor
// semantics, one thread per output pixel.
extern "C" int lsi_b200_conv2d_thin(const lsi_b200_conv_desc* d, const float* in, const float* w, float* out, void* stream) {
  LSI_REQUIRE(d && in && w && out, "NULL pointer argument");
  LSI_REQUIRE(lsi_b200_conv2d_thin_supported(d), "shape not supported by the thin-contraction kernel");
  ThinParams p{in, w, out, d->batch, d->h_in, d->w_in, d->c_in, d->h_out, d->w_out, d->c_out, d->kh, d->kw, d->pad_top, d->pad_left,
               d->mode, d->w_tap_stride, d->w_ci_stride, d->w_co_stride, d->in_c_stride, d->out_c_stride, d->accumulate};
  const long long total = (long long)d->batch * d->h_out * d->w_out;
  long long grid = (total + 127) / 128;
  if (grid > 148 * 16) grid = 148 * 16;
  const bool quad = d->c_in == 4 && d->c_out == kThinCo && d->kh == 3 && d->kw == 3 && d->in_c_stride == 4 && (d->out_c_stride & 3) == 0 &&
                    ((uintptr_t)in & 15) == 0 && ((uintptr_t)out & 15) == 0;
  {
    ScopedTiming tm(kConvFp32, as_stream(stream));
    if (quad) conv_thin4_kernel<<<(unsigned)grid, 128, 0, as_stream(stream)>>>(p);
    else conv_thin_kernel<<<(unsigned)grid, 128, 0, as_stream(stream)>>>(p);
  }
  LSI_LAUNCH_CHECK();
  return LSI_B200_OK;
}

extern "C" int lsi_b200_conv2d_stem_wgrad_supported(const lsi_b200_conv_desc* d) {
  if (!d) return 0;
  return (d->mode == 0 && d->c_in >= 1 && d->c_in <= 4 && d->c_out == 32 && d->kh <= 7 && d->kw <= 7 && d->stride >= 1 &&
          d->stride <= 2) ? 1 : 0;
}

// lsi_b200_conv2d_wgrad for the stem shape family (c_in <= 4, c_out == 32, <= 7x7 taps, stride 1 or 2): dw [kh,kw,c_in,32] dense.
extern "C" int lsi_b200_conv2d_stem_wgrad(const lsi_b200_conv_desc* d, const float* big, const float* small, float* dw, void* stream) {
  LSI_REQUIRE(d && big && small && dw, "NULL pointer argument");
  LSI_REQUIRE(lsi_b200_conv2d_stem_wgrad_supported(d), "shape not supported by the stem weight-gradient kernel");
  cudaStream_t st = as_stream(stream);
  StemWgradParams p{big, small, dw, d->batch, d->h_in, d->w_in, d->c_in, d->h_out, d->w_out, d->kh, d->kw, d->stride, d->pad_top,
                    d->pad_left, d->in_c_stride, d->out_c_stride, 0, 0};
  p.tiles_x = (d->w_out + kSwTile - 1) / kSwTile;
  p.tiles = (long long)d->batch * d->h_out * p.tiles_x;
  if (!d->accumulate) LSI_CUDA(cudaMemsetAsync(dw, 0, (size_t)d->kh * d->kw * d->c_in * 32 * sizeof(float), st));
  const int patch_w = (kSwTile - 1) * d->stride + d->kw;
  const size_t smem = ((size_t)d->kh * patch_w * d->c_in + kSwTile * 32) * sizeof(float);
  long long grid = 148 * 4;
  if (grid > p.tiles) grid = p.tiles;
  {
    ScopedTiming tm(kWgrad, st);
    stem_wgrad_kernel<<<(unsigned)grid, kSwWarps * 32, smem, st>>>(p);
  }
  LSI_LAUNCH_CHECK();
  return LSI_B200_OK;
}
