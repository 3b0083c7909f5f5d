// Convolution engine of the encoder-decoder CNN (lsi/nnutils/nets.py:244-348, :73-161 in the reference tree), fp32 NHWC.
//
// One implicit-GEMM gather kernel covers conv forward, transposed-conv forward and both data gradients:
//   out[n,oy,ox,co] = sum_{ky,kx,ci} in[n,iy,ix,ci] * W(ky,kx,ci,co)
//     gather mode 0 (conv fwd, convT dgrad):  iy = oy*stride - pad + ky
//     gather mode 1 (convT fwd, conv dgrad):  iy = (oy + pad - ky) / stride  when divisible
// with the weight tensor addressed through (tap, ci, co) strides, so TF's [kh,kw,cin,cout] conv layout and
// [kh,kw,cout,cin] transposed-conv layout (and their transposes for the gradients) need no re-layout.  Mode 1 is
// phase-decomposed: a CTA tile holds output pixels of one (oy % stride, ox % stride) phase and visits only the taps
// that are valid for it (4 of 16 for the 4x4 stride-2 up-convolution).
// A second kernel computes weight gradients (reduction over pixels, split across CTAs, fp32 atomics).
//
// These are CUDA-core (FFMA) kernels: the first correct path for every layer shape of the network (3-channel stem,
// 4-channel sigmoid head, 7x7/5x5, stride 2, up-convolutions).  The tcgen05/TMA kernels for the 3x3 stride-1 layers
// that carry the FLOPs replace them layer by layer.
#include "capi_common.h"
#include "common.cuh"

namespace lsi {

struct ConvParams {
  const float* in; const float* w; const float* bias; float* out;
  int N, Hi, Wi, Ci, Ho, Wo, Co;
  int kh, kw, stride, pad_t, pad_l, mode;
  int w_tap, w_ci, w_co;    // element strides of the weight tensor
  int in_cs, out_cs;        // pixel strides (>= Ci / Co): channel slices of wider buffers (concat) without copies
  int epilogue;             // 0: none, 1: + bias, 2: sigmoid(. + bias)
  int accumulate;           // out += result
  int Hp, Wp;               // per-phase output extent (mode 1), else Ho, Wo
};

constexpr int kBM = 128, kBK = 16;

template <int BN, int TN>
__global__ void __launch_bounds__(256) conv_gather_kernel(const ConvParams p) {
  constexpr int TM = 8;
  __shared__ __align__(16) float As[kBK][kBM + 4];
  __shared__ __align__(16) float Bs[kBK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * kBM, n0 = blockIdx.y * BN;
  const int s = (p.mode == 1) ? p.stride : 1;
  const int py = (p.mode == 1) ? (int)blockIdx.z / s : 0, px = (p.mode == 1) ? (int)blockIdx.z % s : 0;
  const int M = p.N * p.Hp * p.Wp;

  // the pixel this thread gathers for the A tile
  const int lp = tid & 127, lhalf = tid >> 7;
  const int lm = m0 + lp;
  const bool lvalid = lm < M;
  int ln = 0, loy = 0, lox = 0;
  if (lvalid) {
    ln = lm / (p.Hp * p.Wp);
    const int r = lm - ln * p.Hp * p.Wp;
    loy = r / p.Wp; lox = r - loy * p.Wp;
    if (p.mode == 1) { loy = loy * s + py; lox = lox * s + px; }
  }
  // B tile: thread loads Bs[bk][bn4..bn4+3]
  const int bk = tid / (BN / 4), bn4 = (tid % (BN / 4)) * 4;
  const bool b_loader = bk < kBK;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float a_reg[8], b_reg[4];
  const int cchunks = (p.Ci + kBK - 1) / kBK;
  // tap list: mode 0 -> all taps; mode 1 -> taps with (oy + pad - ky) % stride == 0 for this phase
  const int ky0 = (p.mode == 1) ? ((py + p.pad_t) % s) : 0, kx0 = (p.mode == 1) ? ((px + p.pad_l) % s) : 0;
  const int nky = (p.kh - ky0 + s - 1) / s, nkx = (p.kw - kx0 + s - 1) / s;
  const int ksteps = nky * nkx * cchunks;

  auto load_tiles = [&](int step) {
    const int t = step / cchunks, c0 = (step - t * cchunks) * kBK;
    const int ky = ky0 + (t / nkx) * s, kx = kx0 + (t % nkx) * s;
    int iy, ix;
    if (p.mode == 0) { iy = loy * p.stride - p.pad_t + ky; ix = lox * p.stride - p.pad_l + kx; }
    else { iy = (loy + p.pad_t - ky) / s; ix = (lox + p.pad_l - kx) / s; }   // exact by construction of the tap list
    const bool ok = lvalid && iy >= 0 && iy < p.Hi && ix >= 0 && ix < p.Wi &&
                    (p.mode == 0 || (loy + p.pad_t - ky >= 0 && lox + p.pad_l - kx >= 0));
    const int c = c0 + lhalf * 8;
    if (ok) {
      const float* src = p.in + ((size_t)(ln * p.Hi + iy) * p.Wi + ix) * p.in_cs + c;
      if (c + 8 <= p.Ci && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
        const float4 v0 = __ldg(reinterpret_cast<const float4*>(src)), v1 = __ldg(reinterpret_cast<const float4*>(src) + 1);
        a_reg[0] = v0.x; a_reg[1] = v0.y; a_reg[2] = v0.z; a_reg[3] = v0.w;
        a_reg[4] = v1.x; a_reg[5] = v1.y; a_reg[6] = v1.z; a_reg[7] = v1.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) a_reg[j] = (c + j < p.Ci) ? __ldg(src + j) : 0.f;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) a_reg[j] = 0.f;
    }
    if (b_loader) {
      const int ci = c0 + bk;
      const float* wsrc = p.w + (size_t)(ky * p.kw + kx) * p.w_tap + (size_t)ci * p.w_ci;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int co = n0 + bn4 + j;
        b_reg[j] = (ci < p.Ci && co < p.Co) ? __ldg(wsrc + (size_t)co * p.w_co) : 0.f;
      }
    }
  };
  auto store_tiles = [&]() {
#pragma unroll
    for (int j = 0; j < 8; ++j) As[lhalf * 8 + j][lp] = a_reg[j];
    if (b_loader) *reinterpret_cast<float4*>(&Bs[bk][bn4]) = make_float4(b_reg[0], b_reg[1], b_reg[2], b_reg[3]);
  };

  if (ksteps > 0) { load_tiles(0); store_tiles(); }
  __syncthreads();
  for (int step = 0; step < ksteps; ++step) {
    if (step + 1 < ksteps) load_tiles(step + 1);
#pragma unroll
    for (int k = 0; k < kBK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * TM]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * TM + 4]);
      const float a[TM] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[TN];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[k][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
    if (step + 1 < ksteps) store_tiles();
    __syncthreads();
  }

  // epilogue
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty * TM + i;
    if (m >= M) continue;
    const int n = m / (p.Hp * p.Wp);
    const int r = m - n * p.Hp * p.Wp;
    int oy = r / p.Wp, ox = r - oy * p.Wp;
    if (p.mode == 1) { oy = oy * s + py; ox = ox * s + px; }
    float* dst = p.out + ((size_t)(n * p.Ho + oy) * p.Wo + ox) * p.out_cs + n0 + tx * TN;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int co = n0 + tx * TN + j;
      if (co >= p.Co) continue;
      float v = acc[i][j];
      if (p.epilogue >= 1) v += __ldg(p.bias + co);
      if (p.epilogue == 2) v = 1.f / (1.f + expf(-v));
      if (p.accumulate) v += dst[j];
      dst[j] = v;
    }
  }
}

// dw[tap][a][b] (+)= sum_{n,oy,ox} big[n, oy*stride - pad + ky, ox*stride - pad + kx, a] * small[n, oy, ox, b]
// conv:  big = layer input (a = cin),  small = dout (b = cout)  -> dw in TF conv layout  [kh,kw,cin,cout]
// convT: big = dout (a = cout),        small = layer input (b = cin) -> TF transposed layout [kh,kw,cout,cin]
struct WgradParams {
  const float* big; const float* small; float* dw;
  int N, Hb, Wb, Ca, Hs, Ws, Cb;
  int kh, kw, stride, pad_t, pad_l;
  int big_cs, small_cs;
  int pix_per_split;
};

__global__ void __launch_bounds__(256) conv_wgrad_kernel(const WgradParams p) {
  constexpr int BA = 64, BB = 64, BKP = 16;
  __shared__ __align__(16) float As[BKP][BA + 4];
  __shared__ __align__(16) float Bs[BKP][BB + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int tiles_b = (p.Cb + BB - 1) / BB;
  const int a0 = ((int)blockIdx.x / tiles_b) * BA, b0 = ((int)blockIdx.x % tiles_b) * BB;
  const int tap = blockIdx.y, ky = tap / p.kw, kx = tap % p.kw;
  const long long M = (long long)p.N * p.Hs * p.Ws;
  const long long m_begin = (long long)blockIdx.z * p.pix_per_split;
  const long long m_end = (m_begin + p.pix_per_split < M) ? m_begin + p.pix_per_split : M;
  const int lk = tid >> 4, lc4 = (tid & 15) * 4;   // loader: pixel lk of the chunk, channels lc4..lc4+3
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (long long mc = m_begin; mc < m_end; mc += BKP) {
    const long long m = mc + lk;
    float4 av = make_float4(0.f, 0.f, 0.f, 0.f), bv = av;
    if (m < m_end) {
      const int n = (int)(m / ((long long)p.Hs * p.Ws));
      const int r = (int)(m - (long long)n * p.Hs * p.Ws);
      const int oy = r / p.Ws, ox = r - oy * p.Ws;
      const int iy = oy * p.stride - p.pad_t + ky, ix = ox * p.stride - p.pad_l + kx;
      const float* bsrc = p.small + (size_t)m * p.small_cs + b0 + lc4;
      bv.x = (b0 + lc4 + 0 < p.Cb) ? __ldg(bsrc + 0) : 0.f; bv.y = (b0 + lc4 + 1 < p.Cb) ? __ldg(bsrc + 1) : 0.f;
      bv.z = (b0 + lc4 + 2 < p.Cb) ? __ldg(bsrc + 2) : 0.f; bv.w = (b0 + lc4 + 3 < p.Cb) ? __ldg(bsrc + 3) : 0.f;
      if (iy >= 0 && iy < p.Hb && ix >= 0 && ix < p.Wb) {
        const float* asrc = p.big + ((size_t)(n * p.Hb + iy) * p.Wb + ix) * p.big_cs + a0 + lc4;
        av.x = (a0 + lc4 + 0 < p.Ca) ? __ldg(asrc + 0) : 0.f; av.y = (a0 + lc4 + 1 < p.Ca) ? __ldg(asrc + 1) : 0.f;
        av.z = (a0 + lc4 + 2 < p.Ca) ? __ldg(asrc + 2) : 0.f; av.w = (a0 + lc4 + 3 < p.Ca) ? __ldg(asrc + 3) : 0.f;
      }
    }
    __syncthreads();
    *reinterpret_cast<float4*>(&As[lk][lc4]) = av;
    *reinterpret_cast<float4*>(&Bs[lk][lc4]) = bv;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BKP; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float aa[4] = {a.x, a.y, a.z, a.w}, bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
    }
  }
  float* dst = p.dw + (size_t)tap * p.Ca * p.Cb;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int a = a0 + ty * 4 + i;
    if (a >= p.Ca) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int b = b0 + tx * 4 + j;
      if (b < p.Cb) atomicAdd(dst + (size_t)a * p.Cb + b, acc[i][j]);
    }
  }
}

// Direct convolution for the 3-channel stem (nets.py:273: 7x7 stride 2, 3 -> 32): one thread per output pixel holding
// all 32 output channels in registers, the whole filter bank (kh*kw*Cin x 32 floats, 18.8 KB for the stem) in shared
// memory and read as broadcast LDS.128.  TMA cannot feed this layer (12-byte pixel stride) and the generic gather kernel
// wastes 13/16 of its K tile on it (Cin = 3): 6.5 ms -> ~1 ms per 64-image step.
struct StemParams {
  const float* in; const float* w; float* out;
  int N, Hi, Wi, Ci, Ho, Wo, kh, kw, stride, pad_t, pad_l, in_cs, out_cs;
  int w_tap, w_ci, w_co;
};

__global__ void __launch_bounds__(128) stem_conv_kernel(const StemParams p) {
  extern __shared__ __align__(16) float wsm[];          // [kh*kw*Ci][32]
  const int K = p.kh * p.kw * p.Ci;
  for (int i = threadIdx.x; i < K * 32; i += blockDim.x) {
    const int k = i >> 5, co = i & 31;
    const int tap = k / p.Ci, ci = k - tap * p.Ci;
    wsm[i] = p.w[(size_t)tap * p.w_tap + (size_t)ci * p.w_ci + (size_t)co * p.w_co];
  }
  __syncthreads();
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long M = (long long)p.N * p.Ho * p.Wo;
  if (m >= M) return;
  const int n = (int)(m / ((long long)p.Ho * p.Wo));
  const int r = (int)(m - (long long)n * p.Ho * p.Wo);
  const int oy = r / p.Wo, ox = r - oy * p.Wo;
  float acc[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) acc[c] = 0.f;
  const int iy0 = oy * p.stride - p.pad_t, ix0 = ox * p.stride - p.pad_l;
  for (int ky = 0; ky < p.kh; ++ky) {
    const int iy = iy0 + ky;
    if (iy < 0 || iy >= p.Hi) continue;
    for (int kx = 0; kx < p.kw; ++kx) {
      const int ix = ix0 + kx;
      if (ix < 0 || ix >= p.Wi) continue;
      const float* src = p.in + ((size_t)(n * p.Hi + iy) * p.Wi + ix) * p.in_cs;
      const float* wk = wsm + (size_t)((ky * p.kw + kx) * p.Ci) * 32;
      for (int ci = 0; ci < p.Ci; ++ci) {
        const float a = __ldg(src + ci);
        const float4* w4 = reinterpret_cast<const float4*>(wk + ci * 32);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 wv = w4[q];
          acc[4 * q] = fmaf(a, wv.x, acc[4 * q]); acc[4 * q + 1] = fmaf(a, wv.y, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(a, wv.z, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(a, wv.w, acc[4 * q + 3]);
        }
      }
    }
  }
  float4* dst = reinterpret_cast<float4*>(p.out + (size_t)m * p.out_cs);
#pragma unroll
  for (int q = 0; q < 8; ++q) dst[q] = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
}

}  // namespace lsi

using namespace lsi;

static int check_conv(const lsi_b200_conv_desc* d) {
  LSI_REQUIRE(d != nullptr, "descriptor is NULL");
  LSI_REQUIRE(d->batch >= 1 && d->h_in >= 1 && d->w_in >= 1 && d->c_in >= 1 && d->h_out >= 1 && d->w_out >= 1 && d->c_out >= 1,
              "conv sizes must be >= 1");
  LSI_REQUIRE(d->kh >= 1 && d->kw >= 1 && d->stride >= 1 && d->stride <= 2, "kernel/stride out of range");
  LSI_REQUIRE(d->mode == 0 || d->mode == 1, "mode must be 0 or 1");
  LSI_REQUIRE(d->in_c_stride >= d->c_in && d->out_c_stride >= d->c_out, "pixel strides smaller than channel counts");
  LSI_REQUIRE((long long)d->batch * d->h_out * d->w_out < (1ll << 31) && (long long)d->batch * d->h_in * d->w_in < (1ll << 31),
              "too many pixels for 32-bit indices");
  if (d->mode == 1) LSI_REQUIRE(d->h_out % d->stride == 0 && d->w_out % d->stride == 0, "mode 1 needs output sizes divisible by the stride");
  return LSI_B200_OK;
}

extern "C" int lsi_b200_conv2d(const lsi_b200_conv_desc* d, const float* in, const float* w, const float* bias, float* out,
                               void* stream) {
  if (int rc = check_conv(d)) return rc;
  LSI_REQUIRE(in && w && out, "NULL pointer argument");
  LSI_REQUIRE(d->epilogue == 0 || bias, "epilogue needs a bias pointer");
  ConvParams p;
  p.in = in; p.w = w; p.bias = bias; p.out = out;
  p.N = d->batch; p.Hi = d->h_in; p.Wi = d->w_in; p.Ci = d->c_in; p.Ho = d->h_out; p.Wo = d->w_out; p.Co = d->c_out;
  p.kh = d->kh; p.kw = d->kw; p.stride = d->stride; p.pad_t = d->pad_top; p.pad_l = d->pad_left; p.mode = d->mode;
  p.w_tap = d->w_tap_stride; p.w_ci = d->w_ci_stride; p.w_co = d->w_co_stride;
  p.in_cs = d->in_c_stride; p.out_cs = d->out_c_stride; p.epilogue = d->epilogue; p.accumulate = d->accumulate;
  const int s = d->mode == 1 ? d->stride : 1;
  p.Hp = d->h_out / s; p.Wp = d->w_out / s;
  const long long M = (long long)p.N * p.Hp * p.Wp;
  if (d->mode == 0 && d->c_in <= 4 && d->c_out == 32 && d->epilogue == 0 && !d->accumulate && (d->out_c_stride & 3) == 0 &&
      ((uintptr_t)out & 15) == 0 && (size_t)d->kh * d->kw * d->c_in * 32 * 4 <= 48 * 1024) {
    StemParams sp{in, w, out, d->batch, d->h_in, d->w_in, d->c_in, d->h_out, d->w_out, d->kh, d->kw, d->stride, d->pad_top,
                  d->pad_left, d->in_c_stride, d->out_c_stride, d->w_tap_stride, d->w_ci_stride, d->w_co_stride};
    {
      ScopedTiming tm(kConvFp32, as_stream(stream));
      stem_conv_kernel<<<(unsigned)((M + 127) / 128), 128, (size_t)d->kh * d->kw * d->c_in * 32 * 4, as_stream(stream)>>>(sp);
    }
    LSI_LAUNCH_CHECK();
    return LSI_B200_OK;
  }
  const unsigned gm = (unsigned)((M + kBM - 1) / kBM);
  {
    ScopedTiming tm(kConvFp32, as_stream(stream));
    if (d->c_out > 32) {
      dim3 grid(gm, (d->c_out + 63) / 64, s * s);
      conv_gather_kernel<64, 4><<<grid, 256, 0, as_stream(stream)>>>(p);
    } else {
      dim3 grid(gm, 1, s * s);
      conv_gather_kernel<32, 2><<<grid, 256, 0, as_stream(stream)>>>(p);
    }
  }
  LSI_LAUNCH_CHECK();
  return LSI_B200_OK;
}

extern "C" int lsi_b200_conv2d_wgrad(const lsi_b200_conv_desc* d, const float* big, const float* small, float* dw,
                                     void* stream) {
  // the descriptor describes the strided-gather side: `big` = [batch,h_in,w_in,c_in] (pixel stride in_c_stride),
  // `small` = [batch,h_out,w_out,c_out] (pixel stride out_c_stride), small(oy) pairs with big(oy*stride - pad + ky)
  if (int rc = check_conv(d)) return rc;
  LSI_REQUIRE(big && small && dw, "NULL pointer argument");
  LSI_REQUIRE(d->mode == 0, "wgrad is expressed in gather mode 0");
  WgradParams p;
  p.big = big; p.small = small; p.dw = dw;
  p.N = d->batch; p.Hb = d->h_in; p.Wb = d->w_in; p.Ca = d->c_in; p.Hs = d->h_out; p.Ws = d->w_out; p.Cb = d->c_out;
  p.kh = d->kh; p.kw = d->kw; p.stride = d->stride; p.pad_t = d->pad_top; p.pad_l = d->pad_left;
  p.big_cs = d->in_c_stride; p.small_cs = d->out_c_stride;
  const long long M = (long long)p.N * p.Hs * p.Ws;
  const int tiles = ((p.Ca + 63) / 64) * ((p.Cb + 63) / 64);
  const int taps = p.kh * p.kw;
  long long splits = (148 * 8 + (long long)tiles * taps - 1) / ((long long)tiles * taps);   // aim for ~8 CTAs per SM
  const long long max_splits = (M + 255) / 256;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  long long pps = (M + splits - 1) / splits;
  pps = (pps + 15) / 16 * 16;
  splits = (M + pps - 1) / pps;
  p.pix_per_split = (int)pps;
  if (!d->accumulate) LSI_CUDA(cudaMemsetAsync(dw, 0, (size_t)taps * p.Ca * p.Cb * sizeof(float), as_stream(stream)));
  dim3 grid(tiles, taps, (unsigned)splits);
  {
    ScopedTiming tm(kWgrad, as_stream(stream));
    conv_wgrad_kernel<<<grid, 256, 0, as_stream(stream)>>>(p);
  }
  LSI_LAUNCH_CHECK();
  return LSI_B200_OK;
}
