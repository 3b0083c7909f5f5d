// Normalisation / activation / optimiser kernels of the CNN slice (fp32, NHWC, channel-last so that a warp reads
// consecutive channels of consecutive pixels).
//   batch-stat BN + ReLU  slim.batch_norm(center=True, scale=False, eps=1e-3, is_training=True) (nets.py:263-272):
//                          y = relu((x - mean_B) * rsqrt(var_B + eps) + beta), biased variance over (N,H,W)
//   sigmoid head           pixelwise_predictor's activation (nets.py:143); forward is fused into the conv epilogue
//   Adam                   tf.train.AdamOptimizer(lr, beta1) (train_utils.py:112): m,v update with the
//                          lr*sqrt(1-b2^t)/(1-b1^t) step and epsilon outside the square root
#include <cuda_fp16.h>

#include "capi_common.h"
#include "common.cuh"
#include "split.cuh"

namespace lsi {

constexpr int kStatBlocks = 512;

// partial[blk][c] = (sum_a, sum_b) over the block's pixel range; a/b chosen by mode:
//   mode 0 (forward stats):   a = x,            b = x*x
//   mode 1 (backward stats):  a = dz,           b = dz * xhat    with dz = dy * [y > 0], xhat = (x - mean) * invstd
struct StatParams {
  const float* x; const float* y; const float* dy; const float* stats;   // stats [C][2] = (mean, invstd)
  double* partial;                                                          // [kStatBlocks][C][2]
  long long P; int C, x_cs, y_cs, dy_cs, mode, relu;
};

__global__ void __launch_bounds__(256) channel_stats_kernel(const StatParams p) {
  // thread -> (pixel lane, channel): channels fastest so that loads coalesce
  const int lanes = 256 / min(p.C, 256) > 0 ? 256 / min(p.C, 256) : 1;
  const int cpt = (p.C + 255) / 256;                 // channels per thread when C > 256
  const int c_lo = (p.C >= 256) ? threadIdx.x : (int)threadIdx.x % p.C;
  const int lane = (p.C >= 256) ? 0 : (int)threadIdx.x / p.C;
  const bool active = (p.C >= 256) || lane < lanes;
  const long long per_blk = (p.P + gridDim.x - 1) / gridDim.x;
  const long long p0 = (long long)blockIdx.x * per_blk, p1 = (p0 + per_blk < p.P) ? p0 + per_blk : p.P;
  for (int k = 0; k < cpt; ++k) {
    const int c = c_lo + k * 256;
    double sa = 0.0, sb = 0.0;
    if (active && c < p.C) {
      float mean = 0.f, invstd = 0.f;
      if (p.mode == 1) { mean = p.stats[2 * c]; invstd = p.stats[2 * c + 1]; }
      float fa = 0.f, fb = 0.f; int cnt = 0;
      for (long long q = p0 + lane; q < p1; q += lanes) {
        const float xv = p.x[q * p.x_cs + c];
        if (p.mode == 0) { fa += xv; fb = fmaf(xv, xv, fb); }
        else {
          float dz = p.dy[q * p.dy_cs + c];
          if (p.relu && !(p.y[q * p.y_cs + c] > 0.f)) dz = 0.f;
          fa += dz; fb = fmaf(dz, (xv - mean) * invstd, fb);
        }
        if (++cnt == 256) { sa += fa; sb += fb; fa = fb = 0.f; cnt = 0; }   // fp32 runs of 256, fp64 across runs
      }
      sa += fa; sb += fb;
    }
    // combine the pixel lanes that share a channel (C < 256): shared-memory tree over lanes
    __shared__ double sh[2][256];
    sh[0][threadIdx.x] = sa; sh[1][threadIdx.x] = sb;
    __syncthreads();
    if (p.C < 256) {
      if (lane == 0 && c < p.C) {
        for (int l = 1; l < lanes; ++l) { sa += sh[0][l * p.C + c]; sb += sh[1][l * p.C + c]; }
      }
    }
    if (lane == 0 && c < p.C && active) {
      p.partial[((size_t)blockIdx.x * p.C + c) * 2] = sa;
      p.partial[((size_t)blockIdx.x * p.C + c) * 2 + 1] = sb;
    }
    __syncthreads();
  }
}

// forward: stats[c] = (mean, rsqrt(var + eps)); backward: stats_out[c] = (sum dz, sum dz*xhat).  One warp per channel.
__global__ void __launch_bounds__(256) finalize_stats_kernel(const double* __restrict__ partial, int nblk, int C, long long P,
                                                             float eps, int mode, float* __restrict__ out) {
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (c >= C) return;
  double a = 0.0, b = 0.0;
  for (int i = lane; i < nblk; i += 32) { a += partial[((size_t)i * C + c) * 2]; b += partial[((size_t)i * C + c) * 2 + 1]; }
  for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
  if (lane != 0) return;
  if (mode == 0) {
    const double mean = a / (double)P;
    double var = b / (double)P - mean * mean;
    if (var < 0.0) var = 0.0;
    out[2 * c] = (float)mean; out[2 * c + 1] = (float)(1.0 / sqrt(var + (double)eps));
  } else if (mode == 2) {      // raw fp64 sums (synchronised batch norm: the ranks' sums are all-reduced before mean / variance)
    reinterpret_cast<double*>(out)[2 * c] = a; reinterpret_cast<double*>(out)[2 * c + 1] = b;
  } else {
    out[2 * c] = (float)a; out[2 * c + 1] = (float)b;
  }
}

struct BnApplyParams {
  const float* x; const float* stats; const float* beta; float* y;
  long long P; int C, x_cs, y_cs, relu;
};

__global__ void __launch_bounds__(256) bn_apply_kernel(const BnApplyParams p) {
  const long long total = p.P * p.C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long q = i / p.C; const int c = (int)(i - q * p.C);
    float v = fmaf(p.x[q * p.x_cs + c] - __ldg(p.stats + 2 * c), __ldg(p.stats + 2 * c + 1), __ldg(p.beta + c));
    if (p.relu) v = fmaxf(v, 0.f);
    p.y[q * p.y_cs + c] = v;
  }
}

// fast path (x dense: x_cs == C; y dense or a channel slice of a wider tensor: y_cs % 4 == 0; C % 4 == 0, 16-byte aligned):
// 128-bit loads/stores, 32-bit channel arithmetic
__global__ void __launch_bounds__(256) bn_apply_vec4_kernel(const BnApplyParams p) {
  const long long total4 = p.P * p.C / 4;
  const unsigned c4n = (unsigned)p.C / 4;
  const float4* x = reinterpret_cast<const float4*>(p.x);
  float4* y = reinterpret_cast<float4*>(p.y);
  const unsigned y_s4 = (unsigned)p.y_cs / 4;      // == c4n when y is dense; larger when y is a channel slice of a wider tensor
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    const unsigned long long q = (unsigned long long)i / c4n;
    const unsigned c4 = (unsigned)((unsigned long long)i - q * c4n);
    const int c = (int)c4 * 4;
    const float4 v = __ldcs(x + i);
    const float4 s0 = __ldg(reinterpret_cast<const float4*>(p.stats + 2 * c)), s1 = __ldg(reinterpret_cast<const float4*>(p.stats + 2 * c) + 1);
    const float4 b = __ldg(reinterpret_cast<const float4*>(p.beta + c));
    float4 o;
    o.x = fmaf(v.x - s0.x, s0.y, b.x); o.y = fmaf(v.y - s0.z, s0.w, b.y);      // explicit fma: bn_bwd_z_* recompute the ReLU mask
    o.z = fmaf(v.z - s1.x, s1.y, b.z); o.w = fmaf(v.w - s1.z, s1.w, b.w);      // from z with exactly this expression
    if (p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
    y[q * y_s4 + c4] = o;
  }
}

// dx = invstd * (dz - mean(dz) - xhat * mean(dz * xhat)),  dz = dy * [y > 0];  dbeta = sum dz (from the stats pass)
struct BnBwdParams {
  const float* x; const float* y; const float* dy; const float* stats; const float* sums; float* dx;
  long long P; int C, x_cs, y_cs, dy_cs, dx_cs, relu, accumulate;
  long long P_stat;      // pixels the statistics (and `sums`) were reduced over: P, or the global count under synchronised BN
};

__global__ void __launch_bounds__(256) bn_backward_kernel(const BnBwdParams p) {
  const long long total = p.P * p.C;
  const float invP = 1.f / (float)p.P_stat;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long q = i / p.C; const int c = (int)(i - q * p.C);
    const float mean = __ldg(p.stats + 2 * c), invstd = __ldg(p.stats + 2 * c + 1);
    float dz = p.dy[q * p.dy_cs + c];
    if (p.relu && !(p.y[q * p.y_cs + c] > 0.f)) dz = 0.f;
    const float xhat = (p.x[q * p.x_cs + c] - mean) * invstd;
    float v = invstd * (dz - __ldg(p.sums + 2 * c) * invP - xhat * __ldg(p.sums + 2 * c + 1) * invP);
    if (p.accumulate) v += p.dx[q * p.dx_cs + c];
    p.dx[q * p.dx_cs + c] = v;
  }
}

// dst[q][0..C) (+)= src[q][0..C)   with independent pixel strides (concat / gradient split without torch ops)
__global__ void __launch_bounds__(256) copy_channels_kernel(const float* __restrict__ src, float* __restrict__ dst, long long P,
                                                            int C, int src_cs, int dst_cs, int accumulate) {
  const long long total = P * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long q = i / C; const int c = (int)(i - q * C);
    const float v = src[q * src_cs + c];
    if (accumulate) dst[q * dst_cs + c] += v; else dst[q * dst_cs + c] = v;
  }
}

// the same for 4-channel-aligned, 16-byte-aligned tensors: 128-bit accesses, 32-bit index arithmetic per row
__global__ void __launch_bounds__(256) copy_channels_vec4_kernel(const float4* __restrict__ src, float4* __restrict__ dst, long long P,
                                                                 int C4, int src_cs4, int dst_cs4, int accumulate) {
  const long long total = P * C4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long q = i / C4; const int c = (int)(i - q * C4);
    float4 v = __ldcs(src + q * src_cs4 + c);
    if (accumulate) { const float4 a = dst[q * dst_cs4 + c]; v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w; }
    dst[q * dst_cs4 + c] = v;
  }
}

// Dense batch-norm + ReLU backward from the raw conv output z and the upstream gradient alone: the ReLU mask is recomputed as
// fmaf(z - mean, rstd, beta) > 0 -- the very expression the forward kernels evaluate -- so the normalised output y is not read.
//   stats pass   partial[blk][c] = (sum dz, sum dz * xhat) over the block's items        (2 tensor reads)
//   apply pass   dx = rstd * (dz - sum_dz / P - xhat * sum_dzx / P)                        (2 reads, 1 write)
// against 3 + 3 reads and 1 write of the strided general kernels.  Items are float4s of the flattened [P, C] tensors; the grid
// stride is a multiple of C / 4, so a thread's four channels -- mean, rstd, beta and the two sums -- are loop invariants.
struct BnBwdZParams {
  const float4* z; const float4* dy; float4* dx;
  const float* beta; const float* stats; const float* sums; double* partial;
  long long total4, P; unsigned c4n; int C;
  unsigned dy_s4;   // pixel stride of dy in float4 units (c4n when dy is dense; larger when dy is a channel slice of a concat gradient)
};

__global__ void __launch_bounds__(256) bn_bwd_z_stats_kernel(const BnBwdZParams p) {
  __shared__ float sh[256][9];
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int c = (int)((unsigned long long)i0 % p.c4n) * 4;
  float mean[4], rstd[4], beta[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) { mean[k] = __ldg(p.stats + 2 * (c + k)); rstd[k] = __ldg(p.stats + 2 * (c + k) + 1); beta[k] = __ldg(p.beta + c + k); }
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  // (the grid stride is a multiple of c4n: a thread keeps its channels, its pixel advances by a constant)
  long long j = (i0 / p.c4n) * p.dy_s4 + (i0 % p.c4n);
  const long long jstep = ((long long)gridDim.x * blockDim.x / p.c4n) * p.dy_s4;
  for (long long i = i0; i < p.total4; i += (long long)gridDim.x * blockDim.x, j += jstep) {
    const float4 zv = __ldcs(p.z + i), g = __ldcs(p.dy + j);
    const float zz[4] = {zv.x, zv.y, zv.z, zv.w}, gg[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float d = zz[k] - mean[k];
      const float dz = fmaf(d, rstd[k], beta[k]) > 0.f ? gg[k] : 0.f;
      s[k] += dz; s[4 + k] = fmaf(dz, d * rstd[k], s[4 + k]);
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) sh[threadIdx.x][k] = s[k];
  __syncthreads();
  // thread t < c4n gathers the threads t, t + c4n, ... of this block: they all hold the same four channels
  if (threadIdx.x < p.c4n) {
    double a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (unsigned t = threadIdx.x; t < 256; t += p.c4n)
#pragma unroll
      for (int k = 0; k < 8; ++k) a[k] += (double)sh[t][k];
    double* out = p.partial + ((size_t)blockIdx.x * p.C + c) * 2;
#pragma unroll
    for (int k = 0; k < 4; ++k) { out[2 * k] = a[k]; out[2 * k + 1] = a[4 + k]; }
  }
}

__global__ void __launch_bounds__(256) bn_bwd_z_apply_kernel(const BnBwdZParams p) {
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int c = (int)((unsigned long long)i0 % p.c4n) * 4;
  const float invP = 1.f / (float)p.P;
  float mean[4], rstd[4], beta[4], m0[4], m1[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    mean[k] = __ldg(p.stats + 2 * (c + k)); rstd[k] = __ldg(p.stats + 2 * (c + k) + 1); beta[k] = __ldg(p.beta + c + k);
    m0[k] = __ldg(p.sums + 2 * (c + k)) * invP; m1[k] = __ldg(p.sums + 2 * (c + k) + 1) * invP;
  }
  long long j = (i0 / p.c4n) * p.dy_s4 + (i0 % p.c4n);
  const long long jstep = ((long long)gridDim.x * blockDim.x / p.c4n) * p.dy_s4;
  for (long long i = i0; i < p.total4; i += (long long)gridDim.x * blockDim.x, j += jstep) {
    const float4 zv = __ldcs(p.z + i), g = __ldcs(p.dy + j);
    const float zz[4] = {zv.x, zv.y, zv.z, zv.w}, gg[4] = {g.x, g.y, g.z, g.w};
    float o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float d = zz[k] - mean[k];
      const float dz = fmaf(d, rstd[k], beta[k]) > 0.f ? gg[k] : 0.f;
      o[k] = rstd[k] * (dz - m0[k] - (d * rstd[k]) * m1[k]);
    }
    p.dx[i] = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// sigmoid head backward: dz = dy * y * (1 - y)  (y = sigmoid output), written in place of a fresh buffer
__global__ void __launch_bounds__(256) sigmoid_backward_kernel(const float* __restrict__ y, const float* __restrict__ dy,
                                                               float* __restrict__ dz, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = y[i];
    dz[i] = dy[i] * v * (1.f - v);
  }
}

struct AdamParams {
  float* p; const float* g; float* m; float* v;
  long long n; float lr_t, beta1, beta2, eps, grad_scale;
};

__global__ void __launch_bounds__(256) adam_kernel(const AdamParams a) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (long long)gridDim.x * blockDim.x) {
    const float g = a.g[i] * a.grad_scale;
    const float m = a.beta1 * a.m[i] + (1.f - a.beta1) * g;
    const float v = a.beta2 * a.v[i] + (1.f - a.beta2) * g * g;
    a.m[i] = m; a.v[i] = v;
    a.p[i] -= a.lr_t * m / (sqrtf(v) + a.eps);
  }
}

static unsigned ew_grid(long long n) {
  long long g = (n + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  if (g < 1) g = 1;
  return (unsigned)g;
}

// fp16 activations (inference-only 'f16' mode): y(half) = relu((x - mean) * rstd + beta), x stored as half or float;
// dense, C % 8 == 0, 8 channels per thread (128-bit half loads/stores)
template <bool kInF16>
__global__ void __launch_bounds__(256) bn_apply_h_kernel(const void* __restrict__ xv, const float* __restrict__ stats,
                                                         const float* __restrict__ beta, __half* __restrict__ y, long long total8,
                                                         unsigned c8n) {
  // the grid-stride (gridDim.x * 256) is a multiple of c8n (a power of two <= 256 or a divisor the host checked), so a
  // thread meets the same eight channels in every iteration: scale/shift are loop invariants in registers
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int c = (int)((unsigned long long)i0 % c8n) * 8;
  float sa[8], sb[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float mean = __ldg(stats + 2 * (c + k)), rstd = __ldg(stats + 2 * (c + k) + 1);
    sa[k] = rstd; sb[k] = fmaf(-mean, rstd, __ldg(beta + c + k));
  }
  for (long long i = i0; i < total8; i += (long long)gridDim.x * blockDim.x) {
    float v[8];
    if (kInF16) {
      const uint4 u = __ldcs(reinterpret_cast<const uint4*>(xv) + i);
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[k]));
        v[2 * k] = f.x; v[2 * k + 1] = f.y;
      }
    } else {
      const float4 a = __ldcs(reinterpret_cast<const float4*>(xv) + 2 * i), b = __ldcs(reinterpret_cast<const float4*>(xv) + 2 * i + 1);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const __half2 h = __floats2half2_rn(fmaxf(fmaf(v[2 * k], sa[2 * k], sb[2 * k]), 0.f),
                                          fmaxf(fmaf(v[2 * k + 1], sa[2 * k + 1], sb[2 * k + 1]), 0.f));
      o[k] = *reinterpret_cast<const uint32_t*>(&h);
    }
    reinterpret_cast<uint4*>(y)[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}


// Split-precision activations (split.cuh): one thread per (pixel, 32-channel chunk, 8-channel group): its hi values are the
// 16 bytes at ((pixel * C/32 + chunk) * 128 + group * 16), its lo values 64 bytes further.
//   kIn  0: x is fp32 [P, C]      1: x is a split tensor
//   kOut 0: y is fp32 [P, C]      1: y is a split tensor (may alias x when kIn == 1)
//   kBn  : y = relu((x - mean) * rstd + beta) (slim.batch_norm + ReLU, nets.py:263-272), else y = x
template <int kIn, int kOut, bool kBn>
__global__ void __launch_bounds__(256) split_convert_kernel(const void* __restrict__ xv, const float* __restrict__ stats,
                                                            const float* __restrict__ beta, void* __restrict__ yv, long long total8,
                                                            unsigned c8n) {
  // the grid stride is a multiple of c8n = C / 8, so a thread meets the same eight channels in every iteration
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int c = (int)((unsigned long long)i0 % c8n) * 8;
  float sa[8], sb[8];
  if (kBn) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float mean = __ldg(stats + 2 * (c + k)), rstd = __ldg(stats + 2 * (c + k) + 1);
      sa[k] = rstd; sb[k] = fmaf(-mean, rstd, __ldg(beta + c + k));
    }
  }
  for (long long i = i0; i < total8; i += (long long)gridDim.x * blockDim.x) {
    float v[8];
    const long long s_off = (i >> 2) * 8 + (i & 3);      // uint4 index of the hi values in a split tensor (lo: + 4)
    if (kIn == 1) {
      const uint4 h = __ldcs(reinterpret_cast<const uint4*>(xv) + s_off), l = __ldcs(reinterpret_cast<const uint4*>(xv) + s_off + 4);
      const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) { const float2 f = split_unpack2(hw[k], lw[k]); v[2 * k] = f.x; v[2 * k + 1] = f.y; }
    } else {
      const float4 a = __ldcs(reinterpret_cast<const float4*>(xv) + 2 * i), b = __ldcs(reinterpret_cast<const float4*>(xv) + 2 * i + 1);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    if (kBn) {
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = fmaxf(fmaf(v[k], sa[k], sb[k]), 0.f);
    }
    if (kOut == 1) {
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) split_pack2(v[2 * k], v[2 * k + 1], hi[k], lo[k]);
      reinterpret_cast<uint4*>(yv)[s_off] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      reinterpret_cast<uint4*>(yv)[s_off + 4] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    } else {
      reinterpret_cast<float4*>(yv)[2 * i] = make_float4(v[0], v[1], v[2], v[3]);
      reinterpret_cast<float4*>(yv)[2 * i + 1] = make_float4(v[4], v[5], v[6], v[7]);
    }
  }
}

}  // namespace lsi

using namespace lsi;

// x -> y between fp32 [n_pixels, channels] and split fp16-pair tensors (csrc/split.cuh), optionally through batch norm + ReLU
// with given stats[c] = (mean, rstd) and beta.  x_split / y_split select the layouts; in-place (y == x) is allowed when both
// sides have the same layout.  channels % 32 == 0.
extern "C" int lsi_b200_split_convert(const void* x, int x_split, const float* beta, const float* stats, void* y, int y_split,
                                      long long n_pixels, int channels, void* stream) {
  LSI_REQUIRE(x && y, "NULL pointer argument");
  LSI_REQUIRE((beta == nullptr) == (stats == nullptr), "beta and stats go together");
  LSI_REQUIRE(n_pixels >= 1 && channels >= 32 && channels % 32 == 0, "channels must be a multiple of 32");
  LSI_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0, "tensors must be 16-byte aligned");
  LSI_REQUIRE(x != y || (x_split != 0) == (y_split != 0), "in-place conversion needs the same layout on both sides");
  cudaStream_t st = as_stream(stream);
  const long long total8 = n_pixels * channels / 8;
  const unsigned c8n = (unsigned)channels / 8;
  unsigned m = c8n, g256 = 256;
  while (g256) { const unsigned t = m % g256; m = g256; g256 = t; }   // m = gcd(c8n, 256)
  m = c8n / m;
  unsigned grid = ew_grid(total8);
  grid = (grid + m - 1) / m * m;
  const int sel = (x_split ? 4 : 0) | (y_split ? 2 : 0) | (stats ? 1 : 0);
  switch (sel) {
    case 0: split_convert_kernel<0, 0, false><<<grid, 256, 0, st>>>(x, stats, beta, y, total8, c8n); break;
    case 1: split_convert_kernel<0, 0, true><<<grid, 256, 0, st>>>(x, stats, beta, y, total8, c8n); break;
    case 2: split_convert_kernel<0, 1, false><<<grid, 256, 0, st>>>(x, stats, beta, y, total8, c8n); break;
    case 3: split_convert_kernel<0, 1, true><<<grid, 256, 0, st>>>(x, stats, beta, y, total8, c8n); break;
    case 4: split_convert_kernel<1, 0, false><<<grid, 256, 0, st>>>(x, stats, beta, y, total8, c8n); break;
    case 5: split_convert_kernel<1, 0, true><<<grid, 256, 0, st>>>(x, stats, beta, y, total8, c8n); break;
    case 6: split_convert_kernel<1, 1, false><<<grid, 256, 0, st>>>(x, stats, beta, y, total8, c8n); break;
    default: split_convert_kernel<1, 1, true><<<grid, 256, 0, st>>>(x, stats, beta, y, total8, c8n); break;
  }
  LSI_LAUNCH_CHECK();
  return LSI_B200_OK;
}

namespace lsi {
}  // namespace lsi

using namespace lsi;

extern "C" int lsi_b200_bn_relu_apply_h(const void* x, int x_f16, const float* beta, const float* stats, void* y, long long n_pixels,
                                        int channels, void* stream) {
  LSI_REQUIRE(x && beta && stats && y, "NULL pointer argument");
  LSI_REQUIRE(n_pixels >= 1 && channels >= 8 && channels % 8 == 0, "channels must be a multiple of 8");
  LSI_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0 && ((uintptr_t)beta & 7) == 0 && ((uintptr_t)stats & 15) == 0,
              "tensors must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  const long long total8 = n_pixels * channels / 8;
  const unsigned c8n = (unsigned)channels / 8;
  // grid such that gridDim.x * 256 is a multiple of c8n (each thread then keeps its channels): blocks in multiples of m
  unsigned m = c8n, g256 = 256;
  while (g256) { const unsigned t = m % g256; m = g256; g256 = t; }   // m = gcd(c8n, 256)
  m = c8n / m;
  unsigned grid = ew_grid(total8);
  grid = (grid + m - 1) / m * m;
  if (x_f16) bn_apply_h_kernel<true><<<grid, 256, 0, st>>>(x, stats, beta, static_cast<__half*>(y), total8, c8n);
  else bn_apply_h_kernel<false><<<grid, 256, 0, st>>>(x, stats, beta, static_cast<__half*>(y), total8, c8n);
  LSI_LAUNCH_CHECK();
  return LSI_B200_OK;
}

extern "C" size_t lsi_b200_bn_workspace_bytes(int channels) {
  return (size_t)kStatBlocks * (size_t)(channels > 0 ? channels : 1) * 2 * sizeof(double);
}

static int stat_blocks(long long P) {
  long long b = (P + 63) / 64;
  if (b > kStatBlocks) b = kStatBlocks;
  if (b < 1) b = 1;
  return (int)b;
}

extern "C" int lsi_b200_bn_relu_forward(const float* x, const float* beta, float* y, float* stats, long long n_pixels,
                                        int channels, int x_c_stride, int y_c_stride, float eps, int relu, int stats_given,
                                        void* workspace, void* stream) {
  LSI_REQUIRE(x && beta && y && stats && workspace, "NULL pointer argument");
  LSI_REQUIRE(n_pixels >= 1 && channels >= 1 && x_c_stride >= channels && y_c_stride >= channels, "bad sizes");
  cudaStream_t st = as_stream(stream);
  if (!stats_given) {   // otherwise `stats` was produced by the conv epilogue (lsi_b200_conv2d_tc_bnstats)
    const int nb = stat_blocks(n_pixels);
    StatParams sp{x, nullptr, nullptr, nullptr, static_cast<double*>(workspace), n_pixels, channels, x_c_stride, 0, 0, 0, 0};
    channel_stats_kernel<<<nb, 256, 0, st>>>(sp);
    LSI_LAUNCH_CHECK();
    finalize_stats_kernel<<<(channels + 7) / 8, 256, 0, st>>>(static_cast<double*>(workspace), nb, channels, n_pixels, eps, 0, stats);
    LSI_LAUNCH_CHECK();
  }
  BnApplyParams ap{x, stats, beta, y, n_pixels, channels, x_c_stride, y_c_stride, relu};
  if (x_c_stride == channels && y_c_stride % 4 == 0 && channels % 4 == 0 && ((uintptr_t)x & 15) == 0 &&
      ((uintptr_t)y & 15) == 0 && ((uintptr_t)beta & 15) == 0)
    bn_apply_vec4_kernel<<<ew_grid(n_pixels * channels / 4), 256, 0, st>>>(ap);
  else
    bn_apply_kernel<<<ew_grid(n_pixels * channels), 256, 0, st>>>(ap);
  LSI_LAUNCH_CHECK();
  return LSI_B200_OK;
}

extern "C" int lsi_b200_bn_relu_backward(const float* x, const float* y, const float* dy, const float* stats, float* dx,
                                         float* dbeta_sums, long long n_pixels, int channels, int x_c_stride, int y_c_stride,
                                         int dy_c_stride, int dx_c_stride, int relu, int accumulate, void* workspace,
                                         void* stream) {
  LSI_REQUIRE(x && y && dy && stats && dx && dbeta_sums && workspace, "NULL pointer argument");
  LSI_REQUIRE(n_pixels >= 1 && channels >= 1, "bad sizes");
  cudaStream_t st = as_stream(stream);
  const int nb = stat_blocks(n_pixels);
  StatParams sp{x, y, dy, stats, static_cast<double*>(workspace), n_pixels, channels, x_c_stride, y_c_stride, dy_c_stride, 1, relu};
  channel_stats_kernel<<<nb, 256, 0, st>>>(sp);
  LSI_LAUNCH_CHECK();
  finalize_stats_kernel<<<(channels + 7) / 8, 256, 0, st>>>(static_cast<double*>(workspace), nb, channels, n_pixels, 0.f, 1, dbeta_sums);
  LSI_LAUNCH_CHECK();
  BnBwdParams bp{x, y, dy, stats, dbeta_sums, dx, n_pixels, channels, x_c_stride, y_c_stride, dy_c_stride, dx_c_stride, relu, accumulate, n_pixels};
  bn_backward_kernel<<<ew_grid(n_pixels * channels), 256, 0, st>>>(bp);
  LSI_LAUNCH_CHECK();
  return LSI_B200_OK;
}

// The same in two stages, for batch norm whose statistics span several ranks (the reference normalises over the whole batch on
// one device, nets.py:263-272): stage 1 reduces this rank's (sum dz, sum dz*xhat) into dbeta_sums; the caller all-reduces them;
// stage 2 applies dx with the given sums and 1 / n_pixels_stat (the global pixel count).
extern "C" int lsi_b200_bn_relu_backward_staged(const float* x, const float* y, const float* dy, const float* stats, float* dx,
                                                float* dbeta_sums, long long n_pixels, long long n_pixels_stat, int channels,
                                                int x_c_stride, int y_c_stride, int dy_c_stride, int dx_c_stride, int relu,
                                                int accumulate, int stage, void* workspace, void* stream) {
  LSI_REQUIRE(x && y && dy && stats && dbeta_sums && workspace, "NULL pointer argument");
  LSI_REQUIRE(n_pixels >= 1 && n_pixels_stat >= n_pixels && channels >= 1 && (stage == 1 || stage == 2), "bad sizes / stage");
  cudaStream_t st = as_stream(stream);
  if (stage == 1) {
    const int nb = stat_blocks(n_pixels);
    StatParams sp{x, y, dy, stats, static_cast<double*>(workspace), n_pixels, channels, x_c_stride, y_c_stride, dy_c_stride, 1, relu};
    channel_stats_kernel<<<nb, 256, 0, st>>>(sp);
    LSI_LAUNCH_CHECK();
    finalize_stats_kernel<<<(channels + 7) / 8, 256, 0, st>>>(static_cast<double*>(workspace), nb, channels, n_pixels, 0.f, 1, dbeta_sums);
    LSI_LAUNCH_CHECK();
    return LSI_B200_OK;
  }
  LSI_REQUIRE(dx != nullptr, "NULL pointer argument");
  BnBwdParams bp{x, y, dy, stats, dbeta_sums, dx, n_pixels, channels, x_c_stride, y_c_stride, dy_c_stride, dx_c_stride, relu, accumulate,
                 n_pixels_stat};
  bn_backward_kernel<<<ew_grid(n_pixels * channels), 256, 0, st>>>(bp);
  LSI_LAUNCH_CHECK();
  return LSI_B200_OK;
}

extern "C" int lsi_b200_channel_sums(const float* x, float* sums, long long n_pixels, int channels, int x_c_stride,
                                     void* workspace, void* stream) {
  // sums[c][0] = sum over pixels of x[., c] (bias gradients); sums[c][1] = sum of squares
  LSI_REQUIRE(x && sums && workspace, "NULL pointer argument");
  LSI_REQUIRE(n_pixels >= 1 && channels >= 1 && x_c_stride >= channels, "bad sizes");
  cudaStream_t st = as_stream(stream);
  const int nb = stat_blocks(n_pixels);
  StatParams sp{x, nullptr, nullptr, nullptr, static_cast<double*>(workspace), n_pixels, channels, x_c_stride, 0, 0, 0, 0};
  channel_stats_kernel<<<nb, 256, 0, st>>>(sp);
  LSI_LAUNCH_CHECK();
  finalize_stats_kernel<<<(channels + 7) / 8, 256, 0, st>>>(static_cast<double*>(workspace), nb, channels, n_pixels, 0.f, 1, sums);
  LSI_LAUNCH_CHECK();
  return LSI_B200_OK;
}

// Dense fast path of lsi_b200_bn_relu_backward (contiguous [P, C] tensors, C % 4 == 0, ReLU): reads z and dy only.
extern "C" int lsi_b200_bn_relu_backward_z(const float* z, const float* beta, const float* dy, const float* stats, float* dx,
                                           float* dbeta_sums, long long n_pixels, int channels, void* workspace, void* stream) {
  return lsi_b200_bn_relu_backward_zs(z, beta, dy, channels, stats, dx, dbeta_sums, n_pixels, channels, workspace, stream);
}

extern "C" int lsi_b200_bn_relu_backward_zs(const float* z, const float* beta, const float* dy, int dy_c_stride, const float* stats,
                                            float* dx, float* dbeta_sums, long long n_pixels, int channels, void* workspace,
                                            void* stream) {
  LSI_REQUIRE(z && beta && dy && stats && dx && dbeta_sums && workspace, "NULL pointer argument");
  LSI_REQUIRE(dy_c_stride >= channels && dy_c_stride % 4 == 0, "dy_c_stride must be a multiple of 4, >= channels");
  LSI_REQUIRE(n_pixels >= 1 && channels >= 4 && channels % 4 == 0 && channels <= 1024, "channels must be a multiple of 4 (<= 1024)");
  LSI_REQUIRE(((uintptr_t)z & 15) == 0 && ((uintptr_t)dy & 15) == 0 && ((uintptr_t)dx & 15) == 0, "tensors must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  BnBwdZParams p;
  p.z = reinterpret_cast<const float4*>(z); p.dy = reinterpret_cast<const float4*>(dy); p.dx = reinterpret_cast<float4*>(dx);
  p.beta = beta; p.stats = stats; p.sums = dbeta_sums; p.partial = static_cast<double*>(workspace);
  p.total4 = n_pixels * channels / 4; p.P = n_pixels; p.c4n = (unsigned)channels / 4; p.C = channels; p.dy_s4 = (unsigned)dy_c_stride / 4;
  // grid: a multiple of m blocks so that gridDim.x * 256 is a multiple of c4n (threads keep their channels across iterations)
  unsigned m = p.c4n, g256 = 256;
  while (g256) { const unsigned t = m % g256; m = g256; g256 = t; }   // m = gcd(c4n, 256)
  m = p.c4n / m;
  long long want = (p.total4 + 256 * 8 - 1) / (256 * 8);
  if (want > kStatBlocks) want = kStatBlocks;
  unsigned nb = (unsigned)((want + m - 1) / m * m);
  if (nb > (unsigned)kStatBlocks) nb = (unsigned)kStatBlocks / m * m;
  LSI_REQUIRE(nb >= 1, "channel count %d not supported by the dense batch-norm backward", channels);
  bn_bwd_z_stats_kernel<<<nb, 256, 0, st>>>(p);
  LSI_LAUNCH_CHECK();
  finalize_stats_kernel<<<(channels + 7) / 8, 256, 0, st>>>(static_cast<double*>(workspace), (int)nb, channels, n_pixels, 0.f, 1, dbeta_sums);
  LSI_LAUNCH_CHECK();
  unsigned ga = ew_grid(p.total4);
  ga = (ga + m - 1) / m * m;
  bn_bwd_z_apply_kernel<<<ga, 256, 0, st>>>(p);
  LSI_LAUNCH_CHECK();
  return LSI_B200_OK;
}

// fp64 variant: sums[c] = (sum_x, sum_x^2) as doubles -- E[x^2] - mean^2 cancels catastrophically in fp32 for channels whose mean
// dominates their spread, so the sums that cross ranks under synchronised batch norm stay in fp64
extern "C" int lsi_b200_channel_sums_f64(const float* x, double* sums, long long n_pixels, int channels, int x_c_stride,
                                         void* workspace, void* stream) {
  LSI_REQUIRE(x && sums && workspace, "NULL pointer argument");
  LSI_REQUIRE(n_pixels >= 1 && channels >= 1 && x_c_stride >= channels, "bad sizes");
  cudaStream_t st = as_stream(stream);
  const int nb = stat_blocks(n_pixels);
  StatParams sp{x, nullptr, nullptr, nullptr, static_cast<double*>(workspace), n_pixels, channels, x_c_stride, 0, 0, 0, 0};
  channel_stats_kernel<<<nb, 256, 0, st>>>(sp);
  LSI_LAUNCH_CHECK();
  finalize_stats_kernel<<<(channels + 7) / 8, 256, 0, st>>>(static_cast<double*>(workspace), nb, channels, n_pixels, 0.f, 2,
                                                           reinterpret_cast<float*>(sums));
  LSI_LAUNCH_CHECK();
  return LSI_B200_OK;
}

extern "C" int lsi_b200_copy_channels(const float* src, float* dst, long long n_pixels, int channels, int src_c_stride,
                                      int dst_c_stride, int accumulate, void* stream) {
  LSI_REQUIRE(src && dst, "NULL pointer argument");
  LSI_REQUIRE(n_pixels >= 1 && channels >= 1 && src_c_stride >= channels && dst_c_stride >= channels, "bad sizes");
  if (channels % 4 == 0 && src_c_stride % 4 == 0 && dst_c_stride % 4 == 0 && ((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 15) == 0)
    copy_channels_vec4_kernel<<<ew_grid(n_pixels * channels / 4), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const float4*>(src), reinterpret_cast<float4*>(dst), n_pixels, channels / 4, src_c_stride / 4, dst_c_stride / 4,
        accumulate);
  else
    copy_channels_kernel<<<ew_grid(n_pixels * channels), 256, 0, as_stream(stream)>>>(src, dst, n_pixels, channels, src_c_stride,
                                                                                       dst_c_stride, accumulate);
  LSI_LAUNCH_CHECK();
  return LSI_B200_OK;
}

extern "C" int lsi_b200_sigmoid_backward(const float* y, const float* dy, float* dz, long long n, void* stream) {
  LSI_REQUIRE(y && dy && dz && n >= 1, "bad arguments");
  sigmoid_backward_kernel<<<ew_grid(n), 256, 0, as_stream(stream)>>>(y, dy, dz, n);
  LSI_LAUNCH_CHECK();
  return LSI_B200_OK;
}

extern "C" int lsi_b200_adam_step(float* params, const float* grads, float* m, float* v, long long n, float learning_rate,
                                  float beta1, float beta2, float epsilon, long long step, float grad_scale, void* stream) {
  LSI_REQUIRE(params && grads && m && v && n >= 1 && step >= 1, "bad arguments");
  AdamParams a;
  a.p = params; a.g = grads; a.m = m; a.v = v; a.n = n;
  a.lr_t = (float)((double)learning_rate * sqrt(1.0 - pow((double)beta2, (double)step)) / (1.0 - pow((double)beta1, (double)step)));
  a.beta1 = beta1; a.beta2 = beta2; a.eps = epsilon; a.grad_scale = grad_scale;
  adam_kernel<<<ew_grid(n), 256, 0, as_stream(stream)>>>(a);
  LSI_LAUNCH_CHECK();
  return LSI_B200_OK;
}
