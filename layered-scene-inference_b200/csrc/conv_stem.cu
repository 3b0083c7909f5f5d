// Tensor-core stem convolution: cnv1 of the U-Net (nets.py:273: 7x7, stride 2, 3 -> 32 channels, SAME padding) as an
// im2col GEMM on tcgen05.  TMA cannot feed this layer (a 3-channel fp32 pixel is 12 bytes; tensor maps need 16-byte pixel
// strides), and on CUDA cores it was the sixth most expensive launch of the inference step (1.8 ms at 64 x 256 x 896, 19
// TFLOP/s).  Here the threads build the im2col operand themselves:
//
//   tile        2 output rows x 64 output columns = 128 pixels = 128 TMEM lanes;  D[128 x 32] += A[128 x 160] * B[32 x 160]^T
//   K           (ky, kx, c) = 147 values, zero-padded to 160 = 5 blocks of 32 fp16 (64-byte rows, 64B swizzle, K-major)
//   staging     the 9 input rows x 133 pixels the tile needs -> shared memory with coalesced loads (zero outside the image)
//   A operand   thread r converts its pixel's 7 x 21 contiguous floats to fp16 and stores them as 16-byte chunks at the
//               swizzled positions (chunk ^ ((r >> 1) & 3)) -- conflict-free -- then fence.proxy.async
//   B operand   the whole filter bank (10 KB as fp16), swizzled into shared memory once per CTA
//   MMA         10 x tcgen05.mma.kind::f16 (M 128, N 32, K 16) per tile from one elected thread, completion on an mbarrier
//   epilogue    tcgen05.ld -> registers -> per-thread batch-statistics partial sums + 64-byte (fp16) / 128-byte stores
// Persistent CTAs, three per SM, so one CTA's staging/building overlaps another's MMAs and stores.
#include <cuda_fp16.h>

#include "capi_common.h"
#include "common.cuh"
#include "split.cuh"

namespace lsi {

// mbarrier.try_wait suspend-time hint: a waiting thread sleeps until the phase completes (or this many ns pass) instead
// of re-polling -- in the halo kernel 27 % of all issued instructions were YIELD/TRYWAIT/BRA of waiting warps
#ifndef LSI_SUSPEND_HINT_DEFINED
#define LSI_SUSPEND_HINT_DEFINED
constexpr unsigned kSuspendHintNs = 0x989680u;
#endif

namespace {

constexpr int kK = 7, kS = 2, kCin = 3, kN = 32;
constexpr int kTileRows = 2, kTileCols = 64;
constexpr int kInRows = (kTileRows - 1) * kS + kK;                 // 9
constexpr int kInCols = (kTileCols - 1) * kS + kK;                 // 133 pixels
constexpr int kRowF = 400;                                          // floats per staged row (133 * 3 = 399, padded)
constexpr int kKPad = 160, kKBlocks = kKPad / 32;
constexpr uint32_t kABlock = 128 * 64, kBBlock = kN * 64;          // bytes of one 32-wide K block of A / B
constexpr uint32_t kBBlockS = 2 * kN * 64;                          // split mode: (Whi ; Wlo) rows per K block
constexpr int kThreads = 128;

struct StemParams {
  const float* in; void* out; const __half* wk; float* stat_part;
  int H, W, Ho, Wo, out_cs, pad_t, pad_l, batch;
  int tiles_x, tiles_y, total_tiles, out_f16;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(smem_u32(bar)), "r"(parity), "r"(kSuspendHintNs) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n.reg .pred P1;\nelect.sync _|P1, 0xffffffff;\n@P1 mov.s32 %0, 1;\n}" : "+r"(pred));
  return pred != 0;
}
// K-major operand with the 64-byte swizzle, dense 8-row groups (512 bytes apart)
__device__ __forceinline__ uint64_t desc_sw64(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ uint32_t pack_half2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// kSplit: the split-precision mode (split.cuh).  The im2col operand is built twice per tile into the same buffer -- first the hi
// halves (MMAs against (Whi ; Wlo), N' = 64, into TMEM columns [0, 64)), then, once those MMAs have completed, the lo halves
// (v - hi) * 2^11 (MMAs against Whi, N' = 32, into columns [32, 64)) -- so the shared-memory footprint, hence three CTAs per SM,
// stays that of the fp16 kernel.  out_f16 == 2 stores split pairs.
template <bool kSplit>
__global__ void __launch_bounds__(kThreads, 3) conv_stem_kernel(const StemParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 64B-swizzle atoms are 512 bytes: 512-byte alignment is enough (and keeps the split variant at three CTAs per SM)
  uint8_t* smem = smem_raw + ((512u - (smem_u32(smem_raw) & 511u)) & 511u);
  constexpr uint32_t kBB = kSplit ? kBBlockS : kBBlock;
  constexpr int kBRows = kSplit ? 2 * kN : kN;
  uint8_t* s_a = smem;                                          // 5 x 8 KB
  uint8_t* s_b = s_a + kKBlocks * kABlock;                      // 5 x 2 KB (split: 5 x 4 KB)
  float* s_in = reinterpret_cast<float*>(s_b + kKBlocks * kBB);   // 9 x 400 floats
  uint64_t* bar = reinterpret_cast<uint64_t*>(s_in + kInRows * kRowF);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kSplit ? 64u : 32u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // filter bank -> swizzled shared memory: wk is [kb][n][32] fp16; 16-byte chunk c of row n goes to chunk c ^ ((n >> 1) & 3)
  for (int i = tid; i < kKBlocks * kBRows * 4; i += kThreads) {
    const int c = i & 3, n = (i >> 2) % kBRows, kb = i / (4 * kBRows);
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(p.wk) + i);
    *reinterpret_cast<uint4*>(s_b + kb * kBB + n * 64 + ((c ^ ((n >> 1) & 3)) << 4)) = v;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  // instruction descriptor: D = F32, A = B = F16, K-major both, N >> 3, M >> 4
  const uint32_t idesc = (1u << 4) | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t idesc2 = (1u << 4) | ((uint32_t)((2 * kN) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // split, first pass: N' = 64

  const int ry = tid >> 6, rx = tid & 63;                       // this thread's pixel of the tile = A row = TMEM lane
  float ssum[kN], ssq[kN];
#pragma unroll
  for (int j = 0; j < kN; ++j) { ssum[j] = 0.f; ssq[j] = 0.f; }
  uint32_t parity = 0;
  const int per_img = p.tiles_x * p.tiles_y;
  for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
    const int n_img = tile / per_img;
    const int r = tile - n_img * per_img;
    const int oy0 = (r / p.tiles_x) * kTileRows, ox0 = (r % p.tiles_x) * kTileCols;
    // ---- stage the input patch (coalesced; zero outside the image = SAME padding) ----
    const int iy0 = oy0 * kS - p.pad_t, ix0 = ox0 * kS - p.pad_l;
    const float* img = p.in + (size_t)n_img * p.H * p.W * kCin;
    for (int i = tid; i < kInRows * kRowF; i += kThreads) {
      const int row = i / kRowF, col = i - row * kRowF;
      const int y = iy0 + row, xf = ix0 * kCin + col;           // float offset along the image row
      float v = 0.f;
      if (col < kInCols * kCin && (unsigned)y < (unsigned)p.H && xf >= 0 && xf < p.W * kCin) v = __ldg(img + (size_t)y * p.W * kCin + xf);
      s_in[i] = v;
    }
    __syncthreads();   // (also: every thread has drained the previous tile's accumulator, see the fence before it)
    // ---- im2col row of this thread's pixel -> fp16 (split: hi, then lo), swizzled 16-byte chunks; MMAs; wait ----
#pragma unroll 1
    for (int pass = 0; pass < (kSplit ? 2 : 1); ++pass) {
      {
        const float* src = s_in + (ry * kS) * kRowF + rx * kS * kCin;
        const uint32_t sw = (uint32_t)((tid >> 1) & 3);
        uint8_t* arow = s_a + tid * 64;
#pragma unroll
        for (int kb = 0; kb < kKBlocks; ++kb) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t h[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int k0 = kb * 32 + c * 8 + e * 2, k1 = k0 + 1;
              const float f0 = (k0 < kK * kK * kCin) ? src[(k0 / 21) * kRowF + (k0 % 21)] : 0.f;
              const float f1 = (k1 < kK * kK * kCin) ? src[(k1 / 21) * kRowF + (k1 % 21)] : 0.f;
              if (kSplit) {
                uint32_t hi, lo;
                split_pack2(f0, f1, hi, lo);
                h[e] = pass == 0 ? hi : lo;
              } else {
                h[e] = pack_half2(f0, f1);
              }
            }
            *reinterpret_cast<uint4*>(arow + kb * kABlock + ((c ^ sw) << 4)) = make_uint4(h[0], h[1], h[2], h[3]);
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();
      if (warp == 0 && elect_one()) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t a0 = desc_sw64(smem_u32(s_a)), b0 = desc_sw64(smem_u32(s_b));
#pragma unroll
        for (int kb = 0; kb < kKBlocks; ++kb)
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) {
            const uint64_t ad = a0 + (uint64_t)((kb * kABlock) >> 4) + 2 * kk, bd = b0 + (uint64_t)((kb * kBB) >> 4) + 2 * kk;
            if (!kSplit) umma_f16(tmem_base, ad, bd, idesc, (kb | kk) != 0);
            else if (pass == 0) umma_f16(tmem_base, ad, bd, idesc2, (kb | kk) != 0);     // a_hi x (Whi ; Wlo) -> columns [0, 64)
            else umma_f16(tmem_base + kN, ad, bd, idesc, 1u);                            // a_lo x Whi -> columns [32, 64)
          }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
      }
      __syncwarp();
      mbar_wait(bar, parity);        // (split: the hi operand may be overwritten only after its MMAs have read it)
      parity ^= 1;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    // ---- epilogue ----
    uint32_t acc[32];
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(acc[0]), "=r"(acc[1]), "=r"(acc[2]), "=r"(acc[3]), "=r"(acc[4]), "=r"(acc[5]), "=r"(acc[6]), "=r"(acc[7]), "=r"(acc[8]),
          "=r"(acc[9]), "=r"(acc[10]), "=r"(acc[11]), "=r"(acc[12]), "=r"(acc[13]), "=r"(acc[14]), "=r"(acc[15]), "=r"(acc[16]),
          "=r"(acc[17]), "=r"(acc[18]), "=r"(acc[19]), "=r"(acc[20]), "=r"(acc[21]), "=r"(acc[22]), "=r"(acc[23]), "=r"(acc[24]),
          "=r"(acc[25]), "=r"(acc[26]), "=r"(acc[27]), "=r"(acc[28]), "=r"(acc[29]), "=r"(acc[30]), "=r"(acc[31])
        : "r"(taddr));
    if (kSplit) {
      uint32_t acc1[32];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(acc1[0]), "=r"(acc1[1]), "=r"(acc1[2]), "=r"(acc1[3]), "=r"(acc1[4]), "=r"(acc1[5]), "=r"(acc1[6]), "=r"(acc1[7]), "=r"(acc1[8]),
            "=r"(acc1[9]), "=r"(acc1[10]), "=r"(acc1[11]), "=r"(acc1[12]), "=r"(acc1[13]), "=r"(acc1[14]), "=r"(acc1[15]), "=r"(acc1[16]),
            "=r"(acc1[17]), "=r"(acc1[18]), "=r"(acc1[19]), "=r"(acc1[20]), "=r"(acc1[21]), "=r"(acc1[22]), "=r"(acc1[23]), "=r"(acc1[24]),
            "=r"(acc1[25]), "=r"(acc1[26]), "=r"(acc1[27]), "=r"(acc1[28]), "=r"(acc1[29]), "=r"(acc1[30]), "=r"(acc1[31])
          : "r"(taddr + (uint32_t)kN));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[j] = __float_as_uint(fmaf(__uint_as_float(acc1[j]), kSplitInvScale, __uint_as_float(acc[j])));
    } else {
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");   // ordered before the next tile's first __syncthreads
    const int oy = oy0 + ry, ox = ox0 + rx;
    if (oy < p.Ho && ox < p.Wo) {
#pragma unroll
      for (int j = 0; j < kN; ++j) {
        const float v = __uint_as_float(acc[j]);
        ssum[j] += v; ssq[j] = fmaf(v, v, ssq[j]);
      }
      const size_t e = ((size_t)(n_img * p.Ho + oy) * p.Wo + ox) * p.out_cs;
      if (kSplit && p.out_f16 == 2) {      // split pairs: the pixel's one 32-channel chunk = [hi 64 B | lo 64 B] at the fp32 chunk address
        uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<float*>(p.out) + e);
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) split_pack2(__uint_as_float(acc[2 * j]), __uint_as_float(acc[2 * j + 1]), hi[j], lo[j]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          dst[j] = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
          dst[4 + j] = make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
        }
      } else if (p.out_f16) {
        uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(p.out) + e);
#pragma unroll
        for (int j = 0; j < kN; j += 8)
          dst[j >> 3] = make_uint4(pack_half2(__uint_as_float(acc[j]), __uint_as_float(acc[j + 1])),
                                   pack_half2(__uint_as_float(acc[j + 2]), __uint_as_float(acc[j + 3])),
                                   pack_half2(__uint_as_float(acc[j + 4]), __uint_as_float(acc[j + 5])),
                                   pack_half2(__uint_as_float(acc[j + 6]), __uint_as_float(acc[j + 7])));
      } else {
        float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + e);
#pragma unroll
        for (int j = 0; j < kN; j += 4)
          dst[j >> 2] = make_float4(__uint_as_float(acc[j]), __uint_as_float(acc[j + 1]), __uint_as_float(acc[j + 2]), __uint_as_float(acc[j + 3]));
      }
    }
  }
  // batch statistics: per-warp partial sums (finalised by stem_finalize_stats_kernel)
  if (p.stat_part) {
#pragma unroll
    for (int j = 0; j < kN; ++j) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        ssum[j] += __shfl_xor_sync(0xffffffffu, ssum[j], o);
        ssq[j] += __shfl_xor_sync(0xffffffffu, ssq[j], o);
      }
    }
    if (lane == 0) {
      float* part = p.stat_part + ((size_t)blockIdx.x * 4 + warp) * kN * 2;
#pragma unroll
      for (int j = 0; j < kN; ++j) { part[2 * j] = ssum[j]; part[2 * j + 1] = ssq[j]; }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kSplit ? 64u : 32u) : "memory");
  }
}

// weights (any strides) -> [kb][n][32] fp16 with k = tap * 3 + c, zero for k >= 147
__global__ void __launch_bounds__(256) stem_prep_weights_kernel(const float* __restrict__ w, __half* __restrict__ wk, int w_tap, int w_ci,
                                                                int w_co) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kKBlocks * kN * 32) return;
  const int kk = i & 31, n = (i >> 5) % kN, kb = i / (32 * kN);
  const int k = kb * 32 + kk;
  float v = 0.f;
  if (k < kK * kK * kCin) v = w[(size_t)(k / kCin) * w_tap + (size_t)(k % kCin) * w_ci + (size_t)n * w_co];
  wk[i] = __float2half_rn(v);
}

// split mode: [kb][64][32] fp16, rows 0..31 = w_hi, rows 32..63 = w_lo = rn16((w - w_hi) * 2^11)
__global__ void __launch_bounds__(256) stem_prep_weights_split_kernel(const float* __restrict__ w, __half* __restrict__ wk, int w_tap, int w_ci,
                                                                      int w_co) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kKBlocks * 2 * kN * 32) return;
  const int kk = i & 31, r = (i >> 5) % (2 * kN), kb = i / (32 * 2 * kN);
  const int n = r % kN, k = kb * 32 + kk;
  float v = 0.f;
  if (k < kK * kK * kCin) v = w[(size_t)(k / kCin) * w_tap + (size_t)(k % kCin) * w_ci + (size_t)n * w_co];
  const __half hi = __float2half_rn(fminf(fmaxf(v, -kSplitMax), kSplitMax));
  wk[i] = r < kN ? hi : __float2half_rn((v - __half2float(hi)) * kSplitScale);
}

__global__ void __launch_bounds__(256) stem_finalize_stats_kernel(const float* __restrict__ partial, int nparts, long long P, float eps,
                                                                  float* __restrict__ out) {
  const int c = threadIdx.x >> 3, sub = threadIdx.x & 7;   // 32 channels x 8 lanes
  double a = 0.0, b = 0.0;
  for (int i = sub; i < nparts; i += 8) { a += (double)partial[((size_t)i * kN + c) * 2]; b += (double)partial[((size_t)i * kN + c) * 2 + 1]; }
  for (int o = 4; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
  if (sub != 0) return;
  const double mean = a / (double)P;
  double var = b / (double)P - mean * mean;
  if (var < 0.0) var = 0.0;
  out[2 * c] = (float)mean; out[2 * c + 1] = (float)(1.0 / sqrt(var + (double)eps));
}

int stem_num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

constexpr size_t kStemSmem = 1024 + kKBlocks * (kABlock + kBBlock) + kInRows * kRowF * sizeof(float) + 64;
constexpr size_t kStemSmemS = 512 + kKBlocks * (kABlock + kBBlockS) + kInRows * kRowF * sizeof(float) + 64;
constexpr size_t kWkBytes = (size_t)kKBlocks * 2 * kN * 32 * sizeof(__half);      // sized for the split bank (Whi ; Wlo)

}  // namespace
}  // namespace lsi

using namespace lsi;

extern "C" int lsi_b200_conv2d_stem_tc_supported(const lsi_b200_conv_desc* d) {
  if (!d) return 0;
  return (d->c_in == kCin && d->in_c_stride == kCin && d->c_out == kN && d->kh == kK && d->kw == kK && d->stride == kS && d->mode == 0 &&
          d->epilogue == 0 && d->accumulate == 0 && d->out_c_stride % 8 == 0 && d->h_out == (d->h_in + 1) / 2 && d->w_out == (d->w_in + 1) / 2 &&
          d->pad_top >= 0 && d->pad_top < kK && d->pad_left >= 0 && d->pad_left < kK && (long long)d->w_in * kCin < (1ll << 30))
             ? 1 : 0;
}

extern "C" size_t lsi_b200_conv2d_stem_tc_workspace_bytes(void) {
  return 256 + kWkBytes + 256 + (size_t)148 * 4 * 4 * kN * 2 * sizeof(float);
}

static int stem_tc_impl(const lsi_b200_conv_desc* d, const float* in, const float* w, void* out, int out_f16, float* bn_stats, float bn_eps,
                        void* workspace, size_t workspace_bytes, void* stream, bool split);

extern "C" int lsi_b200_conv2d_stem_tc(const lsi_b200_conv_desc* d, const float* in, const float* w, void* out, int out_f16,
                                       float* bn_stats, float bn_eps, void* workspace, size_t workspace_bytes, void* stream) {
  LSI_REQUIRE(out_f16 == 0 || out_f16 == 1, "out_f16 must be 0 (fp32) or 1 (fp16)");
  return stem_tc_impl(d, in, w, out, out_f16, bn_stats, bn_eps, workspace, workspace_bytes, stream, false);
}

// Split-precision stem (csrc/split.cuh): fp32 image in, split fp16-pair im2col operand and weights, three exact fp16 products per
// fp32 product; out_kind 0: fp32 output, 2: split output.
extern "C" int lsi_b200_conv2d_stem_tc_s(const lsi_b200_conv_desc* d, const float* in, const float* w, void* out, int out_kind,
                                         float* bn_stats, float bn_eps, void* workspace, size_t workspace_bytes, void* stream) {
  LSI_REQUIRE(out_kind == 0 || out_kind == 2, "out_kind must be 0 (fp32) or 2 (split)");
  LSI_REQUIRE(out_kind != 2 || (d && d->out_c_stride % 32 == 0), "split output needs a 32-channel-aligned pixel stride");
  return stem_tc_impl(d, in, w, out, out_kind, bn_stats, bn_eps, workspace, workspace_bytes, stream, true);
}

static int stem_tc_impl(const lsi_b200_conv_desc* d, const float* in, const float* w, void* out, int out_f16, float* bn_stats, float bn_eps,
                        void* workspace, size_t workspace_bytes, void* stream, bool split) {
  LSI_REQUIRE(d && in && w && out && workspace, "NULL pointer argument");
  LSI_REQUIRE(lsi_b200_conv2d_stem_tc_supported(d), "shape not supported by the tensor-core stem (7x7 stride-2 conv, 3 -> 32 channels)");
  LSI_REQUIRE(workspace_bytes >= lsi_b200_conv2d_stem_tc_workspace_bytes(), "workspace too small");
  LSI_REQUIRE(((uintptr_t)out & 15) == 0, "output must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  __half* wk = reinterpret_cast<__half*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
  float* part = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(wk) + kWkBytes + 255) & ~uintptr_t(255));
  if (split) stem_prep_weights_split_kernel<<<(kKBlocks * 2 * kN * 32 + 255) / 256, 256, 0, st>>>(w, wk, d->w_tap_stride, d->w_ci_stride, d->w_co_stride);
  else stem_prep_weights_kernel<<<(kKBlocks * kN * 32 + 255) / 256, 256, 0, st>>>(w, wk, d->w_tap_stride, d->w_ci_stride, d->w_co_stride);
  LSI_LAUNCH_CHECK();
  StemParams p;
  p.in = in; p.out = out; p.wk = wk; p.stat_part = bn_stats ? part : nullptr;
  p.H = d->h_in; p.W = d->w_in; p.Ho = d->h_out; p.Wo = d->w_out; p.out_cs = d->out_c_stride; p.pad_t = d->pad_top; p.pad_l = d->pad_left;
  p.batch = d->batch; p.out_f16 = out_f16;
  p.tiles_x = (p.Wo + kTileCols - 1) / kTileCols; p.tiles_y = (p.Ho + kTileRows - 1) / kTileRows;
  p.total_tiles = p.tiles_x * p.tiles_y * d->batch;
  static bool attr_set = false;
  if (!attr_set) {
    LSI_CUDA(cudaFuncSetAttribute(conv_stem_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kStemSmem));
    LSI_CUDA(cudaFuncSetAttribute(conv_stem_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kStemSmemS));
    attr_set = true;
  }
  int n_ctas = stem_num_sms() * 3;
  if (n_ctas > 148 * 4) n_ctas = 148 * 4;
  if (n_ctas > p.total_tiles) n_ctas = p.total_tiles;
  {
    ScopedTiming tm(kConvTc, st);
    if (split) conv_stem_kernel<true><<<n_ctas, kThreads, kStemSmemS, st>>>(p);
    else conv_stem_kernel<false><<<n_ctas, kThreads, kStemSmem, st>>>(p);
  }
  LSI_LAUNCH_CHECK();
  if (bn_stats) {
    stem_finalize_stats_kernel<<<1, 256, 0, st>>>(part, n_ctas * 4, (long long)d->batch * d->h_out * d->w_out, bn_eps, bn_stats);
    LSI_LAUNCH_CHECK();
  }
  return LSI_B200_OK;
}
