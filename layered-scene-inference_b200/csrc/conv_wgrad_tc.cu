// Tensor-core weight gradient (tcgen05 + TMA, TF32 inputs, fp32 TMEM accumulation) for the CNN's backward pass.
//
//   dW[tap][a][b] = sum over pixels p of  big(p*stride - pad + tap)[a] * small(p)[b]        (same contract as
//   lsi_b200_conv2d_wgrad: conv -> big = layer input, small = dout; transposed conv -> big = dout, small = layer input)
//
// GEMM view: D[(tap,a) rows x b columns] += A[(tap,a) x pixels] * B[b x pixels]^T with K = pixels.  Both operands are
// "MN-major" in shared memory: a TMA box of an NHWC patch lands as [128 pixels][32 channels] = 128-byte rows (K) with
// the 32 channels (M or N) contiguous, exactly the canonical MN-major atom for 32-bit operands (128-byte swizzle with 32-byte
// atoms, 4 K-rows x 128 B, 512 B apart = SBO); further 32-channel blocks of M / N sit 16 KB apart (= LBO).  The 128 rows of an M tile are four such
// blocks, enumerating (tap, 32-channel chunk) pairs, i.e. four differently SHIFTED boxes of `big` -- so D's rows are
// directly rows of dW viewed as [taps*Ca][Cb].  One UMMA (K = 8) consumes one 8-pixel group: 16 UMMAs per 128-pixel
// stage.  Pixels are split across CTAs (split-K); each CTA adds its TMEM tile into dW with fp32 reductions.
#include <cuda.h>

#include "capi_common.h"
#include "common.cuh"

namespace lsi {

// mbarrier.try_wait suspend-time hint: a waiting thread sleeps until the phase completes (or this many ns pass) instead
// of re-polling -- in the halo kernel 27 % of all issued instructions were YIELD/TRYWAIT/BRA of waiting warps
#ifndef LSI_SUSPEND_HINT_DEFINED
#define LSI_SUSPEND_HINT_DEFINED
constexpr unsigned kSuspendHintNs = 0x989680u;
#endif


namespace wg {

constexpr int kTileH = 8, kTileW = 16, kPix = 128;       // pixels per stage (K per stage)
constexpr int kBlk = 32;                                  // channels per operand block (one 128-byte row)
constexpr int kBlkBytes = kPix * 128;                     // 16 KB
constexpr int kMBlocks = 4;                               // M tile = 128 rows
constexpr int kStages = 2;
constexpr int kThreads = 192;

struct Params {
  float* dw;
  int Ca, Cb, taps, kw;
  int chunks_a;               // ceil(Ca / 32)
  int n_blocks;               // B blocks per N tile (1 or 2)
  int stride, pad_t, pad_l;
  int tiles_x, tiles_y, batch;
  int pix_tiles, tiles_per_split;
  int m_tiles, n_tiles;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity), "r"(kSuspendHintNs) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// MN-major TF32 operand, 128B swizzle with 32-byte atoms (cute Layout_MN_SW128_32B_Atom; TMA SWIZZLE_128B_ATOM_32B --
// with the plain 128B swizzle the MMA silently produces zeros): 4-row K groups 512 B apart (SBO), 32-element MN blocks
// 16 KB apart (LBO)
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)(kBlkBytes >> 4) << 16;          // leading byte offset: next 32-channel block
  d |= (uint64_t)(512 >> 4) << 32;                // stride byte offset: next group of 4 pixels (the 32B-atom swizzle repeats every 4 rows)
  d |= (uint64_t)1 << 46;                         // descriptor version (Blackwell)
  d |= (uint64_t)1 << 61;                         // SWIZZLE_128B_BASE32B: what TF32 MN-major operands require
  return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// one lane of a converged warp, in a form that lets ptxas keep the guarded region's tcgen05/TMA operands in uniform
// registers (see conv_tc.cu)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "elect.sync _|P1, 0xffffffff;\n"
      "@P1 mov.s32 %0, 1;\n"
      "}\n" : "+r"(pred));
  return pred != 0;
}

__global__ void __launch_bounds__(kThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_big, const __grid_constant__ CUtensorMap map_small, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t stage_bytes = (uint32_t)(kMBlocks + p.n_blocks) * kBlkBytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kStages * stage_bytes);
  uint64_t* empty = full + kStages;
  uint64_t* tmem_full = empty + kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tile = blockIdx.x % p.m_tiles, n_tile = blockIdx.x / p.m_tiles;
  const int split = blockIdx.y;
  const int t_begin = split * p.tiles_per_split;
  const int t_end = min(t_begin + p.tiles_per_split, p.pix_tiles);
  const int n_iters = t_end - t_begin;
  const int n_cols = p.n_blocks * kBlk;
  const int total_blocks = p.taps * p.chunks_a;
  const int a_blocks = min(kMBlocks, total_blocks - m_tile * kMBlocks);   // valid 32-row blocks of this M tile

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_big) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_small) : "memory");
    for (int i = 0; i < kStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const uint32_t tmem_cols = n_cols <= 32 ? 32 : 64;
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (n_iters > 0) {
    if (warp == 0) {
      if (elect_one()) {
        // ---------------- TMA producer ----------------
        const int per_img = p.tiles_x * p.tiles_y;
        for (int it = 0; it < n_iters; ++it) {
          const int st = it % kStages;
          const uint32_t ph = (it / kStages) & 1;
          mbar_wait(&empty[st], ph ^ 1);
          const int t = t_begin + it;
          const int n_img = t / per_img, r = t - n_img * per_img;
          const int y0 = (r / p.tiles_x) * kTileH, x0 = (r % p.tiles_x) * kTileW;   // patch of `small` pixels
          uint8_t* sa = smem + st * stage_bytes;
          uint8_t* sb = sa + kMBlocks * kBlkBytes;
          mbar_expect_tx(&full[st], (uint32_t)(a_blocks + p.n_blocks) * kBlkBytes);
          for (int j = 0; j < a_blocks; ++j) {
            const int g = m_tile * kMBlocks + j;
            const int tap = g / p.chunks_a, chunk = g - tap * p.chunks_a;
            const int ky = tap / p.kw, kx = tap - ky * p.kw;
            tma_load_4d(sa + j * kBlkBytes, &map_big, &full[st], chunk * kBlk, x0 * p.stride - p.pad_l + kx,
                        y0 * p.stride - p.pad_t + ky, n_img);
          }
          for (int j = 0; j < p.n_blocks; ++j)
            tma_load_4d(sb + j * kBlkBytes, &map_small, &full[st], (n_tile * p.n_blocks + j) * kBlk, x0, y0, n_img);
        }
      }
    } else if (warp == 1) {
      if (elect_one()) {
        // ---------------- MMA issuer ----------------
        // D=F32, A=B=TF32, both MN-major (bits 15, 16), N>>3, M>>4
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n_cols >> 3) << 17) |
                               ((uint32_t)(128 >> 4) << 24);
        for (int it = 0; it < n_iters; ++it) {
          const int st = it % kStages;
          const uint32_t ph = (it / kStages) & 1;
          mbar_wait(&full[st], ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = smem_u32(smem + st * stage_bytes), sb = sa + kMBlocks * kBlkBytes;
          const uint64_t ad = umma_desc_mn(sa), bd = umma_desc_mn(sb);   // + (bytes >> 4) advances the start address
#pragma unroll
          for (int kk = 0; kk < kPix / 8; ++kk)     // one UMMA per group of 8 pixels (1024 bytes down the tile)
            umma_tf32(tmem_base, ad + 64 * kk, bd + 64 * kk, idesc, (it | kk) != 0);
          umma_commit(&empty[st]);
        }
        umma_commit(tmem_full);
      }
    } else {
      // ---------------- epilogue: TMEM -> fp32 reductions into dW ----------------
      mbar_wait(tmem_full, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int lg = warp & 3;
      const int row = lg * 32 + lane;                         // row of the M tile = (block, channel within block)
      const int blk = row >> 5, ch = row & 31;
      const int g = m_tile * kMBlocks + blk;
      const int tap = g / p.chunks_a, chunk = g - tap * p.chunks_a;
      const int a = chunk * kBlk + ch;
      const bool row_ok = blk < a_blocks && a < p.Ca;
      float* dst = p.dw + ((size_t)tap * p.Ca + a) * p.Cb + (size_t)n_tile * n_cols;
      for (int c = 0; c < n_cols; c += 32) {
        uint32_t r[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)c;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (row_ok) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int b = n_tile * n_cols + c + j;
            if (b < p.Cb) atomicAdd(dst + c + j, __uint_as_float(r[j]));
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

}  // namespace wg
}  // namespace lsi

using namespace lsi;

extern "C" int lsi_b200_conv2d_wgrad_tc_supported(const lsi_b200_conv_desc* d) {
  if (!d) return 0;
  if (d->mode != 0 || d->stride < 1 || d->stride > 2) return 0;
  if (d->in_c_stride % 4 != 0 || d->out_c_stride % 4 != 0 || d->c_in % 4 != 0 || d->c_out % 4 != 0) return 0;
  if (d->c_in < 16) return 0;                                   // the 3-channel stem stays on the fp32 kernel
  return 1;
}

// Same contract as lsi_b200_conv2d_wgrad (dw is zero-filled first unless d->accumulate).
extern "C" int lsi_b200_conv2d_wgrad_tc(const lsi_b200_conv_desc* d, const float* big, const float* small, float* dw,
                                        void* stream) {
  LSI_REQUIRE(d && big && small && dw, "NULL pointer argument");
  LSI_REQUIRE(lsi_b200_conv2d_wgrad_tc_supported(d), "shape not supported by the tensor-core weight-gradient path");
  LSI_REQUIRE(((uintptr_t)big & 15) == 0 && ((uintptr_t)small & 15) == 0, "inputs must be 16-byte aligned");
  wg::EncodeTiledFn encode = wg::get_encode();
  LSI_REQUIRE(encode != nullptr, "cuTensorMapEncodeTiled is unavailable");
  cudaStream_t st = as_stream(stream);
  wg::Params p;
  p.dw = dw; p.Ca = d->c_in; p.Cb = d->c_out; p.taps = d->kh * d->kw; p.kw = d->kw;
  p.chunks_a = (p.Ca + wg::kBlk - 1) / wg::kBlk;
  p.n_blocks = p.Cb > 32 ? 2 : 1;
  p.stride = d->stride; p.pad_t = d->pad_top; p.pad_l = d->pad_left;
  p.tiles_x = (d->w_out + wg::kTileW - 1) / wg::kTileW; p.tiles_y = (d->h_out + wg::kTileH - 1) / wg::kTileH; p.batch = d->batch;
  p.pix_tiles = p.tiles_x * p.tiles_y * d->batch;
  p.m_tiles = (p.taps * p.chunks_a + wg::kMBlocks - 1) / wg::kMBlocks;
  p.n_tiles = (p.Cb + p.n_blocks * wg::kBlk - 1) / (p.n_blocks * wg::kBlk);
  const int mn = p.m_tiles * p.n_tiles;
  int splits = (148 * 2 + mn - 1) / mn;
  if (splits > p.pix_tiles) splits = p.pix_tiles;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  p.tiles_per_split = (p.pix_tiles + splits - 1) / splits;
  splits = (p.pix_tiles + p.tiles_per_split - 1) / p.tiles_per_split;

  auto make_map = [&](CUtensorMap* m, const float* base, int channels, int cs, int h, int w, int es) -> int {
    cuuint64_t dims[4] = {(cuuint64_t)channels, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)d->batch};
    cuuint64_t strides[3] = {(cuuint64_t)cs * 4, (cuuint64_t)w * cs * 4, (cuuint64_t)h * w * cs * 4};
    cuuint32_t box[4] = {(cuuint32_t)wg::kBlk, (cuuint32_t)((wg::kTileW - 1) * es + 1), (cuuint32_t)((wg::kTileH - 1) * es + 1), 1};
    cuuint32_t estr[4] = {1, (cuuint32_t)es, (cuuint32_t)es, 1};
    CUresult r = encode(m, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 4, const_cast<float*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed: %d", (int)r); return LSI_B200_ECUDA; }
    return LSI_B200_OK;
  };
  CUtensorMap map_big, map_small;
  if (int rc = make_map(&map_big, big, d->c_in, d->in_c_stride, d->h_in, d->w_in, d->stride)) return rc;
  if (int rc = make_map(&map_small, small, d->c_out, d->out_c_stride, d->h_out, d->w_out, 1)) return rc;
  if (!d->accumulate) LSI_CUDA(cudaMemsetAsync(dw, 0, (size_t)p.taps * p.Ca * p.Cb * sizeof(float), st));
  const size_t smem = (size_t)wg::kStages * (wg::kMBlocks + p.n_blocks) * wg::kBlkBytes + 256 + 1024;
  static size_t smem_set = 0;
  if (smem > smem_set) {
    LSI_CUDA(cudaFuncSetAttribute(wg::wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  dim3 grid((unsigned)mn, (unsigned)splits);
  {
    ScopedTiming tm(kWgrad, st);
    wg::wgrad_tc_kernel<<<grid, wg::kThreads, smem, st>>>(map_big, map_small, p);
  }
  LSI_LAUNCH_CHECK();
  return LSI_B200_OK;
}
