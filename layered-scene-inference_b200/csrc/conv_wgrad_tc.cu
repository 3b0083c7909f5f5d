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

#include <cstdlib>

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
  int vec;                    // dW rows are 16-byte aligned: vector reductions
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


// 32 consecutive columns of one dW row: 16-byte vector reductions where the row segment allows (4x fewer L2 operations than
// scalar atomics; the epilogues of all CTAs land on the L2 reduction units at once, and with split-K every dW element is hit
// once per split)
__device__ __forceinline__ void red_row32(float* dst, const uint32_t (&r)[32], int b0, int Cb, bool vec) {
  if (vec && b0 + 32 <= Cb) {
#pragma unroll
    for (int j = 0; j < 32; j += 4)
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(__uint_as_float(r[j])),
                   "f"(__uint_as_float(r[j + 1])), "f"(__uint_as_float(r[j + 2])), "f"(__uint_as_float(r[j + 3])) : "memory");
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (b0 + j < Cb) atomicAdd(dst + j, __uint_as_float(r[j]));
  }
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
        if (row_ok) red_row32(dst + c, r, n_tile * n_cols + c, p.Cb, p.vec != 0);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}


// ---------------------------------------------------------------------------------------------------------------------------
// Halo-tile variant (stride 1): the kernel above loads one shifted box of `big` per (tap, chunk) and one box of `small` per M
// tile -- at full resolution that is 12 x 16 KB from L2 per 128 pixels for 32 KB of distinct data, and the layer runs at the
// L2 -> SM limit (~12 TB/s), not at the tensor pipe.  Here a CTA owns (32-channel chunk of `big`, N tile of `small`, a group
// of tap rows) and loads, per 16 x 8 pixel tile, ONE halo box of `big` ((16 + kw - 1) x (8 + kh - 1) pixels x 32 channels)
// and ONE set of `small` boxes.  Every tap is then a descriptor that STARTS (ky * halo_w + kx) pixel rows into the halo box
// (the swizzle is a function of absolute shared-memory address bits, as for the K-major halo conv), and the four 32-row
// blocks of an M tile are four consecutive kx taps: leading-byte-offset = ONE pixel row (128 B).  An M tile is (ky, 4 kx);
// for 3 x 3 filters its fourth block reads the next pixel (junk rows 96..127 of D, never stored).  Accumulators: one
// [128 x N] TMEM tile per M tile of the CTA's group (G * N <= 512 columns).
struct HParams {
  float* dw;
  int Ca, Cb, kh, kw;
  int chunks_a, n_blocks;
  int pad_t, pad_l;
  int tiles_x, tiles_y;
  int pix_tiles, tiles_per_split;
  int mt_per_row, m_tiles, G, m_groups, n_tiles;
  int halo_w, halo_bytes, a_bytes;     // a_bytes: halo_bytes rounded up to 1024 (+ slack for the junk block's reads)
  int stages;
  int vec;
};

__device__ __forceinline__ uint64_t umma_desc_mn_lbo(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;
  return d;
}

__global__ void __launch_bounds__(kThreads, 1)
wgrad_halo_kernel(const __grid_constant__ CUtensorMap map_big, const __grid_constant__ CUtensorMap map_small, const HParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t stage_bytes = (uint32_t)p.a_bytes + (uint32_t)p.n_blocks * kBlkBytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + p.stages * stage_bytes);
  uint64_t* empty = full + p.stages;
  uint64_t* tmem_full = empty + p.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int bx = blockIdx.x;
  const int mg = bx % p.m_groups; bx /= p.m_groups;
  const int chunk = bx % p.chunks_a;
  const int n_tile = bx / p.chunks_a;
  const int split = blockIdx.y;
  const int t_begin = split * p.tiles_per_split;
  const int t_end = min(t_begin + p.tiles_per_split, p.pix_tiles);
  const int n_iters = t_end - t_begin;
  const int n_cols = p.n_blocks * kBlk;
  const int g_cnt = min(p.G, p.m_tiles - mg * p.G);          // M tiles (accumulators) of this CTA

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_big) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_small) : "memory");
    for (int i = 0; i < p.stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < p.G * n_cols) tmem_cols <<= 1;
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
        // ---------------- TMA producer: one halo box of `big`, n_blocks boxes of `small` per pixel tile ----------------
        const int per_img = p.tiles_x * p.tiles_y;
        int t = t_begin;
        int n_img = t / per_img, r = t - n_img * per_img;
        int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
        int st = 0; uint32_t ph = 0;
        for (int it = 0; it < n_iters; ++it) {
          mbar_wait(&empty[st], ph ^ 1);
          uint8_t* sa = smem + st * stage_bytes;
          uint8_t* sb = sa + p.a_bytes;
          mbar_expect_tx(&full[st], (uint32_t)p.halo_bytes + (uint32_t)p.n_blocks * kBlkBytes);
          const int x0 = tx * kTileW, y0 = ty * kTileH;
          tma_load_4d(sa, &map_big, &full[st], chunk * kBlk, x0 - p.pad_l, y0 - p.pad_t, n_img);
          for (int j = 0; j < p.n_blocks; ++j)
            tma_load_4d(sb + j * kBlkBytes, &map_small, &full[st], (n_tile * p.n_blocks + j) * kBlk, x0, y0, n_img);
          if (++tx == p.tiles_x) { tx = 0; if (++ty == p.tiles_y) { ty = 0; ++n_img; } }
          if (++st == p.stages) { st = 0; ph ^= 1; }
        }
      }
    } else if (warp == 1) {
      if (elect_one()) {
        // ---------------- MMA issuer ----------------
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n_cols >> 3) << 17) |
                               ((uint32_t)(128 >> 4) << 24);
        // first pixel row (in 16-byte units of the descriptor's start address: 128 B per row = 8) of this CTA's first M tile
        const int mt0 = mg * p.G;
        const int ky0 = mt0 / p.mt_per_row, kg0 = mt0 - ky0 * p.mt_per_row;
        int st = 0; uint32_t ph = 0;
        for (int it = 0; it < n_iters; ++it) {
          mbar_wait(&full[st], ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = smem_u32(smem + st * stage_bytes), sb = sa + p.a_bytes;
          const uint64_t ad = umma_desc_mn_lbo(sa, 128), bd = umma_desc_mn(sb);
          int ky = ky0, kg = kg0;
          for (int g = 0; g < g_cnt; ++g) {
            const uint32_t row0 = (uint32_t)(ky * p.halo_w + kg * 4) * 8;       // (ky, kx0) shift of the halo box
            const uint32_t td = tmem_base + (uint32_t)(g * n_cols);
#pragma unroll
            for (int kk = 0; kk < kPix / 8; ++kk) {     // tile row kk >> 1, pixels 8 * (kk & 1) .. + 7
              const uint32_t row = row0 + (uint32_t)((kk >> 1) * p.halo_w + (kk & 1) * 8) * 8;
              umma_tf32(td, ad + row, bd + 64 * kk, idesc, (it | kk) != 0);
            }
            if (++kg == p.mt_per_row) { kg = 0; ++ky; }
          }
          umma_commit(&empty[st]);
          if (++st == p.stages) { st = 0; ph ^= 1; }
        }
        umma_commit(tmem_full);
      }
    } else {
      // ---------------- epilogue: TMEM -> fp32 reductions into dW ----------------
      mbar_wait(tmem_full, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int lg = warp & 3;
      const int row = lg * 32 + lane;                         // row of the M tile = (kx within the group of 4, channel)
      const int blk = row >> 5, ch = row & 31;
      const int a = chunk * kBlk + ch;
      for (int g = 0; g < g_cnt; ++g) {
        const int mt = mg * p.G + g;
        const int ky = mt / p.mt_per_row, kx = (mt - ky * p.mt_per_row) * 4 + blk;
        const bool row_ok = kx < p.kw && a < p.Ca;
        float* dst = p.dw + ((size_t)(ky * p.kw + kx) * p.Ca + a) * p.Cb + (size_t)n_tile * n_cols;
        for (int c = 0; c < n_cols; c += 32) {
          uint32_t r[32];
          const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(g * n_cols + c);
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
              "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
              : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
              : "r"(taddr));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (row_ok) red_row32(dst + c, r, n_tile * n_cols + c, p.Cb, p.vec != 0);
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
  const char* halo_env = getenv("LSI_B200_WGRAD_HALO");      // "0": the per-tap kernel for every shape (A/B runs, tests)
  const bool use_halo = !(halo_env && halo_env[0] == '0');
  if (use_halo && d->stride == 1) {
    wg::HParams h;
    h.vec = (((uintptr_t)dw & 15) == 0 && d->c_out % 4 == 0) ? 1 : 0;
    h.dw = dw; h.Ca = d->c_in; h.Cb = d->c_out; h.kh = d->kh; h.kw = d->kw;
    h.chunks_a = (h.Ca + wg::kBlk - 1) / wg::kBlk;
    h.n_blocks = (h.Cb + wg::kBlk - 1) / wg::kBlk;
    if (h.n_blocks > 4) h.n_blocks = 4;
    if (h.n_blocks == 3) h.n_blocks = 4;
    const int n_cols = h.n_blocks * wg::kBlk;
    h.pad_t = d->pad_top; h.pad_l = d->pad_left;
    h.tiles_x = (d->w_out + wg::kTileW - 1) / wg::kTileW; h.tiles_y = (d->h_out + wg::kTileH - 1) / wg::kTileH;
    h.pix_tiles = h.tiles_x * h.tiles_y * d->batch;
    h.mt_per_row = (h.kw + 3) / 4;
    h.m_tiles = h.kh * h.mt_per_row;
    h.G = 512 / n_cols;
    if (h.G > h.m_tiles) h.G = h.m_tiles;
    h.m_groups = (h.m_tiles + h.G - 1) / h.G;
    h.G = (h.m_tiles + h.m_groups - 1) / h.m_groups;          // balance the groups
    h.n_tiles = (h.Cb + n_cols - 1) / n_cols;
    h.halo_w = wg::kTileW + h.kw - 1;
    const int halo_h = wg::kTileH + h.kh - 1;
    h.halo_bytes = h.halo_w * halo_h * 128;
    h.a_bytes = (h.halo_bytes + 512 + 1023) & ~1023;          // the junk fourth block reads up to 3 rows past the box
    const size_t stage_bytes = (size_t)h.a_bytes + (size_t)h.n_blocks * wg::kBlkBytes;
    // small stages: two co-resident CTAs per SM (one's epilogue / pipeline fill overlaps the other's MMAs), else one; the
    // grid is ONE wave of resident CTAs -- every extra split costs a full dW tile of reductions
    const bool two_per_sm = stage_bytes * 2 + 2048 <= 100 * 1024 && h.G * n_cols <= 256;
    h.stages = two_per_sm ? 2 : (int)((200 * 1024) / stage_bytes);
    if (h.stages > 4) h.stages = 4;
    if (h.stages >= 2 && h.halo_w <= 256 && halo_h <= 256) {
      const int mn = h.m_groups * h.chunks_a * h.n_tiles;
      const int resident = two_per_sm ? 296 : 148;
      int splits = resident / mn;
      if (splits > h.pix_tiles) splits = h.pix_tiles;
      if (splits < 1) splits = 1;
      if (splits > 65535) splits = 65535;
      h.tiles_per_split = (h.pix_tiles + splits - 1) / splits;
      splits = (h.pix_tiles + h.tiles_per_split - 1) / h.tiles_per_split;
      auto make = [&](CUtensorMap* m, const float* base, int channels, int cs, int hh, int ww, int bw, int bh) -> int {
        cuuint64_t dims[4] = {(cuuint64_t)channels, (cuuint64_t)ww, (cuuint64_t)hh, (cuuint64_t)d->batch};
        cuuint64_t strides[3] = {(cuuint64_t)cs * 4, (cuuint64_t)ww * cs * 4, (cuuint64_t)hh * ww * cs * 4};
        cuuint32_t box[4] = {(cuuint32_t)wg::kBlk, (cuuint32_t)bw, (cuuint32_t)bh, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = encode(m, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 4, const_cast<float*>(base), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed: %d", (int)r); return LSI_B200_ECUDA; }
        return LSI_B200_OK;
      };
      CUtensorMap map_big, map_small;
      if (int rc = make(&map_big, big, d->c_in, d->in_c_stride, d->h_in, d->w_in, h.halo_w, halo_h)) return rc;
      if (int rc = make(&map_small, small, d->c_out, d->out_c_stride, d->h_out, d->w_out, wg::kTileW, wg::kTileH)) return rc;
      if (!d->accumulate) LSI_CUDA(cudaMemsetAsync(dw, 0, (size_t)h.kh * h.kw * h.Ca * h.Cb * sizeof(float), st));
      const size_t smem = (size_t)h.stages * stage_bytes + 256 + 1024;
      static size_t smem_set_h = 0;
      if (smem > smem_set_h) {
        LSI_CUDA(cudaFuncSetAttribute(wg::wgrad_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set_h = smem;
      }
      dim3 grid((unsigned)mn, (unsigned)splits);
      {
        ScopedTiming tm(kWgrad, st);
        wg::wgrad_halo_kernel<<<grid, wg::kThreads, smem, st>>>(map_big, map_small, h);
      }
      LSI_LAUNCH_CHECK();
      return LSI_B200_OK;
    }
  }
  wg::Params p;
  p.vec = (((uintptr_t)dw & 15) == 0 && d->c_out % 4 == 0) ? 1 : 0;
  p.dw = dw; p.Ca = d->c_in; p.Cb = d->c_out; p.taps = d->kh * d->kw; p.kw = d->kw;
  p.chunks_a = (p.Ca + wg::kBlk - 1) / wg::kBlk;
  p.n_blocks = p.Cb > 32 ? 2 : 1;
  p.stride = d->stride; p.pad_t = d->pad_top; p.pad_l = d->pad_left;
  p.tiles_x = (d->w_out + wg::kTileW - 1) / wg::kTileW; p.tiles_y = (d->h_out + wg::kTileH - 1) / wg::kTileH; p.batch = d->batch;
  p.pix_tiles = p.tiles_x * p.tiles_y * d->batch;
  p.m_tiles = (p.taps * p.chunks_a + wg::kMBlocks - 1) / wg::kMBlocks;
  p.n_tiles = (p.Cb + p.n_blocks * wg::kBlk - 1) / (p.n_blocks * wg::kBlk);
  const int mn = p.m_tiles * p.n_tiles;
  int splits = 148 / mn;                      // one wave of resident CTAs (192 KB of stages: one CTA per SM)
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
