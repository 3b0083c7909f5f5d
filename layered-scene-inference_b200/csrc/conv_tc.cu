// Tensor-core convolution for B200 (sm_100a): implicit GEMM on tcgen05 with TMA-fed shared-memory operands and
// TMEM accumulators.  Same descriptor and semantics as lsi_b200_conv2d (conv.cu), used for every layer whose summed
// channel count is a multiple of 32 -- i.e. everything but the 3-channel stem (nets.py:273).
//
//   GEMM view   D[128 pixels x N channels] += A[128 x K] * B[N x K]^T,  K = taps x Cin, fp32 data fed as TF32
//   A operand   activations, NHWC fp32.  A CTA owns an 8 x 16 patch of output pixels of one image; for tap (ky,kx) and
//               channel chunk c0 the im2col tile is the same patch shifted by the tap, i.e. ONE 4-D TMA box
//               (32 ch x 16 x 8 x 1) whose out-of-bounds elements TMA zero-fills (= TF SAME padding, asymmetric or
//               not); stride-2 convs use the tensor map's element strides; the 4x4 stride-2 up-convolution and conv
//               data gradients run phase-decomposed (4 valid taps per output phase).  Two tensor maps implement the
//               channel concat of the U-Net skip connections without materialising it.
//   B operand   weights re-laid out once per call as [tap][N_pad][Cin] (K-major) by a small kernel, 2-D TMA box.
//   shared mem  128-byte rows with the 128B swizzle (TMA writes it, the UMMA descriptors read it), 4-stage
//               full/empty mbarrier ring;  warp 0 = TMA producer, warp 1 = TMEM owner + single-thread MMA issuer,
//               warps 2-5 = epilogue (tcgen05.ld -> bias/sigmoid/accumulate -> 128-bit global stores).
//   accumulate  TMEM, 128 lanes x N fp32 columns.
#include <cuda.h>
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


constexpr int kTileH = 8, kTileW = 16, kTileM = kTileH * kTileW;   // 128 output pixels = 128 TMEM lanes
constexpr int kKC = 32;                                            // fp32 channels per K chunk = one 128-byte row
constexpr int kMaxStages = 8;
constexpr int kThreads = 192;

struct TcParams {
  float* out; const float* bias;
  int Ho, Wo, Co, out_cs;
  int Hp, Wp;                 // per-phase output extent
  int tiles_x, tiles_y;       // patches per image (phase space)
  int Ca, Cb;                 // channels of source A / source B (concat), Cb may be 0
  int kh, kw, stride, pad_t, pad_l, mode;
  int n_tile, n_pad;          // UMMA N, padded Cout
  int epilogue, accumulate;
  int batch, total_tiles;     // persistent tile loop
  int xm;                     // x-merge: 16x8 pixel tiles, ONE halo load per (ky, chunk) feeds all kx taps through shifted UMMA descriptors
  int halo_w;                 // xm: pixels per halo row (tile width 8 + taps along x - 1, or padded to 16)
  int th, tw;                 // tile height / width in pixels (8x16 default, 16x8 with xm)
  int tn;                     // images per tile (th * tw * tn = 128): the 2x7 / 4x14 layers at the U-Net bottleneck fill a tile
                              // with 8 / 2 whole images instead of leaving 89 % / 56 % of its rows empty
  float* stat_part;           // [gridDim.x*4][n_pad][2] per-(CTA,warp) channel sums of the output (batch-norm statistics), or NULL
  int h16;                    // 1: fp16 activations and weights (64-byte rows, 64B swizzle, kind::f16 MMAs, K = 16)
                              // 2: split-precision fp16 pairs (split.cuh): 128-byte rows [hi 32 ch | lo 32 ch], 128B swizzle, kind::f16
                              //    MMAs over K = 64 against the doubled filter tile (rows [Whi|0] then [Wlo|Whi]); two accumulators
  int out_f16;                // 1: plain outputs are stored as fp16; 2: as split fp16 pairs (chunk-interleaved, same bytes as fp32)
  const float* out_scale;     // optional per-channel factor applied after the activation (bias / sigmoid epilogues), or NULL
  int stages;                 // smem ring depth (2..4): shallower rings let 2-3 CTAs share an SM so that one CTA's
                              // prologue/epilogue overlaps another's main loop
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
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// K-major, 128B-swizzled operand tile: 8-row groups 1024 bytes apart (cute::UMMA::SmemDescriptor, version 1)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);         // start address
  d |= (uint64_t)1 << 16;                         // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;               // stride byte offset
  d |= (uint64_t)1 << 46;                         // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                         // SWIZZLE_128B
  return d;
}
// same, for a tile whose 8-row groups are `sbo` bytes apart and whose start address is shifted by whole 128-byte rows
// (x-merged halo tiles).  Measured on B200: the tensor core derives the 128B-swizzle phase from the absolute shared-memory
// address bits, exactly like TMA wrote it, so the base-offset field stays 0 (setting it to the row phase gives wrong
// results) and SBO need not be a multiple of 1024 (1280 = a 10-pixel halo row works).
__device__ __forceinline__ uint64_t umma_desc_shifted(uint32_t saddr, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// generic form: layout 2 = SWIZZLE_128B (128-byte rows), 4 = SWIZZLE_64B (64-byte rows: 32 fp16 channels); the shifted-start
// property above holds for both (tests/test_gpu_conv_tc.py, test_gpu_conv_halo.py)
__device__ __forceinline__ uint64_t umma_desc_any(uint32_t saddr, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ uint32_t pack_half2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
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

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// true for exactly one lane of a converged warp; unlike `lane == 0` it tells ptxas that a single thread runs the
// guarded region, so the tcgen05/TMA operands are built in uniform registers without a per-lane waterfall loop
// (measured: ~30 SASS instructions and ~130 cycles per tcgen05.mma with `lane == 0`, 2-6 instructions with elect)
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

struct TileCoord {
  int n_img, y0, x0, n0, py, px, ky0, kx0, nky, nkx;
};

// persistent tile index -> (phase, Cout tile, image, patch); phases and Cout tiles are the slow dimensions so that
// concurrently running CTAs share the same weights
__device__ __forceinline__ TileCoord tile_coord(const TcParams& p, int tile, int s) {
  TileCoord c;
  const int per_img = p.tiles_x * p.tiles_y;
  const int spatial = per_img * ((p.batch + p.tn - 1) / p.tn);
  const int n_tiles_n = p.n_pad / p.n_tile;
  const int sp = tile % spatial; int rest = tile / spatial;
  const int nt = rest % n_tiles_n; const int ph = rest / n_tiles_n;
  c.n_img = sp / per_img;
  const int r = sp - c.n_img * per_img;
  c.n_img *= p.tn;
  c.y0 = (r / p.tiles_x) * p.th; c.x0 = (r % p.tiles_x) * p.tw;
  c.n0 = nt * p.n_tile;
  c.py = (p.mode == 1) ? ph / s : 0; c.px = (p.mode == 1) ? ph % s : 0;
  c.ky0 = (p.mode == 1) ? ((c.py + p.pad_t) % s) : 0; c.kx0 = (p.mode == 1) ? ((c.px + p.pad_l) % s) : 0;
  c.nky = (p.kh - c.ky0 + s - 1) / s; c.nkx = (p.kw - c.kx0 + s - 1) / s;
  return c;
}

// Persistent CTAs: each loops over output tiles.  The TMA producer runs ahead across tile boundaries through the
// shared-memory ring; the MMA issuer alternates between two TMEM accumulator buffers so that the epilogue warps drain
// tile i while tile i+1 is being accumulated (small-K layers -- 3x3x32 -- are prologue/epilogue bound otherwise).
// kSplit: the split-precision mode (p.h16 == 2 / 3) is compiled separately so that the TF32 / fp16 instantiation keeps its code
template <bool kSplit>
__global__ void __launch_bounds__(kThreads)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ CUtensorMap map_w, const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages][A 16 KB][B n_tile*128 B] then barriers
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t rb = (p.h16 == 1) ? 64u : 128u; // bytes of one 32-channel activation row
  const bool split = kSplit;                     // p.h16 2: variant S (128-byte filter rows [Whi|0];[Wlo|Whi]), 3: variant L (64-byte filter rows Whi;Wlo)
  const bool split_l = kSplit && p.h16 == 3;
  const int n_mma = split ? 2 * p.n_tile : p.n_tile;   // UMMA N (split: D0 = hi*Whi in columns [0,N), D1 = cross terms in [N,2N))
  const int cm = split ? 2 : 1;                  // fp16 elements per channel in the activation tensor maps of the split layout
  const int cmw = (p.h16 == 2) ? 2 : 1;          // ... and in the filter map (variant S only)
  const uint32_t rbw = (p.h16 == 1 || split_l) ? 64u : 128u;   // bytes of one filter row
  const uint32_t b_tap_bytes = ((uint32_t)n_mma * rbw + 1023) & ~1023u;
  const uint32_t a_bytes = p.xm ? (uint32_t)p.halo_w * p.th * rb : kTileM * rb;
  const uint32_t b_bytes = (uint32_t)n_mma * rbw;
  const uint32_t stage_bytes = p.xm ? a_bytes + (uint32_t)p.kw * b_tap_bytes : a_bytes + b_tap_bytes;   // xm: up to kw weight tiles
  const int kStages = p.stages;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kStages * stage_bytes);
  uint64_t* empty = full + kMaxStages;
  uint64_t* tmem_full = empty + kMaxStages;      // [2]
  uint64_t* tmem_empty = tmem_full + 2;          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* stat_stage = reinterpret_cast<float*>(smem + kStages * stage_bytes + 256);   // [4 warps][32][33]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s = (p.mode == 1) ? p.stride : 1;
  const int chunks = (p.Ca + p.Cb) / kKC;

  uint32_t acc_cols = 32;                        // columns of one accumulator buffer
  while ((int)acc_cols < n_mma) acc_cols <<= 1;
  const uint32_t tmem_cols = acc_cols * 2;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    for (int i = 0; i < kStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      // ---------------- TMA producer ----------------
      int st = 0; uint32_t ph = 0;               // ring position, continues across tiles (no integer division per stage:
                                                 // this thread is one dependent instruction stream)
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const TileCoord c = tile_coord(p, tile, s);
        const int k_iters = p.xm ? c.nky * chunks : c.nky * c.nkx * chunks;
        int kyi_c = 0, ch_c = 0;                 // k = kyi_c * chunks + ch_c (xm) / tap * chunks + ch_c
        for (int k = 0; k < k_iters; ++k, st = (st + 1 == kStages ? 0 : st + 1), ph ^= (st == 0 ? 1u : 0u), ch_c = (ch_c + 1 == chunks ? 0 : ch_c + 1), kyi_c += (ch_c == 0 ? 1 : 0)) {
          mbar_wait(&empty[st], ph ^ 1);
          uint8_t* sa = smem + st * stage_bytes;
          uint8_t* sb = sa + a_bytes;
          if (p.xm) {
            // one halo tile (halo_w x th pixels) per (ky, chunk); tap j along x reads it shifted by off_j pixels
            const int kyi = kyi_c, c0 = ch_c * kKC;
            const int ky = c.ky0 + kyi * s;
            int ys, xs_min;
            if (p.mode == 0) { ys = c.y0 - p.pad_t + ky; xs_min = c.x0 - p.pad_l; }
            else { ys = c.y0 + (c.py + p.pad_t - ky) / s; xs_min = c.x0 + (c.px + p.pad_l - (c.kx0 + (c.nkx - 1) * s)) / s; }
            mbar_expect_tx(&full[st], a_bytes + (uint32_t)c.nkx * b_bytes);
            if (c0 < p.Ca) tma_load_4d(sa, &map_a, &full[st], c0 * cm, xs_min, ys, c.n_img);
            else tma_load_4d(sa, &map_b, &full[st], (c0 - p.Ca) * cm, xs_min, ys, c.n_img);
            for (int j = 0; j < c.nkx; ++j) {
              const int kx = c.kx0 + j * s;
              tma_load_2d(sb + j * b_tap_bytes, &map_w, &full[st], c0 * cmw, ((ky * p.kw + kx) * p.n_pad + c.n0) * cm);
            }
            continue;
          }
          const int tap = kyi_c, c0 = ch_c * kKC;
          const int ky = c.ky0 + (tap / c.nkx) * s, kx = c.kx0 + (tap % c.nkx) * s;
          int ys, xs;
          if (p.mode == 0) { ys = c.y0 * p.stride - p.pad_t + ky; xs = c.x0 * p.stride - p.pad_l + kx; }
          else { ys = c.y0 + (c.py + p.pad_t - ky) / s; xs = c.x0 + (c.px + p.pad_l - kx) / s; }   // exact: tap list matches the phase
          mbar_expect_tx(&full[st], a_bytes + b_bytes);
          if (c0 < p.Ca) tma_load_4d(sa, &map_a, &full[st], c0 * cm, xs, ys, c.n_img);
          else tma_load_4d(sa, &map_b, &full[st], (c0 - p.Ca) * cm, xs, ys, c.n_img);
          tma_load_2d(sb, &map_w, &full[st], c0 * cmw, ((ky * p.kw + kx) * p.n_pad + c.n0) * cm);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ---------------- MMA issuer ----------------
      // instruction descriptor (cute::UMMA::InstrDescriptor): D=F32, A=B=TF32, K-major both, N>>3, M>>4
      // (kind::f16: A/B format F16 = 0, two K = 16 steps per 32-channel chunk)
      const uint32_t idesc = p.h16 ? ((1u << 4) | ((uint32_t)(n_mma >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24))
                                   : ((1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.n_tile >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24));
      const uint32_t layout = (p.h16 == 1) ? 4u : 2u, sbo_dense = (p.h16 == 1) ? 512u : 1024u;
      // variant L: filter tile = 2N rows of 64 bytes (64B swizzle): rows [0,N) = Whi, [N,2N) = Wlo.  Per chunk: a_hi x (Whi;Wlo) with
      // N' = 2N into columns [0,2N), then a_lo x Whi with N' = N into columns [N,2N): 3N instead of 4N columns of MMA work, half the
      // filter bytes.  The activation operand keeps its 128-byte rows / 128B swizzle (each descriptor carries its own layout).
      const uint32_t idesc_n = (1u << 4) | ((uint32_t)(p.n_tile >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
      const uint32_t layout_b = split_l ? 4u : layout, sbo_b = split_l ? 512u : sbo_dense;
      int st = 0, tcount = 0; uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tcount) {
        const TileCoord c = tile_coord(p, tile, s);
        const int k_iters = p.xm ? c.nky * chunks : c.nky * c.nkx * chunks;
        const int buf = tcount & 1;
        mbar_wait(&tmem_empty[buf], ((tcount >> 1) & 1) ^ 1);       // epilogue has drained this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_d = tmem_base + (uint32_t)buf * acc_cols;
        for (int k = 0; k < k_iters; ++k, st = (st + 1 == kStages ? 0 : st + 1), ph ^= (st == 0 ? 1u : 0u)) {
          mbar_wait(&full[st], ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          // descriptors advance by adding (byte offset >> 4) to the start-address field: every operand lives below
          // 256 KB, so there is no carry out of its 14 bits
          const uint32_t sa = smem_u32(smem + st * stage_bytes), sb = sa + a_bytes;
          if (p.xm) {
            const uint64_t ad0 = umma_desc_any(sa, (uint32_t)p.halo_w * rb, layout), bd0 = umma_desc_any(sb, sbo_b, layout_b);
            for (int j = 0; j < c.nkx; ++j) {
              const uint32_t off = (p.mode == 0) ? (uint32_t)j : (uint32_t)(c.nkx - 1 - j);   // pixels into the halo row
              const uint64_t ad = ad0 + (uint64_t)(off * (rb >> 4)), bd = bd0 + (uint64_t)(((uint32_t)j * b_tap_bytes) >> 4);
              if (split_l) {
#pragma unroll
                for (int kk = 0; kk < 2; ++kk)
                  umma_f16(tmem_d, ad + 2 * kk, bd + 2 * kk, idesc, (k | j | kk) != 0);
#pragma unroll
                for (int kk = 0; kk < 2; ++kk)
                  umma_f16(tmem_d + (uint32_t)p.n_tile, ad + 4 + 2 * kk, bd + 2 * kk, idesc_n, 1u);
              } else if (split) {        // K = 64 fp16 per 128-byte row: [hi | lo] x ([Whi | 0] ; [Wlo | Whi])
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                  umma_f16(tmem_d, ad + 2 * kk, bd + 2 * kk, idesc, (k | j | kk) != 0);
              } else if (p.h16) {
#pragma unroll
                for (int kk = 0; kk < kKC / 16; ++kk)
                  umma_f16(tmem_d, ad + 2 * kk, bd + 2 * kk, idesc, (k | j | kk) != 0);
              } else {
#pragma unroll
                for (int kk = 0; kk < kKC / 8; ++kk)
                  umma_tf32(tmem_d, ad + 2 * kk, bd + 2 * kk, idesc, (k | j | kk) != 0);
              }
            }
          } else {
            const uint64_t ad = umma_desc_any(sa, sbo_dense, layout), bd = umma_desc_any(sb, sbo_b, layout_b);
            if (split_l) {
#pragma unroll
              for (int kk = 0; kk < 2; ++kk)
                umma_f16(tmem_d, ad + 2 * kk, bd + 2 * kk, idesc, (k | kk) != 0);
#pragma unroll
              for (int kk = 0; kk < 2; ++kk)
                umma_f16(tmem_d + (uint32_t)p.n_tile, ad + 4 + 2 * kk, bd + 2 * kk, idesc_n, 1u);
            } else if (split) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                umma_f16(tmem_d, ad + 2 * kk, bd + 2 * kk, idesc, (k | kk) != 0);
            } else if (p.h16) {
#pragma unroll
              for (int kk = 0; kk < kKC / 16; ++kk)     // UMMA K = 16 for fp16: 32 bytes along the swizzled row
                umma_f16(tmem_d, ad + 2 * kk, bd + 2 * kk, idesc, (k | kk) != 0);
            } else {
#pragma unroll
              for (int kk = 0; kk < kKC / 8; ++kk)      // UMMA K = 8 for TF32: 32 bytes along the swizzled row
                umma_tf32(tmem_d, ad + 2 * kk, bd + 2 * kk, idesc, (k | kk) != 0);
            }
          }
          umma_commit(&empty[st]);                 // frees the stage once these MMAs have read it
        }
        umma_commit(&tmem_full[buf]);              // accumulator complete
      }
    }
  } else {
    // ---------------- epilogue: TMEM -> registers -> global ----------------
    const int lg = warp & 3;                     // TMEM lane group this warp may access
    const int row = lg * 32 + lane;              // = A tile row = pixel within the patch
    const int hy_full = row / p.tw, wx = row % p.tw;
    const int n_off = hy_full / p.th, hy = hy_full - n_off * p.th;   // (image within the tile, row, column)
    float ssum[4] = {0.f, 0.f, 0.f, 0.f}, ssq[4] = {0.f, 0.f, 0.f, 0.f};   // channel (n0 + 32*i + lane) sums of this warp's pixels
    int stat_n0 = -1;
    float* stg = stat_stage + (size_t)lg * 32 * 33;
    float* my_part = p.stat_part ? p.stat_part + ((size_t)blockIdx.x * 4 + lg) * p.n_pad * 2 : nullptr;
    auto flush_stats = [&]() {
      if (my_part && stat_n0 >= 0) {
        for (int i = 0; i < 4; ++i) {
          const int ch = stat_n0 + 32 * i + lane;
          if (32 * i < p.n_tile && ch < p.n_pad && (32 * i + lane) < p.n_tile) { my_part[2 * ch] += ssum[i]; my_part[2 * ch + 1] += ssq[i]; }
          ssum[i] = 0.f; ssq[i] = 0.f;
        }
      }
    };
    int tcount = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tcount) {
      const TileCoord c = tile_coord(p, tile, s);
      const int buf = tcount & 1;
      if (my_part && c.n0 != stat_n0) { flush_stats(); stat_n0 = c.n0; }
      mbar_wait(&tmem_full[buf], (tcount >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      int oy = c.y0 + hy, ox = c.x0 + wx;
      const int n_out = c.n_img + n_off;
      const bool in_range = oy < p.Hp && ox < p.Wp && n_out < p.batch;
      if (p.mode == 1) { oy = oy * s + c.py; ox = ox * s + c.px; }
      float* dst = p.out + ((size_t)(n_out * p.Ho + oy) * p.Wo + ox) * p.out_cs + c.n0;
      for (int cc = 0; cc < p.n_tile; cc += 32) {
        uint32_t r[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)buf * acc_cols + (uint32_t)cc;
        if (p.n_tile - cc >= 32) {
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
              "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
              : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
              : "r"(taddr));
        } else {   // n_tile == 16
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
              : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
              : "r"(taddr));
#pragma unroll
          for (int j = 16; j < 32; ++j) r[j] = 0u;
        }
        if (split) {   // second accumulator (cross terms, scaled by 2^11): columns [n_tile + cc, ...)
          uint32_t r1[32];
          const uint32_t taddr1 = taddr + (uint32_t)p.n_tile;
          if (p.n_tile - cc >= 32) {
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(r1[0]), "=r"(r1[1]), "=r"(r1[2]), "=r"(r1[3]), "=r"(r1[4]), "=r"(r1[5]), "=r"(r1[6]), "=r"(r1[7]), "=r"(r1[8]),
                  "=r"(r1[9]), "=r"(r1[10]), "=r"(r1[11]), "=r"(r1[12]), "=r"(r1[13]), "=r"(r1[14]), "=r"(r1[15]), "=r"(r1[16]),
                  "=r"(r1[17]), "=r"(r1[18]), "=r"(r1[19]), "=r"(r1[20]), "=r"(r1[21]), "=r"(r1[22]), "=r"(r1[23]), "=r"(r1[24]),
                  "=r"(r1[25]), "=r"(r1[26]), "=r"(r1[27]), "=r"(r1[28]), "=r"(r1[29]), "=r"(r1[30]), "=r"(r1[31])
                : "r"(taddr1));
          } else {   // n_tile == 16
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                : "=r"(r1[0]), "=r"(r1[1]), "=r"(r1[2]), "=r"(r1[3]), "=r"(r1[4]), "=r"(r1[5]), "=r"(r1[6]), "=r"(r1[7]), "=r"(r1[8]),
                  "=r"(r1[9]), "=r"(r1[10]), "=r"(r1[11]), "=r"(r1[12]), "=r"(r1[13]), "=r"(r1[14]), "=r"(r1[15])
                : "r"(taddr1));
#pragma unroll
            for (int j = 16; j < 32; ++j) r1[j] = 0u;
          }
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 32; ++j)
            r[j] = __float_as_uint(fmaf(__uint_as_float(r1[j]), kSplitInvScale, __uint_as_float(r[j])));
        } else {
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        }
        if (cc + 32 >= p.n_tile) {               // last read of this accumulator: hand the buffer back to the MMA issuer
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          mbar_arrive(&tmem_empty[buf]);
        }
        if (my_part) {   // per-channel sum / sum of squares over this warp's 32 pixels, via a padded smem transpose
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 32; ++j) stg[lane * 33 + j] = (in_range && j < p.n_tile - cc) ? __uint_as_float(r[j]) : 0.f;
          __syncwarp();
          float a = 0.f, b2 = 0.f;
#pragma unroll 8
          for (int rr = 0; rr < 32; ++rr) { const float v = stg[rr * 33 + lane]; a += v; b2 = fmaf(v, v, b2); }
          ssum[cc >> 5] += a; ssq[cc >> 5] += b2;
        }
        if (in_range && p.out_f16 == 2) {   // plain conv output as split fp16 pairs: this chunk's 128 bytes = [hi 32 ch | lo 32 ch]
          uint4* d4 = reinterpret_cast<uint4*>(dst + cc);     // byte address of (pixel, chunk) is the same as in an fp32 tensor
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) split_pack2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]), hi[j], lo[j]);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            d4[j] = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
            d4[4 + j] = make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
          }
        } else if (in_range && p.out_f16) {   // plain conv output (host guarantees epilogue 0, no accumulate, Co % 8 == 0), stored as fp16
          const int nvalid = min(min(32, p.n_tile - cc), p.Co - (c.n0 + cc));
          __half* dh = reinterpret_cast<__half*>(p.out) + ((size_t)(n_out * p.Ho + oy) * p.Wo + ox) * p.out_cs + c.n0 + cc;
#pragma unroll
          for (int j = 0; j < 32; j += 8)
            if (j < nvalid)
              *reinterpret_cast<uint4*>(dh + j) = make_uint4(pack_half2(__uint_as_float(r[j]), __uint_as_float(r[j + 1])),
                                                             pack_half2(__uint_as_float(r[j + 2]), __uint_as_float(r[j + 3])),
                                                             pack_half2(__uint_as_float(r[j + 4]), __uint_as_float(r[j + 5])),
                                                             pack_half2(__uint_as_float(r[j + 6]), __uint_as_float(r[j + 7])));
        } else if (in_range) {
          const int nvalid = min(min(32, p.n_tile - cc), p.Co - (c.n0 + cc));
          if (nvalid == 32 && p.epilogue == 0 && !p.accumulate && ((reinterpret_cast<uintptr_t>(dst + cc) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(dst + cc + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                                     __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (j < nvalid) {
                float v = __uint_as_float(r[j]);
                if (p.epilogue >= 1) v += __ldg(p.bias + c.n0 + cc + j);
                if (p.epilogue == 2) v = 1.f / (1.f + expf(-v));
                if (p.out_scale) v *= __ldg(p.out_scale + c.n0 + cc + j);
                if (p.accumulate) v += dst[cc + j];
                dst[cc + j] = v;
              }
            }
          }
        }
      }
    }
    flush_stats();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// weights (any strides) -> [tap][n_pad][cin] fp32, zero rows for co >= Cout
__global__ void __launch_bounds__(256) prep_weights_kernel(const float* __restrict__ w, float* __restrict__ wk, int taps, int cin,
                                                           int cout, int n_pad, int w_tap, int w_ci, int w_co) {
  const long long total = (long long)taps * n_pad * cin;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % cin);
    const int co = (int)((i / cin) % n_pad);
    const int tap = (int)(i / ((long long)cin * n_pad));
    wk[i] = (co < cout) ? w[(size_t)tap * w_tap + (size_t)ci * w_ci + (size_t)co * w_co] : 0.f;
  }
}

// same, rounded to fp16 (operand type of the kind::f16 path)
__global__ void __launch_bounds__(256) prep_weights_f16_kernel(const float* __restrict__ w, __half* __restrict__ wk, int taps, int cin,
                                                               int cout, int n_pad, int w_tap, int w_ci, int w_co) {
  const long long total = (long long)taps * n_pad * cin;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % cin);
    const int co = (int)((i / cin) % n_pad);
    const int tap = (int)(i / ((long long)cin * n_pad));
    wk[i] = __float2half_rn((co < cout) ? w[(size_t)tap * w_tap + (size_t)ci * w_ci + (size_t)co * w_co] : 0.f);
  }
}

// split mode (split.cuh): [tap][Cout tile][half][n in tile][chunk][64 fp16]; half 0 rows = [w_hi | 0], half 1 rows = [w_lo | w_hi]
__global__ void __launch_bounds__(256) prep_weights_split_kernel(const float* __restrict__ w, __half* __restrict__ wk, int taps, int cin,
                                                                 int cout, int n_pad, int n_tile, int w_tap, int w_ci, int w_co) {
  const long long total = (long long)taps * 2 * n_pad * 2 * cin;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % (2 * cin));
    const long long row = i / (2 * cin);
    const int ci = (k >> 6) * 32 + (k & 31), part = (k >> 5) & 1;
    const int tap = (int)(row / (2 * n_pad)), rr = (int)(row % (2 * n_pad));
    const int nt = rr / (2 * n_tile), r2 = rr % (2 * n_tile);
    const int half = r2 / n_tile, co = nt * n_tile + (r2 % n_tile);
    const float v = (co < cout) ? w[(size_t)tap * w_tap + (size_t)ci * w_ci + (size_t)co * w_co] : 0.f;
    const __half hi = __float2half_rn(fminf(fmaxf(v, -kSplitMax), kSplitMax));
    const __half lo = __float2half_rn((v - __half2float(hi)) * kSplitScale);
    wk[i] = half == 0 ? (part == 0 ? hi : __float2half_rn(0.f)) : (part == 0 ? lo : hi);
  }
}

// split mode, variant L: [tap][Cout tile][half][n in tile][cin] fp16; half 0 rows = w_hi, half 1 rows = w_lo
__global__ void __launch_bounds__(256) prep_weights_split_l_kernel(const float* __restrict__ w, __half* __restrict__ wk, int taps, int cin,
                                                                   int cout, int n_pad, int n_tile, int w_tap, int w_ci, int w_co) {
  const long long total = (long long)taps * 2 * n_pad * cin;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % cin);
    const long long row = i / cin;
    const int tap = (int)(row / (2 * n_pad)), rr = (int)(row % (2 * n_pad));
    const int nt = rr / (2 * n_tile), r2 = rr % (2 * n_tile);
    const int half = r2 / n_tile, co = nt * n_tile + (r2 % n_tile);
    const float v = (co < cout) ? w[(size_t)tap * w_tap + (size_t)ci * w_ci + (size_t)co * w_co] : 0.f;
    const __half hi = __float2half_rn(fminf(fmaxf(v, -kSplitMax), kSplitMax));
    wk[i] = half == 0 ? hi : __float2half_rn((v - __half2float(hi)) * kSplitScale);
  }
}

static bool split_variant_l() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("LSI_B200_SPLIT_VARIANT"); on = (e && (e[0] == 'S' || e[0] == 's')) ? 0 : 1; }
  return on == 1;
}

static bool xmerge_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("LSI_B200_CONV_XMERGE"); on = (e && atoi(e) == 0) ? 0 : 1; }
  return on == 1;
}
static bool batch_tiles_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("LSI_B200_CONV_BATCH_TILES"); on = (e && atoi(e) == 0) ? 0 : 1; }
  return on == 1;
}
static bool xmerge_tight() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("LSI_B200_CONV_XMERGE_TIGHT"); on = (e && atoi(e) == 0) ? 0 : 1; }
  return on == 1;
}

static int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
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

}  // namespace lsi

using namespace lsi;

static size_t stat_part_bytes(size_t n_pad) { return (size_t)148 * 4 * 4 * n_pad * 2 * sizeof(float) * 2; }   // generous: <= 4 CTAs/SM

extern "C" size_t lsi_b200_conv2d_tc_workspace_bytes(const lsi_b200_conv_desc* d) {
  if (!d) return 0;
  const size_t n_pad = (size_t)(d->c_out + 15) / 16 * 16;
  return 2 * (size_t)d->kh * d->kw * n_pad * (size_t)d->c_in * sizeof(float) + 512 + stat_part_bytes(n_pad);   // x2: split-mode filter tiles
}

extern "C" int lsi_b200_conv2d_tc_supported(const lsi_b200_conv_desc* d, int c_in_a) {
  if (!d) return 0;
  if (d->c_in % kKC != 0 || c_in_a % kKC != 0 || c_in_a > d->c_in || c_in_a < 1) return 0;
  if (d->stride < 1 || d->stride > 2 || (d->mode != 0 && d->mode != 1)) return 0;
  if (d->mode == 1 && (d->h_out % d->stride || d->w_out % d->stride)) return 0;
  if (d->in_c_stride % 4 != 0) return 0;
  const int n_pad = (d->c_out + 15) / 16 * 16;
  int n_tile = n_pad <= 128 ? n_pad : 128;
  if (n_pad > 128 && n_pad % 128 != 0) n_tile = 64;
  if (n_pad % n_tile != 0 || !(n_tile == 16 || n_tile % 32 == 0)) return 0;
  return 1;
}

// Same contract as lsi_b200_conv2d, plus an optional second input source: channels [0, c_in_a) come from `in_a`
// (pixel stride in_c_stride), channels [c_in_a, c_in) from `in_b` (pixel stride in_b_c_stride) -- tf.concat on the fly.
namespace lsi {
__global__ void __launch_bounds__(256) finalize_stats_f32_kernel(const float* __restrict__ partial, int nparts, int n_pad, int C,
                                                                 long long P, float eps, float* __restrict__ out) {
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (c >= C) return;
  double a = 0.0, b = 0.0;
  for (int i = lane; i < nparts; i += 32) { a += (double)partial[((size_t)i * n_pad + c) * 2]; b += (double)partial[((size_t)i * n_pad + c) * 2 + 1]; }
  for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
  if (lane != 0) return;
  const double mean = a / (double)P;
  double var = b / (double)P - mean * mean;
  if (var < 0.0) var = 0.0;
  out[2 * c] = (float)mean; out[2 * c + 1] = (float)(1.0 / sqrt(var + (double)eps));
}
}  // namespace lsi

static int conv2d_tc_impl(const lsi_b200_conv_desc* d, const void* in_a, int c_in_a, const void* in_b, int in_b_c_stride,
                          const float* w, const float* bias, void* out, float* bn_stats, float bn_eps, void* workspace,
                          size_t workspace_bytes, void* stream, int h16 = 0, int out_f16 = 0, const float* out_scale = nullptr);

extern "C" int lsi_b200_conv2d_tc(const lsi_b200_conv_desc* d, const float* in_a, int c_in_a, const float* in_b,
                                  int in_b_c_stride, const float* w, const float* bias, float* out, void* workspace,
                                  size_t workspace_bytes, void* stream) {
  return conv2d_tc_impl(d, in_a, c_in_a, in_b, in_b_c_stride, w, bias, out, nullptr, 0.f, workspace, workspace_bytes, stream);
}

// conv + batch statistics of its output in one pass: bn_stats[c] = (mean, rsqrt(biased var + eps)) over all output pixels
extern "C" int lsi_b200_conv2d_tc_bnstats(const lsi_b200_conv_desc* d, const float* in_a, int c_in_a, const float* in_b,
                                          int in_b_c_stride, const float* w, float* out, float* bn_stats, float bn_eps,
                                          void* workspace, size_t workspace_bytes, void* stream) {
  LSI_REQUIRE(bn_stats != nullptr, "NULL pointer argument");
  LSI_REQUIRE(d && d->epilogue == 0 && d->accumulate == 0, "bn statistics need a plain conv output");
  return conv2d_tc_impl(d, in_a, c_in_a, in_b, in_b_c_stride, w, nullptr, out, bn_stats, bn_eps, workspace, workspace_bytes, stream);
}

static int conv2d_tc_impl(const lsi_b200_conv_desc* d, const void* in_a, int c_in_a, const void* in_b, int in_b_c_stride,
                          const float* w, const float* bias, void* out, float* bn_stats, float bn_eps, void* workspace,
                          size_t workspace_bytes, void* stream, int h16, int out_f16, const float* out_scale) {
  const unsigned long long wver = take_weight_version();   // consumed by this call whatever happens next
  LSI_REQUIRE(d && in_a && w && out && workspace, "NULL pointer argument");
  if (h16 == 2 && split_variant_l()) h16 = 3;
  LSI_REQUIRE(h16 != 1 || (d->in_c_stride % 8 == 0 && (!in_b || in_b_c_stride % 8 == 0)), "fp16 activations need 8-channel-aligned pixel strides");
  LSI_REQUIRE(h16 < 2 || (d->in_c_stride % 32 == 0 && (!in_b || in_b_c_stride % 32 == 0)), "split activations need 32-channel-aligned pixel strides");
  LSI_REQUIRE(out_f16 != 1 || (d->epilogue == 0 && d->accumulate == 0 && d->c_out % 8 == 0 && d->out_c_stride % 8 == 0),
              "fp16 output is for plain conv outputs with a multiple of 8 channels");
  LSI_REQUIRE(out_f16 != 2 || (h16 >= 2 && d->epilogue == 0 && d->accumulate == 0 && d->c_out % 32 == 0 && d->out_c_stride % 32 == 0),
              "split output is for plain split-mode conv outputs with a multiple of 32 channels");
  LSI_REQUIRE(!out_scale || d->epilogue >= 1, "out_scale goes with the bias / sigmoid epilogues");
  LSI_REQUIRE(lsi_b200_conv2d_tc_supported(d, c_in_a), "shape not supported by the tensor-core path");
  LSI_REQUIRE(c_in_a == d->c_in || (in_b && in_b_c_stride % 4 == 0 && in_b_c_stride >= d->c_in - c_in_a), "bad second source");
  LSI_REQUIRE(d->epilogue == 0 || bias, "epilogue needs a bias pointer");
  LSI_REQUIRE(workspace_bytes >= lsi_b200_conv2d_tc_workspace_bytes(d), "workspace too small");
  LSI_REQUIRE(((uintptr_t)in_a & 15) == 0 && (!in_b || ((uintptr_t)in_b & 15) == 0), "inputs must be 16-byte aligned");
  EncodeTiledFn encode = get_encode();
  LSI_REQUIRE(encode != nullptr, "cuTensorMapEncodeTiled is unavailable (driver too old?)");
  cudaStream_t st = as_stream(stream);

  TcParams p;
  p.out = static_cast<float*>(out); p.bias = bias; p.out_scale = out_scale; p.h16 = h16; p.out_f16 = out_f16; p.Ho = d->h_out; p.Wo = d->w_out; p.Co = d->c_out; p.out_cs = d->out_c_stride;
  const int s = d->mode == 1 ? d->stride : 1;
  p.Hp = d->h_out / s; p.Wp = d->w_out / s;
  p.n_pad = (d->c_out + 15) / 16 * 16;
  p.n_tile = p.n_pad <= 128 ? p.n_pad : 128;
  if (p.n_pad > 128 && p.n_pad % 128 != 0) p.n_tile = 64;
  LSI_REQUIRE(p.n_pad % p.n_tile == 0 && (p.n_tile == 16 || p.n_tile % 32 == 0), "unsupported output channel count %d", d->c_out);
  const uint32_t rb = (h16 == 1) ? 64u : 128u;
  const uint32_t rbw = (h16 == 1 || h16 == 3) ? 64u : 128u;
  const int n_mma = (h16 >= 2) ? 2 * p.n_tile : p.n_tile;
  const uint32_t b_bytes = ((uint32_t)n_mma * rbw + 1023) & ~1023u;
  // x-merge: unit-stride gathers with more than one tap along x, on images wide enough for 16x8 tiles to make sense
  const int nkx_max = (d->mode == 1) ? (d->kw + s - 1) / s : d->kw;
  p.xm = (xmerge_enabled() && (d->mode == 1 || d->stride == 1) && nkx_max >= 2 && nkx_max <= 9 && p.Hp >= 16) ? 1 : 0;
  if (p.xm) {   // at least two ring stages of (halo tile + one filter tile per tap along x) must fit
    const uint32_t hw = xmerge_tight() ? 8 + nkx_max - 1 : 16;
    if (2 * (hw * 16 * rb + (uint32_t)d->kw * b_bytes) > 200u * 1024u) p.xm = 0;
  }
  p.th = p.xm ? 16 : kTileH; p.tw = p.xm ? 8 : kTileW; p.tn = 1;
  if (!p.xm && p.Wp <= 16 && batch_tiles_enabled()) {   // small images: whole images side by side in the 128 rows of a tile
    const int tw = p.Wp <= 8 ? 8 : 16;
    int th = 1;
    while (th < p.Hp) th <<= 1;
    if (tw * th < kTileM) { p.tw = tw; p.th = th; p.tn = kTileM / (tw * th); }
    else if (tw == 8) { p.tw = 8; p.th = 16; }
  }
  p.halo_w = p.xm ? (xmerge_tight() ? p.tw + nkx_max - 1 : 16) : 0;
  p.tiles_x = (p.Wp + p.tw - 1) / p.tw; p.tiles_y = (p.Hp + p.th - 1) / p.th;
  p.Ca = c_in_a; p.Cb = d->c_in - c_in_a;
  p.kh = d->kh; p.kw = d->kw; p.stride = d->stride; p.pad_t = d->pad_top; p.pad_l = d->pad_left; p.mode = d->mode;
  p.epilogue = d->epilogue; p.accumulate = d->accumulate;

  // weights -> K-major [tap][n_pad][cin]: into the workspace, or -- when the caller vouched for a weight version -- into the memo
  float* const wk_ws = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
  float* wk = wk_ws;
  const int taps = d->kh * d->kw;
  bool prep_needed = true;
  if (wver) {
    const int sig[12] = {1, h16, taps, d->c_in, d->c_out, p.n_pad, p.n_tile, d->w_tap_stride, d->w_ci_stride, d->w_co_stride, 0, 0};
    bool hit = false;
    void* buf = prep_cache_get(w, wver, sig, (size_t)taps * p.n_pad * d->c_in * sizeof(float) * ((h16 == 2) ? 2 : 1), &hit);
    if (buf) { wk = static_cast<float*>(buf); prep_needed = !hit; }
  }
  if (prep_needed) {
    const long long total = (long long)taps * p.n_pad * d->c_in;
    long long g = (total + 255) / 256; if (g > 148 * 8) g = 148 * 8;
    if (h16 == 3)
      prep_weights_split_l_kernel<<<(unsigned)(g * 2 > 148 * 8 ? 148 * 8 : g * 2), 256, 0, st>>>(
          w, reinterpret_cast<__half*>(wk), taps, d->c_in, d->c_out, p.n_pad, p.n_tile, d->w_tap_stride, d->w_ci_stride, d->w_co_stride);
    else if (h16 == 2)
      prep_weights_split_kernel<<<(unsigned)(g * 4 > 148 * 8 ? 148 * 8 : g * 4), 256, 0, st>>>(
          w, reinterpret_cast<__half*>(wk), taps, d->c_in, d->c_out, p.n_pad, p.n_tile, d->w_tap_stride, d->w_ci_stride, d->w_co_stride);
    else if (h16)
      prep_weights_f16_kernel<<<(unsigned)g, 256, 0, st>>>(w, reinterpret_cast<__half*>(wk), taps, d->c_in, d->c_out, p.n_pad,
                                                           d->w_tap_stride, d->w_ci_stride, d->w_co_stride);
    else
      prep_weights_kernel<<<(unsigned)g, 256, 0, st>>>(w, wk, taps, d->c_in, d->c_out, p.n_pad, d->w_tap_stride, d->w_ci_stride,
                                                       d->w_co_stride);
    LSI_LAUNCH_CHECK();
  }

  // tensor maps
  const cuuint64_t eb = h16 ? 2 : 4;
  const cuuint64_t cm = (h16 >= 2) ? 2 : 1;    // split: 2 fp16 elements per channel, 64-element (128-byte) box rows = [hi | lo]
  const cuuint64_t cmw = (h16 == 2) ? 2 : 1;   // filter map: variant S rows are [Whi|0] / [Wlo|Whi] (64 elements per chunk), variant L 32
  const CUtensorMapDataType dt = h16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32;
  const CUtensorMapSwizzle sw = (h16 == 1) ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
  auto make_act_map = [&](CUtensorMap* m, const void* base, int channels, int cs) -> int {
    const int es = (d->mode == 0) ? d->stride : 1;      // element (traversal) stride of the gather
    cuuint64_t dims[4] = {(cuuint64_t)channels * cm, (cuuint64_t)d->w_in, (cuuint64_t)d->h_in, (cuuint64_t)d->batch};
    cuuint64_t strides[3] = {(cuuint64_t)cs * cm * eb, (cuuint64_t)d->w_in * cs * cm * eb, (cuuint64_t)d->h_in * d->w_in * cs * cm * eb};
    cuuint32_t box[4] = {(cuuint32_t)(kKC * cm), (cuuint32_t)((p.tw - 1) * es + 1), (cuuint32_t)((p.th - 1) * es + 1), (cuuint32_t)p.tn};
    if (p.xm) { box[1] = (cuuint32_t)p.halo_w; box[2] = (cuuint32_t)p.th; }
    cuuint32_t estr[4] = {1, (cuuint32_t)es, (cuuint32_t)es, 1};
    CUresult r = encode(m, dt, 4, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(activations) failed: %d", (int)r); return LSI_B200_ECUDA; }
    return LSI_B200_OK;
  };
  CUtensorMap map_a, map_b, map_w;
  if (int rc = make_act_map(&map_a, in_a, p.Ca, d->in_c_stride)) return rc;
  if (p.Cb > 0) { if (int rc = make_act_map(&map_b, in_b, p.Cb, in_b_c_stride)) return rc; }
  else map_b = map_a;
  {
    cuuint64_t dims[2] = {(cuuint64_t)d->c_in * cmw, (cuuint64_t)taps * p.n_pad * cm};
    cuuint64_t strides[1] = {(cuuint64_t)d->c_in * cmw * eb};
    cuuint32_t box[2] = {(cuuint32_t)(kKC * cmw), (cuuint32_t)n_mma};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&map_w, dt, 2, wk, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        (h16 == 3) ? CU_TENSOR_MAP_SWIZZLE_64B : sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(weights) failed: %d", (int)r); return LSI_B200_ECUDA; }
  }
  const uint32_t stage_bytes = p.xm ? (uint32_t)p.halo_w * p.th * rb + (uint32_t)d->kw * b_bytes : kTileM * rb + b_bytes;
  p.batch = d->batch;
  p.total_tiles = p.tiles_x * p.tiles_y * ((d->batch + p.tn - 1) / p.tn) * (p.n_pad / p.n_tile) * s * s;
  // ring depth: 74 KB of stages (up to 4) lets 2-3 CTAs share an SM; the small-spatial layers of the trunk (fewer tiles than
  // SMs, K = 9 x 512 .. 1024) are one long TMA -> MMA latency chain per CTA, so they get the whole SM: up to 8 stages
  const bool latency_bound = p.total_tiles <= num_sms();
  int stages = (int)(((latency_bound ? 200u : 74u) * 1024u) / stage_bytes);
  if (stages > (latency_bound ? kMaxStages : 4)) stages = latency_bound ? kMaxStages : 4;
  if (stages < 2) stages = 2;
  p.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + 256 + 1024 + (bn_stats ? 4 * 32 * 33 * sizeof(float) : 0);
  static size_t smem_set[2] = {0, 0};
  const int ks = h16 >= 2 ? 1 : 0;
  if (smem > smem_set[ks]) {
    if (ks) LSI_CUDA(cudaFuncSetAttribute(conv_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else LSI_CUDA(cudaFuncSetAttribute(conv_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set[ks] = smem;
  }
  int ctas_per_sm = (int)((220u * 1024u) / (smem + 1024));
  const int tmem_per_cta = 2 * (n_mma <= 32 ? 32 : n_mma <= 64 ? 64 : n_mma <= 128 ? 128 : 256);
  if (ctas_per_sm > 512 / tmem_per_cta) ctas_per_sm = 512 / tmem_per_cta;
  if (ctas_per_sm < 1) ctas_per_sm = 1;
  if (ctas_per_sm > 4) ctas_per_sm = 4;
  int n_ctas = num_sms() * ctas_per_sm;
  if (n_ctas > p.total_tiles) n_ctas = p.total_tiles;
  dim3 grid((unsigned)n_ctas);
  p.stat_part = nullptr;
  if (bn_stats) {
    p.stat_part = wk_ws + (size_t)taps * p.n_pad * d->c_in * ((h16 == 2) ? 2 : 1);   // (variant L: the fp16 pair rows take the fp32 size)
    p.stat_part = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(p.stat_part) + 255) & ~uintptr_t(255));
    LSI_CUDA(cudaMemsetAsync(p.stat_part, 0, (size_t)n_ctas * 4 * p.n_pad * 2 * sizeof(float), st));
  }
  {
    ScopedTiming tm(kConvTc, st);
    if (ks) conv_tc_kernel<true><<<grid, kThreads, smem, st>>>(map_a, map_b, map_w, p);
    else conv_tc_kernel<false><<<grid, kThreads, smem, st>>>(map_a, map_b, map_w, p);
  }
  LSI_LAUNCH_CHECK();
  if (bn_stats) {
    finalize_stats_f32_kernel<<<(d->c_out + 7) / 8, 256, 0, st>>>(p.stat_part, n_ctas * 4, p.n_pad, d->c_out,
                                                                 (long long)d->batch * d->h_out * d->w_out, bn_eps, bn_stats);
    LSI_LAUNCH_CHECK();
  }
  return LSI_B200_OK;
}

// fp16 activations (in_a / in_b / out point to __half tensors, strides in elements), fp16-rounded weights, fp32 accumulation,
// batch statistics from the fp32 accumulators: the inference-only 'f16' mode of lsi.nnutils.nets.  out_f16 == 0 writes fp32.
extern "C" int lsi_b200_conv2d_tc_h(const lsi_b200_conv_desc* d, const void* in_a, int c_in_a, const void* in_b, int in_b_c_stride,
                                    const float* w, const float* bias, void* out, int out_f16, float* bn_stats, float bn_eps,
                                    void* workspace, size_t workspace_bytes, void* stream) {
  LSI_REQUIRE(!bn_stats || (d && d->epilogue == 0 && d->accumulate == 0), "bn statistics need a plain conv output");
  return conv2d_tc_impl(d, in_a, c_in_a, in_b, in_b_c_stride, w, bias, out, bn_stats, bn_eps, workspace, workspace_bytes, stream, 1,
                        out_f16);
}

// Split-precision mode (csrc/split.cuh, lsi.nnutils.nets.set_conv_mode('split')): in_a / in_b are split fp16-pair tensors
// (chunk-interleaved [hi 32 ch | lo 32 ch], same bytes as fp32; strides in channels, multiples of 32); weights are split on the fly;
// three exact fp16 products per fp32 product accumulate in two fp32 TMEM accumulators.  out_kind 0: fp32 output (any
// epilogue, optional out_scale per channel after the activation); 2: split output (plain convs).  bn_stats as in
// lsi_b200_conv2d_tc_bnstats.
extern "C" int lsi_b200_conv2d_tc_s(const lsi_b200_conv_desc* d, const void* in_a, int c_in_a, const void* in_b, int in_b_c_stride,
                                    const float* w, const float* bias, const float* out_scale, void* out, int out_kind,
                                    float* bn_stats, float bn_eps, void* workspace, size_t workspace_bytes, void* stream) {
  LSI_REQUIRE(out_kind == 0 || out_kind == 2, "out_kind must be 0 (fp32) or 2 (split)");
  LSI_REQUIRE(!bn_stats || (d && d->epilogue == 0 && d->accumulate == 0), "bn statistics need a plain conv output");
  return conv2d_tc_impl(d, in_a, c_in_a, in_b, in_b_c_stride, w, bias, out, bn_stats, bn_eps, workspace, workspace_bytes, stream, 2,
                        out_kind, out_scale);
}
