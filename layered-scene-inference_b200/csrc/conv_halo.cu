// Halo-tile tensor-core convolution for the full-resolution, few-channel layers of the LDI heads (nets.py:87-114,
// 137-159: upcnv1 64->32 4x4/2 up-conv, upcnv1b 32->32 3x3, pred 32->4 3x3 + bias + sigmoid).  These layers are
// HBM-bound by arithmetic intensity (<= 40 flop/B), so the kernel is organised around bytes, not flops:
//
//   weights      the CTA's whole filter bank ([tap][N][Cin] K-major, <= 36 KB) is TMA-loaded ONCE and stays resident in
//                shared memory; an up-conv CTA serves a single output phase (blockIdx % 4), i.e. 4 of the 16 taps.
//   activations  ONE 4-D TMA box per (tile, 32-channel chunk): the (16 + kh - 1) x (8 + kw - 1) pixel halo of a 16 x 8
//                output tile.  Every tap (ky, kx) is an MMA whose A descriptor starts (ky * halo_w + kx) 128-byte rows
//                into that box (8-row groups halo_w rows apart): the tensor core applies the 128B swizzle on absolute
//                shared-memory address bits, exactly as TMA wrote it, so shifted starts need no data movement.  L2->SM
//                traffic per tile drops from 3 x 20 KB activations + 36 KB weights (x-merged kernel) to 23 KB.
//   BN on load   the producing layer's batch-norm + ReLU (slim.batch_norm, nets.py:263-272) is applied to the halo tile
//                IN shared memory by four transform warps (y = max(x * rstd + (beta - mean * rstd), 0); out-of-image
//                pixels stay zero = SAME padding of the normalised tensor), so the normalised activation never
//                exists in HBM: the separate bn_apply pass (read + write of every activation) disappears.
//   epilogue     TMEM -> registers -> per-warp shared-memory staging -> (a) per-channel sum / sum-of-squares for this
//                layer's own batch statistics, (b) 512-byte coalesced global stores (4 pixels x 128 B per instruction).
//   pipeline     persistent CTAs; warp 0 TMA producer, warp 1 single-thread MMA issuer (double-buffered TMEM
//                accumulators), warps 2-5 epilogue, warps 6-9 transform; mbarrier ring over tiles.
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

namespace {

constexpr int kTH = 16, kTW = 8, kTileM = kTH * kTW;   // 128 output pixels = 128 TMEM lanes
constexpr int kKC = 32;                                // fp32 channels per 128-byte row
constexpr int kMaxStages = 8;
constexpr int kThreads = 320;                          // warp 0 TMA, 1 MMA, 2-5 epilogue, 6-9 transform
constexpr int kThreadsWide = 448;                      // 1-CTA/SM variant: 8 transform warps (6-13); its split flavour uses the kThreadsAll layout
constexpr int kThreadsAll = 576;                       // all-phase variant: 8 epilogue warps (2-9, two per TMEM lane group), 8 transform warps (10-17)
constexpr int kStgPitch = 36;                          // floats per staged pixel (144 B: conflict-free 128-bit rows)
constexpr int kMaxCin = 128;

struct HaloParams {
  float* out; const float* bias; const float* out_scale;   // out_scale: per-channel factor after the activation (small-output path), or NULL
  const float* in_stats; const float* in_beta;   // producer's (mean, rstd)[Cin] and beta[Cin]; NULL = input is final
  float* stat_part;                              // [gridDim.x * 4][n_tile][2] channel sums of the output, or NULL
  int Hin, Win, Ho, Wo, Co, out_cs, Hp, Wp;
  int tiles_x, per_img, spatial_tiles;
  int chunks, kh, kw, stride, pad_t, pad_l, mode;
  int chunks_a;              // 32-channel chunks that come from source A (pending batch norm); the rest from source B
                             // (already normalised: tf.concat of the U-Net skip, nets.py:108-109), read through map_b
  int n_tile, epilogue, stages;
  int f16;                   // 1: transform warps also convert the normalised tile to fp16 in place; MMAs run kind::f16
  int in_f16;                // 1: the input tensor is stored as fp16 (RAW output of a producer run with out_f16): 64-byte rows,
                             //    64B swizzle, normalised in place by the transform warps (implies f16)
  int out_f16;               // 1: the RAW output is stored as fp16 (its only consumer normalises on load and feeds fp16 MMAs anyway)
  int split;                 // (= kSplit) split fp16-pair activations (128-byte rows [hi 32 ch | lo 32 ch]) and weights (2N rows of 64 bytes: Whi ; Wlo):
                             //    a_hi x both with N' = 2N, then a_lo x Whi with N' = N; two accumulators (columns [0,N) and [N,2N)); out_f16 == 2 stores split pairs
  int all_phase;             // 1 (kAll): one CTA computes all stride^2 output phases of an up-conv tile from ONE halo load/transform
  int oy_min, ox_min;        // kAll: halo origin relative to the tile origin (minimum over the phases)
  int halo_w, halo_h;
  uint32_t halo_bytes, b_tap_bytes, w_bytes, div_halo_w;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// K-major 128B-swizzled operand: 8-row groups `sbo` bytes apart, start address possibly shifted by whole 128-byte rows
__device__ __forceinline__ uint64_t umma_desc_hi(uint32_t sbo) {
  return ((uint64_t)1 << 16) | ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint64_t umma_desc(uint64_t hi, uint32_t saddr) { return hi | (uint64_t)((saddr >> 4) & 0x3FFF); }
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// K-major operand tile with the 64-byte swizzle (fp16 weights: 32 channels = 64-byte rows), 8-row groups 512 bytes apart
__device__ __forceinline__ uint64_t umma_desc_hi_sw64() {
  return ((uint64_t)1 << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// true for exactly one lane of a converged warp; unlike `lane == 0` it tells ptxas that a single thread runs the
// guarded region, so tcgen05/TMA operands go to uniform registers without a per-lane waterfall loop
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

// ring position that advances without integer division
struct Ring {
  int st; uint32_t ph; int n;
  __device__ __forceinline__ explicit Ring(int stages) : st(0), ph(0), n(stages) {}
  __device__ __forceinline__ void next() { if (++st == n) { st = 0; ph ^= 1; } }
};

// (image, tile row, tile column) of the persistent tile sequence t0, t0 + step, ... advanced without integer division
struct TileIter {
  int n, ty, tx;             // current tile
  int dn, dty, dtx;          // step decomposed in (images, tile rows, tile columns)
  int tiles_x, tiles_y;
  __device__ __forceinline__ TileIter(int t0, int step, int tiles_x_, int per_img) : tiles_x(tiles_x_), tiles_y(per_img / tiles_x_) {
    n = t0 / per_img; int r = t0 - n * per_img; ty = r / tiles_x; tx = r - ty * tiles_x;
    dn = step / per_img; r = step - dn * per_img; dty = r / tiles_x; dtx = r - dty * tiles_x;
  }
  __device__ __forceinline__ void next() {
    tx += dtx; ty += dty; n += dn;
    if (tx >= tiles_x) { tx -= tiles_x; ++ty; }
    if (ty >= tiles_y) { ty -= tiles_y; ++n; }
  }
};

__device__ __forceinline__ uint32_t pack_half2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// One halo tile (halo_w * halo_h pixel rows of 128 bytes = 32 fp32 channels, 128B-swizzled as TMA wrote it), 128 threads:
// y = max(x * a + b, 0) for pixels inside the image, 0 outside (SAME padding of the NORMALISED tensor).
// thread -> (16-byte column j, pixels p0, p0 + kNT/8, ...): the pixel step is a multiple of 8, so the swizzle phase -- hence the four
// channels this thread touches -- is fixed and scale/shift live in registers.
// kF16: additionally convert to fp16 IN PLACE: row p keeps its 128-byte pitch, its 32 channels become four 16-byte chunks
// at swizzled positions (c ^ (p & 7)), i.e. the K-major SWIZZLE_128B layout of a 64-channel-wide fp16 tile of which the MMA
// reads K = 0..31.  The two lanes that hold fp32 chunks 2c and 2c+1 of a row are neighbours (j and j ^ 1): one shuffle
// pairs them, the even one stores.  All 8 lanes of a row belong to one warp instruction, so every read of a row precedes
// every write to it.
template <bool kF16, bool kInterior, int kNT>
__device__ __forceinline__ void transform_tile(float4* tile, const float* bn_a, const float* bn_b, int tid, const HaloParams& p,
                                               int ys0, int xs0) {
  const int j = tid & 7, p0 = tid >> 3, s7 = p0 & 7;
  const int jl = j ^ s7;                                     // logical 16-byte fp32 chunk = channels 4*jl .. 4*jl+3
  const float4 a = *reinterpret_cast<const float4*>(bn_a + (jl << 2));
  const float4 b = *reinterpret_cast<const float4*>(bn_b + (jl << 2));
  const int npx = p.halo_w * p.halo_h;
  constexpr int kRows = kNT / 8;                             // pixel rows covered per pass (16 with 128 threads, 32 with 256)
  const int iters = (npx + kRows - 1) / kRows;               // warp-uniform trip count (shuffles below)
  float4* q = tile + tid;                                    // item k lives at q + kNT * k
  const int wr = ((jl >> 1) ^ s7) - j;                       // fp16 mode: float4 offset of this pair's output chunk from q
  for (int k0 = 0; k0 < iters; k0 += 4) {
    float4 v[4]; bool ok[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int pxl = p0 + kRows * (k0 + u);
      ok[u] = pxl < npx;
      v[u] = ok[u] ? q[kNT * (k0 + u)] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      bool inb = ok[u];
      if (!kInterior) {
        const int pxl = p0 + kRows * (k0 + u);
        const int hy = (int)(((uint32_t)pxl * p.div_halo_w) >> 16), hx = pxl - hy * p.halo_w;
        inb = inb && (unsigned)(ys0 + hy) < (unsigned)p.Hin && (unsigned)(xs0 + hx) < (unsigned)p.Win;
      }
      float4 y;
      y.x = inb ? fmaxf(fmaf(v[u].x, a.x, b.x), 0.f) : 0.f; y.y = inb ? fmaxf(fmaf(v[u].y, a.y, b.y), 0.f) : 0.f;
      y.z = inb ? fmaxf(fmaf(v[u].z, a.z, b.z), 0.f) : 0.f; y.w = inb ? fmaxf(fmaf(v[u].w, a.w, b.w), 0.f) : 0.f;
      if (kF16) {
        const uint32_t h01 = pack_half2(y.x, y.y), h23 = pack_half2(y.z, y.w);
        const uint32_t o01 = __shfl_xor_sync(0xffffffffu, h01, 1), o23 = __shfl_xor_sync(0xffffffffu, h23, 1);
        if (ok[u] && !(jl & 1))
          *reinterpret_cast<uint4*>(q + kNT * (k0 + u) + wr) = make_uint4(h01, h23, o01, o23);
      } else {
        if (ok[u]) q[kNT * (k0 + u)] = y;
      }
    }
  }
}

// fp16 input tile: halo_w * halo_h pixel rows of 64 bytes (32 fp16 channels), 64B-swizzled as TMA wrote it: 16-byte chunk c
// of row p sits at chunk position c ^ ((p >> 1) & 3).  thread -> (chunk position j, pixels p0, p0 + kNT/4, ...): the pixel
// step is a multiple of 8, so the logical chunk -- hence the eight channels this thread touches -- is fixed and their
// scale/shift live in registers.  Normalised in place (no layout change: the MMA reads the same SWIZZLE_64B K-major tile).
template <bool kInterior, int kNT>
__device__ __forceinline__ void transform_tile_h(uint4* tile, const float* bn_a, const float* bn_b, int tid, const HaloParams& p,
                                                 int ys0, int xs0) {
  const int j = tid & 3, p0 = tid >> 2;
  const int jl = j ^ ((p0 >> 1) & 3);                        // logical chunk = channels 8*jl .. 8*jl+7
  float a[8], b[8];
#pragma unroll
  for (int k = 0; k < 8; k += 4) {
    const float4 a4 = *reinterpret_cast<const float4*>(bn_a + (jl << 3) + k);
    const float4 b4 = *reinterpret_cast<const float4*>(bn_b + (jl << 3) + k);
    a[k] = a4.x; a[k + 1] = a4.y; a[k + 2] = a4.z; a[k + 3] = a4.w;
    b[k] = b4.x; b[k + 1] = b4.y; b[k + 2] = b4.z; b[k + 3] = b4.w;
  }
  const int npx = p.halo_w * p.halo_h;
  constexpr int kRows = kNT / 4;                             // pixel rows covered per pass
  uint4* q = tile + tid;                                     // item k lives at q + kNT * k
  for (int k0 = 0; p0 + kRows * k0 < npx; k0 += 2) {
    uint4 v[2]; bool ok[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      ok[u] = p0 + kRows * (k0 + u) < npx;
      v[u] = ok[u] ? q[kNT * (k0 + u)] : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      bool inb = ok[u];
      if (!kInterior) {
        const int pxl = p0 + kRows * (k0 + u);
        const int hy = (int)(((uint32_t)pxl * p.div_halo_w) >> 16), hx = pxl - hy * p.halo_w;
        inb = inb && (unsigned)(ys0 + hy) < (unsigned)p.Hin && (unsigned)(xs0 + hx) < (unsigned)p.Win;
      }
      const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
      uint32_t o[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[k]));
        const float y0 = inb ? fmaxf(fmaf(f.x, a[2 * k], b[2 * k]), 0.f) : 0.f;
        const float y1 = inb ? fmaxf(fmaf(f.y, a[2 * k + 1], b[2 * k + 1]), 0.f) : 0.f;
        o[k] = pack_half2(y0, y1);
      }
      if (ok[u]) q[kNT * (k0 + u)] = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

// split input tile (split.cuh): halo_w * halo_h pixel rows of 128 bytes = [hi of 32 channels | lo of 32 channels], 128B-swizzled as
// TMA wrote it: 16-byte chunk c (0..3 hi, 4..7 lo) of row p sits at position c ^ (p & 7).  thread -> (logical chunk c = 8 channels,
// rows p0, p0 + kNT/4, ...): it owns the hi chunk at position c ^ (p & 7) and the matching lo chunk 4 positions (xor) away; the row
// step is a multiple of 8, so positions and scale/shift are loop invariants.  The eight lanes of a quarter-warp cover rows r and r + 4
// (opposite 64-byte halves): conflict-free 128-bit accesses.  y = max(v * a + b, 0) re-split in place.
template <bool kInterior, int kNT>
__device__ __forceinline__ void transform_tile_s(uint4* tile, const float* bn_a, const float* bn_b, int tid, const HaloParams& p,
                                                 int ys0, int xs0) {
  const int lane = tid & 31, c = lane & 3;
  const int p0 = (tid >> 5) * 8 + ((lane >> 2) & 1) * 4 + (lane >> 3);
  const int pos = c ^ (p0 & 7);
  float a[8], b[8];
#pragma unroll
  for (int k = 0; k < 8; k += 4) {
    const float4 a4 = *reinterpret_cast<const float4*>(bn_a + (c << 3) + k);
    const float4 b4 = *reinterpret_cast<const float4*>(bn_b + (c << 3) + k);
    a[k] = a4.x; a[k + 1] = a4.y; a[k + 2] = a4.z; a[k + 3] = a4.w;
    b[k] = b4.x; b[k + 1] = b4.y; b[k + 2] = b4.z; b[k + 3] = b4.w;
  }
  const int npx = p.halo_w * p.halo_h;
  constexpr int kRows = kNT / 4;                             // pixel rows covered per pass (a multiple of 8)
  uint4* qh = tile + p0 * 8 + pos;                           // row p0 + kRows * k lives kRows * 8 uint4 further per k
  uint4* ql = tile + p0 * 8 + (pos ^ 4);
  for (int k0 = 0; p0 + kRows * k0 < npx; k0 += 2) {
    uint4 vh[2], vl[2]; bool ok[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      ok[u] = p0 + kRows * (k0 + u) < npx;
      vh[u] = ok[u] ? qh[kRows * 8 * (k0 + u)] : make_uint4(0u, 0u, 0u, 0u);
      vl[u] = ok[u] ? ql[kRows * 8 * (k0 + u)] : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      bool inb = ok[u];
      if (!kInterior) {
        const int pxl = p0 + kRows * (k0 + u);
        const int hy = (int)(((uint32_t)pxl * p.div_halo_w) >> 16), hx = pxl - hy * p.halo_w;
        inb = inb && (unsigned)(ys0 + hy) < (unsigned)p.Hin && (unsigned)(xs0 + hx) < (unsigned)p.Win;
      }
      const uint32_t hw[4] = {vh[u].x, vh[u].y, vh[u].z, vh[u].w}, lw[4] = {vl[u].x, vl[u].y, vl[u].z, vl[u].w};
      uint32_t oh[4], ol[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = split_unpack2(hw[k], lw[k]);
        const float y0 = inb ? fmaxf(fmaf(f.x, a[2 * k], b[2 * k]), 0.f) : 0.f;
        const float y1 = inb ? fmaxf(fmaf(f.y, a[2 * k + 1], b[2 * k + 1]), 0.f) : 0.f;
        split_pack2(y0, y1, oh[k], ol[k]);
      }
      if (ok[u]) {
        qh[kRows * 8 * (k0 + u)] = make_uint4(oh[0], oh[1], oh[2], oh[3]);
        ql[kRows * 8 * (k0 + u)] = make_uint4(ol[0], ol[1], ol[2], ol[3]);
      }
    }
  }
}

// tap list and halo offset of output phase (py, px) of a stride-s gather in mode 1 (see the phase setup in the kernel)
struct PhaseGeom { int ky0, kx0, nky, nkx, oy_off, ox_off; };
__device__ __forceinline__ PhaseGeom phase_geom(const HaloParams& p, int s, int py, int px) {
  PhaseGeom g;
  g.ky0 = (py + p.pad_t) % s; g.kx0 = (px + p.pad_l) % s;
  g.nky = (p.kh - g.ky0 + s - 1) / s; g.nkx = (p.kw - g.kx0 + s - 1) / s;
  g.oy_off = (py + p.pad_t - (g.ky0 + (g.nky - 1) * s)) / s;
  g.ox_off = (px + p.pad_l - (g.kx0 + (g.nkx - 1) * s)) / s;
  return g;
}

// kWide: n_tile == 64 (two 32-column passes per accumulator, two sets of statistics registers)
// kAll:  n_tile == 32 up-conv whose stride^2 output phases are all computed by the same CTA: the halo tile is loaded and
//        normalised once instead of once per phase (that redundancy was ~45 % of the shared-memory traffic of upcnv1);
//        the whole 16-tap filter bank is resident, one 32-column accumulator per phase.  Like kWide it runs one CTA per SM
//        with 8 transform warps.
// kSplit: split fp16-pair activations / weights (split.cuh, variant L: 64-byte filter rows Whi ; Wlo), two accumulators; its wide
//        variant runs 8 epilogue warps (two sets taking alternate tiles = alternate TMEM buffers) because the split epilogue
//        (second tcgen05.ld, combine, hi/lo packing) is twice as long.  Compile-time so that the other paths keep their code.
template <bool kWide, bool kAll, bool kSplit>
__global__ void __launch_bounds__((kAll || (kWide && kSplit)) ? kThreadsAll : (kWide ? kThreadsWide : kThreads), (kWide || kAll) ? 1 : 2)
conv_halo_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 const __grid_constant__ CUtensorMap map_w, const HaloParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS/STS)
  // carve: [resident weights][stages x chunks x halo tile][barriers 256 B][bn scale/shift 2 x 512 B][staging 4 x 32 x 36 floats]
  const uint32_t stage_bytes = p.halo_bytes;     // one stage = one 32-channel chunk of one tile's halo
  uint8_t* s_w = smem;
  uint8_t* s_a = smem + p.w_bytes;
  uint8_t* s_ctl = s_a + (size_t)p.stages * stage_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(s_ctl);
  uint64_t* xready = full + kMaxStages;
  uint64_t* empty = xready + kMaxStages;
  uint64_t* tmem_full = empty + kMaxStages;     // [2]
  uint64_t* tmem_empty = tmem_full + 2;         // [2]
  uint64_t* wfull = tmem_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull + 1);
  float* bn_a = reinterpret_cast<float*>(s_ctl + 256);
  float* bn_b = bn_a + kMaxCin;
  float* staging = bn_b + kMaxCin;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool bn_in = p.in_stats != nullptr;

  // phase of this CTA (up-conv / data-gradient mode: one of stride^2 output phases; plain conv: the only one)
  constexpr bool kBig = kWide || kAll;
  constexpr bool kEpi8 = kAll || (kWide && kSplit);   // 8 epilogue warps (2-9), transform warps from warp 10
  const int s = (p.mode == 1) ? p.stride : 1;
  const int G = kAll ? 1 : s * s;
  const int n_ph = kAll ? s * s : 1;           // phases computed per tile by this CTA
  const int phase = (int)(blockIdx.x % G);
  const int cta_in_group = (int)(blockIdx.x / G), ctas_per_group = (int)(gridDim.x / G);
  const int py = phase / s, px = phase % s;
  const int ky0 = (p.mode == 1) ? ((py + p.pad_t) % s) : 0, kx0 = (p.mode == 1) ? ((px + p.pad_l) % s) : 0;
  const int nky = (p.kh - ky0 + s - 1) / s, nkx = (p.kw - kx0 + s - 1) / s;
  // halo origin relative to the tile origin, in source pixels (exact divisions: the tap list matches the phase)
  const int oy_off = kAll ? p.oy_min : ((p.mode == 0) ? -p.pad_t : (py + p.pad_t - (ky0 + (nky - 1) * s)) / s);
  const int ox_off = kAll ? p.ox_min : ((p.mode == 0) ? -p.pad_l : (px + p.pad_l - (kx0 + (nkx - 1) * s)) / s);

  const int n_mma = kSplit ? 2 * p.n_tile : p.n_tile;    // UMMA N (split: D0 in columns [0, n_tile), D1 in [n_tile, 2 n_tile))
  const int cm = kSplit ? 2 : 1;                          // fp16 elements per channel in the activation tensor maps of the split layout
  const uint32_t rbw = (p.f16 || kSplit) ? 64u : 128u;    // bytes of one filter row
  const uint32_t acc_cols = kAll ? 32u * (uint32_t)(s * s) : (n_mma <= 32 ? 32u : (n_mma <= 64 ? 64u : 128u));
  const uint32_t tmem_cols = acc_cols * 2;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    for (int i = 0; i < kMaxStages; ++i) { mbar_init(&full[i], 1); mbar_init(&xready[i], kBig ? 8 : 4); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], kAll ? 256 : 128); }
    mbar_init(wfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (bn_in) {   // y = max(x * a + b, 0) with a = rstd, b = beta - mean * rstd
    for (int c = threadIdx.x; c < p.chunks * kKC; c += blockDim.x) {
      if (c < p.chunks_a * kKC) {
        const float mean = __ldg(p.in_stats + 2 * c), rstd = __ldg(p.in_stats + 2 * c + 1);
        bn_a[c] = rstd; bn_b[c] = fmaf(-mean, rstd, __ldg(p.in_beta + c));
      } else {   // source B is already normalised and >= 0: identity (the ReLU is a no-op)
        bn_a[c] = 1.f; bn_b[c] = 0.f;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      // ---------------- TMA producer ----------------
      if (kAll) {   // the whole filter bank, indexed by (ky * kw + kx)
        mbar_expect_tx(wfull, (uint32_t)(p.kh * p.kw * p.chunks) * (uint32_t)p.n_tile * (p.f16 ? 64u : 128u));
        for (int t = 0; t < p.kh * p.kw; ++t)
          for (int ch = 0; ch < p.chunks; ++ch)
            tma_load_2d(smem_u32(s_w) + (uint32_t)(t * p.chunks + ch) * p.b_tap_bytes, &map_w, wfull, ch * kKC, t * p.n_tile);
      } else {
        mbar_expect_tx(wfull, (uint32_t)(nky * nkx * p.chunks) * (uint32_t)n_mma * rbw);
        for (int i = 0; i < nky; ++i)
          for (int j = 0; j < nkx; ++j)
            for (int ch = 0; ch < p.chunks; ++ch)
              tma_load_2d(smem_u32(s_w) + (uint32_t)((i * nkx + j) * p.chunks + ch) * p.b_tap_bytes, &map_w, wfull, ch * kKC,
                          ((ky0 + i * s) * p.kw + (kx0 + j * s)) * n_mma);
      }
      Ring r(p.stages);
      const uint32_t tx = (uint32_t)(p.halo_w * p.halo_h) * (p.in_f16 ? 64u : 128u);
      TileIter ti(cta_in_group, ctas_per_group, p.tiles_x, p.per_img);
      for (int t = cta_in_group; t < p.spatial_tiles; t += ctas_per_group, ti.next()) {
        const int n_img = ti.n, y0 = ti.ty * kTH, x0 = ti.tx * kTW;
        for (int ch = 0; ch < p.chunks; ++ch, r.next()) {
          mbar_wait(&empty[r.st], r.ph ^ 1);
          mbar_expect_tx(&full[r.st], tx);
          if (ch < p.chunks_a) tma_load_4d(smem_u32(s_a) + (uint32_t)r.st * stage_bytes, &map_a, &full[r.st], ch * kKC * cm, x0 + ox_off, y0 + oy_off, n_img);
          else tma_load_4d(smem_u32(s_a) + (uint32_t)r.st * stage_bytes, &map_b, &full[r.st], (ch - p.chunks_a) * kKC * cm, x0 + ox_off, y0 + oy_off, n_img);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ---------------- MMA issuer ----------------
      // instruction descriptor (cute::UMMA::InstrDescriptor): D=F32, A=B=TF32, K-major both, N>>3, M>>4
      // (kind::f16: A/B format F16 = 0, two K = 16 steps per 32-channel chunk)
      const uint32_t idesc = (p.f16 || kSplit) ? ((1u << 4) | ((uint32_t)(n_mma >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24))
                                   : ((1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.n_tile >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24));
      // fp16-stored input: 64-byte rows with the 64B swizzle, 8-row groups halo_w rows apart; same shifted-start trick
      const uint32_t row_b = p.in_f16 ? 64u : 128u;
      const uint64_t hi_a = p.in_f16 ? (((uint64_t)1 << 16) | ((uint64_t)(((uint32_t)p.halo_w * 64u) >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)4 << 61))
                                     : umma_desc_hi((uint32_t)p.halo_w * 128u);
      const uint64_t hi_b = (p.f16 || kSplit) ? umma_desc_hi_sw64() : umma_desc_hi(1024u);
      const uint32_t idesc_n = (1u << 4) | ((uint32_t)(p.n_tile >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);   // variant L: a_lo x Whi, N' = N
      mbar_wait(wfull, 0);
      const uint64_t b_desc0 = umma_desc(hi_b, smem_u32(s_w));
      // kAll (4x4 stride-2 up-conv: 4 phases x 2x2 taps): descriptor offsets of every (phase, tap) once, in registers -- the
      // issuing thread is a single dependent instruction stream, per-tile integer divisions would dominate it
      uint32_t a_off[16], b_off[16];
      if (kAll) {
#pragma unroll
        for (int ph = 0; ph < 4; ++ph) {
          const PhaseGeom g = phase_geom(p, s, ph >> 1, ph & 1);
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const int i = t >> 1, j = t & 1;
            const int dy = g.oy_off - p.oy_min + g.nky - 1 - i, dx = g.ox_off - p.ox_min + g.nkx - 1 - j;
            a_off[ph * 4 + t] = ((uint32_t)(dy * p.halo_w + dx) * row_b) >> 4;
            b_off[ph * 4 + t] = ((uint32_t)(((g.ky0 + i * s) * p.kw + (g.kx0 + j * s)) * p.chunks) * p.b_tap_bytes) >> 4;
          }
        }
      }
      // single-phase path: the same, for the <= 9 taps of this CTA's phase
      uint32_t a_tap[9], b_tap[9];
      const int ntaps = nky * nkx;
      if (!kAll) {
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const int i = t / nkx, j = t - i * nkx;
          const int dy = (p.mode == 0) ? i : nky - 1 - i, dx = (p.mode == 0) ? j : nkx - 1 - j;
          a_tap[t] = ((uint32_t)(dy * p.halo_w + dx) * row_b) >> 4;
          b_tap[t] = ((uint32_t)(t * p.chunks) * p.b_tap_bytes) >> 4;
        }
      }
      Ring r(p.stages);
      int tcount = 0;
      for (int t = cta_in_group; t < p.spatial_tiles; t += ctas_per_group, ++tcount) {
        const int buf = tcount & 1;
        mbar_wait(&tmem_empty[buf], ((tcount >> 1) & 1) ^ 1);     // epilogue has drained this accumulator
        const uint32_t tmem_d = tmem_base + (uint32_t)buf * acc_cols;
        uint32_t acc = 0;
        for (int ch = 0; ch < p.chunks; ++ch, r.next()) {
          mbar_wait(bn_in ? &xready[r.st] : &full[r.st], r.ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          // descriptors advance by adding byte offsets >> 4 to the start-address field (no carry out of its 14 bits:
          // every operand lives below 256 KB)
          const uint64_t a_desc0 = umma_desc(hi_a, smem_u32(s_a) + (uint32_t)r.st * stage_bytes);
          if (kAll) {                               // every output phase from the same (normalised) halo tile
            const uint32_t ch_off = ((uint32_t)ch * p.b_tap_bytes) >> 4;
#pragma unroll
            for (int ph = 0; ph < 4; ++ph) {
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const uint64_t ad = a_desc0 + (uint64_t)a_off[ph * 4 + t], bd = b_desc0 + (uint64_t)(b_off[ph * 4 + t] + ch_off);
#pragma unroll
                for (int kk = 0; kk < kKC / 16; ++kk)
                  umma_f16(tmem_d + (uint32_t)ph * 32u, ad + 2 * kk, bd + 2 * kk, idesc, (uint32_t)((ch | t | kk) != 0));
              }
            }
            umma_commit(&empty[r.st]);
            continue;
          }
          {
            const uint32_t ch_off = ((uint32_t)ch * p.b_tap_bytes) >> 4;
#pragma unroll
            for (int tp = 0; tp < 9; ++tp) {
              if (tp < ntaps) {
                const uint64_t ad = a_desc0 + (uint64_t)a_tap[tp], bd = b_desc0 + (uint64_t)(b_tap[tp] + ch_off);
                if (kSplit) {                              // a_hi x (Whi ; Wlo) into columns [0, 2N), a_lo x Whi into [N, 2N)
#pragma unroll
                  for (int kk = 0; kk < 2; ++kk) {
                    umma_f16(tmem_d, ad + 2 * kk, bd + 2 * kk, idesc, acc);
                    acc = 1;
                  }
#pragma unroll
                  for (int kk = 0; kk < 2; ++kk)
                    umma_f16(tmem_d + (uint32_t)p.n_tile, ad + 4 + 2 * kk, bd + 2 * kk, idesc_n, 1u);
                } else if (p.f16) {
#pragma unroll
                  for (int kk = 0; kk < kKC / 16; ++kk) {  // UMMA K = 16 for fp16: 32 bytes along the row
                    umma_f16(tmem_d, ad + 2 * kk, bd + 2 * kk, idesc, acc);
                    acc = 1;
                  }
                } else {
#pragma unroll
                  for (int kk = 0; kk < kKC / 8; ++kk) {   // UMMA K = 8 for TF32: 32 bytes along the swizzled row
                    umma_tf32(tmem_d, ad + 2 * kk, bd + 2 * kk, idesc, acc);
                    acc = 1;
                  }
                }
              }
            }
          }
          umma_commit(&empty[r.st]);               // frees the stage once these MMAs have read it
        }
        umma_commit(&tmem_full[buf]);              // accumulator complete
      }
    }
  } else if (warp < (kEpi8 ? 10 : 6)) {
    // ---------------- epilogue: TMEM -> registers -> staging -> global ----------------
    const int lg = warp & 3;                     // TMEM lane group this warp may access
    const int eset = kEpi8 ? ((warp - 2) >> 2) : 0;   // two warps per lane group: kAll -- each takes every other phase; kWide -- every other tile
    const int row = lg * 32 + lane;              // A-tile row = pixel within the 16 x 8 patch
    const int hy = row >> 3, wx = row & 7;
    float* stg = staging + (size_t)(eset * 4 + lg) * 32 * kStgPitch;
    // batch statistics: lane (q4, c4) meets channels cc + c4 .. c4 + 3 of pixels q4, q4 + 4, ... in the store loop below and
    // sums them there (zero rows for pixels outside the output), so the statistics cost no extra shared-memory reads
    float4 ssum0 = make_float4(0.f, 0.f, 0.f, 0.f), ssum1 = ssum0, ssq0 = ssum0, ssq1 = ssum0;   // channels cc = 0 / cc = 32
    const bool direct4 = (p.Co <= 4 && p.out_cs == 4);   // prediction head: one 16-byte store per pixel
    int tcount = 0;
    TileIter ti(cta_in_group, ctas_per_group, p.tiles_x, p.per_img);
    for (int t = cta_in_group; t < p.spatial_tiles; t += ctas_per_group, ++tcount, ti.next()) {
      const int n_img = ti.n, y0 = ti.ty * kTH, x0 = ti.tx * kTW;
      const int buf = tcount & 1;
      if (kWide && kSplit && buf != eset) continue;        // this tile's accumulator buffer belongs to the other set of epilogue warps
      mbar_wait(&tmem_full[buf], (tcount >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const bool in_range = (y0 + hy) < p.Hp && (x0 + wx) < p.Wp;
      for (int ph = (kAll ? eset : 0); ph < n_ph; ph += (kAll ? 2 : 1)) {   // kAll: one 32-column accumulator per output phase
      const int e_py = kAll ? ph / s : py, e_px = kAll ? ph % s : px;
      for (int cc = 0; cc < p.n_tile; cc += 32) {
        uint32_t r[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)buf * acc_cols + (uint32_t)cc + (kAll ? (uint32_t)ph * 32u : 0u);
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
        if (kSplit) {   // second accumulator (cross terms, scaled by 2^11): columns [n_tile + cc, ...)
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
        if (cc + 32 >= p.n_tile && ph + (kAll ? 2 : 1) >= n_ph) {   // this warp's last read of the accumulator: hand the buffer back
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          mbar_arrive(&tmem_empty[buf]);
        }
        if (direct4) {
          if (in_range) {
            int oy = y0 + hy, ox = x0 + wx;
            if (p.mode == 1) { oy = oy * s + e_py; ox = ox * s + e_px; }
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              v[j] = __uint_as_float(r[j]);
              if (p.epilogue >= 1 && j < p.Co) v[j] += __ldg(p.bias + j);
              if (p.epilogue == 2) v[j] = 1.f / (1.f + expf(-v[j]));
              if (p.out_scale && j < p.Co) v[j] *= __ldg(p.out_scale + j);
            }
            float* dst = p.out + ((size_t)(n_img * p.Ho + oy) * p.Wo + ox) * 4;
            if (p.Co == 4) *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
            else {
#pragma unroll
              for (int j = 0; j < 4; ++j) if (j < p.Co) dst[j] = v[j];
            }
          }
          continue;
        }
        // stage this warp's 32 pixels x 32 channels (zero rows for pixels outside the output)
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 v = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
          if (p.epilogue >= 1) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + cc + j));
            v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w;
          }
          if (!in_range) v = make_float4(0.f, 0.f, 0.f, 0.f);
          *reinterpret_cast<float4*>(stg + lane * kStgPitch + j) = v;
        }
        __syncwarp();
        // coalesced stores: each instruction writes 4 pixels x 128 bytes
        const int q4 = lane >> 3, c4 = (lane & 7) * 4;
        float4 sa = make_float4(0.f, 0.f, 0.f, 0.f), sq = sa;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int pxi = i * 4 + q4;
          int oy = y0 + lg * 4 + (pxi >> 3), ox = x0 + (pxi & 7);
          const float4 v = *reinterpret_cast<const float4*>(stg + pxi * kStgPitch + c4);
          sa.x += v.x; sa.y += v.y; sa.z += v.z; sa.w += v.w;
          sq.x = fmaf(v.x, v.x, sq.x); sq.y = fmaf(v.y, v.y, sq.y); sq.z = fmaf(v.z, v.z, sq.z); sq.w = fmaf(v.w, v.w, sq.w);
          if (oy < p.Hp && ox < p.Wp) {
            if (p.mode == 1) { oy = oy * s + e_py; ox = ox * s + e_px; }
            const size_t e = ((size_t)(n_img * p.Ho + oy) * p.Wo + ox) * p.out_cs + cc + c4;
            if (kSplit && p.out_f16 == 2) {   // split pairs: this chunk's 128 bytes = [hi 32 ch | lo 32 ch] at the fp32 tensor's chunk address
              uint2 h2, l2;
              split_pack2(v.x, v.y, h2.x, l2.x); split_pack2(v.z, v.w, h2.y, l2.y);
              uint8_t* cb = reinterpret_cast<uint8_t*>(p.out + (e - c4)) + c4 * 2;
              *reinterpret_cast<uint2*>(cb) = h2;
              *reinterpret_cast<uint2*>(cb + 64) = l2;
            } else if (p.out_f16) *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(p.out) + e) = make_uint2(pack_half2(v.x, v.y), pack_half2(v.z, v.w));
            else *reinterpret_cast<float4*>(p.out + e) = v;
          }
        }
        if (cc == 0) {
          ssum0.x += sa.x; ssum0.y += sa.y; ssum0.z += sa.z; ssum0.w += sa.w;
          ssq0.x += sq.x; ssq0.y += sq.y; ssq0.z += sq.z; ssq0.w += sq.w;
        } else if (kWide) {
          ssum1.x += sa.x; ssum1.y += sa.y; ssum1.z += sa.z; ssum1.w += sa.w;
          ssq1.x += sq.x; ssq1.y += sq.y; ssq1.z += sq.z; ssq1.w += sq.w;
        }
      }
      }
    }
    if (p.stat_part) {
      float* my_part = p.stat_part + ((size_t)blockIdx.x * (kEpi8 ? 8 : 4) + eset * 4 + lg) * p.n_tile * 2;
#pragma unroll
      for (int i = 0; i < (kWide ? 2 : 1); ++i) {
        const float4 fs = i ? ssum1 : ssum0, fq = i ? ssq1 : ssq0;
        float v[8] = {fs.x, fs.y, fs.z, fs.w, fq.x, fq.y, fq.z, fq.w};
#pragma unroll
        for (int k = 0; k < 8; ++k) {       // lanes that share c4 (q4 = 0..3) hold partial sums of the same channels
          v[k] += __shfl_xor_sync(0xffffffffu, v[k], 8);
          v[k] += __shfl_xor_sync(0xffffffffu, v[k], 16);
        }
        const int ch = 32 * i + (lane & 7) * 4;
        if (lane < 8 && ch < p.n_tile) {
#pragma unroll
          for (int k = 0; k < 4; ++k) { my_part[2 * (ch + k)] = v[k]; my_part[2 * (ch + k) + 1] = v[4 + k]; }
        }
      }
    }
  } else if (bn_in) {
    // ---------------- transform: producer's batch-norm + ReLU applied to the halo tile in shared memory ----------------
    const int tid = threadIdx.x - (kEpi8 ? 320 : 192);
    Ring r(p.stages);
    TileIter ti(cta_in_group, ctas_per_group, p.tiles_x, p.per_img);
    for (int t = cta_in_group; t < p.spatial_tiles; t += ctas_per_group, ti.next()) {
      const int ys0 = ti.ty * kTH + oy_off, xs0 = ti.tx * kTW + ox_off;
      const bool interior = ys0 >= 0 && xs0 >= 0 && ys0 + p.halo_h <= p.Hin && xs0 + p.halo_w <= p.Win;
      for (int ch = 0; ch < p.chunks; ++ch, r.next()) {
        mbar_wait(&full[r.st], r.ph);
        float4* tile = reinterpret_cast<float4*>(s_a + (size_t)r.st * stage_bytes);
        const float* ta = bn_a + ch * kKC; const float* tb = bn_b + ch * kKC;
        constexpr int kNT = kBig ? 256 : 128;
        if (kSplit) {
          if (interior) transform_tile_s<true, kNT>(reinterpret_cast<uint4*>(tile), ta, tb, tid, p, ys0, xs0);
          else transform_tile_s<false, kNT>(reinterpret_cast<uint4*>(tile), ta, tb, tid, p, ys0, xs0);
        } else if (p.in_f16) {
          if (interior) transform_tile_h<true, kNT>(reinterpret_cast<uint4*>(tile), ta, tb, tid, p, ys0, xs0);
          else transform_tile_h<false, kNT>(reinterpret_cast<uint4*>(tile), ta, tb, tid, p, ys0, xs0);
        } else if (p.f16) {
          if (interior) transform_tile<true, true, kNT>(tile, ta, tb, tid, p, ys0, xs0);
          else transform_tile<true, false, kNT>(tile, ta, tb, tid, p, ys0, xs0);
        } else {
          if (interior) transform_tile<false, true, kNT>(tile, ta, tb, tid, p, ys0, xs0);
          else transform_tile<false, false, kNT>(tile, ta, tb, tid, p, ys0, xs0);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
        __syncwarp();
        if (lane == 0) mbar_arrive(&xready[r.st]);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// weights (any strides) -> [tap][n_pad][cin] fp32, zero rows for co >= Cout
__global__ void __launch_bounds__(256) halo_prep_weights_kernel(const float* __restrict__ w, float* __restrict__ wk, int taps, int cin,
                                                                int cout, int n_pad, int w_tap, int w_ci, int w_co) {
  const long long total = (long long)taps * n_pad * cin;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % cin);
    const int co = (int)((i / cin) % n_pad);
    const int tap = (int)(i / ((long long)cin * n_pad));
    wk[i] = (co < cout) ? w[(size_t)tap * w_tap + (size_t)ci * w_ci + (size_t)co * w_co] : 0.f;
  }
}

// same, rounded to fp16 (the operand type of the kind::f16 path)
__global__ void __launch_bounds__(256) halo_prep_weights_f16_kernel(const float* __restrict__ w, __half* __restrict__ wk, int taps, int cin,
                                                                    int cout, int n_pad, int w_tap, int w_ci, int w_co) {
  const long long total = (long long)taps * n_pad * cin;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % cin);
    const int co = (int)((i / cin) % n_pad);
    const int tap = (int)(i / ((long long)cin * n_pad));
    wk[i] = __float2half_rn((co < cout) ? w[(size_t)tap * w_tap + (size_t)ci * w_ci + (size_t)co * w_co] : 0.f);
  }
}

// split mode, variant L: [tap][half][n][cin] fp16; half 0 rows = w_hi, half 1 rows = w_lo
__global__ void __launch_bounds__(256) halo_prep_weights_split_l_kernel(const float* __restrict__ w, __half* __restrict__ wk, int taps, int cin,
                                                                        int cout, int n_tile, int w_tap, int w_ci, int w_co) {
  const long long total = (long long)taps * 2 * n_tile * cin;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % cin);
    const long long row = i / cin;
    const int tap = (int)(row / (2 * n_tile)), r2 = (int)(row % (2 * n_tile));
    const int half = r2 / n_tile, co = r2 % n_tile;
    const float v = (co < cout) ? w[(size_t)tap * w_tap + (size_t)ci * w_ci + (size_t)co * w_co] : 0.f;
    const __half hi = __float2half_rn(fminf(fmaxf(v, -kSplitMax), kSplitMax));
    wk[i] = half == 0 ? hi : __float2half_rn((v - __half2float(hi)) * kSplitScale);
  }
}

__global__ void __launch_bounds__(256) halo_finalize_stats_kernel(const float* __restrict__ partial, int nparts, int n_pad, int C,
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

int halo_num_sms() {
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

EncodeTiledFn halo_get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// shared-memory plan of one launch; returns false when the layer does not fit
struct HaloPlan {
  int all_phase, oy_min, ox_min;
  int n_tile, chunks, nky_max, nkx_max, halo_w, halo_h, stages, ctas_per_sm;
  uint32_t halo_bytes, b_tap_bytes, w_bytes;
  size_t smem;
};

bool halo_plan(const lsi_b200_conv_desc* d, HaloPlan* pl, bool f16 = false, bool in_f16 = false, bool all_phase = false, bool split = false) {
  if (!d) return false;
  pl->all_phase = 0; pl->oy_min = 0; pl->ox_min = 0;
  if (d->c_in % kKC != 0 || d->c_in > kMaxCin || d->c_in < kKC) return false;
  if (d->mode != 0 && d->mode != 1) return false;
  if (d->mode == 0 && d->stride != 1) return false;
  if (d->mode == 1 && (d->stride < 1 || d->stride > 2 || d->h_out % d->stride || d->w_out % d->stride)) return false;
  if (d->in_c_stride % 4 != 0 || d->accumulate != 0 || d->epilogue < 0 || d->epilogue > 2) return false;
  const int s = d->mode == 1 ? d->stride : 1;
  const bool small_out = d->c_out <= 4 && d->out_c_stride == 4;                       // prediction head
  const bool wide_out = (d->c_out == 32 || d->c_out == 64) && d->out_c_stride % 4 == 0 && d->epilogue <= 1;
  if (!small_out && !wide_out) return false;
  pl->n_tile = small_out ? 16 : d->c_out;
  pl->chunks = d->c_in / kKC;
  pl->nky_max = (d->kh + s - 1) / s; pl->nkx_max = (d->kw + s - 1) / s;
  if (pl->nky_max * pl->nkx_max > 9) return false;
  pl->halo_w = kTW + pl->nkx_max - 1; pl->halo_h = kTH + pl->nky_max - 1;
  int w_taps = pl->nky_max * pl->nkx_max;
  if (split && (all_phase || f16 || in_f16 || d->in_c_stride % 32 != 0)) return false;
  if (all_phase) {
    // one CTA computes every output phase: the halo is the union of the phases' windows, the whole filter bank is resident
    if (!(d->mode == 1 && s == 2 && f16 && !small_out && d->c_out == 32)) return false;
    int y_lo = 1 << 30, y_hi = -(1 << 30), x_lo = 1 << 30, x_hi = -(1 << 30);
    for (int ph = 0; ph < s * s; ++ph) {
      const int py = ph / s, px = ph % s;
      const int ky0 = (py + d->pad_top) % s, kx0 = (px + d->pad_left) % s;
      const int nky = (d->kh - ky0 + s - 1) / s, nkx = (d->kw - kx0 + s - 1) / s;
      const int oy = (py + d->pad_top - (ky0 + (nky - 1) * s)) / s, ox = (px + d->pad_left - (kx0 + (nkx - 1) * s)) / s;
      if ((py + d->pad_top - (ky0 + (nky - 1) * s)) % s || (px + d->pad_left - (kx0 + (nkx - 1) * s)) % s) return false;
      y_lo = oy < y_lo ? oy : y_lo; y_hi = oy + nky - 1 > y_hi ? oy + nky - 1 : y_hi;
      x_lo = ox < x_lo ? ox : x_lo; x_hi = ox + nkx - 1 > x_hi ? ox + nkx - 1 : x_hi;
    }
    pl->all_phase = 1; pl->oy_min = y_lo; pl->ox_min = x_lo;
    pl->halo_h = kTH + (y_hi - y_lo); pl->halo_w = kTW + (x_hi - x_lo);
    w_taps = d->kh * d->kw;
  }
  pl->halo_bytes = ((uint32_t)(pl->halo_w * pl->halo_h) * (in_f16 ? 64u : 128u) + 1023u) & ~1023u;
  pl->b_tap_bytes = ((uint32_t)pl->n_tile * ((split || !f16) ? 128u : 64u) + 1023u) & ~1023u;   // fp16 weights: 64-byte rows; split: 2 n_tile rows of 64 (L) / 128 (S) bytes
  pl->w_bytes = (uint32_t)(w_taps * pl->chunks) * pl->b_tap_bytes;
  const size_t fixed4 = 1024 + pl->w_bytes + 256 + 2 * kMaxCin * sizeof(float) + 4 * 32 * kStgPitch * sizeof(float);
  const size_t fixed8 = fixed4 + 4 * 32 * kStgPitch * sizeof(float);      // 8 epilogue warps (all-phase and 1-CTA/SM wide variants)
  size_t fixed = all_phase ? fixed8 : fixed4;
  const size_t stage = (size_t)pl->halo_bytes;   // one 32-channel chunk of one tile's halo
  const size_t budget2 = 112 * 1024, budget1 = 224 * 1024;
  static int force_ctas = -1, max_stages = -1;   // measurement knobs
  if (force_ctas < 0) { const char* e = getenv("LSI_B200_HALO_CTAS"); force_ctas = e ? atoi(e) : 0; }
  if (max_stages < 0) { const char* e = getenv("LSI_B200_HALO_STAGES"); max_stages = e ? atoi(e) : 0; }
  if (force_ctas != 1 && !all_phase && pl->n_tile <= 32 && fixed + 2 * stage <= budget2) {   // the 64-column / all-phase variants are built for 1 CTA/SM
    pl->ctas_per_sm = 2;
    pl->stages = (int)((budget2 - fixed) / stage);
  } else if ((fixed = ((all_phase || split) ? fixed8 : fixed4)) + 2 * stage <= budget1) {
    pl->ctas_per_sm = 1;
    pl->stages = (int)((budget1 - fixed) / stage);
  } else {
    return false;
  }
  if (pl->stages > kMaxStages) pl->stages = kMaxStages;
  if (max_stages >= 2 && pl->stages > max_stages) pl->stages = max_stages;
  pl->smem = fixed + (size_t)pl->stages * stage;
  return true;
}

size_t halo_stat_part_bytes(int n_tile) { return (size_t)148 * 2 * 4 * n_tile * 2 * sizeof(float) * 2; }

}  // namespace
}  // namespace lsi

using namespace lsi;

extern "C" int lsi_b200_conv2d_halo_supported(const lsi_b200_conv_desc* d) {
  HaloPlan pl;
  return halo_plan(d, &pl) ? 1 : 0;
}

// fp16-stored input (normalised on load, fp16 operands): layers whose fp16 filter bank fits although the TF32 one does not
extern "C" int lsi_b200_conv2d_halo_h_supported(const lsi_b200_conv_desc* d) {
  HaloPlan pl;
  return (halo_plan(d, &pl) || halo_plan(d, &pl, true, true)) ? 1 : 0;
}

extern "C" size_t lsi_b200_conv2d_halo_workspace_bytes(const lsi_b200_conv_desc* d) {
  HaloPlan pl;
  if (!halo_plan(d, &pl) && !halo_plan(d, &pl, true, true) && !halo_plan(d, &pl, false, false, false, true)) return 0;
  return 2 * (size_t)d->kh * d->kw * pl.n_tile * (size_t)d->c_in * sizeof(float) + 512 + halo_stat_part_bytes(pl.n_tile);   // x2: split filter tiles
}

static int conv2d_halo_impl(const lsi_b200_conv_desc* d, const void* in, const void* in_b, int c_in_a, int in_b_c_stride, int in_f16,
                           const float* in_bn_stats, const float* in_bn_beta,
                           const float* w, const float* bias, const float* out_scale, void* out, int out_f16, float* out_bn_stats,
                           float bn_eps, void* workspace, size_t workspace_bytes, void* stream, int split = 0) {
  const unsigned long long wver = take_weight_version();   // consumed by this call whatever happens next
  LSI_REQUIRE(d && in && w && out && workspace, "NULL pointer argument");
  HaloPlan pl;
  LSI_REQUIRE(split ? halo_plan(d, &pl, false, false, false, true) : (halo_plan(d, &pl) || (in_f16 && halo_plan(d, &pl, true, true))),
              "shape not supported by the halo-tile tensor-core path");
  LSI_REQUIRE(!split || (!in_f16 && !in_b && (out_f16 == 0 || out_f16 == 2)), "split mode: one split source, fp32 or split output");
  LSI_REQUIRE(split || out_f16 != 2, "split output needs the split mode");
  LSI_REQUIRE(!in_f16 || (in_bn_stats && d->in_c_stride % 8 == 0), "fp16-stored input needs the producer's batch-norm statistics and an 8-channel-aligned pixel stride");
  if (!in_b) c_in_a = d->c_in;
  LSI_REQUIRE(c_in_a >= kKC && c_in_a % kKC == 0 && c_in_a <= d->c_in, "bad source split %d of %d channels", c_in_a, d->c_in);
  LSI_REQUIRE(!in_b || (in_bn_stats && in_b_c_stride >= d->c_in - c_in_a && in_b_c_stride % (in_f16 ? 8 : 4) == 0 && ((uintptr_t)in_b & 15) == 0),
              "second source needs pending statistics for the first one and an aligned pixel stride");
  LSI_REQUIRE(!out_f16 || (pl.n_tile == d->c_out && d->epilogue == 0), "fp16-stored / split output is for plain 32/64-channel conv outputs");
  LSI_REQUIRE(out_f16 != 2 || d->out_c_stride % 32 == 0, "split output needs a 32-channel-aligned pixel stride");
  LSI_REQUIRE((in_bn_stats == nullptr) == (in_bn_beta == nullptr), "in_bn_stats and in_bn_beta go together");
  // fp16 operands (same 10-bit mantissa as TF32, fp32 accumulation) whenever the transform warps rewrite the tile anyway:
  // normalised post-ReLU activations are O(1), far inside fp16's range; halves the operand bytes the MMAs pull from
  // shared memory, the resource these layers are bound by
  static int f16_on = -1;
  if (f16_on < 0) { const char* e = getenv("LSI_B200_HALO_F16"); f16_on = (e && atoi(e) == 0) ? 0 : 1; }
  const bool f16 = !split && in_bn_stats != nullptr && (f16_on == 1 || in_f16);
  if (f16) { LSI_REQUIRE(halo_plan(d, &pl, true, in_f16 != 0), "halo plan (fp16) failed"); }
  static int all_on = -1;
  if (all_on < 0) { const char* e = getenv("LSI_B200_HALO_ALLPHASE"); all_on = (e && atoi(e) == 0) ? 0 : 1; }
  if (f16 && all_on == 1) {   // 32-channel up-conv (upcnv1): all output phases from one halo load + transform
    HaloPlan pa;
    if (halo_plan(d, &pa, true, in_f16 != 0, true) && pa.stages >= 2 * pa.chunks) pl = pa;
  }
  LSI_REQUIRE(d->epilogue == 0 || bias, "epilogue needs a bias pointer");
  LSI_REQUIRE(!out_bn_stats || (d->epilogue == 0 && pl.n_tile == d->c_out), "bn statistics need a plain 32/64-channel conv output");
  LSI_REQUIRE(workspace_bytes >= lsi_b200_conv2d_halo_workspace_bytes(d), "workspace too small");
  LSI_REQUIRE(((uintptr_t)in & 15) == 0 && ((uintptr_t)out & 15) == 0, "tensors must be 16-byte aligned");
  LSI_REQUIRE(d->epilogue == 0 || pl.n_tile == 16 || ((uintptr_t)bias & 15) == 0, "bias must be 16-byte aligned");
  EncodeTiledFn encode = halo_get_encode();
  LSI_REQUIRE(encode != nullptr, "cuTensorMapEncodeTiled is unavailable (driver too old?)");
  cudaStream_t st = as_stream(stream);

  const int s = d->mode == 1 ? d->stride : 1;
  HaloParams p;
  LSI_REQUIRE(!out_scale || pl.n_tile == 16, "out_scale is supported on the <= 4-channel output path only");
  p.out = static_cast<float*>(out); p.bias = bias; p.out_scale = out_scale; p.in_stats = in_bn_stats; p.in_beta = in_bn_beta; p.stat_part = nullptr;
  p.Hin = d->h_in; p.Win = d->w_in; p.Ho = d->h_out; p.Wo = d->w_out; p.Co = d->c_out; p.out_cs = d->out_c_stride;
  p.Hp = d->h_out / s; p.Wp = d->w_out / s;
  p.tiles_x = (p.Wp + kTW - 1) / kTW;
  const int tiles_y = (p.Hp + kTH - 1) / kTH;
  p.per_img = p.tiles_x * tiles_y; p.spatial_tiles = p.per_img * d->batch;
  p.chunks = pl.chunks; p.chunks_a = c_in_a / kKC; p.kh = d->kh; p.kw = d->kw; p.stride = d->stride; p.pad_t = d->pad_top; p.pad_l = d->pad_left; p.mode = d->mode;
  p.n_tile = pl.n_tile; p.epilogue = d->epilogue; p.stages = pl.stages; p.f16 = f16 ? 1 : 0; p.in_f16 = in_f16 ? 1 : 0; p.out_f16 = out_f16;
  p.split = split ? 1 : 0;
  p.all_phase = pl.all_phase; p.oy_min = pl.oy_min; p.ox_min = pl.ox_min;
  p.halo_w = pl.halo_w; p.halo_h = pl.halo_h; p.halo_bytes = pl.halo_bytes; p.b_tap_bytes = pl.b_tap_bytes; p.w_bytes = pl.w_bytes;
  p.div_halo_w = 65536u / (uint32_t)pl.halo_w + 1u;

  // weights -> K-major [tap][n_tile][cin]: into the workspace, or -- when the caller vouched for a weight version -- into the memo
  float* const wk_ws = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
  float* wk = wk_ws;
  const int taps = d->kh * d->kw;
  bool prep_needed = true;
  if (wver) {
    const int sig[12] = {2, split ? 2 : (f16 ? 1 : 0), taps, d->c_in, d->c_out, pl.n_tile, d->mode, d->w_tap_stride, d->w_ci_stride, d->w_co_stride, 0, 0};
    bool hit = false;
    void* buf = prep_cache_get(w, wver, sig, (size_t)taps * pl.n_tile * d->c_in * sizeof(float) * (split ? 2 : 1), &hit);
    if (buf) { wk = static_cast<float*>(buf); prep_needed = !hit; }
  }
  if (prep_needed) {
    const long long total = (long long)taps * pl.n_tile * d->c_in;
    long long g = (total + 255) / 256; if (g > 148 * 8) g = 148 * 8;
    if (split)
      halo_prep_weights_split_l_kernel<<<(unsigned)(g * 2 > 148 * 8 ? 148 * 8 : g * 2), 256, 0, st>>>(
          w, reinterpret_cast<__half*>(wk), taps, d->c_in, d->c_out, pl.n_tile, d->w_tap_stride, d->w_ci_stride, d->w_co_stride);
    else if (f16)
      halo_prep_weights_f16_kernel<<<(unsigned)g, 256, 0, st>>>(w, reinterpret_cast<__half*>(wk), taps, d->c_in, d->c_out, pl.n_tile,
                                                                d->w_tap_stride, d->w_ci_stride, d->w_co_stride);
    else
      halo_prep_weights_kernel<<<(unsigned)g, 256, 0, st>>>(w, wk, taps, d->c_in, d->c_out, pl.n_tile, d->w_tap_stride, d->w_ci_stride,
                                                            d->w_co_stride);
    LSI_LAUNCH_CHECK();
  }
  CUtensorMap map_a, map_b, map_w;
  auto make_act_map = [&](CUtensorMap* m, const void* base, int channels, int cs) -> int {
    const cuuint64_t cm = split ? 2 : 1;   // split: 2 fp16 elements per channel, 64-element (128-byte) box rows = [hi | lo]
    cuuint64_t dims[4] = {(cuuint64_t)channels * cm, (cuuint64_t)d->w_in, (cuuint64_t)d->h_in, (cuuint64_t)d->batch};
    const cuuint64_t eb = in_f16 ? 2 : 4;   // bytes per channel (split: 2 x fp16)
    cuuint64_t strides[3] = {(cuuint64_t)cs * eb, (cuuint64_t)d->w_in * cs * eb, (cuuint64_t)d->h_in * d->w_in * cs * eb};
    cuuint32_t box[4] = {(cuuint32_t)(kKC * cm), (cuuint32_t)pl.halo_w, (cuuint32_t)pl.halo_h, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(m, (in_f16 || split) ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 4, const_cast<void*>(base), dims,
                        strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, in_f16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(activations) failed: %d", (int)r); return LSI_B200_ECUDA; }
    return LSI_B200_OK;
  };
  if (int rc = make_act_map(&map_a, in, c_in_a, d->in_c_stride)) return rc;
  if (in_b) { if (int rc = make_act_map(&map_b, in_b, d->c_in - c_in_a, in_b_c_stride)) return rc; }
  else map_b = map_a;
  {
    const cuuint64_t cm = split ? 2 : 1;   // split: 2 n_tile filter rows (Whi ; Wlo) of 32 fp16 channels per tap
    cuuint64_t dims[2] = {(cuuint64_t)d->c_in, (cuuint64_t)taps * pl.n_tile * cm};
    cuuint64_t strides[1] = {(cuuint64_t)d->c_in * ((f16 || split) ? 2 : 4)};
    cuuint32_t box[2] = {(cuuint32_t)kKC, (cuuint32_t)(pl.n_tile * cm)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&map_w, (f16 || split) ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 2, wk, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, (f16 || split) ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(weights) failed: %d", (int)r); return LSI_B200_ECUDA; }
  }
  const bool wide = pl.n_tile > 32 || (split && pl.ctas_per_sm == 1);   // the 448-thread variant: 1 CTA/SM, 8 transform warps
  const int kv = split ? (wide ? 4 : 3) : (pl.all_phase ? 2 : (wide ? 1 : 0));
  static size_t smem_set[5] = {0, 0, 0, 0, 0};
  if (pl.smem > smem_set[kv]) {
    if (kv == 4) LSI_CUDA(cudaFuncSetAttribute(conv_halo_kernel<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
    else if (kv == 3) LSI_CUDA(cudaFuncSetAttribute(conv_halo_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
    else if (kv == 2) LSI_CUDA(cudaFuncSetAttribute(conv_halo_kernel<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
    else if (kv == 1) LSI_CUDA(cudaFuncSetAttribute(conv_halo_kernel<true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
    else LSI_CUDA(cudaFuncSetAttribute(conv_halo_kernel<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
    smem_set[kv] = pl.smem;
  }
  const int G = pl.all_phase ? 1 : s * s;
  int n_ctas = halo_num_sms() * pl.ctas_per_sm;
  if (n_ctas > p.spatial_tiles * G) n_ctas = p.spatial_tiles * G;
  n_ctas = n_ctas / G * G;
  if (n_ctas < G) n_ctas = G;
  if (out_bn_stats) {
    p.stat_part = wk_ws + (size_t)taps * pl.n_tile * d->c_in * (split ? 2 : 1);
    p.stat_part = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(p.stat_part) + 255) & ~uintptr_t(255));
  }
  {
    ScopedTiming tm(kConvTc, st);
    if (kv == 4) conv_halo_kernel<true, false, true><<<dim3((unsigned)n_ctas), kThreadsAll, pl.smem, st>>>(map_a, map_b, map_w, p);
    else if (kv == 3) conv_halo_kernel<false, false, true><<<dim3((unsigned)n_ctas), kThreads, pl.smem, st>>>(map_a, map_b, map_w, p);
    else if (kv == 2) conv_halo_kernel<false, true, false><<<dim3((unsigned)n_ctas), kThreadsAll, pl.smem, st>>>(map_a, map_b, map_w, p);
    else if (kv == 1) conv_halo_kernel<true, false, false><<<dim3((unsigned)n_ctas), kThreadsWide, pl.smem, st>>>(map_a, map_b, map_w, p);
    else conv_halo_kernel<false, false, false><<<dim3((unsigned)n_ctas), kThreads, pl.smem, st>>>(map_a, map_b, map_w, p);
  }
  LSI_LAUNCH_CHECK();
  if (out_bn_stats) {
    halo_finalize_stats_kernel<<<(d->c_out + 7) / 8, 256, 0, st>>>(p.stat_part, n_ctas * ((pl.all_phase || kv == 4) ? 8 : 4), pl.n_tile, d->c_out,
                                                                  (long long)d->batch * d->h_out * d->w_out, bn_eps, out_bn_stats);
    LSI_LAUNCH_CHECK();
  }
  return LSI_B200_OK;
}

extern "C" int lsi_b200_conv2d_halo(const lsi_b200_conv_desc* d, const float* in, const float* in_bn_stats, const float* in_bn_beta,
                                    const float* w, const float* bias, const float* out_scale, float* out, float* out_bn_stats,
                                    float bn_eps, void* workspace, size_t workspace_bytes, void* stream) {
  return conv2d_halo_impl(d, in, nullptr, 0, 0, 0, in_bn_stats, in_bn_beta, w, bias, out_scale, out, 0, out_bn_stats, bn_eps, workspace,
                          workspace_bytes, stream);
}

extern "C" int lsi_b200_conv2d_halo_h(const lsi_b200_conv_desc* d, const void* in, const void* in_b, int c_in_a, int in_b_c_stride,
                                      int in_f16, const float* in_bn_stats, const float* in_bn_beta, const float* w, const float* bias,
                                      const float* out_scale, void* out, int out_f16, float* out_bn_stats, float bn_eps,
                                      void* workspace, size_t workspace_bytes, void* stream) {
  return conv2d_halo_impl(d, in, in_b, c_in_a, in_b_c_stride, in_f16, in_bn_stats, in_bn_beta, w, bias, out_scale, out, out_f16,
                          out_bn_stats, bn_eps, workspace, workspace_bytes, stream);
}

// Split-precision mode (csrc/split.cuh): lsi_b200_conv2d_halo on a split fp16-pair input (in_c_stride % 32 == 0), split on-the-fly
// weights, three exact fp16 products per fp32 product in two TMEM accumulators.  in_bn_stats/in_bn_beta as above: the producer's
// batch norm + ReLU is applied to the (hi, lo) pairs of the halo tile in shared memory.  out_kind 0: fp32 output (<= 4-channel
// prediction head with bias / sigmoid / out_scale, or plain 32/64-channel); 2: split output (plain 32/64-channel convs).
extern "C" int lsi_b200_conv2d_halo_s_supported(const lsi_b200_conv_desc* d) {
  HaloPlan pl;
  return halo_plan(d, &pl, false, false, false, true) ? 1 : 0;
}

extern "C" int lsi_b200_conv2d_halo_s(const lsi_b200_conv_desc* d, const void* in, const float* in_bn_stats, const float* in_bn_beta,
                                      const float* w, const float* bias, const float* out_scale, void* out, int out_kind,
                                      float* out_bn_stats, float bn_eps, void* workspace, size_t workspace_bytes, void* stream) {
  LSI_REQUIRE(out_kind == 0 || out_kind == 2, "out_kind must be 0 (fp32) or 2 (split)");
  return conv2d_halo_impl(d, in, nullptr, 0, 0, 0, in_bn_stats, in_bn_beta, w, bias, out_scale, out, out_kind, out_bn_stats, bn_eps,
                          workspace, workspace_bytes, stream, 1);
}
