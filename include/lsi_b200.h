/* lsi_b200.h -- C ABI of the B200-native LDI view-synthesis hot path.
 *
 * Drop-in boundary for the renderer/loss slice of google/layered-scene-inference.  The reference has no
 * FFI of its own (it is TF-1 graph Python); the functions below are what the Python call sites of
 *   lsi/geometry/ldi.py, lsi/geometry/sampling.py, lsi/geometry/projection.py, lsi/nnutils/helpers.py,
 *   lsi/loss/loss.py and the loss glue of ldi_enc_dec.py
 * bind to (through ctypes, see layered-scene-inference_b200/lsi/_b200.py and INTEGRATION.md).  Each entry
 * point cites the reference interface (file:line under the reference tree) it replaces.
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is a DEVICE pointer (fp32, contiguous) unless its name
 *    ends in `_host`; `stream` is a cudaStream_t passed as void* (NULL = default stream).
 *  - images NHWC [B,H,W,C]; LDI tensors layer-first [L,B,H,W,C]; pixel centres at +0.5; intrinsics [B,3,3]
 *    row-major; translations [B,3] (the reference's [B,3,1]).
 *  - purely functional: inputs are never written; outputs/workspaces are caller-allocated.
 *  - return value: 0 on success, LSI_B200_EINVAL (-1) for a rejected argument, LSI_B200_ECUDA (-2) for a CUDA
 *    error; lsi_b200_last_error() gives the message (thread-local).  There is NO CPU fallback: without a
 *    CUDA device every compute entry point fails with LSI_B200_ECUDA.
 *  - all launches are asynchronous on `stream`; nothing here synchronises except the *_host entry points.
 */
#ifndef LSI_B200_H_
#define LSI_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define LSI_B200_API __attribute__((visibility("default")))
#else
#define LSI_B200_API
#endif

#define LSI_B200_OK 0
#define LSI_B200_EINVAL (-1)
#define LSI_B200_ECUDA (-2)

LSI_B200_API int lsi_b200_version(void);
LSI_B200_API const char* lsi_b200_last_error(void);
/* Number of this library's kernels launched since load (bench.py's `gpu_launches`). */
LSI_B200_API unsigned long long lsi_b200_launch_count(void);

/* Measurement aid (bench.py's roofline): when enabled, CUDA events on the launching stream bracket every launch
 * of the renderer kernels.  collect() waits for them and returns, per kernel kind (0 = forward splat,
 * 1 = normalise/compose, 2 = backward target stage, 3 = backward source stage, 4 = tcgen05 conv, 5 = fp32 conv,
 * 6 = weight gradient, 7 = reserved), the summed milliseconds and the launch count since the last collect.
 * Arrays of 8. */
LSI_B200_API int lsi_b200_kernel_timing_enable(int on);
LSI_B200_API int lsi_b200_kernel_timing_collect(double* ms_by_kind, int* launches_by_kind);

/* Prepared-weights memo of the tensor-core convolutions (lsi_b200_conv2d_tc*, lsi_b200_conv2d_halo*): the K-major / fp16 / split
 * re-layout of a filter bank depends on the weights only.  A caller that knows a weight tensor has not changed since it last passed
 * `version` for it sets that (non-zero) version immediately before the conv call; the call consumes it, keeps the re-laid-out filter in
 * library-owned device memory per (weight pointer, layout) and rebuilds it only when the version differs.  Without a version (the
 * default, and every training-path call) the filter is re-laid-out into the caller's workspace as before.  The reference has no
 * counterpart: slim.conv2d reads its variable directly (nets.py:263-348). */
LSI_B200_API void lsi_b200_set_weight_version(unsigned long long version);
LSI_B200_API void lsi_b200_weight_cache_clear(void);

/* ------------------------------------------------------------------------------------------------
 * Renderer: lsi/geometry/ldi.py:71-182 forward_splat (+ projection.py:71-86, helpers.py:82-85,116-137,
 * 180-193, sampling.py:171-313 fused inside).
 * ---------------------------------------------------------------------------------------------- */
typedef struct lsi_b200_splat_desc {
  int n_layers, batch, h_s, w_s, h_t, w_t;     /* h_t = h_s*trg_downsampling (ldi.py:113-114)               */
  float trg_downsampling;                      /* ldi.py:139                                               */
  float bg_layer_disp, max_disp, zbuf_scale;   /* ldi.py:115,145                                           */
  int compose_layers;                          /* ldi.py:167-171 ; nl_out = compose ? 1 : n_layers          */
  int compute_trg_disp;                        /* ldi.py:179-182                                           */
  /* element strides between consecutive pixels of one (layer,batch) image: 3/1/1 for the reference's
   * separate tex/disp/mask tensors, 4/4 when tex and disp are views into the packed [L,B,H,W,4] head
   * output (nets.py:204).  (layer,batch) images must be densely packed: image stride = h_s*w_s*px_stride. */
  int tex_px_stride, disp_px_stride, mask_px_stride;
  int variant;                                 /* 0 = default; 1 = plain global-atomic kernel (ablation);   */
                                               /* 2 = deterministic row-owner kernel for rectified poses;    */
                                               /* 3 = block-per-segment reduction kernel without the bulk-   */
                                               /* copy ring (the default before the streaming kernel);       */
                                               /* 5 / 6 = default path + a hint from the caller, who has read */
                                               /* this call's per-image pose-class flags back before: 5 = all */
                                               /* images are rectified (row-gather kernel only), 6 = none is  */
} lsi_b200_splat_desc;

/* src->trg (inverse==0, projection.py:71-86) or trg->src (inverse!=0, projection.py:89-106) 4x4 matrices.
 * k_s,k_t,rot: [B,3,3]; t: [B,3]; out: [B,4,4] row-major. */
LSI_B200_API int lsi_b200_projection_matrix(const float* k_s, const float* k_t, const float* rot, const float* t, int batch,
                               int inverse, float* out, void* stream);

/* Scratch bytes needed by lsi_b200_forward_splat / _backward for this descriptor. */
LSI_B200_API size_t lsi_b200_forward_splat_workspace_bytes(const lsi_b200_splat_desc* d);
LSI_B200_API size_t lsi_b200_forward_splat_backward_workspace_bytes(const lsi_b200_splat_desc* d);

/* ldi.py:71-182.  tex [L,B,H,W,3], mask [L,B,H,W,1] (NULL = all ones, nets.py:205), disp [L,B,H,W,1];
 * pixel_coords [B,H,W,3] or NULL for the standard (x+0.5,y+0.5,1) grid of helpers.py:88-113;
 * focal_disps [B] or NULL (ldi.py:131-132,142-143).
 * Outputs: trg_img [nl_out,B,Ht,Wt,3], trg_wts [nl_out,B,Ht,Wt,1], trg_disp [nl_out,B,Ht,Wt,1] (required iff
 * compute_trg_disp).  layer_acc (optional, [L,B,Ht,Wt,2]) keeps the per-layer (sum w, sum w*d) accumulators
 * that the backward of trg_disp needs. */
LSI_B200_API int lsi_b200_forward_splat(const lsi_b200_splat_desc* d, const float* tex, const float* mask, const float* disp,
                           const float* pixel_coords, const float* k_s, const float* k_t, const float* rot,
                           const float* t, const float* focal_disps, float* trg_img, float* trg_wts,
                           float* trg_disp, float* layer_acc, void* workspace, size_t workspace_bytes,
                           void* stream);

/* Gradient of the above w.r.t. tex / mask / disp (what TF autodiff gives train_utils.py:113); gather form,
 * atomic-free.  trg_img/trg_wts are the saved forward outputs; g_img/g_wts/g_disp are the upstream gradients
 * (any may be NULL = zero; g_disp needs layer_acc).  d_mask may be NULL. */
LSI_B200_API int lsi_b200_forward_splat_backward(const lsi_b200_splat_desc* d, const float* tex, const float* mask,
                                    const float* disp, const float* pixel_coords, const float* k_s,
                                    const float* k_t, const float* rot, const float* t, const float* focal_disps,
                                    const float* trg_img, const float* trg_wts, const float* layer_acc,
                                    const float* g_img, const float* g_wts, const float* g_disp, float* d_tex,
                                    float* d_mask, float* d_disp, void* workspace, size_t workspace_bytes,
                                    void* stream);

/* Same as lsi_b200_forward_splat but with HOST buffers (pinned or pageable): copies the inputs to the device,
 * renders, copies trg_img/trg_wts(/trg_disp) back and synchronises.  The reference-facing end-to-end path that
 * bench.py times as `e2e`.  Device scratch is owned by the library and reused across calls. */
LSI_B200_API int lsi_b200_forward_splat_host(const lsi_b200_splat_desc* d, const float* tex_host, const float* mask_host,
                                const float* disp_host, const float* k_s_host, const float* k_t_host,
                                const float* rot_host, const float* t_host, float* trg_img_host,
                                float* trg_wts_host, float* trg_disp_host);

/* ------------------------------------------------------------------------------------------------
 * Sampling primitives: lsi/geometry/sampling.py
 * ---------------------------------------------------------------------------------------------- */
/* sampling.py:171-254 splat: out = init + scatter-add of src at coords (bilinear, 4 corners, weights <= 1e-3
 * dropped).  src [B,Hs,Ws,C], coords [B,Hs,Ws,2] (x,y), init/out [B,Ht,Wt,C]. */
LSI_B200_API int lsi_b200_splat(const float* src, const float* coords, const float* init, float* out, int batch, int h_s,
                   int w_s, int h_t, int w_t, int channels, void* stream);
/* gradient: g [B,Ht,Wt,C] -> d_src [B,Hs,Ws,C], d_coords [B,Hs,Ws,2] (d_init == g). */
LSI_B200_API int lsi_b200_splat_backward(const float* src, const float* coords, const float* g, float* d_src, float* d_coords,
                            int batch, int h_s, int w_s, int h_t, int w_t, int channels, void* stream);
/* sampling.py:41-132 bilinear (compose=True): imgs [B,Hs,Ws,C], coords [B,Ht,Wt,2] -> out [B,Ht,Wt,C]. */
LSI_B200_API int lsi_b200_bilinear(const float* imgs, const float* coords, float* out, int batch, int h_s, int w_s, int h_t,
                      int w_t, int channels, void* stream);
/* gradient: d_imgs must be zero-filled by the caller (scatter-add), d_coords [B,Ht,Wt,2]. */
LSI_B200_API int lsi_b200_bilinear_backward(const float* imgs, const float* coords, const float* g, float* d_imgs,
                               float* d_coords, int batch, int h_s, int w_s, int h_t, int w_t, int channels,
                               void* stream);
/* sampling.bilinear(imgs, coords, compose=False) (sampling.py:117-131; used by the reference's synthetic-scene generator): the four
 * corner samples masked by validity, out_ims [4,B,Ht,Wt,C], and their raw bilinear weights, out_wts [4,B,Ht,Wt,1], in the
 * reference's order (x0,y0), (x0,y1), (x1,y0), (x1,y1).  Forward only. */
LSI_B200_API int lsi_b200_bilinear_corners(const float* imgs, const float* coords, float* out_ims, float* out_wts, int batch,
                                           int h_s, int w_s, int h_t, int w_t, int channels, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Losses: lsi/loss/loss.py and the loss glue of ldi_enc_dec.py.  Scalars are device floats.
 * `partials` is scratch of lsi_b200_loss_partials_count() floats.
 * ---------------------------------------------------------------------------------------------- */
LSI_B200_API size_t lsi_b200_loss_partials_count(void);

/* loss.py:66-115 zbuffer_composition_loss.  tex [L,N,3], mask [L,N,1] (NULL = ones), disp [L,N,1], trg [N,3]
 * with N = B*H*W pixels. */
LSI_B200_API int lsi_b200_zbuf_composition_loss(const float* tex, const float* mask, const float* disp, const float* trg,
                                   int n_layers, long long n_pixels, float bg_layer_disp, float max_disp,
                                   float zbuf_scale, float* loss, float* partials, void* stream);
LSI_B200_API int lsi_b200_zbuf_composition_loss_backward(const float* tex, const float* mask, const float* disp,
                                            const float* trg, int n_layers, long long n_pixels,
                                            float bg_layer_disp, float max_disp, float zbuf_scale,
                                            const float* g_loss, float* d_tex, float* d_mask, float* d_disp,
                                            void* stream);

/* ldi_enc_dec.py:337-357: AREA-downsample gt [B,H,W,3] to Ht x Wt (integer box), mean_c |gt - render|,
 * min over the nl layers of render [nl,B,Ht,Wt,3], crop round(Wt*bdry)/round(Ht*bdry), mean. */
LSI_B200_API int lsi_b200_photo_loss(const float* render, const float* gt, int n_layers, int batch, int h, int w, int h_t,
                        int w_t, float bdry_ignore, float* loss, float* partials, void* stream);
LSI_B200_API int lsi_b200_photo_loss_backward(const float* render, const float* gt, int n_layers, int batch, int h, int w,
                                 int h_t, int w_t, float bdry_ignore, const float* g_loss, float* d_render,
                                 void* stream);

/* ldi.py:47-68 disp_smoothness_loss on disp [L*B,H,W] and loss.py:48-63 decreasing_disp_loss on [L,N]. */
LSI_B200_API int lsi_b200_disp_smoothness_loss(const float* disp, int n_images, int h, int w, float* loss, float* partials,
                                  void* stream);
LSI_B200_API int lsi_b200_disp_smoothness_loss_backward(const float* disp, int n_images, int h, int w, const float* g_loss,
                                           float* d_disp, void* stream);
LSI_B200_API int lsi_b200_decreasing_disp_loss(const float* disp, int n_layers, long long n_pixels, float* loss,
                                  float* partials, void* stream);
LSI_B200_API int lsi_b200_decreasing_disp_loss_backward(const float* disp, int n_layers, long long n_pixels,
                                           const float* g_loss, float* d_disp, void* stream);


/* ------------------------------------------------------------------------------------------------
 * CNN slice: lsi/nnutils/nets.py (slim.conv2d / conv2d_transpose / batch_norm), train_utils.py:107-117 (Adam).
 * fp32 NHWC; weights in TF layouts ([kh,kw,cin,cout] conv, [kh,kw,cout,cin] transposed conv) addressed by strides.
 * ---------------------------------------------------------------------------------------------- */
typedef struct lsi_b200_conv_desc {
  int batch, h_in, w_in, c_in, h_out, w_out, c_out;
  int kh, kw, stride, pad_top, pad_left;        /* TF SAME: pad_before = total/2 (asymmetric at stride 2)            */
  int mode;                                     /* 0: out(oy) gathers in(oy*stride - pad + ky)   (conv fwd, convT dgrad) */
                                                /* 1: out(oy) gathers in((oy + pad - ky)/stride) (convT fwd, conv dgrad) */
  int w_tap_stride, w_ci_stride, w_co_stride;   /* element strides of the weights for (tap, summed ch, produced ch)  */
  int in_c_stride, out_c_stride;                /* pixel strides (>= channels): channel slices of concat buffers      */
  int epilogue;                                 /* 0 none; 1 + bias; 2 sigmoid(. + bias) (nets.py:143,154-155)        */
  int accumulate;                               /* out += result (gradient accumulation into shared skip tensors)    */
} lsi_b200_conv_desc;

/* slim.conv2d (nets.py:273-286,...) / slim.conv2d_transpose (nets.py:296,...) forward and both data gradients. */
LSI_B200_API int lsi_b200_conv2d(const lsi_b200_conv_desc* d, const float* in, const float* w, const float* bias,
                                 float* out, void* stream);
/* Weight gradient: dw[tap][a][b] = sum big(oy*stride - pad + ky)[a] * small(oy)[b]; `big` is described by the
 * descriptor's *_in fields, `small` by its *_out fields (conv: big = input, small = dout; convT: big = dout,
 * small = input -- which yields the TF layout of either op directly). */
LSI_B200_API int lsi_b200_conv2d_wgrad(const lsi_b200_conv_desc* d, const float* big, const float* small, float* dw,
                                       void* stream);

/* CUDA-core kernels for the two contractions that are too thin for the tensor-core paths and badly shaped for the generic
 * implicit GEMM (csrc/conv_small.cu): lsi_b200_conv2d for c_in <= 8, c_out <= 32, unit stride, <= 9 taps, no epilogue (the data
 * gradient of the prediction conv, nets.py:139-155), and lsi_b200_conv2d_wgrad for c_in <= 4, c_out == 32, <= 7x7 taps,
 * stride 1 / 2 (the weight gradient of the stem cnv1, nets.py:273).  Same descriptor semantics as the generic entry points. */
LSI_B200_API int lsi_b200_conv2d_thin_supported(const lsi_b200_conv_desc* d);
LSI_B200_API int lsi_b200_conv2d_thin(const lsi_b200_conv_desc* d, const float* in, const float* w, float* out, void* stream);
LSI_B200_API int lsi_b200_conv2d_stem_wgrad_supported(const lsi_b200_conv_desc* d);
LSI_B200_API int lsi_b200_conv2d_stem_wgrad(const lsi_b200_conv_desc* d, const float* big, const float* small, float* dw,
                                            void* stream);

/* Tensor-core (tcgen05 + TMA, TF32 inputs / fp32 accumulate) version of lsi_b200_conv2d for layers whose summed
 * channel count is a multiple of 32, with an optional second input source that implements tf.concat on the fly:
 * channels [0, c_in_a) come from in_a (pixel stride d->in_c_stride), [c_in_a, c_in) from in_b (pixel stride
 * in_b_c_stride; NULL when c_in_a == c_in).  workspace: lsi_b200_conv2d_tc_workspace_bytes(d). */
LSI_B200_API int lsi_b200_conv2d_tc_supported(const lsi_b200_conv_desc* d, int c_in_a);
LSI_B200_API size_t lsi_b200_conv2d_tc_workspace_bytes(const lsi_b200_conv_desc* d);
LSI_B200_API int lsi_b200_conv2d_tc(const lsi_b200_conv_desc* d, const float* in_a, int c_in_a, const float* in_b,
                                    int in_b_c_stride, const float* w, const float* bias, float* out, void* workspace,
                                    size_t workspace_bytes, void* stream);

/* Tensor-core weight gradient (tcgen05, both operands MN-major straight from NHWC TMA boxes, split-K over pixels,
 * fp32 reductions into dw); same contract as lsi_b200_conv2d_wgrad. */
LSI_B200_API int lsi_b200_conv2d_wgrad_tc_supported(const lsi_b200_conv_desc* d);
LSI_B200_API int lsi_b200_conv2d_wgrad_tc(const lsi_b200_conv_desc* d, const float* big, const float* small, float* dw,
                                          void* stream);

/* lsi_b200_conv2d_tc that also reduces the batch-norm statistics of its output in the epilogue (per-CTA partial sums
 * from the TMEM registers, then one finalise kernel): bn_stats[c] = (mean, rsqrt(biased var + eps)). */
LSI_B200_API int lsi_b200_conv2d_tc_bnstats(const lsi_b200_conv_desc* d, const float* in_a, int c_in_a, const float* in_b,
                                            int in_b_c_stride, const float* w, float* out, float* bn_stats, float bn_eps,
                                            void* workspace, size_t workspace_bytes, void* stream);

/* Inference-only fp16 mode (lsi.nnutils.nets.set_conv_mode('f16')): lsi_b200_conv2d_tc / _bnstats with fp16 activations
 * (in_a, in_b and -- when out_f16 != 0 -- out are __half tensors; strides in elements, multiples of 8), weights rounded to
 * fp16 on the fly, kind::f16 tcgen05 MMAs with fp32 accumulation; bn_stats (optional) as in lsi_b200_conv2d_tc_bnstats, reduced
 * from the fp32 accumulators.  bias may be NULL when epilogue == 0. */
LSI_B200_API int lsi_b200_conv2d_tc_h(const lsi_b200_conv_desc* d, const void* in_a, int c_in_a, const void* in_b,
                                      int in_b_c_stride, const float* w, const float* bias, void* out, int out_f16,
                                      float* bn_stats, float bn_eps, void* workspace, size_t workspace_bytes, void* stream);

/* Split-precision mode (lsi.nnutils.nets.set_conv_mode('split'); layout and error analysis in csrc/split.cuh): the reference
 * CNN is fp32 (lsi/nnutils/nets.py:263-348) and tcgen05 has no fp32 MMA, so every activation and weight is carried as a pair of
 * fp16 numbers (hi, lo * 2^11) -- 22 mantissa bits -- and every fp32 product becomes three exact fp16 products accumulated in
 * fp32 TMEM.  A split tensor [pixels][channels] (channels % 32 == 0) has the byte size and pixel/chunk addresses of the fp32
 * tensor: per pixel and 32-channel chunk 64 bytes of hi values then 64 bytes of lo values.
 *
 * lsi_b200_conv2d_tc_s: lsi_b200_conv2d_tc on split inputs (strides in channels, multiples of 32).  out_kind 0: fp32 output,
 * any epilogue, optional out_scale[c_out] applied after the activation (`disps *= max_disp`, ldi_enc_dec.py:213);
 * out_kind 2: split output (plain convs, c_out % 32 == 0).  bn_stats (optional) as in lsi_b200_conv2d_tc_bnstats.
 * Workspace: lsi_b200_conv2d_tc_workspace_bytes. */
LSI_B200_API int lsi_b200_conv2d_tc_s(const lsi_b200_conv_desc* d, const void* in_a, int c_in_a, const void* in_b,
                                      int in_b_c_stride, const float* w, const float* bias, const float* out_scale, void* out,
                                      int out_kind, float* bn_stats, float bn_eps, void* workspace, size_t workspace_bytes,
                                      void* stream);

/* lsi_b200_conv2d_halo in the split-precision mode: one split input (in_c_stride % 32 == 0), producer's batch norm + ReLU applied
 * to the (hi, lo) pairs of the halo tile in shared memory when in_bn_stats/in_bn_beta are given.  out_kind 0: fp32 output (the
 * <= 4-channel prediction head with bias / sigmoid / out_scale, nets.py:139-155, or a plain 32/64-channel conv); 2: split output
 * (plain 32/64-channel convs, out_c_stride % 32 == 0).  Workspace: lsi_b200_conv2d_halo_workspace_bytes. */
LSI_B200_API int lsi_b200_conv2d_halo_s_supported(const lsi_b200_conv_desc* d);
LSI_B200_API int lsi_b200_conv2d_halo_s(const lsi_b200_conv_desc* d, const void* in, const float* in_bn_stats,
                                        const float* in_bn_beta, const float* w, const float* bias, const float* out_scale,
                                        void* out, int out_kind, float* out_bn_stats, float bn_eps, void* workspace,
                                        size_t workspace_bytes, void* stream);

/* x -> y between fp32 [n_pixels, channels] and split tensors (x_split / y_split != 0), optionally through
 * y = relu((x - mean) * rstd + beta) with stats[c] = (mean, rstd) (slim.batch_norm + ReLU, nets.py:263-272; beta and stats both
 * or neither).  y == x is allowed when both sides have the same layout.  channels % 32 == 0. */
LSI_B200_API int lsi_b200_split_convert(const void* x, int x_split, const float* beta, const float* stats, void* y, int y_split,
                                        long long n_pixels, int channels, void* stream);

/* Tensor-core stem (nets.py:273, cnv1: 7x7 stride-2 conv, 3 -> 32 channels, fp32 NHWC input with 12-byte pixels that TMA
 * cannot address): im2col built by the threads in shared memory (fp16, 64B swizzle), kind::f16 tcgen05 MMAs, fp32
 * accumulation; out is __half (out_f16 != 0) or float; bn_stats (optional) = (mean, rsqrt(biased var + eps)) of the raw
 * output.  Same descriptor semantics as lsi_b200_conv2d. */
LSI_B200_API int lsi_b200_conv2d_stem_tc_supported(const lsi_b200_conv_desc* d);
LSI_B200_API size_t lsi_b200_conv2d_stem_tc_workspace_bytes(void);
LSI_B200_API int lsi_b200_conv2d_stem_tc(const lsi_b200_conv_desc* d, const float* in, const float* w, void* out, int out_f16,
                                         float* bn_stats, float bn_eps, void* workspace, size_t workspace_bytes, void* stream);

/* lsi_b200_conv2d_stem_tc in the split-precision mode: the im2col operand is built as (hi, lo) fp16 pairs from the fp32 image, the
 * filter bank is split on the fly; out_kind 0: fp32 output, 2: split output (out_c_stride % 32 == 0).  Same workspace. */
LSI_B200_API int lsi_b200_conv2d_stem_tc_s(const lsi_b200_conv_desc* d, const float* in, const float* w, void* out, int out_kind,
                                           float* bn_stats, float bn_eps, void* workspace, size_t workspace_bytes, void* stream);

/* y (__half, dense [n_pixels, channels]) = relu((x - mean) * rstd + beta) with given stats[c] = (mean, rstd); x is __half
 * (x_f16 != 0) or float; channels % 8 == 0.  The normalise pass of the fp16 mode (slim.batch_norm + ReLU, nets.py:263-272). */
LSI_B200_API int lsi_b200_bn_relu_apply_h(const void* x, int x_f16, const float* beta, const float* stats, void* y,
                                          long long n_pixels, int channels, void* stream);

/* Halo-tile tensor-core convolution for the full-resolution few-channel layers of the LDI heads (nets.py:87-114: upcnv1
 * 4x4/2 up-conv 64->32, upcnv1b 3x3 32->32; nets.py:137-159: pred 3x3 32->4 + bias + sigmoid): unit-stride gathers
 * (mode 0 stride 1, or mode 1), c_in in {32,64,96,128}, c_out in {32,64} or <= 4 with out_c_stride 4.  The filter bank
 * stays resident in shared memory and each 16x8 output tile reads ONE TMA halo box per 32-channel chunk.
 * in_bn_stats/in_bn_beta (both or neither): the input is the RAW output of a slim.conv2d whose batch_norm + ReLU
 * (nets.py:263-272) has not been applied yet -- (mean, rstd)[c_in] and beta[c_in]; it is applied on load, in shared
 * memory, so the normalised tensor never exists in HBM.  out_bn_stats (optional, plain c_out in {32,64} convs): receives
 * (mean, rsqrt(biased var + eps)) of this layer's raw output, reduced in the epilogue.  out_scale (optional, <= 4-channel
 * outputs): per-channel factor applied after the activation -- `disps *= max_disp` of ldi_enc_dec.py:213 fused into the
 * prediction conv. */
LSI_B200_API int lsi_b200_conv2d_halo_supported(const lsi_b200_conv_desc* d);
LSI_B200_API size_t lsi_b200_conv2d_halo_workspace_bytes(const lsi_b200_conv_desc* d);
LSI_B200_API int lsi_b200_conv2d_halo(const lsi_b200_conv_desc* d, const float* in, const float* in_bn_stats,
                                      const float* in_bn_beta, const float* w, const float* bias, const float* out_scale,
                                      float* out, float* out_bn_stats, float bn_eps, void* workspace,
                                      size_t workspace_bytes, void* stream);

/* Same, with the RAW intermediate activations optionally stored as fp16 (in_f16 / out_f16 != 0: `in` / `out` point to
 * __half tensors, strides in elements as before).  Used between halo layers of one head (upcnv1 -> upcnv1b -> pred,
 * nets.py:87-114,137-159): the consumer normalises on load and feeds fp16 MMAs anyway, so storing the raw output in
 * fp16 halves the HBM bytes of these byte-bound layers.  in_f16 requires in_bn_stats; out_f16 a plain 32/64-channel
 * output (epilogue 0).  Statistics are reduced from the fp32 accumulators, before the rounding.
 * in_b (optional): second input source = tf.concat([in, in_b], axis=3) on the fly (the U-Net skip, nets.py:108-109):
 * channels [0, c_in_a) come from `in` (pending batch norm, in_bn_stats/in_bn_beta have c_in_a entries), channels
 * [c_in_a, c_in) from `in_b` (pixel stride in_b_c_stride), which is already normalised; same element type as `in`. */
LSI_B200_API int lsi_b200_conv2d_halo_h_supported(const lsi_b200_conv_desc* d);   /* with in_f16 != 0 (fp16 filter bank) */
LSI_B200_API int lsi_b200_conv2d_halo_h(const lsi_b200_conv_desc* d, const void* in, const void* in_b, int c_in_a,
                                        int in_b_c_stride, int in_f16, const float* in_bn_stats, const float* in_bn_beta,
                                        const float* w, const float* bias, const float* out_scale, void* out, int out_f16,
                                        float* out_bn_stats, float bn_eps, void* workspace, size_t workspace_bytes,
                                        void* stream);

/* slim.batch_norm(is_training=True, center=True, scale=False) + ReLU (nets.py:263-272): batch statistics over the
 * n_pixels = B*H*W rows; stats[c] = (mean, rsqrt(var + eps)) is kept for the backward (stats_given != 0: they were
 * already produced by lsi_b200_conv2d_tc_bnstats and only the normalise + ReLU pass runs).  workspace:
 * lsi_b200_bn_workspace_bytes(channels). */
LSI_B200_API size_t lsi_b200_bn_workspace_bytes(int channels);
LSI_B200_API int lsi_b200_bn_relu_forward(const float* x, const float* beta, float* y, float* stats, long long n_pixels,
                                          int channels, int x_c_stride, int y_c_stride, float eps, int relu,
                                          int stats_given, void* workspace, void* stream);
/* dx (and dbeta_sums[c] = (sum dz, sum dz*xhat); dbeta = the first) from dy, with dz = dy*[y>0] when relu. */
LSI_B200_API int lsi_b200_bn_relu_backward(const float* x, const float* y, const float* dy, const float* stats, float* dx,
                                           float* dbeta_sums, long long n_pixels, int channels, int x_c_stride,
                                           int y_c_stride, int dy_c_stride, int dx_c_stride, int relu, int accumulate,
                                           void* workspace, void* stream);
/* Dense fast path of lsi_b200_bn_relu_backward for contiguous [n_pixels, channels] tensors, channels % 4 == 0, with ReLU: reads
 * the raw conv output z and dy only (the mask is recomputed from z with the forward kernels' exact expression). */
LSI_B200_API int lsi_b200_bn_relu_backward_z(const float* z, const float* beta, const float* dy, const float* stats, float* dx,
                                             float* dbeta_sums, long long n_pixels, int channels, void* workspace, void* stream);
/* The same with dy read through a pixel stride (dy_c_stride >= channels, a multiple of 4): dy is the first `channels` channels of the
 * gradient of tf.concat([this layer's output, skip], axis=3) (nets.py:108-109, 300), consumed where it lies. */
LSI_B200_API int lsi_b200_bn_relu_backward_zs(const float* z, const float* beta, const float* dy, int dy_c_stride, const float* stats,
                                              float* dx, float* dbeta_sums, long long n_pixels, int channels, void* workspace,
                                              void* stream);
/* lsi_b200_bn_relu_backward in two stages, for batch statistics that span several data-parallel ranks (the reference
 * normalises over the whole batch on one device, nets.py:263-272): stage 1 writes this rank's (sum dz, sum dz*xhat) to
 * dbeta_sums; the caller all-reduces them; stage 2 computes dx from the given sums with 1 / n_pixels_stat (global count). */
LSI_B200_API int lsi_b200_bn_relu_backward_staged(const float* x, const float* y, const float* dy, const float* stats,
                                                  float* dx, float* dbeta_sums, long long n_pixels, long long n_pixels_stat,
                                                  int channels, int x_c_stride, int y_c_stride, int dy_c_stride,
                                                  int dx_c_stride, int relu, int accumulate, int stage, void* workspace,
                                                  void* stream);
/* sums[c] = (sum_x, sum_x^2) over pixels (bias gradients of the prediction conv). */
LSI_B200_API int lsi_b200_channel_sums(const float* x, float* sums, long long n_pixels, int channels, int x_c_stride,
                                       void* workspace, void* stream);
/* The same with fp64 results (the sums that cross ranks under synchronised batch norm, lsi.nnutils.nets.set_sync_bn). */
LSI_B200_API int lsi_b200_channel_sums_f64(const float* x, double* sums, long long n_pixels, int channels, int x_c_stride,
                                           void* workspace, void* stream);
/* dst[q][0..C) (+)= src[q][0..C) with independent pixel strides: tf.concat (nets.py:300,...) and its gradient. */
LSI_B200_API int lsi_b200_copy_channels(const float* src, float* dst, long long n_pixels, int channels, int src_c_stride,
                                        int dst_c_stride, int accumulate, void* stream);
/* dz = dy * y * (1 - y) for the sigmoid prediction head. */
LSI_B200_API int lsi_b200_sigmoid_backward(const float* y, const float* dy, float* dz, long long n, void* stream);
/* tf.train.AdamOptimizer.apply_gradients (train_utils.py:112-115) on one flat parameter buffer; grads are scaled by
 * grad_scale first (1/world_size after the all-reduce). step counts from 1. */
LSI_B200_API int lsi_b200_adam_step(float* params, const float* grads, float* m, float* v, long long n, float learning_rate,
                                    float beta1, float beta2, float epsilon, long long step, float grad_scale, void* stream);

/* GPU-native synthetic planar-room generator (lsi/data/syntheticPlanes/data.py:372-420 with lsi/geometry/homography.py:95-156 and
 * lsi/geometry/layers.py:29-118): for `batch` views of worlds of n_planes (<= 16) textured planes, one fused pass per target
 * pixel over the planes -- hom_t2w [batch,n,9] takes target pixels to texture pixels (homography.inv_homography), dmat_t
 * [batch,n,3] gives the plane's disparity at a target pixel (inv_homography_dmat); bilinear texture + mask sample, hard soft-z
 * selection with a white background layer at min_disp.  Outputs: out_img [batch,h_out,w_out,3] (layers.compose, soft=False),
 * out_disp_fg / out_disp_bg [batch,h_out,w_out,1] (layers.compose_depth, bg_layer False / True; either may be NULL).
 * scratch_gmax: batch floats. */
LSI_B200_API int lsi_b200_render_planes(const float* imgs_w, const float* masks_w, const float* hom_t2w, const float* dmat_t,
                                        int batch, int n_planes, int h_tex, int w_tex, int h_out, int w_out, float min_disp,
                                        float depth_softmax_temp, float* out_img, float* out_disp_fg, float* out_disp_bg,
                                        float* scratch_gmax, void* stream);
/* Procedural stand-ins for the PASCAL / SUN textures of the reference's generator: img [n,h,w,3] = clamp(0.5 + sum_k amp sin(fx x +
 * fy y + phase)) per channel from params [n,3,n_waves,4]; mask [n,h,w,1] = 1 (kind 0) or a soft super-ellipse alpha (kind != 0). */
LSI_B200_API int lsi_b200_procedural_texture(const float* params, const int* kind, int n_textures, int n_waves, int h, int w,
                                             float* img, float* mask, void* stream);
/* tf.image.resize_images(AREA) by an integer factor on float NHWC (data.py:364-368): box mean. */
LSI_B200_API int lsi_b200_box_downsample(const float* in, float* out, int batch, int h, int w, int channels, int factor,
                                         void* stream);

/* KITTI loader data path (lsi/data/kitti/data.py:247-266): 8-bit HWC image (c_in channels, the first nc are used) -> float
 * [h_out, w_out, nc] in [0, 1], resized with the semantics of tf.image.resize_images(method=AREA) (exact area weighting,
 * any scale factor). */
LSI_B200_API int lsi_b200_area_resize_u8(const unsigned char* in, int h_in, int w_in, int c_in, float* out, int h_out,
                                         int w_out, int nc, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LSI_B200_H_ */
