/*
 * havc_b200.h — C ABI of libhavc_b200.so, the B200 (sm_100a) device library behind the
 * vsdeoldify_b200 Python plugin surface (HAVC_main / HAVC_colorizer / HAVC_deoldify /
 * HAVC_ddeoldify / HAVC_merge).
 *
 * The reference (dan64/vs-deoldify 5.6.7) has NO FFI boundary of its own: its hot path is Python
 * calling torch (cuDNN/cuBLAS) and Pillow/OpenCV/numpy.  Each entry point below names the
 * reference call site whose device (or host-pixel) work it replaces.  All pointers are plain
 * device pointers unless a parameter says "host"; sizes are plain integers; no torch types.
 * Every function returns 0 on success and a negative havc_status on failure; the text of the
 * last failure (per calling thread) is available from havc_last_error().
 * `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream).
 *
 * Activations are NHWC ("pixel-major, channel-contiguous") 16-bit (fp16 or bf16) tensors whose
 * channel count and all strides are multiples of 8 elements (16 bytes, a TMA requirement).
 */
#ifndef HAVC_B200_H
#define HAVC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    HAVC_OK = 0,
    HAVC_ERR_ARG = -1,      /* invalid argument / unsupported shape            */
    HAVC_ERR_CUDA = -2,     /* a CUDA runtime / driver call failed             */
    HAVC_ERR_NO_DEVICE = -3 /* no sm_100 device / driver entry point available */
} havc_status;

enum { HAVC_F16 = 0, HAVC_BF16 = 1, HAVC_F32 = 2 };

#define HAVC_MAX_TAPS 16

/* A strided view of an NHWC activation tensor with an optional leading "phase" dimension
 * (P > 1 only for the phase-split inputs of stride-2 convolutions).  Strides in ELEMENTS. */
typedef struct {
    const void *ptr;
    int32_t C, W, H, B, P;
    int64_t stride_w, stride_h, stride_b, stride_p;
} havc_act_view;

/* One implicit-GEMM convolution / batched GEMM launch:
 *   acc[m, n] = sum_{tap, c} A[pixel(m) + (dh,dw)[tap], c] * Wt[n, tap, c]     (fp32 accumulate in TMEM)
 *   y = acc + bias[n];  relu1;  y = y*scale[n] + shift[n];  y += residual[pixel, n];  relu2;  store.
 * Replaces every torch.nn.Conv2d / Conv1d / bmm call of the reference's networks
 * (vsdeoldify/deoldify/unet.py:24-285, vsdeoldify/fastai/layers.py:81-96,
 *  torchvision resnet blocks via vsdeoldify/fastai/vision/learner.py:54-63,
 *  vsdeoldify/colorization/colorizers/eccv16.py:87-98, siggraph17.py:133-161) together with the
 * BatchNorm / ReLU / residual-add / PixelShuffle / torch.cat modules fused around them.
 */
typedef struct {
    havc_act_view src0;         /* first K source                                              */
    havc_act_view src1;         /* second K source (torch.cat partner); ptr == NULL if unused   */
    const void *weight;         /* [w_batches][w_rows][w_taps][w_cin] 16-bit, K(c)-contiguous   */
    int32_t w_rows, w_taps, w_cin, w_batches;
    int32_t w_c1_off;           /* position of src1's first channel on the weight's cin axis    */
    int32_t n_taps;
    int8_t tap_dh[HAVC_MAX_TAPS], tap_dw[HAVC_MAX_TAPS], tap_p[HAVC_MAX_TAPS], tap_wi[HAVC_MAX_TAPS];
    int32_t out_B, out_H, out_W;/* iteration space of the GEMM M dimension (output pixels)      */
    int32_t box_w, box_h, box_b;/* M tile = box_w*box_h*box_b = 128 output pixels               */
    int32_t a_batched, b_batched;
    int32_t BN;                 /* N tile (multiple of 16, <= 320; > 256 is issued as 256 + rest) */
    int32_t N_total;            /* GEMM N including padding (multiple of 16)                    */
    const float *bias, *scale, *shift; /* per GEMM column, >= ceil(N_total/BN)*BN entries; NULL = skip */
    int32_t relu1, relu2;
    const void *residual;       /* same dtype as the activations; NULL = none                   */
    int64_t res_stride_w, res_stride_h, res_stride_b;
    void *out;
    int32_t out_dtype;          /* HAVC_F16 / HAVC_BF16 / HAVC_F32                              */
    int64_t out_stride_w, out_stride_h, out_stride_b;
    int32_t up, oy, ox;         /* out pixel = (h*up + oy, w*up + ox)                           */
    int32_t shuffle;            /* 1: PixelShuffle(2) store, column n -> group n/group_n        */
    int32_t group_n;            /* padded columns per shuffle group (multiple of 16)            */
    int32_t c_store;            /* channels written per pixel (multiple of 8)                   */
    int32_t dtype;              /* HAVC_F16 or HAVC_BF16 operands                               */
    /* K-loop variant: src1 is an im2col'd side tensor that contributes ONE tap (offset 0) after all taps of src0;
     * its weights live at weight tap index src1_wi, channels [w_c1_off, ...).  Used for the 3 image channels of
     * MergeLayer(dense=True) (unet.py:273): 9 taps x 3 channels become one 64-wide K chunk instead of nine.   */
    int32_t src1_single_tap, src1_wi;
    /* Column split: GEMM columns n >= split_n (multiple of 16; 0 = off) are stored to out2 at channel n - split_n
     * and take their residual from residual2.                                                                  */
    int32_t split_n;
    void *out2;
    int64_t out2_stride_w, out2_stride_h, out2_stride_b;
    int32_t c_store2;
    const void *residual2;
    int64_t res2_stride_w, res2_stride_h, res2_stride_b;
    /* Fused 1x1 head (unet.py:276-278 `custom_conv_layer(ni, 3, ks=1)`): instead of storing the tile, write
     * head_out[pixel][o] = sum_n y[n] * head_w[o][n] (o < 3; fp32, 4 floats per pixel; head_w is [3][N_total]).
     * Requires a single N tile (BN == N_total).                                                                */
    const float *head_w;
    float *head_out;
    int64_t head_stride_w, head_stride_h, head_stride_b;
    /* 1: stage the output tile in shared memory and write it with TMA bulk tensor stores (coalesced, clipped at the
     * tensor bounds).  Honoured for 16-bit NHWC / PixelShuffle outputs with BN % 64 == 0; otherwise the epilogue falls
     * back to per-thread vector stores.                                                                           */
    int32_t tma_store;
    /* relu1 becomes LeakyReLU(leaky1) when leaky1 > 0 (siggraph17.py:98 `nn.LeakyReLU(negative_slope=.2)`); 0 = ReLU. */
    float leaky1;
    /* CTA pairs: two SMs of a cluster share one 256-pixel M tile pair through tcgen05.mma.cta_group::2, each loading half of
     * the weight rows (halves the dominant L2 -> SM operand traffic).  0 = library default (on), 1 = force, -1 = off.   */
    int32_t pair;
    /* Split-precision operands ("x3"): a value is stored as hi + lo, two 16-bit planes of identical geometry (hi = the value
     * rounded to 16 bit, lo = the rounded remainder), and the contraction issues hi*hi + lo*hi + hi*lo (three MMAs per K step,
     * fp32 accumulate: ~2^-20 relative operand precision instead of 2^-11).  Any of the lo pointers may be NULL (that product
     * is skipped).  src0_lo / src1_lo / weight_lo / residual_lo have the strides of their hi tensors; out_lo (fast TMA-store
     * epilogue only, BN <= 128, no PixelShuffle) receives the lo plane of the result.  Used where the error budget
     * (tools/precision_emulator.py) shows 16-bit storage dominates the logit error: the ResNet encoders.             */
    const void *src0_lo, *src1_lo, *weight_lo, *residual_lo;
    void *out_lo;
    /* blur = 1 (with shuffle = 1): the ICNR blur that follows the PixelShuffle in (Custom)PixelShuffle_ICNR
     * (vsdeoldify/deoldify/unet.py:47-52: ReplicationPad2d((1,0,1,0)) + AvgPool2d(2, stride=1)) is applied in the epilogue, so the
     * shuffled tensor never goes to HBM un-blurred.  Column layout of the packed weight / bias: N tile t holds channels
     * [t*BN/4, (t+1)*BN/4) as four consecutive sub-pixel groups g = 2*dy + dx, column = t*BN + g*BN/4 + (c - t*BN/4); the M tiles
     * overlap by one halo row / column (recomputed), box_b must be 1; epilogue = bias + ReLU only.                          */
    int32_t blur;
} havc_conv_desc;

const char *havc_last_error(void);
int havc_version(void);
/* Number of kernels this library has launched since load / last reset (bench.py "gpu_launches"). */
int64_t havc_launch_count(void);
void havc_launch_count_reset(void);

int havc_conv_gemm(const havc_conv_desc *d, void *stream);

/* ---- memory-bound network ops (NHWC 16-bit, C multiple of 8) -------------------------------- */

/* im2col for tiny-Cin convolutions over 8-channel-wide input pixels.  K layout: filter row kh occupies
 * [kh*RW, kh*RW + ks*cin) with RW = round_up(ks*cin, 8): out[b,oy,ox, kh*RW + kw*cin + c] =
 * in[b, oy*stride-pad+kh, ox*stride-pad+kw, c0 + c] (c0 = 0 or 4), zero outside the image; other K positions are never written
 * (zero-initialise the buffer once).  Feeds havc_conv_gemm for the ResNet 7x7/s2 stem (torchvision conv1 behind
 * vsdeoldify/fastai/vision/learner.py:54-63), for the 3 image channels of MergeLayer(dense=True)
 * (vsdeoldify/deoldify/unet.py:273) and for Zhang's model1.0 (colorizers/eccv16.py:16, siggraph17.py:20).
 * in: [B,H,W,8], out: [B,OH,OW,Kp]. */
int havc_im2col_small(const void *in, void *out, int B, int H, int W, int Cs, int c0, int cin, int ks, int stride,
                      int pad, int Kp, int dtype, void *stream);
/* nn.MaxPool2d(3, 2, 1) of the torchvision resnet stem.  in_lo / out_lo: optional lo planes of a split-precision tensor
 * (value = hi + lo, see havc_conv_desc), NULL for plain 16-bit tensors. */
int havc_maxpool3x3s2(const void *in, const void *in_lo, void *out, void *out_lo, int B, int H, int W, int C, int dtype,
                      void *stream);
/* out[p=a*2+b][n][i][j][:] = in[n][2i+a][2j+b][:] — the input layout of a stride-2 havc_conv_gemm. */
int havc_phase_split(const void *in, void *out, int B, int H, int W, int C, int n_phases, void *stream);
/* y = x*scale[c]+shift[c] (+ReLU): eval BatchNorm on U-Net skips / encoder output
 * (vsdeoldify/deoldify/unet.py:203, :244).  Pixel strides (elements) let it read/write a channel slice of a
 * wider NHWC buffer, which is how torch.cat partners are written in place (MergeLayer, fastai/layers.py:149-152).
 * in_lo / out_lo: optional lo planes (same strides) of split-precision tensors. */
int havc_affine_act(const void *in, const void *in_lo, void *out, void *out_lo, long long n_pixels, int C, int in_pix_stride,
                    int out_pix_stride, const float *scale, const float *shift, int relu, int dtype, void *stream);
/* ReplicationPad2d((1,0,1,0)) + AvgPool2d(2,1) of (Custom)PixelShuffle_ICNR
 * (vsdeoldify/deoldify/unet.py:47-52, vsdeoldify/fastai/layers.py:214-220). */
int havc_blur2x2(const void *in, void *out, int B, int H, int W, int C, int out_pix_stride, int dtype, void *stream);
/* Row soft-max of fp32 (in_dtype = HAVC_F32) or fp16 (HAVC_F16) attention logits -> 16-bit probabilities (F.softmax(.., dim=1) of
 * vsdeoldify/fastai/layers.py:94, stored transposed so the reduction runs along contiguous rows). */
int havc_softmax_rows(const void *in, int in_dtype, void *out, long long rows, int cols, int in_stride, int out_stride,
                      int out_dtype, void *stream);

/* ---- frame pre / post pixel passes (planar u8 RGB frames [B][3][H][W]) ----------------------- */

/* Phase-periodic form of the two horizontal passes (integer ratios: 1920 <-> 384 / 480, 3840 <-> 640): every interior column of a
 * phase has the same weights and a window that moves by a fixed step, so the kernels keep the weights in their parameter bank
 * and a run of input pixels in registers (no table look-ups, no shared-memory load per FMA).  The host proves the periodicity
 * of the table bit for bit (vsdeoldify_b200/resample.py:periodic_plan) and passes the interior range; border columns, whose
 * windows are folded at the image edge, take the table path inside the same kernel.  Results are bit-identical to
 * havc_resample_h / havc_post_horizontal (same fmaf chain per output). */
typedef struct havc_periodic_plan {
    int ratio;   /* Win / Wout (squeeze) or W / S (way back): 4, 5 or 6 */
    int taps;    /* padded taps per output: squeeze = (window start mod 4) leading zeros + table taps; back = table taps + spread of the phase starts */
    int offset;  /* squeeze: aligned window start of column block 0 (input pixels, multiple of 4, may be negative); back: -min phase start */
    int lo, hi;  /* interior range: squeeze in blocks of 8 output columns, back in units of 4 input positions */
    float w[72]; /* squeeze: w[t]; back: w[phase * taps + t] */
} havc_periodic_plan;
int havc_resample_h_periodic(const uint8_t *in, float *out, long long rows, int Win, int Wout, const int *start,
                             const float *weights, int taps, const havc_periodic_plan *plan, void *stream);
int havc_post_horizontal_periodic(const float *in, const uint8_t *orig, uint8_t *out, int B, int S, int H, int W,
                                  const int *start, const float *weights, int taps, int transplant,
                                  const havc_periodic_plan *plan, void *stream);

/* Table-driven separable resampling, horizontal pass: out[row][o] = sum_t weights_t[t][o]*in[row][start[o]+t]
 * (horizontal passes take the weight table TRANSPOSED, [taps][Wout], so a warp reads it coalesced; vertical
 * passes take it as [Hout][taps]).  With Spline64 tables this is zimg's resize.Spline64 of
 * vsdeoldify/__init__.py:2504 and :3547. */
int havc_resample_h(const uint8_t *in, float *out, long long rows, int Win, int Wout, const int *start,
                    const float *weights, int taps, void *stream);
/* Vertical pass of the squeeze to S x S, fused with ColorizerFilter._transform (Pillow 'L' luma,
 * vsdeoldify/deoldify/filters.py:92-93) and ImageNet normalisation (filters.py:50-53).
 * in: float [B][3][Hin][S]; rgb_small: u8 [B][3][S][S]; x: 16-bit NHWC [B][S][S][8]. */
int havc_pre_vertical(const float *in, uint8_t *rgb_small, void *x, int B, int Hin, int S, const int *start,
                      const float *weights, int taps, int dtype, void *stream);
/* ColorizerFilter._transform (Pillow 'L' gray, filters.py:92-93) + ImageNet normalisation (filters.py:50-53) of an
 * already S x S image (the direct ModelImageRender path, where Pillow did the squeeze): rgb u8 [B][3][n] -> x 16-bit
 * NHWC [B][n][8]. */
int havc_gray_normalize(const uint8_t *rgb, void *x, int B, long long n_pixels, int dtype, void *stream);
/* Network head: 1x1 conv 259->3 + SigmoidRange(-3,3) (unet.py:276-281), de-normalise, clamp, *255, truncate
 * to u8 (filters.py:64-67), and — if transplant — ColorizerFilter._post_process (filters.py:100-110) at S x S
 * against rgb_small.  res: [B,S,S,Cs] 16-bit; w11: fp32 [3][Cs]; colored: u8 [B][3][S][S];
 * With Cs == 0, `res` is instead fp32 logits [B][S][S][4] already produced by havc_conv_gemm's fused head
 * (bias b11 not yet added) and w11 is ignored.
 * net_out (optional): fp32 [B][3][S][S] network output for parity tests; skip (optional): u8 [B], 1 = the
 * scene-change gate of vsslib/vsmodels.py:221-224 returned this frame uncoloured (colored := rgb_small). */
int havc_head(const void *res, int Cs, const float *w11, const float *b11, const uint8_t *rgb_small,
              uint8_t *colored, float *net_out, const uint8_t *skip, int B, int S, int dtype, int transplant,
              void *stream);
/* PIL.Image.blend(a, b, alpha) = trunc(a + alpha*(b-a)) on n u8 values (n % 16 == 0): the 50/50 mix of the
 * 'stable'/'artistic' generator with the 'video' one (vsdeoldify/deoldify/visualize.py:129,135) and
 * image_weighted_merge (vsdeoldify/vsslib/imfilters.py:113-124). */
int havc_blend_u8(const uint8_t *a, const uint8_t *b, uint8_t *out, long long n, float alpha, void *stream);
/* Vertical pass on u8 planes (first pass of the resize back to W x H, run on the S-wide image):
 * out[plane][oy][x] = sum_t weights[oy][t] * in[plane][start[oy]+t][x].  in: u8 [planes][Hin][W]; out: float. */
int havc_resample_v(const uint8_t *in, float *out, long long planes, int Hin, int Hout, int W, const int *start,
                    const float *weights, int taps, void *stream);
/* Vertical pass float -> u8 (round half to even, clamp): the second pass of a horizontal-first resize to a rectangular size
 * (vsslib/vsresize.py:30-75 resize_min_HW -> resize.Spline36).  in float [planes][Hin][W]; out u8 [planes][Hout][W];
 * weights [Hout][taps]. */
int havc_resample_v_f32_u8(const float *in, uint8_t *out, long long planes, int Hin, int Hout, int W, const int *start,
                           const float *weights, int taps, void *stream);
/* Final horizontal pass of the resize back to W x H fused with vs_recover_clip_luma / chroma_post_process
 * (vsdeoldify/vsslib/vsfilters.py:863-899, imfilters.py:312-321): keep the luma of `orig`, the chroma of the
 * upscaled colour image (OpenCV Q14 8-bit YUV).  in: float [B][3][H][S]; orig/out: u8 [B][3][H][W];
 * weights transposed [taps][W]. */
int havc_post_horizontal(const float *in, const uint8_t *orig, uint8_t *out, int B, int S, int H, int W,
                         const int *start, const float *weights, int taps, int transplant, void *stream);

/* ---- vsslib model merges and chroma-adjust filters (planar u8 RGB frames [B][3][H][W]) ---------------------
 * Per-frame statistics live on the device as exact integer sums: stats[2*b] = sum of OpenCV 8-bit Y,
 * stats[2*b+1] = sum of Pillow 'L' of frame b; the kernels derive get_image_luma() = round(mean(Y)/255, 6)
 * (vsdeoldify/vsslib/imfilters.py:597-601) and ImageStat means from them, so no host round trip is needed.
 * `simd_width` is the per-row vector block of the host OpenCV build whose 8-bit HSV->RGB the reference calls
 * (its SIMD body truncates x*255, the scalar row tail rounds; 32 for an AVX2 build, 0 = round everywhere). */
#define HAVC_MAX_HUE_RANGES 8
typedef struct {
    int32_t n;                                   /* number of (min, max) hue ranges in DEGREES (0..360), strict bounds */
    double lo_deg[HAVC_MAX_HUE_RANGES], hi_deg[HAVC_MAX_HUE_RANGES];
} havc_hue_ranges;

/* Frame sums of `img` (or of ImageEnhance.Brightness(img).enhance(bright) when bright != 1). */
int havc_frame_stats(const uint8_t *img, int B, int H, int W, float bright, unsigned long long *stats, void *stream);
/* chroma_stabilizer (imfilters.py:160-200; adaptive=0) / chroma_stabilizer_adaptive (imfilters.py:202-269; adaptive=1):
 * U,V of b clamped around those of a, Y of a, Image.blend(a, result, weight) when weight < 1.  stats_out (optional)
 * receives the Y sums of the result for the red fix. */
int havc_chroma_stabilizer(const uint8_t *a, const uint8_t *b, uint8_t *out, int B, int H, int W, int adaptive, double alpha,
                           int base_tol, int max_extra, float weight, unsigned long long *stats_out, void *stream);
/* Dark-frame red-shift adjustment of ConstrainedChromaMerge / ChromaBoundAdaptiveMerge (mcomb.py:351-362, :405-416):
 * image_tweak(sat, hue_range="280:360,0:30") + w_image_luma_merge chosen by the frame luma in `stats`. */
int havc_red_fix(const uint8_t *stab, uint8_t *out, int B, int H, int W, const unsigned long long *stats, void *stream);
/* LumaMaskedMerge.merge_frame (mcomb.py:238-271): image_luma_merge / w_image_luma_merge of (c, b), then
 * image_weighted_merge(a, masked, weight). */
int havc_luma_masked_merge(const uint8_t *a, const uint8_t *b, const uint8_t *c, uint8_t *out, int B, int H, int W,
                           double luma_limit, double white_limit, float weight, void *stream);
/* AdaptiveLumaMerge.merge_frame (mcomb.py:289-314): Image.blend(a, b, w') with w' from the frame luma of b. */
int havc_adaptive_luma_merge(const uint8_t *a, const uint8_t *b, uint8_t *out, int B, int H, int W,
                             const unsigned long long *stats_b, double luma_threshold, double alpha, double weight,
                             double min_weight, void *stream);
/* restore_color_gradient (restcolor.py:98-134) with the luma gating of vs_sc_recover_gradient_color
 * (vsfilters.py:403-409; stats_gray != NULL) and the final std.Merge(gray, restored, merge_weight) of
 * ChromaRetentionMerge (mcomb.py:511, vsfilters.py:730-739; merge_weight < 0 = none).  lut / lut_gated: device
 * tables [256] of w_np_gradient_mask(S) for the normal and the gated alpha (built by the host). */
int havc_restore_color_gradient(const uint8_t *color, const uint8_t *gray, uint8_t *out, int B, int H, int W, double sat,
                                const uint8_t *lut, const uint8_t *lut_gated, double weight, double weight_gated,
                                const unsigned long long *stats_gray, double merge_weight, int simd_width, void *stream);
/* The temporal chroma stabiliser (scope row N3: vs_chroma_stabilizer_ex, vsslib/vsfilters.py:84-287).
 * havc_gray_mask_stats: stats[b] = { sum of OpenCV Y (get_image_luma, imfilters.py:597-601), number of pixels whose OpenCV
 *   HSV saturation is < tht (the gray mask of restore_color, restcolor.py:51-55) }.
 * havc_restore_color: restore_color (restcolor.py:38-83) as vs_sc_recover_clip_color calls it (vsfilters.py:327-351): gray pixels
 *   of `gray` take the (desaturated) colours of `color`; weight > 0 merges with `gray`, < 0 with the colours; frames with a luma
 *   outside [0.22, 0.78] use min(weight, -0.8); a frame whose gray share exceeds tht_scen, or with active[b] == 0 (frame number
 *   < 15), comes back untouched.  lut: device table [256], lut[s] = s < tht ? 255 : 0.
 * havc_average_frames_u8: std.AverageFrames on 8-bit planes (vsfilters.py:58,242,272; VapourSynth's filter restated, unpinned):
 *   out[b] = clamp((sum_k w[k] * src[k * clip_stride + b * frame_elems + ..] + scale / 2) / scale); clip_stride = frame_elems
 *   gives the temporal form on a sequence that carries its halo frames; weights are [n_clips] or per frame [B][n_clips]. */
int havc_gray_mask_stats(const uint8_t *img, int B, int H, int W, int tht, unsigned long long *stats, void *stream);
int havc_restore_color(const uint8_t *color, const uint8_t *gray, uint8_t *out, int B, int H, int W, double sat, int tht,
                       double weight, double tht_scen, const unsigned long long *stats_gray, const uint8_t *active,
                       const uint8_t *lut, int simd_width, void *stream);
int havc_average_frames_u8(const uint8_t *src, long long clip_stride, int n_clips, const int *weights, int weights_per_frame,
                           int scale, uint8_t *out, int B, long long frame_elems, void *stream);
/* adjust_chroma (restcolor.py:239-286) behind adjust_hue_range / vs_sc_adjust_clip_hue (vsfilters.py:435-455). */
int havc_adjust_chroma(const uint8_t *img, uint8_t *out, int B, int H, int W, const havc_hue_ranges *ranges, double sat, int hue,
                       double weight, int simd_width, void *stream);
/* np_image_chroma_tweak (restcolor.py:288-342): cv2 HSV hue add / S scale / V scale and, when `ranges` != NULL, the
 * "chroma adjustment" stage (mask from the tweaked hue, unmasked pixels from the original image, weight blend), optionally
 * fused with the luma merge that follows it in vs_sc_chroma_bright_tweak (vsfilters.py:525-552): luma_merge != 0 ->
 * out = image_luma_merge / w_image_luma_merge(img_dark = tweaked, img_white = img, luma_limit, white_limit)
 * (imfilters.py:66-100).  luma_merge == 0 with ranges is _vs_sc_colormap (vsfilters.py:577-590).  These are the `smooth` and
 * `colormap` stages of HAVC_stabilizer (vsdeoldify/__init__.py:2852-2858). */
int havc_chroma_tweak(const uint8_t *img, uint8_t *out, int B, int H, int W, double sat, double bright, int hue,
                      const havc_hue_ranges *ranges, double sat2, int hue2, double weight, int luma_merge, double luma_limit,
                      double white_limit, int simd_width, void *stream);
/* image_tweak (imfilters.py:463-504) without gamma / hue (gamma raises in the reference, see oracle/filters_oracle.py):
 * ImageEnhance Brightness -> Contrast -> Color, optionally restricted to hue ranges of the input.  stats_scratch
 * (2*B u64) is needed when cont != 1. */
int havc_image_tweak(const uint8_t *img, uint8_t *out, int B, int H, int W, double sat, double cont, double bright,
                     const havc_hue_ranges *ranges, unsigned long long *stats_scratch, void *stream);
/* luma_adjusted_levels (imfilters.py:335-372) behind sc_constrained_tweak (vsfilters.py:656-675). */
int havc_luma_adjusted_levels(const uint8_t *img, uint8_t *out, int B, int H, int W, const unsigned long long *stats,
                              double luma_min, double gamma, double gamma_luma_min, double gamma_alpha, double gamma_min,
                              void *stream);

/* ---- Zhang et al. colorizers: pre/post passes (vsdeoldify/colorization/__init__.py:76-95, colorizers/util.py:21-55) ---- */

/* One separable pass of Pillow's Image.resize on 8-bit planes (ImagingResample, libImaging/Resample.c): 22-bit
 * fixed-point coefficients, (sum + 2^21) >> 22, clip.  horizontal=1: in [planes][Hin][Win] -> out [planes][Hin][out_size];
 * horizontal=0: -> out [planes][out_size][Win].  bounds: int [out_size][2] = (first tap, count); coeffs: int [out_size][ksize].
 * Pillow runs the horizontal pass first and stores a uint8 intermediate; BICUBIC feeds preprocess_img (util.py:21-30),
 * BILINEAR is BaseFilter._scale_to_square / _unsquare (vsdeoldify/deoldify/filters.py:37-41,70-73). */
int havc_pil_resample_u8(const uint8_t *in, uint8_t *out, long long planes, int Hin, int Win, int out_size, int horizontal,
                         const int *bounds, const int *coeffs, int ksize, void *stream);
/* skimage.color.rgb2lab L channel (float64, stored float32 like torch.Tensor(img_l), util.py:29-36) of planar u8 RGB
 * [B][3][n]; L_out: float [B][n] (optional); x: 16-bit NHWC [B][n][8] whose channel 0 receives (L-50)/100
 * (BaseColor.normalize_l, base_color.py:13-14) (optional). */
int havc_zhang_pre(const uint8_t *rgb, int B, long long n_pixels, float *L_out, void *x, int dtype, void *stream);
/* eccv16 tail (eccv16.py:94): softmax over n_classes logits (fp32, row stride ld) and the bias-free 1x1 conv
 * n_classes -> 2 (w_out: float [2][n_classes]); out: float [pixels][2]. */
int havc_eccv_head(const float *logits, int ld, int n_classes, const float *w_out, float *out, long long pixels, void *stream);
/* siggraph17 tail (siggraph17.py:101-102,159-161): out[p] = tanh(head[p][0..1] + bias) * mul on the fused-head output
 * ([pixels][4] fp32) of the last havc_conv_gemm. */
int havc_zhang_tanh(const float *head, const float *bias, float *out, long long pixels, float mul, void *stream);
/* torch bilinear resize, align_corners=False, of 2-channel interleaved float maps [B][h][w][2] -> [B][H][W][2], times mul
 * (nn.Upsample(scale_factor=4) + unnormalize_ab, eccv16.py:84,96). */
int havc_bilinear_ab(const float *in, float *out, int B, int h, int w, int H, int W, float mul, void *stream);
/* postprocess_tens + uint8 conversion (util.py:38-55, colorization/__init__.py:93-95): ab [B][h][w][2] resized
 * bilinearly to H x W, concatenated with L [B][H][W], skimage lab2rgb in float64, uint8(clip(x*255, 0, 255)) ->
 * planar u8 RGB [B][3][H][W]. */
int havc_zhang_post(const float *ab, int h, int w, const float *L, uint8_t *out, int B, int H, int W, void *stream);

/* chroma_post_process (vsslib/imfilters.py:312-321) == ColorizerFilter._post_process (deoldify/filters.py:100-110) ==
 * vs_recover_clip_luma (vsfilters.py:863-899): luma of `orig`, chroma of `color` through OpenCV's 8-bit Q14 YUV. */
int havc_chroma_post_process(const uint8_t *color, const uint8_t *orig, uint8_t *out, int B, int H, int W, void *stream);
/* Scene-change gate of the per-frame selectors (vsslib/vsmodels.py:221-224, mcomb.py:210-213: `return f[0].copy()`):
 * frames b with skip[b] != 0 are overwritten by the same frame of `src`. */
int havc_select_frames(uint8_t *dst, const uint8_t *src, const uint8_t *skip, int B, long long frame_bytes, void *stream);
/* VapourSynth std.Merge(clipa, clipb, weight) on 8-bit samples (vs_simple_merge, vsfilters.py:730-739; HAVC_merge
 * method 2, vsdeoldify/__init__.py:2648): 15-bit fixed-point weight, out = a + (((b - a)*w15 + 2^14) >> 15).
 * Restated from the VapourSynth sources; the library is absent here, so this edge is parity-unpinned. */
int havc_vs_merge_u8(const uint8_t *a, const uint8_t *b, uint8_t *out, long long n, double weight, void *stream);

/* ---- zimg-backed steps of vs_tweak (vsdeoldify/vsslib/vsfilters.py:753-850; also vs_clip_color_stabilizer :38-61 and the
 * luma_mask_sat branch of LumaMaskedMerge, mcomb.py:243), restated from the published zimg algorithm (zimg / VapourSynth are not
 * available here: parity unpinned against the real library, bit-exact against oracle/zimg_oracle.py).  Planar u8 batches. -------- */

/* clip.resize.Bicubic(format=YUV420P8, matrix_s=..., range_s=..., [dither_type="error_diffusion"]): vs_tweak (vsfilters.py:790:
 * matrix 709, full range, no dither) and restore_format (havc_utils.py:199-222: the clip's matrix / range, dither on Y, U and V).
 * rgb u8 [B][3][H][W] -> y u8 [B][H][W], uv u8 [B][2][H/2][W/2].  scratch444: float [B][2][H][W]; scratch_v: float [B][2][H/2][W];
 * scratch_q (dither only): float [B][H][W] + [B][2][H/2][W/2].  Tables (host built, resample.chroma420_tables): vertical H -> H/2 and
 * horizontal W -> W/2 Bicubic with 'left' chroma siting, weights [out][T].  matrix: 0 = BT.709, 1 = BT.601 (470bg / 170m);
 * limited: 1 = 16-235 / 16-240, 0 = full range. */
int havc_zimg_rgb_to_yuv420p8(const uint8_t *rgb, uint8_t *y, uint8_t *uv, float *scratch444, float *scratch_v, float *scratch_q, int B,
                              int H, int W, const int *start_v, const float *w_v, int Tv, const int *start_h, const float *w_h, int Th,
                              int matrix, int limited, int dither, void *stream);
/* The std.Expr hue / saturation rotation of (U, V) (vsfilters.py:797-826; c1 = cos(hue) * sat, c2 = sin(hue) * sat; do_uv = 0
 * skips it) and the std.Lut brightness / contrast table on Y (:828-845; lut = 256 device bytes or NULL), in place. */
int havc_zimg_tweak_yuv(uint8_t *y, uint8_t *uv, int B, int H, int W, float c1, float c2, int do_uv, const uint8_t *lut, void *stream);
/* clip.resize.Bicubic(format=RGB24, matrix_in=..., range_in_s=..., range_s="full", dither_type="error_diffusion"): the way back of
 * vs_tweak (vsfilters.py:848) and convert_format_RGB24 for YUV clips (havc_utils.py:133-143); dither = 1 runs zimg's Floyd-Steinberg
 * error diffusion per plane (a warp-level wavefront), 0 rounds.  scratch_h: float [B][2][H/2][W]; scratch_rgb: float [B][3][H][W]
 * (dither only). */
int havc_zimg_yuv420p8_to_rgb(const uint8_t *y, const uint8_t *uv, uint8_t *rgb, float *scratch_h, float *scratch_rgb, int B, int H, int W,
                              const int *start_h, const float *w_h, int Th, const int *start_v, const float *w_v, int Tv, int matrix,
                              int limited, int dither, void *stream);
/* convert_format_RGB24 for GRAY8 clips (havc_utils.py:145-151): R = G = B = the luma expanded to full range, no dither. */
int havc_zimg_gray8_to_rgb(const uint8_t *y, uint8_t *rgb, int B, int H, int W, int limited, void *stream);
/* The BT.709 RGB -> YCbCr matrix (inverse = 0) or its inverse (1) as 9 floats, row major (host memory). */
int havc_zimg_inverse_matrix(float *out9, int inverse);

#ifdef __cplusplus
}
#endif
#endif /* HAVC_B200_H */
