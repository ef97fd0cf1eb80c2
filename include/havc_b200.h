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
    int32_t BN;                 /* N tile (multiple of 16, <= 272)                              */
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
} havc_conv_desc;

const char *havc_last_error(void);
int havc_version(void);
/* Number of kernels this library has launched since load / last reset (bench.py "gpu_launches"). */
int64_t havc_launch_count(void);
void havc_launch_count_reset(void);

int havc_conv_gemm(const havc_conv_desc *d, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* HAVC_B200_H */
