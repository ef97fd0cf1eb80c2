// zhang.cu — pre/post pixel passes around the Zhang et al. colorizers (eccv16 / siggraph17) and the Pillow-exact
// integer resampler they (and BaseFilter._scale_to_square / _unsquare) use.
//
// Reference: vsdeoldify/colorization/__init__.py:76-95 (ModelColorization.colorize_frame),
// colorizers/util.py:21-55 (Pillow BICUBIC resize to 256 x 256, skimage rgb2lab / lab2rgb, bilinear F.interpolate of ab),
// colorizers/eccv16.py:82-98 (softmax over 313 classes -> 1x1 conv 313->2 -> x4 bilinear -> *110),
// colorizers/siggraph17.py:123-161 (1x1 conv 128->2 -> tanh -> *110), deoldify/filters.py:37-41,70-73 (Pillow BILINEAR).
// LAB runs in float64 like scikit-image; the L handed to the network is narrowed to float32 (torch.Tensor(img_l)).
#include "pixel_math.cuh"

namespace havc {

// ---- Pillow ImagingResample, 8 bits per channel: 22-bit fixed-point coefficients, rounding add, clip (Resample.c) ----
__global__ void pil_resample_u8_kernel(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, long long planes, int Hin, int Win,
                                       int out_size, int horizontal, const int *__restrict__ bounds, const int *__restrict__ coeffs,
                                       int ksize) {
    const int Hout = horizontal ? Hin : out_size, Wout = horizontal ? out_size : Win;
    const long long total = planes * Hout * Wout;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % Wout);
        const long long t = i / Wout;
        const int y = (int)(t % Hout);
        const long long p = t / Hout;
        const int o = horizontal ? x : y;
        const int xmin = __ldg(bounds + 2 * o), cnt = __ldg(bounds + 2 * o + 1);
        const int *k = coeffs + (long long)o * ksize;
        const uint8_t *src = in + p * Hin * Win;
        int acc = 1 << 21;
        if (horizontal) {
            const uint8_t *row = src + (long long)y * Win + xmin;
            for (int j = 0; j < cnt; ++j) acc += (int)__ldg(row + j) * __ldg(k + j);
        } else {
            const uint8_t *col = src + (long long)xmin * Win + x;
            for (int j = 0; j < cnt; ++j) acc += (int)__ldg(col + (long long)j * Win) * __ldg(k + j);
        }
        out[i] = (uint8_t)sat8(acc >> 22);
    }
}

// ---- skimage.color.rgb2lab, L channel only (float64), narrowed to float32; optional network input (L-50)/100 ---------
__device__ __forceinline__ double srgb_to_linear(int c8) {
    const double c = (double)c8 / 255.0;
    return c > 0.04045 ? pow((c + 0.055) / 1.055, 2.4) : c / 12.92;
}
__device__ __forceinline__ double lab_f(double t) { return t > 0.008856 ? cbrt(t) : 7.787 * t + 16.0 / 116.0; }

__global__ void zhang_pre_kernel(const uint8_t *__restrict__ rgb, int B, long long n, float *__restrict__ L_out, void *__restrict__ x,
                                 int dtype) {
    const long long total = (long long)B * n;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / n, px = i - b * n;
        const uint8_t *q = rgb + b * 3 * n + px;
        const double r = srgb_to_linear(__ldg(q)), g = srgb_to_linear(__ldg(q + n)), bl = srgb_to_linear(__ldg(q + 2 * n));
        // Y row of the sRGB -> XYZ matrix, Y white = 1 (skimage xyz_from_rgb / D65 2-degree observer)
        const double Y = __dadd_rn(__dadd_rn(__dmul_rn(r, 0.212671), __dmul_rn(g, 0.715160)), __dmul_rn(bl, 0.072169));
        const float L = (float)(116.0 * lab_f(Y) - 16.0);
        if (L_out) L_out[i] = L;
        if (x) {        // BaseColor.normalize_l in float32 (base_color.py:13-14); NHWC with 8-channel storage, channel 0
            const float v = __fdiv_rn(__fsub_rn(L, 50.0f), 100.0f);
            store16(x, i * 8, v, dtype);
            // channel 4: the remainder of the 16-bit rounding (lo plane of the split-precision first convolution)
            const float hi = load16(x, i * 8, dtype);
            store16(x, i * 8 + 4, v - hi, dtype);
        }
    }
}

// ---- eccv16 tail: softmax over the class logits, then the 313 -> 2 1x1 conv (no bias); one warp per pixel ------------
__global__ void eccv_head_kernel(const float *__restrict__ logits, int ld, int n_classes, const float *__restrict__ w_out,
                                 float *__restrict__ out, long long pixels) {
    const int lane = threadIdx.x & 31;
    const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long p = warp; p < pixels; p += n_warps) {
        const float *row = logits + p * ld;
        float m = -INFINITY;
        for (int c = lane; c < n_classes; c += 32) m = fmaxf(m, __ldg(row + c));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        float s = 0.f, a = 0.f, b = 0.f;
        for (int c = lane; c < n_classes; c += 32) {
            const float e = expf(__ldg(row + c) - m);
            s += e;
            a = fmaf(e, __ldg(w_out + c), a);
            b = fmaf(e, __ldg(w_out + n_classes + c), b);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, o);
            a += __shfl_xor_sync(0xffffffffu, a, o);
            b += __shfl_xor_sync(0xffffffffu, b, o);
        }
        if (lane == 0) {
            out[2 * p] = a / s;
            out[2 * p + 1] = b / s;
        }
    }
}

// ---- siggraph17 tail: tanh(head + bias) * 110 on the fused-head output of the last GEMM ([pix][4] fp32) ---------------
__global__ void zhang_tanh_kernel(const float4 *__restrict__ head, const float *__restrict__ bias, float2 *__restrict__ out,
                                  long long pixels, float mul) {
    const float b0 = bias[0], b1 = bias[1];
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < pixels; i += (long long)gridDim.x * blockDim.x) {
        const float4 h = __ldg(head + i);
        out[i] = make_float2(tanhf(h.x + b0) * mul, tanhf(h.y + b1) * mul);
    }
}

// ---- torch upsample_bilinear2d, align_corners=False, 2 channels interleaved -----------------------------------------
struct BilinearAxis { int i0, i1; float l0, l1; };
__device__ __forceinline__ BilinearAxis bilinear_axis(int dst, float scale, int in_size) {
    float src = fmaf(scale, (float)dst + 0.5f, -0.5f);
    src = src < 0.f ? 0.f : src;
    BilinearAxis a;
    a.i0 = min((int)src, in_size - 1);
    a.i1 = a.i0 + (a.i0 < in_size - 1 ? 1 : 0);
    a.l1 = src - (float)a.i0;
    a.l0 = 1.0f - a.l1;
    return a;
}
__device__ __forceinline__ float2 bilinear_ab(const float2 *__restrict__ img, int h, int w, int oy, int ox, float sh, float sw) {
    const BilinearAxis ay = bilinear_axis(oy, sh, h), ax = bilinear_axis(ox, sw, w);
    const float2 v00 = __ldg(img + (long long)ay.i0 * w + ax.i0), v01 = __ldg(img + (long long)ay.i0 * w + ax.i1);
    const float2 v10 = __ldg(img + (long long)ay.i1 * w + ax.i0), v11 = __ldg(img + (long long)ay.i1 * w + ax.i1);
    float2 r;
    r.x = ay.l0 * (ax.l0 * v00.x + ax.l1 * v01.x) + ay.l1 * (ax.l0 * v10.x + ax.l1 * v11.x);
    r.y = ay.l0 * (ax.l0 * v00.y + ax.l1 * v01.y) + ay.l1 * (ax.l0 * v10.y + ax.l1 * v11.y);
    return r;
}
__global__ void bilinear_ab_kernel(const float2 *__restrict__ in, float2 *__restrict__ out, int B, int h, int w, int H, int W, float mul) {
    const long long total = (long long)B * H * W;
    const float sh = (float)h / (float)H, sw = (float)w / (float)W;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ox = (int)(i % W);
        const long long t = i / W;
        const int oy = (int)(t % H);
        const long long b = t / H;
        const float2 v = bilinear_ab(in + b * h * w, h, w, oy, ox, sh, sw);
        out[i] = make_float2(v.x * mul, v.y * mul);
    }
}

// ---- postprocess_tens + uint8 conversion: ab resized to the frame, cat with L, skimage lab2rgb (float64), trunc -------
__device__ __forceinline__ double lab_finv(double f) { return f > 0.2068966 ? f * f * f : (f - 16.0 / 116.0) / 7.787; }
__device__ __forceinline__ int linear_to_srgb8(double c) {
    double v = c > 0.0031308 ? 1.055 * pow(c < 0.0 ? 0.0 : c, 1.0 / 2.4) - 0.055 : 12.92 * c;
    v = v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v);
    v = v * 255.0;
    return (int)(v < 0.0 ? 0.0 : (v > 255.0 ? 255.0 : v));      // np.uint8(np.clip(x*255, 0, 255)): truncation
}
struct XyzToRgb { double m[9]; };
__global__ void zhang_post_kernel(const float2 *__restrict__ ab, int h, int w, const float *__restrict__ L, uint8_t *__restrict__ out,
                                  int B, int H, int W, XyzToRgb M) {
    const long long n = (long long)H * W, total = (long long)B * n;
    const bool resize = (h != H) || (w != W);
    const float sh = (float)h / (float)H, sw = (float)w / (float)W;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / n, px = i - b * n;
        const int oy = (int)(px / W), ox = (int)(px - (long long)oy * W);
        const float2 v = resize ? bilinear_ab(ab + b * h * w, h, w, oy, ox, sh, sw) : __ldg(ab + b * h * w + px);
        const double fy = ((double)__ldg(L + i) + 16.0) / 116.0;
        const double fx = (double)v.x / 500.0 + fy;
        double fz = fy - (double)v.y / 200.0;
        fz = fz < 0.0 ? 0.0 : fz;
        const double X = lab_finv(fx) * 0.95047, Y = lab_finv(fy), Z = lab_finv(fz) * 1.08883;
        const double r = M.m[0] * X + M.m[1] * Y + M.m[2] * Z;
        const double g = M.m[3] * X + M.m[4] * Y + M.m[5] * Z;
        const double bl = M.m[6] * X + M.m[7] * Y + M.m[8] * Z;
        uint8_t *q = out + b * 3 * n + px;
        q[0] = (uint8_t)linear_to_srgb8(r);
        q[n] = (uint8_t)linear_to_srgb8(g);
        q[2 * n] = (uint8_t)linear_to_srgb8(bl);
    }
}

static int grid_for(long long n, int block = 256) {
    long long g = (n + block - 1) / block;
    const long long cap = (long long)num_sms() * 32;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace havc

using namespace havc;

extern "C" int havc_pil_resample_u8(const uint8_t *in, uint8_t *out, long long planes, int Hin, int Win, int out_size, int horizontal,
                                    const int *bounds, const int *coeffs, int ksize, void *stream) {
    HAVC_CHECK_ARG(in && out && bounds && coeffs && planes > 0 && Hin > 0 && Win > 0 && out_size > 0 && ksize > 0,
                   "havc_pil_resample_u8: bad arguments");
    const long long total = planes * (horizontal ? (long long)Hin * out_size : (long long)out_size * Win);
    pil_resample_u8_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(in, out, planes, Hin, Win, out_size, horizontal, bounds,
                                                                           coeffs, ksize);
    HAVC_LAUNCHED();
    return HAVC_OK;
}

extern "C" int havc_zhang_pre(const uint8_t *rgb, int B, long long n_pixels, float *L_out, void *x, int dtype, void *stream) {
    HAVC_CHECK_ARG(rgb && B > 0 && n_pixels > 0 && (L_out || x) && (dtype == HAVC_F16 || dtype == HAVC_BF16), "havc_zhang_pre: bad arguments");
    zhang_pre_kernel<<<grid_for((long long)B * n_pixels), 256, 0, (cudaStream_t)stream>>>(rgb, B, n_pixels, L_out, x, dtype);
    HAVC_LAUNCHED();
    return HAVC_OK;
}

extern "C" int havc_eccv_head(const float *logits, int ld, int n_classes, const float *w_out, float *out, long long pixels, void *stream) {
    HAVC_CHECK_ARG(logits && w_out && out && pixels > 0 && n_classes > 0 && ld >= n_classes, "havc_eccv_head: bad arguments");
    eccv_head_kernel<<<grid_for(pixels * 32), 256, 0, (cudaStream_t)stream>>>(logits, ld, n_classes, w_out, out, pixels);
    HAVC_LAUNCHED();
    return HAVC_OK;
}

extern "C" int havc_zhang_tanh(const float *head, const float *bias, float *out, long long pixels, float mul, void *stream) {
    HAVC_CHECK_ARG(head && bias && out && pixels > 0, "havc_zhang_tanh: bad arguments");
    zhang_tanh_kernel<<<grid_for(pixels), 256, 0, (cudaStream_t)stream>>>((const float4 *)head, bias, (float2 *)out, pixels, mul);
    HAVC_LAUNCHED();
    return HAVC_OK;
}

extern "C" int havc_bilinear_ab(const float *in, float *out, int B, int h, int w, int H, int W, float mul, void *stream) {
    HAVC_CHECK_ARG(in && out && B > 0 && h > 0 && w > 0 && H > 0 && W > 0, "havc_bilinear_ab: bad arguments");
    bilinear_ab_kernel<<<grid_for((long long)B * H * W), 256, 0, (cudaStream_t)stream>>>((const float2 *)in, (float2 *)out, B, h, w, H, W, mul);
    HAVC_LAUNCHED();
    return HAVC_OK;
}

extern "C" int havc_zhang_post(const float *ab, int h, int w, const float *L, uint8_t *out, int B, int H, int W, void *stream) {
    HAVC_CHECK_ARG(ab && L && out && B > 0 && h > 0 && w > 0 && H > 0 && W > 0, "havc_zhang_post: bad arguments");
    // inverse of skimage's xyz_from_rgb, computed once in double (Gauss-Jordan on the 3x3)
    static XyzToRgb M;
    static std::atomic<unsigned long long> have{0ull};
    unsigned long long have_bit;
    if (device_pending(have, &have_bit)) {
        const double a[9] = {0.412453, 0.357580, 0.180423, 0.212671, 0.715160, 0.072169, 0.019334, 0.119193, 0.950227};
        const double det = a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) + a[2] * (a[3] * a[7] - a[4] * a[6]);
        M.m[0] = (a[4] * a[8] - a[5] * a[7]) / det; M.m[1] = (a[2] * a[7] - a[1] * a[8]) / det; M.m[2] = (a[1] * a[5] - a[2] * a[4]) / det;
        M.m[3] = (a[5] * a[6] - a[3] * a[8]) / det; M.m[4] = (a[0] * a[8] - a[2] * a[6]) / det; M.m[5] = (a[2] * a[3] - a[0] * a[5]) / det;
        M.m[6] = (a[3] * a[7] - a[4] * a[6]) / det; M.m[7] = (a[1] * a[6] - a[0] * a[7]) / det; M.m[8] = (a[0] * a[4] - a[1] * a[3]) / det;
        device_done(have, have_bit);
    }
    zhang_post_kernel<<<grid_for((long long)B * H * W), 256, 0, (cudaStream_t)stream>>>((const float2 *)ab, h, w, L, out, B, H, W, M);
    HAVC_LAUNCHED();
    return HAVC_OK;
}
