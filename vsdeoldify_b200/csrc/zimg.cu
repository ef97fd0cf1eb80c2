// zimg.cu — the VapourSynth / zimg steps the reference's vs_tweak chains (vsdeoldify/vsslib/vsfilters.py:753-850), on planar u8
// batches: RGB24 <-> YUV420P8 (BT.709, full range, Bicubic chroma resampling with "left" chroma siting), the std.Expr hue /
// saturation rotation of (U, V), the std.Lut brightness / contrast table on Y, and the Floyd-Steinberg error-diffusion dither of
// the way back.  zimg itself is not available in this environment: the arithmetic follows the published zimg algorithm as restated
// in oracle/zimg_oracle.py (PARITY UNPINNED against the real library; bit-exact against that restatement): float32 throughout,
// separate multiplies and adds in the stated order (no FMA contraction), taps accumulated in ascending order.
#include "common.cuh"
#include "pixel_math.cuh"

namespace havc {

struct Mat3 { float m[3][3]; };

// RGB u8 planes -> Y u8 (rint(y * 255)) + float Cb / Cr planes at full resolution.
struct Quant { float y_scale, y_off, c_scale, c_off; };     // float -> integer: v * scale + off (full range: 255 / 0 / 255 / 128)

__global__ void zimg_rgb_to_yuv444_kernel(const uint8_t *__restrict__ rgb, uint8_t *__restrict__ y, float *__restrict__ yf,
                                          float *__restrict__ cbcr, int B, long long n, Mat3 fwd, Quant qn) {
    const long long total = (long long)B * n;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / n, px = i - b * n;
        const uint8_t *q = rgb + b * 3 * n + px;
        const float r = __fmul_rn((float)__ldg(q), 1.0f / 255.0f), g = __fmul_rn((float)__ldg(q + n), 1.0f / 255.0f),
                    bl = __fmul_rn((float)__ldg(q + 2 * n), 1.0f / 255.0f);
        float p[3];
#pragma unroll
        for (int k = 0; k < 3; ++k)
            p[k] = __fadd_rn(__fadd_rn(__fmul_rn(fwd.m[k][0], r), __fmul_rn(fwd.m[k][1], g)), __fmul_rn(fwd.m[k][2], bl));
        const float ys = __fadd_rn(__fmul_rn(p[0], qn.y_scale), qn.y_off);
        if (yf) yf[i] = ys;                                   // quantised by the error-diffusion pass
        else y[i] = (uint8_t)sat8(__float2int_rn(ys));
        cbcr[(b * 2 + 0) * n + px] = p[1];
        cbcr[(b * 2 + 1) * n + px] = p[2];
    }
}

// Generic separable pass on float planes: vertical (axis 0) or horizontal (axis 1); out = sum_t w[o][t] * in[start[o] + t].
//   mode 0: float out;  mode 1: u8 out = rint(clamp(acc * c_scale + c_off)) (chroma quantisation);  mode 2: the same, kept as float
template <bool kVertical>
__global__ void zimg_resample_kernel(const float *__restrict__ in, void *__restrict__ out, long long planes, int Hin, int Win,
                                     int Hout, int Wout, const int *__restrict__ start, const float *__restrict__ wts, int T,
                                     int mode, float c_scale, float c_off) {
    const long long total = planes * Hout * Wout;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ox = (int)(i % Wout);
        const int oy = (int)((i / Wout) % Hout);
        const long long pl = i / ((long long)Wout * Hout);
        const int o = kVertical ? oy : ox;
        const int s0 = __ldg(start + o);
        const int nsrc = kVertical ? Hin : Win;
        const float *w = wts + (long long)o * T;
        const float *src = in + pl * Hin * (long long)Win;
        float acc = 0.f;
        for (int t = 0; t < T; ++t) {
            const int idx = min(s0 + t, nsrc - 1);
            const float v = kVertical ? __ldg(src + (long long)idx * Win + ox) : __ldg(src + (long long)oy * Win + idx);
            acc = __fadd_rn(acc, __fmul_rn(v, __ldg(w + t)));
        }
        if (mode == 0) {
            reinterpret_cast<float *>(out)[i] = acc;
        } else if (mode == 1) {
            reinterpret_cast<uint8_t *>(out)[i] = (uint8_t)sat8(__float2int_rn(__fadd_rn(__fmul_rn(acc, c_scale), c_off)));
        } else {                                              // scaled float for the error-diffusion pass
            reinterpret_cast<float *>(out)[i] = __fadd_rn(__fmul_rn(acc, c_scale), c_off);
        }
    }
}

// u8 chroma -> float ((p - 128) / 255), then the horizontal up-sampling pass.
__global__ void zimg_chroma_up_h_kernel(const uint8_t *__restrict__ in, float *__restrict__ out, long long planes, int Hc, int Wc,
                                        int Wout, const int *__restrict__ start, const float *__restrict__ wts, int T, float c_off,
                                        float c_inv) {
    const long long total = planes * Hc * Wout;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ox = (int)(i % Wout);
        const int oy = (int)((i / Wout) % Hc);
        const long long pl = i / ((long long)Wout * Hc);
        const int s0 = __ldg(start + ox);
        const float *w = wts + (long long)ox * T;
        const uint8_t *src = in + (pl * Hc + oy) * (long long)Wc;
        float acc = 0.f;
        for (int t = 0; t < T; ++t) {
            const int idx = min(s0 + t, Wc - 1);
            const float v = __fmul_rn(__fsub_rn((float)__ldg(src + idx), c_off), c_inv);
            acc = __fadd_rn(acc, __fmul_rn(v, __ldg(w + t)));
        }
        out[i] = acc;
    }
}

// Vertical chroma up-sampling + inverse matrix + * 255: float RGB planes (for the dither) or rounded u8.
__global__ void zimg_chroma_up_v_matrix_kernel(const uint8_t *__restrict__ y, const float *__restrict__ ch /*[B][2][Hc][W]*/,
                                               float *__restrict__ rgbf, uint8_t *__restrict__ rgb8, int B, int H, int W, int Hc,
                                               const int *__restrict__ start, const float *__restrict__ wts, int T, Mat3 inv, float y_off,
                                               float y_inv) {
    const long long n = (long long)H * W, total = (long long)B * n;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / n, px = i - b * n;
        const int oy = (int)(px / W), ox = (int)(px - (long long)oy * W);
        const int s0 = __ldg(start + oy);
        const float *w = wts + (long long)oy * T;
        float c[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const float *src = ch + ((b * 2 + k) * Hc) * (long long)W + ox;
            float acc = 0.f;
            for (int t = 0; t < T; ++t) acc = __fadd_rn(acc, __fmul_rn(__ldg(src + (long long)min(s0 + t, Hc - 1) * W), __ldg(w + t)));
            c[k] = acc;
        }
        const float yf = __fmul_rn(__fsub_rn((float)__ldg(y + i), y_off), y_inv);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float p = __fadd_rn(__fadd_rn(__fmul_rn(inv.m[k][0], yf), __fmul_rn(inv.m[k][1], c[0])), __fmul_rn(inv.m[k][2], c[1]));
            const float s = __fmul_rn(p, 255.0f);
            if (rgbf) rgbf[(b * 3 + k) * n + px] = s;
            else rgb8[(b * 3 + k) * n + px] = (uint8_t)sat8(__float2int_rn(s));
        }
    }
}

// std.Expr hue / saturation rotation of the 4:2:0 chroma planes (vsfilters.py:809-826), in place:
//   U' = (U - 128) * c1 + (V - 128) * c2 + 128,  V' = (V - 128) * c1 - (U - 128) * c2 + 128, clamp [0, 255], round half even.
__global__ void zimg_uv_expr_kernel(uint8_t *__restrict__ uv, int B, long long nc, float c1, float c2) {
    const long long total = (long long)B * nc;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / nc, px = i - b * nc;
        uint8_t *pu = uv + (b * 2) * nc + px, *pv = pu + nc;
        const float u = __fsub_rn((float)*pu, 128.0f), v = __fsub_rn((float)*pv, 128.0f);
        const float nu = __fadd_rn(__fadd_rn(__fmul_rn(u, c1), __fmul_rn(v, c2)), 128.0f);
        const float nv = __fadd_rn(__fsub_rn(__fmul_rn(v, c1), __fmul_rn(u, c2)), 128.0f);
        *pu = (uint8_t)__float2int_rn(fminf(fmaxf(nu, 0.f), 255.f));
        *pv = (uint8_t)__float2int_rn(fminf(fmaxf(nv, 0.f), 255.f));
    }
}

__global__ void zimg_lut_kernel(uint8_t *__restrict__ y, long long n, const uint8_t *__restrict__ lut) {
    __shared__ uint8_t s[256];
    if (threadIdx.x < 256) s[threadIdx.x] = lut[threadIdx.x];
    __syncthreads();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) y[i] = s[y[i]];
}

// Floyd-Steinberg error diffusion (zimg depth/dither.cpp dither_ed): one block per plane, raster order.  The dependency chain is
// serial along a row and row to row; a warp runs it as a WAVEFRONT: lane l works on row r0 + l, two columns behind lane l - 1, so
// that the three errors it needs from the row above (upper-left, upper, upper-right) have just been produced; they travel down
// through __shfl_up.  The last lane's error row is parked in shared memory for lane 0 of the next band of 32 rows.
__global__ void zimg_error_diffusion_kernel(const float *__restrict__ in, uint8_t *__restrict__ out, int H, int W) {
    extern __shared__ float band_err[];          // [W + 2]: the error row of the last row of the previous band (padded by one each side)
    const int lane = threadIdx.x;
    const float *src = in + (long long)blockIdx.x * H * W;
    uint8_t *dst = out + (long long)blockIdx.x * H * W;
    for (int i = lane; i < W + 2; i += 32) band_err[i] = 0.f;
    __syncwarp();
    for (int r0 = 0; r0 < H; r0 += 32) {
        const int row = r0 + lane;
        const bool live = row < H;
        // history of MY row's errors at columns j-1 (left), and the upper row's errors as seen through the lane above
        float e_left = 0.f;                       // my error at column j - 1
        float mine_m0 = 0.f, mine_m1 = 0.f, mine_m2 = 0.f;   // my errors at my columns jj, jj-1, jj-2 (after the step)
        const int steps = W + 2 * 31 + 2;
        for (int s = 0; s < steps; ++s) {
            const int j = s - 2 * lane;           // my column at this step
            // the lane above is at column j + 2 this step; BEFORE this step it had finished columns <= j + 1:
            // its history holds errors at (j + 1, j, j - 1) = (mine_m0, mine_m1, mine_m2) of that lane
            float up_r = __shfl_up_sync(0xffffffffu, mine_m0, 1);
            float up_c = __shfl_up_sync(0xffffffffu, mine_m1, 1);
            float up_l = __shfl_up_sync(0xffffffffu, mine_m2, 1);
            if (lane == 0) {                      // row above = last row of the previous band (zeros for the first band)
                const bool in_rng = j >= 0 && j < W;
                up_l = in_rng ? band_err[j] : 0.f;        // padded index: column j - 1 -> j
                up_c = in_rng ? band_err[j + 1] : 0.f;
                up_r = in_rng ? band_err[j + 2] : 0.f;
            }
            float e_new = 0.f;
            const bool act = live && j >= 0 && j < W;
            if (act) {
                if (j + 1 >= W) up_r = 0.f;       // padded entry right of the row above
                if (j == 0) up_l = 0.f;
                float err = 0.f;
                err = __fadd_rn(err, __fmul_rn(e_left, 7.0f / 16.0f));
                err = __fadd_rn(err, __fmul_rn(up_r, 3.0f / 16.0f));
                err = __fadd_rn(err, __fmul_rn(up_c, 5.0f / 16.0f));
                err = __fadd_rn(err, __fmul_rn(up_l, 1.0f / 16.0f));
                float v = __fadd_rn(__ldg(src + (long long)row * W + j), err);
                v = fminf(fmaxf(v, 0.f), 255.f);
                const int q = __float2int_rn(v);
                dst[(long long)row * W + j] = (uint8_t)q;
                e_new = __fsub_rn(v, (float)q);
                e_left = e_new;
            }
            // shift my history: after this step my newest finished column is j
            if (j >= 0) { mine_m2 = mine_m1; mine_m1 = mine_m0; mine_m0 = (j < W) ? e_new : 0.f; }
            if (lane == 31 && act) band_err[W + 2 + j + 1] = e_new;     // staged behind the live row (second half), swapped below
        }
        __syncwarp();
        // the last row of this band becomes the "row above" of the next band
        for (int i = lane; i < W + 2; i += 32) band_err[i] = (i >= 1 && i <= W && r0 + 31 < H) ? band_err[W + 2 + i] : 0.f;
        __syncwarp();
    }
}

// GRAY8 -> RGB24 planes: R = G = B = rint((y - y_off) * y_inv * 255) (resize.Bicubic(format=RGB24, range_in_s="limited", range_s="full"),
// havc_utils.py:145-151: no dither).
__global__ void zimg_gray_to_rgb_kernel(const uint8_t *__restrict__ y, uint8_t *__restrict__ rgb, int B, long long n, float y_off, float y_inv) {
    const long long total = (long long)B * n;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / n, px = i - b * n;
        const float v = __fmul_rn(__fmul_rn(__fsub_rn((float)__ldg(y + i), y_off), y_inv), 255.0f);
        const uint8_t q = (uint8_t)sat8(__float2int_rn(v));
        rgb[(b * 3) * n + px] = q; rgb[(b * 3 + 1) * n + px] = q; rgb[(b * 3 + 2) * n + px] = q;
    }
}

}  // namespace havc

using namespace havc;

// matrix: 0 = BT.709, 1 = BT.601 (470bg / 170m)
static Mat3 ycbcr_matrix(bool inverse, int matrix) {
    const double kr = matrix == 1 ? 0.299 : 0.2126, kb = matrix == 1 ? 0.114 : 0.0722, kg = 1.0 - kr - kb;
    double m[3][3] = {{kr, kg, kb}, {-kr / (2 * (1 - kb)), -kg / (2 * (1 - kb)), 0.5}, {0.5, -kg / (2 * (1 - kr)), -kb / (2 * (1 - kr))}};
    Mat3 o;
    if (!inverse) {
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) o.m[i][j] = (float)m[i][j];
        return o;
    }
    // inverse in closed form (R = Y + 2(1-Kr) Cr, B = Y + 2(1-Kb) Cb, G from the luma equation), evaluated in double
    const double inv[3][3] = {{1.0, 0.0, 2 * (1 - kr)}, {1.0, -2 * kb * (1 - kb) / kg, -2 * kr * (1 - kr) / kg}, {1.0, 2 * (1 - kb), 0.0}};
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) o.m[i][j] = (float)inv[i][j];
    return o;
}

static Quant quant_of(int limited) {
    Quant q;
    q.y_scale = limited ? 219.f : 255.f; q.y_off = limited ? 16.f : 0.f; q.c_scale = limited ? 224.f : 255.f; q.c_off = 128.f;
    return q;
}

extern "C" int havc_zimg_inverse_matrix(float *out9, int inverse) {
    HAVC_CHECK_ARG(out9 != nullptr, "havc_zimg_inverse_matrix: null output");
    const Mat3 m = ycbcr_matrix(inverse != 0, 0);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) out9[3 * i + j] = m.m[i][j];
    return HAVC_OK;
}

extern "C" int havc_zimg_rgb_to_yuv420p8(const uint8_t *rgb, uint8_t *y, uint8_t *uv, float *scratch444, float *scratch_v, float *scratch_q,
                                         int B, int H, int W, const int *start_v, const float *w_v, int Tv, const int *start_h,
                                         const float *w_h, int Th, int matrix, int limited, int dither, void *stream) {
    HAVC_CHECK_ARG(rgb && y && uv && scratch444 && scratch_v && (scratch_q || !dither) && B > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0 &&
                       start_v && w_v && start_h && w_h && Tv > 0 && Th > 0 && (matrix == 0 || matrix == 1) &&
                       (W + 2) * 2 * sizeof(float) <= 48 * 1024,
                   "havc_zimg_rgb_to_yuv420p8: bad arguments (4:2:0 needs even width and height; matrix 0 = 709, 1 = 601)");
    cudaStream_t st = (cudaStream_t)stream;
    const long long n = (long long)H * W, nc = (long long)(H / 2) * (W / 2);
    const Quant qn = quant_of(limited);
    float *yf = dither ? scratch_q : nullptr;                     // [B][H][W], then [B][2][H/2][W/2]
    float *cf = dither ? scratch_q + (long long)B * n : nullptr;
    zimg_rgb_to_yuv444_kernel<<<grid1d((long long)B * n, 256), 256, 0, st>>>(rgb, y, yf, scratch444, B, n, ycbcr_matrix(false, matrix), qn);
    HAVC_LAUNCHED();
    // vertical pass first (zimg orders the passes by cost; for a 2:1 reduction in both directions that is the vertical one)
    zimg_resample_kernel<true><<<grid1d((long long)B * 2 * (H / 2) * W, 256), 256, 0, st>>>(scratch444, scratch_v, (long long)B * 2, H, W, H / 2, W,
                                                                                          start_v, w_v, Tv, 0, 0.f, 0.f);
    HAVC_LAUNCHED();
    zimg_resample_kernel<false><<<grid1d((long long)B * 2 * nc, 256), 256, 0, st>>>(scratch_v, dither ? (void *)cf : (void *)uv, (long long)B * 2, H / 2,
                                                                                  W, H / 2, W / 2, start_h, w_h, Th, dither ? 2 : 1, qn.c_scale,
                                                                                  qn.c_off);
    HAVC_LAUNCHED();
    if (dither) {
        zimg_error_diffusion_kernel<<<B, 32, (size_t)(W + 2) * 2 * sizeof(float), st>>>(yf, y, H, W);
        HAVC_LAUNCHED();
        zimg_error_diffusion_kernel<<<B * 2, 32, (size_t)(W / 2 + 2) * 2 * sizeof(float), st>>>(cf, uv, H / 2, W / 2);
        HAVC_LAUNCHED();
    }
    return HAVC_OK;
}

extern "C" int havc_zimg_tweak_yuv(uint8_t *y, uint8_t *uv, int B, int H, int W, float c1, float c2, int do_uv, const uint8_t *lut,
                                   void *stream) {
    HAVC_CHECK_ARG(y && uv && B > 0 && H > 0 && W > 0, "havc_zimg_tweak_yuv: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    if (do_uv) {
        const long long nc = (long long)(H / 2) * (W / 2);
        zimg_uv_expr_kernel<<<grid1d((long long)B * nc, 256), 256, 0, st>>>(uv, B, nc, c1, c2);
        HAVC_LAUNCHED();
    }
    if (lut != nullptr) {
        zimg_lut_kernel<<<grid1d((long long)B * H * W, 256), 256, 0, st>>>(y, (long long)B * H * W, lut);
        HAVC_LAUNCHED();
    }
    return HAVC_OK;
}

extern "C" int havc_zimg_yuv420p8_to_rgb(const uint8_t *y, const uint8_t *uv, uint8_t *rgb, float *scratch_h, float *scratch_rgb, int B,
                                         int H, int W, const int *start_h, const float *w_h, int Th, const int *start_v, const float *w_v,
                                         int Tv, int matrix, int limited, int dither, void *stream) {
    HAVC_CHECK_ARG(y && uv && rgb && scratch_h && (scratch_rgb || !dither) && B > 0 && H % 2 == 0 && W % 2 == 0 && H > 0 && W > 0 &&
                       (matrix == 0 || matrix == 1) && (W + 2) * 2 * sizeof(float) <= 48 * 1024,
                   "havc_zimg_yuv420p8_to_rgb: bad arguments (even sizes, width <= 6142; matrix 0 = 709, 1 = 601)");
    cudaStream_t st = (cudaStream_t)stream;
    const Quant qn = quant_of(limited);
    zimg_chroma_up_h_kernel<<<grid1d((long long)B * 2 * (H / 2) * W, 256), 256, 0, st>>>(uv, scratch_h, (long long)B * 2, H / 2, W / 2, W, start_h,
                                                                                       w_h, Th, qn.c_off, 1.0f / qn.c_scale);
    HAVC_LAUNCHED();
    zimg_chroma_up_v_matrix_kernel<<<grid1d((long long)B * H * W, 256), 256, 0, st>>>(y, scratch_h, dither ? scratch_rgb : nullptr, rgb, B, H, W,
                                                                                    H / 2, start_v, w_v, Tv, ycbcr_matrix(true, matrix),
                                                                                    qn.y_off, 1.0f / qn.y_scale);
    HAVC_LAUNCHED();
    if (dither) {
        zimg_error_diffusion_kernel<<<B * 3, 32, (size_t)(W + 2) * 2 * sizeof(float), st>>>(scratch_rgb, rgb, H, W);
        HAVC_LAUNCHED();
    }
    return HAVC_OK;
}

/* GRAY8 (limited or full range) -> RGB24 planes, no dither (havc_utils.py:145-151). */
extern "C" int havc_zimg_gray8_to_rgb(const uint8_t *y, uint8_t *rgb, int B, int H, int W, int limited, void *stream) {
    HAVC_CHECK_ARG(y && rgb && B > 0 && H > 0 && W > 0, "havc_zimg_gray8_to_rgb: bad arguments");
    const Quant qn = quant_of(limited);
    zimg_gray_to_rgb_kernel<<<grid1d((long long)B * H * W, 256), 256, 0, (cudaStream_t)stream>>>(y, rgb, B, (long long)H * W, qn.y_off,
                                                                                               1.0f / qn.y_scale);
    HAVC_LAUNCHED();
    return HAVC_OK;
}
