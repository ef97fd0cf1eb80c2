// pixel_math.cuh — device-side integer / float32 colour arithmetic shared by the pixel passes (pixel.cu) and the
// merge / chroma-adjust filters (filters.cu).  Every formula is the bit-exact model of the library call the
// reference makes (OpenCV 4.x 8-bit cvtColor, Pillow Blend.c / convert('L')), pinned by tests/test_pixel_oracle.py
// and tests/test_filters_oracle.py.  Floating-point steps use explicit _rn intrinsics so that nvcc never contracts
// a multiply-add the CPU library does not contract (and fuses exactly where the library's compiled code does).
#pragma once
#include "common.cuh"

namespace havc {

static inline int grid1d(long long n, int block) {
    long long g = (n + block - 1) / block;
    long long cap = (long long)num_sms() * 32;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

// clamp to [0, 255]: one VIMNMX with the relu modifier (max(min(v, 255), 0))
__device__ __forceinline__ int sat8(int v) { return __vimin_s32_relu(v, 255); }

// OpenCV 8-bit COLOR_RGB2YUV / COLOR_YUV2RGB, Q14 fixed point (pinned against
// cv2 4.13 by tests/test_pixel_oracle.py).
__device__ __forceinline__ void rgb2yuv(int r, int g, int b, int &y, int &u, int &v) {
    y = (4899 * r + 9617 * g + 1868 * b + 8192) >> 14;
    u = sat8(((b - y) * 8061 + (128 << 14) + 8192) >> 14);
    v = sat8(((r - y) * 14369 + (128 << 14) + 8192) >> 14);
}
__device__ __forceinline__ void yuv2rgb(int y, int u, int v, int &r, int &g, int &b) {
    r = sat8(y + (((v - 128) * 18678 + 8192) >> 14));
    g = sat8(y + (((u - 128) * -6472 + (v - 128) * -9519 + 8192) >> 14));
    b = sat8(y + (((u - 128) * 33292 + 8192) >> 14));
}
// Keep the luma of `o` and the chroma of `c` (ColorizerFilter._post_process, filters.py:100-110).
__device__ __forceinline__ void luma_transplant(int orr, int og, int ob, int cr, int cg, int cb, int &r, int &g,
                                                int &b) {
    int y, u, v, y2, u2, v2;
    rgb2yuv(orr, og, ob, y, u, v);
    rgb2yuv(cr, cg, cb, y2, u2, v2);
    yuv2rgb(y, u2, v2, r, g, b);
}

// Pillow convert('L'): (19595 R + 38470 G + 7471 B + 0x8000) >> 16.
__device__ __forceinline__ int pil_luma(int r, int g, int b) { return (19595 * r + 38470 * g + 7471 * b + 0x8000) >> 16; }

// PIL.Image.blend(a, b, alpha): float32 a + alpha*(b-a) (separate multiply and add); alpha in [0,1] truncates,
// outside it the result is clipped to [0,255] first (libImaging/Blend.c).
__device__ __forceinline__ int pil_blend(int a, int b, float alpha) {
    const float t = __fadd_rn((float)a, __fmul_rn(alpha, __fsub_rn((float)b, (float)a)));
    if (alpha >= 0.f && alpha <= 1.f) return sat8((int)t);
    return t <= 0.f ? 0 : (t >= 255.f ? 255 : (int)t);
}

// cv2 COLOR_RGB2HSV on 8-bit data (H in [0,179]): integer division tables, 12-bit fixed point.
__device__ __forceinline__ int cv_div_table(int num, int i) {   // round-half-even of (num << 12) / i, i in [1,255]
    return __double2int_rn((double)(num << 12) / (double)i);
}
__device__ __forceinline__ void rgb2hsv(int r, int g, int b, int &h, int &s, int &v) {
    v = max(max(r, g), b);
    const int vmin = min(min(r, g), b);
    const int d = v - vmin;
    s = v == 0 ? 0 : (d * cv_div_table(255, v) + 2048) >> 12;
    int hh = (v == r) ? (g - b) : ((v == g) ? (b - r + 2 * d) : (r - g + 4 * d));
    hh = d == 0 ? 0 : (hh * __double2int_rn((double)(180 << 12) / (6.0 * (double)d)) + 2048) >> 12;
    h = hh < 0 ? hh + 180 : hh;
}
// cv2 COLOR_HSV2RGB on 8-bit data: float32 sector model; 1 - s*f and 1 - s*(1-f) are fused (as in the compiled
// library), x*255 is truncated inside OpenCV's per-row SIMD blocks and rounded half-to-even in the scalar row tail.
__device__ __forceinline__ void hsv2rgb(int h, int s, int v, bool simd_body, int &r, int &g, int &b) {
    const float hf = __fmul_rn((float)h, 6.0f / 180.0f);
    const float sf = __fmul_rn((float)s, 1.0f / 255.0f), vf = __fmul_rn((float)v, 1.0f / 255.0f);
    const float fl = floorf(hf);
    const float fr = __fsub_rn(hf, fl);
    int sector = (int)fl % 6;
    float tab[4];
    tab[0] = vf;
    tab[1] = __fmul_rn(vf, __fsub_rn(1.0f, sf));
    tab[2] = __fmul_rn(vf, __fmaf_rn(-sf, fr, 1.0f));
    tab[3] = __fmul_rn(vf, __fmaf_rn(-sf, __fsub_rn(1.0f, fr), 1.0f));
    // (b, g, r) table index per sector
    int kb, kg, kr;
    switch (sector) {
        case 0: kb = 1; kg = 3; kr = 0; break;
        case 1: kb = 1; kg = 0; kr = 2; break;
        case 2: kb = 3; kg = 0; kr = 1; break;
        case 3: kb = 0; kg = 2; kr = 1; break;
        case 4: kb = 0; kg = 1; kr = 3; break;
        default: kb = 2; kg = 1; kr = 0; break;
    }
    const float xb = __fmul_rn(tab[kb], 255.0f), xg = __fmul_rn(tab[kg], 255.0f), xr = __fmul_rn(tab[kr], 255.0f);
    if (simd_body) {
        r = sat8((int)floorf(xr)); g = sat8((int)floorf(xg)); b = sat8((int)floorf(xb));
    } else {
        r = sat8(__float2int_rn(xr)); g = sat8(__float2int_rn(xg)); b = sat8(__float2int_rn(xb));
    }
}

}  // namespace havc
