// filters.cu — the vsslib model-merge and chroma-adjust filters of the HAVC hot path on planar RGB24 frames
// [B][3][H][W].  HBM-bound per-pixel passes; frame-global quantities (mean OpenCV luma,
// mean Pillow 'L') are exact integer sums produced by warp-shuffle + atomic reductions and consumed on the device, so a
// whole merge is graph-capturable with no host round trip.
//
// Reference: vsdeoldify/vsslib/mcomb.py:125-516 (merge dispatch and the per-frame selectors),
// imfilters.py:66-372,463-504 (chroma_stabilizer[_adaptive], luma merges, image_tweak, luma_adjusted_levels),
// restcolor.py:98-470 (restore_color_gradient, adjust_chroma, hue-range language), nputils.py:27-283,
// vsfilters.py:366-455,656-739.  numpy computes these in float64 and truncates to uint8; the kernels do the same
// arithmetic in the same order in double precision (explicit _rn intrinsics: no contraction numpy does not do).
#include "pixel_math.cuh"

namespace havc {

struct Img {           // planar RGB24 batch
    const uint8_t *p;
    long long plane;   // H*W
    __device__ __forceinline__ void load(int b, long long i, int &r, int &g, int &bl) const {
        const uint8_t *q = p + (long long)b * 3 * plane + i;
        r = __ldg(q); g = __ldg(q + plane); bl = __ldg(q + 2 * plane);
    }
};
__device__ __forceinline__ void store_px(uint8_t *out, long long plane, int b, long long i, int r, int g, int bl) {
    uint8_t *q = out + (long long)b * 3 * plane + i;
    q[0] = (uint8_t)r; q[plane] = (uint8_t)g; q[2 * plane] = (uint8_t)bl;
}

__device__ __forceinline__ double clip255(double v) { return v < 0.0 ? 0.0 : (v > 255.0 ? 255.0 : v); }
// float64 -> uint8 of a value already clipped to [0,255]: truncation
__device__ __forceinline__ int trunc8(double v) { return (int)v; }
// np luma: (R*0.299 + G*0.587) + B*0.114 in float64, clipped to [0,255] (nputils.py:101-113)
__device__ __forceinline__ double np_luma(int r, int g, int b) {
    return clip255(__dadd_rn(__dadd_rn(__dmul_rn((double)r, 0.299), __dmul_rn((double)g, 0.587)), __dmul_rn((double)b, 0.114)));
}
// img1*(1-w) + img2*w in float64, clip, truncate (np_weighted_merge nputils.py:265-283; `omw` = the host's 1-w)
__device__ __forceinline__ int np_wmerge(int a, int b, double omw, double w) {
    return trunc8(clip255(__dadd_rn(__dmul_rn((double)a, omw), __dmul_rn((double)b, w))));
}
// frame-mean luma rounded to 6 decimals: round(mean(Y)/255, 6) (imfilters.py:597-601)
__device__ __forceinline__ double frame_luma(unsigned long long sum_y, long long n) {
    const double mean = (double)sum_y / (double)n;
    return rint(mean / 255.0 * 1e6) / 1e6;
}

// block-wide sum -> one atomicAdd per block (all threads of a block belong to frame blockIdx.y)
__device__ __forceinline__ void block_add(unsigned long long v, unsigned long long *dst) {
    __shared__ unsigned long long warp_sums[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) warp_sums[w] = v;
    __syncthreads();
    if (w == 0) {
        v = lane < (blockDim.x >> 5) ? warp_sums[lane] : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0 && v) atomicAdd(dst, v);
    }
    __syncthreads();
}

struct HueRanges {
    int n;
    double lo[HAVC_MAX_HUE_RANGES], hi[HAVC_MAX_HUE_RANGES];   // already halved: cv hue units
    __device__ __forceinline__ bool hit(int h) const {          // _build_hue_conditions restcolor.py:412-428 (strict)
        bool c = false;
        for (int i = 0; i < n; ++i) c |= ((double)h > lo[i]) && ((double)h < hi[i]);
        return c;
    }
};

// ---- frame statistics ------------------------------------------------------------------------------------
// stats[b][0] = sum of OpenCV Y, stats[b][1] = sum of Pillow L over frame b.  `bright` != 1: statistics of
// ImageEnhance.Brightness(img).enhance(bright) instead (the input ImageEnhance.Contrast sees inside image_tweak).
__global__ void frame_stats_kernel(Img img, unsigned long long *stats, float bright, int use_bright) {
    const int b = blockIdx.y;
    unsigned long long sy = 0, sl = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < img.plane; i += (long long)gridDim.x * blockDim.x) {
        int r, g, bl;
        img.load(b, i, r, g, bl);
        if (use_bright) { r = pil_blend(0, r, bright); g = pil_blend(0, g, bright); bl = pil_blend(0, bl, bright); }
        int y, u, v;
        rgb2yuv(r, g, bl, y, u, v);
        sy += (unsigned)y;
        sl += (unsigned)pil_luma(r, g, bl);
    }
    block_add(sy, stats + 2 * b);
    block_add(sl, stats + 2 * b + 1);
}

// stats[b][0] = sum of OpenCV Y (get_image_luma, imfilters.py:597-601), stats[b][1] = number of pixels whose OpenCV HSV
// saturation is below `tht` (the gray mask of restore_color, restcolor.py:51-55)
__global__ void gray_mask_stats_kernel(Img img, unsigned long long *stats, int tht) {
    const int b = blockIdx.y;
    unsigned long long sy = 0, cnt = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < img.plane; i += (long long)gridDim.x * blockDim.x) {
        int r, g, bl, y, u, v, h, s, vv;
        img.load(b, i, r, g, bl);
        rgb2yuv(r, g, bl, y, u, v);
        rgb2hsv(r, g, bl, h, s, vv);
        sy += (unsigned)y;
        cnt += s < tht ? 1u : 0u;
    }
    block_add(sy, stats + 2 * b);
    block_add(cnt, stats + 2 * b + 1);
}

// std.AverageFrames on 8-bit planes (VapourSynth averageframes restated; unpinned): out = clamp((sum_k w_k * src_k + scale/2) / scale).
// Source k of frame b is src + k * clip_stride + b * frame_elems: clip_stride = a whole clip for the multi-clip form,
// = frame_elems for the temporal form on a sequence that carries its halo frames.  Weights are per frame (scene-change folding).
__global__ void average_frames_u8_kernel(const uint8_t *__restrict__ src, long long clip_stride, int n_clips, const int *__restrict__ weights,
                                         int weights_per_frame, int scale, uint8_t *__restrict__ out, long long frame_elems) {
    const int b = blockIdx.y;
    const int *w = weights + (weights_per_frame ? (long long)b * n_clips : 0);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < frame_elems; i += (long long)gridDim.x * blockDim.x) {
        int acc = 0;
        for (int k = 0; k < n_clips; ++k) acc += w[k] * (int)src[k * clip_stride + (long long)b * frame_elems + i];
        acc = (acc + scale / 2) / scale;
        out[(long long)b * frame_elems + i] = (uint8_t)(acc < 0 ? 0 : (acc > 255 ? 255 : acc));
    }
}

// ---- chroma_stabilizer / chroma_stabilizer_adaptive (imfilters.py:160-269) -------------------------------------
struct StabParams {
    int adaptive;
    double up, dn;          // 1 + alpha, 1 - alpha (host doubles)
    int base_tol, max_extra;
    float weight;           // < 1: Image.blend(a, out, weight)
    int H, W;
};
__global__ void chroma_stabilizer_kernel(Img a, Img b, uint8_t *out, StabParams p, unsigned long long *stats) {
    const int fb = blockIdx.y;
    unsigned long long sy = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < a.plane; i += (long long)gridDim.x * blockDim.x) {
        int r1, g1, b1, r2, g2, b2;
        a.load(fb, i, r1, g1, b1);
        b.load(fb, i, r2, g2, b2);
        int y1, u1, v1, y2, u2, v2;
        rgb2yuv(r1, g1, b1, y1, u1, v1);
        rgb2yuv(r2, g2, b2, y2, u2, v2);
        int um, vm;
        if (!p.adaptive) {
            // bounds = uint8(clip(u1 * (1 +- alpha), 0, 255)); array_max then array_min (nputils.py:27-79)
            const int u_up = trunc8(clip255(__dmul_rn((double)u1, p.up))), u_dn = trunc8(clip255(__dmul_rn((double)u1, p.dn)));
            const int v_up = trunc8(clip255(__dmul_rn((double)v1, p.up))), v_dn = trunc8(clip255(__dmul_rn((double)v1, p.dn)));
            um = u2 > u_up ? u_up : u2; um = um < u_dn ? u_dn : um;
            vm = v2 > v_up ? v_up : v2; vm = vm < v_dn ? v_dn : vm;
        } else {
            // texture = clip(|Laplacian(Y1)| / 255, 0, 1): 4-neighbour stencil on float32 Y, BORDER_REFLECT_101
            const int y = (int)(i / p.W), x = (int)(i - (long long)y * p.W);
            auto ya = [&](int yy, int xx) {
                yy = yy < 0 ? -yy : (yy >= p.H ? 2 * p.H - 2 - yy : yy);
                xx = xx < 0 ? -xx : (xx >= p.W ? 2 * p.W - 2 - xx : xx);
                if (p.H == 1) yy = 0;
                if (p.W == 1) xx = 0;
                int rr, gg, bb, yv, uv, vv;
                a.load(fb, (long long)yy * p.W + xx, rr, gg, bb);
                rgb2yuv(rr, gg, bb, yv, uv, vv);
                return (float)yv;
            };
            const float sum = __fadd_rn(__fadd_rn(__fadd_rn(ya(y - 1, x), ya(y + 1, x)), ya(y, x - 1)), ya(y, x + 1));
            const float lap = __fsub_rn(sum, __fmul_rn(4.0f, (float)y1));
            float tex = __fdiv_rn(fabsf(lap), 255.0f);
            tex = tex < 0.f ? 0.f : (tex > 1.f ? 1.f : tex);
            const float tol = __fadd_rn((float)p.base_tol, __fmul_rn((float)p.max_extra, tex));
            auto bound = [&](int c1, int c2) {
                const float lo = fminf(fmaxf(__fsub_rn((float)(c1 - 128), tol), -128.f), 127.f);
                const float hi = fminf(fmaxf(__fadd_rn((float)(c1 - 128), tol), -128.f), 127.f);
                const float m = fminf(fmaxf((float)(c2 - 128), lo), hi);
                return (int)__fadd_rn(m, 128.f);        // astype(uint8): truncation of a value in [0,255]
            };
            um = bound(u1, u2);
            vm = bound(v1, v2);
        }
        int r, g, bl;
        yuv2rgb(y1, um, vm, r, g, bl);
        if (p.weight < 1.0f) { r = pil_blend(r1, r, p.weight); g = pil_blend(g1, g, p.weight); bl = pil_blend(b1, bl, p.weight); }
        store_px(out, a.plane, fb, i, r, g, bl);
        if (stats) { int yo, uo, vo; rgb2yuv(r, g, bl, yo, uo, vo); sy += (unsigned)yo; }
    }
    if (stats) block_add(sy, stats + 2 * fb);
}

// ---- w_image_luma_merge (imfilters.py:80-100, nputils.py:140-185,228-253) ----------------------------------------
struct Ramp { double tresh, grad; };    // host: tresh = min(round(dark*255), round(white*255) - 10), grad = round(1/(max_white - tresh), 3)
__device__ __forceinline__ void ramp_merge(const Ramp &rp, int dr, int dg, int db, int wr, int wg, int wb, int &r, int &g, int &b) {
    const double lum = np_luma(wr, wg, wb);
    double gq = __dmul_rn(__dsub_rn(lum, rp.tresh), rp.grad);
    gq = gq > 1.0 ? 1.0 : (gq < 0.0 ? 0.0 : gq);
    const double mw = (double)(float)gq;          // array_clip(.., np.float32)
    const double mb = __dsub_rn(1.0, mw);
    r = trunc8(clip255(__dadd_rn(__dmul_rn((double)dr, mb), __dmul_rn((double)wr, mw))));
    g = trunc8(clip255(__dadd_rn(__dmul_rn((double)dg, mb), __dmul_rn((double)wg, mw))));
    b = trunc8(clip255(__dadd_rn(__dmul_rn((double)db, mb), __dmul_rn((double)wb, mw))));
}

// ---- red fix (mcomb.py:351-362): dark frames get the hue range 280-360 / 0-30 desaturated ---------------------
__global__ void red_fix_kernel(Img stab, uint8_t *out, const unsigned long long *stats, Ramp r23, Ramp r12) {
    const int fb = blockIdx.y;
    const double luma = frame_luma(stats[2 * fb], stab.plane);
    const int mode = luma > 0.3 ? 0 : (luma > 0.2 ? 1 : (luma > 0.1 ? 2 : 3));
    const float sat = mode == 1 ? 0.9f : (mode == 2 ? 0.8f : 0.7f);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < stab.plane; i += (long long)gridDim.x * blockDim.x) {
        int r0, g0, b0;
        stab.load(fb, i, r0, g0, b0);
        int r = r0, g = g0, b = b0;
        if (mode != 0) {
            const int L = pil_luma(r0, g0, b0);                      // ImageEnhance.Color: blend(L-gray, img, sat)
            int dr = pil_blend(L, r0, sat), dg = pil_blend(L, g0, sat), db = pil_blend(L, b0, sat);
            if (mode == 3) {
                r = dr; g = dg; b = db;
            } else {
                int h, s, v;
                rgb2hsv(r0, g0, b0, h, s, v);
                const bool in = ((double)h > 140.0 && (double)h < 180.0) || ((double)h > 0.0 && (double)h < 15.0);
                if (!in) { dr = r0; dg = g0; db = b0; }              // np_adjust_chroma2 (restcolor.py:344-370)
                ramp_merge(mode == 1 ? r23 : r12, dr, dg, db, r0, g0, b0, r, g, b);
            }
        }
        store_px(out, stab.plane, fb, i, r, g, b);
    }
}

// ---- LumaMaskedMerge (mcomb.py:238-271) -----------------------------------------------------------------------
struct LumaMaskParams { int hard; double hard_thr; Ramp ramp; float weight; int zero_limit; };
__global__ void luma_masked_merge_kernel(Img a, Img b, Img c, uint8_t *out, LumaMaskParams p) {
    const int fb = blockIdx.y;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < a.plane; i += (long long)gridDim.x * blockDim.x) {
        int ra, ga, ba, rb, gb, bb, rc, gc, bc;
        a.load(fb, i, ra, ga, ba);
        b.load(fb, i, rb, gb, bb);
        c.load(fb, i, rc, gc, bc);
        int r, g, bl;
        if (p.hard) {
            if (p.zero_limit) {       // threshold 0: mask = uint8(luma) / 255 (np_rgb_to_gray + np_image_mask_merge)
                const double mw = (double)trunc8(np_luma(rb, gb, bb)) / 255.0, mk = __dsub_rn(1.0, mw);
                r = trunc8(clip255(__dadd_rn(__dmul_rn((double)rc, mk), __dmul_rn((double)rb, mw))));
                g = trunc8(clip255(__dadd_rn(__dmul_rn((double)gc, mk), __dmul_rn((double)gb, mw))));
                bl = trunc8(clip255(__dadd_rn(__dmul_rn((double)bc, mk), __dmul_rn((double)bb, mw))));
            } else {
                const bool white = np_luma(rb, gb, bb) > p.hard_thr;
                r = white ? rb : rc; g = white ? gb : gc; bl = white ? bb : bc;
            }
        } else {
            ramp_merge(p.ramp, rc, gc, bc, rb, gb, bb, r, g, bl);
        }
        if (p.weight < 1.0f && p.weight != 0.0f) { r = pil_blend(ra, r, p.weight); g = pil_blend(ga, g, p.weight); bl = pil_blend(ba, bl, p.weight); }
        if (p.weight == 0.0f) { r = ra; g = ga; bl = ba; }
        store_px(out, a.plane, fb, i, r, g, bl);
    }
}

// ---- AdaptiveLumaMerge (mcomb.py:289-314) -----------------------------------------------------------------------
__global__ void adaptive_luma_merge_kernel(Img a, Img b, uint8_t *out, const unsigned long long *stats_b, double luma_limit,
                                           double alpha, double weight, double min_w) {
    const int fb = blockIdx.y;
    const double luma = frame_luma(stats_b[2 * fb], a.plane);
    double w = weight;
    if (luma < luma_limit) w = fmax(__dmul_rn(weight, pow(luma / luma_limit, alpha)), min_w);
    const float wf = (float)w;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < a.plane; i += (long long)gridDim.x * blockDim.x) {
        int ra, ga, ba, rb, gb, bb;
        a.load(fb, i, ra, ga, ba);
        b.load(fb, i, rb, gb, bb);
        store_px(out, a.plane, fb, i, pil_blend(ra, rb, wf), pil_blend(ga, gb, wf), pil_blend(ba, bb, wf));
    }
}

// ---- restore_color_gradient + std.Merge (restcolor.py:98-134, vsfilters.py:366-422,730-739) --------------------
struct RestoreParams {
    double sat; int scale_sat;
    const uint8_t *lut, *lut_gated;      // gradient mask per saturation value (host-built: w_np_gradient_mask)
    double w, omw, wg, omwg;             // |weight| and 1-|weight|, normal / luma-gated
    int wsign, wgsign;                   // sign of the weight: >0 merge with colour, <0 merge with gray, 0 none
    int gate; double dark, bright;       // DEF_STANDARD_DARK / BRIGHT gating on the frame luma of `gray`
    int merge_w15;                       // std.Merge(gray, restored, w): 15-bit weight, < 0 = no merge
    int W, simd_width;
    // restore_color (restcolor.py:38-83, the temporal stabiliser's binary restore): frame-level bypasses -> out = gray
    const uint8_t *active;               // per frame; 0 = the selector returns the frame untouched (n < 15, vsfilters.py:337)
    double tht_scen;                     // > 0: a frame whose share of gray pixels (stats[2b+1] / plane) exceeds it is returned as is
};
__global__ void restore_color_gradient_kernel(Img color, Img gray, uint8_t *out, RestoreParams p, const unsigned long long *stats) {
    const int fb = blockIdx.y;
    bool bypass = p.active != nullptr && p.active[fb] == 0;
    if (!bypass && p.tht_scen > 0.0 && p.tht_scen < 1.0) {                 // np.mean(mask) / 255 > tht_scen (restcolor.py:55-61)
        const double share = __ddiv_rn(__ddiv_rn((double)(255ull * stats[2 * fb + 1]), (double)gray.plane), 255.0);
        bypass = share > p.tht_scen;
    }
    if (bypass) {
        for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < gray.plane; i += (long long)gridDim.x * blockDim.x) {
            int rg, gg, bg;
            gray.load(fb, i, rg, gg, bg);
            store_px(out, gray.plane, fb, i, rg, gg, bg);
        }
        return;
    }
    bool gated = false;
    if (p.gate) {
        const double luma = frame_luma(stats[2 * fb], gray.plane);
        gated = !(p.dark <= luma && luma <= p.bright);
    }
    const uint8_t *lut = gated ? p.lut_gated : p.lut;
    const double w = gated ? p.wg : p.w, omw = gated ? p.omwg : p.omw;
    const int ws = gated ? p.wgsign : p.wsign;
    const int body = p.simd_width > 0 ? (p.W / p.simd_width) * p.simd_width : 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < gray.plane; i += (long long)gridDim.x * blockDim.x) {
        int rc, gc, bc, rg, gg, bg;
        color.load(fb, i, rc, gc, bc);
        gray.load(fb, i, rg, gg, bg);
        int h, s, v, hg, sg, vg;
        rgb2hsv(rc, gc, bc, h, s, v);
        rgb2hsv(rg, gg, bg, hg, sg, vg);
        if (p.scale_sat) s = (int)((long long)__dmul_rn((double)s, p.sat) & 0xff);
        const int x = (int)(i % p.W);
        int rs, gs, bs;
        hsv2rgb(h, s, v, x < body, rs, gs, bs);
        const double mw = (double)__ldg(lut + sg) / 255.0, mb = __dsub_rn(1.0, mw);
        int r = trunc8(clip255(__dadd_rn(__dmul_rn((double)rg, mb), __dmul_rn((double)rs, mw))));
        int g = trunc8(clip255(__dadd_rn(__dmul_rn((double)gg, mb), __dmul_rn((double)gs, mw))));
        int b = trunc8(clip255(__dadd_rn(__dmul_rn((double)bg, mb), __dmul_rn((double)bs, mw))));
        if (ws > 0) { r = np_wmerge(r, rs, omw, w); g = np_wmerge(g, gs, omw, w); b = np_wmerge(b, bs, omw, w); }
        if (ws < 0) { r = np_wmerge(r, rg, omw, w); g = np_wmerge(g, gg, omw, w); b = np_wmerge(b, bg, omw, w); }
        if (p.merge_w15 >= 0) {     // VapourSynth std.Merge, 8-bit: a + (((b - a) * w15 + 2^14) >> 15)   (restated; unpinned)
            r = rg + (((r - rg) * p.merge_w15 + (1 << 14)) >> 15);
            g = gg + (((g - gg) * p.merge_w15 + (1 << 14)) >> 15);
            b = bg + (((b - bg) * p.merge_w15 + (1 << 14)) >> 15);
        }
        store_px(out, gray.plane, fb, i, r, g, b);
    }
}

// ---- adjust_chroma (restcolor.py:239-286; vs_sc_adjust_clip_hue vsfilters.py:435-455) -----------------------------
struct AdjustParams { HueRanges rng; double sat; int scale_sat; double hue_half; int hue_on; double w, omw; int wsign; int W, simd_width; };
__global__ void adjust_chroma_kernel(Img img, uint8_t *out, AdjustParams p) {
    const int fb = blockIdx.y;
    const int body = p.simd_width > 0 ? (p.W / p.simd_width) * p.simd_width : 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < img.plane; i += (long long)gridDim.x * blockDim.x) {
        int r0, g0, b0;
        img.load(fb, i, r0, g0, b0);
        int h, s, v;
        rgb2hsv(r0, g0, b0, h, s, v);
        int h2 = h, s2 = s;
        if (p.hue_on) {     // np_hue_add (nputils.py:330-340), stored back into a uint8 plane
            double x = __dadd_rn((double)h, p.hue_half);
            x = x > 180.0 ? x - 180.0 : x;
            x = x < 0.0 ? x + 180.0 : x;
            h2 = (int)x & 0xff;
        }
        if (p.scale_sat) s2 = (int)((long long)__dmul_rn((double)s, p.sat) & 0xff);
        int rg, gg, bg;
        hsv2rgb(h2, s2, v, (int)(i % p.W) < body, rg, gg, bg);
        const bool in = p.rng.hit(h);
        int r = in ? rg : r0, g = in ? gg : g0, b = in ? bg : b0;
        if (p.wsign > 0) {
            if (!p.hue_on) { r = np_wmerge(r, rg, p.omw, p.w); g = np_wmerge(g, gg, p.omw, p.w); b = np_wmerge(b, bg, p.omw, p.w); }
            else { r = np_wmerge(r, r0, p.omw, p.w); g = np_wmerge(g, g0, p.omw, p.w); b = np_wmerge(b, b0, p.omw, p.w); }
        }
        if (p.wsign < 0) { r = np_wmerge(r, r0, p.omw, p.w); g = np_wmerge(g, g0, p.omw, p.w); b = np_wmerge(b, b0, p.omw, p.w); }
        store_px(out, img.plane, fb, i, r, g, b);
    }
}

// ---- np_image_chroma_tweak (restcolor.py:288-342) fused with the luma merge of vs_sc_chroma_bright_tweak ------------
// (vsfilters.py:525-552) / _vs_sc_colormap (vsfilters.py:577-590): HAVC_stabilizer's `smooth` and `colormap` stages.
struct ChromaTweakParams {
    double sat, val, hue_half; int hue_on;
    int stage2; HueRanges rng; double sat2; int scale_sat2; double hue2_half; int hue2_on; double w, omw; int wsign;
    int merge; LumaMaskParams lm;      // merge != 0: out = luma_merge(img_dark = tweaked, img_white = img)
    int W, simd_width;
};
__device__ __forceinline__ int cv_hue_add(int h, double half) {     // np_hue_add (nputils.py:330-340) + the uint8 store
    double x = __dadd_rn((double)h, half);
    x = x > 180.0 ? x - 180.0 : x;
    x = x < 0.0 ? x + 180.0 : x;
    return (int)x & 0xff;
}
__global__ void chroma_tweak_kernel(Img img, uint8_t *out, ChromaTweakParams p) {
    const int fb = blockIdx.y;
    const int body = p.simd_width > 0 ? (p.W / p.simd_width) * p.simd_width : 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < img.plane; i += (long long)gridDim.x * blockDim.x) {
        int r0, g0, b0;
        img.load(fb, i, r0, g0, b0);
        const bool simd = (int)(i % p.W) < body;
        int h, s, v;
        rgb2hsv(r0, g0, b0, h, s, v);
        if (p.hue_on) h = cv_hue_add(h, p.hue_half);
        s = (int)((long long)__dmul_rn((double)s, p.sat) & 0xff);      // float64 product stored into a uint8 plane
        v = (int)((long long)__dmul_rn((double)v, p.val) & 0xff);
        int r, g, b;
        hsv2rgb(h, s, v, simd, r, g, b);
        if (p.stage2) {           // "chroma adjustment": the mask comes from the tweaked hue, unmasked pixels from the ORIGINAL image
            int hg, sg, vg, rg, gg, bg;
            rgb2hsv(r, g, b, hg, sg, vg);
            if (p.hue2_on) hg = cv_hue_add(hg, p.hue2_half);
            if (p.scale_sat2) sg = (int)((long long)__dmul_rn((double)sg, p.sat2) & 0xff);
            hsv2rgb(hg, sg, vg, simd, rg, gg, bg);
            const bool in = p.rng.hit(h);
            r = in ? rg : r0; g = in ? gg : g0; b = in ? bg : b0;
            if (p.wsign > 0) {
                if (!p.hue2_on) { r = np_wmerge(r, rg, p.omw, p.w); g = np_wmerge(g, gg, p.omw, p.w); b = np_wmerge(b, bg, p.omw, p.w); }
                else { r = np_wmerge(r, r0, p.omw, p.w); g = np_wmerge(g, g0, p.omw, p.w); b = np_wmerge(b, b0, p.omw, p.w); }
            }
            if (p.wsign < 0) { r = np_wmerge(r, r0, p.omw, p.w); g = np_wmerge(g, g0, p.omw, p.w); b = np_wmerge(b, b0, p.omw, p.w); }
        }
        if (p.merge) {
            int ro, go, bo;
            if (p.lm.hard) {
                if (p.lm.zero_limit) {
                    const double mw = (double)trunc8(np_luma(r0, g0, b0)) / 255.0, mk = __dsub_rn(1.0, mw);
                    ro = trunc8(clip255(__dadd_rn(__dmul_rn((double)r, mk), __dmul_rn((double)r0, mw))));
                    go = trunc8(clip255(__dadd_rn(__dmul_rn((double)g, mk), __dmul_rn((double)g0, mw))));
                    bo = trunc8(clip255(__dadd_rn(__dmul_rn((double)b, mk), __dmul_rn((double)b0, mw))));
                } else {
                    const bool white = np_luma(r0, g0, b0) > p.lm.hard_thr;
                    ro = white ? r0 : r; go = white ? g0 : g; bo = white ? b0 : b;
                }
            } else {
                ramp_merge(p.lm.ramp, r, g, b, r0, g0, b0, ro, go, bo);
            }
            r = ro; g = go; b = bo;
        }
        store_px(out, img.plane, fb, i, r, g, b);
    }
}

// ---- image_tweak (imfilters.py:463-504): Brightness -> Contrast -> Color, optional hue-range restriction --------
struct TweakParams { float bright; int use_bright; float cont; int use_cont; float sat; int use_sat; HueRanges rng; };
__global__ void image_tweak_kernel(Img img, uint8_t *out, TweakParams p, const unsigned long long *stats) {
    const int fb = blockIdx.y;
    int mean = 0;
    if (p.use_cont) mean = (int)((double)stats[2 * fb + 1] / (double)img.plane + 0.5);   // int(ImageStat mean(L) + 0.5)
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < img.plane; i += (long long)gridDim.x * blockDim.x) {
        int r0, g0, b0;
        img.load(fb, i, r0, g0, b0);
        int r = r0, g = g0, b = b0;
        if (p.use_bright) { r = pil_blend(0, r, p.bright); g = pil_blend(0, g, p.bright); b = pil_blend(0, b, p.bright); }
        if (p.use_cont) { r = pil_blend(mean, r, p.cont); g = pil_blend(mean, g, p.cont); b = pil_blend(mean, b, p.cont); }
        if (p.use_sat) {
            const int L = pil_luma(r, g, b);
            r = pil_blend(L, r, p.sat); g = pil_blend(L, g, p.sat); b = pil_blend(L, b, p.sat);
        }
        if (p.rng.n > 0) {
            int h, s, v;
            rgb2hsv(r0, g0, b0, h, s, v);
            if (!p.rng.hit(h)) { r = r0; g = g0; b = b0; }
        }
        store_px(out, img.plane, fb, i, r, g, b);
    }
}

// ---- luma_adjusted_levels (imfilters.py:335-372; sc_constrained_tweak vsfilters.py:656-675) ----------------------
struct LevelsParams { double luma_min, gamma, gamma_luma_min, gamma_alpha, gamma_min; };
__global__ void luma_adjusted_levels_kernel(Img img, uint8_t *out, LevelsParams p, const unsigned long long *stats) {
    const int fb = blockIdx.y;
    const double luma = ((double)stats[2 * fb] / (double)img.plane) / 255.0;     // np.mean(Y) / 255, not rounded here
    int i_alpha = 0;
    if (luma < p.luma_min) i_alpha = (int)(255.0 * (p.luma_min - luma));
    const bool do_gamma = p.gamma != 1.0 && luma < p.gamma_luma_min;
    double inv_g = 1.0;
    if (do_gamma) {
        const double g_new = p.gamma_alpha != 0.0 ? fmax(p.gamma * pow(luma / p.gamma_luma_min, p.gamma_alpha), p.gamma_min) : p.gamma;
        inv_g = 1.0 / g_new;
    }
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < img.plane; i += (long long)gridDim.x * blockDim.x) {
        int r0, g0, b0;
        img.load(fb, i, r0, g0, b0);
        int y, u, v;
        rgb2yuv(r0, g0, b0, y, u, v);
        if (i_alpha > 1) y = sat8(y + i_alpha);
        if (do_gamma) y = trunc8(clip255(__dmul_rn(pow((double)y / 255.0, inv_g), 255.0)));
        int r, g, b;
        yuv2rgb(y, u, v, r, g, b);
        store_px(out, img.plane, fb, i, r, g, b);
    }
}

// ---- VapourSynth std.Merge on 8-bit planes (vs_simple_merge, vsfilters.py:730-739): restated, parity unpinned ------
__global__ void vs_merge_u8_kernel(const uint8_t *__restrict__ a, const uint8_t *__restrict__ b, uint8_t *__restrict__ out, long long n,
                                   int w15) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int va = a[i], vb = b[i];
        out[i] = (uint8_t)(va + (((vb - va) * w15 + (1 << 14)) >> 15));
    }
}

// ---- chroma_post_process (imfilters.py:312-321; ColorizerFilter._post_process filters.py:100-110) --------------------------
__global__ void chroma_post_process_kernel(Img color, Img orig, uint8_t *out) {
    const int fb = blockIdx.y;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < orig.plane; i += (long long)gridDim.x * blockDim.x) {
        int rc, gc, bc, ro, go, bo, r, g, b;
        color.load(fb, i, rc, gc, bc);
        orig.load(fb, i, ro, go, bo);
        luma_transplant(ro, go, bo, rc, gc, bc, r, g, b);
        store_px(out, orig.plane, fb, i, r, g, b);
    }
}

// ---- scene-change gate (vsslib/vsmodels.py:221-224, mcomb.py:210-213): frames with skip[b] != 0 take `src` unchanged ------
__global__ void select_frames_kernel(uint8_t *__restrict__ dst, const uint8_t *__restrict__ src, const uint8_t *__restrict__ skip,
                                     long long frame_bytes) {
    const int b = blockIdx.y;
    if (!skip[b]) return;
    uint8_t *d = dst + (long long)b * frame_bytes;
    const uint8_t *s = src + (long long)b * frame_bytes;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < frame_bytes; i += (long long)gridDim.x * blockDim.x) d[i] = s[i];
}

static dim3 frame_grid(long long plane, int B, int block = 256) {
    long long g = (plane + block - 1) / block;
    const long long cap = (long long)num_sms() * 8;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return dim3((unsigned)g, (unsigned)B, 1);
}

static bool ranges_ok(const havc_hue_ranges *r) { return r != nullptr && r->n >= 0 && r->n <= HAVC_MAX_HUE_RANGES; }
static HueRanges to_dev_ranges(const havc_hue_ranges *r) {
    HueRanges d;
    d.n = r ? r->n : 0;
    for (int i = 0; i < HAVC_MAX_HUE_RANGES; ++i) {
        d.lo[i] = (r && i < r->n) ? r->lo_deg[i] * 0.5 : 0.0;
        d.hi[i] = (r && i < r->n) ? r->hi_deg[i] * 0.5 : 0.0;
    }
    return d;
}
// tresh / grad of w_np_rgb_to_gray (nputils.py:160-166); Python round() = round-half-even
static Ramp make_ramp(double dark, double white) {
    const double max_white = nearbyint(white * 255.0);
    double tresh = nearbyint(dark * 255.0);
    if (max_white - 10.0 < tresh) tresh = max_white - 10.0;
    const double inv = 1.0 / (max_white - tresh);
    Ramp r;
    r.tresh = tresh;
    r.grad = nearbyint(inv * 1000.0) / 1000.0;       // round(x, 3)
    return r;
}

}  // namespace havc

using namespace havc;

#define HAVC_IMG_ARGS_OK(B, H, W) ((B) > 0 && (H) > 0 && (W) > 0)

extern "C" int havc_frame_stats(const uint8_t *img, int B, int H, int W, float bright, unsigned long long *stats, void *stream) {
    HAVC_CHECK_ARG(img && stats && HAVC_IMG_ARGS_OK(B, H, W), "havc_frame_stats: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    HAVC_CHECK_CUDA(cudaMemsetAsync(stats, 0, sizeof(unsigned long long) * 2 * B, st));
    Img im{img, (long long)H * W};
    frame_stats_kernel<<<frame_grid(im.plane, B), 256, 0, st>>>(im, stats, bright, bright != 1.0f);
    HAVC_LAUNCHED();
    return HAVC_OK;
}

extern "C" int havc_chroma_stabilizer(const uint8_t *a, const uint8_t *b, uint8_t *out, int B, int H, int W, int adaptive,
                                      double alpha, int base_tol, int max_extra, float weight, unsigned long long *stats_out,
                                      void *stream) {
    HAVC_CHECK_ARG(a && b && out && HAVC_IMG_ARGS_OK(B, H, W), "havc_chroma_stabilizer: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    if (stats_out) HAVC_CHECK_CUDA(cudaMemsetAsync(stats_out, 0, sizeof(unsigned long long) * 2 * B, st));
    StabParams p;
    p.adaptive = adaptive; p.up = 1.0 + alpha; p.dn = 1.0 - alpha; p.base_tol = base_tol; p.max_extra = max_extra;
    p.weight = weight; p.H = H; p.W = W;
    Img ia{a, (long long)H * W}, ib{b, (long long)H * W};
    chroma_stabilizer_kernel<<<frame_grid(ia.plane, B), 256, 0, st>>>(ia, ib, out, p, stats_out);
    HAVC_LAUNCHED();
    return HAVC_OK;
}

extern "C" int havc_red_fix(const uint8_t *stab, uint8_t *out, int B, int H, int W, const unsigned long long *stats, void *stream) {
    HAVC_CHECK_ARG(stab && out && stats && HAVC_IMG_ARGS_OK(B, H, W), "havc_red_fix: bad arguments");
    Img im{stab, (long long)H * W};
    red_fix_kernel<<<frame_grid(im.plane, B), 256, 0, (cudaStream_t)stream>>>(im, out, stats, make_ramp(0.2, 0.3), make_ramp(0.1, 0.2));
    HAVC_LAUNCHED();
    return HAVC_OK;
}

// Fills the hard-mask / ramp description shared by havc_luma_masked_merge and havc_chroma_tweak; false = unsupported.
static bool make_luma_mask(double luma_limit, double white_limit, LumaMaskParams &p) {
    if (luma_limit == white_limit) {                    // image_luma_merge: hard mask (imfilters.py:66-78)
        p.hard = 1;
        p.zero_limit = !(luma_limit > 0);
        p.hard_thr = nearbyint(luma_limit * 255.0);
    } else if (luma_limit > white_limit) {              // w_image_luma_merge returns img_dark
        p.hard = 1; p.zero_limit = 0; p.hard_thr = 1e30;
    } else if (!(luma_limit > 0)) {                     // ramp with dark_luma == 0: weight = luma / 255 (nputils.py:178-183)
        return false;
    } else {
        p.ramp = make_ramp(luma_limit, white_limit);
    }
    return true;
}

extern "C" int havc_luma_masked_merge(const uint8_t *a, const uint8_t *b, const uint8_t *c, uint8_t *out, int B, int H, int W,
                                      double luma_limit, double white_limit, float weight, void *stream) {
    HAVC_CHECK_ARG(a && b && c && out && HAVC_IMG_ARGS_OK(B, H, W), "havc_luma_masked_merge: bad arguments");
    LumaMaskParams p;
    memset(&p, 0, sizeof(p));
    p.weight = weight;
    HAVC_CHECK_ARG(make_luma_mask(luma_limit, white_limit, p), "havc_luma_masked_merge: luma_limit = 0 with a white limit is not supported");
    Img ia{a, (long long)H * W}, ib{b, (long long)H * W}, ic{c, (long long)H * W};
    luma_masked_merge_kernel<<<frame_grid(ia.plane, B), 256, 0, (cudaStream_t)stream>>>(ia, ib, ic, out, p);
    HAVC_LAUNCHED();
    return HAVC_OK;
}

extern "C" int havc_adaptive_luma_merge(const uint8_t *a, const uint8_t *b, uint8_t *out, int B, int H, int W,
                                        const unsigned long long *stats_b, double luma_threshold, double alpha, double weight,
                                        double min_weight, void *stream) {
    HAVC_CHECK_ARG(a && b && out && stats_b && HAVC_IMG_ARGS_OK(B, H, W), "havc_adaptive_luma_merge: bad arguments");
    Img ia{a, (long long)H * W}, ib{b, (long long)H * W};
    adaptive_luma_merge_kernel<<<frame_grid(ia.plane, B), 256, 0, (cudaStream_t)stream>>>(ia, ib, out, stats_b, luma_threshold, alpha,
                                                                                         weight, min_weight);
    HAVC_LAUNCHED();
    return HAVC_OK;
}

extern "C" int havc_restore_color_gradient(const uint8_t *color, const uint8_t *gray, uint8_t *out, int B, int H, int W, double sat,
                                           const uint8_t *lut, const uint8_t *lut_gated, double weight, double weight_gated,
                                           const unsigned long long *stats_gray, double merge_weight, int simd_width, void *stream) {
    HAVC_CHECK_ARG(color && gray && out && lut && HAVC_IMG_ARGS_OK(B, H, W), "havc_restore_color_gradient: bad arguments");
    HAVC_CHECK_ARG(stats_gray == nullptr || lut_gated != nullptr, "havc_restore_color_gradient: luma gating needs lut_gated");
    RestoreParams p;
    memset(&p, 0, sizeof(p));
    p.scale_sat = sat != 1.0;
    p.sat = sat < 0 ? 0 : (sat > 10 ? 10 : sat);
    p.lut = lut; p.lut_gated = lut_gated ? lut_gated : lut;
    p.w = fabs(weight); p.omw = 1.0 - p.w; p.wsign = weight > 0 ? 1 : (weight < 0 ? -1 : 0);
    p.wg = fabs(weight_gated); p.omwg = 1.0 - p.wg; p.wgsign = weight_gated > 0 ? 1 : (weight_gated < 0 ? -1 : 0);
    p.gate = stats_gray != nullptr; p.dark = 0.22; p.bright = 0.78;       // vsslib/constants.py:28-29
    p.merge_w15 = -1;
    if (merge_weight >= 0) {                                              // vs_simple_merge (vsfilters.py:730-739)
        int w15 = (int)(merge_weight * 32768.0 + 0.5);
        p.merge_w15 = w15 < 0 ? 0 : (w15 > 32768 ? 32768 : w15);
    }
    p.W = W; p.simd_width = simd_width;
    Img ic{color, (long long)H * W}, ig{gray, (long long)H * W};
    restore_color_gradient_kernel<<<frame_grid(ic.plane, B), 256, 0, (cudaStream_t)stream>>>(ic, ig, out, p, stats_gray);
    HAVC_LAUNCHED();
    return HAVC_OK;
}

/* restore_color (restcolor.py:38-83) as vs_sc_recover_clip_color calls it (vsfilters.py:327-351): binary gray mask (OpenCV S of
   `gray` < tht) takes the (desaturated) colours of `color`; weight > 0 merges the result with `gray`, < 0 with the colours;
   frames whose luma is outside [0.22, 0.78] use min(weight, -0.8); a frame whose gray share exceeds tht_scen, or whose
   active[b] == 0, is returned untouched.  stats = havc_gray_mask_stats(gray). */
extern "C" int havc_restore_color(const uint8_t *color, const uint8_t *gray, uint8_t *out, int B, int H, int W, double sat, int tht,
                                  double weight, double tht_scen, const unsigned long long *stats_gray, const uint8_t *active,
                                  const uint8_t *lut, int simd_width, void *stream) {
    HAVC_CHECK_ARG(color && gray && out && lut && stats_gray && HAVC_IMG_ARGS_OK(B, H, W) && tht >= 0 && tht <= 256,
                   "havc_restore_color: bad arguments");
    RestoreParams p;
    memset(&p, 0, sizeof(p));
    p.scale_sat = sat != 1.0;
    p.sat = sat < 0 ? 0 : (sat > 10 ? 10 : sat);
    p.lut = lut; p.lut_gated = lut;                                       // lut[s] = s < tht ? 255 : 0 (host-built)
    const double wg = weight < -0.8 ? weight : -0.8;                      // vsfilters.py:347-348
    // the kernel's sign convention is restore_color_gradient's (> 0: merge with the colours), restore_color's is the opposite
    p.w = fabs(weight); p.omw = 1.0 - p.w; p.wsign = weight > 0 ? -1 : (weight < 0 ? 1 : 0);
    p.wg = fabs(wg); p.omwg = 1.0 - p.wg; p.wgsign = 1;
    p.gate = 1; p.dark = 0.22; p.bright = 0.78;
    p.merge_w15 = -1;
    p.W = W; p.simd_width = simd_width;
    p.active = active; p.tht_scen = tht_scen;
    Img ic{color, (long long)H * W}, ig{gray, (long long)H * W};
    restore_color_gradient_kernel<<<frame_grid(ic.plane, B), 256, 0, (cudaStream_t)stream>>>(ic, ig, out, p, stats_gray);
    HAVC_LAUNCHED();
    return HAVC_OK;
}

extern "C" int havc_gray_mask_stats(const uint8_t *img, int B, int H, int W, int tht, unsigned long long *stats, void *stream) {
    HAVC_CHECK_ARG(img && stats && HAVC_IMG_ARGS_OK(B, H, W), "havc_gray_mask_stats: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    HAVC_CHECK_CUDA(cudaMemsetAsync(stats, 0, sizeof(unsigned long long) * 2 * B, st));
    Img im{img, (long long)H * W};
    gray_mask_stats_kernel<<<frame_grid(im.plane, B), 256, 0, st>>>(im, stats, tht);
    HAVC_LAUNCHED();
    return HAVC_OK;
}

extern "C" int havc_average_frames_u8(const uint8_t *src, long long clip_stride, int n_clips, const int *weights, int weights_per_frame,
                                      int scale, uint8_t *out, int B, long long frame_elems, void *stream) {
    HAVC_CHECK_ARG(src && weights && out && B > 0 && frame_elems > 0 && n_clips > 0 && n_clips <= 31 && scale > 0 && clip_stride > 0,
                   "havc_average_frames_u8: bad arguments");
    average_frames_u8_kernel<<<frame_grid(frame_elems, B), 256, 0, (cudaStream_t)stream>>>(src, clip_stride, n_clips, weights,
                                                                                           weights_per_frame, scale, out, frame_elems);
    HAVC_LAUNCHED();
    return HAVC_OK;
}

extern "C" int havc_adjust_chroma(const uint8_t *img, uint8_t *out, int B, int H, int W, const havc_hue_ranges *ranges, double sat,
                                  int hue, double weight, int simd_width, void *stream) {
    HAVC_CHECK_ARG(img && out && HAVC_IMG_ARGS_OK(B, H, W) && ranges_ok(ranges) && ranges->n > 0, "havc_adjust_chroma: bad arguments");
    AdjustParams p;
    memset(&p, 0, sizeof(p));
    p.rng = to_dev_ranges(ranges);
    p.scale_sat = sat != 1.0;
    p.sat = sat < 0 ? 0 : (sat > 10 ? 10 : sat);
    p.hue_on = hue != 0;
    const int hc = hue < -360 ? -360 : (hue > 360 ? 360 : hue);
    p.hue_half = 0.5 * hc;
    p.w = fabs(weight); p.omw = 1.0 - p.w; p.wsign = weight > 0 ? 1 : (weight < 0 ? -1 : 0);
    p.W = W; p.simd_width = simd_width;
    Img im{img, (long long)H * W};
    adjust_chroma_kernel<<<frame_grid(im.plane, B), 256, 0, (cudaStream_t)stream>>>(im, out, p);
    HAVC_LAUNCHED();
    return HAVC_OK;
}

extern "C" int havc_chroma_tweak(const uint8_t *img, uint8_t *out, int B, int H, int W, double sat, double bright, int hue,
                                 const havc_hue_ranges *ranges, double sat2, int hue2, double weight, int luma_merge,
                                 double luma_limit, double white_limit, int simd_width, void *stream) {
    HAVC_CHECK_ARG(img && out && HAVC_IMG_ARGS_OK(B, H, W) && (ranges == nullptr || (ranges_ok(ranges) && ranges->n > 0)),
                   "havc_chroma_tweak: bad arguments");
    ChromaTweakParams p;
    memset(&p, 0, sizeof(p));
    auto clamp10 = [](double x) { return x < 0 ? 0.0 : (x > 10 ? 10.0 : x); };
    auto half_hue = [](int h) { return 0.5 * (h < -360 ? -360 : (h > 360 ? 360 : h)); };
    p.sat = clamp10(sat); p.val = clamp10(1.0 + bright);
    p.hue_on = hue != 0; p.hue_half = half_hue(hue);
    if (ranges != nullptr) {
        p.stage2 = 1;
        p.rng = to_dev_ranges(ranges);
        p.scale_sat2 = sat2 != 1.0; p.sat2 = clamp10(sat2);
        p.hue2_on = hue2 != 0; p.hue2_half = half_hue(hue2);
        p.w = fabs(weight); p.omw = 1.0 - p.w; p.wsign = weight > 0 ? 1 : (weight < 0 ? -1 : 0);
    }
    if (luma_merge) {
        p.merge = 1;
        HAVC_CHECK_ARG(make_luma_mask(luma_limit, white_limit, p.lm), "havc_chroma_tweak: luma_limit = 0 with a white limit is not supported");
    }
    p.W = W; p.simd_width = simd_width;
    Img im{img, (long long)H * W};
    chroma_tweak_kernel<<<frame_grid(im.plane, B), 256, 0, (cudaStream_t)stream>>>(im, out, p);
    HAVC_LAUNCHED();
    return HAVC_OK;
}

extern "C" int havc_image_tweak(const uint8_t *img, uint8_t *out, int B, int H, int W, double sat, double cont, double bright,
                                const havc_hue_ranges *ranges, unsigned long long *stats_scratch, void *stream) {
    HAVC_CHECK_ARG(img && out && HAVC_IMG_ARGS_OK(B, H, W) && (ranges == nullptr || ranges_ok(ranges)), "havc_image_tweak: bad arguments");
    HAVC_CHECK_ARG(cont == 1.0 || stats_scratch != nullptr, "havc_image_tweak: contrast needs a stats scratch buffer (2*B u64)");
    TweakParams p;
    memset(&p, 0, sizeof(p));
    p.use_bright = bright != 0.0;
    p.bright = (float)(1.0 + bright / 255.0);       // ImageEnhance factors are Python floats, narrowed to C float by Image.blend
    p.use_cont = cont != 1.0; p.cont = (float)cont;
    p.use_sat = sat != 1.0; p.sat = (float)sat;
    p.rng = to_dev_ranges(ranges);
    if (p.use_cont) {
        int rc = havc_frame_stats(img, B, H, W, p.use_bright ? p.bright : 1.0f, stats_scratch, stream);
        if (rc) return rc;
    }
    Img im{img, (long long)H * W};
    image_tweak_kernel<<<frame_grid(im.plane, B), 256, 0, (cudaStream_t)stream>>>(im, out, p, stats_scratch);
    HAVC_LAUNCHED();
    return HAVC_OK;
}

extern "C" int havc_luma_adjusted_levels(const uint8_t *img, uint8_t *out, int B, int H, int W, const unsigned long long *stats,
                                         double luma_min, double gamma, double gamma_luma_min, double gamma_alpha, double gamma_min,
                                         void *stream) {
    HAVC_CHECK_ARG(img && out && stats && HAVC_IMG_ARGS_OK(B, H, W), "havc_luma_adjusted_levels: bad arguments");
    LevelsParams p{luma_min, gamma, gamma_luma_min, gamma_alpha, gamma_min};
    Img im{img, (long long)H * W};
    luma_adjusted_levels_kernel<<<frame_grid(im.plane, B), 256, 0, (cudaStream_t)stream>>>(im, out, p, stats);
    HAVC_LAUNCHED();
    return HAVC_OK;
}

extern "C" int havc_vs_merge_u8(const uint8_t *a, const uint8_t *b, uint8_t *out, long long n, double weight, void *stream) {
    HAVC_CHECK_ARG(a && b && out && n > 0 && weight >= 0.0 && weight <= 1.0, "havc_vs_merge_u8: bad arguments");
    int w15 = (int)(weight * 32768.0 + 0.5);
    w15 = w15 < 0 ? 0 : (w15 > 32768 ? 32768 : w15);
    long long g = (n + 255) / 256;
    const long long cap = (long long)num_sms() * 32;
    if (g > cap) g = cap;
    vs_merge_u8_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(a, b, out, n, w15);
    HAVC_LAUNCHED();
    return HAVC_OK;
}

extern "C" int havc_select_frames(uint8_t *dst, const uint8_t *src, const uint8_t *skip, int B, long long frame_bytes, void *stream) {
    HAVC_CHECK_ARG(dst && src && skip && B > 0 && frame_bytes > 0, "havc_select_frames: bad arguments");
    select_frames_kernel<<<frame_grid(frame_bytes, B), 256, 0, (cudaStream_t)stream>>>(dst, src, skip, frame_bytes);
    HAVC_LAUNCHED();
    return HAVC_OK;
}

extern "C" int havc_chroma_post_process(const uint8_t *color, const uint8_t *orig, uint8_t *out, int B, int H, int W, void *stream) {
    HAVC_CHECK_ARG(color && orig && out && HAVC_IMG_ARGS_OK(B, H, W), "havc_chroma_post_process: bad arguments");
    Img ic{color, (long long)H * W}, io{orig, (long long)H * W};
    chroma_post_process_kernel<<<frame_grid(io.plane, B), 256, 0, (cudaStream_t)stream>>>(ic, io, out);
    HAVC_LAUNCHED();
    return HAVC_OK;
}
