// pixel.cu — frame pre/post pixel passes (HBM-bound, one thread per output element, coalesced rows).
//
// Pre  (reference: HAVC_colorizer Spline64 squeeze vsdeoldify/__init__.py:2504; ColorizerFilter._transform
//       filters.py:92-93; normalize filters.py:50-53 + fastai/vision/data.py:56-79):
//       planar RGB u8 W x H --separable resample--> RGB u8 S x S --L--> (L/255 - mean)/std  NHWC 16-bit.
// Head (unet.py:276-281 final 1x1 conv + SigmoidRange; denorm/clamp/*255/trunc filters.py:64-67 +
//       fastai/basic_train.py:358-362 + vision/data.py:300; _post_process filters.py:100-110 at S x S).
// Post (_clip_chroma_resize vsdeoldify/__init__.py:3545-3554: Spline64 back to W x H, then
//       chroma_post_process imfilters.py:312-321 with the original frame).
//
// Resampling is table driven: for output index o, out[o] = sum_t w[o][t] * in[start[o] + t]; the host
// builds the tables (Spline64 / Spline36 / Pillow triangle), so every resampler shares these kernels.
#include "common.cuh"
#include "pixel_math.cuh"
#include <stdlib.h>

namespace havc {

// Horizontal pass: u8 rows -> float rows.  One block per input row: the row is staged in shared memory as
// floats (coalesced 16-byte loads), every thread then produces outputs ox = tid, tid+blockDim, ...
__global__ void resample_h_kernel(const uint8_t *__restrict__ in, float *__restrict__ out, long long rows, int Win,
                                  int Wout, const int *__restrict__ start, const float *__restrict__ wts, int T) {
    extern __shared__ float srow[];
    for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
        const uint8_t *src = in + row * Win;
        if ((Win & 3) == 0 && ((reinterpret_cast<uintptr_t>(src) & 3) == 0)) {
            // 4 pixels per thread: coalesced 4-byte loads, one conflict-free 16-byte shared store (lane i -> bytes 16 i)
            for (int i = threadIdx.x; i < Win / 4; i += blockDim.x) {
                const uint32_t w = __ldg(reinterpret_cast<const uint32_t *>(src) + i);
                reinterpret_cast<float4 *>(srow)[i] = make_float4((float)(w & 0xff), (float)((w >> 8) & 0xff), (float)((w >> 16) & 0xff),
                                                                  (float)(w >> 24));
            }
        } else {
            for (int i = threadIdx.x; i < Win; i += blockDim.x) srow[i] = (float)src[i];
        }
        __syncthreads();
        for (int ox = threadIdx.x; ox < Wout; ox += blockDim.x) {
            const float *sp = srow + __ldg(start + ox);
            const float *w = wts + ox;                       // transposed table [T][Wout]: coalesced across the warp
            float acc = 0.f;
            for (int t = 0; t < T; ++t) acc = fmaf(__ldg(w + (long long)t * Wout), sp[t], acc);
            out[row * Wout + ox] = acc;
        }
        __syncthreads();
    }
}

// Horizontal pass, register-resident weights: one thread per OUTPUT COLUMN keeps its kMaxT weights in registers and walks
// the rows of its block, so the inner loop is one shared-memory load + one FMA per tap (the table is read once per block
// instead of once per row).  Rows are double-buffered in shared memory: row r+1 is staged while row r is computed.
template <int kMaxT>
__global__ void resample_h_regw_kernel(const uint8_t *__restrict__ in, float *__restrict__ out, long long rows, int Win,
                                       int Wout, const int *__restrict__ start, const float *__restrict__ wts, int T) {
    extern __shared__ float srow[];          // [2][Win + kMaxT]; the kMaxT floats behind each row stay zero
    const int rowlen = Win + kMaxT;
    for (int i = threadIdx.x; i < 2 * kMaxT; i += blockDim.x) srow[(i / kMaxT) * rowlen + Win + (i % kMaxT)] = 0.f;
    const int ox = threadIdx.x;
    const bool active = ox < Wout;
    float w[kMaxT];
    int s0 = 0;
    if (active) s0 = __ldg(start + ox);
#pragma unroll
    for (int t = 0; t < kMaxT; ++t) w[t] = (active && t < T) ? __ldg(wts + (long long)t * Wout + ox) : 0.f;
    // taps t >= T carry zero weights and read either real pixels or the zero pad behind the row
    auto stage = [&](long long row, float *dst) {
        const uint8_t *src = in + row * Win;
        for (int i = threadIdx.x; i < Win / 4; i += blockDim.x) {
            const uint32_t v = __ldg(reinterpret_cast<const uint32_t *>(src) + i);
            reinterpret_cast<float4 *>(dst)[i] = make_float4((float)(v & 0xff), (float)((v >> 8) & 0xff), (float)((v >> 16) & 0xff),
                                                             (float)(v >> 24));
        }
    };
    const long long r0 = blockIdx.x;
    if (r0 < rows) stage(r0, srow);
    __syncthreads();
    int buf = 0;
    for (long long row = r0; row < rows; row += gridDim.x) {
        const long long nxt = row + gridDim.x;
        if (nxt < rows) stage(nxt, srow + (buf ^ 1) * rowlen);
        if (active) {
            const float *sp = srow + buf * rowlen + s0;
            float acc = 0.f;
#pragma unroll
            for (int t = 0; t < kMaxT; ++t) acc = fmaf(w[t], sp[t], acc);
            out[row * Wout + ox] = acc;
        }
        __syncthreads();
        buf ^= 1;
    }
}

// Horizontal pass, register-resident weights, kR rows per step: like resample_h_regw_kernel, but a block stages kR rows at a
// time and the NEXT group's pixels are already in flight (registers) while the current group is computed, so the global-load
// latency never sits on the critical path; each weight register feeds kR independent FMAs.  One block per SM (the two
// kR-row buffers take most of the shared memory), bound by the one-LDS-per-FMA rate of the inner loop.
template <int kMaxT, int kR>
__global__ void __launch_bounds__(640, 1)
resample_h_rows_kernel(const uint8_t *__restrict__ in, float *__restrict__ out, long long rows, int Win, int Wout,
                       const int *__restrict__ start, const float *__restrict__ wts, int T) {
    extern __shared__ float srow[];          // [2][kR][Win + kMaxT]; the kMaxT floats behind each row stay zero
    constexpr int kPre = 6;                  // uint32 (4-pixel) loads in flight per thread
    const int rowlen = Win + kMaxT;
    const int w4 = Win / 4;                  // 4-pixel words per row
    const int words = kR * w4;               // per group
    for (int i = threadIdx.x; i < 2 * kR * kMaxT; i += blockDim.x) srow[(i / kMaxT) * rowlen + Win + (i % kMaxT)] = 0.f;
    const int ox = threadIdx.x;
    const bool active = ox < Wout;
    float w[kMaxT];
    int s0 = 0;
    if (active) s0 = __ldg(start + ox);
#pragma unroll
    for (int t = 0; t < kMaxT; ++t) w[t] = (active && t < T) ? __ldg(wts + (long long)t * Wout + ox) : 0.f;
    const long long groups = (rows + kR - 1) / kR;
    uint32_t pre[kPre];
    auto fetch = [&](long long g) {          // rows of a group are consecutive in memory: one flat run of `words` words
        const uint32_t *src = reinterpret_cast<const uint32_t *>(in + g * kR * (long long)Win);
        const long long lim = (rows - g * kR) * w4;      // words that exist (last group may be short)
#pragma unroll
        for (int k = 0; k < kPre; ++k) {
            const int i = threadIdx.x + k * blockDim.x;
            pre[k] = (i < words && i < lim) ? __ldg(src + i) : 0u;
        }
    };
    auto stash = [&](float *dst) {
#pragma unroll
        for (int k = 0; k < kPre; ++k) {
            const int i = threadIdx.x + k * blockDim.x;
            if (i < words) {
                const int r = i / w4, c = i - r * w4;
                const uint32_t v = pre[k];
                *reinterpret_cast<float4 *>(dst + r * rowlen + 4 * c) =
                    make_float4((float)(v & 0xff), (float)((v >> 8) & 0xff), (float)((v >> 16) & 0xff), (float)(v >> 24));
            }
        }
    };
    long long g = blockIdx.x;
    if (g < groups) { fetch(g); stash(srow); }
    __syncthreads();
    int buf = 0;
    for (; g < groups; g += gridDim.x) {
        const long long nxt = g + gridDim.x;
        if (nxt < groups) fetch(nxt);
        if (active) {
            const float *sp = srow + buf * (kR * rowlen) + s0;
            float acc[kR];
#pragma unroll
            for (int r = 0; r < kR; ++r) acc[r] = 0.f;
#pragma unroll
            for (int t = 0; t < kMaxT; ++t)
#pragma unroll
                for (int r = 0; r < kR; ++r) acc[r] = fmaf(w[t], sp[r * rowlen + t], acc[r]);
#pragma unroll
            for (int r = 0; r < kR; ++r)
                if (g * kR + r < rows) out[(g * kR + r) * Wout + ox] = acc[r];
        }
        if (nxt < groups) stash(srow + (buf ^ 1) * (kR * rowlen));
        __syncthreads();
        buf ^= 1;
    }
}

// Vertical pass on u8 planes: out[plane][oy][x] = sum_t w[oy][t] * in[plane][start[oy]+t][x]  (float out).
__global__ void resample_v_kernel(const uint8_t *__restrict__ in, float *__restrict__ out, long long planes, int Hin,
                                  int Hout, int W, const int *__restrict__ start, const float *__restrict__ wts, int T) {
    const long long total = planes * Hout * W;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % W);
        const int oy = (int)((i / W) % Hout);
        const long long pl = i / ((long long)W * Hout);
        const uint8_t *src = in + (pl * Hin + __ldg(start + oy)) * W + x;
        const float *w = wts + (long long)oy * T;
        float acc = 0.f;
        for (int t = 0; t < T; ++t) acc = fmaf(__ldg(w + t), (float)__ldg(src + (long long)t * W), acc);
        out[i] = acc;
    }
}

// Same pass, four adjacent columns per thread (W % 4 == 0): one 4-byte load per tap instead of four 1-byte loads, a
// warp-uniform weight load, one 16-byte store.
__global__ void resample_v4_kernel(const uint8_t *__restrict__ in, float4 *__restrict__ out, long long planes, int Hin,
                                   int Hout, int W, const int *__restrict__ start, const float *__restrict__ wts, int T) {
    const int W4 = W >> 2;
    const long long total = planes * Hout * W4;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x4 = (int)(i % W4);
        const int oy = (int)((i / W4) % Hout);
        const long long pl = i / ((long long)W4 * Hout);
        const uint32_t *src = reinterpret_cast<const uint32_t *>(in + (pl * Hin + __ldg(start + oy)) * W) + x4;
        const float *w = wts + (long long)oy * T;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        for (int t = 0; t < T; ++t) {
            const float wt = __ldg(w + t);
            const uint32_t v = __ldg(src + (long long)t * W4);
            a0 = fmaf(wt, (float)(v & 0xff), a0);
            a1 = fmaf(wt, (float)((v >> 8) & 0xff), a1);
            a2 = fmaf(wt, (float)((v >> 16) & 0xff), a2);
            a3 = fmaf(wt, (float)(v >> 24), a3);
        }
        out[i] = make_float4(a0, a1, a2, a3);
    }
}

// round half to even and clamp to [0, 255] in one conversion (float -> integer conversions saturate; NaN -> 0): the value
// sat8(__float2int_rn(v)) of the two-step form
__device__ __forceinline__ int round_u8(float v) {
    uint32_t r;
    asm("cvt.rni.u8.f32 %0, %1;" : "=r"(r) : "f"(v));
    return (int)r;
}

// Vertical pass of the pre-resize fused with gray conversion + ImageNet normalisation.
// in: float [B][3][Hin][S]; out: rgb_small u8 [B][3][S][S]; x: 16-bit [B][S][S][8] (channels 3..7 = 0).
__global__ void pre_vertical_kernel(const float *__restrict__ in, uint8_t *__restrict__ rgb_small, void *__restrict__ x,
                                    int B, int Hin, int S, const int *__restrict__ start,
                                    const float *__restrict__ wts, int T, int dtype) {
    const long long total = (long long)B * S * S;
    const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int ox = (int)(i % S);
        const int oy = (int)((i / S) % S);
        const int b = (int)(i / ((long long)S * S));
        const int s0 = __ldg(start + oy);
        const float *w = wts + (long long)oy * T;
        int q[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float *src = in + (((long long)b * 3 + c) * Hin + s0) * S + ox;
            float acc = 0.f;
            for (int t = 0; t < T; ++t) acc = fmaf(__ldg(w + t), __ldg(src + (long long)t * S), acc);
            q[c] = round_u8(acc);
            rgb_small[(((long long)b * 3 + c) * S + oy) * S + ox] = (uint8_t)q[c];
        }
        // Pillow convert('L'): (19595 R + 38470 G + 7471 B + 0x8000) >> 16, replicated to 3 channels
        const int L = (19595 * q[0] + 38470 * q[1] + 7471 * q[2] + 0x8000) >> 16;
        const float lf = __fdiv_rn((float)L, 255.f);
        float n[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) n[c] = __fdiv_rn(__fsub_rn(lf, mean[c]), stdv[c]);
        uint4 pk = make_uint4(pack2(n[0], n[1], dtype), pack2(n[2], 0.f, dtype), pack2((float)L, 1.f, dtype), 0u);   // ch 4,5: (L, 1) exact
        *reinterpret_cast<uint4 *>(reinterpret_cast<uint16_t *>(x) + i * 8) = pk;
    }
}

// Same pass, four adjacent output columns per thread (S % 4 == 0): 16-byte loads of the fp32 rows, warp-uniform weights.
__global__ void pre_vertical4_kernel(const float *__restrict__ in, uint8_t *__restrict__ rgb_small, void *__restrict__ x,
                                     int B, int Hin, int S, const int *__restrict__ start,
                                     const float *__restrict__ wts, int T, int dtype) {
    const int S4 = S >> 2;
    const long long total = (long long)B * S * S4;
    const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x4 = (int)(i % S4);
        const int oy = (int)((i / S4) % S);
        const int b = (int)(i / ((long long)S4 * S));
        const int s0 = __ldg(start + oy);
        const float *w = wts + (long long)oy * T;
        int q[3][4];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float4 *src = reinterpret_cast<const float4 *>(in + (((long long)b * 3 + c) * Hin + s0) * S) + x4;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
            for (int t = 0; t < T; ++t) {
                const float wt = __ldg(w + t);
                const float4 v = __ldg(src + (long long)t * S4);
                a0 = fmaf(wt, v.x, a0); a1 = fmaf(wt, v.y, a1); a2 = fmaf(wt, v.z, a2); a3 = fmaf(wt, v.w, a3);
            }
            q[c][0] = round_u8(a0); q[c][1] = round_u8(a1); q[c][2] = round_u8(a2); q[c][3] = round_u8(a3);
            *reinterpret_cast<uint32_t *>(rgb_small + (((long long)b * 3 + c) * S + oy) * S + 4 * x4) =
                (uint32_t)q[c][0] | ((uint32_t)q[c][1] << 8) | ((uint32_t)q[c][2] << 16) | ((uint32_t)q[c][3] << 24);
        }
        uint4 *xo = reinterpret_cast<uint4 *>(reinterpret_cast<uint16_t *>(x) + (((long long)b * S + oy) * S + 4 * x4) * 8);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int L = (19595 * q[0][k] + 38470 * q[1][k] + 7471 * q[2][k] + 0x8000) >> 16;   // Pillow convert('L')
            const float lf = __fdiv_rn((float)L, 255.f);
            float n[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) n[c] = __fdiv_rn(__fsub_rn(lf, mean[c]), stdv[c]);
            xo[k] = make_uint4(pack2(n[0], n[1], dtype), pack2(n[2], 0.f, dtype), pack2((float)L, 1.f, dtype), 0u);
        }
    }
}

// Vertical pass float rows -> u8 (second pass of a horizontal-first resize to a rectangular size, e.g. resize_min_HW's Spline36):
// in float [planes][Hin][W] -> out u8 [planes][Hout][W], round half to even + clamp.
__global__ void resample_v_f32_u8_kernel(const float *__restrict__ in, uint8_t *__restrict__ out, long long planes, int Hin, int Hout,
                                         int W, const int *__restrict__ start, const float *__restrict__ wts, int T) {
    const long long total = planes * Hout * W;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ox = (int)(i % W);
        const int oy = (int)((i / W) % Hout);
        const long long pl = i / ((long long)W * Hout);
        const float *src = in + (pl * Hin + __ldg(start + oy)) * W + ox;
        const float *w = wts + (long long)oy * T;
        float acc = 0.f;
        for (int t = 0; t < T; ++t) acc = fmaf(__ldg(w + t), __ldg(src + (long long)t * W), acc);
        out[i] = (uint8_t)round_u8(acc);
    }
}

// ColorizerFilter._transform (convert('LA').convert('RGB'), filters.py:92-93) + ImageNet normalisation (filters.py:50-53)
// of an already square u8 image: rgb u8 [B][3][n] -> x 16-bit [B][n][8].
__global__ void gray_normalize_kernel(const uint8_t *__restrict__ rgb, void *__restrict__ x, int B, long long n, int dtype) {
    const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
    const long long total = (long long)B * n;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / n, px = i - b * n;
        const uint8_t *q = rgb + b * 3 * n + px;
        const int L = pil_luma(__ldg(q), __ldg(q + n), __ldg(q + 2 * n));
        const float lf = __fdiv_rn((float)L, 255.f);
        float v[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) v[c] = __fdiv_rn(__fsub_rn(lf, mean[c]), stdv[c]);
        *reinterpret_cast<uint4 *>(reinterpret_cast<uint16_t *>(x) + i * 8) = make_uint4(pack2(v[0], v[1], dtype), pack2(v[2], 0.f, dtype), pack2((float)L, 1.f, dtype), 0u);
    }
}

// Shared tail of the head: logits -> SigmoidRange(-3,3) -> de-normalise -> clamp -> *255 -> truncate -> (skip |
// S x S luma transplant) -> colored u8 planes.
__device__ __forceinline__ void head_finish_pixel(const float lg[3], long long pix, const uint8_t *__restrict__ rgb_small,
                                                  uint8_t *__restrict__ colored, float *__restrict__ net_out,
                                                  const uint8_t *__restrict__ skip, int S, int transplant) {
    const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
    const int ox = (int)(pix % S);
    const int oy = (int)((pix / S) % S);
    const int b = (int)(pix / ((long long)S * S));
    int q[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float sg = 1.f / (1.f + expf(-lg[c]));
        const float y = __fadd_rn(__fmul_rn(sg, 6.f), -3.f);                 // SigmoidRange(-3,3)
        if (net_out) net_out[(((long long)b * 3 + c) * S + oy) * S + ox] = y;
        float d = __fadd_rn(__fmul_rn(y, stdv[c]), mean[c]);                 // denorm
        d = fminf(fmaxf(d, 0.f), 1.f);                                       // reconstruct: clamp(0,1)
        q[c] = (int)__fmul_rn(d, 255.f);                                     // astype(uint8): truncation
    }
    int r = q[0], g = q[1], bl = q[2];
    const long long o = ((long long)b * 3 * S + oy) * S + ox;
    if (skip != nullptr && skip[b]) {   // scene-change gate: the selector returned the frame unchanged
        r = rgb_small[o]; g = rgb_small[o + (long long)S * S]; bl = rgb_small[o + 2ll * S * S];
    } else if (transplant) {
        luma_transplant(rgb_small[o], rgb_small[o + (long long)S * S], rgb_small[o + 2ll * S * S], q[0], q[1], q[2], r, g,
                        bl);
    }
    colored[o] = (uint8_t)r;
    colored[o + (long long)S * S] = (uint8_t)g;
    colored[o + 2ll * S * S] = (uint8_t)bl;
}

// Head on pre-computed logits (the 1x1 conv ran inside the last GEMM's epilogue): fp32 [B*S*S][4], one thread/pixel.
__global__ void head_from_logits_kernel(const float4 *__restrict__ logits, const float *__restrict__ b11,
                                        const uint8_t *__restrict__ rgb_small, uint8_t *__restrict__ colored,
                                        float *__restrict__ net_out, const uint8_t *__restrict__ skip, int B, int S,
                                        int transplant) {
    const long long total = (long long)B * S * S;
    const float b0 = __ldg(b11), b1 = __ldg(b11 + 1), b2 = __ldg(b11 + 2);
    for (long long pix = blockIdx.x * (long long)blockDim.x + threadIdx.x; pix < total;
         pix += (long long)gridDim.x * blockDim.x) {
        const float4 l = __ldg(logits + pix);
        const float lg[3] = {l.x + b0, l.y + b1, l.z + b2};
        head_finish_pixel(lg, pix, rgb_small, colored, net_out, skip, S, transplant);
    }
}

// Output head: one warp per pixel.  logits = W11 . res + b; y = sigmoid*6-3; rgb = trunc(clamp(y*std+mean)*255);
// then the S x S luma transplant against the resized source.  Optionally dumps the fp32 net output.
__global__ void head_kernel(const void *__restrict__ res, int Cs, const float *__restrict__ w11 /*[3][Cs]*/,
                            const float *__restrict__ b11, const uint8_t *__restrict__ rgb_small,
                            uint8_t *__restrict__ colored, float *__restrict__ net_out,
                            const uint8_t *__restrict__ skip, int B, int S, int dtype, int transplant) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long total = (long long)B * S * S;
    // each lane owns channel groups g = lane, lane+32 (Cs <= 512): its slice of the 1x1 weights stays in registers
    constexpr int kG = 2;
    float wr[kG][3][8];
#pragma unroll
    for (int k = 0; k < kG; ++k) {
        const int g = lane + 32 * k;
#pragma unroll
        for (int o = 0; o < 3; ++o)
#pragma unroll
            for (int j = 0; j < 8; ++j) wr[k][o][j] = (g < Cs / 8) ? __ldg(w11 + o * Cs + g * 8 + j) : 0.f;
    }
    for (long long pix = warp0; pix < total; pix += nwarps) {
        const uint4 *row = reinterpret_cast<const uint4 *>(reinterpret_cast<const uint16_t *>(res) + pix * Cs);
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
        for (int k = 0; k < kG; ++k) {
            const int g = lane + 32 * k;
            if (g >= Cs / 8) break;
            const uint4 v = __ldg(row + g);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = unpack2(w[j], dtype);
                a0 = fmaf(f.x, wr[k][0][2 * j], a0); a0 = fmaf(f.y, wr[k][0][2 * j + 1], a0);
                a1 = fmaf(f.x, wr[k][1][2 * j], a1); a1 = fmaf(f.y, wr[k][1][2 * j + 1], a1);
                a2 = fmaf(f.x, wr[k][2][2 * j], a2); a2 = fmaf(f.y, wr[k][2][2 * j + 1], a2);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a0 += __shfl_xor_sync(0xffffffffu, a0, o);
            a1 += __shfl_xor_sync(0xffffffffu, a1, o);
            a2 += __shfl_xor_sync(0xffffffffu, a2, o);
        }
        if (lane == 0) {
            const float lg[3] = {a0 + __ldg(b11), a1 + __ldg(b11 + 1), a2 + __ldg(b11 + 2)};
            head_finish_pixel(lg, pix, rgb_small, colored, net_out, skip, S, transplant);
        }
    }
}

// Final (horizontal) pass of the resize back to W x H fused with the full-resolution luma transplant.
// in: float [B][3][H][S] (already resampled vertically); orig/out: u8 [B][3][H][W].  One block per output row:
// the three S-wide source rows sit in shared memory; consecutive threads produce consecutive pixels, so the
// transposed weight table, the start table and the u8 planes are all read/written coalesced.
__global__ void post_horizontal_kernel(const float *__restrict__ in, const uint8_t *__restrict__ orig,
                                       uint8_t *__restrict__ out, int B, int S, int H, int W,
                                       const int *__restrict__ start, const float *__restrict__ wts, int T,
                                       int transplant, int vec_tables) {
    extern __shared__ float srows[];   // [3][S]
    const long long ps = (long long)H * W;
    for (long long r = blockIdx.x; r < (long long)B * H; r += gridDim.x) {
        const int oy = (int)(r % H);
        const int b = (int)(r / H);
        for (int i = threadIdx.x; i < 3 * S; i += blockDim.x) {
            const int c = i / S, x = i - c * S;
            srows[i] = __ldg(in + (((long long)b * 3 + c) * H + oy) * S + x);
        }
        __syncthreads();
        const long long base = ((long long)b * 3 * H + oy) * W;
        auto pixel = [&](int ox, int o_r, int o_g, int o_b, int &rr, int &gg, int &bb) {
            const int s0 = __ldg(start + ox);
            const float *w = wts + ox;                       // transposed table [T][W]
            float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f;
            for (int t = 0; t < T; ++t) {
                const float wt = __ldg(w + (long long)t * W);
                acc0 = fmaf(wt, srows[s0 + t], acc0);
                acc1 = fmaf(wt, srows[S + s0 + t], acc1);
                acc2 = fmaf(wt, srows[2 * S + s0 + t], acc2);
            }
            const int q0 = round_u8(acc0), q1 = round_u8(acc1), q2 = round_u8(acc2);
            rr = q0; gg = q1; bb = q2;
            if (transplant) luma_transplant(o_r, o_g, o_b, q0, q1, q2, rr, gg, bb);
        };
        if ((W & 3) == 0 && transplant && ((reinterpret_cast<uintptr_t>(orig) | reinterpret_cast<uintptr_t>(out)) & 3) == 0) {
            // 4 pixels per thread: 4-byte loads of the original planes and 4-byte stores (128 B per warp and plane)
            for (int x4 = threadIdx.x; x4 < W / 4; x4 += blockDim.x) {
                const int ox = 4 * x4;
                const uint32_t o0 = __ldg(reinterpret_cast<const uint32_t *>(orig + base + ox));
                const uint32_t o1 = __ldg(reinterpret_cast<const uint32_t *>(orig + base + ps + ox));
                const uint32_t o2 = __ldg(reinterpret_cast<const uint32_t *>(orig + base + 2 * ps + ox));
                uint32_t p0 = 0, p1 = 0, p2 = 0;
                if (vec_tables && T <= 8) {
                    // the four pixels' weights arrive as ONE 16-byte load per tap and their window starts as one int4 (the
                    // per-pixel path issues 36 scalar table loads per item and the pass is L1 / load-store bound); same FMA order
                    const int4 s4 = __ldg(reinterpret_cast<const int4 *>(start) + x4);
                    const int s0k[4] = {s4.x, s4.y, s4.z, s4.w};
                    float wk[8][4];
#pragma unroll
                    for (int t = 0; t < 8; ++t) {
                        const float4 w4 = t < T ? __ldg(reinterpret_cast<const float4 *>(wts + (long long)t * W) + x4) : make_float4(0.f, 0.f, 0.f, 0.f);
                        wk[t][0] = w4.x; wk[t][1] = w4.y; wk[t][2] = w4.z; wk[t][3] = w4.w;
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f;
                        const float *sp = srows + s0k[k];
#pragma unroll
                        for (int t = 0; t < 8; ++t) {
                            if (t < T) {
                                acc0 = fmaf(wk[t][k], sp[t], acc0);
                                acc1 = fmaf(wk[t][k], sp[S + t], acc1);
                                acc2 = fmaf(wk[t][k], sp[2 * S + t], acc2);
                            }
                        }
                        const int q0 = round_u8(acc0), q1 = round_u8(acc1), q2 = round_u8(acc2);
                        int rr = q0, gg = q1, bb = q2;
                        luma_transplant((o0 >> (8 * k)) & 0xff, (o1 >> (8 * k)) & 0xff, (o2 >> (8 * k)) & 0xff, q0, q1, q2, rr, gg, bb);
                        p0 |= (uint32_t)rr << (8 * k); p1 |= (uint32_t)gg << (8 * k); p2 |= (uint32_t)bb << (8 * k);
                    }
                } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    int rr, gg, bb;
                    pixel(ox + k, (o0 >> (8 * k)) & 0xff, (o1 >> (8 * k)) & 0xff, (o2 >> (8 * k)) & 0xff, rr, gg, bb);
                    p0 |= (uint32_t)rr << (8 * k); p1 |= (uint32_t)gg << (8 * k); p2 |= (uint32_t)bb << (8 * k);
                }
                }
                *reinterpret_cast<uint32_t *>(out + base + ox) = p0;
                *reinterpret_cast<uint32_t *>(out + base + ps + ox) = p1;
                *reinterpret_cast<uint32_t *>(out + base + 2 * ps + ox) = p2;
            }
        } else {
            for (int ox = threadIdx.x; ox < W; ox += blockDim.x) {
                int rr, gg, bb, o_r = 0, o_g = 0, o_b = 0;
                if (transplant) { o_r = __ldg(orig + base + ox); o_g = __ldg(orig + base + ps + ox); o_b = __ldg(orig + base + 2 * ps + ox); }
                pixel(ox, o_r, o_g, o_b, rr, gg, bb);
                out[base + ox] = (uint8_t)rr;
                out[base + ps + ox] = (uint8_t)gg;
                out[base + 2 * ps + ox] = (uint8_t)bb;
            }
        }
        __syncthreads();
    }
}

// PIL.Image.blend(a, b, alpha) on u8 data: trunc(a + alpha*(b - a)) in float32 (Pillow's Blend.c), 16 bytes/thread.
// `out` may alias `a` or `b` (the engines blend in place): no __restrict__ / non-coherent loads; element i is read before it
// is written and no other thread touches it.
__global__ void blend_u8_kernel(const uint4 *a, const uint4 *b, uint4 *out, long long n16, float alpha) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n16; i += (long long)gridDim.x * blockDim.x) {
        const uint4 va = a[i], vb = b[i];
        const uint32_t wa[4] = {va.x, va.y, va.z, va.w}, wb[4] = {vb.x, vb.y, vb.z, vb.w};
        uint32_t wo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t o = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float fa = (float)((wa[j] >> (8 * k)) & 0xff), fb = (float)((wb[j] >> (8 * k)) & 0xff);
                const float t = __fadd_rn(fa, __fmul_rn(alpha, __fsub_rn(fb, fa)));
                o |= (uint32_t)sat8((int)t) << (8 * k);
            }
            wo[j] = o;
        }
        out[i] = make_uint4(wo[0], wo[1], wo[2], wo[3]);
    }
}

// ---- phase-periodic horizontal passes ------------------------------------------------------------------------------------
// When the resampling ratio is an integer (1920 <-> 384, 1920 <-> 480, 3840 <-> 640) every interior output column of a phase has
// the SAME weights and its window moves by a fixed step: the table lookups disappear (the weights sit in the kernel's parameter
// bank and feed the FMAs as constant operands) and one thread keeps a run of input pixels in registers for several outputs, so
// the one-shared-memory-load-per-FMA bound of the table kernels (profiles/r01_ncu_post_horizontal_b32.txt) turns into an
// FMA / ALU bound.  The host checks the periodicity of the table bit for bit (resample.periodic_plan) and passes the interior
// range; the few border columns, whose windows are folded at the image edge, go through the table path inside the same
// kernel.  Every output is the same fmaf chain over the taps in the same order as the table kernels: bit-identical results.
struct PeriodicW { float w[72]; };

// Down-scaling (the squeeze): u8 rows -> float rows, ratio kRt = Win / Wout.  A thread owns 8 adjacent output columns of one
// row: its inputs are kN aligned 16-byte shared-memory loads, its 8 x kTP FMAs take their weights from the parameter bank
// (pw.w = the kT taps preceded by `sh` zeros, sh = window start mod 4, so that the run starts on a 16-byte boundary).
// Small blocks (2 rows each), many per SM: the phases of one block (stage, compute, border) overlap with the other blocks'.
template <int kTP, int kRt>
__global__ void __launch_bounds__(160)
resample_h_periodic_kernel(const uint8_t *__restrict__ in, float *__restrict__ out, long long rows, int Win, int Wout,
                           const int *__restrict__ start, const float *__restrict__ wts, int T,
                           const __grid_constant__ PeriodicW pw, int a_off, int lo8, int hi8) {
    extern __shared__ float srow[];          // [kRows][Win + 64] (the 64 floats behind each row stay zero), then the border table
    constexpr int kRows = 2;
    constexpr int kSpan = kRt * 7 + kTP;     // input floats behind one thread's 8 outputs
    constexpr int kN = (kSpan + 3) / 4;
    const int rowlen = Win + 64;
    const int w4 = Win / 4;
    const int words = kRows * w4;
    const int nblk = hi8 - lo8;                               // interior column blocks (8 columns each)
    const int nborder = 8 * lo8 + (Wout - 8 * hi8);           // border columns per row
    float *bw = srow + kRows * rowlen;                        // [nborder][T] weights of the border columns
    int *bs = reinterpret_cast<int *>(bw + nborder * T);      // [nborder] window starts
    for (int i = threadIdx.x; i < kRows * 64; i += blockDim.x) srow[(i / 64) * rowlen + Win + (i % 64)] = 0.f;
    for (int i = threadIdx.x; i < nborder * T; i += blockDim.x) {
        const int k = i / T, t = i - k * T;
        const int ox = k < 8 * lo8 ? k : 8 * hi8 + (k - 8 * lo8);
        bw[i] = __ldg(wts + (long long)t * Wout + ox);
        if (t == 0) bs[k] = __ldg(start + ox);
    }
    const int ur = threadIdx.x / nblk, ucb = lo8 + threadIdx.x % nblk;
    const bool unit = threadIdx.x < kRows * nblk;
    const int bit = (int)blockDim.x - 1 - (int)threadIdx.x;   // border item of this thread (the last warp's lanes take them)
    const long long groups = (rows + kRows - 1) / kRows;
    for (long long g = blockIdx.x; g < groups; g += gridDim.x) {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(in + g * kRows * (long long)Win);
        const long long lim = (rows - g * kRows) * w4;      // words that exist (the last group may be short)
        __syncthreads();                                     // the previous group's readers are done
        for (int i = threadIdx.x; i < words; i += blockDim.x) {
            const uint32_t v = i < lim ? __ldg(src + i) : 0u;
            const int r = i >= w4 ? 1 : 0, c = i - r * w4;      // kRows == 2
            *reinterpret_cast<float4 *>(srow + r * rowlen + 4 * c) =
                make_float4((float)(v & 0xff), (float)((v >> 8) & 0xff), (float)((v >> 16) & 0xff), (float)(v >> 24));
        }
        __syncthreads();
        if (unit) {
            const float4 *sp = reinterpret_cast<const float4 *>(srow + ur * rowlen + a_off + kRt * 8 * ucb);
            float x[4 * kN];
#pragma unroll
            for (int k = 0; k < kN; ++k) {
                const float4 v = sp[k];
                x[4 * k] = v.x; x[4 * k + 1] = v.y; x[4 * k + 2] = v.z; x[4 * k + 3] = v.w;
            }
            float acc[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
            for (int t = 0; t < kTP; ++t)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] = fmaf(pw.w[t], x[kRt * j + t], acc[j]);
            const long long row = g * kRows + ur;
            if (row < rows) {
                float4 *op = reinterpret_cast<float4 *>(out + row * Wout + 8 * ucb);
                op[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
                op[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
            }
        }
        for (int it = bit; it < kRows * nborder; it += blockDim.x) {      // border columns: the table path (table in shared memory)
            const int r = it / nborder, k = it - r * nborder;
            const int ox = k < 8 * lo8 ? k : 8 * hi8 + (k - 8 * lo8);
            const long long row = g * kRows + r;
            if (row < rows) {
                const float *sp = srow + r * rowlen + bs[k];
                const float *wk = bw + k * T;
                float acc = 0.f;
                for (int t = 0; t < T; ++t) acc = fmaf(wk[t], sp[t], acc);
                out[row * Wout + ox] = acc;
            }
        }
    }
}

// Up-scaling (the way back) fused with the full-resolution luma transplant, ratio kRt = W / S.  A thread owns 4 input
// positions = 4 kRt output pixels of one row: 3 + kTP input floats per channel in registers (kTP = taps + the spread of the
// per-phase window starts; pw.w[p * kTP + t] = the weights of phase p, zero-padded to the common window), the original pixels
// read as 32-bit words, the result staged per warp in shared memory and written back as full 128-byte lines.  One row per block
// and iteration, several blocks per SM.
template <int kRt, int kTP>
__global__ void __launch_bounds__(160)
post_horizontal_periodic_kernel(const float *__restrict__ in, const uint8_t *__restrict__ orig, uint8_t *__restrict__ out, int B, int S,
                                int H, int W, const int *__restrict__ start, const float *__restrict__ wts, int T, int transplant,
                                const __grid_constant__ PeriodicW pw, int base_off, int ulo, int uhi) {
    extern __shared__ float sm_f[];
    constexpr int kX = 3 + kTP;                 // inputs behind one thread's 4 positions
    constexpr int kN = (kX + 3) / 4;
    constexpr int kWords = kRt;                 // 4 kRt output bytes per plane = kRt words per thread
    const int spitch = S + 16;                  // row pitch in floats (data at +base_off, zeros around it)
    const int upr = S / 4;                      // units per row
    const int nbu = ulo + (upr - uhi);          // border units per row
    const int nbpx = nbu * 4 * kRt;             // border pixels per row
    float *srows = sm_f;                                                        // [3][spitch]
    uint32_t *stage = reinterpret_cast<uint32_t *>(sm_f + 3 * spitch);          // [warps][3][32 kRt]
    float *bw = reinterpret_cast<float *>(stage + (blockDim.x >> 5) * 3 * 32 * kWords);   // [nbpx][T]
    int *bs = reinterpret_cast<int *>(bw + nbpx * T);                           // [nbpx] window starts
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int u = threadIdx.x;
    const long long ps = (long long)H * W;
    const long long nrows = (long long)B * H;
    uint32_t *stg = stage + warp * (3 * 32 * kWords);
    for (int i = threadIdx.x; i < 3 * spitch; i += blockDim.x) srows[i] = 0.f;
    for (int i = threadIdx.x; i < nbpx * T; i += blockDim.x) {
        const int k = i / T, t = i - k * T;
        const int bu = k / (4 * kRt), uu = bu < ulo ? bu : uhi + (bu - ulo);
        const int ox = uu * 4 * kRt + (k - bu * 4 * kRt);
        bw[i] = __ldg(wts + (long long)t * W + ox);
        if (t == 0) bs[k] = __ldg(start + ox);
    }
    const int bit = (int)blockDim.x - 1 - (int)threadIdx.x;
    for (long long r = blockIdx.x; r < nrows; r += gridDim.x) {
        const int oy = (int)(r % H), b = (int)(r / H);
        __syncthreads();
        if ((base_off & 3) == 0) {
            for (int i = threadIdx.x; i < 3 * upr; i += blockDim.x) {
                const int c = i / upr, x4 = i - c * upr;
                const float4 v = __ldg(reinterpret_cast<const float4 *>(in + (((long long)b * 3 + c) * H + oy) * S) + x4);
                *reinterpret_cast<float4 *>(srows + c * spitch + base_off + 4 * x4) = v;
            }
        } else {
            for (int i = threadIdx.x; i < 3 * S; i += blockDim.x) {
                const int c = i / S, x = i - c * S;
                srows[c * spitch + base_off + x] = __ldg(in + (((long long)b * 3 + c) * H + oy) * S + x);
            }
        }
        __syncthreads();
        const long long base = ((long long)b * 3 * H + oy) * W;
        if (u >= ulo && u < uhi) {
            uint32_t ow[3][kWords];
#pragma unroll
            for (int pl = 0; pl < 3; ++pl)
#pragma unroll
                for (int k = 0; k < kWords; ++k)
                    ow[pl][k] = transplant ? __ldg(reinterpret_cast<const uint32_t *>(orig + base + pl * ps) + u * kWords + k) : 0u;
            float x[3][4 * kN];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float4 *sp = reinterpret_cast<const float4 *>(srows + c * spitch + 4 * u);
#pragma unroll
                for (int k = 0; k < kN; ++k) {
                    const float4 v = sp[k];
                    x[c][4 * k] = v.x; x[c][4 * k + 1] = v.y; x[c][4 * k + 2] = v.z; x[c][4 * k + 3] = v.w;
                }
            }
            uint32_t res[3][kWords];
#pragma unroll
            for (int pl = 0; pl < 3; ++pl)
#pragma unroll
                for (int k = 0; k < kWords; ++k) res[pl][k] = 0u;
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int p = 0; p < kRt; ++p) {
                    const int j = q * kRt + p;                    // pixel within the unit
                    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
                    for (int t = 0; t < kTP; ++t) {
                        const float wt = pw.w[p * kTP + t];
                        a0 = fmaf(wt, x[0][q + t], a0);
                        a1 = fmaf(wt, x[1][q + t], a1);
                        a2 = fmaf(wt, x[2][q + t], a2);
                    }
                    const int q0 = round_u8(a0), q1 = round_u8(a1), q2 = round_u8(a2);
                    int rr = q0, gg = q1, bb = q2;
                    const int sh = 8 * (j & 3), wd = j >> 2;
                    if (transplant)
                        luma_transplant((ow[0][wd] >> sh) & 0xff, (ow[1][wd] >> sh) & 0xff, (ow[2][wd] >> sh) & 0xff, q0, q1, q2, rr, gg, bb);
                    res[0][wd] |= (uint32_t)rr << sh; res[1][wd] |= (uint32_t)gg << sh; res[2][wd] |= (uint32_t)bb << sh;
                }
#pragma unroll
            for (int pl = 0; pl < 3; ++pl)
#pragma unroll
                for (int k = 0; k < kWords; ++k) stg[pl * 32 * kWords + lane * kWords + k] = res[pl][k];
        }
        __syncwarp();
        {   // write-back of the warp's interior units as whole words, consecutive lanes -> consecutive words
            const int ufirst = max(ulo, warp * 32), ulast = min(uhi, min(upr, warp * 32 + 32));
            const int w0 = (ufirst - warp * 32) * kWords, w1 = (ulast - warp * 32) * kWords;
            for (int pl = 0; pl < 3; ++pl)
                for (int k = w0 + lane; k < w1; k += 32)
                    reinterpret_cast<uint32_t *>(out + base + pl * ps)[warp * 32 * kWords + k] = stg[pl * 32 * kWords + k];
        }
        // border pixels (windows folded at the image edge): the table path with the table in shared memory, one pixel per thread
        for (int it = bit; it < nbpx; it += blockDim.x) {
            const int bu = it / (4 * kRt), uu = bu < ulo ? bu : uhi + (bu - ulo);
            const int ox = uu * 4 * kRt + (it - bu * 4 * kRt);
            const float *sp = srows + base_off + bs[it];
            const float *wk = bw + it * T;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f;
            for (int t = 0; t < T; ++t) {
                a0 = fmaf(wk[t], sp[t], a0);
                a1 = fmaf(wk[t], sp[spitch + t], a1);
                a2 = fmaf(wk[t], sp[2 * spitch + t], a2);
            }
            const int q0 = round_u8(a0), q1 = round_u8(a1), q2 = round_u8(a2);
            int cr = q0, cg = q1, cb = q2;
            if (transplant) luma_transplant(__ldg(orig + base + ox), __ldg(orig + base + ps + ox), __ldg(orig + base + 2 * ps + ox), q0, q1, q2, cr, cg, cb);
            out[base + ox] = (uint8_t)cr;
            out[base + ps + ox] = (uint8_t)cg;
            out[base + 2 * ps + ox] = (uint8_t)cb;
        }
    }
}

}  // namespace havc

using namespace havc;

__global__ void blend_u8_tail_kernel(const uint8_t *a, const uint8_t *b, uint8_t *out, long long n, float alpha) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) out[i] = (uint8_t)pil_blend(a[i], b[i], alpha);
}

extern "C" int havc_blend_u8(const uint8_t *a, const uint8_t *b, uint8_t *out, long long n, float alpha, void *stream) {
    HAVC_CHECK_ARG(a && b && out && n > 0 && alpha >= 0.f && alpha <= 1.f, "havc_blend_u8: alpha must be in [0,1]");
    const bool aligned = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    const long long n16 = aligned ? n / 16 : 0;
    if (n16 > 0) {
        blend_u8_kernel<<<grid1d(n16, 256), 256, 0, (cudaStream_t)stream>>>((const uint4 *)a, (const uint4 *)b, (uint4 *)out, n16, alpha);
        HAVC_LAUNCHED();
    }
    const long long rest = n - 16 * n16;
    if (rest > 0) {   // unaligned buffers or the last n % 16 values
        blend_u8_tail_kernel<<<(unsigned)((rest + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a + 16 * n16, b + 16 * n16, out + 16 * n16,
                                                                                                rest, alpha);
        HAVC_LAUNCHED();
    }
    return HAVC_OK;
}

extern "C" int havc_resample_h(const uint8_t *in, float *out, long long rows, int Win, int Wout, const int *start,
                               const float *weights, int taps, void *stream) {
    HAVC_CHECK_ARG(in && out && start && weights && taps > 0, "havc_resample_h: bad arguments");
    HAVC_CHECK_ARG(Win * sizeof(float) <= 96 * 1024, "havc_resample_h: row too wide for shared memory");
    static const bool legacy_h = getenv("HAVC_B200_LEGACY_PIXEL") != nullptr;   // A/B switch for profiling
    constexpr int kR = 4;
    if (!legacy_h && Wout <= 640 && taps <= 48 && (Win & 3) == 0 && (reinterpret_cast<uintptr_t>(in) & 3) == 0 && rows >= kR) {
        const int kt = taps <= 8 ? 8 : (taps <= 24 ? 24 : (taps <= 40 ? 40 : 48));
        const int block = ((Wout + 31) / 32) * 32;
        const size_t sm = 2 * (size_t)kR * (Win + kt) * sizeof(float);
        if (sm <= 200 * 1024 && kR * (Win / 4) <= 6 * block) {
            static std::atomic<unsigned long long> attr3{0ull};
            unsigned long long attr3_bit;
            if (device_pending(attr3, &attr3_bit)) {
                HAVC_CHECK_CUDA(cudaFuncSetAttribute(resample_h_rows_kernel<8, kR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                HAVC_CHECK_CUDA(cudaFuncSetAttribute(resample_h_rows_kernel<24, kR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                HAVC_CHECK_CUDA(cudaFuncSetAttribute(resample_h_rows_kernel<40, kR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                HAVC_CHECK_CUDA(cudaFuncSetAttribute(resample_h_rows_kernel<48, kR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                device_done(attr3, attr3_bit);
            }
            const long long groups = (rows + kR - 1) / kR;
            const int g3 = (int)(groups < (long long)num_sms() ? groups : (long long)num_sms());
            cudaStream_t st = (cudaStream_t)stream;
            if (kt == 8) resample_h_rows_kernel<8, kR><<<g3, block, sm, st>>>(in, out, rows, Win, Wout, start, weights, taps);
            else if (kt == 24) resample_h_rows_kernel<24, kR><<<g3, block, sm, st>>>(in, out, rows, Win, Wout, start, weights, taps);
            else if (kt == 40) resample_h_rows_kernel<40, kR><<<g3, block, sm, st>>>(in, out, rows, Win, Wout, start, weights, taps);
            else resample_h_rows_kernel<48, kR><<<g3, block, sm, st>>>(in, out, rows, Win, Wout, start, weights, taps);
            HAVC_LAUNCHED();
            return HAVC_OK;
        }
    }
    if (Wout <= 512 && taps <= 48 && (Win & 3) == 0 && (reinterpret_cast<uintptr_t>(in) & 3) == 0 && 2 * (Win + 48) * sizeof(float) <= 96 * 1024) {
        static std::atomic<unsigned long long> attr2{0ull};
        unsigned long long attr2_bit;
        if (device_pending(attr2, &attr2_bit)) {
            HAVC_CHECK_CUDA(cudaFuncSetAttribute(resample_h_regw_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
            HAVC_CHECK_CUDA(cudaFuncSetAttribute(resample_h_regw_kernel<24>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
            HAVC_CHECK_CUDA(cudaFuncSetAttribute(resample_h_regw_kernel<40>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
            HAVC_CHECK_CUDA(cudaFuncSetAttribute(resample_h_regw_kernel<48>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
            device_done(attr2, attr2_bit);
        }
        const int block = ((Wout + 31) / 32) * 32;
        long long g2 = rows < (long long)num_sms() * 4 ? rows : (long long)num_sms() * 4;
        const int kt = taps <= 8 ? 8 : (taps <= 24 ? 24 : (taps <= 40 ? 40 : 48));
        const size_t sm = 2 * (size_t)(Win + kt) * sizeof(float);
        cudaStream_t st = (cudaStream_t)stream;
        if (kt == 8) resample_h_regw_kernel<8><<<(int)g2, block, sm, st>>>(in, out, rows, Win, Wout, start, weights, taps);
        else if (kt == 24) resample_h_regw_kernel<24><<<(int)g2, block, sm, st>>>(in, out, rows, Win, Wout, start, weights, taps);
        else if (kt == 40) resample_h_regw_kernel<40><<<(int)g2, block, sm, st>>>(in, out, rows, Win, Wout, start, weights, taps);
        else resample_h_regw_kernel<48><<<(int)g2, block, sm, st>>>(in, out, rows, Win, Wout, start, weights, taps);
        HAVC_LAUNCHED();
        return HAVC_OK;
    }
    static std::atomic<unsigned long long> attr{0ull};
    unsigned long long attr_bit;
    if (device_pending(attr, &attr_bit)) {
        HAVC_CHECK_CUDA(cudaFuncSetAttribute(resample_h_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        device_done(attr, attr_bit);
    }
    long long g = rows < (long long)num_sms() * 16 ? rows : (long long)num_sms() * 16;
    resample_h_kernel<<<(int)g, 128, Win * sizeof(float), (cudaStream_t)stream>>>(in, out, rows, Win, Wout, start, weights,
                                                                               taps);
    HAVC_LAUNCHED();
    return HAVC_OK;
}

extern "C" int havc_pre_vertical(const float *in, uint8_t *rgb_small, void *x, int B, int Hin, int S, const int *start,
                                 const float *weights, int taps, int dtype, void *stream) {
    HAVC_CHECK_ARG(in && rgb_small && x && start && weights && taps > 0 && (dtype == HAVC_F16 || dtype == HAVC_BF16),
                   "havc_pre_vertical: bad arguments");
    static const bool legacy_v = getenv("HAVC_B200_LEGACY_PIXEL") != nullptr;   // A/B switch for profiling
    if (!legacy_v && (S & 3) == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(rgb_small) & 3) == 0)
        pre_vertical4_kernel<<<grid1d((long long)B * S * (S / 4), 256), 256, 0, (cudaStream_t)stream>>>(
            in, rgb_small, x, B, Hin, S, start, weights, taps, dtype);
    else
        pre_vertical_kernel<<<grid1d((long long)B * S * S, 256), 256, 0, (cudaStream_t)stream>>>(
            in, rgb_small, x, B, Hin, S, start, weights, taps, dtype);
    HAVC_LAUNCHED();
    return HAVC_OK;
}

extern "C" int havc_resample_v_f32_u8(const float *in, uint8_t *out, long long planes, int Hin, int Hout, int W, const int *start,
                                      const float *weights, int taps, void *stream) {
    HAVC_CHECK_ARG(in && out && start && weights && taps > 0 && planes > 0 && Hin > 0 && Hout > 0 && W > 0, "havc_resample_v_f32_u8: bad arguments");
    resample_v_f32_u8_kernel<<<grid1d(planes * Hout * W, 256), 256, 0, (cudaStream_t)stream>>>(in, out, planes, Hin, Hout, W, start, weights, taps);
    HAVC_LAUNCHED();
    return HAVC_OK;
}

extern "C" int havc_gray_normalize(const uint8_t *rgb, void *x, int B, long long n_pixels, int dtype, void *stream) {
    HAVC_CHECK_ARG(rgb && x && B > 0 && n_pixels > 0 && (dtype == HAVC_F16 || dtype == HAVC_BF16), "havc_gray_normalize: bad arguments");
    gray_normalize_kernel<<<grid1d((long long)B * n_pixels, 256), 256, 0, (cudaStream_t)stream>>>(rgb, x, B, n_pixels, dtype);
    HAVC_LAUNCHED();
    return HAVC_OK;
}

template <int kTP, int kRt>
static int launch_h_periodic(const uint8_t *in, float *out, long long rows, int Win, int Wout, const int *start, const float *weights,
                             int taps, const havc_periodic_plan *pl, cudaStream_t st) {
    const int nblk = pl->hi - pl->lo, nborder = 8 * pl->lo + (Wout - 8 * pl->hi);
    const int block = ((2 * nblk + 31) / 32) * 32;
    const size_t sm = (size_t)2 * (Win + 64) * sizeof(float) + (size_t)nborder * (taps + 1) * sizeof(float);
    HAVC_CHECK_ARG(block <= 160 && sm <= 96 * 1024, "havc_resample_h_periodic: row too wide for the periodic kernel");
    static std::atomic<unsigned long long> attr{0ull};
    unsigned long long bit;
    if (device_pending(attr, &bit)) {
        HAVC_CHECK_CUDA(cudaFuncSetAttribute(resample_h_periodic_kernel<kTP, kRt>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        device_done(attr, bit);
    }
    PeriodicW pw;
    memcpy(pw.w, pl->w, sizeof(pw.w));
    const long long groups = (rows + 1) / 2;
    static const int per_sm = getenv("HAVC_B200_PX_BLOCKS") ? atoi(getenv("HAVC_B200_PX_BLOCKS")) : 12;  // tuning knob
    const long long cap = (long long)num_sms() * per_sm;
    const int grid = (int)(groups < cap ? groups : cap);
    resample_h_periodic_kernel<kTP, kRt><<<grid, block, sm, st>>>(in, out, rows, Win, Wout, start, weights, taps, pw, pl->offset, pl->lo, pl->hi);
    HAVC_LAUNCHED();
    return HAVC_OK;
}

/* havc_resample_h for an integer ratio Win / Wout whose interior columns share one weight vector (resample.periodic_plan). */
extern "C" int havc_resample_h_periodic(const uint8_t *in, float *out, long long rows, int Win, int Wout, const int *start,
                                        const float *weights, int taps, const havc_periodic_plan *plan, void *stream) {
    HAVC_CHECK_ARG(in && out && start && weights && plan && taps > 0 && rows > 0, "havc_resample_h_periodic: bad arguments");
    HAVC_CHECK_ARG(Win == plan->ratio * Wout && (Win & 3) == 0 && (Wout & 7) == 0 && (reinterpret_cast<uintptr_t>(in) & 3) == 0 &&
                       (reinterpret_cast<uintptr_t>(out) & 15) == 0 && plan->lo >= 0 && plan->lo < plan->hi && 8 * plan->hi <= Wout &&
                       (plan->offset & 3) == 0 && plan->offset + plan->ratio * 8 * plan->lo >= 0 && taps <= 48,
                   "havc_resample_h_periodic: plan does not fit the call");
    cudaStream_t st = (cudaStream_t)stream;
    const int r = plan->ratio, tp = plan->taps;
    if (r == 5 && tp <= 43) return launch_h_periodic<43, 5>(in, out, rows, Win, Wout, start, weights, taps, plan, st);
    if (r == 4 && tp <= 35) return launch_h_periodic<35, 4>(in, out, rows, Win, Wout, start, weights, taps, plan, st);
    if (r == 6 && tp <= 51) return launch_h_periodic<51, 6>(in, out, rows, Win, Wout, start, weights, taps, plan, st);
    HAVC_CHECK_ARG(false, "havc_resample_h_periodic: ratio / taps not instantiated (4: <= 35, 5: <= 43, 6: <= 51)");
    return HAVC_OK;
}

template <int kRt, int kTP>
static int launch_post_periodic(const float *in, const uint8_t *orig, uint8_t *out, int B, int S, int H, int W, const int *start,
                                const float *weights, int taps, int transplant, const havc_periodic_plan *pl, cudaStream_t st) {
    const int upr = S / 4, warps = (upr + 31) / 32;
    const int nbpx = (pl->lo + (upr - pl->hi)) * 4 * kRt;
    const size_t sm = (size_t)3 * (S + 16) * sizeof(float) + (size_t)warps * 3 * 32 * kRt * sizeof(uint32_t) + (size_t)nbpx * (taps + 1) * sizeof(float);
    HAVC_CHECK_ARG(warps <= 5 && sm <= 96 * 1024, "havc_post_horizontal_periodic: row too wide for the periodic kernel");
    static std::atomic<unsigned long long> attr{0ull};
    unsigned long long bit;
    if (device_pending(attr, &bit)) {
        HAVC_CHECK_CUDA(cudaFuncSetAttribute(post_horizontal_periodic_kernel<kRt, kTP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        device_done(attr, bit);
    }
    PeriodicW pw;
    memcpy(pw.w, pl->w, sizeof(pw.w));
    const long long nrows = (long long)B * H;
    static const int per_sm = getenv("HAVC_B200_PX_BLOCKS") ? atoi(getenv("HAVC_B200_PX_BLOCKS")) : 12;  // tuning knob
    const long long cap = (long long)num_sms() * per_sm;
    const int grid = (int)(nrows < cap ? nrows : cap);
    post_horizontal_periodic_kernel<kRt, kTP><<<grid, warps * 32, sm, st>>>(in, orig, out, B, S, H, W, start, weights, taps, transplant, pw,
                                                                           pl->offset, pl->lo, pl->hi);
    HAVC_LAUNCHED();
    return HAVC_OK;
}

/* havc_post_horizontal for an integer ratio W / S whose interior columns are phase-periodic (resample.periodic_plan). */
extern "C" int havc_post_horizontal_periodic(const float *in, const uint8_t *orig, uint8_t *out, int B, int S, int H, int W,
                                             const int *start, const float *weights, int taps, int transplant,
                                             const havc_periodic_plan *plan, void *stream) {
    HAVC_CHECK_ARG(in && out && start && weights && plan && taps > 0 && (!transplant || orig), "havc_post_horizontal_periodic: bad arguments");
    HAVC_CHECK_ARG(W == plan->ratio * S && (S & 3) == 0 && ((reinterpret_cast<uintptr_t>(orig) | reinterpret_cast<uintptr_t>(out)) & 3) == 0 &&
                       (reinterpret_cast<uintptr_t>(in) & 15) == 0 && plan->lo >= 0 && plan->lo < plan->hi && plan->hi <= S / 4 &&
                       plan->offset >= 0 && plan->offset <= 8 && plan->ratio * plan->taps <= 72,
                   "havc_post_horizontal_periodic: plan does not fit the call");
    cudaStream_t st = (cudaStream_t)stream;
    const int r = plan->ratio, tp = plan->taps;
#define HAVC_POST_P(R, TP) \
    if (r == R && tp == TP) return launch_post_periodic<R, TP>(in, orig, out, B, S, H, W, start, weights, taps, transplant, plan, st);
    HAVC_POST_P(5, 8) HAVC_POST_P(5, 9) HAVC_POST_P(5, 10) HAVC_POST_P(4, 8) HAVC_POST_P(4, 9) HAVC_POST_P(4, 10)
    HAVC_POST_P(6, 8) HAVC_POST_P(6, 9) HAVC_POST_P(6, 10)
#undef HAVC_POST_P
    HAVC_CHECK_ARG(false, "havc_post_horizontal_periodic: ratio / taps not instantiated (ratio 4-6, taps 8-10)");
    return HAVC_OK;
}

extern "C" int havc_head(const void *res, int Cs, const float *w11, const float *b11, const uint8_t *rgb_small,
                         uint8_t *colored, float *net_out, const uint8_t *skip, int B, int S, int dtype, int transplant,
                         void *stream) {
    HAVC_CHECK_ARG(res && b11 && colored && Cs % 8 == 0 && Cs <= 512 && ((!transplant && !skip) || rgb_small) &&
                       (dtype == HAVC_F16 || dtype == HAVC_BF16),
                   "havc_head: bad arguments");
    if (Cs == 0) {   // `res` holds fp32 logits [B*S*S][4] produced by the fused-head GEMM epilogue
        head_from_logits_kernel<<<grid1d((long long)B * S * S, 256), 256, 0, (cudaStream_t)stream>>>(
            (const float4 *)res, b11, rgb_small, colored, net_out, skip, B, S, transplant);
        HAVC_LAUNCHED();
        return HAVC_OK;
    }
    HAVC_CHECK_ARG(w11 != nullptr, "havc_head: w11 missing");
    head_kernel<<<grid1d((long long)B * S * S * 32, 256), 256, 0, (cudaStream_t)stream>>>(
        res, Cs, w11, b11, rgb_small, colored, net_out, skip, B, S, dtype, transplant);
    HAVC_LAUNCHED();
    return HAVC_OK;
}

extern "C" int havc_resample_v(const uint8_t *in, float *out, long long planes, int Hin, int Hout, int W,
                               const int *start, const float *weights, int taps, void *stream) {
    HAVC_CHECK_ARG(in && out && start && weights && taps > 0, "havc_resample_v: bad arguments");
    static const bool legacy_v = getenv("HAVC_B200_LEGACY_PIXEL") != nullptr;   // A/B switch for profiling
    if (!legacy_v && (W & 3) == 0 && (reinterpret_cast<uintptr_t>(in) & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0)
        resample_v4_kernel<<<grid1d(planes * Hout * (W / 4), 256), 256, 0, (cudaStream_t)stream>>>(
            in, reinterpret_cast<float4 *>(out), planes, Hin, Hout, W, start, weights, taps);
    else
        resample_v_kernel<<<grid1d(planes * Hout * W, 256), 256, 0, (cudaStream_t)stream>>>(in, out, planes, Hin, Hout, W,
                                                                                         start, weights, taps);
    HAVC_LAUNCHED();
    return HAVC_OK;
}

extern "C" int havc_post_horizontal(const float *in, const uint8_t *orig, uint8_t *out, int B, int S, int H, int W,
                                    const int *start, const float *weights, int taps, int transplant, void *stream) {
    HAVC_CHECK_ARG(in && out && start && weights && taps > 0 && (!transplant || orig) && 3 * S * sizeof(float) <= 48 * 1024,
                   "havc_post_horizontal: bad arguments");
    long long rows = (long long)B * H;
    long long g = rows < (long long)num_sms() * 16 ? rows : (long long)num_sms() * 16;
    static const bool legacy_p = getenv("HAVC_B200_LEGACY_PIXEL") != nullptr;   // A/B switch for profiling
    // 16-byte loads of the start / weight tables need aligned tables and rows (W % 4 == 0)
    const bool vec = !legacy_p && (W & 3) == 0 && ((reinterpret_cast<uintptr_t>(start) | reinterpret_cast<uintptr_t>(weights)) & 15) == 0;
    post_horizontal_kernel<<<(int)g, 256, 3 * S * sizeof(float), (cudaStream_t)stream>>>(in, orig, out, B, S, H, W, start,
                                                                                     weights, taps, transplant, vec ? 1 : 0);
    HAVC_LAUNCHED();
    return HAVC_OK;
}
