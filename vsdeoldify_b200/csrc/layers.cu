// layers.cu — the memory-bound network ops around the tcgen05 GEMM: small-Cin im2col (ResNet stem,
// Zhang first conv), max-pool, stride-2 phase split, eval-BatchNorm(+ReLU), the PixelShuffle_ICNR blur,
// the attention row soft-max.  All NHWC 16-bit, 8 channels (16 B) per thread access, grid-stride loops.
#include "common.cuh"
#include <stdlib.h>

namespace havc {

static inline int grid_for(long long n_threads, int block) {
    long long g = (n_threads + block - 1) / block;
    long long cap = (long long)num_sms() * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

// ---------------------------------------------------------------------------------------------
// im2col for convolutions with a tiny channel count (7x7 s2 stem on the 3-channel normalised image,
// torchvision resnet conv1 via vsdeoldify/fastai/vision/learner.py:54-63; the 3 image channels of
// MergeLayer(dense=True) in the res_block; Zhang model1.0 on the 1-channel L image).
// Input pixels are 8 channels (16 B) wide.  K layout of the output: filter row kh occupies
// [kh*RW, kh*RW + ks*cin) with RW = round_up(ks*cin, 8), i.e. k = kh*RW + kw*cin + c; everything else is zero and
// is never written (the buffer is zero-initialised once).  One thread = one (output pixel, filter row):
// ks 16-byte loads, RW/8 16-byte stores.
// ---------------------------------------------------------------------------------------------
// CIN > 0 fixes the channel count at compile time: the K-row scratch `v` is then indexed statically and lives in registers
// (with a run-time cin it sits in local memory and the kernel is bound by those load/store-unit round trips).
template <int KS, int CIN>
__global__ void im2col_rows_kernel(const uint4 *__restrict__ in, uint16_t *__restrict__ out, int B, int H, int W,
                                   int cin_rt, int stride, int pad, int OH, int OW, int Kp, int c0) {
    constexpr int kMaxRW = ((KS * (CIN > 0 ? CIN : 8) + 7) / 8) * 8;
    const int cin = CIN > 0 ? CIN : cin_rt;
    const int RW = ((KS * cin + 7) / 8) * 8;
    const long long total = (long long)B * OH * OW * KS;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int kh = (int)(i % KS);
        long long pix = i / KS;
        const int ox = (int)(pix % OW);
        const int oy = (int)((pix / OW) % OH);
        const int b = (int)(pix / ((long long)OW * OH));
        uint16_t v[kMaxRW];
#pragma unroll
        for (int j = 0; j < kMaxRW; ++j) v[j] = 0;
        const int iy = oy * stride - pad + kh;
        if (iy >= 0 && iy < H) {
#pragma unroll
            for (int kw = 0; kw < KS; ++kw) {
                const int ix = ox * stride - pad + kw;
                if (ix < 0 || ix >= W) continue;
                const uint4 q = __ldg(in + ((long long)b * H + iy) * W + ix);
                const uint32_t w[4] = {c0 ? q.z : q.x, c0 ? q.w : q.y, q.z, q.w};     // c0 = 4: channels 4.. of the pixel
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    if (c < cin) v[kw * cin + c] = (uint16_t)((w[c >> 1] >> ((c & 1) * 16)) & 0xffff);
            }
        }
        uint16_t *dst = out + pix * Kp + kh * RW;
#pragma unroll
        for (int g = 0; g < kMaxRW / 8; ++g) {
            if (g * 8 < RW) {
                uint4 o;
                o.x = v[g * 8 + 0] | ((uint32_t)v[g * 8 + 1] << 16);
                o.y = v[g * 8 + 2] | ((uint32_t)v[g * 8 + 3] << 16);
                o.z = v[g * 8 + 4] | ((uint32_t)v[g * 8 + 5] << 16);
                o.w = v[g * 8 + 6] | ((uint32_t)v[g * 8 + 7] << 16);
                *reinterpret_cast<uint4 *>(dst + g * 8) = o;
            }
        }
    }
}

// Same im2col, tiled: a block builds the K rows of 128 consecutive output pixels in shared memory (16-byte chunks XOR-swizzled
// so both sides are bank-conflict free) and then streams the tile out as ONE contiguous run of whole 128-byte lines, zero tail
// included.  The per-(pixel, filter row) kernel above leaves a never-written tail in every K row and issues half-sector
// stores; partially written lines cost a read-modify-write at the (ECC) HBM and held it to 1-1.7 TB/s.  Needs Kp % 64 == 0.
template <int KS, int CIN>
__global__ void im2col_tile_kernel(const uint4 *__restrict__ in, uint4 *__restrict__ out, int H, int W, int cin_rt, int stride,
                                   int pad, int OH, int OW, int Kp, long long total_pix, int c0) {
    extern __shared__ uint4 tile[];          // [128][Kp / 8]
    constexpr int kMaxRW = ((KS * (CIN > 0 ? CIN : 8) + 7) / 8) * 8;
    const int cin = CIN > 0 ? CIN : cin_rt;
    const int CH = Kp >> 3;                  // 16-byte chunks per pixel (multiple of 8)
    const int RW = ((KS * cin + 7) / 8) * 8, RC = RW >> 3;
    const int t = threadIdx.x;
    for (long long p0 = blockIdx.x * 128ll; p0 < total_pix; p0 += gridDim.x * 128ll) {
        const long long pix = p0 + t;
        if (pix < total_pix) {
            const int ox = (int)(pix % OW);
            const int oy = (int)((pix / OW) % OH);
            const long long b = pix / ((long long)OW * OH);
            uint4 *row = tile + t * CH;
#pragma unroll
            for (int kh = 0; kh < KS; ++kh) {
                uint16_t v[kMaxRW];
#pragma unroll
                for (int j = 0; j < kMaxRW; ++j) v[j] = 0;
                const int iy = oy * stride - pad + kh;
                if (iy >= 0 && iy < H) {
#pragma unroll
                    for (int kw = 0; kw < KS; ++kw) {
                        const int ix = ox * stride - pad + kw;
                        if (ix < 0 || ix >= W) continue;
                        const uint4 q = __ldg(in + (b * H + iy) * W + ix);
                        const uint32_t w[4] = {c0 ? q.z : q.x, c0 ? q.w : q.y, q.z, q.w};
#pragma unroll
                        for (int c = 0; c < 8; ++c)
                            if (c < cin) v[kw * cin + c] = (uint16_t)((w[c >> 1] >> ((c & 1) * 16)) & 0xffff);
                    }
                }
#pragma unroll
                for (int g = 0; g < kMaxRW / 8; ++g) {
                    if (g < RC) {
                        const int chunk = kh * RC + g;
                        row[(chunk & ~7) | ((chunk ^ t) & 7)] =
                            make_uint4(v[g * 8 + 0] | ((uint32_t)v[g * 8 + 1] << 16), v[g * 8 + 2] | ((uint32_t)v[g * 8 + 3] << 16),
                                       v[g * 8 + 4] | ((uint32_t)v[g * 8 + 5] << 16), v[g * 8 + 6] | ((uint32_t)v[g * 8 + 7] << 16));
                    }
                }
            }
            for (int chunk = KS * RC; chunk < CH; ++chunk) row[(chunk & ~7) | ((chunk ^ t) & 7)] = make_uint4(0u, 0u, 0u, 0u);
        }
        __syncthreads();
        const long long left = total_pix - p0;
        const int n = (int)(left < 128 ? left : 128) * CH;
        uint4 *dst = out + p0 * CH;
        for (int f = t; f < n; f += 128) {
            const int pl = f / CH, slot = f - pl * CH;
            dst[pl * CH + ((slot & ~7) | ((slot ^ pl) & 7))] = tile[f];
        }
        __syncthreads();
    }
}

// The stem's im2col (7 x 7, stride 2, pad 3, two channels = one 32-bit word per input pixel, Kp = 128): one block per output
// row.  The seven input rows are staged once in shared memory as words (the generic kernel issues 49 16-byte loads per output
// pixel to use 4 bytes of each); a thread assembles its pixel's 14 chunks from 7 consecutive words per filter row and the tile
// leaves as whole lines, exactly like im2col_tile_kernel (same layout, same bytes).
__global__ void im2col_stem_kernel(const uint4 *__restrict__ in, uint4 *__restrict__ out, int H, int W, int OH, int OW,
                                   long long rows_total, int c0) {
    extern __shared__ uint4 tile[];                       // [OW][16] swizzled, then the 7 staged input rows
    constexpr int KS = 7, CH = 16;
    const int wp = W + 8;                                 // staged row pitch in words: 4 zero words on either side
    uint32_t *srow = reinterpret_cast<uint32_t *>(tile + OW * CH);
    const int t = threadIdx.x;
    const uint32_t *in32 = reinterpret_cast<const uint32_t *>(in) + (c0 ? 2 : 0);
    for (long long r = blockIdx.x; r < rows_total; r += gridDim.x) {
        const int oy = (int)(r % OH);
        const long long b = r / OH;
        for (int i = t; i < KS * wp; i += blockDim.x) {
            const int kh = i / wp, ix = i - kh * wp - 4;
            const int iy = oy * 2 - 3 + kh;
            srow[i] = (iy >= 0 && iy < H && ix >= 0 && ix < W) ? __ldg(in32 + ((b * H + iy) * W + ix) * 4) : 0u;
        }
        __syncthreads();
        if (t < OW) {
            uint4 *row = tile + t * CH;
#pragma unroll
            for (int kh = 0; kh < KS; ++kh) {
                const uint32_t *sp = srow + kh * wp + 2 * t + 1;          // input column 2 t - 3 sits at word 2 t - 3 + 4
                const int c0k = kh * 2, c1k = kh * 2 + 1;
                row[(c0k & ~7) | ((c0k ^ t) & 7)] = make_uint4(sp[0], sp[1], sp[2], sp[3]);
                row[(c1k & ~7) | ((c1k ^ t) & 7)] = make_uint4(sp[4], sp[5], sp[6], 0u);
            }
#pragma unroll
            for (int chunk = 2 * KS; chunk < CH; ++chunk) row[(chunk & ~7) | ((chunk ^ t) & 7)] = make_uint4(0u, 0u, 0u, 0u);
        }
        __syncthreads();
        uint4 *dst = out + r * OW * CH;
        for (int f = t; f < OW * CH; f += blockDim.x) {
            const int pl = f / CH, slot = f - pl * CH;
            dst[pl * CH + ((slot & ~7) | ((slot ^ pl) & 7))] = tile[f];
        }
        __syncthreads();
    }
}

// 3x3 stride-2 pad-1 max-pool (torchvision resnet maxpool).  Split-precision tensors (in_lo / out_lo != NULL) are compared as
// hi + lo in fp32 (exact: 11 + 11 significant bits) and stored as hi / lo planes again.
__global__ void maxpool_kernel(const uint4 *__restrict__ in, const uint4 *__restrict__ in_lo, uint4 *__restrict__ out,
                               uint4 *__restrict__ out_lo, int B, int H, int W, int C8, int OH, int OW, int dtype) {
    const long long total = (long long)B * OH * OW * C8;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int cg = (int)(i % C8);
        long long pix = i / C8;
        const int ox = (int)(pix % OW);
        const int oy = (int)((pix / OW) % OH);
        const int b = (int)(pix / ((long long)OW * OH));
        float m[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
        for (int dy = -1; dy <= 1; ++dy) {
            const int iy = oy * 2 + dy;
            if (iy < 0 || iy >= H) continue;
            for (int dx = -1; dx <= 1; ++dx) {
                const int ix = ox * 2 + dx;
                if (ix < 0 || ix >= W) continue;
                const long long off = (((long long)b * H + iy) * W + ix) * C8 + cg;
                const uint4 v = __ldg(in + off);
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
                float f[8];
#pragma unroll
                for (int j = 0; j < 4; ++j) { const float2 t = unpack2(w[j], dtype); f[2 * j] = t.x; f[2 * j + 1] = t.y; }
                if (in_lo != nullptr) {
                    const uint4 l = __ldg(in_lo + off);
                    const uint32_t wl[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) { const float2 t = unpack2(wl[j], dtype); f[2 * j] += t.x; f[2 * j + 1] += t.y; }
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], f[j]);
            }
        }
        uint32_t hp[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) hp[j] = pack2(m[2 * j], m[2 * j + 1], dtype);
        out[i] = make_uint4(hp[0], hp[1], hp[2], hp[3]);
        if (out_lo != nullptr) {
            uint32_t lp[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) { const float2 h = unpack2(hp[j], dtype); lp[j] = pack2(m[2 * j] - h.x, m[2 * j + 1] - h.y, dtype); }
            out_lo[i] = make_uint4(lp[0], lp[1], lp[2], lp[3]);
        }
    }
}

// Phase split for stride-2 convolutions: out[p=(a*2+b)][n][i][j][c] = in[n][2i+a][2j+b][c] (zero past the edge).
__global__ void phase_split_kernel(const uint4 *__restrict__ in, uint4 *__restrict__ out, int B, int H, int W,
                                   int C8, int OH, int OW, int P) {
    const long long total = (long long)P * B * OH * OW * C8;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int cg = (int)(i % C8);
        long long r = i / C8;
        const int ox = (int)(r % OW); r /= OW;
        const int oy = (int)(r % OH); r /= OH;
        const int b = (int)(r % B);
        const int p = (int)(r / B);
        const int iy = oy * 2 + (p >> 1), ix = ox * 2 + (p & 1);
        uint4 v = make_uint4(0, 0, 0, 0);
        if (iy < H && ix < W) v = __ldg(in + (((long long)b * H + iy) * W + ix) * C8 + cg);
        out[i] = v;
    }
}

// y = x*scale[c] + shift[c], optional ReLU (eval BatchNorm on skips / encoder output:
// unet.py:203 `self.bn(s)`, unet.py:244 layers[1..2]).
__global__ void affine_act_kernel(const uint4 *__restrict__ in, const uint4 *__restrict__ in_lo, uint4 *__restrict__ out,
                                  uint4 *__restrict__ out_lo, long long n_pix, int C8, int in_stride8, int out_stride8,
                                  const float *__restrict__ scale, const float *__restrict__ shift, int relu, int dtype) {
    const long long total = n_pix * C8;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int cg = (int)(i % C8);
        const long long pix = i / C8;
        uint4 v = __ldg(in + pix * in_stride8 + cg);
        uint32_t w[4] = {v.x, v.y, v.z, v.w};
        uint32_t wl[4] = {0u, 0u, 0u, 0u};
        if (in_lo != nullptr) {     // split-precision input: value = hi + lo
            const uint4 l = __ldg(in_lo + pix * in_stride8 + cg);
            wl[0] = l.x; wl[1] = l.y; wl[2] = l.z; wl[3] = l.w;
        }
        // the 8 scales / shifts of this channel group as two 16-byte loads each (16 scalar loads per 16 bytes of data made
        // the kernel load/store-unit bound for wide tensors: 1.7 TB/s at C = 256 against 4.4 TB/s at C = 64)
        const float4 s0 = __ldg(reinterpret_cast<const float4 *>(scale) + 2 * cg), s1 = __ldg(reinterpret_cast<const float4 *>(scale) + 2 * cg + 1);
        const float4 t0 = __ldg(reinterpret_cast<const float4 *>(shift) + 2 * cg), t1 = __ldg(reinterpret_cast<const float4 *>(shift) + 2 * cg + 1);
        const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w}, sh[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float2 f = unpack2(w[j], dtype);
            const float2 fl = unpack2(wl[j], dtype);
            f.x = fmaf(f.x + fl.x, sc[2 * j], sh[2 * j]);
            f.y = fmaf(f.y + fl.y, sc[2 * j + 1], sh[2 * j + 1]);
            if (relu) { f.x = fmaxf(f.x, 0.f); f.y = fmaxf(f.y, 0.f); }
            w[j] = pack2(f.x, f.y, dtype);
            if (out_lo != nullptr) {
                const float2 h = unpack2(w[j], dtype);
                wl[j] = pack2(f.x - h.x, f.y - h.y, dtype);
            }
        }
        out[pix * out_stride8 + cg] = make_uint4(w[0], w[1], w[2], w[3]);
        if (out_lo != nullptr) out_lo[pix * out_stride8 + cg] = make_uint4(wl[0], wl[1], wl[2], wl[3]);
    }
}

// ReplicationPad2d((1,0,1,0)) + AvgPool2d(2, stride=1): out[y][x] = mean(in[y-1..y][x-1..x]) with the
// index clamped at 0 (CustomPixelShuffle_ICNR, unet.py:47-52; PixelShuffle_ICNR, fastai/layers.py:214-220).
// One thread walks a strip of kBlurRows output rows for a fixed (x, 8-channel group) and keeps the previous input row's two
// taps in registers, so every output costs two 16-byte loads instead of four.
static constexpr int kBlurRows = 8;
static constexpr int kBlurCols = 4;   // adjacent output columns per thread: kBlurCols + 1 loads per row for kBlurCols outputs
// ((a + b) + (c + d)) * 0.25 in packed 16-bit arithmetic (three roundings in the storage type; the inputs are non-negative
// ReLU outputs, so there is no cancellation): the same operations, in the same order, as the fused epilogue of conv_gemm.cu
__device__ __forceinline__ uint4 blur_avg4(const uint4 &a, const uint4 &b, const uint4 &c, const uint4 &d, int dtype) {
    const uint32_t pa[4] = {a.x, a.y, a.z, a.w}, pb[4] = {b.x, b.y, b.z, b.w}, pc[4] = {c.x, c.y, c.z, c.w}, pd[4] = {d.x, d.y, d.z, d.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = quarter2(add2(add2(pa[j], pb[j], dtype), add2(pc[j], pd[j], dtype), dtype), dtype);
    return make_uint4(o[0], o[1], o[2], o[3]);
}
__global__ void blur2x2_kernel(const uint4 *__restrict__ in, uint4 *__restrict__ out, int B, int H, int W, int C8,
                               int out_stride8, int dtype) {
    const int strips = (H + kBlurRows - 1) / kBlurRows;
    const int xg = (W + kBlurCols - 1) / kBlurCols;
    const long long total = (long long)B * strips * xg * C8;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int cg = (int)(i % C8);
        long long t = i / C8;
        const int x = (int)(t % xg) * kBlurCols;
        t /= xg;
        const int ys = (int)(t % strips) * kBlurRows;
        const int b = (int)(t / strips);
        // input columns x-1 (clamped at 0), x, x+1, ... (clamped at W-1: those outputs are not stored)
        int xc[kBlurCols + 1];
        xc[0] = x > 0 ? x - 1 : 0;
#pragma unroll
        for (int k = 0; k < kBlurCols; ++k) xc[k + 1] = x + k < W ? x + k : W - 1;
        const uint4 *base = in + (long long)b * H * W * C8 + cg;
        const int yp = ys > 0 ? ys - 1 : 0;
        uint4 prev[kBlurCols + 1], cur[kBlurCols + 1];
#pragma unroll
        for (int k = 0; k <= kBlurCols; ++k) prev[k] = __ldg(base + ((long long)yp * W + xc[k]) * C8);
        const int yend = min(ys + kBlurRows, H);
        for (int y = ys; y < yend; ++y) {
#pragma unroll
            for (int k = 0; k <= kBlurCols; ++k) cur[k] = __ldg(base + ((long long)y * W + xc[k]) * C8);
            uint4 *dst = out + (((long long)b * H + y) * W + x) * out_stride8 + cg;
#pragma unroll
            for (int k = 0; k < kBlurCols; ++k)
                if (x + k < W) dst[(long long)k * out_stride8] = blur_avg4(prev[k], prev[k + 1], cur[k], cur[k + 1], dtype);
#pragma unroll
            for (int k = 0; k <= kBlurCols; ++k) prev[k] = cur[k];
        }
    }
}

// Row soft-max of the attention logits: P[j, :] = softmax_i S[j, i] (SelfAttention, fastai/layers.py:94:
// softmax over dim=1 of beta[b,i,j] == over the contiguous row of the transposed logits we store).
// One warp per row; the row lives in registers across the three passes when cols <= 32*kMaxPerLane.
// kIn16: the logits were stored as fp16 (attention logits of the fp16 path: |S| stays far below the fp16 range and the 2^-11
// rounding disappears in the soft-max, tools/precision_emulator.py) instead of fp32 - half the bytes of the N x N round trip.
template <bool kIn16>
__global__ void softmax_rows_kernel(const void *__restrict__ in, void *__restrict__ out, long long rows, int cols,
                                    int in_stride, int out_stride, int out_dtype) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    auto load4 = [&](long long row, int c) -> float4 {
        if constexpr (kIn16) {
            const uint2 u = *reinterpret_cast<const uint2 *>(reinterpret_cast<const uint16_t *>(in) + row * in_stride + c);
            const float2 a = unpack2(u.x, HAVC_F16), b = unpack2(u.y, HAVC_F16);
            return make_float4(a.x, a.y, b.x, b.y);
        } else {
            return *reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(in) + row * in_stride + c);
        }
    };
    for (long long row = warp0; row < rows; row += nwarps) {
        float m = -INFINITY;
        for (int c = lane * 4; c < cols; c += 128) {
            const float4 v = load4(row, c);
            m = fmaxf(fmaxf(m, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        float s = 0.f;
        for (int c = lane * 4; c < cols; c += 128) {
            const float4 v = load4(row, c);
            s += __expf(v.x - m) + __expf(v.y - m) + __expf(v.z - m) + __expf(v.w - m);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float inv = 1.f / s;
        for (int c = lane * 4; c < cols; c += 128) {
            const float4 v = load4(row, c);
            const float e0 = __expf(v.x - m) * inv, e1 = __expf(v.y - m) * inv, e2 = __expf(v.z - m) * inv,
                        e3 = __expf(v.w - m) * inv;
            uint2 pk = make_uint2(pack2(e0, e1, out_dtype), pack2(e2, e3, out_dtype));
            *reinterpret_cast<uint2 *>(reinterpret_cast<uint16_t *>(out) + row * out_stride + c) = pk;
        }
    }
}

}  // namespace havc

using namespace havc;

static bool dt16(int d) { return d == HAVC_F16 || d == HAVC_BF16; }

extern "C" int havc_im2col_small(const void *in, void *out, int B, int H, int W, int Cs, int c0, int cin, int ks, int stride,
                                 int pad, int Kp, int dtype, void *stream) {
    const int RW = ((ks * cin + 7) / 8) * 8;
    HAVC_CHECK_ARG(in && out && dt16(dtype) && Kp % 8 == 0 && Kp >= ks * RW && Cs == 8 && cin >= 1 && cin <= 8 &&
                       (c0 == 0 || (c0 == 4 && cin <= 4)) && (ks == 1 || ks == 3 || ks == 7),
                   "havc_im2col_small: needs 8-channel input pixels, c0 in {0,4}, c0 + cin <= 8, ks in {1,3,7}, Kp >= ks*round_up(ks*cin,8)");
    const int OH = (H + 2 * pad - ks) / stride + 1, OW = (W + 2 * pad - ks) / stride + 1;
    const long long n = (long long)B * OH * OW * ks;
    const int grid = grid_for(n, 256);
    cudaStream_t st = (cudaStream_t)stream;
    static const bool legacy = getenv("HAVC_B200_LEGACY_IM2COL") != nullptr;   // A/B switch for profiling
    const bool tiled = !legacy && Kp % 64 == 0 && 128 * Kp * 2 <= 48 * 1024;
    const long long pixels = (long long)B * OH * OW;
    const long long blocks = (pixels + 127) / 128;
    const int g2 = (int)(blocks < (long long)num_sms() * 8 ? blocks : (long long)num_sms() * 8);
    const size_t sm = (size_t)128 * Kp * 2;
#define HAVC_IM2COL(KS_, CIN_)                                                                                                   \
    do {                                                                                                                         \
        if (tiled)                                                                                                               \
            im2col_tile_kernel<KS_, CIN_><<<g2, 128, sm, st>>>((const uint4 *)in, (uint4 *)out, H, W, cin, stride, pad, OH, OW, Kp, pixels, c0); \
        else                                                                                                                     \
            im2col_rows_kernel<KS_, CIN_><<<grid, 256, 0, st>>>((const uint4 *)in, (uint16_t *)out, B, H, W, cin, stride, pad, OH, OW, Kp, c0); \
    } while (0)
    if (tiled && ks == 7 && cin == 2 && stride == 2 && pad == 3 && Kp == 128 && OW <= 256 && (size_t)OW * 256 + 7 * (W + 8) * 4 <= 96 * 1024) {
        // the stem of the exact-input path: one block per output row
        const size_t sm2 = (size_t)OW * 256 + (size_t)7 * (W + 8) * 4;
        static std::atomic<unsigned long long> attr{0ull};
        unsigned long long bit;
        if (device_pending(attr, &bit)) {
            HAVC_CHECK_CUDA(cudaFuncSetAttribute(im2col_stem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
            device_done(attr, bit);
        }
        const long long rows_total = (long long)B * OH;
        const int g3 = (int)(rows_total < (long long)num_sms() * 4 ? rows_total : (long long)num_sms() * 4);
        im2col_stem_kernel<<<g3, 256, sm2, st>>>((const uint4 *)in, (uint4 *)out, H, W, OH, OW, rows_total, c0);
        HAVC_LAUNCHED();
        return HAVC_OK;
    }
    if (ks == 7 && cin == 3) HAVC_IM2COL(7, 3);
    else if (ks == 7 && cin == 2) HAVC_IM2COL(7, 2);
    else if (ks == 3 && cin == 3) HAVC_IM2COL(3, 3);
    else if (ks == 3 && cin == 1) HAVC_IM2COL(3, 1);
    else if (ks == 7) HAVC_IM2COL(7, 0);
    else if (ks == 3) HAVC_IM2COL(3, 0);
    else HAVC_IM2COL(1, 0);
#undef HAVC_IM2COL
    HAVC_LAUNCHED();
    return HAVC_OK;
}

extern "C" int havc_maxpool3x3s2(const void *in, const void *in_lo, void *out, void *out_lo, int B, int H, int W, int C, int dtype,
                                 void *stream) {
    HAVC_CHECK_ARG(in && out && dt16(dtype) && C % 8 == 0, "havc_maxpool3x3s2: bad arguments");
    const int OH = (H + 2 - 3) / 2 + 1, OW = (W + 2 - 3) / 2 + 1;
    const long long n = (long long)B * OH * OW * (C / 8);
    maxpool_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>((const uint4 *)in, (const uint4 *)in_lo, (uint4 *)out,
                                                                      (uint4 *)out_lo, B, H, W, C / 8, OH, OW, dtype);
    HAVC_LAUNCHED();
    return HAVC_OK;
}

extern "C" int havc_phase_split(const void *in, void *out, int B, int H, int W, int C, int n_phases, void *stream) {
    HAVC_CHECK_ARG(in && out && C % 8 == 0 && (n_phases == 1 || n_phases == 4), "havc_phase_split: bad arguments");
    const int OH = (H + 1) / 2, OW = (W + 1) / 2;
    const long long n = (long long)n_phases * B * OH * OW * (C / 8);
    phase_split_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>((const uint4 *)in, (uint4 *)out, B, H, W,
                                                                          C / 8, OH, OW, n_phases);
    HAVC_LAUNCHED();
    return HAVC_OK;
}

extern "C" int havc_affine_act(const void *in, const void *in_lo, void *out, void *out_lo, long long n_pixels, int C,
                               int in_pix_stride, int out_pix_stride, const float *scale, const float *shift, int relu, int dtype,
                               void *stream) {
    HAVC_CHECK_ARG(in && out && scale && shift && dt16(dtype) && C % 8 == 0 && in_pix_stride % 8 == 0 &&
                       out_pix_stride % 8 == 0 && in_pix_stride >= C && out_pix_stride >= C &&
                       ((reinterpret_cast<uintptr_t>(scale) | reinterpret_cast<uintptr_t>(shift)) & 15) == 0,
                   "havc_affine_act: bad arguments (scale / shift hold C floats, 16-byte aligned)");
    const long long n = n_pixels * (C / 8);
    affine_act_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(
        (const uint4 *)in, (const uint4 *)in_lo, (uint4 *)out, (uint4 *)out_lo, n_pixels, C / 8, in_pix_stride / 8,
        out_pix_stride / 8, scale, shift, relu, dtype);
    HAVC_LAUNCHED();
    return HAVC_OK;
}

extern "C" int havc_blur2x2(const void *in, void *out, int B, int H, int W, int C, int out_pix_stride, int dtype,
                            void *stream) {
    HAVC_CHECK_ARG(in && out && in != out && dt16(dtype) && C % 8 == 0 && out_pix_stride % 8 == 0 && out_pix_stride >= C,
                   "havc_blur2x2: bad arguments");
    const long long n = (long long)B * ((H + kBlurRows - 1) / kBlurRows) * ((W + kBlurCols - 1) / kBlurCols) * (C / 8);
    blur2x2_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>((const uint4 *)in, (uint4 *)out, B, H, W, C / 8,
                                                                      out_pix_stride / 8, dtype);
    HAVC_LAUNCHED();
    return HAVC_OK;
}

extern "C" int havc_softmax_rows(const void *in, int in_dtype, void *out, long long rows, int cols, int in_stride, int out_stride,
                                 int out_dtype, void *stream) {
    HAVC_CHECK_ARG(in && out && dt16(out_dtype) && (in_dtype == HAVC_F32 || in_dtype == HAVC_F16) && cols % 4 == 0 &&
                       in_stride % 4 == 0 && out_stride % 4 == 0 && in_stride >= cols && out_stride >= cols,
                   "havc_softmax_rows: cols and strides must be multiples of 4, input fp32 or fp16");
    const long long nthreads = rows * 32;
    if (in_dtype == HAVC_F16)
        softmax_rows_kernel<true><<<grid_for(nthreads, 256), 256, 0, (cudaStream_t)stream>>>(in, out, rows, cols, in_stride, out_stride, out_dtype);
    else
        softmax_rows_kernel<false><<<grid_for(nthreads, 256), 256, 0, (cudaStream_t)stream>>>(in, out, rows, cols, in_stride, out_stride, out_dtype);
    HAVC_LAUNCHED();
    return HAVC_OK;
}
