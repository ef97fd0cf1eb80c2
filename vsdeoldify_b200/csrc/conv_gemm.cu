// conv_gemm.cu — the implicit-GEMM convolution / batched GEMM kernel of libhavc_b200 (sm_100a).
//
// One persistent, warp-specialised kernel covers every dense contraction on the HAVC hot path
// (reference: torch Conv2d/Conv1d/bmm calls in vsdeoldify/deoldify/unet.py:24-285,
// vsdeoldify/fastai/layers.py:81-96, torchvision resnet via fastai/vision/learner.py:54-63,
// vsdeoldify/colorization/colorizers/{eccv16,siggraph17}.py):
//
//   * activations are NHWC 16-bit; a 128-pixel M tile is a (box_w x box_h x box_b) spatial box that
//     TMA (cp.async.bulk.tensor.5d) fetches once per filter tap and 64-channel K chunk, shifted by the
//     tap offset — out-of-image taps are zero-filled by the TMA unit, so padding costs nothing and no
//     im2col buffer exists;
//   * the K loop can walk two source tensors (torch.cat is never materialised);
//   * weights [Cout][tap][Cin] arrive through a second tensor map; both operands land in shared memory
//     in the 128-byte-swizzled K-major layout tcgen05.mma consumes directly;
//   * tcgen05.mma (M=128, or M=256 across a CTA pair with cta_group::2; N<=256, K=16, kind::f16, fp32 accumulate) into a
//     double-buffered TMEM accumulator; tcgen05.commit releases smem stages / publishes accumulators.  The producer and
//     the issuer are convergent whole-warp code with warp-uniform operands; only the TMA / MMA / commit instructions are
//     predicated on one elected lane (issued from a divergent `if (lane == 0)` region every instruction paid an R2UR
//     waterfall and the issuing thread, not the tensor pipe, was the limiter);
//   * the epilogue warps drain TMEM with tcgen05.ld and apply, in fp32,
//         +bias -> (Leaky)ReLU -> *scale+shift (eval BatchNorm) -> +residual -> ReLU
//     then store fp16/bf16/fp32, optionally in PixelShuffle(2) order or to a strided sub-pixel phase.
//
// Warp roles (1 CTA/SM): warp0 = TMA producer, warp1 = MMA issuer, warp2 = TMEM allocator, warps 4.. = epilogue (TMEM lane
// quarter = warp % 4).  Two epilogues: the generic one (8 warps, 384 threads: fp32 / split / strided outputs, fused head,
// staggered accumulators) and the fast one (16 warps, 640 threads: warp-private parameter slices, 2 KB staging blocks, one
// TMA store per 32x32 chunk, TMA-loaded residual).  Launches are programmatic-dependent (griddepcontrol) so a kernel's
// prologue overlaps its predecessor's tail.
#include "common.cuh"
#include <stdlib.h>
#include <type_traits>

namespace havc {

static constexpr int kTileM = 128;
static constexpr int kChunkK = 64;  // 64 x 16-bit = 128 B = one swizzle row
static constexpr int kABytes = kTileM * kChunkK * 2;
static constexpr int kMaxStages = 8;
static constexpr int kThreads = 384;       // generic epilogue: 4 control warps + 8 epilogue warps
static constexpr int kThreadsFast = 640;   // fast epilogue: 4 control warps + 16 epilogue warps (four per TMEM lane quarter)
static constexpr int kMaxBN = 320;      // parameter staging rows (BN <= 320)
// epilogue shared memory: 2 x {bias,scale,shift}[kMaxBN] + head weights [3][kMaxBN] + head exchange [128][4]
static constexpr int kWarpCols = 64;      // fast epilogue: columns one epilogue warp owns (BN <= 256, every fourth 32-column chunk)
// the generic epilogue's arrays and the fast epilogue's per-warp slices (8 x 3 x kWarpCols floats) share the same bytes
static constexpr int kEpiSmemFloats = 2 * 3 * kMaxBN + 3 * kMaxBN + 128 * 4;
static_assert(16 * 3 * kWarpCols <= kEpiSmemFloats, "per-warp parameter slices must fit the epilogue area");
static constexpr int kBarBytes = 512;    // mbarrier area in front of the epilogue parameters
static constexpr uint32_t kTmemCols = 512;

struct ConvParams {
    int out_B, out_H, out_W;
    int bw, bh, bb;
    int tiles_w, tiles_h, tiles_b, tiles_n, total_tiles;
    int BN, n_wloads, w_box_rows;
    // weight loads of one K step: load l fetches the rows [n0 + wl_row0[l] + rank*wl_rank_rows[l], ...) of the N tile through
    // tensor map wl_map[l] (its box height) to the byte offset wl_smem[l] of the stage's B area; rank = CTA rank in a pair
    int wl_row0[2], wl_rank_rows[2], wl_map[2];
    uint32_t wl_smem[2];
    uint32_t b1_smem_off;         // byte offset of the second MMA's B rows (n_part1 > 0) in the stage's B area
    int n_taps;
    int8_t dh[HAVC_MAX_TAPS], dw[HAVC_MAX_TAPS], tp[HAVC_MAX_TAPS], twi[HAVC_MAX_TAPS];
    int chunks0, chunks1, w_c1_off;
    int a_batched, b_batched;
    int acc_stages, acc_stride;   // acc_stride = TMEM column offset of accumulator stage 1
    int staggered;                // BN > 256: the two accumulators overlap in columns [acc_stride, BN) (see MMA issuer)
    int num_stages;
    uint32_t stage_bytes;
    int n_part0, n_part1;
    uint32_t idesc0, idesc1;
    int N_total;
    const float *bias, *scale, *shift;
    int relu1, relu2;
    float slope1;                 // first activation as max(t, t*slope1): 0 = ReLU, 1 = none, 0.2 = LeakyReLU(0.2)
    const void *residual;
    long long rsw, rsh, rsb;
    void *out;
    int out_dtype;
    long long osw, osh, osb;
    int up, oy, ox, shuffle, group_n, c_store;
    int dtype;
    // K loop variant: src1 contributes once (tap offset 0) after all taps of src0 (an im2col'd side tensor)
    int src1_single_tap, src1_wi;
    // column split: GEMM columns >= split_n are stored to out2 / take their residual from residual2
    int split_n;
    void *out2;
    long long o2sw, o2sh, o2sb;
    int c_store2;
    const void *residual2;
    long long r2sw, r2sh, r2sb;
    // TMA-store epilogue: the tile is staged in shared memory (128B-swizzled 64-channel sub-tiles) and written with
    // cp.async.bulk.tensor stores (coalesced, clipped at the tensor bounds); stage_out_bytes = 128*BN*2
    int tma_store;
    uint32_t stage_out_bytes;
    // fast epilogue only: the residual tile is TMA-loaded (tensor map tmO.m[1], same box and swizzle as the output) into the
    // staging buffer ahead of the accumulator and read back from exactly the shared-memory words the result then overwrites
    int res_tma;
    uint32_t wait_hint_ns;        // suspend-time hint of the long mbarrier waits (0 = plain polling)
    // fused 1x1 head: logits[pix][o] = sum_n y[n] * head_w[o][n]  (o < 3), written as fp32 [pix][4]; no tile store
    const float *head_w;
    float *head_out;
    long long hsw, hsh, hsb;
    // split-precision operands ("x3": value = hi + lo, both 16-bit): K sub-steps per (tap, chunk) of source 0 / 1 =
    // 1 (hi*hi) + [A has a lo plane] (lo*hi) + [W has a lo plane] (hi*lo); x3_out: the fast epilogue stores hi and lo planes
    int a0_lo, a1_lo, w_lo, nsub0, nsub1;
    int x3_out, res_lo;
    uint32_t lo_stage_off;        // byte offset of the lo plane's staging blocks inside the output staging area
    // fused PixelShuffle(2) + ICNR blur epilogue (blur = 1): M tiles overlap by one halo row / column (halo = 1), the N tile holds
    // the four sub-pixel groups of BN/4 channels, the tile is staged in shared memory and every output pixel averages its
    // 2 x 2 neighbourhood of the shuffled image there (ReplicationPad2d((1,0,1,0)) + AvgPool2d(2, 1), unet.py:47-52)
    int blur, halo;
    uint32_t blur_row_bytes;
};

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must trap (launch failure) instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (!mbar_try_wait(bar, parity)) {
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 4000000000ull) {  // 4 s
            printf("havc conv_gemm: mbarrier timeout (block %d thread %d bar %u parity %u)\n",
                   blockIdx.x, threadIdx.x, bar, parity);
            __trap();
        }
    }
}
// Waits that are expected to be long (the epilogue waiting for an accumulator, the producer for a free stage): try_wait with a
// suspend-time hint parks the warp in hardware instead of re-issuing the poll every ~100 cycles (hint = 0: plain polling).
__device__ __forceinline__ bool mbar_try_wait_hint(uint32_t bar, uint32_t parity, uint32_t hint_ns) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(hint_ns)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_long(uint32_t bar, uint32_t parity, uint32_t hint_ns) {
    if (hint_ns == 0) { mbar_wait(bar, parity); return; }
    if (mbar_try_wait(bar, parity)) return;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (!mbar_try_wait_hint(bar, parity, hint_ns)) {
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 4000000000ull) {  // 4 s
            printf("havc conv_gemm: mbarrier timeout (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
            __trap();
        }
    }
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap *tm, uint32_t bar, int c0,
                                            int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        :
        : "r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap *tm, uint32_t bar, int c0,
                                            int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2];"
        :
        : "r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *tm) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16/fp16 operands, fp32 accumulate, single-CTA group.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                         uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        :
        : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
                 : "memory");
}
// One deterministic leader lane of the (fully active) warp.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
// ---- CTA-pair (cta_group::2) helpers ---------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// shared::cluster address of `addr` (a shared::cta address of this CTA's layout) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    __syncwarp();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA loads of a CTA pair: the data lands in the executing CTA's shared memory, the transaction bytes are counted on
// `bar`, a shared::cluster address that may belong to the peer (the leader's full barrier).
__device__ __forceinline__ void tma_load_5d_pair(uint32_t dst, const CUtensorMap *tm, uint32_t bar, int c0,
                                                 int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        :
        : "r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d_pair(uint32_t dst, const CUtensorMap *tm, uint32_t bar, int c0,
                                                 int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2];"
        :
        : "r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
// D[tmem of both CTAs] (+)= A[256 rows: 128 per CTA] * B[N rows: N/2 per CTA]^T, issued by the leader CTA only.
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                              uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        :
        : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives on the barrier at the same shared-memory offset in BOTH CTAs of the pair when the MMAs issued so far retire
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
    asm volatile(
        "{\n\t.reg .b16 m;\n\t"
        "mov.b16 m, 3;\n\t"
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}"
        ::"r"(bar)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),
          "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
// 32 accumulator columns (or the last 16 of a tile whose width is 16 mod 32)
__device__ __forceinline__ void tmem_ld_chunk(uint32_t taddr, uint32_t (&v)[32], int cols_left) {
    if (cols_left >= 32) {
        tmem_ld32(taddr, v);
    } else {
        uint32_t(&lo)[16] = *reinterpret_cast<uint32_t(*)[16]>(&v[0]);
        tmem_ld16(taddr, lo);
    }
}
__device__ __forceinline__ void tmem_ld_wait() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128-byte-swizzled shared-memory matrix descriptor (8-row groups 1024 B apart).
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
    const uint32_t lo = ((smem_addr >> 4) & 0x3FFFu) | (1u << 16);            // LBO = 1 (unused)
    const uint32_t hi = (1024u >> 4) | (1u << 14) /* version 1 */ | (2u << 29) /* SWIZZLE_128B */;
    return (static_cast<uint64_t>(hi) << 32) | lo;
}

// ---------------------------------------------------------------------------------------------
// The kernel
// ---------------------------------------------------------------------------------------------
struct StoreMaps {
    CUtensorMap m[4];   // [0] plain NHWC output; [0..3] the four PixelShuffle sub-pixel views; non-shuffle: [1] residual,
                        // [2] lo plane of the output, [3] lo plane of the residual (split-precision launches)
};
struct LoMaps {
    CUtensorMap a0, a1, w, w1;   // lo planes of src0 / src1 / weight (same geometry as the hi maps)
};

__device__ __forceinline__ void tma_store_5d(const CUtensorMap *tm, uint32_t src, int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
                 :
                 : "l"(tm), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
                 : "memory");
}

template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_read_n(int n) {   // n = bulk groups that may still be reading shared memory
    switch (n) {
        case 0: bulk_wait_read<0>(); break;
        case 1: bulk_wait_read<1>(); break;
        case 2: bulk_wait_read<2>(); break;
        default: bulk_wait_read<3>(); break;
    }
}

// kDT: operand / 16-bit output type (HAVC_F16 or HAVC_BF16), fixed at compile time so the pack / unpack helpers fold.
// kFast: the streamlined epilogue for the common case (16-bit output through the TMA-store staging buffer, whole N tiles,
// no head / column split / staggered accumulators); the generic epilogue covers everything else.
// kPair: two CTAs of a cluster (one TPC) work on two adjacent M tiles of the same N tile with cta_group::2 MMAs
// (M = 256): each CTA loads its own 128 pixels of A and only HALF of the weight rows, the leader's MMA reads both
// shared memories and writes both TMEMs.  Per CTA and K step 16 KB + BN*64 B arrive from L2 instead of 16 KB + BN*128 B
// (the 3x3 decoder convs are bound by exactly that L2 -> SM traffic).
template <int kDT, bool kFast, bool kPair>
__global__ void __launch_bounds__(kFast ? kThreadsFast : kThreads, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                 const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmW1,
                 const __grid_constant__ LoMaps tmL,
                 const __grid_constant__ StoreMaps tmO, const __grid_constant__ ConvParams p) {
    extern __shared__ uint8_t smem_raw[];
    // let the next kernel of the stream (if launched with programmatic serialization) start its own prologue as SMs free up
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t stage_out = smem_base + p.num_stages * p.stage_bytes;        // 1024-aligned output staging
    const uint32_t bar_base = stage_out + p.stage_out_bytes;
    // barrier layout: full[kMaxStages], empty[kMaxStages], tmem_full[2], tmem_empty[2], tmem_ptr
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (kMaxStages + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + 2 + s); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * kMaxStages + 4);
    auto tearly_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + 5 + s); };
    auto wres_bar = [&](int w) { return bar_base + 8u * (2 * kMaxStages + 7 + w); };   // one per epilogue warp (fast epilogue)
    constexpr uint32_t kEpiThreads = kFast ? (kThreadsFast - 128) : (kThreads - 128);
    volatile uint32_t *tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t *>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    // Warp-uniform role / rank values are broadcast from lane 0 so that the compiler KNOWS they are uniform: the producer and the
    // MMA issuer then run as convergent whole-warp code whose operands live in uniform registers, and only the TMA / MMA /
    // commit instructions themselves are predicated on one elected lane.  (Issued from inside an `if (lane == 0)` region every
    // tcgen05.mma cost a 20-instruction R2UR "waterfall"; ~190 instructions per K step made the issuing thread, not the
    // tensor pipe, the limiter of the 3x3 convolutions: profiles/r01_ncu_resconv0_pair_issue_bound.txt.)
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    // work distribution: a "group" is one CTA (or one CTA pair); tiles are dealt round-robin to groups
    const uint32_t rank = kPair ? __shfl_sync(0xffffffffu, cluster_ctarank(), 0) : 0u;
    const int group = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int ngroups = kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    auto decode_tile = [&](int tile, int &nt, int &wt, int &ht, int &bt) {
        int t = tile;
        nt = t % p.tiles_n; t /= p.tiles_n;
        if (kPair) t = 2 * t + (int)rank;    // the pair's two M tiles are neighbours; an odd tail lands outside the batch
        wt = t % p.tiles_w; t /= p.tiles_w;
        ht = t % p.tiles_h; t /= p.tiles_h;
        bt = t;
    };

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA0);
        tma_prefetch_desc(&tmA1);
        tma_prefetch_desc(&tmW);
        if (kPair) tma_prefetch_desc(&tmW1);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < p.num_stages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(tfull_bar(s), 1);
            mbar_init(tempty_bar(s), kPair ? 2 * kEpiThreads : kEpiThreads);   // pair: the leader's barrier collects both CTAs' epilogues
            mbar_init(tearly_bar(s), kPair ? 2 * kEpiThreads : kEpiThreads);
        }
        for (int w = 0; w < 16; ++w) mbar_init(wres_bar(w), 1);
        fence_barrier_init();
    }
    if (warp == 2) {
        if (kPair) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(kTmemCols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(kTmemCols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    if (kPair) cluster_sync_all(); else __syncthreads();   // pair: the peer's barriers must be initialised before any remote signal
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot_ptr, 0);
    // Programmatic dependent launch: everything above (barrier init, TMEM allocation, tensor-map prefetch) touched no global
    // memory and may overlap the tail of the previous kernel in the stream; from here on its results are needed.
    asm volatile("griddepcontrol.wait;" ::: "memory");

    const int ksteps_per_tap = p.src1_single_tap ? p.chunks0 : (p.chunks0 + p.chunks1);
    const int ksteps_main = p.n_taps * ksteps_per_tap;
    const int ksteps = ksteps_main + (p.src1_single_tap ? p.chunks1 : 0);
    // MMA-side K steps: every (tap, chunk) of source s is issued nsub_s times (split-precision products)
    const int ksteps_mma = p.n_taps * (p.chunks0 * p.nsub0 + (p.src1_single_tap ? 0 : p.chunks1 * p.nsub1)) +
                           (p.src1_single_tap ? p.chunks1 * p.nsub1 : 0);

    if (warp == 0) {
        // ===================== TMA producer (whole warp, one elected lane issues) =====================
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = group; tile < p.total_tiles; tile += ngroups) {
            int nt, wt, ht, bt;
            decode_tile(tile, nt, wt, ht, bt);
            const int n0 = nt * p.BN;
            const int w0 = wt * (p.bw - p.halo) - p.halo, h0 = ht * (p.bh - p.halo) - p.halo, b0 = bt * p.bb;
            const int ab = p.a_batched ? b0 : 0, wb = p.b_batched ? b0 : 0;
            int tap = 0, kc_in_tap = 0;
            for (int ks = 0; ks < ksteps; ++ks) {
                int cw = w0, ch = h0, cp = 0, wi = p.src1_wi, kc;
                bool from1;
                if (ks < ksteps_main) {
                    kc = kc_in_tap;
                    cw += p.dw[tap]; ch += p.dh[tap]; cp = p.tp[tap]; wi = p.twi[tap];
                    from1 = kc >= p.chunks0;
                    if (from1) kc -= p.chunks0;
                    if (++kc_in_tap == ksteps_per_tap) { kc_in_tap = 0; ++tap; }
                } else {   // single-tap side source
                    kc = ks - ksteps_main;
                    from1 = true;
                }
                const int nsub = from1 ? p.nsub1 : p.nsub0;
                const bool a_has_lo = (from1 ? p.a1_lo : p.a0_lo) != 0;
                for (int sub = 0; sub < nsub; ++sub) {
                // sub 0: A.hi x W.hi;  then A.lo x W.hi (if the source has a lo plane);  then A.hi x W.lo (if the weight has one)
                const bool use_a_lo = sub == 1 && a_has_lo;
                const bool use_w_lo = sub >= 1 && !use_a_lo;
                mbar_wait_long(empty_bar(stage), phase ^ 1u, p.wait_hint_ns);
                const uint32_t sa = smem_base + stage * p.stage_bytes;
                const uint32_t sb = sa + kABytes;
                const int wc = (from1 ? p.w_c1_off : 0) + kc * kChunkK;
                const CUtensorMap *ta = from1 ? (use_a_lo ? &tmL.a1 : &tmA1) : (use_a_lo ? &tmL.a0 : &tmA0);
                const CUtensorMap *tw = use_w_lo ? &tmL.w : &tmW;
                const CUtensorMap *tw1 = use_w_lo ? &tmL.w1 : &tmW1;
                if (kPair) {
                    // both CTAs' loads complete on the LEADER's full barrier, which expects the bytes of both.  Within a CTA
                    // pair the shared::cluster address of the leader's copy of a barrier is this CTA's address with the rank
                    // bit (bit 24) cleared
                    const uint32_t fb = full_bar(stage) & 0xFEFFFFFFu;
                    if (elect_one()) {
                        if (rank == 0) mbar_arrive_expect_tx(full_bar(stage), 2u * p.stage_bytes);
                        tma_load_5d_pair(sa, ta, fb, kc * kChunkK, cw, ch, ab, cp);
                        tma_load_4d_pair(sb + p.wl_smem[0], tw, fb, wc, wi, n0 + p.wl_row0[0] + (int)rank * p.wl_rank_rows[0], wb);
                        if (p.n_wloads > 1)
                            tma_load_4d_pair(sb + p.wl_smem[1], tw1, fb, wc, wi, n0 + p.wl_row0[1] + (int)rank * p.wl_rank_rows[1], wb);
                    }
                } else {
                    if (elect_one()) {
                        mbar_arrive_expect_tx(full_bar(stage), p.stage_bytes);
                        tma_load_5d(sa, ta, full_bar(stage), kc * kChunkK, cw, ch, ab, cp);
                        tma_load_4d(sb + p.wl_smem[0], tw, full_bar(stage), wc, wi, n0 + p.wl_row0[0], wb);
                        if (p.n_wloads > 1) tma_load_4d(sb + p.wl_smem[1], tw, full_bar(stage), wc, wi, n0 + p.wl_row0[1], wb);
                    }
                }
                __syncwarp();
                if (++stage == p.num_stages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (whole warp, one elected lane issues; pair: the leader CTA issues for both) ======
        if (rank == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int as = 0;
            uint32_t aphase = 0;
            int seq = 0;   // index of this tile in the CTA's own sequence
            const uint32_t idesc0 = p.idesc0, idesc1 = p.idesc1;
            const bool two_parts = p.n_part1 > 0;
            constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) /* version 1 */ | (2u << 29) /* SWIZZLE_128B */;
            auto desc = [&](uint32_t lo) { return (static_cast<uint64_t>(kDescHi) << 32) | lo; };
            for (int tile = group; tile < p.total_tiles; tile += ngroups, ++seq) {
                // Accumulator hand-over.  Normal double buffering: wait until the epilogue has fully drained this
                // stage (its tile before last).  Staggered (BN > 256, 2*BN > 512 TMEM columns): stage 1 starts at
                // column 512-BN, so the stages share columns [512-BN, BN); the epilogue drains that shared range
                // of the PREVIOUS tile first and signals `tearly`, after which this tile may start while the rest of
                // the previous tile is still being drained.
                mbar_wait(tempty_bar(as), aphase ^ 1u);
                if (p.staggered && seq > 0) mbar_wait(tearly_bar(as ^ 1), ((uint32_t)(seq - 1) >> 1) & 1u);
                tc_fence_after();
                const uint32_t acc0 = tmem_base + as * p.acc_stride;
                const uint32_t acc1 = acc0 + p.n_part0;
                for (int ks = 0; ks < ksteps_mma; ++ks) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t sa = smem_base + stage * p.stage_bytes;
                    // K-major SWIZZLE_128B descriptors (make_sw128_desc): only the 14-bit address field changes per stage / k
                    const uint32_t a_lo = ((sa >> 4) & 0x3FFFu) | (1u << 16);
                    const uint32_t b0_lo = (((sa + kABytes) >> 4) & 0x3FFFu) | (1u << 16);
                    const uint32_t b1_lo = (((sa + kABytes + p.b1_smem_off) >> 4) & 0x3FFFu) | (1u << 16);
                    const uint32_t first = ks > 0 ? 1u : 0u;
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < kChunkK / 16; ++k) {
                            const uint32_t accum = k > 0 ? 1u : first;
                            if (kPair) umma_f16_pair(acc0, desc(a_lo + 2u * k), desc(b0_lo + 2u * k), idesc0, accum);
                            else umma_f16(acc0, desc(a_lo + 2u * k), desc(b0_lo + 2u * k), idesc0, accum);
                            if (two_parts) {
                                if (kPair) umma_f16_pair(acc1, desc(a_lo + 2u * k), desc(b1_lo + 2u * k), idesc1, accum);
                                else umma_f16(acc1, desc(a_lo + 2u * k), desc(b1_lo + 2u * k), idesc1, accum);
                            }
                        }
                        // frees this smem stage (in both CTAs of a pair) when the MMAs retire
                        if (kPair) umma_commit_pair(empty_bar(stage)); else umma_commit(empty_bar(stage));
                    }
                    __syncwarp();
                    if (++stage == p.num_stages) { stage = 0; phase ^= 1u; }
                }
                // accumulator complete -> epilogue (of both CTAs)
                if (elect_one()) {
                    if (kPair) umma_commit_pair(tfull_bar(as)); else umma_commit(tfull_bar(as));
                }
                __syncwarp();
                if (++as == p.acc_stages) { as = 0; aphase ^= 1u; }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue (8 warps; 16 in the fast variant) =====================
        // warp w owns TMEM lanes 32*(w%4)..+31 (hardware rule); the two warps that share a lane quarter
        // split the accumulator columns in interleaved 32-column chunks.
        const int q = warp & 3;
        const int half = (warp - 4) >> 2;
        const int etid = threadIdx.x - 128;              // 0..255
        const int r = q * 32 + lane;
        const int rw = r % p.bw;
        const int rh = (r / p.bw) % p.bh;
        const int rb = r / (p.bw * p.bh);
        float *sparams = reinterpret_cast<float *>(smem_raw + (bar_base + kBarBytes - smem_u32(smem_raw)));
        int as = 0;
        uint32_t aphase = 0;
        int pbuf = 0, last_nt = -1;
        uint32_t rphase = 0;
        const int nchunks = (p.BN + 31) >> 5;
        if (p.head_w != nullptr) {   // 1x1 head weights [3][BN] -> shared memory, once per CTA
            float *hw = sparams + 2 * 3 * kMaxBN;
            for (int i = etid; i < 3 * kMaxBN; i += 256) {
                const int o = i / kMaxBN, c = i - o * kMaxBN;
                hw[i] = (c < p.N_total) ? __ldg(p.head_w + o * p.N_total + c) : 0.f;
            }
            asm volatile("bar.sync 2, 256;" ::: "memory");
        }
        // accumulator hand-back barriers live in the leader CTA
        const uint32_t tempty_sig[2] = {kPair ? mapa_rank(tempty_bar(0), 0) : tempty_bar(0), kPair ? mapa_rank(tempty_bar(1), 0) : tempty_bar(1)};
        const uint32_t tearly_sig[2] = {kPair ? mapa_rank(tearly_bar(0), 0) : tearly_bar(0), kPair ? mapa_rank(tearly_bar(1), 0) : tearly_bar(1)};
        auto acc_signal = [&](uint32_t bar) { if (kPair) mbar_arrive_cluster(bar); else mbar_arrive(bar); };
        for (int tile = group; tile < p.total_tiles; tile += ngroups) {
            int nt, wt, ht, bt;
            decode_tile(tile, nt, wt, ht, bt);
            const int n0 = nt * p.BN;
            if constexpr (kFast) {
                // ---------------- warp-private fast epilogue ----------------
                // Sixteen warps: every warp owns the 32 rows of its TMEM lane quarter and every fourth 32-column chunk (four warps per
                // scheduler hide the TMEM-load / shared-memory latencies that two could not): its slice of the per-
                // column parameters, its 2 KB staging blocks (64B-swizzled), its own TMA stores (one per chunk, issued as soon
                // as the chunk is staged), its own residual loads and barrier.  No CTA-wide barrier is left in the tile loop.
                const int part = (warp - 4) >> 2;                              // 0..3: which chunks of the lane quarter are mine
                const int n_my = (nchunks - part + 3) >> 2;                    // <= 2 (can be 0 for narrow N tiles)
                float *wp = sparams + (warp - 4) * (3 * kWarpCols);
                const uint32_t my_res_bar = wres_bar(warp - 4);
                const bool res_tma = p.res_tma != 0, has_scale = p.scale != nullptr, shuffle = p.shuffle != 0;
                const bool x3_out = p.x3_out != 0, res_lo = p.res_lo != 0;      // split-precision output / residual (hi + lo planes)
                const uint32_t lo_off = p.lo_stage_off;
                const int gn = p.group_n;
                const float slope1 = p.slope1;
                // origin of this warp's 32-row slab inside the (bw x bh x bb) box (all box extents are powers of two)
                const int row0 = q * 32;
                const int sw0 = wt * p.bw + row0 % p.bw, sh0 = ht * p.bh + (row0 / p.bw) % p.bh, sb0 = bt * p.bb + row0 / (p.bw * p.bh);
                auto block = [&](int ci) { return stage_out + (uint32_t)(ci * 4 + q) * 2048u; };
                // parameters of my columns: fetched now (when the N tile changed), parked in shared memory after the accumulator wait
                float pb[2], ps[2], pt[2];
                const bool reload = nt != last_nt;
                if (reload) {
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        const int n = n0 + (part + 4 * k) * 32 + lane;
                        const bool in = k < n_my;
                        pb[k] = (in && p.bias) ? __ldg(p.bias + n) : 0.f;
                        ps[k] = (in && p.scale) ? __ldg(p.scale + n) : 1.f;
                        pt[k] = (in && p.shift) ? __ldg(p.shift + n) : 0.f;
                    }
                    last_nt = nt;
                }
                if (lane == 0) {
                    bulk_wait_read<0>();       // my stores of the previous tile have finished reading my staging blocks
                    if (res_tma && n_my > 0) { // residual slab -> the same blocks (rows / channels outside the tensor arrive as zeros)
                        mbar_arrive_expect_tx(my_res_bar, (uint32_t)n_my * (res_lo ? 4096u : 2048u));
                        for (int k = 0; k < n_my; ++k) {
                            tma_load_5d(block(part + 4 * k), &tmO.m[1], my_res_bar, n0 + (part + 4 * k) * 32, sw0, sh0, sb0, 0);
                            if (res_lo) tma_load_5d(block(part + 4 * k) + lo_off, &tmO.m[3], my_res_bar, n0 + (part + 4 * k) * 32, sw0, sh0, sb0, 0);
                        }
                    }
                }
                __syncwarp();
                if (p.residual != nullptr) {   // pull the NEXT tile's residual rows towards L2 while this tile is processed
                    const int tn = tile + ngroups;
                    if (tn < p.total_tiles) {
                        int nt2, wt2, ht2, bt2;
                        decode_tile(tn, nt2, wt2, ht2, bt2);
                        const int ow2 = wt2 * p.bw + rw, oh2 = ht2 * p.bh + rh, ob2 = bt2 * p.bb + rb;
                        if (ow2 < p.out_W && oh2 < p.out_H && ob2 < p.out_B) {
                            const uint8_t *row = reinterpret_cast<const uint8_t *>(p.residual) + 2ll * (ob2 * p.rsb + oh2 * p.rsh + ow2 * p.rsw);
                            const int cbeg = nt2 * p.BN, cend = min(cbeg + p.BN, p.c_store);
                            for (int c = cbeg + part * 64; c < cend; c += 256)
                                asm volatile("prefetch.global.L2 [%0];" ::"l"(row + 2 * c));
                        }
                    }
                }
                mbar_wait_long(tfull_bar(as), aphase, p.wait_hint_ns);
                tc_fence_after();
                const uint32_t tbase = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * p.acc_stride;
                uint32_t va[32];
                if (reload) {
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        wp[k * 32 + lane] = pb[k];
                        wp[kWarpCols + k * 32 + lane] = ps[k];
                        wp[2 * kWarpCols + k * 32 + lane] = pt[k];
                    }
                    __syncwarp();
                }
                if (res_tma && n_my > 0) { mbar_wait(my_res_bar, rphase); rphase ^= 1u; }
                const uint32_t swz = ((uint32_t)lane >> 1) & 3u;               // SWIZZLE_64B: 16-byte chunk index ^= address bits [7:8]
                // kAct: 0 = none, 1 = ReLU, 2 = run-time slope (LeakyReLU); the other flags switch whole stages off at compile time
                auto chunk = [&](int k, uint32_t(&vc)[32], auto kActT, auto kScaleT, auto kResT, auto kRelu2T, auto kX3T) {
                    constexpr int kAct = decltype(kActT)::value;
                    constexpr bool kScale = decltype(kScaleT)::value, kRes = decltype(kResT)::value, kRelu2 = decltype(kRelu2T)::value;
                    constexpr bool kX3 = decltype(kX3T)::value;
                    const int ci = part + 4 * k;
                    __syncwarp();
                    tmem_ld32(tbase + ci * 32, vc);
                    const uint32_t rowaddr = block(ci) + (uint32_t)lane * 64u;
                    uint4 rres[4], rlo[4];
                    if constexpr (kRes) {
#pragma unroll
                        for (int g = 0; g < 4; ++g)
                            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(rres[g].x), "=r"(rres[g].y), "=r"(rres[g].z), "=r"(rres[g].w)
                                         : "r"(rowaddr + (((uint32_t)g ^ swz) << 4)) : "memory");
                        if constexpr (kX3) {
#pragma unroll
                            for (int g = 0; g < 4; ++g) {
                                rlo[g] = make_uint4(0u, 0u, 0u, 0u);
                                if (res_lo)
                                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(rlo[g].x), "=r"(rlo[g].y), "=r"(rlo[g].z), "=r"(rlo[g].w)
                                                 : "r"(rowaddr + lo_off + (((uint32_t)g ^ swz) << 4)) : "memory");
                            }
                        }
                    }
                    tmem_ld_wait();
                    const float *sb = wp + k * 32;
                    auto act = [&](float t) {
                        if constexpr (kAct == 0) return t;
                        else if constexpr (kAct == 1) return fmaxf(t, 0.f);
                        else return fmaxf(t, t * slope1);
                    };
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        float y[16];
                        const float4 *b4 = reinterpret_cast<const float4 *>(sb + 16 * hh);
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            const float4 bv = b4[g];
                            const float t0 = __uint_as_float(vc[16 * hh + 4 * g + 0]) + bv.x, t1 = __uint_as_float(vc[16 * hh + 4 * g + 1]) + bv.y;
                            const float t2 = __uint_as_float(vc[16 * hh + 4 * g + 2]) + bv.z, t3 = __uint_as_float(vc[16 * hh + 4 * g + 3]) + bv.w;
                            y[4 * g + 0] = act(t0);
                            y[4 * g + 1] = act(t1);
                            y[4 * g + 2] = act(t2);
                            y[4 * g + 3] = act(t3);
                        }
                        if constexpr (kScale) {
                            const float4 *s4 = reinterpret_cast<const float4 *>(sb + kWarpCols + 16 * hh);
                            const float4 *t4 = reinterpret_cast<const float4 *>(sb + 2 * kWarpCols + 16 * hh);
#pragma unroll
                            for (int g = 0; g < 4; ++g) {
                                const float4 sv = s4[g], tv = t4[g];
                                y[4 * g + 0] = fmaf(y[4 * g + 0], sv.x, tv.x);
                                y[4 * g + 1] = fmaf(y[4 * g + 1], sv.y, tv.y);
                                y[4 * g + 2] = fmaf(y[4 * g + 2], sv.z, tv.z);
                                y[4 * g + 3] = fmaf(y[4 * g + 3], sv.w, tv.w);
                            }
                        }
                        if constexpr (kRes) {
                            const uint32_t rr[8] = {rres[2 * hh].x, rres[2 * hh].y, rres[2 * hh].z, rres[2 * hh].w,
                                                    rres[2 * hh + 1].x, rres[2 * hh + 1].y, rres[2 * hh + 1].z, rres[2 * hh + 1].w};
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float2 f = unpack2(rr[j], kDT);
                                y[2 * j] += f.x;
                                y[2 * j + 1] += f.y;
                            }
                            if constexpr (kX3) {
                                const uint32_t rl[8] = {rlo[2 * hh].x, rlo[2 * hh].y, rlo[2 * hh].z, rlo[2 * hh].w,
                                                        rlo[2 * hh + 1].x, rlo[2 * hh + 1].y, rlo[2 * hh + 1].z, rlo[2 * hh + 1].w};
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    const float2 f = unpack2(rl[j], kDT);
                                    y[2 * j] += f.x;
                                    y[2 * j + 1] += f.y;
                                }
                            }
                        }
                        if constexpr (kRelu2) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) y[j] = fmaxf(y[j], 0.f);
                        }
                        uint32_t hp[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) hp[j] = pack2(y[2 * j], y[2 * j + 1], kDT);
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowaddr + (((uint32_t)(2 * hh) ^ swz) << 4)), "r"(hp[0]),
                                     "r"(hp[1]), "r"(hp[2]), "r"(hp[3]) : "memory");
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowaddr + (((uint32_t)(2 * hh + 1) ^ swz) << 4)), "r"(hp[4]),
                                     "r"(hp[5]), "r"(hp[6]), "r"(hp[7]) : "memory");
                        if constexpr (kX3) {     // lo plane: what the 16-bit rounding of the hi plane dropped, rounded to 16 bit itself
                            uint32_t lp[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float2 h = unpack2(hp[j], kDT);
                                lp[j] = pack2(y[2 * j] - h.x, y[2 * j + 1] - h.y, kDT);
                            }
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowaddr + lo_off + (((uint32_t)(2 * hh) ^ swz) << 4)), "r"(lp[0]),
                                         "r"(lp[1]), "r"(lp[2]), "r"(lp[3]) : "memory");
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowaddr + lo_off + (((uint32_t)(2 * hh + 1) ^ swz) << 4)), "r"(lp[4]),
                                         "r"(lp[5]), "r"(lp[6]), "r"(lp[7]) : "memory");
                        }
                    }
                    // hand the staged chunk to the TMA unit: generic-proxy writes -> async proxy, then one bulk store
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) {
                        const int nn = n0 + ci * 32;
                        if (shuffle) {
                            const int g = (nn >= gn) + (nn >= 2 * gn) + (nn >= 3 * gn);
                            tma_store_5d(&tmO.m[g], block(ci), nn - g * gn, sw0, sh0, sb0, 0);
                        } else {
                            tma_store_5d(&tmO.m[0], block(ci), nn, sw0, sh0, sb0, 0);
                            if constexpr (kX3) tma_store_5d(&tmO.m[2], block(ci) + lo_off, nn, sw0, sh0, sb0, 0);
                        }
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                };
                // the four stage combinations the networks use get their own straight-line body; anything else takes the general one
                using T = std::true_type;
                using F = std::false_type;
                using A0 = std::integral_constant<int, 0>;
                using A1 = std::integral_constant<int, 1>;
                using A2 = std::integral_constant<int, 2>;
                const bool relu1 = slope1 == 0.f, noact = slope1 == 1.f, relu2 = p.relu2 != 0;
                if (x3_out) {   // split-precision output (encoders): the same stage combinations, storing hi and lo planes
                    if (relu1 && !has_scale && !res_tma && !relu2) {
                        for (int k = 0; k < n_my; ++k) chunk(k, va, A1{}, F{}, F{}, F{}, T{});
                    } else if (noact && !has_scale && res_tma && relu2) {
                        for (int k = 0; k < n_my; ++k) chunk(k, va, A0{}, F{}, T{}, T{}, T{});
                    } else if (!res_tma) {
                        if (relu2) { for (int k = 0; k < n_my; ++k) chunk(k, va, A2{}, T{}, F{}, T{}, T{}); }
                        else { for (int k = 0; k < n_my; ++k) chunk(k, va, A2{}, T{}, F{}, F{}, T{}); }
                    } else {
                        if (relu2) { for (int k = 0; k < n_my; ++k) chunk(k, va, A2{}, T{}, T{}, T{}, T{}); }
                        else { for (int k = 0; k < n_my; ++k) chunk(k, va, A2{}, T{}, T{}, F{}, T{}); }
                    }
                } else if (relu1 && !has_scale && !res_tma && !relu2) {            // conv + folded BN + ReLU, PixelShuffle convs
                    for (int k = 0; k < n_my; ++k) chunk(k, va, A1{}, F{}, F{}, F{}, F{});
                } else if (relu1 && has_scale && !res_tma && !relu2) {      // custom_conv_layer: conv -> ReLU -> BN
                    for (int k = 0; k < n_my; ++k) chunk(k, va, A1{}, T{}, F{}, F{}, F{});
                } else if (noact && !has_scale && res_tma && relu2) {       // bottleneck conv3: + identity, ReLU
                    for (int k = 0; k < n_my; ++k) chunk(k, va, A0{}, F{}, T{}, T{}, F{});
                } else if (!res_tma) {
                    if (relu2) { for (int k = 0; k < n_my; ++k) chunk(k, va, A2{}, T{}, F{}, T{}, F{}); }
                    else { for (int k = 0; k < n_my; ++k) chunk(k, va, A2{}, T{}, F{}, F{}, F{}); }
                } else {
                    if (relu2) { for (int k = 0; k < n_my; ++k) chunk(k, va, A2{}, T{}, T{}, T{}, F{}); }
                    else { for (int k = 0; k < n_my; ++k) chunk(k, va, A2{}, T{}, T{}, F{}, F{}); }
                }
                tc_fence_before();
                acc_signal(tempty_sig[as]);
                if (++as == p.acc_stages) { as = 0; aphase ^= 1u; }
                continue;
            }
            if (p.blur) {
                // ---------------- fused PixelShuffle + blur epilogue (generic kernel, 8 epilogue warps) ----------------
                // N tile = [4 sub-pixel groups][CW = BN/4 channels].  Phase 1: +bias, ReLU, 16-bit, whole tile -> shared memory
                // (row = low-resolution pixel incl. the halo row / column; padded row pitch: conflict-free 16-byte accesses).
                // Phase 2: every (output pixel, 8-channel chunk) averages its four contributors; 8 consecutive lanes write one
                // full 128-byte line of the NHWC output.
                const uint32_t pitch = p.blur_row_bytes;
                if (last_nt == -1) {     // once per CTA: the bias of EVERY N tile (N_total <= kEpiSmemFloats) - the N tile changes with every tile
                    for (int i = etid; i < p.N_total; i += 256) sparams[i] = p.bias ? __ldg(p.bias + i) : 0.f;
                    last_nt = 0;
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");      // phase 2 of the previous tile has finished with the staging tile (+ bias visible)
                mbar_wait_long(tfull_bar(as), aphase, p.wait_hint_ns);
                tc_fence_after();
                const uint32_t tb = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * p.acc_stride;
                const uint32_t myrow = stage_out + (uint32_t)r * pitch;
                uint32_t va[32], vb[32];
                // this warp's chunks: half, half + 2, ... (four of the eight); the TMEM load of the next one is in flight while one is converted
                auto emit = [&](int ci, uint32_t(&v)[32]) {
                    const float4 *bs4 = reinterpret_cast<const float4 *>(sparams + n0 + ci * 32);
#pragma unroll
                    for (int g4 = 0; g4 < 4; ++g4) {
                        const float4 b0 = bs4[2 * g4], b1 = bs4[2 * g4 + 1];
                        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                        uint32_t o[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float y0 = fmaxf(__uint_as_float(v[8 * g4 + 2 * j]) + bb[2 * j], 0.f);
                            const float y1 = fmaxf(__uint_as_float(v[8 * g4 + 2 * j + 1]) + bb[2 * j + 1], 0.f);
                            o[j] = pack2(y0, y1, kDT);
                        }
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(myrow + (uint32_t)(ci * 64 + g4 * 16)), "r"(o[0]),
                                     "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
                    }
                };
                __syncwarp();
                tmem_ld32(tb + half * 32, va);
#pragma unroll
                for (int kk = 0; kk < 4; kk += 2) {
                    tmem_ld_wait();
                    __syncwarp();
                    tmem_ld32(tb + (half + 2 * (kk + 1)) * 32, vb);
                    emit(half + 2 * kk, va);
                    tmem_ld_wait();
                    if (kk + 2 < 4) {
                        __syncwarp();
                        tmem_ld32(tb + (half + 2 * (kk + 2)) * 32, va);
                    }
                    emit(half + 2 * (kk + 1), vb);
                }
                tc_fence_before();
                acc_signal(tempty_sig[as]);                          // the accumulator is out of TMEM: the next tile may start
                if (++as == p.acc_stages) { as = 0; aphase ^= 1u; }
                asm volatile("bar.sync 2, 256;" ::: "memory");      // the whole tile is staged
                // phase 2 (BN = 256, box 16 x 8: compile-time geometry): thread = 16-byte channel chunk k = etid & 7 of the low-resolution
                // pixels row = etid >> 3, + 32, ...  It loads the 3 x 3 high-resolution neighbourhood S[-1..1][-1..1] of the pixel's 2 x 2
                // outputs (own four sub-pixels, two of the left neighbour, two of the upper one, one of the upper-left one; redirected
                // onto the pixel's own values on the top / left image edge = replication padding), forms the horizontal pair sums once
                // and writes the four outputs: ((S[y-1][x-1] + S[y-1][x]) + (S[y][x-1] + S[y][x])) / 4, havc_blur2x2's association and
                // arithmetic (packed 16-bit adds).
                {
                    constexpr int kBW = 16, kCW = 64;
                    const int k = etid & 7;
                    const int w_base = wt * (kBW - 1) - 1, h_base = ht * (p.bh - 1) - 1;
                    const int Wl = p.out_W, Hl = p.out_H;
                    const int chan = nt * kCW + k * 8;
                    const bool chan_ok = chan < p.c_store;
                    uint16_t *obase = reinterpret_cast<uint16_t *>(p.out) + bt * p.osb + chan;
                    auto grp = [&](int g) -> uint32_t { return (uint32_t)((g * kCW + k * 8) * 2); };   // byte offset of sub-pixel group g in a row
                    auto lds4 = [&](uint32_t addr, uint32_t(&w)[4]) {
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(addr) : "memory");
                    };
                    for (int row = (etid >> 3); row < kTileM; row += 32) {
                        const int trw = row & (kBW - 1), trh = row >> 4;
                        const int pw = w_base + trw, ph = h_base + trh;
                        if (trw == 0 || trh == 0 || pw >= Wl || ph >= Hl || !chan_ok) continue;   // halo rows / columns only contribute
                        const uint32_t own = stage_out + (uint32_t)row * pitch;
                        const uint32_t left = pw > 0 ? own - pitch : own;                 // X - 1 clamps onto X at the left image edge
                        const uint32_t up = ph > 0 ? own - kBW * pitch : own;             // Y - 1 clamps onto Y at the top image edge
                        const uint32_t upleft = ph > 0 ? left - kBW * pitch : left;
                        // sub-pixel groups g = 2 a + b.  Column X-1 is sub-column b = 1 of the left neighbour (b = 0 of the pixel itself
                        // when clamped); row Y-1 is sub-row a = 1 of the upper neighbour (a = 0 of the pixel itself when clamped)
                        const int gl = pw > 0 ? 1 : 0, gu = ph > 0 ? 2 : 0;
                        // packed 16-bit sums (add2: two channels per instruction, no conversions), havc_blur2x2's operations and order
                        uint32_t s[4], t[4], hm[2][4], h0[2][4], h1[2][4];
                        // row Y-1: S[-1][-1], S[-1][0], S[-1][1]
                        lds4(upleft + grp(gu + gl), s); lds4(up + grp(gu), t);
#pragma unroll
                        for (int j = 0; j < 4; ++j) hm[0][j] = add2(s[j], t[j], kDT);
                        lds4(up + grp(gu + 1), s);
#pragma unroll
                        for (int j = 0; j < 4; ++j) hm[1][j] = add2(t[j], s[j], kDT);
                        // row Y (a = 0): S[0][-1], S[0][0], S[0][1]
                        lds4(left + grp(gl), s); lds4(own + grp(0), t);
#pragma unroll
                        for (int j = 0; j < 4; ++j) h0[0][j] = add2(s[j], t[j], kDT);
                        lds4(own + grp(1), s);
#pragma unroll
                        for (int j = 0; j < 4; ++j) h0[1][j] = add2(t[j], s[j], kDT);
                        // row Y+1 (a = 1): S[1][-1], S[1][0], S[1][1]
                        lds4(left + grp(2 + gl), s); lds4(own + grp(2), t);
#pragma unroll
                        for (int j = 0; j < 4; ++j) h1[0][j] = add2(s[j], t[j], kDT);
                        lds4(own + grp(3), s);
#pragma unroll
                        for (int j = 0; j < 4; ++j) h1[1][j] = add2(t[j], s[j], kDT);
                        uint16_t *dst = obase + (long long)(2 * ph) * p.osh + (long long)(2 * pw) * p.osw;
#pragma unroll
                        for (int bq = 0; bq < 2; ++bq) {
                            uint32_t o0[4], o1[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                o0[j] = quarter2(add2(hm[bq][j], h0[bq][j], kDT), kDT);
                                o1[j] = quarter2(add2(h0[bq][j], h1[bq][j], kDT), kDT);
                            }
                            *reinterpret_cast<uint4 *>(dst + bq * p.osw) = make_uint4(o0[0], o0[1], o0[2], o0[3]);
                            *reinterpret_cast<uint4 *>(dst + p.osh + bq * p.osw) = make_uint4(o1[0], o1[1], o1[2], o1[3]);
                        }
                    }
                }
                continue;
            }
            const int ow = wt * p.bw + rw, oh = ht * p.bh + rh, ob = bt * p.bb + rb;
            const bool valid = (ow < p.out_W) && (oh < p.out_H) && (ob < p.out_B);

            if (p.tma_store && etid == 0)   // previous tile's bulk stores must have finished reading the staging buffer
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            // stage this tile's per-column parameters in shared memory (double-buffered; reloaded only when the N tile
            // changes, i.e. never after the first tile of a launch with a single N tile)
            if (nt != last_nt) {
                pbuf ^= 1;
                float *sbw = sparams + pbuf * (3 * kMaxBN);
                for (int i = etid; i < p.BN; i += 256) {
                    const int n = n0 + i;
                    const bool in = n < p.N_total;
                    sbw[i] = (in && p.bias) ? __ldg(p.bias + n) : 0.f;
                    sbw[kMaxBN + i] = (in && p.scale) ? __ldg(p.scale + n) : 1.f;
                    sbw[2 * kMaxBN + i] = (in && p.shift) ? __ldg(p.shift + n) : 0.f;
                }
                last_nt = nt;
                asm volatile("bar.sync 1, 256;" ::: "memory");
            } else if (p.tma_store) {
                // no parameter reload: the barrier is still what orders thread 0's wait on the previous tile's bulk
                // store (above) before any warp overwrites the staging buffer
                asm volatile("bar.sync 1, 256;" ::: "memory");
            }
            const float *sb = sparams + pbuf * (3 * kMaxBN);

            const uint8_t *res_row = nullptr, *res_row2 = nullptr;
            if (p.residual != nullptr && valid)
                res_row = reinterpret_cast<const uint8_t *>(p.residual) + 2ll * (ob * p.rsb + oh * p.rsh + ow * p.rsw);
            if (p.residual2 != nullptr && valid)
                res_row2 = reinterpret_cast<const uint8_t *>(p.residual2) + 2ll * (ob * p.r2sb + oh * p.r2sh + ow * p.r2sw);
            // Fused-head partial sums.  The order in which a warp walks its chunks depends on the accumulator stage the tile
            // happens to use (staggered mode drains the shared columns first), i.e. on the tile's position in the launch: a
            // plain fp32 running sum would make a pixel's logits depend on the batch slot of its frame in the last bits.  Every
            // chunk therefore sums into its own register slot (chunk index -> slot through predicated static selects) and the
            // slots are added in ascending order at the end of the tile, whatever order they were filled in.
            float hp0[5] = {0.f, 0.f, 0.f, 0.f, 0.f}, hp1[5] = {0.f, 0.f, 0.f, 0.f, 0.f}, hp2[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
            if (p.residual != nullptr) {
                // pull the NEXT tile's residual rows towards L2 while this tile is processed: each thread covers the
                // 128-byte lines of its own row that its warp half will read
                const int tn = tile + ngroups;
                if (tn < p.total_tiles) {
                    int nt2, wt2, ht2, bt2;
                    decode_tile(tn, nt2, wt2, ht2, bt2);
                    const int ow2 = wt2 * p.bw + rw, oh2 = ht2 * p.bh + rh, ob2 = bt2 * p.bb + rb;
                    if (ow2 < p.out_W && oh2 < p.out_H && ob2 < p.out_B) {
                        const uint8_t *row = reinterpret_cast<const uint8_t *>(p.residual) + 2ll * (ob2 * p.rsb + oh2 * p.rsh + ow2 * p.rsw);
                        const int cbeg = nt2 * p.BN, cend = min(cbeg + p.BN, p.c_store);
                        for (int c = cbeg + half * 64; c < cend; c += 128)
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(row + 2 * c));
                    }
                }
            }

            mbar_wait_long(tfull_bar(as), aphase, p.wait_hint_ns);
            tc_fence_after();
            const uint32_t tbase = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * p.acc_stride;

            uint32_t va[32], vb[32];
            // ---- per-tile invariants of this thread (hoisted out of the chunk loop) ----
            const int BN = p.BN, N_total = p.N_total, split_n = p.split_n, c_store = p.c_store, c_store2 = p.c_store2;
            constexpr int dt = kDT;
            const int od = p.out_dtype == HAVC_F32 ? HAVC_F32 : kDT, gn = p.group_n;
            const bool shuffle = p.shuffle != 0, has_scale = p.scale != nullptr, head = p.head_w != nullptr;
            const bool tma_store = p.tma_store != 0;
            const float slope1 = p.slope1, lo2 = p.relu2 ? 0.f : -INFINITY;
            const long long pix_main = ob * p.osb + (long long)(oh * p.up + p.oy) * p.osh + (long long)(ow * p.up + p.ox) * p.osw;
            const long long pix_shuf = ob * p.osb + (long long)(oh * 2) * p.osh + (long long)(ow * 2) * p.osw;
            const long long pix_out2 = ob * p.o2sb + (long long)(oh * p.up + p.oy) * p.o2sh + (long long)(ow * p.up + p.ox) * p.o2sw;
            const float *hw = sparams + 2 * 3 * kMaxBN;
            // One chunk: wait for its TMEM load, kick off the load of the next chunk of this warp into `vn`,
            // then run the fp32 epilogue on `vc` and store.
            auto process = [&](int ci, int nxt, uint32_t(&vc)[32], uint32_t(&vn)[32]) {
                const int c0 = ci * 32;
                uint4 rres[4];
                const int n = n0 + c0;
                const bool do_res = res_row != nullptr && n < N_total;
                if (do_res) {  // residual prefetch: global loads overlap the TMEM load latency
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        rres[g] = make_uint4(0, 0, 0, 0);
                        const int ng = n + 8 * g;
                        if (c0 + 8 * g >= BN) continue;
                        if (ng < split_n) {
                            if (ng < c_store) rres[g] = __ldg(reinterpret_cast<const uint4 *>(res_row + 2 * ng));
                        } else if (res_row2 != nullptr && ng - split_n < c_store2) {
                            rres[g] = __ldg(reinterpret_cast<const uint4 *>(res_row2 + 2 * (ng - split_n)));
                        }
                    }
                }
                tmem_ld_wait();
                if (nxt >= 0) {
                    __syncwarp();
                    tmem_ld_chunk(tbase + nxt * 32, vn, BN - nxt * 32);
                }
                if (!valid) return;
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const int cc = c0 + 16 * hh;       // column within the tile
                    const int nn = n0 + cc;            // GEMM column
                    if (cc >= BN || nn >= N_total) continue;
                    float y[16];
                    {   // +bias, ReLU  (parameters are read as 4 x LDS.128 broadcasts)
                        const float4 *b4 = reinterpret_cast<const float4 *>(sb + cc);
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            const float4 bv = b4[g];
                            const float t0 = __uint_as_float(vc[16 * hh + 4 * g + 0]) + bv.x, t1 = __uint_as_float(vc[16 * hh + 4 * g + 1]) + bv.y;
                            const float t2 = __uint_as_float(vc[16 * hh + 4 * g + 2]) + bv.z, t3 = __uint_as_float(vc[16 * hh + 4 * g + 3]) + bv.w;
                            y[4 * g + 0] = fmaxf(t0, t0 * slope1);
                            y[4 * g + 1] = fmaxf(t1, t1 * slope1);
                            y[4 * g + 2] = fmaxf(t2, t2 * slope1);
                            y[4 * g + 3] = fmaxf(t3, t3 * slope1);
                        }
                    }
                    if (has_scale) {
                        const float4 *s4 = reinterpret_cast<const float4 *>(sb + kMaxBN + cc);
                        const float4 *t4 = reinterpret_cast<const float4 *>(sb + 2 * kMaxBN + cc);
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            const float4 sv = s4[g], tv = t4[g];
                            y[4 * g + 0] = fmaf(y[4 * g + 0], sv.x, tv.x);
                            y[4 * g + 1] = fmaf(y[4 * g + 1], sv.y, tv.y);
                            y[4 * g + 2] = fmaf(y[4 * g + 2], sv.z, tv.z);
                            y[4 * g + 3] = fmaf(y[4 * g + 3], sv.w, tv.w);
                        }
                    }
                    if (do_res) {
                        const uint32_t rr[8] = {rres[2 * hh].x, rres[2 * hh].y, rres[2 * hh].z, rres[2 * hh].w,
                                                rres[2 * hh + 1].x, rres[2 * hh + 1].y, rres[2 * hh + 1].z,
                                                rres[2 * hh + 1].w};
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float2 f = unpack2(rr[j], dt);
                            y[2 * j] += f.x;
                            y[2 * j + 1] += f.y;
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 16; ++j) y[j] = fmaxf(y[j], lo2);
                    if (head) {   // fused 1x1 head: three dot products over this thread's columns
                        const float4 *h0 = reinterpret_cast<const float4 *>(hw + cc);
                        const float4 *h1 = reinterpret_cast<const float4 *>(hw + kMaxBN + cc);
                        const float4 *h2 = reinterpret_cast<const float4 *>(hw + 2 * kMaxBN + cc);
                        float l0 = 0.f, l1 = 0.f, l2 = 0.f;
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            const float4 a = h0[g], b = h1[g], c = h2[g];
                            l0 = fmaf(y[4 * g + 3], a.w, fmaf(y[4 * g + 2], a.z, fmaf(y[4 * g + 1], a.y, fmaf(y[4 * g], a.x, l0))));
                            l1 = fmaf(y[4 * g + 3], b.w, fmaf(y[4 * g + 2], b.z, fmaf(y[4 * g + 1], b.y, fmaf(y[4 * g], b.x, l1))));
                            l2 = fmaf(y[4 * g + 3], c.w, fmaf(y[4 * g + 2], c.z, fmaf(y[4 * g + 1], c.y, fmaf(y[4 * g], c.x, l2))));
                        }
                        const int slot = (ci - half) >> 1;      // this warp's chunks are half, half + 2, ...: at most 5 per tile
#pragma unroll
                        for (int j = 0; j < 5; ++j)
                            if (j == slot) { hp0[j] += l0; hp1[j] += l1; hp2[j] += l2; }
                        continue;
                    }
                    if (tma_store) {   // stage into the swizzled 64-channel sub-tile; garbage rows/columns are clipped by TMA
                        const uint32_t sub = stage_out + (uint32_t)(cc >> 6) * (kTileM * 128u) + (uint32_t)r * 128u;
                        const uint32_t k16 = (uint32_t)(cc & 63) >> 3;
                        const uint4 v0 = make_uint4(pack2(y[0], y[1], od), pack2(y[2], y[3], od), pack2(y[4], y[5], od),
                                                    pack2(y[6], y[7], od));
                        const uint4 v1 = make_uint4(pack2(y[8], y[9], od), pack2(y[10], y[11], od), pack2(y[12], y[13], od),
                                                    pack2(y[14], y[15], od));
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sub + ((k16 ^ (r & 7u)) << 4)), "r"(v0.x),
                                     "r"(v0.y), "r"(v0.z), "r"(v0.w) : "memory");
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sub + (((k16 + 1u) ^ (r & 7u)) << 4)),
                                     "r"(v1.x), "r"(v1.y), "r"(v1.z), "r"(v1.w) : "memory");
                        continue;
                    }
                    // destination
                    long long off;
                    int chan, cs;
                    void *obase = p.out;
                    if (shuffle) {   // column group g -> sub-pixel (g>>1, g&1); no division: groups are >= 16 wide
                        const int g = (nn >= gn) + (nn >= 2 * gn) + (nn >= 3 * gn);
                        chan = nn - g * gn;
                        cs = c_store;
                        off = pix_shuf + (g >> 1) * p.osh + (g & 1) * p.osw + chan;
                    } else if (nn >= split_n) {
                        chan = nn - split_n;
                        cs = c_store2;
                        off = pix_out2 + chan;
                        obase = p.out2;
                    } else {
                        chan = nn;
                        cs = c_store;
                        off = pix_main + chan;
                    }
                    if (chan >= cs) continue;
                    const bool hi_ok = (chan + 8) < cs;
                    if (od == HAVC_F32) {
                        float4 *dst = reinterpret_cast<float4 *>(reinterpret_cast<float *>(obase) + off);
                        dst[0] = make_float4(y[0], y[1], y[2], y[3]);
                        dst[1] = make_float4(y[4], y[5], y[6], y[7]);
                        if (hi_ok) {
                            dst[2] = make_float4(y[8], y[9], y[10], y[11]);
                            dst[3] = make_float4(y[12], y[13], y[14], y[15]);
                        }
                    } else {
                        uint4 *dst = reinterpret_cast<uint4 *>(reinterpret_cast<uint16_t *>(obase) + off);
                        dst[0] = make_uint4(pack2(y[0], y[1], od), pack2(y[2], y[3], od), pack2(y[4], y[5], od),
                                            pack2(y[6], y[7], od));
                        if (hi_ok)
                            dst[1] = make_uint4(pack2(y[8], y[9], od), pack2(y[10], y[11], od),
                                                pack2(y[12], y[13], od), pack2(y[14], y[15], od));
                    }
                }
            };
            // This warp's chunks are ci = half, half+2, ...; in staggered mode the columns shared with the other
            // accumulator stage are drained first (stage 0: the highest chunks, stage 1: the lowest) and `tearly`
            // is signalled as soon as they are out of TMEM.
            const int n_my = (nchunks - half + 1) >> 1;
            const bool desc = p.staggered && as == 0;
            auto chunk_at = [&](int k) { return desc ? (half + 2 * (n_my - 1 - k)) : (half + 2 * k); };
            auto shared_cols = [&](int ci) {   // does local chunk ci touch columns [acc_stride, BN) of the other stage?
                if (!p.staggered) return false;
                return as == 0 ? (ci * 32 + 32 > p.acc_stride) : (ci * 32 < p.BN - p.acc_stride);
            };
            bool early_done = !p.staggered;
            if (n_my > 0) {
                __syncwarp();
                tmem_ld_chunk(tbase + chunk_at(0) * 32, va, p.BN - chunk_at(0) * 32);
            }
            for (int k = 0; k < n_my; k += 2) {
                if (!early_done && !shared_cols(chunk_at(k))) {
                    tc_fence_before();
                    acc_signal(tearly_sig[as]);
                    early_done = true;
                }
                process(chunk_at(k), k + 1 < n_my ? chunk_at(k + 1) : -1, va, vb);
                if (k + 1 < n_my) {
                    if (!early_done && !shared_cols(chunk_at(k + 1))) {
                        tc_fence_before();
                        acc_signal(tearly_sig[as]);
                        early_done = true;
                    }
                    process(chunk_at(k + 1), k + 2 < n_my ? chunk_at(k + 2) : -1, vb, va);
                }
            }
            if (!early_done) {
                tc_fence_before();
                acc_signal(tearly_sig[as]);
            }
            if (tma_store) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> async proxy
                asm volatile("bar.sync 2, 256;" ::: "memory");
                if (etid == 0) {
                    const int w0 = wt * p.bw, h0 = ht * p.bh, b0 = bt * p.bb;
                    for (int sub = 0; sub < (BN >> 6); ++sub) {
                        const int nn = n0 + sub * 64;
                        if (nn >= N_total) break;
                        if (shuffle) {
                            const int g = (nn >= gn) + (nn >= 2 * gn) + (nn >= 3 * gn);
                            tma_store_5d(&tmO.m[g], stage_out + sub * (kTileM * 128u), nn - g * gn, w0, h0, b0, 0);
                        } else {
                            tma_store_5d(&tmO.m[0], stage_out + sub * (kTileM * 128u), nn, w0, h0, b0, 0);
                        }
                    }
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
            if (p.head_w != nullptr) {   // the two warps of a lane quarter each hold half of the columns
                float *hx = sparams + (2 * 3 + 3) * kMaxBN;   // [128 rows][4]
                const float hacc0 = (((hp0[0] + hp0[1]) + hp0[2]) + hp0[3]) + hp0[4];
                const float hacc1 = (((hp1[0] + hp1[1]) + hp1[2]) + hp1[3]) + hp1[4];
                const float hacc2 = (((hp2[0] + hp2[1]) + hp2[2]) + hp2[3]) + hp2[4];
                if (half == 1) {
                    hx[r * 4 + 0] = hacc0; hx[r * 4 + 1] = hacc1; hx[r * 4 + 2] = hacc2;
                }
                asm volatile("bar.sync 2, 256;" ::: "memory");
                if (half == 0 && valid) {
                    const float4 o = make_float4(hacc0 + hx[r * 4 + 0], hacc1 + hx[r * 4 + 1], hacc2 + hx[r * 4 + 2], 0.f);
                    *reinterpret_cast<float4 *>(p.head_out + ob * p.hsb + oh * p.hsh + ow * p.hsw) = o;
                }
                asm volatile("bar.sync 2, 256;" ::: "memory");
            }
            tc_fence_before();
            acc_signal(tempty_sig[as]);
            if (++as == p.acc_stages) { as = 0; aphase ^= 1u; }
        }
    }

    // shared memory must outlive the bulk stores that read it: generic epilogue = one issuing thread, fast = one per epilogue warp
    if (p.tma_store && (kFast ? (warp >= 4 && lane == 0) : (threadIdx.x == 128))) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    tc_fence_before();
    if (kPair) cluster_sync_all(); else __syncthreads();   // pair: neither CTA may exit while the peer can still signal it
    if (warp == 2) {
        tc_fence_after();
        if (kPair)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// Host side: tensor maps + launch
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void *sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || sym == nullptr) return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(sym);
    return fn;
}

static int encode_map(CUtensorMap *tm, int dtype, int rank, const void *ptr, const uint64_t *dims,
                      const uint64_t *strides_elems, const uint32_t *box, bool swizzle64 = false) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled driver entry point unavailable");
        return HAVC_ERR_NO_DEVICE;
    }
    cuuint64_t gdim[5], gstr[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bx[i] = box[i];
        es[i] = 1;
        if (i > 0) gstr[i - 1] = strides_elems[i] * 2ull;
    }
    CUresult r = fn(tm, dtype == HAVC_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                    rank, const_cast<void *>(ptr), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    swizzle64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu %llu %llu %llu %llu)",
                  (int)r, rank, (unsigned long long)dims[0], (unsigned long long)dims[1],
                  (unsigned long long)dims[2], (unsigned long long)(rank > 3 ? dims[3] : 0),
                  (unsigned long long)(rank > 4 ? dims[4] : 0));
        return HAVC_ERR_CUDA;
    }
    return HAVC_OK;
}

// chunk32: the box of the fast epilogue's per-warp stores / residual loads (32 channels x a 32-pixel slab, 64-byte swizzle)
static int encode_act(CUtensorMap *tm, const havc_act_view &a, int dtype, int bw, int bh, int bb, bool chunk32 = false) {
    const int P = a.P > 0 ? a.P : 1;
    uint64_t dims[5] = {(uint64_t)a.C, (uint64_t)a.W, (uint64_t)a.H, (uint64_t)a.B, (uint64_t)P};
    // size-1 dimensions still need a legal (non-zero, 16 B multiple) stride
    uint64_t sw = a.stride_w, sh = a.stride_h, sb = a.stride_b, sp = a.stride_p;
    if (sh == 0) sh = sw * a.W;
    if (sb == 0) sb = sh * a.H;
    if (sp == 0) sp = sb * a.B;
    uint64_t strides[5] = {1, sw, sh, sb, sp};
    uint32_t box[5] = {(uint32_t)(chunk32 ? 32 : kChunkK), (uint32_t)bw, (uint32_t)bh, (uint32_t)bb, 1u};
    return encode_map(tm, dtype, 5, a.ptr, dims, strides, box, chunk32);
}

static bool act_ok(const havc_act_view &a) {
    return a.ptr != nullptr && a.C > 0 && a.C % 8 == 0 && a.W > 0 && a.H > 0 && a.B > 0 &&
           a.stride_w % 8 == 0 && a.stride_h % 8 == 0 && a.stride_b % 8 == 0 && a.stride_p % 8 == 0 &&
           (reinterpret_cast<uintptr_t>(a.ptr) & 15) == 0;
}

}  // namespace havc

using namespace havc;

extern "C" int havc_conv_gemm(const havc_conv_desc *d, void *stream) {
    HAVC_CHECK_ARG(d != nullptr, "havc_conv_gemm: null descriptor");
    HAVC_CHECK_ARG(d->dtype == HAVC_F16 || d->dtype == HAVC_BF16, "havc_conv_gemm: dtype must be f16/bf16");
    HAVC_CHECK_ARG(act_ok(d->src0), "havc_conv_gemm: src0 view invalid (C/strides must be multiples of 8, ptr 16 B aligned)");
    const bool two = d->src1.ptr != nullptr;
    if (two) HAVC_CHECK_ARG(act_ok(d->src1), "havc_conv_gemm: src1 view invalid");
    HAVC_CHECK_ARG(d->box_w > 0 && d->box_h > 0 && d->box_b > 0 && d->box_w * d->box_h * d->box_b == kTileM &&
                       d->box_w <= 256 && d->box_h <= 256 && d->box_b <= 256,
                   "havc_conv_gemm: box %dx%dx%d must multiply to 128", d->box_w, d->box_h, d->box_b);
    HAVC_CHECK_ARG(d->BN >= 16 && d->BN % 16 == 0 && d->BN <= kMaxBN, "havc_conv_gemm: BN=%d unsupported", d->BN);
    HAVC_CHECK_ARG(d->N_total > 0 && d->N_total % 16 == 0, "havc_conv_gemm: N_total=%d must be a multiple of 16", d->N_total);
    HAVC_CHECK_ARG(d->n_taps >= 1 && d->n_taps <= HAVC_MAX_TAPS, "havc_conv_gemm: n_taps=%d", d->n_taps);
    HAVC_CHECK_ARG(d->weight != nullptr && d->w_cin % 8 == 0 && d->w_rows > 0 && d->w_taps > 0 && d->w_batches > 0 &&
                       (reinterpret_cast<uintptr_t>(d->weight) & 15) == 0,
                   "havc_conv_gemm: weight tensor invalid");
    HAVC_CHECK_ARG((d->out != nullptr || d->head_w != nullptr) && d->c_store > 0 && d->c_store % 8 == 0, "havc_conv_gemm: out/c_store invalid");
    HAVC_CHECK_ARG(d->out_dtype == HAVC_F16 || d->out_dtype == HAVC_BF16 || d->out_dtype == HAVC_F32, "havc_conv_gemm: out_dtype");
    HAVC_CHECK_ARG(d->out_stride_w % 8 == 0 && d->out_stride_h % 8 == 0 && d->out_stride_b % 8 == 0 &&
                       (reinterpret_cast<uintptr_t>(d->out) & 15) == 0,
                   "havc_conv_gemm: output strides must be multiples of 8 elements");
    HAVC_CHECK_ARG(d->up >= 1, "havc_conv_gemm: up must be >= 1");
    if (d->shuffle && !d->blur) HAVC_CHECK_ARG(d->group_n > 0 && d->group_n % 16 == 0 && d->N_total == 4 * d->group_n,
                                               "havc_conv_gemm: shuffle needs N_total == 4*group_n, group_n %% 16 == 0");
    if (d->residual) HAVC_CHECK_ARG(d->res_stride_w % 8 == 0 && d->res_stride_h % 8 == 0 && d->res_stride_b % 8 == 0 &&
                                        (reinterpret_cast<uintptr_t>(d->residual) & 15) == 0 && !d->shuffle,
                                    "havc_conv_gemm: residual strides invalid");
    HAVC_CHECK_ARG((d->scale == nullptr) == (d->shift == nullptr), "havc_conv_gemm: scale and shift go together");

    ConvParams p;
    memset(&p, 0, sizeof(p));
    p.out_B = d->out_B; p.out_H = d->out_H; p.out_W = d->out_W;
    p.bw = d->box_w; p.bh = d->box_h; p.bb = d->box_b;
    p.blur = d->blur ? 1 : 0;
    p.halo = p.blur;
    if (p.blur) HAVC_CHECK_ARG(d->shuffle && d->box_b == 1 && d->box_w == 16 && d->box_h == 8 && d->BN == 256 && d->N_total <= kEpiSmemFloats &&
                                   d->N_total % d->BN == 0 && d->out_dtype == d->dtype && d->residual == nullptr && d->scale == nullptr &&
                                   d->relu1 && d->leaky1 == 0.f && d->split_n == 0 && d->head_w == nullptr && d->out_lo == nullptr && d->BN <= 256,
                               "havc_conv_gemm: blur needs a PixelShuffle launch (bias + ReLU only) with box_b = 1 and BN = 4 * channels per tile");
    p.tiles_w = ceil_div(d->out_W, d->box_w - p.halo);
    p.tiles_h = ceil_div(d->out_H, d->box_h - p.halo);
    p.tiles_b = ceil_div(d->out_B, d->box_b);
    p.BN = d->BN;
    p.tiles_n = ceil_div(d->N_total, d->BN);
    const int tiles_m = p.tiles_w * p.tiles_h * p.tiles_b;
    p.N_total = d->N_total;
    // BN > 256 is issued as two MMAs per K step.  Balanced halves (272 = 144 + 128) keep both compute-bound; a 256 + 16
    // split makes the narrow one shared-memory-bound (it re-reads the whole 128 x 16 A slice for 16 columns).
    static const bool legacy_split = getenv("HAVC_B200_SPLIT256") != nullptr;   // A/B switch for profiling
    p.n_part0 = d->BN > 256 ? (legacy_split ? 256 : ((d->BN / 2 + 15) / 16) * 16) : d->BN;
    p.n_part1 = d->BN - p.n_part0;
    // CTA pairs (cta_group::2): d->pair = 1 forces, -1 forbids, 0 = library default (env HAVC_B200_PAIR=0 turns it off, =1 pairs every launch)
    static const char *pair_env = getenv("HAVC_B200_PAIR");
    const bool pair_default = !(pair_env && pair_env[0] == '0');
    // a pair shares ONE weight tile: with per-batch weights (b_batched) both M tiles must lie in the same batch
    // Default policy (d->pair == 0), measured per layer shape on B200 (profiles/r01_pair_vs_single.txt): with the lean issue loop
    // pairs win 2-10 % on every 3x3 convolution (MMA-bound: the halved weight reads relieve shared memory) and on long 1x1
    // launches; short epilogue-bound 1x1 launches lose 3-5 % to the cross-SM barrier round trips.
    const bool pair_everywhere = pair_env && pair_env[0] == '1';      // A/B switch for profiling
    const bool pair_auto = pair_default && (pair_everywhere || d->n_taps > 1 || (long long)tiles_m * p.tiles_n >= 4096);
    const bool pair = (d->pair > 0 || (d->pair == 0 && pair_auto)) && tiles_m >= 2 && num_sms() >= 2 &&
                      (!d->b_batched || (p.tiles_w * p.tiles_h) % 2 == 0);
    p.total_tiles = (pair ? ceil_div(tiles_m, 2) : tiles_m) * p.tiles_n;
    if (pair) {      // each CTA loads half of the rows of each MMA's B operand
        p.n_wloads = p.n_part1 > 0 ? 2 : 1;
        p.wl_row0[0] = 0;          p.wl_rank_rows[0] = p.n_part0 / 2; p.wl_map[0] = 0; p.wl_smem[0] = 0;
        p.wl_row0[1] = p.n_part0;  p.wl_rank_rows[1] = p.n_part1 / 2; p.wl_map[1] = 1; p.wl_smem[1] = (uint32_t)(p.n_part0 / 2) * 128u;
        p.b1_smem_off = p.wl_smem[1];
        p.w_box_rows = p.n_part0 / 2;
    } else {
        p.n_wloads = d->BN > 256 ? 2 : 1;
        p.w_box_rows = d->BN / p.n_wloads;
        for (int l = 0; l < p.n_wloads; ++l) {
            p.wl_row0[l] = l * p.w_box_rows; p.wl_rank_rows[l] = 0; p.wl_map[l] = 0; p.wl_smem[l] = (uint32_t)(l * p.w_box_rows) * 128u;
        }
        p.b1_smem_off = (uint32_t)p.n_part0 * 128u;
    }
    HAVC_CHECK_ARG(p.w_box_rows % 8 == 0 && (p.n_part1 / 2) % 8 == 0, "havc_conv_gemm: weight box rows %d not a multiple of 8", p.w_box_rows);
    p.n_taps = d->n_taps;
    for (int i = 0; i < d->n_taps; ++i) {
        p.dh[i] = d->tap_dh[i]; p.dw[i] = d->tap_dw[i]; p.tp[i] = d->tap_p[i]; p.twi[i] = d->tap_wi[i];
        HAVC_CHECK_ARG(d->tap_wi[i] >= 0 && d->tap_wi[i] < d->w_taps, "havc_conv_gemm: tap_wi out of range");
    }
    p.chunks0 = ceil_div(d->src0.C, kChunkK);
    p.chunks1 = two ? ceil_div(d->src1.C, kChunkK) : 0;
    p.w_c1_off = d->w_c1_off;
    p.a_batched = d->a_batched; p.b_batched = d->b_batched;
    p.acc_stages = 2;
    p.staggered = (2 * d->BN > (int)kTmemCols) ? 1 : 0;
    p.acc_stride = p.staggered ? (int)kTmemCols - d->BN : d->BN;
    p.stage_bytes = kABytes + (pair ? d->BN / 2 : d->BN) * 128;     // per CTA
    const bool out16 = d->out_dtype == HAVC_F16 || d->out_dtype == HAVC_BF16;
    p.tma_store = (d->tma_store && !d->blur && out16 && d->out != nullptr && d->BN % 64 == 0 && d->N_total % 64 == 0 && d->split_n == 0 &&
                   d->head_w == nullptr && d->up == 1 && d->oy == 0 && d->ox == 0 && (!d->shuffle || d->group_n % 64 == 0))
                      ? 1 : 0;
    // split precision (hi + lo planes)
    p.a0_lo = d->src0_lo != nullptr; p.a1_lo = (two && d->src1_lo != nullptr); p.w_lo = d->weight_lo != nullptr;
    p.nsub0 = 1 + p.a0_lo + p.w_lo; p.nsub1 = 1 + p.a1_lo + p.w_lo;
    p.x3_out = d->out_lo != nullptr; p.res_lo = (d->residual != nullptr && d->residual_lo != nullptr);
    if (p.x3_out) HAVC_CHECK_ARG(p.tma_store && !d->shuffle && d->BN <= 128 && (reinterpret_cast<uintptr_t>(d->out_lo) & 15) == 0,
                                 "havc_conv_gemm: out_lo needs the TMA-store epilogue (16-bit NHWC output, BN %% 64 == 0, BN <= 128, no PixelShuffle)");
    for (const void *lp : {d->src0_lo, d->src1_lo, d->weight_lo, d->residual_lo})
        HAVC_CHECK_ARG((reinterpret_cast<uintptr_t>(lp) & 15) == 0, "havc_conv_gemm: lo planes must be 16-byte aligned");
    p.lo_stage_off = (uint32_t)(kTileM * d->BN * 2);
    p.stage_out_bytes = p.tma_store ? (uint32_t)(kTileM * d->BN * 2) * (p.x3_out ? 2u : 1u) : 0u;
    if (p.blur) {       // staging tile of the fused blur: 128 rows, padded pitch (BN*2 + 16 bytes) for conflict-free 16-byte accesses
        p.blur_row_bytes = (uint32_t)d->BN * 2u + 16u;
        p.stage_out_bytes = ((uint32_t)kTileM * p.blur_row_bytes + 1023u) & ~1023u;
    }
    static const char *hint_env = getenv("HAVC_B200_WAIT_HINT");            // ns; A/B switch for profiling
    p.wait_hint_ns = hint_env ? (uint32_t)atoi(hint_env) : 0u;

    int stages = (227 * 1024 - 1024 - kBarBytes - kEpiSmemFloats * (int)sizeof(float) - (int)p.stage_out_bytes) / (int)p.stage_bytes;
    if (stages > kMaxStages) stages = kMaxStages;
    HAVC_CHECK_ARG(stages >= 2, "havc_conv_gemm: tile too large for shared memory");
    p.num_stages = stages;
    const uint32_t fmt = d->dtype == HAVC_F16 ? 0u : 1u;
    auto idesc = [&](int n) -> uint32_t {
        return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)((pair ? 2 * kTileM : kTileM) >> 4) << 24);
    };
    p.idesc0 = idesc(p.n_part0);
    p.idesc1 = p.n_part1 > 0 ? idesc(p.n_part1) : 0u;
    p.bias = d->bias; p.scale = d->scale; p.shift = d->shift;
    p.relu1 = d->relu1; p.relu2 = d->relu2;
    p.slope1 = d->relu1 ? d->leaky1 : 1.0f;
    HAVC_CHECK_ARG(d->leaky1 >= 0.f && d->leaky1 < 1.f, "havc_conv_gemm: leaky1 must be in [0,1)");
    p.residual = d->residual; p.rsw = d->res_stride_w; p.rsh = d->res_stride_h; p.rsb = d->res_stride_b;
    p.out = d->out; p.out_dtype = d->out_dtype;
    p.osw = d->out_stride_w; p.osh = d->out_stride_h; p.osb = d->out_stride_b;
    p.up = d->up; p.oy = d->oy; p.ox = d->ox; p.shuffle = d->shuffle; p.group_n = d->group_n > 0 ? d->group_n : 16;
    p.c_store = d->c_store; p.dtype = d->dtype;
    p.src1_single_tap = d->src1_single_tap; p.src1_wi = d->src1_wi;
    p.split_n = d->split_n > 0 ? d->split_n : 0x7fffffff;
    p.out2 = d->out2; p.o2sw = d->out2_stride_w; p.o2sh = d->out2_stride_h; p.o2sb = d->out2_stride_b;
    p.c_store2 = d->c_store2;
    p.residual2 = d->residual2; p.r2sw = d->res2_stride_w; p.r2sh = d->res2_stride_h; p.r2sb = d->res2_stride_b;
    p.head_w = d->head_w; p.head_out = d->head_out;
    p.hsw = d->head_stride_w; p.hsh = d->head_stride_h; p.hsb = d->head_stride_b;
    if (d->src1_single_tap) HAVC_CHECK_ARG(two && d->src1_wi >= 0 && d->src1_wi < d->w_taps, "havc_conv_gemm: src1_single_tap needs src1 and a valid src1_wi");
    if (d->split_n > 0) HAVC_CHECK_ARG(d->split_n % 16 == 0 && !d->shuffle && (d->out2 != nullptr || d->head_w != nullptr || d->residual2 != nullptr) &&
                                           d->c_store2 % 8 == 0, "havc_conv_gemm: split_n must be a multiple of 16 with out2/residual2");
    if (d->head_w) HAVC_CHECK_ARG(d->head_out != nullptr && d->BN == d->N_total && !d->shuffle && d->head_stride_w % 4 == 0,
                                  "havc_conv_gemm: the fused head needs a single N tile (BN == N_total) and head_out");

    CUtensorMap tmA0, tmA1, tmW, tmW1;
    int rc = encode_act(&tmA0, d->src0, d->dtype, d->box_w, d->box_h, d->a_batched ? d->box_b : 1);
    if (rc) return rc;
    if (!d->a_batched) HAVC_CHECK_ARG(d->box_b == 1, "havc_conv_gemm: a_batched=0 needs box_b=1");
    if (two) {
        rc = encode_act(&tmA1, d->src1, d->dtype, d->box_w, d->box_h, d->a_batched ? d->box_b : 1);
        if (rc) return rc;
    } else {
        tmA1 = tmA0;
    }
    {
        uint64_t dims[4] = {(uint64_t)d->w_cin, (uint64_t)d->w_taps, (uint64_t)d->w_rows, (uint64_t)d->w_batches};
        uint64_t strides[4] = {1, (uint64_t)d->w_cin, (uint64_t)d->w_cin * d->w_taps,
                               (uint64_t)d->w_cin * d->w_taps * d->w_rows};
        uint32_t box[4] = {(uint32_t)kChunkK, 1u, (uint32_t)p.w_box_rows, 1u};
        rc = encode_map(&tmW, d->dtype, 4, d->weight, dims, strides, box);
        if (rc) return rc;
        tmW1 = tmW;
        if (pair && p.n_part1 > 0) {
            box[2] = (uint32_t)(p.n_part1 / 2);
            rc = encode_map(&tmW1, d->dtype, 4, d->weight, dims, strides, box);
            if (rc) return rc;
        }
    }
    if (d->b_batched) HAVC_CHECK_ARG(d->box_b == 1, "havc_conv_gemm: b_batched=1 needs box_b=1");
    LoMaps tmL;
    tmL.a0 = tmA0; tmL.a1 = tmA1; tmL.w = tmW; tmL.w1 = tmW1;
    if (p.a0_lo) {
        havc_act_view v = d->src0;
        v.ptr = d->src0_lo;
        rc = encode_act(&tmL.a0, v, d->dtype, d->box_w, d->box_h, d->a_batched ? d->box_b : 1);
        if (rc) return rc;
    }
    if (p.a1_lo) {
        havc_act_view v = d->src1;
        v.ptr = d->src1_lo;
        rc = encode_act(&tmL.a1, v, d->dtype, d->box_w, d->box_h, d->a_batched ? d->box_b : 1);
        if (rc) return rc;
    }
    if (p.w_lo) {
        uint64_t dims[4] = {(uint64_t)d->w_cin, (uint64_t)d->w_taps, (uint64_t)d->w_rows, (uint64_t)d->w_batches};
        uint64_t strides[4] = {1, (uint64_t)d->w_cin, (uint64_t)d->w_cin * d->w_taps,
                               (uint64_t)d->w_cin * d->w_taps * d->w_rows};
        uint32_t box[4] = {(uint32_t)kChunkK, 1u, (uint32_t)p.w_box_rows, 1u};
        rc = encode_map(&tmL.w, d->dtype, 4, d->weight_lo, dims, strides, box);
        if (rc) return rc;
        tmL.w1 = tmL.w;
        if (pair && p.n_part1 > 0) {
            box[2] = (uint32_t)(p.n_part1 / 2);
            rc = encode_map(&tmL.w1, d->dtype, 4, d->weight_lo, dims, strides, box);
            if (rc) return rc;
        }
    }

    const size_t smem = (size_t)p.num_stages * p.stage_bytes + p.stage_out_bytes + 1024 + kBarBytes + kEpiSmemFloats * sizeof(float);
    static const bool no_fast = getenv("HAVC_B200_NO_FAST_EPILOGUE") != nullptr;   // A/B switch for profiling
    static const bool no_res_tma = getenv("HAVC_B200_NO_RES_TMA") != nullptr;      // A/B switch: residual convs take the generic epilogue
    auto is_pow2 = [](int v) { return v > 0 && (v & (v - 1)) == 0; };
    const bool fast = !no_fast && p.tma_store && !p.staggered && d->head_w == nullptr && d->split_n == 0 && d->residual2 == nullptr &&
                      d->out_dtype == d->dtype && d->BN % 32 == 0 && d->N_total % d->BN == 0 &&
                      is_pow2(d->box_w) && is_pow2(d->box_h) && is_pow2(d->box_b) &&      // 32-pixel slabs must be sub-boxes
                      (d->residual == nullptr || !no_res_tma);                            // the fast epilogue takes its residual by TMA
    if (p.x3_out) HAVC_CHECK_ARG(fast, "havc_conv_gemm: out_lo is only implemented in the fast (TMA-store) epilogue");
    typedef void (*KernelFn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const LoMaps, const StoreMaps, const ConvParams);
    static const KernelFn kernels[2][2][2] = {
        {{conv_gemm_kernel<HAVC_F16, false, false>, conv_gemm_kernel<HAVC_F16, false, true>},
         {conv_gemm_kernel<HAVC_F16, true, false>, conv_gemm_kernel<HAVC_F16, true, true>}},
        {{conv_gemm_kernel<HAVC_BF16, false, false>, conv_gemm_kernel<HAVC_BF16, false, true>},
         {conv_gemm_kernel<HAVC_BF16, true, false>, conv_gemm_kernel<HAVC_BF16, true, true>}}};
    KernelFn kern = kernels[d->dtype == HAVC_F16 ? 0 : 1][fast ? 1 : 0][pair ? 1 : 0];
    {   // the opt-in to > 48 KB of dynamic shared memory is per device (one engine per GPU may live in this process)
        static std::atomic<unsigned long long> attr_done{0ull};
        int dev = 0;
        HAVC_CHECK_CUDA(cudaGetDevice(&dev));
        const unsigned long long bit = 1ull << (dev & 63);
        if (!(attr_done.load(std::memory_order_acquire) & bit)) {
            for (int i = 0; i < 8; ++i)
                HAVC_CHECK_CUDA(cudaFuncSetAttribute(reinterpret_cast<const void *>(kernels[i >> 2][(i >> 1) & 1][i & 1]),
                                                     cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            attr_done.fetch_or(bit, std::memory_order_release);
        }
    }
    StoreMaps tmO;
    memset(&tmO, 0, sizeof(tmO));
    // fast epilogue: every warp stores / loads its own 32-pixel slab of the box, 32 channels at a time
    const int slab_w = d->box_w < 32 ? d->box_w : 32;
    const int slab_h = d->box_h < 32 / slab_w ? d->box_h : 32 / slab_w;
    const int slab_b = 32 / (slab_w * slab_h);
    const int obw = fast ? slab_w : d->box_w, obh = fast ? slab_h : d->box_h, obb = fast ? slab_b : d->box_b;
    p.res_tma = (fast && d->residual != nullptr && !d->shuffle && !no_res_tma) ? 1 : 0;
    if (p.res_tma) {     // same geometry as the output map
        havc_act_view v;
        memset(&v, 0, sizeof(v));
        v.ptr = d->residual;
        v.C = d->c_store; v.W = d->out_W; v.H = d->out_H; v.B = d->out_B; v.P = 1;
        v.stride_w = d->res_stride_w; v.stride_h = d->res_stride_h; v.stride_b = d->res_stride_b;
        v.stride_p = d->res_stride_b * d->out_B;
        rc = encode_act(&tmO.m[1], v, d->dtype, obw, obh, obb, fast);
        if (rc) return rc;
    }
    if (p.res_tma && p.res_lo) {
        havc_act_view v;
        memset(&v, 0, sizeof(v));
        v.ptr = d->residual_lo;
        v.C = d->c_store; v.W = d->out_W; v.H = d->out_H; v.B = d->out_B; v.P = 1;
        v.stride_w = d->res_stride_w; v.stride_h = d->res_stride_h; v.stride_b = d->res_stride_b;
        v.stride_p = d->res_stride_b * d->out_B;
        rc = encode_act(&tmO.m[3], v, d->dtype, obw, obh, obb, fast);
        if (rc) return rc;
    } else {
        p.res_lo = 0;
    }
    if (p.x3_out) {
        havc_act_view v;
        memset(&v, 0, sizeof(v));
        v.ptr = d->out_lo;
        v.C = d->c_store; v.W = d->out_W; v.H = d->out_H; v.B = d->out_B; v.P = 1;
        v.stride_w = d->out_stride_w; v.stride_h = d->out_stride_h; v.stride_b = d->out_stride_b;
        v.stride_p = d->out_stride_b * d->out_B;
        rc = encode_act(&tmO.m[2], v, d->out_dtype, obw, obh, obb, fast);
        if (rc) return rc;
    }
    if (p.tma_store) {
        const int nmaps = d->shuffle ? 4 : 1;
        for (int g = 0; g < nmaps; ++g) {
            havc_act_view v;
            memset(&v, 0, sizeof(v));
            const long long off = d->shuffle ? ((g >> 1) * d->out_stride_h + (g & 1) * d->out_stride_w) : 0;
            v.ptr = reinterpret_cast<const uint16_t *>(d->out) + off;
            v.C = d->c_store; v.W = d->out_W; v.H = d->out_H; v.B = d->out_B; v.P = 1;
            const int m = d->shuffle ? 2 : 1;
            v.stride_w = m * d->out_stride_w; v.stride_h = m * d->out_stride_h; v.stride_b = d->out_stride_b;
            v.stride_p = d->out_stride_b * d->out_B;
            rc = encode_act(&tmO.m[g], v, d->out_dtype, obw, obh, obb, fast);
            if (rc) return rc;
        }
    }
    {
        static const bool no_pdl = getenv("HAVC_B200_NO_PDL") != nullptr;   // A/B switch for profiling
        int grid;
        if (pair) {
            const int max_groups = num_sms() / 2;
            grid = 2 * (p.total_tiles < max_groups ? p.total_tiles : max_groups);
        } else {
            grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
        }
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(fast ? kThreadsFast : kThreads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = (cudaStream_t)stream;
        cudaLaunchAttribute attrs[2];
        int na = 0;
        if (pair) {
            attrs[na].id = cudaLaunchAttributeClusterDimension;
            attrs[na].val.clusterDim.x = 2; attrs[na].val.clusterDim.y = 1; attrs[na].val.clusterDim.z = 1;
            ++na;
        }
        if (!no_pdl) {     // the kernel waits (griddepcontrol.wait) before it reads anything a predecessor wrote
            attrs[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attrs[na].val.programmaticStreamSerializationAllowed = 1;
            ++na;
        }
        cfg.attrs = attrs;
        cfg.numAttrs = na;
        HAVC_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA0, tmA1, tmW, tmW1, tmL, tmO, p));
    }
    HAVC_LAUNCHED();
    return HAVC_OK;
}
