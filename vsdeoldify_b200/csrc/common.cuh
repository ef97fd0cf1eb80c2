// Shared host/device helpers for libhavc_b200 (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include "../../include/havc_b200.h"

namespace havc {

void set_error(const char *fmt, ...);
extern std::atomic<long long> g_launches;

#define HAVC_CHECK_ARG(cond, ...)                                                                  \
    do {                                                                                           \
        if (!(cond)) {                                                                             \
            havc::set_error(__VA_ARGS__);                                                          \
            return HAVC_ERR_ARG;                                                                   \
        }                                                                                          \
    } while (0)

#define HAVC_CHECK_CUDA(expr)                                                                      \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            havc::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,      \
                            __LINE__);                                                             \
            return HAVC_ERR_CUDA;                                                                  \
        }                                                                                          \
    } while (0)

// Called after every kernel launch: counts it and surfaces launch-configuration errors.
#define HAVC_LAUNCHED()                                                                            \
    do {                                                                                           \
        havc::g_launches.fetch_add(1, std::memory_order_relaxed);                                  \
        HAVC_CHECK_CUDA(cudaGetLastError());                                                       \
    } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
int num_sms();
// Per-DEVICE one-time setup (cudaFuncSetAttribute opt-ins are per device, and one process may drive several GPUs):
// true while the calling thread's current device has not been marked in `mask`; mark it with device_done() afterwards.
static inline bool device_pending(std::atomic<unsigned long long> &mask, unsigned long long *bit) {
    int dev = 0;
    cudaGetDevice(&dev);
    *bit = 1ull << (dev & 63);
    return (mask.load(std::memory_order_acquire) & *bit) == 0;
}
static inline void device_done(std::atomic<unsigned long long> &mask, unsigned long long bit) {
    mask.fetch_or(bit, std::memory_order_release);
}

// ---- device-side numeric helpers -------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t pack2(float a, float b, int dtype) {
    if (dtype == HAVC_F16) {
        // saturating conversion (F2FP.SATFINITE, one instruction): a value beyond fp16's 65504 is stored as +-65504 instead of
        // inf, which the next layer would turn into NaN (inf - inf, 0 * inf) for the whole receptive field
        uint32_t r;
        asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
        return r;
    } else {
        __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
        return *reinterpret_cast<uint32_t *>(&h);
    }
}
__device__ __forceinline__ float2 unpack2(uint32_t v, int dtype) {
    if (dtype == HAVC_F16) {
        return __half22float2(*reinterpret_cast<__half2 *>(&v));
    } else {
        return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162 *>(&v));
    }
}
// packed 16-bit arithmetic on two lanes (round to nearest even in the storage type): the ICNR blur's sums
__device__ __forceinline__ uint32_t add2(uint32_t a, uint32_t b, int dtype) {
    uint32_t r;
    if (dtype == HAVC_F16) asm("add.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    else asm("add.rn.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
__device__ __forceinline__ uint32_t quarter2(uint32_t a, int dtype) {      // a * 0.25 (exact unless the result is subnormal)
    uint32_t r;
    if (dtype == HAVC_F16) asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(0x34003400u));
    else asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(0x3E803E80u));
    return r;
}
__device__ __forceinline__ float load16(const void *p, int64_t idx, int dtype) {
    if (dtype == HAVC_F16) return __half2float(reinterpret_cast<const __half *>(p)[idx]);
    return __bfloat162float(reinterpret_cast<const __nv_bfloat16 *>(p)[idx]);
}
__device__ __forceinline__ void store16(void *p, int64_t idx, float v, int dtype) {
    if (dtype == HAVC_F16)
        reinterpret_cast<__half *>(p)[idx] = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
    else
        reinterpret_cast<__nv_bfloat16 *>(p)[idx] = __float2bfloat16_rn(v);
}
#endif

}  // namespace havc
