// capi.cu — error reporting, launch accounting and device queries of libhavc_b200.
#include "common.cuh"
#include <stdarg.h>
#include <string.h>

namespace havc {

static thread_local char g_err[1024] = "";
std::atomic<long long> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int num_sms() {
    // per device: engines for several GPUs may live in one process
    static std::atomic<int> cache[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    int n = cache[dev & 63].load(std::memory_order_relaxed);
    if (n == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cache[dev & 63].store(n, std::memory_order_relaxed);
    }
    return n;
}

}  // namespace havc

extern "C" const char *havc_last_error(void) { return havc::g_err; }
extern "C" int havc_version(void) { return 100; }
extern "C" int64_t havc_launch_count(void) { return havc::g_launches.load(); }
extern "C" void havc_launch_count_reset(void) { havc::g_launches.store(0); }
