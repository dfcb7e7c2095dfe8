// yb_api.cu — library-wide plumbing: thread-local error text, launch counter, device query.
#include <atomic>
#include <cstdarg>
#include <cstring>
#include "yb_common.cuh"

namespace yb {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

int sm_count() {
    static thread_local int cached_dev = -1, cached = 148;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) {
            cached = n;
            cached_dev = dev;
        }
    }
    return cached;
}

}  // namespace yb

extern "C" int yb_version(void) { return 100; }
extern "C" const char* yb_last_error(void) { return yb::g_err; }
extern "C" unsigned long long yb_launch_count(void) { return yb::g_launches.load(); }
