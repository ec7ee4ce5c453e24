// Library-wide plumbing of libgapart_b200.so: error text, device queries.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"
#include "../../include/gapart_b200.h"

static thread_local char g_err[512] = "";

void gp_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* gp_last_error(void) { return g_err; }
extern "C" int gp_version(void) { return GP_ABI_VERSION; }

int gp_num_sms() {
    static thread_local int cached_dev = -1, cached = 148;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return cached;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            cached = n;
        cached_dev = dev;
    }
    return cached;
}

extern "C" int gp_device_sms(void) { return gp_num_sms(); }

#include <stdlib.h>
bool gp_pdl_enabled() {
    static int cached = -1;
    if (cached < 0) {
        const char* e = getenv("GAPART_PDL");
        cached = (e && e[0] == '0') ? 0 : 1;
    }
    return cached == 1;
}

#include <atomic>
static std::atomic<long long> g_launches{0};
void gp_note_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
extern "C" long long gp_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

// fill an int32 / fp32 buffer (stream ordered); avoids a torch op inside graph-captured plans
__global__ void k_fill_i32(int* p, long long n, int v) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        p[i] = v;
}
extern "C" int gp_fill_i32(int* p, long long n, int v, void* stream_) {
    if (n <= 0) return GP_OK;
    long long b = (n + 255) / 256;
    if (b > 4096) b = 4096;
    k_fill_i32<<<(int)b, 256, 0, (cudaStream_t)stream_>>>(p, n, v);
    gp_note_launch(1);
    GP_LAUNCH_CHECK();
    return GP_OK;
}
extern "C" int gp_memset(void* p, int byte, long long nbytes, void* stream_) {
    if (nbytes <= 0) return GP_OK;
    GP_CUDA(cudaMemsetAsync(p, byte, (size_t)nbytes, (cudaStream_t)stream_));
    return GP_OK;
}
