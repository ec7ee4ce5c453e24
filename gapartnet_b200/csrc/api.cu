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

// ---- optimizer -------------------------------------------------------------------------------------------------------
// torch.optim.Adam (no amsgrad / weight decay: GAPartNet.configure_optimizers, gapartnet/network/model.py:1051-1055) over
// flat parameter / gradient arenas in ONE launch.  The step count lives on the device (the launch can sit in a CUDA graph).
__global__ void __launch_bounds__(256) k_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                              float* __restrict__ v, long long n, float lr, float b1, float b2, float eps,
                                              float grad_scale, const int* __restrict__ d_step) {
    const int t = *d_step;
    const double bc1 = 1.0 - pow((double)b1, (double)t), bc2 = 1.0 - pow((double)b2, (double)t);
    const float step_size = (float)((double)lr / bc1), inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
    for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += (long long)gridDim.x * blockDim.x * 4) {
        if (i + 4 <= n) {
            float4 pv = *reinterpret_cast<float4*>(p + i), mv = *reinterpret_cast<float4*>(m + i),
                   vv = *reinterpret_cast<float4*>(v + i);
            const float4 gv = *reinterpret_cast<const float4*>(g + i);
            float* pp = &pv.x; float* mm = &mv.x; float* vq = &vv.x; const float* gg = &gv.x;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float gr = gg[e] * grad_scale;
                mm[e] = b1 * mm[e] + (1.f - b1) * gr;
                vq[e] = b2 * vq[e] + (1.f - b2) * gr * gr;
                pp[e] -= step_size * mm[e] / (sqrtf(vq[e]) * inv_sqrt_bc2 + eps);
            }
            *reinterpret_cast<float4*>(p + i) = pv;
            *reinterpret_cast<float4*>(m + i) = mv;
            *reinterpret_cast<float4*>(v + i) = vv;
        } else {
            for (long long j = i; j < n; ++j) {
                const float gr = g[j] * grad_scale;
                m[j] = b1 * m[j] + (1.f - b1) * gr;
                v[j] = b2 * v[j] + (1.f - b2) * gr * gr;
                p[j] -= step_size * m[j] / (sqrtf(v[j]) * inv_sqrt_bc2 + eps);
            }
        }
    }
}
extern "C" int gp_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, float lr,
                            float beta1, float beta2, float eps, float grad_scale, const int* d_step, void* stream_) {
    GP_CHECK_ARG(n >= 0 && d_step != nullptr, "gp_adam_step: bad arguments");
    GP_CHECK_ARG(((reinterpret_cast<size_t>(param) | reinterpret_cast<size_t>(grad) | reinterpret_cast<size_t>(exp_avg) |
                   reinterpret_cast<size_t>(exp_avg_sq)) & 15) == 0, "gp_adam_step: arenas must be 16-byte aligned");
    if (n == 0) return GP_OK;
    long long blocks = (n / 4 + 255) / 256;
    const long long cap = (long long)gp_num_sms() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    k_adam<<<(int)blocks, 256, 0, (cudaStream_t)stream_>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                                                           grad_scale, d_step);
    gp_note_launch(1);
    GP_LAUNCH_CHECK();
    return GP_OK;
}
