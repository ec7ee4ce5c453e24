// Shared device/host helpers for libgapart_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define GP_OK 0
#define GP_ERR_INVALID (-1)
#define GP_ERR_CUDA (-2)
#define GP_ERR_UNSUPPORTED (-3)
#define GP_ERR_OVERFLOW (-4)

// thread-local error text (api.cu)
void gp_set_error(const char* fmt, ...);

#define GP_CHECK_ARG(cond, ...)                 \
    do {                                        \
        if (!(cond)) {                          \
            gp_set_error(__VA_ARGS__);          \
            return GP_ERR_INVALID;              \
        }                                       \
    } while (0)

#define GP_CUDA(expr)                                                                 \
    do {                                                                              \
        cudaError_t _e = (expr);                                                      \
        if (_e != cudaSuccess) {                                                      \
            gp_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr,                 \
                         cudaGetErrorString(_e));                                     \
            return GP_ERR_CUDA;                                                       \
        }                                                                             \
    } while (0)

#define GP_LAUNCH_CHECK()                                                             \
    do {                                                                              \
        cudaError_t _e = cudaPeekAtLastError();                                       \
        if (_e != cudaSuccess) {                                                      \
            cudaGetLastError();                                                       \
            gp_set_error("%s:%d kernel launch -> %s", __FILE__, __LINE__,             \
                         cudaGetErrorString(_e));                                     \
            return GP_ERR_CUDA;                                                       \
        }                                                                             \
    } while (0)

static inline int gp_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// kernel-launch accounting (bench.py reports gpu_launches from it)
void gp_note_launch(int n);

// number of SMs of the current device (cached); B200 = 148
int gp_num_sms();

// Row counts live on the device so that no host sync is needed between
// voxelize -> rulebook -> conv.  A NULL pointer means "use the host bound".
__device__ __forceinline__ int gp_rows(const int* __restrict__ d_n, int bound) {
    if (d_n == nullptr) return bound;
    int n = *d_n;
    return n < bound ? n : bound;
}

__device__ __forceinline__ float4 ldg4(const float* p) {
    return __ldg(reinterpret_cast<const float4*>(p));
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// inclusive warp scan
__device__ __forceinline__ int warp_scan_incl(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// ---- programmatic dependent launch (PDL) ----------------------------------------------------------
// The step is a chain of ~500 short dependent kernels; at the deep U-Net levels a kernel runs 8-15 us and the
// launch gap between dependent graph nodes is a visible fraction of that.  Kernels launched through gp_launch()
// carry cudaLaunchAttributeProgrammaticStreamSerialization: the next grid may be scheduled while the current one
// drains (its CTAs take an SM as soon as one frees up) and runs its prologue; gp_pdl_wait() - executed by every such
// kernel before it touches global memory - blocks until the predecessor grid has completed and flushed, so the
// data dependencies (RAW and WAR) are those of ordinary stream order.  GAPART_PDL=0 launches without the attribute
// (the two instructions are then no-ops).
__device__ __forceinline__ void gp_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void gp_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool gp_pdl_enabled();

template <typename... KArgs, typename... Args>
static inline cudaError_t gp_launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                    Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = gp_pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}
