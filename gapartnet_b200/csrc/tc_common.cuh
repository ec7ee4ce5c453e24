// tcgen05 / TMEM / mbarrier / cp.async helpers shared by the tensor-core kernels (conv_tc.cu,
// conv_wgrad_tc.cu).  Inline PTX only - no CUTLASS/CuTe.
#pragma once
#include "common.cuh"

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ int lds_i32(uint32_t addr) {
    int v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void lds_i32x4(uint32_t addr, int* v) {
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(addr));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(
                     smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
// Blocking wait.  A bare try_wait loop returns every few cycles: in the first tcgen05 conv kernel 58 % of
// all issued warp instructions were these polls (ncu source page, profiles/r1_summary.md), starving the
// producer warps that share the schedulers; ptxas drops try_wait's suspend-time hint (same SASS), so the
// back-off is an explicit nanosleep between polls.  ns ~ how long the role can afford to oversleep.
// Every blocking wait carries a watchdog: a protocol bug must surface as a trapped kernel with a message, never
// as a hung GPU (legitimate waits are micro-seconds; the limit is ~1 s of polling).
#define GP_MBAR_WATCHDOG_POLLS 40000000u
#ifdef GP_MBAR_TRAP_INLINE
// k_wgrad_win: the watchdog is a bare trap, NOT the printf helper: a (never taken) function call in a polling loop made
// ptxas keep loop counters of the feeder role in local memory across it, and with the 227 KB shared-memory carve-out there
// is no L1 left - every such LDL/STL is an L2 round trip (~1400 idle cycles per stage in the clock64 trace).  (The same
// change made k_wgrad_tc 45 % slower - its polling loops got laid out differently -, hence opt-in per file.)
__device__ __forceinline__ void mbar_timeout(uint32_t addr, uint32_t parity) {
    (void)addr;
    (void)parity;
    asm volatile("trap;");
}
#else
static __device__ __noinline__ void mbar_timeout(uint32_t addr, uint32_t parity) {
    printf("gapart_b200: mbarrier wait timed out (block %d thread %d smem 0x%x parity %u)\n", (int)blockIdx.x,
           (int)threadIdx.x, addr, parity);
    __trap();
}
#endif
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity, uint32_t ns) {
    uint32_t addr = smem_u32(bar);
    uint32_t ok = 0, polls = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
        if (ok) break;
        __nanosleep(ns);
        if (++polls > GP_MBAR_WATCHDOG_POLLS) mbar_timeout(addr, parity);
    }
}
__device__ __forceinline__ void mbar_wait_addr(uint32_t addr, uint32_t parity) {
    uint32_t ok = 0, polls = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
        if (ok) break;
        __nanosleep(20);
        if (++polls > GP_MBAR_WATCHDOG_POLLS) mbar_timeout(addr, parity);
    }
}
__device__ __forceinline__ void mbar_wait_addr_sleep(uint32_t addr, uint32_t parity, uint32_t ns) {
    uint32_t ok = 0, polls = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
        if (ok) break;
        __nanosleep(ns);
        if (++polls > GP_MBAR_WATCHDOG_POLLS) mbar_timeout(addr, parity);
    }
}
#ifndef GP_MBAR_BACKOFF_NS
#define GP_MBAR_BACKOFF_NS 32
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    mbar_wait_sleep(bar, parity, GP_MBAR_BACKOFF_NS);
}
// non-blocking probe: true if the phase with this parity has completed
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// mbarrier operations are per-thread shared-memory transactions: 128 threads polling one barrier
// cost ~1.4k cycles per chunk (measured).  Waits are therefore warp-uniform: lane 0 polls, the
// warp re-converges on __syncwarp (which also orders the other lanes' later accesses).
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity, int lane) {
    if (lane == 0) mbar_wait(bar, parity);
    __syncwarp();
}
// one elected lane of a converged warp (elect.sync): unlike `lane == 0` the compiler keeps the
// operands of the guarded tcgen05 instructions in uniform registers (no R2UR waterfall loop)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred px;\n\telect.sync _|px, 0xffffffff;\n\tselp.u32 %0, 1, 0, px;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint32_t uniform(uint32_t v) { return __shfl_sync(0xffffffffu, v, 0); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major operand tile, 128-byte swizzle, 8-row groups 1024 bytes apart (SBO), sm100 descriptor v1
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);      // start address
    d |= (uint64_t)0 << 16;                          // leading byte offset (unused: one atom along K)
    d |= (uint64_t)((1024 >> 4) & 0x3FFF) << 32;     // stride byte offset
    d |= (uint64_t)1 << 46;                          // descriptor version (sm100)
    d |= (uint64_t)2 << 61;                          // SWIZZLE_128B
    return d;
}
// byte offset of 16-byte piece j (0..7) of row r inside a 128B-swizzled K-major tile
__device__ __forceinline__ uint32_t swz128(int r, int j) {
    return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((j ^ (r & 7)) << 4));
}


__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                               uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
        "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
        "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]), "f"(v[9]),
        "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15]), "f"(v[16]), "f"(v[17]), "f"(v[18]),
        "f"(v[19]), "f"(v[20]), "f"(v[21]), "f"(v[22]), "f"(v[23]), "f"(v[24]), "f"(v[25]), "f"(v[26]), "f"(v[27]),
        "f"(v[28]), "f"(v[29]), "f"(v[30]), "f"(v[31])
        : "memory");
}


__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]), "f"(v[9]),
        "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15])
        : "memory");
}
