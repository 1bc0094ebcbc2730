// Shared device/host helpers for the cellregmap_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

namespace crm {

// ---- status codes of the C ABI (include/crm_b200.h) ----
enum : int {
    CRM_OK = 0,
    CRM_ERR_INVALID = -1,      // invalid argument
    CRM_ERR_UNSUPPORTED = -2,  // shape outside the compiled limits
    CRM_ERR_STATE = -3,        // handle not set up
    CRM_ERR_NONFINITE = -4,    // non-finite values in an input matrix
    CRM_ERR_CUDA = 1,          // CUDA runtime / driver error
    CRM_ERR_SOLVER = 2,        // cuSOLVER error or non-converged eigendecomposition
};

void set_error(const char* fmt, ...);
void count_launch();   // bumps the process-wide count of kernels launched by this library

#define CRM_CUDA(expr)                                                                         \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            crm::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return crm::CRM_ERR_CUDA;                                                          \
        }                                                                                      \
    } while (0)

#define CRM_CHECK(expr)                   \
    do {                                  \
        int _s = (expr);                  \
        if (_s != crm::CRM_OK) return _s; \
    } while (0)

// numpy_sugar.epsilon constants used by the reference path (oracle/sugar_port.py)
#define CRM_EPS_TINY 2.220446049250313e-16
#define CRM_EPS_SMALL 1.4901161193847656e-08
#define CRM_LOG2PI 1.8378770664093453
#define CRM_LOGMAX 709.782712893384

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier / TMA wrappers (PTX ISA: mbarrier.*, cp.async.bulk.tensor) ----
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t addr = smem_u32(bar), done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}
// polling wait with back-off, for the producer lane (keeps its spin loop off the issue ports the consumers use)
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
    uint32_t addr = smem_u32(bar), done;
    for (;;) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) break;
        __nanosleep(128);
    }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)map) : "memory");
}

// FP64 tensor-core MMA, the native sm_100a shape (SASS DMMA.8x8x4): D(8x8) += A(8x4) * B(4x8).
// lane = 4*g + t :  a = A[g][t],  b = B[t][g],  d = {D[g][2t], D[g][2t+1]}.
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

}  // namespace crm
