// Shared device/host helpers for the fadernets_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

#include "../../include/fadernets_b200.h"

// ---- error state (fn_last_error) -------------------------------------------
void fn_set_error(const char* fmt, ...);

#define FN_CHECK_CUDA(expr)                                                              \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            fn_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return FN_ERR_CUDA;                                                          \
        }                                                                                \
    } while (0)

#define FN_REQUIRE(cond, ...)                                                            \
    do {                                                                                 \
        if (!(cond)) {                                                                   \
            fn_set_error(__VA_ARGS__);                                                   \
            return FN_ERR_ARG;                                                           \
        }                                                                                \
    } while (0)

#define FN_LAUNCH_CHECK() FN_CHECK_CUDA(cudaGetLastError())

static inline int fn_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

int fn_num_sms();          // cached cudaDevAttrMultiProcessorCount of the current device
int fn_max_smem_optin();   // cached cudaDevAttrMaxSharedMemoryPerBlockOptin

// ---- device helpers -----------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ float fn_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float fn_warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double fn_warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum; result valid in every thread.  `red` = >= 33 floats of smem.
__device__ __forceinline__ float fn_block_sum(float v, float* red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = fn_warp_sum(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    if (w == 0) {
        float t = lane < nw ? red[lane] : 0.f;
        t = fn_warp_sum(t);
        if (lane == 0) red[32] = t;
    }
    __syncthreads();
    return red[32];
}
__device__ __forceinline__ double fn_block_sum_d(double v, double* red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = fn_warp_sum_d(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    if (w == 0) {
        double t = lane < nw ? red[lane] : 0.0;
        t = fn_warp_sum_d(t);
        if (lane == 0) red[32] = t;
    }
    __syncthreads();
    return red[32];
}
__device__ __forceinline__ float fn_block_max(float v, float* red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = fn_warp_max(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    if (w == 0) {
        float t = lane < nw ? red[lane] : -INFINITY;
        t = fn_warp_max(t);
        if (lane == 0) red[32] = t;
    }
    __syncthreads();
    return red[32];
}

__device__ __forceinline__ float fn_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }

// acquire load / release-ish increment used by the per-chain step barriers
__device__ __forceinline__ unsigned fn_ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fn_red_release(unsigned* p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Wait until *ctr >= target.  Bounded: a lost arrival traps instead of hanging the GPU (the report is out of
// line so the inlined wait stays a few instructions long).
static __device__ __noinline__ void fn_spin_timeout(const unsigned* ctr, unsigned target) {
    printf("fadernets_b200: step barrier timeout (block %d target %u have %u)\n", blockIdx.x, target, fn_ld_acquire(ctr));
    __trap();
}
__device__ __forceinline__ void fn_spin_until(const unsigned* ctr, unsigned target) {
    unsigned spins = 0;
    while ((int)(fn_ld_acquire(ctr) - target) < 0) {
        if (++spins > (1u << 24)) fn_spin_timeout(ctr, target);
    }
}

__device__ __forceinline__ void fn_cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void fn_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void fn_cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

#endif  // __CUDACC__
