// Micro-benchmark (profiling aid): tcgen05.mma throughput, cta_group::1 (M=128) vs cta_group::2 (M=256), as a function
// of N; static (garbage) operands in shared memory, unrolled issue from one elected thread, one commit per 64 MMAs.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I music-fader-nets_b200/csrc -o tools/ubench_mma2.bin tools/ubench_mma2.cu
#include <cstdio>
#include <cstdlib>
#include "fn_tc.cuh"
void fn_set_error(const char*, ...) {}
int fn_num_sms() { return 148; }
int fn_max_smem_optin() { return 232448; }
fn_PFN_encodeTiled fn_get_encode_tiled() { return nullptr; }
int fn_make_tmap_bf16_2d(CUtensorMap*, const void*, unsigned long long, unsigned long long, unsigned long long, unsigned, unsigned) { return 0; }

template <int CL, int N>
__global__ void __launch_bounds__(128, 1) k(int reps, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t done;
    __shared__ uint32_t tslot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = CL == 2 ? tc::cluster_ctarank() : 0;
    for (int i = threadIdx.x; i < (200 * 1024) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { tc::mbar_init(&done, 1); tc::fence_barrier_init(); }
    if (warp == 1) { if (CL == 2) tc::tmem_alloc_2cta(&tslot, 512); else tc::tmem_alloc(&tslot, 512); }
    tc::fence_proxy_async();
    tc::tc_fence_before(); __syncthreads();
    if (CL == 2) tc::cluster_sync();
    tc::tc_fence_after();
    const uint32_t tmem = tslot;
    if (warp == 0 && rank == 0) {
        const uint32_t idesc = tc::make_idesc_bf16(128 * CL, N, 0, 0);
        const uint64_t ad0 = tc::make_sdesc(tc::smem_u32(smem), 16, 1024), bd0 = tc::make_sdesc(tc::smem_u32(smem) + 128 * 1024, 16, 1024);
        long long best = 1ll << 60;
        for (int r = 0; r < reps; ++r) {
            __syncwarp();
            const long long t0 = clock64();
            if (tc::elect_one()) {
#pragma unroll
                for (int c = 0; c < 8; ++c) {                // 8 A chunks of 16 KB, B chunk c & 1
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) {
                        const uint64_t ad = ad0 + (uint64_t)(c * 1024 + 2 * (kk & 3)), bd = bd0 + (uint64_t)((c & 1) * ((N / CL) * 128 >> 4) + 2 * (kk & 3));
                        if (CL == 2) tc::umma_f16_2cta(tmem, ad, bd, idesc, 1u); else tc::umma_f16(tmem, ad, bd, idesc, 1u);
                    }
                }
                if (CL == 2) tc::umma_commit_2cta_mc_u32(tc::smem_u32(&done), 1); else tc::umma_commit(&done);
            }
            __syncwarp();
            tc::mbar_wait(&done, r & 1);
            const long long t1 = clock64();
            if (t1 - t0 < best) best = t1 - t0;
        }
        if (lane == 0 && blockIdx.x == 0) out[0] = best;
    }
    tc::tc_fence_before(); __syncthreads();
    if (CL == 2) tc::cluster_sync();
    if (warp == 1) { tc::tc_fence_after(); if (CL == 2) tc::tmem_dealloc_2cta(tmem, 512); else tc::tmem_dealloc(tmem, 512); }
}

template <int CL, int N>
void run(int ctas) {
    long long* out; cudaMalloc(&out, 64); cudaMemset(out, 0, 64);
    const size_t smem = 1024 + 200 * 1024;
    cudaFuncSetAttribute(k<CL, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int reps = 20; void* args[] = {&reps, &out};
    cudaError_t e = cudaLaunchKernelExC(&cfg, (const void*)k<CL, N>, args);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CL=%d N=%d failed: %s\n", CL, N, cudaGetErrorString(e)); return; }
    long long h; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    printf("cta_group::%d M=%d N=%3d ctas=%3d: %6lld cycles / 64 MMAs = %.1f per MMA (ideal %d)\n", CL, 128 * CL, N, ctas, h, h / 64.0, N / 2);
}
int main() {
    run<1, 64>(128); run<1, 96>(128); run<1, 128>(128); run<1, 192>(128); run<1, 256>(128);
    run<2, 64>(128); run<2, 96>(128); run<2, 128>(128); run<2, 192>(128); run<2, 256>(128);
    run<2, 96>(2); run<2, 192>(2);
    return 0;
}
