// Micro-benchmark (profiling aid): LATENCY of one TMA load (issue -> mbarrier completion), L2-resident data, one CTA (or one
// CTA pair) on an otherwise idle GPU, as a function of the box size and of the instruction form.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I music-fader-nets_b200/csrc -o tools/ubench_tmalat.bin tools/ubench_tmalat.cu
#include <cstdio>
#include <cstdlib>
#include "fn_tc.cuh"
void fn_set_error(const char*, ...) {}
int fn_num_sms() { return 148; }
int fn_max_smem_optin() { return 232448; }
fn_PFN_encodeTiled fn_get_encode_tiled() { void* p = nullptr; cudaDriverEntryPointQueryResult q; cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q); return (fn_PFN_encodeTiled)p; }
int fn_make_tmap_bf16_2d(CUtensorMap*, const void*, unsigned long long, unsigned long long, unsigned long long, unsigned, unsigned) { return 0; }

// mode 0: 4-D tensor box, plain; 1: 4-D tensor box, cta_group::2 (leader barrier); 2: linear cp.async.bulk of `bytes`; 3: LDG.128 loop by one warp (bytes)
__global__ void __launch_bounds__(64, 1) k(const __grid_constant__ CUtensorMap tm, const uint8_t* base, int mode, int bytes, int iters, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    const uint32_t rank = tc::cluster_ctarank();
    if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
    __syncthreads(); tc::cluster_sync();
    if (threadIdx.x < 32 && rank == 0) {
        long long best = 1ll << 60, sum = 0;
        for (int it = 0; it < iters; ++it) {
            __syncwarp();
            const long long t0 = clock64();
            if (mode == 3) {
                uint4 acc = make_uint4(0, 0, 0, 0);
                for (int o = threadIdx.x * 16; o < bytes; o += 512) {
                    const uint4 v = __ldcg(reinterpret_cast<const uint4*>(base + (size_t)(it & 15) * 65536 + o));
                    acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
                }
                if (acc.x == 0x12345678) out[3] = 1;
            } else {
                if (threadIdx.x == 0) {
                    tc::mbar_arrive_expect_tx(&bar, (uint32_t)bytes);
                    if (mode == 0)
                        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                                     ::"r"(tc::smem_u32(smem)), "l"(&tm), "r"(tc::smem_u32(&bar)), "r"(0), "r"(0), "r"(0), "r"(it & 15) : "memory");
                    else if (mode == 1)
                        tc::tma_load_4d_2cta_u32(tc::smem_u32(smem), &tm, tc::smem_u32(&bar) & tc::kPeerBitMask, 0, 0, 0, it & 15);
                    else
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                     ::"r"(tc::smem_u32(smem)), "l"(base + (size_t)(it & 15) * 65536), "r"(bytes), "r"(tc::smem_u32(&bar)) : "memory");
                }
                tc::mbar_wait(&bar, it & 1);
            }
            const long long t1 = clock64();
            if (it >= 16) { sum += t1 - t0; if (t1 - t0 < best) best = t1 - t0; }
        }
        if (threadIdx.x == 0) { out[0] = best; out[1] = sum / (iters - 16); }
    }
    __syncthreads(); tc::cluster_sync();
}

int main() {
    const int H = 1024, B = 128, slabs = 16;
    uint8_t* buf; cudaMalloc(&buf, (size_t)slabs * 65536 * 4); cudaMemset(buf, 1, (size_t)slabs * 65536 * 4);
    long long* out; cudaMalloc(&out, 64);
    for (int mode = 0; mode < 4; ++mode)
        for (int kb : {4, 16, 32, 64}) {
            const int rows = kb >= 16 ? 128 : 32, kch = kb >= 16 ? kb / 16 : 1, bytes = rows * kch * 128;
            CUtensorMap tm;
            cuuint64_t dims[4] = {64, (cuuint64_t)B, (cuuint64_t)H / 64, (cuuint64_t)slabs};
            cuuint64_t str[3] = {(cuuint64_t)H * 2, 128, (cuuint64_t)B * H * 2};
            cuuint32_t box[4] = {64, (cuuint32_t)rows, (cuuint32_t)kch, 1}, estr[4] = {1, 1, 1, 1};
            fn_get_encode_tiled()(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, buf, dims, str, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
            cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(2); cfg.blockDim = dim3(64); cfg.dynamicSmemBytes = 70000;
            cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            int iters = 200; int b2 = bytes; const uint8_t* bp = buf;
            void* args[] = {&tm, &bp, &mode, &b2, &iters, &out};
            cudaError_t e = cudaLaunchKernelExC(&cfg, (const void*)k, args);
            if (e == cudaSuccess) e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("mode %d kb %d failed: %s\n", mode, kb, cudaGetErrorString(e)); return 1; }
            long long h[2]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
            printf("mode %d (%s) %2d KB: best %5lld cycles, mean %5lld\n", mode, mode == 0 ? "tensor 4-D" : mode == 1 ? "tensor 4-D cta_group::2" : mode == 2 ? "linear bulk" : "LDG.128 warp", bytes / 1024, h[0], h[1]);
        }
    return 0;
}
