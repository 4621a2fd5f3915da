// Micro-benchmark (profiling aid, not part of the product): streaming throughput of the CTA-pair pipeline of
// fn_gru_tc2.cu in isolation -- TMA loads of [128 x K] state slabs (cta_group::2, leader barrier) through a ring, consumed
// by tcgen05.mma.cta_group::2 (M = 256, N given) or by a bare release; no recurrence, no epilogue.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I music-fader-nets_b200/csrc -o /tmp/ubench_pair tools/ubench_pair.cu
// usage: ubench_pair pairs N kch S boxmode(0: one box per stage, 1: 8 KB boxes from 2*kch lanes) do_mma same_slab steps cluster(1|2)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <type_traits>
#include "fn_tc.cuh"
void fn_set_error(const char*, ...) {}
int fn_num_sms() { return 148; }
int fn_max_smem_optin() { return 232448; }
fn_PFN_encodeTiled fn_get_encode_tiled() { void* p = nullptr; cudaDriverEntryPointQueryResult q; cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q); return (fn_PFN_encodeTiled)p; }
int fn_make_tmap_bf16_2d(CUtensorMap*, const void*, unsigned long long, unsigned long long, unsigned long long, unsigned, unsigned) { return 0; }

constexpr int kATile = 16384;
struct Params { CUtensorMap tm; int H, kch, S, boxmode, do_mma, same_slab, steps, N, wbytes, cluster; long long* out; };

__global__ void __launch_bounds__(128, 1) k_stream(const __grid_constant__ Params P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t full[8], empty[8], done;
    __shared__ uint32_t tslot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = P.cluster == 2 ? tc::cluster_ctarank() : 0;
    const int pair = P.cluster == 2 ? blockIdx.x >> 1 : blockIdx.x;
    const uint32_t a_stage = P.kch * kATile;
    uint8_t* W = smem;                                   // static B operand (garbage): P.wbytes
    uint8_t* A = smem + P.wbytes;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
        tc::mbar_init(&done, 1);
        tc::fence_barrier_init();
    }
    if (warp == 1) { if (P.cluster == 2) tc::tmem_alloc_2cta(&tslot, 512); else tc::tmem_alloc(&tslot, 512); }
    tc::tc_fence_before(); __syncthreads();
    if (P.cluster == 2) tc::cluster_sync();
    tc::tc_fence_after();
    const uint32_t tmem = tslot;
    const int nkc = P.H / 64, nst = nkc / P.kch;
    const uint32_t mask = P.cluster == 2 ? tc::kPeerBitMask : 0xFFFFFFFFu;
    if (warp == 0) {
        const uint32_t full0 = tc::smem_u32(full), full_l = full0 & mask, empty0 = tc::smem_u32(empty), a0 = tc::smem_u32(A);
        uint32_t st = 0, ph = 1;
        const long long t0 = clock64();
        for (int i = 0; i < P.steps; ++i) {
            const int slab = P.same_slab ? i : i * 64 + (pair & 63);
            for (int j = 0; j < nst; ++j) {
                tc::mbar_wait_u32(empty0 + st * 8u, ph);
                if (rank == 0 && lane == 0) tc::mbar_arrive_expect_tx_u32(full0 + st * 8u, P.cluster * a_stage);
                __syncwarp();
                if (P.boxmode == 0) {
                    if (lane == 0) {
                        if (P.cluster == 2) tc::tma_load_4d_2cta_u32(a0 + st * a_stage, &P.tm, full_l + st * 8u, 0, (int)rank * 128, j * P.kch, slab);
                        else asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                                          ::"r"(a0 + st * a_stage), "l"(&P.tm), "r"(full0 + st * 8u), "r"(0), "r"(0), "r"(j * P.kch), "r"(slab) : "memory");
                    }
                } else if (lane < 2 * P.kch) {
                    const uint32_t dst = a0 + st * a_stage + (uint32_t)lane * (kATile / 2);
                    if (P.cluster == 2) tc::tma_load_4d_2cta_u32(dst, &P.tm, full_l + st * 8u, 0, (int)rank * 128 + (lane & 1) * 64, j * P.kch + (lane >> 1), slab);
                    else asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                                      ::"r"(dst), "l"(&P.tm), "r"(full0 + st * 8u), "r"(0), "r"((lane & 1) * 64), "r"(j * P.kch + (lane >> 1)), "r"(slab) : "memory");
                }
                __syncwarp();
                if (++st == (uint32_t)P.S) { st = 0; ph ^= 1u; }
            }
        }
        if (lane == 0 && blockIdx.x == 0) P.out[1] = clock64() - t0;
    } else if (warp == 1 && rank == 0) {
        const uint32_t idesc = tc::make_idesc_bf16(P.cluster == 2 ? 256 : 128, P.N, 0, 0);
        const uint32_t full0 = tc::smem_u32(full), empty0 = tc::smem_u32(empty);
        const uint64_t adesc0 = tc::make_sdesc(tc::smem_u32(A), 16, 1024), bdesc0 = tc::make_sdesc(tc::smem_u32(W), 16, 1024);
        const uint32_t wch = (uint32_t)(P.N / P.cluster) * 128;
        uint32_t st = 0, ph = 0;
        const long long t0 = clock64();
        for (int i = 0; i < P.steps; ++i) {
            for (int j = 0; j < nst; ++j) {
                tc::mbar_wait_u32(full0 + st * 8u, ph);
                tc::tc_fence_after();
                if (tc::elect_one()) {
                    if (P.do_mma) {
                        const uint64_t ad0 = adesc0 + (uint64_t)(st * (a_stage >> 4)), bd0 = bdesc0 + (uint64_t)((j & 1) * (wch >> 4));
                        auto issue = [&](auto kch_c) {
                            constexpr int KCH = decltype(kch_c)::value;
#pragma unroll
                            for (int q = 0; q < KCH; ++q)
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    tc::umma_f16_2cta(tmem, ad0 + (uint64_t)(q * (kATile >> 4) + 2 * k), bd0 + (uint64_t)(2 * k), idesc, 1u);
                        };
                        if (P.kch == 1) issue(std::integral_constant<int, 1>{});
                        else if (P.kch == 2) issue(std::integral_constant<int, 2>{});
                        else issue(std::integral_constant<int, 4>{});
                    }
                    if (P.cluster == 2) tc::umma_commit_2cta_mc_u32(empty0 + st * 8u, 3); else tc::umma_commit_u32(empty0 + st * 8u);
                }
                __syncwarp();
                if (++st == (uint32_t)P.S) { st = 0; ph ^= 1u; }
            }
        }
        if (tc::elect_one()) { if (P.cluster == 2) tc::umma_commit_2cta_mc_u32(tc::smem_u32(&done), 1); else tc::umma_commit(&done); }
        __syncwarp();
        tc::mbar_wait(&done, 0);
        if (lane == 0 && blockIdx.x == 0) P.out[0] = clock64() - t0;
    }
    tc::tc_fence_before(); __syncthreads();
    if (P.cluster == 2) tc::cluster_sync();
    if (warp == 1) { tc::tc_fence_after(); if (P.cluster == 2) tc::tmem_dealloc_2cta(tmem, 512); else tc::tmem_dealloc(tmem, 512); }
}

int main(int argc, char** argv) {
    if (argc < 10) { printf("usage: pairs N kch S boxmode do_mma same_slab steps cluster\n"); return 1; }
    Params P{};
    const int pairs = atoi(argv[1]);
    P.N = atoi(argv[2]); P.kch = atoi(argv[3]); P.S = atoi(argv[4]); P.boxmode = atoi(argv[5]); P.do_mma = atoi(argv[6]);
    P.same_slab = atoi(argv[7]); P.steps = atoi(argv[8]); P.cluster = atoi(argv[9]);
    P.H = 1024;
    const int B = 256, nslabs = P.same_slab ? P.steps : P.steps * 64;
    void* buf; cudaMalloc(&buf, (size_t)nslabs * B * P.H * 2); cudaMemset(buf, 0, (size_t)nslabs * B * P.H * 2);
    cuuint64_t dims[4] = {64, (cuuint64_t)B, (cuuint64_t)P.H / 64, (cuuint64_t)nslabs};
    cuuint64_t str[3] = {(cuuint64_t)P.H * 2, 128, (cuuint64_t)B * P.H * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)(P.boxmode ? 64 : 128), (cuuint32_t)(P.boxmode ? 1 : P.kch), 1}, estr[4] = {1, 1, 1, 1};
    CUresult r = fn_get_encode_tiled()(&P.tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, buf, dims, str, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    P.wbytes = 2 * (P.N / P.cluster) * 128;
    cudaMalloc(&P.out, 64); cudaMemset(P.out, 0, 64);
    const size_t smem = 1024 + P.wbytes + (size_t)P.S * P.kch * kATile;
    cudaFuncSetAttribute(k_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(pairs * P.cluster); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = P.cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    void* args[] = {&P};
    for (int rep = 0; rep < 2; ++rep) {
        cudaError_t e = cudaLaunchKernelExC(&cfg, (const void*)k_stream, args);
        if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); return 1; }
        e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
    }
    long long h[2]; cudaMemcpy(h, P.out, 16, cudaMemcpyDeviceToHost);
    const double bytes = (double)P.steps * 128 * P.H * 2;
    printf("pairs=%d N=%d kch=%d S=%d box=%d mma=%d same=%d cl=%d: consumer %lld cyc (%.1f B/clk/CTA, %.0f cyc/step)  loader %lld cyc\n", pairs, P.N, P.kch, P.S,
           P.boxmode, P.do_mma, P.same_slab, P.cluster, h[0], bytes / h[0], (double)h[0] / P.steps, h[1]);
    return 0;
}
