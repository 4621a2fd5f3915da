// bf16 tensor-core GEMM on tcgen05 (fn_tc_gemm_bf16): TMA -> 128B-swizzled smem ring -> tcgen05.mma with
// the fp32 accumulator in TMEM -> tcgen05.ld epilogue.  Warp-specialised: warp 0 = TMA producer,
// warp 1 = TMEM allocator + single-thread MMA issuer, warps 2..5 = epilogue (one TMEM lane quarter each).
// One 128x128 output tile per CTA; 3 smem stages (96 KB) so two CTAs share an SM and one tile's epilogue
// overlaps the other's main loop.  Both operands may be K-major or MN-major (transposed in memory), which
// covers forward (x W^T), dgrad (dy W) and wgrad (dy^T x) without materialising any transpose.
#include <mutex>

#include "fn_tc.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 64;
constexpr int kStages = 3;
constexpr int kTileBytes = 128 * BK * 2;                    // 16 KB per operand tile
constexpr int kStageBytes = 2 * kTileBytes;
constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr int kThreads = 192;
constexpr uint32_t kTmemCols = 128;

struct GemmParams {
    void* C;
    const float* bias;
    long long ldc;
    int M, N, K;
    int a_mn, b_mn, c_bf16, accumulate;
};

__global__ void __launch_bounds__(kThreads, 2)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
    uint64_t* empty = full + kStages;
    uint64_t* acc_full = empty + kStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int nkb = (p.K + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        tc::prefetch_tmap(&tmA);
        tc::prefetch_tmap(&tmB);
        for (int s = 0; s < kStages; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
        tc::mbar_init(acc_full, 1);
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc(tmem_slot, kTmemCols);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % kStages;
                const uint32_t ph = (kb / kStages) & 1;
                tc::mbar_wait(&empty[s], ph ^ 1);                      // slot free (passes immediately on first lap)
                uint8_t* sa = smem + s * kStageBytes;
                uint8_t* sb = sa + kTileBytes;
                tc::mbar_arrive_expect_tx(&full[s], kStageBytes);
                const int k0 = kb * BK;
                if (!p.a_mn) {
                    tc::tma_load_2d(sa, &tmA, &full[s], k0, m0);                       // box 64(K) x 128(M)
                } else {
                    tc::tma_load_2d(sa, &tmA, &full[s], m0, k0);                       // box 64(M) x 64(K)
                    tc::tma_load_2d(sa + kTileBytes / 2, &tmA, &full[s], m0 + 64, k0);
                }
                if (!p.b_mn) {
                    tc::tma_load_2d(sb, &tmB, &full[s], k0, n0);
                } else {
                    tc::tma_load_2d(sb, &tmB, &full[s], n0, k0);
                    tc::tma_load_2d(sb + kTileBytes / 2, &tmB, &full[s], n0 + 64, k0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = tc::make_idesc_bf16(BM, BN, p.a_mn, p.b_mn);
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % kStages;
                const uint32_t ph = (kb / kStages) & 1;
                tc::mbar_wait(&full[s], ph);
                tc::tc_fence_after();
                const uint32_t sa = tc::smem_u32(smem + s * kStageBytes), sb = sa + kTileBytes;
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {
                    // K-major: +32 B inside the swizzle atom per 16 K-elements;  MN-major: +2 groups of 8 K-rows
                    const uint64_t da = p.a_mn ? tc::make_sdesc(sa + k * 2048, kTileBytes / 2, 1024)
                                               : tc::make_sdesc(sa + k * 32, 16, 1024);
                    const uint64_t db = p.b_mn ? tc::make_sdesc(sb + k * 2048, kTileBytes / 2, 1024)
                                               : tc::make_sdesc(sb + k * 32, 16, 1024);
                    tc::umma_f16(tmem_base, da, db, idesc, (kb | k) != 0);
                }
                tc::umma_commit(&empty[s]);                           // frees the smem slot when these MMAs retire
            }
            tc::umma_commit(acc_full);                                // accumulator complete
        }
    } else {
        // ---- epilogue: warp w may touch TMEM lanes [32*(w%4), 32*(w%4)+32)
        const int q = warp & 3;
        tc::mbar_wait(acc_full, 0);
        tc::tc_fence_after();
        const int row = m0 + q * 32 + lane;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
            uint32_t r[32];
            tc::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), r);
            tc::tmem_ld_wait();
            const int col0 = n0 + c * 32;
            if (row < p.M && col0 < p.N) {
                if (nkb == 0) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) r[j] = 0u;
                }
                if (p.c_bf16) {
                    __nv_bfloat16* crow = reinterpret_cast<__nv_bfloat16*>(p.C) + (long long)row * p.ldc + col0;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        if (col0 + j < p.N) {
                            float v = __uint_as_float(r[j]);
                            if (p.bias) v += __ldg(p.bias + col0 + j);
                            if (p.accumulate) v += __bfloat162float(crow[j]);
                            crow[j] = __float2bfloat16(v);
                        }
                    }
                } else {
                    float* crow = reinterpret_cast<float*>(p.C) + (long long)row * p.ldc + col0;
                    const bool vec = ((reinterpret_cast<uintptr_t>(crow) & 15) == 0) && (col0 + 32 <= p.N);
                    if (vec) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            float4 v = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                   __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
                            if (p.bias) {
                                const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
                                v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w;
                            }
                            if (p.accumulate) {
                                const float4 o = *reinterpret_cast<const float4*>(crow + j);
                                v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
                            }
                            *reinterpret_cast<float4*>(crow + j) = v;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            if (col0 + j < p.N) {
                                float v = __uint_as_float(r[j]);
                                if (p.bias) v += __ldg(p.bias + col0 + j);
                                if (p.accumulate) v += crow[j];
                                crow[j] = v;
                            }
                        }
                    }
                }
            }
        }
        tc::tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem_base, kTmemCols);
    }
}

}  // namespace

// ---- host helpers shared by the tensor-core kernels -------------------------------------------------
fn_PFN_encodeTiled fn_get_encode_tiled() {
    static fn_PFN_encodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<fn_PFN_encodeTiled>(p);
    });
    return fn;
}

int fn_make_tmap_bf16_2d(CUtensorMap* out, const void* base, unsigned long long rows, unsigned long long cols,
                         unsigned long long ld, unsigned box_rows, unsigned box_cols) {
    fn_PFN_encodeTiled enc = fn_get_encode_tiled();
    FN_REQUIRE(enc, "cuTensorMapEncodeTiled entry point unavailable");
    FN_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld * 2) % 16 == 0,
               "TMA operand needs a 16-byte aligned base and row pitch (ld=%llu)", ld);
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld * 2};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FN_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu ld=%llu", (int)r, rows, cols, ld);
    return FN_OK;
}

extern "C" int fn_tc_gemm_bf16(const void* A, long long lda, int a_mn_major, const void* B, long long ldb,
                               int b_mn_major, void* C, long long ldc, int c_bf16, const float* bias, int M, int N,
                               int K, int accumulate, void* stream) {
    FN_REQUIRE(A && B && C && M > 0 && N > 0 && K >= 0, "fn_tc_gemm_bf16: bad args");
    CUtensorMap tmA, tmB;
    int rc;
    // K-major operand: memory [rows = M|N][cols = K];  MN-major operand: memory [rows = K][cols = M|N]
    rc = a_mn_major ? fn_make_tmap_bf16_2d(&tmA, A, K, M, lda, 64, 64) : fn_make_tmap_bf16_2d(&tmA, A, M, K, lda, 128, 64);
    if (rc) return rc;
    rc = b_mn_major ? fn_make_tmap_bf16_2d(&tmB, B, K, N, ldb, 64, 64) : fn_make_tmap_bf16_2d(&tmB, B, N, K, ldb, 128, 64);
    if (rc) return rc;
    GemmParams p{C, bias, ldc, M, N, K, a_mn_major ? 1 : 0, b_mn_major ? 1 : 0, c_bf16 ? 1 : 0, accumulate ? 1 : 0};
    static bool attr_done = false;
    if (!attr_done) {
        FN_CHECK_CUDA(cudaFuncSetAttribute(tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        attr_done = true;
    }
    dim3 grid(fn_cdiv(N, BN), fn_cdiv(M, BM));
    tc_gemm_kernel<<<grid, kThreads, kSmemBytes, (cudaStream_t)stream>>>(tmA, tmB, p);
    FN_LAUNCH_CHECK();
    return FN_OK;
}
