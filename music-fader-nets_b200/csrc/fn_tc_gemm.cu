// bf16 tensor-core GEMM on tcgen05 (fn_tc_gemm_bf16): TMA -> 128B-swizzled smem ring -> tcgen05.mma with
// the fp32 accumulator in TMEM -> tcgen05.ld epilogue.  Warp-specialised: warp 0 = TMA producer,
// warp 1 = TMEM allocator + single-thread MMA issuer, warps 2..5 = epilogue (one TMEM lane quarter each).
// One 128x128 output tile per CTA; 3 smem stages (96 KB) so two CTAs share an SM and one tile's epilogue
// overlaps the other's main loop.  Both operands may be K-major or MN-major (transposed in memory), which
// covers forward (x W^T), dgrad (dy W) and wgrad (dy^T x) without materialising any transpose.
#include <stdlib.h>

#include <mutex>

#include "fn_tc.cuh"
#include "fn_tc_gemm_epi.cuh"

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
    int splits, kb_per_split;      // split-K: grid.z CTAs per output tile, each owning kb_per_split K blocks
    int issue_lanes;               // lanes of the producer warp that issue TMA boxes (1, 2 or 4)
    float* partial;                // [splits][M][N] fp32 partial products (splits > 1), reduced by splitk_reduce_kernel
    // bf16x3 mode (fn_tc_gemm_bf16x3): an fp32 operand is two bf16 planes (hi, lo) and the K loop runs over `ncombo` plane
    // products of nkb_base K blocks each -- hi*hi, then lo*hi / hi*lo for every operand that has a lo plane -- into ONE fp32
    // accumulator.  combo_sel: two bits per product, bit 0 = A's lo plane, bit 1 = B's lo plane.  Plain bf16: 1 product.
    int ncombo, nkb_base, combo_sel;
};

__global__ void __launch_bounds__(kThreads, 2)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmAlo, const __grid_constant__ CUtensorMap tmBlo, const GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
    uint64_t* empty = full + kStages;
    uint64_t* acc_full = empty + kStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int nkb_total = p.nkb_base * p.ncombo;
    const int kb_first = blockIdx.z * p.kb_per_split;
    const int nkb = max(0, min(p.kb_per_split, nkb_total - kb_first));   // K blocks of this split

    if (warp == 0 && lane == 0) {
        tc::prefetch_tmap(&tmA);
        tc::prefetch_tmap(&tmB);
        if (p.ncombo > 1) { tc::prefetch_tmap(&tmAlo); tc::prefetch_tmap(&tmBlo); }
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
        // control warps run converged; one elected lane issues (see fn_gru_tc.cu)
        const uint32_t full0 = tc::smem_u32(full), empty0 = tc::smem_u32(empty), s0 = tc::smem_u32(smem);
        uint32_t st = 0, ph = 1;
        int combo = 0, kk = kb_first;                                  // plane product and K block inside it
        while (p.nkb_base > 0 && kk >= p.nkb_base) { kk -= p.nkb_base; ++combo; }
        for (int kb = 0; kb < nkb; ++kb) {
            const int sel = (p.combo_sel >> (2 * combo)) & 3;
            tc::mbar_wait_u32(empty0 + st * 8u, ph);                   // slot free (passes immediately on first lap)
            // A stage is 4 boxes of 64 rows x 64 columns (8 KB): lanes 0..3 issue one each.  In tools/ubench_tc.cu one
            // issuing thread completes a TMA op per ~450 cycles whatever the box size while several lanes of one
            // converged instruction scale almost linearly; here it is worth +10-14 % (FN_GEMM_LANES=1 is the old way,
            // 8 boxes of 32 rows measured slower than 4 of 64).
            if (lane == 0) tc::mbar_arrive_expect_tx_u32(full0 + st * 8u, kStageBytes);
            for (int op = lane; op < 4; op += p.issue_lanes) {
                if (lane >= p.issue_lanes) break;
                const int operand = op >> 1, half = op & 1;
                const CUtensorMap* tm = operand ? ((sel & 2) ? &tmBlo : &tmB) : ((sel & 1) ? &tmAlo : &tmA);
                const int mn_major = operand ? p.b_mn : p.a_mn;
                const int r0 = (operand ? n0 : m0) + half * 64, k0 = kk * BK;
                const uint32_t dst = s0 + st * kStageBytes + operand * kTileBytes + half * (kTileBytes / 2);
                // K-major: box 64(K) x 64(M|N rows), the second half of the tile starts 64 rows * 128 B further on;
                // MN-major: box 64(M|N) x 64(K), the second 64-wide M|N block starts kTileBytes / 2 further on
                tc::tma_load_2d_u32(dst, tm, full0 + st * 8u, mn_major ? r0 : k0, mn_major ? k0 : r0);
            }
            __syncwarp();
            if (++st == kStages) { st = 0; ph ^= 1u; }
            if (++kk == p.nkb_base) { kk = 0; ++combo; }
        }
        (void)s0;
    } else if (warp == 1) {
        const uint32_t idesc = tc::make_idesc_bf16(BM, BN, p.a_mn, p.b_mn);
        const uint32_t full0 = tc::smem_u32(full), empty0 = tc::smem_u32(empty), s0 = tc::smem_u32(smem);
        // K-major: +32 B inside the swizzle atom per 16 K-elements;  MN-major: +2 groups of 8 K-rows (2048 B)
        const uint64_t da0 = p.a_mn ? tc::make_sdesc(s0, kTileBytes / 2, 1024) : tc::make_sdesc(s0, 16, 1024);
        const uint64_t db0 = p.b_mn ? tc::make_sdesc(s0 + kTileBytes, kTileBytes / 2, 1024) : tc::make_sdesc(s0 + kTileBytes, 16, 1024);
        const uint32_t ka = p.a_mn ? (2048u >> 4) : (32u >> 4), kbs = p.b_mn ? (2048u >> 4) : (32u >> 4);
        uint32_t st = 0, ph = 0;
        for (int kb = 0; kb < nkb; ++kb) {
            tc::mbar_wait_u32(full0 + st * 8u, ph);
            tc::tc_fence_after();
            if (tc::elect_one()) {
                const uint64_t da = da0 + (uint64_t)(st * (kStageBytes >> 4)), db = db0 + (uint64_t)(st * (kStageBytes >> 4));
#pragma unroll
                for (int k = 0; k < BK / 16; ++k)
                    tc::umma_f16(tmem_base, da + (uint64_t)(k * ka), db + (uint64_t)(k * kbs), idesc, (uint32_t)((kb | k) != 0));
                tc::umma_commit_u32(empty0 + st * 8u);                 // frees the smem slot when these MMAs retire
                if (kb == nkb - 1) tc::umma_commit(acc_full);          // accumulator complete
            }
            __syncwarp();
            if (++st == kStages) { st = 0; ph ^= 1u; }
        }
        if (nkb == 0 && tc::elect_one()) tc::mbar_arrive(acc_full);
    } else {
        // ---- epilogue: warp w may touch TMEM lanes [32*(w%4), 32*(w%4)+32)
        const int q = warp & 3;
        tc::mbar_wait_warp(acc_full, 0);
        tc::tc_fence_after();
        const int row = m0 + q * 32 + lane;
        const GemmEpi ep{p.C, p.bias, p.ldc, p.M, p.N, p.c_bf16, p.accumulate, p.splits, p.partial};
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
                gemm_store_chunk(ep, blockIdx.z, row, col0, p.N, r);
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

// C = (accumulate ? C : 0) + bias + sum_z partial[z]   (fixed summation order: deterministic)
__global__ void splitk_reduce_kernel(const float* __restrict__ partial, int splits, int M, int N, void* C, long long ldc,
                                     int c_bf16, const float* __restrict__ bias, int accumulate) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)M * N) return;
    const int m = (int)(i / N), n = (int)(i % N);
    float v = bias ? __ldg(bias + n) : 0.f;
    for (int z = 0; z < splits; ++z) v += partial[(long long)z * M * N + i];
    if (c_bf16) {
        __nv_bfloat16* c = reinterpret_cast<__nv_bfloat16*>(C) + (long long)m * ldc + n;
        if (accumulate) v += __bfloat162float(*c);
        *c = __float2bfloat16(v);
    } else {
        float* c = reinterpret_cast<float*>(C) + (long long)m * ldc + n;
        if (accumulate) v += *c;
        *c = v;
    }
}

}  // namespace

void fn_splitk_reduce_launch(const float* partial, int splits, int M, int N, void* C, long long ldc, int c_bf16, const float* bias,
                             int accumulate, cudaStream_t st) {
    const long long n = (long long)M * N;
    splitk_reduce_kernel<<<fn_cdiv(n, 256), 256, 0, st>>>(partial, splits, M, N, C, ldc, c_bf16, bias, accumulate);
}
bool fn_tc_gemm2_eligible(int M, int N, int K);
int fn_tc_gemm2_run(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmAlo, const CUtensorMap& tmBlo, int a_mn,
                    int b_mn, void* C, long long ldc, int c_bf16, const float* bias, int M, int N, int K, int accumulate,
                    int ncombo, int combo_sel, int splits, void* workspace, cudaStream_t st);

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

extern "C" size_t fn_tc_gemm_splitk_ws_bytes(int M, int N, int splits) {
    return splits > 1 ? (size_t)splits * M * N * sizeof(float) : 0;
}

extern "C" int fn_tc_gemm_bf16(const void* A, long long lda, int a_mn_major, const void* B, long long ldb,
                               int b_mn_major, void* C, long long ldc, int c_bf16, const float* bias, int M, int N,
                               int K, int accumulate, void* stream) {
    return fn_tc_gemm_bf16_splitk(A, lda, a_mn_major, B, ldb, b_mn_major, C, ldc, c_bf16, bias, M, N, K, accumulate, 1,
                                  nullptr, 0, stream);
}

extern "C" int fn_tc_gemm_bf16_splitk(const void* A, long long lda, int a_mn_major, const void* B, long long ldb,
                                      int b_mn_major, void* C, long long ldc, int c_bf16, const float* bias, int M, int N,
                                      int K, int accumulate, int splits, void* workspace, size_t ws_bytes, void* stream) {
    return fn_tc_gemm_bf16x3(A, lda, 0, a_mn_major, B, ldb, 0, b_mn_major, C, ldc, c_bf16, bias, M, N, K, accumulate, splits,
                             workspace, ws_bytes, stream);
}

extern "C" int fn_tc_gemm_bf16x3(const void* A, long long lda, long long a_lo_off, int a_mn_major, const void* B, long long ldb,
                                 long long b_lo_off, int b_mn_major, void* C, long long ldc, int c_bf16, const float* bias, int M,
                                 int N, int K, int accumulate, int splits, void* workspace, size_t ws_bytes, void* stream) {
    FN_REQUIRE(A && B && C && M > 0 && N > 0 && K >= 0, "fn_tc_gemm_bf16: bad args");
    const int nkb_base = (K + BK - 1) / BK;
    int ncombo = 1, combo_sel = 0;
    if (a_lo_off) { combo_sel |= 1 << (2 * ncombo); ++ncombo; }
    if (b_lo_off) { combo_sel |= 2 << (2 * ncombo); ++ncombo; }
    const int nkb_total = nkb_base * ncombo;
    if (splits < 1) splits = 1;
    if (splits > nkb_total) splits = nkb_total > 0 ? nkb_total : 1;
    FN_REQUIRE(splits == 1 || (workspace && ws_bytes >= fn_tc_gemm_splitk_ws_bytes(M, N, splits)),
               "fn_tc_gemm_bf16_splitk: workspace too small for %d splits", splits);
    CUtensorMap tmA, tmB, tmAlo, tmBlo;
    int rc;
    // K-major operand: memory [rows = M|N][cols = K];  MN-major operand: memory [rows = K][cols = M|N]
    rc = a_mn_major ? fn_make_tmap_bf16_2d(&tmA, A, K, M, lda, 64, 64) : fn_make_tmap_bf16_2d(&tmA, A, M, K, lda, 64, 64);
    if (rc) return rc;
    rc = b_mn_major ? fn_make_tmap_bf16_2d(&tmB, B, K, N, ldb, 64, 64) : fn_make_tmap_bf16_2d(&tmB, B, N, K, ldb, 64, 64);
    if (rc) return rc;
    tmAlo = tmA; tmBlo = tmB;
    if (a_lo_off) {
        const void* Al = reinterpret_cast<const __nv_bfloat16*>(A) + a_lo_off;
        rc = a_mn_major ? fn_make_tmap_bf16_2d(&tmAlo, Al, K, M, lda, 64, 64) : fn_make_tmap_bf16_2d(&tmAlo, Al, M, K, lda, 64, 64);
        if (rc) return rc;
    }
    if (b_lo_off) {
        const void* Bl = reinterpret_cast<const __nv_bfloat16*>(B) + b_lo_off;
        rc = b_mn_major ? fn_make_tmap_bf16_2d(&tmBlo, Bl, K, N, ldb, 64, 64) : fn_make_tmap_bf16_2d(&tmBlo, Bl, N, K, ldb, 64, 64);
        if (rc) return rc;
    }
    if (fn_tc_gemm2_eligible(M, N, K))          // large shapes: 256x256 tiles on CTA pairs (fn_tc_gemm2.cu)
        return fn_tc_gemm2_run(tmA, tmB, tmAlo, tmBlo, a_mn_major, b_mn_major, C, ldc, c_bf16, bias, M, N, K, accumulate, ncombo,
                               combo_sel, splits, workspace, (cudaStream_t)stream);
    static const int issue_lanes = getenv("FN_GEMM_LANES") ? atoi(getenv("FN_GEMM_LANES")) : 4;
    const int kb_per_split = splits > 1 ? (nkb_total + splits - 1) / splits : (nkb_total > 0 ? nkb_total : 1);
    GemmParams p{C, bias, ldc, M, N, K, a_mn_major ? 1 : 0, b_mn_major ? 1 : 0, c_bf16 ? 1 : 0, accumulate ? 1 : 0,
                 splits, kb_per_split, issue_lanes, reinterpret_cast<float*>(workspace), ncombo, nkb_base, combo_sel};
    static bool attr_done = false;
    if (!attr_done) {
        FN_CHECK_CUDA(cudaFuncSetAttribute(tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        attr_done = true;
    }
    dim3 grid(fn_cdiv(N, BN), fn_cdiv(M, BM), splits);
    tc_gemm_kernel<<<grid, kThreads, kSmemBytes, (cudaStream_t)stream>>>(tmA, tmB, tmAlo, tmBlo, p);
    FN_LAUNCH_CHECK();
    if (splits > 1) {
        const long long n = (long long)M * N;
        splitk_reduce_kernel<<<fn_cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(p.partial, splits, M, N, C, ldc, p.c_bf16, bias,
                                                                               p.accumulate);
        FN_LAUNCH_CHECK();
    }
    return FN_OK;
}
