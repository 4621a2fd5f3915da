// bf16 tensor-core GEMM on CTA PAIRS (the large shapes of fn_tc_gemm_bf16 / _splitk / x3): persistent 256x256 output tiles,
// tcgen05.mma.cta_group::2 (M = 256: CTA r of the pair supplies rows [128 r, 128 r + 128) of A and columns
// [128 r, 128 r + 128) of B, and receives the [128 x 256] accumulator of its rows in its own TMEM), TWO accumulators in
// TMEM (2 x 256 columns) so the epilogue of one tile drains while the tensor cores fill the next, 6 x 32 KB stages per CTA.
//
// Why pairs: a 128x128 tile needs 32 KB of operands per 256 MMA cycles (128 B/clk per CTA, two CTAs per SM) -- more than an
// SM ingests -- so fn_tc_gemm.cu tops out at 0.63-0.85 of cuBLAS on the T*B-row products; the pair tile needs 32 KB per 512
// cycles and CTA (64 B/clk).
// Warps: 0 = TMA producer of A, 1 = TMA producer of B (both CTAs; every load completes on the LEADER's barrier), 2 = TMEM
// allocator + MMA issuer (leader only), 3..6 = epilogue (TMEM lane quarter = warp % 4).
// Work items = (split-K slice, tile), dealt round-robin to the pairs; the K loop runs over the bf16x3 plane products like
// fn_tc_gemm.cu.  Split-K partials go to the caller's workspace and are reduced in fixed order (deterministic).
#include <stdlib.h>
#include <string.h>

#include "fn_tc.cuh"
#include "fn_tc_gemm_epi.cuh"

namespace {

constexpr int BM = 256, BK = 64;                             // BN: template parameter (256, or 176 for the vocabulary)
constexpr int kStages = 6;
constexpr int kTileBytes = 128 * BK * 2;                    // 16 KB: one CTA's share of an operand tile
constexpr int kStageBytes = 2 * kTileBytes;                 // per CTA
constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr int kThreads = 7 * 32;
constexpr uint32_t kTmemCols = 512;                         // two [128 x 256] fp32 accumulators

struct Gemm2Params {
    GemmEpi epi;
    int K, a_mn, b_mn;
    int kb_per_split;
    int ncombo, nkb_base, combo_sel;                        // plane products (see GemmParams in fn_tc_gemm.cu)
    int tiles_n, tiles, items;                              // items = splits * tiles
};

// BN = 256, or 176 (each CTA supplies 88 columns of B; its 128-row B tile is loaded whole, the MMA reads the first 88 rows):
// two 176-column tiles cover the 342-wide vocabulary with 3 % padding instead of 33 %.
template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
tc_gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmAlo, const __grid_constant__ CUtensorMap tmBlo, const Gemm2Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);     // used in the leader only
    uint64_t* empty = full + kStages;
    uint64_t* acc_full = empty + kStages;                   // [2]
    uint64_t* acc_empty = acc_full + 2;                     // [2], used in the leader only
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = tc::cluster_ctarank();            // 0 = leader of the pair
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    const int nkb_total = p.nkb_base * p.ncombo;

    if (warp == 0 && lane == 0) {
        tc::prefetch_tmap(&tmA);
        tc::prefetch_tmap(&tmB);
        if (p.ncombo > 1) { tc::prefetch_tmap(&tmAlo); tc::prefetch_tmap(&tmBlo); }
        for (int s = 0; s < kStages; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { tc::mbar_init(&acc_full[a], 1); tc::mbar_init(&acc_empty[a], 8); }   // 4 epilogue warps x 2 CTAs
        tc::fence_barrier_init();
    }
    if (warp == 2) tc::tmem_alloc_2cta(tmem_slot, kTmemCols);
    tc::tc_fence_before();
    __syncthreads();
    tc::cluster_sync();                                     // the peer's barriers exist before anything arrives on them
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 2) {
        // ---- producers: warp 0 streams this CTA's 128 rows of A, warp 1 its 128 columns of B; a tile is two 64-row boxes
        // (8 KB) issued by two lanes.  Loads of both CTAs complete on the leader's `full` barrier.
        const int operand = warp;
        const uint32_t full0 = tc::smem_u32(full), full_l = full0 & tc::kPeerBitMask, empty0 = tc::smem_u32(empty);
        const uint32_t s0 = tc::smem_u32(smem) + operand * kTileBytes;
        const int mn_major = operand ? p.b_mn : p.a_mn;
        uint32_t st = 0, ph = 1;
        for (int item = pair; item < p.items; item += npairs) {
            const int z = item / p.tiles, t = item - z * p.tiles;
            const int tm = t / p.tiles_n, tn = t - tm * p.tiles_n;
            const int r0 = operand ? tn * BN + (int)rank * (BN / 2) : tm * BM + (int)rank * 128;
            const int kb_first = z * p.kb_per_split;
            const int nkb = min(p.kb_per_split, nkb_total - kb_first);
            int combo = 0, kk = kb_first;
            while (p.nkb_base > 0 && kk >= p.nkb_base) { kk -= p.nkb_base; ++combo; }
            for (int kb = 0; kb < nkb; ++kb) {
                const int sel = (p.combo_sel >> (2 * combo)) & 3;
                const CUtensorMap* tmap = operand ? ((sel & 2) ? &tmBlo : &tmB) : ((sel & 1) ? &tmAlo : &tmA);
                tc::mbar_wait_u32(empty0 + st * 8u, ph);
                if (operand == 0 && rank == 0 && lane == 0) tc::mbar_arrive_expect_tx_u32(full0 + st * 8u, 2 * kStageBytes);
                if (lane < 2) {
                    const int rr = r0 + lane * 64, k0 = kk * BK;
                    tc::tma_load_2d_2cta_u32(s0 + st * kStageBytes + lane * (kTileBytes / 2), tmap, full_l + st * 8u,
                                             mn_major ? rr : k0, mn_major ? k0 : rr);
                }
                __syncwarp();
                if (++st == kStages) { st = 0; ph ^= 1u; }
                if (++kk == p.nkb_base) { kk = 0; ++combo; }
            }
        }
    } else if (warp == 2) {
        // ---- MMA issuer (leader only): one M = 256, N = 256 product per 16 K
        if (rank == 0) {
            const uint32_t idesc = tc::make_idesc_bf16(BM, BN, p.a_mn, p.b_mn);
            const uint32_t full0 = tc::smem_u32(full), empty0 = tc::smem_u32(empty), s0 = tc::smem_u32(smem);
            const uint32_t accf0 = tc::smem_u32(acc_full), acce0 = tc::smem_u32(acc_empty);
            const uint64_t da0 = p.a_mn ? tc::make_sdesc(s0, kTileBytes / 2, 1024) : tc::make_sdesc(s0, 16, 1024);
            const uint64_t db0 = p.b_mn ? tc::make_sdesc(s0 + kTileBytes, kTileBytes / 2, 1024) : tc::make_sdesc(s0 + kTileBytes, 16, 1024);
            const uint32_t ka = p.a_mn ? (2048u >> 4) : (32u >> 4), kbs = p.b_mn ? (2048u >> 4) : (32u >> 4);
            uint32_t st = 0, ph = 0, it = 0;
            for (int item = pair; item < p.items; item += npairs, ++it) {
                const int z = item / p.tiles;
                const int kb_first = z * p.kb_per_split;
                const int nkb = min(p.kb_per_split, nkb_total - kb_first);
                const uint32_t a = it & 1u, use = it >> 1;
                tc::mbar_wait_u32(acce0 + a * 8u, (use & 1u) ^ 1u);      // both CTAs' epilogues have drained this accumulator
                tc::tc_fence_after();
                const uint32_t d_tmem = tmem_base + a * 256u;
                for (int kb = 0; kb < nkb; ++kb) {
                    tc::mbar_wait_u32(full0 + st * 8u, ph);
                    tc::tc_fence_after();
                    if (tc::elect_one()) {
                        const uint64_t da = da0 + (uint64_t)(st * (kStageBytes >> 4)), db = db0 + (uint64_t)(st * (kStageBytes >> 4));
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k)
                            tc::umma_f16_2cta(d_tmem, da + (uint64_t)(k * ka), db + (uint64_t)(k * kbs), idesc, (uint32_t)((kb | k) != 0));
                        tc::umma_commit_2cta_mc_u32(empty0 + st * 8u, 3);                 // slot free in both CTAs
                        if (kb == nkb - 1) tc::umma_commit_2cta_mc_u32(accf0 + a * 8u, 3);  // accumulator complete, both CTAs
                    }
                    __syncwarp();
                    if (++st == kStages) { st = 0; ph ^= 1u; }
                }
            }
        }
    } else {
        // ---- epilogue: this CTA's 128 rows x 256 columns of the tile, 32 columns at a time
        const int q = warp & 3;
        const uint32_t accf0 = tc::smem_u32(acc_full), acce0 = tc::smem_u32(acc_empty);
        uint32_t it = 0;
        for (int item = pair; item < p.items; item += npairs, ++it) {
            const int z = item / p.tiles, t = item - z * p.tiles;
            const int tm = t / p.tiles_n, tn = t - tm * p.tiles_n;
            const uint32_t a = it & 1u, use = it >> 1;
            if (lane == 0) tc::mbar_wait_u32(accf0 + a * 8u, use & 1u);
            __syncwarp();
            tc::tc_fence_after();
            const int row = tm * BM + (int)rank * 128 + q * 32 + lane;
            const int n_lim = min(p.epi.N, (tn + 1) * BN);              // BN = 176: the last chunk of a tile is half valid
#pragma unroll 1
            for (int c = 0; c < (BN + 31) / 32; ++c) {
                uint32_t r[32];
                tc::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + a * 256u + (uint32_t)(c * 32), r);
                tc::tmem_ld_wait();
                const int col0 = tn * BN + c * 32;
                if (row < p.epi.M && col0 < n_lim) gemm_store_chunk(p.epi, z, row, col0, n_lim, r);
            }
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive_remote_u32(acce0 + a * 8u, 0);            // the leader's barrier
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::cluster_sync();                                     // the leader's MMAs read the peer's shared memory until the end
    if (warp == 2) {
        tc::tc_fence_after();
        tc::tmem_dealloc_2cta(tmem_base, kTmemCols);
    }
}

}  // namespace

void fn_splitk_reduce_launch(const float* partial, int splits, int M, int N, void* C, long long ldc, int c_bf16, const float* bias,
                             int accumulate, cudaStream_t st);

// Column-tile width for N: 176 where it pads less than 256 (N in (256, 352]: the 342-wide vocabulary), else 256.
static int pick_bn(int N) { return (N > 256 && N <= 352) ? 176 : 256; }

// Shapes the pair kernel takes (the others stay with the 128x128 kernel): at least one full 256-row tile and N wide enough
// that its column tiles are mostly full.
bool fn_tc_gemm2_eligible(int M, int N, int K) {
    static const int on = getenv("FN_GEMM_PAIR") ? atoi(getenv("FN_GEMM_PAIR")) : 1;
    if (!on) return false;
    const int bn = pick_bn(N), waste_n = (N + bn - 1) / bn * bn - N;
    return M >= 256 && N > 256 && waste_n * 4 <= N && K >= 64;
}

int fn_tc_gemm2_run(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmAlo, const CUtensorMap& tmBlo, int a_mn,
                    int b_mn, void* C, long long ldc, int c_bf16, const float* bias, int M, int N, int K, int accumulate,
                    int ncombo, int combo_sel, int splits, void* workspace, cudaStream_t st) {
    const int nkb_base = (K + BK - 1) / BK, nkb_total = nkb_base * ncombo;
    if (splits > nkb_total) splits = nkb_total;
    if (splits < 1) splits = 1;
    const int kb_per_split = (nkb_total + splits - 1) / splits;
    splits = (nkb_total + kb_per_split - 1) / kb_per_split;            // no empty split: every work item issues MMAs
    Gemm2Params p;
    p.epi = GemmEpi{C, bias, ldc, M, N, c_bf16 ? 1 : 0, accumulate ? 1 : 0, splits, reinterpret_cast<float*>(workspace)};
    p.K = K; p.a_mn = a_mn ? 1 : 0; p.b_mn = b_mn ? 1 : 0;
    p.kb_per_split = kb_per_split;
    p.ncombo = ncombo; p.nkb_base = nkb_base; p.combo_sel = combo_sel;
    const int BN = pick_bn(N);
    p.tiles_n = fn_cdiv(N, BN);
    p.tiles = fn_cdiv(M, BM) * p.tiles_n;
    p.items = p.tiles * splits;
    static bool attr_done = false;
    if (!attr_done) {
        FN_CHECK_CUDA(cudaFuncSetAttribute(tc_gemm2_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        FN_CHECK_CUDA(cudaFuncSetAttribute(tc_gemm2_kernel<176>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        attr_done = true;
    }
    const int max_pairs = fn_num_sms() / 2;
    const int pairs = p.items < max_pairs ? p.items : max_pairs;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = kSmemBytes;
    cfg.stream = st;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    if (BN == 256) FN_CHECK_CUDA(cudaLaunchKernelEx(&cfg, tc_gemm2_kernel<256>, tmA, tmB, tmAlo, tmBlo, p));
    else FN_CHECK_CUDA(cudaLaunchKernelEx(&cfg, tc_gemm2_kernel<176>, tmA, tmB, tmAlo, tmBlo, p));
    if (splits > 1) fn_splitk_reduce_launch(p.epi.partial, splits, M, N, C, ldc, p.epi.c_bf16, bias, p.epi.accumulate, st);
    FN_LAUNCH_CHECK();
    return FN_OK;
}
