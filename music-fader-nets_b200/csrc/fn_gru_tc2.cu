// Persistent time-loop GRU gate block, CTA-PAIR decomposition (tcgen05.mma cta_group::2, M = 256):
// the second-generation kernels behind fn_gru_seq_fwd_bf16 / fn_gru_seq_bwd_bf16 for chains of two 128-row batch tiles.
//
// Why.  The first-generation kernel (fn_gru_tc.cu) cuts a chain into 32-unit slices, one CTA each; every CTA streams
// the WHOLE state slab of BOTH batch tiles (512 KB) plus, forward at H = 1024, its whole weight slice twice (384 KB)
// per step from L2: ~900 KB per CTA and step, ~115 MB per step over the machine -- that, not the tensor pipe, set the
// ~19k-cycle step (r01 profiles).  Here a chain is cut into 64-unit slices owned by a PAIR of CTAs on one TPC:
//   * CTA r of the pair (r = 0, 1) owns batch tile r: it streams only ITS 128 x K state slab (A operand, 256 KB at
//     H = 1024) and holds only HALF of the pair's weight slice (B operand: forward 3 x 32 rows of W_hh, BPTT 32 rows of
//     W_hh^T); ONE tcgen05.mma.cta_group::2 of M = 256 multiplies both tiles with the whole slice -- the tensor cores
//     read the peer's half of B through the pair's shared-memory window.  Per CTA and step: 256 KB of state + the
//     non-resident part of 192 KB of weights, i.e. less than half of the first generation, and HALF the MMA
//     instructions (N = 192 | 64 per instruction instead of 96 | 64 on one tile).
//   * the new state / gate gradient tile is staged in 128B-swizzled shared memory by the epilogue warps and written by
//     ONE TMA store (cp.async.bulk.tensor shared -> global), after which a single thread publishes the step with one
//     red.release per CTA (the first generation stored one 128-byte line per lane and published once per warp).
//   * 16 epilogue warps (4 per TMEM lane quarter): a thread owns (row, 16 units) of its CTA's tile.
// Step protocol per batch tile: counter bar[chain][tile] counts the CTAs (one per pair) that have published step s of
// that tile; the state loaders of the 16 (H / 64) CTAs of the same rank wait for it before they stream slab s+1.
//
// Everything in memory is indexed BY TIME exactly as in fn_gru_tc.cu (hsx slabs, blocked saved gates, dg stream).
#include "fn_gru_tc_common.cuh"
#include "fn_gru_tc2.h"

namespace {

constexpr int kEpi0 = 4;                       // warps 0-3: state loader, MMA issuer (+ TMEM), weight loader, store + publish
constexpr int kEW2 = 16;                       // epilogue warps
constexpr int kThreads2 = (kEpi0 + kEW2) * 32; // 640
constexpr int kUP = 64;                        // hidden units per CTA pair
constexpr int kStg = 128 * kUP * 2;            // staging tile of the new state: 128 rows x 64 units bf16 = 16 KB

struct Tc2Chain {
    CUtensorMap tmW;       // fwd: W_hh [3H][H] as (k64, unit, gate, kchunk), box (64, 32, 3, KCH);  bwd: W_hh^T [H][3H] as (k64, unit, kchunk), box (64, 32, KCH)
    CUtensorMap tmA;       // streamed operand as (k64, row, kchunk, slab), box (64, 128, KCH, 1): fwd hsx [T+1][B][H]; bwd dg [T][B][4H]
    CUtensorMap tmS;       // TMA store of the step's result tile, (col, row, slab), box (64, 128, 1): fwd hsx; bwd dg
    const float* b_hh; const __nv_bfloat16* emb; const int32_t* ids; const float* proj; long long proj_ld;
    const __nv_bfloat16* dense;
    __nv_bfloat16* hsx; __nv_bfloat16* gates; float* h_final; long long h_final_ld;
    const void* dhs; const float* dh_final; long long dh_final_ld;
    __nv_bfloat16* dg; float* dh0;
    int reverse, dhs_f32;
};
struct Tc2Launch {
    Tc2Chain c[kMaxChainsTc];
    unsigned* bar;         // per chain 16 counters (one per batch tile), zeroed by the host
    long long* dbg;        // profiling aid (fn_gru_debug_timeline): [step][64] clock64 stamps of CTA 0, or NULL
    int n_chains, npairs, B, T, H;
    int stages, kres, wst; // state-ring stages (of KCH chunks); resident weight chunks; weight-ring slots (of KCH chunks)
};

struct Smem2 {
    uint8_t *W, *WR, *A, *stg;
    uint64_t *full, *empty, *wfull, *wempty, *wbar, *acc_full;
    uint32_t* tmem_slot;
    float* bias;
};
constexpr size_t kSmemTail2 = 1024 /*align*/ + 1024 /*barriers + tmem slot*/ + 3 * kUP * 4 /*bias*/;
__device__ __forceinline__ Smem2 carve2(uint8_t* raw, int w_res_bytes, int w_ring_bytes, int a_ring_bytes) {
    Smem2 s;
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    s.W = base; s.WR = s.W + w_res_bytes; s.A = s.WR + w_ring_bytes; s.stg = s.A + a_ring_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s.stg + kStg);
    s.full = bars; s.empty = s.full + kMaxStages; s.wfull = s.empty + kMaxStages; s.wempty = s.wfull + kMaxWst;
    s.wbar = s.wempty + kMaxWst; s.acc_full = s.wbar + 1;
    s.tmem_slot = reinterpret_cast<uint32_t*>(s.acc_full + 1);
    s.bias = reinterpret_cast<float*>(bars + 128);
    return s;
}

// swizzled (128B) byte offset of the 16-byte chunk `k16` (0..7) of row `row` in a [rows][128 B] tile
__device__ __forceinline__ uint32_t swz(int row, int k16) { return (uint32_t)(row * 128 + ((k16 ^ (row & 7)) << 4)); }
__device__ __forceinline__ void sts16(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&h);
}
// 16 fp32 -> bf16 -> two 16-byte chunks of a swizzled staging row
__device__ __forceinline__ void stage16(uint32_t tile, int row, int k16, const float (&v)[16]) {
    sts16(tile + swz(row, k16), pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
    sts16(tile + swz(row, k16 + 1), pack2(v[8], v[9]), pack2(v[10], v[11]), pack2(v[12], v[13]), pack2(v[14], v[15]));
}
// Step hand-over between CTAs.  The state tile is written by the ASYNC proxy (TMA store) and read by the async proxy (TMA
// loads); cp.async.bulk.wait_group 0 returns once the tile's writes are performed in L2 -- the only level the consumers'
// TMA loads read -- so the counter increment that follows needs neither a cross-proxy fence nor a releasing fence of
// its own (each measured at ~1.4k cycles on the step's critical path; FN_GRU2_FENCES=1 restores both, for A/B runs).
// The consumer side keeps its acquire load: it orders the loader's TMA issue after the observation.
#ifndef FN_GRU2_FENCES
#define FN_GRU2_FENCES 0
#endif
__device__ __forceinline__ void publish2(unsigned* ctr) {
#if FN_GRU2_FENCES
    fn_red_release(ctr, 1u);
#else
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(ctr), "r"(1u) : "memory");
#endif
}
constexpr int kBarStage = 2;
__device__ __forceinline__ long long clk() { long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory"); return t; }
#define FN_STAMP2(i, k)                                                         \
    do {                                                                        \
        if (dbg_on) P.dbg[(long long)(i) * 64 + (k)] = clk();                   \
    } while (0)                   // named barrier: epilogue warps (arrive) -> store warp (sync)

// =====================================================================================================
// Forward.  Accumulator columns of a CTA (its 128 rows x the pair's 192 gate columns):
//   [ r z n of units 0..31 (the B half held by CTA 0) | r z n of units 32..63 (held by CTA 1) ], 32 columns each;
// columns [192, 384) hold the time-invariant part of the pre-activations in the same order.
// =====================================================================================================
template <int KCH>
__global__ void __launch_bounds__(kThreads2, 1) gru2_fwd_kernel(const __grid_constant__ Tc2Launch P) {
    constexpr int N = 3 * kUP;                                 // 192
    constexpr int kWChunk = (N / 2) * 128;                     // 12 KB: one 64-wide K chunk of this CTA's half of the slice
    constexpr uint32_t kTmemCols = 512;                        // 192 accumulator + 192 projection columns
    constexpr uint32_t a_stage = KCH * kATile, w_slot = KCH * kWChunk;
    extern __shared__ uint8_t smem_raw[];
    const int H = P.H, B = P.B, T = P.T, S = P.stages, WST = P.wst, kres = P.kres;
    const int nkc = H / 64, nst = nkc / KCH, nsst = (nkc - kres) / KCH;          // stages per step; the streamed-weight ones come first
    const Smem2 sm = carve2(smem_raw, kres * kWChunk, WST * (int)w_slot, S * (int)a_stage);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = tc::cluster_ctarank();               // 0 = leader (issues the MMAs), owns batch tile `rank`
    const int pair = blockIdx.x >> 1;
    const int chain = pair / P.npairs, slice = pair % P.npairs;
    const Tc2Chain& c = P.c[chain];
    unsigned* gflag = P.bar + chain * 16 + rank;               // publishes of this CTA's batch tile
    const int u0 = slice * kUP;                                // first unit of the pair
    const int b0 = (int)rank * 128;                            // first row of this CTA's batch tile
    const bool dbg_on = P.dbg != nullptr && blockIdx.x == 0 && lane == 0;

    if (warp == 0 && lane == 0) {
        tc::prefetch_tmap(&c.tmW); tc::prefetch_tmap(&c.tmA); tc::prefetch_tmap(&c.tmS);
        for (int i = 0; i < S; ++i) { tc::mbar_init(&sm.full[i], 1); tc::mbar_init(&sm.empty[i], 1); }
        for (int i = 0; i < WST; ++i) { tc::mbar_init(&sm.wfull[i], 1); tc::mbar_init(&sm.wempty[i], 1); }
        tc::mbar_init(sm.wbar, 1); tc::mbar_init(sm.acc_full, 1);
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc_2cta(sm.tmem_slot, kTmemCols);
    if (warp >= kEpi0) {
        for (int i = threadIdx.x - kEpi0 * 32; i < 3 * kUP; i += kEW2 * 32) sm.bias[i] = c.b_hh[(i / kUP) * H + u0 + (i % kUP)];
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::cluster_sync();                                        // the peer's barriers exist before anything arrives on them
    tc::tc_fence_after();
    const uint32_t tmem_base = *sm.tmem_slot;

    if (warp == 0) {
        // ------------------------------- state loader (both CTAs: own batch tile) ------------------------------
        // Every load of the pair completes on the LEADER's barrier, which expects the bytes of both CTAs.
        const uint32_t full_l = tc::smem_u32(sm.full) & tc::kPeerBitMask, empty0 = tc::smem_u32(sm.empty);
        const uint32_t full0 = tc::smem_u32(sm.full);
        const uint32_t a0 = tc::smem_u32(sm.A);
        if (kres > 0 && tc::elect_one()) {
            if (rank == 0) tc::mbar_arrive_expect_tx(sm.wbar, (uint32_t)(2 * kres * kWChunk));
            const uint32_t wbar_l = tc::smem_u32(sm.wbar) & tc::kPeerBitMask;
            for (int kc = 0; kc < kres; kc += KCH)
                tc::tma_load_4d_2cta_u32(tc::smem_u32(sm.W) + kc * kWChunk, &c.tmW, wbar_l, 0, u0 + (int)rank * 32, 0, kc);
        }
        __syncwarp();
        uint32_t st = 0, ph = 1;
        for (int i = 0; i < T; ++i) {
            const int slab = c.reverse ? T - i : i;            // the state before step i
            FN_STAMP2(i, 0);
            if (i > 0) {
                fn_spin_until(gflag, (unsigned)(P.npairs * i));
#if FN_GRU2_FENCES
                asm volatile("fence.proxy.async.global;" ::: "memory");
#endif
            }
            FN_STAMP2(i, 1);
            int kc = kres;                                     // K order: streamed-weight chunks [kres, nkc) first, then [0, kres)
            for (int j = 0; j < nst; ++j) {
                if (j == nsst) kc = 0;
                tc::mbar_wait_u32(empty0 + st * 8u, ph);
                if (tc::elect_one()) {
                    if (rank == 0) tc::mbar_arrive_expect_tx_u32(full0 + st * 8u, 2 * a_stage);
                    tc::tma_load_4d_2cta_u32(a0 + st * a_stage, &c.tmA, full_l + st * 8u, 0, b0, kc, slab);
                }
                __syncwarp();
                kc += KCH;
                if (++st == (uint32_t)S) { st = 0; ph ^= 1u; }
            }
            FN_STAMP2(i, 2);
        }
    } else if (warp == 1) {
        // ------------------------------- MMA issuer (leader only) -----------------------------------------------
        if (rank == 0) {
            const uint32_t idesc = tc::make_idesc_bf16(256, N, 0, 0);
            if (kres > 0) tc::mbar_wait(sm.wbar, 0);
            const uint32_t full0 = tc::smem_u32(sm.full), empty0 = tc::smem_u32(sm.empty);
            const uint32_t wfull0 = tc::smem_u32(sm.wfull), wempty0 = tc::smem_u32(sm.wempty), accf = tc::smem_u32(sm.acc_full);
            const uint64_t adesc0 = tc::make_sdesc(tc::smem_u32(sm.A), 16, 1024);
            const uint64_t bdesc0 = tc::make_sdesc(tc::smem_u32(sm.W), 16, 1024);
            const uint64_t wdesc0 = tc::make_sdesc(tc::smem_u32(sm.WR), 16, 1024);
            uint32_t st = 0, ph = 0, ws = 0, wph = 0;
            auto issue = [&](uint64_t ad, uint64_t bd, bool first) {
#pragma unroll
                for (int q = 0; q < KCH; ++q) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        tc::umma_f16_2cta(tmem_base, ad + (uint64_t)(q * (kATile >> 4) + 2 * k), bd + (uint64_t)(q * (kWChunk >> 4) + 2 * k), idesc,
                                          (q | k) ? 1u : (first ? 0u : 1u));
                }
            };
            for (int i = 0; i < T; ++i) {
                // (the accumulator is free: every chunk of this step's state was published after both CTAs' epilogues
                // had read the previous accumulator -- the step dependency itself orders the reuse)
                for (int j = 0; j < nsst; ++j) {
                    tc::mbar_wait_u32(wfull0 + ws * 8u, wph);
                    FN_STAMP2(i, 32 + j);
                    tc::mbar_wait_u32(full0 + st * 8u, ph);
                    tc::tc_fence_after();
                    FN_STAMP2(i, 16 + j);
                    if (tc::elect_one()) {
                        issue(adesc0 + (uint64_t)(st * (a_stage >> 4)), wdesc0 + (uint64_t)(ws * (w_slot >> 4)), j == 0);
                        tc::umma_commit_2cta_mc_u32(wempty0 + ws * 8u, 3);
                        tc::umma_commit_2cta_mc_u32(empty0 + st * 8u, 3);
                        if (j == nst - 1) tc::umma_commit_2cta_mc_u32(accf, 3);
                    }
                    __syncwarp();
                    if (++ws == (uint32_t)WST) { ws = 0; wph ^= 1u; }
                    if (++st == (uint32_t)S) { st = 0; ph ^= 1u; }
                }
                uint64_t bd = bdesc0;
                for (int j = nsst; j < nst; ++j) {
                    tc::mbar_wait_u32(full0 + st * 8u, ph);
                    tc::tc_fence_after();
                    FN_STAMP2(i, 16 + j);
                    if (tc::elect_one()) {
                        issue(adesc0 + (uint64_t)(st * (a_stage >> 4)), bd, j == 0);
                        tc::umma_commit_2cta_mc_u32(empty0 + st * 8u, 3);
                        if (j == nst - 1) tc::umma_commit_2cta_mc_u32(accf, 3);
                    }
                    __syncwarp();
                    bd += (uint64_t)(w_slot >> 4);
                    if (++st == (uint32_t)S) { st = 0; ph ^= 1u; }
                }
                FN_STAMP2(i, 4);
            }
        }
    } else if (warp == 2) {
        // ------------------------------- streamed part of the weight half (both CTAs) ---------------------------
        // Independent of the recurrence: runs ahead of the step barrier, up to the ring depth.
        if (nsst > 0) {
            const uint32_t wr0 = tc::smem_u32(sm.WR), wfull0 = tc::smem_u32(sm.wfull), wempty0 = tc::smem_u32(sm.wempty);
            const uint32_t wfull_l = wfull0 & tc::kPeerBitMask;
            uint32_t ws = 0, wph = 1;
            for (int i = 0; i < T; ++i) {
                for (int j = 0; j < nsst; ++j) {
                    tc::mbar_wait_u32(wempty0 + ws * 8u, wph);
                    if (tc::elect_one()) {
                        if (rank == 0) tc::mbar_arrive_expect_tx_u32(wfull0 + ws * 8u, 2 * w_slot);
                        tc::tma_load_4d_2cta_u32(wr0 + ws * w_slot, &c.tmW, wfull_l + ws * 8u, 0, u0 + (int)rank * 32, 0, kres + j * KCH);
                    }
                    __syncwarp();
                    if (++ws == (uint32_t)WST) { ws = 0; wph ^= 1u; }
                }
            }
        }
    } else if (warp == 3) {
        // ------------------------------- store + publish ---------------------------------------------------------
        const uint32_t stg = tc::smem_u32(sm.stg);
        for (int s = 0; s < T; ++s) {
            const int tau = c.reverse ? T - 1 - s : s;
            asm volatile("bar.sync %0, %1;" ::"n"(kBarStage), "n"((kEW2 + 1) * 32) : "memory");
            long long t8 = 0, t9 = 0, t12 = 0, t10 = 0;
            if (dbg_on) t8 = clk();
            if (tc::elect_one()) {
                tc::tma_store_3d_u32(&c.tmS, stg, u0, b0, c.reverse ? tau : tau + 1);
                tc::bulk_commit_group();
                tc::bulk_wait_group<0>();                      // the tile is in global memory (and the staging tile reusable)
            }
            __syncwarp();
            if (dbg_on) t9 = clk();
            if (tc::elect_one()) {
                if (s + 1 < T) {
#if FN_GRU2_FENCES
                    asm volatile("fence.proxy.async.global;" ::: "memory");
#endif
                }
            }
            __syncwarp();
            if (dbg_on) t12 = clk();
            if (tc::elect_one()) {
                if (s + 1 < T) publish2(gflag);
            }
            __syncwarp();
            if (dbg_on) {
                t10 = clk();
                long long* dp = P.dbg + (long long)s * 64;
                dp[8] = t8; dp[9] = t9; dp[12] = t12; dp[10] = t10;
            }
        }
    } else {
        // ------------------------------- gate epilogue (16 warps) ------------------------------------------------
        constexpr int UT = 16;
        const int q = warp & 3;                                // TMEM lane quarter
        const int grp = (warp - kEpi0) >> 2;                   // 16-unit group of the pair's 64 units
        const int uu = grp * UT, u = u0 + uu;
        const uint32_t col = (uint32_t)((grp >> 1) * 96 + (grp & 1) * UT);       // column of gate r of this group; z: +32, n: +64
        const uint32_t t_acc = tmem_base + ((uint32_t)(q * 32) << 16) + col, t_prj = t_acc + N;
        const int rl = q * 32 + lane, b = b0 + rl;
        const bool row_ok = b < B;
        const uint32_t stg = tc::smem_u32(sm.stg);
        float hreg[UT];
        {
            float pr[UT], pz[UT], pn[UT];
#pragma unroll
            for (int j = 0; j < UT; ++j) { pr[j] = 0.f; pz[j] = 0.f; pn[j] = 0.f; hreg[j] = 0.f; }
            if (row_ok) {
                if (c.proj) {
                    const float* pj = c.proj + (long long)b * c.proj_ld + u;
                    ldf<UT>(pj, pr); ldf<UT>(pj + H, pz); ldf<UT>(pj + 2 * H, pn);
                }
                uint32_t hw[UT / 2];
                ldb_raw<UT>(c.hsx + ((long long)(c.reverse ? T : 0) * B + b) * H + u, hw, false);
                unpack<UT>(hw, hreg);
            }
#pragma unroll
            for (int j = 0; j < UT; ++j) { pr[j] += sm.bias[uu + j]; pz[j] += sm.bias[kUP + uu + j]; }
            tmem_st<UT>(t_prj, pr); tmem_st<UT>(t_prj + 32, pz); tmem_st<UT>(t_prj + 64, pn);
            tmem_st_wait();
        }
        int id_next = (c.emb && row_ok) ? c.ids[(long long)(c.reverse ? T - 1 : 0) * B + b] : 0;
        const bool has_in = (c.emb != nullptr) || (c.dense != nullptr);
        for (int s = 0; s < T; ++s) {
            const int tau = c.reverse ? T - 1 - s : s;
            const int tau_n = c.reverse ? tau - 1 : tau + 1;
            // ---- operand that does not depend on the recurrence: fetch before waiting for the MMAs
            uint32_t ir[UT / 2], iz[UT / 2], in_[UT / 2];
#pragma unroll
            for (int j = 0; j < UT / 2; ++j) { ir[j] = 0u; iz[j] = 0u; in_[j] = 0u; }
            if (row_ok && has_in) {
                const __nv_bfloat16* src = c.emb ? c.emb + (long long)id_next * 3 * H + u : c.dense + ((long long)tau * B + b) * 3 * H + u;
                ldb_raw<UT>(src, ir, false); ldb_raw<UT>(src + H, iz, false); ldb_raw<UT>(src + 2 * H, in_, false);
                if (c.emb && s + 1 < T) id_next = c.ids[(long long)tau_n * B + b];
            }
            const bool dbg_e = dbg_on && warp == kEpi0;
            if (dbg_e) P.dbg[(long long)s * 64 + 5] = clk();
            tc::mbar_wait_warp(sm.acc_full, s & 1);
            tc::tc_fence_after();
            if (dbg_e) P.dbg[(long long)s * 64 + 6] = clk();
            float r[UT], z[UT], n[UT], g[UT];
            {
                float a[UT], p[UT], x[UT];
                tmem_ld2<UT>(t_acc, a, t_prj, p);
                unpack<UT>(ir, x);
#pragma unroll
                for (int j = 0; j < UT; ++j) r[j] = fast_sigmoid(a[j] + p[j] + x[j]);
                tmem_ld2<UT>(t_acc + 32, a, t_prj + 32, p);
                unpack<UT>(iz, x);
#pragma unroll
                for (int j = 0; j < UT; ++j) z[j] = fast_sigmoid(a[j] + p[j] + x[j]);
            }
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {                   // the n gate in halves of 8 units (register pressure)
                float a[8], p[8];
                tmem_ld2<8>(t_acc + 64 + hh * 8, a, t_prj + 64 + hh * 8, p);
                if (hh == 1) tc::tc_fence_before();
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int jj = hh * 8 + j;
                    const uint32_t w = in_[jj >> 1];
                    const float x = (jj & 1) ? bf_hi(w) : bf_lo(w);
                    g[jj] = a[j] + sm.bias[2 * kUP + uu + jj];
                    n[jj] = fast_tanh(p[j] + x + r[jj] * g[jj]);
                    hreg[jj] = (1.f - z[jj]) * n[jj] + z[jj] * hreg[jj];
                }
            }
            // new state -> swizzled staging tile -> (store warp) one TMA store + one publish per CTA
            stage16(stg, rl, uu >> 3, hreg);
            tc::fence_proxy_async();
            if (P.dbg != nullptr && blockIdx.x == 0 && lane == 0) P.dbg[(long long)s * 64 + 40 + (warp - kEpi0)] = clk();
            asm volatile("bar.arrive %0, %1;" ::"n"(kBarStage), "n"((kEW2 + 1) * 32) : "memory");
            if (row_ok) {
                if (c.gates) {                                 // off the critical path
                    stb<UT>(c.gates + gate_off(tau, b, u, B, 4 * H), r);
                    stb<UT>(c.gates + gate_off(tau, b, H + u, B, 4 * H), z);
                    stb<UT>(c.gates + gate_off(tau, b, 2 * H + u, B, 4 * H), n);
                    stb<UT>(c.gates + gate_off(tau, b, 3 * H + u, B, 4 * H), g);
                }
                if (s == T - 1 && c.h_final) {
                    float* hf = c.h_final + (long long)b * c.h_final_ld + u;
#pragma unroll
                    for (int j = 0; j < UT; ++j) hf[j] = hreg[j];
                }
            }
            if (dbg_e) P.dbg[(long long)s * 64 + 11] = clk();
        }
        tc::tc_fence_before();
    }
    __syncthreads();
    tc::cluster_sync();                                        // the leader's MMAs read the peer's shared memory until the end
    if (warp == 1) {
        tc::tc_fence_after();
        tc::tmem_dealloc_2cta(tmem_base, kTmemCols);
    }
}

// ---- host -------------------------------------------------------------------------------------------
struct Plan2 { int kch, stages, kres, wst; size_t smem; bool ok; };

// nkc K chunks of `w_chunk` bytes per CTA; the state ring wants >= 2 stages of kch * 16 KB.
Plan2 plan2(int nkc, int w_chunk, const char* tag) {
    Plan2 pl{};
    static const int kch_env = env_int("FN_GRU2_KCH", 2), s_env = env_int("FN_GRU2_S", 3), wst_env = env_int("FN_GRU2_WST", 2);
    int kch = kch_env == 1 || kch_env == 2 || kch_env == 4 ? kch_env : 2;
    while (kch > 1 && nkc % kch) kch >>= 1;
    pl.kch = kch;
    const long long budget = (long long)fn_max_smem_optin() - (long long)kSmemTail2 - kStg;
    const long long a_stage = (long long)kch * kATile, w_slot = (long long)kch * w_chunk;
    long long room = budget - (long long)nkc * w_chunk;
    if (room >= 3 * a_stage) {                                  // the whole half-slice stays resident
        pl.kres = nkc; pl.wst = 0;
        pl.stages = (int)(room / a_stage);
    } else {
        int S = s_env < 2 ? 2 : s_env, wst = wst_env < 2 ? 2 : wst_env;
        if (wst > kMaxWst) wst = kMaxWst;
        long long kres = (budget - S * a_stage - wst * w_slot) / w_chunk;
        if (kres > nkc - kch) kres = nkc - kch;
        kres -= kres % kch;
        if (kres < 0) { pl.ok = false; return pl; }
        pl.kres = (int)kres; pl.wst = wst; pl.stages = S;
    }
    if (pl.stages > kMaxStages) pl.stages = kMaxStages;
    if (pl.stages > nkc / kch * 2) pl.stages = nkc / kch * 2;
    pl.ok = pl.stages >= 2;
    pl.smem = (size_t)((long long)pl.kres * w_chunk + pl.wst * w_slot + pl.stages * a_stage) + kStg + kSmemTail2;
    static const int verbose = env_int("FN_GRU_VERBOSE", 0);
    if (verbose && pl.ok)
        fprintf(stderr, "plan2 %s nkc=%d: kch=%d stages=%d kres=%d wst=%d smem=%zu\n", tag, nkc, pl.kch, pl.stages, pl.kres, pl.wst, pl.smem);
    return pl;
}

int make_tmap_bf16_4d(CUtensorMap* out, const void* base, const cuuint64_t (&dims)[4], const cuuint64_t (&strides_bytes)[3],
                      const cuuint32_t (&box)[4]) {
    fn_PFN_encodeTiled enc = fn_get_encode_tiled();
    FN_REQUIRE(enc, "cuTensorMapEncodeTiled entry point unavailable");
    FN_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA 4-D operand alignment");
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides_bytes, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FN_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(4d) failed (%d)", (int)r);
    return FN_OK;
}

template <typename K>
int launch2(K kernel, const Tc2Launch& P, size_t smem, cudaStream_t st) {
    const void* fn = (const void*)kernel;
    FN_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    const int ctas = P.n_chains * P.npairs * 2;
    cfg.gridDim = dim3(ctas);
    cfg.blockDim = dim3(kThreads2);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attrs[2];
    attrs[0].id = cudaLaunchAttributeClusterDimension;
    attrs[0].val.clusterDim.x = 2; attrs[0].val.clusterDim.y = 1; attrs[0].val.clusterDim.z = 1;
    attrs[1].id = cudaLaunchAttributeCooperative;           // every CTA must be co-resident: they wait on each other
    attrs[1].val.cooperative = 1;
    cfg.attrs = attrs;
    cfg.numAttrs = 2;
    int max_clusters = 0;
    FN_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, fn, &cfg));
    FN_REQUIRE(max_clusters * 2 >= ctas, "fn_gru_seq_bf16: %d CTAs in pairs are not co-resident (max %d pairs)", ctas, max_clusters);
    void* args[] = {(void*)&P};
    cudaError_t e = cudaLaunchKernelExC(&cfg, fn, args);
    if (e != cudaSuccess) {                                  // cooperative + cluster refused: residency was checked above
        (void)cudaGetLastError();
        cfg.numAttrs = 1;
        e = cudaLaunchKernelExC(&cfg, fn, args);
    }
    FN_CHECK_CUDA(e);
    return FN_OK;
}

}  // namespace

bool fn_gru2_eligible(bool bwd, int n_chains, int B, int H) {
    static const int on = env_int("FN_GRU_V2", 1);
    if (!on || (bwd && !(on & 2))) return false;
    if (B <= 128 || B > 256 || H % kUP || H < 128) return false;
    if (2 * (H / kUP) > fn_num_sms()) return false;
    return plan2(H / 64, 96 * 128, "fwd").ok;
}

int fn_gru2_run(bool bwd, const FnGruChainBf16* chains, int n_chains, int B, int T, int H, void* barrier_ws, size_t ws_bytes,
                cudaStream_t st) {
    FN_REQUIRE(!bwd, "fn_gru2_run: backward not built yet");
    FN_REQUIRE(chains && n_chains > 0 && barrier_ws && ws_bytes >= (size_t)64 * n_chains, "fn_gru_seq_bf16: bad arguments");
    FN_CHECK_CUDA(cudaMemsetAsync(barrier_ws, 0, (size_t)64 * n_chains, st));
    const int npairs = H / kUP, nkc = H / 64;
    const Plan2 pl = plan2(nkc, 96 * 128, "fwd");
    FN_REQUIRE(pl.ok, "fn_gru_seq_bf16: no shared-memory plan for H=%d", H);
    const int per_launch = fn_num_sms() / (2 * npairs) < kMaxChainsTc ? fn_num_sms() / (2 * npairs) : kMaxChainsTc;
    int done = 0;
    while (done < n_chains) {
        const int group = n_chains - done < per_launch ? n_chains - done : per_launch;
        Tc2Launch P;
        memset(&P, 0, sizeof(P));
        for (int i = 0; i < group; ++i) {
            const FnGruChainBf16& s = chains[done + i];
            Tc2Chain& d = P.c[i];
            FN_REQUIRE(s.hsx && s.w_hh && s.b_hh, "fn_gru_seq_fwd_bf16: chain %d misses buffers", done + i);
            FN_REQUIRE(!s.emb || s.ids, "fn_gru_seq_fwd_bf16: chain %d has emb without ids", done + i);
            FN_REQUIRE(!(s.emb && s.dense), "fn_gru_seq_fwd_bf16: chain %d has both a token and a dense input", done + i);
            int rc;
            {   // W_hh [3H][H] as (k in chunk, unit, gate, K chunk): one box = this CTA's 3 x 32 rows of KCH chunks
                const cuuint64_t dims[4] = {64, (cuuint64_t)H, 3, (cuuint64_t)nkc};
                const cuuint64_t str[3] = {(cuuint64_t)H * 2, (cuuint64_t)H * H * 2, 128};
                const cuuint32_t box[4] = {64, 32, 3, (cuuint32_t)pl.kch};
                if ((rc = make_tmap_bf16_4d(&d.tmW, s.w_hh, dims, str, box))) return rc;
            }
            {   // hsx [T+1][B][H] as (k in chunk, row, K chunk, slab)
                const cuuint64_t dims[4] = {64, (cuuint64_t)B, (cuuint64_t)nkc, (cuuint64_t)T + 1};
                const cuuint64_t str[3] = {(cuuint64_t)H * 2, 128, (cuuint64_t)B * H * 2};
                const cuuint32_t box[4] = {64, 128, (cuuint32_t)pl.kch, 1};
                if ((rc = make_tmap_bf16_4d(&d.tmA, s.hsx, dims, str, box))) return rc;
            }
            if ((rc = fn_make_tmap_bf16_3d(&d.tmS, s.hsx, T + 1, B, H, H, 128, 64))) return rc;
            d.b_hh = s.b_hh; d.emb = (const __nv_bfloat16*)s.emb; d.ids = s.ids; d.proj = s.proj; d.proj_ld = s.proj_ld;
            d.dense = (const __nv_bfloat16*)s.dense;
            d.hsx = (__nv_bfloat16*)s.hsx; d.gates = (__nv_bfloat16*)s.gates;
            d.h_final = s.h_final; d.h_final_ld = s.h_final_ld;
            d.reverse = s.reverse;
        }
        P.bar = reinterpret_cast<unsigned*>(barrier_ws) + done * 16;
        P.n_chains = group; P.npairs = npairs; P.B = B; P.T = T; P.H = H;
        P.stages = pl.stages; P.kres = pl.kres; P.wst = pl.wst;
        P.dbg = fn_gru_dbg_ptr();
        int rc;
        if (pl.kch == 4) rc = launch2(gru2_fwd_kernel<4>, P, pl.smem, st);
        else if (pl.kch == 2) rc = launch2(gru2_fwd_kernel<2>, P, pl.smem, st);
        else rc = launch2(gru2_fwd_kernel<1>, P, pl.smem, st);
        if (rc != FN_OK) return rc;
        done += group;
    }
    return FN_OK;
}
