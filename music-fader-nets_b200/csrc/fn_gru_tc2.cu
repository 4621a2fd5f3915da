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
    int n_chains, ppc, cpp, B, T, H;   // ppc: pairs per chain (H / UP); cpp: chains per pair (1 or 2)
    int mc;                            // pairs per cluster (1, 2 or 4): same-rank CTAs of a cluster share ONE multicast load of each state chunk
    int stages, kres, wst; // state-ring stages (of KCH chunks); resident weight chunks; weight-ring slots (of KCH chunks)
};

struct Smem2 {
    uint8_t *W, *WR, *A, *stg;
    uint64_t *full, *empty, *wfull, *wempty, *wbar, *acc_full;    // acc_full: one per chain of the pair
    uint32_t* tmem_slot;
    float* bias;
};
constexpr size_t kSmemTail2 = 1024 /*align*/ + 1024 /*barriers + tmem slot*/ + 1024 /*bias*/;
__device__ __forceinline__ Smem2 carve2(uint8_t* raw, int w_res_bytes, int w_ring_bytes, int a_ring_bytes, int stg_bytes) {
    Smem2 s;
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    s.W = base; s.WR = s.W + w_res_bytes; s.A = s.WR + w_ring_bytes; s.stg = s.A + a_ring_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s.stg + stg_bytes);
    s.full = bars; s.empty = s.full + kMaxStages; s.wfull = s.empty + kMaxStages; s.wempty = s.wfull + kMaxWst;
    s.wbar = s.wempty + kMaxWst; s.acc_full = s.wbar + 1;
    s.tmem_slot = reinterpret_cast<uint32_t*>(s.acc_full + 2);
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
// loads); cp.async.bulk.wait_group 0 returns once the tile's writes are complete for the issuing thread.  The counter increment
// that follows is a RELEASING red (bit 0 of FN_GRU2_FENCES, the default): with a relaxed increment (an earlier build of this
// round) about 3 % of fresh processes showed a run-to-run difference in the states of a few sequences during their first steps
// (tools/determinism_probe.py in a loop of fresh processes; 0 of 130 with the release) -- the tile's writes are not guaranteed
// to be visible GPU-wide before a relaxed increment is.  The release costs ~1 ms per config-3 step.  The cross-proxy fences on
// either side (bits 1 and 2, ~1.4 k cycles each on the step's critical path) stay off: the consumers' TMA loads are issued after
// the acquire load observed the counter, and both the writes and the reads of the tile go through the async proxy and L2.
#ifndef FN_GRU2_FENCES
#define FN_GRU2_FENCES 1      /* bit 0: releasing counter increment; bit 1: producer-side proxy fence; bit 2: consumer-side proxy fence */
#endif
__device__ __forceinline__ void publish2(unsigned* ctr) {
#if FN_GRU2_FENCES & 1
    fn_red_release(ctr, 1u);
#else
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(ctr), "r"(1u) : "memory");
#endif
}
constexpr int kBarStage = 2;                   // named barriers 2, 3: epilogue warps (arrive) -> store warp (sync), per chain of the pair
constexpr int kBarPub = 4;                     // named barriers 4, 5: store warp (arrive, after the publish) -> epilogue warps (sync)
__device__ __forceinline__ long long clk() { long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory"); return t; }
#define FN_STAMP2(i, k)                                                         \
    do {                                                                        \
        if (dbg_on) P.dbg[(long long)(i) * 64 + (k)] = clk();                   \
    } while (0)                   // named barrier: epilogue warps (arrive) -> store warp (sync)

// =====================================================================================================
// Forward.  Geometry: a pair owns kUP = 32 units of a chain (MMA N = 96: r, z, n), CTA r of the pair holds the
// 3 x 16 weight rows of units [u0 + 16 r, +16) (6 KB per 64-wide K chunk) and batch tile r.  A pair serves ONE or TWO
// chains ("chains per pair", cpp): with two, the chains are independent recurrences that take turns on the tensor
// cores -- while chain A's gate epilogue, TMA store, hand-over and the first loads of its next step run (the part of a
// step that is pure latency), the pair multiplies chain B.  Accumulator columns of chain k: [k*96, +96) =
// [ r z n of the leader's 16 units | r z n of the peer's 16 units ]; columns 192 + the same hold the time-invariant
// part of the pre-activations.
// =====================================================================================================
// Two geometries (template parameter UP = units of a chain per pair): UP = 32 with up to two chains per pair (above), and
// UP = 64 with one chain per pair (N = 192: half the state bytes per FLOP, for launches whose chains each get H/64 pairs).
template <int UP> struct Geo {
    static constexpr int kN = 3 * UP, kNW = kN / 2, kWCh = kNW * 128, kUT = UP / 4;
    static constexpr int kStg = 128 * UP * 2;                  // staging tile of a chain's new state: 128 rows x UP units bf16
    static constexpr int kMaxCpp = 64 / UP;
};
constexpr int kStgAll = 16384;                                 // staging bytes of a CTA (both geometries)
constexpr int kAccCols = 192;                                  // accumulator columns of a CTA (both geometries); projections follow

// swizzled (64B) byte offset of the 16-byte chunk `k16` (0..3) of row `row` in a [rows][64 B] tile (TMA SWIZZLE_64B)
__device__ __forceinline__ uint32_t swz64(int row, int k16) { return (uint32_t)(row * 64 + ((k16 ^ ((row >> 1) & 3)) << 4)); }

// sigmoid of two pre-activations with ONE SFU instruction (tanh.approx.f16x2): the gate epilogue is bound by the SFU
// (3 transcendentals per unit and row); r and z are stored in bf16 (2^-9) and tolerate the f16 evaluation (2^-11).
__device__ __forceinline__ void sigmoid2(float x0, float x1, float& y0, float& y1) {
    uint32_t h, t;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(0.5f * x1), "f"(0.5f * x0));
    asm("tanh.approx.f16x2 %0, %1;" : "=r"(t) : "r"(h));
    asm("fma.rn.f16x2 %0, %1, %2, %2;" : "=r"(h) : "r"(t), "r"(0x38003800u));      // 0.5 t + 0.5
    asm("{\n\t.reg .f16 lo, hi;\n\tmov.b32 {lo, hi}, %2;\n\tcvt.f32.f16 %0, lo;\n\tcvt.f32.f16 %1, hi;\n\t}" : "=f"(y0), "=f"(y1) : "r"(h));
}

template <int UP, int KCH>
__global__ void __launch_bounds__(kThreads2, 1) gru2_fwd_kernel(const __grid_constant__ Tc2Launch P) {
    constexpr int kUP = UP, kN = Geo<UP>::kN, kNW = Geo<UP>::kNW, kWCh = Geo<UP>::kWCh, kUT = Geo<UP>::kUT, kStgF = Geo<UP>::kStg;
    constexpr int kMaxCpp = Geo<UP>::kMaxCpp;
    constexpr uint32_t kTmemCols = 512;                        // 192 accumulator + 192 projection columns
    constexpr uint32_t a_stage = KCH * kATile, w_slot = KCH * kWCh;
    extern __shared__ uint8_t smem_raw[];
    const int H = P.H, B = P.B, T = P.T, S = P.stages, WST = P.wst, kres = P.kres, cpp = P.cpp;
    const int nkc = H / 64, nst = nkc / KCH, nsst = (nkc - kres) / KCH;          // stages per chain step; the streamed-weight ones come first
    const Smem2 sm = carve2(smem_raw, cpp * kres * kWCh, WST * (int)w_slot, S * (int)a_stage, kStgAll);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t crank = tc::cluster_ctarank();
    const uint32_t rank = crank & 1u;                          // 0 = leader of its pair (issues the MMAs); this CTA owns batch tile `rank`
    const int pp = (int)(crank >> 1), mc = P.mc;               // pair index inside the cluster; pairs per cluster
    const uint16_t pair_mask = (uint16_t)(3u << (2 * pp));     // both CTAs of this pair
    const uint16_t all_mask = (uint16_t)((1u << (2 * mc)) - 1u);
    const int pair = blockIdx.x >> 1;
    const int group = pair / P.ppc, slice = pair % P.ppc;
    const int ch0 = group * cpp;                               // first chain of this pair
    const int nact = P.n_chains - ch0 < cpp ? P.n_chains - ch0 : cpp;
    const int u0 = slice * kUP;                                // first unit of the pair (in every chain it serves)
    const int b0 = (int)rank * 128;                            // first row of this CTA's batch tile
    const bool dbg_on = P.dbg != nullptr && blockIdx.x == 0 && lane == 0;

    if (warp == 0 && lane == 0) {
        for (int k = 0; k < nact; ++k) { tc::prefetch_tmap(&P.c[ch0 + k].tmW); tc::prefetch_tmap(&P.c[ch0 + k].tmA); tc::prefetch_tmap(&P.c[ch0 + k].tmS); }
        for (int i = 0; i < S; ++i) { tc::mbar_init(&sm.full[i], 1); tc::mbar_init(&sm.empty[i], (uint32_t)mc); }   // a slot is free when EVERY pair of the cluster has consumed it
        for (int i = 0; i < WST; ++i) { tc::mbar_init(&sm.wfull[i], 1); tc::mbar_init(&sm.wempty[i], 1); }
        tc::mbar_init(sm.wbar, 1);
        for (int k = 0; k < kMaxCpp; ++k) tc::mbar_init(&sm.acc_full[k], 1);
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc_2cta(sm.tmem_slot, kTmemCols);
    if (warp >= kEpi0) {
        for (int i = threadIdx.x - kEpi0 * 32; i < nact * 3 * kUP; i += kEW2 * 32) {
            const int k = i / (3 * kUP), j = i % (3 * kUP);
            sm.bias[i] = P.c[ch0 + k].b_hh[(j / kUP) * H + u0 + (j % kUP)];
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::cluster_sync();                                        // the peer's barriers exist before anything arrives on them
    tc::tc_fence_after();
    const uint32_t tmem_base = *sm.tmem_slot;

    if (warp == 0) {
        // ------------------------------- state loader (both CTAs: own batch tile) ------------------------------
        // Every load of the pair completes on the LEADER's barrier, which expects the bytes of both CTAs.  A stage is ONE
        // box of 128 rows x KCH chunks: tools/ubench_pair.cu measures ~230 cycles + bytes / 140 B/clk per TMA operation of
        // one issuing thread (16 / 32 / 64 KB boxes: 45 / 70 / 94 B/clk per CTA), and SMALLER boxes from several lanes
        // are slower still (8 KB x 4 lanes: 50 B/clk) -- so the boxes are as big as the ring allows.
        const uint32_t full0 = tc::smem_u32(sm.full), full_l = full0 & tc::kPeerBitMask, empty0 = tc::smem_u32(sm.empty);
        const uint32_t a0 = tc::smem_u32(sm.A);
        uint16_t mc_mask = 0;
        for (int q = 0; q < mc; ++q) mc_mask |= (uint16_t)(1u << (2 * q + (int)rank));
        if (kres > 0) {
            if (rank == 0 && lane == 0) tc::mbar_arrive_expect_tx(sm.wbar, (uint32_t)(2 * nact * kres * kWCh));
            __syncwarp();
            const uint32_t wbar_l = tc::smem_u32(sm.wbar) & tc::kPeerBitMask;
            for (int i = lane; i < nact * (kres / KCH); i += 32) {
                const int k = i / (kres / KCH), kc = (i % (kres / KCH)) * KCH;
                tc::tma_load_4d_2cta_u32(tc::smem_u32(sm.W) + (uint32_t)((k * kres + kc) * kWCh), &P.c[ch0 + k].tmW, wbar_l, 0, u0 + (int)rank * (kUP / 2), 0, kc);
            }
        }
        __syncwarp();
        uint32_t st = 0, ph = 1;
        for (int i = 0; i < T; ++i) {
            for (int k = 0; k < nact; ++k) {
                const Tc2Chain& c = P.c[ch0 + k];
                const int slab = c.reverse ? T - i : i;        // the state before step i
                FN_STAMP2(i, k * 32 + 0);
                if (i > 0) {
                    fn_spin_until(P.bar + (ch0 + k) * 16 + rank, (unsigned)(P.ppc * i));
#if FN_GRU2_FENCES & 4
                    asm volatile("fence.proxy.async.global;" ::: "memory");
#endif
                }
                FN_STAMP2(i, k * 32 + 1);
                int kc = kres;                                 // K order: streamed-weight chunks [kres, nkc) first, then [0, kres)
                for (int j = 0; j < nst; ++j) {
                    if (j == nsst) kc = 0;
                    tc::mbar_wait_u32(empty0 + st * 8u, ph);
                    if (tc::elect_one()) {
                        if (rank == 0) tc::mbar_arrive_expect_tx_u32(full0 + st * 8u, 2 * a_stage);
                        if (mc == 1) tc::tma_load_4d_2cta_u32(a0 + st * a_stage, &c.tmA, full_l + st * 8u, 0, b0, kc, slab);
                        else         // this CTA's share of the stage (KCH / mc chunks), delivered to the same-rank CTA of every pair of the cluster
                            tc::tma_load_4d_2cta_mc_u32(a0 + st * a_stage + (uint32_t)pp * (a_stage / (uint32_t)mc), &c.tmA, full_l + st * 8u, 0, b0,
                                                        kc + pp * (KCH / mc), slab, mc_mask);
                    }
                    __syncwarp();
                    kc += KCH;
                    if (++st == (uint32_t)S) { st = 0; ph ^= 1u; }
                }
                FN_STAMP2(i, k * 32 + 2);
            }
        }
    } else if (warp == 1) {
        // ------------------------------- MMA issuer (leader only) -----------------------------------------------
        if (rank == 0) {
            const uint32_t idesc = tc::make_idesc_bf16(256, kN, 0, 0);
            if (kres > 0) tc::mbar_wait(sm.wbar, 0);
            const uint32_t full0 = tc::smem_u32(sm.full), empty0 = tc::smem_u32(sm.empty);
            const uint32_t wfull0 = tc::smem_u32(sm.wfull), wempty0 = tc::smem_u32(sm.wempty), accf0 = tc::smem_u32(sm.acc_full);
            const uint64_t adesc0 = tc::make_sdesc(tc::smem_u32(sm.A), 16, 1024);
            const uint64_t bdesc0 = tc::make_sdesc(tc::smem_u32(sm.W), 16, 1024);
            const uint64_t wdesc0 = tc::make_sdesc(tc::smem_u32(sm.WR), 16, 1024);
            uint32_t st = 0, ph = 0, ws = 0, wph = 0;
            auto issue = [&](uint32_t d_tmem, uint64_t ad, uint64_t bd, bool first) {
#pragma unroll
                for (int q = 0; q < KCH; ++q) {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        tc::umma_f16_2cta(d_tmem, ad + (uint64_t)(q * (kATile >> 4) + 2 * kk), bd + (uint64_t)(q * (kWCh >> 4) + 2 * kk), idesc,
                                          (q | kk) ? 1u : (first ? 0u : 1u));
                }
            };
            for (int i = 0; i < T; ++i) {
                // (a chain's accumulator is free: every chunk of this step's state was published after both CTAs'
                // epilogues had read the previous accumulator -- the step dependency itself orders the reuse)
                for (int k = 0; k < nact; ++k) {
                    const uint32_t d_tmem = tmem_base + (uint32_t)(k * kN);
                    for (int j = 0; j < nsst; ++j) {
                        tc::mbar_wait_u32(wfull0 + ws * 8u, wph);
                        tc::mbar_wait_u32(full0 + st * 8u, ph);
                        tc::tc_fence_after();
                        if (j == 0) FN_STAMP2(i, k * 32 + 3);
                        if (tc::elect_one()) {
                            issue(d_tmem, adesc0 + (uint64_t)(st * (a_stage >> 4)), wdesc0 + (uint64_t)(ws * (w_slot >> 4)), j == 0);
                            tc::umma_commit_2cta_mc_u32(wempty0 + ws * 8u, pair_mask);
                            tc::umma_commit_2cta_mc_u32(empty0 + st * 8u, all_mask);
                            if (j == nst - 1) tc::umma_commit_2cta_mc_u32(accf0 + k * 8u, pair_mask);
                        }
                        __syncwarp();
                        if (++ws == (uint32_t)WST) { ws = 0; wph ^= 1u; }
                        if (++st == (uint32_t)S) { st = 0; ph ^= 1u; }
                    }
                    uint64_t bd = bdesc0 + (uint64_t)(k * kres * (kWCh >> 4));
                    for (int j = nsst; j < nst; ++j) {
                        tc::mbar_wait_u32(full0 + st * 8u, ph);
                        tc::tc_fence_after();
                        if (j == 0) FN_STAMP2(i, k * 32 + 3);
                        if (tc::elect_one()) {
                            issue(d_tmem, adesc0 + (uint64_t)(st * (a_stage >> 4)), bd, j == 0);
                            tc::umma_commit_2cta_mc_u32(empty0 + st * 8u, all_mask);
                            if (j == nst - 1) tc::umma_commit_2cta_mc_u32(accf0 + k * 8u, pair_mask);
                        }
                        __syncwarp();
                        bd += (uint64_t)(w_slot >> 4);
                        if (++st == (uint32_t)S) { st = 0; ph ^= 1u; }
                    }
                    FN_STAMP2(i, k * 32 + 4);
                }
            }
        }
    } else if (warp == 2) {
        // ------------------------------- streamed part of the weight halves (both CTAs) -------------------------
        // Independent of the recurrence: runs ahead of the step barrier, up to the ring depth.
        if (nsst > 0) {
            const uint32_t wr0 = tc::smem_u32(sm.WR), wfull0 = tc::smem_u32(sm.wfull), wempty0 = tc::smem_u32(sm.wempty);
            const uint32_t wfull_l = wfull0 & tc::kPeerBitMask;
            uint32_t ws = 0, wph = 1;
            for (int i = 0; i < T; ++i) {
                for (int k = 0; k < nact; ++k) {
                    const Tc2Chain& c = P.c[ch0 + k];
                    for (int j = 0; j < nsst; ++j) {
                        tc::mbar_wait_u32(wempty0 + ws * 8u, wph);
                        if (tc::elect_one()) {
                            if (rank == 0) tc::mbar_arrive_expect_tx_u32(wfull0 + ws * 8u, 2 * w_slot);
                            tc::tma_load_4d_2cta_u32(wr0 + ws * w_slot, &c.tmW, wfull_l + ws * 8u, 0, u0 + (int)rank * (kUP / 2), 0, kres + j * KCH);
                        }
                        __syncwarp();
                        if (++ws == (uint32_t)WST) { ws = 0; wph ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 3) {
        // ------------------------------- store + publish ---------------------------------------------------------
        for (int s = 0; s < T; ++s) {
            for (int k = 0; k < nact; ++k) {
                const Tc2Chain& c = P.c[ch0 + k];
                const int tau = c.reverse ? T - 1 - s : s;
                asm volatile("bar.sync %0, %1;" ::"r"(kBarStage + k), "n"((kEW2 + 1) * 32) : "memory");
                if (tc::elect_one()) {
                    tc::tma_store_3d_u32(&c.tmS, tc::smem_u32(sm.stg) + (uint32_t)(k * kStgF), u0, b0, c.reverse ? tau : tau + 1);
                    tc::bulk_commit_group();
                    tc::bulk_wait_group<0>();                  // the tile is in L2 (and the staging tile reusable)
                }
                __syncwarp();
                FN_STAMP2(s, k * 32 + 9);
                if (s + 1 < T && tc::elect_one()) {
#if FN_GRU2_FENCES & 2
                    asm volatile("fence.proxy.async.global;" ::: "memory");
#endif
                    publish2(P.bar + (ch0 + k) * 16 + rank);
                }
                __syncwarp();
                asm volatile("bar.arrive %0, %1;" ::"r"(kBarPub + k), "n"((kEW2 + 1) * 32) : "memory");   // the epilogue's bulk stores may go now
                FN_STAMP2(s, k * 32 + 10);
            }
        }
    } else {
        // ------------------------------- gate epilogue (16 warps) ------------------------------------------------
        constexpr int UT = kUT;
        const int q = warp & 3;                                // TMEM lane quarter
        const int grp = (warp - kEpi0) >> 2;                   // 8-unit group of the pair's 32 units
        const int uu = grp * UT, u = u0 + uu;
        const uint32_t col = (uint32_t)((grp >> 1) * kNW + (grp & 1) * UT);      // column of gate r of this group; z: + UP/2, n: + UP
        constexpr uint32_t GS = kUP / 2;                        // column stride between the gates of a half
        const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16) + col;
        const int rl = q * 32 + lane, b = b0 + rl;
        const bool row_ok = b < B;
        const uint32_t stg = tc::smem_u32(sm.stg) + (kUP == 32 ? swz64(rl, grp) : swz(rl, 2 * grp));
        const uint32_t stg_hi = tc::smem_u32(sm.stg) + swz(rl, 2 * grp + 1);      // second 16-byte chunk (UP = 64 only)
        const bool dbg_e = dbg_on && warp == kEpi0;
        float hreg[kMaxCpp][UT];
        int id_next[kMaxCpp];
#pragma unroll
        for (int k = 0; k < kMaxCpp; ++k) {
            id_next[k] = 0;
#pragma unroll
            for (int j = 0; j < UT; ++j) hreg[k][j] = 0.f;
            if (k < nact) {
                const Tc2Chain& c = P.c[ch0 + k];
                float pr[UT], pz[UT], pn[UT];
#pragma unroll
                for (int j = 0; j < UT; ++j) { pr[j] = 0.f; pz[j] = 0.f; pn[j] = 0.f; }
                if (row_ok) {
                    if (c.proj) {
                        const float* pj = c.proj + (long long)b * c.proj_ld + u;
                        ldf<UT>(pj, pr); ldf<UT>(pj + H, pz); ldf<UT>(pj + 2 * H, pn);
                    }
                    uint32_t hw[UT / 2];
                    ldb_raw<UT>(c.hsx + ((long long)(c.reverse ? T : 0) * B + b) * H + u, hw, false);
                    unpack<UT>(hw, hreg[k]);
                    if (c.emb) id_next[k] = c.ids[(long long)(c.reverse ? T - 1 : 0) * B + b];
                }
                const float* bs = sm.bias + k * 3 * kUP + uu;
#pragma unroll
                for (int j = 0; j < UT; ++j) { pr[j] += bs[j]; pz[j] += bs[kUP + j]; }
                const uint32_t tp = t_lane + kAccCols + (uint32_t)(k * kN);
                tmem_st<UT>(tp, pr); tmem_st<UT>(tp + GS, pz); tmem_st<UT>(tp + 2 * GS, pn);
            }
        }
        tmem_st_wait();
        for (int s = 0; s < T; ++s) {
#pragma unroll
            for (int k = 0; k < kMaxCpp; ++k) {
                if (k >= nact) break;
                const Tc2Chain& c = P.c[ch0 + k];
                const int tau = c.reverse ? T - 1 - s : s;
                const int tau_n = c.reverse ? tau - 1 : tau + 1;
                // ---- operand that does not depend on the recurrence: fetch before waiting for the MMAs
                uint32_t ir[UT / 2], iz[UT / 2], in_[UT / 2];
#pragma unroll
                for (int j = 0; j < UT / 2; ++j) { ir[j] = 0u; iz[j] = 0u; in_[j] = 0u; }
                if (row_ok && (c.emb || c.dense)) {
                    const __nv_bfloat16* src = c.emb ? c.emb + (long long)id_next[k] * 3 * H + u : c.dense + ((long long)tau * B + b) * 3 * H + u;
                    ldb_raw<UT>(src, ir, false); ldb_raw<UT>(src + H, iz, false); ldb_raw<UT>(src + 2 * H, in_, false);
                    if (c.emb && s + 1 < T) id_next[k] = c.ids[(long long)tau_n * B + b];
                }
                if (dbg_e) P.dbg[(long long)s * 64 + k * 32 + 5] = clk();
                tc::mbar_wait_warp(&sm.acc_full[k], s & 1);
                tc::tc_fence_after();
                if (dbg_e) P.dbg[(long long)s * 64 + k * 32 + 6] = clk();
                const uint32_t ta = t_lane + (uint32_t)(k * kN), tp = ta + kAccCols;
                float r[UT], z[UT], n[UT], g[UT];
                {
                    float a[UT], p[UT], x[UT];
                    tmem_ld2<UT>(ta, a, tp, p);
                    unpack<UT>(ir, x);
#pragma unroll
                    for (int j = 0; j < UT; j += 2) sigmoid2(a[j] + p[j] + x[j], a[j + 1] + p[j + 1] + x[j + 1], r[j], r[j + 1]);
                    tmem_ld2<UT>(ta + GS, a, tp + GS, p);
                    unpack<UT>(iz, x);
#pragma unroll
                    for (int j = 0; j < UT; j += 2) sigmoid2(a[j] + p[j] + x[j], a[j + 1] + p[j + 1] + x[j + 1], z[j], z[j + 1]);
                    tmem_ld2<UT>(ta + 2 * GS, a, tp + 2 * GS, p);
                    tc::tc_fence_before();
                    unpack<UT>(in_, x);
                    const float* bn = sm.bias + k * 3 * kUP + 2 * kUP + uu;
#pragma unroll
                    for (int j = 0; j < UT; ++j) {
                        g[j] = a[j] + bn[j];
                        n[j] = fast_tanh(p[j] + x[j] + r[j] * g[j]);
                        hreg[k][j] = (1.f - z[j]) * n[j] + z[j] * hreg[k][j];
                    }
                }
                // new state -> swizzled staging tile -> (store warp) one TMA store + one publish per CTA and chain
                sts16(stg + (uint32_t)(k * kStgF), pack2(hreg[k][0], hreg[k][1]), pack2(hreg[k][2], hreg[k][3]), pack2(hreg[k][4], hreg[k][5]),
                      pack2(hreg[k][6], hreg[k][7]));
                if constexpr (UT == 16)
                    sts16(stg_hi, pack2(hreg[k][8], hreg[k][9]), pack2(hreg[k][10], hreg[k][11]), pack2(hreg[k][12], hreg[k][13]),
                          pack2(hreg[k][14], hreg[k][15]));
                tc::fence_proxy_async();
                if (P.dbg != nullptr && blockIdx.x == 0 && lane == 0) P.dbg[(long long)s * 64 + k * 32 + 12 + (warp - kEpi0)] = clk();
                asm volatile("bar.arrive %0, %1;" ::"r"(kBarStage + k), "n"((kEW2 + 1) * 32) : "memory");
                // The saved gates (and the next step's input gather) are off the critical path, but they share the SM's
                // load/store pipe with the hand-over: issued now, 64 store instructions per CTA queue ahead of the store
                // warp's counter update and the loader's polls (measured: +1.5k cycles per step).  Wait for the publish.
                asm volatile("bar.sync %0, %1;" ::"r"(kBarPub + k), "n"((kEW2 + 1) * 32) : "memory");
                if (row_ok) {
                    if (c.gates) {                             // off the critical path
                        stb<UT>(c.gates + gate_off(tau, b, u, B, 4 * H), r);
                        stb<UT>(c.gates + gate_off(tau, b, H + u, B, 4 * H), z);
                        stb<UT>(c.gates + gate_off(tau, b, 2 * H + u, B, 4 * H), n);
                        stb<UT>(c.gates + gate_off(tau, b, 3 * H + u, B, 4 * H), g);
                    }
                    if (s == T - 1 && c.h_final) {
                        float* hf = c.h_final + (long long)b * c.h_final_ld + u;
#pragma unroll
                        for (int j = 0; j < UT; ++j) hf[j] = hreg[k][j];
                    }
                }
                if (dbg_e) P.dbg[(long long)s * 64 + k * 32 + 11] = clk();
            }
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

// =====================================================================================================
// Backward (BPTT).  Iteration i = 0..T handles step s = T-1-i (s = -1 finishes dh0).  A pair owns 64 units: the product
//   dh_prev[256 rows x 64 units] = dg3[256 x 3H] * W_hh[3H x 64 units],  dg3 = (dr, dz, dn*r) of the step above,
// is ONE accumulator of 64 columns (M = 256, N = 64, K = 3H); CTA r holds 32 rows of W_hh^T (4 KB per 64-wide K chunk,
// resident for its first `kres` chunks, the rest re-streamed) and streams the gate gradients of batch tile r
// (768 KB per step at H = 1024 -- shared with the same-rank CTAs of its cluster by multicast).  The epilogue turns the
// accumulator + carry + incoming gradients into (dr, dz, dn, dn*r) of its (row, 16 units), staged as four 128B-swizzled
// [128 x 64] tiles -- in the state ring itself, which is idle between a step's last MMA and the hand-over -- and written
// by ONE 4-D TMA store (64 KB) into the dg stream, followed by one publish per CTA.
// =====================================================================================================
constexpr int kUPB = 64, kNB = 64, kWChB = 32 * 128;          // units per pair, MMA N, bytes of one K chunk of a CTA's 32 weight rows

template <int KCH>
__global__ void __launch_bounds__(kThreads2, 1) gru2_bwd_kernel(const __grid_constant__ Tc2Launch P) {
    constexpr uint32_t kTmemCols = 64;
    constexpr uint32_t a_stage = KCH * kATile, w_slot = KCH * kWChB;
    extern __shared__ uint8_t smem_raw[];
    const int H = P.H, B = P.B, T = P.T, S = P.stages, WST = P.wst, kres = P.kres;
    const int nkc = 3 * H / 64, nst = nkc / KCH, nsst = (nkc - kres) / KCH;
    const int gap0 = 2 * H / 64, gap = H / 64;                 // K chunk kc >= gap0 reads dg column chunk kc + gap (skips dn)
    const Smem2 sm = carve2(smem_raw, kres * kWChB, WST * (int)w_slot, S * (int)a_stage, 0);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t crank = tc::cluster_ctarank();
    const uint32_t rank = crank & 1u;
    const int pp = (int)(crank >> 1), mc = P.mc;
    const uint16_t pair_mask = (uint16_t)(3u << (2 * pp));
    const uint16_t all_mask = (uint16_t)((1u << (2 * mc)) - 1u);
    const int pair = blockIdx.x >> 1;
    const int chain = pair / P.ppc, slice = pair % P.ppc;
    const Tc2Chain& c = P.c[chain];
    unsigned* gflag = P.bar + chain * 16 + rank;
    const int u0 = slice * kUPB, b0 = (int)rank * 128;

    if (warp == 0 && lane == 0) {
        tc::prefetch_tmap(&c.tmW); tc::prefetch_tmap(&c.tmA); tc::prefetch_tmap(&c.tmS);
        for (int i = 0; i < S; ++i) { tc::mbar_init(&sm.full[i], 1); tc::mbar_init(&sm.empty[i], (uint32_t)mc); }
        for (int i = 0; i < WST; ++i) { tc::mbar_init(&sm.wfull[i], 1); tc::mbar_init(&sm.wempty[i], 1); }
        tc::mbar_init(sm.wbar, 1); tc::mbar_init(&sm.acc_full[0], 1);
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc_2cta(sm.tmem_slot, kTmemCols);
    tc::tc_fence_before();
    __syncthreads();
    tc::cluster_sync();
    tc::tc_fence_after();
    const uint32_t tmem_base = *sm.tmem_slot;

    if (warp == 0) {
        // ------------------------------- gate-gradient loader (both CTAs: own batch tile) -----------------------
        const uint32_t full0 = tc::smem_u32(sm.full), full_l = full0 & tc::kPeerBitMask, empty0 = tc::smem_u32(sm.empty);
        const uint32_t a0 = tc::smem_u32(sm.A);
        uint16_t mc_mask = 0;
        for (int q = 0; q < mc; ++q) mc_mask |= (uint16_t)(1u << (2 * q + (int)rank));
        if (kres > 0) {
            if (rank == 0 && lane == 0) tc::mbar_arrive_expect_tx(sm.wbar, (uint32_t)(2 * kres * kWChB));
            __syncwarp();
            const uint32_t wbar_l = tc::smem_u32(sm.wbar) & tc::kPeerBitMask;
            for (int i = lane; i < kres / KCH; i += 32)
                tc::tma_load_4d_2cta_u32(tc::smem_u32(sm.W) + (uint32_t)(i * KCH * kWChB), &c.tmW, wbar_l, 0, u0 + (int)rank * 32, i * KCH, 0);
        }
        __syncwarp();
        uint32_t st = 0, ph = 1;
        for (int i = 1; i <= T; ++i) {
            const int slab = c.reverse ? i - 1 : T - i;        // the gate gradient of step s+1 (s = T-1-i), by time
            fn_spin_until(gflag, (unsigned)(P.ppc * i));
#if FN_GRU2_FENCES & 4
            asm volatile("fence.proxy.async.global;" ::: "memory");
#endif
            int kc = kres;
            for (int j = 0; j < nst; ++j) {
                if (j == nsst) kc = 0;
                const int cc = (kc >= gap0 ? kc + gap : kc) + pp * (KCH / mc);
                tc::mbar_wait_u32(empty0 + st * 8u, ph);
                if (tc::elect_one()) {
                    if (rank == 0) tc::mbar_arrive_expect_tx_u32(full0 + st * 8u, 2 * a_stage);
                    if (mc == 1) tc::tma_load_4d_2cta_u32(a0 + st * a_stage, &c.tmA, full_l + st * 8u, 0, b0, cc, slab);
                    else tc::tma_load_4d_2cta_mc_u32(a0 + st * a_stage + (uint32_t)pp * (a_stage / (uint32_t)mc), &c.tmA, full_l + st * 8u, 0, b0, cc,
                                                     slab, mc_mask);
                }
                __syncwarp();
                kc += KCH;
                if (++st == (uint32_t)S) { st = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ------------------------------- MMA issuer (leader only) -----------------------------------------------
        if (rank == 0) {
            const uint32_t idesc = tc::make_idesc_bf16(256, kNB, 0, 0);
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
                    for (int kk = 0; kk < 4; ++kk)
                        tc::umma_f16_2cta(tmem_base, ad + (uint64_t)(q * (kATile >> 4) + 2 * kk), bd + (uint64_t)(q * (kWChB >> 4) + 2 * kk), idesc,
                                          (q | kk) ? 1u : (first ? 0u : 1u));
                }
            };
            for (int i = 1; i <= T; ++i) {
                for (int j = 0; j < nsst; ++j) {
                    tc::mbar_wait_u32(wfull0 + ws * 8u, wph);
                    tc::mbar_wait_u32(full0 + st * 8u, ph);
                    tc::tc_fence_after();
                    if (tc::elect_one()) {
                        issue(adesc0 + (uint64_t)(st * (a_stage >> 4)), wdesc0 + (uint64_t)(ws * (w_slot >> 4)), j == 0);
                        tc::umma_commit_2cta_mc_u32(wempty0 + ws * 8u, pair_mask);
                        tc::umma_commit_2cta_mc_u32(empty0 + st * 8u, all_mask);
                        if (j == nst - 1) tc::umma_commit_2cta_mc_u32(accf, pair_mask);
                    }
                    __syncwarp();
                    if (++ws == (uint32_t)WST) { ws = 0; wph ^= 1u; }
                    if (++st == (uint32_t)S) { st = 0; ph ^= 1u; }
                }
                uint64_t bd = bdesc0;
                for (int j = nsst; j < nst; ++j) {
                    tc::mbar_wait_u32(full0 + st * 8u, ph);
                    tc::tc_fence_after();
                    if (tc::elect_one()) {
                        issue(adesc0 + (uint64_t)(st * (a_stage >> 4)), bd, j == 0);
                        tc::umma_commit_2cta_mc_u32(empty0 + st * 8u, all_mask);
                        if (j == nst - 1) tc::umma_commit_2cta_mc_u32(accf, pair_mask);
                    }
                    __syncwarp();
                    bd += (uint64_t)(w_slot >> 4);
                    if (++st == (uint32_t)S) { st = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 2) {
        // ------------------------------- streamed part of the weight half (both CTAs) ---------------------------
        if (nsst > 0) {
            const uint32_t wr0 = tc::smem_u32(sm.WR), wfull0 = tc::smem_u32(sm.wfull), wempty0 = tc::smem_u32(sm.wempty);
            const uint32_t wfull_l = wfull0 & tc::kPeerBitMask;
            uint32_t ws = 0, wph = 1;
            for (int i = 1; i <= T; ++i) {
                for (int j = 0; j < nsst; ++j) {
                    tc::mbar_wait_u32(wempty0 + ws * 8u, wph);
                    if (tc::elect_one()) {
                        if (rank == 0) tc::mbar_arrive_expect_tx_u32(wfull0 + ws * 8u, 2 * w_slot);
                        tc::tma_load_4d_2cta_u32(wr0 + ws * w_slot, &c.tmW, wfull_l + ws * 8u, 0, u0 + (int)rank * 32, kres + j * KCH, 0);
                    }
                    __syncwarp();
                    if (++ws == (uint32_t)WST) { ws = 0; wph ^= 1u; }
                }
            }
        }
    } else if (warp == 3) {
        // ------------------------------- store + publish (iterations 0 .. T-1 produce a gate-gradient tile) ------
        for (int i = 0; i < T; ++i) {
            const int s = T - 1 - i;
            const int tau = c.reverse ? T - 1 - s : s;
            asm volatile("bar.sync %0, %1;" ::"n"(kBarStage), "n"((kEW2 + 1) * 32) : "memory");
            if (tc::elect_one()) {
                tc::tma_store_4d_u32(&c.tmS, tc::smem_u32(sm.A), u0, b0, 0, tau);
                tc::bulk_commit_group();
                tc::bulk_wait_group<0>();
#if FN_GRU2_FENCES & 2
                asm volatile("fence.proxy.async.global;" ::: "memory");
#endif
                publish2(gflag);
            }
            __syncwarp();
            asm volatile("bar.arrive %0, %1;" ::"n"(kBarPub), "n"((kEW2 + 1) * 32) : "memory");
        }
    } else {
        // ------------------------------- gate-gradient epilogue (16 warps) --------------------------------------
        constexpr int UT = 16;
        const int q = warp & 3, grp = (warp - kEpi0) >> 2;
        const int uu = grp * UT, u = u0 + uu;
        const uint32_t t_acc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)uu;
        const int rl = q * 32 + lane, b = b0 + rl;
        const bool row_ok = b < B;
        const uint32_t stg0 = tc::smem_u32(sm.A) + swz(rl, 2 * grp), stg1 = tc::smem_u32(sm.A) + swz(rl, 2 * grp + 1);
        float carry[UT];
#pragma unroll
        for (int j = 0; j < UT; ++j) carry[j] = 0.f;
        for (int i = 0; i <= T; ++i) {
            const int s = T - 1 - i;
            const int tau = c.reverse ? T - 1 - s : s;
            const long long row = (long long)tau * B + b;
            // ---- saved forward values and incoming gradients: fetch before waiting for the MMAs -- but after the previous
            // iteration's hand-over (these 6 x 16 warps of loads would queue ahead of the publish and the loader's polls)
            if (i > 0) asm volatile("bar.sync %0, %1;" ::"n"(kBarPub), "n"((kEW2 + 1) * 32) : "memory");
            uint32_t wr[UT / 2], wz[UT / 2], wn[UT / 2], wg[UT / 2], wh[UT / 2];
            float din[UT];
#pragma unroll
            for (int j = 0; j < UT / 2; ++j) { wr[j] = 0; wz[j] = 0; wn[j] = 0; wg[j] = 0; wh[j] = 0; }
#pragma unroll
            for (int j = 0; j < UT; ++j) din[j] = 0.f;
            if (row_ok && s >= 0) {
                ldb_raw<UT>(c.gates + gate_off(tau, b, u, B, 4 * H), wr, false);
                ldb_raw<UT>(c.gates + gate_off(tau, b, H + u, B, 4 * H), wz, false);
                ldb_raw<UT>(c.gates + gate_off(tau, b, 2 * H + u, B, 4 * H), wn, false);
                ldb_raw<UT>(c.gates + gate_off(tau, b, 3 * H + u, B, 4 * H), wg, false);
                ldb_raw<UT>(c.hsx + (row + (c.reverse ? B : 0)) * H + u, wh, false);       // the state before step s
                if (c.dhs) {
                    if (c.dhs_f32) ldf<UT>(reinterpret_cast<const float*>(c.dhs) + row * H + u, din);
                    else {
                        uint32_t wd[UT / 2];
                        ldb_raw<UT>(reinterpret_cast<const __nv_bfloat16*>(c.dhs) + row * H + u, wd, false);
                        unpack<UT>(wd, din);
                    }
                }
                if (s == T - 1 && c.dh_final) {
                    const float* df = c.dh_final + (long long)b * c.dh_final_ld + u;
#pragma unroll
                    for (int j = 0; j < UT; ++j) din[j] += __ldg(df + j);
                }
            }
            float dh[UT];
            if (i > 0) {
                tc::mbar_wait_warp(&sm.acc_full[0], (i - 1) & 1);
                tc::tc_fence_after();
                tmem_ld<UT>(t_acc, dh);
                tc::tc_fence_before();
            } else {
#pragma unroll
                for (int j = 0; j < UT; ++j) dh[j] = 0.f;
            }
#pragma unroll
            for (int j = 0; j < UT; ++j) dh[j] += carry[j] + din[j];
            if (s < 0) {
                if (row_ok) stf<UT>(c.dh0 + (long long)b * H + u, dh);
                break;
            }
            // (dr, dz, dn, dn*r) of this thread's 16 units, gate by gate straight into the staging tiles
            uint32_t o[UT / 2];
#define FN_STAGE_GATE(gidx)                                                                 \
            sts16(stg0 + (gidx) * kATile, o[0], o[1], o[2], o[3]);                          \
            sts16(stg1 + (gidx) * kATile, o[4], o[5], o[6], o[7]);
#pragma unroll
            for (int j = 0; j < UT; j += 2) {                  // dz = dh (h_prev - n) z (1 - z);   carry = dh z
                const float z0 = bf_lo(wz[j >> 1]), z1 = bf_hi(wz[j >> 1]), n0 = bf_lo(wn[j >> 1]), n1 = bf_hi(wn[j >> 1]);
                const float h0 = bf_lo(wh[j >> 1]), h1 = bf_hi(wh[j >> 1]);
                o[j >> 1] = pack2(dh[j] * (h0 - n0) * z0 * (1.f - z0), dh[j + 1] * (h1 - n1) * z1 * (1.f - z1));
                carry[j] = dh[j] * z0; carry[j + 1] = dh[j + 1] * z1;
                dh[j] = dh[j] * (1.f - z0) * (1.f - n0 * n0);  // dn_pre
                dh[j + 1] = dh[j + 1] * (1.f - z1) * (1.f - n1 * n1);
            }
            FN_STAGE_GATE(1)
#pragma unroll
            for (int j = 0; j < UT; j += 2) o[j >> 1] = pack2(dh[j], dh[j + 1]);                       // dn_pre
            FN_STAGE_GATE(2)
#pragma unroll
            for (int j = 0; j < UT; j += 2) {                  // dn_pre * r
                const float r0 = bf_lo(wr[j >> 1]), r1 = bf_hi(wr[j >> 1]);
                o[j >> 1] = pack2(dh[j] * r0, dh[j + 1] * r1);
            }
            FN_STAGE_GATE(3)
#pragma unroll
            for (int j = 0; j < UT; j += 2) {                  // dr = dn_pre g r (1 - r)
                const float r0 = bf_lo(wr[j >> 1]), r1 = bf_hi(wr[j >> 1]), g0 = bf_lo(wg[j >> 1]), g1 = bf_hi(wg[j >> 1]);
                o[j >> 1] = pack2(dh[j] * g0 * r0 * (1.f - r0), dh[j + 1] * g1 * r1 * (1.f - r1));
            }
            FN_STAGE_GATE(0)
#undef FN_STAGE_GATE
            tc::fence_proxy_async();
            asm volatile("bar.arrive %0, %1;" ::"n"(kBarStage), "n"((kEW2 + 1) * 32) : "memory");
        }
        tc::tc_fence_before();
    }
    __syncthreads();
    tc::cluster_sync();
    if (warp == 1) {
        tc::tc_fence_after();
        tc::tmem_dealloc_2cta(tmem_base, kTmemCols);
    }
}

// ---- host -------------------------------------------------------------------------------------------
struct Plan2 { int kch, stages, kres, wst; size_t smem; bool ok; };

// Shared-memory plan: `cpp` chains per pair, nkc K chunks of `w_chunk` bytes per chain and CTA, `stg` staging bytes.
Plan2 plan2(int nkc, int w_chunk, int cpp, int stg, const char* tag, int kch_dflt = 4, int s_dflt = 2, int wst_dflt = 2) {
    Plan2 pl{};
    const bool is_bwd = tag[0] == 'b';
    static const int kch_f = env_int("FN_GRU2_KCH", 0), s_f = env_int("FN_GRU2_S", 0), wst_f = env_int("FN_GRU2_WST", 0);
    static const int kch_b = env_int("FN_GRU2_KCH_BWD", 0), s_b = env_int("FN_GRU2_S_BWD", 0), wst_b = env_int("FN_GRU2_WST_BWD", 0);
    const int kch_e = is_bwd ? kch_b : kch_f, s_e = is_bwd ? s_b : s_f, wst_e = is_bwd ? wst_b : wst_f;
    const int kch_env = kch_e ? kch_e : kch_dflt, s_env = s_e ? s_e : s_dflt, wst_env = wst_e ? wst_e : wst_dflt;
    int kch = kch_env == 1 || kch_env == 2 || kch_env == 4 ? kch_env : 2;
    while (kch > 1 && nkc % kch) kch >>= 1;
    pl.kch = kch;
    const long long budget = (long long)fn_max_smem_optin() - (long long)kSmemTail2 - stg;
    const long long a_stage = (long long)kch * kATile, w_slot = (long long)kch * w_chunk;
    long long room = budget - (long long)cpp * nkc * w_chunk;
    if (room >= 3 * a_stage) {                                  // the whole half-slices stay resident
        pl.kres = nkc; pl.wst = 0;
        pl.stages = (int)(room / a_stage);
    } else {
        int S = s_env < 2 ? 2 : s_env, wst = wst_env < 2 ? 2 : wst_env;
        if (wst > kMaxWst) wst = kMaxWst;
        long long kres = (budget - S * a_stage - wst * w_slot) / ((long long)cpp * w_chunk);
        if (kres > nkc - kch) kres = nkc - kch;
        kres -= kres % kch;
        if (kres < 0) { pl.ok = false; return pl; }
        pl.kres = (int)kres; pl.wst = wst; pl.stages = S;
    }
    if (pl.stages > kMaxStages) pl.stages = kMaxStages;
    pl.ok = pl.stages >= 2;
    pl.smem = (size_t)((long long)cpp * pl.kres * w_chunk + pl.wst * w_slot + pl.stages * a_stage) + stg + kSmemTail2;
    static const int verbose = env_int("FN_GRU_VERBOSE", 0);
    if (verbose && pl.ok)
        fprintf(stderr, "plan2 %s nkc=%d cpp=%d: kch=%d stages=%d kres=%d wst=%d smem=%zu\n", tag, nkc, cpp, pl.kch, pl.stages, pl.kres, pl.wst, pl.smem);
    return pl;
}

int make_tmap_bf16_nd(CUtensorMap* out, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                      const cuuint32_t* box, CUtensorMapSwizzle swz) {
    fn_PFN_encodeTiled enc = fn_get_encode_tiled();
    FN_REQUIRE(enc, "cuTensorMapEncodeTiled entry point unavailable");
    FN_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA operand alignment");
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), dims, strides_bytes, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FN_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(%d-d) failed (%d)", rank, (int)r);
    return FN_OK;
}

template <typename K>
int launch2(K kernel, const Tc2Launch& P, int pairs, size_t smem, cudaStream_t st, int* fits = nullptr) {
    const void* fn = (const void*)kernel;
    FN_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    const int ctas = pairs * 2;
    cfg.gridDim = dim3(ctas);
    cfg.blockDim = dim3(kThreads2);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attrs[2];
    attrs[0].id = cudaLaunchAttributeClusterDimension;
    attrs[0].val.clusterDim.x = 2 * P.mc; attrs[0].val.clusterDim.y = 1; attrs[0].val.clusterDim.z = 1;
    attrs[1].id = cudaLaunchAttributeCooperative;           // every CTA must be co-resident: they wait on each other
    attrs[1].val.cooperative = 1;
    cfg.attrs = attrs;
    static const int coop = env_int("FN_GRU2_COOP", 1);     // 0: no cooperative attribute (profilers that cannot replay cooperative cluster launches)
    cfg.numAttrs = coop ? 2 : 1;
    int max_clusters = 0;
    FN_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, fn, &cfg));
    if (fits) { *fits = max_clusters * P.mc >= pairs; if (!*fits) return FN_OK; }
    FN_REQUIRE(max_clusters * P.mc >= pairs, "fn_gru_seq_bf16: %d CTA pairs are not co-resident (max %d)", pairs, max_clusters);
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

// Geometry of a launch of n chains of hidden size H: UP = 32 with one chain per pair if every chain can have its own H/32
// pairs; else UP = 64 (one chain per pair, H/64 pairs each) or UP = 32 with two chains per pair (FN_GRU2_GEO = 64 | 32).
struct Geom { int up, cpp; };
Geom pick_geo(int n, int H) {
    static const int force = env_int("FN_GRU2_GEO", 32);
    const int max_pairs = fn_num_sms() / 2;
    if (n * (H / 32) <= max_pairs) return {32, 1};
    if (force == 64 && n * (H / 64) <= max_pairs) return {64, 1};
    return {32, 2};
}
template <int UP>
int launch_fwd(int kch, const Tc2Launch& P, int pairs, size_t smem, cudaStream_t st, int* fits) {
    if (kch == 4) return launch2(gru2_fwd_kernel<UP, 4>, P, pairs, smem, st, fits);
    if (kch == 2) return launch2(gru2_fwd_kernel<UP, 2>, P, pairs, smem, st, fits);
    return launch2(gru2_fwd_kernel<UP, 1>, P, pairs, smem, st, fits);
}

int gru2_run_bwd(const FnGruChainBf16* chains, int n_chains, int B, int T, int H, void* barrier_ws, size_t ws_bytes, cudaStream_t st) {
    FN_REQUIRE(chains && n_chains > 0 && barrier_ws && ws_bytes >= (size_t)64 * n_chains, "fn_gru_seq_bf16: bad arguments");
    FN_CHECK_CUDA(cudaMemsetAsync(barrier_ws, 0, (size_t)64 * n_chains, st));
    const int nkc = 3 * H / 64, ppc = H / kUPB, max_pairs = fn_num_sms() / 2;
    const Plan2 pl = plan2(nkc, kWChB, 1, 0, "bwd", 4, 3, 2);
    FN_REQUIRE(pl.ok && pl.stages * pl.kch >= 4, "fn_gru_seq_bwd_bf16: no shared-memory plan for H=%d", H);   // (the ring doubles as the 64 KB staging area)
    static const int mc_env = env_int("FN_GRU2_MC_BWD", 1);
    int done = 0;
    while (done < n_chains) {
        int group = n_chains - done < kMaxChainsTc ? n_chains - done : kMaxChainsTc;
        while (group * ppc > max_pairs) --group;
        int mc = mc_env == 4 || mc_env == 2 ? mc_env : 1;
        while (mc > 1 && (pl.kch % mc || ppc % mc)) mc >>= 1;
        Tc2Launch P;
        memset(&P, 0, sizeof(P));
        auto encode_a = [&](int i, int mcv) {
            const cuuint64_t dims[4] = {64, (cuuint64_t)B, (cuuint64_t)(4 * H / 64), (cuuint64_t)T};
            const cuuint64_t str[3] = {(cuuint64_t)4 * H * 2, 128, (cuuint64_t)B * 4 * H * 2};
            const cuuint32_t box[4] = {64, 128, (cuuint32_t)(pl.kch / mcv), 1};
            return make_tmap_bf16_nd(&P.c[i].tmA, chains[done + i].dg, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
        };
        for (int i = 0; i < group; ++i) {
            const FnGruChainBf16& s = chains[done + i];
            Tc2Chain& d = P.c[i];
            FN_REQUIRE(s.hsx && s.w_hh_t && s.gates && s.dg && s.dh0, "fn_gru_seq_bwd_bf16: chain %d misses buffers", done + i);
            int rc;
            {   // W_hh^T [H][3H] as (k in chunk, unit, K chunk, -): one box = this CTA's 32 rows of KCH chunks
                const cuuint64_t dims[4] = {64, (cuuint64_t)H, (cuuint64_t)nkc, 1};
                const cuuint64_t str[3] = {(cuuint64_t)3 * H * 2, 128, (cuuint64_t)3 * H * H * 2};
                const cuuint32_t box[4] = {64, 32, (cuuint32_t)pl.kch, 1};
                if ((rc = make_tmap_bf16_nd(&d.tmW, s.w_hh_t, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
            }
            if ((rc = encode_a(i, mc))) return rc;
            {   // store of (dr, dz, dn, dn*r) of 64 units x 128 rows: (unit, row, gate, slab), one 64 KB box
                const cuuint64_t dims[4] = {(cuuint64_t)H, (cuuint64_t)B, 4, (cuuint64_t)T};
                const cuuint64_t str[3] = {(cuuint64_t)4 * H * 2, (cuuint64_t)H * 2, (cuuint64_t)B * 4 * H * 2};
                const cuuint32_t box[4] = {64, 128, 4, 1};
                if ((rc = make_tmap_bf16_nd(&d.tmS, s.dg, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
            }
            d.hsx = (__nv_bfloat16*)s.hsx; d.gates = (__nv_bfloat16*)s.gates;
            d.dhs = s.dhs; d.dh_final = s.dh_final; d.dh_final_ld = s.dh_final_ld;
            d.dg = (__nv_bfloat16*)s.dg; d.dh0 = s.dh0;
            d.reverse = s.reverse; d.dhs_f32 = s.dhs_f32;
        }
        P.bar = reinterpret_cast<unsigned*>(barrier_ws) + done * 16;
        P.n_chains = group; P.ppc = ppc; P.cpp = 1; P.B = B; P.T = T; P.H = H;
        P.stages = pl.stages; P.kres = pl.kres; P.wst = pl.wst;
        P.dbg = nullptr;
        P.mc = mc;
        const int pairs = group * ppc;
        int rc = FN_OK, fits = 0;
        for (;;) {
            rc = pl.kch == 4 ? launch2(gru2_bwd_kernel<4>, P, pairs, pl.smem, st, &fits)
               : pl.kch == 2 ? launch2(gru2_bwd_kernel<2>, P, pairs, pl.smem, st, &fits) : launch2(gru2_bwd_kernel<1>, P, pairs, pl.smem, st, &fits);
            if (rc != FN_OK || fits || P.mc == 1) break;
            P.mc >>= 1;
            for (int i = 0; i < group; ++i)
                if ((rc = encode_a(i, P.mc))) return rc;
        }
        if (rc != FN_OK) return rc;
        FN_REQUIRE(fits, "fn_gru_seq_bwd_bf16: %d CTA pairs are not co-resident", pairs);
        done += group;
    }
    return FN_OK;
}

}  // namespace

bool fn_gru2_eligible(bool bwd, int n_chains, int B, int H) {
    static const int on = env_int("FN_GRU_V2", 3);                 // bit 0: forward, bit 1: backward
    if (!(on & (bwd ? 2 : 1))) return false;
    if (B <= 128 || B > 256 || H % 64 || H < 128) return false;
    if (H / 64 > fn_num_sms() / 2) return false;               // one chain must fit the machine
    if (bwd) return plan2(3 * H / 64, kWChB, 1, 0, "bwd", 4, 3, 2).ok;
    return plan2(H / 64, Geo<32>::kWCh, 2, kStgAll, "fwd").ok;
}

int fn_gru2_run(bool bwd, const FnGruChainBf16* chains, int n_chains, int B, int T, int H, void* barrier_ws, size_t ws_bytes,
                cudaStream_t st) {
    if (bwd) return gru2_run_bwd(chains, n_chains, B, T, H, barrier_ws, ws_bytes, st);
    FN_REQUIRE(chains && n_chains > 0 && barrier_ws && ws_bytes >= (size_t)64 * n_chains, "fn_gru_seq_bf16: bad arguments");
    FN_CHECK_CUDA(cudaMemsetAsync(barrier_ws, 0, (size_t)64 * n_chains, st));
    const int nkc = H / 64, max_pairs = fn_num_sms() / 2;
    int done = 0;
    while (done < n_chains) {
        int group = n_chains - done < kMaxChainsTc ? n_chains - done : kMaxChainsTc;
        Geom g = pick_geo(group, H);
        while (((group + g.cpp - 1) / g.cpp) * (H / g.up) > max_pairs) { --group; g = pick_geo(group, H); }   // (group >= 1 fits: fn_gru2_eligible)
        const int ppc = H / g.up, cpp = g.cpp, w_chunk = g.up == 64 ? Geo<64>::kWCh : Geo<32>::kWCh;
        // ring defaults from the config-3 sweeps: N = 96 wants 64 KB state boxes (everything streamed), N = 192 32 KB boxes
        const Plan2 pl = g.up == 64 ? plan2(nkc, w_chunk, cpp, kStgAll, "fwd64", 2, 3, 2) : plan2(nkc, w_chunk, cpp, kStgAll, "fwd32", 4, 2, 2);
        FN_REQUIRE(pl.ok, "fn_gru_seq_bf16: no shared-memory plan for H=%d", H);
        // Pairs per cluster: the same-rank CTAs of a cluster need the same state slab, so each loads 1/mc of every stage and
        // multicasts it -- L2 reads of the state (the bulk of the step's traffic; the kernels are bound by the ~6 kB/clk the L2
        // delivers to the SMs) drop by mc.  Largest of 4, 2, 1 that divides the stage and the pairs of a chain and is co-resident.
        static const int mc_env = env_int("FN_GRU2_MC", 4);
        int mc = mc_env == 4 || mc_env == 2 ? mc_env : 1;
        while (mc > 1 && (pl.kch % mc || ppc % mc)) mc >>= 1;
        Tc2Launch P;
        memset(&P, 0, sizeof(P));
        for (int i = 0; i < group; ++i) {
            const FnGruChainBf16& s = chains[done + i];
            Tc2Chain& d = P.c[i];
            FN_REQUIRE(s.hsx && s.w_hh && s.b_hh, "fn_gru_seq_fwd_bf16: chain %d misses buffers", done + i);
            FN_REQUIRE(!s.emb || s.ids, "fn_gru_seq_fwd_bf16: chain %d has emb without ids", done + i);
            FN_REQUIRE(!(s.emb && s.dense), "fn_gru_seq_fwd_bf16: chain %d has both a token and a dense input", done + i);
            int rc;
            {   // W_hh [3H][H] as (k in chunk, unit, gate, K chunk): one box = this CTA's 3 x 16 rows of KCH chunks (KCH x 6 KB)
                const cuuint64_t dims[4] = {64, (cuuint64_t)H, 3, (cuuint64_t)nkc};
                const cuuint64_t str[3] = {(cuuint64_t)H * 2, (cuuint64_t)H * H * 2, 128};
                const cuuint32_t box[4] = {64, (cuuint32_t)g.up / 2, 3, (cuuint32_t)pl.kch};
                if ((rc = make_tmap_bf16_nd(&d.tmW, s.w_hh, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
            }
            {   // hsx [T+1][B][H] as (k in chunk, row, K chunk, slab): one box = 128 rows of KCH chunks (KCH x 16 KB)
                const cuuint64_t dims[4] = {64, (cuuint64_t)B, (cuuint64_t)nkc, (cuuint64_t)T + 1};
                const cuuint64_t str[3] = {(cuuint64_t)H * 2, 128, (cuuint64_t)B * H * 2};
                const cuuint32_t box[4] = {64, 128, (cuuint32_t)(pl.kch / mc), 1};
                if ((rc = make_tmap_bf16_nd(&d.tmA, s.hsx, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
            }
            {   // store of the new state tile: (unit, row, slab), box 32 units x 128 rows (64-byte rows)
                const cuuint64_t dims[3] = {(cuuint64_t)H, (cuuint64_t)B, (cuuint64_t)T + 1};
                const cuuint64_t str[2] = {(cuuint64_t)H * 2, (cuuint64_t)B * H * 2};
                const cuuint32_t box[3] = {(cuuint32_t)g.up, 128, 1};
                if ((rc = make_tmap_bf16_nd(&d.tmS, s.hsx, 3, dims, str, box, g.up == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
            }
            d.b_hh = s.b_hh; d.emb = (const __nv_bfloat16*)s.emb; d.ids = s.ids; d.proj = s.proj; d.proj_ld = s.proj_ld;
            d.dense = (const __nv_bfloat16*)s.dense;
            d.hsx = (__nv_bfloat16*)s.hsx; d.gates = (__nv_bfloat16*)s.gates;
            d.h_final = s.h_final; d.h_final_ld = s.h_final_ld;
            d.reverse = s.reverse;
        }
        P.bar = reinterpret_cast<unsigned*>(barrier_ws) + done * 16;
        P.n_chains = group; P.ppc = ppc; P.cpp = cpp; P.B = B; P.T = T; P.H = H;
        P.stages = pl.stages; P.kres = pl.kres; P.wst = pl.wst;
        P.dbg = fn_gru_dbg_ptr();
        const int pairs = ((group + cpp - 1) / cpp) * ppc;
        P.mc = mc;
        int rc = FN_OK, fits = 0;
        for (;;) {                                               // (the state boxes above were sized for `mc`: re-encode when it shrinks)
            rc = g.up == 64 ? launch_fwd<64>(pl.kch, P, pairs, pl.smem, st, &fits) : launch_fwd<32>(pl.kch, P, pairs, pl.smem, st, &fits);
            if (rc != FN_OK || fits || P.mc == 1) break;
            P.mc >>= 1;
            for (int i = 0; i < group; ++i) {
                const cuuint64_t dims[4] = {64, (cuuint64_t)B, (cuuint64_t)nkc, (cuuint64_t)T + 1};
                const cuuint64_t str[3] = {(cuuint64_t)H * 2, 128, (cuuint64_t)B * H * 2};
                const cuuint32_t box[4] = {64, 128, (cuuint32_t)(pl.kch / P.mc), 1};
                if ((rc = make_tmap_bf16_nd(&P.c[i].tmA, chains[done + i].hsx, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
            }
        }
        if (rc != FN_OK) return rc;
        FN_REQUIRE(fits, "fn_gru_seq_bf16: %d CTA pairs are not co-resident", pairs);
        done += group;
    }
    return FN_OK;
}
