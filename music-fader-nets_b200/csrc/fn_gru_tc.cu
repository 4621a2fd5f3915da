// Persistent time-loop GRU gate block on tcgen05 (bf16 operands, fp32 accumulation in TMEM):
// fn_gru_seq_fwd_bf16 / fn_gru_seq_bwd_bf16.
//
// Decomposition.  A chain (one recurrence) is cut into H/U hidden-unit slices; ONE CTA owns one slice of
// one chain for all T steps and keeps its weight slice RESIDENT in 128B-swizzled shared memory, loaded
// once by TMA (forward: the 3U rows of W_hh that produce its units' r/z/n gates; BPTT: the U rows of
// W_hh^T that produce its units' state gradient).  The batch is cut into 128-row tiles (the MMA M); the
// tiles of a chain are INDEPENDENT recurrences that share the resident weights, so the CTA ping-pongs
// between them: while the gate epilogue of tile 0 runs, the tensor core works on tile 1.
//
// Per (step, batch tile):
//   warp 0  (1 thread)  waits until every slice of the chain has published the previous step of this
//                       batch tile (release/acquire counter in global memory), then streams the
//                       [128][K] state slab (K = H forward, 3H backward) through a TMA ring;
//   warp 1  (1 thread)  issues tcgen05.mma (M = 128 batch rows, N = 3U | U, K = 16) into the tile's
//                       TMEM accumulator and commits ring slots / the accumulator to mbarriers;
//   warps 2..9          gate epilogue straight out of TMEM (tcgen05.ld): token-embedding gather,
//                       sigma / tanh / Hadamard, state + gate saves (forward) or the gate-gradient chain
//                       rule (backward); the running state (forward) / carried gradient (backward) of
//                       the thread's (row, units) stays in REGISTERS for the whole sequence; the
//                       time-invariant input projection (+ biases) lives in TMEM next to the accumulator.
// Everything in memory is indexed BY TIME: step s of a chain works on time tau = s (forward in time) or
// tau = T-1-s (reverse chain).  hsx has T+1 slabs: a forward chain keeps its initial state in slab 0 and
// h_tau in slab tau+1; a reverse chain keeps its initial state in slab T and h_tau in slab tau -- so
// "the state before step s" is slab tau (forward) / tau+1 (reverse) and lines up row for row with
// gates[tau] / dg[tau] in the batched weight-gradient GEMMs.
#include "fn_gru_tc_common.cuh"
#include "fn_gru_tc2.h"

namespace {

struct TcChain {
    CUtensorMap tmW;       // resident operand: fwd W_hh [3H][H]; bwd W_hh^T [H][3H]   (box 64 x U)
    CUtensorMap tmA;       // streamed operand, 3-D [slabs][B][K]                      (box 64 x 128 x 1)
    const float* b_hh; const __nv_bfloat16* emb; const int32_t* ids; const float* proj; long long proj_ld;
    const __nv_bfloat16* dense;
    __nv_bfloat16* hsx; __nv_bfloat16* gates; float* h_final; long long h_final_ld;
    const void* dhs; const float* dh_final; long long dh_final_ld;
    __nv_bfloat16* dg; float* dh0;
    int reverse, dhs_f32;
};

// Warp roles (15 warps; 23 in the tile-split BPTT, see Roles): warp 0 = state loader 0 (+ resident weights), warp 1 = MMA issuer, warps 2-3 = state loaders 1-2,
// warps 4-11 = the 8 gate-epilogue warps, warp 12 = state loader 3, warps 13-14 = weight-tail loaders 0-1.
// tools/ubench_tc.cu gets 30 / 73 / 110 / 126 B/clk of TMA ingest per SM with 1 / 2 / 3 / 4 issuing warps, so the ring
// stages CAN be dealt round-robin to several loader warps (FN_GRU_LS / FN_GRU_LW); in this kernel it was measured
// neutral (one loader thread sustains ~56 B/clk here and the MMA warp / the step chain pace it): one loader each.
// (setmaxnreg re-balancing between warpgroups was tried: ptxas then spills in the 80-register control roles.)
#ifndef FN_EPI16
#define FN_EPI16 0      /* 1: sixteen epilogue warps in the tile-split BPTT (measured slower once the epilogue accesses were coalesced: 18.6 vs 17.3 ms) */
#endif
#ifndef FN_PUB_CTA
#define FN_PUB_CTA 0     /* measured neutral (forward 11.6 vs 11.4 ms), off.  1: ONE release per (CTA, tile, step) after a named barrier of the tile's epilogue warps; 0: one per warp */
#endif
constexpr int kEpiWarp0 = 4;
// Publish a finished (step, tile) of this CTA to the other slices of the chain.  `nthr` epilogue threads work on the tile
// (barrier `id`); their stores happen-before the barrier, the barrier before the elected thread's release (cumulative at
// gpu scope), so ONE red.release per CTA is enough -- 8-16x fewer atomics on the step counter the consumers poll.
__device__ __forceinline__ void publish_tile(unsigned* ctr, int id, int nthr, bool leader) {
#if FN_PUB_CTA
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthr) : "memory");
    if (leader) {
        asm volatile("fence.proxy.async.global;" ::: "memory");
        fn_red_release(ctr, 1u);
    }
#else
    publish(ctr);
#endif
}
template <int EW, int NBT, bool BWD>
__host__ __device__ constexpr unsigned publishes_per_step() { return FN_PUB_CTA ? 1u : (unsigned)(BWD ? EW / NBT : EW); }
// Epilogue warps of an instantiation: 8, or 16 for the tile-split BPTT (64-unit slices, one batch tile per CTA): there a
// thread then owns ONE 16-unit chunk of its row and fetches all of its saved gates before the accumulator is ready; with
// 8 warps the second chunk's loads sat on the step's critical path (gate-gradient epilogue 8.4k of the 30k-cycle step).
template <int U, int NBT, bool BWD> struct Roles {
    static constexpr int kEpi = (BWD && U == 64 && NBT == 1 && FN_EPI16) ? 16 : 8;       // epilogue warps 4 .. 4 + kEpi - 1
    // 8 epilogue warps: + state loader 3 and tail loaders 0-1 after them (15 warps, 128 registers per thread).
    // 16 epilogue warps: no optional loaders, the tail loader is warp 2 (20 warps, 96 registers per thread).
    static constexpr bool kLean = kEpi == 16;
    static constexpr int kWarps = kLean ? kEpiWarp0 + kEpi : kEpiWarp0 + kEpi + 3;
    static constexpr int kThreads = kWarps * 32;
};
constexpr int kMaxStateLoaders = 4, kMaxTailLoaders = 2;
// `after` = first warp after the epilogue warps; lean = the 20-warp layout without optional loaders
__device__ __forceinline__ int state_loader_rank(int warp, int after, bool lean) {
    return warp == 0 ? 0 : lean ? -1 : warp == 2 ? 1 : warp == 3 ? 2 : warp == after ? 3 : -1;
}
__device__ __forceinline__ int tail_loader_rank(int warp, int after, bool lean) {
    return lean ? (warp == 2 ? 0 : -1) : warp == after + 1 ? 0 : warp == after + 2 ? 1 : -1;
}

struct TcLaunch {
    TcChain c[kMaxChainsTc];
    unsigned* bar;         // per chain 16 counters (one per batch tile), zeroed by the host
    long long* dbg;        // profiling aid (fn_gru_debug_timeline): [iteration][bt][16] clock64 stamps of CTA 0, or NULL
    int n_chains, nslices, B, T, H, stages, cluster;
    int tsplit;            // CTAs per (chain, slice): each owns NBT of the chain's 128-row batch tiles (tsplit * NBT tiles)
    int kres, wst;         // resident K chunks of the weight slice; ring slots for the streamed rest (0 = all resident)
    int ls, lw;            // issuing warps of the state stream (1..4) and of the weight-tail stream (1..2)
    int x3c;               // bf16x3, compact K loop (whole weight slice resident): the state stream carries each plane ONCE
                           // ([hi | lo], K'' = 2K) and a hi stage is multiplied with the hi AND the lo weight chunks -- a third
                           // fewer ring stages (the MMA-issuing warp's per-stage cost paces the kernel) and state bytes
    int box_rows;          // rows of a streamed state box: 128, or 64 / 32 when the chain's only batch tile has no more rows (the
                           // MMA still reads 128 rows of the slot; rows of one product are independent and the others are never stored)
};
#define FN_STAMP(i, bt, k)                                                                       \
    do {                                                                                         \
        if (dbg_on) P.dbg[((long long)(i) * kMaxNbt + (bt)) * 64 + (k)] = clock64(); \
    } while (0)

// =====================================================================================================
// Gate epilogue (8 warps, two warpgroups), see the kernel header below.
// =====================================================================================================
// The saved gates are private to the forward / backward kernels, so they are stored in [32 rows][16 columns] blocks
// (1 KB; block order: time slab, 32-row block, 16-column block) instead of [T][B][4H] rows: a warp of the epilogue owns
// 32 consecutive rows x 16 (or 8) consecutive units, i.e. exactly one block per gate, so its loads and stores are
// contiguous 512-1024 B instead of 32 sectors 8 KB apart (the epilogues are LSU-transaction bound).

// ---- bf16x3 mode (X3): every fp32 value that feeds a tensor-core product is carried as two bf16 planes, hi = bf16(x) and
// lo = bf16(x - hi) (16 mantissa bits), and the K loop runs over the three plane products hi*hi + lo*hi + hi*lo with fp32
// accumulation -- fp32-level results (~2^-16 relative per product) at bf16 tensor-core rate / 3.  State slabs are
// [B][2H] (hi | lo), saved gates [..][8H] (the 4H hi columns, then the 4H lo columns, same 32x16 blocks), the gate-gradient
// stream [B][8H]; token embeddings, the dense input stream and dhs are fp32; sigma / tanh use exp (not tanh.approx).
template <int W>
__device__ __forceinline__ void ld_split(const __nv_bfloat16* p, long long lo_off, float (&v)[W]) {
    uint32_t a[W / 2], b[W / 2];
    ldb_raw<W>(p, a, false); ldb_raw<W>(p + lo_off, b, false);
#pragma unroll
    for (int i = 0; i < W / 2; ++i) { v[2 * i] = bf_lo(a[i]) + bf_lo(b[i]); v[2 * i + 1] = bf_hi(a[i]) + bf_hi(b[i]); }
}
template <int W>
__device__ __forceinline__ void st_split(__nv_bfloat16* p, long long lo_off, const float (&v)[W]) {
    float lo[W];
#pragma unroll
    for (int i = 0; i < W; ++i) lo[i] = v[i] - __bfloat162float(__float2bfloat16(v[i]));
    stb<W>(p, v);
    stb<W>(p + lo_off, lo);
}
__device__ __forceinline__ float acc_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float acc_tanh(float x) { return fmaf(2.f, acc_sigmoid(2.f * x), -1.f); }

template <int U, int NBT, bool BWD, bool X3>
__device__ __forceinline__ void epilogue_role(const TcLaunch& P, const TcChain& c, const Smem& sm, const uint32_t tmem_base, unsigned* gbar,
                                              const int u0, const int tile0, const int warp, const int lane) {
    constexpr int N = BWD ? U : 3 * U;
    constexpr uint32_t kAccCols = NBT * N;
    constexpr int EW = Roles<U, NBT, BWD>::kEpi;
    const int H = P.H, B = P.B, T = P.T;
    const int HP = X3 ? 2 * H : H;             // pitch of a state row
    const int G4 = X3 ? 8 * H : 4 * H;         // pitch of a saved-gates / gate-gradient row
    const bool dbg_on = P.dbg != nullptr && blockIdx.x == 0 && lane == 0;
    {
        // ------------------------------- epilogue warps --------------------------------------------
        // All 8 warps serve batch tile 0, then tile 1: warp w reads TMEM lane quarter w % 4 (32 batch rows), the
        // two warp groups split the slice's units; each thread owns (row, UT units) of every tile for the whole
        // sequence.  The input-side operand of a chain is EITHER the token-embedding gather (bf16 table) or the
        // dense bf16 stream -- both have gate stride H, so one prefetch path serves both.
        constexpr int UT = U / 2;
        const int q = warp & 3;                  // TMEM lane quarter this warp may read (warp id % 4)
        const int uu = ((warp - kEpiWarp0) >> 2) * UT;   // unit offset inside the slice
        const int u = u0 + uu;
        const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
        const bool stamp = (threadIdx.x == kEpiWarp0 * 32);

        if constexpr (!BWD) {
            float hreg[NBT][UT];
            int id_next[NBT];
            // time-invariant part of the gate pre-activations -> TMEM columns [kAccCols + bt*N, +N)
#pragma unroll
            for (int bt = 0; bt < NBT; ++bt) {
                const int b = (tile0 + bt) * 128 + q * 32 + lane;
                const bool row_ok = b < B;
                float pr[UT], pz[UT], pn[UT];
#pragma unroll
                for (int j = 0; j < UT; ++j) { pr[j] = 0.f; pz[j] = 0.f; pn[j] = 0.f; hreg[bt][j] = 0.f; }
                if (row_ok) {
                    if (c.proj) {
                        const float* pj = c.proj + (long long)b * c.proj_ld + u;
                        ldf<UT>(pj, pr); ldf<UT>(pj + H, pz); ldf<UT>(pj + 2 * H, pn);
                    }
                    const __nv_bfloat16* hp0 = c.hsx + ((long long)(c.reverse ? T : 0) * B + b) * HP + u;
                    if constexpr (X3) {
                        ld_split<UT>(hp0, H, hreg[bt]);
                    } else {
                        uint32_t hw[UT / 2];
                        ldb_raw<UT>(hp0, hw, false);
                        unpack<UT>(hw, hreg[bt]);
                    }
                }
#pragma unroll
                for (int j = 0; j < UT; ++j) { pr[j] += sm.bias[uu + j]; pz[j] += sm.bias[U + uu + j]; }
                const uint32_t tp = tmem_base + lane_sel + kAccCols + (uint32_t)(bt * N);
                tmem_st<UT>(tp + uu, pr); tmem_st<UT>(tp + U + uu, pz); tmem_st<UT>(tp + 2 * U + uu, pn);
                id_next[bt] = (c.emb && row_ok) ? c.ids[(long long)(c.reverse ? T - 1 : 0) * B + b] : 0;
            }
            tmem_st_wait();
            const bool has_in = (c.emb != nullptr) || (c.dense != nullptr);

            for (int s = 0; s < T; ++s) {
                const int tau = c.reverse ? T - 1 - s : s;
                const int tau_n = c.reverse ? tau - 1 : tau + 1;
#pragma unroll
                for (int bt = 0; bt < NBT; ++bt) {
                    const int b = (tile0 + bt) * 128 + q * 32 + lane;
                    const bool row_ok = b < B;
                    const long long row_in = (long long)tau * B + b;       // input side is indexed by time
                    // ---- operand that does not depend on the recurrence: fetch before waiting for the MMA
                    constexpr int IW = X3 ? UT : UT / 2;                   // X3: fp32 inputs, else packed bf16
                    uint32_t ir[IW], iz[IW], in_[IW];
#pragma unroll
                    for (int j = 0; j < IW; ++j) { ir[j] = 0u; iz[j] = 0u; in_[j] = 0u; }
                    if (row_ok && has_in) {
                        if constexpr (X3) {
                            const float* src = c.emb ? reinterpret_cast<const float*>(c.emb) + (long long)id_next[bt] * 3 * H + u
                                                     : reinterpret_cast<const float*>(c.dense) + row_in * 3 * H + u;
                            ldf<UT>(src, reinterpret_cast<float(&)[UT]>(ir)); ldf<UT>(src + H, reinterpret_cast<float(&)[UT]>(iz));
                            ldf<UT>(src + 2 * H, reinterpret_cast<float(&)[UT]>(in_));
                        } else {
                            const __nv_bfloat16* src = c.emb ? c.emb + (long long)id_next[bt] * 3 * H + u : c.dense + row_in * 3 * H + u;
                            ldb_raw<UT>(src, ir, false); ldb_raw<UT>(src + H, iz, false); ldb_raw<UT>(src + 2 * H, in_, false);
                        }
                        if (c.emb && s + 1 < T) id_next[bt] = c.ids[(long long)tau_n * B + b];
                    }
                    if (stamp) FN_STAMP(s, bt, 6);
                    tc::mbar_wait_warp(&sm.acc_full[bt], s & 1);
                    tc::tc_fence_after();
                    if (stamp) FN_STAMP(s, bt, 7);
                    const uint32_t ta = tmem_base + lane_sel + (uint32_t)(bt * N) + uu;
                    const uint32_t tp = ta + kAccCols;
                    float a[UT], p[UT], x[UT], r[UT], z[UT], n[UT], g[UT];
                    auto input = [&](const uint32_t (&w)[IW]) {
                        if constexpr (X3) {
#pragma unroll
                            for (int j = 0; j < UT; ++j) x[j] = __uint_as_float(w[j]);
                        } else {
                            unpack<UT>(w, x);
                        }
                    };
                    tmem_ld<UT>(ta, a); tmem_ld<UT>(tp, p);
                    input(ir);
#pragma unroll
                    for (int j = 0; j < UT; ++j) r[j] = X3 ? acc_sigmoid(a[j] + p[j] + x[j]) : fast_sigmoid(a[j] + p[j] + x[j]);
                    tmem_ld<UT>(ta + U, a); tmem_ld<UT>(tp + U, p);
                    input(iz);
#pragma unroll
                    for (int j = 0; j < UT; ++j) z[j] = X3 ? acc_sigmoid(a[j] + p[j] + x[j]) : fast_sigmoid(a[j] + p[j] + x[j]);
                    tmem_ld<UT>(ta + 2 * U, a); tmem_ld<UT>(tp + 2 * U, p);
                    // accumulator consumed: the tensor core may start the next step of this tile
                    tc::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(&sm.acc_empty[bt]);
                    input(in_);
#pragma unroll
                    for (int j = 0; j < UT; ++j) {
                        g[j] = a[j] + sm.bias[2 * U + uu + j];
                        n[j] = X3 ? acc_tanh(p[j] + x[j] + r[j] * g[j]) : fast_tanh(p[j] + x[j] + r[j] * g[j]);
                        hreg[bt][j] = (1.f - z[j]) * n[j] + z[j] * hreg[bt][j];
                    }
                    if (row_ok) {
                        __nv_bfloat16* hdst = c.hsx + ((long long)(c.reverse ? tau : tau + 1) * B + b) * HP + u;
                        if constexpr (X3) st_split<UT>(hdst, H, hreg[bt]);
                        else stb<UT>(hdst, hreg[bt]);
                    }
                    if (stamp) FN_STAMP(s, bt, 9);
                    if (s + 1 < T) publish_tile(gbar + tile0 + bt, 1, EW * 32, warp == kEpiWarp0 && lane == 0);   // the next step only needs the state
                    if (stamp) FN_STAMP(s, bt, 10);
                    if (row_ok) {
                        if (c.gates) {                               // off the critical path: after the publish
                            if constexpr (X3) {
                                const long long lo = gate_off(tau, b, 4 * H + u, B, G4) - gate_off(tau, b, u, B, G4);
                                st_split<UT>(c.gates + gate_off(tau, b, u, B, G4), lo, r);
                                st_split<UT>(c.gates + gate_off(tau, b, H + u, B, G4), lo, z);
                                st_split<UT>(c.gates + gate_off(tau, b, 2 * H + u, B, G4), lo, n);
                                st_split<UT>(c.gates + gate_off(tau, b, 3 * H + u, B, G4), lo, g);
                            } else {
                                stb<UT>(c.gates + gate_off(tau, b, u, B, 4 * H), r);
                                stb<UT>(c.gates + gate_off(tau, b, H + u, B, 4 * H), z);
                                stb<UT>(c.gates + gate_off(tau, b, 2 * H + u, B, 4 * H), n);
                                stb<UT>(c.gates + gate_off(tau, b, 3 * H + u, B, 4 * H), g);
                            }
                        }
                        if (s == T - 1 && c.h_final) {               // caller-chosen offset / pitch: no alignment assumed
                            float* hf = c.h_final + (long long)b * c.h_final_ld + u;
#pragma unroll
                            for (int j = 0; j < UT; ++j) hf[j] = hreg[bt][j];
                        }
                    }
                }
            }
        } else {
            // Backward: the epilogue is load-heavy (saved gates, states, incoming gradients), so the two warp groups
            // each OWN one batch tile (all U units of a row per thread, in register chunks of CH) and run
            // concurrently instead of serving the tiles one after the other.
            constexpr int UB = U * NBT / (EW / 4);                  // units per thread: EW/4 warp groups over NBT tiles x U units
            constexpr int CH = UB > 16 ? 16 : UB;
            constexpr int NCHK = UB / CH;
            const int grp = (warp - kEpiWarp0) >> 2;
            const int bt = NBT == 2 ? grp : 0;
            const int uu0 = NBT == 2 ? 0 : grp * UB;
            static_assert(NBT == 1 || EW == 8, "two tiles per CTA: one 4-warp group per tile");
            const int b = (tile0 + bt) * 128 + q * 32 + lane;
            const bool row_ok = b < B;
            const uint32_t t_acc = tmem_base + lane_sel + (uint32_t)(bt * N);
            unsigned* gflag = gbar + tile0 + bt;
            uint64_t* accf = &sm.acc_full[bt];
            uint64_t* acce = &sm.acc_empty[bt];
            float carry[UB];
#pragma unroll
            for (int j = 0; j < UB; ++j) carry[j] = 0.f;
            for (int i = 0; i <= T; ++i) {
                const int s = T - 1 - i;
                const int tau = c.reverse ? T - 1 - s : s;
                const long long row = (long long)tau * B + b;
                // ---- saved forward values and incoming gradients: fetch (chunk 0) before waiting for the MMA
                // saved values: packed bf16 words, or (X3) fp32 = hi + lo plane, combined while the MMAs of the step run
                constexpr int SW = X3 ? CH : CH / 2;
                uint32_t wr[SW], wz[SW], wn[SW], wg[SW], wh[SW];
                float din[CH];
                auto fetch = [&](int ch) {
                    const int uc = u0 + uu0 + ch * CH;
#pragma unroll
                    for (int j = 0; j < SW; ++j) { wr[j] = 0; wz[j] = 0; wn[j] = 0; wg[j] = 0; wh[j] = 0; }
#pragma unroll
                    for (int j = 0; j < CH; ++j) din[j] = 0.f;
                    if (row_ok && s >= 0) {
                        if constexpr (X3) {
                            const long long lo = 128LL * H;                       // 4H columns further on in the blocked layout
                            ld_split<CH>(c.gates + gate_off(tau, b, uc, B, G4), lo, reinterpret_cast<float(&)[CH]>(wr));
                            ld_split<CH>(c.gates + gate_off(tau, b, H + uc, B, G4), lo, reinterpret_cast<float(&)[CH]>(wz));
                            ld_split<CH>(c.gates + gate_off(tau, b, 2 * H + uc, B, G4), lo, reinterpret_cast<float(&)[CH]>(wn));
                            ld_split<CH>(c.gates + gate_off(tau, b, 3 * H + uc, B, G4), lo, reinterpret_cast<float(&)[CH]>(wg));
                            ld_split<CH>(c.hsx + (row + (c.reverse ? B : 0)) * HP + uc, H, reinterpret_cast<float(&)[CH]>(wh));
                        } else {
                        ldb_raw<CH>(c.gates + gate_off(tau, b, uc, B, 4 * H), wr, false);
                        ldb_raw<CH>(c.gates + gate_off(tau, b, H + uc, B, 4 * H), wz, false);
                        ldb_raw<CH>(c.gates + gate_off(tau, b, 2 * H + uc, B, 4 * H), wn, false);
                        ldb_raw<CH>(c.gates + gate_off(tau, b, 3 * H + uc, B, 4 * H), wg, false);
                        ldb_raw<CH>(c.hsx + (row + (c.reverse ? B : 0)) * H + uc, wh, false);   // the state before step s
                        }
                        if (c.dhs) {
                            if (c.dhs_f32) ldf<CH>(reinterpret_cast<const float*>(c.dhs) + row * H + uc, din);
                            else {
                                uint32_t wd[CH / 2];
                                ldb_raw<CH>(reinterpret_cast<const __nv_bfloat16*>(c.dhs) + row * H + uc, wd, false);
                                unpack<CH>(wd, din);
                            }
                        }
                        if (s == T - 1 && c.dh_final) {
                            const float* df = c.dh_final + (long long)b * c.dh_final_ld + uc;
#pragma unroll
                            for (int j = 0; j < CH; ++j) din[j] += __ldg(df + j);
                        }
                    }
                };
                fetch(0);
                if (stamp) FN_STAMP(i, bt, 6);
                if (i > 0) {
                    tc::mbar_wait_warp(accf, (i - 1) & 1);
                    tc::tc_fence_after();
                }
                if (stamp) FN_STAMP(i, bt, 7);
                float o_i[CH];
#pragma unroll
                for (int ch = 0; ch < NCHK; ++ch) {
                    const int uc = u0 + uu0 + ch * CH;
                    if (ch > 0) fetch(ch);
                    float dh[CH];
                    if (i > 0) {
                        tmem_ld<CH>(t_acc + uu0 + ch * CH, dh);
                        if (ch == NCHK - 1) {
                            tc::tc_fence_before();
                            __syncwarp();
                            if (lane == 0) tc::mbar_arrive(acce);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < CH; ++j) dh[j] = 0.f;
                    }
#pragma unroll
                    for (int j = 0; j < CH; ++j) dh[j] += carry[ch * CH + j] + din[j];
                    if (s < 0) {
                        if (row_ok) stf<CH>(c.dh0 + (long long)b * H + uc, dh);
                        continue;
                    }
                    float r[CH], z[CH], n[CH], g[CH], hp[CH], o_r[CH], o_z[CH], o_n[CH];
                    if constexpr (X3) {
#pragma unroll
                        for (int j = 0; j < CH; ++j) {
                            r[j] = __uint_as_float(wr[j]); z[j] = __uint_as_float(wz[j]); n[j] = __uint_as_float(wn[j]);
                            g[j] = __uint_as_float(wg[j]); hp[j] = __uint_as_float(wh[j]);
                        }
                    } else {
                        unpack<CH>(wr, r); unpack<CH>(wz, z); unpack<CH>(wn, n); unpack<CH>(wg, g); unpack<CH>(wh, hp);
                    }
#pragma unroll
                    for (int j = 0; j < CH; ++j) {
                        const float dnp = dh[j] * (1.f - z[j]) * (1.f - n[j] * n[j]);
                        o_z[j] = dh[j] * (hp[j] - n[j]) * z[j] * (1.f - z[j]);
                        o_r[j] = dnp * g[j] * r[j] * (1.f - r[j]);
                        o_n[j] = dnp * r[j];
                        o_i[j] = dnp;
                        carry[ch * CH + j] = dh[j] * z[j];
                    }
                    if (row_ok) {
                        __nv_bfloat16* dgp = c.dg + row * G4 + uc;
                        if constexpr (X3) {
                            st_split<CH>(dgp, 4 * H, o_r); st_split<CH>(dgp + H, 4 * H, o_z); st_split<CH>(dgp + 3 * H, 4 * H, o_n);
                            if (NCHK > 1) st_split<CH>(dgp + 2 * H, 4 * H, o_i);
                        } else {
                            stb<CH>(dgp, o_r); stb<CH>(dgp + H, o_z); stb<CH>(dgp + 3 * H, o_n);
                            if (NCHK > 1) stb<CH>(dgp + 2 * H, o_i);
                        }
                    }
                }
                if (s < 0) continue;
                if (stamp) FN_STAMP(i, bt, 9);
                publish_tile(gflag, NBT == 2 ? 1 + grp : 1, NBT == 2 ? 128 : EW * 32, (NBT == 2 ? (warp - kEpiWarp0) % 4 == 0 : warp == kEpiWarp0) && lane == 0);   // the recurrence consumes (dr, dz, dn*r) only
                if (stamp) FN_STAMP(i, bt, 10);
                if (NCHK == 1 && row_ok) {
                    if constexpr (X3) st_split<CH>(c.dg + row * G4 + u0 + uu0 + 2 * H, 4 * H, o_i);
                    else stb<CH>(c.dg + row * 4 * H + u0 + uu0 + 2 * H, o_i);
                }
            }
        }
        tc::tc_fence_before();
    }
}

// =====================================================================================================
// U = hidden units per CTA, NBT = 128-row batch tiles per chain.  BWD = false: N = 3U gate columns,
// K = H.  BWD = true: N = U, K = 3H (A = gate gradients of the following step, pitch 4H).
// =====================================================================================================
template <int U, int NBT, bool BWD, int KCH, bool X3>
__global__ void __launch_bounds__((Roles<U, NBT, BWD>::kThreads), 1) gru_tc_kernel(const __grid_constant__ TcLaunch P) {
    constexpr int EW = Roles<U, NBT, BWD>::kEpi;
    constexpr int N = BWD ? U : 3 * U;
    constexpr uint32_t kAccCols = NBT * N;                     // accumulators; forward: + NBT*N projection columns
    constexpr uint32_t kNeedCols = BWD ? kAccCols : 2 * kAccCols;
    constexpr uint32_t kTmemCols = kNeedCols <= 32 ? 32 : kNeedCols <= 64 ? 64 : kNeedCols <= 128 ? 128 : kNeedCols <= 256 ? 256 : 512;
    static_assert(kNeedCols <= 512, "TMEM columns");
    extern __shared__ uint8_t smem_raw[];
    const int H = P.H, B = P.B, T = P.T, S = P.stages;
    const int Kb = BWD ? 3 * H : H;                            // K of one plane product
    const bool x3c = X3 && P.x3c;
    const int K = X3 ? (x3c ? 2 * Kb : 3 * Kb) : Kb;           // X3: the plane products side by side along K (compact: the two planes)
    const int nkc = K / 64;
    const int w_chunk_bytes = N * 128;                         // one 64-wide K chunk of the resident operand
    constexpr uint32_t stage_bytes = KCH * kATile;             // one ring stage: 128 rows x (KCH * 64) K
    // The weight slice is resident for its first `kres` K chunks; the remaining `nstream` chunks are re-streamed
    // every (step, batch tile) through a small ring by a dedicated warp.  They do not depend on the recurrence, so
    // that stream runs ahead of the step barrier; the K loop consumes the streamed chunks FIRST.
    const int kres = P.kres, nstream = nkc - kres, WST = P.wst;
    const Smem sm = carve(smem_raw, kres * w_chunk_bytes, WST * KCH * w_chunk_bytes, S * KCH);   // a tail slot = the KCH chunks of one stage

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool dbg_on = P.dbg != nullptr && blockIdx.x == 0 && lane == 0;      // profiling stamps (fn_gru_debug_timeline)
    const int cta_group = blockIdx.x / P.nslices, slice = blockIdx.x % P.nslices;
    const int chain = cta_group / P.tsplit, tile0 = (cta_group % P.tsplit) * NBT;   // first of this CTA's batch tiles
    const TcChain& c = P.c[chain];
    unsigned* gbar = P.bar + chain * 16;
    const int u0 = slice * U;

    if (warp == 0 && lane == 0) {
        tc::prefetch_tmap(&c.tmW);
        tc::prefetch_tmap(&c.tmA);
        for (int i = 0; i < S; ++i) { tc::mbar_init(&sm.full[i], 1); tc::mbar_init(&sm.empty[i], 1); }
        for (int i = 0; i < NBT; ++i) { tc::mbar_init(&sm.acc_full[i], 1); tc::mbar_init(&sm.acc_empty[i], BWD ? EW / NBT : EW); }
        tc::mbar_init(sm.wbar, 1);
        for (int i = 0; i < WST; ++i) { tc::mbar_init(&sm.wfull[i], 1); tc::mbar_init(&sm.wempty[i], 1); }
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc(sm.tmem_slot, kTmemCols);
    if (!BWD && warp >= kEpiWarp0 && warp < kEpiWarp0 + EW) {
        for (int i = threadIdx.x - kEpiWarp0 * 32; i < 3 * U; i += EW * 32) sm.bias[i] = c.b_hh[(i / U) * H + u0 + (i % U)];
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *sm.tmem_slot;
    constexpr bool kLean = Roles<U, NBT, BWD>::kLean;
    const int lrank = state_loader_rank(warp, kEpiWarp0 + EW, kLean), trank = tail_loader_rank(warp, kEpiWarp0 + EW, kLean);
    const int n_ls = kLean ? 1 : P.ls, n_lw = kLean ? 1 : P.lw;

    // iteration i: forward step s = i (A slab = s: the state before the step);
    // backward s = T-1-i for i = 0..T (s = -1 finishes dh0); the product of iteration i >= 1 reads slab s+1 of dg.
    const int n_iters = BWD ? T + 1 : T;

    auto state_loader_role = [&]() {
        // Control warps run CONVERGED (all 32 lanes execute the loop, one elected lane issues the async
        // instructions): the compiler then keeps the loop state in uniform registers and issues TMA / MMA /
        // commit without per-instruction election loops -- these single-thread loops pace the whole kernel.
        if (warp == 0 && tc::elect_one()) {
            tc::mbar_arrive_expect_tx(sm.wbar, (uint32_t)(kres * w_chunk_bytes));
            for (int kc = 0; kc < kres; ++kc) {
                // compact bf16x3: resident chunks = the hi plane, then the lo plane of the caller's [hi | hi | lo] layout
                const int kcol = (x3c && kc >= Kb / 64) ? 2 * Kb + (kc - Kb / 64) * 64 : kc * 64;
                if (!BWD) {
                    for (int g = 0; g < 3; ++g)
                        tc::tma_load_2d(sm.W + (size_t)kc * w_chunk_bytes + g * U * 128, &c.tmW, sm.wbar, kcol, g * H + u0);
                } else {
                    tc::tma_load_2d(sm.W + (size_t)kc * w_chunk_bytes, &c.tmW, sm.wbar, kcol, u0);
                }
            }
        }
        __syncwarp();
        {
            // ---- state-slab stream.  One thread, so everything per stage is kept to a handful of instructions:
            // raw shared addresses, incremental stage / phase / column counters, no divisions.
            const uint32_t a0 = tc::smem_u32(sm.A);
            const uint32_t full0 = tc::smem_u32(sm.full), empty0 = tc::smem_u32(sm.empty);
            const int nst = nkc / KCH;                                            // stages per (step, batch tile)
            uint32_t st = 0, ph = 1;                                              // ph: parity that means "slot free"
            int turn = 0;                                                         // stages are dealt round-robin to the loaders
            const int LS = n_ls;
            for (int i = BWD ? 1 : 0; i < n_iters; ++i) {
                // forward: the state before step s=i;  backward (s = T-1-i): the gate gradient of step s+1
                const int slab = BWD ? (c.reverse ? i - 1 : T - i) : (c.reverse ? T - i : i);
                for (int bt = 0; bt < NBT; ++bt) {
                    if (lrank == 0) FN_STAMP(i, bt, 0);
                    if (i > 0) {
                        fn_spin_until(gbar + tile0 + bt, (unsigned)(P.nslices * i) * publishes_per_step<EW, NBT, BWD>());
                        if (lrank == 0) FN_STAMP(i, bt, 1);
                        asm volatile("fence.proxy.async.global;" ::: "memory");
                    }
                    if (lrank == 0) FN_STAMP(i, bt, 2);
                    int col = kres * 64;                       // K order: streamed chunks [kres, nkc) first, then [0, kres)
                    for (int j = 0; j < nst; ++j) {
                        if (j * KCH == nstream) col = 0;
                        const uint32_t fb = full0 + st * 8u, sa = a0 + st * stage_bytes;
                        const bool mine = (turn == lrank);
                        if (mine) tc::mbar_wait_u32(empty0 + st * 8u, ph);
                        if (mine && tc::elect_one()) {
                            tc::mbar_arrive_expect_tx_u32(fb, (uint32_t)(KCH * P.box_rows * 128));
#pragma unroll
                            for (int q = 0; q < KCH; ++q) {
                                // dg columns are (dr, dz, dn, dn*r): the recurrent product consumes (dr, dz, dn*r).
                                // X3: K block -> (plane product, column): products 0 / 2 read the hi plane, product 1 the lo
                                // plane (forward: [B][H hi | H lo]; backward: [B][4H hi | 4H lo]); the weights' K axis is
                                // laid out [hi | hi | lo] by the caller, so their column is the K index itself.
                                int cq = col + q * 64, plane = 0;
                                if constexpr (X3) {
                                    const int KB = BWD ? 3 * H : H;
                                    const int prod = cq / KB;
                                    cq -= prod * KB;
                                    plane = prod == 1 ? (BWD ? 4 * H : H) : 0;
                                }
                                const int cc = ((BWD && cq >= 2 * H) ? cq + H : cq) + plane;
                                tc::tma_load_3d_u32(sa + q * kATile, &c.tmA, fb, cc, (tile0 + bt) * 128, slab);
                            }
                        }
                        __syncwarp();
                        col += 64 * KCH;
                        if (++st == (uint32_t)S) { st = 0; ph ^= 1u; }
                        if (++turn == LS) turn = 0;
                    }
                    if (lrank == 0) FN_STAMP(i, bt, 3);
                }
            }
        }
    };
    auto mma_role = [&]() {
        // This single-thread loop paces the whole kernel (ncu: ~1500 cycles per 64 KB stage against ~700 cycles of MMA
        // execution), so it is kept to the bare sequence wait -> MMAs -> commits: the streamed and the resident stages
        // are separate loops (no per-MMA selects), descriptors advance by constants, no profiling code inside.
        const uint32_t idesc = tc::make_idesc_bf16(128, N, 0, 0);
        tc::mbar_wait(sm.wbar, 0);
        const uint32_t full0 = tc::smem_u32(sm.full), empty0 = tc::smem_u32(sm.empty);
        const uint64_t adesc0 = tc::make_sdesc(tc::smem_u32(sm.A), 16, 1024);
        const uint64_t bdesc0 = tc::make_sdesc(tc::smem_u32(sm.W), 16, 1024);
        const uint64_t wdesc0 = tc::make_sdesc(tc::smem_u32(sm.WR), 16, 1024);
        const uint32_t wfull0 = tc::smem_u32(sm.wfull), wempty0 = tc::smem_u32(sm.wempty);
        constexpr uint32_t a_step = stage_bytes >> 4;                          // descriptor address units (16 B)
        const uint32_t b_step = (uint32_t)w_chunk_bytes >> 4, slot_step = KCH * b_step;
        const int nst = nkc / KCH, nsst = nstream / KCH;                       // stages per (step, tile); streamed ones first
        uint32_t st = 0, ph = 0, ws = 0, wph = 0, uses = 0;
        // one stage: KCH chunks x 4 MMAs (K = 16 each); `first` clears the accumulator with the very first MMA
        auto issue = [&](uint32_t d_tmem, uint64_t ad, uint64_t bd, bool first) {
#pragma unroll
            for (int q = 0; q < KCH; ++q) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    tc::umma_f16(d_tmem, ad + (uint64_t)(q * (kATile >> 4) + 2 * k), bd + (uint64_t)(q * b_step + 2 * k), idesc,
                                 (q | k) ? 1u : (first ? 0u : 1u));
            }
        };
        for (int i = BWD ? 1 : 0; i < n_iters; ++i, ++uses) {
            for (int bt = 0; bt < NBT; ++bt) {
                tc::mbar_wait(&sm.acc_empty[bt], (uses & 1) ^ 1);
                tc::tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(bt * N);
                for (int j = 0; j < nsst; ++j) {                               // weights from the tail ring
                    tc::mbar_wait_u32(wfull0 + ws * 8u, wph);
                    tc::mbar_wait_u32(full0 + st * 8u, ph);
                    tc::tc_fence_after();
                    if (j == 0) FN_STAMP(i, bt, 4);
                    if (tc::elect_one()) {
                        issue(d_tmem, adesc0 + (uint64_t)(st * a_step), wdesc0 + (uint64_t)(ws * slot_step), j == 0);
                        tc::umma_commit_u32(wempty0 + ws * 8u);                // weight-ring slot reusable
                        tc::umma_commit_u32(empty0 + st * 8u);                 // state-ring slot reusable
                        if (j == nst - 1) tc::umma_commit(&sm.acc_full[bt]);
                    }
                    __syncwarp();
                    if (++ws == (uint32_t)WST) { ws = 0; wph ^= 1u; }
                    if (++st == (uint32_t)S) { st = 0; ph ^= 1u; }
                }
                uint64_t bd = bdesc0;
                for (int j = nsst; j < nst; ++j) {                             // resident weights
                    tc::mbar_wait_u32(full0 + st * 8u, ph);
                    tc::tc_fence_after();
                    if (j == 0) FN_STAMP(i, bt, 4);
                    if (tc::elect_one()) {
                        if (X3 && x3c) {
                            // a hi-plane stage meets the hi and the lo weight chunks of its K range, a lo-plane stage the hi ones
                            const int kc0 = j * KCH, npl = Kb / 64;
                            const uint64_t ad = adesc0 + (uint64_t)(st * a_step);
                            if (kc0 < npl) {
                                issue(d_tmem, ad, bdesc0 + (uint64_t)(kc0 * b_step), j == 0);
                                issue(d_tmem, ad, bdesc0 + (uint64_t)((npl + kc0) * b_step), false);
                            } else {
                                issue(d_tmem, ad, bdesc0 + (uint64_t)((kc0 - npl) * b_step), false);
                            }
                        } else {
                            issue(d_tmem, adesc0 + (uint64_t)(st * a_step), bd, j == 0);
                        }
                        tc::umma_commit_u32(empty0 + st * 8u);
                        if (j == nst - 1) tc::umma_commit(&sm.acc_full[bt]);
                    }
                    __syncwarp();
                    bd += (uint64_t)slot_step;
                    if (++st == (uint32_t)S) { st = 0; ph ^= 1u; }
                }
                FN_STAMP(i, bt, 5);
            }
        }
    };
    auto tail_loader_role = [&]() {
        // ------------------------------- streamed part of the weight slice -----------------------------
        if (nstream > 0) {
            int turn = 0;
            const int LW = n_lw;
            const uint32_t wr0 = tc::smem_u32(sm.WR), wfull0 = tc::smem_u32(sm.wfull), wempty0 = tc::smem_u32(sm.wempty);
            const int n_tilesteps = T * NBT, nsst = nstream / KCH;           // streamed stages per (step, batch tile)
            const uint32_t slot_bytes = (uint32_t)(KCH * w_chunk_bytes);
            uint32_t ws = 0, wph = 1;
            for (int ts = 0; ts < n_tilesteps; ++ts) {
                for (int pst = 0; pst < nsst; ++pst) {
                    const bool mine = (turn == trank);
                    if (++turn == LW) turn = 0;
                    if (mine) tc::mbar_wait_u32(wempty0 + ws * 8u, wph);
                    if (mine && tc::elect_one()) {
                        const uint32_t dst = wr0 + ws * slot_bytes, fb = wfull0 + ws * 8u;
                        tc::mbar_arrive_expect_tx_u32(fb, slot_bytes);
#pragma unroll
                        for (int q = 0; q < KCH; ++q) {
                            const int kcol = (kres + pst * KCH + q) * 64;
                            if (!BWD) {
#pragma unroll
                                for (int g = 0; g < 3; ++g)
                                    tc::tma_load_2d_u32(dst + q * w_chunk_bytes + g * U * 128, &c.tmW, fb, kcol, g * H + u0);
                            } else {
                                tc::tma_load_2d_u32(dst + q * w_chunk_bytes, &c.tmW, fb, kcol, u0);
                            }
                        }
                    }
                    __syncwarp();
                    if (++ws == (uint32_t)WST) { ws = 0; wph ^= 1u; }
                }
            }
        }
    };

    if (warp >= kEpiWarp0 && warp < kEpiWarp0 + EW) epilogue_role<U, NBT, BWD, X3>(P, c, sm, tmem_base, gbar, u0, tile0, warp, lane);
    else if (warp == 1) mma_role();
    else if (lrank >= 0 && lrank < n_ls) state_loader_role();
    else if (trank >= 0 && trank < n_lw) tail_loader_role();
    __syncthreads();
    if (warp == 1) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ---- host -------------------------------------------------------------------------------------------
long long* g_dbg = nullptr;
// U for a launch of n_chains chains: the widest slice that fits the grid on the machine -- the streamed state
// slab is re-read by every slice of a chain, so fewer, wider slices mean less traffic per FLOP and bigger MMAs.
int pick_u_tc(int n_chains, int H, bool x3) {
    const int sms = fn_num_sms();
    static const int force_u = env_int("FN_GRU_U", 0);
    // bf16x3: the epilogues hold twice the saved values (hi + lo planes), 16-unit slices keep them in registers
    // (c2: 21.3 vs 23.3 ms per train step)
    const int order[2] = {x3 ? 16 : 32, x3 ? 32 : 16};
    for (int U : order) {
        if (force_u && U != force_u) continue;
        if (H % U) continue;
        if (!tc_plan(U, H, false, true, x3 ? 3 : 1).ok || !tc_plan(U, H, true, true, x3 ? 3 : 1).ok) continue;
        if ((long long)n_chains * (H / U) > sms) continue;
        return U;
    }
    return 0;
}

template <int U, int NBT, bool BWD, int KCH, bool X3>
int launch_tc(const TcLaunch& P, size_t smem, cudaStream_t st) {
    const void* fn = (const void*)gru_tc_kernel<U, NBT, BWD, KCH, X3>;
    FN_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(P.n_chains * P.tsplit * P.nslices);
    cfg.blockDim = dim3(Roles<U, NBT, BWD>::kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attrs[2];
    attrs[0].id = cudaLaunchAttributeClusterDimension;
    attrs[0].val.clusterDim.x = P.cluster; attrs[0].val.clusterDim.y = 1; attrs[0].val.clusterDim.z = 1;
    attrs[1].id = cudaLaunchAttributeCooperative;           // every CTA must be co-resident: they wait on each other
    attrs[1].val.cooperative = 1;
    cfg.attrs = attrs;
    cfg.numAttrs = 2;
    // all CTAs must be resident at once (slices of a chain wait for each other every step)
    int max_clusters = 0;
    FN_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, fn, &cfg));
    FN_REQUIRE(max_clusters * P.cluster >= P.n_chains * P.tsplit * P.nslices,
               "fn_gru_seq_bf16: %d CTAs in clusters of %d are not co-resident (max %d clusters)", P.n_chains * P.tsplit * P.nslices,
               P.cluster, max_clusters);
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
template <bool BWD, int KCH, bool X3>
int dispatch_tc2(int U, int nbt, const TcLaunch& P, size_t smem, cudaStream_t st) {
    if constexpr (BWD && !X3) {
        if (U == 64) return launch_tc<64, 1, BWD, KCH, X3>(P, smem, st);      // tile-split BPTT only (see run_tc)
    }
    if (U == 32) return nbt == 1 ? launch_tc<32, 1, BWD, KCH, X3>(P, smem, st) : launch_tc<32, 2, BWD, KCH, X3>(P, smem, st);
    return nbt == 1 ? launch_tc<16, 1, BWD, KCH, X3>(P, smem, st) : launch_tc<16, 2, BWD, KCH, X3>(P, smem, st);
}
template <bool BWD, bool X3>
int dispatch_tc(int U, int nbt, int kch, const TcLaunch& P, size_t smem, cudaStream_t st) {
    return kch == 4 ? dispatch_tc2<BWD, 4, X3>(U, nbt, P, smem, st)
         : kch == 2 ? dispatch_tc2<BWD, 2, X3>(U, nbt, P, smem, st) : dispatch_tc2<BWD, 1, X3>(U, nbt, P, smem, st);
}

// x3: bf16x3 mode (see epilogue_role): hi / lo plane operands, three plane products per recurrent product
int run_tc(bool bwd, bool x3, const FnGruChainBf16* chains, int n_chains, int B, int T, int H, void* barrier_ws, size_t ws_bytes,
           cudaStream_t st) {
    FN_REQUIRE(chains && n_chains > 0, "fn_gru_seq_bf16: no chains");
    FN_REQUIRE(B > 0 && T > 0 && H >= 64 && H % 64 == 0, "fn_gru_seq_bf16: need H %% 64 == 0 (H=%d)", H);
    FN_REQUIRE(B <= 128 * kMaxNbt, "fn_gru_seq_bf16: B=%d > %d rows per chain (split the batch into several chains)", B,
               128 * kMaxNbt);
    FN_REQUIRE(barrier_ws && ws_bytes >= (size_t)64 * n_chains, "fn_gru_seq_bf16: barrier_ws too small");
    FN_CHECK_CUDA(cudaMemsetAsync(barrier_ws, 0, (size_t)64 * n_chains, st));
    const int nbt = (B + 127) / 128;
    int done = 0;
    while (done < n_chains) {
        int group = n_chains - done < kMaxChainsTc ? n_chains - done : kMaxChainsTc, U = 0;
        for (; group >= 1; --group)
            if ((U = pick_u_tc(group, H, x3)) != 0) break;
        FN_REQUIRE(group >= 1, "fn_gru_seq_bf16: H=%d not supported by the tcgen05 path", H);
        // BPTT, tile split: the product has N = U columns and tcgen05.mma costs >= ~43 cycles whatever N <= 64 (it is
        // bound by the 4 KB A-tile read), so 64-unit slices double the work per MMA; each CTA then takes ONE 128-row
        // batch tile of its chain (half the MMAs and ring stages per step for the MMA-issuing warp that paces the
        // kernel) and the CTA count stays n_chains * (H / 32).
        static const int tsplit_env = env_int("FN_GRU_TSPLIT", 1);
        int tsplit = 1, nbt_cta = nbt;
        if (bwd && !x3 && tsplit_env && nbt == 2 && H % 64 == 0 && tc_plan(64, H, true, true).ok &&
            (long long)group * (H / 64) * nbt <= fn_num_sms()) {
            U = 64; tsplit = nbt; nbt_cta = 1;
        }
        TcLaunch P;
        memset(&P, 0, sizeof(P));
        // (cluster multicast of the state tiles between slices of a chain was implemented and measured neutral at 2, 4
        // and 8 CTAs per cluster -- the MMA-issue loop, not L2 reads, paces the kernel -- and removed.)
        const int cs = 1;
        P.cluster = cs;
        // a chain of <= 64 (<= 32) rows streams half (quarter) boxes: the recurrences of small batches (BASELINE config 2: B = 64)
        // are bound by the bytes the state stream moves into shared memory
        static const int small_box = env_int("FN_GRU_SMALL_BOX", 1);
        const int box_rows = (small_box && nbt == 1) ? (B <= 32 ? 32 : B <= 64 ? 64 : 128) : 128;
        P.box_rows = box_rows;
        for (int i = 0; i < group; ++i) {
            const FnGruChainBf16& s = chains[done + i];
            TcChain& d = P.c[i];
            int rc;
            FN_REQUIRE(s.hsx, "fn_gru_seq_bf16: chain %d misses hsx", done + i);
            if (!bwd) {
                FN_REQUIRE(s.w_hh && s.b_hh, "fn_gru_seq_fwd_bf16: chain %d misses weights", done + i);
                FN_REQUIRE(!s.emb || s.ids, "fn_gru_seq_fwd_bf16: chain %d has emb without ids", done + i);
                FN_REQUIRE(!(s.emb && s.dense), "fn_gru_seq_fwd_bf16: chain %d has both a token and a dense input", done + i);
                const unsigned long long kw = x3 ? 3ull * H : H, ka = x3 ? 2ull * H : H;   // x3: W [3H][hi|hi|lo], state [hi|lo]
                rc = fn_make_tmap_bf16_2d(&d.tmW, s.w_hh, 3ull * H, kw, kw, U, 64);
                if (rc) return rc;
                rc = fn_make_tmap_bf16_3d(&d.tmA, s.hsx, T + 1, B, ka, ka, box_rows, 64);
                if (rc) return rc;
            } else {
                FN_REQUIRE(s.w_hh_t && s.gates && s.dg && s.dh0, "fn_gru_seq_bwd_bf16: chain %d misses buffers", done + i);
                const unsigned long long kw = x3 ? 9ull * H : 3ull * H, ka = x3 ? 8ull * H : 4ull * H;
                FN_REQUIRE(!x3 || !s.dhs || s.dhs_f32, "fn_gru_seq_bwd_bf16x3: dhs must be fp32");
                rc = fn_make_tmap_bf16_2d(&d.tmW, s.w_hh_t, H, kw, kw, U, 64);
                if (rc) return rc;
                rc = fn_make_tmap_bf16_3d(&d.tmA, s.dg, T, B, ka, ka, box_rows, 64);
                if (rc) return rc;
            }
            d.b_hh = s.b_hh; d.emb = (const __nv_bfloat16*)s.emb; d.ids = s.ids; d.proj = s.proj; d.proj_ld = s.proj_ld;
            d.dense = (const __nv_bfloat16*)s.dense;
            d.hsx = (__nv_bfloat16*)s.hsx; d.gates = (__nv_bfloat16*)s.gates;
            d.h_final = s.h_final; d.h_final_ld = s.h_final_ld;
            d.dhs = s.dhs; d.dh_final = s.dh_final; d.dh_final_ld = s.dh_final_ld;
            d.dg = (__nv_bfloat16*)s.dg; d.dh0 = s.dh0;
            d.reverse = s.reverse; d.dhs_f32 = s.dhs_f32;
        }
        P.bar = reinterpret_cast<unsigned*>(barrier_ws) + done * 16;
        P.n_chains = group; P.nslices = H / U; P.B = B; P.T = T; P.H = H; P.tsplit = tsplit;
        TcPlan pl = tc_plan(U, H, bwd, true, x3 ? 3 : 1);
        if (x3) {                                            // compact K loop when the two weight planes fit next to a state ring
            static const int compact = env_int("FN_X3_COMPACT", 1);
            const TcPlan pc = tc_plan_x3c(U, H, bwd);
            if (compact && pc.ok) { pl = pc; P.x3c = 1; }
        }
        const int kch = pl.kch;
        P.stages = pl.stages; P.kres = pl.kres; P.wst = pl.wst;
        // Issuing warps per stream.  A ring slot must always be filled by the SAME loader (the "slot free" parity wait
        // is only safe against the loader's own previous fill of that slot), so the count has to divide the slot count.
        // Measured on config 3: 1..3 state loaders and 1..2 tail loaders give the same step period (the recurrence is
        // bound by its synchronisation chain, not by TMA issue), hence the default of one each.
        static const int ls_env = env_int("FN_GRU_LS", 1), lw_env = env_int("FN_GRU_LW", 1);
        P.ls = ls_env < 1 ? 1 : ls_env > kMaxStateLoaders ? kMaxStateLoaders : ls_env;
        P.lw = lw_env < 1 ? 1 : lw_env > kMaxTailLoaders ? kMaxTailLoaders : lw_env;
        while (P.stages % P.ls) --P.ls;
        while (P.wst > 0 && P.wst % P.lw) --P.lw;
        P.dbg = g_dbg;
        const int rc = x3 ? (bwd ? dispatch_tc<true, true>(U, nbt_cta, kch, P, pl.smem, st) : dispatch_tc<false, true>(U, nbt_cta, kch, P, pl.smem, st))
                          : (bwd ? dispatch_tc<true, false>(U, nbt_cta, kch, P, pl.smem, st) : dispatch_tc<false, false>(U, nbt_cta, kch, P, pl.smem, st));
        if (rc != FN_OK) return rc;
        done += group;
    }
    return FN_OK;
}

}  // namespace

// Profiling aid: when set (device buffer of >= (T+1)*2*64 int64), CTA 0 of every following launch records
// clock64 stamps of its pipeline events (see FN_STAMP).  NULL switches it off.  Not part of the product path.
long long* fn_gru_dbg_ptr() { return g_dbg; }
extern "C" int fn_gru_debug_timeline(void* device_buffer) {
    g_dbg = reinterpret_cast<long long*>(device_buffer);
    return FN_OK;
}

extern "C" int fn_gru_seq_fwd_bf16(const FnGruChainBf16* chains, int n_chains, int B, int T, int H, void* barrier_ws,
                                   size_t barrier_ws_bytes, void* stream) {
    if (fn_gru2_eligible(false, n_chains, B, H))          // two batch tiles: the CTA-pair kernel (fn_gru_tc2.cu)
        return fn_gru2_run(false, chains, n_chains, B, T, H, barrier_ws, barrier_ws_bytes, (cudaStream_t)stream);
    return run_tc(false, false, chains, n_chains, B, T, H, barrier_ws, barrier_ws_bytes, (cudaStream_t)stream);
}
extern "C" int fn_gru_seq_bwd_bf16(const FnGruChainBf16* chains, int n_chains, int B, int T, int H, void* barrier_ws,
                                   size_t barrier_ws_bytes, void* stream) {
    if (fn_gru2_eligible(true, n_chains, B, H))
        return fn_gru2_run(true, chains, n_chains, B, T, H, barrier_ws, barrier_ws_bytes, (cudaStream_t)stream);
    return run_tc(true, false, chains, n_chains, B, T, H, barrier_ws, barrier_ws_bytes, (cudaStream_t)stream);
}
extern "C" int fn_gru_seq_fwd_bf16x3(const FnGruChainBf16* chains, int n_chains, int B, int T, int H, void* barrier_ws,
                                     size_t barrier_ws_bytes, void* stream) {
    return run_tc(false, true, chains, n_chains, B, T, H, barrier_ws, barrier_ws_bytes, (cudaStream_t)stream);
}
extern "C" int fn_gru_seq_bwd_bf16x3(const FnGruChainBf16* chains, int n_chains, int B, int T, int H, void* barrier_ws,
                                     size_t barrier_ws_bytes, void* stream) {
    return run_tc(true, true, chains, n_chains, B, T, H, barrier_ws, barrier_ws_bytes, (cudaStream_t)stream);
}
