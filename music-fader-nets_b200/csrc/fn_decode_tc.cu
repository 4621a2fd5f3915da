// Greedy decode of the global decoder as ONE persistent kernel (fn_decode_greedy_bf16): reference
// gmm_model.py:119-149 in eval mode (cell 1 -> cell 2 -> vocabulary projection -> arg-max -> next token), the loop
// that test_class.py:233-254 / arousal_transfer.ipynb drive one sequence at a time.
//
// Four groups of CTAs ("roles") run the four products of a step; each keeps its weight slice resident in shared
// memory for the whole decode (same machinery as fn_gru_tc.cu: partially resident slice + streamed tail, TMA state
// ring, single-thread tcgen05 issue, epilogue out of TMEM), and they hand each other their results through global
// memory with release/acquire step counters:
//   role 0  cell 1     h1_s   = GRU(emb[tok_s] + z-projection, h1_{s-1})          needs: h1_{s-1} (own), arg-max partials of step s-1
//   role 1  input 2    gi2_s  = h1_s W_ih2^T + b_ih2                              needs: h1_s
//   role 2  cell 2     h2_s   = GRU(gi2_s, h2_{s-1}),  h2_{-1} := h1_0            needs: h2_{s-1} (own), gi2_s
//   role 3  logits     l_s    = h2_s W_out^T + b_out -> per-CTA arg-max partials   needs: h2_s
// The recurrent products of the two cells (h_{s-1} W_hh) only need the cell's OWN previous state, so their producers wait for
// nothing else and the tensor cores run them while the rest of the previous token is still in flight; the late operands
// (the token = arg-max partials for cell 1, gi2_s for cell 2) are awaited by the gate EPILOGUE.  Per token the dependent
// chain is therefore two products (roles 1 and 3) and two gate epilogues instead of four products.
// Sequences are independent, so the two 128-row batch tiles of a 256-row batch ping-pong through the roles.
// The token of step s+1 is reduced from the partials by every cell-1 thread for its own row (first maximum, like
// torch.max / the reference's _sampling); a small kernel afterwards turns the partials into the token matrix.
#include "fn_gru_tc_common.cuh"

namespace {

constexpr int kU = 32;                 // hidden units (weight rows per gate) per CTA
constexpr int kN = 3 * kU;             // MMA N: three 32-row groups of the role's weight matrix
constexpr int kRoles = 4;
constexpr int kUT = kU / 2;            // units per epilogue thread

struct DecRole {
    CUtensorMap tmW;        // weights [rows][H] bf16, box 64 x 32
    CUtensorMap tmA;        // streamed operand: state slabs [steps+1][B][H], box 64 x 128 x 1
    CUtensorMap tmA0;       // cell 2 only: cell 1's slabs (its "previous state" of step 0 is h1_0)
    int cta0, nslices;      // CTAs [cta0, cta0 + nslices)
    int w_slice_stride, w_gate_stride;   // weight rows of (slice, group g): slice*w_slice_stride + g*w_gate_stride + [0, 32)
    int a_slab_off;         // step s streams slab s + a_slab_off
    int wait_role[2], wait_off[2];       // producer waits: counter of wait_role >= per_step[wait_role] * (s + wait_off); -1 = none
};

struct DecLaunch {
    DecRole r[kRoles];
    unsigned* bar;          // [kRoles][16] step counters (per batch tile), zeroed by the host
    unsigned per_step[kRoles];
    int B, T, H, V, stages, kres, wst;
    // cell 1
    const float* b_hh1; const float* proj1; const __nv_bfloat16* emb1; int start_token;
    __nv_bfloat16* hs1;     // [T+1][B][H], slab 0 = initial state (caller-filled)
    // input projection of cell 2
    const float* b_ih2; __nv_bfloat16* gi2;       // [T][B][3H]
    // cell 2
    const float* b_hh2; __nv_bfloat16* hs2;       // [T+1][B][H], slab 0 unused (initial state = hs1 slab 1)
    // logits
    const float* b_out; float* pval; int32_t* pidx;   // arg-max partials [T][npart][B]
    float* logits_out;      // fp32 [T][B][V] or NULL
    int npart;
};

__device__ __forceinline__ void ld16_cg_bf(const __nv_bfloat16* p, uint32_t (&w)[8]) { ldb_raw<16>(p, w, true); }

template <int NBT, int KCH>
__global__ void __launch_bounds__(kThreadsTc, 1) decode_tc_kernel(const __grid_constant__ DecLaunch P) {
    constexpr uint32_t kAccCols = NBT * kN;
    constexpr uint32_t kTmemCols = 2 * kAccCols <= 256 ? 256 : 512;
    extern __shared__ uint8_t smem_raw[];
    const int H = P.H, B = P.B, T = P.T, S = P.stages;
    const int nkc = H / 64;
    constexpr int w_chunk_bytes = kN * 128;
    constexpr uint32_t stage_bytes = KCH * kATile;
    const int kres = P.kres, nstream = nkc - kres, WST = P.wst;
    const Smem sm = carve(smem_raw, kres * w_chunk_bytes, WST * KCH * w_chunk_bytes, S * KCH);   // a weight-ring slot = the KCH chunks of one stage

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int role = 0;
#pragma unroll
    for (int i = 1; i < kRoles; ++i)
        if ((int)blockIdx.x >= P.r[i].cta0) role = i;
    const DecRole& R = P.r[role];
    const int slice = blockIdx.x - R.cta0;
    unsigned* gbar = P.bar + role * 16;
    const int wrow0 = slice * R.w_slice_stride;           // first weight row of group 0
    const int u0 = slice * kU;                            // first hidden unit (cells, input projection)

    if (warp == 0 && lane == 0) {
        tc::prefetch_tmap(&R.tmW);
        tc::prefetch_tmap(&R.tmA);
        for (int i = 0; i < S; ++i) { tc::mbar_init(&sm.full[i], 1); tc::mbar_init(&sm.empty[i], 1); }
        for (int i = 0; i < NBT; ++i) { tc::mbar_init(&sm.acc_full[i], 1); tc::mbar_init(&sm.acc_empty[i], kEpiWarps); }
        tc::mbar_init(sm.wbar, 1);
        for (int i = 0; i < WST; ++i) { tc::mbar_init(&sm.wfull[i], 1); tc::mbar_init(&sm.wempty[i], 1); }
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc(sm.tmem_slot, kTmemCols);
    if (threadIdx.x >= 64 && threadIdx.x < 64 + kEpiThreads) {      // recurrent bias of the cells (r, z folded into P later; n kept)
        const float* bh = role == 0 ? P.b_hh1 : (role == 2 ? P.b_hh2 : nullptr);
        for (int i = threadIdx.x - 64; i < kN; i += kEpiThreads) sm.bias[i] = bh ? bh[(i / kU) * H + u0 + (i % kU)] : 0.f;
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *sm.tmem_slot;

    if (warp == 0) {
        // ------------------------------- state-slab producer ----------------------------------------
        if (tc::elect_one()) {
            tc::mbar_arrive_expect_tx(sm.wbar, (uint32_t)(kres * w_chunk_bytes));
            for (int kc = 0; kc < kres; ++kc)
                for (int g = 0; g < 3; ++g)
                    tc::tma_load_2d(sm.W + (size_t)kc * w_chunk_bytes + g * kU * 128, &R.tmW, sm.wbar, kc * 64,
                                    wrow0 + g * R.w_gate_stride);
        }
        __syncwarp();
        const uint32_t a0 = tc::smem_u32(sm.A), full0 = tc::smem_u32(sm.full), empty0 = tc::smem_u32(sm.empty);
        const int nst = nkc / KCH;
        uint32_t st = 0, ph = 1;
        for (int s = 0; s < T; ++s) {
            for (int bt = 0; bt < NBT; ++bt) {
#pragma unroll
                for (int w = 0; w < 2; ++w) {
                    const int wr = R.wait_role[w], steps_done = s + R.wait_off[w];
                    if (wr >= 0 && steps_done > 0) fn_spin_until(P.bar + wr * 16 + bt, P.per_step[wr] * (unsigned)steps_done);
                }
                const bool first_of_cell2 = (role == 2 && s == 0);
                if (first_of_cell2) fn_spin_until(P.bar + 0 * 16 + bt, P.per_step[0]);      // its "previous state" is h1_0
                asm volatile("fence.proxy.async.global;" ::: "memory");
                const CUtensorMap* tm = first_of_cell2 ? &R.tmA0 : &R.tmA;
                const int slab = first_of_cell2 ? 1 : s + R.a_slab_off;
                int col = kres * 64;                       // K order: streamed chunks first (see fn_gru_tc.cu)
                for (int j = 0; j < nst; ++j) {
                    if (j * KCH == nstream) col = 0;
                    const uint32_t fb = full0 + st * 8u, sa = a0 + st * stage_bytes;
                    tc::mbar_wait_u32(empty0 + st * 8u, ph);
                    if (tc::elect_one()) {
                        tc::mbar_arrive_expect_tx_u32(fb, stage_bytes);
#pragma unroll
                        for (int q = 0; q < KCH; ++q) tc::tma_load_3d_u32(sa + q * kATile, tm, fb, col + q * 64, bt * 128, slab);
                    }
                    __syncwarp();
                    col += 64 * KCH;
                    if (++st == (uint32_t)S) { st = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------- MMA issuer ---------------------------------------------------
        const uint32_t idesc = tc::make_idesc_bf16(128, kN, 0, 0);
        tc::mbar_wait(sm.wbar, 0);
        const uint32_t full0 = tc::smem_u32(sm.full), empty0 = tc::smem_u32(sm.empty);
        const uint64_t adesc0 = tc::make_sdesc(tc::smem_u32(sm.A), 16, 1024);
        const uint64_t bdesc0 = tc::make_sdesc(tc::smem_u32(sm.W), 16, 1024);
        const uint64_t wdesc0 = tc::make_sdesc(tc::smem_u32(sm.WR), 16, 1024);
        const uint32_t wfull0 = tc::smem_u32(sm.wfull), wempty0 = tc::smem_u32(sm.wempty);
        constexpr uint32_t a_step = stage_bytes >> 4, b_step = (uint32_t)w_chunk_bytes >> 4, slot_step = KCH * b_step;
        const int nst = nkc / KCH, nsst = nstream / KCH;        // stages per (step, tile); the streamed ones come first
        uint32_t st = 0, ph = 0, ws = 0, wph = 0;
        // same bare wait -> MMAs -> commits sequence as the training kernel (fn_gru_tc.cu): this single-thread loop
        // paces the role
        auto issue = [&](uint32_t d_tmem, uint64_t ad, uint64_t bd, bool first) {
#pragma unroll
            for (int q = 0; q < KCH; ++q) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    tc::umma_f16(d_tmem, ad + (uint64_t)(q * (kATile >> 4) + 2 * k), bd + (uint64_t)(q * b_step + 2 * k), idesc,
                                 (q | k) ? 1u : (first ? 0u : 1u));
            }
        };
        for (int s = 0; s < T; ++s) {
            for (int bt = 0; bt < NBT; ++bt) {
                tc::mbar_wait(&sm.acc_empty[bt], (s & 1) ^ 1);
                tc::tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(bt * kN);
                for (int j = 0; j < nsst; ++j) {                 // weights from the tail ring
                    tc::mbar_wait_u32(wfull0 + ws * 8u, wph);
                    tc::mbar_wait_u32(full0 + st * 8u, ph);
                    tc::tc_fence_after();
                    if (tc::elect_one()) {
                        issue(d_tmem, adesc0 + (uint64_t)(st * a_step), wdesc0 + (uint64_t)(ws * slot_step), j == 0);
                        tc::umma_commit_u32(wempty0 + ws * 8u);
                        tc::umma_commit_u32(empty0 + st * 8u);
                        if (j == nst - 1) tc::umma_commit(&sm.acc_full[bt]);
                    }
                    __syncwarp();
                    if (++ws == (uint32_t)WST) { ws = 0; wph ^= 1u; }
                    if (++st == (uint32_t)S) { st = 0; ph ^= 1u; }
                }
                uint64_t bd = bdesc0;
                for (int j = nsst; j < nst; ++j) {               // resident weights
                    tc::mbar_wait_u32(full0 + st * 8u, ph);
                    tc::tc_fence_after();
                    if (tc::elect_one()) {
                        issue(d_tmem, adesc0 + (uint64_t)(st * a_step), bd, j == 0);
                        tc::umma_commit_u32(empty0 + st * 8u);
                        if (j == nst - 1) tc::umma_commit(&sm.acc_full[bt]);
                    }
                    __syncwarp();
                    bd += (uint64_t)slot_step;
                    if (++st == (uint32_t)S) { st = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == kWTailWarp) {
        // ------------------------------- streamed part of the weight slice ----------------------------
        if (nstream > 0) {
            const uint32_t wr0 = tc::smem_u32(sm.WR), wfull0 = tc::smem_u32(sm.wfull), wempty0 = tc::smem_u32(sm.wempty);
            constexpr uint32_t slot_bytes = (uint32_t)(KCH * w_chunk_bytes);
            const int nsst = nstream / KCH;
            uint32_t ws = 0, wph = 1;
            for (int ts = 0; ts < T * NBT; ++ts) {
                for (int pst = 0; pst < nsst; ++pst) {
                    tc::mbar_wait_u32(wempty0 + ws * 8u, wph);
                    if (tc::elect_one()) {
                        const uint32_t dst = wr0 + ws * slot_bytes, fb = wfull0 + ws * 8u;
                        tc::mbar_arrive_expect_tx_u32(fb, slot_bytes);
#pragma unroll
                        for (int q = 0; q < KCH; ++q) {
#pragma unroll
                            for (int g = 0; g < 3; ++g)
                                tc::tma_load_2d_u32(dst + q * w_chunk_bytes + g * kU * 128, &R.tmW, fb, (kres + pst * KCH + q) * 64,
                                                    wrow0 + g * R.w_gate_stride);
                        }
                    }
                    __syncwarp();
                    if (++ws == (uint32_t)WST) { ws = 0; wph ^= 1u; }
                }
            }
        }
    } else {
        // ------------------------------- epilogue warps --------------------------------------------
        const int q = warp & 3;
        const int grp = (warp - 2) >> 2;
        const int uu = grp * kUT;                 // column offset inside each 32-wide group
        const int u = u0 + uu;
        const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
        const bool is_cell = (role == 0 || role == 2);
        float hreg[NBT][kUT];
        // time-invariant part of the pre-activations -> TMEM columns [kAccCols + bt*kN, +kN)
#pragma unroll
        for (int bt = 0; bt < NBT; ++bt) {
            const int b = bt * 128 + q * 32 + lane;
            const bool row_ok = b < B;
            float p0[kUT], p1[kUT], p2[kUT];
#pragma unroll
            for (int j = 0; j < kUT; ++j) { p0[j] = 0.f; p1[j] = 0.f; p2[j] = 0.f; hreg[bt][j] = 0.f; }
            if (role == 0) {
                if (row_ok) {
                    const float* pj = P.proj1 + (long long)b * 3 * H + u;
                    ldf<kUT>(pj, p0); ldf<kUT>(pj + H, p1); ldf<kUT>(pj + 2 * H, p2);
                    uint32_t hw[kUT / 2];
                    ldb_raw<kUT>(P.hs1 + (long long)b * H + u, hw, false);     // slab 0: initial state
                    unpack<kUT>(hw, hreg[bt]);
                }
#pragma unroll
                for (int j = 0; j < kUT; ++j) { p0[j] += sm.bias[uu + j]; p1[j] += sm.bias[kU + uu + j]; }
            } else if (role == 1) {
                ldf<kUT>(P.b_ih2 + u, p0); ldf<kUT>(P.b_ih2 + H + u, p1); ldf<kUT>(P.b_ih2 + 2 * H + u, p2);
            } else if (role == 2) {
#pragma unroll
                for (int j = 0; j < kUT; ++j) { p0[j] = sm.bias[uu + j]; p1[j] = sm.bias[kU + uu + j]; }
            } else {
#pragma unroll
                for (int g = 0; g < 3; ++g) {
                    float* pg = g == 0 ? p0 : (g == 1 ? p1 : p2);
#pragma unroll
                    for (int j = 0; j < kUT; ++j) {
                        const int col = wrow0 + g * R.w_gate_stride + uu + j;
                        pg[j] = col < P.V ? __ldg(P.b_out + col) : 0.f;
                    }
                }
            }
            const uint32_t tp = tmem_base + lane_sel + kAccCols + (uint32_t)(bt * kN);
            tmem_st<kUT>(tp + uu, p0); tmem_st<kUT>(tp + kU + uu, p1); tmem_st<kUT>(tp + 2 * kU + uu, p2);
        }
        tmem_st_wait();

        for (int s = 0; s < T; ++s) {
#pragma unroll
            for (int bt = 0; bt < NBT; ++bt) {
                const int b = bt * 128 + q * 32 + lane;
                const bool row_ok = b < B;
                const long long row = (long long)s * B + b;
                // ---- late operands of the cells: the token (arg-max partials of step s-1, role 3) / gi2_s (role 1); one lane
                // polls, the acquire + warp barrier order the other lanes' loads after it
                if (role == 0 && s > 0) {
                    if (lane == 0) fn_spin_until(P.bar + 3 * 16 + bt, P.per_step[3] * (unsigned)s);
                    __syncwarp();
                } else if (role == 2) {
                    if (lane == 0) fn_spin_until(P.bar + 1 * 16 + bt, P.per_step[1] * (unsigned)(s + 1));
                    __syncwarp();
                }
                // ---- input-side operand (cells): token-embedding gather (cell 1) / input projection (cell 2)
                uint32_t ir[kUT / 2], iz[kUT / 2], in_[kUT / 2];
#pragma unroll
                for (int j = 0; j < kUT / 2; ++j) { ir[j] = 0u; iz[j] = 0u; in_[j] = 0u; }
                if (row_ok && role == 0) {
                    int tok = P.start_token;
                    if (s > 0) {                                    // first maximum over the partial arg-maxes of step s-1
                        const float* pv = P.pval + ((long long)(s - 1) * P.npart) * B + b;
                        const int32_t* pi = P.pidx + ((long long)(s - 1) * P.npart) * B + b;
                        float best = -INFINITY;
                        tok = 0x7fffffff;
                        for (int pp = 0; pp < P.npart; ++pp) {
                            const float v = __ldcg(pv + (long long)pp * B);
                            const int id = __ldcg(pi + (long long)pp * B);
                            if (v > best || (v == best && id < tok)) { best = v; tok = id; }
                        }
                    }
                    const __nv_bfloat16* e = P.emb1 + (long long)tok * 3 * H + u;
                    ldb_raw<kUT>(e, ir, false); ldb_raw<kUT>(e + H, iz, false); ldb_raw<kUT>(e + 2 * H, in_, false);
                } else if (row_ok && role == 2) {
                    const __nv_bfloat16* g2 = P.gi2 + row * 3 * H + u;
                    ldb_raw<kUT>(g2, ir, true); ldb_raw<kUT>(g2 + H, iz, true); ldb_raw<kUT>(g2 + 2 * H, in_, true);
                    if (s == 0) {                                   // hx[1] <- the new hx[0] (gmm_model.py:134-135)
                        uint32_t hw[kUT / 2];
                        ldb_raw<kUT>(P.hs1 + ((long long)B + b) * H + u, hw, true);
                        unpack<kUT>(hw, hreg[bt]);
                    }
                }
                tc::mbar_wait_warp(&sm.acc_full[bt], s & 1);        // the product of this step (issued as soon as its state operand existed)
                tc::tc_fence_after();
                const uint32_t ta = tmem_base + lane_sel + (uint32_t)(bt * kN) + uu;
                const uint32_t tp = ta + kAccCols;
                float a0[kUT], a1[kUT], a2[kUT], p[kUT], x[kUT];
                tmem_ld<kUT>(ta, a0); tmem_ld<kUT>(tp, p);
#pragma unroll
                for (int j = 0; j < kUT; ++j) a0[j] += p[j];
                tmem_ld<kUT>(ta + kU, a1); tmem_ld<kUT>(tp + kU, p);
#pragma unroll
                for (int j = 0; j < kUT; ++j) a1[j] += p[j];
                tmem_ld<kUT>(ta + 2 * kU, a2); tmem_ld<kUT>(tp + 2 * kU, p);
                tc::tc_fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&sm.acc_empty[bt]);
                if (is_cell) {
                    float r[kUT], z[kUT];
                    unpack<kUT>(ir, x);
#pragma unroll
                    for (int j = 0; j < kUT; ++j) r[j] = fast_sigmoid(a0[j] + x[j]);
                    unpack<kUT>(iz, x);
#pragma unroll
                    for (int j = 0; j < kUT; ++j) z[j] = fast_sigmoid(a1[j] + x[j]);
                    unpack<kUT>(in_, x);
#pragma unroll
                    for (int j = 0; j < kUT; ++j) {
                        const float g = a2[j] + sm.bias[2 * kU + uu + j];
                        const float n = fast_tanh(p[j] + x[j] + r[j] * g);
                        hreg[bt][j] = (1.f - z[j]) * n + z[j] * hreg[bt][j];
                    }
                    __nv_bfloat16* hs = role == 0 ? P.hs1 : P.hs2;
                    if (row_ok) stb<kUT>(hs + ((long long)(s + 1) * B + b) * H + u, hreg[bt]);
                } else if (role == 1) {
#pragma unroll
                    for (int j = 0; j < kUT; ++j) a2[j] += p[j];
                    if (row_ok) {
                        __nv_bfloat16* o = P.gi2 + row * 3 * H + u;
                        stb<kUT>(o, a0); stb<kUT>(o + H, a1); stb<kUT>(o + 2 * H, a2);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < kUT; ++j) a2[j] += p[j];
                    float best = -INFINITY;
                    int bi = 0x7fffffff;
#pragma unroll
                    for (int g = 0; g < 3; ++g) {
                        const float* ag = g == 0 ? a0 : (g == 1 ? a1 : a2);
                        const int c0 = wrow0 + g * R.w_gate_stride + uu;
#pragma unroll
                        for (int j = 0; j < kUT; ++j) {
                            const int col = c0 + j;
                            if (col < P.V && ag[j] > best) { best = ag[j]; bi = col; }      // ascending columns: keeps the first max
                        }
                        if (P.logits_out && row_ok) {
                            float* lo = P.logits_out + row * P.V + c0;
#pragma unroll
                            for (int j = 0; j < kUT; ++j)
                                if (c0 + j < P.V) lo[j] = ag[j];
                        }
                    }
                    if (row_ok) {
                        const long long o = ((long long)s * P.npart + slice * 2 + grp) * B + b;
                        __stcg(P.pval + o, best);
                        __stcg(P.pidx + o, bi);
                    }
                }
                publish(gbar + bt);
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

// tokens[s][b] = first arg-max over the partials of step s
__global__ void decode_tokens_kernel(const float* __restrict__ pval, const int32_t* __restrict__ pidx, int T, int npart, int B,
                                     int32_t* __restrict__ tokens) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)T * B) return;
    const int s = (int)(i / B), b = (int)(i % B);
    float best = -INFINITY;
    int tok = 0x7fffffff;
    for (int pp = 0; pp < npart; ++pp) {
        const long long o = ((long long)s * npart + pp) * B + b;
        const float v = pval[o];
        const int id = pidx[o];
        if (v > best || (v == best && id < tok)) { best = v; tok = id; }
    }
    tokens[i] = tok;
}

template <int NBT, int KCH>
int launch_dec(const DecLaunch& P, int grid, size_t smem, cudaStream_t st) {
    const void* fn = (const void*)decode_tc_kernel<NBT, KCH>;
    FN_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    FN_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kThreadsTc, smem));
    FN_REQUIRE(per_sm * fn_num_sms() >= grid, "fn_decode_greedy_bf16: %d CTAs are not co-resident", grid);
    void* args[] = {(void*)&P};
    FN_CHECK_CUDA(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(kThreadsTc), args, smem, st));
    return FN_OK;
}

}  // namespace

extern "C" size_t fn_decode_greedy_ws_bytes(int B, int steps, int H, int V) {
    const int npart = ((V + kN - 1) / kN) * 2;
    return (size_t)kRoles * 64 + (size_t)steps * npart * B * 8 + 256;
}

extern "C" int fn_decode_greedy_bf16(const void* w_hh1, const float* b_hh1, const void* emb1, const float* proj1,
                                     const void* w_ih2, const float* b_ih2, const void* w_hh2, const float* b_hh2,
                                     const void* w_out, const float* b_out, void* hs1, void* hs2, void* gi2, int B, int steps,
                                     int H, int V, int start_token, int32_t* tokens, float* logits_out, void* workspace,
                                     size_t ws_bytes, void* stream) {
    FN_REQUIRE(w_hh1 && b_hh1 && emb1 && proj1 && w_ih2 && b_ih2 && w_hh2 && b_hh2 && w_out && b_out && hs1 && hs2 && gi2 &&
                   tokens && workspace, "fn_decode_greedy_bf16: null pointer");
    FN_REQUIRE(B > 0 && B <= 128 * kMaxNbt && steps > 0 && H >= 64 && H % 64 == 0 && V > 0,
               "fn_decode_greedy_bf16: need B <= 256, H %% 64 == 0 (B=%d H=%d)", B, H);
    FN_REQUIRE(ws_bytes >= fn_decode_greedy_ws_bytes(B, steps, H, V), "fn_decode_greedy_bf16: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const TcPlan pl = tc_plan(kU, H, false, true);
    FN_REQUIRE(pl.ok, "fn_decode_greedy_bf16: H=%d does not fit", H);
    DecLaunch P;
    memset(&P, 0, sizeof(P));
    const int nsl = H / kU, nlog = (V + kN - 1) / kN;
    const int nslices[kRoles] = {nsl, nsl, nsl, nlog};
    const void* wts[kRoles] = {w_hh1, w_ih2, w_hh2, w_out};
    const unsigned long long wrows[kRoles] = {3ull * H, 3ull * H, 3ull * H, (unsigned long long)V};
    const void* asrc[kRoles] = {hs1, hs1, hs2, hs2};
    int cta = 0;
    for (int i = 0; i < kRoles; ++i) {
        DecRole& R = P.r[i];
        int rc = fn_make_tmap_bf16_2d(&R.tmW, wts[i], wrows[i], H, H, kU, 64);
        if (rc) return rc;
        rc = fn_make_tmap_bf16_3d(&R.tmA, asrc[i], steps + 1, B, H, H, 128, 64);
        if (rc) return rc;
        rc = fn_make_tmap_bf16_3d(&R.tmA0, hs1, steps + 1, B, H, H, 128, 64);
        if (rc) return rc;
        R.cta0 = cta; R.nslices = nslices[i];
        cta += nslices[i];
        R.w_slice_stride = i == 3 ? kN : kU;
        R.w_gate_stride = i == 3 ? kU : H;
        R.a_slab_off = (i == 1 || i == 3) ? 1 : 0;
        R.wait_role[0] = R.wait_role[1] = -1;
        P.per_step[i] = (unsigned)nslices[i] * kEpiWarps;
    }
    FN_REQUIRE(cta <= fn_num_sms(), "fn_decode_greedy_bf16: H=%d needs %d CTAs (> %d SMs)", H, cta, fn_num_sms());
    // cell 1: own previous step (the arg-max partials of the previous step are awaited by its epilogue)
    P.r[0].wait_role[0] = 0; P.r[0].wait_off[0] = 0;
    // input projection of cell 2: cell 1 of this step
    P.r[1].wait_role[0] = 0; P.r[1].wait_off[0] = 1;
    // cell 2: own previous step (its input projection of this step is awaited by its epilogue)
    P.r[2].wait_role[0] = 2; P.r[2].wait_off[0] = 0;
    // logits: cell 2 of this step
    P.r[3].wait_role[0] = 2; P.r[3].wait_off[0] = 1;
    uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
    P.bar = reinterpret_cast<unsigned*>(ws);
    FN_CHECK_CUDA(cudaMemsetAsync(ws, 0, (size_t)kRoles * 64, st));
    P.npart = nlog * 2;
    P.pval = reinterpret_cast<float*>(ws + 256);
    P.pidx = reinterpret_cast<int32_t*>(ws + 256 + (size_t)steps * P.npart * B * 4);
    P.B = B; P.T = steps; P.H = H; P.V = V; P.stages = pl.stages; P.kres = pl.kres; P.wst = pl.wst;
    P.b_hh1 = b_hh1; P.proj1 = proj1; P.emb1 = (const __nv_bfloat16*)emb1; P.start_token = start_token;
    P.hs1 = (__nv_bfloat16*)hs1; P.b_ih2 = b_ih2; P.gi2 = (__nv_bfloat16*)gi2; P.b_hh2 = b_hh2; P.hs2 = (__nv_bfloat16*)hs2;
    P.b_out = b_out; P.logits_out = logits_out;
    const int nbt = (B + 127) / 128;
    int rc;
    if (pl.kch == 4) rc = nbt == 1 ? launch_dec<1, 4>(P, cta, pl.smem, st) : launch_dec<2, 4>(P, cta, pl.smem, st);
    else if (pl.kch == 2) rc = nbt == 1 ? launch_dec<1, 2>(P, cta, pl.smem, st) : launch_dec<2, 2>(P, cta, pl.smem, st);
    else rc = nbt == 1 ? launch_dec<1, 1>(P, cta, pl.smem, st) : launch_dec<2, 1>(P, cta, pl.smem, st);
    if (rc != FN_OK) return rc;
    decode_tokens_kernel<<<fn_cdiv((long long)steps * B, 256), 256, 0, st>>>(P.pval, P.pidx, steps, P.npart, B, tokens);
    FN_LAUNCH_CHECK();
    return FN_OK;
}
