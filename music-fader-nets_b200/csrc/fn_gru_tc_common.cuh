// Shared pieces of the persistent tcgen05 kernels (fn_gru_tc.cu: training recurrences; fn_decode_tc.cu: greedy decode):
// constants, TMEM / vector load-store helpers, the release/acquire publish, the shared-memory carve-up and the
// host-side shared-memory plan.  Everything lives in an anonymous namespace of the including translation unit.
#pragma once
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fn_tc.cuh"

namespace {

constexpr int kMaxChainsTc = 4;
constexpr int kThreadsTc = 352;           // 2 control warps + 8 epilogue warps + 1 weight-tail producer warp
constexpr int kWTailWarp = 10;
constexpr int kMaxWst = 32;               // slots of the streamed-weight ring
constexpr int kEpiThreads = 256;
constexpr int kATile = 128 * 64 * 2;      // 16 KB: 128 batch rows x 64 K (bf16), one 128B-swizzle atom wide
constexpr int kMaxStages = 8;
constexpr int kMaxNbt = 2;

// ---- TMEM <-> registers, 32 lanes x W consecutive fp32 columns (thread i <-> lane base + i) --------
template <int W>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, float (&v)[W]);
template <>
__device__ __forceinline__ void tmem_ld<16>(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    tc::tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
template <>
__device__ __forceinline__ void tmem_ld<8>(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
    tc::tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
// two loads, ONE wait (accumulator + time-invariant columns of a gate)
template <int W>
__device__ __forceinline__ void tmem_ld2(uint32_t ta, float (&a)[W], uint32_t tb, float (&b)[W]) {
    uint32_t ra[W], rb[W];
    if constexpr (W == 16) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(ra[0]), "=r"(ra[1]), "=r"(ra[2]), "=r"(ra[3]), "=r"(ra[4]), "=r"(ra[5]), "=r"(ra[6]), "=r"(ra[7]), "=r"(ra[8]),
              "=r"(ra[9]), "=r"(ra[10]), "=r"(ra[11]), "=r"(ra[12]), "=r"(ra[13]), "=r"(ra[14]), "=r"(ra[15])
            : "r"(ta) : "memory");
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(rb[0]), "=r"(rb[1]), "=r"(rb[2]), "=r"(rb[3]), "=r"(rb[4]), "=r"(rb[5]), "=r"(rb[6]), "=r"(rb[7]), "=r"(rb[8]),
              "=r"(rb[9]), "=r"(rb[10]), "=r"(rb[11]), "=r"(rb[12]), "=r"(rb[13]), "=r"(rb[14]), "=r"(rb[15])
            : "r"(tb) : "memory");
    } else {
        static_assert(W == 8, "tmem_ld2: 8 or 16 columns");
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(ra[0]), "=r"(ra[1]), "=r"(ra[2]), "=r"(ra[3]), "=r"(ra[4]), "=r"(ra[5]), "=r"(ra[6]), "=r"(ra[7])
                     : "r"(ta) : "memory");
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(rb[0]), "=r"(rb[1]), "=r"(rb[2]), "=r"(rb[3]), "=r"(rb[4]), "=r"(rb[5]), "=r"(rb[6]), "=r"(rb[7])
                     : "r"(tb) : "memory");
    }
    tc::tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < W; ++i) { a[i] = __uint_as_float(ra[i]); b[i] = __uint_as_float(rb[i]); }
}
template <int W>
__device__ __forceinline__ void tmem_st(uint32_t taddr, const float (&v)[W]);
template <>
__device__ __forceinline__ void tmem_st<16>(uint32_t taddr, const float (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
          "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])),
          "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])),
          "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])),
          "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
        : "memory");
}
template <>
__device__ __forceinline__ void tmem_st<8>(uint32_t taddr, const float (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                   "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
                   "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- W-wide vector loads / stores (W = 8 or 16 elements, 16-byte aligned) ----------------------------
template <int W>
__device__ __forceinline__ void ldf(const float* __restrict__ p, float (&v)[W]) {      // read-only fp32
    if constexpr (W % 8 == 0) {
        if ((reinterpret_cast<uintptr_t>(p) & 31) == 0) {                              // 256-bit accesses (see ldb_raw)
#pragma unroll
            for (int i = 0; i < W / 8; ++i)
                asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                             : "=f"(v[8 * i]), "=f"(v[8 * i + 1]), "=f"(v[8 * i + 2]), "=f"(v[8 * i + 3]), "=f"(v[8 * i + 4]),
                               "=f"(v[8 * i + 5]), "=f"(v[8 * i + 6]), "=f"(v[8 * i + 7])
                             : "l"(p + 8 * i));
            return;
        }
    }
#pragma unroll
    for (int i = 0; i < W / 4; ++i) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p) + i);
        v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
    }
}
template <int W>
__device__ __forceinline__ void stf(float* p, const float (&v)[W]) {
#pragma unroll
    for (int i = 0; i < W / 4; ++i)
        *(reinterpret_cast<float4*>(p) + i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}
// packed bf16: W/2 32-bit words.  W = 16 (32 bytes, 32-byte aligned) is ONE 256-bit access: the epilogues' rows are 8 KB
// apart, so every lane of a load / store touches its own sector and the LSU cost is per instruction x sector.
template <int W>
__device__ __forceinline__ void ldb_raw(const __nv_bfloat16* p, uint32_t (&w)[W / 2], bool coherent) {
    if constexpr (W == 16) {
        if (coherent)
            asm volatile("ld.global.cg.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(p));
        else
            asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(p));
    } else {
#pragma unroll
        for (int i = 0; i < W / 8; ++i) {
            const uint4 t = coherent ? __ldcg(reinterpret_cast<const uint4*>(p) + i) : __ldg(reinterpret_cast<const uint4*>(p) + i);
            w[4 * i] = t.x; w[4 * i + 1] = t.y; w[4 * i + 2] = t.z; w[4 * i + 3] = t.w;
        }
    }
}
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
template <int W>
__device__ __forceinline__ void unpack(const uint32_t (&w)[W / 2], float (&v)[W]) {
#pragma unroll
    for (int i = 0; i < W / 2; ++i) { v[2 * i] = bf_lo(w[i]); v[2 * i + 1] = bf_hi(w[i]); }
}
template <int W>
__device__ __forceinline__ void stb(__nv_bfloat16* p, const float (&v)[W]) {
    uint32_t w[W / 2];
#pragma unroll
    for (int i = 0; i < W / 2; ++i) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        w[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
    if constexpr (W == 16) {
        asm volatile("st.global.cg.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                     ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
    } else {
#pragma unroll
        for (int i = 0; i < W / 8; ++i)
            __stcg(reinterpret_cast<uint4*>(p) + i, make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]));
    }
}

// The saved gates are private to the forward / backward kernels, so they are stored in [32 rows][16 columns] blocks
// (1 KB; block order: time slab, 32-row block, 16-column block) instead of [T][B][4H] rows: a warp of the epilogue owns
// 32 consecutive rows x 16 (or 8) consecutive units, i.e. exactly one block per gate, so its loads and stores are
// contiguous 512-1024 B instead of 32 sectors 8 KB apart (the epilogues are LSU-transaction bound).
__device__ __forceinline__ long long gate_off(long long t, int b, int col, int B, int H4) {
    const long long row_blocks = (B + 31) >> 5;
    return ((t * row_blocks + (b >> 5)) * (long long)(H4 >> 4) + (col >> 4)) * 512 + (b & 31) * 16 + (col & 15);
}

__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory"); }

// Gate non-linearities of the bf16 path: ONE SFU instruction each (tanh.approx.f32, max relative error 2^-11 --
// a quarter of the bf16 rounding the saved gates and the MMA operand get anyway), sigmoid(x) = 0.5 tanh(0.5 x) + 0.5.
__device__ __forceinline__ float fast_tanh(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_sigmoid(float x) { return fmaf(0.5f, fast_tanh(0.5f * x), 0.5f); }

constexpr unsigned kEpiWarps = kEpiThreads / 32;
// Each epilogue warp publishes its part of a finished (step, batch tile) to the other slices of the chain:
// __syncwarp orders the lanes' stores before lane 0's release (cumulativity), red.release.gpu makes them
// visible GPU-wide before the counter moves; the proxy fence covers the consumers' TMA (async-proxy) reads.
__device__ __forceinline__ void publish(unsigned* ctr) {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) {
        asm volatile("fence.proxy.async.global;" ::: "memory");
        fn_red_release(ctr, 1u);
    }
}

struct Smem {
    uint8_t* W; uint8_t* WR; uint8_t* A;
    uint64_t *full, *empty, *acc_full, *acc_empty, *wbar, *wfull, *wempty;
    uint32_t* tmem_slot;
    float* bias;
};
__device__ __forceinline__ Smem carve(uint8_t* smem_raw, int w_res_bytes, int w_ring_bytes, int stages) {
    Smem s;
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    s.W = base;                                               // resident K chunks of the weight slice
    s.WR = base + w_res_bytes;                                // ring for the streamed chunks (all sizes multiples of 1024)
    s.A = s.WR + w_ring_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s.A + (size_t)stages * kATile);
    s.full = bars; s.empty = bars + kMaxStages; s.acc_full = s.empty + kMaxStages; s.acc_empty = s.acc_full + kMaxNbt;
    s.wbar = s.acc_empty + kMaxNbt;
    s.wfull = s.wbar + 1; s.wempty = s.wfull + kMaxWst;
    s.tmem_slot = reinterpret_cast<uint32_t*>(s.wempty + kMaxWst);
    s.bias = reinterpret_cast<float*>(s.tmem_slot + 2);
    return s;
}

// ---- host: shared-memory plan ---------------------------------------------------------------------------
constexpr size_t kSmemTail = 1024 /*align*/ + 1024 /*barriers*/ + 3 * 64 * 4 /*bias*/;

// Shared-memory plan of one kernel instance: how many K chunks of the weight slice stay resident, the ring
// that re-streams the others, and the state-slab ring.
struct TcPlan {
    int kch;        // K chunks (of 64) per state-ring stage
    int stages;     // state-ring stages (kch * 16 KB each)
    int kres;       // resident weight chunks (of nkc)
    int wst;        // weight-ring slots (0: everything resident)
    size_t smem;    // dynamic shared memory bytes
    bool ok;
};
int env_int(const char* name, int dflt) { const char* v = getenv(name); return v ? atoi(v) : dflt; }

// slot_per_stage: a weight-ring slot holds the kch chunks of one stage (fn_gru_tc.cu) instead of one chunk (fn_decode_tc.cu).
// kmul: plane products along K (3 in bf16x3 mode, see fn_gru_tc.cu)
TcPlan tc_plan_k(int U, int H, bool bwd, bool slot_per_stage, int kch_req, int kmul = 1) {
    TcPlan pl{};
    const int N = bwd ? U : 3 * U, nkc = kmul * (bwd ? 3 * H : H) / 64;
    const long long w_chunk = (long long)N * 128;
    const long long budget = (long long)fn_max_smem_optin() - (long long)kSmemTail;
    // state ring / weight ring when the weights do not all fit.  Defaults from the config-3 sweep (H = 1024, U = 32):
    // the recurrent kernels are paced by the MMA-issuing warp's fixed cost per ring stage (~1000 cycles of waits,
    // descriptor set-up and commits against 350-700 cycles of MMA execution), so FEWER, BIGGER stages win even though
    // less of the weight slice stays resident: 64 KB stages (kch = 4), forward 128 + 96 KB (nothing resident),
    // backward 160 + 32 KB.  The per-chunk rings of the greedy-decode kernel keep the earlier 96 / 32 KB.
    static const int ring_kb_f = env_int("FN_GRU_RING_KB", slot_per_stage ? 128 : 96);
    static const int wring_kb_f = env_int("FN_GRU_WRING_KB", slot_per_stage ? 96 : 32);
    static const int ring_kb_b = env_int("FN_GRU_RING_KB_BWD", slot_per_stage ? 160 : ring_kb_f);
    static const int wring_kb_b = env_int("FN_GRU_WRING_KB_BWD", slot_per_stage ? 32 : wring_kb_f);
    const int min_ring_kb = bwd ? ring_kb_b : ring_kb_f, wring_kb = bwd ? wring_kb_b : wring_kb_f;
    int kch = kch_req;                                                  // K chunks (of 64) per ring stage
    while (kch > 1 && nkc % kch) kch >>= 1;
    pl.kch = kch;
    const long long w_slot = slot_per_stage ? kch * w_chunk : w_chunk;  // bytes of one weight-ring slot
    long long room = budget - nkc * w_chunk;                            // ring space with a fully resident slice
    if (room >= 6LL * kATile) {
        pl.kres = nkc; pl.wst = 0;
    } else {
        // keep as much of the slice resident as leaves a >= min_ring_kb state ring + a small weight ring
        int wst = (int)((wring_kb * 1024LL + w_slot - 1) / w_slot);
        if (wst < 2) wst = 2;
        if (wst > kMaxWst) wst = kMaxWst;
        if (!slot_per_stage && wst < kch) wst = kch;
        long long kres = (budget - min_ring_kb * 1024LL - wst * w_slot) / w_chunk;
        if (kres > nkc - kch) kres = nkc - kch;
        kres -= kres % kch;                                             // streamed / resident parts in whole stages
        if (kres < 0) { pl.ok = false; return pl; }
        pl.kres = (int)kres; pl.wst = wst;
        room = budget - kres * w_chunk - wst * w_slot;
    }
    long long tiles = room / kATile;                                    // 16 KB tiles available to the state ring
    if (!slot_per_stage && tiles < 4) pl.kch = 1;
    if (tiles > kMaxStages * pl.kch) tiles = kMaxStages * pl.kch;
    pl.stages = (int)(tiles / pl.kch);
    if (pl.stages > kMaxStages) pl.stages = kMaxStages;
    pl.ok = pl.stages >= 2;
    static const int verbose = env_int("FN_GRU_VERBOSE", 0);
    if (verbose && pl.ok) fprintf(stderr, "tc_plan U=%d H=%d bwd=%d: kch=%d stages=%d kres=%d/%d wst=%d (slots of %lld B)\n", U, H, (int)bwd, pl.kch, pl.stages, pl.kres, nkc, pl.wst, w_slot);
    pl.smem = (size_t)(pl.kres * w_chunk + pl.wst * w_slot + (long long)pl.stages * pl.kch * kATile) + kSmemTail;
    return pl;
}

TcPlan tc_plan(int U, int H, bool bwd, bool slot_per_stage = false, int kmul = 1) {
    static const int kch_f = env_int("FN_GRU_KCH", 4), kch_b = env_int("FN_GRU_KCH_BWD", kch_f);
    int kch = slot_per_stage ? (bwd ? kch_b : kch_f) : 2;
    if (kch != 1 && kch != 2 && kch != 4) kch = 2;
    TcPlan pl = tc_plan_k(U, H, bwd, slot_per_stage, kch, kmul);
    while (!pl.ok && kch > 1 && slot_per_stage) pl = tc_plan_k(U, H, bwd, slot_per_stage, kch >>= 1, kmul);   // smaller stages fit more often
    return pl;
}

// bf16x3, compact K loop (fn_gru_tc.cu): BOTH planes of the weight slice resident (2 * K/64 chunks), the state ring behind them;
// a stage must not straddle the planes, so the chunks per stage divide K/64.
TcPlan tc_plan_x3c(int U, int H, bool bwd) {
    TcPlan pl{};
    const int N = bwd ? U : 3 * U, npl = (bwd ? 3 * H : H) / 64;
    const long long w_chunk = (long long)N * 128, budget = (long long)fn_max_smem_optin() - (long long)kSmemTail;
    for (int kch = 4; kch >= 1; kch >>= 1) {
        if (npl % kch) continue;
        const long long room = budget - 2LL * npl * w_chunk;
        long long stages = room / ((long long)kch * kATile);
        if (stages > kMaxStages) stages = kMaxStages;
        if (stages < 2) continue;
        pl.kch = kch; pl.stages = (int)stages; pl.kres = 2 * npl; pl.wst = 0; pl.ok = true;
        pl.smem = (size_t)(2LL * npl * w_chunk + stages * kch * kATile) + kSmemTail;
        return pl;
    }
    return pl;
}

int fn_make_tmap_bf16_3d(CUtensorMap* out, const void* base, unsigned long long d2, unsigned long long d1,
                         unsigned long long d0, unsigned long long ld1, unsigned box1, unsigned box0) {
    fn_PFN_encodeTiled enc = fn_get_encode_tiled();
    FN_REQUIRE(enc, "cuTensorMapEncodeTiled entry point unavailable");
    FN_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld1 * 2) % 16 == 0, "TMA 3-D operand alignment");
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t strides[2] = {ld1 * 2, ld1 * d1 * 2};
    cuuint32_t box[3] = {box0, box1, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FN_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(3d) failed (%d)", (int)r);
    return FN_OK;
}

}  // namespace
