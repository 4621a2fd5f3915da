// Persistent time-loop GRU on tcgen05 (bf16 operands, fp32 accumulation in TMEM): fn_gru_seq_fwd_bf16 /
// fn_gru_seq_bwd_bf16.  Same decomposition as the fp32 path (fn_gru_simt.cu): one CTA owns U hidden units
// of one chain for all T steps and keeps its weight slice RESIDENT in 128B-swizzled shared memory
// (forward: the 3U x H rows of W_hh; BPTT: the U x 3H rows of W_hh^T), loaded once by TMA.  Per step
//   warp 0      : waits for the chain's step counter, then streams the [B][K] state slab (K = H forward,
//                 3H backward) through a TMA ring of 128 x 64 K-major tiles;
//   warp 1      : one thread issues tcgen05.mma (M=128 batch rows, N=3U or U, K=16) into a per-batch-tile
//                 TMEM accumulator;
//   warps 2..9  : gate epilogue straight out of TMEM (tcgen05.ld): token gather, sigma/tanh/Hadamard, state
//                 and gate saves (forward) or the gate-gradient chain rule (backward); then publish the step.
// Everything is indexed BY STEP (slab s+1 of hsx = state after step s; slab 0 = h0), so reverse chains
// simply get reversed token / dense streams from the host.
#include "fn_tc.cuh"

namespace {

constexpr int kMaxChainsTc = 4;
constexpr int kThreadsTc = 320;           // 2 control warps + 8 epilogue warps
constexpr int kEpiThreads = 256;
constexpr int kAStages = 2;
constexpr int kATile = 128 * 64 * 2;      // 16 KB: 128 batch rows x 64 K
constexpr int kMaxAcc = 8;

struct TcChain {
    CUtensorMap tmW;       // resident operand: fwd W_hh [3H][H]; bwd W_hh^T [H][3H]   (box 64 x U)
    CUtensorMap tmA;       // streamed operand, 3-D [slabs][B][K]                      (box 64 x 128 x 1)
    // forward
    const float* b_hh; const float* emb; const int32_t* ids; const float* proj; long long proj_ld;
    const __nv_bfloat16* dense; const float* h0;
    __nv_bfloat16* hsx; float* hcur; __nv_bfloat16* gates; float* h_final; long long h_final_ld;
    // backward
    const __nv_bfloat16* dhs; const float* dh_final; long long dh_final_ld;
    __nv_bfloat16* dgh; __nv_bfloat16* dgin; float* dh0; float* carry;
};

struct TcLaunch {
    TcChain c[kMaxChainsTc];
    unsigned* bar;
    int n_chains, nslices, B, T, H;
};

__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, float (&v)[16]) {
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

__device__ __forceinline__ void ld16(const float* __restrict__ p, float (&v)[16]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p) + i);
        v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
    }
}
__device__ __forceinline__ void add16(const float* __restrict__ p, float (&v)[16]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p) + i);
        v[4 * i] += t.x; v[4 * i + 1] += t.y; v[4 * i + 2] += t.z; v[4 * i + 3] += t.w;
    }
}
__device__ __forceinline__ void ld16_cg(const float* p, float (&v)[16]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float4 t = __ldcg(reinterpret_cast<const float4*>(p) + i);
        v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
    }
}
__device__ __forceinline__ void st16_cg(float* p, const float (&v)[16]) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
        __stcg(reinterpret_cast<float4*>(p) + i, make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]));
}
__device__ __forceinline__ void ld16_bf(const __nv_bfloat16* p, float (&v)[16]) {      // 32 B, L2 (written by peers / this kernel)
    const uint4 a = __ldcg(reinterpret_cast<const uint4*>(p)), b = __ldcg(reinterpret_cast<const uint4*>(p) + 1);
    const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        v[2 * i] = __uint_as_float(w[i] << 16);
        v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}
__device__ __forceinline__ void add16_bf(const __nv_bfloat16* p, float (&v)[16]) {
    float t[16];
    ld16_bf(p, t);
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] += t[i];
}
__device__ __forceinline__ void st16_bf(__nv_bfloat16* p, const float (&v)[16]) {
    uint32_t w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        w[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
    __stcg(reinterpret_cast<uint4*>(p), make_uint4(w[0], w[1], w[2], w[3]));
    __stcg(reinterpret_cast<uint4*>(p) + 1, make_uint4(w[4], w[5], w[6], w[7]));
}

__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory"); }

// U = hidden units per CTA.  Forward: N = 3U gate columns, K = H.  Backward: N = U, K = 3H.
template <int U, bool BWD>
__global__ void __launch_bounds__(kThreadsTc, 1) gru_tc_kernel(const __grid_constant__ TcLaunch P) {
    constexpr int N = BWD ? U : 3 * U;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int H = P.H, B = P.B, T = P.T;
    const int K = BWD ? 3 * H : H;
    const int nkc = K / 64;
    const int w_chunk_bytes = N * 128;                        // one 64-wide K chunk of the resident operand
    uint8_t* Wsm = smem;
    uint8_t* Asm = smem + (size_t)nkc * w_chunk_bytes;        // kAStages x 16 KB (1024-aligned: N*128 is a multiple of 1024 for U % 8 == 0 ... checked on host)
    uint64_t* bars = reinterpret_cast<uint64_t*>(Asm + kAStages * kATile);
    uint64_t* full = bars;                   // [kAStages]
    uint64_t* empty = bars + kAStages;       // [kAStages]
    uint64_t* acc_full = empty + kAStages;   // [kMaxAcc]
    uint64_t* acc_empty = acc_full + kMaxAcc;
    uint64_t* wbar = acc_empty + kMaxAcc;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wbar + 1);
    float* bias_sm = reinterpret_cast<float*>(tmem_slot + 2);   // [3U] b_hh slice (forward)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int chain = blockIdx.x / P.nslices, slice = blockIdx.x % P.nslices;
    const TcChain& c = P.c[chain];
    unsigned* gbar = P.bar + chain * 16;
    const int u0 = slice * U;
    const int nbt = (B + 127) / 128;

    if (warp == 0 && lane == 0) {
        tc::prefetch_tmap(&c.tmW);
        tc::prefetch_tmap(&c.tmA);
        for (int i = 0; i < kAStages; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
        for (int i = 0; i < kMaxAcc; ++i) { tc::mbar_init(&acc_full[i], 1); tc::mbar_init(&acc_empty[i], 8); }
        tc::mbar_init(wbar, 1);
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc(tmem_slot, 512);
    if (!BWD && threadIdx.x >= 64) {
        for (int i = threadIdx.x - 64; i < 3 * U; i += kEpiThreads) bias_sm[i] = c.b_hh[(i / U) * H + u0 + (i % U)];
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // step schedule.  forward: s = 0..T-1, A slab = s (state before the step).
    // backward: s = T-1..-1, A slab = s+1 (gate gradient of the following step); s = T-1 has no product.
    const int n_iters = BWD ? T + 1 : T;

    if (warp == 0) {
        if (lane == 0) {
            // resident operand
            tc::mbar_arrive_expect_tx(wbar, (uint32_t)(nkc * w_chunk_bytes));
            for (int kc = 0; kc < nkc; ++kc) {
                if (!BWD) {
                    for (int g = 0; g < 3; ++g)
                        tc::tma_load_2d(Wsm + (size_t)kc * w_chunk_bytes + g * U * 128, &c.tmW, wbar, kc * 64, g * H + u0);
                } else {
                    tc::tma_load_2d(Wsm + (size_t)kc * w_chunk_bytes, &c.tmW, wbar, kc * 64, u0);
                }
            }
            uint32_t it = 0;
            for (int i = 0; i < n_iters; ++i) {
                const int s = BWD ? T - 1 - i : i;
                const bool has_a = BWD ? (i > 0) : true;
                if (!has_a) continue;
                if (i > 0) {
                    fn_spin_until(gbar, (unsigned)(P.nslices * i));
                    asm volatile("fence.proxy.async;" ::: "memory");
                }
                const int slab = BWD ? s + 1 : s;
                for (int bt = 0; bt < nbt; ++bt)
                    for (int kc = 0; kc < nkc; ++kc, ++it) {
                        const int st = it % kAStages;
                        tc::mbar_wait(&empty[st], ((it / kAStages) & 1) ^ 1);
                        tc::mbar_arrive_expect_tx(&full[st], kATile);
                        tc::tma_load_3d(Asm + st * kATile, &c.tmA, &full[st], kc * 64, bt * 128, slab);
                    }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = tc::make_idesc_bf16(128, N, 0, 0);
            tc::mbar_wait(wbar, 0);
            uint32_t it = 0, uses = 0;
            for (int i = 0; i < n_iters; ++i) {
                const bool has_a = BWD ? (i > 0) : true;
                if (!has_a) continue;
                for (int bt = 0; bt < nbt; ++bt) {
                    tc::mbar_wait(&acc_empty[bt], (uses & 1) ^ 1);
                    tc::tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(bt * N);
                    for (int kc = 0; kc < nkc; ++kc, ++it) {
                        const int st = it % kAStages;
                        tc::mbar_wait(&full[st], (it / kAStages) & 1);
                        tc::tc_fence_after();
                        const uint32_t sa = tc::smem_u32(Asm + st * kATile);
                        const uint32_t sb = tc::smem_u32(Wsm + (size_t)kc * w_chunk_bytes);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            tc::umma_f16(d_tmem, tc::make_sdesc(sa + k * 32, 16, 1024), tc::make_sdesc(sb + k * 32, 16, 1024),
                                         idesc, (kc | k) != 0);
                        tc::umma_commit(&empty[st]);
                    }
                    tc::umma_commit(&acc_full[bt]);
                }
                ++uses;
            }
        }
    } else {
        // ------------------------------- epilogue warps --------------------------------------------
        const int ew = warp - 2;                 // 0..7
        const int q = warp & 3;                  // TMEM lane quarter this warp may read
        const int half = ew >> 2;                // which half of the unit chunks
        constexpr int NCH = U / 16;              // 16-unit chunks
        uint32_t uses = 0;
        for (int i = 0; i < n_iters; ++i) {
            const int s = BWD ? T - 1 - i : i;
            const bool has_a = BWD ? (i > 0) : true;
            for (int bt = 0; bt < nbt; ++bt) {
                if (has_a) {
                    tc::mbar_wait(&acc_full[bt], uses & 1);
                    tc::tc_fence_after();
                }
                const int b = bt * 128 + q * 32 + lane;
                const bool row_ok = b < B;
                const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(bt * N);
                for (int ch = half; ch < NCH; ch += 2) {
                    const int uu = ch * 16;                  // unit offset inside the slice
                    const int u = u0 + uu;
                    if (!BWD) {
                        float ar[16], az[16], an[16];
                        tmem_ld_32x16(trow + uu, ar);
                        tmem_ld_32x16(trow + U + uu, az);
                        tmem_ld_32x16(trow + 2 * U + uu, an);
                        if (row_ok) {
                            float gr[16], gz[16], gn[16], hp[16];
#pragma unroll
                            for (int j = 0; j < 16; ++j) { gr[j] = 0.f; gz[j] = 0.f; gn[j] = 0.f; hp[j] = 0.f; }
                            const long long row = (long long)s * B + b;
                            if (c.emb) {
                                const float* e = c.emb + (long long)c.ids[row] * 3 * H + u;
                                add16(e, gr); add16(e + H, gz); add16(e + 2 * H, gn);
                            }
                            if (c.proj) {
                                const float* pj = c.proj + (long long)b * c.proj_ld + u;
                                add16(pj, gr); add16(pj + H, gz); add16(pj + 2 * H, gn);
                            }
                            if (c.dense) {
                                const __nv_bfloat16* dn = c.dense + row * 3 * H + u;
                                add16_bf(dn, gr); add16_bf(dn + H, gz); add16_bf(dn + 2 * H, gn);
                            }
                            if (i > 0) ld16_cg(c.hcur + (long long)b * H + u, hp);
                            else if (c.h0) ld16(c.h0 + (long long)b * H + u, hp);
                            float hn[16], sr[16], sz[16], sn[16], sg[16];
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                const float ghn = an[j] + bias_sm[2 * U + uu + j];
                                const float r = fn_sigmoid(gr[j] + ar[j] + bias_sm[uu + j]);
                                const float z = fn_sigmoid(gz[j] + az[j] + bias_sm[U + uu + j]);
                                const float n = tanhf(gn[j] + r * ghn);
                                hn[j] = (1.f - z) * n + z * hp[j];
                                sr[j] = r; sz[j] = z; sn[j] = n; sg[j] = ghn;
                            }
                            st16_cg(c.hcur + (long long)b * H + u, hn);
                            st16_bf(c.hsx + ((long long)(s + 1) * B + b) * H + u, hn);
                            if (c.gates) {
                                __nv_bfloat16* gsv = c.gates + row * 4 * H + u;
                                st16_bf(gsv, sr); st16_bf(gsv + H, sz); st16_bf(gsv + 2 * H, sn); st16_bf(gsv + 3 * H, sg);
                            }
                            if (s == T - 1 && c.h_final) {
                                float* hf = c.h_final + (long long)b * c.h_final_ld + u;
#pragma unroll
                                for (int j = 0; j < 16; ++j) hf[j] = hn[j];
                            }
                        }
                    } else {
                        float dh[16];
                        if (has_a) tmem_ld_32x16(trow + uu, dh);
                        else {
#pragma unroll
                            for (int j = 0; j < 16; ++j) dh[j] = 0.f;
                        }
                        if (row_ok) {
                            if (has_a) {
                                float cr[16];
                                ld16_cg(c.carry + (long long)b * H + u, cr);
#pragma unroll
                                for (int j = 0; j < 16; ++j) dh[j] += cr[j];
                            }
                            if (s < 0) {
                                st16_cg(c.dh0 + (long long)b * H + u, dh);
                            } else {
                                const long long row = (long long)s * B + b;
                                if (c.dhs) add16_bf(c.dhs + row * H + u, dh);
                                if (s == T - 1 && c.dh_final) add16(c.dh_final + (long long)b * c.dh_final_ld + u, dh);
                                float r[16], z[16], n[16], ghn[16], hp[16];
                                const __nv_bfloat16* gsv = c.gates + row * 4 * H + u;
                                ld16_bf(gsv, r); ld16_bf(gsv + H, z); ld16_bf(gsv + 2 * H, n); ld16_bf(gsv + 3 * H, ghn);
                                ld16_bf(c.hsx + row * H + u, hp);            // slab s = state before step s
                                float o_r[16], o_z[16], o_n[16], o_i[16], o_c[16];
#pragma unroll
                                for (int j = 0; j < 16; ++j) {
                                    const float dnp = dh[j] * (1.f - z[j]) * (1.f - n[j] * n[j]);
                                    o_z[j] = dh[j] * (hp[j] - n[j]) * z[j] * (1.f - z[j]);
                                    o_r[j] = dnp * ghn[j] * r[j] * (1.f - r[j]);
                                    o_n[j] = dnp * r[j];
                                    o_i[j] = dnp;
                                    o_c[j] = dh[j] * z[j];
                                }
                                __nv_bfloat16* dg = c.dgh + row * 3 * H + u;
                                st16_bf(dg, o_r); st16_bf(dg + H, o_z); st16_bf(dg + 2 * H, o_n);
                                st16_bf(c.dgin + row * H + u, o_i);
                                st16_cg(c.carry + (long long)b * H + u, o_c);
                            }
                        }
                    }
                }
                if (has_a) {
                    tc::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(&acc_empty[bt]);
                }
            }
            if (has_a) ++uses;
            // publish this step to the other slices of the chain
            if (i + 1 < n_iters) {
                epi_barrier();
                if (threadIdx.x == 64) {
                    __threadfence();
                    fn_red_release(gbar, 1u);
                }
            }
        }
        tc::tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem_base, 512);
    }
}

size_t tc_smem_bytes(int U, int H, bool bwd) {
    const int N = bwd ? U : 3 * U, K = bwd ? 3 * H : H;
    return (size_t)(K / 64) * N * 128 + kAStages * kATile + 1024 /*align*/ + 512 /*barriers*/ + 3 * U * 4;
}

int fn_make_tmap_bf16_3d(CUtensorMap* out, const void* base, unsigned long long d2, unsigned long long d1,
                         unsigned long long d0, unsigned box1, unsigned box0) {
    fn_PFN_encodeTiled enc = fn_get_encode_tiled();
    FN_REQUIRE(enc, "cuTensorMapEncodeTiled entry point unavailable");
    FN_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (d0 * 2) % 16 == 0, "TMA 3-D operand alignment");
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t strides[2] = {d0 * 2, d0 * d1 * 2};
    cuuint32_t box[3] = {box0, box1, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FN_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(3d) failed (%d)", (int)r);
    return FN_OK;
}

int pick_u_tc(int n_chains, int H, int B) {
    const int sms = fn_num_sms();
    const size_t cap = (size_t)fn_max_smem_optin();
    for (int U : {16, 32, 64}) {
        if (H % U) continue;
        if (tc_smem_bytes(U, H, false) > cap || tc_smem_bytes(U, H, true) > cap) break;
        const int nbt = (B + 127) / 128;
        if (nbt * 3 * U > 512 || nbt > kMaxAcc) break;
        if ((long long)n_chains * (H / U) <= sms) return U;
    }
    return 0;
}

template <int U, bool BWD>
int launch_tc(const TcLaunch& P, cudaStream_t st) {
    const size_t smem = tc_smem_bytes(U, P.H, BWD);
    const void* fn = (const void*)gru_tc_kernel<U, BWD>;
    FN_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    void* args[] = {(void*)&P};
    FN_CHECK_CUDA(cudaLaunchCooperativeKernel(fn, dim3(P.n_chains * P.nslices), dim3(kThreadsTc), args, smem, st));
    return FN_OK;
}

int run_tc(bool bwd, const FnGruChainBf16* chains, int n_chains, int B, int T, int H, void* barrier_ws, size_t ws_bytes,
           cudaStream_t st) {
    FN_REQUIRE(chains && n_chains > 0, "fn_gru_seq_bf16: no chains");
    FN_REQUIRE(B > 0 && T > 0 && H >= 64 && H % 64 == 0, "fn_gru_seq_bf16: need H %% 64 == 0 (H=%d)", H);
    FN_REQUIRE(barrier_ws && ws_bytes >= (size_t)64 * n_chains, "fn_gru_seq_bf16: barrier_ws too small");
    FN_CHECK_CUDA(cudaMemsetAsync(barrier_ws, 0, (size_t)64 * n_chains, st));
    int done = 0;
    while (done < n_chains) {
        int group = n_chains - done < kMaxChainsTc ? n_chains - done : kMaxChainsTc, U = 0;
        for (; group >= 1; --group)
            if ((U = pick_u_tc(group, H, B)) != 0) break;
        FN_REQUIRE(group >= 1, "fn_gru_seq_bf16: shape H=%d B=%d not supported by the tcgen05 path", H, B);
        TcLaunch P;
        memset(&P, 0, sizeof(P));
        for (int i = 0; i < group; ++i) {
            const FnGruChainBf16& s = chains[done + i];
            TcChain& d = P.c[i];
            int rc;
            if (!bwd) {
                FN_REQUIRE(s.w_hh && s.b_hh && s.hsx && s.hcur, "fn_gru_seq_fwd_bf16: chain %d misses buffers", done + i);
                FN_REQUIRE(!s.emb || s.ids, "fn_gru_seq_fwd_bf16: chain %d has emb without ids", done + i);
                rc = fn_make_tmap_bf16_2d(&d.tmW, s.w_hh, 3ull * H, H, H, U, 64);
                if (rc) return rc;
                rc = fn_make_tmap_bf16_3d(&d.tmA, s.hsx, T + 1, B, H, 128, 64);
                if (rc) return rc;
            } else {
                FN_REQUIRE(s.w_hh_t && s.hsx && s.gates && s.dgh && s.dgin && s.dh0 && s.carry,
                           "fn_gru_seq_bwd_bf16: chain %d misses buffers", done + i);
                rc = fn_make_tmap_bf16_2d(&d.tmW, s.w_hh_t, H, 3ull * H, 3ull * H, U, 64);
                if (rc) return rc;
                rc = fn_make_tmap_bf16_3d(&d.tmA, s.dgh, T, B, 3ull * H, 128, 64);
                if (rc) return rc;
            }
            d.b_hh = s.b_hh; d.emb = s.emb; d.ids = s.ids; d.proj = s.proj; d.proj_ld = s.proj_ld;
            d.dense = (const __nv_bfloat16*)s.dense; d.h0 = s.h0;
            d.hsx = (__nv_bfloat16*)s.hsx; d.hcur = s.hcur; d.gates = (__nv_bfloat16*)s.gates;
            d.h_final = s.h_final; d.h_final_ld = s.h_final_ld;
            d.dhs = (const __nv_bfloat16*)s.dhs; d.dh_final = s.dh_final; d.dh_final_ld = s.dh_final_ld;
            d.dgh = (__nv_bfloat16*)s.dgh; d.dgin = (__nv_bfloat16*)s.dgin; d.dh0 = s.dh0; d.carry = s.carry;
        }
        P.bar = reinterpret_cast<unsigned*>(barrier_ws) + done * 16;
        P.n_chains = group; P.nslices = H / U; P.B = B; P.T = T; P.H = H;
        int rc;
        if (U == 16) rc = bwd ? launch_tc<16, true>(P, st) : launch_tc<16, false>(P, st);
        else if (U == 32) rc = bwd ? launch_tc<32, true>(P, st) : launch_tc<32, false>(P, st);
        else rc = bwd ? launch_tc<64, true>(P, st) : launch_tc<64, false>(P, st);
        if (rc != FN_OK) return rc;
        done += group;
    }
    return FN_OK;
}

}  // namespace

extern "C" int fn_gru_seq_fwd_bf16(const FnGruChainBf16* chains, int n_chains, int B, int T, int H, void* barrier_ws,
                                   size_t barrier_ws_bytes, void* stream) {
    return run_tc(false, chains, n_chains, B, T, H, barrier_ws, barrier_ws_bytes, (cudaStream_t)stream);
}
extern "C" int fn_gru_seq_bwd_bf16(const FnGruChainBf16* chains, int n_chains, int B, int T, int H, void* barrier_ws,
                                   size_t barrier_ws_bytes, void* stream) {
    return run_tc(true, chains, n_chains, B, T, H, barrier_ws, barrier_ws_bytes, (cudaStream_t)stream);
}
