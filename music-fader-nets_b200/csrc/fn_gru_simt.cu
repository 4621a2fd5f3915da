// Persistent time-loop GRU, fp32 SIMT path (exact-parity mode): fn_gru_seq_fwd_f32 / _bwd_f32.
//
// Layout of the work: every chain (one recurrence) is cut into H/U hidden-unit slices; one CTA
// owns one slice for the WHOLE sequence and keeps its 3U x H slice of W_hh (forward) or the
// U x 3H slice of W_hh^T (backward) in shared memory, so per step only the [B][H] state (fwd)
// or the [B][3H] gate gradient (bwd) is streamed from L2.  Slices of a chain meet at a per-chain
// release/acquire counter once per step; chains never wait for each other.  The launch is
// cooperative (all CTAs co-resident) so the counters cannot deadlock.
#include "fn_common.cuh"

namespace {

constexpr int kMaxChains = 8;
constexpr int kThreads = 256;
constexpr int kBT = 64;    // batch rows per tile
constexpr int kKC = 64;    // K chunk streamed through smem
constexpr int kLDH = kKC + 4;

struct GruLaunch {
    FnGruChain c[kMaxChains];
    unsigned* bar;   // one counter per chain, 16 uints apart
    int n_chains, nslices, B, T, H;
};

template <int U>
struct Cfg {
    static constexpr int TU = U < 16 ? U : 16;   // threads along units
    static constexpr int UPT = U / TU;           // units per thread
    static constexpr int TB = kThreads / TU;     // threads along batch
    static constexpr int RPT = kBT / TB;         // batch rows per thread
    static_assert(RPT >= 1, "tile");
};

// stream rows [b0, b0+64) x cols [k0, k0+kKC) of src (row stride ld, `ncols` valid columns) into dst
__device__ __forceinline__ void load_chunk(float* dst, const float* __restrict__ src, long long ld, int b0, int B,
                                           int k0, int ncols) {
    for (int idx = threadIdx.x; idx < kBT * (kKC / 4); idx += kThreads) {
        const int row = idx / (kKC / 4), k = (idx % (kKC / 4)) * 4;
        float* d = dst + row * kLDH + k;
        if (b0 + row < B && k0 + k < ncols) fn_cp_async16(d, src + (long long)(b0 + row) * ld + k0 + k);
        else *reinterpret_cast<float4*>(d) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

template <int U>
__global__ void __launch_bounds__(kThreads, 1) gru_fwd_kernel(const GruLaunch P) {
    using C = Cfg<U>;
    extern __shared__ __align__(16) float smem[];
    const int H = P.H, B = P.B, T = P.T;
    const int LDW = H + 4;
    float* Ws = smem;                          // [3U][LDW]
    float* Hs = smem + 3 * U * LDW;            // [2][kBT][kLDH]

    const int chain = blockIdx.x / P.nslices, slice = blockIdx.x % P.nslices;
    const FnGruChain& c = P.c[chain];
    unsigned* bar = P.bar + chain * 16;
    const int u0 = slice * U;
    const int tid = threadIdx.x, tu = tid % C::TU, tb = tid / C::TU;

    for (int idx = tid; idx < 3 * U * (H / 4); idx += kThreads) {
        const int row = idx / (H / 4), k4 = idx % (H / 4);
        const int g = row / U, ul = row % U;
        const float4 v = *reinterpret_cast<const float4*>(c.w_hh + ((long long)(g * H + u0 + ul)) * H + k4 * 4);
        *reinterpret_cast<float4*>(&Ws[row * LDW + k4 * 4]) = v;
    }
    float bh[3][C::UPT];
#pragma unroll
    for (int g = 0; g < 3; ++g)
#pragma unroll
        for (int j = 0; j < C::UPT; ++j) bh[g][j] = c.b_hh[g * H + u0 + tu + C::TU * j];
    __syncthreads();

    const int nkc = (H + kKC - 1) / kKC;
    for (int s = 0; s < T; ++s) {
        const int tau = c.reverse ? T - 1 - s : s;
        const int tau_prev = c.reverse ? tau + 1 : tau - 1;
        const float* hprev = (s == 0) ? c.h0 : c.hs + (long long)tau_prev * B * H;
        if (s > 0) {
            if (tid == 0) fn_spin_until(bar, (unsigned)(P.nslices * s));
            __syncthreads();
        }
        for (int b0 = 0; b0 < B; b0 += kBT) {
            float acc[C::RPT][3][C::UPT];
#pragma unroll
            for (int i = 0; i < C::RPT; ++i)
#pragma unroll
                for (int g = 0; g < 3; ++g)
#pragma unroll
                    for (int j = 0; j < C::UPT; ++j) acc[i][g][j] = 0.f;

            if (hprev) {
                load_chunk(Hs, hprev, H, b0, B, 0, H);
                fn_cp_async_commit();
                for (int kc = 0; kc < nkc; ++kc) {
                    if (kc + 1 < nkc) {
                        load_chunk(Hs + ((kc + 1) & 1) * kBT * kLDH, hprev, H, b0, B, (kc + 1) * kKC, H);
                        fn_cp_async_commit();
                        fn_cp_async_wait<1>();
                    } else {
                        fn_cp_async_wait<0>();
                    }
                    __syncthreads();
                    const float* hb = Hs + (kc & 1) * kBT * kLDH;
                    const int kmax = min(kKC, H - kc * kKC);
                    const float* wb = Ws + kc * kKC;
                    for (int k = 0; k < kmax; k += 4) {
                        float4 hv[C::RPT];
#pragma unroll
                        for (int i = 0; i < C::RPT; ++i)
                            hv[i] = *reinterpret_cast<const float4*>(&hb[(tb + C::TB * i) * kLDH + k]);
#pragma unroll
                        for (int g = 0; g < 3; ++g)
#pragma unroll
                            for (int j = 0; j < C::UPT; ++j) {
                                const float4 wv =
                                    *reinterpret_cast<const float4*>(&wb[(g * U + tu + C::TU * j) * LDW + k]);
#pragma unroll
                                for (int i = 0; i < C::RPT; ++i) {
                                    float a = acc[i][g][j];
                                    a = fmaf(hv[i].x, wv.x, a);
                                    a = fmaf(hv[i].y, wv.y, a);
                                    a = fmaf(hv[i].z, wv.z, a);
                                    a = fmaf(hv[i].w, wv.w, a);
                                    acc[i][g][j] = a;
                                }
                            }
                    }
                    __syncthreads();
                }
            }

            // gate epilogue
#pragma unroll
            for (int i = 0; i < C::RPT; ++i) {
                const int b = b0 + tb + C::TB * i;
                if (b >= B) continue;
                const float* e = nullptr;
                if (c.emb) e = c.emb + (long long)c.ids[(long long)tau * B + b] * 3 * H;
                const float* pj = c.proj ? c.proj + (long long)b * c.proj_ld : nullptr;
                const float* dn = c.dense ? c.dense + ((long long)tau * B + b) * 3 * H : nullptr;
#pragma unroll
                for (int j = 0; j < C::UPT; ++j) {
                    const int u = u0 + tu + C::TU * j;
                    float gi[3] = {0.f, 0.f, 0.f};
#pragma unroll
                    for (int g = 0; g < 3; ++g) {
                        if (e) gi[g] += __ldg(e + g * H + u);
                        if (pj) gi[g] += __ldg(pj + g * H + u);
                        if (dn) gi[g] += __ldg(dn + g * H + u);
                    }
                    const float ghr = acc[i][0][j] + bh[0][j], ghz = acc[i][1][j] + bh[1][j],
                                ghn = acc[i][2][j] + bh[2][j];
                    const float hp = hprev ? __ldcg(hprev + (long long)b * H + u) : 0.f;
                    const float r = fn_sigmoid(gi[0] + ghr), z = fn_sigmoid(gi[1] + ghz);
                    const float n = tanhf(gi[2] + r * ghn);
                    const float h = (1.f - z) * n + z * hp;
                    const long long row = (long long)tau * B + b;
                    __stcg(c.hs + row * H + u, h);
                    if (c.gates) {
                        float* gsv = c.gates + row * 4 * H;
                        gsv[u] = r; gsv[H + u] = z; gsv[2 * H + u] = n; gsv[3 * H + u] = ghn;
                    }
                    if (s == T - 1 && c.h_final) c.h_final[(long long)b * c.h_final_ld + u] = h;
                }
            }
        }
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            fn_red_release(bar, 1u);
        }
    }
}

template <int U>
__global__ void __launch_bounds__(kThreads, 1) gru_bwd_kernel(const GruLaunch P) {
    using C = Cfg<U>;
    extern __shared__ __align__(16) float smem[];
    const int H = P.H, B = P.B, T = P.T;
    const int K3 = 3 * H, LDW = K3 + 4;
    float* Wt = smem;                       // [U][LDW]  Wt[ul][g] = W_hh[g][u0+ul]
    float* Gs = smem + U * LDW;             // [2][kBT][kLDH]

    const int chain = blockIdx.x / P.nslices, slice = blockIdx.x % P.nslices;
    const FnGruChain& c = P.c[chain];
    unsigned* bar = P.bar + chain * 16;
    const int u0 = slice * U;
    const int tid = threadIdx.x, tu = tid % C::TU, tb = tid / C::TU;

    for (int idx = tid; idx < K3 * U; idx += kThreads) {
        const int g = idx / U, ul = idx % U;
        Wt[ul * LDW + g] = c.w_hh[(long long)g * H + u0 + ul];
    }
    __syncthreads();

    const int nkc = (K3 + kKC - 1) / kKC;
    // s == -1 is the extra pass that finishes dh0
    for (int s = T - 1; s >= -1; --s) {
        const int tau = c.reverse ? T - 1 - s : s;
        const int tau_next = c.reverse ? tau - 1 : tau + 1;     // time of step s+1
        const int tau_prev = c.reverse ? tau + 1 : tau - 1;     // time of step s-1
        const bool have_next = (s < T - 1);
        if (have_next) {
            if (tid == 0) fn_spin_until(bar, (unsigned)(P.nslices * (T - 1 - s)));
            __syncthreads();
        }
        const float* gnext = have_next ? c.dgh + (long long)tau_next * B * K3 : nullptr;
        for (int b0 = 0; b0 < B; b0 += kBT) {
            float acc[C::RPT][C::UPT];
#pragma unroll
            for (int i = 0; i < C::RPT; ++i)
#pragma unroll
                for (int j = 0; j < C::UPT; ++j) acc[i][j] = 0.f;
            if (have_next) {
                load_chunk(Gs, gnext, K3, b0, B, 0, K3);
                fn_cp_async_commit();
                for (int kc = 0; kc < nkc; ++kc) {
                    if (kc + 1 < nkc) {
                        load_chunk(Gs + ((kc + 1) & 1) * kBT * kLDH, gnext, K3, b0, B, (kc + 1) * kKC, K3);
                        fn_cp_async_commit();
                        fn_cp_async_wait<1>();
                    } else {
                        fn_cp_async_wait<0>();
                    }
                    __syncthreads();
                    const float* gb = Gs + (kc & 1) * kBT * kLDH;
                    const int kmax = min(kKC, K3 - kc * kKC);
                    const float* wb = Wt + kc * kKC;
                    for (int k = 0; k < kmax; k += 4) {
                        float4 gv[C::RPT];
#pragma unroll
                        for (int i = 0; i < C::RPT; ++i)
                            gv[i] = *reinterpret_cast<const float4*>(&gb[(tb + C::TB * i) * kLDH + k]);
#pragma unroll
                        for (int j = 0; j < C::UPT; ++j) {
                            const float4 wv = *reinterpret_cast<const float4*>(&wb[(tu + C::TU * j) * LDW + k]);
#pragma unroll
                            for (int i = 0; i < C::RPT; ++i) {
                                float a = acc[i][j];
                                a = fmaf(gv[i].x, wv.x, a);
                                a = fmaf(gv[i].y, wv.y, a);
                                a = fmaf(gv[i].z, wv.z, a);
                                a = fmaf(gv[i].w, wv.w, a);
                                acc[i][j] = a;
                            }
                        }
                    }
                    __syncthreads();
                }
            }
#pragma unroll
            for (int i = 0; i < C::RPT; ++i) {
                const int b = b0 + tb + C::TB * i;
                if (b >= B) continue;
#pragma unroll
                for (int j = 0; j < C::UPT; ++j) {
                    const int u = u0 + tu + C::TU * j;
                    float dh = acc[i][j];
                    if (have_next) dh += __ldcg(c.dh_carry + (long long)b * H + u);
                    if (s < 0) {                       // dh0 = z_0 * dh_0 + dgh_0 W_hh
                        c.dh0[(long long)b * H + u] = dh;
                        continue;
                    }
                    const long long row = (long long)tau * B + b;
                    if (c.dhs) dh += __ldg(c.dhs + row * H + u);
                    if (s == T - 1 && c.dh_final) dh += __ldg(c.dh_final + (long long)b * c.dh_final_ld + u);
                    const float* gsv = c.gates + row * 4 * H;
                    const float r = gsv[u], z = gsv[H + u], n = gsv[2 * H + u], ghn = gsv[3 * H + u];
                    float hp = 0.f;
                    if (s > 0) hp = __ldg(c.hs + ((long long)tau_prev * B + b) * H + u);
                    else if (c.h0) hp = __ldg(c.h0 + (long long)b * H + u);
                    const float dnp = dh * (1.f - z) * (1.f - n * n);
                    const float dzp = dh * (hp - n) * z * (1.f - z);
                    const float drp = dnp * ghn * r * (1.f - r);
                    float* dg = c.dgh + row * K3;
                    __stcg(dg + u, drp);
                    __stcg(dg + H + u, dzp);
                    __stcg(dg + 2 * H + u, dnp * r);
                    c.dgin[row * H + u] = dnp;
                    __stcg(c.dh_carry + (long long)b * H + u, dh * z);
                }
            }
        }
        if (s >= 0) {
            __syncthreads();
            if (tid == 0) {
                __threadfence();
                fn_red_release(bar, 1u);
            }
        }
    }
}

size_t smem_bytes(int U, int H) {
    return ((size_t)3 * U * (H + 4) + 2 * kBT * kLDH) * sizeof(float) + 64;
}
size_t smem_bytes_bwd(int U, int H) {
    return ((size_t)U * (3 * H + 4) + 2 * kBT * kLDH) * sizeof(float) + 64;
}

// pick U: the smallest slice width (most CTAs) such that n_chains * H/U <= SMs and the slice fits smem.
// returns 0 if this many chains cannot run in one launch.
int pick_u(int n_chains, int H) {
    const int sms = fn_num_sms();
    const size_t cap = (size_t)fn_max_smem_optin();
    for (int U : {4, 8, 16, 32}) {
        if (H % U) continue;
        if (smem_bytes(U, H) > cap || smem_bytes_bwd(U, H) > cap) break;
        if ((long long)n_chains * (H / U) <= sms) return U;
    }
    return 0;
}

template <int U>
int launch_one(bool bwd, const GruLaunch& P, cudaStream_t st) {
    const size_t smem = bwd ? smem_bytes_bwd(U, P.H) : smem_bytes(U, P.H);
    const void* fn = bwd ? (const void*)gru_bwd_kernel<U> : (const void*)gru_fwd_kernel<U>;
    FN_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    void* args[] = {(void*)&P};
    FN_CHECK_CUDA(cudaLaunchCooperativeKernel(fn, dim3(P.n_chains * P.nslices), dim3(kThreads), args, smem, st));
    return FN_OK;
}

int run(bool bwd, const FnGruChain* chains, int n_chains, int B, int T, int H, void* barrier_ws, size_t ws_bytes,
        cudaStream_t st) {
    FN_REQUIRE(chains && n_chains > 0, "fn_gru_seq: no chains");
    FN_REQUIRE(B > 0 && T > 0 && H > 0 && H % 4 == 0, "fn_gru_seq: need B,T>0 and H %% 4 == 0 (H=%d)", H);
    FN_REQUIRE(barrier_ws && ws_bytes >= (size_t)64 * n_chains, "fn_gru_seq: barrier_ws too small");
    for (int i = 0; i < n_chains; ++i) {
        const FnGruChain& c = chains[i];
        FN_REQUIRE(c.w_hh && c.b_hh && c.hs, "fn_gru_seq: chain %d misses w_hh/b_hh/hs", i);
        FN_REQUIRE(!c.emb || c.ids, "fn_gru_seq: chain %d has emb without ids", i);
        if (bwd) FN_REQUIRE(c.gates && c.dgh && c.dgin && c.dh0 && c.dh_carry, "fn_gru_seq_bwd: chain %d misses buffers", i);
    }
    FN_CHECK_CUDA(cudaMemsetAsync(barrier_ws, 0, (size_t)64 * n_chains, st));
    int done = 0;
    while (done < n_chains) {
        int group = min(n_chains - done, kMaxChains), U = 0;
        for (; group >= 1; --group)
            if ((U = pick_u(group, H)) != 0) break;
        FN_REQUIRE(group >= 1, "fn_gru_seq: H=%d does not fit (needs H/U <= #SMs with the slice in smem)", H);
        GruLaunch P;
        for (int i = 0; i < group; ++i) P.c[i] = chains[done + i];
        P.bar = reinterpret_cast<unsigned*>(barrier_ws) + done * 16;
        P.n_chains = group; P.nslices = H / U; P.B = B; P.T = T; P.H = H;
        int rc;
        switch (U) {
            case 4: rc = launch_one<4>(bwd, P, st); break;
            case 8: rc = launch_one<8>(bwd, P, st); break;
            case 16: rc = launch_one<16>(bwd, P, st); break;
            default: rc = launch_one<32>(bwd, P, st); break;
        }
        if (rc != FN_OK) return rc;
        done += group;
    }
    return FN_OK;
}

}  // namespace

extern "C" int fn_gru_seq_ctas_per_chain(int H) {
    if (H <= 0 || H % 4) return 0;
    const int U = pick_u(1, H);
    return U ? H / U : 0;
}

extern "C" int fn_gru_seq_fwd_f32(const FnGruChain* chains, int n_chains, int B, int T, int H, void* barrier_ws,
                                  size_t barrier_ws_bytes, void* stream) {
    return run(false, chains, n_chains, B, T, H, barrier_ws, barrier_ws_bytes, (cudaStream_t)stream);
}

extern "C" int fn_gru_seq_bwd_f32(const FnGruChain* chains, int n_chains, int B, int T, int H, void* barrier_ws,
                                  size_t barrier_ws_bytes, void* stream) {
    return run(true, chains, n_chains, B, T, H, barrier_ws, barrier_ws_bytes, (cudaStream_t)stream);
}
