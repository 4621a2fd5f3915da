// Latent block: exp/reparameterisation, Gaussian-mixture responsibilities, the categorical-
// Gaussian KL terms, the pairwise latent regulariser and the clip+Adam update.  All of these
// are KB-scale at the reference's shapes (latency-bound); they are warp-shuffle kernels with
// fixed reduction orders so results are run-to-run deterministic.
#include "fn_common.cuh"

namespace {
// supervised component label, clamped to [0, K) (the host mirror counts out-of-range labels and raises)
__device__ __forceinline__ int fn_label(long long y, int K) { return y < 0 ? 0 : y >= K ? K - 1 : (int)y; }


constexpr float kLn2Pi = 1.8378770664093453f;
constexpr int kMaxK = 32;

__global__ void reparam_fwd_kernel(const float* __restrict__ mu, const float* __restrict__ pre,
                                   const float* __restrict__ eps, long long n, float* __restrict__ scale,
                                   float* __restrict__ z) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float s = expf(pre[i]);
    scale[i] = s;
    if (z) z[i] = mu[i] + s * eps[i];
}
__global__ void reparam_bwd_kernel(const float* __restrict__ dz, const float* __restrict__ dmu_in,
                                   const float* __restrict__ dscale_in, const float* __restrict__ eps,
                                   const float* __restrict__ scale, long long n, float* __restrict__ dmu,
                                   float* __restrict__ dpre) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float g = dz ? dz[i] : 0.f;
    if (dmu) dmu[i] = g + (dmu_in ? dmu_in[i] : 0.f);
    dpre[i] = (g * (eps ? eps[i] : 0.f) + (dscale_in ? dscale_in[i] : 0.f)) * scale[i];
}

// ---- approx_qy_x ---------------------------------------------------------------------------
__global__ void qy_fwd_kernel(const float* __restrict__ z, const float* __restrict__ mul, const float* __restrict__ lvl,
                              int B, int Z, int K, float* __restrict__ ll, float* __restrict__ qy,
                              int64_t* __restrict__ y) {
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (b >= B) return;
    float lg[kMaxK];
    const float lnp = logf(1.f / (float)K);
    float mx = -INFINITY;
#pragma unroll 1
    for (int k = 0; k < K; ++k) {
        float s = 0.f;
        for (int d = lane; d < Z; d += 32) {
            const float df = z[(long long)b * Z + d] - mul[k * Z + d], lv = lvl[k * Z + d];
            s += df * df / expf(lv) + lv + kLn2Pi;
        }
        s = fn_warp_sum(s);
        lg[k] = -0.5f * s + lnp;
        mx = fmaxf(mx, lg[k]);
    }
    float den = 0.f;
#pragma unroll 1
    for (int k = 0; k < K; ++k) den += expf(lg[k] - mx);
    if (lane == 0) {
        float best = -INFINITY;
        int bi = 0;
        for (int k = 0; k < K; ++k) {
            const float q = expf(lg[k] - mx) / den;
            ll[(long long)b * K + k] = lg[k];
            qy[(long long)b * K + k] = q;
            if (q > best) { best = q; bi = k; }
        }
        if (y) y[b] = bi;
    }
}

__device__ __forceinline__ float qy_dl(const float* qy, const float* dll, const float* dqy, long long b, int K, int k) {
    float g = dll ? dll[b * K + k] : 0.f;
    if (dqy) {
        float dot = 0.f;
        for (int j = 0; j < K; ++j) dot += dqy[b * K + j] * qy[b * K + j];
        g += qy[b * K + k] * (dqy[b * K + k] - dot);
    }
    return g;
}
__global__ void qy_bwd_dz_kernel(const float* __restrict__ z, const float* __restrict__ mul, const float* __restrict__ lvl,
                                 const float* __restrict__ qy, const float* __restrict__ dll,
                                 const float* __restrict__ dqy, int B, int Z, int K, float* __restrict__ dz) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * Z) return;
    const long long b = i / Z;
    const int d = (int)(i % Z);
    float g = 0.f;
    for (int k = 0; k < K; ++k) g -= qy_dl(qy, dll, dqy, b, K, k) * (z[i] - mul[k * Z + d]) / expf(lvl[k * Z + d]);
    dz[i] = g;
}
__global__ void qy_bwd_dmu_kernel(const float* __restrict__ z, const float* __restrict__ mul,
                                  const float* __restrict__ lvl, const float* __restrict__ qy,
                                  const float* __restrict__ dll, const float* __restrict__ dqy, int B, int Z, int K,
                                  float* __restrict__ dmul) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K * Z) return;
    const int k = i / Z, d = i % Z;
    const float iv = 1.f / expf(lvl[i]), m = mul[i];
    float g = 0.f;
    for (long long b = 0; b < B; ++b) g += qy_dl(qy, dll, dqy, b, K, k) * (z[b * Z + d] - m) * iv;
    dmul[i] = g;
}

// ---- KL block ---------------------------------------------------------------------------------
// single CTA of 32 warps; warp w owns rows w, w+32, ... ; fixed-order final reduction.
__global__ void __launch_bounds__(1024) gm_kl_fwd_kernel(const float* __restrict__ mu, const float* __restrict__ sc,
                                                         const float* __restrict__ mul, const float* __restrict__ lvl,
                                                         const float* __restrict__ qy, const float* __restrict__ ll,
                                                         const int64_t* __restrict__ ylab, int mode, int B, int Z, int K,
                                                         float* __restrict__ out3) {
    __shared__ float acc[3][32];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int b = w; b < B; b += 32) {
        const float* q = qy + (long long)b * K;
        const int k_lo = mode ? fn_label(ylab[b], K) : 0, k_hi = mode ? k_lo + 1 : K;
        for (int k = k_lo; k < k_hi; ++k) {
            float s = 0.f;
            for (int d = lane; d < Z; d += 32) {
                const float sp = expf(lvl[k * Z + d]), sq = sc[(long long)b * Z + d];
                const float ratio = sq / sp, rho = ratio * ratio;
                const float df = (mu[(long long)b * Z + d] - mul[k * Z + d]) / sp;
                s += 0.5f * (rho + df * df - 1.f - logf(rho));
            }
            s = fn_warp_sum(s) / (float)Z;
            a0 += mode ? s : s * q[k];
        }
        if (mode == 0) {
            const float* l = ll + (long long)b * K;
            float mx = -INFINITY, den = 0.f, e = 0.f;
            for (int k = 0; k < K; ++k) mx = fmaxf(mx, l[k]);
            for (int k = 0; k < K; ++k) den += expf(l[k] - mx);
            const float lse = mx + logf(den);
            for (int k = 0; k < K; ++k) e += q[k] * (l[k] - lse);
            a1 += e / (float)K;
        } else {
            float mx = -INFINITY, den = 0.f;
            for (int k = 0; k < K; ++k) mx = fmaxf(mx, q[k]);
            for (int k = 0; k < K; ++k) den += expf(q[k] - mx);
            a2 += (mx + logf(den)) - q[fn_label(ylab[b], K)];
        }
    }
    if (lane == 0) { acc[0][w] = a0; acc[1][w] = a1; acc[2][w] = a2; }
    __syncthreads();
    if (threadIdx.x < 3) {
        float s = 0.f;
        for (int i = 0; i < 32; ++i) s += acc[threadIdx.x][i];
        s /= (float)B;
        if (threadIdx.x == 1 && mode == 0) s -= logf(1.f / (float)K);
        if (threadIdx.x == 1 && mode != 0) s = 0.f;
        if (threadIdx.x == 2 && mode == 0) s = 0.f;
        out3[threadIdx.x] = s;
    }
}

__global__ void gm_kl_bwd_rows_kernel(const float* __restrict__ mu, const float* __restrict__ sc,
                                      const float* __restrict__ mul, const float* __restrict__ lvl,
                                      const float* __restrict__ qy, const float* __restrict__ ll,
                                      const int64_t* __restrict__ ylab, int mode, int B, int Z, int K,
                                      const float* __restrict__ dout3, float* __restrict__ dmu, float* __restrict__ dsc,
                                      float* __restrict__ dqy, float* __restrict__ dll) {
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (b >= B) return;
    const float d0 = dout3[0] / (float)B, d1 = dout3[1] / (float)B, d2 = dout3[2] / (float)B;
    const float* q = qy + (long long)b * K;
    const int k_lo = mode ? fn_label(ylab[b], K) : 0, k_hi = mode ? k_lo + 1 : K;
    // per-dim grads
    for (int d = lane; d < Z; d += 32) {
        const float m = mu[(long long)b * Z + d], sq = sc[(long long)b * Z + d];
        float gm = 0.f, gs = 0.f;
        for (int k = k_lo; k < k_hi; ++k) {
            const float sp = expf(lvl[k * Z + d]), ip2 = 1.f / (sp * sp);
            const float wgt = (mode ? 1.f : q[k]) * d0 / (float)Z;
            gm += wgt * (m - mul[k * Z + d]) * ip2;
            gs += wgt * (sq * ip2 - 1.f / sq);
        }
        dmu[(long long)b * Z + d] = gm;
        dsc[(long long)b * Z + d] = gs;
    }
    // per-component grads
    float lse = 0.f, sumq = 0.f, mxq = -INFINITY, denq = 0.f;
    if (mode == 0) {
        const float* l = ll + (long long)b * K;
        float mx = -INFINITY, den = 0.f;
        for (int k = 0; k < K; ++k) mx = fmaxf(mx, l[k]);
        for (int k = 0; k < K; ++k) den += expf(l[k] - mx);
        lse = mx + logf(den);
        for (int k = 0; k < K; ++k) sumq += q[k];
    } else {
        for (int k = 0; k < K; ++k) mxq = fmaxf(mxq, q[k]);
        for (int k = 0; k < K; ++k) denq += expf(q[k] - mxq);
    }
    for (int k = 0; k < K; ++k) {
        float gq = 0.f, gl = 0.f;
        if (mode == 0) {
            float s = 0.f;
            for (int d = lane; d < Z; d += 32) {
                const float sp = expf(lvl[k * Z + d]), sq = sc[(long long)b * Z + d];
                const float ratio = sq / sp, rho = ratio * ratio;
                const float df = (mu[(long long)b * Z + d] - mul[k * Z + d]) / sp;
                s += 0.5f * (rho + df * df - 1.f - logf(rho));
            }
            s = fn_warp_sum(s) / (float)Z;
            const float lk = ll[(long long)b * K + k] - lse;
            gq = d0 * s + d1 * lk / (float)K;
            gl = d1 / (float)K * (q[k] - expf(lk) * sumq);
        } else {
            gq = d2 * (expf(q[k] - mxq) / denq - (k == fn_label(ylab[b], K) ? 1.f : 0.f));
        }
        if (lane == 0) {
            dqy[(long long)b * K + k] = gq;
            dll[(long long)b * K + k] = gl;
        }
    }
}
__global__ void gm_kl_bwd_lookup_kernel(const float* __restrict__ mu, const float* __restrict__ mul,
                                        const float* __restrict__ lvl, const float* __restrict__ qy,
                                        const int64_t* __restrict__ ylab, int mode, int B, int Z, int K,
                                        const float* __restrict__ dout3, float* __restrict__ dmul) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K * Z) return;
    const int k = i / Z, d = i % Z;
    const float sp = expf(lvl[i]), ip2 = 1.f / (sp * sp), m = mul[i];
    const float d0 = dout3[0] / ((float)B * (float)Z);
    float g = 0.f;
    for (long long b = 0; b < B; ++b) {
        const float wgt = mode ? (fn_label(ylab[b], K) == k ? 1.f : 0.f) : qy[b * K + k];
        g -= wgt * (mu[b * Z + d] - m) * ip2;
    }
    dmul[i] = g * d0;
}

__global__ void __launch_bounds__(1024) std_kl_fwd_kernel(const float* __restrict__ mu, const float* __restrict__ sc,
                                                          long long n, float* __restrict__ out) {
    __shared__ float red[33];
    float s = 0.f;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        const float v = sc[i] * sc[i], m = mu[i];
        s += 0.5f * (v + m * m - 1.f - logf(v));
    }
    s = fn_block_sum(s, red);
    if (threadIdx.x == 0) out[0] = s / (float)n;
}
__global__ void std_kl_bwd_kernel(const float* __restrict__ mu, const float* __restrict__ sc, long long n,
                                  const float* __restrict__ dout, float* __restrict__ dmu, float* __restrict__ dsc) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float g = dout[0] / (float)n;
    dmu[i] = g * mu[i];
    dsc[i] = g * (sc[i] - 1.f / sc[i]);
}

// ---- pairwise latent regulariser --------------------------------------------------------------
__global__ void latent_reg_rows_kernel(const float* __restrict__ z, long long z_ld, const double* __restrict__ attr,
                                       int B, float* __restrict__ rowsum, float* __restrict__ dz0) {
    __shared__ float red[33];
    const int i = blockIdx.x;
    const float zi = z[(long long)i * z_ld];
    const double ai = attr[i];
    float s = 0.f, g = 0.f;
    for (int j = threadIdx.x; j < B; j += blockDim.x) {
        const float th = tanhf(zi - z[(long long)j * z_ld]);
        const float df = (float)(ai - attr[j]);                     // float64 difference cast to float, as the reference
        const float sg = (df > 0.f) ? 1.f : ((df < 0.f) ? -1.f : 0.f);
        const float e = th - sg;
        s += e * e;
        g += e * (1.f - th * th);
    }
    s = fn_block_sum(s, red);
    g = fn_block_sum(g, red);
    if (threadIdx.x == 0) {
        rowsum[i] = s;
        dz0[i] = 4.f * g / ((float)B * (float)B);
    }
}
__global__ void latent_reg_final_kernel(const float* __restrict__ rowsum, int B, float* __restrict__ loss) {
    __shared__ float red[33];
    float s = 0.f;
    for (int i = threadIdx.x; i < B; i += blockDim.x) s += rowsum[i];
    s = fn_block_sum(s, red);
    if (threadIdx.x == 0) loss[0] = s / ((float)B * (float)B);
}
__global__ void latent_reg_bwd_kernel(const float* __restrict__ dz0, const float* __restrict__ dloss, int B, int Z,
                                      float* __restrict__ dz) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * Z) return;
    dz[i] = (i % Z == 0) ? dloss[0] * dz0[i / Z] : 0.f;
}

__global__ void clip_adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                 float* __restrict__ v, long long n, const float* __restrict__ norm, float max_norm,
                                 float step_size, float beta1, float beta2, float eps, float inv_sqrt_bc2) {
    const float coef = fminf(max_norm / (norm[0] + 1e-6f), 1.f);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float gi = g[i] * coef;
        const float mi = beta1 * m[i] + (1.f - beta1) * gi;
        const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        p[i] -= step_size * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
    }
}


// =================================================================================================
// Streaming ("fast") variants: Z % 4 == 0, Z <= 512, K * Z <= kFastKZ.  One warp per row with float4 loads (a
// 128-wide row is ONE 512-byte warp transaction), grid-stride over rows, every row-invariant term of the mixture
// (component means, 1/scale, 1/variance, log-normaliser) hoisted into shared memory once per CTA, reductions in two
// fixed-order stages through a caller-owned scratch buffer: bandwidth-bound at large B, deterministic at any B.
// =================================================================================================
constexpr int kFastKZ = 2048;
constexpr int kRowThreads = 256, kRowWarps = kRowThreads / 32;

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ld4s(const float* p) { return *reinterpret_cast<const float4*>(p); }

// fixed-order sum of per-CTA partials: out[j] = post(sum_g partial[g][j]) for j < ncols
template <int MODE>   // 0: plain float columns (lookup-table gradients); 1: the 3 KL scalars; 2: std KL scalar
__global__ void latent_stage2_kernel(const double* __restrict__ partial, int nblk, int ncols, float* __restrict__ out,
                                     double inv_b, int mode, int K) {
    __shared__ double red[33];
    if (MODE == 0) {
        const int j = blockIdx.x * blockDim.x + threadIdx.x;
        if (j >= ncols) return;
        double s = 0.0;
        for (int g = 0; g < nblk; ++g) s += partial[(long long)g * ncols + j];
        out[j] = (float)s;
    } else {
        for (int j = 0; j < ncols; ++j) {
            double s = 0.0;
            for (int g = threadIdx.x; g < nblk; g += blockDim.x) s += partial[(long long)g * ncols + j];
            s = fn_block_sum_d(s, red);
            if (threadIdx.x == 0) {
                float v = (float)(s * inv_b);
                if (MODE == 1) {
                    if (j == 1) v = mode == 0 ? v - logf(1.f / (float)K) : 0.f;
                    if (j == 2 && mode == 0) v = 0.f;
                }
                out[j] = v;
            }
            __syncthreads();
        }
    }
}

// shared-memory tables of a mixture: s_mu[k*Z+d], s_is[k*Z+d] = exp(-logvar_k) -- the reference uses exp(logvar) as
// the VARIANCE in approx_qy_x and as the SCALE in the KL (sic), so one table serves both -- and s_lv[k*Z+d] = logvar_k
__device__ __forceinline__ void load_tables(const float* mul, const float* lvl, int KZ, float* s_mu, float* s_is, float* s_lv) {
    for (int i = threadIdx.x; i < KZ; i += blockDim.x) {
        const float lv = lvl[i];
        s_mu[i] = mul[i];
        s_lv[i] = lv;
        s_is[i] = 1.f / expf(lv);
    }
    __syncthreads();
}

// Row kernels.  LPR lanes share a row (32 / LPR rows per warp pass): LPR = 8 for Z <= 128 -- four rows per warp, 3-step
// shuffles, the per-row scalar work (soft-max over K, class terms) done for four rows at once -- and LPR = 32 for wider
// rows.  Lane `sl` of a row owns the float4 chunks sl + LPR*j, j < NJ.  Two passes are in flight per warp.  These
// kernels are instruction-issue bound before they are bandwidth bound (ncu: 57-70 % issue slots at 1-2 TB/s with one
// row per warp), hence fast-math logarithms on the elementwise part and no redundant per-row work.
// KT > 0: K == KT known at compile time (component loops unrolled, per-component values in registers); KT = 0: any K.
template <int LPR>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <int LPR, int NJ>
__device__ __forceinline__ void load_chunks(const float* x, long long b, int Z, int sl, float4 (&r)[NJ]) {
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const int d = (sl + LPR * j) * 4;
        if (d < Z) r[j] = ld4(x + b * Z + d);
    }
}

template <int KT, int LPR, int NJ>
__device__ __forceinline__ void qy_row(const float4 (&zr)[NJ], long long b, bool live, int Z, int K, int sl, const float* s_mu,
                                       const float* s_iv, const float* s_c, float lnp, float* __restrict__ ll,
                                       float* __restrict__ qy, int64_t* __restrict__ y) {
    float lg[KT > 0 ? KT : kMaxK];
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < (KT > 0 ? KT : K); ++k) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int d = (sl + LPR * j) * 4;
            if (d < Z) {
                const float4 m = ld4s(s_mu + k * Z + d), iv = ld4s(s_iv + k * Z + d);
                float df;
                df = zr[j].x - m.x; s += df * df * iv.x;
                df = zr[j].y - m.y; s += df * df * iv.y;
                df = zr[j].z - m.z; s += df * df * iv.z;
                df = zr[j].w - m.w; s += df * df * iv.w;
            }
        }
        s = group_sum<LPR>(s);
        lg[k] = -0.5f * (s + s_c[k]) + lnp;
        mx = fmaxf(mx, lg[k]);
    }
    if (sl == 0 && live) {
        float den = 0.f, e[KT > 0 ? KT : kMaxK];
#pragma unroll
        for (int k = 0; k < (KT > 0 ? KT : K); ++k) { e[k] = expf(lg[k] - mx); den += e[k]; }
        const float inv = 1.f / den;
        float best = -INFINITY;
        int bi = 0;
#pragma unroll
        for (int k = 0; k < (KT > 0 ? KT : K); ++k) {
            const float q = e[k] * inv;
            ll[b * K + k] = lg[k];
            qy[b * K + k] = q;
            if (q > best) { best = q; bi = k; }
        }
        if (y) y[b] = bi;
    }
}
template <int KT, int LPR, int NJ>
__global__ void __launch_bounds__(kRowThreads) qy_fwd_fast_kernel(const float* __restrict__ z, const float* __restrict__ mul,
                                                                  const float* __restrict__ lvl, int B, int Z, int K,
                                                                  float* __restrict__ ll, float* __restrict__ qy,
                                                                  int64_t* __restrict__ y) {
    extern __shared__ float sm_f[];
    float* s_mu = sm_f; float* s_iv = sm_f + K * Z; float* s_lv = s_iv + K * Z;
    __shared__ float s_c[kMaxK];                         // sum_d (logvar_k[d] + ln 2 pi)
    load_tables(mul, lvl, K * Z, s_mu, s_iv, s_lv);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int k = w; k < K; k += kRowWarps) {
        float c = 0.f;
        for (int d = lane; d < Z; d += 32) c += s_lv[k * Z + d] + kLn2Pi;
        c = fn_warp_sum(c);
        if (lane == 0) s_c[k] = c;
    }
    __syncthreads();
    constexpr int RPW = 32 / LPR;
    const int sl = lane % LPR, rs = lane / LPR;
    const float lnp = logf(1.f / (float)K);
    const long long stride = (long long)gridDim.x * kRowWarps * RPW;
    for (long long r0 = ((long long)blockIdx.x * kRowWarps + w) * RPW; r0 < B; r0 += 2 * stride) {   // r0 is warp-uniform
        const long long b0 = r0 + rs, b1 = b0 + stride;
        const bool l0 = b0 < B, l1 = b1 < B;
        const long long c0 = l0 ? b0 : B - 1, c1 = l1 ? b1 : B - 1;                    // dead lanes shadow the last row
        float4 z0[NJ], z1[NJ];
        load_chunks<LPR, NJ>(z, c0, Z, sl, z0);
        load_chunks<LPR, NJ>(z, c1, Z, sl, z1);
        qy_row<KT, LPR, NJ>(z0, c0, l0, Z, K, sl, s_mu, s_iv, s_c, lnp, ll, qy, y);
        if (r0 + stride < B) qy_row<KT, LPR, NJ>(z1, c1, l1, Z, K, sl, s_mu, s_iv, s_c, lnp, ll, qy, y);
    }
}

// mean over Z of KL(N(mu, sq) || N(mu_k, sp_k)) of one row against ALL components (only_k < 0) or one (only_k); every
// lane of the row gets the results.  Chunk-outer / component-inner: the row is touched once, a handful of registers.
// log(rho) = 2 (log sq - log sp): ONE logarithm per element instead of one per element and component.
template <int KT, int LPR, int NJ>
__device__ __forceinline__ void row_kl_all(const float4 (&mr)[NJ], const float4 (&sr)[NJ], int Z, int K, int sl, const float* s_mu,
                                           const float* s_is, const float* s_lv, float inv_z, int only_k,
                                           float (&out)[KT > 0 ? KT : kMaxK]) {
    constexpr int KA = KT > 0 ? KT : kMaxK;
#pragma unroll
    for (int k = 0; k < KA; ++k) out[k] = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const int d = (sl + LPR * j) * 4;
        if (d < Z) {
            const float4 lq = make_float4(__logf(sr[j].x), __logf(sr[j].y), __logf(sr[j].z), __logf(sr[j].w));
            auto one = [&](int k) {
                const float4 m = ld4s(s_mu + k * Z + d), is = ld4s(s_is + k * Z + d), lv = ld4s(s_lv + k * Z + d);
                float r, df, t = 0.f;
                r = sr[j].x * is.x; df = (mr[j].x - m.x) * is.x; t += r * r + df * df - 1.f - 2.f * (lq.x - lv.x);
                r = sr[j].y * is.y; df = (mr[j].y - m.y) * is.y; t += r * r + df * df - 1.f - 2.f * (lq.y - lv.y);
                r = sr[j].z * is.z; df = (mr[j].z - m.z) * is.z; t += r * r + df * df - 1.f - 2.f * (lq.z - lv.z);
                r = sr[j].w * is.w; df = (mr[j].w - m.w) * is.w; t += r * r + df * df - 1.f - 2.f * (lq.w - lv.w);
                return t;
            };
            if (only_k >= 0) {
                out[0] += one(only_k);
            } else {
#pragma unroll
                for (int k = 0; k < (KT > 0 ? KT : K); ++k) out[k] += one(k);
            }
        }
    }
    if (only_k >= 0) {
        out[0] = group_sum<LPR>(out[0]) * (0.5f * inv_z);
    } else {
#pragma unroll
        for (int k = 0; k < (KT > 0 ? KT : K); ++k) out[k] = group_sum<LPR>(out[k]) * (0.5f * inv_z);
    }
}

template <int KT, int LPR, int NJ>
__global__ void __launch_bounds__(kRowThreads, 3) gm_kl_fwd_fast_kernel(const float* __restrict__ mu, const float* __restrict__ sc,
                                                                        const float* __restrict__ mul, const float* __restrict__ lvl,
                                                                        const float* __restrict__ qy, const float* __restrict__ ll,
                                                                        const int64_t* __restrict__ ylab, int mode, int B, int Z, int K,
                                                                        double* __restrict__ partial) {
    extern __shared__ float sm_f[];
    float* s_mu = sm_f; float* s_is = sm_f + K * Z; float* s_lv = s_is + K * Z;
    __shared__ float acc[3][kRowWarps];
    load_tables(mul, lvl, K * Z, s_mu, s_is, s_lv);
    constexpr int RPW = 32 / LPR, KA = KT > 0 ? KT : kMaxK;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, sl = lane % LPR, rs = lane / LPR;
    const float inv_z = 1.f / (float)Z, inv_k = 1.f / (float)K;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;                   // per ROW SLOT (valid in lane sl == 0 of the slot)
    const long long stride = (long long)gridDim.x * kRowWarps * RPW;
    for (long long r0 = ((long long)blockIdx.x * kRowWarps + w) * RPW; r0 < B; r0 += stride) {     // r0 is warp-uniform
        const bool live = r0 + rs < B;
        const long long b = live ? r0 + rs : B - 1;        // dead lanes shadow the last row (the shuffles need them)
        float4 mr[NJ], sr[NJ];
        load_chunks<LPR, NJ>(mu, b, Z, sl, mr); load_chunks<LPR, NJ>(sc, b, Z, sl, sr);
        // per-row scalars fetched with the row, before any arithmetic (a load in the middle of the dependent chain costs
        // a memory round trip per row)
        float q[KA], l[KA];
        const int yl = mode ? fn_label(ylab[b], K) : -1;
#pragma unroll
        for (int k = 0; k < (KT > 0 ? KT : K); ++k) { q[k] = __ldg(qy + b * K + k); l[k] = mode == 0 ? __ldg(ll + b * K + k) : 0.f; }
        float kl[KA];
        row_kl_all<KT, LPR, NJ>(mr, sr, Z, K, sl, s_mu, s_is, s_lv, inv_z, yl, kl);
        if (sl == 0 && live) {
            if (mode == 0) {
                float t = 0.f, mx = -INFINITY, den = 0.f, e = 0.f;
#pragma unroll
                for (int k = 0; k < (KT > 0 ? KT : K); ++k) { t += kl[k] * q[k]; mx = fmaxf(mx, l[k]); }
#pragma unroll
                for (int k = 0; k < (KT > 0 ? KT : K); ++k) den += expf(l[k] - mx);
                const float lse = mx + logf(den);
#pragma unroll
                for (int k = 0; k < (KT > 0 ? KT : K); ++k) e += q[k] * (l[k] - lse);
                a0 += t;
                a1 += e * inv_k;
            } else {
                float mx = -INFINITY, den = 0.f, qyl = 0.f;
#pragma unroll
                for (int k = 0; k < (KT > 0 ? KT : K); ++k) { mx = fmaxf(mx, q[k]); qyl = (k == yl) ? q[k] : qyl; }
#pragma unroll
                for (int k = 0; k < (KT > 0 ? KT : K); ++k) den += expf(q[k] - mx);
                a0 += kl[0];
                a2 += (mx + logf(den)) - qyl;
            }
        }
    }
    // fixed-order reduction: row slots of a warp, then warps of the CTA
    if (sl != 0) { a0 = 0.f; a1 = 0.f; a2 = 0.f; }
    a0 = fn_warp_sum(a0); a1 = fn_warp_sum(a1); a2 = fn_warp_sum(a2);
    if (lane == 0) { acc[0][w] = a0; acc[1][w] = a1; acc[2][w] = a2; }
    __syncthreads();
    if (threadIdx.x < 3) {
        double s = 0.0;
        for (int i = 0; i < kRowWarps; ++i) s += (double)acc[threadIdx.x][i];
        partial[(long long)blockIdx.x * 3 + threadIdx.x] = s;
    }
}

template <int KT, int LPR, int NJ>
__global__ void __launch_bounds__(kRowThreads, 3) gm_kl_bwd_rows_fast_kernel(
    const float* __restrict__ mu, const float* __restrict__ sc, const float* __restrict__ mul, const float* __restrict__ lvl,
    const float* __restrict__ qy, const float* __restrict__ ll, const int64_t* __restrict__ ylab, int mode, int B, int Z, int K,
    const float* __restrict__ dout3, float* __restrict__ dmu, float* __restrict__ dsc, float* __restrict__ dqy,
    float* __restrict__ dll) {
    extern __shared__ float sm_f[];
    float* s_mu = sm_f; float* s_is = sm_f + K * Z; float* s_lv = s_is + K * Z;
    load_tables(mul, lvl, K * Z, s_mu, s_is, s_lv);
    constexpr int RPW = 32 / LPR, KA = KT > 0 ? KT : kMaxK;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, sl = lane % LPR, rs = lane / LPR;
    const float inv_z = 1.f / (float)Z, inv_k = 1.f / (float)K;
    const float d0 = dout3[0] / (float)B, d1 = dout3[1] / (float)B, d2 = dout3[2] / (float)B;
    const long long stride = (long long)gridDim.x * kRowWarps * RPW;
    for (long long r0 = ((long long)blockIdx.x * kRowWarps + w) * RPW; r0 < B; r0 += stride) {
        const bool live = r0 + rs < B;
        const long long b = live ? r0 + rs : B - 1;
        float4 mr[NJ], sr[NJ];
        load_chunks<LPR, NJ>(mu, b, Z, sl, mr); load_chunks<LPR, NJ>(sc, b, Z, sl, sr);
        float q[KA], l[KA];
        const int yl = mode ? fn_label(ylab[b], K) : -1;
#pragma unroll
        for (int k = 0; k < (KT > 0 ? KT : K); ++k) { q[k] = __ldg(qy + b * K + k); l[k] = mode == 0 ? __ldg(ll + b * K + k) : 0.f; }
        // per-dimension gradients
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int d = (sl + LPR * j) * 4;
            if (d < Z) {
                float4 gm = make_float4(0.f, 0.f, 0.f, 0.f), gs = gm;
                const float4 ri = make_float4(__frcp_rn(sr[j].x), __frcp_rn(sr[j].y), __frcp_rn(sr[j].z), __frcp_rn(sr[j].w));
                auto one = [&](int k, float wgt) {
                    const float4 m = ld4s(s_mu + k * Z + d), is = ld4s(s_is + k * Z + d);
                    float i2;
                    i2 = is.x * is.x; gm.x += wgt * (mr[j].x - m.x) * i2; gs.x += wgt * (sr[j].x * i2 - ri.x);
                    i2 = is.y * is.y; gm.y += wgt * (mr[j].y - m.y) * i2; gs.y += wgt * (sr[j].y * i2 - ri.y);
                    i2 = is.z * is.z; gm.z += wgt * (mr[j].z - m.z) * i2; gs.z += wgt * (sr[j].z * i2 - ri.z);
                    i2 = is.w * is.w; gm.w += wgt * (mr[j].w - m.w) * i2; gs.w += wgt * (sr[j].w * i2 - ri.w);
                };
                if (mode) {
                    one(yl, d0 * inv_z);
                } else {
#pragma unroll
                    for (int k = 0; k < (KT > 0 ? KT : K); ++k) one(k, q[k] * d0 * inv_z);
                }
                if (live) {
                    *reinterpret_cast<float4*>(dmu + b * Z + d) = gm;
                    *reinterpret_cast<float4*>(dsc + b * Z + d) = gs;
                }
            }
        }
        // per-component gradients
        if (mode == 0) {
            float kl[KA];
            row_kl_all<KT, LPR, NJ>(mr, sr, Z, K, sl, s_mu, s_is, s_lv, inv_z, -1, kl);
            if (sl == 0 && live) {
                float mx = -INFINITY, den = 0.f, sumq = 0.f;
#pragma unroll
                for (int k = 0; k < (KT > 0 ? KT : K); ++k) { mx = fmaxf(mx, l[k]); sumq += q[k]; }
#pragma unroll
                for (int k = 0; k < (KT > 0 ? KT : K); ++k) den += expf(l[k] - mx);
                const float lse = mx + logf(den);
#pragma unroll
                for (int k = 0; k < (KT > 0 ? KT : K); ++k) {
                    const float lk = l[k] - lse;
                    dqy[b * K + k] = d0 * kl[k] + d1 * lk * inv_k;
                    dll[b * K + k] = d1 * inv_k * (q[k] - expf(lk) * sumq);
                }
            }
        } else if (sl == 0 && live) {
            float mxq = -INFINITY, denq = 0.f;
#pragma unroll
            for (int k = 0; k < (KT > 0 ? KT : K); ++k) mxq = fmaxf(mxq, q[k]);
#pragma unroll
            for (int k = 0; k < (KT > 0 ? KT : K); ++k) denq += expf(q[k] - mxq);
#pragma unroll
            for (int k = 0; k < (KT > 0 ? KT : K); ++k) {
                dqy[b * K + k] = d2 * (expf(q[k] - mxq) / denq - (k == yl ? 1.f : 0.f));
                dll[b * K + k] = 0.f;
            }
        }
    }
}

// Lookup-table gradients: column sums over the batch.  CTA g owns the rows [g*chunk, (g+1)*chunk); 4 row groups of
// KZ/4 (<= 64... any) float4 columns each; partial[g][k*Z+d] in double.
// WHICH 0: gm_kl (dmul = -d0/(B Z) * sum_b w(b,k) (mu - m) / sp^2);  WHICH 1: qy (dmul = sum_b dl(b,k) (z - m) / var)
template <int WHICH>
__global__ void __launch_bounds__(256) lookup_grad_stage1_kernel(const float* __restrict__ x, const float* __restrict__ mul,
                                                                 const float* __restrict__ lvl, const float* __restrict__ qy,
                                                                 const float* __restrict__ dll, const float* __restrict__ dqy,
                                                                 const int64_t* __restrict__ ylab, int mode, int B, int Z, int K,
                                                                 int chunk, double* __restrict__ partial) {
    extern __shared__ float sm_f[];                      // [4 groups][K*Z] partial sums
    const int KZ = K * Z, nq = KZ / 4;                   // float4 columns
    const long long r0 = (long long)blockIdx.x * chunk, r1 = min((long long)B, r0 + chunk);
    for (int i = threadIdx.x; i < 4 * KZ; i += blockDim.x) sm_f[i] = 0.f;
    __syncthreads();
    const int grp = threadIdx.x >> 6, t = threadIdx.x & 63;
    for (int c = t; c < nq; c += 64) {
        const int k = (c * 4) / Z, d = (c * 4) % Z;
        const float4 m = ld4(mul + k * Z + d);
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        for (long long b = r0 + grp; b < r1; b += 4) {
            float wgt;
            if (WHICH == 0) wgt = mode ? (fn_label(ylab[b], K) == k ? 1.f : 0.f) : qy[b * K + k];
            else wgt = qy_dl(qy, dll, dqy, b, K, k);
            const float4 v = ld4(x + b * Z + d);
            g.x += wgt * (v.x - m.x); g.y += wgt * (v.y - m.y); g.z += wgt * (v.z - m.z); g.w += wgt * (v.w - m.w);
        }
        *reinterpret_cast<float4*>(sm_f + grp * KZ + c * 4) = g;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < KZ; i += blockDim.x)
        partial[(long long)blockIdx.x * KZ + i] = (double)sm_f[i] + (double)sm_f[KZ + i] + (double)sm_f[2 * KZ + i] + (double)sm_f[3 * KZ + i];
}
// out[i] = scale_i * sum_g partial[g][i]
template <int WHICH>
__global__ void lookup_grad_stage2_kernel(const double* __restrict__ partial, int nblk, const float* __restrict__ lvl,
                                          const float* __restrict__ dout3, int B, int Z, int K, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K * Z) return;
    double s = 0.0;
    for (int g = 0; g < nblk; ++g) s += partial[(long long)g * K * Z + i];
    const float e = expf(lvl[i]);
    if (WHICH == 0) out[i] = (float)(-s) * (1.f / (e * e)) * (dout3[0] / ((float)B * (float)Z));
    else out[i] = (float)s * (1.f / e);
}

__global__ void __launch_bounds__(256) std_kl_fwd_fast_kernel(const float* __restrict__ mu, const float* __restrict__ sc, long long n4,
                                                              double* __restrict__ partial) {
    __shared__ double red[33];
    double s = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 m = ld4(mu + 4 * i), c = ld4(sc + 4 * i);
        float a = 0.f, v;
        v = c.x * c.x; a += 0.5f * (v + m.x * m.x - 1.f - logf(v));
        v = c.y * c.y; a += 0.5f * (v + m.y * m.y - 1.f - logf(v));
        v = c.z * c.z; a += 0.5f * (v + m.z * m.z - 1.f - logf(v));
        v = c.w * c.w; a += 0.5f * (v + m.w * m.w - 1.f - logf(v));
        s += (double)a;
    }
    s = fn_block_sum_d(s, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// elementwise float4 variants
__global__ void reparam_fwd4_kernel(const float4* __restrict__ mu, const float4* __restrict__ pre, const float4* __restrict__ eps,
                                    long long n4, float4* __restrict__ scale, float4* __restrict__ z) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const float4 p = __ldg(pre + i);
    const float4 s = make_float4(expf(p.x), expf(p.y), expf(p.z), expf(p.w));
    scale[i] = s;
    if (z) {
        const float4 m = __ldg(mu + i), e = __ldg(eps + i);
        z[i] = make_float4(m.x + s.x * e.x, m.y + s.y * e.y, m.z + s.z * e.z, m.w + s.w * e.w);
    }
}
__global__ void clip_adam4_kernel(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m, float4* __restrict__ v,
                                  long long n4, const float* __restrict__ norm, float max_norm, float step_size, float beta1,
                                  float beta2, float eps, float inv_sqrt_bc2) {
    const float coef = fminf(max_norm / (norm[0] + 1e-6f), 1.f);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 g4 = __ldg(g + i);
        float4 m4 = m[i], v4 = v[i], p4 = p[i];
#define FN_ADAM1(c)                                                        \
        {                                                                  \
            const float gi = g4.c * coef;                                  \
            m4.c = beta1 * m4.c + (1.f - beta1) * gi;                      \
            v4.c = beta2 * v4.c + (1.f - beta2) * gi * gi;                 \
            p4.c -= step_size * m4.c / (sqrtf(v4.c) * inv_sqrt_bc2 + eps); \
        }
        FN_ADAM1(x) FN_ADAM1(y) FN_ADAM1(z) FN_ADAM1(w)
#undef FN_ADAM1
        m[i] = m4; v[i] = v4; p[i] = p4;
    }
}

__host__ bool fast_ok(int Z, int K) { return Z % 4 == 0 && Z <= 512 && K * Z <= kFastKZ && K <= kMaxK; }
__host__ bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }   // NULL counts as aligned
__host__ int row_grid(long long B) {
    int sms = fn_num_sms();
    if (sms <= 0) sms = 148;
    const long long need = (B + kRowWarps - 1) / kRowWarps;
    return (int)(need < (long long)sms * 8 ? need : (long long)sms * 8);
}
__host__ int col_grid(long long B) {          // CTAs of the lookup-gradient stage 1 (>= 64 rows each)
    int sms = fn_num_sms();
    if (sms <= 0) sms = 148;
    const long long need = (B + 63) / 64;
    return (int)(need < (long long)sms * 4 ? need : (long long)sms * 4);
}

}  // namespace

#define EW_GRID(n) fn_cdiv((n), 256), 256, 0, (cudaStream_t)stream

extern "C" size_t fn_latent_scratch_bytes(int B, int Z, int K) {
    if (B <= 0 || Z <= 0 || K <= 0) return 0;
    const size_t a = (size_t)row_grid(B) * 3 * sizeof(double), c = (size_t)col_grid(B) * K * Z * sizeof(double);
    const size_t d = (size_t)1184 * sizeof(double);              // std KL partials
    size_t m = a > c ? a : c;
    return m > d ? m : d;
}
extern "C" int fn_reparam_fwd(const float* mu, const float* pre_scale, const float* eps, long long n, float* scale,
                              float* z, void* stream) {
    FN_REQUIRE(pre_scale && scale && n > 0 && (!z || (mu && eps)), "fn_reparam_fwd: bad args");
    if (n % 4 == 0 && aligned16(pre_scale) && aligned16(scale) && aligned16(mu) && aligned16(eps) && aligned16(z))
        reparam_fwd4_kernel<<<EW_GRID(n / 4)>>>((const float4*)mu, (const float4*)pre_scale, (const float4*)eps, n / 4,
                                                (float4*)scale, (float4*)z);
    else
        reparam_fwd_kernel<<<EW_GRID(n)>>>(mu, pre_scale, eps, n, scale, z);
    FN_LAUNCH_CHECK();
    return FN_OK;
}
extern "C" int fn_reparam_bwd(const float* dz, const float* dmu_in, const float* dscale_in, const float* eps,
                              const float* scale, long long n, float* dmu, float* dpre, void* stream) {
    FN_REQUIRE(scale && dpre && n > 0 && (!dz || eps), "fn_reparam_bwd: bad args");
    reparam_bwd_kernel<<<EW_GRID(n)>>>(dz, dmu_in, dscale_in, eps, scale, n, dmu, dpre);
    FN_LAUNCH_CHECK();
    return FN_OK;
}
extern "C" int fn_qy_x_fwd(const float* z, const float* mu_lookup, const float* logvar_lookup, int B, int Z, int K,
                           float* logLogit, float* qy, int64_t* y, void* stream) {
    FN_REQUIRE(z && mu_lookup && logvar_lookup && logLogit && qy && B > 0 && Z > 0, "fn_qy_x_fwd: bad args");
    FN_REQUIRE(K >= 1 && K <= kMaxK, "fn_qy_x_fwd: K=%d outside [1,%d]", K, kMaxK);
    if (fast_ok(Z, K) && aligned16(z))
    {
#define FN_QY(KT) if (Z <= 128) FN_QY2(KT, 8, 4); else FN_QY2(KT, 32, 4)
#define FN_QY2(KT, LPR, NJ) qy_fwd_fast_kernel<KT, LPR, NJ><<<row_grid(B), kRowThreads, (size_t)3 * K * Z * sizeof(float), (cudaStream_t)stream>>>( \
        z, mu_lookup, logvar_lookup, B, Z, K, logLogit, qy, y)
        if (K == 1) { FN_QY(1); } else if (K == 2) { FN_QY(2); } else if (K == 4) { FN_QY(4); } else { FN_QY(0); }
#undef FN_QY
#undef FN_QY2
    }
    else
        qy_fwd_kernel<<<fn_cdiv((long long)B * 32, 128), 128, 0, (cudaStream_t)stream>>>(z, mu_lookup, logvar_lookup, B, Z, K, logLogit, qy, y);
    FN_LAUNCH_CHECK();
    return FN_OK;
}
extern "C" int fn_qy_x_bwd(const float* z, const float* mu_lookup, const float* logvar_lookup, const float* qy,
                           const float* dlogLogit, const float* dqy, int B, int Z, int K, float* dz, float* dmu_lookup,
                           void* scratch, size_t scratch_bytes, void* stream) {
    FN_REQUIRE(z && mu_lookup && logvar_lookup && qy && dz && dmu_lookup && B > 0 && Z > 0 && K >= 1, "fn_qy_x_bwd: bad args");
    qy_bwd_dz_kernel<<<EW_GRID((long long)B * Z)>>>(z, mu_lookup, logvar_lookup, qy, dlogLogit, dqy, B, Z, K, dz);
    FN_LAUNCH_CHECK();
    if (fast_ok(Z, K) && aligned16(z) && aligned16(mu_lookup) && scratch && scratch_bytes >= fn_latent_scratch_bytes(B, Z, K)) {
        const int G = col_grid(B), chunk = (int)fn_cdiv((long long)B, G);
        lookup_grad_stage1_kernel<1><<<G, 256, (size_t)4 * K * Z * sizeof(float), (cudaStream_t)stream>>>(
            z, mu_lookup, logvar_lookup, qy, dlogLogit, dqy, nullptr, 0, B, Z, K, chunk, (double*)scratch);
        FN_LAUNCH_CHECK();
        lookup_grad_stage2_kernel<1><<<EW_GRID((long long)K * Z)>>>((const double*)scratch, G, logvar_lookup, nullptr, B, Z, K, dmu_lookup);
    } else {
        qy_bwd_dmu_kernel<<<EW_GRID((long long)K * Z)>>>(z, mu_lookup, logvar_lookup, qy, dlogLogit, dqy, B, Z, K, dmu_lookup);
    }
    FN_LAUNCH_CHECK();
    return FN_OK;
}
extern "C" int fn_gm_kl_fwd(const float* mu, const float* scale, const float* mu_lookup, const float* logvar_lookup,
                            const float* qy, const float* logLogit, const int64_t* y_label, int mode, int B, int Z,
                            int K, float* out3, void* scratch, size_t scratch_bytes, void* stream) {
    FN_REQUIRE(mu && scale && mu_lookup && logvar_lookup && qy && logLogit && out3, "fn_gm_kl_fwd: null pointer");
    FN_REQUIRE(B > 0 && Z > 0 && K >= 1 && (mode == 0 || y_label), "fn_gm_kl_fwd: bad args");
    if (fast_ok(Z, K) && aligned16(mu) && aligned16(scale) && scratch && scratch_bytes >= fn_latent_scratch_bytes(B, Z, K)) {
        const int G = row_grid(B);
#define FN_KLF(KT, LPR) gm_kl_fwd_fast_kernel<KT, LPR, 4><<<G, kRowThreads, (size_t)3 * K * Z * sizeof(float), (cudaStream_t)stream>>>( \
    mu, scale, mu_lookup, logvar_lookup, qy, logLogit, y_label, mode, B, Z, K, (double*)scratch)
#define FN_KLF2(KT) do { if (Z <= 128) FN_KLF(KT, 8); else FN_KLF(KT, 32); } while (0)
        if (K == 1) FN_KLF2(1); else if (K == 2) FN_KLF2(2); else if (K == 4) FN_KLF2(4); else FN_KLF2(0);
#undef FN_KLF2
#undef FN_KLF
        FN_LAUNCH_CHECK();
        latent_stage2_kernel<1><<<1, 256, 0, (cudaStream_t)stream>>>((const double*)scratch, G, 3, out3, 1.0 / (double)B, mode, K);
    } else {
        gm_kl_fwd_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(mu, scale, mu_lookup, logvar_lookup, qy, logLogit, y_label, mode, B, Z, K, out3);
    }
    FN_LAUNCH_CHECK();
    return FN_OK;
}
extern "C" int fn_gm_kl_bwd(const float* mu, const float* scale, const float* mu_lookup, const float* logvar_lookup,
                            const float* qy, const float* logLogit, const int64_t* y_label, int mode, int B, int Z,
                            int K, const float* dout3, float* dmu, float* dscale, float* dqy, float* dlogLogit,
                            float* dmu_lookup, void* scratch, size_t scratch_bytes, void* stream) {
    FN_REQUIRE(mu && scale && mu_lookup && logvar_lookup && qy && logLogit && dout3 && dmu && dscale && dqy && dlogLogit && dmu_lookup,
               "fn_gm_kl_bwd: null pointer");
    FN_REQUIRE(B > 0 && Z > 0 && K >= 1 && (mode == 0 || y_label), "fn_gm_kl_bwd: bad args");
    const bool fast = fast_ok(Z, K) && aligned16(mu) && aligned16(scale) && aligned16(dmu) && aligned16(dscale) && aligned16(mu_lookup);
#define FN_KLB(KT, LPR) gm_kl_bwd_rows_fast_kernel<KT, LPR, 4><<<row_grid(B), kRowThreads, (size_t)3 * K * Z * sizeof(float), (cudaStream_t)stream>>>( \
    mu, scale, mu_lookup, logvar_lookup, qy, logLogit, y_label, mode, B, Z, K, dout3, dmu, dscale, dqy, dlogLogit)
#define FN_KLB2(KT) do { if (Z <= 128) FN_KLB(KT, 8); else FN_KLB(KT, 32); } while (0)
    if (fast) {
        if (K == 1) FN_KLB2(1); else if (K == 2) FN_KLB2(2); else if (K == 4) FN_KLB2(4); else FN_KLB2(0);
    } else
        gm_kl_bwd_rows_kernel<<<fn_cdiv((long long)B * 32, 128), 128, 0, (cudaStream_t)stream>>>(
            mu, scale, mu_lookup, logvar_lookup, qy, logLogit, y_label, mode, B, Z, K, dout3, dmu, dscale, dqy, dlogLogit);
#undef FN_KLB2
#undef FN_KLB
    FN_LAUNCH_CHECK();
    if (fast && scratch && scratch_bytes >= fn_latent_scratch_bytes(B, Z, K)) {
        const int G = col_grid(B), chunk = (int)fn_cdiv((long long)B, G);
        lookup_grad_stage1_kernel<0><<<G, 256, (size_t)4 * K * Z * sizeof(float), (cudaStream_t)stream>>>(
            mu, mu_lookup, logvar_lookup, qy, nullptr, nullptr, y_label, mode, B, Z, K, chunk, (double*)scratch);
        FN_LAUNCH_CHECK();
        lookup_grad_stage2_kernel<0><<<EW_GRID((long long)K * Z)>>>((const double*)scratch, G, logvar_lookup, dout3, B, Z, K, dmu_lookup);
    } else {
        gm_kl_bwd_lookup_kernel<<<EW_GRID((long long)K * Z)>>>(mu, mu_lookup, logvar_lookup, qy, y_label, mode, B, Z, K, dout3, dmu_lookup);
    }
    FN_LAUNCH_CHECK();
    return FN_OK;
}
extern "C" int fn_std_kl_fwd(const float* mu, const float* scale, long long n, float* out, void* scratch, size_t scratch_bytes,
                             void* stream) {
    FN_REQUIRE(mu && scale && out && n > 0, "fn_std_kl_fwd: bad args");
    if (n % 4 == 0 && aligned16(mu) && aligned16(scale) && scratch && scratch_bytes >= 1184 * sizeof(double)) {
        int sms = fn_num_sms();
        if (sms <= 0) sms = 148;
        const long long need = (n / 4 + 255) / 256;
        const int G = (int)(need < (long long)sms * 8 ? need : (long long)sms * 8);       // <= 1184
        std_kl_fwd_fast_kernel<<<G, 256, 0, (cudaStream_t)stream>>>(mu, scale, n / 4, (double*)scratch);
        FN_LAUNCH_CHECK();
        latent_stage2_kernel<2><<<1, 256, 0, (cudaStream_t)stream>>>((const double*)scratch, G, 1, out, 1.0 / (double)n, 0, 1);
    } else {
        std_kl_fwd_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(mu, scale, n, out);
    }
    FN_LAUNCH_CHECK();
    return FN_OK;
}
extern "C" int fn_std_kl_bwd(const float* mu, const float* scale, long long n, const float* dout, float* dmu,
                             float* dscale, void* stream) {
    FN_REQUIRE(mu && scale && dout && dmu && dscale && n > 0, "fn_std_kl_bwd: bad args");
    std_kl_bwd_kernel<<<EW_GRID(n)>>>(mu, scale, n, dout, dmu, dscale);
    FN_LAUNCH_CHECK();
    return FN_OK;
}
extern "C" int fn_latent_reg_fwd(const float* z, long long z_ld, const double* attr, int B, float* loss, float* dz0,
                                 float* row_scratch, void* stream) {
    FN_REQUIRE(z && attr && loss && dz0 && row_scratch && B > 0, "fn_latent_reg_fwd: bad args");
    latent_reg_rows_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(z, z_ld, attr, B, row_scratch, dz0);
    FN_LAUNCH_CHECK();
    latent_reg_final_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(row_scratch, B, loss);
    FN_LAUNCH_CHECK();
    return FN_OK;
}
extern "C" int fn_latent_reg_bwd(const float* dz0, const float* dloss, int B, int Z, float* dz, void* stream) {
    FN_REQUIRE(dz0 && dloss && dz && B > 0 && Z > 0, "fn_latent_reg_bwd: bad args");
    latent_reg_bwd_kernel<<<EW_GRID((long long)B * Z)>>>(dz0, dloss, B, Z, dz);
    FN_LAUNCH_CHECK();
    return FN_OK;
}
extern "C" int fn_clip_adam(float* p, const float* g, float* m, float* v, long long n, const float* norm,
                            float max_norm, float lr, float beta1, float beta2, float eps, int step, void* stream) {
    FN_REQUIRE(p && g && m && v && norm && n > 0 && step >= 1, "fn_clip_adam: bad args");
    const double bc1 = 1.0 - pow((double)beta1, step), bc2 = 1.0 - pow((double)beta2, step);
    int sms = fn_num_sms();
    if (sms <= 0) sms = 148;
    const float step_size = (float)((double)lr / bc1), isb2 = (float)(1.0 / sqrt(bc2));
    if (n % 4 == 0 && aligned16(p) && aligned16(g) && aligned16(m) && aligned16(v)) {
        const int grid = (int)min((long long)sms * 8, (n / 4 + 255) / 256);
        clip_adam4_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((float4*)p, (const float4*)g, (float4*)m, (float4*)v, n / 4, norm,
                                                                  max_norm, step_size, beta1, beta2, eps, isb2);
    } else {
        const int grid = (int)min((long long)sms * 8, (n + 255) / 256);
        clip_adam_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, norm, max_norm, step_size, beta1, beta2, eps, isb2);
    }
    FN_LAUNCH_CHECK();
    return FN_OK;
}
