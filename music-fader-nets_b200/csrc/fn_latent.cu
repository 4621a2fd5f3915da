// Latent block: exp/reparameterisation, Gaussian-mixture responsibilities, the categorical-
// Gaussian KL terms, the pairwise latent regulariser and the clip+Adam update.  All of these
// are KB-scale at the reference's shapes (latency-bound); they are warp-shuffle kernels with
// fixed reduction orders so results are run-to-run deterministic.
#include "fn_common.cuh"

namespace {

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
        const int k_lo = mode ? (int)ylab[b] : 0, k_hi = mode ? k_lo + 1 : K;
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
            a2 += (mx + logf(den)) - q[(int)ylab[b]];
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
    const int k_lo = mode ? (int)ylab[b] : 0, k_hi = mode ? k_lo + 1 : K;
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
            gq = d2 * (expf(q[k] - mxq) / denq - (k == (int)ylab[b] ? 1.f : 0.f));
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
        const float wgt = mode ? ((int)ylab[b] == k ? 1.f : 0.f) : qy[b * K + k];
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

}  // namespace

#define EW_GRID(n) fn_cdiv((n), 256), 256, 0, (cudaStream_t)stream

extern "C" int fn_reparam_fwd(const float* mu, const float* pre_scale, const float* eps, long long n, float* scale,
                              float* z, void* stream) {
    FN_REQUIRE(pre_scale && scale && n > 0 && (!z || (mu && eps)), "fn_reparam_fwd: bad args");
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
    qy_fwd_kernel<<<fn_cdiv((long long)B * 32, 128), 128, 0, (cudaStream_t)stream>>>(z, mu_lookup, logvar_lookup, B, Z, K, logLogit, qy, y);
    FN_LAUNCH_CHECK();
    return FN_OK;
}
extern "C" int fn_qy_x_bwd(const float* z, const float* mu_lookup, const float* logvar_lookup, const float* qy,
                           const float* dlogLogit, const float* dqy, int B, int Z, int K, float* dz, float* dmu_lookup,
                           void* stream) {
    FN_REQUIRE(z && mu_lookup && logvar_lookup && qy && dz && dmu_lookup && B > 0 && Z > 0 && K >= 1, "fn_qy_x_bwd: bad args");
    qy_bwd_dz_kernel<<<EW_GRID((long long)B * Z)>>>(z, mu_lookup, logvar_lookup, qy, dlogLogit, dqy, B, Z, K, dz);
    FN_LAUNCH_CHECK();
    qy_bwd_dmu_kernel<<<EW_GRID((long long)K * Z)>>>(z, mu_lookup, logvar_lookup, qy, dlogLogit, dqy, B, Z, K, dmu_lookup);
    FN_LAUNCH_CHECK();
    return FN_OK;
}
extern "C" int fn_gm_kl_fwd(const float* mu, const float* scale, const float* mu_lookup, const float* logvar_lookup,
                            const float* qy, const float* logLogit, const int64_t* y_label, int mode, int B, int Z,
                            int K, float* out3, void* stream) {
    FN_REQUIRE(mu && scale && mu_lookup && logvar_lookup && qy && logLogit && out3, "fn_gm_kl_fwd: null pointer");
    FN_REQUIRE(B > 0 && Z > 0 && K >= 1 && (mode == 0 || y_label), "fn_gm_kl_fwd: bad args");
    gm_kl_fwd_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(mu, scale, mu_lookup, logvar_lookup, qy, logLogit, y_label, mode, B, Z, K, out3);
    FN_LAUNCH_CHECK();
    return FN_OK;
}
extern "C" int fn_gm_kl_bwd(const float* mu, const float* scale, const float* mu_lookup, const float* logvar_lookup,
                            const float* qy, const float* logLogit, const int64_t* y_label, int mode, int B, int Z,
                            int K, const float* dout3, float* dmu, float* dscale, float* dqy, float* dlogLogit,
                            float* dmu_lookup, void* stream) {
    FN_REQUIRE(mu && scale && mu_lookup && logvar_lookup && qy && logLogit && dout3 && dmu && dscale && dqy && dlogLogit && dmu_lookup,
               "fn_gm_kl_bwd: null pointer");
    FN_REQUIRE(B > 0 && Z > 0 && K >= 1 && (mode == 0 || y_label), "fn_gm_kl_bwd: bad args");
    gm_kl_bwd_rows_kernel<<<fn_cdiv((long long)B * 32, 128), 128, 0, (cudaStream_t)stream>>>(
        mu, scale, mu_lookup, logvar_lookup, qy, logLogit, y_label, mode, B, Z, K, dout3, dmu, dscale, dqy, dlogLogit);
    FN_LAUNCH_CHECK();
    gm_kl_bwd_lookup_kernel<<<EW_GRID((long long)K * Z)>>>(mu, mu_lookup, logvar_lookup, qy, y_label, mode, B, Z, K, dout3, dmu_lookup);
    FN_LAUNCH_CHECK();
    return FN_OK;
}
extern "C" int fn_std_kl_fwd(const float* mu, const float* scale, long long n, float* out, void* stream) {
    FN_REQUIRE(mu && scale && out && n > 0, "fn_std_kl_fwd: bad args");
    std_kl_fwd_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(mu, scale, n, out);
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
    const int grid = (int)min((long long)sms * 8, (n + 255) / 256);
    clip_adam_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, norm, max_norm, (float)((double)lr / bc1), beta1,
                                                             beta2, eps, (float)(1.0 / sqrt(bc2)));
    FN_LAUNCH_CHECK();
    return FN_OK;
}
