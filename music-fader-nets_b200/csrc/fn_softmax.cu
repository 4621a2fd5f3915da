// Soft-max heads (vocabulary axis / time axis) and NLL reductions.  Warp-per-row, shuffle
// reductions, HBM-bound: one read of the logits and one write of the log-probs per element.
#include "fn_common.cuh"

namespace {

constexpr int kRedBlocks = 1024;

// ---- vocabulary log-softmax: logits [T][B][V] -> out (B,T,V) --------------------------------
// One warp per row, the row held in REGISTERS (NV values per lane, V <= 32 * NV): one coalesced read of the logits,
// one write of the result, all loads of a row in flight at once.  NV = 0 is the multi-pass fallback for wider rows.
// MODE 0: plain log-softmax;  MODE 1: + NLL row loss and saved log-sum-exp (fused train path)
template <int NV>
__device__ __forceinline__ void row_load(const float* __restrict__ x, int V, int lane, float (&r)[NV > 0 ? NV : 1], float fill) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int v = lane + 32 * i;
        r[i] = v < V ? __ldg(x + v) : fill;
    }
}
template <int MODE, int NV>
__global__ void vocab_lsm_fwd_kernel(const float* __restrict__ logits, const int64_t* __restrict__ target, int B, int T,
                                     int V, float* __restrict__ out, float* __restrict__ lse_tm,
                                     float* __restrict__ loss_rows) {
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;   // = t*B + b
    const int lane = threadIdx.x & 31;
    if (row >= (long long)B * T) return;
    const int t = (int)(row / B), b = (int)(row % B);
    const float* x = logits + row * V;
    float r[NV > 0 ? NV : 1];
    float mx = -INFINITY, s = 0.f;
    if (NV > 0) {
        row_load<NV>(x, V, lane, r, -INFINITY);
#pragma unroll
        for (int i = 0; i < NV; ++i) mx = fmaxf(mx, r[i]);
        mx = fn_warp_max(mx);
#pragma unroll
        for (int i = 0; i < NV; ++i) s += expf(r[i] - mx);           // exp(-inf) = 0 for the padding
    } else {
        for (int v = lane; v < V; v += 32) mx = fmaxf(mx, x[v]);
        mx = fn_warp_max(mx);
        for (int v = lane; v < V; v += 32) s += expf(x[v] - mx);
    }
    s = fn_warp_sum(s);
    const float lse = mx + logf(s);
    if (out) {
        float* o = out + ((long long)b * T + t) * V;
        if (NV > 0) {
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const int v = lane + 32 * i;
                if (v < V) o[v] = r[i] - lse;
            }
        } else {
            for (int v = lane; v < V; v += 32) o[v] = x[v] - lse;
        }
    }
    if (MODE == 1 && lane == 0) {
        const int tg = (int)min((long long)V - 1, max(0LL, (long long)target[(long long)b * T + t]));
        lse_tm[row] = lse;
        loss_rows[row] = lse - x[tg];
    }
}

template <int NV>
__global__ void vocab_lsm_bwd_kernel(const float* __restrict__ out, const float* __restrict__ dout, int B, int T, int V,
                                     float* __restrict__ dlogits) {
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;   // = t*B + b
    const int lane = threadIdx.x & 31;
    if (row >= (long long)B * T) return;
    const int t = (int)(row / B), b = (int)(row % B);
    const float* o = out + ((long long)b * T + t) * V;
    const float* g = dout + ((long long)b * T + t) * V;
    float* d = dlogits + row * V;
    float s = 0.f;
    if (NV > 0) {
        float ro[NV > 0 ? NV : 1], rg[NV > 0 ? NV : 1];
        row_load<NV>(o, V, lane, ro, -INFINITY);
        row_load<NV>(g, V, lane, rg, 0.f);
#pragma unroll
        for (int i = 0; i < NV; ++i) s += rg[i];
        s = fn_warp_sum(s);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int v = lane + 32 * i;
            if (v < V) d[v] = rg[i] - expf(ro[i]) * s;
        }
    } else {
        for (int v = lane; v < V; v += 32) s += g[v];
        s = fn_warp_sum(s);
        for (int v = lane; v < V; v += 32) d[v] = g[v] - expf(o[v]) * s;
    }
}

template <int NV>
__global__ void vocab_nll_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ lse_tm,
                                     const int64_t* __restrict__ target, const float* __restrict__ scale_dev,
                                     float scale_host, int B, int T, int V, float* __restrict__ dlogits) {
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= (long long)B * T) return;
    const int t = (int)(row / B), b = (int)(row % B);
    const float sc = (scale_dev ? *scale_dev : 1.f) * scale_host;
    const float lse = lse_tm[row];
    const int tg = (int)min((long long)V - 1, max(0LL, (long long)target[(long long)b * T + t]));
    const float* x = logits + row * V;
    float* d = dlogits + row * V;
    if (NV > 0) {
        float r[NV > 0 ? NV : 1];
        row_load<NV>(x, V, lane, r, -INFINITY);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int v = lane + 32 * i;
            if (v < V) d[v] = sc * (expf(r[i] - lse) - (v == tg ? 1.f : 0.f));
        }
    } else {
        for (int v = lane; v < V; v += 32) d[v] = sc * (expf(x[v] - lse) - (v == tg ? 1.f : 0.f));
    }
}
// rows of up to 384 (the 342-token vocabulary) / 1024 values live in registers; wider rows take the multi-pass kernel
#define FN_VOCAB_DISPATCH(V, CALL12, CALL32, CALL0) \
    do {                                            \
        if ((V) <= 384) { CALL12; }                 \
        else if ((V) <= 1024) { CALL32; }           \
        else { CALL0; }                             \
    } while (0)

// ---- time-axis log-softmax: logits [T][B][C] -> out (B,T,C), normalised over T per (b,c) ------
__global__ void time_lsm_fwd_kernel(const float* __restrict__ logits, int B, int T, int C, float* __restrict__ out) {
    const int col = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;   // = b*C + c
    const int lane = threadIdx.x & 31;
    if (col >= B * C) return;
    const int b = col / C, c = col % C;
    const long long st = (long long)B * C;
    const float* x = logits + (long long)b * C + c;
    float mx = -INFINITY;
    for (int t = lane; t < T; t += 32) mx = fmaxf(mx, x[t * st]);
    mx = fn_warp_max(mx);
    float s = 0.f;
    for (int t = lane; t < T; t += 32) s += expf(x[t * st] - mx);
    s = fn_warp_sum(s);
    const float lse = mx + logf(s);
    float* o = out + (long long)b * T * C + c;
    for (int t = lane; t < T; t += 32) o[(long long)t * C] = x[t * st] - lse;
}
__global__ void time_lsm_bwd_kernel(const float* __restrict__ out, const float* __restrict__ dout, int B, int T, int C,
                                    float* __restrict__ dlogits) {
    const int col = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (col >= B * C) return;
    const int b = col / C, c = col % C;
    const float* o = out + (long long)b * T * C + c;
    const float* g = dout + (long long)b * T * C + c;
    float s = 0.f;
    for (int t = lane; t < T; t += 32) s += g[(long long)t * C];
    s = fn_warp_sum(s);
    const long long st = (long long)B * C;
    float* d = dlogits + (long long)b * C + c;
    for (int t = lane; t < T; t += 32) d[t * st] = g[(long long)t * C] - expf(o[(long long)t * C]) * s;
}

// ---- deterministic reductions ---------------------------------------------------------------
// GATHER 0: x[i];  GATHER 1: -logp[i*C + target[i]]
template <int GATHER>
__global__ void reduce_stage1(const float* __restrict__ x, const int64_t* __restrict__ target, int C, long long n,
                              double* __restrict__ partial) {
    __shared__ double red[33];
    double s = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        if (GATHER == 0) s += (double)x[i];
        else {
            const long long tg = target[i];
            s -= (double)x[i * C + (tg < 0 ? 0 : tg >= C ? C - 1 : tg)];      // clamped: never out of bounds (ops.check_index counts + raises)
        }
    }
    s = fn_block_sum_d(s, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}
// SQRT 1: out = sqrt(sum)  (gradient norm)
template <int SQRT>
__global__ void reduce_stage2(const double* __restrict__ partial, int nblocks, double scale, float* __restrict__ out) {
    __shared__ double red[33];
    double s = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += blockDim.x) s += partial[i];
    s = fn_block_sum_d(s, red);
    if (threadIdx.x == 0) out[0] = (float)(SQRT ? sqrt(s) : s * scale);
}
__global__ void sq_stage1(const float* __restrict__ x, long long n, double* __restrict__ partial) {
    __shared__ double red[33];
    double s = 0.0;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
    if ((reinterpret_cast<uintptr_t>(x) & 15) == 0) {
        // 16-byte loads, two in flight per thread; squares of a quad summed in double like the scalar tail
        const long long n4 = n >> 2;
        const float4* x4 = reinterpret_cast<const float4*>(x);
        long long i = tid;
        for (; i + nth < n4; i += 2 * nth) {
            const float4 a = __ldg(x4 + i), b = __ldg(x4 + i + nth);
            s += (double)a.x * a.x + (double)a.y * a.y + (double)a.z * a.z + (double)a.w * a.w;
            s += (double)b.x * b.x + (double)b.y * b.y + (double)b.z * b.z + (double)b.w * b.w;
        }
        for (; i < n4; i += nth) {
            const float4 a = __ldg(x4 + i);
            s += (double)a.x * a.x + (double)a.y * a.y + (double)a.z * a.z + (double)a.w * a.w;
        }
        for (long long j = (n4 << 2) + tid; j < n; j += nth) s += (double)x[j] * x[j];
    } else {
        for (long long i = tid; i < n; i += nth) {
            const double v = (double)x[i];
            s += v * v;
        }
    }
    s = fn_block_sum_d(s, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

__global__ void nll_bwd_kernel(const int64_t* __restrict__ target, long long rows, int C, const float* __restrict__ dloss,
                               float* __restrict__ dlogp) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    const long long tg = target[i];
    dlogp[i * C + (tg < 0 ? 0 : tg >= C ? C - 1 : tg)] += -dloss[0] / (float)rows;
}

int red_blocks(long long n) { return (int)max(1LL, min((long long)kRedBlocks, (n + 1023) / 1024)); }

}  // namespace

extern "C" size_t fn_reduce_scratch_bytes(long long n) {
    (void)n;
    return (size_t)kRedBlocks * sizeof(double);
}

extern "C" int fn_vocab_logsoftmax_fwd(const float* logits_tm, int B, int T, int V, float* out_bm, void* stream) {
    FN_REQUIRE(logits_tm && out_bm && B > 0 && T > 0 && V > 0, "fn_vocab_logsoftmax_fwd: bad args");
    const long long rows = (long long)B * T;
#define FN_L(NV) vocab_lsm_fwd_kernel<0, NV><<<fn_cdiv(rows * 32, 256), 256, 0, (cudaStream_t)stream>>>(logits_tm, nullptr, B, T, V, out_bm, nullptr, nullptr)
    FN_VOCAB_DISPATCH(V, FN_L(12), FN_L(32), FN_L(0));
#undef FN_L
    FN_LAUNCH_CHECK();
    return FN_OK;
}
extern "C" int fn_vocab_logsoftmax_bwd(const float* out_bm, const float* dout_bm, int B, int T, int V,
                                       float* dlogits_tm, void* stream) {
    FN_REQUIRE(out_bm && dout_bm && dlogits_tm && B > 0 && T > 0 && V > 0, "fn_vocab_logsoftmax_bwd: bad args");
    const long long rows = (long long)B * T;
#define FN_L(NV) vocab_lsm_bwd_kernel<NV><<<fn_cdiv(rows * 32, 256), 256, 0, (cudaStream_t)stream>>>(out_bm, dout_bm, B, T, V, dlogits_tm)
    FN_VOCAB_DISPATCH(V, FN_L(12), FN_L(32), FN_L(0));
#undef FN_L
    FN_LAUNCH_CHECK();
    return FN_OK;
}
extern "C" int fn_vocab_nll_fwd(const float* logits_tm, const int64_t* target_bm, int B, int T, int V, float* out_bm,
                                float* lse_tm, float* loss_rows, void* stream) {
    FN_REQUIRE(logits_tm && target_bm && lse_tm && loss_rows && B > 0 && T > 0 && V > 0, "fn_vocab_nll_fwd: bad args");
    const long long rows = (long long)B * T;
#define FN_L(NV) vocab_lsm_fwd_kernel<1, NV><<<fn_cdiv(rows * 32, 256), 256, 0, (cudaStream_t)stream>>>(logits_tm, target_bm, B, T, V, out_bm, lse_tm, loss_rows)
    FN_VOCAB_DISPATCH(V, FN_L(12), FN_L(32), FN_L(0));
#undef FN_L
    FN_LAUNCH_CHECK();
    return FN_OK;
}
extern "C" int fn_vocab_nll_bwd(const float* logits_tm, const float* lse_tm, const int64_t* target_bm,
                                const float* scale_dev, float scale_host, int B, int T, int V, float* dlogits_tm,
                                void* stream) {
    FN_REQUIRE(logits_tm && lse_tm && target_bm && dlogits_tm && B > 0 && T > 0 && V > 0, "fn_vocab_nll_bwd: bad args");
    const long long rows = (long long)B * T;
#define FN_L(NV) vocab_nll_bwd_kernel<NV><<<fn_cdiv(rows * 32, 256), 256, 0, (cudaStream_t)stream>>>(logits_tm, lse_tm, target_bm, scale_dev, scale_host, B, T, V, dlogits_tm)
    FN_VOCAB_DISPATCH(V, FN_L(12), FN_L(32), FN_L(0));
#undef FN_L
    FN_LAUNCH_CHECK();
    return FN_OK;
}

extern "C" int fn_time_logsoftmax_fwd(const float* logits_tm, int B, int T, int C, float* out_bm, void* stream) {
    FN_REQUIRE(logits_tm && out_bm && B > 0 && T > 0 && C > 0, "fn_time_logsoftmax_fwd: bad args");
    time_lsm_fwd_kernel<<<fn_cdiv((long long)B * C * 32, 128), 128, 0, (cudaStream_t)stream>>>(logits_tm, B, T, C, out_bm);
    FN_LAUNCH_CHECK();
    return FN_OK;
}
extern "C" int fn_time_logsoftmax_bwd(const float* out_bm, const float* dout_bm, int B, int T, int C,
                                      float* dlogits_tm, void* stream) {
    FN_REQUIRE(out_bm && dout_bm && dlogits_tm && B > 0 && T > 0 && C > 0, "fn_time_logsoftmax_bwd: bad args");
    time_lsm_bwd_kernel<<<fn_cdiv((long long)B * C * 32, 128), 128, 0, (cudaStream_t)stream>>>(out_bm, dout_bm, B, T, C, dlogits_tm);
    FN_LAUNCH_CHECK();
    return FN_OK;
}

extern "C" int fn_nll_mean_fwd(const float* logp, const int64_t* target, long long rows, int C, float* loss,
                               void* scratch, size_t scratch_bytes, void* stream) {
    FN_REQUIRE(logp && target && loss && scratch && rows > 0 && C > 0, "fn_nll_mean_fwd: bad args");
    FN_REQUIRE(scratch_bytes >= fn_reduce_scratch_bytes(rows), "fn_nll_mean_fwd: scratch too small");
    cudaStream_t st = (cudaStream_t)stream;
    const int nb = red_blocks(rows);
    reduce_stage1<1><<<nb, 256, 0, st>>>(logp, target, C, rows, (double*)scratch);
    FN_LAUNCH_CHECK();
    reduce_stage2<0><<<1, 256, 0, st>>>((const double*)scratch, nb, 1.0 / (double)rows, loss);
    FN_LAUNCH_CHECK();
    return FN_OK;
}
extern "C" int fn_nll_mean_bwd(const int64_t* target, long long rows, int C, const float* dloss, float* dlogp,
                               int zero_fill, void* stream) {
    FN_REQUIRE(target && dloss && dlogp && rows > 0 && C > 0, "fn_nll_mean_bwd: bad args");
    cudaStream_t st = (cudaStream_t)stream;
    if (zero_fill) FN_CHECK_CUDA(cudaMemsetAsync(dlogp, 0, (size_t)rows * C * sizeof(float), st));
    nll_bwd_kernel<<<fn_cdiv(rows, 256), 256, 0, st>>>(target, rows, C, dloss, dlogp);
    FN_LAUNCH_CHECK();
    return FN_OK;
}
extern "C" int fn_sum_f32(const float* x, long long n, float scale, float* out, void* scratch, size_t scratch_bytes,
                          void* stream) {
    FN_REQUIRE(x && out && scratch && n > 0, "fn_sum_f32: bad args");
    FN_REQUIRE(scratch_bytes >= fn_reduce_scratch_bytes(n), "fn_sum_f32: scratch too small");
    cudaStream_t st = (cudaStream_t)stream;
    const int nb = red_blocks(n);
    reduce_stage1<0><<<nb, 256, 0, st>>>(x, nullptr, 0, n, (double*)scratch);
    FN_LAUNCH_CHECK();
    reduce_stage2<0><<<1, 256, 0, st>>>((const double*)scratch, nb, (double)scale, out);
    FN_LAUNCH_CHECK();
    return FN_OK;
}
extern "C" int fn_grad_norm(const float* g, long long n, float* norm_out, void* scratch, size_t scratch_bytes,
                            void* stream) {
    FN_REQUIRE(g && norm_out && scratch && n > 0, "fn_grad_norm: bad args");
    FN_REQUIRE(scratch_bytes >= fn_reduce_scratch_bytes(n), "fn_grad_norm: scratch too small");
    cudaStream_t st = (cudaStream_t)stream;
    const int nb = red_blocks(n);
    sq_stage1<<<nb, 256, 0, st>>>(g, n, (double*)scratch);
    FN_LAUNCH_CHECK();
    reduce_stage2<1><<<1, 256, 0, st>>>((const double*)scratch, nb, 1.0, norm_out);
    FN_LAUNCH_CHECK();
    return FN_OK;
}
