// Token plumbing, layout changes and the reductions that finish the GRU parameter gradients.
#include <stdarg.h>
#include <mutex>

#include "fn_common.cuh"

// ---- library-wide state: last error + cached device attributes ---------------------------
static thread_local char g_err[512] = "";
void fn_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
extern "C" const char* fn_last_error(void) { return g_err; }
extern "C" int fn_abi_version(void) { return FN_ABI_VERSION; }
#ifndef FN_SOURCE_HASH
#define FN_SOURCE_HASH "unknown"
#endif
extern "C" const char* fn_source_hash(void) { return FN_SOURCE_HASH; }

static int g_attr_dev = -1, g_sms = 0, g_smem = 0;
static void refresh_attrs() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return;
    if (dev == g_attr_dev) return;
    cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&g_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    g_attr_dev = dev;
}
int fn_num_sms() { refresh_attrs(); return g_sms; }
int fn_max_smem_optin() { refresh_attrs(); return g_smem; }

extern "C" int fn_device_info(int* sm_count, int* cc_major, int* cc_minor, int* max_smem_optin) {
    int dev = 0;
    FN_CHECK_CUDA(cudaGetDevice(&dev));
    if (sm_count) FN_CHECK_CUDA(cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, dev));
    if (cc_major) FN_CHECK_CUDA(cudaDeviceGetAttribute(cc_major, cudaDevAttrComputeCapabilityMajor, dev));
    if (cc_minor) FN_CHECK_CUDA(cudaDeviceGetAttribute(cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
    if (max_smem_optin) FN_CHECK_CUDA(cudaDeviceGetAttribute(max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    return FN_OK;
}

namespace {

// ---- one-hot <-> ids ------------------------------------------------------------------------
__global__ void onehot_to_ids_kernel(const float* __restrict__ oh, int B, int T, int V, int32_t* __restrict__ ids_tm) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= B * T) return;
    const int b = warp / T, t = warp % T;
    const float* row = oh + (long long)warp * V;
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int v = lane; v < V; v += 32) {
        const float x = row[v];
        if (x > best) { best = x; bi = v; }          // ascending v per lane: keeps the first max
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (lane == 0) ids_tm[(long long)t * B + b] = bi == 0x7fffffff ? 0 : bi;     // a row without a maximum (all NaN / -inf) still yields a valid id
}

__global__ void ids_to_onehot_kernel(const int64_t* __restrict__ ids, long long rows, int V, float* __restrict__ oh) {
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= rows) return;
    const int id = (int)ids[warp];
    float* row = oh + warp * V;
    for (int v = lane; v < V; v += 32) row[v] = (v == id) ? 1.f : 0.f;
}

__global__ void ids_to_tm_kernel(const int64_t* __restrict__ ids, int B, int T, int shift, int start,
                                 int32_t* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * T) return;
    const int t = (int)(i / B), b = (int)(i % B);
    const int ts = t - shift;
    out[i] = ts < 0 ? start : (int)ids[(long long)b * T + ts];
}

// Index validation (the reference's F.nll_loss / nn.Embedding raise on an out-of-range index; here the ids are clamped so
// that no kernel ever reads or writes out of bounds, and counted so that the host mirror can raise at its next sync).
__global__ void check_index_i64_kernel(const int64_t* __restrict__ idx, long long n, long long hi, int32_t* __restrict__ bad) {
    int c = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long v = idx[i];
        c += (v < 0 || v >= hi) ? 1 : 0;
    }
    if (__syncthreads_or(c)) {
        if (c) atomicAdd(bad, c);
    }
}
__global__ void clamp_index_i32_kernel(int32_t* __restrict__ idx, long long n, int hi, int32_t* __restrict__ bad) {
    int c = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int v = idx[i];
        if (v < 0 || v >= hi) { idx[i] = v < 0 ? 0 : hi - 1; ++c; }
    }
    if (__syncthreads_or(c)) {
        if (c) atomicAdd(bad, c);
    }
}

// ---- small element-wise pieces of the sibling models (MusicAttrFaderNets, model_v2.py:426-435, 574-575; trainer_fader.py:105-110)
// y = relu(x) * mask   (mask = the dropout keep mask already divided by 1 - p, or NULL)
__global__ void relu_mask_fwd_kernel(const float* __restrict__ x, const float* __restrict__ mask, float* __restrict__ y, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = fmaxf(x[i], 0.f) * (mask ? mask[i] : 1.f);
}
__global__ void relu_mask_bwd_kernel(const float* __restrict__ x, const float* __restrict__ mask, const float* __restrict__ dy,
                                     float* __restrict__ dx, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dx[i] = x[i] > 0.f ? dy[i] * (mask ? mask[i] : 1.f) : 0.f;
}
__global__ void scale_kernel(const float* __restrict__ src, float* __restrict__ dst, float alpha, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = alpha * src[i];
}
// loss = mean_i (x_i - y_i)^2 : ONE block, fixed summation order (deterministic); n is a batch size
__global__ void mse_mean_fwd_kernel(const float* __restrict__ x, const float* __restrict__ y, long long n, float* __restrict__ loss) {
    __shared__ double red[33];
    double s = 0.0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) { const double d = (double)x[i] - (double)y[i]; s += d * d; }
    s = fn_block_sum_d(s, red);
    if (threadIdx.x == 0) loss[0] = (float)(s / (double)n);
}
__global__ void mse_mean_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y, long long n, const float* __restrict__ dloss,
                                    float* __restrict__ dx) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dx[i] = 2.f * (x[i] - y[i]) / (float)n * dloss[0];
}

__global__ void transpose_kernel(const float* __restrict__ src, long long ld_src, float* __restrict__ dst,
                                 long long ld_dst, int rows, int cols, int accumulate) {
    __shared__ float tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < rows && c < cols) ? src[(long long)r * ld_src + c] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        if (r < rows && c < cols) {
            float* d = dst + (long long)c * ld_dst + r;
            *d = accumulate ? *d + tile[threadIdx.x][i] : tile[threadIdx.x][i];
        }
    }
}

// clean_output (test_class.py:44-50) per row: trim leading/trailing zeros, cut at the first EOS (1) of what is left.
// One warp per row; start/len describe the kept span of the ORIGINAL row.
__global__ void clean_tokens_kernel(const int64_t* __restrict__ tok, int rows, int S, int32_t* __restrict__ start,
                                    int32_t* __restrict__ len) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= rows) return;
    const int64_t* t = tok + (long long)row * S;
    int first = S, last = -1;
    for (int i = lane; i < S; i += 32)
        if (t[i] != 0) { first = min(first, i); last = max(last, i); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
        last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
    }
    int eos = S;
    for (int i = first + lane; i <= last; i += 32)
        if (t[i] == 1) eos = min(eos, i);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) eos = min(eos, __shfl_xor_sync(0xffffffffu, eos, o));
    if (lane == 0) {
        if (last < 0) { start[row] = 0; len[row] = 0; }
        else { start[row] = first; len[row] = (eos <= last ? eos : last + 1) - first; }
    }
}

__global__ void add_kernel(float* __restrict__ dst, const float* __restrict__ src, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] += src[i];
}

// ---- embedding-table gradient --------------------------------------------------------------
// stage 1: block (col chunk of 128, row chunk) accumulates a private [V][128] table in smem
// walking its rows in order -> deterministic.  RS row-lanes (each with its own table) when V is small.
constexpr int kEC = 128;
__global__ void emb_grad_stage1(const int32_t* __restrict__ ids, const float* __restrict__ dgh,
                                const float* __restrict__ dgin, long long rows, int H, int V, int rows_per_chunk,
                                int RS, float* __restrict__ partial) {
    extern __shared__ float tab[];              // [RS][V][kEC]
    const int K3 = 3 * H;
    const int col = blockIdx.x * kEC + (threadIdx.x % kEC);
    const int lane_r = threadIdx.x / kEC;
    const bool colok = col < K3;
    float* my = tab + (long long)lane_r * V * kEC + (threadIdx.x % kEC);
    for (int v = 0; v < V; ++v) my[v * kEC] = 0.f;
    const long long r_begin = (long long)blockIdx.y * rows_per_chunk;
    const long long r_end = min(rows, r_begin + rows_per_chunk);
    // column -> source stream
    const float* src;
    long long ld;
    int c;
    if (col < 2 * H) { src = dgh; ld = K3; c = col; } else { src = dgin; ld = H; c = col - 2 * H; }
    constexpr int UN = 8;
    for (long long r = r_begin + lane_r; r < r_end; r += (long long)RS * UN) {
        float v[UN];
        int id[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const long long rr = r + (long long)u * RS;
            const bool ok = rr < r_end;
            id[u] = ok ? ids[rr] : -1;
            v[u] = (ok && colok) ? __ldg(src + rr * ld + c) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < UN; ++u)
            if (id[u] >= 0) my[id[u] * kEC] += v[u];
    }
    __syncthreads();
    if (lane_r == 0 && colok) {
        float* out = partial + (long long)blockIdx.y * V * K3 + col;
        for (int v = 0; v < V; ++v) {
            float s = 0.f;
            for (int q = 0; q < RS; ++q) s += tab[((long long)q * V + v) * kEC + (threadIdx.x % kEC)];
            out[(long long)v * K3] = s;
        }
    }
}
__global__ void emb_grad_stage2(const float* __restrict__ partial, int nchunks, long long n, float* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s = 0.f;
    for (int q = 0; q < nchunks; ++q) s += partial[(long long)q * n + i];
    out[i] = s;
}

// ---- sums over time / rows -------------------------------------------------------------------
__global__ void time_sum_kernel(const float* __restrict__ dgh, const float* __restrict__ dgin, int B, int T, int H,
                                float* __restrict__ dproj, float* __restrict__ dghsum) {
    const int K3 = 3 * H;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * K3) return;
    const int b = (int)(i / K3), c = (int)(i % K3);
    float sgh = 0.f, sgi = 0.f;
    const bool npart = c >= 2 * H;
    for (int t = 0; t < T; ++t) {
        const long long row = (long long)t * B + b;
        sgh += dgh[row * K3 + c];
        if (npart) sgi += dgin[row * H + (c - 2 * H)];
    }
    if (dghsum) dghsum[i] = sgh;
    if (dproj) dproj[i] = npart ? sgi : sgh;
}

constexpr int kCS_R = 32;
__global__ void col_sum_stage1(const float* __restrict__ x, long long ld, long long rows, int cols,
                               long long rows_per_chunk, float* __restrict__ partial) {
    __shared__ float red[32][33];
    const int c = blockIdx.x * 32 + threadIdx.x;
    const long long r0 = (long long)blockIdx.y * rows_per_chunk, r1 = min(rows, r0 + rows_per_chunk);
    float s = 0.f;
    if (c < cols)
        for (long long r = r0 + threadIdx.y; r < r1; r += 32) s += x[r * ld + c];
    red[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && c < cols) {
        float t = 0.f;
        for (int q = 0; q < 32; ++q) t += red[q][threadIdx.x];
        partial[(long long)blockIdx.y * cols + c] = t;
    }
}
__global__ void col_sum_stage2(const float* __restrict__ partial, int nchunks, int cols, float* __restrict__ out,
                               int accumulate) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    float s = 0.f;
    for (int q = 0; q < nchunks; ++q) s += partial[(long long)q * cols + c];
    out[c] = accumulate ? out[c] + s : s;
}

int emb_cfg(int B, int T, int H, int V, int* RS, int* nchunks, int* rows_per_chunk) {
    const long long rows = (long long)B * T;
    const size_t cap = 200 * 1024;
    int rs = (int)(cap / ((size_t)V * kEC * sizeof(float)));
    if (rs < 1) return -1;
    if (rs > 8) rs = 8;
    const int colchunks = fn_cdiv(3 * H, kEC);
    int sms = fn_num_sms();
    if (sms <= 0) sms = 148;
    int nc = max(1, (2 * sms) / colchunks);
    const long long min_rows = 64LL * rs;                  // do not cut thinner than this
    if ((long long)nc * min_rows > rows) nc = (int)max(1LL, rows / min_rows);
    *RS = rs;
    *nchunks = nc;
    *rows_per_chunk = (int)((rows + nc - 1) / nc);
    return 0;
}

}  // namespace

extern "C" int fn_onehot_to_ids(const float* onehot, int B, int T, int V, int32_t* ids_tm, void* stream) {
    FN_REQUIRE(onehot && ids_tm && B > 0 && T > 0 && V > 0, "fn_onehot_to_ids: bad args");
    const long long warps = (long long)B * T;
    onehot_to_ids_kernel<<<fn_cdiv(warps * 32, 256), 256, 0, (cudaStream_t)stream>>>(onehot, B, T, V, ids_tm);
    FN_LAUNCH_CHECK();
    return FN_OK;
}
extern "C" int fn_ids_to_onehot(const int64_t* ids, int B, int T, int V, float* onehot, void* stream) {
    FN_REQUIRE(ids && onehot && B > 0 && T > 0 && V > 0, "fn_ids_to_onehot: bad args");
    const long long rows = (long long)B * T;
    ids_to_onehot_kernel<<<fn_cdiv(rows * 32, 256), 256, 0, (cudaStream_t)stream>>>(ids, rows, V, onehot);
    FN_LAUNCH_CHECK();
    return FN_OK;
}
extern "C" int fn_ids_to_time_major(const int64_t* ids, int B, int T, int shift, int start_token, int32_t* ids_tm,
                                    void* stream) {
    FN_REQUIRE(ids && ids_tm && B > 0 && T > 0 && shift >= 0, "fn_ids_to_time_major: bad args");
    ids_to_tm_kernel<<<fn_cdiv((long long)B * T, 256), 256, 0, (cudaStream_t)stream>>>(ids, B, T, shift, start_token, ids_tm);
    FN_LAUNCH_CHECK();
    return FN_OK;
}
extern "C" int fn_check_index_i64(const int64_t* idx, long long n, long long hi, int32_t* bad_count, void* stream) {
    FN_REQUIRE(idx && bad_count && n >= 0 && hi > 0, "fn_check_index_i64: bad args");
    if (n == 0) return FN_OK;
    const int blocks = (int)((n + 1023) / 1024 < 592 ? (n + 1023) / 1024 : 592);
    check_index_i64_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(idx, n, hi, bad_count);
    FN_LAUNCH_CHECK();
    return FN_OK;
}
extern "C" int fn_clamp_index_i32(int32_t* idx, long long n, int hi, int32_t* bad_count, void* stream) {
    FN_REQUIRE(idx && bad_count && n >= 0 && hi > 0, "fn_clamp_index_i32: bad args");
    if (n == 0) return FN_OK;
    const int blocks = (int)((n + 1023) / 1024 < 592 ? (n + 1023) / 1024 : 592);
    clamp_index_i32_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(idx, n, hi, bad_count);
    FN_LAUNCH_CHECK();
    return FN_OK;
}

extern "C" int fn_relu_mask_fwd(const float* x, const float* mask, float* y, long long n, void* stream) {
    FN_REQUIRE(x && y && n >= 0, "fn_relu_mask_fwd: bad args");
    if (n) relu_mask_fwd_kernel<<<fn_cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(x, mask, y, n);
    FN_LAUNCH_CHECK();
    return FN_OK;
}
extern "C" int fn_relu_mask_bwd(const float* x, const float* mask, const float* dy, float* dx, long long n, void* stream) {
    FN_REQUIRE(x && dy && dx && n >= 0, "fn_relu_mask_bwd: bad args");
    if (n) relu_mask_bwd_kernel<<<fn_cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(x, mask, dy, dx, n);
    FN_LAUNCH_CHECK();
    return FN_OK;
}
extern "C" int fn_scale_f32(const float* src, float* dst, float alpha, long long n, void* stream) {
    FN_REQUIRE(src && dst && n >= 0, "fn_scale_f32: bad args");
    if (n) scale_kernel<<<fn_cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(src, dst, alpha, n);
    FN_LAUNCH_CHECK();
    return FN_OK;
}
extern "C" int fn_mse_mean_fwd(const float* x, const float* y, long long n, float* loss, void* stream) {
    FN_REQUIRE(x && y && loss && n > 0, "fn_mse_mean_fwd: bad args");
    mse_mean_fwd_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(x, y, n, loss);
    FN_LAUNCH_CHECK();
    return FN_OK;
}
extern "C" int fn_mse_mean_bwd(const float* x, const float* y, long long n, const float* dloss, float* dx, void* stream) {
    FN_REQUIRE(x && y && dloss && dx && n > 0, "fn_mse_mean_bwd: bad args");
    mse_mean_bwd_kernel<<<fn_cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(x, y, n, dloss, dx);
    FN_LAUNCH_CHECK();
    return FN_OK;
}

extern "C" int fn_transpose_f32(const float* src, long long ld_src, float* dst, long long ld_dst, int rows, int cols,
                                int accumulate, void* stream) {
    FN_REQUIRE(src && dst && rows > 0 && cols > 0, "fn_transpose_f32: bad args");
    dim3 grid(fn_cdiv(cols, 32), fn_cdiv(rows, 32)), block(32, 8);
    transpose_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(src, ld_src, dst, ld_dst, rows, cols, accumulate);
    FN_LAUNCH_CHECK();
    return FN_OK;
}

extern "C" int fn_clean_tokens(const int64_t* tokens, int rows, int steps, int32_t* start, int32_t* len, void* stream) {
    FN_REQUIRE(tokens && start && len && rows > 0 && steps > 0, "fn_clean_tokens: bad args");
    clean_tokens_kernel<<<fn_cdiv((long long)rows * 32, 256), 256, 0, (cudaStream_t)stream>>>(tokens, rows, steps, start, len);
    FN_LAUNCH_CHECK();
    return FN_OK;
}

extern "C" int fn_add_f32(float* dst, const float* src, long long n, void* stream) {
    FN_REQUIRE(dst && src && n > 0, "fn_add_f32: bad args");
    add_kernel<<<fn_cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(dst, src, n);
    FN_LAUNCH_CHECK();
    return FN_OK;
}

extern "C" size_t fn_emb_grad_scratch_bytes(int B, int T, int H, int V) {
    int rs, nc, rpc;
    if (emb_cfg(B, T, H, V, &rs, &nc, &rpc)) return 0;
    return (size_t)nc * V * 3 * H * sizeof(float);
}
extern "C" int fn_emb_grad_f32(const int32_t* ids_tm, const float* dgh, const float* dgin, int B, int T, int H, int V,
                               float* demb, void* scratch, size_t scratch_bytes, void* stream) {
    FN_REQUIRE(ids_tm && dgh && dgin && demb && scratch, "fn_emb_grad_f32: null pointer");
    int rs, nc, rpc;
    FN_REQUIRE(emb_cfg(B, T, H, V, &rs, &nc, &rpc) == 0, "fn_emb_grad_f32: V=%d too large for the smem table", V);
    FN_REQUIRE(scratch_bytes >= (size_t)nc * V * 3 * H * sizeof(float), "fn_emb_grad_f32: scratch too small");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = (size_t)rs * V * kEC * sizeof(float);
    FN_CHECK_CUDA(cudaFuncSetAttribute(emb_grad_stage1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(fn_cdiv(3 * H, kEC), nc);
    emb_grad_stage1<<<grid, kEC * rs, smem, st>>>(ids_tm, dgh, dgin, (long long)B * T, H, V, rpc, rs, (float*)scratch);
    FN_LAUNCH_CHECK();
    const long long n = (long long)V * 3 * H;
    emb_grad_stage2<<<fn_cdiv(n, 256), 256, 0, st>>>((const float*)scratch, nc, n, demb);
    FN_LAUNCH_CHECK();
    return FN_OK;
}

extern "C" int fn_time_sum_f32(const float* dgh, const float* dgin, int B, int T, int H, float* dproj, float* dghsum,
                               void* stream) {
    FN_REQUIRE(dgh && dgin && (dproj || dghsum), "fn_time_sum_f32: null pointer");
    time_sum_kernel<<<fn_cdiv((long long)B * 3 * H, 128), 128, 0, (cudaStream_t)stream>>>(dgh, dgin, B, T, H, dproj, dghsum);
    FN_LAUNCH_CHECK();
    return FN_OK;
}

extern "C" size_t fn_col_sum_scratch_bytes(long long rows, int cols) {
    (void)rows;
    return (size_t)kCS_R * cols * sizeof(float);
}
extern "C" int fn_col_sum_f32(const float* x, long long ld, long long rows, int cols, float* out, int accumulate,
                              void* scratch, size_t scratch_bytes, void* stream) {
    FN_REQUIRE(x && out && scratch && rows >= 0 && cols > 0, "fn_col_sum_f32: bad args");
    FN_REQUIRE(scratch_bytes >= (size_t)kCS_R * cols * sizeof(float), "fn_col_sum_f32: scratch too small");
    cudaStream_t st = (cudaStream_t)stream;
    int nchunks = (int)min((long long)kCS_R, max(1LL, rows / 64));
    const long long rpc = (rows + nchunks - 1) / max(nchunks, 1);
    dim3 grid(fn_cdiv(cols, 32), nchunks), block(32, 32);
    col_sum_stage1<<<grid, block, 0, st>>>(x, ld, rows, cols, rpc, (float*)scratch);
    FN_LAUNCH_CHECK();
    col_sum_stage2<<<fn_cdiv(cols, 128), 128, 0, st>>>((const float*)scratch, nchunks, cols, out, accumulate);
    FN_LAUNCH_CHECK();
    return FN_OK;
}
