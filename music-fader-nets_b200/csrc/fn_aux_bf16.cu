// Bandwidth-bound helpers of the bf16 tensor-core path: fp32 -> bf16 casts (optionally transposing),
// bf16 one-hot operands, and the time / column sums over the bf16 gate-gradient stream.
#include "fn_common.cuh"

namespace {

// dst[r][c] = bf16(src[r*s_r + c*s_c]).  32x32 tiles through smem so both sides stay coalesced whichever
// stride is the unit one.
__global__ void cast_bf16_kernel(const float* __restrict__ src, long long s_r, long long s_c,
                                 __nv_bfloat16* __restrict__ dst, long long ld_dst, long long rows, long long cols) {
    __shared__ float tile[32][33];
    const long long c0 = (long long)blockIdx.x * 32, r0 = (long long)blockIdx.y * 32;
    if (s_c == 1) {
        for (int i = threadIdx.y; i < 32; i += blockDim.y) {
            const long long r = r0 + i, c = c0 + threadIdx.x;
            if (r < rows && c < cols) dst[r * ld_dst + c] = __float2bfloat16(src[r * s_r + c]);
        }
        return;
    }
    // column-strided source (transposing copy): read along rows (s_r is the small stride), write along cols
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const long long c = c0 + i, r = r0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < rows && c < cols) ? src[r * s_r + c * s_c] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const long long r = r0 + i, c = c0 + threadIdx.x;
        if (r < rows && c < cols) dst[r * ld_dst + c] = __float2bfloat16(tile[threadIdx.x][i]);
    }
}

// vectorised contiguous cast: 8 elements per thread
__global__ void cast_bf16_vec_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n8) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n8) return;
    const float4 a = __ldg(reinterpret_cast<const float4*>(src) + 2 * i), b = __ldg(reinterpret_cast<const float4*>(src) + 2 * i + 1);
    const __nv_bfloat162 p0 = __floats2bfloat162_rn(a.x, a.y), p1 = __floats2bfloat162_rn(a.z, a.w);
    const __nv_bfloat162 p2 = __floats2bfloat162_rn(b.x, b.y), p3 = __floats2bfloat162_rn(b.z, b.w);
    uint4 o;
    o.x = *reinterpret_cast<const uint32_t*>(&p0); o.y = *reinterpret_cast<const uint32_t*>(&p1);
    o.z = *reinterpret_cast<const uint32_t*>(&p2); o.w = *reinterpret_cast<const uint32_t*>(&p3);
    reinterpret_cast<uint4*>(dst)[i] = o;
}

// bf16x3 mode: an fp32 value is carried as TWO bf16 planes, hi = bf16(x) and lo = bf16(x - hi) (16 mantissa bits together);
// dst[r*ld + c] = hi, dst[r*ld + lo_off + c] = lo, and (weights of the recurrent kernels, whose K loop runs over the three
// plane products hi*hi, lo*hi, hi*lo) optionally a second copy of hi at hi2_off.  Same tiling as cast_bf16_kernel.
__device__ __forceinline__ void split_store(__nv_bfloat16* d, long long lo_off, long long hi2_off, float v) {
    const __nv_bfloat16 hi = __float2bfloat16(v);
    d[0] = hi;
    d[lo_off] = __float2bfloat16(v - __bfloat162float(hi));
    if (hi2_off >= 0) d[hi2_off] = hi;
}
__global__ void split_bf16_kernel(const float* __restrict__ src, long long s_r, long long s_c, __nv_bfloat16* __restrict__ dst,
                                  long long ld_dst, long long rows, long long cols, long long lo_off, long long hi2_off) {
    __shared__ float tile[32][33];
    const long long c0 = (long long)blockIdx.x * 32, r0 = (long long)blockIdx.y * 32;
    if (s_c == 1) {
        for (int i = threadIdx.y; i < 32; i += blockDim.y) {
            const long long r = r0 + i, c = c0 + threadIdx.x;
            if (r < rows && c < cols) split_store(dst + r * ld_dst + c, lo_off, hi2_off, src[r * s_r + c]);
        }
        return;
    }
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const long long c = c0 + i, r = r0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < rows && c < cols) ? src[r * s_r + c * s_c] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const long long r = r0 + i, c = c0 + threadIdx.x;
        if (r < rows && c < cols) split_store(dst + r * ld_dst + c, lo_off, hi2_off, tile[threadIdx.x][i]);
    }
}

__global__ void ids_to_onehot_bf16_kernel(const int32_t* __restrict__ ids, long long rows, int V, long long ld,
                                          __nv_bfloat16* __restrict__ oh) {
    // one warp per row, 8 bf16 (16 B) per lane per pass
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const int id = ids[row];
    __nv_bfloat16* o = oh + row * ld;
    for (long long c = lane * 8LL; c < ld; c += 256) {
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        const int d = id - (int)c;
        if (d >= 0 && d < 8 && id < V) {
            const uint32_t one = 0x3f80u << ((d & 1) * 16);     // bf16(1.0) = 0x3f80
            if ((d >> 1) == 0) v.x = one; else if ((d >> 1) == 1) v.y = one; else if ((d >> 1) == 2) v.z = one; else v.w = one;
        }
        *reinterpret_cast<uint4*>(o + c) = v;
    }
}

// dg [T][B][4H] bf16, columns (dr, dz, dn, dn*r).  One thread per (b, column pair).
// X3: rows of 8H = the hi planes followed by the lo planes (bf16x3 mode); the sum adds both.
template <bool X3>
__global__ void time_sum_bf16_kernel(const __nv_bfloat16* __restrict__ dg, int B, int T, int H, float* __restrict__ dproj,
                                     float* __restrict__ dghsum) {
    const int H4 = 4 * H, K3 = 3 * H, P = X3 ? 8 * H : 4 * H;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * (H4 / 2)) return;
    const int b = (int)(i / (H4 / 2)), c = (int)(i % (H4 / 2)) * 2;
    const __nv_bfloat16* p = dg + (long long)b * P + c;
    const long long stride = (long long)B * P;
    float s0 = 0.f, s1 = 0.f;
    int t = 0;
    for (; t + 4 <= T; t += 4) {
        uint32_t w[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            w[j] = __ldg(reinterpret_cast<const uint32_t*>(p + (t + j) * stride));
            if (X3) l[j] = __ldg(reinterpret_cast<const uint32_t*>(p + (t + j) * stride + H4));
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float a0 = __uint_as_float(w[j] << 16), a1 = __uint_as_float(w[j] & 0xffff0000u);
            if (X3) { a0 += __uint_as_float(l[j] << 16); a1 += __uint_as_float(l[j] & 0xffff0000u); }
            s0 += a0; s1 += a1;
        }
    }
    for (; t < T; ++t) {
        const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(p + t * stride));
        float a0 = __uint_as_float(w << 16), a1 = __uint_as_float(w & 0xffff0000u);
        if (X3) {
            const uint32_t l = __ldg(reinterpret_cast<const uint32_t*>(p + t * stride + H4));
            a0 += __uint_as_float(l << 16); a1 += __uint_as_float(l & 0xffff0000u);
        }
        s0 += a0; s1 += a1;
    }
    float* rowp = dproj ? dproj + (long long)b * K3 : nullptr;
    float* rowh = dghsum ? dghsum + (long long)b * K3 : nullptr;
    if (c < 2 * H) {
        if (rowp) { rowp[c] = s0; rowp[c + 1] = s1; }
        if (rowh) { rowh[c] = s0; rowh[c + 1] = s1; }
    } else if (c < K3) {
        if (rowp) { rowp[c] = s0; rowp[c + 1] = s1; }
    } else {
        if (rowh) { rowh[c - H] = s0; rowh[c - H + 1] = s1; }
    }
}

__global__ void add_f32_to_bf16_kernel(__nv_bfloat16* __restrict__ dst, const float* __restrict__ src, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = __float2bfloat16(__bfloat162float(dst[i]) + src[i]);
}

constexpr int kCSB_R = 64;
__global__ void col_sum_bf16_stage1(const __nv_bfloat16* __restrict__ x, long long ld, long long rows, int cols,
                                    long long rows_per_chunk, float* __restrict__ partial) {
    __shared__ float red[32][33];
    const int c = blockIdx.x * 32 + threadIdx.x;
    const long long r0 = (long long)blockIdx.y * rows_per_chunk, r1 = min(rows, r0 + rows_per_chunk);
    float s = 0.f;
    if (c < cols)
        for (long long r = r0 + threadIdx.y; r < r1; r += 32) s += __bfloat162float(x[r * ld + c]);
    red[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && c < cols) {
        float t = 0.f;
        for (int q = 0; q < 32; ++q) t += red[q][threadIdx.x];
        partial[(long long)blockIdx.y * cols + c] = t;
    }
}
__global__ void col_sum_bf16_stage2(const float* __restrict__ partial, int nchunks, int cols, float* __restrict__ out,
                                    int accumulate) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    float s = 0.f;
    for (int q = 0; q < nchunks; ++q) s += partial[(long long)q * cols + c];
    out[c] = accumulate ? out[c] + s : s;
}

}  // namespace

extern "C" int fn_cast_bf16(const float* src, long long s_r, long long s_c, void* dst, long long ld_dst, long long rows,
                            long long cols, void* stream) {
    FN_REQUIRE(src && dst && rows > 0 && cols > 0, "fn_cast_bf16: bad args");
    cudaStream_t st = (cudaStream_t)stream;
    if (s_c == 1 && s_r == cols && ld_dst == cols && (rows * cols) % 8 == 0 &&
        (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        const long long n8 = rows * cols / 8;
        cast_bf16_vec_kernel<<<fn_cdiv(n8, 256), 256, 0, st>>>(src, (__nv_bfloat16*)dst, n8);
    } else {
        dim3 grid(fn_cdiv(cols, 32), fn_cdiv(rows, 32)), block(32, 8);
        FN_REQUIRE(grid.y <= 65535, "fn_cast_bf16: too many rows for the strided path (%lld)", rows);
        cast_bf16_kernel<<<grid, block, 0, st>>>(src, s_r, s_c, (__nv_bfloat16*)dst, ld_dst, rows, cols);
    }
    FN_LAUNCH_CHECK();
    return FN_OK;
}

extern "C" int fn_ids_to_onehot_bf16(const int32_t* ids, long long rows, int V, long long ld, void* onehot, void* stream) {
    FN_REQUIRE(ids && onehot && rows > 0 && V > 0 && ld >= V && ld % 8 == 0, "fn_ids_to_onehot_bf16: bad args (ld %% 8)");
    ids_to_onehot_bf16_kernel<<<fn_cdiv(rows * 32, 256), 256, 0, (cudaStream_t)stream>>>(ids, rows, V, ld,
                                                                                       (__nv_bfloat16*)onehot);
    FN_LAUNCH_CHECK();
    return FN_OK;
}

extern "C" int fn_time_sum_bf16(const void* dg, int B, int T, int H, float* dproj, float* dghsum, void* stream) {
    FN_REQUIRE(dg && (dproj || dghsum) && B > 0 && T > 0 && H > 0 && H % 2 == 0, "fn_time_sum_bf16: bad args");
    const long long n = (long long)B * 2 * H;
    time_sum_bf16_kernel<false><<<fn_cdiv(n, 128), 128, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)dg, B, T, H, dproj, dghsum);
    FN_LAUNCH_CHECK();
    return FN_OK;
}

extern "C" int fn_time_sum_bf16x3(const void* dg, int B, int T, int H, float* dproj, float* dghsum, void* stream) {
    FN_REQUIRE(dg && (dproj || dghsum) && B > 0 && T > 0 && H > 0 && H % 2 == 0, "fn_time_sum_bf16x3: bad args");
    const long long n = (long long)B * 2 * H;
    time_sum_bf16_kernel<true><<<fn_cdiv(n, 128), 128, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)dg, B, T, H, dproj, dghsum);
    FN_LAUNCH_CHECK();
    return FN_OK;
}

extern "C" int fn_split_bf16(const float* src, long long s_r, long long s_c, void* dst, long long ld_dst, long long rows,
                             long long cols, long long lo_off, long long hi2_off, void* stream) {
    FN_REQUIRE(src && dst && rows > 0 && cols > 0 && lo_off != 0, "fn_split_bf16: bad args");
    dim3 block(32, 8);
    // rows can be T*B (hundreds of thousands): launch in groups of 65535 row blocks
    const long long row_blocks = fn_cdiv(rows, 32);
    for (long long rb0 = 0; rb0 < row_blocks; rb0 += 65535) {
        const long long nb = row_blocks - rb0 < 65535 ? row_blocks - rb0 : 65535;
        dim3 g(fn_cdiv(cols, 32), (unsigned)nb);
        const long long r0 = rb0 * 32;
        split_bf16_kernel<<<g, block, 0, (cudaStream_t)stream>>>(src + r0 * s_r, s_r, s_c, (__nv_bfloat16*)dst + r0 * ld_dst, ld_dst,
                                                                 rows - r0 < nb * 32 ? rows - r0 : nb * 32, cols, lo_off, hi2_off);
        FN_LAUNCH_CHECK();
    }
    return FN_OK;
}

extern "C" int fn_add_f32_to_bf16(void* dst, const float* src, long long n, void* stream) {
    FN_REQUIRE(dst && src && n > 0, "fn_add_f32_to_bf16: bad args");
    add_f32_to_bf16_kernel<<<fn_cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)dst, src, n);
    FN_LAUNCH_CHECK();
    return FN_OK;
}

extern "C" int fn_col_sum_bf16(const void* x, long long ld, long long rows, int cols, float* out, int accumulate,
                               void* scratch, size_t scratch_bytes, void* stream) {
    FN_REQUIRE(x && out && scratch && rows >= 0 && cols > 0, "fn_col_sum_bf16: bad args");
    FN_REQUIRE(scratch_bytes >= (size_t)kCSB_R * cols * sizeof(float), "fn_col_sum_bf16: scratch too small (need %zu)",
               (size_t)kCSB_R * cols * sizeof(float));
    cudaStream_t st = (cudaStream_t)stream;
    int nchunks = (int)min((long long)kCSB_R, max(1LL, rows / 64));
    const long long rpc = (rows + nchunks - 1) / max(nchunks, 1);
    dim3 grid(fn_cdiv(cols, 32), nchunks), block(32, 32);
    col_sum_bf16_stage1<<<grid, block, 0, st>>>((const __nv_bfloat16*)x, ld, rows, cols, rpc, (float*)scratch);
    FN_LAUNCH_CHECK();
    col_sum_bf16_stage2<<<fn_cdiv(cols, 128), 128, 0, st>>>((const float*)scratch, nchunks, cols, out, accumulate);
    FN_LAUNCH_CHECK();
    return FN_OK;
}
