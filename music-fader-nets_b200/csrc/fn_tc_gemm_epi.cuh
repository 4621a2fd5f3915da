// Epilogue store of the tcgen05 GEMMs (fn_tc_gemm.cu: 128x128 tiles, one CTA; fn_tc_gemm2.cu: 256x256 tiles, CTA pairs):
// one thread owns 32 consecutive fp32 accumulator columns of one output row.
#pragma once
#include "fn_common.cuh"

namespace {

struct GemmEpi {
    void* C;
    const float* bias;
    long long ldc;
    int M, N, c_bf16, accumulate, splits;
    float* partial;                // [splits][M][N] fp32 partial products (splits > 1)
};

// r = the thread's 32 accumulator columns [col0, col0 + 32) of `row` (row < M, col0 < n_lim); columns >= n_lim (<= N: the end of
// the matrix or of this tile) are not stored; z = split-K index
__device__ __forceinline__ void gemm_store_chunk(const GemmEpi& p, int z, int row, int col0, int n_lim, uint32_t (&r)[32]) {
    if (p.splits > 1) {
        float* prow = p.partial + ((long long)z * p.M + row) * p.N + col0;
        if (((reinterpret_cast<uintptr_t>(prow) & 15) == 0) && (col0 + 32 <= n_lim)) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<uint4*>(prow + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (col0 + j < n_lim) prow[j] = __uint_as_float(r[j]);
        }
    } else if (p.c_bf16) {
        __nv_bfloat16* crow = reinterpret_cast<__nv_bfloat16*>(p.C) + (long long)row * p.ldc + col0;
        const bool vec = ((reinterpret_cast<uintptr_t>(crow) & 15) == 0) && (col0 + 32 <= n_lim);
        if (vec) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
                float v[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(r[j + e]);
                if (p.bias) {
                    const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
                    const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j + 4));
                    v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
                    v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
                }
                uint4* dst = reinterpret_cast<uint4*>(crow + j);
                if (p.accumulate) {
                    const uint4 o = *dst;
                    const uint32_t ow[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        v[2 * e] += __uint_as_float(ow[e] << 16);
                        v[2 * e + 1] += __uint_as_float(ow[e] & 0xffff0000u);
                    }
                }
                uint32_t w[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
                    w[e] = *reinterpret_cast<const uint32_t*>(&h);
                }
                *dst = make_uint4(w[0], w[1], w[2], w[3]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                if (col0 + j < n_lim) {
                    float v = __uint_as_float(r[j]);
                    if (p.bias) v += __ldg(p.bias + col0 + j);
                    if (p.accumulate) v += __bfloat162float(crow[j]);
                    crow[j] = __float2bfloat16(v);
                }
            }
        }
    } else {
        float* crow = reinterpret_cast<float*>(p.C) + (long long)row * p.ldc + col0;
        const bool vec = ((reinterpret_cast<uintptr_t>(crow) & 15) == 0) && (col0 + 32 <= n_lim);
        if (vec) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                float4 v = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                       __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
                if (p.bias) {
                    const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
                    v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w;
                }
                if (p.accumulate) {
                    const float4 o = *reinterpret_cast<const float4*>(crow + j);
                    v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
                }
                *reinterpret_cast<float4*>(crow + j) = v;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                if (col0 + j < n_lim) {
                    float v = __uint_as_float(r[j]);
                    if (p.bias) v += __ldg(p.bias + col0 + j);
                    if (p.accumulate) v += crow[j];
                    crow[j] = v;
                }
            }
        }
    }
}

}  // namespace
