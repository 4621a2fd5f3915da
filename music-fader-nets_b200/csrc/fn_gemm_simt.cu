// fp32 SIMT GEMM with arbitrary operand strides -- the exact-parity path (fn_gemm_f32).
// Register-blocked tiles staged through shared memory with register prefetch of the next
// K-slab; 128x128x16 tiles (8x8 per thread) for large problems, 64x64x16 (4x4) otherwise.
#include "fn_common.cuh"

namespace {

template <int BM, int BN, int BK, int TM, int TN, bool A_KC, bool B_KC>
__global__ void __launch_bounds__(256)
gemm_f32_kernel(const float* __restrict__ A, long long sam, long long sak, const float* __restrict__ B,
                long long sbk, long long sbn, float* __restrict__ C, long long ldc, const float* __restrict__ bias,
                int M, int N, int K, int accumulate, int k_per_split, float* __restrict__ partial) {
    // split-K (fn_gemm_f32_splitk): CTA z reduces K range [z * k_per_split, +k_per_split) into partial[z][M][N]
    const int kbeg = blockIdx.z * k_per_split, Kend = min(K, kbeg + k_per_split);
    constexpr int NT = 256;
    static_assert((BM / TM) * (BN / TN) == NT, "thread tiling");
    constexpr int LDA = BM + 4, LDB = BN + 4;
    __shared__ __align__(16) float As[BK][LDA];
    __shared__ __align__(16) float Bs[BK][LDB];

    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int tx = tid % (BN / TN), ty = tid / (BN / TN);

    constexpr int A_V = BM * BK / 4 / NT;   // float4 loads per thread
    constexpr int B_V = BN * BK / 4 / NT;
    static_assert(A_V >= 1 && B_V >= 1, "tile too small");
    float4 ra[A_V], rb[B_V];

    const bool a_vec = A_KC ? ((sam & 3) == 0 && ((uintptr_t)A & 15) == 0) : ((sak & 3) == 0 && ((uintptr_t)A & 15) == 0);
    const bool b_vec = B_KC ? ((sbn & 3) == 0 && ((uintptr_t)B & 15) == 0) : ((sbk & 3) == 0 && ((uintptr_t)B & 15) == 0);

    auto load_a = [&](int k0) {
#pragma unroll
        for (int i = 0; i < A_V; ++i) {
            const int idx = tid + i * NT;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (A_KC) {
                const int m = m0 + idx / (BK / 4), k = k0 + (idx % (BK / 4)) * 4;
                if (m < M) {
                    const float* p = A + (long long)m * sam + k;
                    if (a_vec && k + 3 < Kend) v = *reinterpret_cast<const float4*>(p);
                    else {
                        if (k < Kend) v.x = p[0];
                        if (k + 1 < Kend) v.y = p[1];
                        if (k + 2 < Kend) v.z = p[2];
                        if (k + 3 < Kend) v.w = p[3];
                    }
                }
            } else {
                const int k = k0 + idx / (BM / 4), m = m0 + (idx % (BM / 4)) * 4;
                if (k < Kend) {
                    const float* p = A + (long long)k * sak + m;
                    if (a_vec && m + 3 < M) v = *reinterpret_cast<const float4*>(p);
                    else {
                        if (m < M) v.x = p[0];
                        if (m + 1 < M) v.y = p[1];
                        if (m + 2 < M) v.z = p[2];
                        if (m + 3 < M) v.w = p[3];
                    }
                }
            }
            ra[i] = v;
        }
    };
    auto load_b = [&](int k0) {
#pragma unroll
        for (int i = 0; i < B_V; ++i) {
            const int idx = tid + i * NT;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (B_KC) {
                const int n = n0 + idx / (BK / 4), k = k0 + (idx % (BK / 4)) * 4;
                if (n < N) {
                    const float* p = B + (long long)n * sbn + k;
                    if (b_vec && k + 3 < Kend) v = *reinterpret_cast<const float4*>(p);
                    else {
                        if (k < Kend) v.x = p[0];
                        if (k + 1 < Kend) v.y = p[1];
                        if (k + 2 < Kend) v.z = p[2];
                        if (k + 3 < Kend) v.w = p[3];
                    }
                }
            } else {
                const int k = k0 + idx / (BN / 4), n = n0 + (idx % (BN / 4)) * 4;
                if (k < Kend) {
                    const float* p = B + (long long)k * sbk + n;
                    if (b_vec && n + 3 < N) v = *reinterpret_cast<const float4*>(p);
                    else {
                        if (n < N) v.x = p[0];
                        if (n + 1 < N) v.y = p[1];
                        if (n + 2 < N) v.z = p[2];
                        if (n + 3 < N) v.w = p[3];
                    }
                }
            }
            rb[i] = v;
        }
    };
    auto store_ab = [&]() {
#pragma unroll
        for (int i = 0; i < A_V; ++i) {
            const int idx = tid + i * NT;
            if (A_KC) {
                const int m = idx / (BK / 4), k = (idx % (BK / 4)) * 4;
                As[k][m] = ra[i].x; As[k + 1][m] = ra[i].y; As[k + 2][m] = ra[i].z; As[k + 3][m] = ra[i].w;
            } else {
                const int k = idx / (BM / 4), m = (idx % (BM / 4)) * 4;
                *reinterpret_cast<float4*>(&As[k][m]) = ra[i];
            }
        }
#pragma unroll
        for (int i = 0; i < B_V; ++i) {
            const int idx = tid + i * NT;
            if (B_KC) {
                const int n = idx / (BK / 4), k = (idx % (BK / 4)) * 4;
                Bs[k][n] = rb[i].x; Bs[k + 1][n] = rb[i].y; Bs[k + 2][n] = rb[i].z; Bs[k + 3][n] = rb[i].w;
            } else {
                const int k = idx / (BN / 4), n = (idx % (BN / 4)) * 4;
                *reinterpret_cast<float4*>(&Bs[k][n]) = rb[i];
            }
        }
    };

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    // thread's rows: TM/4 groups of 4 at ty*4 + g*(BM/(TM/4)) ; cols likewise
    constexpr int GM = TM / 4, GN = TN / 4;
    load_a(kbeg);
    load_b(kbeg);
    for (int k0 = kbeg; k0 < Kend; k0 += BK) {
        store_ab();
        __syncthreads();
        if (k0 + BK < Kend) { load_a(k0 + BK); load_b(k0 + BK); }
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[TM], b[TN];
#pragma unroll
            for (int g = 0; g < GM; ++g) {
                const float4 v = *reinterpret_cast<const float4*>(&As[k][ty * 4 + g * (BM / GM)]);
                a[g * 4] = v.x; a[g * 4 + 1] = v.y; a[g * 4 + 2] = v.z; a[g * 4 + 3] = v.w;
            }
#pragma unroll
            for (int g = 0; g < GN; ++g) {
                const float4 v = *reinterpret_cast<const float4*>(&Bs[k][tx * 4 + g * (BN / GN)]);
                b[g * 4] = v.x; b[g * 4 + 1] = v.y; b[g * 4 + 2] = v.z; b[g * 4 + 3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int m = m0 + ty * 4 + (i / 4) * (BM / GM) + (i % 4);
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int n = n0 + tx * 4 + (j / 4) * (BN / GN) + (j % 4);
            if (n >= N) continue;
            float v = acc[i][j];
            if (partial) { partial[((long long)blockIdx.z * M + m) * N + n] = v; continue; }
            if (bias) v += bias[n];
            float* c = C + (long long)m * ldc + n;
            if (accumulate) v += *c;
            *c = v;
        }
    }
}

template <int BM, int BN, int TM, int TN>
int launch(const float* A, long long sam, long long sak, const float* B, long long sbk, long long sbn, float* C,
           long long ldc, const float* bias, int M, int N, int K, int accumulate, cudaStream_t st, int splits = 1,
           float* partial = nullptr) {
    dim3 grid(fn_cdiv(N, BN), fn_cdiv(M, BM), splits);
    const int kps = splits > 1 ? (fn_cdiv(K, splits) + 15) / 16 * 16 : (K > 0 ? K : 1);
    const bool akc = (sak == 1), bkc = (sbk == 1);
    if (akc && bkc)
        gemm_f32_kernel<BM, BN, 16, TM, TN, true, true><<<grid, 256, 0, st>>>(A, sam, sak, B, sbk, sbn, C, ldc, bias, M, N, K, accumulate, kps, partial);
    else if (akc && !bkc)
        gemm_f32_kernel<BM, BN, 16, TM, TN, true, false><<<grid, 256, 0, st>>>(A, sam, sak, B, sbk, sbn, C, ldc, bias, M, N, K, accumulate, kps, partial);
    else if (!akc && bkc)
        gemm_f32_kernel<BM, BN, 16, TM, TN, false, true><<<grid, 256, 0, st>>>(A, sam, sak, B, sbk, sbn, C, ldc, bias, M, N, K, accumulate, kps, partial);
    else
        gemm_f32_kernel<BM, BN, 16, TM, TN, false, false><<<grid, 256, 0, st>>>(A, sam, sak, B, sbk, sbn, C, ldc, bias, M, N, K, accumulate, kps, partial);
    FN_LAUNCH_CHECK();
    return FN_OK;
}

// C = (accumulate ? C : 0) + bias + sum_z partial[z]  (fixed order: deterministic)
__global__ void gemm_f32_splitk_reduce(const float* __restrict__ partial, int splits, int M, int N, float* __restrict__ C,
                                       long long ldc, const float* __restrict__ bias, int accumulate) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)M * N) return;
    const int m = (int)(i / N), n = (int)(i % N);
    float v = bias ? __ldg(bias + n) : 0.f;
    for (int z = 0; z < splits; ++z) v += partial[(long long)z * M * N + i];
    float* c = C + (long long)m * ldc + n;
    *c = accumulate ? *c + v : v;
}

}  // namespace

// Products with a long K and few output tiles (the latent heads: [B x 2H] x [2H x Z]; dproj W_ih: [B x 3H] x [3H x G]) occupy
// a handful of CTAs whose K loops are latency chains; `splits` CTAs per tile each reduce a K range into the caller's
// workspace (splits * M * N floats), then a fixed-order reduction applies bias / accumulate -- deterministic.
extern "C" size_t fn_gemm_f32_splitk_ws_bytes(int M, int N, int splits) {
    return splits > 1 ? (size_t)splits * M * N * sizeof(float) : 0;
}
extern "C" int fn_gemm_f32_splitk(const float* A, long long sam, long long sak, const float* B, long long sbk, long long sbn,
                                  float* C, long long ldc, const float* bias, int M, int N, int K, int accumulate, int splits,
                                  void* workspace, size_t ws_bytes, void* stream) {
    if (splits <= 1 || K < 32) return fn_gemm_f32(A, sam, sak, B, sbk, sbn, C, ldc, bias, M, N, K, accumulate, stream);
    FN_REQUIRE(C && A && B && M > 0 && N > 0, "fn_gemm_f32_splitk: bad args");
    FN_REQUIRE(sam == 1 || sak == 1, "fn_gemm_f32_splitk: A needs a unit stride");
    FN_REQUIRE(sbk == 1 || sbn == 1, "fn_gemm_f32_splitk: B needs a unit stride");
    if (splits > K / 16) splits = K / 16;
    const int kps = (fn_cdiv(K, splits) + 15) / 16 * 16;
    splits = fn_cdiv(K, kps);                                        // no empty split
    FN_REQUIRE(workspace && ws_bytes >= fn_gemm_f32_splitk_ws_bytes(M, N, splits), "fn_gemm_f32_splitk: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const int rc = launch<64, 64, 4, 4>(A, sam, sak, B, sbk, sbn, C, ldc, bias, M, N, K, accumulate, st, splits, (float*)workspace);
    if (rc) return rc;
    gemm_f32_splitk_reduce<<<fn_cdiv((long long)M * N, 256), 256, 0, st>>>((const float*)workspace, splits, M, N, C, ldc, bias,
                                                                           accumulate);
    FN_LAUNCH_CHECK();
    return FN_OK;
}

extern "C" int fn_gemm_f32(const float* A, long long sam, long long sak, const float* B, long long sbk,
                           long long sbn, float* C, long long ldc, const float* bias, int M, int N, int K,
                           int accumulate, void* stream) {
    FN_REQUIRE(C && (K == 0 || (A && B)), "fn_gemm_f32: null operand");
    FN_REQUIRE(M >= 0 && N >= 0 && K >= 0, "fn_gemm_f32: negative size");
    FN_REQUIRE(sam == 1 || sak == 1, "fn_gemm_f32: A needs a unit stride (sam=%lld sak=%lld)", sam, sak);
    FN_REQUIRE(sbk == 1 || sbn == 1, "fn_gemm_f32: B needs a unit stride (sbk=%lld sbn=%lld)", sbk, sbn);
    if (M == 0 || N == 0) return FN_OK;
    cudaStream_t st = (cudaStream_t)stream;
    // a unit stride on both axes (vector operand) -> treat as k-contiguous
    const long long tiles128 = (long long)fn_cdiv(M, 128) * fn_cdiv(N, 128);
    if (tiles128 >= 120) return launch<128, 128, 8, 8>(A, sam, sak, B, sbk, sbn, C, ldc, bias, M, N, K, accumulate, st);
    return launch<64, 64, 4, 4>(A, sam, sak, B, sbk, sbn, C, ldc, bias, M, N, K, accumulate, st);
}
