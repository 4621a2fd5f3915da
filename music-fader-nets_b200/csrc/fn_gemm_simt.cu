// fp32 SIMT GEMM with arbitrary operand strides -- the exact-parity path (fn_gemm_f32).
// Register-blocked tiles staged through shared memory with register prefetch of the next
// K-slab; 128x128x16 tiles (8x8 per thread) for large problems, 64x64x16 (4x4) otherwise.
#include "fn_common.cuh"

namespace {

template <int BM, int BN, int BK, int TM, int TN, bool A_KC, bool B_KC>
__global__ void __launch_bounds__(256)
gemm_f32_kernel(const float* __restrict__ A, long long sam, long long sak, const float* __restrict__ B,
                long long sbk, long long sbn, float* __restrict__ C, long long ldc, const float* __restrict__ bias,
                int M, int N, int K, int accumulate) {
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
                    if (a_vec && k + 3 < K) v = *reinterpret_cast<const float4*>(p);
                    else {
                        if (k < K) v.x = p[0];
                        if (k + 1 < K) v.y = p[1];
                        if (k + 2 < K) v.z = p[2];
                        if (k + 3 < K) v.w = p[3];
                    }
                }
            } else {
                const int k = k0 + idx / (BM / 4), m = m0 + (idx % (BM / 4)) * 4;
                if (k < K) {
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
                    if (b_vec && k + 3 < K) v = *reinterpret_cast<const float4*>(p);
                    else {
                        if (k < K) v.x = p[0];
                        if (k + 1 < K) v.y = p[1];
                        if (k + 2 < K) v.z = p[2];
                        if (k + 3 < K) v.w = p[3];
                    }
                }
            } else {
                const int k = k0 + idx / (BN / 4), n = n0 + (idx % (BN / 4)) * 4;
                if (k < K) {
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
    load_a(0);
    load_b(0);
    for (int k0 = 0; k0 < K; k0 += BK) {
        store_ab();
        __syncthreads();
        if (k0 + BK < K) { load_a(k0 + BK); load_b(k0 + BK); }
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
            if (bias) v += bias[n];
            float* c = C + (long long)m * ldc + n;
            if (accumulate) v += *c;
            *c = v;
        }
    }
}

template <int BM, int BN, int TM, int TN>
int launch(const float* A, long long sam, long long sak, const float* B, long long sbk, long long sbn, float* C,
           long long ldc, const float* bias, int M, int N, int K, int accumulate, cudaStream_t st) {
    dim3 grid(fn_cdiv(N, BN), fn_cdiv(M, BM));
    const bool akc = (sak == 1), bkc = (sbk == 1);
    if (akc && bkc)
        gemm_f32_kernel<BM, BN, 16, TM, TN, true, true><<<grid, 256, 0, st>>>(A, sam, sak, B, sbk, sbn, C, ldc, bias, M, N, K, accumulate);
    else if (akc && !bkc)
        gemm_f32_kernel<BM, BN, 16, TM, TN, true, false><<<grid, 256, 0, st>>>(A, sam, sak, B, sbk, sbn, C, ldc, bias, M, N, K, accumulate);
    else if (!akc && bkc)
        gemm_f32_kernel<BM, BN, 16, TM, TN, false, true><<<grid, 256, 0, st>>>(A, sam, sak, B, sbk, sbn, C, ldc, bias, M, N, K, accumulate);
    else
        gemm_f32_kernel<BM, BN, 16, TM, TN, false, false><<<grid, 256, 0, st>>>(A, sam, sak, B, sbk, sbn, C, ldc, bias, M, N, K, accumulate);
    FN_LAUNCH_CHECK();
    return FN_OK;
}

}  // namespace

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
