// Micro-benchmark (profiling aid, not part of the product): issue cost / throughput of tcgen05.mma as a function
// of N, from ONE issuing thread of one CTA per SM, operands in (garbage) shared memory, accumulator in TMEM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I music-fader-nets_b200/csrc -o /tmp/ubench_tc tools/ubench_tc.cu
#include <cstdio>
#include <cstdlib>
#include "fn_tc.cuh"
#include <mutex>
void fn_set_error(const char*, ...) {}
int fn_num_sms() { return 148; }
int fn_max_smem_optin() { return 232448; }
fn_PFN_encodeTiled fn_get_encode_tiled() { void* p = nullptr; cudaDriverEntryPointQueryResult q; cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q); return (fn_PFN_encodeTiled)p; }
int fn_make_tmap_bf16_2d(CUtensorMap* out, const void* base, unsigned long long rows, unsigned long long cols, unsigned long long ld, unsigned box_rows, unsigned box_cols) {
    cuuint64_t dims[2] = {cols, rows}; cuuint64_t strides[1] = {ld * 2}; cuuint32_t box[2] = {box_cols, box_rows}; cuuint32_t estr[2] = {1, 1};
    return (int)fn_get_encode_tiled()(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); }

template <int CONVERGED>
__global__ void __launch_bounds__(128, 1) k_mma(int N, int n_mma, int reps, long long* out, int a_from_tmem) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
    if (warp == 1) tc::tmem_alloc(&tslot, 512);
    tc::tc_fence_before(); __syncthreads(); tc::tc_fence_after();
    const uint32_t tmem = tslot;
    if (warp == 0) {
        const uint32_t idesc = tc::make_idesc_bf16(128, N, 0, 0);
        const uint64_t ad = tc::make_sdesc(tc::smem_u32(smem), 16, 1024), bd = tc::make_sdesc(tc::smem_u32(smem) + 65536, 16, 1024);
        long long best = 1ll << 60, best_issue = 1ll << 60;
        for (int r = 0; r < reps; ++r) {
            __syncwarp();
            const long long t0 = clock64();
            if (CONVERGED) {
                for (int i = 0; i < n_mma; ++i) {
                    uint32_t pred;
                    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
                    if (pred) tc::umma_f16(tmem, ad + (uint64_t)(2 * (i & 3) + 1024 * ((i >> 2) & 3)), bd + (uint64_t)(2 * (i & 3)), idesc, i != 0);
                    __syncwarp();
                }
            } else if (lane == 0) {
                for (int i = 0; i < n_mma; ++i)
                    tc::umma_f16(tmem, ad + (uint64_t)(2 * (i & 3) + 1024 * ((i >> 2) & 3)), bd + (uint64_t)(2 * (i & 3)), idesc, i != 0);
            }
            const long long t1 = clock64();
            if (lane == 0) { tc::umma_commit(&bar); }
            tc::mbar_wait(&bar, r & 1);
            const long long t2 = clock64();
            if (t2 - t0 < best) best = t2 - t0;
            if (t1 - t0 < best_issue) best_issue = t1 - t0;
        }
        if (lane == 0 && blockIdx.x == 0) { out[0] = best; out[1] = best_issue; }
    }
    tc::tc_fence_before(); __syncthreads();
    if (warp == 1) { tc::tc_fence_after(); tc::tmem_dealloc(tmem, 512); }
}

// cost of mbarrier ops / TMA-less handshakes from one thread
__global__ void k_bar(int n, long long* out) {
    __shared__ uint64_t bar[2];
    if (threadIdx.x == 0) {
        tc::mbar_init(&bar[0], 1); tc::mbar_init(&bar[1], 1); tc::fence_barrier_init();
        long long t0 = clock64();
        uint32_t ph = 0;
        for (int i = 0; i < n; ++i) { tc::mbar_arrive(&bar[0]); tc::mbar_wait(&bar[0], ph); ph ^= 1; }
        long long t1 = clock64();
        out[0] = t1 - t0;
        t0 = clock64();
        for (int i = 0; i < n; ++i) { tc::mbar_arrive_expect_tx(&bar[1], 0); }
        t1 = clock64();
        out[1] = t1 - t0;
    }
}

// producer <-> MMA-thread ring handshake, optional MMAs, optional bystander warps waiting on another barrier
__global__ void __launch_bounds__(320, 1) k_ring(int n, int stages, int mmas, int N, int bystanders, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t full[8], empty[8], other, done;
    __shared__ uint32_t tslot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
        tc::mbar_init(&other, 1); tc::mbar_init(&done, 1);
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc(&tslot, 512);
    tc::tc_fence_before(); __syncthreads(); tc::tc_fence_after();
    const uint32_t tmem = tslot;
    if (warp == 0) {
        if (lane == 0) {
            uint32_t st = 0, ph = 1;
            const uint32_t f0 = tc::smem_u32(full), e0 = tc::smem_u32(empty);
            const long long t0 = clock64();
            for (int i = 0; i < n; ++i) {
                tc::mbar_wait_u32(e0 + st * 8, ph);
                tc::mbar_arrive_u32(f0 + st * 8);
                if (++st == (uint32_t)stages) { st = 0; ph ^= 1; }
            }
            out[2] = clock64() - t0;
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = tc::make_idesc_bf16(128, N, 0, 0);
            const uint64_t ad = tc::make_sdesc(tc::smem_u32(smem), 16, 1024), bd = tc::make_sdesc(tc::smem_u32(smem) + 65536, 16, 1024);
            uint32_t st = 0, ph = 0;
            const uint32_t f0 = tc::smem_u32(full), e0 = tc::smem_u32(empty);
            const long long t0 = clock64();
            for (int i = 0; i < n; ++i) {
                tc::mbar_wait_u32(f0 + st * 8, ph);
                tc::tc_fence_after();
                for (int k = 0; k < mmas; ++k) tc::umma_f16(tmem, ad + (uint64_t)(2 * (k & 3)), bd + (uint64_t)(2 * (k & 3)), idesc, 1);
                tc::umma_commit_u32(e0 + st * 8);
                if (++st == (uint32_t)stages) { st = 0; ph ^= 1; }
            }
            tc::umma_commit(&done);
            tc::mbar_wait(&done, 0);
            out[0] = clock64() - t0;
            tc::mbar_arrive(&other);
        }
    } else {
        if (bystanders == 1) tc::mbar_wait_warp(&other, 0);
        else if (bystanders == 2) tc::mbar_wait(&other, 0);
    }
    tc::tc_fence_before(); __syncthreads();
    if (warp == 1) { tc::tc_fence_after(); tc::tmem_dealloc(tmem, 512); }
}

// MMA stream from one thread with a tcgen05.commit (to a barrier nobody waits on) every `every` groups of `mmas`
__global__ void __launch_bounds__(128, 1) k_commit(int n, int mmas, int every, int N, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t dummy, done;
    __shared__ uint32_t tslot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { tc::mbar_init(&dummy, 1 << 20); tc::mbar_init(&done, 1); tc::fence_barrier_init(); }
    if (warp == 1) tc::tmem_alloc(&tslot, 512);
    tc::tc_fence_before(); __syncthreads(); tc::tc_fence_after();
    const uint32_t tmem = tslot;
    if (warp == 0 && lane == 0) {
        const uint32_t idesc = tc::make_idesc_bf16(128, N, 0, 0);
        const uint64_t ad = tc::make_sdesc(tc::smem_u32(smem), 16, 1024), bd = tc::make_sdesc(tc::smem_u32(smem) + 65536, 16, 1024);
        const uint32_t d0 = tc::smem_u32(&dummy);
        const long long t0 = clock64();
        for (int i = 0; i < n; ++i) {
            for (int k = 0; k < mmas; ++k) tc::umma_f16(tmem, ad + (uint64_t)(2 * (k & 3)), bd + (uint64_t)(2 * (k & 3)), idesc, 1);
            if (every && (i % every) == every - 1) tc::umma_commit_u32(d0);
        }
        tc::umma_commit(&done);
        tc::mbar_wait(&done, 0);
        out[0] = clock64() - t0;
    }
    tc::tc_fence_before(); __syncthreads();
    if (warp == 1) { tc::tc_fence_after(); tc::tmem_dealloc(tmem, 512); }
}

// L2 -> shared-memory ingest rate per SM: TMA (16 KB swizzled boxes, `depth` in flight) and/or cp.async from 8 warps
__global__ void __launch_bounds__(320, 1) k_ingest(const __grid_constant__ CUtensorMap tm, const uint4* src, int rows_total, int n_boxes,
                                                   int depth, int use_tma, int use_cpasync, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t full[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) tc::mbar_init(&full[i], 1); tc::fence_barrier_init(); }
    __syncthreads();
    const long long t0 = clock64();
    if (warp == 0 && use_tma) {
        if (lane == 0) {
            // keep `depth` boxes in flight: issue box i after box i-depth has landed
            for (int i = 0; i < n_boxes + depth; ++i) {
                const int st = i % depth;
                if (i >= depth) tc::mbar_wait(&full[st], ((i / depth) - 1) & 1);
                if (i < n_boxes) {
                    tc::mbar_arrive_expect_tx(&full[st], 16384);
                    const int row = ((blockIdx.x * 7 + i) * 128) % (rows_total - 128);
                    tc::tma_load_2d(smem + st * 16384, &tm, &full[st], (i % 16) * 64, row);
                }
            }
        }
    } else if (warp >= 2 && use_cpasync) {
        // 8 warps, each thread 16 B per op; 16 KB "boxes" = 128 rows x 128 B: one warp op covers 4 rows
        const int w = warp - 2;
        uint8_t* dst = smem + 8 * 16384 + w * 8192;
        for (int i = 0; i < n_boxes; ++i) {
            const int row = ((blockIdx.x * 7 + i) * 128) % (rows_total - 128);
            const uint4* base = src + ((long long)(row + w * 16) * 1024 + (i % 16) * 64) * 2 / 16;   // 16 rows per warp per box
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int rr = r * 4 + (lane >> 3);
                fn_cp_async16(dst + ((i & 1) * 4096) + rr * 128 + (lane & 7) * 16, base + (long long)rr * 128 + (lane & 7));
            }
            fn_cp_async_commit();
            fn_cp_async_wait<4>();
        }
        fn_cp_async_wait<0>();
    }
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = clock64() - t0;
}

// 1-D bulk copies (cp.async.bulk, no tensor map): `bytes` contiguous per op, `depth` in flight
__global__ void __launch_bounds__(128, 1) k_bulk(const uint8_t* src, long long src_bytes, int n_ops, int bytes, int depth, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t full[8];
    if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) tc::mbar_init(&full[i], 1); tc::fence_barrier_init(); }
    __syncthreads();
    const long long t0 = clock64();
    if (threadIdx.x == 0) {
        for (int i = 0; i < n_ops + depth; ++i) {
            const int st = i % depth;
            if (i >= depth) tc::mbar_wait(&full[st], ((i / depth) - 1) & 1);
            if (i < n_ops) {
                tc::mbar_arrive_expect_tx(&full[st], bytes);
                const long long off = (((long long)blockIdx.x * 7 + i) * bytes) % (src_bytes - bytes);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(tc::smem_u32(smem + st * 16384)), "l"(src + (off & ~15LL)), "r"(bytes), "r"(tc::smem_u32(&full[st])) : "memory");
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = clock64() - t0;
}

// tensor-TMA ingest with several ISSUING warps (or, lanes_mode, several lanes of warp 0), each its own 2-deep ring of boxes
__global__ void __launch_bounds__(256, 1) k_ingest_mw(const __grid_constant__ CUtensorMap tm, int rows_total, int n_boxes, int nprod, int box_rows,
                                                      int lanes_mode, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t full[8][2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) { tc::mbar_init(&full[i][0], 1); tc::mbar_init(&full[i][1], 1); } tc::fence_barrier_init(); }
    __syncthreads();
    const long long t0 = clock64();
    const int me = lanes_mode ? lane : warp;
    const bool active = lanes_mode ? (warp == 0 && lane < nprod) : (warp < nprod && lane == 0);
    const int box_bytes = box_rows * 128;
    if (active) {
        for (int i = 0; i < n_boxes + 2; ++i) {
            const int st = i & 1;
            if (i >= 2) tc::mbar_wait(&full[me][st], ((i >> 1) - 1) & 1);
            if (i < n_boxes) {
                tc::mbar_arrive_expect_tx(&full[me][st], box_bytes);
                const int row = ((blockIdx.x * 7 + i * nprod + me) * box_rows) % (rows_total - box_rows);
                tc::tma_load_2d(smem + (me * 2 + st) * box_bytes, &tm, &full[me][st], (i % 16) * 64, row);
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = clock64() - t0;
}

int main() {
    {
        long long* o; cudaMallocManaged(&o, 64);
        const int rows_total = 8192;
        __nv_bfloat16* buf; cudaMalloc(&buf, (size_t)rows_total * 1024 * 2); cudaMemset(buf, 0, (size_t)rows_total * 1024 * 2);
        const int smem = 12 * 16384 + 1024;
        cudaFuncSetAttribute(k_ingest_mw, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        for (int box_rows : {32, 128, 256}) {
            CUtensorMap tm;
            fn_make_tmap_bf16_2d(&tm, buf, rows_total, 1024, 1024, box_rows, 64);
            for (int lanes_mode : {0, 1})
                for (int grid : {1, 128})
                    for (int np : {1, 2, 3, 4, 6}) {
                        if (np * 2 * box_rows * 128 > 12 * 16384) continue;
                        k_ingest_mw<<<grid, 256, smem>>>(tm, rows_total, 1000, np, box_rows, lanes_mode, o); cudaDeviceSynchronize();
                        printf("tma-multi: box %3d rows, grid %3d issuing %s %d: %.1f B/clk per SM (%.0f clk per op per issuer) (%s)\n", box_rows, grid,
                               lanes_mode ? "lanes" : "warps", np, 1000.0 * np * box_rows * 128 / (double)o[0], (double)o[0] / 1000.0,
                               cudaGetErrorString(cudaGetLastError()));
                    }
        }
    }
    {
        long long* o; cudaMallocManaged(&o, 64);
        const long long nb = 16ll << 20;
        uint8_t* buf; cudaMalloc(&buf, nb); cudaMemset(buf, 0, nb);
        const int smem = 8 * 16384 + 1024;
        cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        for (int grid : {1, 128})
            for (int bytes : {2048, 8192, 16384})
                for (int depth : {2, 8}) {
                    k_bulk<<<grid, 128, smem>>>(buf, nb, 2000, bytes, depth, o); cudaDeviceSynchronize();
                    printf("bulk1d: grid %3d %5d B per op depth %d: %.1f B/clk per SM (%s)\n", grid, bytes, depth, 2000.0 * bytes / (double)o[0],
                           cudaGetErrorString(cudaGetLastError()));
                }
    }
    {
        long long* o; cudaMallocManaged(&o, 64);
        const int rows_total = 8192;      // 8192 x 1024 bf16 = 16 MB: L2 resident
        __nv_bfloat16* buf; cudaMalloc(&buf, (size_t)rows_total * 1024 * 2); cudaMemset(buf, 0, (size_t)rows_total * 1024 * 2);
        CUtensorMap tm;
        fn_make_tmap_bf16_2d(&tm, buf, rows_total, 1024, 1024, 128, 64);
        const int smem = 8 * 16384 + 8 * 8192 + 1024;
        cudaFuncSetAttribute(k_ingest, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        const int n_boxes = 2000;
        for (int grid : {1, 32, 128, 148})
            for (int mode : {1, 2, 3})
                for (int depth : {2, 4, 8}) {
                    if (mode == 2 && depth != 8) continue;
                    k_ingest<<<grid, 320, smem>>>(tm, (const uint4*)buf, rows_total, n_boxes, depth, mode & 1, (mode >> 1) & 1, o);
                    cudaDeviceSynchronize();
                    const double bytes = (double)n_boxes * 16384 * ((mode & 1) + ((mode >> 1) & 1));
                    printf("ingest: grid %3d  %s depth %2d: %.1f B/clk per SM  (%s)\n", grid, mode == 1 ? "TMA only     " : mode == 2 ? "cp.async only" : "TMA+cp.async ",
                           depth, bytes / (double)o[0], cudaGetErrorString(cudaGetLastError()));
                }
    }
    {
        long long* o; cudaMallocManaged(&o, 64);
        cudaFuncSetAttribute(k_commit, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        for (int N : {48, 96})
            for (int ev : {0, 1, 2, 4})
                for (int mm : {4, 8}) {
                    k_commit<<<1, 128, 200 * 1024>>>(1000, mm, ev, N, o); cudaDeviceSynchronize();
                    printf("commit: N=%d mmas/group=%d commit every %d groups: %.1f cyc/group = %.1f cyc/MMA  %s\n", N, mm, ev, o[0] / 1000.0,
                           o[0] / 1000.0 / mm, cudaGetErrorString(cudaGetLastError()));
                }
    }
    {
        long long* o; cudaMallocManaged(&o, 64);
        cudaFuncSetAttribute(k_ring, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        for (int by : {0, 1, 2})
            for (int mm : {0, 8})
                for (int stg : {1, 4}) {
                    k_ring<<<1, 320, 200 * 1024>>>(2000, stg, mm, 48, by, o); cudaDeviceSynchronize();
                    printf("ring: bystanders=%d (0 none,1 lane0-poll,2 all-lane poll) mmas/stage=%d stages=%d: %.1f cyc/iter (producer loop %.1f)  %s\n", by, mm, stg,
                           o[0] / 2000.0, o[2] / 2000.0, cudaGetErrorString(cudaGetLastError()));
                }
    }
    long long* out; cudaMallocManaged(&out, 64);
    const int smem = 200 * 1024;
    cudaFuncSetAttribute(k_mma<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k_mma<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int n_mma = 256;
    for (int grid : {1, 148}) {
        for (int N : {16, 32, 48, 96, 128, 192, 256}) {
            k_mma<0><<<grid, 128, smem>>>(N, n_mma, 5, out, 0); cudaDeviceSynchronize();
            long long a = out[0], ai = out[1];
            k_mma<1><<<grid, 128, smem>>>(N, n_mma, 5, out, 0); cudaDeviceSynchronize();
            printf("grid %3d  M=128 N=%3d K=16: lane0-only %.1f cyc/MMA (issue %.1f) | elect-converged %.1f cyc/MMA (issue %.1f) | ideal %.1f  err=%s\n",
                   grid, N, (double)a / n_mma, (double)ai / n_mma, (double)out[0] / n_mma, (double)out[1] / n_mma, 128.0 * N / 256, cudaGetErrorString(cudaGetLastError()));
        }
    }
    k_bar<<<1, 32>>>(1000, out); cudaDeviceSynchronize();
    printf("mbarrier arrive+try_wait round trip: %.1f cyc; arrive.expect_tx: %.1f cyc\n", out[0] / 1000.0, out[1] / 1000.0);
    return 0;
}
