// Self-test of the tcgen05 building blocks in tc_common.cuh: D[128,N] = A[128,K] * B[N,K]^T on the tensor
// cores (kind::tf32, fp32 accumulation in TMEM), single TF32 pass or the 3xTF32 split
// (A_hi*B_hi + A_lo*B_hi + A_hi*B_lo) that the refinement projection uses to stay at fp32-level accuracy.
#include "das_common.cuh"
#include "tc_common.cuh"

namespace das {

template <int N>
__global__ void __launch_bounds__(128, 1)
tc_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int K, int split) {
    extern __shared__ __align__(1024) unsigned char smem[];
    // carve (1024-B aligned): A_hi 16 KB | A_lo 16 KB | B_hi N*128 | B_lo N*128
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem) + 1023) & ~static_cast<uintptr_t>(1023));
    unsigned char* sA = base;
    unsigned char* sAl = base + 16384;
    unsigned char* sB = base + 32768;
    unsigned char* sBl = sB + ((N * 128 + 1023) & ~1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;

    if (tid == 0) { tc::mbar_init(&bar, 1); tc::mbar_fence_init(); }
    if (warp == 0) tc::tmem_alloc(&tmem_base, 128);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_d = tmem_base;
    const uint32_t tmem_ahi = tmem_base + 32, tmem_alo = tmem_base + 64;   // A through TMEM (split & 2)
    const bool a_tmem = (split & 2) != 0;
    split &= 1;
    constexpr uint32_t idesc = tc::instr_desc_tf32(128, N);
    uint32_t phase = 0;

    for (int kb = 0; kb < K / 32; ++kb) {
        // gather one k-block: 128 rows x 8 chunks (A), N rows x 8 chunks (B); 8 consecutive threads = one 128-B row
        for (int i = tid; i < 128 * 8; i += 128) {
            const int row = i >> 3, ch = i & 7;
            tc::cp_async16(tc::smem_u32(sA) + tc::swz128(row, ch), A + static_cast<size_t>(row) * K + kb * 32 + ch * 4, true);
        }
        for (int i = tid; i < N * 8; i += 128) {
            const int row = i >> 3, ch = i & 7;
            tc::cp_async16(tc::smem_u32(sB) + tc::swz128(row, ch), B + static_cast<size_t>(row) * K + kb * 32 + ch * 4, true);
        }
        tc::cp_async_commit();
        tc::cp_async_wait<0>();
        __syncthreads();
        if (split) {   // lo = a - tf32(a); the position inside the tile does not matter for an elementwise pass
            for (int i = tid; i < 128 * 32; i += 128) {
                const float a = reinterpret_cast<const float*>(sA)[i];
                reinterpret_cast<float*>(sAl)[i] = a - tc::tf32_hi(a);
            }
            for (int i = tid; i < N * 32; i += 128) {
                const float b = reinterpret_cast<const float*>(sB)[i];
                reinterpret_cast<float*>(sBl)[i] = b - tc::tf32_hi(b);
            }
        }
        if (a_tmem) {
            // thread t owns row t: read its 128 B of this k-block from the swizzled tile, write hi/lo to TMEM
            float hi[32], lo[32];
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) {
                const float4 a = *reinterpret_cast<const float4*>(sA + tc::swz128(tid, ch));
                hi[4 * ch] = a.x; hi[4 * ch + 1] = a.y; hi[4 * ch + 2] = a.z; hi[4 * ch + 3] = a.w;
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) lo[i] = hi[i] - tc::tf32_hi(hi[i]);
            const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
            tc::tmem_st32(tmem_ahi + lane_base, hi);
            if (split) tc::tmem_st32(tmem_alo + lane_base, lo);
            tc::tmem_st_wait();
            tc::tc_fence_before();
        }
        tc::fence_proxy_async();
        __syncthreads();
        if (a_tmem) {
            if (warp == 0) {
                tc::tc_fence_after();
                for (int k = 0; k < 4; ++k) {
                    const uint64_t db = tc::smem_desc_sw128(tc::smem_u32(sB) + k * 32);
                    tc::umma_tf32_ts_elect(tmem_d, tmem_ahi + k * 8, db, idesc, (kb | k) != 0);
                    if (split) {
                        const uint64_t dbl = tc::smem_desc_sw128(tc::smem_u32(sBl) + k * 32);
                        tc::umma_tf32_ts_elect(tmem_d, tmem_alo + k * 8, db, idesc, true);
                        tc::umma_tf32_ts_elect(tmem_d, tmem_ahi + k * 8, dbl, idesc, true);
                    }
                }
                tc::umma_commit_elect(&bar);
            }
        } else if (tid == 0) {
            tc::tc_fence_after();
            for (int k = 0; k < 4; ++k) {
                const uint64_t da = tc::smem_desc_sw128(tc::smem_u32(sA) + k * 32);
                const uint64_t db = tc::smem_desc_sw128(tc::smem_u32(sB) + k * 32);
                tc::umma_tf32(tmem_d, da, db, idesc, (kb | k) != 0);
                if (split) {
                    const uint64_t dal = tc::smem_desc_sw128(tc::smem_u32(sAl) + k * 32);
                    const uint64_t dbl = tc::smem_desc_sw128(tc::smem_u32(sBl) + k * 32);
                    tc::umma_tf32(tmem_d, dal, db, idesc, true);
                    tc::umma_tf32(tmem_d, da, dbl, idesc, true);
                }
            }
            tc::umma_commit(&bar);     // implies tcgen05.fence::before_thread_sync
        }
        tc::mbar_wait(&bar, phase);    // smem of this k-block may be overwritten once the MMAs have read it
        phase ^= 1;
        tc::tc_fence_after();
    }
    // epilogue: thread t <-> TMEM lane (row) t
    float v[16];
#pragma unroll
    for (int c0 = 0; c0 < N; c0 += 16) {
        tc::tmem_ld16(tmem_d + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) D[static_cast<size_t>(tid) * N + c0 + i] = v[i];
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, 128);
}

// Micro-benchmark: one thread issues `iters` back-to-back MMAs (M=128, N, K=8) on resident smem tiles and
// waits for their completion; reports SM cycles for (a) issue only and (b) issue + completion.
template <int N>
__global__ void __launch_bounds__(128, 1) tc_mma_bench_kernel(int iters, long long* out, int a_tmem) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem) + 1023) & ~static_cast<uintptr_t>(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<float*>(base)[i] = 1.0f;
    if (tid == 0) { tc::mbar_init(&bar, 1); tc::mbar_fence_init(); }
    if (warp == 0) tc::tmem_alloc(&tmem_base, 256);
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    if (tid == 0) {
        constexpr uint32_t idesc = tc::instr_desc_tf32(128, N);
        const uint32_t a = tc::smem_u32(base), b = tc::smem_u32(base + 16384);
        const long long t0 = clock64();
        if (a_tmem) {
            for (int i = 0; i < iters; ++i) {
                const uint64_t db = tc::smem_desc_sw128(b + (i & 3) * 32);
                const uint32_t acc = i != 0;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_base),
                             "r"(tmem_base + 128 + (i & 3) * 8), "l"(db), "r"(idesc), "r"(acc) : "memory");
            }
        } else {
            for (int i = 0; i < iters; ++i) {
                const uint64_t da = tc::smem_desc_sw128(a + (i & 3) * 32);
                const uint64_t db = tc::smem_desc_sw128(b + (i & 3) * 32);
                tc::umma_tf32(tmem_base, da, db, idesc, i != 0);
            }
        }
        const long long t1 = clock64();
        tc::umma_commit(&bar);
        tc::mbar_wait(&bar, 0);
        const long long t2 = clock64();
        out[0] = t1 - t0;
        out[1] = t2 - t0;
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, 256);
}

}  // namespace das

extern "C" int das_tc_mma_bench(int32_t N, int32_t iters, long long* out_cycles, void* stream) {
    using namespace das;
    DAS_REQUIRE(out_cycles && iters > 0, DAS_ERR_ARG, "das_tc_mma_bench: bad argument");
    const int a_tmem = N >= 1000;      // N + 1000: A operand from TMEM
    if (a_tmem) N -= 1000;
    const size_t smem = 1024 + 16384 + 32768;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
#define DAS_BENCH_CASE(NN)                                                                                              \
    case NN:                                                                                                            \
        DAS_CUDA_CHECK(cudaFuncSetAttribute(tc_mma_bench_kernel<NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem))); \
        tc_mma_bench_kernel<NN><<<1, 128, smem, st>>>(iters, out_cycles, a_tmem);                                              \
        break;
    switch (N) {
        DAS_BENCH_CASE(16)
        DAS_BENCH_CASE(32)
        DAS_BENCH_CASE(64)
        DAS_BENCH_CASE(128)
        DAS_BENCH_CASE(256)
        default: set_error("das_tc_mma_bench: N=%d", N); return DAS_ERR_ARG;
    }
#undef DAS_BENCH_CASE
    DAS_CUDA_CHECK(cudaGetLastError());
    return DAS_OK;
}

extern "C" int das_tc_selftest(const float* A, const float* B, float* D, int32_t N, int32_t K, int32_t split, void* stream) {
    using namespace das;
    DAS_REQUIRE(A && B && D, DAS_ERR_ARG, "das_tc_selftest: null pointer");
    DAS_REQUIRE((N == 16 || N == 32) && K >= 32 && K % 32 == 0, DAS_ERR_ARG, "das_tc_selftest: N=%d K=%d", N, K);
    const size_t smem = 1024 + 32768 + 2 * 4096;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (N == 16) {
        DAS_CUDA_CHECK(cudaFuncSetAttribute(tc_selftest_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        tc_selftest_kernel<16><<<1, 128, smem, st>>>(A, B, D, K, split);
    } else {
        DAS_CUDA_CHECK(cudaFuncSetAttribute(tc_selftest_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        tc_selftest_kernel<32><<<1, 128, smem, st>>>(A, B, D, K, split);
    }
    DAS_CUDA_CHECK(cudaGetLastError());
    return DAS_OK;
}
