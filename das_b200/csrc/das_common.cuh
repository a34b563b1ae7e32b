// Shared helpers for the DAS decode kernels (sm_100a).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "das_decode.h"

namespace das {

constexpr int kSMs = 148;  // B200: 2 dies x 74 SMs

// cudaFuncSetAttribute is per device: each launcher remembers which devices it has configured, so a
// process that drives several GPUs (one plan per device) sets the attribute on each of them.
struct DeviceOnce {
    unsigned long long done[4] = {0, 0, 0, 0};
    bool need() {
        int d = 0;
        if (cudaGetDevice(&d) != cudaSuccess) return true;
        d &= 255;
        const unsigned long long bit = 1ull << (d & 63);
        const bool first = !(done[d >> 6] & bit);
        done[d >> 6] |= bit;
        return first;
    }
};

void set_error(const char* fmt, ...);

// ---- programmatic dependent launch (PDL) -------------------------------------------------------------------------------
// The kernels of one decode form a chain on one stream.  Launched with the programmatic-stream-serialization attribute,
// kernel n+1 may be scheduled (and run the part of its prologue that touches no predecessor data: barrier / TMEM set-up,
// weight-panel loads) as soon as every CTA of kernel n has called pdl_trigger(); pdl_wait() then blocks until kernel n has
// COMPLETED and its writes are visible -- so correctness only needs pdl_wait() before the first dependent access.  Both
// are no-ops in a kernel launched without the attribute (the stand-alone stage launchers of the C ABI).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// What das_plan's enqueue tells the stage launchers it calls (thread-local; default = stand-alone behaviour).
struct ChainCtx {
    bool pdl = false;                 // launch with the PDL attribute (the previous node on the stream is a kernel of this chain)
    int32_t* zero_counters = nullptr; // das_score_topk: refine counters to clear ([0], [1], [4 .. 4+DAS_MAX_JOINTS)) for the chain
    bool counters_cleared = false;    // das_refine_heads: the counters were cleared by das_score_topk of this chain -> no memset nodes
    int extra_launches = 0;           // kernels a stage launcher enqueued beyond its usual one (das_plan's launch count)
};
ChainCtx& chain_ctx();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_chain(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

#define DAS_CUDA_CHECK(expr)                                                              \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            ::das::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return DAS_ERR_CUDA;                                                          \
        }                                                                                 \
    } while (0)

#define DAS_REQUIRE(cond, code, ...)      \
    do {                                  \
        if (!(cond)) {                    \
            ::das::set_error(__VA_ARGS__); \
            return (code);                \
        }                                 \
    } while (0)

#define DAS_TRY(expr)                 \
    do {                              \
        int _s = (expr);              \
        if (_s != DAS_OK) return _s;  \
    } while (0)

__host__ __device__ __forceinline__ int level_slots(int hw, int nms_pre) {
    // das_head.py:716-717: top-k only `if nms_pre > 0 and N > nms_pre`, else every cell passes through
    return (nms_pre > 0 && hw > nms_pre) ? nms_pre : hw;
}

// Accurate (not __expf) so the score stays within ~2 ulp of a correctly rounded sigmoid; the CPU
// reference's own sigmoid is only that accurate (hard parts).
__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ int block_sum_1024(int v, int* red /* >= 33 ints smem */) {
    v = __reduce_add_sync(0xffffffffu, v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    int nw = (blockDim.x + 31) >> 5;
    int t = (lane < nw) ? red[lane] : 0;
    t = __reduce_add_sync(0xffffffffu, t);
    return t;
}

// A head-output map (cls / ctr / pose) in the element type the caller bound (das_levels.in_dtype): map(i) = element i as
// fp32.  The branch is uniform; fp16 / bf16 values convert exactly, so every result equals the up-cast path bit for bit.
struct InMap {
    const void* p;
    int dt;
    __device__ __forceinline__ InMap(const float* base, int dtype) : p(base), dt(dtype) {}
    __device__ __forceinline__ float operator()(size_t i) const {
        if (dt == DAS_DTYPE_F16) return __half2float(__ldg(reinterpret_cast<const __half*>(p) + i));
        if (dt == DAS_DTYPE_BF16) return __bfloat162float(__ldg(reinterpret_cast<const __nv_bfloat16*>(p) + i));
        return __ldg(reinterpret_cast<const float*>(p) + i);
    }
};
__host__ __device__ __forceinline__ int dtype_bytes(int dt) { return dt == DAS_DTYPE_F32 ? 4 : 2; }

__device__ __forceinline__ float4 ldg_f4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// One 256-bit read-only load (LDG.E.256, sm_100+): a 32-byte record in a single request instead of two 128-bit ones.
struct F8 { float4 lo, hi; };
__device__ __forceinline__ F8 ldg_f8(const float4* p32) {
    F8 r;
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.lo.x), "=f"(r.lo.y), "=f"(r.lo.z), "=f"(r.lo.w), "=f"(r.hi.x), "=f"(r.hi.y), "=f"(r.hi.z), "=f"(r.hi.w)
                 : "l"(p32));
    return r;
}

// streaming 128-bit load that does not pollute L1 (score planes are read once per pass)
__device__ __forceinline__ float4 ldg_f4_stream(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}

}  // namespace das
