// Stage 1+2: fused centre-score scan and exact per-level top-k.
//
// Replaces DASHead._get_poses_single's per-level head (reference das_head.py:708-723): two sigmoids,
// their product, `topk(nms_pre)` -- plus the optional north-star 3x3 max-pool peak mask (SURVEY.md
// 8.0, divergence A).  One CTA per (image, level); the score plane is read with 128-bit coalesced
// loads, the peak mask is evaluated from shared-memory row strips, and selection is exact with ties
// broken towards the lower cell index (composite 64-bit keys: score bits << 32 | ~index).
//
// Selection strategy (K = nms_pre is small against H*W):
//   Q  (no peak mask, K <= 128) bounded fast path: score <= sigmoid(cls), so after a first sweep over the cls
//      plane alone a logit threshold leaves a few dozen cells for which the exact score is computed;
//   otherwise, or if Q's list overflows:
//   A  every thread scans its cells, writes the 32-bit rank keys to an L2-resident scratch plane and
//      keeps its own maximum;
//   B  tau = K-th largest of the 1024 per-thread maxima (31-step bitwise search with
//      __syncthreads_count) -- a lower bound of the true K-th score that only a handful of cells beat;
//   C  cells with key >= tau are compacted into a shared-memory list, which is bitonic-sorted and the
//      first K entries emitted.
//   F  fallback (K > 128, or the list overflows): exact bitwise radix search over the scratch keys.
#include "das_common.cuh"

namespace das {

constexpr int TK_THREADS = 1024;
constexpr int TK_LIST_CAP = 4096;   // >= 2 * DAS_MAX_NMS_PRE
constexpr int TK_FAST_K = 128;
constexpr int TK_TILE_FLOATS = 16384;  // peak-mode row strip (64 KB)

// Element loads of a logit plane in its bound dtype (das_levels.in_dtype): one element, or 4 consecutive ones as one
// 128-bit (fp32) / 64-bit (fp16, bf16) read.  STREAM: bypass L1 (planes that are swept once).
template <typename T> __device__ __forceinline__ float ld1(const T* p, size_t i);
template <> __device__ __forceinline__ float ld1<float>(const float* p, size_t i) { return __ldg(p + i); }
template <> __device__ __forceinline__ float ld1<__half>(const __half* p, size_t i) { return __half2float(__ldg(p + i)); }
template <> __device__ __forceinline__ float ld1<__nv_bfloat16>(const __nv_bfloat16* p, size_t i) { return __bfloat162float(__ldg(p + i)); }

template <typename T, bool STREAM> __device__ __forceinline__ float4 ld4(const T* p, int q);
template <> __device__ __forceinline__ float4 ld4<float, false>(const float* p, int q) { return ldg_f4(p + 4 * q); }
template <> __device__ __forceinline__ float4 ld4<float, true>(const float* p, int q) { return ldg_f4_stream(p + 4 * q); }
__device__ __forceinline__ uint2 ldg_u2(const void* p, int q, bool stream) {
    uint2 r;
    if (stream) asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(reinterpret_cast<const uint2*>(p) + q));
    else r = __ldg(reinterpret_cast<const uint2*>(p) + q);
    return r;
}
template <> __device__ __forceinline__ float4 ld4<__half, false>(const __half* p, int q) {
    const uint2 r = ldg_u2(p, q, false);
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&r.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
    return make_float4(a.x, a.y, b.x, b.y);
}
template <> __device__ __forceinline__ float4 ld4<__half, true>(const __half* p, int q) {
    const uint2 r = ldg_u2(p, q, true);
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&r.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
    return make_float4(a.x, a.y, b.x, b.y);
}
template <> __device__ __forceinline__ float4 ld4<__nv_bfloat16, false>(const __nv_bfloat16* p, int q) {
    const uint2 r = ldg_u2(p, q, false);      // bf16 -> fp32 is a 16-bit shift
    return make_float4(__uint_as_float(r.x << 16), __uint_as_float(r.x & 0xFFFF0000u), __uint_as_float(r.y << 16), __uint_as_float(r.y & 0xFFFF0000u));
}
template <> __device__ __forceinline__ float4 ld4<__nv_bfloat16, true>(const __nv_bfloat16* p, int q) {
    const uint2 r = ldg_u2(p, q, true);
    return make_float4(__uint_as_float(r.x << 16), __uint_as_float(r.x & 0xFFFF0000u), __uint_as_float(r.y << 16), __uint_as_float(r.y & 0xFFFF0000u));
}

__device__ __forceinline__ uint64_t compose(uint32_t key, uint32_t idx) {
    return (static_cast<uint64_t>(key) << 32) | static_cast<uint64_t>(0xFFFFFFFFu - idx);
}

// in-place descending bitonic sort of n2 (power of two) u64 keys in shared memory
__device__ void bitonic_desc(uint64_t* a, int n2) {
    for (int k = 2; k <= n2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (n2 >> 1); t += blockDim.x) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int p = i | j;
                const bool desc = ((i & k) == 0);
                const uint64_t x = a[i], y = a[p];
                if ((x < y) == desc) { a[i] = y; a[p] = x; }
            }
            __syncthreads();
        }
    }
}

template <typename T>
__device__ __forceinline__ void score_topk_body(const das_levels* __restrict__ lvp, int nms_pre, int peak,
                                                float* __restrict__ cand_score, int32_t* __restrict__ cand_index, int cand_slots,
                                                uint32_t* __restrict__ scratch, int scratch_per_image, int32_t* __restrict__ zero_counters) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    pdl_trigger();       // first kernel of the chain: the phase 1-2 kernel may be scheduled now; it waits for this grid to complete
    if (zero_counters && blockIdx.x == 0 && threadIdx.x < 2 + DAS_MAX_JOINTS) {
        // the refinement's work-queue / valid / per-joint row counters of this decode ([2] = peer-store ticket: left alone);
        // every kernel of the previous decode that read them has completed (stream order)
        zero_counters[threadIdx.x < 2 ? threadIdx.x : threadIdx.x + 2] = 0;
    }
    uint64_t* list = reinterpret_cast<uint64_t*>(smem_raw);                       // TK_LIST_CAP
    float* tile = reinterpret_cast<float*>(smem_raw + TK_LIST_CAP * sizeof(uint64_t));  // peak mode only
    __shared__ int red[33];
    __shared__ int list_n;

    const int nl = lvp->n_levels;
    const int b = blockIdx.x / nl, l = blockIdx.x - b * nl;
    const int H = lvp->lv[l].H, W = lvp->lv[l].W;
    const int HW = H * W;
    int slot0 = 0, sc0 = 0;
    for (int i = 0; i < l; ++i) {
        const int hw = lvp->lv[i].H * lvp->lv[i].W;
        slot0 += level_slots(hw, nms_pre);
        sc0 += hw;
    }
    const int K = level_slots(HW, nms_pre);
    const T* __restrict__ cls = reinterpret_cast<const T*>(lvp->lv[l].cls) + static_cast<size_t>(b) * HW;
    const T* __restrict__ ctr = reinterpret_cast<const T*>(lvp->lv[l].ctr) + static_cast<size_t>(b) * HW;
    constexpr uintptr_t VMASK = 4 * sizeof(T) - 1;     // alignment of a 4-element vector load
    float* oscore = cand_score + static_cast<size_t>(b) * cand_slots + slot0;
    int32_t* oidx = cand_index + static_cast<size_t>(b) * cand_slots + slot0;
    const int tid = threadIdx.x;

    if (K == HW) {  // pass-through: raster order, no ranking (das_head.py:717 false branch)
        for (int i = tid; i < HW; i += TK_THREADS) {
            oscore[i] = sigmoid_acc(ld1<T>(cls, i)) * sigmoid_acc(ld1<T>(ctr, i));
            oidx[i] = i;
        }
        return;
    }
    uint32_t* __restrict__ keys = scratch + static_cast<size_t>(b) * scratch_per_image + sc0;

    if (tid == 0) list_n = 0;
    __syncthreads();
    bool done = false;
    if (!peak && K <= TK_FAST_K) {
        // ---- Q: bounded fast path ---------------------------------------------------------------------------
        // score = sigmoid(cls) * sigmoid(ctr) <= sigmoid(cls), so once a lower bound tau of the K-th best score is
        // known, only cells with cls >= logit(tau) can matter: the exp/div work and the centerness reads shrink from
        // every cell to a few dozen.  tau = K-th largest of the exact scores of each thread's best-cls cell (1024
        // distinct cells => a valid lower bound).
        const bool vec = ((HW & 3) == 0) && ((reinterpret_cast<uintptr_t>(cls) & VMASK) == 0);
        float bc = -INFINITY;
        int bi = -1;
        if (vec) {
            const int n4 = HW >> 2;
            for (int q = tid; q < n4; q += TK_THREADS) {
                const float4 a = ld4<T, false>(cls, q);        // stays in L1 for the second sweep
                if (a.x > bc) { bc = a.x; bi = 4 * q; }
                if (a.y > bc) { bc = a.y; bi = 4 * q + 1; }
                if (a.z > bc) { bc = a.z; bi = 4 * q + 2; }
                if (a.w > bc) { bc = a.w; bi = 4 * q + 3; }
            }
        } else {
            for (int i = tid; i < HW; i += TK_THREADS) {
                const float a = ld1<T>(cls, i);
                if (a > bc) { bc = a; bi = i; }
            }
        }
        uint32_t tau = 0;
        if (K <= 32) {
            // tau = K-th largest of the 32 exact scores of each warp's best-cls cell: 32 distinct cells, so still a
            // valid lower bound, and it needs one shuffle sort in warp 0 instead of 31 block-wide barriers.
            const int lane = tid & 31, warp = tid >> 5;
            float wb = bc;
            int wi = bi;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ob = __shfl_xor_sync(0xffffffffu, wb, o);
                const int oi = __shfl_xor_sync(0xffffffffu, wi, o);
                if (ob > wb || (ob == wb && oi >= 0 && (wi < 0 || oi < wi))) { wb = ob; wi = oi; }
            }
            if (lane == 0) red[warp] = (wi >= 0) ? static_cast<int>(__float_as_uint(sigmoid_acc(wb) * sigmoid_acc(ld1<T>(ctr, wi)))) : 0;
            __syncthreads();
            if (warp == 0) {
                uint32_t v = static_cast<uint32_t>(red[lane]);
                // bitonic sort of 32 keys across the lanes, descending
#pragma unroll
                for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
                    for (int j = k >> 1; j > 0; j >>= 1) {
                        const uint32_t o = __shfl_xor_sync(0xffffffffu, v, j);
                        const bool up = ((lane & k) == 0) == ((lane & j) == 0);   // keep the larger one here?
                        v = up ? max(v, o) : min(v, o);
                    }
                }
                if (lane == K - 1) red[32] = static_cast<int>(v);
            }
            __syncthreads();
            tau = static_cast<uint32_t>(red[32]);
        } else {
            uint32_t mykey = 0;
            if (bi >= 0) mykey = __float_as_uint(sigmoid_acc(bc) * sigmoid_acc(ld1<T>(ctr, bi)));
            for (int bit = 30; bit >= 0; --bit) {
                const uint32_t trial = tau | (1u << bit);
                if (__syncthreads_count(mykey >= trial) >= K) tau = trial;
            }
        }
        if (tau > 0) {
            const float tf = fminf(__uint_as_float(tau), 0.999999f);
            const float a_thr = fminf(logf(tf / (1.0f - tf)) - 0.01f, 13.0f);   // conservative logit(tau)
            auto consider = [&](float a, int i) {
                if (a >= a_thr) {
                    const uint32_t k = __float_as_uint(sigmoid_acc(a) * sigmoid_acc(ld1<T>(ctr, i)));
                    if (k >= tau) {
                        const int pos = atomicAdd(&list_n, 1);
                        if (pos < TK_LIST_CAP) list[pos] = compose(k, i);
                    }
                }
            };
            if (vec) {
                const int n4 = HW >> 2;
                for (int q = tid; q < n4; q += TK_THREADS) {
                    const float4 a = ld4<T, false>(cls, q);
                    consider(a.x, 4 * q); consider(a.y, 4 * q + 1); consider(a.z, 4 * q + 2); consider(a.w, 4 * q + 3);
                }
            } else {
                for (int i = tid; i < HW; i += TK_THREADS) consider(ld1<T>(cls, i), i);
            }
            __syncthreads();
            done = (list_n <= TK_LIST_CAP && list_n >= K);   // block-uniform
        }
    }
    if (!done) {
        __syncthreads();
        // ---- A: rank keys -> scratch, per-thread maximum ---------------------------------------------
        uint64_t best = 0;
        if (!peak) {
            const bool vec = ((HW & 3) == 0) && ((reinterpret_cast<uintptr_t>(cls) & VMASK) == 0) &&
                             ((reinterpret_cast<uintptr_t>(ctr) & VMASK) == 0) &&
                             ((reinterpret_cast<uintptr_t>(keys) & 15) == 0);
            if (vec) {
                const int n4 = HW >> 2;
                for (int q = tid; q < n4; q += TK_THREADS) {
                    const float4 a = ld4<T, true>(cls, q);
                    const float4 c = ld4<T, true>(ctr, q);
                    uint4 k;
                    k.x = __float_as_uint(sigmoid_acc(a.x) * sigmoid_acc(c.x));
                    k.y = __float_as_uint(sigmoid_acc(a.y) * sigmoid_acc(c.y));
                    k.z = __float_as_uint(sigmoid_acc(a.z) * sigmoid_acc(c.z));
                    k.w = __float_as_uint(sigmoid_acc(a.w) * sigmoid_acc(c.w));
                    reinterpret_cast<uint4*>(keys)[q] = k;
                    const uint32_t i0 = 4u * q;
                    uint64_t m = compose(k.x, i0);
                    uint64_t t1 = compose(k.y, i0 + 1); m = t1 > m ? t1 : m;
                    t1 = compose(k.z, i0 + 2); m = t1 > m ? t1 : m;
                    t1 = compose(k.w, i0 + 3); m = t1 > m ? t1 : m;
                    best = m > best ? m : best;
                }
            } else {
                for (int i = tid; i < HW; i += TK_THREADS) {
                    const uint32_t k = __float_as_uint(sigmoid_acc(ld1<T>(cls, i)) * sigmoid_acc(ld1<T>(ctr, i)));
                    keys[i] = k;
                    const uint64_t c = compose(k, i);
                    best = c > best ? c : best;
                }
            }
        } else {
            // 3x3 peak mask from shared-memory row strips: rows [r0-1, r0+R] staged, rows [r0, r0+R) ranked
            const int R = max(1, min(H, TK_TILE_FLOATS / W - 2));
            for (int r0 = 0; r0 < H; r0 += R) {
                const int rows = min(R, H - r0);
                const int n_stage = (rows + 2) * W;
                for (int e = tid; e < n_stage; e += TK_THREADS) {
                    const int ry = e / W, x = e - ry * W;
                    const int y = r0 - 1 + ry;
                    float s = 0.0f;  // scores are >= 0, so 0 stands in for "outside the map" (max-pool pads -inf)
                    if (y >= 0 && y < H) s = sigmoid_acc(ld1<T>(cls, y * W + x)) * sigmoid_acc(ld1<T>(ctr, y * W + x));
                    tile[e] = s;
                }
                __syncthreads();
                for (int e = tid; e < rows * W; e += TK_THREADS) {
                    const int ry = e / W, x = e - ry * W;
                    const float* c = tile + (ry + 1) * W + x;
                    const float s = c[0];
                    float m = fmaxf(c[-W], c[W]);
                    if (x > 0) m = fmaxf(m, fmaxf(c[-1], fmaxf(c[-W - 1], c[W - 1])));
                    if (x < W - 1) m = fmaxf(m, fmaxf(c[1], fmaxf(c[-W + 1], c[W + 1])));
                    const uint32_t k = (s >= m) ? __float_as_uint(s) : 0u;
                    const int i = (r0 + ry) * W + x;
                    keys[i] = k;
                    const uint64_t cc = compose(k, i);
                    best = cc > best ? cc : best;
                }
                __syncthreads();
            }
        }
        if (tid == 0) list_n = 0;
        __syncthreads();  // also makes this block's scratch writes visible to the whole block

        bool need_fallback = (K > TK_FAST_K);
        if (!need_fallback) {
            // ---- B: tau = K-th largest per-thread maximum (score bits only) ---------------------------
            const uint32_t mykey = static_cast<uint32_t>(best >> 32);
            uint32_t tau = 0;
            for (int bit = 30; bit >= 0; --bit) {
                const uint32_t trial = tau | (1u << bit);
                if (__syncthreads_count(mykey >= trial) >= K) tau = trial;
            }
            // ---- C: compaction of cells with key >= tau ------------------------------------------------
            for (int i = tid; i < HW; i += TK_THREADS) {
                const uint32_t k = keys[i];
                if (k >= tau) {
                    const int pos = atomicAdd(&list_n, 1);
                    if (pos < TK_LIST_CAP) list[pos] = compose(k, i);
                }
            }
            __syncthreads();
            if (list_n > TK_LIST_CAP || list_n < K) need_fallback = true;  // block-uniform
        }

        if (need_fallback) {
            // ---- F: exact bitwise search over all keys -------------------------------------------------
            __syncthreads();
            if (tid == 0) list_n = 0;
            uint32_t T = 0;
            for (int bit = 30; bit >= 0; --bit) {
                const uint32_t trial = T | (1u << bit);
                int c = 0;
                for (int i = tid; i < HW; i += TK_THREADS) c += (keys[i] >= trial);
                if (block_sum_1024(c, red) >= K) T = trial;
            }
            int cg = 0;
            for (int i = tid; i < HW; i += TK_THREADS) cg += (keys[i] > T);
            cg = block_sum_1024(cg, red);
            const int r = K - cg;  // how many cells with key == T are taken, lowest indices first (r >= 1)
            // smallest I with count(key == T && idx <= I) >= r  <=>  largest prefix P with count(idx < P) < r
            uint32_t P = 0;
            for (int bit = 30; bit >= 0; --bit) {
                const uint32_t trial = P | (1u << bit);
                int c = 0;
                for (int i = tid; i < HW; i += TK_THREADS) c += (keys[i] == T && static_cast<uint32_t>(i) < trial);
                if (block_sum_1024(c, red) < r) P = trial;
            }
            __syncthreads();
            for (int i = tid; i < HW; i += TK_THREADS) {
                const uint32_t k = keys[i];
                if (k > T || (k == T && static_cast<uint32_t>(i) <= P)) {
                    const int pos = atomicAdd(&list_n, 1);
                    if (pos < TK_LIST_CAP) list[pos] = compose(k, i);
                }
            }
            __syncthreads();
        }

    }

    // ---- D: order the short list, emit the first K -------------------------------------------------
    const int n = min(list_n, TK_LIST_CAP);
    if (n <= 256) {
        // rank by counting (keys are distinct): no sort, no further barriers
        for (int i = tid; i < n; i += TK_THREADS) {
            const uint64_t c = list[i];
            int rnk = 0;
            for (int j = 0; j < n; ++j) rnk += (list[j] > c);
            if (rnk < K) {
                const uint32_t idx = 0xFFFFFFFFu - static_cast<uint32_t>(c & 0xFFFFFFFFull);
                float sc = __uint_as_float(static_cast<uint32_t>(c >> 32));
                if (peak) sc = sigmoid_acc(ld1<T>(cls, idx)) * sigmoid_acc(ld1<T>(ctr, idx));
                oscore[rnk] = sc;
                oidx[rnk] = static_cast<int32_t>(idx);
            }
        }
        return;
    }
    int n2 = 1;
    while (n2 < n) n2 <<= 1;
    for (int i = n + tid; i < n2; i += TK_THREADS) list[i] = 0;
    __syncthreads();
    bitonic_desc(list, n2);
    for (int rnk = tid; rnk < K; rnk += TK_THREADS) {
        const uint64_t c = list[rnk];
        const uint32_t idx = 0xFFFFFFFFu - static_cast<uint32_t>(c & 0xFFFFFFFFull);
        float s = __uint_as_float(static_cast<uint32_t>(c >> 32));
        if (peak) s = sigmoid_acc(ld1<T>(cls, idx)) * sigmoid_acc(ld1<T>(ctr, idx));  // masked cells rank as 0
        oscore[rnk] = s;
        oidx[rnk] = static_cast<int32_t>(idx);
    }
}

// One instantiation per element type of the logit planes (das_levels.in_dtype); the launcher picks it from the HOST copy
// of the level table, so a plan whose inputs change type re-captures its graph (das_plan_bind).
template <typename T>
__global__ void __launch_bounds__(TK_THREADS, 1)
score_topk_kernel(const das_levels* __restrict__ lvp, int nms_pre, int peak,
                  float* __restrict__ cand_score, int32_t* __restrict__ cand_index, int cand_slots,
                  uint32_t* __restrict__ scratch, int scratch_per_image, int32_t* __restrict__ zero_counters) {
    score_topk_body<T>(lvp, nms_pre, peak, cand_score, cand_index, cand_slots, scratch, scratch_per_image, zero_counters);
}

}  // namespace das

extern "C" int das_score_topk(const das_levels* d_levels, const das_levels* h_levels, int32_t nms_pre,
                              int32_t peak_kernel, float* cand_score, int32_t* cand_index,
                              int32_t cand_slots, uint32_t* scratch, void* stream) {
    using namespace das;
    DAS_REQUIRE(d_levels && h_levels && cand_score && cand_index && scratch, DAS_ERR_ARG, "das_score_topk: null pointer");
    DAS_REQUIRE(h_levels->n_levels >= 1 && h_levels->n_levels <= DAS_MAX_LEVELS && h_levels->batch >= 1, DAS_ERR_ARG,
                "das_score_topk: n_levels=%d batch=%d out of range", h_levels->n_levels, h_levels->batch);
    DAS_REQUIRE(nms_pre <= DAS_MAX_NMS_PRE, DAS_ERR_CAPACITY, "nms_pre=%d exceeds capacity %d", nms_pre, DAS_MAX_NMS_PRE);
    DAS_REQUIRE(peak_kernel == 0 || peak_kernel == 1 || peak_kernel == 3, DAS_ERR_UNSUPPORTED,
                "peak_kernel=%d: only 0/1 (off) and 3 are built", peak_kernel);
    const int peak = (peak_kernel == 3);
    int total = 0, per_image = 0;
    for (int l = 0; l < h_levels->n_levels; ++l) {
        const int hw = h_levels->lv[l].H * h_levels->lv[l].W;
        DAS_REQUIRE(hw > 0, DAS_ERR_ARG, "level %d has empty map", l);
        DAS_REQUIRE(!peak || (h_levels->lv[l].W + 0) * 3 <= TK_TILE_FLOATS, DAS_ERR_CAPACITY, "map too wide for the peak tile");
        total += level_slots(hw, nms_pre);
        per_image += hw;
    }
    DAS_REQUIRE(total == cand_slots, DAS_ERR_ARG, "cand_slots=%d but levels give %d", cand_slots, total);
    const size_t smem = TK_LIST_CAP * sizeof(uint64_t) + (peak ? TK_TILE_FLOATS * sizeof(float) : 0);
    static DeviceOnce attr_done;
    if (attr_done.need()) {
        const int max_smem = static_cast<int>(TK_LIST_CAP * sizeof(uint64_t) + TK_TILE_FLOATS * sizeof(float));
        DAS_CUDA_CHECK(cudaFuncSetAttribute(score_topk_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        DAS_CUDA_CHECK(cudaFuncSetAttribute(score_topk_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        DAS_CUDA_CHECK(cudaFuncSetAttribute(score_topk_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    }
    const int grid = h_levels->batch * h_levels->n_levels;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int32_t* zc = chain_ctx().zero_counters;     // inside das_plan's chain: clear the refinement counters here (no memset nodes)
    switch (h_levels->in_dtype) {
        case DAS_DTYPE_F32:
            score_topk_kernel<float><<<grid, TK_THREADS, smem, st>>>(d_levels, nms_pre, peak, cand_score, cand_index, cand_slots, scratch, per_image, zc);
            break;
        case DAS_DTYPE_F16:
            score_topk_kernel<__half><<<grid, TK_THREADS, smem, st>>>(d_levels, nms_pre, peak, cand_score, cand_index, cand_slots, scratch, per_image, zc);
            break;
        case DAS_DTYPE_BF16:
            score_topk_kernel<__nv_bfloat16><<<grid, TK_THREADS, smem, st>>>(d_levels, nms_pre, peak, cand_score, cand_index, cand_slots, scratch, per_image, zc);
            break;
        default:
            set_error("das_score_topk: in_dtype=%d", h_levels->in_dtype);
            return DAS_ERR_ARG;
    }
    DAS_CUDA_CHECK(cudaGetLastError());
    return DAS_OK;
}
