// Stage 5: score threshold, OKS-NMS, nms_post, output packing, depth de-normalisation and
// back-projection to camera / world space.
//
// Reference semantics (file:line relative to the reference root):
//   * concat / score_thr / areas / keep[:nms_post] ... das_head.py:751-794
//   * oks_iou / oks_nms ............................... mmdet3d/core/post_processing/pose_nms.py:51-126
//       float32 keypoints and areas, float64 exponent math, float32 OKS compared with `<= thr`;
//       sigmas = COCO-17 table / 10 iff J == 17 else 0.08
//   * depth de-normalisation .......................... mmdet3d/datasets/cmupanoptic_mono_dataset.py:391-401
//   * pixel2world ...................................... mytools/vis_3d.py:16-26 (float64)
// Ordering rule of this implementation: candidates are ranked by score, equal scores by the lower
// candidate slot (the reference's `argsort()[::-1]` has no defined tie order).
//
// One CTA per image.  The reference does this on the host with 3 device->host copies per candidate
// and an O(N^2) python loop; here the candidates never leave the GPU.
#include "das_common.cuh"

namespace das {

constexpr int NM_THREADS = 1024;
constexpr int NM_MAX_CAND = 8192;
constexpr int NM_MATRIX_N = 64;   // up to this many candidates: all-pairs OKS + bitmask greedy

struct NmsParams {
    int B, CT, P, J, root, nms_post, soft;
    float nms_thr, score_thr;
    double ddf;
    const float* cand_score;
    const float* cand_pose;
    const float* cand_center;
    const double* cam;
    das_buffers out;
    das_peer_blocks peers;     // n == 0: local stores only
};

// Fused result all-gather: once an image's part of the packed output block is complete, its CTA copies it to the same
// offsets of every peer's copy of this rank's block -- NVLink P2P stores to IPC-mapped memory, 16 bytes wide where the
// chunk allows it, fire-and-forget ("only the final small pose lists are gathered over NVLink").
template <int NT>
__device__ __forceinline__ void copy_chunk_to_peers(const das_peer_blocks& pe, const void* local, int words) {
    const uint32_t* src = static_cast<const uint32_t*>(local);
    const int head = min(words, static_cast<int>(((16u - (reinterpret_cast<uintptr_t>(src) & 15u)) & 15u) >> 2));
    const int nvec = (words - head) >> 2;
    const int tail0 = head + (nvec << 2);
    // every value is loaded once and stored to all peers back to back (independent posted writes over NVLink)
    auto peer = [&](int q) { return reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(const_cast<uint32_t*>(src)) + pe.delta[q]); };
    for (int i = threadIdx.x; i < nvec; i += NT) {
        const uint4 v = __ldcg(reinterpret_cast<const uint4*>(src + head) + i);
        for (int q = 0; q < pe.n; ++q) reinterpret_cast<uint4*>(peer(q) + head)[i] = v;
    }
    for (int i = threadIdx.x; i < head; i += NT) {
        const uint32_t v = __ldcg(src + i);
        for (int q = 0; q < pe.n; ++q) peer(q)[i] = v;
    }
    for (int i = tail0 + threadIdx.x; i < words; i += NT) {
        const uint32_t v = __ldcg(src + i);
        for (int q = 0; q < pe.n; ++q) peer(q)[i] = v;
    }
}

__device__ __forceinline__ double oks_var(int j, int J) {
    // pose_nms.py:65-73: vars = (sigmas * 2) ** 2
    const double tbl[17] = {.26, .25, .25, .35, .35, .79, .79, .72, .72, .62, .62, 1.07, 1.07, .87, .87, .89, .89};
    const double s = (J == 17) ? tbl[j] / 10.0 : 0.08;
    return (s * 2.0) * (s * 2.0);
}

// One joint's term exp(-e_j) of OKS(g, d); pose rows are [J,3] float32.
__device__ __forceinline__ double oks_term(const float* __restrict__ pg, const float* __restrict__ pd,
                                           float ag, float ad, int j, int J) {
    const float dx = __fsub_rn(pd[3 * j], pg[3 * j]);
    const float dy = __fsub_rn(pd[3 * j + 1], pg[3 * j + 1]);
    const float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
    const float am = __fmul_rn(__fadd_rn(ag, ad), 0.5f);            // (a_g + a_d) / 2 in float32
    const double e = static_cast<double>(d2) / oks_var(j, J) / (static_cast<double>(am) + 2.220446049250313e-16) / 2.0;
    return exp(-e);
}

// group of G lanes (16 or 32) cooperates on one pair, one joint per lane; every lane of the group returns the float32 OKS.
// The J float64 terms are added in the order NumPy's np.sum adds them (pose_nms.py:91; pairwise_sum for fewer than 128
// elements): eight running sums over the first 8 * (J / 8) terms -- lane k < 8 adds terms k, k + 8, k + 16, .. -- combined as
// the balanced tree ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7)) (three xor steps inside the first 8 lanes; a + b == b + a
// exactly), then the remaining terms one by one; fewer than 8 terms: left to right.  Must be called by all 32 lanes.
template <int G>
__device__ __forceinline__ float oks_pair(const float* pg, const float* pd, float ag, float ad, int J, int gl) {
    const double t = (gl < J) ? oks_term(pg, pd, ag, ad, gl, J) : 0.0;
    double res;
    if (J < 8) {
        res = 0.0;
        for (int k = 0; k < J; ++k) res += __shfl_sync(0xffffffffu, t, k, G);
    } else {
        const int m = J - (J % 8);
        double r = t;
        for (int i = 8; i < m; i += 8) r += __shfl_down_sync(0xffffffffu, t, i, G);
        r += __shfl_xor_sync(0xffffffffu, r, 1, G);
        r += __shfl_xor_sync(0xffffffffu, r, 2, G);
        r += __shfl_xor_sync(0xffffffffu, r, 4, G);
        res = r;
        for (int i = m; i < J; ++i) res += __shfl_sync(0xffffffffu, t, i, G);
        res = __shfl_sync(0xffffffffu, res, 0, G);
    }
    return static_cast<float>(res / static_cast<double>(J));
}

// The same value computed by ONE lane, the J terms added in the order NumPy's np.sum adds a contiguous float64 array of
// J < 128 elements (pairwise_sum: eight running sums over the first 8 * (J / 8) terms, combined as a balanced tree, the
// remaining terms added one by one; fewer than 8 terms: left to right) -- pose_nms.py:91.
__device__ __forceinline__ float oks_pair_serial(const float* pg, const float* pd, float ag, float ad, int J) {
    double res;
    if (J < 8) {
        res = 0.0;
        for (int k = 0; k < J; ++k) res += oks_term(pg, pd, ag, ad, k, J);
    } else {
        double r[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) r[k] = oks_term(pg, pd, ag, ad, k, J);
        const int m = J - (J % 8);
        for (int i = 8; i < m; i += 8) {
#pragma unroll
            for (int k = 0; k < 8; ++k) r[k] += oks_term(pg, pd, ag, ad, i + k, J);
        }
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (int i = m; i < J; ++i) res += oks_term(pg, pd, ag, ad, i, J);
    }
    return static_cast<float>(res / static_cast<double>(J));
}

// Pair index pr in [0, n(n-1)/2) -> (i < j), rows enumerated as (0,1) (0,2) ... (0,n-1) (1,2) ...; closed form with a
// one-step fix-up instead of walking the rows (the walk was 27 % of the kernel's stall samples at n = 64).
__device__ __forceinline__ void unrank_pair(int pr, int n, int& i, int& j) {
    const float fn = static_cast<float>(2 * n - 1);
    int r = static_cast<int>((fn - sqrtf(fmaxf(fn * fn - 8.0f * static_cast<float>(pr), 0.f))) * 0.5f);
    r = max(0, min(r, n - 2));
    auto start = [n](int k) { return k * (2 * n - k - 1) / 2; };
    while (r + 1 <= n - 2 && start(r + 1) <= pr) ++r;
    while (r > 0 && start(r) > pr) --r;
    i = r;
    j = r + 1 + (pr - start(r));
}

// Hard NMS only needs the DECISION oks > thr.  A float32 estimate of the same expression settles every pair whose estimate
// is further than OKS_BAND from thr -- on either side: pairs of distinct people (estimate ~0) and near-duplicates of one person
// (estimate ~1) -- and only a pair inside the band pays for the reference's float64 divide / exp chain, so the decision is
// still the exact one.
// Error of the estimate: e differs from the float64 value by a relative 4e-7 (float32 variance, two float32 divisions), a
// term exp(-e) therefore by at most max(e exp(-e)) * 4e-7 = 1.5e-7, ex2.approx adds 2e-8: below 1e-6 for the mean -- the band
// is 1000 times that.  NaN-safe: anything that is not clearly on one side takes the exact path.  Must be called by all 32
// lanes; the exact path is taken warp-wide when any of the warp's groups needs it (the float64 butterfly uses full-warp
// shuffles).
constexpr float OKS_BAND = 1e-3f;
template <int G>
__device__ __forceinline__ bool oks_over(const float* pg, const float* pd, float ag, float ad, int J, int gl, float thr) {
    float t = 0.f;
    if (gl < J) {
        const float dx = pd[3 * gl] - pg[3 * gl], dy = pd[3 * gl + 1] - pg[3 * gl + 1];
        const float e = (dx * dx + dy * dy) / static_cast<float>(oks_var(gl, J)) / ((ag + ad) * 0.5f + 2.220446049250313e-16f) * 0.5f;
        t = __expf(-e);
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    const float est = t / static_cast<float>(J);
    const bool below = est < thr - OKS_BAND, above = est > thr + OKS_BAND;
    if (!__any_sync(0xffffffffu, !(below || above))) return above;
    return oks_pair<G>(pg, pd, ag, ad, J, gl) > thr;
}

template <int NT>
__global__ void __launch_bounds__(NT, 1)
nms_backproject_kernel(const NmsParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: keys u64[n2] | area f32[CT] | order i32[CT] | flag u8[CT]
    const int CT = p.CT;
    int n2cap = 1;
    while (n2cap < CT) n2cap <<= 1;
    uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);
    float* area = reinterpret_cast<float*>(keys + n2cap);
    int* order = reinterpret_cast<int*>(area + CT);
    unsigned char* dead = reinterpret_cast<unsigned char*>(order + CT);
    __shared__ int s_n, s_kept;
    __shared__ unsigned long long s_mask[NM_MATRIX_N];
    __shared__ int s_keep[NM_MATRIX_N];
    int* s_keep_soft = reinterpret_cast<int*>(area + CT) + CT + (CT + 3) / 4;   // after order[] and dead[]: kept slots of the soft path

    const int b = blockIdx.x, tid = threadIdx.x, J = p.J;
    const float* __restrict__ score = p.cand_score + static_cast<size_t>(b) * CT;
    const float* __restrict__ pose = p.cand_pose + static_cast<size_t>(b) * CT * J * 3;
    const float* __restrict__ center = p.cand_center + static_cast<size_t>(b) * CT * 3;
    const int P = p.P;
    int* kept_list = order;   // reused after ranking: kept_list[k] = candidate slot of output row k

    if (tid == 0) { s_n = 0; s_kept = 0; }
    pdl_wait();          // candidates' poses come from the refinement kernels in front of this one (last kernel of the chain: no trigger)
    __syncthreads();

    // ---- validity (das_head.py:763-769) + stable compaction in slot order ---------------------------
    // n is small against 1024 threads in every shipped config; a simple ordered scan keeps slot order.
    for (int base = 0; base < CT; base += NT) {
        const int c = base + tid;
        const bool ok = c < CT && (p.score_thr > 0.f ? (score[c] > p.score_thr) : true);
        const unsigned bal = __ballot_sync(0xffffffffu, ok);
        __shared__ int wcount[32];
        const int lane = tid & 31, warp = tid >> 5;
        if (lane == 0) wcount[warp] = __popc(bal);
        __syncthreads();
        int off = s_n;
        for (int w = 0; w < warp; ++w) off += wcount[w];
        if (ok) order[off + __popc(bal & ((1u << lane) - 1))] = c;
        __syncthreads();
        if (tid == 0) {
            int t = 0;
            for (int w = 0; w < NT / 32; ++w) t += wcount[w];
            s_n += t;
        }
        __syncthreads();
    }
    const int n = s_n;

    int kept = 0;
    if (p.nms_post > 0 && n > 0) {
        // ---- rank by (score desc, slot asc); areas (das_head.py:773-775) ---------------------------
        int n2 = 1;
        while (n2 < n) n2 <<= 1;
        for (int i = tid; i < n2; i += NT) {
            uint64_t k = 0;
            if (i < n) {
                const int c = order[i];
                k = (static_cast<uint64_t>(__float_as_uint(score[c])) << 32) | (0xFFFFFFFFu - static_cast<uint32_t>(c));
                const float* pc = pose + static_cast<size_t>(c) * J * 3;
                float x0 = pc[0], x1 = pc[0], y0 = pc[1], y1 = pc[1];
                for (int j = 1; j < J; ++j) {
                    x0 = fminf(x0, pc[3 * j]); x1 = fmaxf(x1, pc[3 * j]);
                    y0 = fminf(y0, pc[3 * j + 1]); y1 = fmaxf(y1, pc[3 * j + 1]);
                }
                area[c] = __fmul_rn(__fsub_rn(x1, x0), __fsub_rn(y1, y0));
            }
            keys[i] = k;
        }
        __syncthreads();
        if (n <= NT) {
            // few candidates: rank by counting (keys are distinct) -- three barriers instead of a bitonic network
            int rnk = 0, c = 0;
            if (tid < n) {
                const uint64_t k = keys[tid];
                c = order[tid];
                for (int j = 0; j < n; ++j) rnk += (keys[j] > k);
            }
            __syncthreads();
            if (tid < n) { order[rnk] = c; dead[tid] = 0; }
            __syncthreads();
        } else {
            // bitonic sort, descending
            for (int k = 2; k <= n2; k <<= 1) {
                for (int j = k >> 1; j > 0; j >>= 1) {
                    for (int t = tid; t < (n2 >> 1); t += NT) {
                        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                        const int q = i | j;
                        const bool desc = ((i & k) == 0);
                        const uint64_t x = keys[i], y = keys[q];
                        if ((x < y) == desc) { keys[i] = y; keys[q] = x; }
                    }
                    __syncthreads();
                }
            }
            for (int i = tid; i < n; i += NT) {
                order[i] = static_cast<int>(0xFFFFFFFFu - static_cast<uint32_t>(keys[i] & 0xFFFFFFFFull));
                dead[i] = 0;
            }
            __syncthreads();
        }

        const int limit = min(p.nms_post, n);
        if (p.soft) {
            // ---- soft OKS-NMS (pose_nms.py:129-194, gaussian rescoring): repeatedly take the best remaining
            // candidate and multiply every other remaining score by exp(-oks^2 / thr) (float32 like NumPy's).
            // `keys` is reused as the float32 working scores; ties go to the lower slot.
            float* cur = reinterpret_cast<float*>(keys);
            __shared__ unsigned long long s_best;
            for (int i = tid; i < n; i += NT) { cur[i] = score[order[i]]; dead[i] = 0; }
            __syncthreads();
            const int grp = tid >> 5, gl = tid & 31;
            for (int k = 0; k < limit; ++k) {
                if (tid == 0) s_best = 0ull;
                __syncthreads();
                unsigned long long mine = 0ull;
                for (int i = tid; i < n; i += NT)
                    if (!dead[i]) {
                        // positive floats order like their bit patterns; lower sorted position (= lower slot on ties) wins
                        const unsigned long long key = (static_cast<unsigned long long>(__float_as_uint(cur[i])) << 32) |
                                                       (0xFFFFFFFFu - static_cast<unsigned>(i));
                        mine = key > mine ? key : mine;
                    }
                if (mine) atomicMax(&s_best, mine | (1ull << 63));
                __syncthreads();
                const unsigned long long best = s_best;
                if (!best) break;
                const int bi = static_cast<int>(0xFFFFFFFFu - static_cast<unsigned>(best & 0xFFFFFFFFull));
                const int ci = order[bi];
                __syncthreads();
                if (tid == 0) { s_keep_soft[k] = ci; dead[bi] = 1; }
                kept = k + 1;
                for (int j = grp; j < n; j += NT / 32) {
                    if (dead[j] || j == bi) continue;
                    const int cj = order[j];
                    const float v = oks_pair<32>(pose + static_cast<size_t>(ci) * J * 3, pose + static_cast<size_t>(cj) * J * 3,
                                                 area[ci], area[cj], J, gl);
                    if (gl == 0) cur[j] = __fmul_rn(cur[j], expf(__fdiv_rn(-__fmul_rn(v, v), p.nms_thr)));
                }
                __syncthreads();
            }
            __syncthreads();
            for (int k = tid; k < kept; k += NT) kept_list[k] = s_keep_soft[k];
            __syncthreads();
        } else if (n <= NM_MATRIX_N) {
            // ---- all pairs in parallel, then a 64-bit mask greedy pass by one thread ---------------
            // ncu on the K = 64 decode (BASELINE config #4, 47 us per launch) showed the pair loop ISSUE-bound: with 16 lanes
            // per pair it spent 223 instructions per pair and warp -- a third of them the closed-form pair un-ranking that
            // every lane of a group repeated, a quarter the two IEEE divisions and the 64-bit pose addressing, then the
            // shuffle reduction.  Now ONE LANE owns a pair: the candidates' (x, y) are staged joint-major in shared memory (lanes
            // of a warp hold consecutive j: conflict-free), the lane walks the J joints with the float32 ESTIMATE (reciprocal
            // multiplies: its error stays ~1e-6, the band that sends a pair to the exact chain is 1e-3) and only a pair inside
            // the band runs the reference's float64 chain, serially in that lane: ~15 instructions per joint and pair instead
            // of ~220 per pair and warp-round.
            __shared__ float2 s_xy[DAS_MAX_JOINTS * NM_MATRIX_N];     // [joint][rank]
            __shared__ float s_ar[NM_MATRIX_N], s_iv[DAS_MAX_JOINTS];
            const int npairs = n * (n - 1) / 2;
            if (tid < n) { s_mask[tid] = 0ull; s_ar[tid] = area[order[tid]]; }
            if (tid < J) s_iv[tid] = 1.0f / static_cast<float>(oks_var(tid, J));
            for (int e = tid; e < n * J; e += NT) {
                const int i = e / J, j = e - i * J;
                const float* pc = pose + (static_cast<size_t>(order[i]) * J + j) * 3;
                s_xy[j * NM_MATRIX_N + i] = make_float2(pc[0], pc[1]);
            }
            __syncthreads();
            const float inv_J = 1.0f / static_cast<float>(J);
            for (int pr = tid; pr < npairs; pr += NT) {
                int i, j;
                unrank_pair(pr, n, i, j);
                const float inv_am = __fdividef(0.5f, (s_ar[i] + s_ar[j]) * 0.5f + 2.220446049250313e-16f);
                float t = 0.f;
                for (int k = 0; k < J; ++k) {
                    const float2 a = s_xy[k * NM_MATRIX_N + i], c2 = s_xy[k * NM_MATRIX_N + j];
                    const float dx = c2.x - a.x, dy = c2.y - a.y;
                    t += __expf(-((dx * dx + dy * dy) * s_iv[k] * inv_am));
                }
                const float est = t * inv_J;
                bool over = est > p.nms_thr + OKS_BAND;
                if (!over && !(est < p.nms_thr - OKS_BAND)) {
                    // inside the band (or NaN): pose_nms.py:84-91 in the reference's own precision
                    const int ci = order[i], cj = order[j];
                    const float* pg = pose + static_cast<size_t>(ci) * J * 3;
                    const float* pd = pose + static_cast<size_t>(cj) * J * 3;
                    over = oks_pair_serial(pg, pd, area[ci], area[cj], J) > p.nms_thr;
                }
                if (over) atomicOr(&s_mask[i], 1ull << j);
            }
            __syncthreads();
            if (tid == 0) {
                unsigned long long alive = (n == 64) ? ~0ull : ((1ull << n) - 1ull);
                int k = 0;
                for (int i = 0; i < n && k < limit; ++i) {
                    if ((alive >> i) & 1ull) {
                        s_keep[k++] = order[i];
                        alive &= ~s_mask[i];
                    }
                }
                s_kept = k;
            }
            __syncthreads();
            kept = s_kept;
            __syncthreads();
            if (tid < kept) kept_list[tid] = s_keep[tid];
            __syncthreads();
        } else {
            // ---- iterative greedy: one pass over the survivors per pick --------------------------------
            // kept slots are written over the front of `order` (k <= i always, and order[i] is read first)
            const int grp = tid >> 5, gl = tid & 31;
            for (int i = 0; i < n; ++i) {
                if (dead[i]) continue;                 // block-uniform (shared flag, synced below)
                const int ci = order[i];
                __syncthreads();
                if (tid == 0) kept_list[kept] = ci;
                ++kept;
                if (kept >= limit) break;
                for (int j = i + 1 + grp; j < n; j += NT / 32) {
                    if (dead[j]) continue;             // warp-uniform
                    const int cj = order[j];
                    const bool over = oks_over<32>(pose + static_cast<size_t>(ci) * J * 3, pose + static_cast<size_t>(cj) * J * 3,
                                                   area[ci], area[cj], J, gl, p.nms_thr);
                    if (gl == 0 && over) dead[j] = 1;
                }
                __syncthreads();
            }
            __syncthreads();
        }
    } else {
        // nms_post <= 0 (or nothing valid): survivors stay in slot order (das_head.py:770-772)
        kept = n;
        // order[] already holds the slots in ascending order == kept_list
    }

    // ---- outputs -----------------------------------------------------------------------------------
    const das_buffers& o = p.out;
    const das_peer_blocks& pe = p.peers;
    if (tid == 0) o.out_count[b] = kept;
    const double* __restrict__ cam = p.cam + static_cast<size_t>(b) * DAS_CAM_DOUBLES;
    const double K00 = cam[0], K01 = cam[1], K02 = cam[2], K10 = cam[3], K11 = cam[4], K12 = cam[5];
    const double* R = cam + 6;
    const double* T = cam + 15;
    const double detK = K00 * K11 - K01 * K10;
    const double nd = sqrt(K00 * K11);
    // inverse of R (adjugate / determinant)
    const double c00 = R[4] * R[8] - R[5] * R[7], c01 = R[2] * R[7] - R[1] * R[8], c02 = R[1] * R[5] - R[2] * R[4];
    const double c10 = R[5] * R[6] - R[3] * R[8], c11 = R[0] * R[8] - R[2] * R[6], c12 = R[2] * R[3] - R[0] * R[5];
    const double c20 = R[3] * R[7] - R[4] * R[6], c21 = R[1] * R[6] - R[0] * R[7], c22 = R[0] * R[4] - R[1] * R[3];
    const double detR = R[0] * c00 + R[1] * c10 + R[2] * c20;

    for (int k = tid; k < P; k += NT) {
        const bool live = k < kept;
        const int c = live ? kept_list[k] : 0;
        *(o.out_score + static_cast<size_t>(b) * P + k) = (live ? score[c] : 0.f);
        *(o.out_slot + static_cast<size_t>(b) * P + k) = (live ? c : -1);
        for (int d = 0; d < 3; ++d) *(o.out_center + (static_cast<size_t>(b) * P + k) * 3 + d) = (live ? center[c * 3 + d] : 0.f);
    }
    for (int e = tid; e < P * J; e += NT) {
        const int k = e / J, j = e - k * J;
        const size_t ob = ((static_cast<size_t>(b) * P + k) * J + j) * 3;
        if (k < kept) {
            const int c = kept_list[k];
            const float* pc = pose + static_cast<size_t>(c) * J * 3;
            const float fx = pc[3 * j], fy = pc[3 * j + 1], fz = pc[3 * j + 2];
            *(o.out_pose + ob) = (fx); *(o.out_pose + ob + 1) = (fy); *(o.out_pose + ob + 2) = (fz);
            const double zr = static_cast<double>(pc[3 * p.root + 2]);
            double Z = zr * nd + (static_cast<double>(fz) - zr);
            Z *= p.ddf;
            const double X0 = static_cast<double>(fx) - K02, X1 = static_cast<double>(fy) - K12;
            const double a = (K11 * X0 - K01 * X1) / detK;
            const double bb = (-K10 * X0 + K00 * X1) / detK;
            const double cx = a * Z, cy = bb * Z, cz = Z;
            *(o.out_cam + ob) = (cx); *(o.out_cam + ob + 1) = (cy); *(o.out_cam + ob + 2) = (cz);
            const double dx = cx - T[0], dy = cy - T[1], dz = cz - T[2];
            *(o.out_world + ob) = ((c00 * dx + c01 * dy + c02 * dz) / detR);
            *(o.out_world + ob + 1) = ((c10 * dx + c11 * dy + c12 * dz) / detR);
            *(o.out_world + ob + 2) = ((c20 * dx + c21 * dy + c22 * dz) / detR);
        } else {
            for (int d = 0; d < 3; ++d) { *(o.out_pose + ob + d) = (0.f); *(o.out_cam + ob + d) = (0.0); *(o.out_world + ob + d) = (0.0); }
        }
    }
    if (pe.n > 0) {
        __syncthreads();          // this image's slice of the local block is complete
        const size_t bP = static_cast<size_t>(b) * P;
        copy_chunk_to_peers<NT>(pe, o.out_count + b, 1);
        copy_chunk_to_peers<NT>(pe, o.out_score + bP, P);
        copy_chunk_to_peers<NT>(pe, o.out_slot + bP, P);
        copy_chunk_to_peers<NT>(pe, o.out_center + bP * 3, 3 * P);
        copy_chunk_to_peers<NT>(pe, o.out_pose + bP * J * 3, 3 * P * J);
        copy_chunk_to_peers<NT>(pe, o.out_cam + bP * J * 3, 6 * P * J);
        copy_chunk_to_peers<NT>(pe, o.out_world + bP * J * 3, 6 * P * J);
        if (pe.seq) {
            // publish: the last CTA to finish bumps the block's sequence number on every peer, after every CTA's stores
            // have been made visible system-wide -- a consumer on the peer that sees seq == s may read step s's results.
            // ONE system-scope fence per CTA, by the thread that takes the ticket, after the block barrier: the fence is
            // cumulative over the stores it observed through the barrier (the pattern of cooperative groups' grid sync).
            // A fence in every thread made this kernel 2x (1 peer) to 3x (7 peers) slower: 13 -> 28 -> 43 us per launch.
            __syncthreads();
            if (tid == 0) {
                __threadfence_system();
                const int ticket = atomicAdd(pe.ticket, 1);
                if (ticket == static_cast<int>(gridDim.x) - 1) {
                    *pe.ticket = 0;
                    const int sq = *pe.seq + 1;
                    *pe.seq = sq;
                    __threadfence_system();
                    for (int q = 0; q < pe.n; ++q)
                        *reinterpret_cast<int32_t*>(reinterpret_cast<unsigned char*>(pe.seq) + pe.delta[q]) = sq;
                }
            }
        }
    }
}

// Result all-gather as its own small kernel right behind nms_backproject_kernel (the default of das_plan): a few CTAs copy
// the rank's whole packed block to the same offsets of every peer's gathered buffer -- NVLink P2P stores to IPC-mapped
// memory, 16 bytes wide, every value loaded once and stored to all peers back to back -- then publish the step's sequence
// number on every peer behind ONE system-scope fence per CTA (cumulative over the stores observed through the block
// barrier; the last CTA to take a ticket writes the sequence words).  Measured against the stores fused into the NMS
// kernel's 64 CTAs: the system-scope fence and the posted stores stretched that kernel from 13 us to 28 us (1 peer) and
// 43 us (7 peers) ON 64 SMs; here they occupy PUB_CTAS SMs, next to the other streams' kernels.
constexpr int PUB_CTAS = 16;
__global__ void __launch_bounds__(256)
peer_publish_kernel(const das_peer_blocks pe, const uint4* __restrict__ block, int n16) {
    pdl_wait();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += gridDim.x * blockDim.x) {
        const uint4 v = __ldcg(block + i);
        for (int q = 0; q < pe.n; ++q)
            reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(const_cast<uint4*>(block)) + pe.delta[q])[i] = v;
    }
    if (!pe.seq) return;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        const int ticket = atomicAdd(pe.ticket, 1);
        if (ticket == static_cast<int>(gridDim.x) - 1) {
            *pe.ticket = 0;
            const int sq = *pe.seq + 1;
            *pe.seq = sq;
            __threadfence_system();
            for (int q = 0; q < pe.n; ++q)
                *reinterpret_cast<int32_t*>(reinterpret_cast<unsigned char*>(pe.seq) + pe.delta[q]) = sq;
        }
    }
}

}  // namespace das

extern "C" int32_t das_output_slots(int32_t cand_slots, int32_t nms_post) {
    return (nms_post > 0 && nms_post < cand_slots) ? nms_post : cand_slots;
}

extern "C" int das_nms_backproject(const das_decode_cfg* cfg, int32_t batch, int32_t cand_slots,
                                   const float* cand_score, const float* cand_pose, const float* cand_center,
                                   const double* cam, das_buffers out, void* stream) {
    return das_nms_backproject_peers(cfg, batch, cand_slots, cand_score, cand_pose, cand_center, cam, out, nullptr, stream);
}

extern "C" int das_nms_backproject_peers(const das_decode_cfg* cfg, int32_t batch, int32_t cand_slots,
                                         const float* cand_score, const float* cand_pose, const float* cand_center,
                                         const double* cam, das_buffers out, const das_peer_blocks* peers, void* stream) {
    using namespace das;
    DAS_REQUIRE(cfg && cand_score && cand_pose && cand_center && cam, DAS_ERR_ARG, "das_nms_backproject: null pointer");
    DAS_REQUIRE(!peers || (peers->n >= 0 && peers->n <= DAS_MAX_PEERS), DAS_ERR_ARG, "das_nms_backproject: %d peers", peers ? peers->n : 0);
    DAS_REQUIRE(out.out_count && out.out_score && out.out_slot && out.out_pose && out.out_center && out.out_cam && out.out_world,
                DAS_ERR_ARG, "das_nms_backproject: null output pointer");
    DAS_REQUIRE(batch >= 1 && cand_slots >= 1, DAS_ERR_ARG, "batch=%d cand_slots=%d", batch, cand_slots);
    DAS_REQUIRE(cand_slots <= NM_MAX_CAND, DAS_ERR_CAPACITY, "cand_slots=%d exceeds capacity %d", cand_slots, NM_MAX_CAND);
    DAS_REQUIRE(cfg->num_joints >= 1 && cfg->num_joints <= DAS_MAX_JOINTS, DAS_ERR_CAPACITY, "num_joints=%d", cfg->num_joints);
    NmsParams p{};
    p.B = batch; p.CT = cand_slots; p.P = das_output_slots(cand_slots, cfg->nms_post);
    p.J = cfg->num_joints; p.root = cfg->root_idx; p.nms_post = cfg->nms_post; p.soft = cfg->nms_soft;
    p.nms_thr = cfg->nms_thr; p.score_thr = cfg->score_thr;
    p.ddf = cfg->dataset_depth_factor == 0.0 ? 1.0 : cfg->dataset_depth_factor;
    p.cand_score = cand_score; p.cand_pose = cand_pose; p.cand_center = cand_center; p.cam = cam;
    p.out = out;
    if (peers) p.peers = *peers;
    int n2 = 1;
    while (n2 < cand_slots) n2 <<= 1;
    const size_t smem = static_cast<size_t>(n2) * 8 + static_cast<size_t>(cand_slots) * (4 + 4 + 1 + 4) + 32;
    static DeviceOnce attr_done;
    if (attr_done.need()) {
        DAS_CUDA_CHECK(cudaFuncSetAttribute(nms_backproject_kernel<NM_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            NM_MAX_CAND * 8 + NM_MAX_CAND * 13 + 32));
    }
    // few candidates (the usual case): 8 warps keep the ~20 block barriers of this latency-bound kernel cheap
    if (cand_slots <= 32) DAS_CUDA_CHECK(launch_chain(nms_backproject_kernel<256>, dim3(batch), dim3(256), smem, static_cast<cudaStream_t>(stream), chain_ctx().pdl, p));
    else DAS_CUDA_CHECK(launch_chain(nms_backproject_kernel<NM_THREADS>, dim3(batch), dim3(NM_THREADS), smem, static_cast<cudaStream_t>(stream), chain_ctx().pdl, p));
    DAS_CUDA_CHECK(cudaGetLastError());
    return DAS_OK;
}

// Copies `bytes` (a multiple of 16, 16-byte aligned) of the rank's packed output block to every peer's copy of it and
// publishes the sequence word (das_peer_blocks); enqueue right behind das_nms_backproject on the same stream.
extern "C" int das_peer_publish(const das_peer_blocks* peers, const void* local_block, int64_t bytes, void* stream) {
    using namespace das;
    DAS_REQUIRE(peers && local_block, DAS_ERR_ARG, "das_peer_publish: null pointer");
    DAS_REQUIRE(peers->n >= 0 && peers->n <= DAS_MAX_PEERS, DAS_ERR_ARG, "das_peer_publish: %d peers", peers->n);
    DAS_REQUIRE(bytes > 0 && (bytes & 15) == 0 && (reinterpret_cast<uintptr_t>(local_block) & 15) == 0, DAS_ERR_ARG,
                "das_peer_publish: block must be 16-byte aligned and sized");
    if (peers->n == 0) return DAS_OK;
    DAS_CUDA_CHECK(launch_chain(peer_publish_kernel, dim3(PUB_CTAS), dim3(256), 0, static_cast<cudaStream_t>(stream), chain_ctx().pdl,
                                *peers, static_cast<const uint4*>(local_block), static_cast<int>(bytes >> 4)));
    return DAS_OK;
}
