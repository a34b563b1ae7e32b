// Stage 3+4: gather at the selected centres, sparse evaluation of the LAST RecursiveUpdateLayer,
// head eval tail and joint assembly.
//
// Reference semantics (file:line relative to the reference root):
//   * gated blend + 1x1 projections ...... recursive_update.py:186-197
//   * progressive sampling ................ recursive_update.py:34-82, 9-31 (F.grid_sample bilinear,
//                                           padding_mode='zeros', align_corners=False)
//   * eval tail ........................... das_head.py:252-262
//   * gather + joint assembly ............. das_head.py:720-749
// The reference runs the refinement densely over every cell and gathers K cells afterwards; scores
// do not depend on it, so this kernel refines only the selected cells (divergence B).
// For one (centre p, joint j) that needs 37 feature rows of the NHWC map: p itself, the 4 bilinear
// corners of the current joint estimate t = p + O_j(p).xy, and the 4 corners of each of the 2*nh
// sampling heads.  Every row is projected with that joint's 17 1x1-conv rows (8 sampling offsets,
// 3 gate, 3 value, 3 confidence).
//
// Mapping: one warp per (centre, joint) work item, handed out by an atomic counter.  A lane owns
// C/32 channels of every row (two coalesced 128-bit loads per 1 KB row at C=256); dot products are
// finished with a transposing butterfly so 8 rows cost 9 shuffles per output instead of 40; rows are loaded
// in a lane-permuted order so that butterfly needs no selects, and FMAs are packed fma.rn.f32x2.
#include <algorithm>
#include <cstdlib>

#include "refine_common.cuh"
#include "row_cache.cuh"

namespace das {

constexpr int RS_WARPS = 8;  // warps per CTA

struct RefineParams {
    const das_levels* lv;
    const float* wpack;            // [J][NOUT][C] then [J][NOUT] biases
    const float* const* prev_uvd;  // nullptr or device array [n_levels] of joint-major maps [B][J][HW][4] (u, v, d, -)
    const float* const* prev_planes;  // LAZY kernels: device array [n_levels] of layer L-2's projection planes (DensePlanes); the
                                   // previous offsets are evaluated on demand at the sampled cells (dense_sample_cell)
    int batch;
    const float* scale_xy;         // [B,2]
    const float* cand_score;
    const int32_t* cand_index;
    float* cand_pose;
    float* cand_center;
    int* work_counter;             // [0] work queue head, [1] number of valid candidates, [4 + j] distinct rows of joint j (heads-only mode)
    float* urow;                   // heads-only mode: distinct feature rows per joint [J][row_cap][8] = {ptr lo, ptr hi, prev u, v, d, -, -, -}
    float* lrow;                   // heads-only mode: row records [B*CT*J][32 rows][4] = {distinct-row index (int bits; -1 = zero padding), bilinear weight, head offset x, y}
    float* item_asm;               // heads-only mode: per-item assembly record [B*CT*J][8] = {Px, Py, zq, sx, sy, stride, -, -}
    int32_t* valid_list;           // heads-only mode: candidate slots (b*CT+slot) that pass score_thr
    int row_cap;                   // capacity of one joint's distinct-row list (B*CT*32)
    int CT, J, root, nms_pre, layer;
    float depth_factor, z_norm, score_thr;
    int n_items;
    RowCacheView rc;               // heads-only mode, host zero-copy: device row cache (keys == nullptr: off)
};

// Previous-layer offsets (u, v, d) of joint j at cell `pix` of image b, level l.  LAZY: layer L-2's progressive sampling is
// evaluated here, from its projection planes, instead of being read from a map that dense_sample2_kernel filled for every
// cell -- the last layer looks at <= 33 cells per (candidate, joint), a few per cent of the map (recursive_update.py:220-235
// evaluates every layer densely; the values are the same ones, produced by the same device function).
template <bool LAZY>
__device__ __forceinline__ void prev_offsets(const RefineParams& p, const das_levels* __restrict__ lvp, int l, int b, int j, int pix,
                                             float (&pv)[3]) {
    const das_level_desc& d = lvp->lv[l];
    const int HW = d.H * d.W, J = p.J;
    if constexpr (LAZY) {
        const DensePlanes pl = dense_planes(const_cast<float*>(p.prev_planes[l]), p.batch, J, HW);
        const size_t bj = (static_cast<size_t>(b) * J + j) * HW;
        const float3 r = dense_sample_cell<4, false>(pl.s0 + bj, pl.s1 + bj, pl.oa + bj, pl.cb + bj, pix, d.W, d.H);
        pv[0] = r.x; pv[1] = r.y; pv[2] = r.z;
    } else {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            pv[k] = 0.f;
            if (p.prev_uvd) pv[k] = __ldg(p.prev_uvd[l] + ((static_cast<size_t>(b) * J + j) * HW + pix) * 4 + k);
            else if (!(k == 2 && j == p.root))
                pv[k] = InMap(d.pose, lvp->in_dtype)((static_cast<size_t>(b) * (3 + 6 * J) + 3 + 3 * j + k) * HW + pix) * (k < 2 ? d.scale_uv : d.scale_d);
        }
    }
}

template <int CPL, int NH, int MINB, bool HEADS_ONLY, bool ROW_CACHE = false, bool LAZY = false>
__global__ void __launch_bounds__(RS_WARPS * 32, MINB)
refine_sparse_kernel(const RefineParams p) {
    constexpr int C = CPL * 32;
    constexpr int NOUT = 2 * NH + 9;
    constexpr int O_GATE = 2 * NH, O_VAL = 2 * NH + 3, O_CONF = 2 * NH + 6;
    const int lane = threadIdx.x & 31;
    pdl_wait();          // candidates (and, for L > 1, the dense layers' maps) come from the kernels in front of this one
    pdl_trigger();
    const das_levels* __restrict__ lvp = p.lv;
    const int nl = lvp->n_levels;
    const int J = p.J;

    for (;;) {
        int item = 0;
        if (lane == 0) item = atomicAdd(p.work_counter, 1);
        item = __shfl_sync(FULL, item, 0);
        if (item >= p.n_items) break;
        const int j = item % J;
        const int cs = item / J;          // b * CT + slot
        const int b = cs / p.CT, slot = cs - b * p.CT;
        const float score = __ldg(p.cand_score + cs);
        if (p.score_thr > 0.f && !(score > p.score_thr)) continue;   // dropped by das_head.py:763-769 later

        int l = 0, s0 = 0;
        for (; l < nl - 1; ++l) {
            const int ns = level_slots(lvp->lv[l].H * lvp->lv[l].W, p.nms_pre);
            if (slot < s0 + ns) break;
            s0 += ns;
        }
        const das_level_desc& d = lvp->lv[l];
        const int H = d.H, W = d.W, HW = H * W;
        const int idx = __ldg(p.cand_index + cs);
        const int y = idx / W, x = idx - y * W;
        const float* __restrict__ F = d.feats[p.layer] + static_cast<size_t>(b) * HW * C;
        const InMap pose(d.pose, lvp->in_dtype);
        const size_t pb = static_cast<size_t>(b) * (3 + 6 * J) * HW;     // first element of image b in the pose map
        const float* __restrict__ prev = p.prev_uvd ? p.prev_uvd[l] : nullptr;
        if (prev) prev += (static_cast<size_t>(b) * J + j) * HW * 4;
        const float* __restrict__ Wj = p.wpack + static_cast<size_t>(j) * NOUT * C;
        const float* __restrict__ Bj = p.wpack + static_cast<size_t>(J) * NOUT * C + j * NOUT;
        const float fW = static_cast<float>(W), fH = static_cast<float>(H);

        // previous-layer offset of joint j, dim k at cell `pix` (das_head.py:243-249 for layer 0)
        auto prev_at = [&](int pix, int k) -> float {
            if (prev) return __ldg(prev + static_cast<size_t>(pix) * 4 + k);
            if (k == 2 && j == p.root) return 0.0f;
            const float raw = pose(pb + static_cast<size_t>(3 + 3 * j + k) * HW + pix);
            return raw * (k < 2 ? d.scale_uv : d.scale_d);
        };

        // ---- phase 1: cell p -> S_j(p) (2*NH), gate, value -> blended offset O_j(p) ---------------
        float S[2 * NH];
        float O[3];
        {
            float pvc[3];
            if constexpr (LAZY) prev_offsets<true>(p, lvp, l, b, j, idx, pvc);
            else { pvc[0] = prev_at(idx, 0); pvc[1] = prev_at(idx, 1); pvc[2] = prev_at(idx, 2); }
            const float* prow = (ROW_CACHE && p.rc.cand_rows) ? p.rc.cand_rows + static_cast<size_t>(cs) * C : F + static_cast<size_t>(idx) * C;
            const Row<CPL> f = load_row<CPL>(prow, lane, true);
            float acc[2 * NH + 6];
#pragma unroll
            for (int o = 0; o < 2 * NH + 6; ++o) {
                const Row<CPL> w = load_row<CPL>(Wj + o * C, lane, true);
                acc[o] = warp_allsum(dot_row<CPL>(f, w));
            }
#pragma unroll
            for (int o = 0; o < 2 * NH; ++o) S[o] = acc[o] + __ldg(Bj + o);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float g = sigmoid_acc(acc[O_GATE + k] + __ldg(Bj + O_GATE + k));
                const float n = acc[O_VAL + k] + __ldg(Bj + O_VAL + k);
                O[k] = __fadd_rn(__fmul_rn(1.0f - g, pvc[k]), __fmul_rn(g, n));
            }
        }

        // ---- phase 2: sampling offsets bilinearly read at t = p + O.xy ("from target" heads) -----
        float hx[2 * NH], hy[2 * NH];   // per-head sampling offset (x, y)
        {
            const Corner ct = make_corner(sample_coord(x, O[0], fW), sample_coord(y, O[1], fH), W, H);
            Row<CPL> fi;
#pragma unroll
            for (int q = 0; q < CPL / 4; ++q) fi.v[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            float wsum = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const bool ok = corner_ok(ct, k, W, H);
                const float wk = ok ? corner_wgt(ct, k) : 0.f;
                const float* rowp = F + static_cast<size_t>(ok ? corner_pix(ct, k, W) : 0) * C;
                Row<CPL> f;
                if constexpr (ROW_CACHE) {
                    // host zero-copy mode: neighbouring candidates and the sampling phase want these very rows again, so
                    // whoever asks first fetches the row over PCIe into the device row buffer and publishes it; everybody
                    // else waits for that copy instead of fetching the row a second time
                    const float* src = rowp;
                    int slot = -1;
                    uint32_t hh = 0;
                    bool won = false;
                    if (p.rc.keys && ok) {
                        if (lane == 0) {
                            const int ins = row_cache_insert(p.rc, reinterpret_cast<unsigned long long>(rowp), hh, slot, false);
                            won = ins > 0;
                            if (ins == 0) slot = row_cache_wait(p.rc, hh);      // bounded; -2 = read the row at home
                        }
                        won = __shfl_sync(FULL, won, 0);
                        slot = __shfl_sync(FULL, slot, 0);
                        if (!won && slot >= 0) src = p.rc.rows + static_cast<size_t>(slot) * C;
                    }
                    if (src == rowp) {
                        f = load_row<CPL>(rowp, lane, ok);
                    } else {                                  // written by another SM during this kernel: no .nc path
#pragma unroll
                        for (int q = 0; q < CPL / 4; ++q) f.v[q] = __ldcg(reinterpret_cast<const float4*>(src + q * 128 + 4 * lane));
                    }
                    if (won && slot >= 0) {
#pragma unroll
                        for (int q = 0; q < CPL / 4; ++q)
                            *reinterpret_cast<float4*>(p.rc.rows + static_cast<size_t>(slot) * C + q * 128 + 4 * lane) = f.v[q];
                        __threadfence();       // every lane's part of the row is visible device-wide before the slot is
                        __syncwarp();
                        if (lane == 0) row_cache_publish(p.rc, hh, slot);
                    }
                } else {
                    f = load_row<CPL>(rowp, lane, ok);
                }
                wsum += wk;
#pragma unroll
                for (int q = 0; q < CPL / 4; ++q) {
                    fi.v[q].x = fmaf(wk, f.v[q].x, fi.v[q].x);
                    fi.v[q].y = fmaf(wk, f.v[q].y, fi.v[q].y);
                    fi.v[q].z = fmaf(wk, f.v[q].z, fi.v[q].z);
                    fi.v[q].w = fmaf(wk, f.v[q].w, fi.v[q].w);
                }
            }
            // the projection is linear, so Bil(W f + b) = W Bil(f) + b * (in-bounds weight sum)
#pragma unroll
            for (int h = 0; h < NH; ++h) {
                const Row<CPL> wx = load_row<CPL>(Wj + (2 * h) * C, lane, true);
                const Row<CPL> wy = load_row<CPL>(Wj + (2 * h + 1) * C, lane, true);
                const float sx = warp_allsum(dot_row<CPL>(fi, wx)) + wsum * __ldg(Bj + 2 * h);
                const float sy = warp_allsum(dot_row<CPL>(fi, wy)) + wsum * __ldg(Bj + 2 * h + 1);
                hx[h] = sx + O[0];           // recursive_update.py:59
                hy[h] = sy + O[1];
                hx[NH + h] = S[2 * h];       // "from source" heads, recursive_update.py:62
                hy[NH + h] = S[2 * h + 1];
            }
        }

        if constexpr (HEADS_ONLY) {
            // phases 1-2 only.  Everything the tensor-core kernel (refine_tc.cu) needs per gathered row is prepared
            // here, one lane per row (head = lane >> 2, corner = lane & 3): feature-row pointer (null = zero
            // padding), bilinear weight, previous offset at that cell, the head's sampling offset.
            static_assert(NH == 4, "32 rows per item = 8 heads x 4 corners");
            {
                const int h = lane >> 2, ck2 = lane & 3;
                float hxv = hx[0], hyv = hy[0];
#pragma unroll
                for (int i = 1; i < 2 * NH; ++i) { hxv = (h == i) ? hx[i] : hxv; hyv = (h == i) ? hy[i] : hyv; }
                const Corner c = make_corner(sample_coord(x, hxv, fW), sample_coord(y, hyv, fH), W, H);
                const bool ok = corner_ok(c, ck2, W, H);
                const int pix = ok ? corner_pix(c, ck2, W) : -1;
                const float wk = ok ? corner_wgt(c, ck2) : 0.f;
                // The gate / value / confidence projections and the blended offset depend on (cell, joint) only, and the
                // 8 heads x 4 corners of an item mostly land on a handful of cells: keep ONE entry per distinct cell in the
                // joint's row list (what the tensor-core kernel multiplies) and let the 32 row records point at it.
                const unsigned same = __match_any_sync(FULL, pix);
                const int leader = __ffs(same) - 1;
                const bool is_leader = ok && lane == leader;
                const unsigned lead_mask = __ballot_sync(FULL, is_leader);
                const int n_u = __popc(lead_mask);
                int base = 0;
                if (lane == 0 && n_u) base = atomicAdd(p.work_counter + 4 + j, n_u);
                base = __shfl_sync(FULL, base, 0);
                const int my_u = __popc(lead_mask & ((1u << lane) - 1u));
                const int u_of_leader = __shfl_sync(FULL, my_u, leader);
                const int gidx = ok ? base + u_of_leader : -1;
                if (is_leader) {
                    const float* ptr = F + static_cast<size_t>(pix) * C;
                    float pvl[3];
                    if constexpr (LAZY) prev_offsets<true>(p, lvp, l, b, j, pix, pvl);
                    else { pvl[0] = prev_at(pix, 0); pvl[1] = prev_at(pix, 1); pvl[2] = prev_at(pix, 2); }
                    const float pv0 = pvl[0], pv1 = pvl[1], pv2 = pvl[2];
                    const unsigned long long pb = reinterpret_cast<unsigned long long>(ptr);
                    float4* dst = reinterpret_cast<float4*>(p.urow + (static_cast<size_t>(j) * p.row_cap + base + my_u) * 8);
                    dst[0] = make_float4(__uint_as_float(static_cast<unsigned>(pb)), __uint_as_float(static_cast<unsigned>(pb >> 32)), pv0, pv1);
                    dst[1] = make_float4(pv2, 0.f, 0.f, 0.f);
                }
                reinterpret_cast<float4*>(p.lrow)[static_cast<size_t>(item) * 32 + lane] =
                    make_float4(__int_as_float(gidx), wk, hxv, hyv);
            }
            if (lane == 0) {
                // eval-tail / assembly inputs of this item (das_head.py:254-262, 725-743), and the centre for joint 0
                const float sx = __ldg(p.scale_xy + 2 * b), sy = __ldg(p.scale_xy + 2 * b + 1);
                const float qf = sqrtf(sx * sy);
                const float st = static_cast<float>(d.stride), half = static_cast<float>(d.stride / 2);
                float z = pose(pb + 2 * static_cast<size_t>(HW) + idx) * d.scale_depth;
                z = __fdiv_rn(z, p.depth_factor);
                const float zq = __fmul_rn(z, qf);
                const float Px = static_cast<float>(x) * st + half, Py = static_cast<float>(y) * st + half;
                float4* a = reinterpret_cast<float4*>(p.item_asm + static_cast<size_t>(item) * 8);
                a[0] = make_float4(Px, Py, zq, sx);
                a[1] = make_float4(sy, st, 0.f, 0.f);
                if (j == 0) {
                    const float offx = pose(pb + idx) * d.scale_offset;
                    const float offy = pose(pb + static_cast<size_t>(HW) + idx) * d.scale_offset;
                    p.cand_center[static_cast<size_t>(cs) * 3 + 0] = __fdiv_rn(__fsub_rn(Px, offx), sx);
                    p.cand_center[static_cast<size_t>(cs) * 3 + 1] = __fdiv_rn(__fsub_rn(Py, offy), sy);
                    p.cand_center[static_cast<size_t>(cs) * 3 + 2] = zq;
                    p.valid_list[atomicAdd(p.work_counter + 1, 1)] = cs;
                }
            }
            continue;
        }

        // ---- phase 3: 2*NH heads, one head (4 corner rows) per batch; blended offset + confidence per corner ----
        const int ck = (lane >> 3) & 3;     // corner of the batch this lane post-processes (lanes 8ck..8ck+7)
        const int dim = lane & 7;           // u, v, d for dim < 3; the other lanes of the group idle in the epilogue
        float sm_m = -INFINITY, sm_e = 0.f, sm_v = 0.f;   // online softmax over the heads (recursive_update.py:29-31)
#pragma unroll
        for (int h = 0; h < 2 * NH; ++h) {
            const Corner c = make_corner(sample_coord(x, hx[h], fW), sample_coord(y, hy[h], fH), W, H);
            // this lane's own corner: issue the previous-offset gather now so it overlaps the row loads
            const bool ok_me = corner_ok(c, ck, W, H) && dim < 3;
            const int pix_me = ok_me ? corner_pix(c, ck, W) : 0;
            const float prev_me = ok_me ? prev_at(pix_me, dim) : 0.f;
            // lane-permuted batch: row (k ^ ck) goes to f[k] (see reduce4_permuted); a warp-wide load then reads
            // one full 128 B line from each of the 4 corner cells
            Row<CPL> f[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int kk = k ^ ck;
                const bool ok = corner_ok(c, kk, W, H);
                f[k] = load_row<CPL>(F + static_cast<size_t>(ok ? corner_pix(c, kk, W) : 0) * C, lane, ok);
            }
            float res[9];
#pragma unroll
            for (int o = 0; o < 9; ++o) {
                const Row<CPL> w = load_row<CPL>(Wj + (O_GATE + o) * C, lane, true);
                float acc[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) acc[k] = dot_row<CPL>(f[k], w);
                res[o] = reduce4_permuted(acc);
            }
            float val = 0.f, cf = 0.f;
            if (ok_me) {
                const float rg = dim == 0 ? res[0] : (dim == 1 ? res[1] : res[2]);
                const float rn = dim == 0 ? res[3] : (dim == 1 ? res[4] : res[5]);
                const float rc = dim == 0 ? res[6] : (dim == 1 ? res[7] : res[8]);
                const float g = sigmoid_acc(rg + __ldg(Bj + O_GATE + dim));
                const float n = rn + __ldg(Bj + O_VAL + dim);
                const float o = __fadd_rn(__fmul_rn(1.0f - g, prev_me), __fmul_rn(g, n));
                const float wk = corner_wgt(c, ck);
                val = o * wk;
                cf = (rc + __ldg(Bj + O_CONF + dim)) * wk;
            }
            // bilinear sum over the 4 corners (lane bits 3,4)
            val += __shfl_xor_sync(FULL, val, 8);
            cf += __shfl_xor_sync(FULL, cf, 8);
            val += __shfl_xor_sync(FULL, val, 16);
            cf += __shfl_xor_sync(FULL, cf, 16);
            const float hv = val + (dim == 0 ? hx[h] : (dim == 1 ? hy[h] : 0.f));   // + diff, recursive_update.py:72-75, 28
            const float m_new = fmaxf(sm_m, cf);
            const float s_old = expf(sm_m - m_new), e_new = expf(cf - m_new);
            sm_e = sm_e * s_old + e_new;
            sm_v = sm_v * s_old + hv * e_new;
            sm_m = m_new;
        }
        const float out = sm_v / sm_e;

        // ---- eval tail + assembly (das_head.py:254-262, 725-743) ----------------------------------
        if (lane < 3) {
            const float sx = __ldg(p.scale_xy + 2 * b), sy = __ldg(p.scale_xy + 2 * b + 1);
            const float qf = sqrtf(sx * sy);
            const float st = static_cast<float>(d.stride);
            const float half = static_cast<float>(d.stride / 2);
            float z = pose(pb + 2 * static_cast<size_t>(HW) + idx) * d.scale_depth;
            z = __fdiv_rn(z, p.depth_factor);
            const float zq = __fmul_rn(z, qf);
            float v;
            if (lane == 0) v = __fdiv_rn(__fadd_rn(__fmul_rn(out, st), static_cast<float>(x) * st + half), sx);
            else if (lane == 1) v = __fdiv_rn(__fadd_rn(__fmul_rn(out, st), static_cast<float>(y) * st + half), sy);
            else v = __fadd_rn((j == p.root) ? 0.0f : __fmul_rn(out, p.z_norm), zq);
            p.cand_pose[(static_cast<size_t>(cs) * J + j) * 3 + lane] = v;
            if (j == 0) {
                float c;
                if (lane == 2) c = zq;
                else {
                    const float off = pose(pb + static_cast<size_t>(lane) * HW + idx) * d.scale_offset;
                    const float P = static_cast<float>(lane == 0 ? x : y) * st + half;
                    c = __fdiv_rn(__fsub_rn(P, off), lane == 0 ? sx : sy);
                }
                p.cand_center[static_cast<size_t>(cs) * 3 + lane] = c;
            }
        }
    }
}

// Phases 1-2 for NB (4 or 8) candidates at a time (the production path of das_refine_heads on device-resident maps).
//
// The warp-per-item kernel above reads its joint's 22 weight rows (22 KB) for every item, and with the items of all
// joints interleaved those rows come from L2 every time: ncu showed 117 MB of L2 -> L1 traffic per launch for 48 MB of
// feature rows, and 34 us.  Here a task is (joint j, block of NB consecutive candidates): a group of 32/NB lanes owns one
// candidate; the NB feature rows live in registers, every weight row is loaded ONCE per NB items and the NB dot products
// are finished by one transposing butterfly (reduce8_permuted / reduce4_permuted).  Tasks are handed out joint-major, so
// the warps of an SM work on the same joint and its weight rows stay in L1.  The row records are then written for the NB
// items with ONE reservation in the joint's distinct-row list per task (a per-item atomic would put NB dependent global
// round trips on the warp's critical path).
constexpr int H8_WARPS = 4;

template <int NB>
__device__ __forceinline__ float reduce_nb(const float (&a)[NB]) {
    if constexpr (NB == 8) return reduce8_permuted(a);
    else if constexpr (NB == 4) return reduce4_permuted(a);
    else return reduce2_permuted(a);
}

// SPLIT: the kernel stops after phase 2 and leaves each item's 16 head offsets in the first 64 bytes of the item's row-record
// slot; refine_records_kernel (one warp per item, 4x the parallelism, no batch-serial passes) writes the records.
template <int CPL, int NH, int NB, bool SPLIT, bool LAZY = false>
__global__ void __launch_bounds__(H8_WARPS * 32, NB == 4 ? 5 : (NB == 2 ? 8 : 4))     // NB = 4: 96 registers, 20 warps per SM (2 960 task slots: BASELINE config #2 has 2 400 tasks -> one wave)
refine_heads8_kernel(const RefineParams p) {
    static_assert(NH == 4, "32 rows per item = 8 heads x 4 corners");
    static_assert(NB == 2 || NB == 4 || NB == 8, "candidates per task");
    static_assert(!LAZY || SPLIT, "on-demand previous offsets: the split path only (refine_records_kernel evaluates the leaders')");
    constexpr int C = CPL * 32;
    constexpr int NOUT = 2 * NH + 9;
    constexpr int O_GATE = 2 * NH, O_VAL = 2 * NH + 3;
    constexpr int GL = 32 / NB;                        // lanes per candidate
    constexpr int SH = NB == 8 ? 2 : (NB == 4 ? 3 : 4); // log2(GL)
    constexpr int NW = 2 * NH + 6;                     // weight rows phases 1-2 use: S (2*NH), gate 3, value 3
    __shared__ float s_head[H8_WARPS][NB][4 * NH];     // per candidate of the block: hx[2*NH], hy[2*NH]
    __shared__ __align__(16) float s_w[NW * C];        // this CTA's joint: its weight rows, staged once (the global copies sat
                                                       // on every task's critical path: ~22 dependent L1 / L2 round trips)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int J = p.J;
    // A CTA serves ONE joint: blockIdx = j * cpj + g; its warps walk the candidate blocks g*H8_WARPS + warp, + cpj*H8_WARPS, ..
    const int cpj = gridDim.x / J;
    const int j = blockIdx.x / cpj, g = blockIdx.x - j * cpj;
    if (j >= J) return;
    {
        // weights are static: staged BEFORE the dependency wait, so the copy overlaps the predecessor kernel's tail
        const float* __restrict__ Wsrc = p.wpack + static_cast<size_t>(j) * NOUT * C;
        for (int c = threadIdx.x; c < NW * C / 4; c += H8_WARPS * 32)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(s_w + 4 * c))), "l"(Wsrc + 4 * c) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    pdl_wait();          // candidates (and, for L > 1, the dense layers' maps) come from the kernels in front of this one
    pdl_trigger();
    const int r = (lane >> SH) & (NB - 1);             // candidate of the block this lane's group owns
    const das_levels* __restrict__ lvp = p.lv;
    const int nl = lvp->n_levels;
    const int n_cand = p.n_items / J;
    const int n_blocks = (n_cand + NB - 1) / NB;
    bool weights_ready = false;

    auto level_of = [&](int slot) {
        int l = 0, s0 = 0;
        for (; l < nl - 1; ++l) {
            const int ns = level_slots(lvp->lv[l].H * lvp->lv[l].W, p.nms_pre);
            if (slot < s0 + ns) break;
            s0 += ns;
        }
        return l;
    };

    // static hand-out (a ticket counter would put one more global round trip in front of every task).  The trip count is
    // CTA-uniform, so the one block barrier behind the weight copy is reached by every warp.
    const int trips = (n_blocks - g * H8_WARPS + cpj * H8_WARPS - 1) / (cpj * H8_WARPS);
    for (int t = 0; t < trips; ++t) {
        const int cb = g * H8_WARPS + warp + t * cpj * H8_WARPS;
        const int cs = cb * NB + r;
        bool valid = cb < n_blocks && cs < n_cand;
        if (valid && p.score_thr > 0.f && !(__ldg(p.cand_score + cs) > p.score_thr)) valid = false;   // dropped by das_head.py:763-769 later
        const unsigned vmask = __ballot_sync(FULL, valid);
        const float* __restrict__ Bj = p.wpack + static_cast<size_t>(J) * NOUT * C + j * NOUT;

        // ---- this group's candidate ------------------------------------------------------------------------------
        const int csv = valid ? cs : 0;
        const int b = csv / p.CT, slot = csv - b * p.CT;
        const int l = level_of(slot);
        const das_level_desc& d = lvp->lv[l];
        const int H = d.H, W = d.W, HW = H * W;
        const int idx = valid ? __ldg(p.cand_index + csv) : 0;
        const int y = idx / W, x = idx - y * W;
        const float* __restrict__ F = d.feats[p.layer] + static_cast<size_t>(b) * HW * C;
        const float fW = static_cast<float>(W), fH = static_cast<float>(H);
        const float rW = __frcp_rn(fW), rH = __frcp_rn(fH);

        // ---- phase 1: the NB rows F(p) x {S 2*NH, gate 3, value 3} -------------------------------------------------
        float S[2 * NH], O[3];
        {
            float prev[3] = {0.f, 0.f, 0.f};
            if (valid) prev_offsets<LAZY>(p, lvp, l, b, j, idx, prev);
            const float* own = valid ? F + static_cast<size_t>(idx) * C : nullptr;
            Row<CPL> f[NB];
#pragma unroll
            for (int k = 0; k < NB; ++k) {                // lane-permuted: candidate (k ^ r) goes to f[k]
                const float* src = reinterpret_cast<const float*>(__shfl_xor_sync(FULL, reinterpret_cast<unsigned long long>(own), k << SH));
                f[k] = load_row<CPL>(src ? src : p.wpack, lane, src != nullptr);
            }
            if (!weights_ready) {                         // first trip: the joint's weight rows have landed for the whole CTA
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                __syncthreads();
                weights_ready = true;
            }
            float res[2 * NH + 6];
#pragma unroll
            for (int o = 0; o < 2 * NH + 6; ++o) {
                const Row<CPL> w = load_row_smem<CPL>(s_w + o * C, lane);
                float acc[NB];
#pragma unroll
                for (int k = 0; k < NB; ++k) acc[k] = dot_row<CPL>(f[k], w);
                res[o] = reduce_nb<NB>(acc) + __ldg(Bj + o);
            }
#pragma unroll
            for (int o = 0; o < 2 * NH; ++o) S[o] = res[o];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float g = sigmoid_acc(res[O_GATE + k]);
                O[k] = __fadd_rn(__fmul_rn(1.0f - g, prev[k]), __fmul_rn(g, res[O_VAL + k]));
            }
        }

        if (vmask == 0u) continue;         // nothing above score_thr in this block (after the CTA-wide barrier of the first trip)

        // ---- phase 2: sampling offsets read bilinearly at t = p + O.xy -------------------------------------------
        {
            const Corner ct = make_corner(sample_coord(x, O[0], fW, rW), sample_coord(y, O[1], fH, rH), W, H);
            const float* cptr[4];
            float cw[4];
            float wsum = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const bool ok = valid && corner_ok(ct, k, W, H);
                cw[k] = ok ? corner_wgt(ct, k) : 0.f;
                cptr[k] = ok ? F + static_cast<size_t>(corner_pix(ct, k, W)) * C : nullptr;
                wsum += cw[k];
            }
            // pull the 4 corner rows of this candidate towards L2 now: the accumulation below can only keep a few rows in
            // flight (registers), so without the hint every round of it waits for DRAM
            {
                const float* mine = cptr[lane & 3];
                if (mine) {
#pragma unroll
                    for (int q = (lane >> 2) & (GL / 4 - 1); q < C * 4 / 128; q += GL / 4)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(mine + q * 32));
                }
            }
            Row<CPL> fi[NB];
#pragma unroll
            for (int k = 0; k < NB; ++k) {
#pragma unroll
                for (int q = 0; q < CPL / 4; ++q) fi[k].v[q] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float* src = reinterpret_cast<const float*>(__shfl_xor_sync(FULL, reinterpret_cast<unsigned long long>(cptr[c]), k << SH));
                    const float wk = __shfl_xor_sync(FULL, cw[c], k << SH);
                    const Row<CPL> f = load_row<CPL>(src ? src : p.wpack, lane, src != nullptr);
#pragma unroll
                    for (int q = 0; q < CPL / 4; ++q) {
                        fi[k].v[q].x = fmaf(wk, f.v[q].x, fi[k].v[q].x);
                        fi[k].v[q].y = fmaf(wk, f.v[q].y, fi[k].v[q].y);
                        fi[k].v[q].z = fmaf(wk, f.v[q].z, fi[k].v[q].z);
                        fi[k].v[q].w = fmaf(wk, f.v[q].w, fi[k].v[q].w);
                    }
                }
            }
            // the projection is linear, so Bil(W f + b) = W Bil(f) + b * (in-bounds weight sum)
            float st[2 * NH];
#pragma unroll
            for (int o = 0; o < 2 * NH; ++o) {
                const Row<CPL> w = load_row_smem<CPL>(s_w + o * C, lane);
                float acc[NB];
#pragma unroll
                for (int k = 0; k < NB; ++k) acc[k] = dot_row<CPL>(fi[k], w);
                st[o] = reduce_nb<NB>(acc) + wsum * __ldg(Bj + o);
            }
            if ((lane & (GL - 1)) == 0) {
                if constexpr (SPLIT) {
                    if (valid) {
                        // {hx[0..7], hy[0..7]} of item (cs, j): heads 0..3 "from target" (recursive_update.py:59), 4..7 "from source" (:62)
                        float4* dst = reinterpret_cast<float4*>(p.lrow) + (static_cast<size_t>(cs) * J + j) * 32;
                        dst[0] = make_float4(st[0] + O[0], st[2] + O[0], st[4] + O[0], st[6] + O[0]);
                        dst[1] = make_float4(S[0], S[2], S[4], S[6]);
                        dst[2] = make_float4(st[1] + O[1], st[3] + O[1], st[5] + O[1], st[7] + O[1]);
                        dst[3] = make_float4(S[1], S[3], S[5], S[7]);
                    }
                } else {
                    float* sh = s_head[warp][r];
#pragma unroll
                    for (int h = 0; h < NH; ++h) {
                        sh[h] = st[2 * h] + O[0];                 // "from target" heads, recursive_update.py:59
                        sh[2 * NH + h] = st[2 * h + 1] + O[1];
                        sh[NH + h] = S[2 * h];                    // "from source" heads, recursive_update.py:62
                        sh[2 * NH + NH + h] = S[2 * h + 1];
                    }
                }
            }
        }
        if constexpr (SPLIT) continue;
        __syncwarp();

        // ---- row records: lane = (head = lane >> 2, corner = lane & 3) of one candidate at a time -----------------
        // Straight-line over the NB candidates (validity is a predicate, not a branch) so that their dependent chains --
        // coordinate chain, match, loads -- interleave; everything the owner group already knows about its candidate
        // (cell, map size, reciprocals) arrives by shuffle instead of being recomputed (an integer division and two
        // reciprocals per candidate sat on the critical path: 37 % of the kernel's stall samples were in this pass).
        // pass 1: every candidate's sampled cells and their de-duplication inside the item (registers)
        int r_pix[NB], r_g[NB];                       // cell index (-1 = outside the map), position among the item's distinct cells
        float r_w[NB], r_hx[NB], r_hy[NB];
        unsigned r_lead[NB];                          // leader lanes of the item's distinct cells
        int m_idx[NB], m_b[NB], m_l[NB], m_W[NB], m_HW[NB];
        float r_pv[NB][3];                            // leaders: previous offset (u, v, d) at their cell
        int total_u = 0;
#pragma unroll
        for (int rr = 0; rr < NB; ++rr) {
            const int src = rr * GL;
            const bool v = (vmask >> src) & 1u;
            m_idx[rr] = __shfl_sync(FULL, idx, src);
            m_b[rr] = __shfl_sync(FULL, b, src);
            m_l[rr] = __shfl_sync(FULL, l, src);
            const int x2 = __shfl_sync(FULL, x, src), y2 = __shfl_sync(FULL, y, src);
            const int W2 = __shfl_sync(FULL, W, src), H2 = __shfl_sync(FULL, H, src);
            const float rW2 = __shfl_sync(FULL, rW, src), rH2 = __shfl_sync(FULL, rH, src);
            m_W[rr] = W2; m_HW[rr] = H2 * W2;
            const das_level_desc& d2 = lvp->lv[m_l[rr]];
            const int h = lane >> 2, ck2 = lane & 3;
            const float hxv = s_head[warp][rr][h], hyv = s_head[warp][rr][2 * NH + h];
            const Corner c = make_corner(sample_coord(x2, hxv, static_cast<float>(W2), rW2), sample_coord(y2, hyv, static_cast<float>(H2), rH2), W2, H2);
            const bool ok = v && corner_ok(c, ck2, W2, H2);
            const int pix = ok ? corner_pix(c, ck2, W2) : -1;
            // one entry per DISTINCT sampled cell in the joint's row list (see the warp-per-item kernel above)
            const unsigned same = __match_any_sync(FULL, pix);
            const int leader = __ffs(same) - 1;
            const unsigned lead_mask = __ballot_sync(FULL, ok && lane == leader);
            const int my_u = __popc(lead_mask & ((1u << lane) - 1u));
            const int u_of_leader = __shfl_sync(FULL, my_u, leader);
            r_pix[rr] = pix; r_w[rr] = ok ? corner_wgt(c, ck2) : 0.f; r_hx[rr] = hxv; r_hy[rr] = hyv;
            r_lead[rr] = lead_mask;
            r_g[rr] = ok ? total_u + u_of_leader : -1;
            total_u += __popc(lead_mask);
            // the leaders' previous offsets: requested now, so the loads fly while the reservation below makes its round trip
            const bool lead = (lead_mask >> lane) & 1u;
            const int pixs = lead ? pix : 0;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                float pv = 0.f;
                if (p.prev_uvd) {
                    if (lead) pv = __ldg(p.prev_uvd[m_l[rr]] + ((static_cast<size_t>(m_b[rr]) * J + j) * m_HW[rr] + pixs) * 4 + k);
                } else if (!(k == 2 && j == p.root)) {
                    if (lead) pv = InMap(d2.pose, lvp->in_dtype)((static_cast<size_t>(m_b[rr]) * (3 + 6 * J) + 3 + 3 * j + k) * m_HW[rr] + pixs) *
                                   (k < 2 ? d2.scale_uv : d2.scale_d);
                }
                r_pv[rr][k] = pv;
            }
        }
        int base = 0;
        if (lane == 0 && total_u) base = atomicAdd(p.work_counter + 4 + j, total_u);
        base = __shfl_sync(FULL, base, 0);
        // pass 2: the records
#pragma unroll
        for (int rr = 0; rr < NB; ++rr) {
            const bool v = (vmask >> (rr * GL)) & 1u;
            const int cs2 = cb * NB + rr;
            const int item = cs2 * J + j;
            const int b2 = m_b[rr];
            const das_level_desc& d2 = lvp->lv[m_l[rr]];
            const int W2 = m_W[rr], HW2 = m_HW[rr];
            const int idx2 = m_idx[rr];
            const int pix = r_pix[rr];
            if (v && ((r_lead[rr] >> lane) & 1u)) {
                const unsigned long long pb = reinterpret_cast<unsigned long long>(d2.feats[p.layer] + (static_cast<size_t>(b2) * HW2 + pix) * C);
                float4* dst = reinterpret_cast<float4*>(p.urow + (static_cast<size_t>(j) * p.row_cap + base + r_g[rr]) * 8);
                dst[0] = make_float4(__uint_as_float(static_cast<unsigned>(pb)), __uint_as_float(static_cast<unsigned>(pb >> 32)), r_pv[rr][0], r_pv[rr][1]);
                dst[1] = make_float4(r_pv[rr][2], 0.f, 0.f, 0.f);
            }
            if (v)
                reinterpret_cast<float4*>(p.lrow)[static_cast<size_t>(item) * 32 + lane] =
                    make_float4(__int_as_float(pix >= 0 ? base + r_g[rr] : -1), r_w[rr], r_hx[rr], r_hy[rr]);
            if (v && lane == rr) {
                // eval-tail / assembly inputs of this item (das_head.py:254-262, 725-743), and the centre for joint 0;
                // lane rr takes candidate rr, so the NB assembly records are written side by side, not one after the other
                const InMap pose2(d2.pose, lvp->in_dtype);
                const size_t pb2 = static_cast<size_t>(b2) * (3 + 6 * J) * HW2;
                const int y2 = idx2 / W2, x2 = idx2 - y2 * W2;
                const float sx = __ldg(p.scale_xy + 2 * b2), sy = __ldg(p.scale_xy + 2 * b2 + 1);
                const float qf = sqrtf(sx * sy);
                const float stv = static_cast<float>(d2.stride), half = static_cast<float>(d2.stride / 2);
                float z = pose2(pb2 + 2 * static_cast<size_t>(HW2) + idx2) * d2.scale_depth;
                z = __fdiv_rn(z, p.depth_factor);
                const float zq = __fmul_rn(z, qf);
                const float Px = static_cast<float>(x2) * stv + half, Py = static_cast<float>(y2) * stv + half;
                float4* a = reinterpret_cast<float4*>(p.item_asm + static_cast<size_t>(item) * 8);
                a[0] = make_float4(Px, Py, zq, sx);
                a[1] = make_float4(sy, stv, 0.f, 0.f);
                if (j == 0) {
                    const float offx = pose2(pb2 + idx2) * d2.scale_offset;
                    const float offy = pose2(pb2 + static_cast<size_t>(HW2) + idx2) * d2.scale_offset;
                    p.cand_center[static_cast<size_t>(cs2) * 3 + 0] = __fdiv_rn(__fsub_rn(Px, offx), sx);
                    p.cand_center[static_cast<size_t>(cs2) * 3 + 1] = __fdiv_rn(__fsub_rn(Py, offy), sy);
                    p.cand_center[static_cast<size_t>(cs2) * 3 + 2] = zq;
                    p.valid_list[atomicAdd(p.work_counter + 1, 1)] = cs2;
                }
            }
        }
        __syncwarp();
    }
}

// Row records of the split phase 1-2 path: one warp per (candidate, joint) item, lane = (head = lane >> 2, corner = lane & 3).
// Reads the item's 16 head offsets left by refine_heads8_kernel<.., SPLIT> in the first 64 bytes of its row-record slot, then
// overwrites the slot with the 32 records; distinct sampled cells -> the joint's row list; assembly record; centre; valid_list.
// 32 registers (8 CTAs per SM; 36 bytes of spills): the kernel is a chain of three dependent global round trips per item, so
// what pays is warps in flight and room for the other streams' kernels -- single-stream stage time unchanged, pipelined
// throughput +3 % (config #2: 1.23 -> 1.27 M images/s, config #4: 239 -> 249 k).  The LAZY variant inlines dense_sample_cell
// and keeps its registers.
template <int NH, bool LAZY = false>
__global__ void __launch_bounds__(256, LAZY ? 1 : 8)
refine_records_kernel(const RefineParams p) {
    static_assert(NH == 4, "32 rows per item = 8 heads x 4 corners");
    constexpr int C = 256;
    const int lane = threadIdx.x & 31;
    pdl_wait();
    pdl_trigger();
    const das_levels* __restrict__ lvp = p.lv;
    const int nl = lvp->n_levels, J = p.J;
    for (int item = blockIdx.x * 8 + (threadIdx.x >> 5); item < p.n_items; item += gridDim.x * 8) {
        const int cs = item / J, j = item - cs * J;
        if (p.score_thr > 0.f && !(__ldg(p.cand_score + cs) > p.score_thr)) continue;
        const int b = cs / p.CT, slot = cs - b * p.CT;
        int l = 0, s0 = 0;
        for (; l < nl - 1; ++l) {
            const int ns = level_slots(lvp->lv[l].H * lvp->lv[l].W, p.nms_pre);
            if (slot < s0 + ns) break;
            s0 += ns;
        }
        const das_level_desc& d = lvp->lv[l];
        const int H = d.H, W = d.W, HW = H * W;
        const int idx = __ldg(p.cand_index + cs);
        const int y = idx / W, x = idx - y * W;
        const float fW = static_cast<float>(W), fH = static_cast<float>(H);
        float4* slotp = reinterpret_cast<float4*>(p.lrow) + static_cast<size_t>(item) * 32;
        const int h = lane >> 2, ck2 = lane & 3;
        // written by the previous kernel: plain (coherent) loads, not the read-only path
        const float hxv = reinterpret_cast<const volatile float*>(slotp)[h], hyv = reinterpret_cast<const volatile float*>(slotp)[8 + h];
        __syncwarp();                                  // every lane has its offsets before the slot is overwritten
        const Corner c = make_corner(sample_coord(x, hxv, fW, __frcp_rn(fW)), sample_coord(y, hyv, fH, __frcp_rn(fH)), W, H);
        const bool ok = corner_ok(c, ck2, W, H);
        const int pix = ok ? corner_pix(c, ck2, W) : -1;
        const float wk = ok ? corner_wgt(c, ck2) : 0.f;
        const unsigned same = __match_any_sync(FULL, pix);
        const int leader = __ffs(same) - 1;
        const bool is_leader = ok && lane == leader;
        const unsigned lead_mask = __ballot_sync(FULL, is_leader);
        const int n_u = __popc(lead_mask);
        // the leaders' previous offsets fly while the reservation makes its round trip
        float pv[3] = {0.f, 0.f, 0.f};
        if (is_leader) prev_offsets<LAZY>(p, lvp, l, b, j, pix, pv);
        int base = 0;
        if (lane == 0 && n_u) base = atomicAdd(p.work_counter + 4 + j, n_u);
        base = __shfl_sync(FULL, base, 0);
        const int my_u = __popc(lead_mask & ((1u << lane) - 1u));
        const int u_of_leader = __shfl_sync(FULL, my_u, leader);
        if (is_leader) {
            const unsigned long long pb = reinterpret_cast<unsigned long long>(d.feats[p.layer] + (static_cast<size_t>(b) * HW + pix) * C);
            float4* dst = reinterpret_cast<float4*>(p.urow + (static_cast<size_t>(j) * p.row_cap + base + my_u) * 8);
            dst[0] = make_float4(__uint_as_float(static_cast<unsigned>(pb)), __uint_as_float(static_cast<unsigned>(pb >> 32)), pv[0], pv[1]);
            dst[1] = make_float4(pv[2], 0.f, 0.f, 0.f);
        }
        slotp[lane] = make_float4(__int_as_float(ok ? base + u_of_leader : -1), wk, hxv, hyv);
        if (lane == 0) {
            // eval-tail / assembly inputs of this item (das_head.py:254-262, 725-743), and the centre for joint 0
            const InMap pose(d.pose, lvp->in_dtype);
            const size_t pb = static_cast<size_t>(b) * (3 + 6 * J) * HW;
            const float sx = __ldg(p.scale_xy + 2 * b), sy = __ldg(p.scale_xy + 2 * b + 1);
            const float qf = sqrtf(sx * sy);
            const float stv = static_cast<float>(d.stride), half = static_cast<float>(d.stride / 2);
            float z = pose(pb + 2 * static_cast<size_t>(HW) + idx) * d.scale_depth;
            z = __fdiv_rn(z, p.depth_factor);
            const float zq = __fmul_rn(z, qf);
            const float Px = static_cast<float>(x) * stv + half, Py = static_cast<float>(y) * stv + half;
            float4* a = reinterpret_cast<float4*>(p.item_asm + static_cast<size_t>(item) * 8);
            a[0] = make_float4(Px, Py, zq, sx);
            a[1] = make_float4(sy, stv, 0.f, 0.f);
            if (j == 0) {
                const float offx = pose(pb + idx) * d.scale_offset;
                const float offy = pose(pb + static_cast<size_t>(HW) + idx) * d.scale_offset;
                p.cand_center[static_cast<size_t>(cs) * 3 + 0] = __fdiv_rn(__fsub_rn(Px, offx), sx);
                p.cand_center[static_cast<size_t>(cs) * 3 + 1] = __fdiv_rn(__fsub_rn(Py, offy), sy);
                p.cand_center[static_cast<size_t>(cs) * 3 + 2] = zq;
                p.valid_list[atomicAdd(p.work_counter + 1, 1)] = cs;
            }
        }
    }
}

// refine = 0: the pose maps are already final (reference get_poses contract): plain gather + assembly.
__global__ void gather_assemble_kernel(const RefineParams p, int batch) {
    const int J = p.J;
    const int per = 3 * J + 3;
    const long long total = static_cast<long long>(batch) * p.CT * per;
    const das_levels* __restrict__ lvp = p.lv;
    for (long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; t < total;
         t += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int e = static_cast<int>(t % per);
        const int cs = static_cast<int>(t / per);
        const int b = cs / p.CT, slot = cs - b * p.CT;
        int l = 0, s0 = 0;
        for (; l < lvp->n_levels - 1; ++l) {
            const int ns = level_slots(lvp->lv[l].H * lvp->lv[l].W, p.nms_pre);
            if (slot < s0 + ns) break;
            s0 += ns;
        }
        const das_level_desc& d = lvp->lv[l];
        const int W = d.W, HW = d.H * d.W;
        const int idx = __ldg(p.cand_index + cs);
        const int y = idx / W, x = idx - y * W;
        const InMap pose(d.pose, lvp->in_dtype);
        const size_t pb = static_cast<size_t>(b) * (3 + 6 * J) * HW;
        const float sx = __ldg(p.scale_xy + 2 * b), sy = __ldg(p.scale_xy + 2 * b + 1);
        const float qf = sqrtf(sx * sy);
        const float st = static_cast<float>(d.stride), half = static_cast<float>(d.stride / 2);
        const float zq = __fmul_rn(pose(pb + 2 * static_cast<size_t>(HW) + idx), qf);
        if (e < 3 * J) {
            const int k = e % 3;
            const float raw = pose(pb + static_cast<size_t>(3 + e) * HW + idx);
            float v;
            if (k == 0) v = __fdiv_rn(__fadd_rn(raw, static_cast<float>(x) * st + half), sx);
            else if (k == 1) v = __fdiv_rn(__fadd_rn(raw, static_cast<float>(y) * st + half), sy);
            else v = __fadd_rn(raw, zq);
            p.cand_pose[static_cast<size_t>(cs) * 3 * J + e] = v;
        } else {
            const int k = e - 3 * J;
            float c;
            if (k == 2) c = zq;
            else {
                const float off = pose(pb + static_cast<size_t>(k) * HW + idx);
                const float P = static_cast<float>(k == 0 ? x : y) * st + half;
                c = __fdiv_rn(__fsub_rn(P, off), k == 0 ? sx : sy);
            }
            p.cand_center[static_cast<size_t>(cs) * 3 + k] = c;
        }
    }
}

// nn.Conv2d layouts -> joint-major rows {so(2nh), uw(3), uv(3), sc(3)} x C, then biases
__global__ void pack_weights_kernel(const float* so_w, const float* so_b, const float* sc_w, const float* sc_b,
                                    const float* uw_w, const float* uw_b, const float* uv_w, const float* uv_b,
                                    float* dst, int J, int nh, int C) {
    const int NOUT = 2 * nh + 9;
    const long long nW = static_cast<long long>(J) * NOUT * C;
    const long long total = nW + static_cast<long long>(J) * NOUT;
    for (long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; t < total;
         t += static_cast<long long>(gridDim.x) * blockDim.x) {
        const bool is_b = t >= nW;
        const long long u = is_b ? t - nW : t / C;
        const int c = is_b ? 0 : static_cast<int>(t % C);
        const int j = static_cast<int>(u / NOUT), o = static_cast<int>(u % NOUT);
        const float *w, *bb;
        int row;
        if (o < 2 * nh) { w = so_w; bb = so_b; row = j * 2 * nh + o; }
        else if (o < 2 * nh + 3) { w = uw_w; bb = uw_b; row = j * 3 + (o - 2 * nh); }
        else if (o < 2 * nh + 6) { w = uv_w; bb = uv_b; row = j * 3 + (o - 2 * nh - 3); }
        else { w = sc_w; bb = sc_b; row = j * 3 + (o - 2 * nh - 6); }
        dst[t] = is_b ? bb[row] : w[static_cast<long long>(row) * C + c];
    }
}

}  // namespace das

extern "C" int64_t das_packed_weight_floats(const das_decode_cfg* cfg) {
    if (!cfg) return 0;
    return static_cast<int64_t>(cfg->num_joints) * (2 * cfg->num_heads + 9) * (cfg->feat_channels + 1);
}

extern "C" int das_pack_weights(const das_decode_cfg* cfg, const float* so_w, const float* so_b,
                                const float* sc_w, const float* sc_b, const float* uw_w, const float* uw_b,
                                const float* uv_w, const float* uv_b, float* dst, void* stream) {
    using namespace das;
    DAS_REQUIRE(cfg && so_w && so_b && sc_w && sc_b && uw_w && uw_b && uv_w && uv_b && dst, DAS_ERR_ARG,
                "das_pack_weights: null pointer");
    pack_weights_kernel<<<kSMs, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        so_w, so_b, sc_w, sc_b, uw_w, uw_b, uv_w, uv_b, dst, cfg->num_joints, cfg->num_heads, cfg->feat_channels);
    DAS_CUDA_CHECK(cudaGetLastError());
    return DAS_OK;
}

extern "C" int das_gather_refine_assemble(const das_levels* d_levels, const das_levels* h_levels,
                                          const das_decode_cfg* cfg, const float* weights,
                                          const float* const* prev_uvd, const float* scale_xy,
                                          const float* cand_score, const int32_t* cand_index,
                                          int32_t cand_slots, float* cand_pose, float* cand_center,
                                          int32_t* work_counter, void* stream) {
    using namespace das;
    DAS_REQUIRE(d_levels && h_levels && cfg && scale_xy && cand_score && cand_index && cand_pose && cand_center,
                DAS_ERR_ARG, "das_gather_refine_assemble: null pointer");
    DAS_REQUIRE(cfg->num_joints >= 1 && cfg->num_joints <= DAS_MAX_JOINTS, DAS_ERR_CAPACITY, "num_joints=%d", cfg->num_joints);
    DAS_REQUIRE(cfg->root_idx >= 0 && cfg->root_idx < cfg->num_joints, DAS_ERR_ARG, "root_idx=%d", cfg->root_idx);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    RefineParams p{};
    p.lv = d_levels;
    p.wpack = weights;
    p.prev_uvd = prev_uvd;
    p.scale_xy = scale_xy;
    p.cand_score = cand_score;
    p.cand_index = cand_index;
    p.cand_pose = cand_pose;
    p.cand_center = cand_center;
    p.work_counter = work_counter;
    p.CT = cand_slots;
    p.J = cfg->num_joints;
    p.root = cfg->root_idx;
    p.nms_pre = cfg->nms_pre;
    p.layer = cfg->num_layers - 1;
    p.depth_factor = cfg->depth_factor;
    p.z_norm = cfg->z_norm;
    p.score_thr = cfg->score_thr;
    const long long items = static_cast<long long>(h_levels->batch) * cand_slots * cfg->num_joints;
    DAS_REQUIRE(items < (1ll << 31), DAS_ERR_CAPACITY, "too many work items");
    p.n_items = static_cast<int>(items);
    if (!cfg->refine) {
        const long long total = static_cast<long long>(h_levels->batch) * cand_slots * (3 * cfg->num_joints + 3);
        const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, 4 * kSMs));
        gather_assemble_kernel<<<grid, 256, 0, st>>>(p, h_levels->batch);
        DAS_CUDA_CHECK(cudaGetLastError());
        return DAS_OK;
    }
    DAS_REQUIRE(weights && work_counter, DAS_ERR_ARG, "refine=1 needs packed weights and a work counter");
    DAS_REQUIRE(cfg->num_layers >= 1 && cfg->num_layers <= DAS_MAX_LAYERS, DAS_ERR_CAPACITY, "num_layers=%d", cfg->num_layers);
    DAS_REQUIRE(cfg->num_heads == 4, DAS_ERR_UNSUPPORTED, "num_heads=%d: only 4 is built", cfg->num_heads);
    DAS_CUDA_CHECK(cudaMemsetAsync(work_counter, 0, sizeof(int32_t), st));
    const int threads = RS_WARPS * 32;
    switch (cfg->feat_channels) {
        case 128: refine_sparse_kernel<4, 4, 3, false><<<kSMs * 3, threads, 0, st>>>(p); break;
        case 256: refine_sparse_kernel<8, 4, 3, false><<<kSMs * 3, threads, 0, st>>>(p); break;
        case 512: refine_sparse_kernel<16, 4, 2, false><<<kSMs * 2, threads, 0, st>>>(p); break;
        default:
            set_error("feat_channels=%d: only 128/256/512 are built", cfg->feat_channels);
            return DAS_ERR_UNSUPPORTED;
    }
    DAS_CUDA_CHECK(cudaGetLastError());
    return DAS_OK;
}

// Phases 1-2 of the sparse refinement only (feeds das_refine_tc / das_refine_finish): per (candidate, joint) item it
// appends the item's DISTINCT sampled cells to that joint's row list (scratch->unique_rows), writes the 32 row records
// that point into it (scratch->row_records), the item's assembly record (scratch->item_records) and, for joint 0, the
// centre and the candidate's entry in valid_list.  counters: [0] work-queue head, [1] number of valid candidates,
// [4 + j] distinct rows of joint j.
static int g_heads_force = 0;
// diagnostics: 0 = automatic choice, 1 = warp-per-item kernel, 2 = batched kernel (process-wide; for tests and A/B timing)
extern "C" int das_debug_force_heads_kernel(int32_t mode) {
    if (mode < 0 || mode > 2) return DAS_ERR_ARG;
    g_heads_force = mode;
    return DAS_OK;
}

extern "C" int das_refine_heads(const das_levels* d_levels, const das_levels* h_levels, const das_decode_cfg* cfg,
                                const float* weights, const float* const* prev_uvd, const float* scale_xy,
                                const float* cand_score, const int32_t* cand_index, int32_t cand_slots,
                                const das_refine_scratch* scratch, float* cand_center, const das_row_cache* rc, void* stream) {
    using namespace das;
    DAS_REQUIRE(d_levels && h_levels && cfg && weights && scale_xy && cand_score && cand_index && scratch && cand_center,
                DAS_ERR_ARG, "das_refine_heads: null pointer");
    DAS_REQUIRE(scratch->unique_rows && scratch->row_records && scratch->item_records && scratch->valid_list && scratch->counters,
                DAS_ERR_ARG, "das_refine_heads: null scratch buffer");
    DAS_REQUIRE(cfg->feat_channels == 256 && cfg->num_heads == 4, DAS_ERR_UNSUPPORTED,
                "das_refine_heads is built for feat_channels=256, num_heads=4");
    DAS_REQUIRE(cfg->num_joints >= 1 && cfg->num_joints <= DAS_MAX_JOINTS, DAS_ERR_CAPACITY, "num_joints=%d", cfg->num_joints);
    const long long want_cap = static_cast<long long>(h_levels->batch) * cand_slots * 32;
    DAS_REQUIRE(scratch->row_cap >= want_cap, DAS_ERR_ARG, "das_refine_heads: row_cap=%d < batch*cand_slots*32=%lld", scratch->row_cap, want_cap);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    RefineParams p{};
    p.lv = d_levels; p.wpack = weights; p.prev_uvd = prev_uvd; p.cand_score = cand_score; p.cand_index = cand_index;
    p.work_counter = scratch->counters; p.urow = scratch->unique_rows; p.lrow = scratch->row_records;
    p.item_asm = scratch->item_records; p.valid_list = scratch->valid_list; p.row_cap = scratch->row_cap;
    p.scale_xy = scale_xy; p.cand_center = cand_center;
    p.CT = cand_slots; p.J = cfg->num_joints; p.root = cfg->root_idx; p.nms_pre = cfg->nms_pre; p.layer = cfg->num_layers - 1;
    p.depth_factor = cfg->depth_factor; p.z_norm = cfg->z_norm; p.score_thr = cfg->score_thr;
    const long long items = static_cast<long long>(h_levels->batch) * cand_slots * cfg->num_joints;
    DAS_REQUIRE(items < (1ll << 31), DAS_ERR_CAPACITY, "too many work items");
    p.n_items = static_cast<int>(items);
    p.rc = row_cache_view(rc);
    // scratch->prev_planes: layer L-2's sampling is evaluated on demand from its projection planes (prev_uvd is ignored)
    const bool lazy = scratch->prev_planes != nullptr;
    p.prev_planes = scratch->prev_planes;
    p.batch = h_levels->batch;
    DAS_REQUIRE(!lazy || cfg->num_layers > 1, DAS_ERR_ARG, "das_refine_heads: prev_planes given but num_layers=%d", cfg->num_layers);
    // [0] queue head, [1] n_valid, [4..4+J) per-joint distinct-row counts; [2] (the peer-store ticket) is left alone.
    // Inside das_plan's chain das_score_topk has already cleared them (no memset nodes between the kernels).
    const ChainCtx& cx = chain_ctx();
    if (!cx.counters_cleared) {
        DAS_CUDA_CHECK(cudaMemsetAsync(scratch->counters, 0, 2 * sizeof(int32_t), st));
        DAS_CUDA_CHECK(cudaMemsetAsync(scratch->counters + 4, 0, DAS_MAX_JOINTS * sizeof(int32_t), st));
    }
    // DAS_HEADS_KERNEL=item / =batch, or das_debug_force_heads_kernel(), forces one of the two kernels (A/B timing, parity tests)
    static const char* force_env = std::getenv("DAS_HEADS_KERNEL");
    const int force = g_heads_force ? g_heads_force : (force_env ? (force_env[0] == 'i' ? 1 : 2) : 0);
    // fewer items than warp slots (one image, a few centres): one warp per item finishes sooner than 4 items per warp
    const bool per_item = force ? force == 1 : items <= 24LL * kSMs;
    if (p.rc.keys) {
        if (lazy) DAS_CUDA_CHECK(launch_chain(refine_sparse_kernel<8, 4, 3, true, true, true>, dim3(kSMs * 3), dim3(RS_WARPS * 32), 0, st, cx.pdl, p));
        else DAS_CUDA_CHECK(launch_chain(refine_sparse_kernel<8, 4, 3, true, true>, dim3(kSMs * 3), dim3(RS_WARPS * 32), 0, st, cx.pdl, p));
    } else if (per_item) {
        if (lazy) DAS_CUDA_CHECK(launch_chain(refine_sparse_kernel<8, 4, 3, true, false, true>, dim3(kSMs * 3), dim3(RS_WARPS * 32), 0, st, cx.pdl, p));
        else DAS_CUDA_CHECK(launch_chain(refine_sparse_kernel<8, 4, 3, true, false>, dim3(kSMs * 3), dim3(RS_WARPS * 32), 0, st, cx.pdl, p));
    } else {
        static const int nb = std::getenv("DAS_HEADS_NB") ? std::atoi(std::getenv("DAS_HEADS_NB")) : 4;
        // grid = J x cpj CTAs: CTA (j, g) serves joint j, its H8_WARPS warps walk candidate blocks g*H8_WARPS + warp, + cpj*H8_WARPS, ...
        const long long n_blocks = (items / cfg->num_joints + nb - 1) / nb;
        const long long cap = std::max<long long>(1, ((nb == 8 ? 4LL : (nb == 2 ? 8LL : 5LL)) * kSMs) / cfg->num_joints);
        const int cpj = static_cast<int>(std::max<long long>(1, std::min<long long>((n_blocks + H8_WARPS - 1) / H8_WARPS, cap)));
        const int grid = cpj * cfg->num_joints;
        // DAS_HEADS_SPLIT=0: records inside the batched kernel (one launch); default: phases 1-2 batched, records by a warp per item
        static const bool split = !(std::getenv("DAS_HEADS_SPLIT") && std::getenv("DAS_HEADS_SPLIT")[0] == '0');
        DAS_REQUIRE(!lazy || (split && nb == 4), DAS_ERR_UNSUPPORTED,
                    "das_refine_heads: prev_planes needs the default split kernels (unset DAS_HEADS_SPLIT / DAS_HEADS_NB)");
        if (split && nb == 4) {
            const int grid_r = static_cast<int>(std::max<long long>(1, std::min<long long>((items + 7) / 8, 8LL * kSMs)));
            if (lazy) {
                DAS_CUDA_CHECK(launch_chain(refine_heads8_kernel<8, 4, 4, true, true>, dim3(grid), dim3(H8_WARPS * 32), 0, st, cx.pdl, p));
                DAS_CUDA_CHECK(launch_chain(refine_records_kernel<4, true>, dim3(grid_r), dim3(256), 0, st, cx.pdl, p));
            } else {
                DAS_CUDA_CHECK(launch_chain(refine_heads8_kernel<8, 4, 4, true>, dim3(grid), dim3(H8_WARPS * 32), 0, st, cx.pdl, p));
                DAS_CUDA_CHECK(launch_chain(refine_records_kernel<4>, dim3(grid_r), dim3(256), 0, st, cx.pdl, p));
            }
            chain_ctx().extra_launches += 1;
        } else if (nb == 8) DAS_CUDA_CHECK(launch_chain(refine_heads8_kernel<8, 4, 8, false>, dim3(grid), dim3(H8_WARPS * 32), 0, st, cx.pdl, p));
        else if (nb == 2) DAS_CUDA_CHECK(launch_chain(refine_heads8_kernel<8, 4, 2, false>, dim3(grid), dim3(H8_WARPS * 32), 0, st, cx.pdl, p));
        else DAS_CUDA_CHECK(launch_chain(refine_heads8_kernel<8, 4, 4, false>, dim3(grid), dim3(H8_WARPS * 32), 0, st, cx.pdl, p));
    }
    DAS_CUDA_CHECK(cudaGetLastError());
    return DAS_OK;
}

// Host zero-copy mode: F(p) of every candidate above score_thr -> rc->cand_rows[b*CT+slot][C], once per candidate
// instead of once per (candidate, joint) warp of das_refine_heads.
namespace das {
__global__ void __launch_bounds__(256)
cand_rows_kernel(const das_levels* __restrict__ lvp, const float* __restrict__ cand_score, const int32_t* __restrict__ cand_index,
                 int CT, int n_cand, int nms_pre, int layer, int C, float score_thr, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    for (int cs = blockIdx.x * 8 + (threadIdx.x >> 5); cs < n_cand; cs += gridDim.x * 8) {
        if (score_thr > 0.f && !(__ldg(cand_score + cs) > score_thr)) continue;
        const int b = cs / CT, slot = cs - b * CT;
        int l = 0, s0 = 0;
        for (; l < lvp->n_levels - 1; ++l) {
            const int ns = level_slots(lvp->lv[l].H * lvp->lv[l].W, nms_pre);
            if (slot < s0 + ns) break;
            s0 += ns;
        }
        const das_level_desc& d = lvp->lv[l];
        const int idx = __ldg(cand_index + cs);
        if (idx < 0) continue;
        const float4* src = reinterpret_cast<const float4*>(d.feats[layer] + (static_cast<size_t>(b) * d.H * d.W + idx) * C);
        float4* dst = reinterpret_cast<float4*>(out + static_cast<size_t>(cs) * C);
        for (int q = lane; q < C / 4; q += 32) dst[q] = __ldg(src + q);
    }
}
}  // namespace das

extern "C" int das_refine_cand_rows(const das_levels* d_levels, const das_levels* h_levels, const das_decode_cfg* cfg,
                                    const float* cand_score, const int32_t* cand_index, int32_t cand_slots,
                                    const das_row_cache* rc, void* stream) {
    using namespace das;
    DAS_REQUIRE(d_levels && h_levels && cfg && cand_score && cand_index && rc && rc->cand_rows, DAS_ERR_ARG,
                "das_refine_cand_rows: null pointer");
    const int n = h_levels->batch * cand_slots;
    cand_rows_kernel<<<std::min(kSMs * 4, (n + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        d_levels, cand_score, cand_index, cand_slots, n, cfg->nms_pre, cfg->num_layers - 1, cfg->feat_channels, cfg->score_thr,
        rc->cand_rows);
    DAS_CUDA_CHECK(cudaGetLastError());
    return DAS_OK;
}
