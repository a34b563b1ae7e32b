// Dense RecursiveUpdateLayer over a whole level -- layers 1..L-1 when recursive_update.num_layers > 1
// (the last layer is always evaluated sparsely by refine_sparse.cu).
//
// Reference: recursive_update.py:220-235 (layer), :186-197 (four 1x1 projections + gated blend),
// :34-82 and :9-31 (progressive sampling).  The reference materialises a [B*J*2nh, 6, H, W] tensor with
// repeat_interleave/cat and runs two full grid_samples over it; here the layer is two kernels:
//   dense_project_kernel  F[B,HW,C] (NHWC) x W[17J,C] -> proj[B,HW,J,14] = {S(2nh), conf(3), blended O(3)}
//   dense_sample_kernel   one thread per (cell, joint): 4 + 2nh*4 bilinear taps in proj -> uvd_out[B,HW,3J]
// This first version of the projection is fp32 SIMT (warp per 8 cells, same building blocks as the sparse
// kernel); it is the parity reference for the tcgen05 GEMM that replaces it.
#include <algorithm>
#include <cstdlib>

#include "refine_common.cuh"

namespace das {

constexpr int DP_WARPS = 8;

struct DenseParams {
    const das_levels* lv;
    const float* wpack;
    const float* uvd_in;   // nullptr -> scaled raw uvd from lv.pose (layer 0); else joint-major [B][J][HW][4]
    float* uvd_out;        // joint-major [B][J][HW][4] (u, v, d, -)
    float* proj;           // four joint-major planes, see DensePlanes (refine_common.cuh)
    int level, layer, J, root, B;
};

template <int CPL, int NH>
__global__ void __launch_bounds__(DP_WARPS * 32)
dense_project_kernel(const DenseParams p) {
    constexpr int C = CPL * 32;
    constexpr int NOUT = 2 * NH + 9;
    constexpr int O_GATE = 2 * NH, O_VAL = 2 * NH + 3, O_CONF = 2 * NH + 6;
    const das_level_desc& d = p.lv->lv[p.level];
    const int HW = d.H * d.W, J = p.J;
    const long long total = static_cast<long long>(p.B) * HW;
    const int lane = threadIdx.x & 31;
    const int r = (lane >> 2) & 7, q = lane & 3;
    const long long n_groups = (total + 7) / 8;
    const float* __restrict__ F = d.feats[p.layer];
    for (long long g = static_cast<long long>(blockIdx.x) * DP_WARPS + (threadIdx.x >> 5); g < n_groups;
         g += static_cast<long long>(gridDim.x) * DP_WARPS) {
        const long long base = g * 8;
        Row<CPL> f[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {          // lane-permuted: row (k ^ r) goes to f[k] (see reduce8_permuted)
            const long long row = base + (k ^ r);
            const bool ok = row < total;
            f[k] = load_row<CPL>(F + (ok ? row : 0) * C, lane, ok);
        }
        const long long cell = base + r;             // the cell this lane post-processes
        const bool live = cell < total;
        const int b = live ? static_cast<int>(cell / HW) : 0;
        const int pix = live ? static_cast<int>(cell - static_cast<long long>(b) * HW) : 0;
        for (int j = 0; j < J; ++j) {
            const float* __restrict__ Wj = p.wpack + static_cast<size_t>(j) * NOUT * C;
            const float* __restrict__ Bj = p.wpack + static_cast<size_t>(J) * NOUT * C + j * NOUT;
            float res[NOUT];
#pragma unroll
            for (int o = 0; o < NOUT; ++o) {
                const Row<CPL> w = load_row<CPL>(Wj + o * C, lane, true);
                float acc[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[k] = dot_row<CPL>(f[k], w);
                res[o] = reduce8_permuted(acc) + __ldg(Bj + o);
            }
            if (!live) continue;
            const DensePlanes pl = dense_planes(p.proj, p.B, J, HW);
            const size_t cellj = (static_cast<size_t>(b) * J + j) * HW + pix;
            // the quad of lanes owning this cell shares the stores: lane q writes S[2q], S[2q+1] and dim q
            {
                float* outS = reinterpret_cast<float*>(q < 2 ? pl.s0 + cellj : pl.s1 + cellj) + 2 * (q & 1);
#pragma unroll
                for (int o = 0; o < 2 * NH; ++o)
                    if ((o >> 1) == q) outS[o & 1] = res[o];
            }
            if (q < 3) {
                const float rg = q == 0 ? res[O_GATE] : (q == 1 ? res[O_GATE + 1] : res[O_GATE + 2]);
                const float rn = q == 0 ? res[O_VAL] : (q == 1 ? res[O_VAL + 1] : res[O_VAL + 2]);
                const float rc = q == 0 ? res[O_CONF] : (q == 1 ? res[O_CONF + 1] : res[O_CONF + 2]);
                float prev;
                if (p.uvd_in) prev = __ldg(p.uvd_in + ((static_cast<size_t>(b) * J + j) * HW + pix) * 4 + q);
                else if (q == 2 && j == p.root) prev = 0.f;
                else prev = InMap(d.pose, p.lv->in_dtype)((static_cast<size_t>(b) * (3 + 6 * J) + 3 + 3 * j + q) * HW + pix) *
                            (q < 2 ? d.scale_uv : d.scale_d);
                const float gate = sigmoid_acc(rg);
                reinterpret_cast<float*>(pl.oa + cellj)[q] = __fadd_rn(__fmul_rn(1.0f - gate, prev), __fmul_rn(gate, rn));  // blended offset
                if (q == 0) reinterpret_cast<float*>(pl.oa + cellj)[3] = rc;             // confidence logits: x next to O, y / z apart
                else reinterpret_cast<float*>(pl.cb + cellj)[q - 1] = rc;
            }
        }
    }
}

template <int NH>
__global__ void __launch_bounds__(256, 3)
dense_sample_kernel(const DenseParams p) {
    static_assert(NH == 4, "plane layout below is for 2*NH = 8 sampling offsets");
    // Four joint-major planes (refine_common.cuh: DensePlanes) of 16-byte (8-byte) entries: a warp = 32 consecutive cells of
    // one joint, whose bilinear taps are (nearly) consecutive entries, so a 128-bit load per lane fills whole 128-byte
    // lines -- with the earlier 32-byte records every L1 wavefront carried half a line (ncu: l1tex 88 % busy, 2.5 ms).
    const das_level_desc& d = p.lv->lv[p.level];
    const int H = d.H, W = d.W, HW = H * W, J = p.J;
    const float fW = static_cast<float>(W), fH = static_cast<float>(H);
    const long long total = static_cast<long long>(p.B) * HW * J;
    const DensePlanes pl = dense_planes(p.proj, p.B, J, HW);
    for (long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; t < total;
         t += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int pix = static_cast<int>(t % HW);
        const long long bj = t / HW;                 // b * J + j
        const int y = pix / W, x = pix - y * W;
        const float4* __restrict__ pS0 = pl.s0 + static_cast<size_t>(bj) * HW;
        const float4* __restrict__ pS1 = pl.s1 + static_cast<size_t>(bj) * HW;
        const float4* __restrict__ pOA = pl.oa + static_cast<size_t>(bj) * HW;
        const float2* __restrict__ pCB = pl.cb + static_cast<size_t>(bj) * HW;
        const float4 s0 = __ldg(pS0 + pix), s1 = __ldg(pS1 + pix), om = __ldg(pOA + pix);
        const float ox = om.x, oy = om.y;
        float hx[2 * NH], hy[2 * NH];
        {
            const Corner ct = make_corner(sample_coord(x, ox, fW), sample_coord(y, oy, fH), W, H);
            float s[2 * NH];
#pragma unroll
            for (int o = 0; o < 2 * NH; ++o) s[o] = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (!corner_ok(ct, k, W, H)) continue;
                const float wk = corner_wgt(ct, k);
                const int cp = corner_pix(ct, k, W);
                const float4 a0 = __ldg(pS0 + cp), a1 = __ldg(pS1 + cp);
                s[0] = fmaf(a0.x, wk, s[0]); s[1] = fmaf(a0.y, wk, s[1]);
                s[2] = fmaf(a0.z, wk, s[2]); s[3] = fmaf(a0.w, wk, s[3]);
                s[4] = fmaf(a1.x, wk, s[4]); s[5] = fmaf(a1.y, wk, s[5]);
                s[6] = fmaf(a1.z, wk, s[6]); s[7] = fmaf(a1.w, wk, s[7]);
            }
            const float sm[2 * NH] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
            for (int h = 0; h < NH; ++h) {
                hx[h] = s[2 * h] + ox;
                hy[h] = s[2 * h + 1] + oy;
                hx[NH + h] = sm[2 * h];
                hy[NH + h] = sm[2 * h + 1];
            }
        }
        float hv[2 * NH][3], hc[2 * NH][3];
#pragma unroll
        for (int h = 0; h < 2 * NH; ++h) {
            const Corner ch = make_corner(sample_coord(x, hx[h], fW), sample_coord(y, hy[h], fH), W, H);
            float v[3] = {0.f, 0.f, 0.f}, cf[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (!corner_ok(ch, k, W, H)) continue;
                const float wk = corner_wgt(ch, k);
                const int cp = corner_pix(ch, k, W);
                const float4 oo = __ldg(pOA + cp);                 // {O.x, O.y, O.z, cf.x}
                const float2 cc = __ldg(pCB + cp);                 // {cf.y, cf.z}
                v[0] = fmaf(oo.x, wk, v[0]); v[1] = fmaf(oo.y, wk, v[1]); v[2] = fmaf(oo.z, wk, v[2]);
                cf[0] = fmaf(oo.w, wk, cf[0]); cf[1] = fmaf(cc.x, wk, cf[1]); cf[2] = fmaf(cc.y, wk, cf[2]);
            }
            hv[h][0] = v[0] + hx[h];
            hv[h][1] = v[1] + hy[h];
            hv[h][2] = v[2];
            hc[h][0] = cf[0]; hc[h][1] = cf[1]; hc[h][2] = cf[2];
        }
        float res[3];
#pragma unroll
        for (int e = 0; e < 3; ++e) {
            float m = hc[0][e];
#pragma unroll
            for (int h = 1; h < 2 * NH; ++h) m = fmaxf(m, hc[h][e]);
            float ex[2 * NH], se = 0.f;
#pragma unroll
            for (int h = 0; h < 2 * NH; ++h) { ex[h] = expf(hc[h][e] - m); se += ex[h]; }
            const float inv = __frcp_rn(se);
            float o = 0.f;
#pragma unroll
            for (int h = 0; h < 2 * NH; ++h) o = fmaf(hv[h][e], ex[h] * inv, o);
            res[e] = o;
        }
        reinterpret_cast<float4*>(p.uvd_out)[static_cast<size_t>(bj) * HW + pix] = make_float4(res[0], res[1], res[2], 0.f);
    }
}

// Second version of the sampling kernel; its body is dense_sample_cell (refine_common.cuh), which the sparse last layer also
// calls for the cells it needs (on-demand sampling of the last dense layer).  Same arithmetic and accumulation order as the
// first version -> bit-identical results; FASTEXP
// swaps the softmax's expf for ex2.approx): ncu showed the first one ISSUE-bound next to its L1 load (1 680 instructions per
// (cell, joint), issue 64 % busy, l1tex 86 %), so this one removes instructions:
//   * the 18 divisions of the coordinate chains use the precomputed reciprocal (div_by: 3 instead of ~10 instructions);
//   * a head whose 4 corners are all inside the map (everywhere but at the border) takes a branch-free path: one cell
//     index, the other three are +1, +W, +W+1 as immediate offsets, no per-corner predicates;
//   * the 6 interpolation FMAs per corner are 3 packed fma.rn.f32x2 on the register pairs the 128-bit loads deliver.
template <int NH, bool FASTEXP>
__global__ void __launch_bounds__(256, 3)
dense_sample2_kernel(const DenseParams p) {
    const das_level_desc& d = p.lv->lv[p.level];
    const int H = d.H, W = d.W, HW = H * W, J = p.J;
    const long long total = static_cast<long long>(p.B) * HW * J;
    const DensePlanes pl = dense_planes(p.proj, p.B, J, HW);
    for (long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; t < total;
         t += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int pix = static_cast<int>(t % HW);
        const size_t bj = static_cast<size_t>(t / HW);   // b * J + j
        const float3 r = dense_sample_cell<NH, FASTEXP>(pl.s0 + bj * HW, pl.s1 + bj * HW, pl.oa + bj * HW, pl.cb + bj * HW, pix, W, H);
        reinterpret_cast<float4*>(p.uvd_out)[bj * HW + pix] = make_float4(r.x, r.y, r.z, 0.f);
    }
}

}  // namespace das

extern "C" int das_dense_project_tc(const das_levels* d_levels, const das_levels* h_levels, int32_t level, int32_t layer,
                                    const das_decode_cfg* cfg, const float* weights, const void* panels,
                                    const float* uvd_in, float* proj, void* stream);

extern "C" int das_refine_dense_layer(const das_levels* d_levels, const das_levels* h_levels, int32_t level,
                                      int32_t layer, const das_decode_cfg* cfg, const float* weights,
                                      const void* tc_panels, const float* uvd_in, float* uvd_out, float* proj,
                                      void* stream) {
    using namespace das;
    DAS_REQUIRE(d_levels && h_levels && cfg && weights && uvd_out && proj, DAS_ERR_ARG, "das_refine_dense_layer: null pointer");
    DAS_REQUIRE(level >= 0 && level < h_levels->n_levels, DAS_ERR_ARG, "level=%d", level);
    DAS_REQUIRE(layer >= 0 && layer < cfg->num_layers, DAS_ERR_ARG, "layer=%d", layer);
    DAS_REQUIRE(cfg->num_heads == 4, DAS_ERR_UNSUPPORTED, "num_heads=%d: only 4 is built", cfg->num_heads);
    DAS_REQUIRE(cfg->num_joints >= 1 && cfg->num_joints <= DAS_MAX_JOINTS, DAS_ERR_CAPACITY, "num_joints=%d", cfg->num_joints);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    DenseParams p{};
    p.lv = d_levels; p.wpack = weights; p.uvd_in = uvd_in; p.uvd_out = uvd_out; p.proj = proj;
    p.level = level; p.layer = layer; p.J = cfg->num_joints; p.root = cfg->root_idx; p.B = h_levels->batch;
    const long long cells = static_cast<long long>(h_levels->batch) * h_levels->lv[level].H * h_levels->lv[level].W;
    const int grid1 = static_cast<int>(std::min<long long>((cells / 8 + DP_WARPS) / DP_WARPS, 4LL * kSMs));
    if (tc_panels && cfg->feat_channels == 256) {
        const int st_ = das_dense_project_tc(d_levels, h_levels, level, layer, cfg, weights, tc_panels, uvd_in, proj, stream);
        if (st_ != DAS_OK) return st_;
    } else
    switch (cfg->feat_channels) {
        case 128: dense_project_kernel<4, 4><<<grid1, DP_WARPS * 32, 0, st>>>(p); break;
        case 256: dense_project_kernel<8, 4><<<grid1, DP_WARPS * 32, 0, st>>>(p); break;
        case 512: dense_project_kernel<16, 4><<<grid1, DP_WARPS * 32, 0, st>>>(p); break;
        default:
            set_error("feat_channels=%d: only 128/256/512 are built", cfg->feat_channels);
            return DAS_ERR_UNSUPPORTED;
    }
    DAS_CUDA_CHECK(cudaGetLastError());
    const long long threads = cells * cfg->num_joints;
    const int grid2 = static_cast<int>(std::min<long long>((threads + 255) / 256, 16LL * kSMs));
    // DAS_SAMPLE_VARIANT: 0 = first version, 1 = dense_sample2 with the accurate expf (default), 2 = with ex2.approx
    static const int variant = std::getenv("DAS_SAMPLE_VARIANT") ? std::atoi(std::getenv("DAS_SAMPLE_VARIANT")) : 1;
    if (variant == 0) dense_sample_kernel<4><<<grid2, 256, 0, st>>>(p);
    else if (variant == 2) dense_sample2_kernel<4, true><<<grid2, 256, 0, st>>>(p);
    else dense_sample2_kernel<4, false><<<grid2, 256, 0, st>>>(p);
    DAS_CUDA_CHECK(cudaGetLastError());
    return DAS_OK;
}
