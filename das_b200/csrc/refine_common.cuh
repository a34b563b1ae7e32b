// Device helpers shared by the sparse (refine_sparse.cu) and dense (refine_dense.cu) refinement kernels.
#pragma once

#include "das_common.cuh"

namespace das {

constexpr unsigned FULL = 0xffffffffu;

template <int CPL>
struct Row {
    float4 v[CPL / 4];
};

template <int CPL>
__device__ __forceinline__ Row<CPL> load_row(const float* __restrict__ base, int lane, bool ok) {
    Row<CPL> r;
#pragma unroll
    for (int q = 0; q < CPL / 4; ++q) {
        r.v[q] = ok ? __ldg(reinterpret_cast<const float4*>(base + q * 128 + 4 * lane)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    return r;
}

// the same lane slice of a row staged in shared memory (conflict-free: consecutive lanes read consecutive 16-byte chunks)
template <int CPL>
__device__ __forceinline__ Row<CPL> load_row_smem(const float* base, int lane) {
    Row<CPL> r;
#pragma unroll
    for (int q = 0; q < CPL / 4; ++q) r.v[q] = *reinterpret_cast<const float4*>(base + q * 128 + 4 * lane);
    return r;
}

// Packed fp32x2 FMA (sm_100a: one FFMA2 issue slot does two fp32 FMAs). d = a * b + c, lane-wise.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long ra, rb, rc, rd;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    float2 d;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
    return d;
}

template <int CPL>
__device__ __forceinline__ float dot_row(const Row<CPL>& f, const Row<CPL>& w) {
    float2 a = make_float2(0.f, 0.f);
#pragma unroll
    for (int q = 0; q < CPL / 4; ++q) {
        a = ffma2(make_float2(f.v[q].x, f.v[q].y), make_float2(w.v[q].x, w.v[q].y), a);
        a = ffma2(make_float2(f.v[q].z, f.v[q].w), make_float2(w.v[q].z, w.v[q].w), a);
    }
    return a.x + a.y;
}

__device__ __forceinline__ float warp_allsum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

// Sum over the 32 lanes for 8 rows at once.  PRECONDITION: lane L accumulated row (k ^ r(L)) into a[k], with
// r(L) = (L >> 2) & 7 (the callers load their rows in that lane-permuted order), which makes every exchange of
// the transposing butterfly static -- no per-lane selects.  Lane L returns the total of row r(L).
__device__ __forceinline__ float reduce8_permuted(const float (&a)[8]) {
    float b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = a[i] + __shfl_xor_sync(FULL, a[i + 4], 16);
    float c[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) c[i] = b[i] + __shfl_xor_sync(FULL, b[i + 2], 8);
    float d = c[0] + __shfl_xor_sync(FULL, c[1], 4);
    d += __shfl_xor_sync(FULL, d, 2);
    d += __shfl_xor_sync(FULL, d, 1);
    return d;
}

// 4-row variant: lane L accumulated row (k ^ c(L)) into a[k], c(L) = (L >> 3) & 3; returns the total of row c(L).
__device__ __forceinline__ float reduce4_permuted(const float (&a)[4]) {
    const float b0 = a[0] + __shfl_xor_sync(FULL, a[2], 16);
    const float b1 = a[1] + __shfl_xor_sync(FULL, a[3], 16);
    float c = b0 + __shfl_xor_sync(FULL, b1, 8);
    c += __shfl_xor_sync(FULL, c, 4);
    c += __shfl_xor_sync(FULL, c, 2);
    c += __shfl_xor_sync(FULL, c, 1);
    return c;
}

// 2-row variant: lane L accumulated row (k ^ h(L)) into a[k], h(L) = (L >> 4) & 1; returns the total of row h(L).
__device__ __forceinline__ float reduce2_permuted(const float (&a)[2]) {
    float b = a[0] + __shfl_xor_sync(FULL, a[1], 16);
    b += __shfl_xor_sync(FULL, b, 8);
    b += __shfl_xor_sync(FULL, b, 4);
    b += __shfl_xor_sync(FULL, b, 2);
    b += __shfl_xor_sync(FULL, b, 1);
    return b;
}

// Projection scratch of a dense layer: four joint-major planes [B][J][HW] inside one allocation of 16 floats per
// (cell, joint).  Entries are 16 (8) bytes so that a warp of 32 consecutive cells reads whole 128-byte lines per tap.
//   s0 = {S head0.x, head0.y, head1.x, head1.y}   s1 = {S head2.x, head2.y, head3.x, head3.y}   (sampling offsets)
//   oa = {blended O.u, O.v, O.d, conf.u}          cb = {conf.v, conf.d}
struct DensePlanes {
    float4* s0;
    float4* s1;
    float4* oa;
    float2* cb;
};
__host__ __device__ __forceinline__ DensePlanes dense_planes(float* proj, int B, int J, int HW) {
    const size_t n = static_cast<size_t>(B) * J * HW;
    DensePlanes pl;
    pl.s0 = reinterpret_cast<float4*>(proj);
    pl.s1 = pl.s0 + n;
    pl.oa = pl.s1 + n;
    pl.cb = reinterpret_cast<float2*>(pl.oa + n);
    return pl;
}

// Index-space sample coordinate of `cell + 0.5 + off` after the reference's normalise ->
// grid_sample un-normalise chain (recursive_update.py:52-54, ATen align_corners=False).
__device__ __forceinline__ float sample_coord(int cell, float off, float size) {
    const float loc = __fdiv_rn(__fadd_rn(static_cast<float>(cell) + 0.5f, off), size);
    const float g = __fadd_rn(__fmul_rn(2.0f, loc), -1.0f);
    return __fadd_rn(__fmul_rn(__fadd_rn(g, 1.0f), size * 0.5f), -0.5f);
}

// a / size with the reciprocal precomputed (rsize = __frcp_rn(size)): q0 = a * rsize, one exact-remainder step, one
// correction -- Markstein's sequence, which returns the correctly rounded quotient (checked against IEEE division for every
// map size up to 600 and 240 k numerators each, tools/check_div.py), so the coordinate chain stays bit-identical to
// sample_coord() at 3 instructions per division instead of the ~10 of __fdiv_rn.
__device__ __forceinline__ float div_by(float a, float size, float rsize) {
    const float q0 = __fmul_rn(a, rsize);
    const float r = __fmaf_rn(-q0, size, a);
    return __fmaf_rn(r, rsize, q0);
}
__device__ __forceinline__ float sample_coord(int cell, float off, float size, float rsize) {
    const float loc = div_by(__fadd_rn(static_cast<float>(cell) + 0.5f, off), size, rsize);
    const float g = __fadd_rn(__fmul_rn(2.0f, loc), -1.0f);
    return __fadd_rn(__fmul_rn(__fadd_rn(g, 1.0f), size * 0.5f), -0.5f);
}

struct Corner {
    int x0, y0;        // north-west corner (clamped to a safe int range)
    float w, n;        // distance to the west / north side
};

__device__ __forceinline__ Corner make_corner(float ix, float iy, int W, int H) {
    Corner c;
    const float fx = floorf(ix), fy = floorf(iy);
    c.w = ix - fx;
    c.n = iy - fy;
    // clamp before the int conversion; anything outside [-1, size] has no in-bounds corner anyway
    c.x0 = static_cast<int>(fminf(fmaxf(fx, -2.0f), static_cast<float>(W) + 1.0f));
    c.y0 = static_cast<int>(fminf(fmaxf(fy, -2.0f), static_cast<float>(H) + 1.0f));
    if (!(ix == ix) || !(iy == iy)) { c.x0 = -2; c.y0 = -2; c.w = 0.f; c.n = 0.f; }  // NaN -> nothing sampled
    return c;
}

__device__ __forceinline__ bool corner_ok(const Corner& c, int k, int W, int H) {
    const int x = c.x0 + (k & 1), y = c.y0 + (k >> 1);
    return x >= 0 && x < W && y >= 0 && y < H;
}
__device__ __forceinline__ int corner_pix(const Corner& c, int k, int W) { return (c.y0 + (k >> 1)) * W + c.x0 + (k & 1); }
__device__ __forceinline__ float corner_wgt(const Corner& c, int k) {
    // ATen: nw = s*e, ne = s*w, sw = n*e, se = n*w with e = 1-w, s = 1-n
    const float wx = (k & 1) ? c.w : (1.0f - c.w);
    const float wy = (k >> 1) ? c.n : (1.0f - c.n);
    return wy * wx;
}

// One (joint, cell) of a dense layer's progressive sampling (recursive_update.py:34-82, 9-31) from the layer's projection
// planes: 4 bilinear taps of the sampling offsets at t = p + O.xy, 8 heads x 4 taps of {O, conf}, softmax over the heads.
// Shared by dense_sample2_kernel (every cell of the map) and the sparse last layer's on-demand evaluation (refine_sparse.cu:
// only the cells the selected candidates sample), so both produce the same bits.  pS0 / pS1 / pOA / pCB point at the
// (image, joint) slice of the four planes.  The 18 divisions of the coordinate chains use the precomputed reciprocal
// (div_by); a head whose 4 corners are all inside the map takes a branch-free path (one cell index, +1, +W, +W+1); the 6
// interpolation FMAs per corner are 3 packed fma.rn.f32x2.
template <int NH, bool FASTEXP>
__device__ __forceinline__ float3 dense_sample_cell(const float4* __restrict__ pS0, const float4* __restrict__ pS1,
                                                    const float4* __restrict__ pOA, const float2* __restrict__ pCB,
                                                    int pix, int W, int H) {
    static_assert(NH == 4, "plane layout is for 2*NH = 8 sampling offsets");
    const float fW = static_cast<float>(W), fH = static_cast<float>(H);
    const float rW = __frcp_rn(fW), rH = __frcp_rn(fH);
    const int y = pix / W, x = pix - y * W;
    const float4 s0 = __ldg(pS0 + pix), s1 = __ldg(pS1 + pix), om = __ldg(pOA + pix);
    const float ox = om.x, oy = om.y;
    float hx[2 * NH], hy[2 * NH];
    {
        const float ix = sample_coord(x, ox, fW, rW), iy = sample_coord(y, oy, fH, rH);
        const float fx = floorf(ix), fy = floorf(iy);
        const float ww = ix - fx, wn = iy - fy;
        float2 a01 = make_float2(0.f, 0.f), a23 = a01, a45 = a01, a67 = a01;
        if (fx >= 0.f && fx <= fW - 2.f && fy >= 0.f && fy <= fH - 2.f) {
            const int cp = static_cast<int>(fy) * W + static_cast<int>(fx);
            const float we = 1.0f - ww, ws = 1.0f - wn;
            const float wk4[4] = {ws * we, ws * ww, wn * we, wn * ww};
            const int off4[4] = {0, 1, W, W + 1};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float4 b0 = __ldg(pS0 + cp + off4[k]), b1 = __ldg(pS1 + cp + off4[k]);
                const float2 wk2 = make_float2(wk4[k], wk4[k]);
                a01 = ffma2(make_float2(b0.x, b0.y), wk2, a01); a23 = ffma2(make_float2(b0.z, b0.w), wk2, a23);
                a45 = ffma2(make_float2(b1.x, b1.y), wk2, a45); a67 = ffma2(make_float2(b1.z, b1.w), wk2, a67);
            }
        } else {
            const Corner ct = make_corner(ix, iy, W, H);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (!corner_ok(ct, k, W, H)) continue;
                const float wk = corner_wgt(ct, k);
                const int cp = corner_pix(ct, k, W);
                const float4 b0 = __ldg(pS0 + cp), b1 = __ldg(pS1 + cp);
                const float2 wk2 = make_float2(wk, wk);
                a01 = ffma2(make_float2(b0.x, b0.y), wk2, a01); a23 = ffma2(make_float2(b0.z, b0.w), wk2, a23);
                a45 = ffma2(make_float2(b1.x, b1.y), wk2, a45); a67 = ffma2(make_float2(b1.z, b1.w), wk2, a67);
            }
        }
        hx[0] = a01.x + ox; hy[0] = a01.y + oy;
        hx[1] = a23.x + ox; hy[1] = a23.y + oy;
        hx[2] = a45.x + ox; hy[2] = a45.y + oy;
        hx[3] = a67.x + ox; hy[3] = a67.y + oy;
        hx[4] = s0.x; hy[4] = s0.y; hx[5] = s0.z; hy[5] = s0.w;
        hx[6] = s1.x; hy[6] = s1.y; hx[7] = s1.z; hy[7] = s1.w;
    }
    float hv[2 * NH][3], hc[2 * NH][3];
#pragma unroll
    for (int h = 0; h < 2 * NH; ++h) {
        const float ix = sample_coord(x, hx[h], fW, rW), iy = sample_coord(y, hy[h], fH, rH);
        const float fx = floorf(ix), fy = floorf(iy);
        const float ww = ix - fx, wn = iy - fy;
        float2 v01 = make_float2(0.f, 0.f), v2c = v01, c12 = v01;      // {O.x, O.y}, {O.z, cf.x}, {cf.y, cf.z}
        if (fx >= 0.f && fx <= fW - 2.f && fy >= 0.f && fy <= fH - 2.f) {
            const int cp = static_cast<int>(fy) * W + static_cast<int>(fx);
            const float we = 1.0f - ww, ws = 1.0f - wn;
            const float wk4[4] = {ws * we, ws * ww, wn * we, wn * ww};
            const int off4[4] = {0, 1, W, W + 1};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float4 oo = __ldg(pOA + cp + off4[k]);
                const float2 cc = __ldg(pCB + cp + off4[k]);
                const float2 wk2 = make_float2(wk4[k], wk4[k]);
                v01 = ffma2(make_float2(oo.x, oo.y), wk2, v01);
                v2c = ffma2(make_float2(oo.z, oo.w), wk2, v2c);
                c12 = ffma2(cc, wk2, c12);
            }
        } else {
            const Corner ch = make_corner(ix, iy, W, H);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (!corner_ok(ch, k, W, H)) continue;
                const float wk = corner_wgt(ch, k);
                const int cp = corner_pix(ch, k, W);
                const float4 oo = __ldg(pOA + cp);
                const float2 cc = __ldg(pCB + cp);
                const float2 wk2 = make_float2(wk, wk);
                v01 = ffma2(make_float2(oo.x, oo.y), wk2, v01);
                v2c = ffma2(make_float2(oo.z, oo.w), wk2, v2c);
                c12 = ffma2(cc, wk2, c12);
            }
        }
        hv[h][0] = v01.x + hx[h];
        hv[h][1] = v01.y + hy[h];
        hv[h][2] = v2c.x;
        hc[h][0] = v2c.y; hc[h][1] = c12.x; hc[h][2] = c12.y;
    }
    float res[3];
#pragma unroll
    for (int e = 0; e < 3; ++e) {
        float m = hc[0][e];
#pragma unroll
        for (int h = 1; h < 2 * NH; ++h) m = fmaxf(m, hc[h][e]);
        float ex[2 * NH], se = 0.f;
#pragma unroll
        for (int h = 0; h < 2 * NH; ++h) { ex[h] = FASTEXP ? __expf(hc[h][e] - m) : expf(hc[h][e] - m); se += ex[h]; }
        const float inv = __frcp_rn(se);
        float o = 0.f;
#pragma unroll
        for (int h = 0; h < 2 * NH; ++h) o = fmaf(hv[h][e], ex[h] * inv, o);
        res[e] = o;
    }
    return make_float3(res[0], res[1], res[2]);
}

}  // namespace das
