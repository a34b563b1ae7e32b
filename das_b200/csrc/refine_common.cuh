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

template <int CPL>
__device__ __forceinline__ float dot_row(const Row<CPL>& f, const Row<CPL>& w) {
    float a = 0.f;
#pragma unroll
    for (int q = 0; q < CPL / 4; ++q) {
        a = fmaf(f.v[q].x, w.v[q].x, a);
        a = fmaf(f.v[q].y, w.v[q].y, a);
        a = fmaf(f.v[q].z, w.v[q].z, a);
        a = fmaf(f.v[q].w, w.v[q].w, a);
    }
    return a;
}

__device__ __forceinline__ float warp_allsum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

// Sum a[r] over the 32 lanes for 8 rows at once; lane L returns the total of row (L >> 2) & 7.
__device__ __forceinline__ float reduce8_transposed(const float (&a)[8], int lane) {
    const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
    float b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float send = h16 ? a[i] : a[i + 4];
        const float keep = h16 ? a[i + 4] : a[i];
        b[i] = keep + __shfl_xor_sync(FULL, send, 16);
    }
    float c[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float send = h8 ? b[i] : b[i + 2];
        const float keep = h8 ? b[i + 2] : b[i];
        c[i] = keep + __shfl_xor_sync(FULL, send, 8);
    }
    const float send = h4 ? c[0] : c[1];
    const float keep = h4 ? c[1] : c[0];
    float d = keep + __shfl_xor_sync(FULL, send, 4);
    d += __shfl_xor_sync(FULL, d, 2);
    d += __shfl_xor_sync(FULL, d, 1);
    return d;
}

// Index-space sample coordinate of `cell + 0.5 + off` after the reference's normalise ->
// grid_sample un-normalise chain (recursive_update.py:52-54, ATen align_corners=False).
__device__ __forceinline__ float sample_coord(int cell, float off, float size) {
    const float loc = __fdiv_rn(__fadd_rn(static_cast<float>(cell) + 0.5f, off), size);
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

}  // namespace das
