// Dense 1x1 projection of a RecursiveUpdateLayer on the tensor cores (layers 1..L-1 when num_layers > 1):
//   F[B*H*W, 256] (NHWC feature rows) x W[17*J, 256]^T  ->  proj planes S[B][J][H*W][8] and OC[B][J][H*W][8] = {blended O 3, conf 3, -, -}
// Reference: recursive_update.py:186-197 (the four nn.Conv2d(C, ., 1) of NextLevelOffset + the gated blend).
// This is the one genuinely dense contraction of the path (2*256*17J flops per cell, 3.5-4.7 GFLOP per image and
// layer); everything else of the layer is the gather-bound dense_sample_kernel.
//
// 3xTF32 on tcgen05 (fp32-level accuracy), same machinery as refine_tc.cu: 128-cell tiles, A operand staged
// through TMEM by row-owning producer groups, fused MMA bursts, double-buffered accumulators drained by two
// epilogue groups.  The weights of all joints (17J x 256 x {hi, lo} = 0.6 MB) do not fit in shared memory, so a
// CTA keeps the panel of ONE joint group (3 joints = 51 outputs, padded to 64; 128 KB with the lo half) resident
// for its whole life and walks the cell tiles; CTA c serves joint group c % Q, and the Q CTAs c, c+1, .. work on
// the same cell tile at the same time, so the 128-KB feature tile is fetched from HBM once and re-read from L2.
#include <algorithm>

#include "refine_common.cuh"
#include "tc_common.cuh"

namespace das {

constexpr int DT_JG = 3;                         // joints per CTA
constexpr int DT_N = 64;                         // MMA N per half (17 * 3 = 51 real rows)
constexpr int DT_KB = 8;
constexpr int DT_C = 256;
constexpr int DT_NH = 4;
constexpr int DT_NOUT = 2 * DT_NH + 9;           // 17
constexpr int DT_PANEL_KB = 2 * DT_N * 128;      // bytes per k-block: 64 hi rows then 64 lo rows
constexpr int DT_PANEL = DT_KB * DT_PANEL_KB;    // 128 KB
constexpr int DT_EG = 2;                         // epilogue groups / accumulator buffers
constexpr int DT_PG = 3;                         // producer groups
constexpr int DT_STAGES = 2;                     // smem stages per producer group
constexpr int DT_SLOTS = 4;                      // TMEM A slots (64 columns: 32 hi + 32 lo)
constexpr int DT_A_BYTES = 128 * 128;
constexpr int DT_FIRST_PRODUCER = 4 * DT_EG;     // warp 8
constexpr int DT_MMA_WARP = DT_FIRST_PRODUCER + 4 * DT_PG;   // warp 20
constexpr int DT_THREADS = 32 * (DT_MMA_WARP + 1);
constexpr int DT_D_COL = 0;                      // accumulators: 2 x 128 columns
constexpr int DT_A_COL = 2 * DT_N * DT_EG;       // 256: A ring 4 x 64 columns
constexpr int DT_SMEM = 1024 + DT_PG * DT_STAGES * DT_A_BYTES + DT_PANEL;

struct DenseTcParams {
    const das_levels* lv;
    const float* wpack;            // [J][17][C] + biases
    const unsigned char* panels;   // [Q][DT_PANEL]
    const float* uvd_in;           // nullptr -> scaled raw uvd from lv.pose; else joint-major [B][J][HW][4]
    float* proj;                   // two joint-major planes: S [B][J][HW][8], then OC [B][J][HW][8] = {O 3, conf 3, -, -}
    int level, layer, J, root, B, Q;
};

__global__ void __launch_bounds__(DT_THREADS, 1)
dense_project_tc_kernel(const DenseTcParams p) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    unsigned char* sA = base;                                      // [group][stage] k-block tiles (128-B swizzle)
    unsigned char* sB = sA + DT_PG * DT_STAGES * DT_A_BYTES;       // resident panel of this CTA's joint group
    __shared__ uint64_t a_full[DT_SLOTS], a_empty[DT_SLOTS], acc_full[DT_EG], acc_free[DT_EG];
    __shared__ uint32_t tmem_base;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const das_level_desc& d = p.lv->lv[p.level];
    const int HW = d.H * d.W, J = p.J;
    const long long cells = static_cast<long long>(p.B) * HW;
    const int n_tiles = static_cast<int>((cells + 127) / 128);
    const int Q = p.Q;
    const int G = gridDim.x / Q;                   // CTAs per joint group
    const int q = blockIdx.x % Q, m = blockIdx.x / Q;
    if (m >= G) return;
    const int my_tiles = (n_tiles - m + G - 1) / G;          // tiles m, m+G, ...
    if (my_tiles <= 0) return;
    const int total_kb = my_tiles * DT_KB;
    const float* __restrict__ F = d.feats[p.layer];

    if (tid == 0) {
        for (int s = 0; s < DT_SLOTS; ++s) { tc::mbar_init(&a_full[s], 4); tc::mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < DT_EG; ++s) { tc::mbar_init(&acc_full[s], 1); tc::mbar_init(&acc_free[s], 4); }
        tc::mbar_fence_init();
    }
    if (warp == 0) tc::tmem_alloc(&tmem_base, 512);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem0 = tmem_base;
    const uint32_t sB_u = tc::smem_u32(sB);

    if (warp == DT_MMA_WARP) {
        // ===== MMA issuer warp: loads this joint group's panel once, then one burst per k-block =====================
        {
            const unsigned char* src = p.panels + static_cast<size_t>(q) * DT_PANEL;
            for (int c = lane; c < DT_PANEL / 16; c += 32) tc::cp_async16(sB_u + c * 16, src + c * 16, true);
            tc::cp_async_commit();
            tc::cp_async_wait<0>();
            tc::fence_proxy_async();
            __syncwarp();
        }
        constexpr uint32_t idesc_hi = tc::instr_desc_tf32(128, 2 * DT_N);   // A_hi x [B_hi ; B_lo] -> D[:, 0:128]
        constexpr uint32_t idesc_lo = tc::instr_desc_tf32(128, DT_N);       // A_lo x  B_hi         -> D[:, 0:64]
        int g = 0;
        for (int i = 0; i < my_tiles; ++i) {
            const uint32_t dcol = tmem0 + DT_D_COL + (i % DT_EG) * (2 * DT_N);
            if (i >= DT_EG) tc::mbar_wait(&acc_free[i % DT_EG], ((i / DT_EG) - 1) & 1);
            for (int kb = 0; kb < DT_KB; ++kb, ++g) {
                const int slot = g % DT_SLOTS;
                tc::mbar_wait(&a_full[slot], (g / DT_SLOTS) & 1);
                tc::tc_fence_after();
                const uint32_t a_hi = tmem0 + DT_A_COL + slot * 64;
                tc::umma_kblock_3xtf32_ts(dcol, a_hi, a_hi + 32, tc::smem_desc_sw128(sB_u + kb * DT_PANEL_KB), idesc_hi, idesc_lo, kb != 0);
                tc::umma_commit_elect(&a_empty[slot]);
                if (kb == DT_KB - 1) tc::umma_commit_elect(&acc_full[i % DT_EG]);
            }
        }
    } else if (warp >= DT_FIRST_PRODUCER) {
        // ===== producer groups: contiguous 128-cell k-block -> smem (cp.async) -> own row -> hi/lo -> TMEM ===========
        const int pg = (warp - DT_FIRST_PRODUCER) >> 2;
        const int qw = warp & 3;
        const int gt = (qw << 5) | lane;
        const uint32_t sG_u = tc::smem_u32(sA) + pg * DT_STAGES * DT_A_BYTES;
        unsigned char* sG = sA + pg * DT_STAGES * DT_A_BYTES;
        const int bar_id = 1 + pg;
        const int my_n = (total_kb - pg + DT_PG - 1) / DT_PG;
        auto gather = [&](int n) {
            const int g = pg + n * DT_PG;
            const int i = g / DT_KB, kb = g - i * DT_KB, st = n % DT_STAGES;
            const long long cell0 = static_cast<long long>(m + static_cast<long long>(i) * G) * 128;
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int row = (gt >> 3) + 16 * it, ch = gt & 7;
                const bool ok = cell0 + row < cells;
                tc::cp_async16(sG_u + st * DT_A_BYTES + tc::swz128(row, ch), F + (ok ? (cell0 + row) : 0) * DT_C + kb * 32 + ch * 4, ok);
            }
        };
        for (int n = 0; n < DT_STAGES - 1; ++n) { if (n < my_n) gather(n); tc::cp_async_commit(); }
        for (int n = 0; n < my_n; ++n) {
            const int g = pg + n * DT_PG;
            const int slot = g % DT_SLOTS, st = n % DT_STAGES;
            tc::cp_async_wait<DT_STAGES - 2>();
            asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
            float hi[32];
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) {
                const float4 a = *reinterpret_cast<const float4*>(sG + st * DT_A_BYTES + tc::swz128(gt, ch));
                hi[4 * ch] = a.x; hi[4 * ch + 1] = a.y; hi[4 * ch + 2] = a.z; hi[4 * ch + 3] = a.w;
            }
            asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
            if (n + DT_STAGES - 1 < my_n) gather(n + DT_STAGES - 1);
            tc::cp_async_commit();
            if (g >= DT_SLOTS) tc::mbar_wait(&a_empty[slot], ((g / DT_SLOTS) - 1) & 1);
            tc::tc_fence_after();
            const uint32_t taddr = tmem0 + (static_cast<uint32_t>(qw * 32) << 16) + DT_A_COL + slot * 64;
            tc::tmem_st32(taddr, hi);
#pragma unroll
            for (int e = 0; e < 32; ++e) hi[e] = hi[e] - tc::tf32_hi(hi[e]);
            tc::tmem_st32(taddr + 32, hi);
            tc::tmem_st_wait();
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&a_full[slot]);
        }
        tc::cp_async_wait<0>();
    } else {
        // ===== epilogue groups: thread = cell; bias, gate, blend with the previous offset, 64-B records ==============
        const int e = warp >> 2;
        const int qw = warp & 3;
        const int row = (qw << 5) | lane;
        const float* __restrict__ Bias = p.wpack + static_cast<size_t>(J) * DT_NOUT * DT_C;
        for (int i = e; i < my_tiles; i += DT_EG) {
            const long long cell = (static_cast<long long>(m) + static_cast<long long>(i) * G) * 128 + row;
            const bool live = cell < cells;
            const int b = live ? static_cast<int>(cell / HW) : 0;
            const int pix = live ? static_cast<int>(cell - static_cast<long long>(b) * HW) : 0;
            tc::mbar_wait(&acc_full[e], (i / DT_EG) & 1);
            tc::tc_fence_after();
            const uint32_t tbase = tmem0 + (static_cast<uint32_t>(qw * 32) << 16) + DT_D_COL + e * (2 * DT_N);
#pragma unroll 1
            for (int jj = 0; jj < DT_JG; ++jj) {
                const int j = q * DT_JG + jj;
                if (j >= J) break;                                // uniform across the CTA
                float v[32], w2[32];
                {
                    float t0[16], t1[16], t2[16], t3[16];
                    tc::tmem_ld16(tbase + jj * DT_NOUT, t0);
                    tc::tmem_ld16(tbase + jj * DT_NOUT + 16, t1);
                    tc::tmem_ld16(tbase + DT_N + jj * DT_NOUT, t2);
                    tc::tmem_ld16(tbase + DT_N + jj * DT_NOUT + 16, t3);
#pragma unroll
                    for (int k = 0; k < 16; ++k) { v[k] = t0[k]; v[16 + k] = t1[k]; w2[k] = t2[k]; w2[16 + k] = t3[k]; }
                }
                if (!live) continue;
                const float* bj = Bias + j * DT_NOUT;
                float o[DT_NOUT];
#pragma unroll
                for (int k = 0; k < DT_NOUT; ++k) o[k] = (v[k] + w2[k]) + __ldg(bj + k);
                float prev[3];
                if (p.uvd_in) {
                    const float4 qv = __ldg(reinterpret_cast<const float4*>(p.uvd_in) + (static_cast<size_t>(b) * J + j) * HW + pix);
                    prev[0] = qv.x; prev[1] = qv.y; prev[2] = qv.z;
                } else {
                    const float* qv = d.pose + (static_cast<size_t>(b) * (3 + 6 * J) + 3 + 3 * j) * HW + pix;
                    prev[0] = __ldg(qv) * d.scale_uv;
                    prev[1] = __ldg(qv + HW) * d.scale_uv;
                    prev[2] = (j == p.root) ? 0.f : __ldg(qv + 2 * static_cast<size_t>(HW)) * d.scale_d;
                }
                float blend[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float gate = sigmoid_acc(o[2 * DT_NH + k]);
                    blend[k] = __fadd_rn(__fmul_rn(1.0f - gate, prev[k]), __fmul_rn(gate, o[2 * DT_NH + 3 + k]));
                }
                const size_t rec = ((static_cast<size_t>(b) * J + j) * HW + pix) * 2;            // float4 index into a plane
                float4* outS = reinterpret_cast<float4*>(p.proj) + rec;
                float4* outOC = reinterpret_cast<float4*>(p.proj) + static_cast<size_t>(p.B) * J * HW * 2 + rec;
                outS[0] = make_float4(o[0], o[1], o[2], o[3]);
                outS[1] = make_float4(o[4], o[5], o[6], o[7]);
                outOC[0] = make_float4(blend[0], blend[1], blend[2], o[2 * DT_NH + 6]);          // blended offset | conf.x
                outOC[1] = make_float4(o[2 * DT_NH + 7], o[2 * DT_NH + 8], 0.f, 0.f);            // conf.y, conf.z
            }
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&acc_free[e]);
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, 512);
}

// [J][17][C] packed weights -> per joint group [8 k-blocks][128 rows = 64 hi + 64 lo][128 B swizzled]
__global__ void pack_dense_panels_kernel(const float* __restrict__ wpack, unsigned char* __restrict__ dst, int J, int Q) {
    const long long total = static_cast<long long>(Q) * DT_KB * 2 * DT_N * 32;
    for (long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; t < total; t += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int e = static_cast<int>(t & 31);
        const int prow = static_cast<int>((t >> 5) % (2 * DT_N));
        const int kb = static_cast<int>((t / (32 * 2 * DT_N)) % DT_KB);
        const int q = static_cast<int>(t / (32LL * 2 * DT_N * DT_KB));
        const int part = prow / DT_N, r = prow % DT_N;
        const int jj = r / DT_NOUT, o = r % DT_NOUT;
        const int j = q * DT_JG + jj;
        float w = 0.f;
        if (r < DT_JG * DT_NOUT && j < J) w = wpack[(static_cast<size_t>(j) * DT_NOUT + o) * DT_C + kb * 32 + e];
        const float hi = tc::tf32_hi(w);
        const float v = part == 0 ? hi : (w - hi);
        const uint32_t off = static_cast<uint32_t>(prow) * 128u + ((static_cast<uint32_t>(e >> 2) ^ (prow & 7)) << 4) + (e & 3) * 4;
        *reinterpret_cast<float*>(dst + static_cast<size_t>(q) * DT_PANEL + kb * DT_PANEL_KB + off) = v;
    }
}

}  // namespace das

extern "C" int64_t das_dense_panel_bytes(const das_decode_cfg* cfg) {
    if (!cfg) return 0;
    const int Q = (cfg->num_joints + das::DT_JG - 1) / das::DT_JG;
    return static_cast<int64_t>(Q) * das::DT_PANEL;
}

extern "C" int das_pack_dense_panels(const das_decode_cfg* cfg, const float* packed_weights, void* panels, void* stream) {
    using namespace das;
    DAS_REQUIRE(cfg && packed_weights && panels, DAS_ERR_ARG, "das_pack_dense_panels: null pointer");
    DAS_REQUIRE(cfg->feat_channels == DT_C && cfg->num_heads == DT_NH, DAS_ERR_UNSUPPORTED,
                "tensor-core dense projection is built for feat_channels=256, num_heads=4");
    const int Q = (cfg->num_joints + DT_JG - 1) / DT_JG;
    pack_dense_panels_kernel<<<kSMs, 256, 0, static_cast<cudaStream_t>(stream)>>>(packed_weights, static_cast<unsigned char*>(panels),
                                                                                 cfg->num_joints, Q);
    DAS_CUDA_CHECK(cudaGetLastError());
    return DAS_OK;
}

// Tensor-core projection of one dense layer: proj[B*HW][J][16] (see das_refine_dense_layer for the sampling half).
extern "C" int das_dense_project_tc(const das_levels* d_levels, const das_levels* h_levels, int32_t level, int32_t layer,
                                    const das_decode_cfg* cfg, const float* weights, const void* panels,
                                    const float* uvd_in, float* proj, void* stream) {
    using namespace das;
    DAS_REQUIRE(d_levels && h_levels && cfg && weights && panels && proj, DAS_ERR_ARG, "das_dense_project_tc: null pointer");
    DAS_REQUIRE(cfg->feat_channels == DT_C && cfg->num_heads == DT_NH, DAS_ERR_UNSUPPORTED,
                "tensor-core dense projection is built for feat_channels=256, num_heads=4");
    DAS_REQUIRE(level >= 0 && level < h_levels->n_levels && layer >= 0 && layer < cfg->num_layers, DAS_ERR_ARG, "level/layer out of range");
    DenseTcParams p{};
    p.lv = d_levels; p.wpack = weights; p.panels = static_cast<const unsigned char*>(panels); p.uvd_in = uvd_in; p.proj = proj;
    p.level = level; p.layer = layer; p.J = cfg->num_joints; p.root = cfg->root_idx; p.B = h_levels->batch;
    p.Q = (cfg->num_joints + DT_JG - 1) / DT_JG;
    DAS_REQUIRE(p.Q <= kSMs, DAS_ERR_CAPACITY, "too many joint groups");
    static DeviceOnce attr_done;
    if (attr_done.need()) {
        DAS_CUDA_CHECK(cudaFuncSetAttribute(dense_project_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DT_SMEM));
    }
    const int grid = (kSMs / p.Q) * p.Q;
    dense_project_tc_kernel<<<grid, DT_THREADS, DT_SMEM, static_cast<cudaStream_t>(stream)>>>(p);
    DAS_CUDA_CHECK(cudaGetLastError());
    return DAS_OK;
}
