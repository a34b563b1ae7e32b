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

#include <cuda.h>

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
constexpr int DT_PG = 2;                         // producer groups (k-block g goes to group g % 2)
constexpr int DT_STAGES = 6;                     // TMA ring of A k-block tiles: 5 k-blocks (~2300 MMA cycles) of look-ahead
constexpr int DT_SLOTS = 4;                      // TMEM A slots (64 columns: 32 hi + 32 lo)
constexpr int DT_A_BYTES = 128 * 128;
constexpr int DT_FIRST_PRODUCER = 4 * DT_EG;     // warp 8
constexpr int DT_MMA_WARP = DT_FIRST_PRODUCER + 4 * DT_PG;   // warp 16
constexpr int DT_TMA_WARP = DT_MMA_WARP + 1;     // warp 17: one elected lane issues the TMA loads
constexpr int DT_THREADS = 32 * (DT_TMA_WARP + 1);
constexpr int DT_D_COL = 0;                      // accumulators: 2 x 128 columns
constexpr int DT_A_COL = 2 * DT_N * DT_EG;       // 256: A ring 4 x 64 columns
constexpr int DT_SMEM = 1024 + DT_STAGES * DT_A_BYTES + DT_PANEL;

struct DenseTcParams {
    const das_levels* lv;
    const float* wpack;            // [J][17][C] + biases
    const unsigned char* panels;   // [Q][DT_PANEL]
    const float* uvd_in;           // nullptr -> scaled raw uvd from lv.pose; else joint-major [B][J][HW][4]
    float* proj;                   // four joint-major planes, see DensePlanes (refine_common.cuh)
    int* progress;                 // [G] tiles issued by the Q CTAs that share tile sequence m (zeroed before the launch)
    int level, layer, J, root, B, Q;
    long long* dbg;                // optional [gridDim.x][8] cycle counters (tools/dense_role_cycles.py); nullptr in production
};

constexpr int DT_LOCKSTEP_TILES = 2;             // the Q CTAs that read the same feature tiles stay within this many tiles

// The feature map of a level as a 2-D tensor {C = 256 floats, B*H*W cells}: one TMA box = 32 channels x 128 cells
// (a k-block of a 128-cell tile), 128-byte swizzle = the K-major operand layout of tc_common.cuh.
__global__ void __maxnreg__(96)                // 18 warps are allocated as 20 (granularity 4): 20 x 32 x 96 = 60 K of the 64 K registers (104 and 112 fail to launch); one CTA per SM anyway (225 KB of shared memory)
dense_project_tc_kernel(const DenseTcParams p, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    unsigned char* sA = base;                                      // [stage] k-block tiles (128-B swizzle), filled by TMA
    unsigned char* sB = sA + DT_STAGES * DT_A_BYTES;               // resident panel of this CTA's joint group
    __shared__ uint64_t st_full[DT_STAGES], st_empty[DT_STAGES];   // TMA -> producers, producers -> TMA
    __shared__ uint64_t a_full[DT_SLOTS], a_empty[DT_SLOTS], acc_full[DT_EG], acc_free[DT_EG];
    __shared__ uint32_t tmem_base;
    __shared__ float s_bias[DT_JG * DT_NOUT];                      // biases of this CTA's joint group

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long kernel_t0 = p.dbg ? clock64() : 0ll;
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

    if (tid < DT_JG * DT_NOUT) {
        const int j = q * DT_JG + tid / DT_NOUT;
        s_bias[tid] = j < J ? p.wpack[static_cast<size_t>(J) * DT_NOUT * DT_C + j * DT_NOUT + tid % DT_NOUT] : 0.f;
    }
    if (tid == 0) {
        for (int s = 0; s < DT_STAGES; ++s) { tc::mbar_init(&st_full[s], 1); tc::mbar_init(&st_empty[s], 4); }
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
    const uint32_t sA_u = tc::smem_u32(sA);

    if (warp == DT_TMA_WARP) {
        // ===== TMA issuer: streams the k-blocks of this CTA's tiles through the ring, DT_STAGES - 1 ahead ============
        if (lane == 0) {
            tc::tma_prefetch_desc(&tmap);
            volatile int* prog = p.progress + m;
            bool lockstep = true;
            for (int g = 0; g < total_kb; ++g) {
                const int s = g % DT_STAGES;
                if (g >= DT_STAGES) tc::mbar_wait(&st_empty[s], ((g / DT_STAGES) - 1) & 1);   // the 4 warps that read it are done
                const int i = g / DT_KB, kb = g - i * DT_KB;
                if (kb == 0) {
                    // Loose lock-step of the Q CTAs (one per joint group) that walk the same tiles: nobody starts tile i
                    // before everybody has issued tile i - DT_LOCKSTEP_TILES, so a feature tile is fetched from HBM by
                    // the first CTA and found in L2 by the other Q - 1 (unsynchronised they drift apart by more than
                    // the L2 holds: ncu showed 40 % L2 hits and 2.8x the map's bytes from DRAM).  All CTAs of the grid
                    // are co-resident (grid <= SM count, one CTA per SM), so the wait cannot deadlock.
                    if (i > 0) atomicAdd(p.progress + m, 1);                                   // tile i - 1 fully issued
                    const int need = Q * (i - DT_LOCKSTEP_TILES);
                    // bounded: if a sibling CTA is not resident (two such grids sharing the SMs from different
                    // streams), give the lock-step up after ~2 ms instead of waiting for a CTA that cannot start
                    for (int spin = 0; lockstep && *prog < need; ++spin) {
                        __nanosleep(64);
                        if (spin > (1 << 15)) lockstep = false;
                    }
                }
                const long long cell0 = (static_cast<long long>(m) + static_cast<long long>(i) * G) * 128;
                tc::mbar_arrive_expect_tx(&st_full[s], DT_A_BYTES);
                // rows beyond the map (last tile) are zero-filled by the TMA unit and still count towards the byte total
                tc::tma_load_2d(sA_u + s * DT_A_BYTES, &tmap, &st_full[s], kb * 32, static_cast<int>(cell0), tc::kEvictNormal);
            }
        }
    } else if (warp == DT_MMA_WARP) {
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
        long long m_free = 0, m_full = 0, m_issue = 0;
        const bool prof = p.dbg != nullptr;
        for (int i = 0; i < my_tiles; ++i) {
            const uint32_t dcol = tmem0 + DT_D_COL + (i % DT_EG) * (2 * DT_N);
            const long long c0 = prof ? clock64() : 0ll;
            if (i >= DT_EG) tc::mbar_wait(&acc_free[i % DT_EG], ((i / DT_EG) - 1) & 1);
            if (prof) m_free += clock64() - c0;
            for (int kb = 0; kb < DT_KB; ++kb, ++g) {
                const int slot = g % DT_SLOTS;
                const long long c1 = prof ? clock64() : 0ll;
                tc::mbar_wait(&a_full[slot], (g / DT_SLOTS) & 1);
                tc::tc_fence_after();
                const long long c2 = prof ? clock64() : 0ll;
                const uint32_t a_hi = tmem0 + DT_A_COL + slot * 64;
                tc::umma_kblock_3xtf32_ts(dcol, a_hi, a_hi + 32, tc::smem_desc_sw128(sB_u + kb * DT_PANEL_KB), idesc_hi, idesc_lo, kb != 0);
                tc::umma_commit_elect(&a_empty[slot]);
                if (kb == DT_KB - 1) tc::umma_commit_elect(&acc_full[i % DT_EG]);
                if (prof) { m_full += c2 - c1; m_issue += clock64() - c2; }
            }
        }
        if (prof && lane == 0) { long long* o = p.dbg + blockIdx.x * 8; o[0] += m_free; o[1] += m_full; o[2] += m_issue; }
    } else if (warp >= DT_FIRST_PRODUCER) {
        // ===== producer groups: own row of the TMA-staged k-block -> hi/lo split in registers -> TMEM ===============
        const int pg = (warp - DT_FIRST_PRODUCER) >> 2;
        const int qw = warp & 3;
        const int gt = (qw << 5) | lane;
        long long d_tma = 0, d_slot = 0, d_work = 0;
        const bool prof = p.dbg != nullptr && pg == 0 && gt == 0;
        for (int g = pg; g < total_kb; g += DT_PG) {
            const int slot = g % DT_SLOTS, s = g % DT_STAGES;
            const long long c0 = prof ? clock64() : 0ll;
            tc::mbar_wait(&st_full[s], (g / DT_STAGES) & 1);
            const long long c1 = prof ? clock64() : 0ll;
            float hi[32];
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) {
                const float4 a = *reinterpret_cast<const float4*>(sA + s * DT_A_BYTES + tc::swz128(gt, ch));
                hi[4 * ch] = a.x; hi[4 * ch + 1] = a.y; hi[4 * ch + 2] = a.z; hi[4 * ch + 3] = a.w;
            }
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&st_empty[s]);          // this warp's 32 rows are in registers
            const long long c2 = prof ? clock64() : 0ll;
            if (g >= DT_SLOTS) tc::mbar_wait(&a_empty[slot], ((g / DT_SLOTS) - 1) & 1);
            tc::tc_fence_after();
            const long long c3 = prof ? clock64() : 0ll;
            const uint32_t taddr = tmem0 + (static_cast<uint32_t>(qw * 32) << 16) + DT_A_COL + slot * 64;
            tc::tmem_st32(taddr, hi);
#pragma unroll
            for (int e = 0; e < 32; ++e) hi[e] = hi[e] - tc::tf32_hi(hi[e]);
            tc::tmem_st32(taddr + 32, hi);
            tc::tmem_st_wait();
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&a_full[slot]);
            if (prof) { d_tma += c1 - c0; d_slot += c3 - c2; d_work += (c2 - c1) + (clock64() - c3); }
        }
        if (prof) { long long* o = p.dbg + blockIdx.x * 8; o[3] += d_tma; o[4] += d_slot; o[5] += d_work; }
    } else {
        // ===== epilogue groups: thread = cell; bias, gate, blend with the previous offset -> the four planes =============
        const int e = warp >> 2;
        const int qw = warp & 3;
        const int row = (qw << 5) | lane;
        const int nj = min(DT_JG, J - q * DT_JG);                  // joints of this group (uniform across the CTA)
        const DensePlanes pl = dense_planes(p.proj, p.B, J, HW);
        // the previous offsets of a cell for the group's joints, RAW (the Scale factors are applied at use): loaded one
        // own tile ahead, so the global-load latency hides behind two tiles of MMAs instead of stalling the drain
        auto load_prev = [&](int i, float (&pv)[DT_JG][3]) {
            const long long cell = (static_cast<long long>(m) + static_cast<long long>(i) * G) * 128 + row;
#pragma unroll
            for (int jj = 0; jj < DT_JG; ++jj) pv[jj][0] = pv[jj][1] = pv[jj][2] = 0.f;
            if (i >= my_tiles || cell >= cells) return;
            const int b = static_cast<int>(cell / HW);
            const int pix = static_cast<int>(cell - static_cast<long long>(b) * HW);
#pragma unroll
            for (int jj = 0; jj < DT_JG; ++jj) {
                if (jj >= nj) break;
                const int j = q * DT_JG + jj;
                if (p.uvd_in) {
                    const float4 qv = __ldg(reinterpret_cast<const float4*>(p.uvd_in) + (static_cast<size_t>(b) * J + j) * HW + pix);
                    pv[jj][0] = qv.x; pv[jj][1] = qv.y; pv[jj][2] = qv.z;
                } else {
                    const InMap pose(d.pose, p.lv->in_dtype);
                    const size_t qv = (static_cast<size_t>(b) * (3 + 6 * J) + 3 + 3 * j) * HW + pix;
                    pv[jj][0] = pose(qv);
                    pv[jj][1] = pose(qv + HW);
                    pv[jj][2] = (j == p.root) ? 0.f : pose(qv + 2 * static_cast<size_t>(HW));
                }
            }
        };
        const float sc_uv = p.uvd_in ? 1.0f : d.scale_uv, sc_d = p.uvd_in ? 1.0f : d.scale_d;
        float prev[DT_JG][3], nxt[DT_JG][3];
        load_prev(e, prev);
        for (int i = e; i < my_tiles; i += DT_EG) {
            const long long cell = (static_cast<long long>(m) + static_cast<long long>(i) * G) * 128 + row;
            const bool live = cell < cells;
            const int b = live ? static_cast<int>(cell / HW) : 0;
            const int pix = live ? static_cast<int>(cell - static_cast<long long>(b) * HW) : 0;
            load_prev(i + DT_EG, nxt);
            const long long e0 = (p.dbg && e == 0 && row == 0) ? clock64() : 0ll;
            tc::mbar_wait(&acc_full[e], (i / DT_EG) & 1);
            tc::tc_fence_after();
            if (p.dbg && e == 0 && row == 0) p.dbg[blockIdx.x * 8 + 6] += clock64() - e0;
            const uint32_t tbase = tmem0 + (static_cast<uint32_t>(qw * 32) << 16) + DT_D_COL + e * (2 * DT_N);
#pragma unroll
            for (int jj = 0; jj < DT_JG; ++jj) {
                if (jj >= nj) break;                               // uniform across the CTA
                const int j = q * DT_JG + jj;
                // outputs 0..15 (+ their lo-pass partners in the second half of the accumulator), then output 16
                float t0[16], t2[16], t1[16], t3[16];
                tc::tmem_ld16_nowait(tbase + jj * DT_NOUT, t0);
                tc::tmem_ld16_nowait(tbase + DT_N + jj * DT_NOUT, t2);
                tc::tmem_ld16_nowait(tbase + jj * DT_NOUT + 16, t1);
                tc::tmem_ld16_nowait(tbase + DT_N + jj * DT_NOUT + 16, t3);
                tc::tmem_ld_wait();
                if (jj == nj - 1) {                                // accumulator drained: hand it back before the stores
                    tc::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(&acc_free[e]);
                }
                if (!live) continue;
                const float* bj = s_bias + jj * DT_NOUT;
                float o[DT_NOUT];
#pragma unroll
                for (int k = 0; k < 16; ++k) o[k] = (t0[k] + t2[k]) + bj[k];
                o[16] = (t1[0] + t3[0]) + bj[16];
                // layer 0 reads the raw predictor output: das_head.py:243-249 Scale factors here (exact: one multiply each)
                const float pv0 = prev[jj][0] * sc_uv, pv1 = prev[jj][1] * sc_uv, pv2 = prev[jj][2] * sc_d;
                const float pvk[3] = {pv0, pv1, pv2};
                float blend[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float gate = sigmoid_acc(o[2 * DT_NH + k]);
                    blend[k] = __fadd_rn(__fmul_rn(1.0f - gate, pvk[k]), __fmul_rn(gate, o[2 * DT_NH + 3 + k]));
                }
                const size_t cellj = (static_cast<size_t>(b) * J + j) * HW + pix;
                // 16-byte plane entries: the 32 threads of a warp (consecutive cells) fill whole 128-byte lines
                pl.s0[cellj] = make_float4(o[0], o[1], o[2], o[3]);
                pl.s1[cellj] = make_float4(o[4], o[5], o[6], o[7]);
                pl.oa[cellj] = make_float4(blend[0], blend[1], blend[2], o[2 * DT_NH + 6]);           // blended offset | conf.u
                pl.cb[cellj] = make_float2(o[2 * DT_NH + 7], o[2 * DT_NH + 8]);                       // conf.v, conf.d
            }
#pragma unroll
            for (int jj = 0; jj < DT_JG; ++jj) { prev[jj][0] = nxt[jj][0]; prev[jj][1] = nxt[jj][1]; prev[jj][2] = nxt[jj][2]; }
        }
    }
    if (p.dbg && tid == 0) p.dbg[blockIdx.x * 8 + 7] += clock64() - kernel_t0;
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

static long long* g_dense_dbg = nullptr;
// profiling aid: per-CTA cycle counters of the dense projection's warp roles ([148][8] int64 device buffer, zeroed by the
// caller, accumulated over the launches; NULL = off): {mma: wait acc_free, wait a_full, issue | producer group 0: wait TMA,
// wait TMEM slot, work | epilogue group 0: wait acc_full | CTA total}
extern "C" int das_dense_set_debug_buffer(long long* dev_buf) { g_dense_dbg = dev_buf; return DAS_OK; }

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
    p.dbg = g_dense_dbg;
    DAS_REQUIRE(p.Q <= kSMs, DAS_ERR_CAPACITY, "too many joint groups");
    static DeviceOnce attr_done;
    if (attr_done.need()) {
        DAS_CUDA_CHECK(cudaFuncSetAttribute(dense_project_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DT_SMEM));
    }
    // TMA descriptor of this level's feature map: {256 channels, B*H*W cells}, box = 32 channels x 128 cells, 128-B swizzle
    const float* feat = h_levels->lv[level].feats[layer];
    DAS_REQUIRE(feat && (reinterpret_cast<uintptr_t>(feat) & 15) == 0, DAS_ERR_ARG, "das_dense_project_tc: feature map is null or not 16-byte aligned");
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        DAS_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        DAS_REQUIRE(fn && qres == cudaDriverEntryPointSuccess, DAS_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
        encode = reinterpret_cast<EncodeFn>(fn);
    }
    const cuuint64_t cells = static_cast<cuuint64_t>(h_levels->batch) * h_levels->lv[level].H * h_levels->lv[level].W;
    const cuuint64_t gdim[2] = {DT_C, cells};
    const cuuint64_t gstride[1] = {DT_C * sizeof(float)};
    const cuuint32_t box[2] = {32, 128};
    const cuuint32_t estr[2] = {1, 1};
    CUtensorMap tmap;
    const CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(feat), gdim, gstride, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DAS_REQUIRE(cr == CUDA_SUCCESS, DAS_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(cr));
    const int grid = (kSMs / p.Q) * p.Q;
    {
        // lock-step counters live in the unused tail of the projection scratch (16 floats are allocated per (cell, joint),
        // the four planes use 14)
        const size_t n = static_cast<size_t>(p.B) * p.J * h_levels->lv[level].H * h_levels->lv[level].W;
        DAS_REQUIRE(n * 8 >= static_cast<size_t>(kSMs) * sizeof(int), DAS_ERR_ARG, "das_dense_project_tc: map too small for the tensor-core path");
        const DensePlanes pl = dense_planes(proj, p.B, p.J, h_levels->lv[level].H * h_levels->lv[level].W);
        p.progress = reinterpret_cast<int*>(pl.cb + n);
        DAS_CUDA_CHECK(cudaMemsetAsync(p.progress, 0, kSMs * sizeof(int), static_cast<cudaStream_t>(stream)));
    }
    dense_project_tc_kernel<<<grid, DT_THREADS, DT_SMEM, static_cast<cudaStream_t>(stream)>>>(p, tmap);
    DAS_CUDA_CHECK(cudaGetLastError());
    return DAS_OK;
}
