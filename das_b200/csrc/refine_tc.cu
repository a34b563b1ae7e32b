// Stage 3+4 on the tensor cores: the sampling phase of the sparse last-layer refinement as a gathered
// 3xTF32 GEMM (tcgen05.mma kind::tf32, A operand and accumulator in TMEM), for C = 256, num_heads = 4.
//
// Reference semantics: recursive_update.py:186-197 (gate / value / confidence 1x1 projections and the gated
// blend), :34-82 + :9-31 (bilinear sampling with zero padding, softmax over the 2*nh heads), das_head.py:252-262
// and :725-743 (eval tail + joint assembly).  refine_sparse.cu (phases 1-2, "heads only" mode) has already
// produced, for every (centre, joint) item, the 2*nh sampling offsets; what is left is, per item, 32 feature
// rows (8 heads x 4 bilinear corners) times that joint's 9 projection rows -- a real dense contraction once
// items of the SAME joint are batched:
//
//   tile   = 4 items of one joint = 128 gathered feature rows  (A: 128 x 256, fp32 split into tf32 hi/lo)
//   B      = that joint's {gate 3, value 3, conf 3} rows, zero-padded to N = 16, pre-split into hi/lo and
//            pre-swizzled by das_pack_tc_panels; per k-block the 16 hi rows are followed by the 16 lo rows
//   D      = A_hi [B_hi;B_lo]^T (N = 32)  +  A_lo B_hi^T (N = 16, into the first 16 columns)
//            -> 3xTF32: fp32-level accuracy, fp32 accumulation in TMEM, 2 MMAs per K = 8 step
//
// One CTA per SM walks a contiguous range of tiles, warp-specialised (measured design history in DESIGN.md):
//   3 producer groups (4 warps each; k-blocks round-robin) gather their k-block with coalesced cp.async into a
//                     private 3-stage smem ring; then every thread reads ITS row (thread = row = TMEM lane, conflict-
//                     free thanks to the 128-B swizzle), splits hi/lo in registers and tcgen05.st's both into a
//                     4-slot TMEM ring -- the tensor core never reads A from shared memory
//   MMA warp          one fused burst of 8 tcgen05.mma per k-block (~47 cycles each: an M=128,K=8 tcgen05.mma costs
//                     ~50 cycles for any N <= 64, A from smem or TMEM alike), tcgen05.commit to the slot's mbarrier;
//                     also (re)loads the B panels when the joint changes
//   3 epilogue groups (4 warps each; thread = row) drain the triple-buffered accumulators: bias, sigmoid gate, blend
//                     with the previous offset, bilinear weight / zero padding, corner sum (2 shuffles), softmax
//                     over the 8 heads (3 shuffles), eval tail, assembly; they also prepare the row pointers and
//                     per-row state 6 tiles ahead and prefetch those rows into L2
#include <algorithm>
#include <cstdlib>

#include "refine_common.cuh"
#include "row_cache.cuh"
#include "tc_common.cuh"

namespace das {

constexpr int TC_EPI_WARPS = 4;                 // one epilogue / producer group = 4 warps = the 128 TMEM lanes
constexpr int TC_N = 16;                        // MMA N: 9 projection rows + zero padding
constexpr int TC_KB = 8;                        // k-blocks of 32 channels (C = 256)
constexpr int TC_C = 256;
constexpr int TC_NH = 4;
constexpr int TC_A_BYTES = 128 * 128;           // one k-block of 128 rows
constexpr int TC_B_BYTES = TC_KB * TC_N * 128;  // hi or lo rows of one joint: 16 KB (a joint's panel = 2x that)
constexpr int TC_BK_BYTES = 2 * TC_N * 128;     // panel bytes per k-block: 16 hi rows followed by 16 lo rows
constexpr int TC_NOUT = 2 * TC_NH + 9;
constexpr int TC_OGATE = 2 * TC_NH;

struct TcParams {
    const das_levels* lv;
    const float* wpack;              // biases live behind the [J][17][C] weights
    const unsigned char* bpanel;     // [J][2][TC_B_BYTES] swizzled hi / lo panels
    const float* urow;               // distinct feature rows per joint [J][row_cap][8] = {ptr lo, ptr hi, prev u, v, d, -, -, -}
    float* ures;                     // per distinct row [J][row_cap][8] = {O u, v, d, conf u, v, d, -, -} (written by the GEMM epilogue)
    const float* lrow;               // row records [B*CT*J][32][4] = {distinct-row index, bilinear weight, head offset x, y}
    const float* item_asm;           // assembly records [B*CT*J][8] = {Px, Py, zq, sx, sy, stride, -, -}
    const int32_t* valid_list;       // candidates that survive score_thr, any order
    const int32_t* n_valid;
    const int32_t* joint_count;      // [J] distinct rows of every joint
    float* cand_pose;
    int CT, J, root, split, row_cap, gather_cg, prefetch_lines;
    float depth_factor, z_norm;
    long long* dbg;                  // optional [gridDim.x][16] cycle counters (profiling builds of the host code)
};

struct RowState {                    // what thread t keeps about row t of a tile until its epilogue
    const float* ptr;                // feature row, nullptr = padding row of a joint's last tile
    float prev0, prev1, prev2;
    int r;                           // index of the row in its joint's list
};

// tile index -> (joint, 128-row chunk of that joint's distinct-row list); pref[j] = first tile of joint j, pref[J] = total
__device__ __forceinline__ void tile_joint(const int* pref, int J, int tile, int& j, int& chunk) {
    j = 0;
    while (j + 1 < J && pref[j + 1] <= tile) ++j;
    chunk = tile - pref[j];
}

// Row t of tile (j, chunk) = entry chunk*128 + t of joint j's distinct-row list.  The phase-1/2 kernel (das_refine_heads)
// has already resolved every distinct sampled cell to a 32-byte record; setting a row up is two loads.
__device__ __forceinline__ RowState setup_row(const TcParams& p, const int* pref, int tile, int tid) {
    RowState r;
    int j, chunk;
    tile_joint(pref, p.J, tile, j, chunk);
    r.r = chunk * 128 + tid;
    r.ptr = nullptr; r.prev0 = r.prev1 = r.prev2 = 0.f;
    if (r.r >= __ldg(p.joint_count + j)) return r;
    const float4* rec = reinterpret_cast<const float4*>(p.urow) + (static_cast<size_t>(j) * p.row_cap + r.r) * 2;
    const float4 a = __ldg(rec), c = __ldg(rec + 1);
    r.ptr = reinterpret_cast<const float*>(static_cast<unsigned long long>(__float_as_uint(a.x)) |
                                           (static_cast<unsigned long long>(__float_as_uint(a.y)) << 32));
    r.prev0 = a.z; r.prev1 = a.w; r.prev2 = c.x;
    if (r.ptr) {
        // Optional (DAS_TC_PREFETCH_LINES = 1..8; default 0): pull the row's lines into L2 six tiles ahead of the gather.  It paid
        // in round 1 (32 rows per item, 2 400 tiles); with the distinct-row lists it costs more than it hides -- the L2 read
        // path is the kernel's limit, and 6 tiles x 128 KB of prefetch per SM are 113 MB in flight against a 126-MB L2:
        // config #2 38.9 -> 36.9 us for the stage (1.19 M -> 1.23 M images/s), spread workload 75.8 -> 68.9 us without it
#pragma unroll
        for (int q = 0; q < TC_C * 4 / 128; ++q)
            if (q < p.prefetch_lines) asm volatile("prefetch.global.L2 [%0];" ::"l"(r.ptr + q * 32));
    }
    return r;
}

// Row epilogue of the tensor-core kernel: v[0..8] = this distinct row's {gate 3, value 3, conf 3} projections ->
// bias, sigmoid gate, blend with the previous offset at that cell (recursive_update.py:193-195) and the confidence
// logits; one 32-byte result per distinct (cell, joint), shared by every head / corner that sampled the cell.
__device__ __forceinline__ void tc_epilogue(const TcParams& p, const RowState& cur, const float (&v)[16], int j) {
    if (!cur.ptr) return;
    const float* Bj = p.wpack + static_cast<size_t>(p.J) * TC_NOUT * TC_C + j * TC_NOUT + TC_OGATE;
    const float pv[3] = {cur.prev0, cur.prev1, cur.prev2};
    float o[3], cf[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float gte = sigmoid_acc(v[k] + __ldg(Bj + k));
        const float n = v[3 + k] + __ldg(Bj + 3 + k);
        o[k] = __fadd_rn(__fmul_rn(1.0f - gte, pv[k]), __fmul_rn(gte, n));
        cf[k] = v[6 + k] + __ldg(Bj + 6 + k);
    }
    float4* dst = reinterpret_cast<float4*>(p.ures) + (static_cast<size_t>(j) * p.row_cap + cur.r) * 2;
    dst[0] = make_float4(o[0], o[1], o[2], cf[0]);
    dst[1] = make_float4(cf[1], cf[2], 0.f, 0.f);
}

constexpr int T2_PGROUPS = 3;                   // producer groups (4 warps each)
constexpr int T2_STAGES = 3;                    // smem stages per group
constexpr int T2_SLOTS = 4;                     // TMEM A slots (64 columns each: 32 hi + 32 lo)
constexpr int T2_EG = 3;                        // epilogue groups: tile i -> group i % 3, accumulator i % 3
constexpr int T2_RB = 2 * T2_EG;                // row-pointer buffers: rows are set up 2 own tiles (= 6 tiles) ahead
constexpr int T2_FIRST_PRODUCER = TC_EPI_WARPS * T2_EG;                         // warp 12
constexpr int T2_MMA_WARP = T2_FIRST_PRODUCER + 4 * T2_PGROUPS;                // warp 24
constexpr int T2_THREADS = 32 * (T2_MMA_WARP + 1);
constexpr int T2_TMEM_COLS = 512;
constexpr int T2_D_COL = 0;                     // accumulators: T2_EG x 32 columns
constexpr int T2_A_COL = 32 * T2_EG;            // A ring: 4 slots x 64 columns
constexpr int T2_SMEM = 1024 + T2_PGROUPS * T2_STAGES * TC_A_BYTES + 2 * 2 * TC_B_BYTES;   // A rings + two joint panels (current, prefetched)

// PROF: per-role cycle counters into p.dbg (tools/tc_role_cycles.py); the production instantiation carries none.
template <bool PROF>
__global__ void __launch_bounds__(T2_THREADS, 1)
refine_tc2_kernel(const TcParams p) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    unsigned char* sA = base;                                              // [group][stage] k-block tiles
    unsigned char* sB = sA + T2_PGROUPS * T2_STAGES * TC_A_BYTES;          // per k-block: 16 hi rows then 16 lo rows
    __shared__ uint64_t a_full[T2_SLOTS], a_empty[T2_SLOTS];               // producers -> MMA, MMA -> producers
    __shared__ uint64_t acc_full[T2_EG], acc_free[T2_EG], rows_ready[T2_RB], b_full[2];
    __shared__ uint32_t tmem_base;
    __shared__ const float* s_rowptr[T2_RB][128];   // row pointers are produced T2_RB tiles ahead of their epilogue

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long kernel_t0 = (PROF ? clock64() : 0ll);
    pdl_wait();          // the row lists and their counters come from das_refine_heads
    pdl_trigger();
    const int J = p.J;
    __shared__ int s_pref[DAS_MAX_JOINTS + 1];      // first tile of every joint (tiles = 128-row chunks of its distinct-row list)
    if (tid == 0) {
        int acc = 0;
        for (int j = 0; j < J; ++j) { s_pref[j] = acc; acc += (__ldg(p.joint_count + j) + 127) >> 7; }
        s_pref[J] = acc;
    }
    __syncthreads();
    const int n_tiles = s_pref[J];
    const int t0 = static_cast<int>(static_cast<long long>(n_tiles) * blockIdx.x / gridDim.x);
    const int t1 = static_cast<int>(static_cast<long long>(n_tiles) * (blockIdx.x + 1) / gridDim.x);
    if (t0 >= t1) return;
    const int my_tiles = t1 - t0;
    const int total_kb = my_tiles * TC_KB;

    if (tid == 0) {
        for (int s = 0; s < T2_SLOTS; ++s) { tc::mbar_init(&a_full[s], 4); tc::mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < T2_EG; ++s) { tc::mbar_init(&acc_full[s], 1); tc::mbar_init(&acc_free[s], TC_EPI_WARPS); }
        for (int s = 0; s < T2_RB; ++s) tc::mbar_init(&rows_ready[s], TC_EPI_WARPS);
        tc::mbar_init(&b_full[0], 1); tc::mbar_init(&b_full[1], 1);
        tc::mbar_fence_init();
    }
    if (warp == 0) tc::tmem_alloc(&tmem_base, T2_TMEM_COLS);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem0 = tmem_base;
    const uint32_t sB_u = tc::smem_u32(sB);

    if (warp == T2_MMA_WARP) {
        // ===== MMA issuer warp (also owns the B panels) ================================================================
        constexpr uint32_t idesc32 = tc::instr_desc_tf32(128, 2 * TC_N);   // A_hi x [B_hi ; B_lo]  -> D[:, 0:32]
        constexpr uint32_t idesc16 = tc::instr_desc_tf32(128, TC_N);       // A_lo x  B_hi          -> D[:, 0:16]
        int g = 0, cur_j = -1, cur_buf = 0;
        long long m_w = 0, m_i = 0, m_b = 0, m_f = 0;
        // Two panel buffers, filled by bulk copies (one instruction per 32-KB panel, no LSU traffic next to the producers'
        // gathers): the first joint's panel and -- a CTA's contiguous tile range spans at most two joints unless the lists
        // are tiny -- the second joint's are requested up front, so neither the start nor the switch inside the range
        // waits for 64 rounds of cp.async (the synchronous reload took 29 % of the kernel: tools/tc_role_cycles.py).
        int buf_joint[2] = {-1, -1};
        uint32_t buf_phase[2] = {0u, 0u};
        auto load_panel = [&](int j, int buf) {
            if (lane == 0) {
                tc::mbar_arrive_expect_tx(&b_full[buf], 2 * TC_B_BYTES);
                tc::bulk_load(sB_u + buf * 2 * TC_B_BYTES, p.bpanel + static_cast<size_t>(j) * 2 * TC_B_BYTES, 2 * TC_B_BYTES, &b_full[buf]);
            }
            buf_joint[buf] = j;
        };
        {
            int j0, j1, chunk_unused;
            tile_joint(s_pref, J, t0, j0, chunk_unused);
            load_panel(j0, 0);
            tile_joint(s_pref, J, t1 - 1, j1, chunk_unused);
            if (j1 != j0) {
                int jn = j0 + 1;
                while (jn < j1 && s_pref[jn + 1] == s_pref[jn]) ++jn;     // next joint that owns a tile
                load_panel(jn, 1);
            }
        }
        for (int i = 0; i < my_tiles; ++i) {
            const uint32_t dcol = tmem0 + T2_D_COL + (i % T2_EG) * 32;
            const long long tb0 = (PROF ? clock64() : 0ll);
            int j, chunk_unused;
            tile_joint(s_pref, J, t0 + i, j, chunk_unused);
            if (j != cur_j) {
                if (buf_joint[0] == j || buf_joint[1] == j) {
                    cur_buf = buf_joint[0] == j ? 0 : 1;
                } else {
                    // third joint of a range (tiny lists only): every earlier MMA must have finished reading the panels
                    if (g > 0) tc::mbar_wait(&a_empty[(g - 1) % T2_SLOTS], ((g - 1) / T2_SLOTS) & 1);
                    cur_buf ^= 1;
                    load_panel(j, cur_buf);
                }
                tc::mbar_wait(&b_full[cur_buf], buf_phase[cur_buf]);
                buf_phase[cur_buf] ^= 1u;
                cur_j = j;
            }
            const long long tb1 = (PROF ? clock64() : 0ll);
            if (i >= T2_EG) tc::mbar_wait(&acc_free[i % T2_EG], ((i / T2_EG) - 1) & 1);   // epilogue of tile i-EG has drained this buffer
            const long long tb2 = (PROF ? clock64() : 0ll);
            m_b += tb1 - tb0; m_f += tb2 - tb1;
            for (int kb = 0; kb < TC_KB; ++kb, ++g) {
                const int slot = g % T2_SLOTS;
                const long long c0 = (PROF ? clock64() : 0ll);
                tc::mbar_wait(&a_full[slot], (g / T2_SLOTS) & 1);
                tc::tc_fence_after();
                const long long c1 = (PROF ? clock64() : 0ll);
                m_w += c1 - c0;
                const uint32_t a_hi = tmem0 + T2_A_COL + slot * 64, a_lo = a_hi + 32;
                const uint32_t b_pk = sB_u + cur_buf * 2 * TC_B_BYTES + kb * TC_BK_BYTES;
                if (p.split & 1) {
                    tc::umma_kblock_3xtf32_ts(dcol, a_hi, a_lo, tc::smem_desc_sw128(b_pk), idesc32, idesc16, kb != 0);
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        tc::umma_tf32_ts_elect(dcol, a_hi + k * 8, tc::smem_desc_sw128(b_pk + k * 32), idesc16, (kb | k) != 0);
                }
                tc::umma_commit_elect(&a_empty[slot]);
                if (kb == TC_KB - 1) tc::umma_commit_elect(&acc_full[i % T2_EG]);
                m_i += (PROF ? clock64() : 0ll) - c1;
            }
        }
        if (PROF && p.dbg && lane == 0) { long long* o = p.dbg + blockIdx.x * 16; o[0] = m_w; o[1] = m_i; o[2] = m_f; o[14] = m_b; }
    } else if (warp >= T2_FIRST_PRODUCER) {
        // ===== producer groups: gather (cp.async, coalesced) -> own row from smem -> hi/lo -> TMEM ====================
        const int pg = (warp - T2_FIRST_PRODUCER) >> 2;          // group: handles k-blocks g with g % 3 == pg
        const int qw = warp & 3;                                 // TMEM lane quadrant (T2_FIRST_PRODUCER % 4 == 0)
        const int gt = (qw << 5) | lane;                         // thread within the group = row it owns
        const uint32_t sG_u = tc::smem_u32(sA) + pg * T2_STAGES * TC_A_BYTES;
        unsigned char* sG = sA + pg * T2_STAGES * TC_A_BYTES;
        const int bar_id = 1 + pg;
        const int my_n = (total_kb - pg + T2_PGROUPS - 1) / T2_PGROUPS;   // number of k-blocks of this group
        // coalesced gather mapping: chunk c = gt + 128 * it  ->  row c >> 3, 16-B chunk c & 7
        auto gather = [&](int n) {
            const int g = pg + n * T2_PGROUPS;
            const int i = g / TC_KB, kb = g - i * TC_KB, st = n % T2_STAGES;
            if (kb < T2_PGROUPS) tc::mbar_wait(&rows_ready[i % T2_RB], (i / T2_RB) & 1);   // first k-block of tile i seen by this group
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int row = (gt >> 3) + 16 * it, ch = gt & 7;
                const float* src = s_rowptr[i % T2_RB][row];
                if (p.gather_cg) tc::cp_async16(sG_u + st * TC_A_BYTES + tc::swz128(row, ch), src ? src + kb * 32 + ch * 4 : p.wpack, src != nullptr);
                else tc::cp_async16_ca(sG_u + st * TC_A_BYTES + tc::swz128(row, ch), src ? src + kb * 32 + ch * 4 : p.wpack, src != nullptr);
            }
        };
        for (int n = 0; n < T2_STAGES - 1; ++n) { if (n < my_n) gather(n); tc::cp_async_commit(); }
        long long d0 = 0, d1 = 0, d2 = 0, d3 = 0, d4 = 0;
        for (int n = 0; n < my_n; ++n) {
            const int g = pg + n * T2_PGROUPS;
            const int slot = g % T2_SLOTS, st = n % T2_STAGES;
            const long long c0 = (PROF ? clock64() : 0ll);
            tc::cp_async_wait<T2_STAGES - 2>();                  // this thread's chunks of k-block n have landed
            asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");   // ... and everybody else's in the group
            const long long c1 = (PROF ? clock64() : 0ll);
            float hi[32];
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) {
                const float4 a = *reinterpret_cast<const float4*>(sG + st * TC_A_BYTES + tc::swz128(gt, ch));
                hi[4 * ch] = a.x; hi[4 * ch + 1] = a.y; hi[4 * ch + 2] = a.z; hi[4 * ch + 3] = a.w;
            }
            asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");   // stage `st` may be overwritten from here on
            const long long c2 = (PROF ? clock64() : 0ll);
            if (n + T2_STAGES - 1 < my_n) gather(n + T2_STAGES - 1);     // into stage (n + 2) % 3 == (n - 1) % 3, read last step
            tc::cp_async_commit();
            const long long c3 = (PROF ? clock64() : 0ll);
            if (g >= T2_SLOTS) tc::mbar_wait(&a_empty[slot], ((g / T2_SLOTS) - 1) & 1);   // MMAs of k-block g-4 done with this TMEM slot
            tc::tc_fence_after();
            const long long c4 = (PROF ? clock64() : 0ll);
            const uint32_t taddr = tmem0 + (static_cast<uint32_t>(qw * 32) << 16) + T2_A_COL + slot * 64;
            tc::tmem_st32(taddr, hi);
            if (p.split & 1) {
                float lo[32];
#pragma unroll
                for (int e = 0; e < 32; ++e) lo[e] = hi[e] - tc::tf32_hi(hi[e]);
                tc::tmem_st32(taddr + 32, lo);
            }
            tc::tmem_st_wait();
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&a_full[slot]);
            d0 += c1 - c0; d1 += c2 - c1; d2 += c3 - c2; d3 += c4 - c3; d4 += (PROF ? clock64() : 0ll) - c4;
        }
        tc::cp_async_wait<0>();
        if (PROF && p.dbg && pg == 0 && gt == 0) { long long* o = p.dbg + blockIdx.x * 16; o[3] = d0; o[4] = d1; o[5] = d2; o[6] = d3; o[7] = d4; o[8] = my_n; }
    } else {
        // ===== epilogue / row-setup warps: thread t <-> row t = (item warp, head lane>>2, corner lane&3) ===========
        const int e = warp / TC_EPI_WARPS;          // epilogue group
        const int rt = tid - e * 32 * TC_EPI_WARPS; // row of the tile this thread owns (= TMEM lane)
        const int qw = warp % TC_EPI_WARPS;         // TMEM lane quadrant this warp may read
        // this group's tiles are e, e+EG, e+2EG, ...; their row state is prepared two of them (= 2*EG tiles) ahead
        RowState cur{}, nx1{};
        if (e < my_tiles) {
            cur = setup_row(p, s_pref, t0 + e, rt);
            s_rowptr[e][rt] = cur.ptr;
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&rows_ready[e]);
        }
        if (e + T2_EG < my_tiles) {
            nx1 = setup_row(p, s_pref, t0 + e + T2_EG, rt);
            s_rowptr[e + T2_EG][rt] = nx1.ptr;
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&rows_ready[e + T2_EG]);
        }
#pragma unroll 1
        for (int i = e; i < my_tiles; i += T2_EG) {
            const int tile = t0 + i;
            int j, chunk_unused;
            tile_joint(s_pref, J, tile, j, chunk_unused);
            const long long e0 = (PROF ? clock64() : 0ll);
            tc::mbar_wait(&acc_full[e], (i / T2_EG) & 1);
            tc::tc_fence_after();
            const long long e1 = (PROF ? clock64() : 0ll);
            float v[16], v2[16];
            const uint32_t taddr = tmem0 + (static_cast<uint32_t>(qw * 32) << 16) + T2_D_COL + e * 32;
            tc::tmem_ld16(taddr, v);
            if (p.split & 1) {
                tc::tmem_ld16(taddr + TC_N, v2);
#pragma unroll
                for (int k = 0; k < 9; ++k) v[k] += v2[k];
            }
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&acc_free[e]);
            // tile i+2EG reuses tile i's pointer buffer: every gather of tile i was issued before its MMAs completed
            RowState nx2{};
            if (i + T2_RB < my_tiles) {
                nx2 = setup_row(p, s_pref, tile + T2_RB, rt);
                s_rowptr[i % T2_RB][rt] = nx2.ptr;
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&rows_ready[i % T2_RB]);
            }
            tc_epilogue(p, cur, v, j);
            cur = nx1;
            nx1 = nx2;
            if (PROF && p.dbg && tid == 0) { long long* o = p.dbg + blockIdx.x * 16; o[9] += e1 - e0; o[10] += (PROF ? clock64() : 0ll) - e1; }
        }
    }
    if (PROF && p.dbg && tid == 0) p.dbg[blockIdx.x * 16 + 13] = (PROF ? clock64() : 0ll) - kernel_t0;
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, T2_TMEM_COLS);
}

// [J][17][C] packed weights -> per joint [8 k-blocks][32 rows = 16 hi + 16 lo][128 B swizzled]
__global__ void pack_tc_panels_kernel(const float* __restrict__ wpack, unsigned char* __restrict__ dst, int J) {
    const int total = J * 2 * TC_KB * TC_N * 32;   // one thread per float
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const int e = t & 31;                      // channel inside the k-block
        const int row = (t >> 5) % TC_N;
        const int kb = (t / (32 * TC_N)) % TC_KB;
        const int part = (t / (32 * TC_N * TC_KB)) & 1;
        const int j = t / (32 * TC_N * TC_KB * 2);
        float w = 0.f;
        if (row < 9) w = wpack[(static_cast<size_t>(j) * TC_NOUT + TC_OGATE + row) * TC_C + kb * 32 + e];
        const float hi = tc::tf32_hi(w);
        const float v = part == 0 ? hi : (w - hi);
        const int prow = part * TC_N + row;          // 16 hi rows then 16 lo rows inside the k-block tile (N = 32 view)
        const uint32_t off = static_cast<uint32_t>(prow) * 128u + ((static_cast<uint32_t>(e >> 2) ^ (prow & 7)) << 4) + (e & 3) * 4;
        *reinterpret_cast<float*>(dst + static_cast<size_t>(j) * 2 * TC_B_BYTES + kb * TC_BK_BYTES + off) = v;
    }
}

}  // namespace das

static long long* g_tc_dbg = nullptr;
// profiling aid: per-CTA cycle counters of the three warp roles ([148][16] int64 device buffer, or NULL to disable)
extern "C" int das_tc_set_debug_buffer(long long* dev_buf) { g_tc_dbg = dev_buf; return DAS_OK; }

extern "C" int64_t das_tc_panel_bytes(const das_decode_cfg* cfg) {
    return cfg ? static_cast<int64_t>(cfg->num_joints) * 2 * das::TC_B_BYTES : 0;
}

extern "C" int das_pack_tc_panels(const das_decode_cfg* cfg, const float* packed_weights, void* panels, void* stream) {
    using namespace das;
    DAS_REQUIRE(cfg && packed_weights && panels, DAS_ERR_ARG, "das_pack_tc_panels: null pointer");
    DAS_REQUIRE(cfg->feat_channels == TC_C && cfg->num_heads == TC_NH, DAS_ERR_UNSUPPORTED,
                "tensor-core refinement is built for feat_channels=256, num_heads=4");
    pack_tc_panels_kernel<<<kSMs, 256, 0, static_cast<cudaStream_t>(stream)>>>(packed_weights, static_cast<unsigned char*>(panels),
                                                                              cfg->num_joints);
    DAS_CUDA_CHECK(cudaGetLastError());
    return DAS_OK;
}

namespace das {
// ---- row cache (host zero-copy mode) ---------------------------------------------------------------------------------
// When the feature maps live in pinned HOST memory the sampling phase would pull every row record's 1 KB row over PCIe,
// although the 32 rows of neighbouring heads / joints / candidates are mostly the same few rows (on device memory L2
// absorbs that 10x re-use; sysmem reads are not deduplicated the same way: measured 117 MB per 64-image batch for
// ~31 MB of distinct rows).  This pass runs between the two refinement kernels: every distinct row address is inserted
// into an open-addressing hash set (row_cache.cuh), the inserting lane's warp copies that row ONCE into a
// device-resident row buffer, and every record is re-pointed at the copy.  Rows the heads kernel already left in the
// cache are found, not fetched again.  A full buffer just leaves the remaining records pointing at the host.
__global__ void __launch_bounds__(256)
row_cache_kernel(float* urow, const int32_t* joint_count, int row_cap, const RowCacheView rc) {
    // grid: x = warps over a joint's distinct-row list (32 records per warp step), y = joint
    const int lane = threadIdx.x & 31;
    const int j = blockIdx.y;
    const int n = __ldg(joint_count + j);
    for (int r0 = (blockIdx.x * 8 + (threadIdx.x >> 5)) * 32; r0 < n; r0 += gridDim.x * 8 * 32) {
        const int r = r0 + lane;
        float4* rec = reinterpret_cast<float4*>(urow) + (static_cast<size_t>(j) * row_cap + min(r, n - 1)) * 2;
        float4 a = *rec;
        const unsigned long long ptr = r < n ? (static_cast<unsigned long long>(__float_as_uint(a.x)) |
                                                (static_cast<unsigned long long>(__float_as_uint(a.y)) << 32)) : 0ull;
        int slot = -1;
        bool won = false;
        uint32_t h = 0;
        int ins = -1;
        if (ptr) { ins = row_cache_insert(rc, ptr, h, slot); won = ins > 0; }
        __syncwarp();
        if (ptr && ins == 0) slot = row_cache_wait(rc, h);       // bounded; a negative slot leaves the record pointing at the host row
        unsigned todo = __ballot_sync(0xffffffffu, won && slot >= 0);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const unsigned long long sp = __shfl_sync(0xffffffffu, ptr, src);
            const int ss = __shfl_sync(0xffffffffu, slot, src);
            const float4* s4 = reinterpret_cast<const float4*>(sp) + lane * 2;        // 32 lanes x 32 B = one row
            const float4 v0 = __ldg(s4), v1 = __ldg(s4 + 1);
            float4* d4 = reinterpret_cast<float4*>(rc.rows + static_cast<size_t>(ss) * TC_C) + lane * 2;
            d4[0] = v0;
            d4[1] = v1;
        }
        if (ptr && slot >= 0) {
            const unsigned long long np = reinterpret_cast<unsigned long long>(rc.rows + static_cast<size_t>(slot) * TC_C);
            a.x = __uint_as_float(static_cast<uint32_t>(np));
            a.y = __uint_as_float(static_cast<uint32_t>(np >> 32));
            *rec = a;
        }
    }
}

// Last step of the sparse refinement: one warp per (candidate, joint) item, lane = (head = lane >> 2, corner = lane & 3).
// Every lane fetches the {O, conf} result of the distinct cell its row record points at, applies its bilinear weight
// (zero padding outside the map), the 4 corners are summed, the head offset added (recursive_update.py:72-75, 28),
// softmax over the 2*nh heads per dim (:29-31), eval tail and joint assembly (das_head.py:254-262, 725-743).
__global__ void __launch_bounds__(256)
refine_finish_kernel(const TcParams p) {
    const int lane = threadIdx.x & 31;
    pdl_wait();          // unique_out comes from das_refine_tc
    pdl_trigger();
    const int J = p.J;
    const int n_items = __ldg(p.n_valid) * J;
    for (int it = blockIdx.x * 8 + (threadIdx.x >> 5); it < n_items; it += gridDim.x * 8) {
        const int cs = __ldg(p.valid_list + it / J);
        const int j = it - (it / J) * J;
        const size_t item = static_cast<size_t>(cs) * J + j;
        const float4 rec = __ldg(reinterpret_cast<const float4*>(p.lrow) + item * 32 + lane);
        const int gidx = __float_as_int(rec.x);
        const float wk = rec.y, hxv = rec.z, hyv = rec.w;
        float val[3] = {0.f, 0.f, 0.f}, cf[3] = {0.f, 0.f, 0.f};
        if (gidx >= 0) {
            const float4* res = reinterpret_cast<const float4*>(p.ures) + (static_cast<size_t>(j) * p.row_cap + gidx) * 2;
            const float4 a = __ldcg(res), c = __ldcg(res + 1);
            val[0] = a.x * wk; val[1] = a.y * wk; val[2] = a.z * wk;
            cf[0] = a.w * wk; cf[1] = c.x * wk; cf[2] = c.y * wk;
        }
        float out[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            // bilinear sum over the 4 corners (lane bits 0,1)
            val[k] += __shfl_xor_sync(FULL, val[k], 1);
            cf[k] += __shfl_xor_sync(FULL, cf[k], 1);
            val[k] += __shfl_xor_sync(FULL, val[k], 2);
            cf[k] += __shfl_xor_sync(FULL, cf[k], 2);
            const float hv = val[k] + (k == 0 ? hxv : (k == 1 ? hyv : 0.f));   // + diff
            // softmax over the 8 heads (lane bits 2,3,4); every corner lane carries the same head value
            float m = cf[k];
            m = fmaxf(m, __shfl_xor_sync(FULL, m, 4));
            m = fmaxf(m, __shfl_xor_sync(FULL, m, 8));
            m = fmaxf(m, __shfl_xor_sync(FULL, m, 16));
            const float e = expf(cf[k] - m);
            float se = e;
            se += __shfl_xor_sync(FULL, se, 4);
            se += __shfl_xor_sync(FULL, se, 8);
            se += __shfl_xor_sync(FULL, se, 16);
            float o = hv * (e / se);
            o += __shfl_xor_sync(FULL, o, 4);
            o += __shfl_xor_sync(FULL, o, 8);
            o += __shfl_xor_sync(FULL, o, 16);
            out[k] = o;
        }
        if (lane < 3) {
            // eval tail + assembly (das_head.py:254-262, 725-743) from the item's record {Px, Py, zq, sx, sy, stride}
            const float4* ar = reinterpret_cast<const float4*>(p.item_asm) + item * 2;
            const float4 a0 = __ldg(ar), a1 = __ldg(ar + 1);
            const float o = lane == 0 ? out[0] : (lane == 1 ? out[1] : out[2]);
            float r;
            if (lane == 0) r = __fdiv_rn(__fadd_rn(__fmul_rn(o, a1.y), a0.x), a0.w);
            else if (lane == 1) r = __fdiv_rn(__fadd_rn(__fmul_rn(o, a1.y), a0.y), a1.x);
            else r = __fadd_rn((j == p.root) ? 0.0f : __fmul_rn(o, p.z_norm), a0.z);
            p.cand_pose[item * 3 + lane] = r;
        }
    }
}
}  // namespace das

static int check_row_cache(const das_row_cache* rc, const char* who) {
    using namespace das;
    DAS_REQUIRE(rc && rc->table && rc->rows, DAS_ERR_ARG, "%s: null row cache", who);
    DAS_REQUIRE(rc->table_bits >= 10 && rc->table_bits <= 26 && rc->max_rows >= 1, DAS_ERR_ARG, "%s: table_bits=%d max_rows=%d",
                who, rc->table_bits, rc->max_rows);
    return DAS_OK;
}

extern "C" int64_t das_row_cache_table_bytes(int32_t table_bits) {
    if (table_bits < 10 || table_bits > 26) return 0;
    return (static_cast<int64_t>(1) << table_bits) * 12 + 64;      // keys (8 B) + slots (4 B) + the row counter
}

extern "C" int das_row_cache_clear(const das_row_cache* rc, void* stream) {
    using namespace das;
    DAS_TRY(check_row_cache(rc, "das_row_cache_clear"));
    DAS_CUDA_CHECK(cudaMemsetAsync(rc->table, 0xFF, static_cast<size_t>(das_row_cache_table_bytes(rc->table_bits)),
                                   static_cast<cudaStream_t>(stream)));
    return DAS_OK;
}

static int check_scratch(const das_refine_scratch* sc, const char* who) {
    using namespace das;
    DAS_REQUIRE(sc && sc->unique_rows && sc->unique_out && sc->row_records && sc->item_records && sc->valid_list && sc->counters &&
                sc->row_cap >= 1, DAS_ERR_ARG, "%s: null / empty scratch", who);
    return DAS_OK;
}

extern "C" int das_refine_row_cache(const das_decode_cfg* cfg, const das_refine_scratch* scratch, const das_row_cache* rc, void* stream) {
    using namespace das;
    DAS_REQUIRE(cfg, DAS_ERR_ARG, "das_refine_row_cache: null pointer");
    DAS_TRY(check_scratch(scratch, "das_refine_row_cache"));
    DAS_TRY(check_row_cache(rc, "das_refine_row_cache"));
    DAS_REQUIRE(cfg->feat_channels == TC_C, DAS_ERR_UNSUPPORTED, "the row cache is built for feat_channels=256");
    const dim3 grid(std::min(kSMs * 2, (scratch->row_cap + 255) / 256), cfg->num_joints);
    row_cache_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(scratch->unique_rows, scratch->counters + 4, scratch->row_cap,
                                                                         row_cache_view(rc));
    DAS_CUDA_CHECK(cudaGetLastError());
    return DAS_OK;
}

static das::TcParams tc_params(const das_levels* d_levels, const das_decode_cfg* cfg, const float* weights, const void* panels,
                               int32_t cand_slots, const das_refine_scratch* sc, float* cand_pose, int32_t split) {
    das::TcParams p{};
    p.lv = d_levels; p.wpack = weights; p.bpanel = static_cast<const unsigned char*>(panels);
    p.urow = sc->unique_rows; p.ures = sc->unique_out; p.lrow = sc->row_records; p.item_asm = sc->item_records;
    p.valid_list = sc->valid_list; p.n_valid = sc->counters + 1; p.joint_count = sc->counters + 4;
    p.cand_pose = cand_pose;
    p.CT = cand_slots; p.J = cfg->num_joints; p.root = cfg->root_idx; p.row_cap = sc->row_cap;
    p.split = split; p.depth_factor = cfg->depth_factor; p.z_norm = cfg->z_norm;
    return p;
}

extern "C" int das_refine_tc(const das_levels* d_levels, const das_levels* h_levels, const das_decode_cfg* cfg,
                             const float* weights, const void* panels, int32_t cand_slots,
                             const das_refine_scratch* scratch, int32_t split, void* stream) {
    using namespace das;
    DAS_REQUIRE(d_levels && h_levels && cfg && weights && panels, DAS_ERR_ARG, "das_refine_tc: null pointer");
    DAS_TRY(check_scratch(scratch, "das_refine_tc"));
    DAS_REQUIRE(cfg->feat_channels == TC_C && cfg->num_heads == TC_NH, DAS_ERR_UNSUPPORTED,
                "tensor-core refinement is built for feat_channels=256, num_heads=4");
    DAS_REQUIRE(cfg->num_joints >= 1 && cfg->num_joints <= DAS_MAX_JOINTS, DAS_ERR_CAPACITY, "num_joints=%d", cfg->num_joints);
    TcParams p = tc_params(d_levels, cfg, weights, panels, cand_slots, scratch, nullptr, split);
    p.dbg = g_tc_dbg;
    static const int gather_cg = std::getenv("DAS_TC_GATHER_CG") ? std::atoi(std::getenv("DAS_TC_GATHER_CG")) : 1;   // .cg: no L1 allocation (the rows are distinct)
    p.gather_cg = gather_cg;
    static const int pf_lines = std::getenv("DAS_TC_PREFETCH_LINES") ? std::atoi(std::getenv("DAS_TC_PREFETCH_LINES")) : 0;
    p.prefetch_lines = pf_lines;
    static DeviceOnce attr2_done;
    if (attr2_done.need()) {
        DAS_CUDA_CHECK(cudaFuncSetAttribute(refine_tc2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, T2_SMEM));
        DAS_CUDA_CHECK(cudaFuncSetAttribute(refine_tc2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, T2_SMEM));
    }
    if (p.dbg) DAS_CUDA_CHECK(launch_chain(refine_tc2_kernel<true>, dim3(kSMs), dim3(T2_THREADS), T2_SMEM, static_cast<cudaStream_t>(stream), chain_ctx().pdl, p));
    else DAS_CUDA_CHECK(launch_chain(refine_tc2_kernel<false>, dim3(kSMs), dim3(T2_THREADS), T2_SMEM, static_cast<cudaStream_t>(stream), chain_ctx().pdl, p));
    DAS_CUDA_CHECK(cudaGetLastError());
    return DAS_OK;
}

extern "C" int das_refine_finish(const das_levels* h_levels, const das_decode_cfg* cfg, int32_t cand_slots,
                                 const das_refine_scratch* scratch, float* cand_pose, void* stream) {
    using namespace das;
    DAS_REQUIRE(h_levels && cfg && cand_pose, DAS_ERR_ARG, "das_refine_finish: null pointer");
    DAS_TRY(check_scratch(scratch, "das_refine_finish"));
    DAS_REQUIRE(cfg->num_heads == TC_NH, DAS_ERR_UNSUPPORTED, "das_refine_finish is built for num_heads=4");
    TcParams p = tc_params(nullptr, cfg, nullptr, nullptr, cand_slots, scratch, cand_pose, 0);
    const long long items = static_cast<long long>(h_levels->batch) * cand_slots * cfg->num_joints;
    const int grid = static_cast<int>(std::max<long long>(1, std::min<long long>((items + 7) / 8, 8LL * kSMs)));
    DAS_CUDA_CHECK(launch_chain(refine_finish_kernel, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream), chain_ctx().pdl, p));
    DAS_CUDA_CHECK(cudaGetLastError());
    return DAS_OK;
}
