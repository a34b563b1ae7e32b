// tcgen05 / TMEM / mbarrier / cp.async building blocks (inline PTX, sm_100a).
//
// Operand layout used everywhere in this repo: K-major tiles with the 128-byte swizzle.  One "k-block" is
// 32 tf32 elements (128 B) wide; rows are 128 B apart, 8 rows form a 1024-B swizzle atom in which the 16-B
// chunk c of row r is stored at chunk position c ^ (r & 7); atoms are stacked along M/N every 1024 B (SBO).
// A tile base must be 1024-B aligned.  Advancing along K inside a k-block = adding the byte offset to the
// descriptor start address; the next k-block is a separate tile.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace das {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// byte offset of the 16-B chunk `chunk` (0..7) of row `row` inside a k-block tile
__device__ __forceinline__ uint32_t swz128(uint32_t row, uint32_t chunk) { return row * 128u + ((chunk ^ (row & 7u)) << 4); }

// ---- cp.async (LDGSTS), 16 B, zero-fill when !valid ------------------------------------------------
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
// same, but allocating in L1 (.ca): gathered feature rows are re-read by neighbouring heads / joints / candidates
__device__ __forceinline__ void cp_async16_ca(uint32_t dst, const void* src, bool valid) {
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// generic-proxy smem writes (st.shared, cp.async) -> visible to the async proxy (tensor core operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- mbarrier ------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"   // suspend-time hint: sleep in HW, do not poll
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)
        : "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

// ---- TMA (cp.async.bulk.tensor): one 2-D box global -> shared, completion on an mbarrier -----------------------
// tmap: CUtensorMap in kernel-parameter (__grid_constant__) space.  The box lands in the layout its swizzle mode
// encodes; with CU_TENSOR_MAP_SWIZZLE_128B and a 128-byte-wide box that is exactly swz128() above.
constexpr uint64_t kEvictNormal = 0x1000000000000000ull, kEvictFirst = 0x12F0000000000000ull, kEvictLast = 0x14F0000000000000ull;
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const void* tmap, uint64_t* bar, int c0, int c1, uint64_t hint) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_dst), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(hint)
        : "memory");
}
// 1-D bulk copy global -> shared (cp.async.bulk): one instruction moves `bytes` (multiple of 16, 16-B aligned on both
// sides) without touching the LSU; completion is signalled on the mbarrier like a tensor load.
__device__ __forceinline__ void bulk_load(uint32_t smem_dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_dst), "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) { asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory"); }

// ---- TMEM ----------------------------------------------------------------------------------------------
// one full warp; writes the allocated base address (lane 0, column base) to *dst (shared memory)
__device__ __forceinline__ void tmem_alloc(uint32_t* dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 lanes x 32 bit x 16 columns: thread t of warp w reads TMEM lane 32*(w%4)+t, columns [col, col+16)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// same load without the wait: issue several, then tmem_ld_wait() once before the first use of any of them
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 bit x 32 columns store: thread t of warp w writes TMEM lane 32*(w%4)+t, columns [col, col+32)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
        "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
        "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
        "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
        "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
        "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
        "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
        "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
        "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- UMMA descriptors --------------------------------------------------------------------------------
// shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row atoms 1024 B apart (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);      // start address, bits [0,14)
    d |= static_cast<uint64_t>(1) << 16;                       // LBO (ignored for swizzled K-major), bits [16,30)
    d |= static_cast<uint64_t>(1024 >> 4) << 32;               // SBO = 1024 B, bits [32,46)
    d |= static_cast<uint64_t>(1) << 46;                       // descriptor version 1 (sm_100)
    d |= static_cast<uint64_t>(2) << 61;                       // layout type SWIZZLE_128B
    return d;
}
// instruction descriptor for kind::tf32, fp32 accumulate, both operands K-major (cute::UMMA::InstrDescriptor)
__host__ __device__ constexpr uint32_t instr_desc_tf32(int M, int N) {
    return (1u << 4)                       // c_format = F32
           | (2u << 7)                     // a_format = TF32
           | (2u << 10)                    // b_format = TF32
           | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T, one thread issues
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
    const uint32_t acc = accumulate ? 1u : 0u;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(acc)
        : "memory");
}
// arrive on an mbarrier once every previously issued MMA of this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Warp-convergent variants: every lane executes the call, elect.sync picks the one lane that issues.  Keeping the
// issue path free of per-lane branches lets ptxas hold descriptors in uniform registers and emit a bare UTCHMMA
// instead of an ELECT / BRA.U.ANY loop around every instruction.  (elect.sync is deterministic for a fixed mask, so
// the MMAs and the commit that tracks them come from the same thread.)
__device__ __forceinline__ void umma_tf32_elect(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
    const uint32_t acc = accumulate ? 1u : 0u;
    asm volatile(
        "{\n\t"
        ".reg .pred p, q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint64_t* bar) {
    asm volatile(
        "{\n\t"
        ".reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
        "}" ::"r"(smem_u32(bar))
        : "memory");
}

// A operand from TMEM (lane = row, one 32-bit column per K element), B from shared memory
__device__ __forceinline__ void umma_tf32_ts_elect(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
    const uint32_t acc = accumulate ? 1u : 0u;
    asm volatile(
        "{\n\t"
        ".reg .pred p, q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc)
        : "memory");
}

// One k-block (4 k-steps of K = 8) of the 3xTF32 scheme in a single issue burst, A from TMEM:
//   D[:, 0:n32] (+)= A_hi[k] x [B_hi ; B_lo][k]     (idesc32)      D[:, 0:n16] += A_lo[k] x B_hi[k]   (idesc16)
// a_hi / a_lo: TMEM column addresses of the 32 hi / lo columns; db0: smem descriptor of the B k-block (k-step 0).
__device__ __forceinline__ void umma_kblock_3xtf32_ts(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint64_t db0,
                                                       uint32_t idesc32, uint32_t idesc16, bool accumulate_first) {
    const uint32_t acc = accumulate_first ? 1u : 0u;
    asm volatile(
        "{\n\t"
        ".reg .pred p, q;\n\t"
        ".reg .b64 d1, d2, d3;\n\t"
        ".reg .b32 h1, h2, h3, l1, l2, l3;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "add.s64 d1, %3, 2;\n\t add.s64 d2, %3, 4;\n\t add.s64 d3, %3, 6;\n\t"      // +32 B per k-step (>> 4)
        "add.u32 h1, %1, 8;\n\t add.u32 h2, %1, 16;\n\t add.u32 h3, %1, 24;\n\t"
        "add.u32 l1, %2, 8;\n\t add.u32 l2, %2, 16;\n\t add.u32 l3, %2, 24;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %3, %4, p;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%2], %3, %5, 1;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [h1], d1, %4, 1;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [l1], d1, %5, 1;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [h2], d2, %4, 1;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [l2], d2, %5, 1;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [h3], d3, %4, 1;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [l3], d3, %5, 1;\n\t"
        "}" ::"r"(tmem_d), "r"(a_hi), "r"(a_lo), "l"(db0), "r"(idesc32), "r"(idesc16), "r"(acc)
        : "memory");
}

// fp32 -> (hi, lo) with hi = the tf32 the tensor core will see (low 13 mantissa bits dropped), lo = a - hi (exact)
__device__ __forceinline__ float tf32_hi(float a) { return __uint_as_float(__float_as_uint(a) & 0xFFFFE000u); }

}  // namespace tc
}  // namespace das
