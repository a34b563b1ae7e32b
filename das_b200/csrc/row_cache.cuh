// Device-resident cache of feature rows for the host zero-copy mode (see das_refine_row_cache in refine_tc.cu):
// an open-addressing hash set keyed by the row's (host) address; the inserting thread gets a slot in the row buffer.
#pragma once
#include <cstdint>

#include "../../include/das_decode.h"

namespace das {

struct RowCacheView {
    unsigned long long* keys = nullptr;   // [mask + 1], all-ones = empty
    int32_t* slots = nullptr;             // [mask + 1], -1 = not assigned yet, -2 = row buffer full
    int32_t* counter = nullptr;           // starts at -1 (the table is cleared with 0xFF bytes)
    float* rows = nullptr;                // [cap][C]
    const float* cand_rows = nullptr;     // [B*CT][C] or null
    uint32_t mask = 0;
    int cap = 0;
};

inline RowCacheView row_cache_view(const das_row_cache* rc) {
    RowCacheView v;
    if (!rc || !rc->table) return v;
    const size_t n = static_cast<size_t>(1) << rc->table_bits;
    v.keys = static_cast<unsigned long long*>(rc->table);
    v.slots = reinterpret_cast<int32_t*>(v.keys + n);
    v.counter = v.slots + n;
    v.rows = rc->rows;
    v.cand_rows = rc->cand_rows;
    v.mask = static_cast<uint32_t>(n - 1);
    v.cap = rc->max_rows;
    return v;
}

// Insert `ptr` (never 0).  Returns true if this thread created the entry: then `slot` is its row-buffer slot (or -2
// when the buffer is full) and is already published.  Otherwise the entry exists at `h`; its slot may still be -1
// for a while (row_cache_wait).  With publish = true the row DATA of a slot is only guaranteed complete after the
// inserting kernel has finished.
// publish = false: the caller fills the row first and then calls row_cache_publish, so that a thread that finds the
// entry (row_cache_wait) may read the row's DATA in the same kernel; a full buffer (-2) is always published at once.
__device__ __forceinline__ bool row_cache_insert(const RowCacheView& rc, unsigned long long ptr, uint32_t& h, int& slot,
                                                 bool publish = true) {
    constexpr unsigned long long EMPTY = ~0ull;
    h = static_cast<uint32_t>(((ptr >> 10) * 0x9E3779B97F4A7C15ull) >> 40) & rc.mask;   // rows are >= 512 B apart
    while (true) {
        const unsigned long long old = atomicCAS(rc.keys + h, EMPTY, ptr);
        if (old == EMPTY) {
            const int sl = atomicAdd(rc.counter, 1) + 1;
            slot = sl < rc.cap ? sl : -2;
            if (publish || slot < 0) atomicExch(rc.slots + h, slot);
            return true;
        }
        if (old == ptr) { slot = -1; return false; }
        h = (h + 1) & rc.mask;
    }
}

__device__ __forceinline__ void row_cache_publish(const RowCacheView& rc, uint32_t h, int slot) {
    __threadfence();                       // the row's stores before the slot becomes visible
    atomicExch(rc.slots + h, slot);
}

__device__ __forceinline__ int row_cache_wait(const RowCacheView& rc, uint32_t h) {
    volatile int32_t* sl = rc.slots + h;
    int v;
    while ((v = *sl) == -1) {}
    return v;
}

}  // namespace das
