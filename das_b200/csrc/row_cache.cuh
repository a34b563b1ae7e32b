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

constexpr int kRowCacheMaxProbes = 512;      // open-addressing probe limit: beyond it the table counts as full
constexpr int kRowCacheMaxPolls = 1 << 16;   // bounded wait for another thread's copy (~tens of ms), then fall back

// Insert `ptr` (never 0).  Returns 1 if this thread created the entry: then `slot` is its row-buffer slot (or -2
// when the buffer is full) and is already published.  Returns 0 if the entry exists at `h`; its slot may still be -1
// for a while (row_cache_wait).  Returns -1 if no free table entry was found within kRowCacheMaxProbes probes (table
// full): nothing was inserted, slot = -2, and the caller keeps reading the row from where it lives (always correct --
// the cache only ever holds copies).  With publish = true the row DATA of a slot is only guaranteed complete after
// the inserting kernel has finished.
// publish = false: the caller fills the row first and then calls row_cache_publish, so that a thread that finds the
// entry (row_cache_wait) may read the row's DATA in the same kernel; a full buffer (-2) is always published at once.
__device__ __forceinline__ int row_cache_insert(const RowCacheView& rc, unsigned long long ptr, uint32_t& h, int& slot,
                                                bool publish = true) {
    constexpr unsigned long long EMPTY = ~0ull;
    h = static_cast<uint32_t>(((ptr >> 10) * 0x9E3779B97F4A7C15ull) >> 40) & rc.mask;   // rows are >= 512 B apart
    for (int probe = 0; probe < kRowCacheMaxProbes; ++probe) {
        const unsigned long long old = atomicCAS(rc.keys + h, EMPTY, ptr);
        if (old == EMPTY) {
            const int sl = atomicAdd(rc.counter, 1) + 1;
            slot = sl < rc.cap ? sl : -2;
            if (publish || slot < 0) atomicExch(rc.slots + h, slot);
            return 1;
        }
        if (old == ptr) { slot = -1; return 0; }
        h = (h + 1) & rc.mask;
    }
    slot = -2;
    return -1;
}

__device__ __forceinline__ void row_cache_publish(const RowCacheView& rc, uint32_t h, int slot) {
    __threadfence();                       // the row's stores before the slot becomes visible
    atomicExch(rc.slots + h, slot);
}

// Wait for the inserting thread to publish the slot of entry `h`.  Bounded: a publisher that is not resident (or is
// itself waiting) must never hang the grid, so after kRowCacheMaxPolls polls the waiter gives up and returns -2 = "read
// the row from its home address", which is always correct.
__device__ __forceinline__ int row_cache_wait(const RowCacheView& rc, uint32_t h) {
    volatile int32_t* sl = rc.slots + h;
    for (int poll = 0; poll < kRowCacheMaxPolls; ++poll) {
        const int v = *sl;
        if (v != -1) return v;
        __nanosleep(64);
    }
    return -2;
}

}  // namespace das
