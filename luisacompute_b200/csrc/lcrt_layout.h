// lcrt_layout.h — blob and scratch memory layouts (host side, shared by the C-ABI and tests/hostsim).
#pragma once
#include "lcrt_common.h"

namespace lcrt {

inline uint64_t align_up(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }

struct BlobLayout {
    uint32_t node_capacity;
    uint32_t prim_stride;
    uint64_t nodes_offset, prims_offset, parent_offset, total_bytes;
};

inline uint32_t prim_stride_of(uint32_t kind) {
    return kind == KIND_TRIANGLES ? (uint32_t)sizeof(TriRecord) : kind == KIND_AABBS ? (uint32_t)sizeof(BoxRecord) : (uint32_t)sizeof(InstanceRecord);
}

inline BlobLayout blob_layout(uint32_t kind, uint32_t prim_count, uint32_t flags) {
    BlobLayout l;
    l.node_capacity = node_capacity_for(prim_count);
    l.prim_stride = prim_stride_of(kind);
    l.nodes_offset = sizeof(BlobHeader);
    l.prims_offset = align_up(l.nodes_offset + (uint64_t)l.node_capacity * NODE_BYTES, 128);
    const uint64_t prims_end = l.prims_offset + (uint64_t)(prim_count ? prim_count : 1u) * l.prim_stride;
    if (flags & BUILD_ALLOW_UPDATE) {
        l.parent_offset = align_up(prims_end, 128);
        l.total_bytes = align_up(l.parent_offset + (uint64_t)l.node_capacity * 4u, 256);
    } else {
        l.parent_offset = 0;
        l.total_bytes = align_up(prims_end, 256);
    }
    return l;
}

// scratch for one build of n primitives
struct TempLayout {
    uint64_t keys0, keys1, vals0, vals1;      // radix sort ping-pong
    uint64_t hist;                            // per-block digit histograms (one pass at a time)
    uint64_t lo, hi, child, range, flag;      // binary hierarchy
    uint64_t task;                            // wide-node work items
    uint64_t small;                           // centroid bounds (6 ordered uints), root id, counters[3]
    uint64_t total_bytes;
    uint32_t sort_blocks;
};

constexpr uint32_t SORT_TILE = 4096;     // keys per block per pass (256 threads x 16)
constexpr uint32_t SMALL_WORDS = 32;

inline TempLayout temp_layout(uint32_t n, uint32_t node_capacity) {
    TempLayout t;
    const uint64_t nn = n ? n : 1u;
    uint64_t off = 0;
    auto take = [&](uint64_t bytes) {
        const uint64_t o = off;
        off = align_up(off + bytes, 256);
        return o;
    };
    t.sort_blocks = (uint32_t)((nn + SORT_TILE - 1) / SORT_TILE);
    t.small = take(SMALL_WORDS * 4);
    t.flag = take(nn * 4);// zeroed together with `small` and `task`... (see build: one memset each)
    t.task = take((uint64_t)node_capacity * 4);
    t.keys0 = take(nn * 8), t.keys1 = take(nn * 8);
    t.vals0 = take(nn * 4), t.vals1 = take(nn * 4);
    {
        const uint64_t m = 256ull * t.sort_blocks;
        t.hist = take((m + (m + 2047) / 2048 + 64) * 4);// digit-major table followed by the scan's tile sums
    }
    t.lo = take((2 * nn) * 16), t.hi = take((2 * nn) * 16);
    t.child = take(nn * 8), t.range = take(nn * 8);
    t.total_bytes = off;
    return t;
}

// scratch for one refit
inline uint64_t refit_temp_bytes(uint32_t node_capacity) { return align_up((uint64_t)node_capacity * (16 + 16 + 4), 256) + 512; }

}// namespace lcrt
