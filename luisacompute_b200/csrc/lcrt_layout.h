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
    uint64_t small;                           // SMALL_WORDS u32: centroid bounds, counters, control words, the root half
    uint64_t flag;                            // [n] arrival counters of the binary nodes
    uint64_t task;                            // [node_capacity] u2 wide-node work items
    uint64_t keys0, keys1;                    // radix sort ping-pong (u64 keys = Morton << index_bits | primitive)
    uint64_t control;                         // sort histograms / tickets / look-back words
    uint64_t pbox;                            // [n] 2 x f4 primitive boxes in input order
    uint64_t node;                            // [n] 96-byte binary nodes
    uint64_t total_bytes;
};

constexpr uint32_t SORT_TILE = 4096;     // keys per block per pass (256 threads x 16)
constexpr uint32_t SORT_CONTROL_PASSES = 8;
constexpr uint32_t SMALL_WORDS = 256;

inline uint64_t sort_control_bytes_for(uint32_t n) {// == sort_control_bytes() of lcrt_sort.cuh (static_assert'ed in lcrt_kernels.cu)
    const uint64_t tiles = ((uint64_t)(n ? n : 1u) + SORT_TILE - 1) / SORT_TILE;
    return ((uint64_t)SORT_CONTROL_PASSES * 256u + SORT_CONTROL_PASSES + 64u + (uint64_t)SORT_CONTROL_PASSES * tiles * 256u) * 4u;
}

inline TempLayout temp_layout(uint32_t n, uint32_t node_capacity) {
    TempLayout t;
    const uint64_t nn = n ? n : 1u;
    uint64_t off = 0;
    auto take = [&](uint64_t bytes) {
        const uint64_t o = off;
        off = align_up(off + bytes, 256);
        return o;
    };
    t.small = take(SMALL_WORDS * 4);
    t.flag = take(nn * 4);
    t.task = take((uint64_t)node_capacity * 8);
    t.keys0 = take(nn * 8), t.keys1 = take(nn * 8);
    t.control = take(sort_control_bytes_for(n));
    t.pbox = take(nn * 32);
    t.node = take(nn * 96);
    t.total_bytes = off;
    return t;
}

// scratch for one refit
inline uint64_t refit_temp_bytes(uint32_t node_capacity) { return align_up((uint64_t)node_capacity * (16 + 16 + 4), 256) + 512; }

}// namespace lcrt
