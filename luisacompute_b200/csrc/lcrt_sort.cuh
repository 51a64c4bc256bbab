// lcrt_sort.cuh — CUB-free LSD radix sort, 8-bit digits, "one sweep" per digit (after Adinets & Merrill, "Onesweep: A Faster
// Least Significant Digit Radix Sort for GPUs", 2022).
//
//   k_sort_histograms   ONE pass over the keys builds the digit histograms of ALL passes (shared-memory atomics, one
//                       global atomic per (pass, digit, block)).
//   k_sort_onesweep     one launch per digit: each block takes the next tile (atomic ticket, so a tile only ever waits
//                       for tiles that are already running), ranks its 4096 keys (stable: every warp owns consecutive
//                       keys, same-digit peers from eight ballots), publishes the tile's digit counts, resolves its
//                       global scatter base by decoupled look-back over the preceding tiles' counts, reorders the
//                       tile by digit in shared memory and writes each digit's run contiguously.
// P passes = P + 1 launches (round 1: 5 P launches — histogram, three-kernel scan, scatter per pass) and per pass the
// keys are read ONCE and written once: 2 x 8 B per key and pass for keys only (the builder packs the primitive index
// into the low bits of the Morton key, lcrt_build.h), 2 x 12 B with a u32 payload (lcrtSortPairs).
// The reference has no sort of its own (its optional `lcub` extension wraps cub::DeviceRadixSort,
// src/backends/cuda/lcub/private/dcub/device_radix_sort.cu) — north_star asks for a CUB-free one.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lcrt {

constexpr int SORT_THREADS = 256;
constexpr int SORT_ITEMS = 16;                          // keys per thread
constexpr int SORT_TILE_KEYS = SORT_THREADS * SORT_ITEMS;// 4096 == SORT_TILE in lcrt_layout.h
constexpr int SORT_MAX_PASSES = 8;
constexpr uint32_t SORT_FLAG_AGGREGATE = 1u << 30, SORT_FLAG_INCLUSIVE = 2u << 30, SORT_VALUE_MASK = (1u << 30) - 1u;
// look-back words carry a 30-bit count: n < 2^30 keys
constexpr uint32_t SORT_MAX_KEYS = (1u << 30) - 1u;

struct SortControl {// device scratch of one sort: zeroed once per sort (one memset)
    // [pass][256] global digit histograms, then [pass] tile tickets, then [pass][tile][256] look-back words
    uint32_t *hist;
    uint32_t *ticket;
    uint32_t *status;
};

__global__ void __launch_bounds__(SORT_THREADS) k_sort_histograms(const uint64_t *__restrict__ keys, uint32_t n, uint32_t shift0, uint32_t passes,
                                                                  uint32_t *__restrict__ hist) {
    __shared__ uint32_t sh[SORT_MAX_PASSES][256];
    for (uint32_t p = 0; p < passes; p++) sh[p][threadIdx.x] = 0u;
    __syncthreads();
    for (uint32_t i = blockIdx.x * SORT_THREADS + threadIdx.x; i < n; i += gridDim.x * SORT_THREADS) {
        const uint64_t k = keys[i] >> shift0;
        for (uint32_t p = 0; p < passes; p++) atomicAdd(&sh[p][(uint32_t)(k >> (8u * p)) & 255u], 1u);
    }
    __syncthreads();
    for (uint32_t p = 0; p < passes; p++) {
        const uint32_t c = sh[p][threadIdx.x];
        if (c) atomicAdd(&hist[p * 256u + threadIdx.x], c);
    }
}

// exclusive scan of one value per thread over a 256-thread block
__device__ __forceinline__ uint32_t block_exclusive_scan_256(uint32_t v, uint32_t *smem /*[8]*/) {
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (uint32_t)o) inc += t;
    }
    if (lane == 31u) smem[warp] = inc;
    __syncthreads();
    uint32_t base = 0u;
#pragma unroll
    for (int w = 0; w < 8; w++)
        if ((uint32_t)w < warp) base += smem[w];
    __syncthreads();
    return inc - v + base;
}

template<bool HAS_VALS>
__global__ void __launch_bounds__(SORT_THREADS, 3)
    k_sort_onesweep(const uint64_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in, uint64_t *__restrict__ keys_out,
                    uint32_t *__restrict__ vals_out, uint32_t n, uint32_t shift, uint32_t ntiles, const uint32_t *__restrict__ hist /*[256] of this pass*/,
                    uint32_t *ticket, uint32_t *status /*[ntiles][256] of this pass*/) {
    __shared__ uint32_t warp_cnt[SORT_THREADS / 32][256];// per warp digit counts, then per (warp, digit) tile-local bases
    __shared__ uint32_t tile_off[256];                    // first tile-local slot of each digit
    __shared__ uint32_t gbase[256];                       // global position of the digit's first key of this tile
    __shared__ uint32_t scan_tmp[8];
    __shared__ uint32_t s_tile;
    __shared__ uint64_t skeys[SORT_TILE_KEYS];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
#pragma unroll
    for (int w = 0; w < SORT_THREADS / 32; w++) warp_cnt[w][threadIdx.x] = 0u;
    __syncthreads();
    const uint32_t tile = s_tile;
    // warp `warp` owns SORT_ITEMS * 32 consecutive keys; round r covers 32 consecutive keys => stable
    const uint32_t tile_base = tile * SORT_TILE_KEYS, base = tile_base + warp * (SORT_ITEMS * 32);
    const uint32_t tile_n = n - tile_base < (uint32_t)SORT_TILE_KEYS ? n - tile_base : (uint32_t)SORT_TILE_KEYS;
    uint64_t key[SORT_ITEMS];
    uint32_t val[HAS_VALS ? SORT_ITEMS : 1];
    uint32_t rank[SORT_ITEMS];
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; r++) {
        const uint32_t i = base + r * 32 + lane;
        const bool valid = i < n;
        key[r] = valid ? keys_in[i] : ~0ull;
        if (HAS_VALS) val[r] = valid ? vals_in[i] : 0u;
    }
    const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; r++) {
        const uint32_t i = base + r * 32 + lane;
        const bool valid = i < n;
        const uint32_t d = valid ? ((uint32_t)(key[r] >> shift) & 255u) : 0u;
        // lanes holding the same digit: eight ballots (one per digit bit) — constant cost, where MATCH.ANY takes one
        // round per distinct value in the warp (up to 32 once the keys are shuffled by the earlier passes)
        uint32_t peers = __ballot_sync(0xffffffffu, valid);
#pragma unroll
        for (int b = 0; b < 8; b++) {
            const uint32_t m = __ballot_sync(0xffffffffu, (d >> b) & 1u);
            peers &= ((d >> b) & 1u) ? m : ~m;
        }
        uint32_t prev = 0u;
        if (valid) prev = warp_cnt[warp][d];
        __syncwarp();
        rank[r] = prev + __popc(peers & lt_mask);
        if (valid && (peers & lt_mask) == 0u) warp_cnt[warp][d] = prev + __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    // ---- thread d: this tile's count of digit d; per-warp bases inside the digit's run
    const uint32_t d = threadIdx.x;
    uint32_t count = 0u;
#pragma unroll
    for (int w = 0; w < SORT_THREADS / 32; w++) {
        const uint32_t c = warp_cnt[w][d];
        warp_cnt[w][d] = count;
        count += c;
    }
    // publish the aggregate first: later tiles can move on while this one is still looking back
    uint32_t *my_status = status + (uint64_t)tile * 256u;
    if (tile == 0u) {
        __stcg(&my_status[d], SORT_FLAG_INCLUSIVE | count);
    } else {
        __stcg(&my_status[d], SORT_FLAG_AGGREGATE | count);
    }
    const uint32_t off = block_exclusive_scan_256(count, scan_tmp);// first tile-local slot of digit d
    tile_off[d] = off;
    // global exclusive prefix of the digit histogram (256 values, every block recomputes it: 1 KB from L2)
    const uint32_t digit_start = block_exclusive_scan_256(hist[d], scan_tmp);
    // ---- decoupled look-back: keys of digit d in the tiles before this one
    uint32_t before = 0u;
    if (tile != 0u) {
        for (int32_t t = (int32_t)tile - 1; t >= 0;) {
            const uint32_t s = __ldcg(&status[(uint64_t)t * 256u + d]);
            if (s & SORT_FLAG_INCLUSIVE) {
                before += s & SORT_VALUE_MASK;
                break;
            }
            if (s & SORT_FLAG_AGGREGATE) {
                before += s & SORT_VALUE_MASK;
                t--;
                continue;
            }
            __nanosleep(32);// the tile holds a ticket, so it is running: its counts arrive
        }
        __stcg(&my_status[d], SORT_FLAG_INCLUSIVE | (before + count));
    }
    gbase[d] = digit_start + before;
    __syncthreads();
    // ---- reorder the tile by digit in shared memory
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; r++) {
        const uint32_t i = base + r * 32 + lane;
        if (i < n) {
            const uint32_t dd = (uint32_t)(key[r] >> shift) & 255u;
            rank[r] = tile_off[dd] + warp_cnt[warp][dd] + rank[r];// tile-local slot
            skeys[rank[r]] = key[r];
        }
    }
    __syncthreads();
    // ---- write: consecutive threads hold consecutive slots, i.e. consecutive addresses within a digit's run
    uint32_t pos[HAS_VALS ? SORT_ITEMS : 1];
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; r++) {
        const uint32_t slot = r * SORT_THREADS + threadIdx.x;
        if (slot < tile_n) {
            const uint64_t k = skeys[slot];
            const uint32_t dd = (uint32_t)(k >> shift) & 255u;
            const uint32_t q = gbase[dd] + (slot - tile_off[dd]);
            keys_out[q] = k;
            if (HAS_VALS) pos[r] = q;
        }
    }
    if (HAS_VALS) {
        // the payload takes the same route through the same shared-memory tile (a second buffer would not fit 48 KB)
        __syncthreads();
        uint32_t *svals = reinterpret_cast<uint32_t *>(skeys);
#pragma unroll
        for (int r = 0; r < SORT_ITEMS; r++)
            if (base + r * 32 + lane < n) svals[rank[r]] = val[r];
        __syncthreads();
#pragma unroll
        for (int r = 0; r < SORT_ITEMS; r++) {
            const uint32_t slot = r * SORT_THREADS + threadIdx.x;
            if (slot < tile_n) vals_out[pos[r]] = svals[slot];
        }
    }
}

struct SortBuffers {
    uint64_t *keys[2];
    uint32_t *vals[2];// both null: keys only
    uint32_t *control;// sort_control_bytes(n) bytes
};

inline uint32_t sort_tiles(uint32_t n) { return (uint32_t)(((uint64_t)(n ? n : 1) + SORT_TILE_KEYS - 1) / SORT_TILE_KEYS); }
inline uint64_t sort_control_bytes(uint32_t n) {
    return ((uint64_t)SORT_MAX_PASSES * 256u + SORT_MAX_PASSES + 64u + (uint64_t)SORT_MAX_PASSES * sort_tiles(n) * 256u) * 4u;
}

// Sorts by the key bits [first_bit, first_bit + key_bits).  Returns the index (0/1) of the buffer pair holding the result,
// or -1 for an unsupported size.
inline int radix_sort(cudaStream_t stream, const SortBuffers &b, uint32_t n, uint32_t first_bit, uint32_t key_bits, int sm_count) {
    if (n == 0) return 0;
    if (n > SORT_MAX_KEYS) return -1;
    const uint32_t passes = (key_bits + 7u) / 8u;
    if (passes == 0u) return 0;
    if (passes > (uint32_t)SORT_MAX_PASSES) return -1;
    const uint32_t ntiles = sort_tiles(n);
    SortControl c;
    c.hist = b.control, c.ticket = c.hist + SORT_MAX_PASSES * 256, c.status = c.ticket + SORT_MAX_PASSES + 64;
    const uint64_t used = ((uint64_t)SORT_MAX_PASSES * 256u + SORT_MAX_PASSES + 64u + (uint64_t)passes * ntiles * 256u) * 4u;
    if (cudaMemsetAsync(b.control, 0, used, stream) != cudaSuccess) return -1;
    uint32_t hb = (n + SORT_THREADS * 8 - 1) / (SORT_THREADS * 8);
    const uint32_t cap = (uint32_t)(sm_count > 0 ? sm_count : 148) * 8u;
    if (hb > cap) hb = cap;
    k_sort_histograms<<<hb, SORT_THREADS, 0, stream>>>(b.keys[0], n, first_bit, passes, c.hist);
    int cur = 0;
    for (uint32_t p = 0; p < passes; p++) {
        const uint32_t shift = first_bit + 8u * p;
        if (b.vals[0])
            k_sort_onesweep<true><<<ntiles, SORT_THREADS, 0, stream>>>(b.keys[cur], b.vals[cur], b.keys[cur ^ 1], b.vals[cur ^ 1], n, shift, ntiles,
                                                                       c.hist + p * 256u, c.ticket + p, c.status + (uint64_t)p * ntiles * 256u);
        else
            k_sort_onesweep<false><<<ntiles, SORT_THREADS, 0, stream>>>(b.keys[cur], nullptr, b.keys[cur ^ 1], nullptr, n, shift, ntiles, c.hist + p * 256u,
                                                                        c.ticket + p, c.status + (uint64_t)p * ntiles * 256u);
        cur ^= 1;
    }
    return cur;
}

}// namespace lcrt
