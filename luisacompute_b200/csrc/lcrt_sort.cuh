// lcrt_sort.cuh — CUB-free LSD radix sort of (u64 key, u32 value) pairs, 8-bit digits.
//
// Per pass: (1) k_sort_hist: per-tile digit histogram, written digit-major so that ONE exclusive scan over the
// flattened [digit][tile] table yields every tile's global scatter base; (2) a three-kernel exclusive scan;
// (3) k_sort_scatter: stable scatter — each warp ranks 32 consecutive keys per round (same-digit peers from eight
// ballots), warp totals are combined across the tile's eight warps in shared memory.
//
// HBM traffic per pass and key: 8 B (hist read) + 12 B (read) + 12 B (write); the table is n/4096 * 1 KiB.
// The reference has no sort of its own (its optional `lcub` extension wraps cub::DeviceRadixSort,
// src/backends/cuda/lcub/private/dcub/device_radix_sort.cu) — north_star asks for a CUB-free one.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lcrt {

constexpr int SORT_THREADS = 256;
constexpr int SORT_ITEMS = 16;                          // keys per thread
constexpr int SORT_TILE_KEYS = SORT_THREADS * SORT_ITEMS;// 4096 == SORT_TILE in lcrt_layout.h
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;// 2048

__global__ void __launch_bounds__(SORT_THREADS) k_sort_hist(const uint64_t *__restrict__ keys, uint32_t n, uint32_t shift, uint32_t ntiles,
                                                            uint32_t *__restrict__ table) {
    __shared__ uint32_t hist[256];
    hist[threadIdx.x] = 0u;
    __syncthreads();
    const uint32_t base = blockIdx.x * SORT_TILE_KEYS;
#pragma unroll 4
    for (int r = 0; r < SORT_ITEMS; r++) {
        const uint32_t i = base + r * SORT_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&hist[(uint32_t)(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    table[threadIdx.x * ntiles + blockIdx.x] = hist[threadIdx.x];
}

// ---- exclusive scan of `m` u32 values (in place): tile sums -> scan of sums (single block) -> rescan with offsets
__device__ __forceinline__ uint32_t block_exclusive_scan_256(uint32_t v, uint32_t *total, uint32_t *smem /*[9]*/) {
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (uint32_t)o) inc += t;
    }
    if (lane == 31u) smem[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < 8u ? smem[lane] : 0u;
        uint32_t winc = w;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= (uint32_t)o) winc += t;
        }
        if (lane < 8u) smem[lane] = winc - w;
        if (lane == 7u) smem[8] = winc;
    }
    __syncthreads();
    const uint32_t r = inc - v + smem[warp];
    *total = smem[8];
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_sums(const uint32_t *__restrict__ data, uint32_t m, uint32_t *__restrict__ sums) {
    __shared__ uint32_t smem[9];
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++)
        if (base + k < m) s += data[base + k];
    uint32_t total;
    block_exclusive_scan_256(s, &total, smem);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_top(uint32_t *__restrict__ sums, uint32_t count) {
    __shared__ uint32_t smem[9];
    uint32_t carry = 0;
    for (uint32_t base = 0; base < count; base += SCAN_THREADS) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < count ? sums[i] : 0u;
        uint32_t total;
        const uint32_t ex = block_exclusive_scan_256(v, &total, smem);
        if (i < count) sums[i] = ex + carry;
        carry += total;
    }
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(uint32_t *__restrict__ data, uint32_t m, const uint32_t *__restrict__ sums) {
    __shared__ uint32_t smem[9];
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        v[k] = base + k < m ? data[base + k] : 0u;
        s += v[k];
    }
    uint32_t total;
    uint32_t run = block_exclusive_scan_256(s, &total, smem) + sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        if (base + k < m) data[base + k] = run;
        run += v[k];
    }
}

// (A variant that first sorts the tile by digit inside shared memory and writes each digit's run contiguously was
// measured on B200: 103 vs 107 us per pass at 5 M keys, 41 vs 35 us at 1.3 M — the extra barriers and the lower
// occupancy cost what the coalescing gains, so the direct scatter stays.)
__global__ void __launch_bounds__(SORT_THREADS)
    k_sort_scatter(const uint64_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in, uint64_t *__restrict__ keys_out,
                   uint32_t *__restrict__ vals_out, uint32_t n, uint32_t shift, uint32_t ntiles, const uint32_t *__restrict__ table) {
    __shared__ uint32_t warp_cnt[SORT_THREADS / 32][256];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
#pragma unroll
    for (int w = 0; w < SORT_THREADS / 32; w++) warp_cnt[w][threadIdx.x] = 0u;
    __syncthreads();
    // warp `warp` owns SORT_ITEMS * 32 consecutive keys; round r covers 32 consecutive keys => stable
    const uint32_t base = blockIdx.x * SORT_TILE_KEYS + warp * (SORT_ITEMS * 32);
    uint64_t key[SORT_ITEMS];
    uint32_t val[SORT_ITEMS];
    uint32_t rank[SORT_ITEMS];
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; r++) {
        const uint32_t i = base + r * 32 + lane;
        const bool valid = i < n;
        key[r] = valid ? keys_in[i] : 0ull;
        val[r] = valid ? vals_in[i] : 0u;
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
    {
        // thread d: exclusive scan of digit d over the warps, seeded with this tile's global base
        const uint32_t d = threadIdx.x;
        uint32_t run = table[d * ntiles + blockIdx.x];
#pragma unroll
        for (int w = 0; w < SORT_THREADS / 32; w++) {
            const uint32_t c = warp_cnt[w][d];
            warp_cnt[w][d] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; r++) {
        const uint32_t i = base + r * 32 + lane;
        if (i < n) {
            const uint32_t d = (uint32_t)(key[r] >> shift) & 255u;
            const uint32_t pos = warp_cnt[warp][d] + rank[r];
            keys_out[pos] = key[r];
            vals_out[pos] = val[r];
        }
    }
}

struct SortBuffers {
    uint64_t *keys[2];
    uint32_t *vals[2];
    uint32_t *table;// [256 * ntiles]
    uint32_t *sums; // [ceil(256 * ntiles / SCAN_TILE)]
};

inline uint64_t sort_table_bytes(uint32_t n) {
    const uint64_t ntiles = ((uint64_t)(n ? n : 1) + SORT_TILE_KEYS - 1) / SORT_TILE_KEYS;
    const uint64_t m = 256 * ntiles;
    const uint64_t nsums = (m + SCAN_TILE - 1) / SCAN_TILE;
    return (m + nsums + 64) * 4;
}

// Sorts by the low `key_bits` bits.  Returns the index (0/1) of the buffer pair holding the result.
inline int radix_sort_pairs(cudaStream_t stream, const SortBuffers &b, uint32_t n, uint32_t key_bits) {
    if (n == 0) return 0;
    const uint32_t ntiles = (n + SORT_TILE_KEYS - 1) / SORT_TILE_KEYS;
    const uint32_t m = 256u * ntiles;
    const uint32_t nsums = (m + SCAN_TILE - 1) / SCAN_TILE;
    int cur = 0;
    for (uint32_t shift = 0; shift < key_bits; shift += 8) {
        k_sort_hist<<<ntiles, SORT_THREADS, 0, stream>>>(b.keys[cur], n, shift, ntiles, b.table);
        k_scan_sums<<<nsums, SCAN_THREADS, 0, stream>>>(b.table, m, b.sums);
        k_scan_top<<<1, SCAN_THREADS, 0, stream>>>(b.sums, nsums);
        k_scan_apply<<<nsums, SCAN_THREADS, 0, stream>>>(b.table, m, b.sums);
        k_sort_scatter<<<ntiles, SORT_THREADS, 0, stream>>>(b.keys[cur], b.vals[cur], b.keys[cur ^ 1], b.vals[cur ^ 1], n, shift, ntiles, b.table);
        cur ^= 1;
    }
    return cur;
}

}// namespace lcrt
