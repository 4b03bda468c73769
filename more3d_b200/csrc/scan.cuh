// Block scan and decoupled look-back helpers shared by the scans, the radix sort and the LOD-tree partitions.
#pragma once
#include "common.cuh"

namespace m3d {

constexpr int SC_THREADS = 256;
constexpr int SC_ITEMS = 8;
constexpr int SC_TILE = SC_THREADS * SC_ITEMS;
constexpr uint64_t LB_AGG = 1ull << 62, LB_PREFIX = 2ull << 62, LB_VAL = (1ull << 62) - 1ull;

template <typename T>
__device__ __forceinline__ T block_exclusive_scan(T v, T* total, T* smem /* 33 entries */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        T t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) smem[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        T w = (lane < (int)(blockDim.x >> 5)) ? smem[lane] : T(0);
        T winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            T t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        smem[lane] = winc - w;  // exclusive warp offsets
        if (lane == 31) smem[32] = winc;
    }
    __syncthreads();
    T res = smem[warp] + inc - v;
    if (total) *total = smem[32];
    __syncthreads();
    return res;
}

// sum of the aggregates of tiles tile-1, tile-2, ... down to (and including) the first one that carries an
// inclusive prefix; executed by one full warp
__device__ __forceinline__ uint64_t lookback_prefix(const volatile uint64_t* desc, int tile, int lane) {
    uint64_t prefix = 0;
    for (int j = tile - 1;; j -= 32) {
        const int idx = j - lane;
        uint64_t d = LB_PREFIX;                    // before the first tile: prefix 0
        for (;;) {
            if (idx >= 0) d = desc[idx];
            if (!__any_sync(0xffffffffu, (d >> 62) == 0)) break;
        }
        const unsigned pmask = __ballot_sync(0xffffffffu, (d >> 62) == 2);
        const int first = pmask ? __ffs(pmask) - 1 : 32;
        uint64_t val = (lane <= first) ? (d & LB_VAL) : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
        prefix += val;
        if (pmask) return prefix;
    }
}


// lanes of the warp whose 8-bit digit equals this lane's (invalid lanes match nobody): eight ballots instead of
// __match_any_sync, whose cost on sm_100 grows with the number of DISTINCT values in the warp (it resolves one
// value per internal round; low radix digits are all different)
__device__ __forceinline__ unsigned match_digit(unsigned d, bool valid) {
    unsigned peers = __ballot_sync(0xffffffffu, valid);
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        const unsigned m = __ballot_sync(0xffffffffu, (d >> b) & 1u);
        peers &= ((d >> b) & 1u) ? m : ~m;
    }
    return valid ? peers : 0u;
}

}  // namespace m3d
