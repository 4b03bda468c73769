// K13b: the per-step strip kernels of the multi-GPU saliency path (no reference equivalent; OPALS tile streaming,
// DM_saliency.py:76-86, is the analogue):
//   more3d_strip_pack       both halo bands of a strip packed in ONE pass over the cloud (stable, look-back scan),
//                           together with the strip's bounding box -- everything stays on the device
//   more3d_histogram256_range   numpy.histogram(v, 256, range=(lo, hi)) with lo / hi read from DEVICE memory (the
//                           all-reduced saliency extrema), so the Otsu chain needs no host round trip
#include "common.cuh"
#include "scan.cuh"

namespace m3d {

constexpr int SP_THREADS = 256;
constexpr int SP_ITEMS = 8;
constexpr int SP_TILE = SP_THREADS * SP_ITEMS;

// info[0..5] = {min x, min y, min z, max x, max y, max z} (as ordered-uint64 atomics while the kernel runs, doubles
// after strip_pack_finish), info[6] = points written to the left band, info[7] = to the right band
__device__ __forceinline__ unsigned long long ord64(double v) {
    const unsigned long long u = (unsigned long long)__double_as_longlong(v);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double unord64(unsigned long long u) {
    return __longlong_as_double((long long)((u >> 63) ? (u & 0x7fffffffffffffffull) : ~u));
}

__global__ void strip_pack_init(unsigned long long* info) {
    if (threadIdx.x < 3) info[threadIdx.x] = ~0ull;
    else if (threadIdx.x < 8) info[threadIdx.x] = 0ull;
}

// Warp w of a tile owns 256 consecutive points and reads them in 8 rounds of 32 (lanes 24 bytes apart: coalesced);
// a ballot per round and band gives every selected point its rank, so the packed bands keep the cloud's order.
__global__ void __launch_bounds__(SP_THREADS) strip_pack_kernel(const double* __restrict__ xyz, int64_t n,
                                                                double left_below, double right_from,
                                                                double* __restrict__ left_out,
                                                                double* __restrict__ right_out, int64_t cap,
                                                                uint64_t* desc, unsigned* ticket, int ntiles,
                                                                unsigned long long* info) {
    __shared__ uint64_t sm[33];
    __shared__ uint64_t s_prefix;
    __shared__ unsigned s_tile;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const int tile = (int)s_tile;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const int64_t wbase = (int64_t)tile * SP_TILE + (int64_t)warp * (32 * SP_ITEMS);
    unsigned ml[SP_ITEMS], mr[SP_ITEMS];              // ballots of round j: lanes in the left / right band
    unsigned cl = 0, cr = 0;                          // warp totals
    double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    // the coordinates are NOT kept in registers across the scan (24 doubles per thread cost two thirds of the resident
    // CTAs): the band members -- under 1 % of the points -- re-read theirs from L1 / L2 when they are written out
#pragma unroll
    for (int j = 0; j < SP_ITEMS; ++j) {
        const int64_t i = wbase + j * 32 + lane;
        bool fl = false, fr = false;
        if (i < n) {
            const double x = __ldg(xyz + 3 * i), y = __ldg(xyz + 3 * i + 1), z = __ldg(xyz + 3 * i + 2);
            fl = x < left_below;
            fr = x >= right_from;
            lo[0] = fmin(lo[0], x); lo[1] = fmin(lo[1], y); lo[2] = fmin(lo[2], z);
            hi[0] = fmax(hi[0], x); hi[1] = fmax(hi[1], y); hi[2] = fmax(hi[2], z);
        }
        ml[j] = __ballot_sync(0xffffffffu, fl);
        mr[j] = __ballot_sync(0xffffffffu, fr);
        cl += __popc(ml[j]);
        cr += __popc(mr[j]);
    }
    // both band counts travel in one 62-bit scan value: left in bits 0..30, right in bits 31..61.  Thread = (warp,
    // lane): lane 0 of every warp carries the warp's totals, the block scan then yields the warps' exclusive offsets.
    const uint64_t mine = lane == 0 ? ((uint64_t)cl | ((uint64_t)cr << 31)) : 0ull;
    uint64_t total;
    uint64_t ex = block_exclusive_scan<uint64_t>(mine, &total, sm);
    ex = __shfl_sync(0xffffffffu, ex, 0);             // the warp's offset inside the tile
    volatile uint64_t* vd = desc;
    if (threadIdx.x == 0) {
        vd[tile] = (tile == 0 ? LB_PREFIX : LB_AGG) | total;
        if (tile == 0) s_prefix = 0;
    }
    if (tile > 0 && threadIdx.x < 32) {
        const uint64_t p = lookback_prefix(vd, tile, threadIdx.x);
        if (threadIdx.x == 0) {
            vd[tile] = LB_PREFIX | (p + total);
            s_prefix = p;
        }
    }
    __syncthreads();
    ex += s_prefix;
    int64_t l = (int64_t)(ex & 0x7fffffffull), r = (int64_t)(ex >> 31);
#pragma unroll
    for (int j = 0; j < SP_ITEMS; ++j) {
        const int64_t i = wbase + j * 32 + lane;
        if (ml[j] & (1u << lane)) {
            const int64_t o = l + __popc(ml[j] & lt);
            if (o < cap) { left_out[3 * o] = __ldg(xyz + 3 * i); left_out[3 * o + 1] = __ldg(xyz + 3 * i + 1); left_out[3 * o + 2] = __ldg(xyz + 3 * i + 2); }
        }
        if (mr[j] & (1u << lane)) {
            const int64_t o = r + __popc(mr[j] & lt);
            if (o < cap) { right_out[3 * o] = __ldg(xyz + 3 * i); right_out[3 * o + 1] = __ldg(xyz + 3 * i + 1); right_out[3 * o + 2] = __ldg(xyz + 3 * i + 2); }
        }
        l += __popc(ml[j]);
        r += __popc(mr[j]);
    }
    if (tile == ntiles - 1 && threadIdx.x == 0) {
        const uint64_t all = s_prefix + total;
        info[6] = all & 0x7fffffffull;
        info[7] = all >> 31;
    }
    // bounding box: warp shuffle, then one ordered-integer atomic per warp and coordinate
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = fmin(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = fmax(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            if (lo[a] <= hi[a]) {
                // racy pre-read: after the first tiles hardly any warp still improves a bound, and thousands of
                // same-address atomics would serialise in L2
                const unsigned long long ol = ord64(lo[a]), oh = ord64(hi[a]);
                if (ol < *(volatile unsigned long long*)(info + a)) atomicMin(info + a, ol);
                if (oh > *(volatile unsigned long long*)(info + 3 + a)) atomicMax(info + 3 + a, oh);
            }
        }
    }
}

__global__ void strip_pack_finish(const unsigned long long* info, double* out) {
    if (threadIdx.x < 6) out[threadIdx.x] = unord64(info[threadIdx.x]);
    else if (threadIdx.x < 8) out[threadIdx.x] = (double)info[threadIdx.x];
}

__global__ void negate_odd_kernel(float* v, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && (i & 1)) v[i] = -v[i];
}

template <typename T>
__global__ void __launch_bounds__(256) hist256_range_kernel(const T* __restrict__ v, int64_t n,
                                                            const void* __restrict__ lohi, int lohi_bytes,
                                                            unsigned long long* __restrict__ hist) {
    __shared__ unsigned h[256];
    __shared__ double e[257];
    h[threadIdx.x] = 0;
    const double first = lohi_bytes == 4 ? (double)((const float*)lohi)[0] : ((const double*)lohi)[0];
    const double last = lohi_bytes == 4 ? (double)((const float*)lohi)[1] : ((const double*)lohi)[1];
    // numpy.linspace(lo, hi, 257): step = (hi - lo) / 256, edge i = i * step + lo (separate multiply and add), last = hi
    const double step = __ddiv_rn(__dsub_rn(last, first), 256.0);
    for (int i = threadIdx.x; i < 257; i += 256) e[i] = (i == 256) ? last : __dadd_rn(__dmul_rn((double)i, step), first);
    __syncthreads();
    const double denom = __dsub_rn(last, first);
    const double scale = 256.0 / denom;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t nround = ((n + stride - 1) / stride) * stride;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nround; i += stride) {
        int bin = -1;
        if (i < n && denom > 0.0) {
            const double a = (double)v[i];
            if (a >= first && a <= last) {
                // numpy divides here ((a - first) / (last - first) * 256); a reciprocal multiply lands within one bin of
                // that, and the two edge fix-ups below move any such index to THE bin whose edges bracket a -- the
                // same final bin numpy reaches
                double f = __dmul_rn(__dsub_rn(a, first), scale);
                int b = (int)f;
                if (b > 255) b = 255;
                if (b < 0) b = 0;
                if (a < e[b]) b -= 1;
                if (a >= e[b + 1] && b != 255) b += 1;
                bin = b;
            }
        }
        const unsigned peers = match_digit((unsigned)bin & 255u, bin >= 0);
        if (bin >= 0 && (peers & ((1u << (threadIdx.x & 31)) - 1u)) == 0u) atomicAdd(&h[bin], __popc(peers));
    }
    __syncthreads();
    if (h[threadIdx.x]) atomicAdd(&hist[threadIdx.x], (unsigned long long)h[threadIdx.x]);
}

}  // namespace m3d

using namespace m3d;

extern "C" int more3d_strip_pack(const double* xyz, int64_t n, double left_below, double right_from,
                                 double* left_out, double* right_out, int64_t cap, double* info, void* stream) {
    M3D_REQUIRE(xyz && left_out && right_out && info && n > 0 && cap >= 0, "NULL / empty argument");
    if (n >= ((int64_t)1 << 31)) { set_error("n >= 2^31 not supported"); return MORE3D_E_RANGE; }
    cudaStream_t s = (cudaStream_t)stream;
    init_device_pool();
    ProfScope prof("strip_pack", s);
    const int ntiles = (int)((n + SP_TILE - 1) / SP_TILE);
    uint64_t* work = nullptr;                       // [ntiles descriptors][ticket][8 info words]
    M3D_TRY(dev_alloc(&work, (size_t)ntiles + 1 + 8, s));
    M3D_CUDA(cudaMemsetAsync(work, 0, ((size_t)ntiles + 1) * sizeof(uint64_t), s));
    unsigned long long* iw = (unsigned long long*)(work + ntiles + 1);
    strip_pack_init<<<1, 32, 0, s>>>(iw);
    strip_pack_kernel<<<ntiles, SP_THREADS, 0, s>>>(xyz, n, left_below, right_from, left_out, right_out, cap, work,
                                                    (unsigned*)(work + ntiles), ntiles, iw);
    strip_pack_finish<<<1, 32, 0, s>>>(iw, info);
    count_launches(3);
    M3D_CHECK_LAUNCH();
    dev_free(work, s);
    return MORE3D_OK;
}

extern "C" int more3d_negate_odd(float* v, int64_t n, void* stream) {
    M3D_REQUIRE(v != nullptr && n >= 0 && n < (1 << 20), "bad argument");
    if (n == 0) return MORE3D_OK;
    negate_odd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(v, (int)n);
    count_launches(1);
    M3D_CHECK_LAUNCH();
    return MORE3D_OK;
}

extern "C" int more3d_histogram256_range(const void* v, int elem_bytes, int64_t n, const void* lohi, int lohi_bytes,
                                         int64_t* hist, void* stream) {
    M3D_REQUIRE(elem_bytes == 4 || elem_bytes == 8, "elem_bytes must be 4 (float32) or 8 (float64)");
    M3D_REQUIRE(lohi_bytes == 4 || lohi_bytes == 8, "lohi_bytes must be 4 (float32) or 8 (float64)");
    M3D_REQUIRE(v && lohi && hist && n > 0, "NULL / empty argument");
    cudaStream_t s = (cudaStream_t)stream;
    M3D_CUDA(cudaMemsetAsync(hist, 0, 256 * sizeof(int64_t), s));
    ProfScope prof("histogram", s);
    if (elem_bytes == 4) hist256_range_kernel<float><<<148 * 4, 256, 0, s>>>((const float*)v, n, lohi, lohi_bytes, (unsigned long long*)hist);
    else hist256_range_kernel<double><<<148 * 4, 256, 0, s>>>((const double*)v, n, lohi, lohi_bytes, (unsigned long long*)hist);
    count_launches(1);
    M3D_CHECK_LAUNCH();
    return MORE3D_OK;
}
