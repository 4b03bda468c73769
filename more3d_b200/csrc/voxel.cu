// K8: voxel-grid downsampling = fp64 voxel key -> stable radix sort -> segmented mean.
// Replaces open3d voxel_down_sample as called by the reference (Downsampling/Downsample.py:86-89):
// voxel index = floor((p - (min_bound - voxel/2)) / voxel) per axis, output = mean of the member points,
// members accumulated in original point order (the stable sort keeps it, so the fp64 sums are the ones a
// sequential host loop produces).
#include "common.cuh"

namespace m3d {

constexpr int VX_THREADS = 256;

// min/max of x, y, z: reuse the pattern of grid.cu with one fmin tree over (v, -v)
__global__ void __launch_bounds__(VX_THREADS) vx_bbox_partial(const double* __restrict__ xyz, int64_t n,
                                                              double* __restrict__ part) {
    double v[6] = {INFINITY, INFINITY, INFINITY, INFINITY, INFINITY, INFINITY};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double c = xyz[3 * i + a];
            v[a] = fmin(v[a], c);
            v[3 + a] = fmin(v[3 + a], -c);
        }
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[a] = fmin(v[a], __shfl_xor_sync(0xffffffffu, v[a], o));
    __shared__ double sm[6][VX_THREADS / 32];
    if ((threadIdx.x & 31) == 0)
        for (int a = 0; a < 6; ++a) sm[a][threadIdx.x >> 5] = v[a];
    __syncthreads();
    if (threadIdx.x < 6) {
        double r = sm[threadIdx.x][0];
        for (int w = 1; w < VX_THREADS / 32; ++w) r = fmin(r, sm[threadIdx.x][w]);
        part[6 * blockIdx.x + threadIdx.x] = r;
    }
}
__global__ void vx_bbox_final(const double* __restrict__ part, int nparts, double* __restrict__ out) {
    double v[6] = {INFINITY, INFINITY, INFINITY, INFINITY, INFINITY, INFINITY};
    for (int i = threadIdx.x; i < nparts; i += 32)
#pragma unroll
        for (int a = 0; a < 6; ++a) v[a] = fmin(v[a], part[6 * i + a]);
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[a] = fmin(v[a], __shfl_xor_sync(0xffffffffu, v[a], o));
    if (threadIdx.x == 0)
        for (int a = 0; a < 6; ++a) out[a] = (a < 3) ? v[a] : -v[a];
}

struct VoxelGeom {
    double vmin[3];
    double voxel;
    long long ny, nz;   // key = (kx * ny + ky) * nz + kz
};

__device__ __forceinline__ void voxel_index(const double* __restrict__ xyz, int64_t i, const VoxelGeom& vg,
                                            long long k[3]) {
#pragma unroll
    for (int a = 0; a < 3; ++a)
        k[a] = (long long)floor(__ddiv_rn(__dsub_rn(xyz[3 * i + a], vg.vmin[a]), vg.voxel));
}

template <typename K>
__global__ void __launch_bounds__(VX_THREADS) vx_keys(const double* __restrict__ xyz, int64_t n, VoxelGeom vg,
                                                      K* __restrict__ keys, uint32_t* __restrict__ vals) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    long long k[3];
    voxel_index(xyz, i, vg, k);
    keys[i] = (K)((k[0] * vg.ny + k[1]) * vg.nz + k[2]);
    vals[i] = (uint32_t)i;
}

template <typename K>
__global__ void __launch_bounds__(VX_THREADS) vx_heads(const K* __restrict__ keys, int64_t n,
                                                       uint32_t* __restrict__ head) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    head[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
}

// seg[i] = exclusive scan of head + head[i] - 1 = voxel row of sorted position i; starts[row] = i at heads
template <typename K>
__global__ void __launch_bounds__(VX_THREADS) vx_starts(const K* __restrict__ keys,
                                                        const uint32_t* __restrict__ scan, int64_t n,
                                                        uint32_t* __restrict__ starts) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (i == 0 || keys[i] != keys[i - 1]) starts[scan[i]] = (uint32_t)i;
    if (i == n - 1) starts[scan[n]] = (uint32_t)n;
}

// One WARP per voxel.  The members' coordinates are gathered 32 at a time (parallel, latency-hidden loads) and then
// added ONE AFTER THE OTHER in original point order through lane shuffles, so the fp64 sums are exactly those of the
// sequential host loop (np.add.at / open3d's per-voxel accumulator) -- a tree reduction would round differently.
template <typename K>
__global__ void __launch_bounds__(VX_THREADS) vx_means(const double* __restrict__ xyz,
                                                       const K* __restrict__ keys,
                                                       const uint32_t* __restrict__ order,
                                                       const uint32_t* __restrict__ starts, int64_t nvox,
                                                       VoxelGeom vg, double* __restrict__ out_xyz,
                                                       int64_t* __restrict__ out_keys,
                                                       int32_t* __restrict__ voxel_of_point) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t v = (((int64_t)blockIdx.x * blockDim.x) + threadIdx.x) >> 5; v < nvox; v += warps) {
        const uint32_t b = starts[v], e = starts[v + 1];
        double sx = 0.0, sy = 0.0, sz = 0.0;
        for (uint32_t t0 = b; t0 < e; t0 += 32) {
            const uint32_t t = t0 + lane;
            double x = 0.0, y = 0.0, z = 0.0;
            if (t < e) {
                const int64_t o = order[t];
                x = xyz[3 * o]; y = xyz[3 * o + 1]; z = xyz[3 * o + 2];
                if (voxel_of_point) voxel_of_point[o] = (int32_t)v;
            }
            const int cnt = (int)min(32u, e - t0);
            for (int l = 0; l < cnt; ++l) {            // original point order inside the voxel
                sx += __shfl_sync(0xffffffffu, x, l);
                sy += __shfl_sync(0xffffffffu, y, l);
                sz += __shfl_sync(0xffffffffu, z, l);
            }
        }
        if (lane == 0) {
            const double cnt = (double)(e - b);
            out_xyz[3 * v] = sx / cnt; out_xyz[3 * v + 1] = sy / cnt; out_xyz[3 * v + 2] = sz / cnt;
            if (out_keys) {
                const long long key = (long long)keys[b];
                const long long kz = key % vg.nz, kxy = key / vg.nz;
                out_keys[3 * v] = kxy / vg.ny; out_keys[3 * v + 1] = kxy % vg.ny; out_keys[3 * v + 2] = kz;
            }
        }
    }
}

}  // namespace m3d

using namespace m3d;

static int voxel_downsample_impl(const double* xyz, int64_t n, double voxel, const double* bounds_h, double* out_xyz,
                                 int64_t* out_keys, int32_t* voxel_of_point, int64_t* nout_h, void* stream);


template <typename K>
static int sort_pairs(K* k1, uint32_t* v1, K* k2, uint32_t* v2, int64_t n, int bits, K** ks, uint32_t** vs, cudaStream_t s);
template <>
int sort_pairs<uint32_t>(uint32_t* k1, uint32_t* v1, uint32_t* k2, uint32_t* v2, int64_t n, int bits, uint32_t** ks, uint32_t** vs, cudaStream_t s) {
    return radix_sort_pairs_u32(k1, v1, k2, v2, n, bits, ks, vs, s);
}
template <>
int sort_pairs<uint64_t>(uint64_t* k1, uint32_t* v1, uint64_t* k2, uint32_t* v2, int64_t n, int bits, uint64_t** ks, uint32_t** vs, cudaStream_t s) {
    return radix_sort_pairs_u64(k1, v1, k2, v2, n, bits, ks, vs, s);
}

template <typename K>
static int voxel_reduce(const double* xyz, int64_t n, const VoxelGeom& vg, int bits, double* out_xyz, int64_t* out_keys,
                        int32_t* voxel_of_point, int64_t* nout_h, cudaStream_t s) {
    K *k1 = nullptr, *k2 = nullptr, *ks = nullptr;
    uint32_t *v1 = nullptr, *v2 = nullptr, *vs = nullptr, *head = nullptr, *starts = nullptr;
    M3D_TRY(dev_alloc(&k1, (size_t)n, s));
    M3D_TRY(dev_alloc(&k2, (size_t)n, s));
    M3D_TRY(dev_alloc(&v1, (size_t)n, s));
    M3D_TRY(dev_alloc(&v2, (size_t)n, s));
    M3D_TRY(dev_alloc(&head, (size_t)n + 1, s));
    M3D_TRY(dev_alloc(&starts, (size_t)n + 1, s));
    const unsigned nb = (unsigned)((n + VX_THREADS - 1) / VX_THREADS);
    vx_keys<K><<<nb, VX_THREADS, 0, s>>>(xyz, n, vg, k1, v1);
    count_launches(1);
    {
        ProfScope p2("voxel_sort", s);
        M3D_TRY(sort_pairs<K>(k1, v1, k2, v2, n, bits, &ks, &vs, s));
    }
    vx_heads<K><<<nb, VX_THREADS, 0, s>>>(ks, n, head);
    M3D_TRY(exclusive_scan_u32(head, head, n, s));
    vx_starts<K><<<nb, VX_THREADS, 0, s>>>(ks, head, n, starts);
    count_launches(2);
    M3D_CHECK_LAUNCH();
    uint32_t nvox = 0;
    M3D_CUDA(cudaMemcpyAsync(&nvox, head + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    M3D_CUDA(cudaStreamSynchronize(s));
    {
        ProfScope p3("voxel_means", s);
        vx_means<K><<<(unsigned)min((int64_t)(((int64_t)nvox * 32 + VX_THREADS - 1) / VX_THREADS), (int64_t)148 * 64), VX_THREADS, 0, s>>>(
            xyz, ks, vs, starts, (int64_t)nvox, vg, out_xyz, out_keys, voxel_of_point);
    }
    count_launches(1);
    M3D_CHECK_LAUNCH();
    dev_free(k1, s); dev_free(k2, s); dev_free(v1, s); dev_free(v2, s); dev_free(head, s); dev_free(starts, s);
    *nout_h = (int64_t)nvox;
    return MORE3D_OK;
}

extern "C" int more3d_voxel_downsample(const double* xyz, int64_t n, double voxel, double* out_xyz,
                                       int64_t* out_keys, int32_t* voxel_of_point, int64_t* nout_h,
                                       void* stream) {
    return voxel_downsample_impl(xyz, n, voxel, nullptr, out_xyz, out_keys, voxel_of_point, nout_h, stream);
}

extern "C" int more3d_voxel_downsample_bounds(const double* xyz, int64_t n, double voxel, const double* bounds_h,
                                              double* out_xyz, int64_t* out_keys, int32_t* voxel_of_point,
                                              int64_t* nout_h, void* stream) {
    M3D_REQUIRE(bounds_h != nullptr, "bounds_h is NULL");
    return voxel_downsample_impl(xyz, n, voxel, bounds_h, out_xyz, out_keys, voxel_of_point, nout_h, stream);
}

static int voxel_downsample_impl(const double* xyz, int64_t n, double voxel, const double* bounds_h, double* out_xyz,
                                 int64_t* out_keys, int32_t* voxel_of_point, int64_t* nout_h, void* stream) {
    M3D_REQUIRE(xyz && out_xyz && nout_h && n > 0, "NULL / empty argument");
    M3D_REQUIRE(voxel > 0 && isfinite(voxel), "voxel_size must be > 0");   // open3d raises on <= 0
    if (n >= ((int64_t)1 << 32)) { set_error("n >= 2^32 not supported"); return MORE3D_E_RANGE; }
    cudaStream_t s = (cudaStream_t)stream;
    init_device_pool();
    ProfScope prof("voxel_downsample", s);

    double bb[6];
    if (bounds_h) {
        // the bounding box of the WHOLE cloud, supplied by a caller that holds one shard of it: every shard then
        // bins on the same voxel lattice and produces the same voxel indices as the unsharded call
        for (int a = 0; a < 6; ++a) bb[a] = bounds_h[a];
    } else {
        const int nparts = 592;
        double* part = nullptr;
        M3D_TRY(dev_alloc(&part, (size_t)6 * nparts + 6, s));
        vx_bbox_partial<<<nparts, VX_THREADS, 0, s>>>(xyz, n, part);
        vx_bbox_final<<<1, 32, 0, s>>>(part, nparts, part + 6 * nparts);
        count_launches(2);
        M3D_CHECK_LAUNCH();
        M3D_CUDA(cudaMemcpyAsync(bb, part + 6 * nparts, sizeof(bb), cudaMemcpyDeviceToHost, s));
        M3D_CUDA(cudaStreamSynchronize(s));
        dev_free(part, s);
    }
    for (int a = 0; a < 6; ++a)
        if (!isfinite(bb[a])) { set_error("non-finite coordinates in cloud"); return MORE3D_E_INVALID; }

    VoxelGeom vg;
    vg.voxel = voxel;
    double dims[3];
    for (int a = 0; a < 3; ++a) {
        vg.vmin[a] = bb[a] - voxel * 0.5;              // min_bound - voxel_size * 0.5
        dims[a] = floor((bb[3 + a] - vg.vmin[a]) / voxel) + 2.0;
    }
    if (dims[0] * dims[1] * dims[2] >= 9.0e18) {
        set_error("voxel_size is too small for this cloud's extent (voxel index overflow)");   // open3d: "voxel_size is too small."
        return MORE3D_E_RANGE;
    }
    vg.ny = (long long)dims[1];
    vg.nz = (long long)dims[2];
    int bits = 1;
    {
        const double total = dims[0] * dims[1] * dims[2];
        while (bits < 63 && ldexp(1.0, bits) < total) ++bits;
    }

    // 32-bit keys whenever the voxel lattice has fewer than 2^32 cells (terrain tiles always do): 8 instead of
    // 12 bytes per pair and pass through the radix sort
    if (bits <= 32) return voxel_reduce<uint32_t>(xyz, n, vg, bits, out_xyz, out_keys, voxel_of_point, nout_h, s);
    return voxel_reduce<uint64_t>(xyz, n, vg, bits, out_xyz, out_keys, voxel_of_point, nout_h, s);
}
