// Counter-based synthetic open-terrain generator (bench / test INPUT generation, not part of the reference path):
// point i of a tile is a pure function of (seed, i), so any id range of a 50 M / 500 M-point tile is produced
// directly in HBM in milliseconds, and every partition of the ids over the GPUs is the same global cloud.
// Model = SURVEY.md §8d (7-octave relief + Gaussian noise + one bump / pit per 50 m block, scanner origin 1000 m
// above the tile); numpy twin: more3d_b200/synthetic.py:hash_terrain_host.
#include "common.cuh"

namespace m3d {

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ double u01(uint64_t seed, uint64_t id, uint64_t stream) {
    const uint64_t k = splitmix64(splitmix64(seed ^ (stream * 0xD1B54A32D192ED03ull)) ^ id);
    return (double)(k >> 11) * (1.0 / 9007199254740992.0);
}

__global__ void __launch_bounds__(256) synth_terrain_kernel(uint64_t seed, int64_t id_start, int64_t count,
                                                            double slab_per_id, double slab_w, double lx, double ly,
                                                            double noise, double* __restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const uint64_t id = (uint64_t)(id_start + t);
    const long long slab = (long long)(__dmul_rn(__dadd_rn((double)id, 0.5), slab_per_id));
    const double x = __dmul_rn(__dadd_rn((double)slab, u01(seed, id, 1)), slab_w);
    const double y = __dmul_rn(u01(seed, id, 2), ly);
    double z = 0.0;
#pragma unroll
    for (int o = 0; o < 7; ++o) {
        const double f = (double)(1 << o);
        z += 2.0 * exp2(-0.8 * o) * sin(f * x / 40.0 + o) * cos(f * y / 37.0 + 2 * o);
    }
    const double u1 = u01(seed, id, 3), u2 = u01(seed, id, 4);
    z += noise * sqrt(-2.0 * log(1.0 - u1)) * cos(2.0 * 3.141592653589793 * u2);
    const double B = 50.0;
    const double bx = floor(x / B), by = floor(y / B);
    const uint64_t bid = (uint64_t)((long long)bx * 1000003ll + (long long)by);
    const double cx = bx * B + 10.0 + 30.0 * u01(seed, bid, 5);
    const double cy = by * B + 10.0 + 30.0 * u01(seed, bid, 6);
    const double amp = u01(seed, bid, 7) < 0.5 ? -0.4 : 0.4;
    const double sig = 1.0 + u01(seed, bid, 8);
    z += amp * exp(-0.5 * ((x - cx) * (x - cx) + (y - cy) * (y - cy)) / (sig * sig));
    out[3 * t] = __dsub_rn(x, 0.5 * lx);
    out[3 * t + 1] = __dsub_rn(y, 0.5 * ly);
    out[3 * t + 2] = z - 1000.0;
}

}  // namespace m3d

extern "C" int more3d_synth_terrain(uint64_t seed, int64_t id_start, int64_t count, int64_t total, double lx,
                                    double ly, int slabs, double noise, double* out_xyz, void* stream) {
    M3D_REQUIRE(out_xyz != nullptr || count == 0, "out_xyz is NULL");
    M3D_REQUIRE(count >= 0 && id_start >= 0 && total > 0 && id_start + count <= total, "bad id range");
    M3D_REQUIRE(slabs > 0 && lx > 0 && ly > 0, "bad tile geometry");
    if (count == 0) return MORE3D_OK;
    cudaStream_t s = (cudaStream_t)stream;
    m3d::synth_terrain_kernel<<<(unsigned)((count + 255) / 256), 256, 0, s>>>(
        seed, id_start, count, (double)slabs / (double)total, lx / (double)slabs, lx, ly, noise, out_xyz);
    m3d::count_launches(1);
    M3D_CHECK_LAUNCH();
    return MORE3D_OK;
}
