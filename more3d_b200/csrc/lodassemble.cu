// K9 tail: assemble the output of the saliency downsampler on the device, in the reference's row order
// (Downsampling/Downsample.py:60-83): cell by cell, the kept points of the cell (flag 1) in ball-tree
// order, then -- for cells that were not kept whole -- the original point nearest to the mean of the
// rest (flag 0).  Two scans (kept points per position, averaged cells per cell) give every row its slot.
#include "common.cuh"

namespace m3d {

__global__ void __launch_bounds__(256) la_keep_u32(const uint8_t* __restrict__ keep, int64_t n, uint32_t* __restrict__ o) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) o[i] = keep[i] ? 1u : 0u;
}
__global__ void __launch_bounds__(256) la_avg_u32(const uint8_t* __restrict__ keep_all, int64_t ncells,
                                                  uint32_t* __restrict__ o) {
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < ncells) o[c] = keep_all[c] ? 0u : 1u;
}
// scatter near[] (one entry per averaged cell, in cell order) to a per-cell table
__global__ void __launch_bounds__(256) la_near_table(const uint8_t* __restrict__ keep_all, int64_t ncells,
                                                     const uint32_t* __restrict__ apos, const int64_t* __restrict__ near,
                                                     int64_t* __restrict__ near_of_cell) {
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < ncells) near_of_cell[c] = keep_all[c] ? (int64_t)-1 : near[apos[c]];
}
// one warp per cell: rows of its kept points (row = kept points before it in the whole cloud + representatives of
// the cells before it), then its representative
__global__ void __launch_bounds__(256) la_rows(const double* __restrict__ xyz, const int64_t* __restrict__ idx_array,
                                               const int64_t* __restrict__ cell_start, int64_t ncells,
                                               const uint8_t* __restrict__ keep, const uint32_t* __restrict__ kpos,
                                               const uint32_t* __restrict__ apos, const int64_t* __restrict__ near_of_cell,
                                               double* __restrict__ out /* M x 4 */, int64_t* __restrict__ out_idx,
                                               int32_t* __restrict__ out_cell) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t c = (((int64_t)blockIdx.x * blockDim.x) + threadIdx.x) >> 5; c < ncells; c += warps) {
        const int64_t b = cell_start[c], e = cell_start[c + 1];
        const int64_t shift = (int64_t)apos[c];
        for (int64_t t = b + lane; t < e; t += 32) {
            if (!keep[t]) continue;
            const int64_t row = (int64_t)kpos[t] + shift;
            const int64_t o = idx_array[t];
            out[4 * row] = xyz[3 * o]; out[4 * row + 1] = xyz[3 * o + 1]; out[4 * row + 2] = xyz[3 * o + 2];
            out[4 * row + 3] = 1.0;                                   // salient_tmp = 1, :63 / :75
            out_idx[row] = o;
            if (out_cell) out_cell[row] = (int32_t)c;
        }
        const int64_t o = near_of_cell[c];
        if (lane == 0 && o >= 0) {                                    // pts_tmp = vstack((salient, GetPoint(nearest))), :71
            const int64_t row = (int64_t)kpos[e] + shift;
            out[4 * row] = xyz[3 * o]; out[4 * row + 1] = xyz[3 * o + 1]; out[4 * row + 2] = xyz[3 * o + 2];
            out[4 * row + 3] = 0.0;                                   // :74
            out_idx[row] = o;
            if (out_cell) out_cell[row] = (int32_t)c;
        }
    }
}

}  // namespace m3d

using namespace m3d;

extern "C" int more3d_lod_assemble(const double* xyz, const int64_t* idx_array, int64_t n, const int64_t* cell_start,
                                   int64_t ncells, const uint8_t* keep, const uint8_t* keep_all, const int64_t* near,
                                   double* out_xyzf, int64_t* out_idx, int32_t* out_cell, int64_t* nrows_h,
                                   void* stream) {
    M3D_REQUIRE(xyz && idx_array && cell_start && keep && keep_all && out_xyzf && out_idx && nrows_h, "NULL argument");
    M3D_REQUIRE(n > 0 && ncells > 0, "empty input");
    cudaStream_t s = (cudaStream_t)stream;
    init_device_pool();
    ProfScope prof("lod_assemble", s);
    uint32_t *kpos = nullptr, *apos = nullptr;
    int64_t* near_of_cell = nullptr;
    M3D_TRY(dev_alloc(&kpos, (size_t)n + 1, s));
    M3D_TRY(dev_alloc(&apos, (size_t)ncells + 1, s));
    M3D_TRY(dev_alloc(&near_of_cell, (size_t)ncells, s));
    la_keep_u32<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(keep, n, kpos);
    M3D_TRY(exclusive_scan_u32(kpos, kpos, n, s));
    const unsigned cb = (unsigned)((ncells + 255) / 256);
    la_avg_u32<<<cb, 256, 0, s>>>(keep_all, ncells, apos);
    M3D_TRY(exclusive_scan_u32(apos, apos, ncells, s));
    uint32_t tot[2] = {0, 0};
    M3D_CUDA(cudaMemcpyAsync(&tot[0], kpos + n, 4, cudaMemcpyDeviceToHost, s));
    M3D_CUDA(cudaMemcpyAsync(&tot[1], apos + ncells, 4, cudaMemcpyDeviceToHost, s));
    M3D_CUDA(cudaStreamSynchronize(s));
    M3D_REQUIRE(tot[1] == 0 || near != nullptr, "near is NULL but some cells are averaged");
    la_near_table<<<cb, 256, 0, s>>>(keep_all, ncells, apos, near, near_of_cell);
    const int64_t wb = (ncells * 32 + 255) / 256;
    la_rows<<<(unsigned)(wb < 148 * 64 ? wb : 148 * 64), 256, 0, s>>>(xyz, idx_array, cell_start, ncells, keep, kpos, apos, near_of_cell, out_xyzf, out_idx,
                               out_cell);
    count_launches(4);
    M3D_CHECK_LAUNCH();
    dev_free(kpos, s); dev_free(apos, s); dev_free(near_of_cell, s);
    *nrows_h = (int64_t)tot[0] + (int64_t)tot[1];
    return MORE3D_OK;
}
