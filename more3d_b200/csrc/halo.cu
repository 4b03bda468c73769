// K13: strip partition helpers -- stable selection of the points inside a coordinate band and
// row gathers for packing halo messages.  (No reference equivalent; OPALS tile streaming,
// DM_saliency.py:76-86, is the analogue.)
#include "common.cuh"

namespace m3d {

__global__ void __launch_bounds__(256) band_flags(const double* __restrict__ xyz, int64_t n, int axis,
                                                  double lo, double hi, uint32_t* __restrict__ flags) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double v = xyz[3 * i + axis];
    flags[i] = (v >= lo && v < hi) ? 1u : 0u;
}

__global__ void __launch_bounds__(256) band_scatter(const double* __restrict__ xyz, int64_t n, int axis,
                                                    double lo, double hi, const uint32_t* __restrict__ pos,
                                                    int64_t* __restrict__ idx) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double v = xyz[3 * i + axis];
    if (v >= lo && v < hi) idx[pos[i]] = i;
}

template <typename T>
__global__ void __launch_bounds__(256) gather_rows(const T* __restrict__ src, const int64_t* __restrict__ idx,
                                                   int64_t m, int ncols, T* __restrict__ dst) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m * ncols) return;
    const int64_t r = t / ncols;
    const int c = (int)(t - r * ncols);
    dst[t] = src[idx[r] * ncols + c];
}

}  // namespace m3d

using namespace m3d;

extern "C" int more3d_band_select(const double* xyz, int64_t n, int axis, double lo, double hi, int64_t* idx,
                                  int64_t* count_h, void* stream) {
    M3D_REQUIRE(xyz && idx && count_h && n > 0, "NULL / empty argument");
    M3D_REQUIRE(axis >= 0 && axis < 3, "axis must be 0, 1 or 2");
    cudaStream_t s = (cudaStream_t)stream;
    init_device_pool();
    uint32_t* flags = nullptr;
    M3D_TRY(dev_alloc(&flags, (size_t)n + 1, s));
    const unsigned nb = (unsigned)((n + 255) / 256);
    band_flags<<<nb, 256, 0, s>>>(xyz, n, axis, lo, hi, flags);
    M3D_TRY(exclusive_scan_u32(flags, flags, n, s));
    band_scatter<<<nb, 256, 0, s>>>(xyz, n, axis, lo, hi, flags, idx);
    count_launches(2);
    M3D_CHECK_LAUNCH();
    uint32_t total = 0;
    M3D_CUDA(cudaMemcpyAsync(&total, flags + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    M3D_CUDA(cudaStreamSynchronize(s));
    dev_free(flags, s);
    *count_h = (int64_t)total;
    return MORE3D_OK;
}

extern "C" int more3d_gather_rows(const void* src, const int64_t* idx, int64_t m, int ncols, int elem_bytes,
                                  void* dst, void* stream) {
    M3D_REQUIRE(m >= 0 && ncols > 0, "bad shape");
    if (m == 0) return MORE3D_OK;
    M3D_REQUIRE(src && idx && dst, "NULL argument");
    M3D_REQUIRE(elem_bytes == 4 || elem_bytes == 8, "elem_bytes must be 4 or 8");
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned nb = (unsigned)((m * ncols + 255) / 256);
    if (elem_bytes == 8) gather_rows<double><<<nb, 256, 0, s>>>((const double*)src, idx, m, ncols, (double*)dst);
    else gather_rows<float><<<nb, 256, 0, s>>>((const float*)src, idx, m, ncols, (float*)dst);
    count_launches(1);
    M3D_CHECK_LAUNCH();
    return MORE3D_OK;
}

namespace m3d {
__global__ void __launch_bounds__(256) flag_to_u32(const uint8_t* __restrict__ f, int64_t n, uint32_t* __restrict__ o) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) o[i] = f[i] ? 1u : 0u;
}
__global__ void __launch_bounds__(256) flag_scatter(const uint8_t* __restrict__ f, int64_t n,
                                                    const uint32_t* __restrict__ pos, int64_t* __restrict__ idx) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && f[i]) idx[pos[i]] = i;
}
}  // namespace m3d

extern "C" int more3d_select_flagged(const uint8_t* flags, int64_t n, int64_t* idx, int64_t* count_h, void* stream) {
    M3D_REQUIRE(flags && idx && count_h && n > 0, "NULL / empty argument");
    cudaStream_t s = (cudaStream_t)stream;
    init_device_pool();
    uint32_t* pos = nullptr;
    M3D_TRY(dev_alloc(&pos, (size_t)n + 1, s));
    const unsigned nb = (unsigned)((n + 255) / 256);
    flag_to_u32<<<nb, 256, 0, s>>>(flags, n, pos);
    M3D_TRY(exclusive_scan_u32(pos, pos, n, s));
    flag_scatter<<<nb, 256, 0, s>>>(flags, n, pos, idx);
    count_launches(2);
    M3D_CHECK_LAUNCH();
    uint32_t total = 0;
    M3D_CUDA(cudaMemcpyAsync(&total, pos + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    M3D_CUDA(cudaStreamSynchronize(s));
    dev_free(pos, s);
    *count_h = (int64_t)total;
    return MORE3D_OK;
}
