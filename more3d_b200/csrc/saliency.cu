// K7: DM_saliency -- min/max (OPALS Histo), normalise (+min, as written) and blend (AddInfo),
// and the 256-bin histogram that feeds Otsu (Downsample.py:54).
// Replaces DM_saliency.py:562-580.
#include "common.cuh"

namespace m3d {

__global__ void mm4_init(unsigned* mm, int n) {
    if ((int)threadIdx.x < n) mm[threadIdx.x] = (threadIdx.x & 1) ? 0u : 0xffffffffu;
}
__global__ void mm4_finish(const unsigned* mm, float* out, int n) {
    if ((int)threadIdx.x < n) out[threadIdx.x] = f32_unordered(mm[threadIdx.x]);
}

__device__ __forceinline__ void warp_minmax_atomic(unsigned lo, unsigned hi, unsigned* mm) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) { atomicMin(mm, lo); atomicMax(mm + 1, hi); }
}

__global__ void __launch_bounds__(256) minmax4_kernel(const float* __restrict__ dk,
                                                      const float* __restrict__ dn, int64_t n,
                                                      unsigned* __restrict__ mm) {
    unsigned lk = 0xffffffffu, hk = 0u, ln = 0xffffffffu, hn = 0u;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const float fk = dk[i], fn = dn[i];       // NaN stays out of the extrema (see dkdn_kernel)
        if (!isnan(fk)) { const unsigned a = f32_ordered(fk); lk = min(lk, a); hk = max(hk, a); }
        if (!isnan(fn)) { const unsigned b = f32_ordered(fn); ln = min(ln, b); hn = max(hn, b); }
    }
    warp_minmax_atomic(lk, hk, mm);
    warp_minmax_atomic(ln, hn, mm + 2);
}

// (x - min)/(max - min) + min evaluated in fp64 and stored as f32 ('(float)' AddInfo columns),
// then (1-w)*ndn + w*ndk; no FMA contraction so the result is reproducible on the host.
__global__ void __launch_bounds__(256) blend_kernel(const float* __restrict__ dk, const float* __restrict__ dn,
                                                    int64_t n, const float* __restrict__ minmax, double cw,
                                                    double nw, float* __restrict__ sal,
                                                    unsigned* __restrict__ smm, const uint32_t* __restrict__ order,
                                                    long long own_lo, long long own_hi) {
    const double min_k = (double)minmax[0], max_k = (double)minmax[1];
    const double min_n = (double)minmax[2], max_n = (double)minmax[3];
    const double diff_k = __dsub_rn(max_k, min_k), diff_n = __dsub_rn(max_n, min_n);   // :570-571
    unsigned lo = 0xffffffffu, hi = 0u;
    // most dk and nearly all dn are exactly 0 (the count rules of Saliency_Kernel.process): their normalised value is
    // one constant, evaluated once with the same expression, and the fp64 division is skipped for them
    const float ndk0 = (float)__dadd_rn(__ddiv_rn(__dsub_rn(0.0, min_k), diff_k), min_k);
    const float ndn0 = (float)__dadd_rn(__ddiv_rn(__dsub_rn(0.0, min_n), diff_n), min_n);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const float vk = dk[i], vn = dn[i];
        float ndk = ndk0, ndn = ndn0;
        if (vk != 0.0f) ndk = (float)__dadd_rn(__ddiv_rn(__dsub_rn((double)vk, min_k), diff_k), min_k);  // :577
        if (vn != 0.0f) ndn = (float)__dadd_rn(__ddiv_rn(__dsub_rn((double)vn, min_n), diff_n), min_n);  // :578
        const float sv = (float)__dadd_rn(__dmul_rn(nw, (double)ndn), __dmul_rn(cw, (double)ndk));      // :579-580
        sal[i] = sv;
        // grid-order columns of a strip: only the rows of its OWN points (original index in [own_lo, own_hi)) count
        const bool mine = order == nullptr || [&] {
            const long long o = (long long)__ldg(order + i);
            return o >= own_lo && o < own_hi;
        }();
        if (mine && !isnan(sv)) {
            const unsigned u = f32_ordered(sv);
            lo = min(lo, u); hi = max(hi, u);
        }
    }
    if (smm) warp_minmax_atomic(lo, hi, smm);
}

// numpy.histogram's uniform-bin path (numpy/lib/_histograms_impl.py): index from the scaled
// offset, then the two edge fix-ups against the linspace edges.  CTA-private smem histogram,
// match_any warp aggregation in front of the shared-memory atomics, one global atomic per bin per CTA.
template <typename T>
__global__ void __launch_bounds__(256) hist256_kernel(const T* __restrict__ v, int64_t n,
                                                      const double* __restrict__ edges,
                                                      unsigned long long* __restrict__ hist) {
    __shared__ unsigned h[256];
    __shared__ double e[257];
    h[threadIdx.x] = 0;
    for (int i = threadIdx.x; i < 257; i += 256) e[i] = edges[i];
    __syncthreads();
    const double first = e[0], last = e[256];
    const double denom = __dsub_rn(last, first);
    const double scale = 256.0 / denom;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t nround = ((n + stride - 1) / stride) * stride;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nround; i += stride) {
        int bin = -1;
        if (i < n) {
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
        const unsigned peers = __match_any_sync(0xffffffffu, bin);
        if (bin >= 0 && (peers & ((1u << (threadIdx.x & 31)) - 1u)) == 0u) atomicAdd(&h[bin], __popc(peers));
    }
    __syncthreads();
    if (h[threadIdx.x]) atomicAdd(&hist[threadIdx.x], (unsigned long long)h[threadIdx.x]);
}

}  // namespace m3d

using namespace m3d;

extern "C" int more3d_minmax4(const float* dk, const float* dn, int64_t n, float* minmax, void* stream) {
    M3D_REQUIRE(dk && dn && minmax && n > 0, "NULL / empty argument");
    cudaStream_t s = (cudaStream_t)stream;
    unsigned* mm = nullptr;
    M3D_TRY(dev_alloc(&mm, 4, s));
    mm4_init<<<1, 32, 0, s>>>(mm, 4);
    minmax4_kernel<<<148 * 8, 256, 0, s>>>(dk, dn, n, mm);
    mm4_finish<<<1, 32, 0, s>>>(mm, minmax, 4);
    count_launches(3);
    M3D_CHECK_LAUNCH();
    dev_free(mm, s);
    return MORE3D_OK;
}

extern "C" int more3d_saliency_blend(const float* dk, const float* dn, int64_t n, const float* minmax,
                                     double curvature_weight, float* sal, float* sal_minmax, void* stream) {
    M3D_REQUIRE(dk && dn && minmax && sal && n > 0, "NULL / empty argument");
    cudaStream_t s = (cudaStream_t)stream;
    unsigned* mm = nullptr;
    if (sal_minmax) {
        M3D_TRY(dev_alloc(&mm, 2, s));
        mm4_init<<<1, 32, 0, s>>>(mm, 2);
    }
    const double nw = 1.0 - curvature_weight;   // DM_saliency.py:572
    ProfScope prof("blend", s);
    blend_kernel<<<148 * 8, 256, 0, s>>>(dk, dn, n, minmax, curvature_weight, nw, sal, mm, nullptr, 0, 0);
    if (sal_minmax) mm4_finish<<<1, 32, 0, s>>>(mm, sal_minmax, 2);
    count_launches(sal_minmax ? 3 : 1);
    M3D_CHECK_LAUNCH();
    dev_free(mm, s);
    return MORE3D_OK;
}

extern "C" int more3d_saliency_blend_owned(const more3d_grid* g, const float* dk, const float* dn, const float* minmax,
                                           double curvature_weight, int64_t own_lo, int64_t n_owned, float* sal,
                                           float* sal_minmax, void* stream) {
    M3D_REQUIRE(g && dk && dn && minmax && sal && sal_minmax, "NULL argument");
    cudaStream_t s = (cudaStream_t)stream;
    unsigned* mm = nullptr;
    M3D_TRY(dev_alloc(&mm, 2, s));
    mm4_init<<<1, 32, 0, s>>>(mm, 2);
    const double nw = 1.0 - curvature_weight;   // DM_saliency.py:572
    const bool all = n_owned <= 0 || (own_lo <= 0 && own_lo + n_owned >= g->n);
    ProfScope prof("blend", s);
    blend_kernel<<<148 * 8, 256, 0, s>>>(dk, dn, g->n, minmax, curvature_weight, nw, sal, mm, all ? nullptr : g->order,
                                         (long long)own_lo, (long long)(own_lo + n_owned));
    mm4_finish<<<1, 32, 0, s>>>(mm, sal_minmax, 2);
    count_launches(3);
    M3D_CHECK_LAUNCH();
    dev_free(mm, s);
    return MORE3D_OK;
}

extern "C" int more3d_histogram256(const void* v, int elem_bytes, int64_t n, const double* edges, int64_t* hist,
                                   void* stream) {
    M3D_REQUIRE(elem_bytes == 4 || elem_bytes == 8, "elem_bytes must be 4 (float32) or 8 (float64)");
    M3D_REQUIRE(v && edges && hist && n > 0, "NULL / empty argument");
    cudaStream_t s = (cudaStream_t)stream;
    M3D_CUDA(cudaMemsetAsync(hist, 0, 256 * sizeof(int64_t), s));
    if (elem_bytes == 4) hist256_kernel<float><<<148 * 4, 256, 0, s>>>((const float*)v, n, edges, (unsigned long long*)hist);
    else hist256_kernel<double><<<148 * 4, 256, 0, s>>>((const double*)v, n, edges, (unsigned long long*)hist);
    count_launches(1);
    M3D_CHECK_LAUNCH();
    return MORE3D_OK;
}
