// Stand-alone level-set helpers that work on the caller's array-of-structures arrays (no graph
// handle): the module-level functions of levelsets_func.py -- smooth (:350-358), heaviside (:319-321),
// delta (:332-333), dp (:336-342) and wendland (:92-97).
#include "common.cuh"

namespace m3d {

// out[i, c] = sum_j u[nbs[i, j], c] * w[i, j] / sum_j w[i, j]   for mask[i] != 0, else u[i, c]
__global__ void __launch_bounds__(256) smooth_raw_kernel(const double* __restrict__ u, int cols,
                                                         const int64_t* __restrict__ nbs, const double* __restrict__ w,
                                                         int64_t n, int k, const uint8_t* __restrict__ mask,
                                                         double* __restrict__ out) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * cols) return;
    const int64_t i = t / cols;
    const int c = (int)(t - i * cols);
    if (mask && !mask[i]) { out[t] = u[t]; return; }
    // numpy's rounding and order: the weights are summed pairwise (contiguous axis); the products pairwise for
    // a 1-D u (shape (m, k, 1): the k axis is contiguous) and left to right for an (N, C >= 2) u
    double pr[64], wt[64];
    for (int j = 0; j < k; ++j) {
        const double wj = w[i * k + j];
        wt[j] = wj;
        pr[j] = __dmul_rn(u[nbs[i * k + j] * cols + c], wj);
    }
    double s;
    if (cols == 1) s = np_pairwise_sum(pr, k);
    else {
        s = 0.0;
        for (int j = 0; j < k; ++j) s = __dadd_rn(s, pr[j]);
    }
    out[t] = __ddiv_rn(s, np_pairwise_sum(wt, k));
}

enum { OP_HEAVISIDE = 0, OP_DELTA = 1, OP_DP = 2, OP_WENDLAND = 3 };

__global__ void __launch_bounds__(256) elementwise_kernel(int op, const double* __restrict__ x, double prm, int64_t n,
                                                          double* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double v = x[i];
    double r;
    if (op == OP_HEAVISIDE) {
        r = 0.5 * (1.0 + 2.0 / M_PI * atan(v / prm));
    } else if (op == OP_DELTA) {
        r = 1.0 / M_PI * prm / (prm * prm + v * v);
    } else if (op == OP_DP) {
        if (v < 1.0) {
            const double a = 2.0 * v;
            r = (a == 0.0) ? 1.0 : sin(M_PI * a) / (M_PI * a);
        } else {
            r = 1.0 - 1.0 / v;
        }
    } else {
        if (v < prm) {
            const double a = 1.0 - v / prm, a2 = a * a;
            r = a2 * a2 * (4.0 * v / prm + 1.0);
        } else {
            r = 0.0;
        }
    }
    out[i] = r;
}

}  // namespace m3d

using namespace m3d;

extern "C" int more3d_ls_smooth_raw(const double* u, int cols, const int64_t* nbs, const double* w, int64_t n, int k,
                                    const uint8_t* mask, double* out, void* stream) {
    M3D_REQUIRE(u && nbs && w && out && u != out, "NULL / aliased argument");
    M3D_REQUIRE(n > 0 && k >= 1 && k <= 64 && cols >= 1, "bad shape (1 <= k <= 64)");
    cudaStream_t s = (cudaStream_t)stream;
    smooth_raw_kernel<<<(unsigned)((n * cols + 255) / 256), 256, 0, s>>>(u, cols, nbs, w, n, k, mask, out);
    count_launches(1);
    M3D_CHECK_LAUNCH();
    return MORE3D_OK;
}

extern "C" int more3d_ls_elementwise(int op, const double* x, double param, int64_t n, double* out, void* stream) {
    M3D_REQUIRE(x && out && n > 0, "NULL / empty argument");
    M3D_REQUIRE(op >= 0 && op <= 3, "unknown op");
    cudaStream_t s = (cudaStream_t)stream;
    elementwise_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(op, x, param, n, out);
    count_launches(1);
    M3D_CHECK_LAUNCH();
    return MORE3D_OK;
}
