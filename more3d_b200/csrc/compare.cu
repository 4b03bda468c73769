// "Next" row f-2: the evaluation arithmetic of Downsampling/downsample_compare.py on the device --
// PCA normals from a kNN table (open3d estimate_normals with KDTreeSearchParamKNN, :57, :112-121) and the
// per-point error terms of :132-150 (cloud-to-plane distance, angle between normals).
#include "common.cuh"
#include "eig3.cuh"

namespace m3d {

// centroid-PCA normal over the listed neighbours (the query itself is in the list, as in open3d)
__global__ void __launch_bounds__(256) normals_from_knn_kernel(const double* __restrict__ xyz, int64_t n, int k,
                                                               const int64_t* __restrict__ knn,
                                                               double* __restrict__ normals,
                                                               double* __restrict__ ratio) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double px = xyz[3 * i], py = xyz[3 * i + 1], pz = xyz[3 * i + 2];
    double sx = 0, sy = 0, sz = 0, sxx = 0, sxy = 0, sxz = 0, syy = 0, syz = 0, szz = 0;
    int m = 0;
    for (int j = 0; j < k; ++j) {
        const int64_t q = knn[i * k + j];
        if (q < 0) continue;
        const double dx = xyz[3 * q] - px, dy = xyz[3 * q + 1] - py, dz = xyz[3 * q + 2] - pz;
        ++m;
        sx += dx; sy += dy; sz += dz;
        sxx += dx * dx; sxy += dx * dy; sxz += dx * dz; syy += dy * dy; syz += dy * dz; szz += dz * dz;
    }
    double nx = 0.0, ny = 0.0, nz = 1.0, rt = 0.0;      // open3d's default normal when PCA is impossible
    if (m >= 3) {
        const double im = 1.0 / (double)m;
        const double cx = sx * im, cy = sy * im, cz = sz * im;
        const double a00 = sxx * im - cx * cx, a01 = sxy * im - cx * cy, a02 = sxz * im - cx * cz;
        const double a11 = syy * im - cy * cy, a12 = syz * im - cy * cz, a22 = szz * im - cz * cz;
        double lmin, v[3];
        sym3_min_eig(a00, a01, a02, a11, a12, a22, lmin, v);
        nx = v[0]; ny = v[1]; nz = v[2];
        const double tr = a00 + a11 + a22;
        rt = (tr > 0.0) ? fmax(lmin, 0.0) / tr : 0.0;
    }
    normals[3 * i] = nx; normals[3 * i + 1] = ny; normals[3 * i + 2] = nz;
    if (ratio) ratio[i] = rt;
}

// n[n[:, 2] < 0] = -n[n[:, 2] < 0]  (:60-62); NaN normals (fewer than 3 neighbours) become (0, 0, 1)
__global__ void __launch_bounds__(256) orient_up_kernel(double* __restrict__ normals, int64_t n, int upwards_z) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double x = normals[3 * i], y = normals[3 * i + 1], z = normals[3 * i + 2];
    if (isnan(x)) { x = 0.0; y = 0.0; z = 1.0; }
    if ((upwards_z & 1) && z < 0.0) { x = -x; y = -y; z = -z; }
    if (upwards_z & 2) {                       // normals.normalize_normals(), downsample_compare.py:122
        const double nn = sqrt(x * x + y * y + z * z);
        if (nn > 0.0) { x /= nn; y /= nn; z /= nn; }
    }
    normals[3 * i] = x; normals[3 * i + 1] = y; normals[3 * i + 2] = z;
}

// c2p[i] = ((a[i] - b[nearest[i]]) . n_a[nearest[i]])^2   (:145, a = original, b = downsampled).
// As written in the reference the ORIGINAL cloud's normals are indexed with closest2, an index into the
// DOWNSAMPLED cloud (it runs because n_down <= n_orig); kept literally.
__global__ void __launch_bounds__(256) plane_dist2_kernel(const double* __restrict__ a, int64_t na,
                                                          const double* __restrict__ b,
                                                          const int64_t* __restrict__ nearest,
                                                          const double* __restrict__ n_a,
                                                          double* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= na) return;
    const int64_t q = nearest[i];
    const double vx = a[3 * i] - b[3 * q], vy = a[3 * i + 1] - b[3 * q + 1], vz = a[3 * i + 2] - b[3 * q + 2];
    const double d = (vx * n_a[3 * q] + vy * n_a[3 * q + 1]) + vz * n_a[3 * q + 2];
    out[i] = d * d;
}

// angle (degrees, folded to [0, 90]) between n_down[i] and n_orig[closest[i]]      (:146-149)
__global__ void __launch_bounds__(256) normal_angle_kernel(const double* __restrict__ n_down, int64_t nd,
                                                           const double* __restrict__ n_orig,
                                                           const int64_t* __restrict__ closest,
                                                           double* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nd) return;
    const int64_t q = closest[i];
    double c = (n_orig[3 * q] * n_down[3 * i] + n_orig[3 * q + 1] * n_down[3 * i + 1]) + n_orig[3 * q + 2] * n_down[3 * i + 2];
    if (fabs(c) > 1.0) c = 1.0;                                   // "numerical correction", :147
    double deg = acos(c) * (180.0 / M_PI);
    if (deg > 90.0) deg = 180.0 - deg;
    out[i] = deg;
}

}  // namespace m3d

using namespace m3d;

extern "C" int more3d_normals_from_knn(const double* xyz, int64_t n, int k, const int64_t* knn, double* normals,
                                       double* lam_ratio, void* stream) {
    M3D_REQUIRE(xyz && knn && normals && n > 0 && k >= 1, "NULL / empty argument");
    cudaStream_t s = (cudaStream_t)stream;
    normals_from_knn_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(xyz, n, k, knn, normals, lam_ratio);
    count_launches(1);
    M3D_CHECK_LAUNCH();
    return MORE3D_OK;
}

extern "C" int more3d_orient_normals(double* normals, int64_t n, int upwards_z, void* stream) {
    M3D_REQUIRE(normals && n > 0, "NULL / empty argument");
    cudaStream_t s = (cudaStream_t)stream;
    orient_up_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(normals, n, upwards_z);
    count_launches(1);
    M3D_CHECK_LAUNCH();
    return MORE3D_OK;
}

extern "C" int more3d_compare_terms(const double* orig, int64_t n_orig, const double* down, int64_t n_down,
                                    const int64_t* closest_of_down, const int64_t* closest_of_orig,
                                    const double* normals_orig, const double* normals_down, double* c2p_sq,
                                    double* angle_deg, void* stream) {
    M3D_REQUIRE(orig && down && closest_of_down && closest_of_orig && normals_orig && normals_down && c2p_sq && angle_deg,
                "NULL argument");
    M3D_REQUIRE(n_orig > 0 && n_down > 0, "empty cloud");
    cudaStream_t s = (cudaStream_t)stream;
    plane_dist2_kernel<<<(unsigned)((n_orig + 255) / 256), 256, 0, s>>>(orig, n_orig, down, closest_of_orig, normals_orig,
                                                                        c2p_sq);
    normal_angle_kernel<<<(unsigned)((n_down + 255) / 256), 256, 0, s>>>(normals_down, n_down, normals_orig,
                                                                         closest_of_down, angle_deg);
    count_launches(2);
    M3D_CHECK_LAUNCH();
    return MORE3D_OK;
}
