// K10-K12: point-based level set (Chan-Vese active contour on a kNN graph with WLS derivatives).
// Replaces levelsets_func.py: compute_normals (:112-129), compute_tangent_planes (:132-151),
// Derivative.__init__/__call__/grad/div (:245-305), heaviside/delta/dp/smooth (:319-358) and
// Solver.run/_chan_vese_step/_lmv_step/zero_crossings/initialize_* (:386-513, :524-642).
//
// Per-point arrays are structure-of-arrays on the device so that thread i (= point i) reads
// consecutive addresses: neighbours nbr[j*n + i] (int32), WLS weights w[j*n + i], the folded derivative
// operator E[(r*k + j)*n + i] (r = 0..3 -> d/dx, d/dy, d2/dx2 /2, d2/dy2 /2 rows of pinv(M) times w_j) with
// its regulariser constant cst[r*n + i], tangents t1[c*n + i], t2[c*n + i].
#include "common.cuh"
#include "eig3.cuh"

#define LS_MAX_K 16
#define LS_THREADS 256

namespace m3d {

// ------------------------------------------------------------------------------------------
// K11a: normals (uncentred scatter about the query point, levelsets_func.py:115-117) + tangent frame
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(LS_THREADS) ls_frames(const double* __restrict__ xyz, int64_t n, int k2,
                                                        const int64_t* __restrict__ knn /* n x k2, self first */,
                                                        int k, int has_normals, double* __restrict__ normals,
                                                        double* __restrict__ t1, double* __restrict__ t2,
                                                        int32_t* __restrict__ nbr) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double px = xyz[3 * i], py = xyz[3 * i + 1], pz = xyz[3 * i + 2];
    double nx, ny, nz;
    if (has_normals) {
        nx = normals[3 * i]; ny = normals[3 * i + 1]; nz = normals[3 * i + 2];
    } else {
        double sxx = 0, sxy = 0, sxz = 0, syy = 0, syz = 0, szz = 0;
        for (int j = 0; j < k2; ++j) {
            const int64_t q = knn[i * k2 + j];
            const double dx = xyz[3 * q] - px, dy = xyz[3 * q + 1] - py, dz = xyz[3 * q + 2] - pz;
            sxx += dx * dx; sxy += dx * dy; sxz += dx * dz; syy += dy * dy; syz += dy * dz; szz += dz * dz;
        }
        double lmin, v[3];
        sym3_min_eig(sxx, sxy, sxz, syy, syz, szz, lmin, v);   // = Vh[-1] of svd(C) up to sign
        nx = v[0]; ny = v[1]; nz = v[2];
    }
    const double nn = sqrt(nx * nx + ny * ny + nz * nz);   // n /= norm(n), :140
    nx /= nn; ny /= nn; nz /= nn;
    const int64_t far = knn[i * k2 + (k2 - 1)];            // neighbors[i][-1], :141
    double ax = px - xyz[3 * far], ay = py - xyz[3 * far + 1], az = pz - xyz[3 * far + 2];
    const double dt = nx * ax + ny * ay + nz * az;         // :143
    ax -= dt * nx; ay -= dt * ny; az -= dt * nz;
    const double an = sqrt(ax * ax + ay * ay + az * az);   // :144
    ax /= an; ay /= an; az /= an;
    double bx = ny * az - nz * ay, by = nz * ax - nx * az, bz = nx * ay - ny * ax;   // cross(n, t1), :145
    const double bn = sqrt(bx * bx + by * by + bz * bz);
    bx /= bn; by /= bn; bz /= bn;
    normals[3 * i] = nx; normals[3 * i + 1] = ny; normals[3 * i + 2] = nz;
    t1[3 * i] = ax; t1[3 * i + 1] = ay; t1[3 * i + 2] = az;
    t2[3 * i] = bx; t2[3 * i + 1] = by; t2[3 * i + 2] = bz;
    for (int j = 0; j < k; ++j) nbr[i * k + j] = (int32_t)knn[i * k2 + 1 + j];   // neighbors[:, 1:k+1], :172
}

// ------------------------------------------------------------------------------------------
// K11b: WLS operator = pseudo-inverse of the (k+3) x 5 design matrix, one-sided Jacobi SVD per point
// ------------------------------------------------------------------------------------------
// (1 - x/h)**4 * (4*x/h + 1) (levelsets_func.py:92-97) with every operation rounded as written and a^4
// correctly rounded (double-double: two FMA error terms, one final rounding) -- what libm pow() returns.
// numpy's ARRAY power is a SIMD pow whose last bits depend on the host CPU's dispatch (measured: 3 % of the
// weights differ by 1-2 ulp from libm on an AVX-512 host), so the reference's own weights are machine
// dependent; callers that need a bit-identical trajectory pass the weights in (weights_given).
__device__ __forceinline__ double wendland(double x, double h) {
    if (!(x < h)) return 0.0;
    const double a = __dsub_rn(1.0, __ddiv_rn(x, h));
    const double p = __dmul_rn(a, a);
    const double e = __fma_rn(a, a, -p);               // a^2 = p + e exactly
    const double q = __dmul_rn(p, p);
    const double e2 = __fma_rn(p, p, -q);              // p^2 = q + e2 exactly
    const double c = __fma_rn(2.0 * p, e, e2);         // a^4 = q + (e2 + 2 p e) + O(e^2)
    const double a4 = __dadd_rn(q, c);
    const double b = __dadd_rn(__ddiv_rn(__dmul_rn(4.0, x), h), 1.0);
    return __dmul_rn(a4, b);
}

__global__ void __launch_bounds__(128) ls_derivative_init(const double* __restrict__ xyz, int64_t n, int k,
                                                          const int32_t* __restrict__ nbr /* n x k */,
                                                          const double* __restrict__ t1, const double* __restrict__ t2,
                                                          double max_d, int weights_given,
                                                          double* __restrict__ w_aos /* n x k, in when given */,
                                                          double* __restrict__ D /* n x 5 x (k+3) or NULL */,
                                                          double* __restrict__ w_soa, double* __restrict__ E,
                                                          double* __restrict__ cst, int32_t* __restrict__ nbr_soa,
                                                          double* __restrict__ t1_soa, double* __restrict__ t2_soa,
                                                          int* __restrict__ status) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int rows = k + 3;
    double A[(LS_MAX_K + 3) * 5];   // design matrix, column c at A[c * rows + j]; becomes U * S
    double V[25];
    double w[LS_MAX_K];
    const double px = xyz[3 * i], py = xyz[3 * i + 1], pz = xyz[3 * i + 2];
    const double ax = t1[3 * i], ay = t1[3 * i + 1], az = t1[3 * i + 2];
    const double bx = t2[3 * i], by = t2[3 * i + 1], bz = t2[3 * i + 2];
    for (int j = 0; j < k; ++j) {
        const int64_t q = nbr[i * k + j];
        const double dx = xyz[3 * q] - px, dy = xyz[3 * q + 1] - py, dz = xyz[3 * q + 2] - pz;
        const double x_ = dx * ax + dy * ay + dz * az;     // :251-254
        const double y_ = dx * bx + dy * by + dz * bz;
        // np.linalg.norm(..., axis=-1) = sqrt((dx^2 + dy^2) + dz^2), every product rounded
        const double dist = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
        const double wj = weights_given ? w_aos[i * k + j]
                                        : __ddiv_rn(__dadd_rn(sqrt(wendland(dist, max_d)), 1e-8), 1.0 + 1e-8);   // :264-266
        w[j] = wj;
        A[0 * rows + j] = x_ * wj;                         // :256-260, :267
        A[1 * rows + j] = y_ * wj;
        A[2 * rows + j] = (x_ * y_) * wj;
        A[3 * rows + j] = (x_ * x_) * wj;
        A[4 * rows + j] = (y_ * y_) * wj;
    }
    for (int c = 0; c < 5; ++c)
        for (int j = k; j < rows; ++j) A[c * rows + j] = (c == 2 + (j - k)) ? 1.0 : 0.0;   // :261-263
    for (int c = 0; c < 25; ++c) V[c] = (c % 6 == 0) ? 1.0 : 0.0;
    // one-sided Jacobi (Hestenes): rotate column pairs until mutually orthogonal; A -> U S, V accumulates
    for (int sweep = 0; sweep < 30; ++sweep) {
        double off = 0.0;
        for (int p = 0; p < 4; ++p)
            for (int q = p + 1; q < 5; ++q) {
                double alpha = 0, beta = 0, gamma = 0;
                for (int j = 0; j < rows; ++j) {
                    const double ap = A[p * rows + j], aq = A[q * rows + j];
                    alpha += ap * ap; beta += aq * aq; gamma += ap * aq;
                }
                if (gamma == 0.0) continue;
                const double lim = 1e-16 * sqrt(alpha * beta);
                if (fabs(gamma) <= lim) continue;
                off = fmax(off, fabs(gamma) / sqrt(alpha * beta));
                const double zeta = (beta - alpha) / (2.0 * gamma);
                const double t = ((zeta >= 0) ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                for (int j = 0; j < rows; ++j) {
                    const double ap = A[p * rows + j], aq = A[q * rows + j];
                    A[p * rows + j] = c * ap - s * aq;
                    A[q * rows + j] = s * ap + c * aq;
                }
                for (int j = 0; j < 5; ++j) {
                    const double vp = V[j * 5 + p], vq = V[j * 5 + q];
                    V[j * 5 + p] = c * vp - s * vq;
                    V[j * 5 + q] = s * vp + c * vq;
                }
            }
        if (off < 1e-15) break;
    }
    double inv_s2[5];
    bool bad = false;
    for (int c = 0; c < 5; ++c) {
        double s2 = 0;
        for (int j = 0; j < rows; ++j) s2 += A[c * rows + j] * A[c * rows + j];
        if (sqrt(s2) < 1e-12) bad = true;                  // the reference drops into pdb here, :269-271
        inv_s2[c] = 1.0 / s2;
    }
    if (bad && nbr[i * k] != (int32_t)i) atomicExch(status, 1);   // rows that list themselves are ghost placeholders
    // D = V S^-1 U^T = V S^-2 (U S)^T ; row r, column j                     :272
    const int rsel[4] = {0, 1, 3, 4};
    for (int r = 0; r < 5; ++r) {
        for (int j = 0; j < rows; ++j) {
            double d = 0;
            for (int c = 0; c < 5; ++c) d += V[r * 5 + c] * inv_s2[c] * A[c * rows + j];
            if (D) D[(i * 5 + r) * rows + j] = d;
            for (int rr = 0; rr < 4; ++rr)
                if (rsel[rr] == r) {
                    if (j < k) E[((int64_t)(rr * k + j)) * n + i] = d * w[j];
                    else if (j == k) cst[(int64_t)rr * n + i] = d;
                    else cst[(int64_t)rr * n + i] += d;
                }
        }
    }
    for (int rr = 0; rr < 4; ++rr) cst[(int64_t)rr * n + i] *= (1.0 / 1000);   // f[:, -3:] = 1/1000, :284
    for (int j = 0; j < k; ++j) {
        w_aos[i * k + j] = w[j];
        w_soa[(int64_t)j * n + i] = w[j];
        nbr_soa[(int64_t)j * n + i] = nbr[i * k + j];
    }
    for (int c = 0; c < 3; ++c) {
        t1_soa[(int64_t)c * n + i] = t1[3 * i + c];
        t2_soa[(int64_t)c * n + i] = t2[3 * i + c];
    }
}

// ------------------------------------------------------------------------------------------
// shared per-point helpers
// ------------------------------------------------------------------------------------------
struct LsGraph {
    int64_t n;
    int k;
    const int32_t* nbr;   // SoA
    const double* w;      // SoA
    const double* E;      // SoA
    const double* cst;    // SoA
    const double* t1;     // SoA
    const double* t2;     // SoA
};

__device__ __forceinline__ double np_sign(double v) { return (v > 0.0) ? 1.0 : ((v < 0.0) ? -1.0 : 0.0); }

// a_r = sum_j E[r][j] (u_j - u_i) + cst_r for r in [0, R)  (Derivative.__call__, :277-291)
template <int R>
__device__ __forceinline__ void wls_apply(const LsGraph& g, int64_t i, const double* __restrict__ u, double a[R]) {
    const double ui = u[i];
#pragma unroll
    for (int r = 0; r < R; ++r) a[r] = 0.0;
    for (int j = 0; j < g.k; ++j) {
        const double du = u[g.nbr[(int64_t)j * g.n + i]] - ui;
#pragma unroll
        for (int r = 0; r < R; ++r) a[r] += g.E[((int64_t)(r * g.k + j)) * g.n + i] * du;
    }
#pragma unroll
    for (int r = 0; r < R; ++r) a[r] += g.cst[(int64_t)r * g.n + i];
}

__device__ __forceinline__ double smooth_at(const LsGraph& g, int64_t i, const double* __restrict__ u) {   // :350-358
    // (u[nbs] * w).sum(axis=-2) / w.sum(axis=-1) with numpy's exact rounding and summation order
    double pr[LS_MAX_K], wt[LS_MAX_K];
    for (int j = 0; j < g.k; ++j) {
        const double wj = g.w[(int64_t)j * g.n + i];
        wt[j] = wj;
        pr[j] = __dmul_rn(u[g.nbr[(int64_t)j * g.n + i]], wj);
    }
    return __ddiv_rn(np_pairwise_sum(pr, g.k), np_pairwise_sum(wt, g.k));
}

__global__ void __launch_bounds__(LS_THREADS) ls_diff_kernel(LsGraph g, const double* __restrict__ u,
                                                             const uint8_t* __restrict__ mask,
                                                             double* __restrict__ out /* n x 4 */) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.n) return;
    double a[4] = {0, 0, 0, 0};
    if (!mask || mask[i]) {
        wls_apply<4>(g, i, u, a);
        a[2] *= 2.0; a[3] *= 2.0;                          // 2 * a[:, 3], 2 * a[:, 4], :291
    }
    out[4 * i] = a[0]; out[4 * i + 1] = a[1]; out[4 * i + 2] = a[2]; out[4 * i + 3] = a[3];
}

__global__ void __launch_bounds__(LS_THREADS) ls_smooth_kernel(LsGraph g, const double* __restrict__ u,
                                                               const uint8_t* __restrict__ mask,
                                                               double* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.n) return;
    out[i] = (!mask || mask[i]) ? smooth_at(g, i, u) : u[i];
}

__global__ void __launch_bounds__(LS_THREADS) ls_zero_cross_kernel(LsGraph g, const double* __restrict__ u,
                                                                   const uint8_t* __restrict__ mask,
                                                                   uint8_t* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.n) return;
    uint8_t zc = 0;
    if (!mask || mask[i]) {
        const double si = np_sign(u[i]);
        for (int j = 0; j < g.k; ++j) zc |= (np_sign(u[g.nbr[(int64_t)j * g.n + i]]) != si) ? 1 : 0;   // :533-534
    }
    out[i] = zc;
}

// numpy's float floor division (npy_divmod): fmod, (a - mod) / b, sign fix-up, floor with a > 0.5 snap
__device__ __forceinline__ double np_floor_divide(double a, double b) {
    double mod = fmod(a, b);
    if (!b) return a / b;
    double div = (a - mod) / b;
    if (mod) {
        if ((b < 0) != (mod < 0)) { mod += b; div -= 1.0; }
    }
    double floordiv;
    if (div) {
        floordiv = floor(div);
        if (div - floordiv > 0.5) floordiv += 1.0;
    } else {
        floordiv = copysign(0.0, a / b);
    }
    return floordiv;
}

__global__ void __launch_bounds__(LS_THREADS) ls_init_voxels_kernel(const double* __restrict__ xyz, int64_t n,
                                                                    double h, double max_val,
                                                                    double* __restrict__ phi) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // i = np.sum(points // h, axis=-1) % 2 == 0                                          :592
    const double s = (np_floor_divide(xyz[3 * i], h) + np_floor_divide(xyz[3 * i + 1], h)) + np_floor_divide(xyz[3 * i + 2], h);
    const bool even = fmod(s, 2.0) == 0.0;
    phi[i] = even ? -max_val : max_val;                                                  // :593-594
}

// ------------------------------------------------------------------------------------------
// K12: one Chan-Vese iteration = the kernels below, launched back to back by more3d_ls_run
// ------------------------------------------------------------------------------------------
struct LsParams {
    double stepsize, nu, mu, lambda1, lambda2, epsilon, max_val;
    int cap, channels;
    int model;   // 0 = Chan-Vese (_chan_vese_step), 1 = local mean and variance (_lmv_step, single-channel cue)
};

constexpr int RED_BLOCKS = 592;   // 4 CTAs per SM; fixed so the reduction tree (and the result) is reproducible
constexpr int STEPB_BLOCKS = 148 * 64;   // step b is a latency-bound gather: more, shorter grid-stride threads

__device__ __forceinline__ double heaviside(double x, double eps) { return 0.5 * (1.0 + 2.0 / M_PI * atan(x / eps)); }   // :319-321
__device__ __forceinline__ double delta_fn(double x, double eps) { return 1.0 / M_PI * eps / (eps * eps + x * x); }     // :332-333
__device__ __forceinline__ double dp_fn(double s) {                                                                      // :336-342
    if (s < 1.0) {
        const double x = 2.0 * s;
        if (x == 0.0) return 1.0;
        const double y = M_PI * x;
        return sin(y) / y;                                  // np.sinc
    }
    return 1.0 - 1.0 / s;
}

template <int NV>
__device__ __forceinline__ void block_sum_store(double v[NV], double* __restrict__ part) {
    __shared__ double sm[NV][LS_THREADS / 32];
#pragma unroll
    for (int a = 0; a < NV; ++a)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[a] += __shfl_xor_sync(0xffffffffu, v[a], o);
    if ((threadIdx.x & 31) == 0)
        for (int a = 0; a < NV; ++a) sm[a][threadIdx.x >> 5] = v[a];
    __syncthreads();
    if (threadIdx.x < NV) {
        double r = 0;
        for (int w = 0; w < LS_THREADS / 32; ++w) r += sm[threadIdx.x][w];
        part[(int64_t)blockIdx.x * NV + threadIdx.x] = r;
    }
}

// deactivated / newly activated points are re-clamped to +-max_val (Solver.run :603-608, initialize_new :542-545)
__global__ void __launch_bounds__(LS_THREADS) ls_reclamp(int64_t n, const uint8_t* __restrict__ active,
                                                         const uint8_t* __restrict__ last_active, double max_val,
                                                         double* __restrict__ phi,
        const int* __restrict__ stop = nullptr) {
    if (stop != nullptr && *stop != 0) return;   // the run has converged (device-side early exit, see ls_run_impl)
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (active[i] != last_active[i]) phi[i] = np_sign(phi[i]) * max_val;
}

// derivatives of phi at the active points; dp_ / ctmp fields for everyone      (_chan_vese_step :394-402, :415-416)
__global__ void __launch_bounds__(LS_THREADS) ls_step_a(LsGraph g, LsParams p, const double* __restrict__ phi,
                                                        const uint8_t* __restrict__ active, uint8_t* __restrict__ idx2,
                                                        double* __restrict__ laplace, double* __restrict__ dphi /* SoA 3n */,
                                                        double* __restrict__ aux /* AoS n x 4: dp_, ctmp xyz */,
        const int* __restrict__ stop = nullptr) {
    if (stop != nullptr && *stop != 0) return;   // the run has converged (device-side early exit, see ls_run_impl)
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.n) return;
    uint8_t on = 0;
    double dpv = 0.0, cx = 0.0, cy = 0.0, cz = 0.0;
    if (active[i]) {
        double a[4];
        wls_apply<4>(g, i, phi, a);
        laplace[i] = 2.0 * a[2] + 2.0 * a[3];               // ddx + ddy, :396
        const double gx = a[0] * g.t1[i] + a[1] * g.t2[i];
        const double gy = a[0] * g.t1[g.n + i] + a[1] * g.t2[g.n + i];
        const double gz = a[0] * g.t1[2 * g.n + i] + a[1] * g.t2[2 * g.n + i];
        dphi[i] = gx; dphi[g.n + i] = gy; dphi[2 * g.n + i] = gz;
        const double ab = sqrt(gx * gx + gy * gy + gz * gz);   // :399
        if (ab > 0.0) {                                     // idx = idx * (absdphi > 0), :400
            on = 1;
            dpv = dp_fn(ab);                                // :402
            cx = gx / ab; cy = gy / ab; cz = gz / ab;       // :416
        }
    }
    idx2[i] = on;
    // dp_[~idx] = 0 (:401) and ctmp[:] = 0 elsewhere (:415); one 32-B record per point so that step b
    // gathers a neighbour's four fields with a single sector instead of four
    double2* a2 = reinterpret_cast<double2*>(aux + 4 * i);
    a2[0] = make_double2(dpv, cx);
    a2[1] = make_double2(cy, cz);
}

// the global region means need sum zeta (1-H), sum (1-H), sum zeta H, sum H over ALL points (:410-412)
__global__ void __launch_bounds__(LS_THREADS) ls_region_partial(int64_t n, int channels, double eps,
                                                                const double* __restrict__ phi,
                                                                const double* __restrict__ zeta, int c,
                                                                double* __restrict__ part /* RED_BLOCKS x 4 */,
        const int* __restrict__ stop = nullptr) {
    if (stop != nullptr && *stop != 0) return;   // the run has converged (device-side early exit, see ls_run_impl)
    double v[4] = {0, 0, 0, 0};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double H = heaviside(phi[i], eps);
        const double z = zeta[i * channels + c];
        v[0] += z * (1.0 - H); v[1] += (1.0 - H); v[2] += z * H; v[3] += H;
    }
    block_sum_store<4>(v, part);
}
// out[0..nv) = sum over blocks.  One CTA: thread t adds blocks t, t + 256, ... in order, then a fixed shuffle /
// shared-memory tree -- the order never depends on timing, so the result is reproducible run to run.  (A single
// thread walking all 9 472 partials of step b took 0.28 ms, 12 % of an iteration.)
constexpr int RF_THREADS = 256;
__global__ void __launch_bounds__(RF_THREADS) ls_reduce_final(const double* __restrict__ part, int nblocks, int nv,
                                                              double* __restrict__ out,
        const int* __restrict__ stop = nullptr) {
    if (stop != nullptr && *stop != 0) return;   // the run has converged (device-side early exit, see ls_run_impl)
    __shared__ double sm[RF_THREADS / 32];
    for (int a = 0; a < nv; ++a) {
        double r = 0;
        for (int b = threadIdx.x; b < nblocks; b += RF_THREADS) r += part[(int64_t)b * nv + a];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
        if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = r;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0;
            for (int w = 0; w < RF_THREADS / 32; ++w) t += sm[w];
            out[a] = t;
        }
        __syncthreads();
    }
}

// ---- 'lmv' model (_lmv_step, levelsets_func.py:439-513): the data term compares each point's cue with the
// LOCAL (k+1-neighbourhood, self first) Heaviside-weighted means and deviations instead of the global region
// means.  The cue is a single channel: the reference's expressions only broadcast for a 1-D zeta.
// idx_ = idx | neighbours of idx (:463-464); idxd is zeroed by the caller
__global__ void __launch_bounds__(LS_THREADS) ls_lmv_dilate(LsGraph g, const uint8_t* __restrict__ idx2,
                                                            uint8_t* __restrict__ idxd,
        const int* __restrict__ stop = nullptr) {
    if (stop != nullptr && *stop != 0) return;   // the run has converged (device-side early exit, see ls_run_impl)
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.n) return;
    if (idx2[i]) {
        idxd[i] = 1;
        for (int j = 0; j < g.k; ++j) idxd[g.nbr[(int64_t)j * g.n + i]] = 1;
    }
}
// mu_in, mu_out, sigma_in, sigma_out of the points in idx_ (:473-495); one 32-B record per point.  Row sums follow
// numpy's pairwise order (np_pairwise_sum), 1 - H and the products are rounded as written.
__global__ void __launch_bounds__(LS_THREADS) ls_lmv_stats(LsGraph g, double eps, const double* __restrict__ phi,
                                                           const double* __restrict__ zeta,
                                                           const uint8_t* __restrict__ idxd,
                                                           const uint8_t* __restrict__ force /* or NULL */,
                                                           double* __restrict__ stats /* AoS n x 4 */,
        const int* __restrict__ stop = nullptr) {
    if (stop != nullptr && *stop != 0) return;   // the run has converged (device-side early exit, see ls_run_impl)
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.n) return;
    double2* out = reinterpret_cast<double2*>(stats + 4 * i);
    if (!idxd[i] && !(force && force[i])) {                                     // np.zeros_like(zeta) elsewhere (:466-469); never read
        out[0] = make_double2(0.0, 0.0);
        out[1] = make_double2(0.0, 0.0);
        return;
    }
    const int m = g.k + 1;
    double z[LS_MAX_K + 1], hin[LS_MAX_K + 1], hout[LS_MAX_K + 1], t[LS_MAX_K + 1];
    for (int j = 0; j < m; ++j) {
        const int64_t q = (j == 0) ? i : (int64_t)g.nbr[(int64_t)(j - 1) * g.n + i];
        const double H = heaviside(phi[q], eps);
        z[j] = zeta[q];
        hout[j] = H;
        hin[j] = __dsub_rn(1.0, H);
    }
    const double s_in = np_pairwise_sum(hin, m), s_out = np_pairwise_sum(hout, m);
    for (int j = 0; j < m; ++j) t[j] = __dmul_rn(z[j], hin[j]);
    const double mu_in = __ddiv_rn(np_pairwise_sum(t, m), s_in);                          // :473-475
    for (int j = 0; j < m; ++j) t[j] = __dmul_rn(z[j], hout[j]);
    const double mu_out = __ddiv_rn(np_pairwise_sum(t, m), s_out);                        // :476-478
    for (int j = 0; j < m; ++j) { const double d = __dsub_rn(z[j], mu_in); t[j] = __dmul_rn(__dmul_rn(d, d), hin[j]); }
    double sig_in = sqrt(__ddiv_rn(np_pairwise_sum(t, m), s_in));                         // :480-483
    for (int j = 0; j < m; ++j) { const double d = __dsub_rn(z[j], mu_out); t[j] = __dmul_rn(__dmul_rn(d, d), hout[j]); }
    double sig_out = sqrt(__ddiv_rn(np_pairwise_sum(t, m), s_out));                       // :484-487
    sig_in = (sig_in >= 0.0001 || isnan(sig_in)) ? sig_in : 0.0001;                       // np.maximum, :492-493
    sig_out = (sig_out >= 0.0001 || isnan(sig_out)) ? sig_out : 0.0001;
    out[0] = make_double2(mu_in, mu_out);
    out[1] = make_double2(sig_in, sig_out);
}
// F = e_in - e_out (:499-512) at point i
__device__ __forceinline__ double lmv_force(const LsGraph& g, int64_t i, const double* __restrict__ zeta,
                                            const double* __restrict__ stats) {
    const int m = g.k + 1;
    const double2 own = reinterpret_cast<const double2*>(stats + 4 * i)[1];
    const double sig_in = own.x, sig_out = own.y;
    const double l_out = log(__dmul_rn(2.5066282746310002, sig_out));    // np.log(np.sqrt(2*np.pi) * sigma_out)
    const double l_in = log(__dmul_rn(2.5066282746310002, sig_in));
    const double d_out = __dmul_rn(2.0, __dmul_rn(sig_out, sig_out));    // 2 * sigma_out**2
    const double d_in = __dmul_rn(2.0, __dmul_rn(sig_in, sig_in));
    double a[LS_MAX_K + 1], b[LS_MAX_K + 1];
    for (int j = 0; j < m; ++j) {
        const int64_t q = (j == 0) ? i : (int64_t)g.nbr[(int64_t)(j - 1) * g.n + i];
        const double2 mq = reinterpret_cast<const double2*>(stats + 4 * q)[0];   // (mu_in, mu_out) of q
        const double zq = zeta[q];
        const double eo = __dsub_rn(zq, mq.y), ei = __dsub_rn(zq, mq.x);
        a[j] = __dadd_rn(l_out, __ddiv_rn(__dmul_rn(eo, eo), d_out));
        b[j] = __dadd_rn(l_in, __ddiv_rn(__dmul_rn(ei, ei), d_in));
    }
    const double im = 1.0 / (double)m;
    const double e_in = __dmul_rn(im, np_pairwise_sum(a, m));
    const double e_out = __dmul_rn(im, np_pairwise_sum(b, m));
    return __dsub_rn(e_in, e_out);
}

// R + delta (C + F), phi update, increment statistics                      (:404-408, :414-418, :433-437, :614-616)
// KT > 0: the neighbour count is a compile-time constant (the common k = 5..8), the gather loop is fully unrolled and
// all of a point's index / operator / neighbour-record loads are in flight together; KT = 0: any k
template <int KT>
__global__ void __launch_bounds__(LS_THREADS) ls_step_b(LsGraph g, LsParams p, double* __restrict__ phi,
                                                        const uint8_t* __restrict__ idx2,
                                                        const double* __restrict__ laplace, const double* __restrict__ dphi,
                                                        const double* __restrict__ aux,
                                                        const double* __restrict__ zeta,
                                                        const double* __restrict__ region /* channels x 4 sums */,
                                                        const double* __restrict__ stats /* lmv: n x 4, else NULL */,
                                                        double* __restrict__ phi_new, double* __restrict__ part /* RED_BLOCKS x 2 */,
        const int* __restrict__ stop = nullptr) {
    if (stop != nullptr && *stop != 0) return;   // the run has converged (device-side early exit, see ls_run_impl)
    double v[2] = {0, 0};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < g.n; i += (int64_t)gridDim.x * blockDim.x) {
        double out = phi[i];
        if (idx2[i]) {
            const double t1x = g.t1[i], t1y = g.t1[g.n + i], t1z = g.t1[2 * g.n + i];
            const double t2x = g.t2[i], t2y = g.t2[g.n + i], t2z = g.t2[2 * g.n + i];
            double R = 0.0, C = 0.0;
            // first derivatives (rows 0, 1 of the WLS operator) of the four fields dp_, ctmp x/y/z in one
            // pass over the neighbours: Derivative.grad / div, :293-305
            const double2* own = reinterpret_cast<const double2*>(aux + 4 * i);
            const double2 o0 = own[0], o1 = own[1];
            double ax[4] = {0, 0, 0, 0}, ay[4] = {0, 0, 0, 0};
            const int kk = KT > 0 ? KT : g.k;
#pragma unroll
            for (int j = 0; j < kk; ++j) {
                const int64_t q = g.nbr[(int64_t)j * g.n + i];
                const double e0 = g.E[((int64_t)j) * g.n + i], e1 = g.E[((int64_t)(kk + j)) * g.n + i];
                const double2* nb2 = reinterpret_cast<const double2*>(aux + 4 * q);
                const double2 v0 = nb2[0], v1 = nb2[1];
                const double d0 = v0.x - o0.x, d1 = v0.y - o0.y, d2 = v1.x - o1.x, d3 = v1.y - o1.y;
                ax[0] += e0 * d0; ay[0] += e1 * d0;
                ax[1] += e0 * d1; ay[1] += e1 * d1;
                ax[2] += e0 * d2; ay[2] += e1 * d2;
                ax[3] += e0 * d3; ay[3] += e1 * d3;
            }
            const double c0 = g.cst[i], c1 = g.cst[g.n + i];
#pragma unroll
            for (int f = 0; f < 4; ++f) { ax[f] += c0; ay[f] += c1; }
            if (p.nu > 0 || p.model == 1) {
                const double gx = ax[0] * t1x + ay[0] * t2x, gy = ax[0] * t1y + ay[0] * t2y, gz = ax[0] * t1z + ay[0] * t2z;
                R = p.nu * (o0.x * laplace[i] + ((gx * dphi[i] + gy * dphi[g.n + i]) + gz * dphi[2 * g.n + i]));
            }
            if (p.mu > 0 || p.model == 1) {
                double dv = ax[1] * t1x + ay[1] * t2x;
                dv += ax[2] * t1y + ay[2] * t2y;
                dv += ax[3] * t1z + ay[3] * t2z;
                C = p.mu * dv;
            }
            double F;
            if (p.model == 1) {
                F = lmv_force(g, i, zeta, stats);
            } else {
                double f1 = 0.0, f2 = 0.0;
                for (int c = 0; c < p.channels; ++c) {
                    const double zin = region[4 * c] / region[4 * c + 1], zout = region[4 * c + 2] / region[4 * c + 3];
                    const double z = zeta[i * p.channels + c];
                    f1 += (z - zin) * (z - zin); f2 += (z - zout) * (z - zout);
                }
                F = p.lambda1 * f1 - p.lambda2 * f2;
            }
            const double inc = p.stepsize * (R + delta_fn(phi[i], p.epsilon) * (C + F));
            out = phi[i] + inc;
            v[0] += inc * inc; v[1] += 1.0;
        }
        phi_new[i] = out;
    }
    block_sum_store<2>(v, part);
}

// active[active] = zero_crossings(phi, active) (:617-619); cap to +-max_val (:620-624)
__global__ void __launch_bounds__(LS_THREADS) ls_step_c(LsGraph g, LsParams p, const double* __restrict__ phi_new,
                                                        const uint8_t* __restrict__ active, uint8_t* __restrict__ active_zc,
                                                        double* __restrict__ phi_cap, uint8_t* __restrict__ outside,
        const int* __restrict__ stop = nullptr) {
    if (stop != nullptr && *stop != 0) return;   // the run has converged (device-side early exit, see ls_run_impl)
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.n) return;
    const double u = phi_new[i];
    uint8_t zc = 0;
    if (active[i]) {
        const double si = np_sign(u);
        for (int j = 0; j < g.k; ++j) zc |= (np_sign(phi_new[g.nbr[(int64_t)j * g.n + i]]) != si) ? 1 : 0;
    }
    active_zc[i] = zc;
    uint8_t o = 0;
    double v = u;
    if (p.cap && (p.mu > 0 || p.nu > 0) && (u > p.max_val || u < -p.max_val)) {
        o = 1;
        v = np_sign(u) * p.max_val;
    }
    outside[i] = o;
    phi_cap[i] = v;
}
// phi[outside] = smooth(phi, outside)  (Jacobi: reads the clamped field, :625-627)
__global__ void __launch_bounds__(LS_THREADS) ls_step_d(LsGraph g, const double* __restrict__ phi_cap,
                                                        const uint8_t* __restrict__ outside, double* __restrict__ phi_s,
        const int* __restrict__ stop = nullptr) {
    if (stop != nullptr && *stop != 0) return;   // the run has converged (device-side early exit, see ls_run_impl)
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.n) return;
    phi_s[i] = outside[i] ? smooth_at(g, i, phi_cap) : phi_cap[i];
}
// lonely points: every neighbour has the other sign -> smoothed (:628-631); result is the new phi
__global__ void __launch_bounds__(LS_THREADS) ls_step_e(LsGraph g, const double* __restrict__ phi_s,
                                                        double* __restrict__ phi_out,
        const int* __restrict__ stop = nullptr) {
    if (stop != nullptr && *stop != 0) return;   // the run has converged (device-side early exit, see ls_run_impl)
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.n) return;
    const double u = phi_s[i];
    const double si = np_sign(u);
    bool lonely = true;
    for (int j = 0; j < g.k; ++j) lonely = lonely && (np_sign(phi_s[g.nbr[(int64_t)j * g.n + i]]) != si);
    phi_out[i] = lonely ? smooth_at(g, i, phi_s) : u;
}
// active[neighbors[active]] = True (:632)
__global__ void __launch_bounds__(LS_THREADS) ls_step_f(LsGraph g, const uint8_t* __restrict__ active_zc,
                                                        uint8_t* __restrict__ active,
        const int* __restrict__ stop = nullptr) {
    if (stop != nullptr && *stop != 0) return;   // the run has converged (device-side early exit, see ls_run_impl)
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.n) return;
    if (active_zc[i]) {
        active[i] = 1;
        for (int j = 0; j < g.k; ++j) active[g.nbr[(int64_t)j * g.n + i]] = 1;
    }
}

// ---- Solver.save(extract=True), the region-growing part (levelsets_func.py:651-661) ----------------------
// sums of the cue over phi > 0 and phi <= 0 (all channels) and the element counts
__global__ void __launch_bounds__(LS_THREADS) ls_extract_partial(int64_t n, int channels, const double* __restrict__ phi,
                                                                 const double* __restrict__ zeta,
                                                                 double* __restrict__ part /* RED_BLOCKS x 4 */) {
    double v[4] = {0, 0, 0, 0};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double z = 0.0;
        for (int c = 0; c < channels; ++c) z += zeta[i * channels + c];
        if (phi[i] > 0.0) { v[0] += z; v[1] += (double)channels; } else { v[2] += z; v[3] += (double)channels; }
    }
    block_sum_store<4>(v, part);
}
// salient = region | (sum(region[neighbors]) >= threshold)                                  :658-659
__global__ void __launch_bounds__(LS_THREADS) ls_extract_grow(LsGraph g, const double* __restrict__ phi,
                                                              const double* __restrict__ sums, int threshold,
                                                              uint8_t* __restrict__ s1) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.n) return;
    // region1 (phi > 0) is the salient one iff its mean cue is larger; an empty region gives NaN -> region2 (:653-656)
    const bool pos = (sums[0] / sums[1]) > (sums[2] / sums[3]);
    int cnt = 0;
    for (int j = 0; j < g.k; ++j) cnt += ((phi[g.nbr[(int64_t)j * g.n + i]] > 0.0) == pos) ? 1 : 0;
    s1[i] = (((phi[i] > 0.0) == pos) || cnt >= threshold) ? 1 : 0;
}
// salient &= any(salient[neighbors])                                                        :660
__global__ void __launch_bounds__(LS_THREADS) ls_extract_prune(LsGraph g, const uint8_t* __restrict__ s1,
                                                               uint8_t* __restrict__ s2) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.n) return;
    uint8_t any = 0;
    for (int j = 0; j < g.k; ++j) any |= s1[g.nbr[(int64_t)j * g.n + i]];
    s2[i] = s1[i] & any;
}

}  // namespace m3d

using namespace m3d;

// byte-buffer clears / copies of the iteration as kernels, so that they too stop once the run has converged
__global__ void __launch_bounds__(LS_THREADS) ls_fill_u8(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, int64_t n,
                                                         const int* __restrict__ stop) {
    if (stop != nullptr && *stop != 0) return;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = src ? src[i] : (uint8_t)0;
}
// end of iteration `it`: mean(increment^2)^0.5 < tolerance (levelsets_func.py:635-638) raises the stop flag; every later
// kernel of the queue then returns at once, and `done` says how many iterations really ran
__global__ void ls_check_tolerance(const double* __restrict__ inc2, double tolerance, int it, int* stop, int* done) {
    if (*stop != 0) return;
    *done = it + 1;
    if (sqrt(inc2[0] / inc2[1]) < tolerance) *stop = 1;
}

static void launch_step_b(const LsGraph& g, const LsParams& p, double* phi, const uint8_t* idx2, const double* laplace,
                          const double* dphi, const double* aux, const double* zeta, const double* region,
                          const double* stats, double* phi_new, double* part, cudaStream_t s, const int* stop = nullptr) {
#define M3D_STEP_B(KT_) ls_step_b<KT_><<<STEPB_BLOCKS, LS_THREADS, 0, s>>>(g, p, phi, idx2, laplace, dphi, aux, zeta, region, stats, phi_new, part, stop)
    switch (g.k) {
        case 5: M3D_STEP_B(5); break;
        case 6: M3D_STEP_B(6); break;
        case 7: M3D_STEP_B(7); break;
        case 8: M3D_STEP_B(8); break;
        default: M3D_STEP_B(0); break;
    }
#undef M3D_STEP_B
}

struct more3d_ls_graph {   // typedef'd in include/more3d_b200.h
    int64_t n;
    int k;
    int32_t* nbr;
    double *w, *E, *cst, *t1, *t2;
    cudaStream_t stream;
};

static LsGraph view(const more3d_ls_graph* G) {
    LsGraph g;
    g.n = G->n; g.k = G->k; g.nbr = G->nbr; g.w = G->w; g.E = G->E; g.cst = G->cst; g.t1 = G->t1; g.t2 = G->t2;
    return g;
}

extern "C" int more3d_ls_frames(const double* xyz, int64_t n, int k, const int64_t* knn2k, int has_normals,
                                double* normals, double* t1, double* t2, int32_t* neighbors, void* stream) {
    M3D_REQUIRE(xyz && knn2k && normals && t1 && t2 && neighbors && n > 0, "NULL / empty argument");
    M3D_REQUIRE(k >= 1 && k <= LS_MAX_K, "k out of range (1..16)");
    cudaStream_t s = (cudaStream_t)stream;
    ProfScope prof("ls_frames", s);
    ls_frames<<<(unsigned)((n + LS_THREADS - 1) / LS_THREADS), LS_THREADS, 0, s>>>(xyz, n, 2 * k, knn2k, k, has_normals,
                                                                                   normals, t1, t2, neighbors);
    count_launches(1);
    M3D_CHECK_LAUNCH();
    return MORE3D_OK;
}

extern "C" int more3d_ls_graph_create(const double* xyz, int64_t n, int k, const int32_t* neighbors, const double* t1,
                                      const double* t2, double max_d, int weights_given, double* weights, double* D,
                                      int* singular_h, void* stream, more3d_ls_graph** out) {
    M3D_REQUIRE(out != nullptr, "out is NULL");
    *out = nullptr;
    M3D_REQUIRE(xyz && neighbors && t1 && t2 && weights && n > 0, "NULL / empty argument");
    M3D_REQUIRE(k >= 1 && k <= LS_MAX_K, "k out of range (1..16)");
    cudaStream_t s = (cudaStream_t)stream;
    init_device_pool();
    ProfScope prof("ls_derivative_init", s);
    more3d_ls_graph* G = new more3d_ls_graph();
    G->n = n; G->k = k; G->stream = s;
    G->nbr = nullptr; G->w = G->E = G->cst = G->t1 = G->t2 = nullptr;
    int* status = nullptr;
    int rc = MORE3D_OK;
    if ((rc = dev_alloc(&G->nbr, (size_t)n * k, s)) || (rc = dev_alloc(&G->w, (size_t)n * k, s)) ||
        (rc = dev_alloc(&G->E, (size_t)n * k * 4, s)) || (rc = dev_alloc(&G->cst, (size_t)n * 4, s)) ||
        (rc = dev_alloc(&G->t1, (size_t)n * 3, s)) || (rc = dev_alloc(&G->t2, (size_t)n * 3, s)) ||
        (rc = dev_alloc(&status, 1, s))) {
        more3d_ls_graph_destroy(G);
        return rc;
    }
    cudaMemsetAsync(status, 0, sizeof(int), s);
    ls_derivative_init<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(xyz, n, k, neighbors, t1, t2, max_d, weights_given, weights, D, G->w,
                                                                   G->E, G->cst, G->nbr, G->t1, G->t2, status);
    count_launches(1);
    int st = 0;
    cudaError_t e = cudaMemcpyAsync(&st, status, sizeof(int), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    dev_free(status, s);
    if (e != cudaSuccess) {
        set_error("ls_graph_create: %s", cudaGetErrorString(e));
        more3d_ls_graph_destroy(G);
        return MORE3D_E_CUDA;
    }
    if (singular_h) *singular_h = st;
    *out = G;
    return MORE3D_OK;
}

extern "C" int more3d_ls_graph_destroy(more3d_ls_graph* G) {
    if (!G) return MORE3D_OK;
    dev_free(G->nbr, G->stream); dev_free(G->w, G->stream); dev_free(G->E, G->stream);
    dev_free(G->cst, G->stream); dev_free(G->t1, G->stream); dev_free(G->t2, G->stream);
    delete G;
    return MORE3D_OK;
}

extern "C" int more3d_ls_diff(const more3d_ls_graph* G, const double* u, const uint8_t* mask, double* out4, void* stream) {
    M3D_REQUIRE(G && u && out4, "NULL argument");
    cudaStream_t s = (cudaStream_t)stream;
    ls_diff_kernel<<<(unsigned)((G->n + LS_THREADS - 1) / LS_THREADS), LS_THREADS, 0, s>>>(view(G), u, mask, out4);
    count_launches(1);
    M3D_CHECK_LAUNCH();
    return MORE3D_OK;
}

extern "C" int more3d_ls_smooth(const more3d_ls_graph* G, const double* u, const uint8_t* mask, double* out, void* stream) {
    M3D_REQUIRE(G && u && out && u != out, "NULL / aliased argument");
    cudaStream_t s = (cudaStream_t)stream;
    ls_smooth_kernel<<<(unsigned)((G->n + LS_THREADS - 1) / LS_THREADS), LS_THREADS, 0, s>>>(view(G), u, mask, out);
    count_launches(1);
    M3D_CHECK_LAUNCH();
    return MORE3D_OK;
}

extern "C" int more3d_ls_zero_crossings(const more3d_ls_graph* G, const double* u, const uint8_t* mask, uint8_t* out,
                                        void* stream) {
    M3D_REQUIRE(G && u && out, "NULL argument");
    cudaStream_t s = (cudaStream_t)stream;
    ls_zero_cross_kernel<<<(unsigned)((G->n + LS_THREADS - 1) / LS_THREADS), LS_THREADS, 0, s>>>(view(G), u, mask, out);
    count_launches(1);
    M3D_CHECK_LAUNCH();
    return MORE3D_OK;
}

extern "C" int more3d_ls_extract(const more3d_ls_graph* G, const double* phi, const double* zeta, int channels,
                                 int extraction_threshold, uint8_t* salient, void* stream) {
    M3D_REQUIRE(G && phi && zeta && salient && channels >= 1, "NULL / bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t n = G->n;
    double *part = nullptr, *sums = nullptr;
    uint8_t* s1 = nullptr;
    M3D_TRY(dev_alloc(&part, (size_t)RED_BLOCKS * 4, s));
    M3D_TRY(dev_alloc(&sums, 4, s));
    M3D_TRY(dev_alloc(&s1, (size_t)n, s));
    const unsigned nb = (unsigned)((n + LS_THREADS - 1) / LS_THREADS);
    ls_extract_partial<<<RED_BLOCKS, LS_THREADS, 0, s>>>(n, channels, phi, zeta, part);
    ls_reduce_final<<<1, RF_THREADS, 0, s>>>(part, RED_BLOCKS, 4, sums);
    ls_extract_grow<<<nb, LS_THREADS, 0, s>>>(view(G), phi, sums, extraction_threshold, s1);
    ls_extract_prune<<<nb, LS_THREADS, 0, s>>>(view(G), s1, salient);
    count_launches(4);
    M3D_CHECK_LAUNCH();
    dev_free(part, s); dev_free(sums, s); dev_free(s1, s);
    return MORE3D_OK;
}

extern "C" int more3d_ls_init_voxels(const double* xyz, int64_t n, double h, double max_val, double* phi, void* stream) {
    M3D_REQUIRE(xyz && phi && n > 0, "NULL / empty argument");
    cudaStream_t s = (cudaStream_t)stream;
    ls_init_voxels_kernel<<<(unsigned)((n + LS_THREADS - 1) / LS_THREADS), LS_THREADS, 0, s>>>(xyz, n, h, max_val, phi);
    count_launches(1);
    M3D_CHECK_LAUNCH();
    return MORE3D_OK;
}

static int ls_run_impl(const more3d_ls_graph* G, int model, double* phi, const double* zeta, int channels, uint8_t* active,
                       uint8_t* last_active, int alias_last_active, int iterations, double stepsize, double nu,
                       double mu, double lambda1, double lambda2, double epsilon, double max_val, int cap,
                       double tolerance, double* rmsc_h, int* iterations_done_h, int* converged_h, void* stream) {
    M3D_REQUIRE(G && phi && zeta && active && last_active, "NULL argument");
    M3D_REQUIRE(channels >= 1 && iterations >= 0, "bad channels / iterations");
    M3D_REQUIRE(model == 0 || model == 1, "model must be 0 (chan_vese) or 1 (lmv)");
    M3D_REQUIRE(model == 0 || channels == 1, "the lmv model takes a single-channel cue");
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t n = G->n;
    LsGraph g = view(G);
    LsParams p;
    p.stepsize = stepsize; p.nu = nu; p.mu = mu; p.lambda1 = lambda1; p.lambda2 = lambda2; p.epsilon = epsilon;
    p.max_val = max_val; p.cap = cap; p.channels = channels; p.model = model;
    ProfScope prof("ls_run", s);
    Scratch sc(s);
    double* stats = nullptr;
    uint8_t* idxd = nullptr;
    if (model == 1) { M3D_TRY(sc.alloc(&stats, (size_t)4 * n)); M3D_TRY(sc.alloc(&idxd, (size_t)n)); }
    double *laplace = nullptr, *dphi = nullptr, *aux = nullptr, *bufA = nullptr, *bufB = nullptr;
    double *part = nullptr, *region = nullptr, *incs = nullptr;
    uint8_t *idx2 = nullptr, *azc = nullptr, *outside = nullptr;
    M3D_TRY(sc.alloc(&laplace, (size_t)n)); M3D_TRY(sc.alloc(&dphi, (size_t)3 * n));
    M3D_TRY(sc.alloc(&aux, (size_t)4 * n));
    M3D_TRY(sc.alloc(&bufA, (size_t)n));    M3D_TRY(sc.alloc(&bufB, (size_t)n));
    M3D_TRY(sc.alloc(&part, (size_t)STEPB_BLOCKS * 4));
    M3D_TRY(sc.alloc(&region, (size_t)4 * channels));
    M3D_TRY(sc.alloc(&incs, (size_t)2 * (iterations + 1)));
    M3D_TRY(sc.alloc(&idx2, (size_t)n)); M3D_TRY(sc.alloc(&azc, (size_t)n)); M3D_TRY(sc.alloc(&outside, (size_t)n));
    const unsigned nb = (unsigned)((n + LS_THREADS - 1) / LS_THREADS);
    int done = 0, conv = 0;
    // tolerance > 0 (the reference's default): convergence is decided ON THE DEVICE -- a stop flag raised by the
    // iteration that meets the tolerance makes every kernel queued after it return at once -- so the whole run is
    // enqueued without a host round trip per iteration and read back with ONE synchronisation at the end.
    int* flags = nullptr;                         // {stop, done}
    M3D_TRY(sc.alloc(&flags, 2));
    M3D_CUDA(cudaMemsetAsync(flags, 0, 2 * sizeof(int), s));
    const int* stop = tolerance > 0 ? flags : nullptr;
    const unsigned nbf = 148 * 8;
    for (int it = 0; it < iterations; ++it) {
        if (!alias_last_active) ls_reclamp<<<nb, LS_THREADS, 0, s>>>(n, active, last_active, max_val, phi, stop);
        if (nu > 0 || mu > 0 || model == 1) {             // _lmv_step differentiates unconditionally, :440-446
            ls_step_a<<<nb, LS_THREADS, 0, s>>>(g, p, phi, active, idx2, laplace, dphi, aux, stop);
        } else {
            ls_fill_u8<<<nbf, LS_THREADS, 0, s>>>(idx2, active, n, stop);      // idx stays the active mask
        }
        if (model == 1) {
            ls_fill_u8<<<nbf, LS_THREADS, 0, s>>>(idxd, nullptr, n, stop);
            ls_lmv_dilate<<<nb, LS_THREADS, 0, s>>>(g, idx2, idxd, stop);
            ls_lmv_stats<<<nb, LS_THREADS, 0, s>>>(g, epsilon, phi, zeta, idxd, nullptr, stats, stop);
            count_launches(3);
        } else {
            for (int c = 0; c < channels; ++c) {
                ls_region_partial<<<RED_BLOCKS, LS_THREADS, 0, s>>>(n, channels, epsilon, phi, zeta, c, part, stop);
                ls_reduce_final<<<1, RF_THREADS, 0, s>>>(part, RED_BLOCKS, 4, region + 4 * c, stop);
            }
        }
        launch_step_b(g, p, phi, idx2, laplace, dphi, aux, zeta, region, stats, bufA, part, s, stop);
        ls_reduce_final<<<1, RF_THREADS, 0, s>>>(part, STEPB_BLOCKS, 2, incs + 2 * it, stop);
        if (!alias_last_active) ls_fill_u8<<<nbf, LS_THREADS, 0, s>>>(last_active, active, n, stop);   // :617
        ls_step_c<<<nb, LS_THREADS, 0, s>>>(g, p, bufA, active, azc, bufB, outside, stop);
        ls_step_d<<<nb, LS_THREADS, 0, s>>>(g, bufB, outside, bufA, stop);
        ls_step_e<<<nb, LS_THREADS, 0, s>>>(g, bufA, phi, stop);
        ls_fill_u8<<<nbf, LS_THREADS, 0, s>>>(active, nullptr, n, stop);
        ls_step_f<<<nb, LS_THREADS, 0, s>>>(g, azc, active, stop);
        if (tolerance > 0) ls_check_tolerance<<<1, 1, 0, s>>>(incs + 2 * it, tolerance, it, flags, flags + 1);   // :635-638
        count_launches(9 + (model == 1 ? 0 : 2 * channels) + (alias_last_active ? 0 : 2) + (tolerance > 0 ? 1 : 0));
        M3D_CHECK_LAUNCH();
        done = it + 1;
    }
    if (tolerance > 0 && iterations > 0) {
        int h[2] = {0, 0};
        M3D_CUDA(cudaMemcpyAsync(h, flags, sizeof(h), cudaMemcpyDeviceToHost, s));
        M3D_CUDA(cudaStreamSynchronize(s));
        conv = h[0];
        done = h[1];
    }
    if (rmsc_h && done > 0) {
        double* tmp = new double[(size_t)2 * done];
        cudaError_t e = cudaMemcpyAsync(tmp, incs, sizeof(double) * 2 * done, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        for (int it = 0; it < done; ++it) rmsc_h[it] = sqrt(tmp[2 * it] / tmp[2 * it + 1]);   // mean(increment**2)**0.5, :616
        delete[] tmp;
        if (e != cudaSuccess) { set_error("ls_run: %s", cudaGetErrorString(e)); return MORE3D_E_CUDA; }
    }
    if (iterations_done_h) *iterations_done_h = done;
    if (converged_h) *converged_h = conv;
    return MORE3D_OK;
}

extern "C" int more3d_ls_run(const more3d_ls_graph* G, double* phi, const double* zeta, int channels, uint8_t* active,
                             uint8_t* last_active, int alias_last_active, int iterations, double stepsize, double nu,
                             double mu, double lambda1, double lambda2, double epsilon, double max_val, int cap,
                             double tolerance, double* rmsc_h, int* iterations_done_h, int* converged_h, void* stream) {
    return ls_run_impl(G, 0, phi, zeta, channels, active, last_active, alias_last_active, iterations, stepsize, nu, mu,
                       lambda1, lambda2, epsilon, max_val, cap, tolerance, rmsc_h, iterations_done_h, converged_h, stream);
}

extern "C" int more3d_ls_run_lmv(const more3d_ls_graph* G, double* phi, const double* zeta, uint8_t* active,
                                 uint8_t* last_active, int alias_last_active, int iterations, double stepsize, double nu,
                                 double mu, double epsilon, double max_val, int cap, double tolerance, double* rmsc_h,
                                 int* iterations_done_h, int* converged_h, void* stream) {
    return ls_run_impl(G, 1, phi, zeta, 1, active, last_active, alias_last_active, iterations, stepsize, nu, mu, 0.0, 0.0,
                       epsilon, max_val, cap, tolerance, rmsc_h, iterations_done_h, converged_h, stream);
}


// ------------------------------------------------------------------------------------------
// One phase of the iteration at a time, for hosts that interleave exchanges between the phases (multi-GPU:
// ghost values of phi / aux travel between ranks, the region sums and the RMS change are all-reduced).
// Points [0, n_own) are owned by this rank, the rest are ghosts whose values the owner supplies.
// ------------------------------------------------------------------------------------------
extern "C" int more3d_ls_phase(const more3d_ls_graph* G, int phase, const more3d_ls_state* st, void* stream) {
    M3D_REQUIRE(G && st, "NULL argument");
    M3D_REQUIRE(st->n_own >= 0 && st->n_own <= G->n, "n_own out of range");
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t n = G->n;
    LsGraph g = view(G);
    LsParams p;
    p.stepsize = st->stepsize; p.nu = st->nu; p.mu = st->mu; p.lambda1 = st->lambda1; p.lambda2 = st->lambda2;
    p.epsilon = st->epsilon; p.max_val = st->max_val; p.cap = st->cap; p.channels = st->channels; p.model = st->model;
    M3D_REQUIRE(st->model == 0 || st->model == 1, "model must be 0 (chan_vese) or 1 (lmv)");
    M3D_REQUIRE(st->model == 0 || (st->channels == 1 && st->stats && st->idxd), "lmv: one channel, stats and idxd needed");
    const unsigned nb = (unsigned)((n + LS_THREADS - 1) / LS_THREADS);
    switch (phase) {
    case 0:   // re-clamp (de)activated points
        ls_reclamp<<<nb, LS_THREADS, 0, s>>>(n, st->active, st->last_active, st->max_val, st->phi);
        count_launches(1);
        break;
    case 1:   // derivatives of phi at the active points -> idx2, laplace, dphi, aux
        if (st->nu > 0 || st->mu > 0 || st->model == 1) {
            ls_step_a<<<nb, LS_THREADS, 0, s>>>(g, p, st->phi, st->active, st->idx2, st->laplace, st->dphi, st->aux);
            count_launches(1);
        } else {
            M3D_CUDA(cudaMemcpyAsync(st->idx2, st->active, (size_t)n, cudaMemcpyDeviceToDevice, s));
        }
        break;
    case 2:   // local region sums over the OWNED points -> region[4 * channels] (caller all-reduces)
        if (st->model == 1) {   // lmv: local statistics of idx_ and of the exported points (caller pulls the ghosts')
            M3D_CUDA(cudaMemsetAsync(st->idxd, 0, (size_t)n, s));
            ls_lmv_dilate<<<nb, LS_THREADS, 0, s>>>(g, st->idx2, st->idxd);
            ls_lmv_stats<<<nb, LS_THREADS, 0, s>>>(g, st->epsilon, st->phi, st->zeta, st->idxd, st->force, st->stats);
            count_launches(2);
            break;
        }
        for (int c = 0; c < st->channels; ++c) {
            ls_region_partial<<<RED_BLOCKS, LS_THREADS, 0, s>>>(st->n_own, st->channels, st->epsilon, st->phi, st->zeta, c,
                                                                st->part);
            ls_reduce_final<<<1, RF_THREADS, 0, s>>>(st->part, RED_BLOCKS, 4, st->region + 4 * c);
        }
        count_launches(2 * st->channels);
        break;
    case 3:   // update -> bufA (= phi + increment), incs[0..1] = local sum inc^2, count
        launch_step_b(g, p, st->phi, st->idx2, st->laplace, st->dphi, st->aux, st->zeta, st->region,
                      st->model == 1 ? st->stats : nullptr, st->bufA, st->part, s);
        ls_reduce_final<<<1, RF_THREADS, 0, s>>>(st->part, STEPB_BLOCKS, 2, st->incs);
        count_launches(2);
        break;
    case 4:   // zero crossings of the active points (-> azc), cap (-> bufB, outside)
        ls_step_c<<<nb, LS_THREADS, 0, s>>>(g, p, st->bufA, st->active, st->azc, st->bufB, st->outside);
        count_launches(1);
        break;
    case 5:   // smooth the capped points: bufB -> bufA
        ls_step_d<<<nb, LS_THREADS, 0, s>>>(g, st->bufB, st->outside, st->bufA);
        count_launches(1);
        break;
    case 6:   // smooth the lonely points: bufA -> phi
        ls_step_e<<<nb, LS_THREADS, 0, s>>>(g, st->bufA, st->phi);
        count_launches(1);
        break;
    case 7:   // active = azc dilated by one neighbourhood (ghost flags are sent to their owners by the caller)
        M3D_CUDA(cudaMemsetAsync(st->active, 0, (size_t)n, s));
        ls_step_f<<<nb, LS_THREADS, 0, s>>>(g, st->azc, st->active);
        count_launches(1);
        break;
    default:
        set_error("unknown level-set phase %d", phase);
        return MORE3D_E_INVALID;
    }
    M3D_CHECK_LAUNCH();
    return MORE3D_OK;
}

extern "C" int64_t more3d_ls_scratch_doubles(void) { return (int64_t)STEPB_BLOCKS * 4; }
