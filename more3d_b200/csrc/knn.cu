// K10: k nearest neighbours on the uniform grid (expanding square rings of columns).
// Replaces BallTreePointSet.query (Downsampling/BallTreePointSet.py:292-315) and the sorted 2k-NN of
// levelsets_func.build_neighborhoods (levelsets_func.py:158-163).
#include "gridview.cuh"

namespace m3d {

template <int KMAX>
struct KnnList {
    double d[KMAX];
    long long id[KMAX];
    int cnt, k;
    __device__ __forceinline__ void init(int k_) { cnt = 0; k = k_; }
    __device__ __forceinline__ double worst() const { return (cnt < k) ? INFINITY : d[k - 1]; }
    // keep the k smallest by (distance, index)
    __device__ __forceinline__ void push(double d2, long long idx) {
        if (cnt == k) {
            if (d2 > d[k - 1] || (d2 == d[k - 1] && idx > id[k - 1])) return;
        } else {
            ++cnt;
        }
        int p = cnt - 1;
        while (p > 0 && (d[p - 1] > d2 || (d[p - 1] == d2 && id[p - 1] > idx))) {
            d[p] = d[p - 1]; id[p] = id[p - 1];
            --p;
        }
        d[p] = d2; id[p] = idx;
    }
};

template <int KMAX, bool SELF>
__global__ void __launch_bounds__(128) knn_kernel(GridView g, const double* __restrict__ q, int64_t nq, int k,
                                                  double cell, const uint32_t* __restrict__ qorder,
                                                  int64_t* __restrict__ idx_out,
                                                  double* __restrict__ dist_out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    double px, py, pz;
    int64_t out;
    if (SELF) {
        PointRec p = load_rec(g.rec, i);
        px = p.x; py = p.y; pz = p.z; out = p.idx;
    } else {
        out = qorder ? (int64_t)qorder[i] : i;      // external queries are walked in grid-column order
        px = q[3 * out]; py = q[3 * out + 1]; pz = q[3 * out + 2];
    }
    KnnList<KMAX> L;
    L.init(k);
    const int ix = min(max(g.col_x(px), -1), g.nx);
    const int iy = min(max(g.col_y(py), -1), g.ny);

    auto scan_range = [&](int y, int xa, int xb) {   // columns [xa, xb] of row y, clipped
        if (y < 0 || y >= g.ny) return;
        xa = max(xa, 0); xb = min(xb, g.nx - 1);
        if (xa > xb) return;
        const size_t row = (size_t)y * g.nx;
        const uint32_t b = __ldg(g.col_start + row + xa), e = __ldg(g.col_start + row + xb + 1);
        for (uint32_t j = b; j < e; ++j) {
            const PointRec c = load_rec(g.rec, j);
            const double d2 = dist2_rule(px - c.x, py - c.y, pz - c.z);
            L.push(d2, c.idx);
        }
    };

    const int smax = max(max(ix + 1, g.nx - ix), max(iy + 1, g.ny - iy));
    for (int s = 0; s <= smax; ++s) {
        if (s == 0) {
            scan_range(iy, ix, ix);
        } else {
            scan_range(iy - s, ix - s, ix + s);
            scan_range(iy + s, ix - s, ix + s);
            for (int y = iy - s + 1; y <= iy + s - 1; ++y) {
                scan_range(y, ix - s, ix - s);
                scan_range(y, ix + s, ix + s);
            }
        }
        if (L.cnt == k) {
            // every point whose column lies outside the searched block is farther than `bound`
            const double lx = px - (g.ox + (double)(ix - s) * cell);
            const double hx = (g.ox + (double)(ix + s + 1) * cell) - px;
            const double ly = py - (g.oy + (double)(iy - s) * cell);
            const double hy = (g.oy + (double)(iy + s + 1) * cell) - py;
            double bound = fmin(fmin(lx, hx), fmin(ly, hy)) - 1e-9 * cell;
            if (bound > 0.0 && L.worst() < bound * bound) break;
        }
    }
    for (int j = 0; j < k; ++j) {
        idx_out[out * k + j] = (j < L.cnt) ? (int64_t)L.id[j] : (int64_t)-1;
        if (dist_out) dist_out[out * k + j] = (j < L.cnt) ? sqrt(L.d[j]) : INFINITY;
    }
}

}  // namespace m3d

using namespace m3d;

template <int KMAX>
static void launch_knn(const more3d_grid* g, const double* q, int64_t nq, int k, const uint32_t* qo, int64_t* idx,
                       double* dist, cudaStream_t s) {
    const unsigned nb = (unsigned)((nq + 127) / 128);
    if (q) knn_kernel<KMAX, false><<<nb, 128, 0, s>>>(make_view(g), q, nq, k, g->cell, qo, idx, dist);
    else knn_kernel<KMAX, true><<<nb, 128, 0, s>>>(make_view(g), nullptr, nq, k, g->cell, nullptr, idx, dist);
}

extern "C" int more3d_knn(const more3d_grid* g, const double* q, int64_t nq, int k, int64_t* idx, double* dist,
                          void* stream) {
    M3D_REQUIRE(g == nullptr || g->x_split == 1, "more3d_knn needs a grid with square columns (x_split = 1)");
    M3D_REQUIRE(g != nullptr, "grid is NULL");
    M3D_REQUIRE(nq >= 0, "nq < 0");
    M3D_REQUIRE(q != nullptr || nq == g->n, "self query (q == NULL) needs nq == n");
    M3D_REQUIRE(k >= 1 && k <= MORE3D_MAX_K, "k out of range (1..MORE3D_MAX_K)");
    M3D_REQUIRE((int64_t)k <= g->n, "k exceeds the number of points");
    if (nq == 0) return MORE3D_OK;
    M3D_REQUIRE(idx != nullptr, "idx is NULL");
    cudaStream_t s = (cudaStream_t)stream;
    uint32_t* qo = nullptr;
    if (q) M3D_TRY(sort_queries_by_column(g, q, nq, s, &qo));
    if (k <= 4) launch_knn<4>(g, q, nq, k, qo, idx, dist, s);
    else if (k <= 16) launch_knn<16>(g, q, nq, k, qo, idx, dist, s);
    else launch_knn<64>(g, q, nq, k, qo, idx, dist, s);
    dev_free(qo, s);
    count_launches(1);
    M3D_CHECK_LAUNCH();
    return MORE3D_OK;
}
