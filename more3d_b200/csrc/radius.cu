// K3: fixed-radius neighbour query -> CSR (count, scan, fill).
// Replaces BallTreePointSet.queryRadius (Downsampling/BallTreePointSet.py:317-347).
#include "gridview.cuh"

namespace m3d {

// SELF: queries are the grid's own points, walked in sorted order (neighbouring threads share
// stencil rows, so candidate loads are L1 broadcasts); results land at the ORIGINAL index.
template <bool SELF, bool TALL>
__global__ void __launch_bounds__(256) radius_count_kernel(GridView g, const double* __restrict__ q,
                                                           int64_t nq, double r, double r2, int s, Filter32 flt,
                                                           const uint32_t* __restrict__ qorder,
                                                           int32_t* __restrict__ counts) {
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
    int c = 0;
    if (!SELF) flt = filter_for_query(g, flt, px, py, pz);
    for_each_neighbour<TALL>(g, flt, r, r2, s, px, py, pz, [&](int64_t) { ++c; });
    counts[out] = c;
}

template <bool SELF, typename IdxT, bool TALL>
__global__ void __launch_bounds__(256) radius_fill_kernel(GridView g, const double* __restrict__ q,
                                                          int64_t nq, double r, double r2, int s, Filter32 flt,
                                                          const uint32_t* __restrict__ qorder,
                                                          const int64_t* __restrict__ offsets,
                                                          IdxT* __restrict__ idx) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    double px, py, pz;
    int64_t out;
    if (SELF) {
        PointRec p = load_rec(g.rec, i);
        px = p.x; py = p.y; pz = p.z; out = p.idx;
    } else {
        out = qorder ? (int64_t)qorder[i] : i;
        px = q[3 * out]; py = q[3 * out + 1]; pz = q[3 * out + 2];
    }
    int64_t w = offsets[out];
    if (!SELF) flt = filter_for_query(g, flt, px, py, pz);
    for_each_neighbour<TALL>(g, flt, r, r2, s, px, py, pz, [&](int64_t j) {
        idx[w++] = (IdxT)__double_as_longlong(__ldg(reinterpret_cast<const double*>(g.rec + j) + 3));
    });
}

}  // namespace m3d

using namespace m3d;

static int check_query(const more3d_grid* g, const double* q, int64_t nq, double r) {
    M3D_REQUIRE(g != nullptr, "grid is NULL");
    M3D_REQUIRE(nq >= 0, "nq < 0");
    M3D_REQUIRE(r >= 0 && isfinite(r), "radius must be finite and >= 0");
    M3D_REQUIRE(q != nullptr || nq == g->n, "self query (q == NULL) needs nq == n");
    return MORE3D_OK;
}

extern "C" int more3d_radius_count(const more3d_grid* g, const double* q, int64_t nq, double r,
                                   int32_t* counts, void* stream) {
    M3D_TRY(check_query(g, q, nq, r));
    if (nq == 0) return MORE3D_OK;
    M3D_REQUIRE(counts != nullptr, "counts is NULL");
    cudaStream_t s = (cudaStream_t)stream;
    const int sw = stencil_halfwidth(g, r);
    const unsigned nb = (unsigned)((nq + 255) / 256);
    ProfScope prof("radius_count", s);
    const Filter32 flt = make_filter(g, r);
    GridView v = make_view(g);
    set_row_clip(v, g, r);
    uint32_t* qo = nullptr;
    if (q) M3D_TRY(sort_queries_by_column(g, q, nq, s, &qo));
    if (q) {
        if (g->tall) radius_count_kernel<false, true><<<nb, 256, 0, s>>>(v, q, nq, r, r * r, sw, flt, qo, counts);
        else radius_count_kernel<false, false><<<nb, 256, 0, s>>>(v, q, nq, r, r * r, sw, flt, qo, counts);
    } else {
        if (g->tall) radius_count_kernel<true, true><<<nb, 256, 0, s>>>(v, nullptr, nq, r, r * r, sw, flt, nullptr, counts);
        else radius_count_kernel<true, false><<<nb, 256, 0, s>>>(v, nullptr, nq, r, r * r, sw, flt, nullptr, counts);
    }
    dev_free(qo, s);
    count_launches(1);
    M3D_CHECK_LAUNCH();
    return MORE3D_OK;
}

extern "C" int more3d_offsets_from_counts(const int32_t* counts, int64_t nq, int64_t* offsets,
                                          int64_t* total_h, void* stream) {
    M3D_REQUIRE(offsets != nullptr && (counts != nullptr || nq == 0), "NULL argument");
    cudaStream_t s = (cudaStream_t)stream;
    M3D_TRY(exclusive_scan_i32_to_i64(counts, offsets, nq, s));
    if (total_h) {
        M3D_CUDA(cudaMemcpyAsync(total_h, offsets + nq, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
        M3D_CUDA(cudaStreamSynchronize(s));
    }
    return MORE3D_OK;
}

extern "C" int more3d_radius_fill(const more3d_grid* g, const double* q, int64_t nq, double r,
                                  const int64_t* offsets, void* idx, int idx_bytes, void* stream) {
    M3D_TRY(check_query(g, q, nq, r));
    if (nq == 0) return MORE3D_OK;
    M3D_REQUIRE(offsets != nullptr && idx != nullptr, "NULL argument");
    M3D_REQUIRE(idx_bytes == 4 || idx_bytes == 8, "idx_bytes must be 4 or 8");
    cudaStream_t s = (cudaStream_t)stream;
    const int sw = stencil_halfwidth(g, r);
    const unsigned nb = (unsigned)((nq + 255) / 256);
    GridView v = make_view(g);
    set_row_clip(v, g, r);
    const double r2 = r * r;
    const Filter32 flt = make_filter(g, r);
    ProfScope prof("radius_fill", s);
    uint32_t* qo = nullptr;
    if (q) M3D_TRY(sort_queries_by_column(g, q, nq, s, &qo));
#define M3D_FILL(SELF_, T_, TALL_, Q_) radius_fill_kernel<SELF_, T_, TALL_><<<nb, 256, 0, s>>>(v, Q_, nq, r, r2, sw, flt, qo, offsets, (T_*)idx)
    if (q) {
        if (idx_bytes == 4) { if (g->tall) M3D_FILL(false, int32_t, true, q); else M3D_FILL(false, int32_t, false, q); }
        else { if (g->tall) M3D_FILL(false, int64_t, true, q); else M3D_FILL(false, int64_t, false, q); }
    } else {
        if (idx_bytes == 4) { if (g->tall) M3D_FILL(true, int32_t, true, nullptr); else M3D_FILL(true, int32_t, false, nullptr); }
        else { if (g->tall) M3D_FILL(true, int64_t, true, nullptr); else M3D_FILL(true, int64_t, false, nullptr); }
    }
#undef M3D_FILL
    dev_free(qo, s);
    count_launches(1);
    M3D_CHECK_LAUNCH();
    return MORE3D_OK;
}
