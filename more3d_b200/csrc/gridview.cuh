// Device-side view of the grid and the stencil walk shared by every neighbourhood kernel.
#pragma once
#include "common.cuh"

namespace m3d {

struct GridView {
    const PointRec* rec;
    const uint32_t* col_start;
    const uint8_t* dup;
    int64_t n;
    int nx, ny;
    double ox, oy, inv;
};

static inline GridView make_view(const more3d_grid* g) {
    GridView v;
    v.rec = g->rec; v.col_start = g->col_start; v.dup = g->dup; v.n = g->n;
    v.nx = g->nx; v.ny = g->ny; v.ox = g->ox; v.oy = g->oy; v.inv = g->inv_cell;
    return v;
}

// stencil half-width (in columns) that is guaranteed to contain every point within r
static inline int stencil_halfwidth(const more3d_grid* g, double r) {
    double s = ceil(r * M3D_CELL_MARGIN / g->cell);
    double lim = (double)((g->nx > g->ny) ? g->nx : g->ny);
    if (s > lim) s = lim;
    if (s < 1) s = 1;
    return (int)s;
}

__device__ __forceinline__ PointRec load_rec(const PointRec* __restrict__ rec, int64_t j) {
    const double2* p = reinterpret_cast<const double2*>(rec + j);
    double2 a = __ldg(p), b = __ldg(p + 1);
    PointRec r;
    r.x = a.x; r.y = a.y; r.z = b.x; r.idx = __double_as_longlong(b.y);
    return r;
}

// Walk the (2s+1) stencil rows of the column containing (qx, qy); each row is one contiguous
// range of sorted records.  f(j) is called with the sorted position of every candidate.
template <typename F>
__device__ __forceinline__ void for_each_candidate(const GridView& g, double qx, double qy, int s, F&& f) {
    const int ix = cell_coord(qx, g.ox, g.inv, g.nx);
    const int iy = cell_coord(qy, g.oy, g.inv, g.ny);
    const int x0 = max(ix - s, 0), x1 = min(ix + s, g.nx - 1);
    const int y0 = max(iy - s, 0), y1 = min(iy + s, g.ny - 1);
    if (x0 > x1) return;
    for (int y = y0; y <= y1; ++y) {
        const size_t row = (size_t)y * g.nx;
        const uint32_t b = __ldg(g.col_start + row + x0);
        const uint32_t e = __ldg(g.col_start + row + x1 + 1);
        for (uint32_t j = b; j < e; ++j) f((int64_t)j);
    }
}

}  // namespace m3d
