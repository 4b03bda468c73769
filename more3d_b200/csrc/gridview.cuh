// Device-side view of the grid and the stencil walks shared by every neighbourhood kernel.
#pragma once
#include "common.cuh"

#ifndef M3D_PREFETCH_DEPTH
#define M3D_PREFETCH_DEPTH 1
#endif

namespace m3d {

struct GridView {
    const PointRec* rec;
    const float4* rec32;
    const uint32_t* col_start;
    const uint8_t* dup;
    int64_t n;
    int nx, ny;
    double ox, oy, oz, inv;   // inv: 1 / row height (y)
    double inv_x;             // 1 / column width (x)
    double lox, loy;      // lattice origin and integer column shift (see more3d_grid)
    int x0, y0;
    double lim32;
    int tall;
    unsigned char sxr[16];   // column half-width of the stencil row at |row offset| j (circle-clipped rows), 255 = unclipped
    int64_t feat_n;       // n when the feature records are split into two half arrays, 0 when interleaved (features.cu)
    __device__ __forceinline__ int col_x(double x) const { return cell_coord(x, lox, inv_x, x0); }
    __device__ __forceinline__ int col_y(double y) const { return cell_coord(y, loy, inv, y0); }
};

static inline GridView make_view(const more3d_grid* g) {
    GridView v;
    v.rec = g->rec; v.rec32 = g->rec32; v.col_start = g->col_start; v.dup = g->dup; v.n = g->n;
    v.nx = g->nx; v.ny = g->ny; v.ox = g->ox; v.oy = g->oy; v.oz = g->oz; v.inv = g->inv_cell;
    v.lim32 = g->lim32;
    v.inv_x = g->inv_x;
    v.lox = g->lox; v.loy = g->loy; v.x0 = g->x0; v.y0 = g->y0;
    v.tall = g->tall;
    v.feat_n = g->feat_split ? g->n : 0;
    for (int j = 0; j < 16; ++j) v.sxr[j] = 255;
    return v;
}

// stencil half-width (in columns) that is guaranteed to contain every point within r
// Returned packed: rows (y) in the low 16 bits, columns (x) in the high 16 bits -- the walkers take one int.
static inline int stencil_halfwidth(const more3d_grid* g, double r) {
    double sy = ceil(r * M3D_CELL_MARGIN / g->cell), sx = ceil(r * M3D_CELL_MARGIN / g->cell_x);
    if (sy > (double)g->ny) sy = (double)g->ny;
    if (sx > (double)g->nx) sx = (double)g->nx;
    if (sy < 1) sy = 1;
    if (sx < 1) sx = 1;
    if (sy > 32767.0) sy = 32767.0;         // wider than any table this library builds (<= 2^28 columns)
    if (sx > 32767.0) sx = 32767.0;
    return (int)sy | ((int)sx << 16);
}
#define M3D_SY(s) ((s) & 0xffff)
#define M3D_SX(s) ((s) >> 16)

// Circle-clipped stencil rows, per ROW OFFSET (not per query: every lane of a warp uses the same widths, so nothing
// diverges and no register is spent).  A point in the row j rows away from the query's row is at least (j - 1) row
// heights away in y, so it can only be within r if |dx| <= sqrt(r^2 - ((j - 1) cell)^2): the outer rows of a wide stencil
// (9 rows of quarter-radius cells at r = 1.0 on the shared grid) need 13 and 15 columns instead of 17.
static inline void set_row_clip(GridView& v, const more3d_grid* g, double r) {
    const int s = stencil_halfwidth(g, r), sy = M3D_SY(s), sx = M3D_SX(s);
    for (int j = 0; j < 16; ++j) {
        int w = sx;
        if (j >= 2 && j <= sy) {
            const double dy = (double)(j - 1) * g->cell * (1.0 - 1e-9);
            const double w2 = r * r - dy * dy;
            const double dx = w2 > 0.0 ? sqrt(w2) * (1.0 + 1e-9) : 0.0;
            const double c = ceil(dx * M3D_CELL_MARGIN / g->cell_x);
            if (c < (double)w) w = (int)c;
        }
        v.sxr[j] = (unsigned char)(w > 254 ? 255 : w);
    }
}
__device__ __forceinline__ int row_halfwidth(const GridView& g, int s, int dy) {
    return min(M3D_SX(s), (int)g.sxr[min(abs(dy), 15)]);
}

// fp32 distance filter.  With every fp32 record coordinate within delta32 of the exact one, the fp32
// squared distance of a pair near the sphere surface is within h of the exact fp64 rule's value:
//   |d2_32 - d2| <= (2 delta + 2^-24 r)(2 sqrt(3) r + ...) + 3 * 2^-24 r^2  <  8 delta r + 1e-6 r^2
// (derivation in DESIGN.md §3).  Twice that is used.  d2_32 < lo: certainly inside; d2_32 > hi:
// certainly outside; otherwise the pair is re-evaluated with the exact fp64 rule (dist2_rule).
struct Filter32 {
    float lo, hi;
};
static inline Filter32 make_filter(const more3d_grid* g, double r) {
    const double r2 = r * r, d = g->delta32;
    const double h = 16.0 * d * r + 4e-6 * r2 + 64.0 * d * d;
    Filter32 f;
    if (!(h < 0.25 * r2)) {          // extent too large for this radius: everything goes the exact way
        f.lo = -1.0f;
        f.hi = INFINITY;
    } else {
        f.lo = (float)(r2 - h) * (1.0f - 1e-6f);
        f.hi = (float)(r2 + h) * (1.0f + 1e-6f);
    }
    return f;
}

// a query far outside the cloud's box has a larger fp32 rounding error than delta32: exact path only
__device__ __forceinline__ Filter32 filter_for_query(const GridView& g, Filter32 f, double px, double py, double pz) {
    const bool ok = fabs(px - g.ox) < g.lim32 && fabs(py - g.oy) < g.lim32 && fabs(pz - g.oz) < g.lim32;
    if (!ok) { f.lo = -1.0f; f.hi = INFINITY; }
    return f;
}

// One 256-bit read-only load per 32-byte record (sm_100: LDG.E.256) instead of two LDG.128.  The hot kernels
// are bound by the L1 data pipe (ncu: l1tex__data_pipe_lsu_wavefronts 84-92 %); measured, the wider load
// changes nothing (8.3 vs 8.5 ms, 15.9 vs 15.8 ms): the pipe moves 32 B per lane per candidate either way
// (1 KiB per warp = 8 wavefronts), so only SMALLER records would help, not fewer instructions.
__device__ __forceinline__ void ldg256(const void* p, double& a, double& b, double& c, double& d) {
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
__device__ __forceinline__ PointRec load_rec(const PointRec* __restrict__ rec, int64_t j) {
    PointRec r;
    double w;
    ldg256(rec + j, r.x, r.y, r.z, w);
    r.idx = __double_as_longlong(w);
    return r;
}

// Walk the (2s+1) stencil rows of the column containing (qx, qy); each row is one contiguous
// range of sorted records.  f(j) is called with the sorted position of every candidate.
template <typename F>
__device__ __forceinline__ void for_each_candidate(const GridView& g, double qx, double qy, int s, F&& f) {
    const int ix = g.col_x(qx);
    const int iy = g.col_y(qy);
    const int y0 = max(iy - M3D_SY(s), 0), y1 = min(iy + M3D_SY(s), g.ny - 1);
    if (ix + M3D_SX(s) < 0 || ix - M3D_SX(s) > g.nx - 1 || y0 > y1) return;
    // the bounds of row y+1 are requested before row y is walked, so their latency hides behind the row
    const uint32_t* cs = g.col_start + (size_t)y0 * g.nx;
    int w = row_halfwidth(g, s, y0 - iy);
    uint32_t nb = __ldg(cs + min(max(ix - w, 0), g.nx)), ne = __ldg(cs + min(max(ix + w + 1, 0), g.nx));
    for (int y = y0; y <= y1; ++y) {
        const uint32_t b = nb, e = ne;
        if (y < y1) {
            cs += g.nx;
            w = row_halfwidth(g, s, y + 1 - iy);
            nb = __ldg(cs + min(max(ix - w, 0), g.nx));
            ne = __ldg(cs + min(max(ix + w + 1, 0), g.nx));
        }
        for (uint32_t j = b; j < e; ++j) f((int64_t)j);
    }
}

// Tall clouds (vertical structures: some column holds > M3D_TALL_COLUMN points; columns are z-sorted):
// walk the stencil column by column and bracket long columns to z in [qz - r, qz + r] by binary search.
// The bracket is widened by one ulp-scale slack and the exact rule still decides, so the candidate set is a
// superset of the neighbourhood exactly as in the row walk.
template <typename F>
__device__ __forceinline__ void for_each_candidate_tall(const GridView& g, double qx, double qy, double qz, double r,
                                                     int s, F&& f) {
    const int ix = g.col_x(qx);
    const int iy = g.col_y(qy);
    const int x0 = max(ix - M3D_SX(s), 0), x1 = min(ix + M3D_SX(s), g.nx - 1);
    const int y0 = max(iy - M3D_SY(s), 0), y1 = min(iy + M3D_SY(s), g.ny - 1);
    const double zlo = qz - r * (1.0 + 1e-12), zhi = qz + r * (1.0 + 1e-12);
    for (int y = y0; y <= y1; ++y)
        for (int x = x0; x <= x1; ++x) {
            const size_t col = (size_t)y * g.nx + x;
            uint32_t b = __ldg(g.col_start + col), e = __ldg(g.col_start + col + 1);
            if (e - b > 32u) {
                uint32_t lo = b, hi = e;                     // first record with z >= zlo
                while (lo < hi) {
                    const uint32_t mid = lo + ((hi - lo) >> 1);
                    if (__ldg(reinterpret_cast<const double*>(g.rec + mid) + 2) < zlo) lo = mid + 1; else hi = mid;
                }
                b = lo;
                hi = e;                                       // first record with z > zhi
                while (lo < hi) {
                    const uint32_t mid = lo + ((hi - lo) >> 1);
                    if (__ldg(reinterpret_cast<const double*>(g.rec + mid) + 2) <= zhi) lo = mid + 1; else hi = mid;
                }
                e = lo;
            }
            for (uint32_t j = b; j < e; ++j) f((int64_t)j);
        }
}

// Software-pipelined variant of the walks: load(j) returns the candidate's record (any POD), process(rec)
// consumes it.  The record of candidate j+1 is requested before candidate j is processed, so a warp never
// sits on the L1 latency of the record it is about to use (one exposed latency per row instead of one per
// candidate).
template <typename L, typename P>
__device__ __forceinline__ void run_range_pipelined(uint32_t b, uint32_t e, L&& load, P&& process) {
    if (b >= e) return;
#if M3D_PREFETCH_DEPTH == 2
    // two records in flight: candidate j + 2 is requested before candidate j is processed (reads past the end of the
    // range are clamped to its last record and dropped); the three buffers rotate by unrolling, not by moves
    const uint32_t last = e - 1;
    auto r0 = load((int64_t)b);
    auto r1 = load((int64_t)min(b + 1, last));
    uint32_t j = b;
    while (true) {
        auto r2 = load((int64_t)min(j + 2, last));
        process(r0);
        if (++j >= e) break;
        r0 = load((int64_t)min(j + 2, last));
        process(r1);
        if (++j >= e) break;
        r1 = load((int64_t)min(j + 2, last));
        process(r2);
        if (++j >= e) break;
    }
#else
    auto cur = load((int64_t)b);
#pragma unroll 2
    for (uint32_t j = b + 1; j < e; ++j) {
        auto nxt = load((int64_t)j);
        process(cur);
        cur = nxt;
    }
    process(cur);
#endif
}

template <bool TALL, typename L, typename P>
__device__ __forceinline__ void walk_candidates_pipelined(const GridView& g, double qx, double qy, double qz, double r,
                                                          int s, L&& load, P&& process) {
    if (TALL) {
        for_each_candidate_tall(g, qx, qy, qz, r, s, [&](int64_t j) { process(load(j)); });
        return;
    }
    const int ix = g.col_x(qx);
    const int iy = g.col_y(qy);
    const int y0 = max(iy - M3D_SY(s), 0), y1 = min(iy + M3D_SY(s), g.ny - 1);
    if (ix + M3D_SX(s) < 0 || ix - M3D_SX(s) > g.nx - 1 || y0 > y1) return;
    const uint32_t* cs = g.col_start + (size_t)y0 * g.nx;
    int w = row_halfwidth(g, s, y0 - iy);
    uint32_t nb = __ldg(cs + min(max(ix - w, 0), g.nx)), ne = __ldg(cs + min(max(ix + w + 1, 0), g.nx));
    for (int y = y0; y <= y1; ++y) {
        const uint32_t b = nb, e = ne;
        if (y < y1) {
            cs += g.nx;
            w = row_halfwidth(g, s, y + 1 - iy);
            nb = __ldg(cs + min(max(ix - w, 0), g.nx));
            ne = __ldg(cs + min(max(ix + w + 1, 0), g.nx));
        }
        run_range_pipelined(b, e, load, process);
    }
}

// compile-time choice between the row walk (thin columns) and the column walk (tall columns): the host
// launches the <TALL> instantiation that matches the grid, so the thin path's code is unchanged
template <bool TALL, typename F>
__device__ __forceinline__ void walk_candidates(const GridView& g, double qx, double qy, double qz, double r, int s,
                                                F&& f) {
    if (TALL) for_each_candidate_tall(g, qx, qy, qz, r, s, f);
    else for_each_candidate(g, qx, qy, s, f);
}

// Exact-set neighbour enumeration with the fp32 filter: accept(j) is called for every sorted position
// j whose point satisfies the fp64 rule ((dx*dx)+(dy*dy))+(dz*dz) <= r2 -- bit-identical to the
// all-fp64 walk, only cheaper: fp64 is touched for the thin shell around the sphere surface only.
template <bool TALL, typename F>
__device__ __forceinline__ void for_each_neighbour(const GridView& g, const Filter32 flt, double r, double r2, int s,
                                                   double px, double py, double pz, F&& accept) {
    const float qx = (float)(px - g.ox), qy = (float)(py - g.oy), qz = (float)(pz - g.oz);
    walk_candidates<TALL>(g, px, py, pz, r, s, [&](int64_t j) {
        const float4 c = __ldg(g.rec32 + j);
        const float dx = c.x - qx, dy = c.y - qy, dz = c.z - qz;
        const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        if (d2 <= flt.hi) {
            bool in = d2 < flt.lo;
            if (!in) {
                const PointRec e = load_rec(g.rec, j);
                in = dist2_rule(px - e.x, py - e.y, pz - e.z) <= r2;
            }
            if (in) accept(j);
        }
    });
}

}  // namespace m3d
