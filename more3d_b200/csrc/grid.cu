// Uniform xy-column grid: the spatial index that replaces the reference's ball tree
// (Downsampling/BallTreePointSet.py:19-53 -> sklearn BallTree build).
//
// Layout in HBM after more3d_grid_create:
//   rec[n]        sorted 32-B records (x, y, z, original index), order = (column key, original index)
//   order[n]      sorted position -> original index (u32)
//   col_start[]   nx*ny + 1 exclusive offsets; column (ix, iy) owns rec[col_start[iy*nx+ix] .. col_start[iy*nx+ix+1])
//   dup[n]        1 if the sorted point coincides (bit-equal xyz) with an earlier sorted point
// Columns are x-fastest, so the (2s+1) columns of one stencil row are ONE contiguous range of rec.
#include "common.cuh"

namespace m3d {

constexpr int BB_THREADS = 256;

__device__ __forceinline__ PointRec load_rec_plain(const PointRec* __restrict__ rec, int64_t j) {
    PointRec r;
    double w;
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(w) : "l"(rec + j));
    r.idx = __double_as_longlong(w);
    return r;
}

// The cloud of a grid is the concatenation of up to three device segments (a strip's left ghosts, its own points,
// its right ghosts): the kernels that read the caller's coordinates index the segments directly, nothing is
// concatenated in memory.  end[k] = number of points in segments 0..k.
struct Segs {
    const double* p[3];
    int64_t end[3];
};
__device__ __forceinline__ const double* seg_point(const Segs& sg, int64_t i) {
    if (i < sg.end[0]) return sg.p[0] + 3 * i;
    if (i < sg.end[1]) return sg.p[1] + 3 * (i - sg.end[0]);
    return sg.p[2] + 3 * (i - sg.end[1]);
}

__global__ void __launch_bounds__(BB_THREADS) bbox_partial(Segs sg, int64_t n,
                                                           double* __restrict__ part) {
    // v[0..2] = min x,y,z   v[3..5] = max x,y,z (stored negated so one fmin reduction serves both)
    double v[6] = {INFINITY, INFINITY, INFINITY, INFINITY, INFINITY, INFINITY};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const double* __restrict__ q = seg_point(sg, i);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double c = q[a];
            v[a] = fmin(v[a], c);
            v[3 + a] = fmin(v[3 + a], -c);
        }
    }
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[a] = fmin(v[a], __shfl_xor_sync(0xffffffffu, v[a], o));
    __shared__ double sm[6][BB_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0)
        for (int a = 0; a < 6; ++a) sm[a][warp] = v[a];
    __syncthreads();
    if (threadIdx.x < 6) {
        double r = sm[threadIdx.x][0];
        for (int w = 1; w < BB_THREADS / 32; ++w) r = fmin(r, sm[threadIdx.x][w]);
        part[6 * blockIdx.x + threadIdx.x] = r;
    }
}

// out = {min x, min y, min z, max x, max y, max z}
__global__ void bbox_final(const double* __restrict__ part, int nparts, double* __restrict__ out) {
    double v[6] = {INFINITY, INFINITY, INFINITY, INFINITY, INFINITY, INFINITY};
    for (int i = threadIdx.x; i < nparts; i += 32)
#pragma unroll
        for (int a = 0; a < 6; ++a) v[a] = fmin(v[a], part[6 * i + a]);
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[a] = fmin(v[a], __shfl_xor_sync(0xffffffffu, v[a], o));
    if (threadIdx.x == 0)
        for (int a = 0; a < 6; ++a) out[a] = (a < 3) ? v[a] : -v[a];
}

// the column lattice as the build kernels need it: column of (x, y) = (floor((x - lox) * inv) - x0, ...), clamped
struct Lattice {
    double lox, loy, inv, inv_x;
    int x0, y0, nx, ny;
    __device__ __forceinline__ uint32_t key(double x, double y) const {
        const int ix = min(max(cell_coord(x, lox, inv_x, x0), 0), nx - 1);
        const int iy = min(max(cell_coord(y, loy, inv, y0), 0), ny - 1);
        return (uint32_t)iy * (uint32_t)nx + (uint32_t)ix;
    }
};
static Lattice make_lattice(const more3d_grid* g) {
    Lattice l;
    l.lox = g->lox; l.loy = g->loy; l.inv = g->inv_cell; l.inv_x = g->inv_x; l.x0 = g->x0; l.y0 = g->y0; l.nx = g->nx; l.ny = g->ny;
    return l;
}

// K1: column key per point + column population (spread-address REDG)
__global__ void __launch_bounds__(256) cell_keys(const double* __restrict__ xyz, int64_t n, Lattice lat,
                                                 uint32_t* __restrict__ keys, uint32_t* __restrict__ vals,
                                                 uint32_t* __restrict__ col_count) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t key = lat.key(xyz[3 * i], xyz[3 * i + 1]);
    keys[i] = key;
    vals[i] = (uint32_t)i;
    atomicAdd(&col_count[key], 1u);
}

// K1 (thin grids): column key per point AND the point's arrival rank in its column -- the value the counting
// atomic returns.  Arrival order is arbitrary; column_fixup below turns it into (column, original index) order.
__global__ void __launch_bounds__(256) cell_keys_rank(Segs sg, int64_t n, Lattice lat,
                                                      uint32_t* __restrict__ keys, uint32_t* __restrict__ ranks,
                                                      uint32_t* __restrict__ col_count) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* __restrict__ q = seg_point(sg, i);
    const uint32_t key = lat.key(q[0], q[1]);
    keys[i] = key;
    ranks[i] = atomicAdd(&col_count[key], 1u);
}

// K2 (thin grids): counting-sort scatter.  Reads the cloud coalesced, writes one full 32-B sector per point at
// col_start[key] + arrival rank: the points of a column are now contiguous, in arrival order.
__global__ void __launch_bounds__(256) scatter_records(Segs sg, int64_t n,
                                                       const uint32_t* __restrict__ keys,
                                                       const uint32_t* __restrict__ ranks,
                                                       const uint32_t* __restrict__ col_start,
                                                       PointRec* __restrict__ tmp) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* __restrict__ q = seg_point(sg, i);
    PointRec r;
    r.x = q[0]; r.y = q[1]; r.z = q[2];
    r.idx = (long long)i;
    tmp[(size_t)__ldg(col_start + keys[i]) + ranks[i]] = r;
}

// K3 (thin grids): one thread per scattered record.  Its final slot is the column start plus the number of column
// members with a smaller original index (columns hold a handful of points: a short loop over L1-resident
// records), which makes the order (column key, original index) -- exactly what a stable sort by key gives, for
// any arrival order.  The same loop finds the np.unique flag (a bit-equal point with a smaller index,
// DM_saliency.py:247); the fp32 record and the order table are written alongside.
__global__ void __launch_bounds__(256) column_fixup(const PointRec* __restrict__ tmp,
                                                    const uint32_t* __restrict__ col_start, int64_t n, Lattice lat,
                                                    double ox, double oy, double oz,
                                                    PointRec* __restrict__ rec, float4* __restrict__ rec32,
                                                    uint32_t* __restrict__ order, uint8_t* __restrict__ dup) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const PointRec p = load_rec_plain(tmp, t);
    const size_t col = lat.key(p.x, p.y);
    const uint32_t b = __ldg(col_start + col), e = __ldg(col_start + col + 1);
    uint32_t rank = 0;
    uint8_t f = 0;
    for (uint32_t j = b; j < e; ++j) {
        const PointRec q = load_rec_plain(tmp, j);
        if (q.idx < p.idx) {
            ++rank;
            if (q.x == p.x && q.y == p.y && q.z == p.z) f = 1;
        }
    }
    const size_t dst = (size_t)b + rank;
    rec[dst] = p;
    rec32[dst] = make_float4((float)(p.x - ox), (float)(p.y - oy), (float)(p.z - oz), 0.0f);
    order[dst] = (uint32_t)p.idx;
    dup[dst] = f;
}

__global__ void __launch_bounds__(256) col_max_pop(const uint32_t* __restrict__ col_start, size_t ncols,
                                                   uint32_t* __restrict__ out) {
    uint32_t m = 0;
    for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < ncols; c += (size_t)gridDim.x * blockDim.x)
        m = max(m, col_start[c + 1] - col_start[c]);
    m = __reduce_max_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

// tall clouds: z pre-sort keys, and the column keys re-read in z order for the second (stable) sort
__global__ void __launch_bounds__(256) z_keys(Segs sg, int64_t n, uint64_t* __restrict__ zk,
                                              uint32_t* __restrict__ vals) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    zk[i] = sortable_key(seg_point(sg, i)[2]);
    vals[i] = (uint32_t)i;
}
__global__ void __launch_bounds__(256) gather_u32(const uint32_t* __restrict__ src, const uint32_t* __restrict__ idx,
                                                  int64_t n, uint32_t* __restrict__ dst) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
}

// K2 tail: gather the sorted records
__global__ void __launch_bounds__(256) gather_records(Segs sg,
                                                      const uint32_t* __restrict__ order, int64_t n,
                                                      PointRec* __restrict__ rec, double ox, double oy,
                                                      double oz, float4* __restrict__ rec32) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t o = order[i];
    const double* __restrict__ q = seg_point(sg, (int64_t)o);
    PointRec r;
    r.x = q[0];
    r.y = q[1];
    r.z = q[2];
    r.idx = (long long)o;
    rec[i] = r;
    // fp32 copy relative to the grid origin: |error| <= 2^-24 * extent per coordinate (delta32)
    rec32[i] = make_float4((float)(r.x - ox), (float)(r.y - oy), (float)(r.z - oz), 0.0f);
}

// coincident-point flags of the tall path: the np.unique(axis=1) of Saliency_Kernel.process (DM_saliency.py:247)
// keeps one representative of every set of bit-equal positions.  z-sorted columns: coincident points are adjacent
// (equal z, then original index), so only the run of equal z before i has to be looked at
__global__ void __launch_bounds__(256) dup_flags_tall(const PointRec* __restrict__ rec,
                                                      const uint32_t* __restrict__ col_start, int64_t n, Lattice lat,
                                                      uint8_t* __restrict__ dup) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    PointRec p = rec[i];
    const int64_t b = col_start[lat.key(p.x, p.y)];
    uint8_t f = 0;
    for (int64_t j = i - 1; j >= b; --j) {
        PointRec q = rec[j];
        if (q.z != p.z) break;
        if (q.x == p.x && q.y == p.y) { f = 1; break; }
    }
    dup[i] = f;
}

}  // namespace m3d

using namespace m3d;

extern "C" int more3d_grid_create(const double* xyz, int64_t n, double cell, void* stream,
                                  more3d_grid** out) {
    const double* seg[1] = {xyz};
    const int64_t cnt[1] = {n};
    return more3d_grid_create_segments(seg, cnt, 1, cell, 1, nullptr, nullptr, stream, out);
}

extern "C" int more3d_grid_create_segments(const double* const* seg_xyz_h, const int64_t* seg_n_h, int nseg, double cell,
                                           int x_split, const double* lattice_h, const double* bounds_h, void* stream,
                                           more3d_grid** out) {
    M3D_REQUIRE(x_split >= 1 && x_split <= 8, "x_split must be 1..8");
    M3D_REQUIRE(out != nullptr, "out is NULL");
    *out = nullptr;
    M3D_REQUIRE(seg_xyz_h != nullptr && seg_n_h != nullptr && nseg >= 1 && nseg <= 3, "1 to 3 segments");
    M3D_REQUIRE(cell > 0 && isfinite(cell), "cell must be > 0");
    Segs sg;
    int64_t n = 0;
    for (int k = 0; k < 3; ++k) {
        const int64_t c = k < nseg ? seg_n_h[k] : 0;
        M3D_REQUIRE(c >= 0 && (c == 0 || seg_xyz_h[k] != nullptr), "bad segment");
        sg.p[k] = k < nseg ? seg_xyz_h[k] : nullptr;
        n += c;
        sg.end[k] = n;
    }
    M3D_REQUIRE(n > 0, "empty cloud");
    if (n >= ((int64_t)1 << 31)) {
        set_error("n >= 2^31 points per grid not supported");
        return MORE3D_E_RANGE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    init_device_pool();
    ProfScope prof("grid_build", s);

    // bounding box (the host needs it to size the table): given by the caller, or reduced here (one small sync)
    double b6[6];
    if (bounds_h) {
        for (int a = 0; a < 6; ++a) b6[a] = bounds_h[a];
    } else {
        const int nparts = 592;  // 4 CTAs per SM on 148 SMs
        double* part = nullptr;
        M3D_TRY(dev_alloc(&part, (size_t)6 * nparts + 6, s));
        bbox_partial<<<nparts, BB_THREADS, 0, s>>>(sg, n, part);
        bbox_final<<<1, 32, 0, s>>>(part, nparts, part + 6 * nparts);
        count_launches(2);
        M3D_CHECK_LAUNCH();
        M3D_TRY(read_back(b6, part + 6 * nparts, sizeof(b6), s));
        dev_free(part, s);
    }
    const double bb[4] = {b6[0], b6[1], b6[3], b6[4]};   // min x, min y, max x, max y
    if (!(isfinite(b6[0]) && isfinite(b6[1]) && isfinite(b6[2]) && isfinite(b6[3]) && isfinite(b6[4]) &&
          isfinite(b6[5]))) {
        set_error("non-finite coordinates in cloud");
        return MORE3D_E_INVALID;
    }

    double ex = bb[2] - bb[0], ey = bb[3] - bb[1];
    more3d_grid* g = new more3d_grid();
    g->n = n;
    double c;
    if (lattice_h) {
        // columns of a lattice shared with other grids (the strips of one tile): origin and cell are given, this
        // grid covers the integer column range its own bounding box needs (one spare column either side)
        c = cell;
        const double inv = 1.0 / c, invx = 1.0 / (c / (double)x_split);
        g->lox = lattice_h[0];
        g->loy = lattice_h[1];
        const double fx0 = floor((bb[0] - g->lox) * invx) - 1.0, fx1 = floor((bb[2] - g->lox) * invx) + 1.0;
        const double fy0 = floor((bb[1] - g->loy) * inv) - 1.0, fy1 = floor((bb[3] - g->loy) * inv) + 1.0;
        if (!(fabs(fx0) < 1e9 && fabs(fx1) < 1e9 && fabs(fy0) < 1e9 && fabs(fy1) < 1e9) ||
            (fx1 - fx0 + 1.0) * (fy1 - fy0 + 1.0) > fmax(268435456.0, 2.0 * (double)n)) {
            delete g;
            set_error("column table of the shared lattice exceeds max(2^28, 2 n) entries: use a wider cell or a smaller x_split for every strip");
            return MORE3D_E_RANGE;
        }
        g->x0 = (int)fx0; g->y0 = (int)fy0;
        g->nx = (int)(fx1 - fx0) + 1; g->ny = (int)(fy1 - fy0) + 1;
        g->ox = g->lox + (double)g->x0 * (c / (double)x_split);
        g->oy = g->loy + (double)g->y0 * c;
    } else {
        // The column table may hold max(2^28, 2 n) entries (1 GiB, or 8 bytes per point on very large tiles).  If the
        // requested geometry does not fit, first give up the x refinement, then widen the cell (stencils stay correct
        // for any cell >= requested).
        double maxcols = fmax(268435456.0, 2.0 * (double)n);
        if (maxcols > 2147483647.0) maxcols = 2147483647.0;
        c = cell;
        auto cols = [&](double cc, int xs) { return (floor((ex + cc) / (cc / (double)xs)) + 2.0) * (floor(ey / cc) + 3.0); };
        while (x_split > 1 && cols(c, x_split) > maxcols) --x_split;
        while (cols(c, x_split) > maxcols) c *= 1.05;
        g->ox = bb[0] - c;
        g->oy = bb[1] - c;
        g->lox = g->ox; g->loy = g->oy; g->x0 = 0; g->y0 = 0;
        g->nx = (int)floor((ex + c) / (c / (double)x_split)) + 2;
        g->ny = (int)floor(ey / c) + 3;
    }
    g->cell = c;
    g->inv_cell = 1.0 / c;
    g->x_split = x_split;
    g->cell_x = c / (double)x_split;
    g->inv_x = 1.0 / g->cell_x;
    cudaGetDevice(&g->device);
    g->stream = s;
    g->oz = b6[2];
    {
        // largest |coordinate - origin| any record (or query clamped to the stencil) can have
        double ext = fmax(fmax(ex + 3.0 * c, ey + 3.0 * c), b6[5] - b6[2]);
        int e2 = 0;
        frexp(ext, &e2);                           // ext < 2^e2
        g->delta32 = ldexp(1.0, e2 - 24);
        g->lim32 = ldexp(1.0, e2);
    }
    g->rec32 = nullptr;
    g->feat = nullptr;
    g->feat_split = ((double)n / ((double)g->nx * (double)g->ny)) < 4.0 ? 1 : 0;
    g->order = nullptr; g->rec = nullptr; g->col_start = nullptr; g->dup = nullptr;
    const size_t ncols = (size_t)g->nx * g->ny;

    uint32_t *keys = nullptr, *vals = nullptr, *keys2 = nullptr, *vals2 = nullptr;
    PointRec* tmp = nullptr;
    int rc = MORE3D_OK;
#define G_TRY(e) do { rc = (e); if (rc != MORE3D_OK) goto fail; } while (0)
#define G_CUDA(e) do { cudaError_t _e = (e); if (_e != cudaSuccess) { set_error("%s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(_e)); rc = (_e == cudaErrorMemoryAllocation) ? MORE3D_E_NOMEM : MORE3D_E_CUDA; goto fail; } } while (0)
    {
        G_CUDA(cudaMallocAsync((void**)&g->col_start, (ncols + 1) * sizeof(uint32_t), s));
        G_CUDA(cudaMallocAsync((void**)&g->order, (size_t)n * sizeof(uint32_t), s));
        G_CUDA(cudaMallocAsync((void**)&g->rec, (size_t)n * sizeof(PointRec), s));
        G_CUDA(cudaMallocAsync((void**)&g->dup, (size_t)n, s));
        G_CUDA(cudaMallocAsync((void**)&g->rec32, (size_t)n * sizeof(float4), s));
        G_TRY(dev_alloc(&keys, (size_t)n, s));
        G_TRY(dev_alloc(&vals, (size_t)n, s));
        G_CUDA(cudaMemsetAsync(g->col_start, 0, (ncols + 1) * sizeof(uint32_t), s));
        const unsigned nb = (unsigned)((n + 255) / 256);
        const Lattice lat = make_lattice(g);
        // keys + arrival ranks (vals holds the ranks on the thin path, the point indices on the tall one)
        cell_keys_rank<<<nb, 256, 0, s>>>(sg, n, lat, keys, vals, g->col_start);
        G_CUDA(cudaGetLastError());
        G_TRY(exclusive_scan_u32(g->col_start, g->col_start, (int64_t)ncols, s));
        // thin or tall?  (one more 4-byte read-back per build)
        uint32_t* d_max = nullptr;
        G_TRY(dev_alloc(&d_max, 1, s));
        G_CUDA(cudaMemsetAsync(d_max, 0, sizeof(uint32_t), s));
        col_max_pop<<<592, 256, 0, s>>>(g->col_start, ncols, d_max);
        G_TRY(read_back(&g->max_col_pop, d_max, sizeof(uint32_t), s));
        dev_free(d_max, s);
        g->tall = g->max_col_pop > M3D_TALL_COLUMN ? 1 : 0;
        // thin grids: the two-pass dk/dn reads only the first halves of the feature records in its screening pass, so the
        // records are always split into two half arrays there (the one-pass kernel keeps the population heuristic)
        if (!g->tall && dkdn_variant() != 2) g->feat_split = 1;
        if (g->tall) {
            // order = (column, z, original index): stable radix sort by z first, then by column key
            int bits = 1;
            while (((size_t)1 << bits) < ncols) ++bits;
            uint32_t *ks = nullptr, *vs = nullptr;
            uint64_t *zk = nullptr, *zk2 = nullptr, *zs = nullptr;
            uint32_t* vz = nullptr;
            G_TRY(dev_alloc(&keys2, (size_t)n, s));
            G_TRY(dev_alloc(&vals2, (size_t)n, s));
            G_TRY(dev_alloc(&zk, (size_t)n, s));
            G_TRY(dev_alloc(&zk2, (size_t)n, s));
            z_keys<<<nb, 256, 0, s>>>(sg, n, zk, vals);
            rc = radix_sort_pairs_u64(zk, vals, zk2, vals2, n, 64, &zs, &vz, s);
            dev_free(zk, s); dev_free(zk2, s);
            if (rc != MORE3D_OK) goto fail;
            uint32_t* vother = (vz == vals) ? vals2 : vals;
            gather_u32<<<nb, 256, 0, s>>>(keys, vz, n, keys2);          // column key of the z-ordered points
            G_TRY(radix_sort_pairs_u32(keys2, vz, keys, vother, n, bits, &ks, &vs, s));
            G_CUDA(cudaMemcpyAsync(g->order, vs, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
            gather_records<<<nb, 256, 0, s>>>(sg, g->order, n, g->rec, g->ox, g->oy, g->oz, g->rec32);
            dup_flags_tall<<<nb, 256, 0, s>>>(g->rec, g->col_start, n, lat, g->dup);
            count_launches(6);  // cell_keys_rank, col_max_pop, z_keys, gather_u32, gather_records, dup_flags_tall
        } else {
            // counting sort: the column table is already the histogram of the WHOLE key, so one scatter places every
            // point in its column (arrival order) and one fix-up pass orders each column by original index -- the
            // order a stable sort by key produces, with no radix passes at all
            G_CUDA(cudaMallocAsync((void**)&tmp, (size_t)n * sizeof(PointRec), s));
            scatter_records<<<nb, 256, 0, s>>>(sg, n, keys, vals, g->col_start, tmp);
            column_fixup<<<nb, 256, 0, s>>>(tmp, g->col_start, n, lat, g->ox, g->oy, g->oz, g->rec, g->rec32, g->order,
                                            g->dup);
            count_launches(4);  // cell_keys_rank, col_max_pop, scatter_records, column_fixup
        }
        G_CUDA(cudaGetLastError());
    }
    dev_free(keys, s); dev_free(vals, s); dev_free(keys2, s); dev_free(vals2, s); dev_free(tmp, s);
    *out = g;
    return MORE3D_OK;
fail:
    dev_free(keys, s); dev_free(vals, s); dev_free(keys2, s); dev_free(vals2, s); dev_free(tmp, s);
    more3d_grid_destroy(g);
    return rc;
#undef G_TRY
#undef G_CUDA
}

int m3d::cloud_bbox(const double* xyz, int64_t n, double box_h[6], cudaStream_t s) {
    Segs sg;
    sg.p[0] = xyz; sg.p[1] = sg.p[2] = nullptr;
    sg.end[0] = sg.end[1] = sg.end[2] = n;
    const int nparts = 592;
    double* part = nullptr;
    M3D_TRY(dev_alloc(&part, (size_t)6 * nparts + 6, s));
    bbox_partial<<<nparts, BB_THREADS, 0, s>>>(sg, n, part);
    bbox_final<<<1, 32, 0, s>>>(part, nparts, part + 6 * nparts);
    count_launches(2);
    M3D_CHECK_LAUNCH();
    M3D_TRY(read_back(box_h, part + 6 * nparts, 6 * sizeof(double), s));
    dev_free(part, s);
    return MORE3D_OK;
}

int m3d::sort_queries_by_column(const more3d_grid* g, const double* q, int64_t nq, cudaStream_t s, uint32_t** order) {
    *order = nullptr;
    if (nq < 4096 || nq >= ((int64_t)1 << 32)) return MORE3D_OK;
    uint32_t *keys = nullptr, *vals = nullptr, *keys2 = nullptr, *vals2 = nullptr, *cnt = nullptr, *ks = nullptr, *vs = nullptr;
    const size_t ncols = (size_t)g->nx * g->ny;
    M3D_TRY(dev_alloc(&keys, (size_t)nq, s));
    M3D_TRY(dev_alloc(&vals, (size_t)nq, s));
    M3D_TRY(dev_alloc(&keys2, (size_t)nq, s));
    M3D_TRY(dev_alloc(&vals2, (size_t)nq, s));
    M3D_TRY(dev_alloc(&cnt, ncols + 1, s));      // cell_keys also counts; the counts are not used here
    M3D_CUDA(cudaMemsetAsync(cnt, 0, (ncols + 1) * sizeof(uint32_t), s));
    cell_keys<<<(unsigned)((nq + 255) / 256), 256, 0, s>>>(q, nq, make_lattice(g), keys, vals, cnt);
    count_launches(1);
    int bits = 1;
    while (((size_t)1 << bits) < ncols) ++bits;
    int rc = radix_sort_pairs_u32(keys, vals, keys2, vals2, nq, bits, &ks, &vs, s);
    if (rc == MORE3D_OK) {
        uint32_t* keep = vs;
        if (vs == vals) { dev_free(vals2, s); } else { dev_free(vals, s); }
        dev_free(keys, s); dev_free(keys2, s); dev_free(cnt, s);
        *order = keep;
        return MORE3D_OK;
    }
    dev_free(keys, s); dev_free(vals, s); dev_free(keys2, s); dev_free(vals2, s); dev_free(cnt, s);
    return rc;
}

extern "C" int more3d_grid_destroy(more3d_grid* g) {
    if (!g) return MORE3D_OK;
    if (g->order) cudaFreeAsync(g->order, g->stream);
    if (g->rec) cudaFreeAsync(g->rec, g->stream);
    if (g->col_start) cudaFreeAsync(g->col_start, g->stream);
    if (g->dup) cudaFreeAsync(g->dup, g->stream);
    if (g->rec32) cudaFreeAsync(g->rec32, g->stream);
    if (g->feat) cudaFreeAsync(g->feat, g->stream);
    delete g;
    return MORE3D_OK;
}

extern "C" int more3d_grid_info(const more3d_grid* g, int64_t* info_h, double* geom_h) {
    M3D_REQUIRE(g != nullptr, "grid is NULL");
    if (info_h) { info_h[0] = g->n; info_h[1] = g->nx; info_h[2] = g->ny; }
    if (geom_h) { geom_h[0] = g->cell; geom_h[1] = g->ox; geom_h[2] = g->oy; }
    return MORE3D_OK;
}

extern "C" int more3d_grid_order(const more3d_grid* g, int32_t* order, void* stream) {
    M3D_REQUIRE(g != nullptr && order != nullptr, "NULL argument");
    M3D_CUDA(cudaMemcpyAsync(order, g->order, (size_t)g->n * sizeof(uint32_t), cudaMemcpyDeviceToDevice,
                             (cudaStream_t)stream));
    return MORE3D_OK;
}

// dst[order[i]] = src[i] for elements of `words` 32-bit words (1: f32 / i32 columns, 2: f64, 6: f64 x 3 normals)
namespace m3d {
template <int W>
__global__ void __launch_bounds__(256) unpermute_kernel(const uint32_t* __restrict__ order,
                                                        const uint32_t* __restrict__ src,
                                                        uint32_t* __restrict__ dst, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const size_t o = (size_t)__ldg(order + i) * W;
    uint32_t v[W];
#pragma unroll
    for (int k = 0; k < W; ++k) v[k] = __ldg(src + (size_t)i * W + k);
#pragma unroll
    for (int k = 0; k < W; ++k) dst[o + k] = v[k];
}
}  // namespace m3d

extern "C" int more3d_unpermute(const int32_t* order, int64_t n, const void* src_gridorder, void* dst,
                                int elem_bytes, void* stream) {
    M3D_REQUIRE(order != nullptr && src_gridorder != nullptr && dst != nullptr, "NULL argument");
    M3D_REQUIRE(src_gridorder != dst, "in-place unpermute is not supported");
    M3D_REQUIRE(n >= 0, "n < 0");
    if (n == 0) return MORE3D_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned nb = (unsigned)((n + 255) / 256);
    const uint32_t* ord = (const uint32_t*)order;
    const uint32_t* src = (const uint32_t*)src_gridorder;
    uint32_t* d = (uint32_t*)dst;
    m3d::ProfScope prof("unpermute", s);
    switch (elem_bytes) {
        case 4: m3d::unpermute_kernel<1><<<nb, 256, 0, s>>>(ord, src, d, n); break;
        case 8: m3d::unpermute_kernel<2><<<nb, 256, 0, s>>>(ord, src, d, n); break;
        case 24: m3d::unpermute_kernel<6><<<nb, 256, 0, s>>>(ord, src, d, n); break;
        default: M3D_REQUIRE(false, "elem_bytes must be 4, 8 or 24");
    }
    m3d::count_launches(1);
    M3D_CHECK_LAUNCH();
    return MORE3D_OK;
}
