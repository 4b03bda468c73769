// Shared internals of libmore3d_b200 (sm_100a).  Not part of the C-ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

#include "../../include/more3d_b200.h"

namespace m3d {

// ------------------------------------------------------------------------------------------
// error plumbing: C-ABI functions return codes, never throw
// ------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);

#define M3D_CUDA(expr)                                                                     \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            m3d::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                   \
                           cudaGetErrorString(_e));                                        \
            return (_e == cudaErrorMemoryAllocation) ? MORE3D_E_NOMEM : MORE3D_E_CUDA;     \
        }                                                                                  \
    } while (0)

// every kernel launch of the library is counted (bench.py reports it as gpu_launches)
void count_launches(int n);
#define M3D_CHECK_LAUNCH() M3D_CUDA(cudaGetLastError())

#define M3D_TRY(expr)                  \
    do {                               \
        int _rc = (expr);              \
        if (_rc != MORE3D_OK) return _rc; \
    } while (0)

#define M3D_REQUIRE(cond, msg)                         \
    do {                                               \
        if (!(cond)) {                                 \
            m3d::set_error("invalid argument: %s", msg); \
            return MORE3D_E_INVALID;                   \
        }                                              \
    } while (0)

// Small device -> host read-back that does NOT go through a copy engine: a one-warp kernel stores the bytes into a
// pinned, device-mapped staging slot and the host waits for an event behind it.  A cudaMemcpyAsync of a few bytes is
// queued in the device-to-host copy engine behind whatever large download is in flight (the previous tile's saliency
// columns: 40-120 MB, 1-6 ms); the kernel's posted writes are not.  bytes <= 4096; returns after dst_h holds the data
// (everything queued earlier on `s` has then completed, like cudaStreamSynchronize would guarantee for that stream).
int read_back(void* dst_h, const void* src_dev, size_t bytes, cudaStream_t s);

// one-time per-device setup: keep stream-ordered pool memory cached instead of returning it to the
// driver at every synchronisation (otherwise each call pays cudaMalloc-class latencies)
void init_device_pool();

// Optional per-kernel timing with CUDA events on the launching stream (bench.py's roofline leg).
// Disabled by default: a scope is then a no-op.
struct ProfScope {
    int slot;
    cudaStream_t s;
    ProfScope(const char* name, cudaStream_t stream);
    ~ProfScope();
};

// stream-ordered scratch allocation
template <typename T>
static inline int dev_alloc(T** p, size_t count, cudaStream_t s) {
    *p = nullptr;
    if (count == 0) count = 1;
    M3D_CUDA(cudaMallocAsync((void**)p, count * sizeof(T), s));
    return MORE3D_OK;
}
static inline void dev_free(void* p, cudaStream_t s) {
    if (p) cudaFreeAsync(p, s);
}

// scratch buffers of one C-ABI call: freed (stream-ordered) when the holder goes out of scope, on every return path
struct Scratch {
    cudaStream_t s;
    void* p[40];
    int n;
    explicit Scratch(cudaStream_t stream) : s(stream), n(0) {}
    Scratch(const Scratch&) = delete;
    Scratch& operator=(const Scratch&) = delete;
    ~Scratch() { for (int i = 0; i < n; ++i) dev_free(p[i], s); }
    template <typename T>
    int alloc(T** out, size_t count) {
        const int rc = dev_alloc(out, count, s);
        if (rc == MORE3D_OK && n < 40) p[n++] = (void*)*out;
        return rc;
    }
};

// Margin between the grid's column width and the largest radius one stencil step covers: rounding
// in floor((x - ox) * inv_cell) is ~1e-13 relative, 2^-16 leaves orders of magnitude of slack.
#define M3D_CELL_MARGIN (1.0 + 1.0 / 65536.0)
// a column with more points than this switches the grid to z-sorted columns (see grid.cu)
#define M3D_TALL_COLUMN 128u

// sorted point record: 32 B, two LDG.128
struct __align__(32) PointRec {
    double x, y, z;
    long long idx;  // original index
};

// sorted feature record for the dk/dn pass: 64 B per point, STORED AS TWO ARRAYS OF 32-BYTE HALVES (first (x, y, z, kw)
// of all points, then (nx, ny, nz, pad) of all points; see ld_d4 / st_feat in features.cu) -- the struct only sizes it
struct __align__(32) FeatRec {
    double x, y, z;
    double K;        // the word `kw` (features.cu, feat_kword): [63..32] float32 bits of _nonparam_K, [31..2] the unit
                     // normal quantised to 3 x 10 bits (screening pass of dk/dn only), [1..0] flags: bit 0 = coincident
                     // with an earlier point (np.unique drop), bit 1 = raw normal was NaN.  Everything the screening
                     // pass needs sits in this FIRST 32-B half: one load per candidate.
    double nx, ny, nz;  // unit normal (n / ||n||), NaN if the point has no normal; read by the refinement of the few points
                        // whose dn the screening pass cannot settle, and by the one-pass kernels
    double pad;
};

}  // namespace m3d

struct more3d_grid {
    int64_t n;
    int nx, ny;
    double cell, inv_cell, ox, oy;   // ox, oy: position of local column (0, 0) -- origin of the fp32 records
    double cell_x, inv_x;            // column WIDTH along x (cell / x_split); `cell` is the row height along y.  Narrower
    int x_split;                     // columns cost nothing per stencil row (a row stays one contiguous range) and trim its
                                     // overhang beyond the query circle: (2r + cell_x) instead of (2r + cell) per row
    double lox, loy;                 // origin of the column LATTICE: column of x = floor((x - lox) / cell) - x0.  A single
    int x0, y0;                      // grid has lox = ox, x0 = 0; strips of one tile share the tile's lattice, so every
                                     // strip bins (and therefore orders and sums) exactly like the whole-tile grid
    int device;
    int tall;               // 1: some xy column holds more than M3D_TALL_COLUMN points -> columns are z-sorted
                            //    and the kernels bracket each column by z (vertical structures); 0: thin columns
    uint32_t max_col_pop;   // largest column population
    cudaStream_t stream;    // creation stream: grid arrays are stream-ordered allocations on it
    double oz;              // z origin of the fp32 records (bbox z-min)
    double lim32;           // delta32 holds for |coordinate - origin| < lim32
    double delta32;         // bound on |fp32 record coordinate - exact coordinate| (2^-24 * extent)
    float4* rec32;          // n sorted fp32 records (x-ox, y-oy, z-oz, 0): the fast distance filter
    uint32_t* order;        // n: sorted pos -> original index
    m3d::PointRec* rec;     // n sorted records
    uint32_t* col_start;    // nx*ny + 1
    uint8_t* dup;           // n (sorted order): 1 if coincident with an earlier sorted point
    m3d::FeatRec* feat;     // n sorted feature records cached by the last more3d_normals_curvature (or NULL)
    int feat_split;         // 1: stored as two arrays of 32-B halves (every thin grid, fine tall grids), 0: interleaved 64-B records
};

namespace m3d {

// sort.cu ------------------------------------------------------------------------------------
// Stable LSD radix sort of (key, val) pairs, keys use bits [0, key_bits).  Results end in
// keys_out/vals_out (which may not alias the inputs); inputs are clobbered (ping-pong).
int radix_sort_pairs_u32(uint32_t* keys, uint32_t* vals, uint32_t* keys_alt, uint32_t* vals_alt,
                         int64_t n, int key_bits, uint32_t** keys_sorted, uint32_t** vals_sorted,
                         cudaStream_t s);
int radix_sort_pairs_u64(uint64_t* keys, uint32_t* vals, uint64_t* keys_alt, uint32_t* vals_alt,
                         int64_t n, int key_bits, uint64_t** keys_sorted, uint32_t** vals_sorted,
                         cudaStream_t s);
// exclusive prefix sums; out[n] = total (out has n+1 entries).  in and out may alias for u32.
int exclusive_scan_u32(const uint32_t* in, uint32_t* out, int64_t n, cudaStream_t s);
int exclusive_scan_i32_to_i64(const int32_t* in, int64_t* out, int64_t n, cudaStream_t s);

// grid.cu: processing order for EXTERNAL query points -- their indices sorted by grid column, so that the 32
// lanes of a warp walk the same stencil rows (the locality the self-query path gets from the sorted cloud).
// *order is a stream-ordered allocation of nq uint32 (free with dev_free); NULL for small query sets.
// bounding box {min x, y, z, max x, y, z} of a cloud, read back to the host (one small synchronisation)
int cloud_bbox(const double* xyz, int64_t n, double box_h[6], cudaStream_t s);
int sort_queries_by_column(const more3d_grid* g, const double* q, int64_t nq, cudaStream_t s, uint32_t** order);

// features.cu: kernel variant of the dk/dn pass (MORE3D_DKDN_VARIANT, read once): 0 = two passes (screen + refine,
// default), 1 = shared-memory staged one-pass walk, 2 = global-memory one-pass walk
int dkdn_variant();

// device helpers -------------------------------------------------------------------------------
__device__ __forceinline__ double dist2_rule(double dx, double dy, double dz) {
    // ((dx*dx)+(dy*dy))+(dz*dz), round-to-nearest each step, never contracted to FMA:
    // sklearn metrics/_dist_metrics.pxd.tp:44-49
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

__device__ __forceinline__ int cell_coord(double v, double o, double inv, int shift) {
    double f = floor((v - o) * inv);
    // clamp far-away queries before the int conversion (ix +- stencil must not overflow int32)
    f = fmin(fmax(f, -1.0e9), 1.0e9);
    return (int)f - shift;
}

// order-preserving double -> uint64 mapping (radix-sortable); -0.0 and +0.0 map to the same key
__device__ __forceinline__ uint64_t sortable_key(double v) {
    v = v + 0.0;
    const uint64_t u = (uint64_t)__double_as_longlong(v);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}

// numpy's float64 add.reduce over a contiguous axis of n <= 128 elements (pairwise_sum in
// numpy/_core/src/umath/loops_utils.h.src): plain left-to-right for n < 8, otherwise eight running sums over
// blocks of 8, combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), then the tail left to right.  Reproduced
// bit for bit because `smooth` feeds the |phi| > max_val test of the next iteration: a weighted mean of equal
// clamped values lands within 1 ulp of +-max_val, and which side decides whether the point is re-smoothed.
__device__ __forceinline__ double np_pairwise_sum(const double* a, int n) {
    if (n < 8) {
        double r = 0.0;
        for (int i = 0; i < n; ++i) r = __dadd_rn(r, a[i]);
        return r;
    }
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    int i = 8;
    for (; i < n - (n % 8); i += 8)
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], a[i + j]);
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                           __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __dadd_rn(res, a[i]);
    return res;
}

// order-preserving float <-> uint mapping for atomic min/max
__device__ __forceinline__ unsigned f32_ordered(float f) {
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float f32_unordered(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

}  // namespace m3d
