// K4/K5/K6: fused neighbourhood moments -> PCA normal / surface variation / non-parametric
// curvature, and the dk/dn centre-surround pass.
// Replaces Curvature_Kernel.process (DM_saliency.py:88-157), Saliency_Kernel.process (:219-324) and the
// normals precondition (DM_saliency.py:329-330).
#include <stdlib.h>
#include "gridview.cuh"
#include "eig3.cuh"

namespace m3d {

// ------------------------------------------------------------------------------------------
// K4 + K5
// ------------------------------------------------------------------------------------------
enum { MODE_NORMALS = 0, MODE_CURV = 1, MODE_FUSED = 2 };

struct CurvParams {
    double r, r2;
    double eps;
    int min_points;
    int s;
    int grid_order;     // 1: per-point outputs are written at the point's GRID position (coalesced) instead of
                        //    its original index (a scatter of six partial sectors per point)
    Filter32 flt;
};

constexpr int WALK_THREADS = 256;

struct D4 { double x, y, z, w; };
struct D4x2 { D4 a, b; };
// The feature records live as TWO arrays of 32-byte halves -- half 0 = (x, y, z, kw) of all n points, then half 1
// = (nx, ny, nz, pad) of all n points -- not as n 64-byte structs.  The screening pass of the two-pass dk/dn reads half 0
// only, so every thin grid uses this layout.  For the one-pass kernel (tall grids, MORE3D_DKDN_VARIANT=2), which reads
// both halves of every candidate, the layout follows the grid's mean column population (feat_split, decided when the
// grid is built): at every candidate step the 32 lanes of a warp read ~13 distinct records out of a window of ~32
// consecutive ones, and with 32-byte elements that window is 8 cache lines per load instead of 16 (14.83 -> 14.41 ms
// over the three radii of configs[1] on fine grids); on a coarse grid (>= 4 points per column: the lanes sit in ~3
// columns and both halves of their few records share cache lines) the interleaved 64-byte layout is 8-12 % faster.
// n == 0 selects the interleaved layout.
template <bool SPLIT>
__device__ __forceinline__ D4 ld_d4(const FeatRec* __restrict__ f, int64_t n, int64_t j, int h) {
    D4 r;      // the layout is a template parameter: one more select + multiply per load costs 7 % in this issue-bound loop
    const size_t slot = SPLIT ? (size_t)h * (size_t)n + (size_t)j : 2 * (size_t)j + (size_t)h;
    ldg256(reinterpret_cast<const char*>(f) + slot * 32, r.x, r.y, r.z, r.w);
    return r;
}
// 4th word of a record's first half: [63..32] the float32 bits of _nonparam_K, [31..2] the unit normal quantised to
// 3 x 10 bits (q = rint(511 c) + 512; a NaN normal is stored as the zero vector), [1..0] the flags (bit 0 = coincident
// with an earlier point, bit 1 = raw normal was NaN).  The quantised normal only serves the SCREENING pass of dk/dn
// (dkdn_screen_kernel), which needs a certified bound on n_p . n_q, never its value.
__device__ __forceinline__ uint32_t q10(double c) {
    double v = rint(c * 511.0);
    v = fmin(fmax(v, -511.0), 511.0);          // NaN -> -511 here; the caller replaces NaN normals as a whole
    return (uint32_t)((int)v + 512);
}
__device__ __forceinline__ double feat_kword(float K, int flags, double ux, double uy, double uz) {
    uint32_t lo = (uint32_t)flags;
    if (isnan(ux) || isnan(uy) || isnan(uz)) lo |= (512u << 2) | (512u << 12) | (512u << 22);
    else lo |= (q10(ux) << 2) | (q10(uy) << 12) | (q10(uz) << 22);
    return __longlong_as_double((long long)(((unsigned long long)__float_as_uint(K) << 32) | lo));
}
__device__ __forceinline__ double kword_K(long long kb) { return (double)__uint_as_float((uint32_t)((unsigned long long)kb >> 32)); }

// (nx, ny, nz) is the RAW normal: the record holds n / ||n||, DM_saliency.py:258-259 (norm = sqrt of the sequential
// sum of squares)
__device__ __forceinline__ void st_feat(FeatRec* __restrict__ f, int64_t n, int64_t i, double x, double y, double z,
                                        float K, bool dup, double nx, double ny, double nz) {
    const double nn = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(nx, nx), __dmul_rn(ny, ny)), __dmul_rn(nz, nz)));
    const double ux = nx / nn, uy = ny / nn, uz = nz / nn;
    const double kw = feat_kword(K, (dup ? 1 : 0) | (isnan(nx) ? 2 : 0), ux, uy, uz);
    const size_t sa = n ? (size_t)i : 2 * (size_t)i, sb = n ? (size_t)n + (size_t)i : 2 * (size_t)i + 1;
    *reinterpret_cast<double4*>(reinterpret_cast<char*>(f) + sa * 32) = make_double4(x, y, z, kw);
    *reinterpret_cast<double4*>(reinterpret_cast<char*>(f) + sb * 32) = make_double4(ux, uy, uz, 0.0);
}


// One thread per sorted point.  Accumulates m, sum(d), sum(d d^T) with d = q - p (fp64, shifted to the
// query so there is no cancellation against the ~1e3 m coordinates), then finishes in registers.
// neighbourhood sums of one query point: count, sum(d), sum(d d^T), d = q - p
struct Moments {
    int m;
    double sx, sy, sz, sxx, sxy, sxz, syy, syz, szz;
    __device__ __forceinline__ void clear() { m = 0; sx = sy = sz = sxx = sxy = sxz = syy = syz = szz = 0.0; }
    template <int MODE>
    __device__ __forceinline__ void add(const PointRec& c, const PointRec& p, double r2) {
        const double dx = c.x - p.x, dy = c.y - p.y, dz = c.z - p.z;
        if (dist2_rule(dx, dy, dz) <= r2) {
            ++m;
            sx += dx; sy += dy; sz += dz;
            if (MODE != MODE_CURV) {
                sxx += dx * dx; sxy += dx * dy; sxz += dx * dz;
                syy += dy * dy; syz += dy * dz; szz += dz * dz;
            }
        }
    }
};

// everything after the walk for sorted point i: PCA normal / surface variation, Curvature_Kernel.process, outputs
template <int MODE>
__device__ __forceinline__ void moments_finish(const GridView& g, const CurvParams& prm, int64_t i, const PointRec& p,
                                               const Moments& s, double* __restrict__ normals,
                                               double* __restrict__ lam_ratio, float* __restrict__ Kout,
                                               int32_t* __restrict__ pcount, FeatRec* __restrict__ feat_out) {
    const int64_t o = prm.grid_order ? i : p.idx;
    const int m = s.m;
    const double sx = s.sx, sy = s.sy, sz = s.sz;
    double nx, ny, nz;
    if (MODE == MODE_CURV) {
        nx = normals[3 * o]; ny = normals[3 * o + 1]; nz = normals[3 * o + 2];
    } else {
        double ratio = 0.0;
        if (m >= 3) {
            const double im = 1.0 / (double)m;
            const double cx = sx * im, cy = sy * im, cz = sz * im;
            const double a00 = s.sxx * im - cx * cx, a01 = s.sxy * im - cx * cy, a02 = s.sxz * im - cx * cz;
            const double a11 = s.syy * im - cy * cy, a12 = s.syz * im - cy * cz, a22 = s.szz * im - cz * cz;
            double lmin, v[3];
            sym3_min_eig(a00, a01, a02, a11, a12, a22, lmin, v);
            nx = v[0]; ny = v[1]; nz = v[2];
            const double tr = a00 + a11 + a22;
            ratio = (tr > 0.0) ? fmax(lmin, 0.0) / tr : 0.0;
        } else {
            nx = ny = nz = __longlong_as_double(0x7ff8000000000000LL);
        }
        if (lam_ratio) lam_ratio[o] = ratio;
        if (MODE == MODE_NORMALS) {
            normals[3 * o] = nx; normals[3 * o + 1] = ny; normals[3 * o + 2] = nz;
            if (pcount) pcount[o] = m;
            return;
        }
    }

    // ---- Curvature_Kernel.process ----
    pcount[o] = m;                                                    // DM_saliency.py:102
    float K = 0.0f;
    bool flipped = false;
    if (!isnan(nx) && m >= prm.min_points) {                          // :107-110, :120-123
        // mean tangential offset, projector built from the un-flipped normal (sign invariant) :117,139-141
        const double im = 1.0 / (double)m;
        const double cx = sx * im, cy = sy * im, cz = sz * im;
        const double nc = nx * cx + ny * cy + nz * cz;
        const double tx = cx - nx * nc, ty = cy - ny * nc, tz = cz - nz * nc;
        if (nx * (-p.x) + ny * (-p.y) + nz * (-p.z) < 0.0) {          // :131-135
            nx = -nx; ny = -ny; nz = -nz; flipped = true;
        }
        if (!(sqrt(tx * tx + ty * ty + tz * tz) > 1.0)) {             // :142-144
            // sum_i n.(p - q_i) / (m - 1) = -n.sum(d) / (m - 1)        :137-146
            const double sum_proj = -(nx * sx + ny * sy + nz * sz) / (double)(m - 1);
            if (!(fabs(sum_proj) < prm.eps)) K = (float)sum_proj;     // :149-153
        }
    }
    Kout[o] = K;
    if (MODE == MODE_FUSED || flipped) {
        normals[3 * o] = nx; normals[3 * o + 1] = ny; normals[3 * o + 2] = nz;
    }
    if (MODE == MODE_FUSED && feat_out) {
        // the dk/dn pass's sorted feature record, written here (coalesced) instead of being gathered
        // from the original-order columns by build_feat_kernel; same arithmetic as there
        st_feat(feat_out, g.feat_n, i, p.x, p.y, p.z, K, g.dup[i] != 0, nx, ny, nz);
    }
}

template <int MODE, bool TALL>
__global__ void __launch_bounds__(WALK_THREADS, 4) moments_kernel(GridView g, CurvParams prm,
                                                               double* __restrict__ normals,
                                                               double* __restrict__ lam_ratio,
                                                               float* __restrict__ Kout,
                                                               int32_t* __restrict__ pcount,
                                                               FeatRec* __restrict__ feat_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.n) return;
    const PointRec p = load_rec(g.rec, i);
    Moments s;
    s.clear();
    // The fp64 differences serve both the exact accept rule and the moments, so an fp32 pre-filter
    // (used by the count/fill kernels) does not pay here: measured 2x slower (conditional fp64 loads
    // lose the unrolled loop's memory-level parallelism).
    walk_candidates_pipelined<TALL>(g, p.x, p.y, p.z, prm.r, prm.s, [&](int64_t j) {
        return load_rec(g.rec, j);
    }, [&](const PointRec& c) { s.add<MODE>(c, p, prm.r2); });
    moments_finish<MODE>(g, prm, i, p, s, normals, lam_ratio, Kout, pcount, feat_out);
}

// ------------------------------------------------------------------------------------------
// K6: dk/dn
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) build_feat_kernel(GridView g, const double* __restrict__ normals,
                                                         const float* __restrict__ K,
                                                         FeatRec* __restrict__ feat) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.n) return;
    const PointRec p = load_rec(g.rec, i);
    const int64_t o = p.idx;
    const double nx = normals[3 * o], ny = normals[3 * o + 1], nz = normals[3 * o + 2];
    st_feat(feat, g.feat_n, i, p.x, p.y, p.z, K[o], g.dup[i] != 0, nx, ny, nz);
}

struct DkDnParams {
    double r, r2, rho2, inv_rho2, sigma_n, sigma_k;
    double zk100_thr;   // largest x with fl(100 * x) <= 0.04 (rounding is monotone, so |100*dk| <= 0.04 <=> |dk| <= x)
    int min_points, s, table_len;
    int grid_order;     // 1: dk / dn are written at the grid position
    long long own_lo, own_hi;  // only points with original index in [own_lo, own_hi) enter the fused min/max (a strip's
                               // own points sit between its left and right ghosts)
    Filter32 flt;
    // screening pass (dkdn_screen_kernel): a candidate is ACTIVE (w > 0.01, :272-279) iff it is unique, inside and
    // act_lo <= d2 <= act_hi, d2 != rho2 -- the weight's rounding is monotone on either side of rho2, so the two
    // thresholds (found by the host with the kernel's own arithmetic) reproduce the comparison bit for bit
    double act_lo, act_hi;
    float dot_thr;      // quantised n_p . n_q above this  =>  certainly |1 - n_p . n_q| < 0.016 (:286)
};

// largest double x >= 0 with fl(x * 100.0) <= 0.04: the rescue test of DM_saliency.py:290 without the multiply
static double zk100_threshold() {
    double x = 0.0004;
    while (x * 100.0 <= 0.04) x = nextafter(x, 1.0);
    while (x * 100.0 > 0.04) x = nextafter(x, 0.0);
    return x;
}

// act_lo = smallest d2 with fl(inv_rho2 * d2) > 0.01, act_hi = largest d2 with fl(2 - fl(inv_rho2 * d2)) > 0.01 (the two
// branches of the ring weight, DkdnAcc::add).  false if rho is degenerate (the screening pass is then not used).
static bool act_interval(double rho2, double inv_rho2, double* lo, double* hi) {
    if (!(rho2 > 0.0) || !isfinite(rho2) || !isfinite(inv_rho2) || !(inv_rho2 > 0.0)) return false;
    // volatile: every product and difference is rounded to double exactly as __dmul_rn / __dsub_rn do (no contraction)
    auto w_lo = [&](double x) { volatile double t = inv_rho2 * x; return (double)t; };
    auto w_hi = [&](double x) { volatile double t = inv_rho2 * x; volatile double u = 2.0 - t; return (double)u; };
    double x = 0.01 * rho2;
    int up = 0, down = 0;
    while (!(w_lo(x) > 0.01) && ++up < 1000) x = nextafter(x, INFINITY);
    while (x > 0.0 && w_lo(nextafter(x, 0.0)) > 0.01 && ++down < 1000) x = nextafter(x, 0.0);
    if (up >= 1000 || down >= 1000 || !(x < rho2) || !(w_lo(x) > 0.01)) return false;
    *lo = x;
    x = 1.99 * rho2;
    up = down = 0;
    while (!(w_hi(x) > 0.01) && ++down < 1000) x = nextafter(x, 0.0);
    while (w_hi(nextafter(x, INFINITY)) > 0.01 && ++up < 1000) x = nextafter(x, INFINITY);
    if (up >= 1000 || down >= 1000 || !(x > rho2) || !(w_hi(x) > 0.01)) return false;
    *hi = x;
    return true;
}

// One query point's state of Saliency_Kernel.process: init() from its own record, add() per candidate, finish().
// Shared by the global-memory walk (dkdn_kernel) and the shared-memory staged walk (dkdn_smem_kernel): the two produce
// bit-identical results because they execute the same expression on the same candidates in the same order.
struct DkdnAcc {
    double px, py, pz, Kp, pnx, pny, pnz;
    long long pflags;
    int m, mu, nact, zn, zk, zk100;
    double s1n, s2n, s1k, s2k, wn, wk, wsum;

    __device__ __forceinline__ void init(const D4& pa, const D4& pb) {   // (x y z K) (nx ny nz flags)
        px = pa.x; py = pa.y; pz = pa.z;
        pflags = __double_as_longlong(pa.w) & 3LL;
        Kp = kword_K(__double_as_longlong(pa.w));
        pnx = pb.x; pny = pb.y; pnz = pb.z;
        m = mu = nact = zn = zk = zk100 = 0;
        s1n = s2n = s1k = s2k = wn = wk = wsum = 0.0;
    }

    // Branch-free candidate body: every lane evaluates the whole expression and the accept / np.unique
    // decisions enter as selects.  With 32 lanes at independent positions of their rows some lane accepts at
    // practically every step, so a branch saves no issue slots -- it only serialises two dependent loads per
    // candidate (ncu: 45 % of the stall samples sat on the first use of each load).  Without control flow both
    // 32-B halves of the unrolled candidates are requested up front.
    __device__ __forceinline__ void add(const D4& qa, const D4& qb, const DkDnParams& prm) {
        const double d2 = dist2_rule(px - qa.x, py - qa.y, pz - qa.z);   // :256-257
        const long long kb = __double_as_longlong(qa.w);
        const bool in = d2 <= prm.r2;
        const bool u = in & !((int)kb & 1);                               // np.unique, :247
        double dn = 1.0 - (pnx * qb.x + pny * qb.y + pnz * qb.z);         // :261
        double dk = Kp - kword_K(kb);                                     // :262
        dn = u ? dn : 0.0;
        dk = u ? dk : 0.0;
        // opaque to the optimiser: keeps |.| as an operand modifier of the consumers instead of a
        // second, separately masked copy of each value
        asm("" : "+d"(dn));
        asm("" : "+d"(dk));
        // ring weight on the squared distance (:272-275): d/rho2 below rho2, 2 - d/rho2 above,
        // 0 at equality, clamped at 0.  (-1/rho2)*d + 2 == 2 - (1/rho2)*d bit for bit.
        const double t = __dmul_rn(prm.inv_rho2, d2);
        double w = (d2 < prm.rho2) ? t : __dsub_rn(2.0, t);
        const bool keep = u & (d2 != prm.rho2) & !(w < 0.0);
        w = keep ? w : 0.0;
        s1n += dn; s2n += dn * dn;
        s1k += dk; s2k += dk * dk;
        // counters as predicated adds (one instruction each): m, mu, then with act = w > 0.01 (:279)
        // nact, zn (|dn| < 0.016, :286), zk (|dk| <= 0.04, :287), zk100 (|100 dk| <= 0.04, :290)
        asm("{ .reg .pred pin, pu, pa, q;\n"
            "  setp.ne.u32 pin, %6, 0;\n"
            "  setp.ne.u32 pu, %7, 0;\n"
            "  @pin add.s32 %0, %0, 1;\n"
            "  @pu add.s32 %1, %1, 1;\n"
            "  setp.gt.f64 pa, %8, %11;\n"
            "  @pa add.s32 %2, %2, 1;\n"
            "  .reg .f64 an, ak;\n"
            "  abs.f64 an, %9;\n"
            "  abs.f64 ak, %10;\n"
            "  setp.lt.and.f64 q, an, %12, pa;\n"
            "  @q add.s32 %3, %3, 1;\n"
            "  setp.le.and.f64 q, ak, %13, pa;\n"
            "  @q add.s32 %4, %4, 1;\n"
            "  setp.le.and.f64 q, ak, %14, pa;\n"
            "  @q add.s32 %5, %5, 1;\n"
            "}"
            : "+r"(m), "+r"(mu), "+r"(nact), "+r"(zn), "+r"(zk), "+r"(zk100)
            : "r"((unsigned)in), "r"((unsigned)u), "d"(w), "d"(dn), "d"(dk), "d"(0.01), "d"(0.016), "d"(0.04),
              "d"(prm.zk100_thr));
        wn += fabs(dn) * w; wk += dk * w; wsum += w;                      // :300, :316
    }

    __device__ __forceinline__ void finish(const DkDnParams& prm, const double* __restrict__ chi2_table, int* err,
                                           float& dkf, float& dnf) {
        dkf = 0.0f; dnf = 0.0f;
        if (!(pflags & 2LL) && m >= prm.min_points) {                         // :225-229, :240-244
            if (nact >= prm.table_len) {
                atomicMax(err, nact + 1);       // the table length that would have sufficed
            } else {
                const double chi_t = chi2_table[nact];                        // :282
                const double imu = 1.0 / (double)mu;
                const double mn = s1n * imu, mk = s1k * imu;
                const double std_n = sqrt(fmax(s2n * imu - mn * mn, 0.0));
                const double std_k = sqrt(fmax(s2k * imu - mk * mk, 0.0));
                // NaN propagates like numpy: fmax(NaN, 0) would hide it, so re-inject
                const double sdn = isnan(s1n) ? s1n : std_n;
                const double sdk = isnan(s1k) ? s1k : std_k;
                const double chi_n = (double)(m - 1) * sdn / prm.sigma_n;     // :283
                const double chi_k = (double)(m - 1) * sdk / prm.sigma_k;     // :284
                if (zk == nact) zk = zk100;                                   // :289-290
                const double lim = 0.6 * (double)nact;
                double dn;
                if (chi_n < chi_t) dn = 0.0;                                  // :292-300
                else if ((double)zn > lim) dn = 0.0;
                else dn = wn / wsum;
                if (isnan(dn)) dn = 0.0;                                      // :303-308
                else if (dn > 2.0) dn = 1.0;
                double dk;
                if (chi_k < chi_t) dk = 0.0;                                  // :309-316
                else if ((double)zk > lim) dk = 0.0;
                else dk = fabs(wk / wsum);
                dkf = (float)dk; dnf = (float)dn;
            }
        }
    }
};

// fused Histo min/max (DM_saliency.py:562-568): warp shuffle, then one atomic per warp.
// NaN (a neighbourhood whose active weights sum to 0 gives dk = |0/0|, DM_saliency.py:316) stays in the column
// but is kept out of the extrema -- declared choice: one NaN must not poison every normalised value
__device__ __forceinline__ void dkdn_minmax(unsigned lo_k, unsigned hi_k, unsigned lo_n, unsigned hi_n,
                                            unsigned* __restrict__ mm) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo_k = min(lo_k, __shfl_xor_sync(0xffffffffu, lo_k, o));
        hi_k = max(hi_k, __shfl_xor_sync(0xffffffffu, hi_k, o));
        lo_n = min(lo_n, __shfl_xor_sync(0xffffffffu, lo_n, o));
        hi_n = max(hi_n, __shfl_xor_sync(0xffffffffu, hi_n, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(mm + 0, lo_k); atomicMax(mm + 1, hi_k);
        atomicMin(mm + 2, lo_n); atomicMax(mm + 3, hi_n);
    }
}

// Two-pass dk/dn: the refinement costs about as much per listed point as the one-pass kernel costs per point (a whole-
// expression entry about three times that), so when the screening pass lists more than ~45 % of the cloud (very rough
// surfaces, tiny sigmas) the one-pass kernel is run over everything instead.  The decision is taken on the device from
// the list counters: both alternatives are launched, one of them returns at once.
__device__ __forceinline__ bool refine_gate_open(const unsigned* __restrict__ counts, unsigned above) {
    return (unsigned long long)__ldg(counts) + 3ull * (unsigned long long)__ldg(counts + 1) > (unsigned long long)above;
}

// one sorted point of the one-pass dk/dn: walk, finish, write, fused extrema (all 32 lanes of a warp must call it)
template <bool TALL, bool SPLIT>
__device__ __forceinline__ void dkdn_point(const GridView& g, const FeatRec* __restrict__ feat, const DkDnParams& prm,
                                           const double* __restrict__ chi2_table, float* __restrict__ dk_out,
                                           float* __restrict__ dn_out, unsigned* __restrict__ mm, int* __restrict__ err,
                                           int64_t i) {
    float dkf = 0.0f, dnf = 0.0f;
    const bool live = i < g.n;
    bool owned = false;
    if (live) {
        DkdnAcc acc;
        acc.init(ld_d4<SPLIT>(feat, g.n, i, 0), ld_d4<SPLIT>(feat, g.n, i, 1));
        walk_candidates_pipelined<TALL>(g, acc.px, acc.py, acc.pz, prm.r, prm.s, [&](int64_t j) {
            D4x2 c;
            c.a = ld_d4<SPLIT>(feat, g.n, j, 0);
            c.b = ld_d4<SPLIT>(feat, g.n, j, 1);
            return c;
        }, [&](const D4x2& c) { acc.add(c.a, c.b, prm); });
        acc.finish(prm, chi2_table, err, dkf, dnf);
        const bool all_owned = prm.own_lo <= 0 && prm.own_hi >= g.n;
        const int64_t orig = (prm.grid_order && all_owned)
                                 ? 0 : __double_as_longlong(__ldg(reinterpret_cast<const double*>(g.rec + i) + 3));
        const int64_t o = prm.grid_order ? i : orig;
        dk_out[o] = dkf;
        dn_out[o] = dnf;
        owned = all_owned || (orig >= prm.own_lo && orig < prm.own_hi);
    }
    if (mm) {
        const bool ck = live && owned && !isnan(dkf), cn = live && owned && !isnan(dnf);
        dkdn_minmax(ck ? f32_ordered(dkf) : 0xffffffffu, ck ? f32_ordered(dkf) : 0u,
                    cn ? f32_ordered(dnf) : 0xffffffffu, cn ? f32_ordered(dnf) : 0u, mm);
    }
}

template <bool TALL, bool SPLIT>
__global__ void __launch_bounds__(WALK_THREADS) dkdn_kernel(GridView g, const FeatRec* __restrict__ feat,
                                                   DkDnParams prm, const double* __restrict__ chi2_table,
                                                   float* __restrict__ dk_out, float* __restrict__ dn_out,
                                                   unsigned* __restrict__ mm /* ordered min/max x4 */,
                                                   int* __restrict__ err) {
    dkdn_point<TALL, SPLIT>(g, feat, prm, chi2_table, dk_out, dn_out, mm, err,
                            (int64_t)blockIdx.x * blockDim.x + threadIdx.x);
}

// The fallback of the two-pass path: the one-pass expression over the whole cloud, as a small persistent grid that
// returns at once unless the screening pass listed more than `gate_above` weighted entries (refine_gate_open).
__global__ void __launch_bounds__(WALK_THREADS) dkdn_fallback_kernel(GridView g, const FeatRec* __restrict__ feat,
                                                            DkDnParams prm, const double* __restrict__ chi2_table,
                                                            float* __restrict__ dk_out, float* __restrict__ dn_out,
                                                            unsigned* __restrict__ mm, int* __restrict__ err,
                                                            const unsigned* __restrict__ gate, unsigned gate_above) {
    if (!refine_gate_open(gate, gate_above)) return;
    const int64_t tiles = (g.n + WALK_THREADS - 1) / WALK_THREADS;
    for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x)
        dkdn_point<false, true>(g, feat, prm, chi2_table, dk_out, dn_out, mm, err, t * WALK_THREADS + threadIdx.x);
}

// ------------------------------------------------------------------------------------------
// K6 in two passes (thin grids, split records; the default).
//
// Saliency_Kernel.process returns dk = 0 whenever more than 60 % of the ACTIVE neighbours have |dk| <= 0.04 and dn = 0
// whenever more than 60 % have |dn| < 0.016 (DM_saliency.py:286-316) -- whatever the chi2 tests say.  On open terrain
// that settles dn for practically every point and dk for 80-95 % of them, and both conditions are pure COUNTS:
//
//   pass 1, dkdn_screen_kernel (every point; thread per point, the same row walk as dkdn_kernel): per candidate only
//     the exact fp64 accept rule, the active test as two thresholds on d2 (act_interval), |dk| <= T as a float32
//     interval test on the neighbour's K (dk_interval: exact, K is a float32 column) and a CERTIFIED lower count of
//     |dn| < 0.016 from the 3 x 10-bit normal packed beside K -- 12 FP64 instructions and ONE 32-byte load per
//     candidate instead of 35 and two.  A point whose counts settle both outputs writes (0, 0); the others are
//     appended to a work list (bit 31 of an entry: the normal test did not settle dn).
//   pass 2, dkdn_refine_kernel (listed points; one WARP per point): the full expression of DkdnAcc on the point's
//     candidates, lanes striding over the concatenated stencil rows (coalesced 32-byte loads), partial sums combined
//     by a fixed shuffle tree.  Counters and decisions are the same integers as in the one-pass kernel; the fp64 sums
//     are added in another (still deterministic, lattice-defined) order, i.e. values agree to rounding.
// ------------------------------------------------------------------------------------------

// { x float32 : |fl64(Kp - x)| <= T } as a closed float32 interval [lo, hi] (the fp64 difference is monotone in x).
// ok = false if the few-ulp search did not pin both ends (the point is then refined unconditionally).
__device__ __forceinline__ bool dk_pred(double Kp, float x, double T) { return fabs(__dsub_rn(Kp, (double)x)) <= T; }
// next float32 above / below a finite x (bit arithmetic; nextafterf costs a branchy library call per use)
__device__ __forceinline__ float f32_up(float x) {
    const int b = __float_as_int(x);
    return (x == 0.0f) ? __int_as_float(1) : __int_as_float(b >= 0 ? b + 1 : b - 1);
}
__device__ __forceinline__ float f32_down(float x) {
    const int b = __float_as_int(x);
    return (x == 0.0f) ? __int_as_float((int)0x80000001u) : __int_as_float(b >= 0 ? b - 1 : b + 1);
}
__device__ __forceinline__ bool dk_interval(double Kp, double T, float& lo, float& hi) {
    // the ends lie within one float32 of Kp -+ T rounded outwards: one corrective step each, then the defining
    // property of both ends is verified (a failed verification only sends the point to the refinement pass)
    float l = __double2float_rd(Kp - T), h = __double2float_ru(Kp + T);
    if (!dk_pred(Kp, l, T)) l = f32_up(l);
    if (!dk_pred(Kp, h, T)) h = f32_down(h);
    lo = l; hi = h;
    return dk_pred(Kp, l, T) && !dk_pred(Kp, f32_down(l), T) && dk_pred(Kp, h, T) && !dk_pred(Kp, f32_up(h), T) &&
           l <= h && isfinite(l) && isfinite(h);
}

struct ScreenAcc {
    double px, py, pz;
    float ax, ay, az, c0;                    // t = 1 - quantised n_p . n_q = c0 - sum g_c a_c (see set_normal / add)
    float k_lo, k_hi, k100_lo, k100_hi;
    int m, mu, nact, zk, zk100;
    float s1, s2;                            // float32 sums of t = max(1 - quantised n_p . n_q, 0) over the unique neighbours

    // The 10-bit fields of a neighbour's packed normal are read IN PLACE as mantissa bits of a float in [1, 2):
    // g = 1 + q 2^-k with k = 21, 11 (fields at bits 2-11 and 12-21: one LOP3 each) and 10 (bits 22-31, one shift
    // more), so the quantised component (q - 512) / 511 = g 2^k / 511 - (2^k + 512) / 511 needs no integer conversion:
    //   1 - n_p . n_q' = [1 + sum_c n_p,c (2^k_c + 512) / 511] - sum_c g_c (n_p,c 2^k_c / 511) = c0 - sum_c g_c a_c.
    // float32 error: c0 ~ 4105 n_p,x and a_x g_x carry 2.5e-4 each (the x scale 2^21 / 511), everything else < 1e-5.
    __device__ __forceinline__ void set_normal(double nx, double ny, double nz) {
        const double kx = 2097152.0, ky = 2048.0, kz = 1024.0;
        ax = (float)(nx * kx / 511.0); ay = (float)(ny * ky / 511.0); az = (float)(nz * kz / 511.0);
        c0 = (float)(1.0 + (nx * (kx + 512.0) + ny * (ky + 512.0) + nz * (kz + 512.0)) / 511.0);
    }

    __device__ __forceinline__ void add(const D4& qa, const DkDnParams& prm) {
        const double d2 = dist2_rule(px - qa.x, py - qa.y, pz - qa.z);
        const unsigned long long kb = (unsigned long long)__double_as_longlong(qa.w);
        const uint32_t lo = (uint32_t)kb;
        const float Kq = __uint_as_float((uint32_t)(kb >> 32));
        uint32_t bx, by, bz;
        const uint32_t one = 0x3F800000u;
        asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(bx) : "r"(lo), "r"(0x00000FFCu), "r"(one));
        asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(by) : "r"(lo), "r"(0x003FF000u), "r"(one));
        asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(bz) : "r"(lo >> 9), "r"(0x007FE000u), "r"(one));
        const float t = fmaxf(fmaf(-__uint_as_float(bz), az, fmaf(-__uint_as_float(by), ay, fmaf(-__uint_as_float(bx), ax, c0))), 0.0f);
        // counters as predicated adds: m (inside), mu and the float32 sums (unique & inside), then with
        // act = unique & inside & act_lo <= d2 <= act_hi & d2 != rho2: nact, zk (|dk| <= 0.04), zk100 (|100 dk| <= 0.04)
        asm("{ .reg .pred pin, pu, pa, q;\n"
            "  setp.le.f64 pin, %7, %8;\n"
            "  @pin add.s32 %0, %0, 1;\n"
            "  setp.eq.and.u32 pu, %12, 0, pin;\n"
            "  @pu add.s32 %1, %1, 1;\n"
            "  @pu add.f32 %5, %5, %18;\n"
            "  @pu fma.rn.f32 %6, %18, %18, %6;\n"
            "  setp.ge.and.f64 pa, %7, %9, pu;\n"
            "  setp.le.and.f64 pa, %7, %10, pa;\n"
            "  setp.ne.and.f64 pa, %7, %11, pa;\n"
            "  @pa add.s32 %2, %2, 1;\n"
            "  setp.ge.and.f32 q, %13, %14, pa;\n"
            "  setp.le.and.f32 q, %13, %15, q;\n"
            "  @q add.s32 %3, %3, 1;\n"
            "  setp.ge.and.f32 q, %13, %16, pa;\n"
            "  setp.le.and.f32 q, %13, %17, q;\n"
            "  @q add.s32 %4, %4, 1;\n"
            "}"
            : "+r"(m), "+r"(mu), "+r"(nact), "+r"(zk), "+r"(zk100), "+f"(s1), "+f"(s2)
            : "d"(d2), "d"(prm.r2), "d"(prm.act_lo), "d"(prm.act_hi), "d"(prm.rho2), "r"(lo & 1u), "f"(Kq), "f"(k_lo),
              "f"(k_hi), "f"(k100_lo), "f"(k100_hi), "f"(t));
    }

    // Certified: chi_n < chi_t (DM_saliency.py:283, :292), from the quantised normals.  Every t_j is within
    // delta = 2.3e-3 of the exact dn_j (quantisation sqrt(3) * 0.5 / 511 = 1.70e-3 plus 5e-4 of float32 evaluation; dn_j >= 0,
    // so clamping t at 0 keeps the bound), std is a seminorm: std(dn) <= std(t) + delta.  The float32 sums of the
    // non-negative t_j and t_j^2 carry a relative error below eps = (mu + 2) * 2^-23 each, which moves the variance by at
    // most eps * (var + 3 mean^2).  A NaN anywhere makes the comparison false (not certified).
    __device__ __forceinline__ bool dn_settled_by_chi2(double chi_t, double sigma_n) const {
        const double imu = 1.0 / (double)mu, mean = (double)s1 * imu;
        const double var = fmax((double)s2 * imu - mean * mean, 0.0);
        const double eps = ((double)mu + 2.0) * (1.0 / 8388608.0);
        const double std_up = sqrt(var * (1.0 + 2.0 * eps) + 4.0 * eps * mean * mean) + 2.3e-3;
        return (double)(m - 1) * std_up / sigma_n * 1.001 < chi_t;
    }
};



__global__ void __launch_bounds__(WALK_THREADS) dkdn_screen_kernel(GridView g, const FeatRec* __restrict__ feat,
                                                          DkDnParams prm, const double* __restrict__ chi2_table,
                                                          float* __restrict__ dk_out,
                                                          float* __restrict__ dn_out, unsigned* __restrict__ mm,
                                                          int* __restrict__ err, uint32_t* __restrict__ list,
                                                          unsigned long long* __restrict__ aux,
                                                          unsigned* __restrict__ list_count) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < g.n;
    bool owned = false, refine = false, need_n = false;
    unsigned long long counts = 0;      // m, mu, nact and the effective zk of a dk-only entry, 16 bits each
    if (live) {
        const D4 pa = ld_d4<true>(feat, g.n, i, 0), pb = ld_d4<true>(feat, g.n, i, 1);
        ScreenAcc acc;
        acc.px = pa.x; acc.py = pa.y; acc.pz = pa.z;
        const long long kb = __double_as_longlong(pa.w);
        const double Kp = kword_K(kb);
        acc.set_normal(pb.x, pb.y, pb.z);
        bool ok = dk_interval(Kp, 0.04, acc.k_lo, acc.k_hi);
        ok = dk_interval(Kp, prm.zk100_thr, acc.k100_lo, acc.k100_hi) && ok;
        acc.m = acc.mu = acc.nact = acc.zk = acc.zk100 = 0;
        acc.s1 = acc.s2 = 0.0f;
        walk_candidates_pipelined<false>(g, acc.px, acc.py, acc.pz, prm.r, prm.s, [&](int64_t j) {
            return ld_d4<true>(feat, g.n, j, 0);
        }, [&](const D4& q) { acc.add(q, prm); });
        if (!(kb & 2LL) && acc.m >= prm.min_points) {                       // DM_saliency.py:225-229, :240-244
            if (acc.nact >= prm.table_len) {
                atomicMax(err, acc.nact + 1);
            } else {
                const int zk = (acc.zk == acc.nact) ? acc.zk100 : acc.zk;   // :289-290
                const double lim = 0.6 * (double)acc.nact;
                const bool k_zero = ok && (double)zk > lim;                 // :309-311
                // dn = 0 if chi_n < chi_t (:292-293): certified upper bound of chi_n from the packed normals
                need_n = !acc.dn_settled_by_chi2(__ldg(chi2_table + acc.nact), prm.sigma_n);
                // the refinement of a dk-only point takes its (exact) counters from here; a counter that does not fit
                // 16 bits sends the point through the whole expression instead
                if (acc.m >= 65536 || !ok) need_n = true;
                refine = !k_zero || need_n;
                counts = (unsigned long long)acc.m | ((unsigned long long)acc.mu << 16) |
                         ((unsigned long long)acc.nact << 32) | ((unsigned long long)zk << 48);
            }
        }
        const bool all_owned = prm.own_lo <= 0 && prm.own_hi >= g.n;
        const int64_t orig = (prm.grid_order && all_owned)
                                 ? 0 : __double_as_longlong(__ldg(reinterpret_cast<const double*>(g.rec + i) + 3));
        if (!refine) {
            const int64_t o = prm.grid_order ? i : orig;
            dk_out[o] = 0.0f;
            dn_out[o] = 0.0f;
        }
        owned = all_owned || (orig >= prm.own_lo && orig < prm.own_hi);
    }
    // work lists: the entries of one CTA (256 consecutive sorted points) are appended as ONE contiguous chunk per list
    // -- shared-memory atomics order the warps inside the chunk, one global atomic per CTA and list places it -- so that
    // the refinement pass, which takes consecutive entries per CTA, works on spatially compact sets (L1 reuse).
    // Points whose dn is settled (dk only) fill the buffer from the front, the rare ones that need the whole expression
    // from the back.
    __shared__ unsigned s_cnt[2], s_base[2];
    if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const unsigned vote_k = __ballot_sync(0xffffffffu, refine && !need_n);
    const unsigned vote_n = __ballot_sync(0xffffffffu, refine && need_n);
    unsigned off_k = 0, off_n = 0;
    if (vote_k) {
        const int leader = __ffs(vote_k) - 1;
        if (lane == leader) off_k = atomicAdd(&s_cnt[0], (unsigned)__popc(vote_k));
        off_k = __shfl_sync(0xffffffffu, off_k, leader) + __popc(vote_k & ((1u << lane) - 1u));
    }
    if (vote_n) {
        const int leader = __ffs(vote_n) - 1;
        if (lane == leader) off_n = atomicAdd(&s_cnt[1], (unsigned)__popc(vote_n));
        off_n = __shfl_sync(0xffffffffu, off_n, leader) + __popc(vote_n & ((1u << lane) - 1u));
    }
    __syncthreads();
    if (threadIdx.x < 2 && s_cnt[threadIdx.x]) s_base[threadIdx.x] = atomicAdd(list_count + threadIdx.x, s_cnt[threadIdx.x]);
    __syncthreads();
    if (refine && !need_n) {
        const unsigned pos = s_base[0] + off_k;
        list[pos] = (uint32_t)i;
        aux[pos] = counts;
    }
    if (refine && need_n) list[(size_t)g.n - 1 - (s_base[1] + off_n)] = (uint32_t)i;
    if (mm) {
        const bool c0 = live && owned && !refine;       // refined points enter the extrema in pass 2
        const unsigned z = f32_ordered(0.0f);
        dkdn_minmax(c0 ? z : 0xffffffffu, c0 ? z : 0u, c0 ? z : 0xffffffffu, c0 ? z : 0u, mm);
    }
}

constexpr int REFINE_WARPS = WALK_THREADS / 32;

// The sums of the dk half of DkdnAcc::add (same expressions, same order: bit-identical dk), for the points whose dn the
// screening pass has settled -- practically all of them: no normal is loaded or multiplied, and the counters (m, mu,
// nact, zk) come from the screening pass, which computed them exactly.
struct DkAcc {
    double px, py, pz, Kp;
    double s1k, s2k, wk, wsum;

    __device__ __forceinline__ void init(const D4& pa) {
        px = pa.x; py = pa.y; pz = pa.z;
        Kp = kword_K(__double_as_longlong(pa.w));
        s1k = s2k = wk = wsum = 0.0;
    }
    // valid = false: the record is a clamped re-read past the end of the candidate list and adds nothing (the loop
    // around this has no branches: every lane of a group executes the same steps)
    __device__ __forceinline__ void add(const D4& qa, bool valid, const DkDnParams& prm) {
        const double d2 = dist2_rule(px - qa.x, py - qa.y, pz - qa.z);
        const long long kb = __double_as_longlong(qa.w);
        const bool in = valid & (d2 <= prm.r2);
        const bool u = in & !((int)kb & 1);
        double dk = Kp - kword_K(kb);
        dk = u ? dk : 0.0;
        const double t = __dmul_rn(prm.inv_rho2, d2);
        double w = (d2 < prm.rho2) ? t : __dsub_rn(2.0, t);
        const bool keep = u & (d2 != prm.rho2) & !(w < 0.0);
        w = keep ? w : 0.0;
        s1k += dk; s2k += dk * dk;
        wk += dk * w; wsum += w;
    }
};

// G lanes per listed point, 32 / G points per warp: the per-point work outside the candidate loop (own record, row
// table, shuffle tree, finish) is shared by the 32 / G points a warp handles at once, and that many independent
// dependency chains are in flight per warp.  The G lanes stride over the point's concatenated stencil rows:
// rows: s_off[k] = first flattened index of row k (s_off[nrows] = total), s_b[k] = its first record.
constexpr int REFINE_G = 4;
constexpr int REFINE_ROWS = 32;        // stencil rows per point (2 sy + 1 <= 31, checked by the host)

struct RefineRows {
    const uint32_t* sb;
    const uint32_t* so;
    uint32_t total, next;     // next: first flattened index of row r + 1
    int64_t delta;            // first record of row r minus its first flattened index
    int r;
    __device__ __forceinline__ void start() {
        r = 0;
        next = so[1];
        delta = (int64_t)sb[0] - (int64_t)so[0];
    }
    // sorted position of flattened candidate t < total (t must not decrease between calls): the common case costs a
    // compare and an add, the row table in shared memory is read only when t crosses into another row
    __device__ __forceinline__ int64_t locate(uint32_t t) {
        while (t >= next) {
            ++r;
            next = so[r + 1];
            delta = (int64_t)sb[r] - (int64_t)so[r];
        }
        return (int64_t)t + delta;
    }
};

// the row table of point (px, py) for the G lanes of its group: the same ranges walk_candidates_pipelined visits
template <int G>
__device__ __forceinline__ RefineRows refine_rows(const GridView& g, const DkDnParams& prm, bool valid, double px, double py,
                                                  uint32_t* sb, uint32_t* so, int gl) {
    const int sx = M3D_SX(prm.s), sy = M3D_SY(prm.s);
    const int ix = g.col_x(px), iy = g.col_y(py);
    const int y0 = max(iy - sy, 0), y1 = min(iy + sy, g.ny - 1);
    const int nrows = (!valid || ix + sx < 0 || ix - sx > g.nx - 1 || y0 > y1) ? 0 : y1 - y0 + 1;
    __syncwarp();
    for (int k = gl; k < nrows; k += G) {
        const int y = y0 + k, w = row_halfwidth(g, prm.s, y - iy);
        const uint32_t* cs = g.col_start + (size_t)y * g.nx;
        const uint32_t b = __ldg(cs + min(max(ix - w, 0), g.nx));
        sb[k] = b;
        so[k + 1] = __ldg(cs + min(max(ix + w + 1, 0), g.nx)) - b;
    }
    __syncwarp();
    if (gl == 0) {
        uint32_t run = 0;
        so[0] = 0;
        for (int k = 1; k <= nrows; ++k) { run += so[k]; so[k] = run; }
    }
    __syncwarp();
    RefineRows rr;
    rr.sb = sb; rr.so = so; rr.total = so[nrows];
    if (nrows == 0) so[1] = 0;                 // start() reads the first row's end
    __syncwarp();
    rr.start();
    return rr;
}

// pass 2a: the points whose dn is settled -- dk only, no normals.  Candidate records are double-buffered two at a time
// (the next pair is requested before the current one is consumed).
template <int G>
__global__ void __launch_bounds__(WALK_THREADS, 3) dk_refine_kernel(GridView g, const FeatRec* __restrict__ feat,
                                                           DkDnParams prm, const double* __restrict__ chi2_table,
                                                           float* __restrict__ dk_out, float* __restrict__ dn_out,
                                                           unsigned* __restrict__ mm, int* __restrict__ err,
                                                           const uint32_t* __restrict__ list,
                                                           const unsigned long long* __restrict__ aux,
                                                           const unsigned* __restrict__ list_count, unsigned gate_above) {
    constexpr int GP = 32 / G;
    __shared__ uint32_t s_b[REFINE_WARPS][GP][REFINE_ROWS], s_off[REFINE_WARPS][GP][REFINE_ROWS + 1];
    if (refine_gate_open(list_count, gate_above)) return;      // the one-pass kernel takes over
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, grp = lane / G, gl = lane % G;
    const unsigned count = list_count[0];
    const unsigned stride = gridDim.x * REFINE_WARPS * GP;
    const bool all_owned = prm.own_lo <= 0 && prm.own_hi >= g.n;
    unsigned lo_k = 0xffffffffu, hi_k = 0u;
    bool any = false;
    unsigned base = (blockIdx.x * REFINE_WARPS + wib) * GP;
    // the next point's list entry and own record are requested one iteration ahead
    uint32_t ent_next = (base + grp < count) ? __ldg(list + base + grp) : 0u;
    D4 pa_next = ld_d4<true>(feat, g.n, (int64_t)ent_next, 0);
    for (; base < count; base += stride) {                      // warp-uniform
        const bool valid = base + grp < count;
        const int64_t i = (int64_t)ent_next;
        const D4 pa = pa_next;
        ent_next = (base + stride + grp < count) ? __ldg(list + base + stride + grp) : 0u;
        RefineRows rr = refine_rows<G>(g, prm, valid, pa.x, pa.y, s_b[wib][grp], s_off[wib][grp], gl);
        pa_next = ld_d4<true>(feat, g.n, (int64_t)ent_next, 0);
        DkAcc dk;
        dk.init(pa);
        if (rr.total) {
            const uint32_t total = rr.total, last = total - 1;
            uint32_t t = gl;
            // two records in flight while two are consumed; reads past the end are clamped to the last candidate and masked
            auto fetch = [&](uint32_t tt) { return ld_d4<true>(feat, g.n, rr.locate(min(tt, last)), 0); };
            D4 c0 = fetch(t), c1 = fetch(t + G), n0, n1;
            while (true) {
                n0 = fetch(t + 2 * G);
                n1 = fetch(t + 3 * G);
                dk.add(c0, t < total, prm);
                dk.add(c1, t + G < total, prm);
                t += 2 * G;
                if (t >= total) break;
                c0 = fetch(t + 2 * G);
                c1 = fetch(t + 3 * G);
                dk.add(n0, t < total, prm);
                dk.add(n1, t + G < total, prm);
                t += 2 * G;
                if (t >= total) break;
            }
        }
        __syncwarp();
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) {
            dk.s1k += __shfl_xor_sync(0xffffffffu, dk.s1k, o);
            dk.s2k += __shfl_xor_sync(0xffffffffu, dk.s2k, o);
            dk.wk += __shfl_xor_sync(0xffffffffu, dk.wk, o);
            dk.wsum += __shfl_xor_sync(0xffffffffu, dk.wsum, o);
        }
        if (gl == 0 && valid) {
            // the dk half of DkdnAcc::finish, same expressions (the screening pass has checked the normal, m and nact)
            const unsigned long long cw = __ldg(aux + base + grp);
            const int m = (int)(cw & 0xffffu), mu = (int)((cw >> 16) & 0xffffu), nact = (int)((cw >> 32) & 0xffffu),
                      zk = (int)(cw >> 48);
            const double chi_t = chi2_table[nact];
            const double imu = 1.0 / (double)mu;
            const double mk = dk.s1k * imu;
            const double std_k = sqrt(fmax(dk.s2k * imu - mk * mk, 0.0));
            const double sdk = isnan(dk.s1k) ? dk.s1k : std_k;
            const double chi_k = (double)(m - 1) * sdk / prm.sigma_k;
            const double lim = 0.6 * (double)nact;
            double v;
            if (chi_k < chi_t) v = 0.0;
            else if ((double)zk > lim) v = 0.0;
            else v = fabs(dk.wk / dk.wsum);
            const float dkf = (float)v;
            const int64_t orig = (prm.grid_order && all_owned)
                                     ? 0 : __double_as_longlong(__ldg(reinterpret_cast<const double*>(g.rec + i) + 3));
            const int64_t o = prm.grid_order ? i : orig;
            dk_out[o] = dkf;
            dn_out[o] = 0.0f;
            const bool owned = all_owned || (orig >= prm.own_lo && orig < prm.own_hi);
            if (owned) {
                any = true;
                if (!isnan(dkf)) { const unsigned u = f32_ordered(dkf); lo_k = min(lo_k, u); hi_k = max(hi_k, u); }
            }
        }
    }
    if (mm && hi_k >= lo_k) { atomicMin(mm + 0, lo_k); atomicMax(mm + 1, hi_k); }
    if (mm && any) { const unsigned z = f32_ordered(0.0f); atomicMin(mm + 2, z); atomicMax(mm + 3, z); }
}

// pass 2a, one THREAD per listed point (the row walk of the one-pass kernels over the compacted list)
__global__ void __launch_bounds__(WALK_THREADS) dk_refine_thread_kernel(GridView g, const FeatRec* __restrict__ feat,
                                                               DkDnParams prm, const double* __restrict__ chi2_table,
                                                               float* __restrict__ dk_out, float* __restrict__ dn_out,
                                                               unsigned* __restrict__ mm,
                                                               const uint32_t* __restrict__ list,
                                                               const unsigned long long* __restrict__ aux,
                                                               const unsigned* __restrict__ list_count, unsigned gate_above) {
    if (refine_gate_open(list_count, gate_above)) return;
    const unsigned count = list_count[0];
    const bool all_owned = prm.own_lo <= 0 && prm.own_hi >= g.n;
    unsigned lo_k = 0xffffffffu, hi_k = 0u;
    bool any = false;
    for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < count; e += gridDim.x * blockDim.x) {
        const int64_t i = (int64_t)__ldg(list + e);
        DkAcc dk;
        dk.init(ld_d4<true>(feat, g.n, i, 0));
        walk_candidates_pipelined<false>(g, dk.px, dk.py, dk.pz, prm.r, prm.s, [&](int64_t j) {
            return ld_d4<true>(feat, g.n, j, 0);
        }, [&](const D4& c) { dk.add(c, true, prm); });
        const unsigned long long cw = __ldg(aux + e);
        const int m = (int)(cw & 0xffffu), mu = (int)((cw >> 16) & 0xffffu), nact = (int)((cw >> 32) & 0xffffu),
                  zk = (int)(cw >> 48);
        const double chi_t = chi2_table[nact];
        const double imu = 1.0 / (double)mu;
        const double mk = dk.s1k * imu;
        const double std_k = sqrt(fmax(dk.s2k * imu - mk * mk, 0.0));
        const double sdk = isnan(dk.s1k) ? dk.s1k : std_k;
        const double chi_k = (double)(m - 1) * sdk / prm.sigma_k;
        const double lim = 0.6 * (double)nact;
        double v;
        if (chi_k < chi_t) v = 0.0;
        else if ((double)zk > lim) v = 0.0;
        else v = fabs(dk.wk / dk.wsum);
        const float dkf = (float)v;
        const int64_t orig = (prm.grid_order && all_owned)
                                 ? 0 : __double_as_longlong(__ldg(reinterpret_cast<const double*>(g.rec + i) + 3));
        const int64_t o = prm.grid_order ? i : orig;
        dk_out[o] = dkf;
        dn_out[o] = 0.0f;
        if (all_owned || (orig >= prm.own_lo && orig < prm.own_hi)) {
            any = true;
            if (!isnan(dkf)) { const unsigned u = f32_ordered(dkf); lo_k = min(lo_k, u); hi_k = max(hi_k, u); }
        }
    }
    if (mm && hi_k >= lo_k) { atomicMin(mm + 0, lo_k); atomicMax(mm + 1, hi_k); }
    if (mm && any) { const unsigned z = f32_ordered(0.0f); atomicMin(mm + 2, z); atomicMax(mm + 3, z); }
}

// pass 2b: the (rare) points that need the whole expression, normals included; entries sit at the END of the list
template <int G>
__global__ void __launch_bounds__(WALK_THREADS, 2) dkdn_refine_kernel(GridView g, const FeatRec* __restrict__ feat,
                                                          DkDnParams prm, const double* __restrict__ chi2_table,
                                                          float* __restrict__ dk_out, float* __restrict__ dn_out,
                                                          unsigned* __restrict__ mm, int* __restrict__ err,
                                                          const uint32_t* __restrict__ list,
                                                          const unsigned* __restrict__ list_count, unsigned gate_above) {
    constexpr int GP = 32 / G;
    __shared__ uint32_t s_b[REFINE_WARPS][GP][REFINE_ROWS], s_off[REFINE_WARPS][GP][REFINE_ROWS + 1];
    if (refine_gate_open(list_count, gate_above)) return;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, grp = lane / G, gl = lane % G;
    const unsigned count = list_count[1];
    const unsigned stride = gridDim.x * REFINE_WARPS * GP;
    const bool all_owned = prm.own_lo <= 0 && prm.own_hi >= g.n;
    unsigned lo_k = 0xffffffffu, hi_k = 0u, lo_n = 0xffffffffu, hi_n = 0u;
    for (unsigned base = (blockIdx.x * REFINE_WARPS + wib) * GP; base < count; base += stride) {   // warp-uniform
        const unsigned e = base + grp;
        const bool valid = e < count;
        const int64_t i = valid ? (int64_t)__ldg(list + ((size_t)g.n - 1 - e)) : 0;
        const D4 pa = ld_d4<true>(feat, g.n, i, 0);
        RefineRows rr = refine_rows<G>(g, prm, valid, pa.x, pa.y, s_b[wib][grp], s_off[wib][grp], gl);
        DkdnAcc acc;
        acc.init(pa, ld_d4<true>(feat, g.n, i, 1));
        for (uint32_t t = gl; t < rr.total; t += G) {
            const int64_t j = rr.locate(t);
            acc.add(ld_d4<true>(feat, g.n, j, 0), ld_d4<true>(feat, g.n, j, 1), prm);
        }
        __syncwarp();
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) {
            acc.m += __shfl_xor_sync(0xffffffffu, acc.m, o);
            acc.mu += __shfl_xor_sync(0xffffffffu, acc.mu, o);
            acc.nact += __shfl_xor_sync(0xffffffffu, acc.nact, o);
            acc.zn += __shfl_xor_sync(0xffffffffu, acc.zn, o);
            acc.zk += __shfl_xor_sync(0xffffffffu, acc.zk, o);
            acc.zk100 += __shfl_xor_sync(0xffffffffu, acc.zk100, o);
            acc.s1k += __shfl_xor_sync(0xffffffffu, acc.s1k, o);
            acc.s2k += __shfl_xor_sync(0xffffffffu, acc.s2k, o);
            acc.wk += __shfl_xor_sync(0xffffffffu, acc.wk, o);
            acc.wsum += __shfl_xor_sync(0xffffffffu, acc.wsum, o);
            acc.s1n += __shfl_xor_sync(0xffffffffu, acc.s1n, o);
            acc.s2n += __shfl_xor_sync(0xffffffffu, acc.s2n, o);
            acc.wn += __shfl_xor_sync(0xffffffffu, acc.wn, o);
        }
        if (gl == 0 && valid) {
            float dkf, dnf;
            acc.finish(prm, chi2_table, err, dkf, dnf);
            const int64_t orig = (prm.grid_order && all_owned)
                                     ? 0 : __double_as_longlong(__ldg(reinterpret_cast<const double*>(g.rec + i) + 3));
            const int64_t o = prm.grid_order ? i : orig;
            dk_out[o] = dkf;
            dn_out[o] = dnf;
            const bool owned = all_owned || (orig >= prm.own_lo && orig < prm.own_hi);
            if (owned && !isnan(dkf)) { const unsigned u = f32_ordered(dkf); lo_k = min(lo_k, u); hi_k = max(hi_k, u); }
            if (owned && !isnan(dnf)) { const unsigned u = f32_ordered(dnf); lo_n = min(lo_n, u); hi_n = max(hi_n, u); }
        }
    }
    if (mm && hi_k >= lo_k) { atomicMin(mm + 0, lo_k); atomicMax(mm + 1, hi_k); }
    if (mm && hi_n >= lo_n) { atomicMin(mm + 2, lo_n); atomicMax(mm + 3, hi_n); }
}

// ------------------------------------------------------------------------------------------
// K6, shared-memory staged variant (BASELINE.json north_star: "shared-memory staging of candidate cells and
// warp-cooperative, vectorised float4 point loads").  One CTA per block of CB columns of one grid row: the CTA's
// queries are the points of those columns (a contiguous run of sorted positions), their candidates the (2s+1)
// contiguous record ranges of the block's stencil rows widened by s columns.  All threads copy those ranges with
// coalesced 16-byte loads into shared memory, split into four arrays of 16-byte quarter records so that lanes at
// neighbouring records hit different banks and lanes at the same record are served by one broadcast; every thread
// then walks ITS OWN sub-range of every row out of shared memory -- same candidates, same order, same arithmetic as
// dkdn_kernel, hence bit-identical outputs.  A block whose ranges exceed the staging capacity walks global memory.
// ------------------------------------------------------------------------------------------
constexpr int SMW_THREADS = 128;
constexpr int SMW_MAX_ROWS = 13;       // s <= 6

__device__ __forceinline__ D4 lds_d4(const double2* __restrict__ lo, const double2* __restrict__ hi, uint32_t j) {
    const double2 a = lo[j], b = hi[j];
    D4 r;
    r.x = a.x; r.y = a.y; r.z = b.x; r.w = b.y;
    return r;
}

template <bool SPLIT>
__global__ void __launch_bounds__(SMW_THREADS) dkdn_smem_kernel(GridView g, const FeatRec* __restrict__ feat,
                                                                DkDnParams prm, const double* __restrict__ chi2_table,
                                                                float* __restrict__ dk_out, float* __restrict__ dn_out,
                                                                unsigned* __restrict__ mm, int* __restrict__ err,
                                                                int cb, int nbx, int cap) {
    extern __shared__ __align__(16) unsigned char smw[];
    double2* q0 = reinterpret_cast<double2*>(smw);          // (x, y)
    double2* q1 = q0 + cap;                                 // (z, K | flags)
    double2* q2 = q1 + cap;                                 // (nx, ny)
    double2* q3 = q2 + cap;                                 // (nz, pad)
    uint32_t* cs = reinterpret_cast<uint32_t*>(q3 + cap);   // rows x csw: staged offsets of the widened columns
    __shared__ uint32_t row_b[SMW_MAX_ROWS], row_off[SMW_MAX_ROWS + 1];

    const int iy = blockIdx.x / nbx, bx = blockIdx.x - iy * nbx;
    const int c0 = bx * cb, c1 = min(c0 + cb, g.nx);
    const uint32_t qa = __ldg(g.col_start + (size_t)iy * g.nx + c0), qb = __ldg(g.col_start + (size_t)iy * g.nx + c1);
    if (qa == qb) return;                                   // no query in this block (uniform for the CTA)
    const int s = M3D_SX(prm.s), sy = M3D_SY(prm.s);
    const int y0 = max(iy - sy, 0), y1 = min(iy + sy, g.ny - 1), nrows = y1 - y0 + 1;
    const int xs0 = max(c0 - s, 0), xs1 = min(c1 - 1 + s, g.nx - 1), csw = cb + 2 * s + 1;
    if ((int)threadIdx.x < nrows) {
        const uint32_t* row = g.col_start + (size_t)(y0 + threadIdx.x) * g.nx;
        row_b[threadIdx.x] = __ldg(row + xs0);
        row_off[threadIdx.x + 1] = __ldg(row + xs1 + 1) - row_b[threadIdx.x];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int r = 0; r < nrows; ++r) { const uint32_t len = row_off[r + 1]; row_off[r] = run; run += len; }
        row_off[nrows] = run;
    }
    __syncthreads();
    const uint32_t total = row_off[nrows];
    const bool staged = total <= (uint32_t)cap;
    if (staged) {
        // column offsets of the widened block, as positions in the staged arrays
        for (int t = threadIdx.x; t < nrows * csw; t += blockDim.x) {
            const int r = t / csw, c = t - r * csw;
            const int x = min(xs0 + c, xs1 + 1);
            cs[t] = __ldg(g.col_start + (size_t)(y0 + r) * g.nx + x) - row_b[r] + row_off[r];
        }
        // warp-cooperative copy: consecutive threads take consecutive 16-byte quarters of the contiguous range
        const double2* __restrict__ src = reinterpret_cast<const double2*>(feat);
        for (int r = 0; r < nrows; ++r) {
            const uint32_t len2 = (row_off[r + 1] - row_off[r]) * 2u;       // 16-byte quarters per half array
            const size_t gbase = (size_t)row_b[r] * 2u;
            const uint32_t sbase = row_off[r];
            for (uint32_t t = threadIdx.x; t < 2u * len2; t += blockDim.x) {
                const uint32_t half = t >= len2 ? 1u : 0u, u = t - half * len2;
                const double2 v = SPLIT ? __ldg(src + (size_t)half * (size_t)g.n * 2u + gbase + u)
                                           : __ldg(src + 2 * (gbase + (u & ~1u)) + 2 * half + (u & 1u));
                const uint32_t rec = u >> 1, k = 2u * half + (u & 1u);
                double2* dst = k == 0 ? q0 : (k == 1 ? q1 : (k == 2 ? q2 : q3));
                dst[sbase + rec] = v;
            }
        }
    }
    __syncthreads();

    unsigned lo_k = 0xffffffffu, hi_k = 0u, lo_n = 0xffffffffu, hi_n = 0u;
    const bool all_owned = prm.own_lo <= 0 && prm.own_hi >= g.n;
    const int ry = iy - y0;
    for (uint32_t q = qa + threadIdx.x; q < qb; q += blockDim.x) {
        DkdnAcc acc;
        float dkf, dnf;
        if (staged) {
            const uint32_t o = row_off[ry] + (q - row_b[ry]);
            acc.init(lds_d4(q0, q1, o), lds_d4(q2, q3, o));
            const int ix = min(max(g.col_x(acc.px), c0), c1 - 1);
            const int xa = max(ix - s, 0) - xs0, xb = min(ix + s, g.nx - 1) - xs0 + 1;
            for (int r = 0; r < nrows; ++r) {
                const uint32_t b = cs[r * csw + xa], e = cs[r * csw + xb];
                run_range_pipelined(b, e, [&](int64_t j) {
                    D4x2 c;
                    c.a = lds_d4(q0, q1, (uint32_t)j);
                    c.b = lds_d4(q2, q3, (uint32_t)j);
                    return c;
                }, [&](const D4x2& c) { acc.add(c.a, c.b, prm); });
            }
        } else {
            acc.init(ld_d4<SPLIT>(feat, g.n, q, 0), ld_d4<SPLIT>(feat, g.n, q, 1));
            walk_candidates_pipelined<false>(g, acc.px, acc.py, acc.pz, prm.r, prm.s, [&](int64_t j) {
                D4x2 c;
                c.a = ld_d4<SPLIT>(feat, g.n, j, 0);
                c.b = ld_d4<SPLIT>(feat, g.n, j, 1);
                return c;
            }, [&](const D4x2& c) { acc.add(c.a, c.b, prm); });
        }
        acc.finish(prm, chi2_table, err, dkf, dnf);
        const int64_t orig = (prm.grid_order && all_owned)
                                 ? 0 : __double_as_longlong(__ldg(reinterpret_cast<const double*>(g.rec + q) + 3));
        const int64_t o = prm.grid_order ? (int64_t)q : orig;
        dk_out[o] = dkf;
        dn_out[o] = dnf;
        const bool owned = all_owned || (orig >= prm.own_lo && orig < prm.own_hi);
        if (owned && !isnan(dkf)) { const unsigned u = f32_ordered(dkf); lo_k = min(lo_k, u); hi_k = max(hi_k, u); }
        if (owned && !isnan(dnf)) { const unsigned u = f32_ordered(dnf); lo_n = min(lo_n, u); hi_n = max(hi_n, u); }
    }
    if (mm) dkdn_minmax(lo_k, hi_k, lo_n, hi_n, mm);
}

__global__ void minmax_init(unsigned* mm) {
    if (threadIdx.x < 4) mm[threadIdx.x] = (threadIdx.x & 1) ? 0u : 0xffffffffu;
}
__global__ void minmax_finish(const unsigned* mm, float* out) {
    if (threadIdx.x < 4) out[threadIdx.x] = f32_unordered(mm[threadIdx.x]);
}

}  // namespace m3d

using namespace m3d;

static int launch_moments(const more3d_grid* g, int mode, double r, double eps, int min_points,
                          double* normals, double* lam_ratio, float* K, int32_t* pcount, cudaStream_t s,
                          int grid_order = 0) {
    M3D_REQUIRE(g != nullptr, "grid is NULL");
    M3D_REQUIRE(r >= 0 && isfinite(r), "radius must be finite and >= 0");
    M3D_REQUIRE(normals != nullptr, "normals is NULL");
    CurvParams prm;
    prm.r = r; prm.r2 = r * r; prm.eps = eps; prm.min_points = min_points; prm.s = stencil_halfwidth(g, r);
    prm.flt = make_filter(g, r);
    prm.grid_order = grid_order;
    const unsigned nb = (unsigned)((g->n + WALK_THREADS - 1) / WALK_THREADS);
    GridView v = make_view(g);
    set_row_clip(v, g, r);
    ProfScope prof("moments", s);
    if (mode == MODE_NORMALS) {
        if (g->tall) moments_kernel<MODE_NORMALS, true><<<nb, WALK_THREADS, 0, s>>>(v, prm, normals, lam_ratio, nullptr, pcount, nullptr);
        else moments_kernel<MODE_NORMALS, false><<<nb, WALK_THREADS, 0, s>>>(v, prm, normals, lam_ratio, nullptr, pcount, nullptr);
    } else if (mode == MODE_CURV) {
        if (g->tall) moments_kernel<MODE_CURV, true><<<nb, WALK_THREADS, 0, s>>>(v, prm, normals, nullptr, K, pcount, nullptr);
        else moments_kernel<MODE_CURV, false><<<nb, WALK_THREADS, 0, s>>>(v, prm, normals, nullptr, K, pcount, nullptr);
    } else {
        more3d_grid* gm = const_cast<more3d_grid*>(g);     // the feature-record cache is mutable state of the grid
        if (!gm->feat && cudaMallocAsync((void**)&gm->feat, (size_t)g->n * sizeof(FeatRec), s) != cudaSuccess) {
            cudaGetLastError();
            gm->feat = nullptr;                            // no cache: dk/dn will gather instead
        }
        if (g->tall) moments_kernel<MODE_FUSED, true><<<nb, WALK_THREADS, 0, s>>>(v, prm, normals, lam_ratio, K, pcount, gm->feat);
        else moments_kernel<MODE_FUSED, false><<<nb, WALK_THREADS, 0, s>>>(v, prm, normals, lam_ratio, K, pcount, gm->feat);
    }
    count_launches(1);
    M3D_CHECK_LAUNCH();
    return MORE3D_OK;
}

extern "C" int more3d_normals_pca(const more3d_grid* g, double r, double* normals, double* lam_ratio,
                                  int32_t* counts, void* stream) {
    return launch_moments(g, MODE_NORMALS, r, 0.0, 0, normals, lam_ratio, nullptr, counts, (cudaStream_t)stream);
}

extern "C" int more3d_nonparam_curvature(const more3d_grid* g, double r, double* normals_inout, double eps,
                                         int min_points, float* K, int32_t* pcount, void* stream) {
    M3D_REQUIRE(K != nullptr && pcount != nullptr, "K / pcount is NULL");
    return launch_moments(g, MODE_CURV, r, eps, min_points, normals_inout, nullptr, K, pcount, (cudaStream_t)stream);
}

extern "C" int more3d_normals_curvature(const more3d_grid* g, double r, double eps, int min_points,
                                        double* normals, double* lam_ratio, float* K, int32_t* pcount,
                                        void* stream) {
    M3D_REQUIRE(K != nullptr && pcount != nullptr, "K / pcount is NULL");
    return launch_moments(g, MODE_FUSED, r, eps, min_points, normals, lam_ratio, K, pcount, (cudaStream_t)stream);
}

// kernel variant of the dk/dn pass: 0 = two passes (screening walk + warp-per-point refinement; default on thin grids),
// 1 = shared-memory staged one-pass walk, 2 = global-memory one-pass walk (always used on tall grids).
// MORE3D_DKDN_VARIANT / MORE3D_SMEM_CAP in the environment select it for A/B measurements (read once).
namespace m3d {
static int g_dkdn_variant = -1;
int dkdn_variant() {
    if (g_dkdn_variant < 0) { const char* e = getenv("MORE3D_DKDN_VARIANT"); g_dkdn_variant = e ? atoi(e) : 0; }
    return g_dkdn_variant;
}
}  // namespace m3d
// host-only: the two thresholds of the screening pass's active test for a given rho (no device needed; the CPU tests
// check them against the reference's expression evaluated in numpy)
extern "C" int more3d_dkdn_act_interval(double rho, double* act_lo_h, double* act_hi_h) {
    M3D_REQUIRE(act_lo_h && act_hi_h, "NULL argument");
    const double rho2 = rho * rho;
    M3D_REQUIRE(act_interval(rho2, 1.0 / rho2, act_lo_h, act_hi_h), "rho is degenerate: no screening thresholds");
    return MORE3D_OK;
}
static double g_dkdn_gate = 0.45;
extern "C" int more3d_set_dkdn_gate(double fraction) {
    M3D_REQUIRE(fraction >= 0.0 && isfinite(fraction), "fraction must be finite and >= 0");
    g_dkdn_gate = fraction;
    return MORE3D_OK;
}
extern "C" int more3d_set_dkdn_variant(int variant) {
    M3D_REQUIRE(variant >= 0 && variant <= 2, "variant must be 0 (two passes), 1 (shared-memory staged) or 2 (one pass)");
    m3d::g_dkdn_variant = variant;
    return MORE3D_OK;
}
// persistent grids of the refinement passes: every SM filled once (queried once per device)
template <typename K>
static int persistent_grid(K kernel, int fallback_per_sm) {
    int dev = 0, sms = 148, per = fallback_per_sm;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, kernel, WALK_THREADS, 0) != cudaSuccess || per < 1) {
        cudaGetLastError();
        per = fallback_per_sm;
    }
    return sms * per;
}
static int smem_walk_cap() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("MORE3D_SMEM_CAP"); v = e ? atoi(e) : 640; if (v < 64) v = 64; if (v > 3000) v = 3000; }
    return v;
}

static int run_dk_dn(const more3d_grid* g, double r, const double* normals, const float* K, double rho,
                     double sigma_normal, double sigma_curvature, const double* chi2_table,
                     int table_len, int min_points, float* dk, float* dn, float* minmax, int32_t* status, void* stream,
                     int grid_order, int64_t own_lo, int64_t n_owned) {
    M3D_REQUIRE(g != nullptr, "grid is NULL");
    M3D_REQUIRE(r >= 0 && isfinite(r), "radius must be finite and >= 0");
    M3D_REQUIRE(chi2_table && dk && dn, "NULL argument");
    M3D_REQUIRE((normals != nullptr) == (K != nullptr), "normals and K must both be given or both be NULL");
    const bool cached = normals == nullptr;
    M3D_REQUIRE(!cached || g->feat != nullptr, "no cached feature records: run more3d_normals_curvature on this grid first");
    M3D_REQUIRE(table_len > 0, "chi2 table is empty");
    cudaStream_t s = (cudaStream_t)stream;
    Scratch sc(s);
    FeatRec* feat = nullptr;
    unsigned* mm = nullptr;
    int* err = nullptr;
    if (cached) feat = g->feat;
    else M3D_TRY(sc.alloc(&feat, (size_t)g->n));
    M3D_TRY(sc.alloc(&mm, 4));
    if (status) {
        err = status;       // the caller's status word: raised on overflow, never cleared, read when the caller likes
    } else {
        M3D_TRY(sc.alloc(&err, 1));
        M3D_CUDA(cudaMemsetAsync(err, 0, sizeof(int), s));
    }
    const unsigned nb = (unsigned)((g->n + 255) / 256);
    GridView v = make_view(g);
    set_row_clip(v, g, r);
    if (!cached) {
        ProfScope prof("build_feat", s);
        build_feat_kernel<<<nb, 256, 0, s>>>(v, normals, K, feat);
    }
    minmax_init<<<1, 32, 0, s>>>(mm);
    DkDnParams prm;
    prm.r = r; prm.r2 = r * r; prm.rho2 = rho * rho; prm.inv_rho2 = 1.0 / (rho * rho);
    prm.sigma_n = sigma_normal; prm.sigma_k = sigma_curvature;
    prm.min_points = min_points; prm.s = stencil_halfwidth(g, r); prm.table_len = table_len;
    prm.flt = make_filter(g, r);
    prm.zk100_thr = zk100_threshold();
    prm.grid_order = grid_order;
    prm.own_lo = (n_owned <= 0) ? 0 : own_lo;
    prm.own_hi = (n_owned <= 0) ? g->n : own_lo + n_owned;
    {
        ProfScope prof("dkdn", s);
        const int variant = dkdn_variant();
        const bool split = g->feat_split != 0;
        if (g->tall) {
            if (split) dkdn_kernel<true, true><<<nb, 256, 0, s>>>(v, feat, prm, chi2_table, dk, dn, minmax ? mm : nullptr, err);
            else dkdn_kernel<true, false><<<nb, 256, 0, s>>>(v, feat, prm, chi2_table, dk, dn, minmax ? mm : nullptr, err);
        } else if (variant == 1 && M3D_SY(prm.s) <= (SMW_MAX_ROWS - 1) / 2) {
            // staging capacity in records and the column-block width that fills ~80 % of it at the grid's mean density
            const int cap = smem_walk_cap();
            const double dens = (double)g->n / ((double)g->nx * (double)g->ny);
            const int sxh = M3D_SX(prm.s), syh = M3D_SY(prm.s);
            int cb = (int)(0.8 * cap / ((2 * syh + 1) * fmax(dens, 1e-3))) - 2 * sxh;
            cb = cb < 8 ? 8 : (cb > 96 ? 96 : cb);
            const int nbx = (g->nx + cb - 1) / cb;
            const size_t smem = (size_t)cap * 64 + (size_t)SMW_MAX_ROWS * (cb + 2 * sxh + 1) * sizeof(uint32_t);
            M3D_CUDA(cudaFuncSetAttribute(dkdn_smem_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            M3D_CUDA(cudaFuncSetAttribute(dkdn_smem_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int threads = ((int)(cb * dens * 1.15) + 31) / 32 * 32;          // one thread per expected query
            threads = threads < 32 ? 32 : (threads > SMW_THREADS ? SMW_THREADS : threads);
            if (split) dkdn_smem_kernel<true><<<(unsigned)((size_t)nbx * g->ny), threads, smem, s>>>(v, feat, prm, chi2_table, dk, dn,
                                                                                               minmax ? mm : nullptr, err, cb, nbx, cap);
            else dkdn_smem_kernel<false><<<(unsigned)((size_t)nbx * g->ny), threads, smem, s>>>(v, feat, prm, chi2_table, dk, dn,
                                                                                            minmax ? mm : nullptr, err, cb, nbx, cap);
        } else if (variant == 0 && split && M3D_SY(prm.s) <= 15 &&
                   act_interval(prm.rho2, prm.inv_rho2, &prm.act_lo, &prm.act_hi)) {
            // |n_p . n_q - quantised| <= sqrt(3) * 0.5 / 511 (quantisation) + float32 evaluation < 1.8e-3:
            // quantised > 0.984 + 1.8e-3  =>  1 - n_p . n_q < 0.016 with a 1e-4 margin over the fp64 rounding
            prm.dot_thr = 0.98581f;
            uint32_t* list = nullptr;
            unsigned* list_count = nullptr;
            unsigned long long* aux = nullptr;
            M3D_TRY(sc.alloc(&list, (size_t)g->n));
            M3D_TRY(sc.alloc(&aux, (size_t)g->n));
            M3D_TRY(sc.alloc(&list_count, 2));
            M3D_CUDA(cudaMemsetAsync(list_count, 0, 2 * sizeof(unsigned), s));
            const unsigned gate_above = (unsigned)fmin(g_dkdn_gate * (double)g->n, 4.0e9);
            static int grid_k = 0, grid_n = 0;
            if (!grid_k) grid_k = persistent_grid(dk_refine_kernel<REFINE_G>, 3);
            if (!grid_n) grid_n = persistent_grid(dkdn_refine_kernel<REFINE_G>, 2);
            {
                ProfScope p1("dkdn_screen", s);
                dkdn_screen_kernel<<<nb, WALK_THREADS, 0, s>>>(v, feat, prm, chi2_table, dk, dn, minmax ? mm : nullptr, err,
                                                               list, aux, list_count);
            }
            {
                ProfScope p2("dkdn_refine", s);
                static int refine_g = -1;
                if (refine_g < 0) { const char* e = getenv("MORE3D_REFINE_G"); refine_g = e ? atoi(e) : REFINE_G; }
                if (refine_g == 1)
                    dk_refine_thread_kernel<<<148 * 8, WALK_THREADS, 0, s>>>(v, feat, prm, chi2_table, dk, dn,
                                                                             minmax ? mm : nullptr, list, aux, list_count, gate_above);
                else if (refine_g == 2)
                    dk_refine_kernel<2><<<grid_k, WALK_THREADS, 0, s>>>(v, feat, prm, chi2_table, dk, dn,
                                                                        minmax ? mm : nullptr, err, list, aux, list_count, gate_above);
                else if (refine_g == 4)
                    dk_refine_kernel<4><<<grid_k, WALK_THREADS, 0, s>>>(v, feat, prm, chi2_table, dk, dn,
                                                                        minmax ? mm : nullptr, err, list, aux, list_count, gate_above);
                else
                    dk_refine_kernel<8><<<grid_k, WALK_THREADS, 0, s>>>(v, feat, prm, chi2_table, dk, dn,
                                                                        minmax ? mm : nullptr, err, list, aux, list_count, gate_above);
                dkdn_refine_kernel<REFINE_G><<<grid_n, WALK_THREADS, 0, s>>>(v, feat, prm, chi2_table, dk, dn,
                                                                            minmax ? mm : nullptr, err, list, list_count, gate_above);
                // the fallback (returns at once unless the lists are long): every point through the one-pass kernel
                dkdn_fallback_kernel<<<148 * 6, WALK_THREADS, 0, s>>>(v, feat, prm, chi2_table, dk, dn, minmax ? mm : nullptr,
                                                                      err, list_count, gate_above);
            }
            count_launches(3);
            if (getenv("MORE3D_DKDN_STATS")) {       // debugging aid: synchronises
                unsigned cnt[2] = {0, 0};
                cudaMemcpyAsync(cnt, list_count, 2 * sizeof(unsigned), cudaMemcpyDeviceToHost, s);
                cudaStreamSynchronize(s);
                fprintf(stderr, "[dkdn] r=%g n=%lld refined: dk only %u (%.2f%%), whole expression %u (%.3f%%)\n", r,
                        (long long)g->n, cnt[0], 100.0 * cnt[0] / (double)g->n, cnt[1], 100.0 * cnt[1] / (double)g->n);
            }
        } else if (split) {
            dkdn_kernel<false, true><<<nb, 256, 0, s>>>(v, feat, prm, chi2_table, dk, dn, minmax ? mm : nullptr, err);
        } else {
            dkdn_kernel<false, false><<<nb, 256, 0, s>>>(v, feat, prm, chi2_table, dk, dn, minmax ? mm : nullptr, err);
        }
    }
    if (minmax) minmax_finish<<<1, 32, 0, s>>>(mm, minmax);
    count_launches(minmax ? 4 : 3);
    M3D_CHECK_LAUNCH();
    int err_h = 0;
    if (!status) {          // no status word: the call itself reports the overflow (one stream synchronisation)
        M3D_CUDA(cudaMemcpyAsync(&err_h, err, sizeof(int), cudaMemcpyDeviceToHost, s));
        M3D_CUDA(cudaStreamSynchronize(s));
    }
    if (err_h) {
        set_error("dk/dn: a neighbourhood has %d active neighbours, chi2_table covers %d (table_len)", err_h - 1, table_len - 1);
        return MORE3D_E_RANGE;
    }
    return MORE3D_OK;
}

extern "C" int more3d_dk_dn(const more3d_grid* g, double r, const double* normals, const float* K, double rho,
                            double sigma_normal, double sigma_curvature, const double* chi2_table,
                            int table_len, int min_points, float* dk, float* dn, float* minmax, int32_t* status,
                            void* stream) {
    return run_dk_dn(g, r, normals, K, rho, sigma_normal, sigma_curvature, chi2_table, table_len, min_points,
                     dk, dn, minmax, status, stream, 0, 0, 0);
}

// Grid-order variants: every per-point output is written at the point's position in the grid's sorted order
// (more3d_grid_order), i.e. fully coalesced.  more3d_grid_unpermute brings a column back to the caller's order
// when -- and only if -- it is asked for.
extern "C" int more3d_normals_curvature_gridorder(const more3d_grid* g, double r, double eps, int min_points,
                                                  double* normals, double* lam_ratio, float* K, int32_t* pcount,
                                                  void* stream) {
    M3D_REQUIRE(K != nullptr && pcount != nullptr, "K / pcount is NULL");
    return launch_moments(g, MODE_FUSED, r, eps, min_points, normals, lam_ratio, K, pcount, (cudaStream_t)stream, 1);
}

extern "C" int more3d_dk_dn_gridorder(const more3d_grid* g, double r, double rho, double sigma_normal,
                                      double sigma_curvature, const double* chi2_table, int table_len,
                                      int min_points, int64_t own_lo, int64_t n_owned, float* dk, float* dn,
                                      float* minmax, int32_t* status, void* stream) {
    return run_dk_dn(g, r, nullptr, nullptr, rho, sigma_normal, sigma_curvature, chi2_table, table_len, min_points,
                     dk, dn, minmax, status, stream, 1, own_lo, n_owned);
}
