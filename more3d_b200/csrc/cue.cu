// Cue pre-processing of the level-set driver on the device ("next" row f-3 of SURVEY.md section 8):
//   solver.zeta = np.abs(solver.zeta); clip(zeta, 0, np.percentile(zeta, pc)); normalize(zeta, 1)
// (run_levelsets.py:215-221; levelsets_func.py:100-109).  Everything is exact: the percentile comes from the two
// neighbouring ORDER STATISTICS found by an 8-pass most-significant-digit radix select on order-preserving 64-bit
// keys (no sort, 8 reads of the column), combined with numpy's `linear` interpolation rule; clip and normalize are
// the reference's elementwise expressions with their separate roundings.
#include "common.cuh"

namespace m3d {

constexpr int SEL_MAXK = 4;

struct SelState {
    unsigned long long prefix[SEL_MAXK];   // the high digits fixed so far (aligned at their final bit positions)
    long long rank[SEL_MAXK];              // rank still to find among the elements that match the prefix
    unsigned hist[SEL_MAXK][256];
};

__global__ void sel_init(SelState* st, const long long* ranks, int nk) {
    if ((int)threadIdx.x < nk) { st->prefix[threadIdx.x] = 0ull; st->rank[threadIdx.x] = ranks[threadIdx.x]; }
    for (int i = threadIdx.x; i < SEL_MAXK * 256; i += blockDim.x) (&st->hist[0][0])[i] = 0u;
}

// pass p fixes byte (7 - p): histogram of that byte over the elements whose higher bytes equal the prefix
__global__ void __launch_bounds__(256) sel_hist(const double* __restrict__ v, int64_t n, int pass, int nk, SelState* st) {
    __shared__ unsigned h[SEL_MAXK][256];
    for (int i = threadIdx.x; i < SEL_MAXK * 256; i += 256) (&h[0][0])[i] = 0u;
    __syncthreads();
    const int shift = 8 * (7 - pass);
    const unsigned long long himask = pass == 0 ? 0ull : (~0ull << (shift + 8));
    unsigned long long pre[SEL_MAXK];
    for (int k = 0; k < nk; ++k) pre[k] = st->prefix[k];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long key = sortable_key(v[i]);
        const unsigned d = (unsigned)(key >> shift) & 255u;
        for (int k = 0; k < nk; ++k)
            if ((key & himask) == pre[k]) atomicAdd(&h[k][d], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nk * 256; i += 256) {
        const unsigned c = (&h[0][0])[i];
        if (c) atomicAdd(&st->hist[0][0] + i, c);
    }
}

// one warp per rank query: walk the 256 bins to the one that holds the rank, fix that digit, clear the histogram
__global__ void sel_pick(SelState* st, int pass, int nk) {
    const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (k >= nk) return;
    const int shift = 8 * (7 - pass);
    long long rank = st->rank[k];
    if (lane == 0) {
        unsigned long long run = 0;
        int d = 0;
        for (; d < 255; ++d) {
            const unsigned long long c = st->hist[k][d];
            if ((unsigned long long)rank < run + c) break;
            run += c;
        }
        st->rank[k] = rank - (long long)run;
        st->prefix[k] |= (unsigned long long)d << shift;
    }
    __syncwarp();
    for (int i = lane; i < 256; i += 32) st->hist[k][i] = 0u;
}

__global__ void sel_finish(const SelState* st, int nk, double* out) {
    if ((int)threadIdx.x < nk) {
        const unsigned long long key = st->prefix[threadIdx.x];
        const unsigned long long u = (key >> 63) ? (key & 0x7fffffffffffffffull) : ~key;      // inverse of sortable_key
        out[threadIdx.x] = __longlong_as_double((long long)u);
    }
}

// numpy's _lerp (lib/_function_base_impl.py): a + (b - a) * t, or b - (b - a) * (1 - t) when t >= 0.5
__global__ void lerp_kernel(const double* ab, double t, double lo_fixed, double* lohi_out) {
    const double a = ab[0], b = ab[1];
    const double diff = __dsub_rn(b, a);
    double r = __dadd_rn(a, __dmul_rn(diff, t));
    if (t >= 0.5) r = __dsub_rn(b, __dmul_rn(diff, __dsub_rn(1.0, t)));
    if (t == 0.0) r = a;                       // numpy takes the lower statistic as is when gamma is 0
    lohi_out[0] = lo_fixed;
    lohi_out[1] = r;
}

__global__ void __launch_bounds__(256) abs_kernel(double* x, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) x[i] = fabs(x[i]);
}
// array[array < a] = a; array[array > b] = b   (levelsets_func.py:107-109; NaN stays)
__global__ void __launch_bounds__(256) clip_kernel(double* x, int64_t n, const double* __restrict__ lohi) {
    const double a = lohi[0], b = lohi[1];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double v = x[i];
        if (v < a) v = a;
        if (v > b) v = b;
        x[i] = v;
    }
}
__global__ void __launch_bounds__(256) minmax_f64_partial(const double* __restrict__ x, int64_t n, double* __restrict__ part) {
    double lo = INFINITY, hi = -INFINITY;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = x[i];
        lo = fmin(lo, v); hi = fmax(hi, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    __shared__ double sl[8], sh[8];
    if ((threadIdx.x & 31) == 0) { sl[threadIdx.x >> 5] = lo; sh[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) { lo = fmin(lo, sl[w]); hi = fmax(hi, sh[w]); }
        part[2 * blockIdx.x] = lo; part[2 * blockIdx.x + 1] = hi;
    }
}
__global__ void minmax_f64_final(const double* __restrict__ part, int nparts, double* __restrict__ out) {
    double lo = INFINITY, hi = -INFINITY;
    for (int i = threadIdx.x; i < nparts; i += 32) { lo = fmin(lo, part[2 * i]); hi = fmax(hi, part[2 * i + 1]); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if (threadIdx.x == 0) { out[0] = lo; out[1] = hi; }
}
__global__ void __launch_bounds__(256) sum_f64_partial(const double* __restrict__ x, int64_t n, double* __restrict__ part) {
    double sum = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) sum += x[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    __shared__ double ss[8];
    if ((threadIdx.x & 31) == 0) ss[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) sum += ss[w];
        part[blockIdx.x] = sum;
    }
}
__global__ void sum_f64_final(const double* __restrict__ part, int nparts, double scale, double* __restrict__ out) {
    double sum = 0.0;
    for (int i = threadIdx.x; i < nparts; i += 32) sum += part[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (threadIdx.x == 0) out[0] = sum * scale;
}
// if np.any(array != 0): array -= array.min(); array /= array.max(); array *= s   (levelsets_func.py:100-104)
__global__ void __launch_bounds__(256) normalize_kernel(double* x, int64_t n, const double* __restrict__ mm, double s) {
    const double mn = mm[0], mx = mm[1];
    if (mn == 0.0 && mx == 0.0) return;                 // all zero: left as it is
    const double top = __dsub_rn(mx, mn);               // the maximum of the shifted array
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        x[i] = __dmul_rn(__ddiv_rn(__dsub_rn(x[i], mn), top), s);
}

// {min, max, sum} partials, then {min, max, mean}; second pass: sum of squared deviations -> population sigma.
// Fixed 592-block tree: reproducible run to run.
__global__ void __launch_bounds__(256) stats_partial(const float* __restrict__ x, int64_t n, const double* __restrict__ mean,
                                                     double* __restrict__ part) {
    double lo = INFINITY, hi = -INFINITY, sum = 0.0;
    const double mu = mean ? mean[2] : 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = (double)x[i];
        if (mean) { const double d = v - mu; sum += d * d; }
        else { lo = fmin(lo, v); hi = fmax(hi, v); sum += v; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
    }
    __shared__ double sl[8], sh[8], ss[8];
    if ((threadIdx.x & 31) == 0) { sl[threadIdx.x >> 5] = lo; sh[threadIdx.x >> 5] = hi; ss[threadIdx.x >> 5] = sum; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) { lo = fmin(lo, sl[w]); hi = fmax(hi, sh[w]); sum += ss[w]; }
        part[3 * blockIdx.x] = lo; part[3 * blockIdx.x + 1] = hi; part[3 * blockIdx.x + 2] = sum;
    }
}
__global__ void stats_final(const double* __restrict__ part, int nparts, int64_t n, int second, double* __restrict__ out) {
    double lo = INFINITY, hi = -INFINITY, sum = 0.0;
    for (int i = threadIdx.x; i < nparts; i += 32) { lo = fmin(lo, part[3 * i]); hi = fmax(hi, part[3 * i + 1]); sum += part[3 * i + 2]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
    }
    if (threadIdx.x == 0) {
        if (second) out[3] = sqrt(sum / (double)n);
        else { out[0] = lo; out[1] = hi; out[2] = sum / (double)n; }
    }
}

}  // namespace m3d

using namespace m3d;

/* out (device, 4 f64) = {min, max, mean, population sigma} of an f32 column (DM_saliency_stats, DM_saliency.py:585-607) */
extern "C" int more3d_column_stats(const float* x, int64_t n, double* out, void* stream) {
    M3D_REQUIRE(x && out && n > 0, "NULL / empty argument");
    cudaStream_t s = (cudaStream_t)stream;
    Scratch sc(s);
    double* part = nullptr;
    M3D_TRY(sc.alloc(&part, (size_t)3 * 592));
    stats_partial<<<592, 256, 0, s>>>(x, n, nullptr, part);
    stats_final<<<1, 32, 0, s>>>(part, 592, n, 0, out);
    stats_partial<<<592, 256, 0, s>>>(x, n, out, part);
    stats_final<<<1, 32, 0, s>>>(part, 592, n, 1, out);
    count_launches(4);
    M3D_CHECK_LAUNCH();
    return MORE3D_OK;
}

extern "C" int more3d_select_kth(const double* v, int64_t n, const int64_t* ranks_h, int nk, double* out, void* stream) {
    M3D_REQUIRE(v && ranks_h && out && n > 0, "NULL / empty argument");
    M3D_REQUIRE(nk >= 1 && nk <= SEL_MAXK, "1 to 4 ranks per call");
    for (int k = 0; k < nk; ++k) M3D_REQUIRE(ranks_h[k] >= 0 && ranks_h[k] < n, "rank out of range");
    cudaStream_t s = (cudaStream_t)stream;
    init_device_pool();
    ProfScope prof("select_kth", s);
    SelState* st = nullptr;
    long long* dr = nullptr;
    M3D_TRY(dev_alloc(&st, 1, s));
    M3D_TRY(dev_alloc(&dr, SEL_MAXK, s));
    long long hr[SEL_MAXK] = {0, 0, 0, 0};
    for (int k = 0; k < nk; ++k) hr[k] = ranks_h[k];
    // pageable source: the copy is staged by the runtime before the call returns
    M3D_CUDA(cudaMemcpyAsync(dr, hr, sizeof(hr), cudaMemcpyHostToDevice, s));
    sel_init<<<1, 256, 0, s>>>(st, dr, nk);
    for (int p = 0; p < 8; ++p) {
        sel_hist<<<148 * 8, 256, 0, s>>>(v, n, p, nk, st);
        sel_pick<<<1, 32 * SEL_MAXK, 0, s>>>(st, p, nk);
    }
    sel_finish<<<1, 32, 0, s>>>(st, nk, out);
    count_launches(18);
    M3D_CHECK_LAUNCH();
    dev_free(st, s); dev_free(dr, s);
    return MORE3D_OK;
}

extern "C" int more3d_cue_ops(int op, double* x, int64_t n, const double* arg_dev, double arg, double* out_dev, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned nb = 148 * 8;
    switch (op) {
    case 0:   // abs in place
        M3D_REQUIRE(x && n > 0, "NULL / empty argument");
        abs_kernel<<<nb, 256, 0, s>>>(x, n);
        count_launches(1);
        break;
    case 1:   // clip in place to [arg_dev[0], arg_dev[1]]
        M3D_REQUIRE(x && arg_dev && n > 0, "NULL / empty argument");
        clip_kernel<<<nb, 256, 0, s>>>(x, n, arg_dev);
        count_launches(1);
        break;
    case 2: { // normalize in place, scale = arg; out_dev (2 doubles, may be NULL) receives {min, max} before the shift
        M3D_REQUIRE(x && n > 0, "NULL / empty argument");
        double* part = nullptr;
        M3D_TRY(dev_alloc(&part, (size_t)2 * 592 + 2, s));
        minmax_f64_partial<<<592, 256, 0, s>>>(x, n, part);
        minmax_f64_final<<<1, 32, 0, s>>>(part, 592, part + 2 * 592);
        normalize_kernel<<<nb, 256, 0, s>>>(x, n, part + 2 * 592, arg);
        if (out_dev) M3D_CUDA(cudaMemcpyAsync(out_dev, part + 2 * 592, 2 * sizeof(double), cudaMemcpyDeviceToDevice, s));
        count_launches(3);
        dev_free(part, s);
        break;
    }
    case 3:   // out_dev[0..1] = {arg_dev[2] (fixed lower bound), numpy-lerp(arg_dev[0], arg_dev[1], arg)}
        M3D_REQUIRE(arg_dev && out_dev, "NULL argument");
        lerp_kernel<<<1, 1, 0, s>>>(arg_dev, arg, 0.0, out_dev);
        count_launches(1);
        break;
    case 4: { // out_dev[0..1] = {min, max} of x (x is only read)
        M3D_REQUIRE(x && out_dev && n > 0, "NULL / empty argument");
        double* part = nullptr;
        M3D_TRY(dev_alloc(&part, (size_t)2 * 592, s));
        minmax_f64_partial<<<592, 256, 0, s>>>(x, n, part);
        minmax_f64_final<<<1, 32, 0, s>>>(part, 592, out_dev);
        count_launches(2);
        dev_free(part, s);
        break;
    }
    case 5: { // out_dev[0] = mean of x (fixed 592-block tree: reproducible); x is only read
        M3D_REQUIRE(x && out_dev && n > 0, "NULL / empty argument");
        double* part = nullptr;
        M3D_TRY(dev_alloc(&part, (size_t)592, s));
        sum_f64_partial<<<592, 256, 0, s>>>(x, n, part);
        sum_f64_final<<<1, 32, 0, s>>>(part, 592, 1.0 / (double)n, out_dev);
        count_launches(2);
        dev_free(part, s);
        break;
    }
    default:
        M3D_REQUIRE(false, "op must be 0 (abs), 1 (clip), 2 (normalize), 3 (percentile interpolation), 4 (min / max) or 5 (mean)");
    }
    M3D_CHECK_LAUNCH();
    return MORE3D_OK;
}
