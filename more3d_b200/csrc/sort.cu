// Hand-written stable LSD radix sort (8-bit digits) and exclusive scans for the grid build,
// voxel keys and CSR offsets.  sm_100a; no CUB on the product path.
#include <stdarg.h>
#include <string.h>
#include <atomic>
#include <mutex>
#include <string>
#include <vector>
#include "common.cuh"

namespace m3d {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* get_error() { return g_err; }

static std::atomic<long long> g_launches{0};
void count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
long long get_launches() { return g_launches.load(); }

void init_device_pool() {
    static std::mutex mu;
    static bool done[64] = {false};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return;
    std::lock_guard<std::mutex> lk(mu);
    if (done[dev]) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        uint64_t thr = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    done[dev] = true;
}

// ---- profiler ---------------------------------------------------------------------------------
struct ProfRec { int name; cudaEvent_t a, b; };
static std::mutex g_prof_mu;
static bool g_prof_on = false;
static std::vector<std::string> g_prof_names;
static std::vector<ProfRec> g_prof_recs;
static std::vector<double> g_prof_ms;
static std::vector<int64_t> g_prof_cnt;

ProfScope::ProfScope(const char* name, cudaStream_t stream) : slot(-1), s(stream) {
    if (!g_prof_on) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    int id = -1;
    for (size_t i = 0; i < g_prof_names.size(); ++i)
        if (g_prof_names[i] == name) { id = (int)i; break; }
    if (id < 0) {
        id = (int)g_prof_names.size();
        g_prof_names.push_back(name);
        g_prof_ms.push_back(0.0);
        g_prof_cnt.push_back(0);
    }
    ProfRec r;
    r.name = id;
    cudaEventCreate(&r.a);
    cudaEventCreate(&r.b);
    cudaEventRecord(r.a, s);
    slot = (int)g_prof_recs.size();
    g_prof_recs.push_back(r);
}
ProfScope::~ProfScope() {
    if (slot < 0) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    cudaEventRecord(g_prof_recs[slot].b, s);
}
static void prof_collect() {
    for (auto& r : g_prof_recs) {
        float ms = 0.f;
        cudaEventSynchronize(r.b);
        if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
            g_prof_ms[r.name] += ms;
            g_prof_cnt[r.name] += 1;
        }
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    g_prof_recs.clear();
}

// ------------------------------------------------------------------------------------------
// exclusive scan: tile sums -> single-block scan of the sums -> tile scan + offset
// ------------------------------------------------------------------------------------------
constexpr int SC_THREADS = 256;
constexpr int SC_ITEMS = 8;
constexpr int SC_TILE = SC_THREADS * SC_ITEMS;

template <typename T>
__device__ __forceinline__ T block_exclusive_scan(T v, T* total, T* smem /* 32 entries */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        T t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) smem[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        T w = (lane < (int)(blockDim.x >> 5)) ? smem[lane] : T(0);
        T winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            T t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        smem[lane] = winc - w;  // exclusive warp offsets
        if (lane == 31) smem[32] = winc;
    }
    __syncthreads();
    T res = smem[warp] + inc - v;
    if (total) *total = smem[32];
    __syncthreads();
    return res;
}

template <typename Tin, typename Tout>
__global__ void __launch_bounds__(SC_THREADS) scan_tile_sums(const Tin* __restrict__ in, int64_t n,
                                                             Tout* __restrict__ sums) {
    __shared__ Tout sm[33];
    int64_t base = (int64_t)blockIdx.x * SC_TILE + (int64_t)threadIdx.x * SC_ITEMS;
    Tout s = 0;
#pragma unroll
    for (int i = 0; i < SC_ITEMS; ++i)
        if (base + i < n) s += (Tout)in[base + i];
    Tout total;
    block_exclusive_scan<Tout>(s, &total, sm);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

template <typename Tout>
__global__ void __launch_bounds__(1024) scan_sums_single(Tout* __restrict__ sums, int64_t m,
                                                         Tout* __restrict__ grand_total) {
    __shared__ Tout sm[33];
    Tout carry = 0;
    for (int64_t base = 0; base < m; base += 1024) {
        int64_t i = base + threadIdx.x;
        Tout v = (i < m) ? sums[i] : Tout(0);
        Tout total;
        Tout ex = block_exclusive_scan<Tout>(v, &total, sm);
        if (i < m) sums[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) *grand_total = carry;
}

template <typename Tin, typename Tout>
__global__ void __launch_bounds__(SC_THREADS) scan_tile_apply(const Tin* in, int64_t n,
                                                              const Tout* __restrict__ sums, Tout* out) {
    __shared__ Tout sm[33];
    int64_t base = (int64_t)blockIdx.x * SC_TILE + (int64_t)threadIdx.x * SC_ITEMS;
    Tout v[SC_ITEMS];
    Tout s = 0;
#pragma unroll
    for (int i = 0; i < SC_ITEMS; ++i) {
        v[i] = (base + i < n) ? (Tout)in[base + i] : Tout(0);
        s += v[i];
    }
    Tout ex = block_exclusive_scan<Tout>(s, nullptr, sm) + sums[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SC_ITEMS; ++i) {
        if (base + i < n) out[base + i] = ex;
        ex += v[i];
    }
}

template <typename Tin, typename Tout>
static int exclusive_scan_impl(const Tin* in, Tout* out, int64_t n, cudaStream_t s) {
    if (n <= 0) {
        M3D_CUDA(cudaMemsetAsync(out, 0, sizeof(Tout), s));
        return MORE3D_OK;
    }
    int64_t tiles = (n + SC_TILE - 1) / SC_TILE;
    Tout* sums = nullptr;
    M3D_TRY(dev_alloc(&sums, (size_t)tiles, s));
    scan_tile_sums<Tin, Tout><<<(unsigned)tiles, SC_THREADS, 0, s>>>(in, n, sums);
    scan_sums_single<Tout><<<1, 1024, 0, s>>>(sums, tiles, out + n);
    scan_tile_apply<Tin, Tout><<<(unsigned)tiles, SC_THREADS, 0, s>>>(in, n, sums, out);
    count_launches(3);
    M3D_CHECK_LAUNCH();
    dev_free(sums, s);
    return MORE3D_OK;
}

int exclusive_scan_u32(const uint32_t* in, uint32_t* out, int64_t n, cudaStream_t s) {
    return exclusive_scan_impl<uint32_t, uint32_t>(in, out, n, s);
}
int exclusive_scan_i32_to_i64(const int32_t* in, int64_t* out, int64_t n, cudaStream_t s) {
    return exclusive_scan_impl<int32_t, int64_t>(in, out, n, s);
}

// ------------------------------------------------------------------------------------------
// radix sort
// ------------------------------------------------------------------------------------------
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 8;                     // keys per thread (8: 4 CTAs per SM; the kernel is latency-bound, 16 was slower)
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;  // 2048 keys per CTA
constexpr int RS_WTILE = RS_TILE / RS_WARPS;    // 256 consecutive keys per warp

template <typename K>
__global__ void __launch_bounds__(RS_THREADS) rs_hist(const K* __restrict__ keys, int64_t n, int shift,
                                                      uint32_t* __restrict__ hist, int nblocks) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * RS_TILE;
#pragma unroll 4
    for (int i = threadIdx.x; i < RS_TILE; i += RS_THREADS) {
        int64_t p = base + i;
        if (p < n) atomicAdd(&h[(unsigned)(keys[p] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];  // digit-major
}

// Stable scatter: warp w owns RS_WTILE consecutive keys of the tile, walks them in RS_ITEMS rounds of 32;
// __match_any_sync gives each key its rank among equal digits of the round, per-warp running counters give the
// rank among the earlier keys of the tile.  The tile is first sorted by digit INTO SHARED MEMORY and then
// written out linearly: consecutive threads hold consecutive keys of one digit run, so the global stores are
// contiguous runs (avg 16 keys = 64 B per digit and tile) instead of 4-byte scatters to 256 places.
template <typename K>
__global__ void __launch_bounds__(RS_THREADS, 4) rs_scatter(const K* __restrict__ keys,
                                                         const uint32_t* __restrict__ vals,
                                                         K* __restrict__ keys_out,
                                                         uint32_t* __restrict__ vals_out, int64_t n,
                                                         int shift, const uint32_t* __restrict__ hist_ex,
                                                         int nblocks) {
    extern __shared__ __align__(16) unsigned char rs_smem[];
    K* keys_s = reinterpret_cast<K*>(rs_smem);                                   // RS_TILE keys
    uint32_t* vals_s = reinterpret_cast<uint32_t*>(keys_s + RS_TILE);             // RS_TILE values
    uint32_t (*wcnt)[256] = reinterpret_cast<uint32_t (*)[256]>(vals_s + RS_TILE);   // RS_WARPS x 256
    uint32_t* tile_start = &wcnt[0][0] + RS_WARPS * 256;                          // 256: first tile slot of a digit
    uint32_t* gbase = tile_start + 256;                                           // 256: first global slot
    uint32_t* scan_sm = gbase + 256;                                              // 33
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&wcnt[0][0])[i] = 0;
    __syncthreads();

    const int64_t tbase = (int64_t)blockIdx.x * RS_TILE;
    const int64_t wbase = tbase + (int64_t)warp * RS_WTILE;
    K k[RS_ITEMS];
    uint32_t v[RS_ITEMS];
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) {
        int64_t p = wbase + j * 32 + lane;
        bool ok = p < n;
        k[j] = ok ? keys[p] : K(0);
        v[j] = ok ? vals[p] : 0u;
    }
    // digits and their match masks once for both phases; the RS_ITEMS matches are independent of each other, so
    // they pipeline instead of each waiting behind the counter update of the previous round
    unsigned dg[RS_ITEMS], pm[RS_ITEMS];
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) {
        const bool ok = (wbase + j * 32 + lane) < n;
        dg[j] = ok ? ((unsigned)(k[j] >> shift) & 255u) : 256u;
        pm[j] = __match_any_sync(0xffffffffu, dg[j]);
    }
    // phase A: per-warp digit counts; every key also takes its rank among the warp's EARLIER keys of its digit
    // (running count before this round + equal-digit lanes below it), so phase B needs no second serial pass
    uint32_t rk[RS_ITEMS];
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) {
        const bool ok = dg[j] < 256u;
        rk[j] = ok ? wcnt[warp][dg[j]] + __popc(pm[j] & lt) : 0u;
        __syncwarp();
        if (ok && (pm[j] & lt) == 0u) wcnt[warp][dg[j]] += __popc(pm[j]);
        __syncwarp();
    }
    __syncthreads();
    {
        const int d = threadIdx.x;  // RS_THREADS == 256 digits
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) {
            uint32_t c = wcnt[w][d];
            wcnt[w][d] = run;           // keys of digit d in lower warps
            run += c;
        }
        const uint32_t ts = block_exclusive_scan<uint32_t>(run, nullptr, scan_sm);   // digits below d in the tile
        tile_start[d] = ts;
        gbase[d] = hist_ex[(size_t)d * nblocks + blockIdx.x];
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) wcnt[w][d] += ts;
    }
    __syncthreads();
    // phase B: keys and values go to their sorted slot of the tile
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) {
        if (dg[j] < 256u) {
            const uint32_t pos = wcnt[warp][dg[j]] + rk[j];      // tile slot of the digit run + warps below + rank
            keys_s[pos] = k[j];
            vals_s[pos] = v[j];
        }
    }
    __syncthreads();
    // phase C: linear write-out
    const int tile_n = (int)min((int64_t)RS_TILE, n - tbase);
    for (int i = threadIdx.x; i < tile_n; i += RS_THREADS) {
        const K kk = keys_s[i];
        const unsigned d = (unsigned)(kk >> shift) & 255u;
        const uint32_t g = gbase[d] + ((uint32_t)i - tile_start[d]);
        keys_out[g] = kk;
        vals_out[g] = vals_s[i];
    }
}

template <typename K>
static size_t rs_scatter_smem() {
    return (size_t)RS_TILE * (sizeof(K) + sizeof(uint32_t)) + (size_t)(RS_WARPS * 256 + 256 + 256 + 33) * sizeof(uint32_t);
}

template <typename K>
static int radix_sort_impl(K* keys, uint32_t* vals, K* keys_alt, uint32_t* vals_alt, int64_t n,
                           int key_bits, K** keys_sorted, uint32_t** vals_sorted, cudaStream_t s) {
    *keys_sorted = keys;
    *vals_sorted = vals;
    if (n <= 1 || key_bits <= 0) return MORE3D_OK;
    if (n >= (int64_t)1 << 32) {
        set_error("radix sort: n >= 2^32 not supported");
        return MORE3D_E_RANGE;
    }
    const int nblocks = (int)((n + RS_TILE - 1) / RS_TILE);
    const size_t smem = rs_scatter_smem<K>();
    M3D_CUDA(cudaFuncSetAttribute(rs_scatter<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    uint32_t* hist = nullptr;
    M3D_TRY(dev_alloc(&hist, (size_t)256 * nblocks + 1, s));
    K* kin = keys;
    K* kout = keys_alt;
    uint32_t* vin = vals;
    uint32_t* vout = vals_alt;
    for (int shift = 0; shift < key_bits; shift += 8) {
        rs_hist<K><<<nblocks, RS_THREADS, 0, s>>>(kin, n, shift, hist, nblocks);
        M3D_TRY(exclusive_scan_u32(hist, hist, (int64_t)256 * nblocks, s));
        rs_scatter<K><<<nblocks, RS_THREADS, smem, s>>>(kin, vin, kout, vout, n, shift, hist, nblocks);
        count_launches(2);
        M3D_CHECK_LAUNCH();
        K* tk = kin; kin = kout; kout = tk;
        uint32_t* tv = vin; vin = vout; vout = tv;
    }
    dev_free(hist, s);
    *keys_sorted = kin;
    *vals_sorted = vin;
    return MORE3D_OK;
}

int radix_sort_pairs_u32(uint32_t* keys, uint32_t* vals, uint32_t* keys_alt, uint32_t* vals_alt,
                         int64_t n, int key_bits, uint32_t** keys_sorted, uint32_t** vals_sorted,
                         cudaStream_t s) {
    return radix_sort_impl<uint32_t>(keys, vals, keys_alt, vals_alt, n, key_bits, keys_sorted, vals_sorted, s);
}
int radix_sort_pairs_u64(uint64_t* keys, uint32_t* vals, uint64_t* keys_alt, uint32_t* vals_alt,
                         int64_t n, int key_bits, uint64_t** keys_sorted, uint32_t** vals_sorted,
                         cudaStream_t s) {
    return radix_sort_impl<uint64_t>(keys, vals, keys_alt, vals_alt, n, key_bits, keys_sorted, vals_sorted, s);
}

}  // namespace m3d

extern "C" int64_t more3d_launch_count(void) { return (int64_t)m3d::get_launches(); }
extern "C" int more3d_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(m3d::g_prof_mu);
    m3d::g_prof_on = on != 0;
    return MORE3D_OK;
}
extern "C" int more3d_profile_reset(void) {
    std::lock_guard<std::mutex> lk(m3d::g_prof_mu);
    m3d::prof_collect();
    for (auto& v : m3d::g_prof_ms) v = 0.0;
    for (auto& v : m3d::g_prof_cnt) v = 0;
    return MORE3D_OK;
}
extern "C" int more3d_profile_get(const char* name, double* ms_h, int64_t* count_h) {
    std::lock_guard<std::mutex> lk(m3d::g_prof_mu);
    m3d::prof_collect();
    for (size_t i = 0; i < m3d::g_prof_names.size(); ++i)
        if (m3d::g_prof_names[i] == name) {
            if (ms_h) *ms_h = m3d::g_prof_ms[i];
            if (count_h) *count_h = m3d::g_prof_cnt[i];
            return MORE3D_OK;
        }
    if (ms_h) *ms_h = 0.0;
    if (count_h) *count_h = 0;
    return MORE3D_OK;
}
extern "C" int more3d_profile_names(char* buf_h, int len) {
    std::lock_guard<std::mutex> lk(m3d::g_prof_mu);
    std::string all;
    for (auto& n : m3d::g_prof_names) { if (!all.empty()) all += ","; all += n; }
    if (!buf_h || len <= 0) return MORE3D_E_INVALID;
    strncpy(buf_h, all.c_str(), (size_t)len - 1);
    buf_h[len - 1] = 0;
    return MORE3D_OK;
}
extern "C" const char* more3d_last_error(void) { return m3d::get_error(); }
extern "C" int more3d_version(void) { return 100; }
