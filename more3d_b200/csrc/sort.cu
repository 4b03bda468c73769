// Hand-written stable LSD radix sort (8-bit digits) and exclusive scans for the grid build,
// voxel keys and CSR offsets.  sm_100a; no CUB on the product path.
#include <stdarg.h>
#include <string.h>
#include <atomic>
#include <mutex>
#include <string>
#include <vector>
#include "common.cuh"
#include "scan.cuh"

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

// ---- small read-backs without the copy engine ---------------------------------------------------
__global__ void store_to_host_kernel(const unsigned char* __restrict__ src, unsigned char* dst_mapped, unsigned bytes) {
    for (unsigned i = threadIdx.x; i < bytes; i += blockDim.x) dst_mapped[i] = src[i];
    __threadfence_system();
}

int read_back(void* dst_h, const void* src_dev, size_t bytes, cudaStream_t s) {
    constexpr size_t CAP = 4096;
    struct Slot { void* host = nullptr; void* dev = nullptr; cudaEvent_t ev = nullptr; int device = -1; };
    static thread_local Slot slot;
    if (bytes == 0) return MORE3D_OK;
    int dev = 0;
    M3D_CUDA(cudaGetDevice(&dev));
    if (bytes > CAP) {          // not a "small" read: the ordinary path
        M3D_CUDA(cudaMemcpyAsync(dst_h, src_dev, bytes, cudaMemcpyDeviceToHost, s));
        M3D_CUDA(cudaStreamSynchronize(s));
        return MORE3D_OK;
    }
    if (slot.host == nullptr || slot.device != dev) {
        if (slot.host) { cudaFreeHost(slot.host); cudaEventDestroy(slot.ev); slot.host = nullptr; }
        M3D_CUDA(cudaHostAlloc(&slot.host, CAP, cudaHostAllocMapped | cudaHostAllocPortable));
        M3D_CUDA(cudaHostGetDevicePointer(&slot.dev, slot.host, 0));
        M3D_CUDA(cudaEventCreateWithFlags(&slot.ev, cudaEventDisableTiming));
        slot.device = dev;
    }
    store_to_host_kernel<<<1, 128, 0, s>>>((const unsigned char*)src_dev, (unsigned char*)slot.dev, (unsigned)bytes);
    count_launches(1);
    M3D_CHECK_LAUNCH();
    M3D_CUDA(cudaEventRecord(slot.ev, s));
    M3D_CUDA(cudaEventSynchronize(slot.ev));
    memcpy(dst_h, slot.host, bytes);
    return MORE3D_OK;
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
// exclusive scan: ONE pass over the data (chained scan with decoupled look-back).  Every CTA takes a tile
// ticket, scans its 2048 elements, publishes its aggregate in a 64-bit descriptor (status in the two top bits,
// value below: one store makes both visible together) and adds up the descriptors of the tiles before it until
// it meets one that already carries an inclusive prefix.  Tickets are handed out in launch order, so a tile only
// ever waits for CTAs that are already running.
// ------------------------------------------------------------------------------------------
template <typename Tin, typename Tout>
__global__ void __launch_bounds__(SC_THREADS) scan_lookback(const Tin* in, Tout* out, int64_t n, uint64_t* desc,
                                                            unsigned* ticket, int ntiles) {
    __shared__ uint64_t sm[33];
    __shared__ uint64_t s_prefix;
    __shared__ unsigned s_tile;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const int tile = (int)s_tile;
    const int64_t base = (int64_t)tile * SC_TILE + (int64_t)threadIdx.x * SC_ITEMS;
    uint64_t v[SC_ITEMS];
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < SC_ITEMS; ++i) {
        v[i] = (base + i < n) ? (uint64_t)in[base + i] : 0ull;
        s += v[i];
    }
    uint64_t total;
    uint64_t ex = block_exclusive_scan<uint64_t>(s, &total, sm);
    volatile uint64_t* vd = desc;
    if (threadIdx.x == 0) {
        vd[tile] = (tile == 0 ? LB_PREFIX : LB_AGG) | total;
        if (tile == 0) s_prefix = 0;
    }
    if (tile > 0 && threadIdx.x < 32) {
        const uint64_t p = lookback_prefix(vd, tile, threadIdx.x);
        if (threadIdx.x == 0) {
            vd[tile] = LB_PREFIX | (p + total);
            s_prefix = p;
        }
    }
    __syncthreads();
    ex += s_prefix;
#pragma unroll
    for (int i = 0; i < SC_ITEMS; ++i) {
        if (base + i < n) out[base + i] = (Tout)ex;
        ex += v[i];
    }
    if (tile == ntiles - 1 && threadIdx.x == 0) out[n] = (Tout)(s_prefix + total);
}

template <typename Tin, typename Tout>
static int exclusive_scan_impl(const Tin* in, Tout* out, int64_t n, cudaStream_t s) {
    if (n <= 0) {
        M3D_CUDA(cudaMemsetAsync(out, 0, sizeof(Tout), s));
        return MORE3D_OK;
    }
    const int64_t tiles = (n + SC_TILE - 1) / SC_TILE;
    uint64_t* desc = nullptr;
    M3D_TRY(dev_alloc(&desc, (size_t)tiles + 1, s));
    M3D_CUDA(cudaMemsetAsync(desc, 0, ((size_t)tiles + 1) * sizeof(uint64_t), s));
    scan_lookback<Tin, Tout><<<(unsigned)tiles, SC_THREADS, 0, s>>>(in, out, n, desc, (unsigned*)(desc + tiles), (int)tiles);
    count_launches(1);
    M3D_CHECK_LAUNCH();
    dev_free(desc, s);
    return MORE3D_OK;
}

int exclusive_scan_u32(const uint32_t* in, uint32_t* out, int64_t n, cudaStream_t s) {
    return exclusive_scan_impl<uint32_t, uint32_t>(in, out, n, s);
}
int exclusive_scan_i32_to_i64(const int32_t* in, int64_t* out, int64_t n, cudaStream_t s) {
    return exclusive_scan_impl<int32_t, int64_t>(in, out, n, s);
}

// ------------------------------------------------------------------------------------------
// radix sort: stable LSD, 8-bit digits, ONE kernel per digit ("onesweep": the per-tile digit counts are chained
// through decoupled look-back descriptors instead of a histogram + scan + scatter triple), one up-front kernel
// that builds the global digit histograms of ALL passes in a single read of the keys.
// ------------------------------------------------------------------------------------------
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_MAX_PASSES = 8;
constexpr uint32_t RS_AGG = 1u << 30, RS_PREFIX = 2u << 30, RS_VAL = (1u << 30) - 1u;

template <typename K>
__global__ void __launch_bounds__(RS_THREADS) rs_hist_all(const K* __restrict__ keys, int64_t n, int npass,
                                                          uint32_t* __restrict__ hist /* npass x 256, zeroed */) {
    __shared__ uint32_t h[RS_MAX_PASSES][256];
    for (int i = threadIdx.x; i < npass * 256; i += RS_THREADS) (&h[0][0])[i] = 0;
    __syncthreads();
    const unsigned lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)gridDim.x * RS_THREADS;
    const int64_t n_round = ((n + 31) / 32) * 32;               // whole warps stay in the loop (match needs them)
    for (int64_t i = (int64_t)blockIdx.x * RS_THREADS + threadIdx.x; i < n_round; i += stride) {
        const bool ok = i < n;
        const K k = ok ? keys[i] : K(0);
        for (int p = 0; p < npass; ++p) {
            // warp-aggregated: equal digits of the warp's 32 keys cost one shared atomic (the high digits of grid
            // and voxel keys take a handful of values, a plain atomic per key would serialise on one bank)
            const unsigned d = (unsigned)(k >> (8 * p)) & 255u;
            const unsigned m = match_digit(d, ok);
            if (ok && (m & ((1u << lane) - 1u)) == 0u) atomicAdd(&h[p][d], (uint32_t)__popc(m));
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < npass * 256; i += RS_THREADS) {
        const uint32_t c = (&h[0][0])[i];
        if (c) atomicAdd(hist + i, c);
    }
}

// one CTA per pass: exclusive scan of its 256 bins (in place)
__global__ void __launch_bounds__(256) rs_scan_bins(uint32_t* __restrict__ hist) {
    __shared__ uint32_t sm[33];
    uint32_t* h = hist + 256 * blockIdx.x;
    const uint32_t c = h[threadIdx.x];
    h[threadIdx.x] = block_exclusive_scan<uint32_t>(c, nullptr, sm);
}

// One tile per CTA (ticket order).  Warp w owns ITEMS * 32 consecutive keys of the tile and walks them in ITEMS
// rounds of 32; __match_any_sync gives each key its rank among equal digits of the round, per-warp running counters
// its rank among the warp's earlier keys -- a stable rank.  Thread d then owns digit d: it publishes the tile's
// count of d, looks back over the earlier tiles' descriptors for the number of keys with digit d before this tile,
// and the tile -- sorted by digit in shared memory -- is written out as contiguous runs.
template <typename K, int ITEMS>
__global__ void __launch_bounds__(RS_THREADS) rs_onesweep(const K* __restrict__ keys, const uint32_t* __restrict__ vals,
                                                          K* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                                                          int64_t n, int shift, const uint32_t* __restrict__ bins,
                                                          uint32_t* desc, unsigned* ticket) {
    constexpr int TILE = RS_THREADS * ITEMS;
    constexpr int WTILE = 32 * ITEMS;
    extern __shared__ __align__(16) unsigned char rs_smem[];
    K* keys_s = reinterpret_cast<K*>(rs_smem);                                    // TILE keys
    uint32_t* vals_s = reinterpret_cast<uint32_t*>(keys_s + TILE);                 // TILE values
    uint32_t (*wcnt)[256] = reinterpret_cast<uint32_t (*)[256]>(vals_s + TILE);    // RS_WARPS x 256
    uint32_t* gbase = &wcnt[0][0] + RS_WARPS * 256;                                // 256: global slot of tile slot 0 of a digit run
    uint32_t* scan_sm = gbase + 256;                                               // 33
    __shared__ unsigned s_tile;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&wcnt[0][0])[i] = 0;
    __syncthreads();
    const int tile = (int)s_tile;

    const int64_t tbase = (int64_t)tile * TILE;
    const int64_t wbase = tbase + (int64_t)warp * WTILE;
    K k[ITEMS];
    uint32_t v[ITEMS];
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        const int64_t p = wbase + j * 32 + lane;
        const bool ok = p < n;
        k[j] = ok ? keys[p] : K(0);
        v[j] = ok ? vals[p] : 0u;
    }
    unsigned dg[ITEMS], pm[ITEMS];
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        const bool ok = (wbase + j * 32 + lane) < n;
        dg[j] = ok ? ((unsigned)(k[j] >> shift) & 255u) : 256u;
        pm[j] = match_digit(dg[j], ok);
    }
    uint32_t rk[ITEMS];
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
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
            const uint32_t c = wcnt[w][d];
            wcnt[w][d] = run;           // keys of digit d in lower warps
            run += c;
        }
        volatile uint32_t* vd = desc;
        vd[(size_t)tile * 256 + d] = (tile == 0 ? RS_PREFIX : RS_AGG) | run;
        const uint32_t ts = block_exclusive_scan<uint32_t>(run, nullptr, scan_sm);   // digits below d in the tile
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) wcnt[w][d] += ts;
        uint32_t excl = 0;
        for (int j = tile - 1; j >= 0; --j) {
            uint32_t e;
            do { e = vd[(size_t)j * 256 + d]; } while ((e >> 30) == 0u);
            excl += e & RS_VAL;
            if ((e >> 30) == 2u) break;
        }
        if (tile > 0) vd[(size_t)tile * 256 + d] = RS_PREFIX | (excl + run);
        gbase[d] = bins[d] + excl - ts;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        if (dg[j] < 256u) {
            const uint32_t pos = wcnt[warp][dg[j]] + rk[j];      // tile slot of the digit run + warps below + rank
            keys_s[pos] = k[j];
            vals_s[pos] = v[j];
        }
    }
    __syncthreads();
    const int tile_n = (int)min((int64_t)TILE, n - tbase);
    for (int i = threadIdx.x; i < tile_n; i += RS_THREADS) {
        const K kk = keys_s[i];
        const unsigned d = (unsigned)(kk >> shift) & 255u;
        const uint32_t g = gbase[d] + (uint32_t)i;
        keys_out[g] = kk;
        vals_out[g] = vals_s[i];
    }
}

template <typename K> struct RsItems { static constexpr int value = 16; };
template <> struct RsItems<uint64_t> { static constexpr int value = 12; };

template <typename K>
static int radix_sort_impl(K* keys, uint32_t* vals, K* keys_alt, uint32_t* vals_alt, int64_t n,
                           int key_bits, K** keys_sorted, uint32_t** vals_sorted, cudaStream_t s) {
    constexpr int ITEMS = RsItems<K>::value;
    constexpr int TILE = RS_THREADS * ITEMS;
    *keys_sorted = keys;
    *vals_sorted = vals;
    if (n <= 1 || key_bits <= 0) return MORE3D_OK;
    if (n >= (int64_t)1 << 30) {
        set_error("radix sort: n >= 2^30 not supported");
        return MORE3D_E_RANGE;
    }
    const int npass = (key_bits + 7) / 8;
    const int ntiles = (int)((n + TILE - 1) / TILE);
    const size_t smem = (size_t)TILE * (sizeof(K) + sizeof(uint32_t)) + (size_t)(RS_WARPS * 256 + 256 + 33) * sizeof(uint32_t);
    M3D_CUDA(cudaFuncSetAttribute(rs_onesweep<K, ITEMS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // [npass x 256 bins][ticket + pad][ntiles x 256 descriptors]
    uint32_t* work = nullptr;
    const size_t n_bins = (size_t)npass * 256, n_desc = (size_t)ntiles * 256;
    M3D_TRY(dev_alloc(&work, n_bins + 4 + n_desc, s));
    uint32_t* bins = work;
    unsigned* ticket = work + n_bins;
    uint32_t* desc = work + n_bins + 4;
    M3D_CUDA(cudaMemsetAsync(work, 0, (n_bins + 4) * sizeof(uint32_t), s));
    int hb = (int)((n + (int64_t)RS_THREADS * 16 - 1) / ((int64_t)RS_THREADS * 16));
    if (hb > 148 * 8) hb = 148 * 8;
    rs_hist_all<K><<<hb, RS_THREADS, 0, s>>>(keys, n, npass, bins);
    rs_scan_bins<<<npass, 256, 0, s>>>(bins);
    count_launches(2);
    K* kin = keys;
    K* kout = keys_alt;
    uint32_t* vin = vals;
    uint32_t* vout = vals_alt;
    for (int p = 0; p < npass; ++p) {
        M3D_CUDA(cudaMemsetAsync(ticket, 0, (4 + n_desc) * sizeof(uint32_t), s));
        rs_onesweep<K, ITEMS><<<ntiles, RS_THREADS, smem, s>>>(kin, vin, kout, vout, n, 8 * p, bins + 256 * p, desc, ticket);
        count_launches(1);
        M3D_CHECK_LAUNCH();
        K* tk = kin; kin = kout; kout = tk;
        uint32_t* tv = vin; vin = vout; vout = tv;
    }
    dev_free(work, s);
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
extern "C" int more3d_read_back(const void* src_dev, void* dst_h, int64_t bytes, void* stream) {
    M3D_REQUIRE(src_dev && dst_h && bytes >= 0, "NULL argument");
    return m3d::read_back(dst_h, src_dev, (size_t)bytes, (cudaStream_t)stream);
}
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
