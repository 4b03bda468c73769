// K9: ball-tree-equivalent node hierarchy for the level-of-detail cells of the saliency downsampler.
// Replaces BallTreePointSet.__init__'s sklearn BallTree build + hierarchy reconstruction
// (Downsampling/BallTreePointSet.py:35-53), getPointsOfNode (:240-254) and getNodeRadius (:226).
//
// sklearn builds the tree by recursive median splits (neighbors/_binary_tree.pxi.tp:1033-1083):
// split dimension = largest (max - min) spread, first maximum wins (:598-645); the left child
// receives the floor(n/2) smallest points by (value on that dimension, point index)
// (neighbors/_partition_nodes.pyx:24-57), heap layout (children of i are 2i+1, 2i+2), all leaves on the
// last of n_levels = int(log2(max(1, (n-1)/leaf_size)) + 1) levels.  Node membership is therefore a
// deterministic function of the cloud, and is what this file reproduces -- level by level for all
// nodes at once, with three index lists kept sorted by x, y and z inside every node (stable segmented
// partitions = one global scan per list per level) instead of a recursive nth_element.
// Node radius = sqrt(max squared distance to the node centroid) (neighbors/_ball_tree.pyx.tp:86-145);
// the centroid here is a pairwise (children-first) fp64 sum, sklearn's is a sequential one in an
// implementation-defined order, so radii agree to a few ulp, not bit for bit.
#include <string.h>
#include <vector>
#include "common.cuh"
#include "scan.cuh"

namespace m3d {

constexpr int LT_THREADS = 256;

// node ranges of sklearn's recursive halving (neighbors/_binary_tree.pxi.tp:1033-1083) computed per node by walking
// down from the root: the path to heap node i is the binary representation of i + 1 below its leading one
// (0 = left child, 1 = right child); a left child takes floor(cnt / 2) positions, its sibling the rest.
// A node that is never visited by the recursive build (its parent is a leaf) gets the empty range [0, 0).
__global__ void __launch_bounds__(LT_THREADS) lt_node_ranges(int64_t n, int64_t n_nodes, int64_t* __restrict__ nstart,
                                                             int64_t* __restrict__ nend, int32_t* __restrict__ is_leaf) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    const unsigned long long path = (unsigned long long)(i + 1);
    const int depth = 63 - __clzll(path);
    int64_t b = 0, e = n;
    bool visited = true;
    for (int k = depth - 1; k >= 0; --k) {
        // the node at the current position is internal only if it has children in the heap and >= 2 points
        if (e - b < 2) { visited = false; break; }
        const int64_t mid = b + (e - b) / 2;
        if ((path >> k) & 1ull) b = mid; else e = mid;
    }
    if (!visited) { nstart[i] = 0; nend[i] = 0; is_leaf[i] = 0; return; }
    nstart[i] = b; nend[i] = e;
    is_leaf[i] = (2 * i + 1 >= n_nodes || e - b < 2) ? 1 : 0;
}

// Coordinate order of one axis = order of the keys k = sortable(v) - sortable(v_min): an exact, order-preserving
// unsigned integer of `bits` significant bits (the axis' bounding interval is known).  The order is produced in two
// steps instead of bits / 8 radix passes over 64-bit keys: a radix sort of the TOP 32 bits (4 passes over 8-byte pairs),
// then a fix-up of the short runs of equal top bits (a few points at 40 points / m^2: the top 32 bits resolve a
// quarter of a millimetre or better) by their low bits -- which are gathered once -- and, last, the point index.
__global__ void __launch_bounds__(LT_THREADS) lt_keys(const double* __restrict__ xyz, int64_t n, int d, uint64_t base,
                                                      uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    keys[i] = sortable_key(xyz[3 * i + d]) - base;
    vals[i] = (uint32_t)i;
}
__global__ void __launch_bounds__(LT_THREADS) lt_keys_hi(const double* __restrict__ xyz, int64_t n, int d, uint64_t base,
                                                         int shift, uint32_t* __restrict__ hi, uint32_t* __restrict__ vals) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    hi[i] = (uint32_t)((sortable_key(xyz[3 * i + d]) - base) >> shift);
    vals[i] = (uint32_t)i;
}
__global__ void __launch_bounds__(LT_THREADS) lt_gather_lo(const double* __restrict__ xyz, int64_t n, int d, uint64_t base,
                                                           int shift, const uint32_t* __restrict__ order,
                                                           uint32_t* __restrict__ lo) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const uint64_t k = sortable_key(xyz[3 * (int64_t)order[t] + d]) - base;
    lo[t] = shift ? (uint32_t)(k & ((1ull << shift) - 1ull)) : 0u;
}
// one thread per sorted position: its final slot inside its run of equal top bits = the number of run members with
// smaller low bits, or equal low bits and an earlier position (the stable sort left ties in index order)
constexpr int LT_MAX_RUN = 4096;
__global__ void __launch_bounds__(LT_THREADS) lt_tie_fix(const uint32_t* __restrict__ hi, const uint32_t* __restrict__ lo,
                                                         const uint32_t* __restrict__ order, int64_t n,
                                                         uint32_t* __restrict__ out, int* __restrict__ overflow) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const uint32_t h = hi[t], l = lo[t];
    int64_t b = t, e = t + 1;
    uint32_t rank = 0;
    while (b > 0 && hi[b - 1] == h) {
        --b;
        if (lo[b] <= l) ++rank;
        if (t - b > LT_MAX_RUN) { *overflow = 1; return; }
    }
    while (e < n && hi[e] == h) {
        if (lo[e] < l) ++rank;
        ++e;
        if (e - t > LT_MAX_RUN) { *overflow = 1; return; }
    }
    out[b + rank] = order[t];
}

struct Lists {
    uint32_t* p[3];
};

// one thread per node of the level: choose the split dimension from the sorted lists' end points, and say how many
// of the node's positions count as "left" in this level's partition (floor(cnt / 2) for a split node, all of
// them for a node that stays as it is) -- the exclusive scan of these counts is the number of lefts before each
// node's first position, known BEFORE the partition runs because a median split is size-deterministic
__global__ void __launch_bounds__(LT_THREADS) lt_split_dim(const double* __restrict__ xyz, Lists L,
                                                           const int64_t* __restrict__ nstart,
                                                           const int64_t* __restrict__ nend,
                                                           const int32_t* __restrict__ is_leaf, int64_t first,
                                                           int64_t count, int8_t* __restrict__ split_dim,
                                                           uint32_t* __restrict__ left_count) {
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const int64_t node = first + k;
    const int64_t b = nstart[node], e = nend[node];
    int8_t best = -1;
    if (!is_leaf[node] && e - b >= 2) {
        double max_spread = 0.0;
        best = 0;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const double lo = xyz[3 * (int64_t)L.p[d][b] + d], hi = xyz[3 * (int64_t)L.p[d][e - 1] + d];
            const double spread = hi - lo;
            if (spread > max_spread) { max_spread = spread; best = (int8_t)d; }   // first maximum wins
        }
    }
    split_dim[k] = best;
    left_count[k] = (uint32_t)(best < 0 ? (e - b) : (e - b) / 2);
}

// per position: mark the side (0 = left child, 1 = right child) of the point sitting at rank t - start
// of its node's split-dimension list
// ... and the node every POSITION belongs to on the next level (the partition keeps a node's range: positions
// [b, b + n_mid) become its left child, the rest its right child), written to a second table because the partition
// of this level still reads the current one
__global__ void __launch_bounds__(LT_THREADS) lt_mark_sides(Lists L, int64_t n, const uint32_t* __restrict__ pos_node,
                                                            const int64_t* __restrict__ nstart,
                                                            const int64_t* __restrict__ nend, int64_t first,
                                                            const int8_t* __restrict__ split_dim,
                                                            uint8_t* __restrict__ side,
                                                            uint32_t* __restrict__ pos_node_next) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int64_t node = pos_node[t];
    int d = -1;
    if (node >= first) d = split_dim[node - first];     // node < first: became a leaf above this level
    if (d < 0) { pos_node_next[t] = (uint32_t)node; return; }    // leaf: stays where it is
    const int64_t b = nstart[node], e = nend[node];
    const int64_t n_mid = (e - b) / 2;
    const bool right = (t - b) >= n_mid;
    side[L.p[d][t]] = (uint8_t)right;
    pos_node_next[t] = (uint32_t)(right ? 2 * node + 2 : 2 * node + 1);
}

// Stable segmented partition of the three lists in ONE launch and ONE pass over each (blockIdx.y = list): the
// left flags are scanned with a chained look-back across the tiles of the list; the rank of a left inside its node
// is the global count of lefts before it minus the (precomputed) count before the node's first position, so no
// segmented scan and no second read of the flags is needed.  side[] is a 1-byte table indexed by point id
// (n bytes: L2-resident up to ~100 M points).
constexpr int LTP_ITEMS = 8;
constexpr int LTP_TILE = LT_THREADS * LTP_ITEMS;
__global__ void __launch_bounds__(LT_THREADS) lt_partition3(Lists L, Lists O, int64_t n,
                                                            const uint32_t* __restrict__ pos_node,
                                                            const int64_t* __restrict__ nstart,
                                                            const int64_t* __restrict__ nend, int64_t first,
                                                            const int8_t* __restrict__ split_dim,
                                                            const uint8_t* __restrict__ side,
                                                            const uint32_t* __restrict__ node_lp,
                                                            uint64_t* desc, unsigned* ticket, int ntiles) {
    __shared__ uint64_t sm[33];
    __shared__ uint64_t s_prefix;
    __shared__ unsigned s_tile;
    const int d = blockIdx.y;
    const uint32_t* __restrict__ list = L.p[d];
    uint32_t* __restrict__ out = O.p[d];
    volatile uint64_t* vd = desc + (size_t)d * ntiles;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket + d, 1u);
    __syncthreads();
    const int tile = (int)s_tile;
    const int64_t base = (int64_t)tile * LTP_TILE + (int64_t)threadIdx.x * LTP_ITEMS;
    uint32_t pt[LTP_ITEMS], nd[LTP_ITEMS];
    uint8_t fl[LTP_ITEMS];           // bit 0: counts as left, bit 1: its node is split at this level, bit 2: side
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < LTP_ITEMS; ++i) {
        const int64_t t = base + i;
        fl[i] = 0;
        if (t < n) {
            pt[i] = list[t];
            nd[i] = pos_node[t];
            const int sdim = (int64_t)nd[i] >= first ? (int)split_dim[(int64_t)nd[i] - first] : -1;
            const bool split = sdim >= 0;
            uint8_t sd = 0;
            if (split) {
                if (sdim == d) {            // the node's split list: the side is positional, no gather
                    const int64_t b = nstart[nd[i]], e = nend[nd[i]];
                    sd = (uint8_t)((t - b) >= (e - b) / 2);
                } else {
                    sd = side[pt[i]];       // 1-byte table indexed by point id (L2-resident)
                }
            }
            fl[i] = (uint8_t)((sd == 0 ? 1 : 0) | (split ? 2 : 0) | (sd ? 4 : 0));
            s += fl[i] & 1u;
        }
    }
    uint64_t total;
    uint64_t ex = block_exclusive_scan<uint64_t>(s, &total, sm);
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
    for (int i = 0; i < LTP_ITEMS; ++i) {
        const int64_t t = base + i;
        if (t < n) {
            int64_t dst = t;
            if (fl[i] & 2u) {
                const int64_t b = nstart[nd[i]], e = nend[nd[i]];
                const int64_t n_mid = (e - b) / 2;
                const int64_t lrank = (int64_t)ex - (int64_t)node_lp[(int64_t)nd[i] - first];
                dst = (fl[i] & 4u) ? b + n_mid + ((t - b) - lrank) : b + lrank;
            }
            out[dst] = pt[i];
        }
        ex += fl[i] & 1u;
    }
}

// sums of the deepest visited nodes (thread per node, sequential over its points), then parents
__global__ void __launch_bounds__(LT_THREADS) lt_leaf_sums(const double* __restrict__ xyz,
                                                           const uint32_t* __restrict__ list,
                                                           const int64_t* __restrict__ nstart,
                                                           const int64_t* __restrict__ nend,
                                                           const int32_t* __restrict__ is_leaf, int64_t n_nodes,
                                                           double* __restrict__ sums) {
    int64_t node = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (node >= n_nodes || !is_leaf[node]) return;
    double sx = 0, sy = 0, sz = 0;
    for (int64_t t = nstart[node]; t < nend[node]; ++t) {
        const int64_t o = list[t];
        sx += xyz[3 * o]; sy += xyz[3 * o + 1]; sz += xyz[3 * o + 2];
    }
    sums[3 * node] = sx; sums[3 * node + 1] = sy; sums[3 * node + 2] = sz;
}
__global__ void __launch_bounds__(LT_THREADS) lt_parent_sums(const int32_t* __restrict__ is_leaf,
                                                             const int64_t* __restrict__ nstart,
                                                             const int64_t* __restrict__ nend, int64_t first,
                                                             int64_t count, double* __restrict__ sums) {
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const int64_t node = first + k;
    if (is_leaf[node] || nend[node] - nstart[node] < 1) return;
#pragma unroll
    for (int a = 0; a < 3; ++a) sums[3 * node + a] = sums[3 * (2 * node + 1) + a] + sums[3 * (2 * node + 2) + a];
}

// every position walks from the root to its leaf, raising the max squared distance of every node on
// the way (non-negative doubles order like their bit patterns, so a 64-bit integer atomicMax serves)
__global__ void __launch_bounds__(LT_THREADS) lt_radii(const double* __restrict__ xyz,
                                                       const uint32_t* __restrict__ list, int64_t n,
                                                       const int64_t* __restrict__ nstart,
                                                       const int64_t* __restrict__ nend,
                                                       const int32_t* __restrict__ is_leaf,
                                                       const double* __restrict__ sums,
                                                       unsigned long long* __restrict__ r2) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int64_t o = list[t];
    const double x = xyz[3 * o], y = xyz[3 * o + 1], z = xyz[3 * o + 2];
    int64_t node = 0;
    while (true) {
        const int64_t b = nstart[node], e = nend[node];
        const double cnt = (double)(e - b);
        const double cx = sums[3 * node] / cnt, cy = sums[3 * node + 1] / cnt, cz = sums[3 * node + 2] / cnt;
        const double d2 = dist2_rule(cx - x, cy - y, cz - z);
        const unsigned long long bits = (unsigned long long)__double_as_longlong(d2);
        if (bits > r2[node]) atomicMax(r2 + node, bits);   // racy pre-read only skips redundant atomics
        if (is_leaf[node]) break;
        node = (t - b) < (e - b) / 2 ? 2 * node + 1 : 2 * node + 2;
    }
}

__global__ void __launch_bounds__(LT_THREADS) lt_finish(const unsigned long long* __restrict__ r2,
                                                        const uint32_t* __restrict__ list, int64_t n,
                                                        int64_t n_nodes, double* __restrict__ radius,
                                                        int64_t* __restrict__ idx_array) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n_nodes) radius[t] = sqrt(__longlong_as_double((long long)r2[t]));
    if (t < n) idx_array[t] = (int64_t)list[t];
}

}  // namespace m3d

using namespace m3d;

extern "C" int more3d_balltree_shape(int64_t n, int leaf_size, int64_t* n_levels_h, int64_t* n_nodes_h) {
    M3D_REQUIRE(n > 0 && leaf_size >= 1 && n_levels_h && n_nodes_h, "bad argument");
    // sklearn/neighbors/_binary_tree.pxi.tp:872-873
    double ratio = (double)(n - 1) / (double)leaf_size;
    if (ratio < 1.0) ratio = 1.0;
    const int64_t levels = (int64_t)(log2(ratio) + 1.0);
    *n_levels_h = levels;
    *n_nodes_h = ((int64_t)1 << levels) - 1;
    return MORE3D_OK;
}

extern "C" int more3d_balltree_build(const double* xyz, int64_t n, int leaf_size, int64_t* idx_array,
                                     int64_t* idx_start, int64_t* idx_end, int32_t* is_leaf, double* radius,
                                     void* stream) {
    M3D_REQUIRE(xyz && idx_array && idx_start && idx_end && is_leaf && radius, "NULL argument");
    int64_t n_levels = 0, n_nodes = 0;
    M3D_TRY(more3d_balltree_shape(n, leaf_size, &n_levels, &n_nodes));
    if (n >= ((int64_t)1 << 32) || n_levels > 31) { set_error("cloud too large for the LOD tree"); return MORE3D_E_RANGE; }
    cudaStream_t s = (cudaStream_t)stream;
    init_device_pool();
    ProfScope prof("balltree_build", s);

    // node ranges are analytic (recursive halving): one thread per node, no host tables
    lt_node_ranges<<<(unsigned)((n_nodes + LT_THREADS - 1) / LT_THREADS), LT_THREADS, 0, s>>>(n, n_nodes, idx_start, idx_end,
                                                                                       is_leaf);
    count_launches(1);
    M3D_CHECK_LAUNCH();

    const unsigned nb = (unsigned)((n + LT_THREADS - 1) / LT_THREADS);
    uint64_t *k1 = nullptr, *k2 = nullptr, *ks = nullptr;
    uint32_t *v1 = nullptr, *v2 = nullptr, *vs = nullptr;
    uint32_t *lists[3] = {nullptr, nullptr, nullptr}, *alts[3] = {nullptr, nullptr, nullptr}, *pos_node = nullptr, *pos_node2 = nullptr, *node_lp = nullptr;
    uint64_t* lb_desc = nullptr;
    uint8_t* side = nullptr;
    int8_t* split_dim = nullptr;
    double* sums = nullptr;
    unsigned long long* r2 = nullptr;
    M3D_TRY(dev_alloc(&k1, (size_t)n, s)); M3D_TRY(dev_alloc(&k2, (size_t)n, s));
    M3D_TRY(dev_alloc(&v1, (size_t)n, s)); M3D_TRY(dev_alloc(&v2, (size_t)n, s));
    for (int d = 0; d < 3; ++d) M3D_TRY(dev_alloc(&lists[d], (size_t)n, s));
    for (int d = 0; d < 3; ++d) M3D_TRY(dev_alloc(&alts[d], (size_t)n, s));
    M3D_TRY(dev_alloc(&pos_node, (size_t)n, s));
    M3D_TRY(dev_alloc(&pos_node2, (size_t)n, s));
    M3D_TRY(dev_alloc(&node_lp, (size_t)((n_nodes + 1) / 2 + 2), s));
    M3D_TRY(dev_alloc(&lb_desc, (size_t)3 * ((n + LTP_TILE - 1) / LTP_TILE) + 4, s));
    M3D_TRY(dev_alloc(&side, (size_t)n, s));
    M3D_TRY(dev_alloc(&split_dim, (size_t)((n_nodes + 1) / 2 + 1), s));
    M3D_TRY(dev_alloc(&sums, (size_t)3 * n_nodes, s));
    M3D_TRY(dev_alloc(&r2, (size_t)n_nodes, s));

    // three lists sorted by (coordinate, index)
    double box[6];
    M3D_TRY(cloud_bbox(xyz, n, box, s));                     // one small host read
    int* overflow = nullptr;
    M3D_TRY(dev_alloc(&overflow, 1, s));
    uint32_t* k32 = reinterpret_cast<uint32_t*>(k1);         // the 64-bit key buffers double as 32-bit ones
    uint32_t* k32b = reinterpret_cast<uint32_t*>(k2);
    uint32_t* lo32 = k32b + n;                               // second half of k2: the gathered low bits
    for (int d = 0; d < 3; ++d) {
        auto skey = [](double v) { v = v + 0.0; uint64_t u; memcpy(&u, &v, 8); return (u >> 63) ? ~u : (u | 0x8000000000000000ull); };
        const uint64_t base = skey(box[d]), range = skey(box[3 + d]) - base;
        int bits = 0;
        while (bits < 64 && (range >> bits) != 0) ++bits;
        if (bits == 0) bits = 1;
        const int shift = bits > 32 ? bits - 32 : 0;
        uint32_t *ks32 = nullptr, *vs32 = nullptr;
        lt_keys_hi<<<nb, LT_THREADS, 0, s>>>(xyz, n, d, base, shift, k32, v1);
        M3D_TRY(radix_sort_pairs_u32(k32, v1, k32b, v2, n, bits - shift, &ks32, &vs32, s));
        int ovf = 0;
        if (shift) {
            // the sorted top bits may sit in either buffer; the low bits go to the other one's spare half
            uint32_t* lo_buf = (ks32 == k32b) ? (k32 + n) : lo32;
            M3D_CUDA(cudaMemsetAsync(overflow, 0, sizeof(int), s));
            lt_gather_lo<<<nb, LT_THREADS, 0, s>>>(xyz, n, d, base, shift, vs32, lo_buf);
            lt_tie_fix<<<nb, LT_THREADS, 0, s>>>(ks32, lo_buf, vs32, n, lists[d], overflow);
            M3D_CUDA(cudaMemcpyAsync(&ovf, overflow, sizeof(int), cudaMemcpyDeviceToHost, s));
            M3D_CUDA(cudaStreamSynchronize(s));
            count_launches(3);
        } else {
            M3D_CUDA(cudaMemcpyAsync(lists[d], vs32, (size_t)n * 4, cudaMemcpyDeviceToDevice, s));
            count_launches(1);
        }
        if (ovf) {
            // a run of more than LT_MAX_RUN equal top bits (many near-coincident coordinates): full-width radix sort
            lt_keys<<<nb, LT_THREADS, 0, s>>>(xyz, n, d, base, k1, v1);
            count_launches(1);
            M3D_TRY(radix_sort_pairs_u64(k1, v1, k2, v2, n, bits, &ks, &vs, s));
            M3D_CUDA(cudaMemcpyAsync(lists[d], vs, (size_t)n * 4, cudaMemcpyDeviceToDevice, s));
        }
    }
    dev_free(overflow, s);
    M3D_CUDA(cudaMemsetAsync(pos_node, 0, (size_t)n * 4, s));
    M3D_CUDA(cudaMemsetAsync(side, 0, (size_t)n, s));

    const int ntiles = (int)((n + LTP_TILE - 1) / LTP_TILE);
    for (int64_t lev = 0; lev + 1 < n_levels; ++lev) {
        const int64_t first = ((int64_t)1 << lev) - 1, count = (int64_t)1 << lev;
        Lists L; L.p[0] = lists[0]; L.p[1] = lists[1]; L.p[2] = lists[2];
        Lists O; O.p[0] = alts[0]; O.p[1] = alts[1]; O.p[2] = alts[2];
        lt_split_dim<<<(unsigned)((count + LT_THREADS - 1) / LT_THREADS), LT_THREADS, 0, s>>>(
            xyz, L, idx_start, idx_end, is_leaf, first, count, split_dim, node_lp);
        M3D_TRY(exclusive_scan_u32(node_lp, node_lp, count, s));
        lt_mark_sides<<<nb, LT_THREADS, 0, s>>>(L, n, pos_node, idx_start, idx_end, first, split_dim, side, pos_node2);
        M3D_CUDA(cudaMemsetAsync(lb_desc, 0, ((size_t)3 * ntiles + 4) * sizeof(uint64_t), s));
        lt_partition3<<<dim3((unsigned)ntiles, 3), LT_THREADS, 0, s>>>(L, O, n, pos_node, idx_start, idx_end, first, split_dim,
                                                                       side, node_lp, lb_desc, (unsigned*)(lb_desc + (size_t)3 * ntiles),
                                                                       ntiles);
        count_launches(3);
        M3D_CHECK_LAUNCH();
        for (int d = 0; d < 3; ++d) { uint32_t* tmp = lists[d]; lists[d] = alts[d]; alts[d] = tmp; }
        { uint32_t* tmp = pos_node; pos_node = pos_node2; pos_node2 = tmp; }
    }

    // centroids bottom-up, radii top-down
    M3D_CUDA(cudaMemsetAsync(sums, 0, (size_t)3 * n_nodes * 8, s));
    M3D_CUDA(cudaMemsetAsync(r2, 0, (size_t)n_nodes * 8, s));
    lt_leaf_sums<<<(unsigned)((n_nodes + LT_THREADS - 1) / LT_THREADS), LT_THREADS, 0, s>>>(
        xyz, lists[0], idx_start, idx_end, is_leaf, n_nodes, sums);
    count_launches(1);
    for (int64_t lev = n_levels - 2; lev >= 0; --lev) {
        const int64_t first = ((int64_t)1 << lev) - 1, count = (int64_t)1 << lev;
        lt_parent_sums<<<(unsigned)((count + LT_THREADS - 1) / LT_THREADS), LT_THREADS, 0, s>>>(
            is_leaf, idx_start, idx_end, first, count, sums);
        count_launches(1);
    }
    lt_radii<<<nb, LT_THREADS, 0, s>>>(xyz, lists[0], n, idx_start, idx_end, is_leaf, sums, r2);
    const int64_t m = n > n_nodes ? n : n_nodes;
    lt_finish<<<(unsigned)((m + LT_THREADS - 1) / LT_THREADS), LT_THREADS, 0, s>>>(r2, lists[0], n, n_nodes, radius,
                                                                                   idx_array);
    count_launches(2);
    M3D_CHECK_LAUNCH();
    dev_free(k1, s); dev_free(k2, s); dev_free(v1, s); dev_free(v2, s);
    for (int d = 0; d < 3; ++d) dev_free(lists[d], s);
    for (int d = 0; d < 3; ++d) dev_free(alts[d], s);
    dev_free(pos_node, s); dev_free(pos_node2, s); dev_free(node_lp, s); dev_free(lb_desc, s); dev_free(side, s); dev_free(split_dim, s);
    dev_free(sums, s); dev_free(r2, s);
    return MORE3D_OK;
}

// ------------------------------------------------------------------------------------------
// per-cell rule of the saliency downsampler (Downsampling/Downsample.py:56-75)
// ------------------------------------------------------------------------------------------
namespace m3d {
// One WARP per cell (cells hold from a handful to thousands of points; a thread per cell walked its points
// serially and uncoalesced: 2.1 ms per call at 10 M points).  Lanes stride over the cell's slice of idx_array; the
// counts and the complement's coordinate sums are combined with a fixed xor-shuffle tree, so the mean is
// reproducible run to run (its summation order differs from numpy's pairwise one by rounding only; the
// representative is the original point nearest to it, Downsample.py:70).
__global__ void __launch_bounds__(LT_THREADS) lt_cell_rule(const double* __restrict__ xyz,
                                                           const double* __restrict__ sal,
                                                           const int64_t* __restrict__ idx_array,
                                                           const int64_t* __restrict__ cell_start, int64_t ncells,
                                                           double thresh, double zero_thresh, int literal,
                                                           uint8_t* __restrict__ keep, uint8_t* __restrict__ keep_all,
                                                           double* __restrict__ mean_xyz) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t c = (((int64_t)blockIdx.x * blockDim.x) + threadIdx.x) >> 5; c < ncells; c += warps) {
        const int64_t b = cell_start[c], e = cell_start[c + 1];
        unsigned nsal = 0;
        double sx = 0, sy = 0, sz = 0;
        for (int64_t t = b + lane; t < e; t += 32) {
            const int64_t o = idx_array[t];
            const double v = sal[o];
            bool m = v > thresh;                                    // :58 (repaired: first conjunct)
            if (literal) m = m && (fabs(v) < zero_thresh);          // :58 literal: both conjuncts, and -> &
            keep[t] = m ? 1 : 0;
            if (m) ++nsal;
            else { sx += xyz[3 * o]; sy += xyz[3 * o + 1]; sz += xyz[3 * o + 2]; }   // :67 mean of the complement
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            nsal += __shfl_xor_sync(0xffffffffu, nsal, o);
            sx += __shfl_xor_sync(0xffffffffu, sx, o);
            sy += __shfl_xor_sync(0xffffffffu, sy, o);
            sz += __shfl_xor_sync(0xffffffffu, sz, o);
        }
        const int64_t cnt = e - b;
        const bool all = (double)((int64_t)nsal * 100) / (double)cnt > 60.0;  // :60
        if (all) {
            __syncwarp();
            for (int64_t t = b + lane; t < e; t += 32) keep[t] = 1;           // :61-63 keep the cell as is
        }
        if (lane == 0) {
            keep_all[c] = all ? 1 : 0;
            const double nan = __longlong_as_double(0x7ff8000000000000LL);
            if (all) {
                mean_xyz[3 * c] = mean_xyz[3 * c + 1] = mean_xyz[3 * c + 2] = nan;
            } else {
                const double k = (double)(cnt - (int64_t)nsal);
                mean_xyz[3 * c] = sx / k; mean_xyz[3 * c + 1] = sy / k; mean_xyz[3 * c + 2] = sz / k;
            }
        }
    }
}
}  // namespace m3d

extern "C" int more3d_lod_cell_rule(const double* xyz, const double* saliency, const int64_t* idx_array,
                                    const int64_t* cell_start, int64_t ncells, double thresh, double zero_thresh,
                                    int literal, uint8_t* keep, uint8_t* keep_all, double* mean_xyz, void* stream) {
    M3D_REQUIRE(xyz && saliency && idx_array && cell_start && keep && keep_all && mean_xyz, "NULL argument");
    if (ncells <= 0) return MORE3D_OK;
    cudaStream_t s = (cudaStream_t)stream;
    ProfScope prof("lod_cell_rule", s);
    const int64_t want = (ncells * 32 + LT_THREADS - 1) / LT_THREADS;      // one warp per cell, capped (grid-stride)
    lt_cell_rule<<<(unsigned)(want < 148 * 64 ? want : 148 * 64), LT_THREADS, 0, s>>>(
        xyz, saliency, idx_array, cell_start, ncells, thresh, zero_thresh, literal, keep, keep_all, mean_xyz);
    count_launches(1);
    M3D_CHECK_LAUNCH();
    return MORE3D_OK;
}

// ------------------------------------------------------------------------------------------
// getSmallestNodesOfSize (Downsampling/BallTreePointSet.py:128-157, :256-276) on the device
// ------------------------------------------------------------------------------------------
namespace m3d {
// A node is a level-of-detail cell iff it stops the descent (radius < lod or leaf) and none of its ancestors
// does.  One thread per node walks up its <= n_levels ancestors (heap layout: parent of i is (i - 1) / 2).
__global__ void __launch_bounds__(LT_THREADS) lod_cell_flags(const int32_t* __restrict__ is_leaf,
                                                             const double* __restrict__ radius, int64_t n_nodes,
                                                             double lod, uint32_t* __restrict__ flag) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    bool sel = (radius[i] < lod) || (is_leaf[i] == 1);
    for (int64_t a = i; sel && a > 0;) {
        a = (a - 1) >> 1;
        if ((radius[a] < lod) || (is_leaf[a] == 1)) sel = false;      // an ancestor already stopped (or is a leaf:
    }                                                                 // the node was never visited by the build)
    flag[i] = sel ? 1u : 0u;
}
__global__ void __launch_bounds__(LT_THREADS) lod_cell_compact(const uint32_t* __restrict__ pos, int64_t n_nodes,
                                                               const int64_t* __restrict__ idx_start,
                                                               uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    if (pos[i + 1] != pos[i]) {             // exclusive scan of the flags: a step marks a selected node
        keys[pos[i]] = (uint32_t)idx_start[i];
        vals[pos[i]] = (uint32_t)i;
    }
}
__global__ void __launch_bounds__(LT_THREADS) lod_cell_emit(const uint32_t* __restrict__ keys,
                                                            const uint32_t* __restrict__ vals, int64_t m, int64_t n,
                                                            int64_t* __restrict__ cells, int64_t* __restrict__ starts) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) { cells[i] = (int64_t)vals[i]; starts[i] = (int64_t)keys[i]; }
    if (i == m) starts[m] = n;
}
}  // namespace m3d

extern "C" int more3d_lod_select(const int64_t* idx_start, const int32_t* is_leaf, const double* radius,
                                 int64_t n_nodes, int64_t n, double lod, int64_t* cells, int64_t* starts,
                                 int64_t* count_h, void* stream) {
    M3D_REQUIRE(idx_start && is_leaf && radius && cells && starts && count_h, "NULL argument");
    M3D_REQUIRE(n_nodes > 0 && n > 0, "empty tree");
    if (n >= ((int64_t)1 << 32) || n_nodes >= ((int64_t)1 << 32)) { set_error("n >= 2^32 not supported"); return MORE3D_E_RANGE; }
    cudaStream_t s = (cudaStream_t)stream;
    init_device_pool();
    ProfScope prof("lod_select", s);
    uint32_t *pos = nullptr, *k1 = nullptr, *k2 = nullptr, *v1 = nullptr, *v2 = nullptr, *ks = nullptr, *vs = nullptr;
    M3D_TRY(dev_alloc(&pos, (size_t)n_nodes + 1, s));
    const unsigned nb = (unsigned)((n_nodes + LT_THREADS - 1) / LT_THREADS);
    lod_cell_flags<<<nb, LT_THREADS, 0, s>>>(is_leaf, radius, n_nodes, lod, pos);
    M3D_TRY(exclusive_scan_u32(pos, pos, n_nodes, s));
    uint32_t m32 = 0;
    M3D_CUDA(cudaMemcpyAsync(&m32, pos + n_nodes, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    M3D_CUDA(cudaStreamSynchronize(s));
    const int64_t m = (int64_t)m32;
    M3D_TRY(dev_alloc(&k1, (size_t)m, s)); M3D_TRY(dev_alloc(&k2, (size_t)m, s));
    M3D_TRY(dev_alloc(&v1, (size_t)m, s)); M3D_TRY(dev_alloc(&v2, (size_t)m, s));
    lod_cell_compact<<<nb, LT_THREADS, 0, s>>>(pos, n_nodes, idx_start, k1, v1);
    int bits = 1;
    while (bits < 32 && ((int64_t)1 << bits) < n) ++bits;
    // depth-first, left-before-right order of the reference == order of the cells' ranges in idx_array
    M3D_TRY(radix_sort_pairs_u32(k1, v1, k2, v2, m, bits, &ks, &vs, s));
    lod_cell_emit<<<(unsigned)((m + 1 + LT_THREADS - 1) / LT_THREADS), LT_THREADS, 0, s>>>(ks, vs, m, n, cells, starts);
    count_launches(3);
    M3D_CHECK_LAUNCH();
    dev_free(pos, s); dev_free(k1, s); dev_free(k2, s); dev_free(v1, s); dev_free(v2, s);
    *count_h = m;
    return MORE3D_OK;
}
