/*
 * more3d_b200 -- C-ABI of the B200 (sm_100a) per-point neighbourhood-feature path of MorE3D.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.  Every entry point
 * cites the reference routine it replaces (paths relative to TUW-GEO/MorE3D).  Conventions:
 *   - every function returns int: 0 = ok, < 0 = error (see MORE3D_E_*); it never throws or aborts;
 *     more3d_last_error() returns a thread-local message for the last failure.
 *   - pointers are DEVICE pointers unless the parameter name ends in _h (host memory).
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Work is enqueued
 *     on that stream; functions that return a host scalar (counts, sizes) synchronise it.
 *   - the caller allocates every output; opaque handles (more3d_grid*) are owned by the library.
 *   - per-point arrays are in the ORIGINAL point order of the cloud the grid was built from.
 *   - one grid per GPU/stream user; functions are not re-entrant on the same grid.
 */
#ifndef MORE3D_B200_H
#define MORE3D_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MORE3D_OK 0
#define MORE3D_E_INVALID (-1)  /* bad argument */
#define MORE3D_E_CUDA (-2)     /* CUDA runtime error (message in more3d_last_error) */
#define MORE3D_E_NOMEM (-3)    /* device allocation failed */
#define MORE3D_E_RANGE (-4)    /* size exceeds an implementation limit */

typedef struct more3d_grid more3d_grid;

int more3d_version(void);
const char* more3d_last_error(void);

/* ---- optional per-kernel timing (CUDA events on the launching stream; used by bench.py) ----------
 * enable(1) makes every library kernel group record an event pair; get() synchronises the recorded
 * events and returns the accumulated milliseconds / launch count of one named group, e.g.
 * "moments", "dkdn" (with its parts "dkdn_screen" and "dkdn_refine"), "grid_build", "build_feat".  names() writes a
 * comma-separated list. */
int64_t more3d_launch_count(void); /* kernels launched by this library since it was loaded */
int more3d_profile_enable(int on);
int more3d_profile_reset(void);
int more3d_profile_get(const char* name, double* ms_h, int64_t* count_h);
int more3d_profile_names(char* buf_h, int len);
/* Read `bytes` (<= 4096 for the fast path) of device memory into host memory, ordered behind everything queued on
 * `stream`, WITHOUT the device-to-host copy engine: a one-warp kernel stores into a pinned mapped slot and the host
 * waits for an event.  A small cudaMemcpy would queue behind any large download in flight (tile streaming). */
int more3d_read_back(const void* src_dev, void* dst_h, int64_t bytes, void* stream);

/* Kernel variant of the dk/dn pass (more3d_dk_dn*), for A/B measurements and the equivalence tests; grids built
 * afterwards follow it (the feature-record layout is chosen per grid).  0 = two passes: a screening walk that settles
 * every point whose outputs the count rules of Saliency_Kernel.process (DM_saliency.py:286-316) force to zero, then a
 * refinement of the listed rest (default); 1 = shared-memory staged one-pass walk; 2 = global-memory one-pass walk.
 * The environment variable MORE3D_DKDN_VARIANT sets the initial value. */
int more3d_set_dkdn_variant(int variant);
/* Two-pass dk/dn only: when the screening pass lists more than `fraction` * n points for refinement (a whole-expression
 * entry counts three times) the one-pass kernel is run over the whole cloud instead -- decided on the device, both
 * alternatives are launched.  Default 0.45; a value >= 4 never falls back, 0 always does (tests). */
int more3d_set_dkdn_gate(double fraction);
/* Host-only helper behind the screening pass: the reference calls a neighbour ACTIVE when its ring weight
 * w = d2/rho^2 (d2 < rho^2) or 2 - d2/rho^2 (d2 > rho^2) exceeds 0.01 (DM_saliency.py:272-279).  Both branches are
 * monotone roundings of d2, so the test equals act_lo <= d2 <= act_hi, d2 != rho^2 for two doubles found with the same
 * arithmetic: *act_lo_h = smallest d2 with fl(d2 / rho^2) > 0.01, *act_hi_h = largest d2 with fl(2 - fl(d2 / rho^2)) > 0.01.
 * Needs no device. */
int more3d_dkdn_act_interval(double rho, double* act_lo_h, double* act_hi_h);

/* ---- spatial index --------------------------------------------------------------------------
 * Replaces BallTreePointSet.__init__ -> sklearn BallTree build (Downsampling/BallTreePointSet.py:19-53).
 * Uniform xy-column grid: fp64 cell keys -> LSD radix sort -> column-start table -> sorted
 * (x, y, z, original index) records.  `cell` is the column width; a query of radius r scans a
 * stencil of ceil(r / cell) columns either side.  xyz: n*3 float64, row-major (x0 y0 z0 x1 ...). */
int more3d_grid_create(const double* xyz, int64_t n, double cell, void* stream, more3d_grid** out);
/* Generalised build (no reference equivalent; multi-GPU strips).  The cloud is the concatenation of 1 to 3 device
 * segments -- a strip's left ghosts, its own points, its right ghosts -- indexed in place (no concatenated copy);
 * a point's original index is its position in that concatenation.
 * lattice_h (host, may be NULL) = {origin_x, origin_y} of a column lattice SHARED with other grids: columns are then
 * floor((x - origin_x) / cell) of that lattice (shifted by an integer), so every strip of a tile bins, orders and
 * sums exactly like a single grid over the whole tile built with the same origin and cell (the cell is not widened
 * in this mode).  bounds_h (host, may be NULL) = {min x, y, z, max x, y, z} of the segments when the caller already
 * has them: skips the bounding-box reduction and its host synchronisation.
 * x_split (1..8): columns are cell / x_split wide along x and `cell` high along y.  A stencil row stays ONE contiguous
 * range whatever the column width, so narrower columns only trim the row's overhang beyond the query circle
 * (2 r + cell / x_split instead of 2 r + cell): fewer rejected candidates in every fixed-radius pass.  more3d_knn needs
 * x_split = 1. */
int more3d_grid_create_segments(const double* const* seg_xyz_h, const int64_t* seg_n_h, int nseg, double cell,
                                int x_split, const double* lattice_h, const double* bounds_h, void* stream,
                                more3d_grid** out);
int more3d_grid_destroy(more3d_grid* g);
/* info[0]=n, info[1]=nx, info[2]=ny (host ints); geom[0]=cell, geom[1]=origin_x, geom[2]=origin_y */
int more3d_grid_info(const more3d_grid* g, int64_t* info_h, double* geom_h);
/* sorted position -> original index, n x int32 (device) */
int more3d_grid_order(const more3d_grid* g, int32_t* order, void* stream);

/* ---- fixed-radius query -----------------------------------------------------------------------
 * Replaces BallTreePointSet.queryRadius (BallTreePointSet.py:317-347 -> sklearn query_radius).
 * Accept rule, bit-exact with the reference: fp64 ((dx*dx)+(dy*dy))+(dz*dz) <= r*r, no FMA
 * (sklearn metrics/_dist_metrics.pxd.tp:39-49, neighbors/_binary_tree.pxi.tp:1951-1957).
 * q == NULL means "the cloud itself" (nq must equal n); results are still in original order.
 * Two-pass CSR: count -> (caller or more3d_offsets_from_counts) -> fill. */
int more3d_radius_count(const more3d_grid* g, const double* q, int64_t nq, double r,
                        int32_t* counts, void* stream);
/* offsets[0..nq] = exclusive prefix sum of counts (int64); total_h (host, may be NULL) = offsets[nq]. */
int more3d_offsets_from_counts(const int32_t* counts, int64_t nq, int64_t* offsets,
                               int64_t* total_h, void* stream);
/* idx_bytes = 4 (int32) or 8 (int64) original point indices; order inside a list is unspecified
 * (the reference's is too: compare as sets). */
int more3d_radius_fill(const more3d_grid* g, const double* q, int64_t nq, double r,
                       const int64_t* offsets, void* idx, int idx_bytes, void* stream);

/* ---- k nearest neighbours ---------------------------------------------------------------------
 * Replaces BallTreePointSet.query (BallTreePointSet.py:292-315) and the 2k-NN of
 * levelsets_func.build_neighborhoods (levelsets_func.py:158-163).  Sorted by (distance, index);
 * idx: nq*k int64, dist: nq*k float64 (may be NULL).  Distances are sqrt of the fp64 rule above.
 * q == NULL -> self query.  Requires k <= MORE3D_MAX_K and k <= n. */
#define MORE3D_MAX_K 64
int more3d_knn(const more3d_grid* g, const double* q, int64_t nq, int k, int64_t* idx, double* dist,
               void* stream);

/* ---- neighbourhood moments: PCA normal, surface variation, non-parametric curvature -------------
 * more3d_normals_pca: centroid-PCA normal (sign arbitrary) and lambda_min/sum(lambda) over the
 * radius-r neighbourhood (self included); precondition of DM_nonparam_curvature (DM_saliency.py:329-330;
 * open3d estimate_normals in downsample_compare.py:53-64).  Fewer than 3 neighbours -> NaN normal.
 * normals: n*3 f64; lam_ratio: n f64 (may be NULL); counts: n int32 (may be NULL). */
int more3d_normals_pca(const more3d_grid* g, double r, double* normals, double* lam_ratio,
                       int32_t* counts, void* stream);
/* Curvature_Kernel.process for every point (DM_saliency.py:88-157).  eps = norm.ppf(1-alpha/2)*roughness
 * is computed by the caller with scipy exactly as the reference does (:99).  normals are read and, where
 * n.(-p) < 0, flipped in place (:131-135).  K: n float32 (_nonparam_K, :381); pcount: n int32 (:382). */
int more3d_nonparam_curvature(const more3d_grid* g, double r, double* normals_inout, double eps,
                              int min_points, float* K, int32_t* pcount, void* stream);
/* Both of the above in one neighbourhood pass (normals at the same radius as the curvature). */
int more3d_normals_curvature(const more3d_grid* g, double r, double eps, int min_points,
                             double* normals, double* lam_ratio, float* K, int32_t* pcount,
                             void* stream);

/* ---- dk / dn centre-surround differences ---------------------------------------------------------
 * Saliency_Kernel.process for every point (DM_saliency.py:219-324).  chi2_table[df] =
 * chi2.ppf(alpha, df) for df = 0..table_len-1, computed by the caller with scipy (:282).  A neighbourhood with
 * more active neighbours than table_len-1 is an overflow: with status == NULL the call synchronises the stream
 * and returns MORE3D_E_RANGE; with status != NULL (device int32, zeroed by the caller) nothing synchronises --
 * the kernel raises *status to (largest active count + 1), i.e. the table length that would have sufficed, and the
 * caller reads it whenever it likes (once per pipeline) and repeats the pass with a longer table.
 * NaN outputs (dk = |0/0| when no neighbour is active, :316) stay in the columns but are kept out of minmax.
 * normals and K may BOTH be NULL: the call then uses the sorted feature records the grid cached during the
 * last more3d_normals_curvature on it (valid only while those outputs are unmodified) and skips the gather.
 * dk, dn: n float32 (_dk, _dn; :491-492).  minmax (may be NULL): 4 float32 {min_dk, max_dk, min_dn,
 * max_dn} over this call's outputs, reduced on the device (the Histo pass of DM_saliency, :562-568). */
int more3d_dk_dn(const more3d_grid* g, double r, const double* normals, const float* K, double rho,
                 double sigma_normal, double sigma_curvature, const double* chi2_table, int table_len,
                 int min_points, float* dk, float* dn, float* minmax, int32_t* status, void* stream);

/* ---- grid-order variants of the two neighbourhood passes ------------------------------------------
 * Same arithmetic, same values; every per-point output is written at the point's position in the grid's
 * sorted order (row i belongs to original point order[i], more3d_grid_order) instead of at its original
 * index.  Sorted writes are coalesced; the original-order calls scatter six partial 32-B sectors per point,
 * which is half of their per-point cost on a cloud in random order.  A pipeline that only needs the final
 * column back in the caller's order (DM_nonparam_curvature -> DM_dk_dn -> DM_saliency, DM_saliency.py:326-583)
 * runs these, blends in grid order (more3d_saliency_blend is order-agnostic) and unpermutes one column (more3d_unpermute).
 * more3d_dk_dn_gridorder always uses the feature records cached by the last
 * more3d_normals_curvature[_gridorder] on the grid.  own_lo, n_owned: only points whose ORIGINAL index lies in
 * [own_lo, own_lo + n_owned) enter `minmax` (a strip's own points sit between its left and right ghosts,
 * more3d_grid_create_segments); n_owned <= 0 means all points. */
int more3d_normals_curvature_gridorder(const more3d_grid* g, double r, double eps, int min_points,
                                       double* normals, double* lam_ratio, float* K, int32_t* pcount,
                                       void* stream);
int more3d_dk_dn_gridorder(const more3d_grid* g, double r, double rho, double sigma_normal,
                           double sigma_curvature, const double* chi2_table, int table_len, int min_points,
                           int64_t own_lo, int64_t n_owned, float* dk, float* dn, float* minmax, int32_t* status,
                           void* stream);
/* dst[order[i]] = src[i], order as returned by more3d_grid_order (the grid itself is not needed any more);
 * elem_bytes in {4, 8, 24} (f32 / i32 column, f64 column, n x 3 f64 normals). */
int more3d_unpermute(const int32_t* order, int64_t n, const void* src_gridorder, void* dst, int elem_bytes,
                     void* stream);

/* ---- DM_saliency: min-max normalise (+min, as written) and blend (DM_saliency.py:565-580) ---------
 * minmax: 4 float32 on the device as produced by more3d_dk_dn / more3d_minmax4.
 * sal: n float32.  sal_minmax (may be NULL): 2 float32 {min, max} of sal. */
int more3d_minmax4(const float* dk, const float* dn, int64_t n, float* minmax, void* stream);
int more3d_saliency_blend(const float* dk, const float* dn, int64_t n, const float* minmax,
                          double curvature_weight, float* sal, float* sal_minmax, void* stream);
/* 256-bin histogram of sal over [lo, hi] with numpy.histogram's binning; edges: 257 float64 from
 * numpy.linspace (host computed, device resident).  hist: 256 int64 (zeroed by the call). */
int more3d_histogram256(const void* v, int elem_bytes, int64_t n, const double* edges, int64_t* hist,
                        void* stream);

/* ---- voxel-grid downsampling -----------------------------------------------------------------------
 * Replaces open3d voxel_down_sample as used by Downsampling/Downsample.py:86-89: voxel index =
 * floor((p - (min_bound - voxel/2)) / voxel) per axis; one output point per occupied voxel = mean of its
 * members (accumulated in original point order).  Outputs are sorted by voxel index (x, then y, then z;
 * open3d's own order is hash-map order).  out_xyz: capacity n*3 f64; out_keys: capacity n*3 int64 or
 * NULL; voxel_of_point: n int32 (output row of every input point) or NULL; *nout_h = voxels. */
int more3d_voxel_downsample(const double* xyz, int64_t n, double voxel, double* out_xyz,
                            int64_t* out_keys, int32_t* voxel_of_point, int64_t* nout_h, void* stream);
/* The same on one SHARD of a cloud: bounds_h = {min x, y, z, max x, y, z} of the whole cloud (6 host doubles), so
 * that every shard bins on the lattice the unsharded call would use.  With voxel-aligned shards (all members of a
 * voxel on one rank, in the whole cloud's point order) the union of the shards' outputs is the unsharded output,
 * bit for bit; all points must lie inside the bounds. */
int more3d_voxel_downsample_bounds(const double* xyz, int64_t n, double voxel, const double* bounds_h,
                                   double* out_xyz, int64_t* out_keys, int32_t* voxel_of_point, int64_t* nout_h,
                                   void* stream);

/* ---- level-of-detail cells: ball-tree-equivalent hierarchy -------------------------------------------
 * Replaces the sklearn BallTree build + the hierarchy bookkeeping of BallTreePointSet.__init__
 * (Downsampling/BallTreePointSet.py:35-53): same node membership as sklearn's recursive median split
 * (heap layout, children of i = 2i+1, 2i+2).  shape(): host helper giving n_levels / n_nodes for a
 * cloud size (sklearn _binary_tree.pxi.tp:872-873).  build(): idx_array n int64; idx_start / idx_end
 * n_nodes int64; is_leaf n_nodes int32; radius n_nodes f64 (node_data of BallTree.get_arrays()). */
int more3d_balltree_shape(int64_t n, int leaf_size, int64_t* n_levels_h, int64_t* n_nodes_h);
/* getSmallestNodesOfSize(lod) (BallTreePointSet.py:128-157, :256-276) on the node tables of more3d_balltree_build:
 * the first node on every branch with radius < lod, or the leaf.  cells: capacity n_nodes int64, in the
 * reference's depth-first left-before-right order (= the order of the cells' ranges in idx_array); starts:
 * capacity n_nodes + 1 int64, idx_start of every cell followed by n; *count_h = number of cells. */
int more3d_lod_select(const int64_t* idx_start, const int32_t* is_leaf, const double* radius, int64_t n_nodes,
                      int64_t n, double lod, int64_t* cells, int64_t* starts, int64_t* count_h, void* stream);
int more3d_balltree_build(const double* xyz, int64_t n, int leaf_size, int64_t* idx_array,
                          int64_t* idx_start, int64_t* idx_end, int32_t* is_leaf, double* radius,
                          void* stream);
/* Per-cell rule of the saliency downsampler (Downsampling/Downsample.py:56-75, with the declared repair
 * of :58/:67 -- see DESIGN.md): cells are ranges [cell_start[c], cell_start[c+1]) of idx_array.
 * mask = saliency > thresh (and |saliency| < zero_thresh when literal != 0); a cell with more than 60 %
 * masked points is kept whole, otherwise its masked points are kept and the rest is replaced by their
 * mean.  keep: n uint8 per idx_array POSITION; keep_all: ncells uint8; mean_xyz: ncells*3 f64 (NaN when
 * unused). */
int more3d_lod_cell_rule(const double* xyz, const double* saliency, const int64_t* idx_array,
                         const int64_t* cell_start, int64_t ncells, double thresh, double zero_thresh,
                         int literal, uint8_t* keep, uint8_t* keep_all, double* mean_xyz, void* stream);
/* Output of the downsampler in the reference's row order (Downsample.py:60-83): per cell its kept points
 * (flag 1) then, for averaged cells, the original point `near[a]` (a = rank of the cell among the averaged
 * ones) with flag 0.  out_xyzf: (n + ncells) x 4 f64 capacity; out_idx: original index per row; out_cell
 * (may be NULL): cell rank per row; *nrows_h = rows written. */
int more3d_lod_assemble(const double* xyz, const int64_t* idx_array, int64_t n, const int64_t* cell_start,
                        int64_t ncells, const uint8_t* keep, const uint8_t* keep_all, const int64_t* near,
                        double* out_xyzf, int64_t* out_idx, int32_t* out_cell, int64_t* nrows_h, void* stream);
/* stable compaction: idx = positions i with flags[i] != 0 (ascending); idx has room for n int64 */
int more3d_select_flagged(const uint8_t* flags, int64_t n, int64_t* idx, int64_t* count_h, void* stream);

/* ---- point-based level set (levelsets_func.py) ------------------------------------------------------
 * more3d_ls_frames: the part of build_neighborhoods (levelsets_func.py:154-172) after the kNN query.
 * knn2k: n x 2k int64 sorted neighbours, self first (more3d_knn with q == NULL).  normals (n x 3 f64): when
 * has_normals == 0 they are computed as in compute_normals (:112-117: PCA of the offsets about the
 * query point), otherwise read; in both cases they are normalised in place (:140).  t1, t2: n x 3 f64
 * tangent frame (:141-146); neighbors: n x k int32 = knn2k[:, 1:k+1] (:172). */
int more3d_ls_frames(const double* xyz, int64_t n, int k, const int64_t* knn2k, int has_normals,
                     double* normals, double* t1, double* t2, int32_t* neighbors, void* stream);

/* Derivative.__init__ (levelsets_func.py:246-275): per-point weighted-least-squares operator.  weights:
 * n x k f64 (self.weights) -- written when weights_given == 0, read when != 0 (numpy's array pow in wendland
 * is host-CPU dependent in the last bit; a caller reproducing a reference run bit for bit supplies the
 * reference's weights); D: n x 5 x (k+3) f64 out or NULL (self.D); *singular_h = 1 if some
 * singular value is < 1e-12 (where the reference drops into pdb, :269-271).  The returned graph keeps the
 * device-resident structure-of-arrays form used by the calls below. */
typedef struct more3d_ls_graph more3d_ls_graph;
int more3d_ls_graph_create(const double* xyz, int64_t n, int k, const int32_t* neighbors, const double* t1,
                           const double* t2, double max_d, int weights_given, double* weights, double* D,
                           int* singular_h, void* stream, more3d_ls_graph** out);
int more3d_ls_graph_destroy(more3d_ls_graph* g);
/* Derivative.__call__ (:277-291): out4[i] = (dx, dy, ddx, ddy) of u at the points with mask[i] != 0
 * (mask NULL = all); zeros elsewhere.  u: n f64; out4: n x 4 f64. */
int more3d_ls_diff(const more3d_ls_graph* g, const double* u, const uint8_t* mask, double* out4, void* stream);
/* smooth (:350-358): out[i] = sum_j u[nbr_ij] w_ij / sum_j w_ij where mask[i] != 0 (NULL = all), else u[i]. */
int more3d_ls_smooth(const more3d_ls_graph* g, const double* u, const uint8_t* mask, double* out, void* stream);
/* Solver.zero_crossings (:524-534): out[i] = any neighbour sign differs, for mask[i] != 0 (NULL = all). */
int more3d_ls_zero_crossings(const more3d_ls_graph* g, const double* u, const uint8_t* mask, uint8_t* out,
                             void* stream);
/* Solver.initialize_from_voxels (:591-594): phi = -max_val where sum(points // h) is even else +max_val,
 * with numpy's float floor division. */
int more3d_ls_init_voxels(const double* xyz, int64_t n, double h, double max_val, double* phi, void* stream);
/* Solver.run with the Chan-Vese model (:598-642 + _chan_vese_step :386-437).  phi: n f64 in/out; zeta: n x
 * channels f64; active / last_active: n uint8 in/out; alias_last_active != 0 reproduces the reference's
 * aliasing of the two masks after initialize_from_voxels / seeds / point (:581, :589, :596).  epsilon is
 * already multiplied by h (:611).  rmsc_h: iterations f64 (host) or NULL; tolerance <= 0 never stops early. */
int more3d_ls_run(const more3d_ls_graph* g, double* phi, const double* zeta, int channels, uint8_t* active,
                  uint8_t* last_active, int alias_last_active, int iterations, double stepsize, double nu,
                  double mu, double lambda1, double lambda2, double epsilon, double max_val, int cap,
                  double tolerance, double* rmsc_h, int* iterations_done_h, int* converged_h, void* stream);

/* The region-growing part of Solver.save(extract=True) (:651-661): the region (phi > 0 or phi <= 0) with the
 * larger mean cue, grown by the points with >= extraction_threshold salient neighbours, pruned of the points
 * without any salient neighbour.  phi: n f64; zeta: n x channels f64; salient: n uint8 out. */
int more3d_ls_extract(const more3d_ls_graph* g, const double* phi, const double* zeta, int channels,
                      int extraction_threshold, uint8_t* salient, void* stream);

/* Solver.run with active_contour_model='lmv' (_lmv_step, :439-513): same regularisation terms, the data term
 * is the Gaussian log-likelihood difference of each point's (k+1)-neighbourhood under its local inside /
 * outside means and deviations.  zeta: n f64, ONE channel (the reference's expressions only broadcast for a
 * 1-D cue; with the (N, C) cue its Solver constructor requires, _lmv_step raises ValueError -- the caller
 * squeezes an (N, 1) cue).  lambda1 / lambda2 do not enter (:520-521). */
int more3d_ls_run_lmv(const more3d_ls_graph* g, double* phi, const double* zeta, uint8_t* active,
                      uint8_t* last_active, int alias_last_active, int iterations, double stepsize, double nu,
                      double mu, double epsilon, double max_val, int cap, double tolerance, double* rmsc_h,
                      int* iterations_done_h, int* converged_h, void* stream);

/* The same iteration one phase at a time, for hosts that exchange ghost values between the phases
 * (multi-GPU x-strips).  Points [0, n_own) of the graph are owned by the caller, the rest are ghosts
 * (their neighbour rows list themselves; their phi / aux are supplied by the owner between phases).
 * phase 0: re-clamp (de)activated points; 1: derivatives -> idx2, laplace, dphi, aux; 2: local region sums of
 * the owned points -> region[4*channels] (all-reduce them; 'lmv': local statistics -> stats, pull the ghosts');
 * 3: update -> bufA, incs[2] = {sum inc^2, count};
 * 4: zero crossings -> azc, cap -> bufB / outside; 5: smooth capped points bufB -> bufA; 6: smooth lonely
 * points bufA -> phi; 7: active = azc dilated (ghost flags go back to their owners).  All arrays have n
 * entries unless noted; part: more3d_ls_scratch_doubles() doubles. */
typedef struct {
    double* phi;
    const double* zeta;
    uint8_t* active;
    uint8_t* last_active;
    uint8_t* idx2;
    double* laplace;
    double* dphi;      /* 3n */
    double* aux;       /* 4n */
    double* bufA;
    double* bufB;
    uint8_t* azc;
    uint8_t* outside;
    double* part;
    double* region;    /* 4 * channels */
    double* incs;      /* 2 */
    int64_t n_own;
    double stepsize, nu, mu, lambda1, lambda2, epsilon, max_val;
    int channels, cap;
    /* 'lmv' model (model = 1, channels = 1; model = 0 is Chan-Vese and ignores the rest): phase 2 then computes the
     * local statistics records stats[4n] = {mu_in, mu_out, sigma_in, sigma_out} of every point of idx_ (:463-495)
     * and of every point flagged in `force` (n uint8 or NULL: the points other ranks hold as ghosts) instead of the
     * region sums; the caller then supplies the ghosts' records (and, once, their zeta) before phase 3. */
    int model, reserved;
    double* stats;          /* 4n */
    uint8_t* idxd;          /* n, scratch */
    const uint8_t* force;   /* n or NULL */
} more3d_ls_state;
int more3d_ls_phase(const more3d_ls_graph* g, int phase, const more3d_ls_state* st, void* stream);
int64_t more3d_ls_scratch_doubles(void);

/* Module-level helpers of levelsets_func.py on the caller's own arrays: smooth (:350-358) with u: n x cols
 * f64, nbs: n x k int64, w: n x k f64, mask: n uint8 or NULL; and the elementwise heaviside (op 0, param =
 * epsilon; :319-321), delta (op 1, param = epsilon; :332-333), dp (op 2; :336-342), wendland (op 3, param = h;
 * :92-97). */
int more3d_ls_smooth_raw(const double* u, int cols, const int64_t* nbs, const double* w, int64_t n, int k,
                         const uint8_t* mask, double* out, void* stream);
int more3d_ls_elementwise(int op, const double* x, double param, int64_t n, double* out, void* stream);

/* ---- cue pre-processing of the level-set driver ("next" row; run_levelsets.py:215-221) -----------------------------
 * select_kth: exact order statistics -- out[i] (device) = the ranks_h[i]-th smallest of v (0-based, 1 <= nk <= 4), by an
 * 8-pass radix select on order-preserving keys (no sort); np.percentile's two neighbouring statistics come from one
 * call.  cue_ops, in place on x (n f64): op 0 = np.abs; op 1 = clip(x, arg_dev[0], arg_dev[1]) (levelsets_func.py:107-109);
 * op 2 = normalize(x, s = arg) (:100-104: skipped when the column is all zero), out_dev (may be NULL) = {min, max};
 * op 3 = out_dev[0..1] = {0, numpy `linear` interpolation of arg_dev[0], arg_dev[1] at gamma = arg}: the clip bounds
 * of `clip(zeta, 0, np.percentile(zeta, pc))` without a host round trip; op 4 = out_dev[0..1] = {min, max} of x; op 5 = out_dev[0] = mean of x. */
/* DM_saliency_stats (DM_saliency.py:585-607): out (device, 4 f64) = {min, max, mean, population sigma} of an f32 column */
int more3d_column_stats(const float* x, int64_t n, double* out, void* stream);
int more3d_select_kth(const double* v, int64_t n, const int64_t* ranks_h, int nk, double* out, void* stream);
int more3d_cue_ops(int op, double* x, int64_t n, const double* arg_dev, double arg, double* out_dev, void* stream);

/* ---- evaluation terms of Downsampling/downsample_compare.py ("next" row) -----------------------------
 * normals_from_knn: centroid-PCA normal over a kNN table (open3d estimate_normals(KDTreeSearchParamKNN),
 * downsample_compare.py:57, :112-121); fewer than 3 valid neighbours -> (0, 0, 1).  orient_normals: NaN ->
 * (0, 0, 1), then n[n.z < 0] = -n when upwards_z != 0 (:60-62).  compare_terms (:132-149): c2p_sq[i] for every
 * original point (n_orig f64), angle_deg[i] for every downsampled point (n_down f64, folded to [0, 90]). */
int more3d_normals_from_knn(const double* xyz, int64_t n, int k, const int64_t* knn, double* normals,
                            double* lam_ratio, void* stream);
int more3d_orient_normals(double* normals, int64_t n, int upwards_z, void* stream);
int more3d_compare_terms(const double* orig, int64_t n_orig, const double* down, int64_t n_down,
                         const int64_t* closest_of_down, const int64_t* closest_of_orig,
                         const double* normals_orig, const double* normals_down, double* c2p_sq,
                         double* angle_deg, void* stream);

/* ---- multi-GPU strip partition helpers (no reference equivalent; OPALS tile streaming,
 * DM_saliency.py:76-86, is the analogue) ----------------------------------------------------------
 * band_select: indices i (ascending, stable) with lo <= xyz[i][axis] < hi; idx has room for n int64;
 * *count_h receives how many were written.  gather_rows: dst[r][:] = src[idx[r]][:] for m rows of
 * ncols elements of elem_bytes (4 or 8). */
int more3d_band_select(const double* xyz, int64_t n, int axis, double lo, double hi, int64_t* idx,
                       int64_t* count_h, void* stream);
int more3d_gather_rows(const void* src, const int64_t* idx, int64_t m, int ncols, int elem_bytes,
                       void* dst, void* stream);

/* strip_pack: ONE pass over a strip's cloud that packs both halo bands -- the points with x < left_below into
 * left_out and those with x >= right_from into right_out (cap rows of 3 f64 each, original order kept) -- and reduces
 * the strip's bounding box.  info (device, 8 f64) = {min x, y, z, max x, y, z, rows in left band, rows in right band};
 * a band count above cap means rows were dropped.  Nothing synchronises: the caller ships info with its collective.
 * histogram256_range: numpy.histogram(v, bins=256, range=(lo, hi)) with lohi = {lo, hi} in DEVICE memory (the
 * all-reduced extrema, f32 or f64 per lohi_bytes; Downsample.py:54 -> skimage threshold_otsu), hist = 256 int64.
 * negate_odd: v[i] = -v[i] for odd i (rows of {min, max}: one MIN all-reduce then serves minima and maxima). */
int more3d_strip_pack(const double* xyz, int64_t n, double left_below, double right_from, double* left_out,
                      double* right_out, int64_t cap, double* info, void* stream);
/* saliency_blend for the grid-order columns of a strip (all g->n rows are blended): sal_minmax (2 f32, device) covers
 * only the rows whose ORIGINAL index lies in [own_lo, own_lo + n_owned) -- the strip's own points. */
int more3d_saliency_blend_owned(const more3d_grid* g, const float* dk, const float* dn, const float* minmax,
                                double curvature_weight, int64_t own_lo, int64_t n_owned, float* sal,
                                float* sal_minmax, void* stream);
int more3d_histogram256_range(const void* v, int elem_bytes, int64_t n, const void* lohi, int lohi_bytes,
                              int64_t* hist, void* stream);
int more3d_negate_odd(float* v, int64_t n, void* stream);

/* ---- synthetic input (bench / test infrastructure; no reference equivalent) ----------------------------
 * Counter-based open-terrain generator of SURVEY.md section 8d's model: point i of a `total`-point tile of
 * lx x ly metres is a pure function of (seed, i); ids are slab-major along x (slab = i * slabs / total), so an id
 * range is an x-strip.  Writes points [id_start, id_start + count) as count*3 f64 (x, y centred on the tile,
 * z - 1000 m).  numpy twin: more3d_b200/synthetic.py:hash_terrain_host. */
int more3d_synth_terrain(uint64_t seed, int64_t id_start, int64_t count, int64_t total, double lx, double ly,
                         int slabs, double noise, double* out_xyz, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MORE3D_B200_H */
