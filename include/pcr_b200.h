/*
 * pcr_b200.h -- C ABI of libpcr_b200.so: the B200 (sm_100a) implementation of the
 * per-iteration hot path of scomup/point-cloud-registration and of its once-per-target
 * builds.  Plain C symbols, plain pointers and sizes, no torch/numpy types.
 *
 * The reference has no FFI layer: its "operator API" is the Python class surface
 * re-exported at point_cloud_registration/__init__.py:1-10.  Each entry point below names
 * the reference interface it replaces (paths relative to
 * /root/reference/point_cloud_registration/).  The Python classes in
 * point_cloud_registration_b200/ bind these symbols with ctypes.
 *
 * Conventions
 *   - every function returns 0 on success or a negative PCR_ERR_* code; the message is
 *     available from pcr_last_error(ctx) (ctx may be NULL for creation failures);
 *   - the caller owns all host buffers; the library owns all device buffers;
 *   - xyz inputs are C-contiguous (n,3) float32 unless stated; a pointer may be a HOST
 *     pointer (pageable or pinned) or a CUDA DEVICE pointer -- the library checks with
 *     cudaPointerGetAttributes and copies accordingly;
 *   - one context = one GPU + one CUDA stream; calls on a context are serialised by the
 *     caller; every entry point is synchronous with respect to its host outputs;
 *   - 4x4 transforms are row-major float64; the state vector is dx = [dt(3), dtheta(3)]
 *     with the right-multiplicative update of reference math_tools.py:101-108.
 */
#ifndef PCR_B200_H
#define PCR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pcr_ctx pcr_ctx;

#define PCR_OK 0
#define PCR_ERR_CUDA (-1)      /* CUDA runtime failure (message carries the call)            */
#define PCR_ERR_ARG (-2)       /* bad argument                                               */
#define PCR_ERR_STATE (-3)     /* required structure not built (e.g. target not set)         */
#define PCR_ERR_SINGULAR (-4)  /* singular normal equations (np.linalg.LinAlgError)          */
#define PCR_ERR_NCCL (-5)      /* NCCL failure / library not found                           */
#define PCR_ERR_LIMIT (-6)     /* structure would exceed the implementation's limits         */

/* method ids for pcr_linearize / pcr_align */
#define PCR_ICP 0     /* icp.py:24-57                 */
#define PCR_PLANE 1   /* plane_icp.py:30-69           */
#define PCR_VPLANE 2  /* voxelized_plane_icp.py:23-64 */
#define PCR_NDT 3     /* ndt.py:24-57                 */

/* Layout of the 29-double normal-equation record returned by pcr_linearize:
 * [0..20] upper triangle of the 6x6 H, row major; [21..26] g; [27] e2; [28] inlier count. */
#define PCR_RECORD_LEN 29

/* ---- life cycle ---------------------------------------------------------------------- */
int pcr_version(void);
int pcr_device_count(int* count);
int pcr_create(int device_id, pcr_ctx** out);
int pcr_destroy(pcr_ctx* ctx);
const char* pcr_last_error(const pcr_ctx* ctx);

/* ---- target side (set_target) -------------------------------------------------------- */

/* Upload the target cloud.  Replaces `self.target = target.astype(np.float32)`
 * (icp.py:17-22, plane_icp.py:19-20). */
int pcr_set_target_points(pcr_ctx* ctx, const float* xyz, int64_t n);

/* Append points to the resident target cloud (the old part stays on the GPU); every structure over
 * the target is void afterwards and is rebuilt by the usual build calls, so the result equals
 * set_target(concatenate(old, new)).  Backs Registration.update_target, which the reference declares
 * and leaves unimplemented (registration.py:36-43). */
int pcr_append_target_points(pcr_ctx* ctx, const float* xyz, int64_t n);

/* Build the exact nearest-neighbour index over the target points.  Replaces
 * `KDTree(target)` (kdtree.py:18-25 -> pykdtree; icp.py:20, plane_icp.py:22). */
int pcr_build_nn_index(pcr_ctx* ctx);
/* Build the per-cell shell lists the ICP / PlaneICP correspondence pass streams (needs the NN
 * index; part of ICP.set_target / PlaneICP.set_target, not of KDTree()).  Built implicitly by the
 * first linearisation if it was not called. */
int pcr_build_correspondence_lists(pcr_ctx* ctx);

/* k-NN normal estimation on the device; k includes the point itself; float32 moment
 * formula of the reference replayed (quirk Q5).  Replaces estimate_norm_with_tree
 * (estimate_normals.py:27-87) as called from plane_icp.py:23-24. */
int pcr_estimate_normals(pcr_ctx* ctx, int k);

/* Inject / read back per-target-point normals, (n,3) float32, caller's point order.
 * Injection mirrors `PlaneICP.set_target(target, kdree, norm)` (plane_icp.py:25-27). */
int pcr_set_normals(pcr_ctx* ctx, const float* normals);
int pcr_get_normals(pcr_ctx* ctx, float* normals);

/* Voxel statistics of a cloud: group by floor(p / voxel_size), per-voxel mean, two-pass
 * sample covariance, drop voxels with fewer than min_points points, plane normal =
 * eigenvector of the smallest eigenvalue, optional closed-form inverse covariance, NN index
 * over the kept means.  `xyz` is (n,3) float32 (is_f64 = 0) or float64 (is_f64 = 1); the
 * statistics are accumulated in float64 exactly as the reference does for either dtype.
 * Replaces VoxelGrid.set_points (voxel.py:104-165) and calc_icov (voxel.py:69-102) as
 * called from voxelized_plane_icp.py:18-21 and ndt.py:18-22.  xyz = NULL, n = 0: use the resident
 * target points (pcr_set_target_points / pcr_append_target_points; float32). */
int pcr_build_voxels(pcr_ctx* ctx, const void* xyz, int64_t n, int is_f64, double voxel_size,
                     int min_points, int with_icov);

/* Number of kept voxels and of occupied voxels before the min_points filter. */
int pcr_get_voxel_count(pcr_ctx* ctx, int64_t* n_kept, int64_t* n_occupied);

/* Read back kept-voxel attributes (any pointer may be NULL): mean (n,3), cov (n,3,3),
 * norm (n,3), icov (n,3,3) float64, count (n) int64.  Voxel order = the library's own
 * (brick-major); the reference's order (ascending hash key, voxel.py:109) is not part of
 * its API.  Mirrors the attributes VoxelGrid.mean/.cov/.norm/.icov (voxel.py:160-164). */
int pcr_get_voxels(pcr_ctx* ctx, double* mean, double* cov, double* norm, double* icov, int64_t* count);

/* ---- scan side ------------------------------------------------------------------------- */

/* Upload the scan once per align().  sort > 0 re-orders it along a Morton curve on the device, so
 * that the lanes of a warp query neighbouring cells (they then stream the same candidate lists);
 * sort == 0 keeps the caller's order; sort < 0 keeps the order, the caller promises it is
 * spatially coherent.  Results are identical up to float64 summation order.  With a host source
 * the call returns once the caller's buffer has been read (the re-layout continues on the
 * context's stream, every later call is ordered behind it).  Replaces
 * `source.astype(np.float32)` (registration.py:83).
 * sort == 2: like 1, and the caller vouches that this is the same cloud as the previous upload of this size
 * (e.g. the same array passed to calc_H_g_e2 in every iteration of a host loop): the permutation of that upload
 * may be reused instead of being recomputed -- any order yields the same correspondences. */
int pcr_set_scan(pcr_ctx* ctx, const float* xyz, int64_t n, int sort);
/* Same, with the pose the iteration will start from and the method (PCR_ICP .. PCR_NDT, or -1)
 * whose correspondence grid is meant: with sort > 0 the scan is ordered by the grid cell its posed
 * points fall into, so that the 32 slots of a warp row share one or two candidate lists (a rigid
 * motion keeps that order coherent over the iterations).  Falls back to the Morton order of
 * pcr_set_scan when the grid does not exist yet.  T = NULL means identity. */
int pcr_set_scan_posed(pcr_ctx* ctx, const float* xyz, int64_t n, int sort, const double T[16], int method);

/* Device-resident inputs are read on the context's own (non-blocking) stream: before passing a
 * device pointer that another stream may still be writing, order the two.  has_stream = 1: wait
 * for `stream` (the "stream" entry of __cuda_array_interface__ v3: 1 = legacy default, 2 = per-thread
 * default, else a cudaStream_t); has_stream = 0: wait for the whole device.  The Python layer calls
 * this for every GPU array it is handed. */
int pcr_sync_producer(pcr_ctx* ctx, int has_stream, void* stream);

/* ---- per iteration --------------------------------------------------------------------- */

/* One linearisation at transform T: SE(3) transform of the scan, exact correspondence
 * search, residual + Jacobian, reduction to the normal equations -- a correspondence and an
 * accumulate kernel on the default (list) path, one fused kernel on the tile-stream path.
 * Replaces ICP/PlaneICP/VPlaneICP/NDT.calc_H_g_e2 (icp.py:24, plane_icp.py:30,
 * voxelized_plane_icp.py:23, ndt.py:24).  With a communicator attached (pcr_comm_init_rank)
 * the record is summed over all ranks before it is returned. */
int pcr_linearize(pcr_ctx* ctx, int method, const double T[16], double max_dist, double out[PCR_RECORD_LEN]);

/* calc_H_g_e2(cur_T, source) with `source` in HOST memory, as one call: pcr_set_scan_posed + pcr_linearize
 * with the host->device copy cut into chunks that overlap the ordering and the kernels of the chunk before
 * (the end-to-end step is bound by that copy).  The scan becomes the resident scan, as after pcr_set_scan_posed.
 * Small scans, device pointers, multi-GPU contexts and the tile-stream path take the two calls internally.
 * Replaces the same reference lines as pcr_linearize. */
int pcr_linearize_host(pcr_ctx* ctx, int method, const double T[16], double max_dist, const float* xyz, int64_t n, int sort,
                       double out[PCR_RECORD_LEN]);

/* Whole Gauss-Newton solve on the device: linearise, 6x6 solve, stop test BEFORE the update
 * (quirk Q8), T <- T [+] dx, without a host round trip per iteration.  Replaces
 * Registration.align (registration.py:71-113).  `e2_trace` (may be NULL) receives the squared
 * error of each executed linearisation (at most max_iter values); `iters` their number.
 * Returns PCR_ERR_SINGULAR if a singular H was met (T_out = last transform). */
int pcr_align(pcr_ctx* ctx, int method, const double T0[16], int max_iter, double tol, double max_dist,
              double T_out[16], int* iters, double* e2_trace);

/* The same loop, piecewise and asynchronous (used by pcr_align itself and by the benchmark to
 * time single iterations with CUDA events on pcr_stream): begin resets the device loop state
 * to T0; step_async enqueues `reps` fused iterations without synchronising (iterations after
 * convergence are no-ops); state synchronises and reads the loop state back (`done`: 0 running,
 * 1 converged, 2 singular, 3 max_iter reached). */
int pcr_loop_begin(pcr_ctx* ctx, const double T0[16]);
int pcr_loop_step_async(pcr_ctx* ctx, int method, int max_iter, double tol, double max_dist, int reps);
int pcr_loop_state(pcr_ctx* ctx, double T_out[16], int* iters, int* done, double* e2_trace, int trace_cap);

/* ---- utilities with reference-visible semantics --------------------------------------- */

/* Exact k-NN of m query points against the target points: Euclidean distances (m,k)
 * ascending, indices into the caller's target order; slots beyond the number of target
 * points get dist = inf, idx = n.  Replaces KDTree.query (kdtree.py:18-25). */
int pcr_knn(pcr_ctx* ctx, const float* queries, int64_t m, int k, float* dist, int64_t* idx);

/* Nearest kept voxel MEAN for each query (quirk Q2): voxel ordinal (into pcr_get_voxels
 * order) and float64 distance.  Replaces VoxelGrid.query (voxel.py:171-179). */
int pcr_voxel_query(pcr_ctx* ctx, const float* queries, int64_t m, int64_t* vidx, double* dist);

/* Per-voxel centroid down-sampling (voxel.py:209-241): writes at most n centroids
 * (float32) to `out` and their number to n_out.  Voxel order = library's own. */
int pcr_voxel_filter(pcr_ctx* ctx, const void* xyz, int64_t n, int is_f64, double voxel_size,
                     float* out, int64_t* n_out);

/* Voxel membership of every point (voxel.py:183-206, color_by_voxel's np.unique(..., return_inverse)):
 * labels[i] = ordinal of point i's voxel (library order), coords[3 v .. 3 v + 2] = integer
 * coordinate floor(p / voxel_size) of voxel v (room for n voxels), n_voxels = their number. */
int pcr_voxel_labels(pcr_ctx* ctx, const void* xyz, int64_t n, int is_f64, double voxel_size,
                     int64_t* labels, int32_t* coords, int64_t* n_voxels);

/* Per-correspondence Gauss-Newton rows at transform T for the scalar-residual methods (PCR_PLANE,
 * PCR_VPLANE): rows is (n_scan, 28) float64, C-contiguous, HOST or DEVICE memory, in the scan's storage
 * order (upload with sort = 0 to keep the caller's order): [0..20] upper triangle of J^T J, [21..26] J r,
 * [27] r^2 of every scan point (zeros without correspondence); their column sums are the record of
 * pcr_linearize.  Equals caratheodory.create_gn_set(J, r).T of the reference (caratheodory.py:118-138),
 * the input of its exact coreset extraction. */
int pcr_export_gn_rows(pcr_ctx* ctx, int method, const double T[16], double max_dist, double* rows);

/* ---- multi-GPU (scan tile-sharded, target replicated) ----------- */

/* 128-byte NCCL unique id, created on rank 0 and shipped to the other ranks by the caller. */
int pcr_comm_unique_id(void* id128);
/* Attach this context to a communicator of `nranks` processes (one GPU each).  After this
 * pcr_linearize / pcr_align all-reduce the 29-double record (ncclSum, float64). */
int pcr_comm_init_rank(pcr_ctx* ctx, int nranks, int rank, const void* id128);
int pcr_comm_destroy(pcr_ctx* ctx);
/* Hand the communicator of `src` over to `dst` (same device): a registration object whose target is
 * replaced gets a NEW context, and an NCCL unique id cannot be used twice.  `src` becomes single-GPU. */
int pcr_comm_move(pcr_ctx* dst, pcr_ctx* src);

/* ---- instrumentation -------------------------------------------------------------------- */

/* CUDA-event duration (ms) of the kernels launched by the last pcr_linearize / pcr_align. */
int pcr_last_kernel_ms(pcr_ctx* ctx, float* ms);
/* Total number of kernels launched by this context so far. */
int pcr_launch_count(pcr_ctx* ctx, int64_t* launches);
/* Raw CUDA stream of the context (cudaStream_t as void*) for callers that time with events. */
int pcr_stream(pcr_ctx* ctx, void** stream);
/* Enqueue `reps` linearisations back to back WITHOUT host synchronisation (benchmark aid:
 * lets the caller bracket them with its own CUDA events on pcr_stream). */
int pcr_linearize_async(pcr_ctx* ctx, int method, const double T[16], double max_dist, int reps);
/* Voxel-mean correspondences are read from exact per-cell candidate lists built with the voxels
 * (default on); 0 falls back to the general grid search everywhere (A/B and test hook). */
int pcr_set_voxel_lists(pcr_ctx* ctx, int enable);
int pcr_voxel_list_stats(pcr_ctx* ctx, int64_t* band_cells, int64_t* entries);
/* Opt-in (PCR_VOXEL_SHELL=1 when the voxels are built): the correspondence pass of VPlaneICP / NDT
 * streams margin-ordered shell lists over the kept voxel means (the structure pcr_shell_list_stats
 * describes for target points) instead of the exact candidate lists above.  Measured: equal on late
 * iterations, slower on the first ones (profiles/r2_notes.md). */
int pcr_voxel_shell_stats(pcr_ctx* ctx, int64_t* band_cells, int64_t* entries, double* margin_cells);
/* Target-point correspondences (ICP / PlaneICP) are read from per-cell "shell lists" built with
 * the NN index (every point within `margin` cell edges of the cell, ordered by its distance to
 * the cell; default on, margin 2 reduced until the lists fit the memory cap); 0 walks the brick
 * grid instead (A/B and test hook).  Same exact nearest neighbour either way.
 * PCR_SHELL_LISTS=0 / PCR_SHELL_DMAX / PCR_SHELL_MAX_GIB tune the build. */
int pcr_set_shell_lists(pcr_ctx* ctx, int enable);
int pcr_shell_list_stats(pcr_ctx* ctx, int64_t* band_cells, int64_t* entries, double* margin_cells);
/* Which implementation of the per-iteration path runs (both exact, same results up to float64
 * summation order):
 *   0 = list stream (default): correspondences from per-cell shell lists / voxel candidate lists
 *       (correspond kernel) + accumulate kernel;
 *   1 = tile stream: one fused kernel per linearisation; every warp stages the cell box around its
 *       32 scan points from a row-major dense grid into shared memory with cp.async.bulk (TMA engine,
 *       mbarrier completion) and every lane scans every staged candidate with packed f32x2 arithmetic
 *       (csrc/pcr_tile.cuh, pcr_tile_kernel.cuh).  Needs no per-point list structure (16 B per target
 *       point instead of ~1.4 KB) but is measured 1.6-2x slower per iteration on the round-2 workloads
 *       (profiles/r2_notes.md); kept selectable for targets whose lists do not fit.
 * Structures are built for the path that is selected when set_target runs: choose it right after
 * pcr_create (or with PCR_PATH=tile in the environment). */
int pcr_set_path(pcr_ctx* ctx, int path);
/* Row grid of the tile-stream path (which: 0 target points, 1 kept voxel means): cell edge, number
 * of cells of the dense table, occupied cells, bytes of the whole structure. */
int pcr_tile_stats(pcr_ctx* ctx, int which, double* cell_edge, int64_t* cells, int64_t* occupied, int64_t* bytes);
/* 1: every linearisation also parks the matched positions for pcr_debug_matches (4 B per scan
 * point of extra traffic on the tile-stream path; the list kernels always park them). */
int pcr_set_record_matches(pcr_ctx* ctx, int enable);
/* Test hook: correspondences parked by the LAST linearisation (any search variant): per resident
 * scan point, in storage order (upload with sort <= 0 to keep the caller's order), the caller
 * index of the matched target point (which = 0) / kept voxel (which = 1) or -1. */
int pcr_debug_matches(pcr_ctx* ctx, int which, int64_t* idx);
/* Grid statistics of the NN indices (cells, bricks, cell edge) for diagnostics. */
int pcr_index_stats(pcr_ctx* ctx, int which /*0 target, 1 voxel*/, double* cell_edge, int64_t* n_cells,
                    int64_t* n_bricks, int64_t* n_points);

#ifdef __cplusplus
}
#endif
#endif /* PCR_B200_H */
