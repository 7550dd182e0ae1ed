/*
 * libtriangl_cuda -- C ABI of the B200 (sm_100a) batched two-view triangulation hot path.
 *
 * Each entry point replaces one native/third-party call of the reference
 * (Eliasvan/Multiple-Quadrotor-SLAM, paths relative to the reference root):
 *
 *   trgl_linear_ls      <- triangulation_ext.linear_LS_triangulation(u1,P1,u2,P2,x)
 *                          Work/python_libs/triangulation_c/triangulation.c:48-83
 *                          (Python fallback Work/python_libs/triangulation.py:31-94)
 *   trgl_iterative_ls   <- triangulation_ext.iterative_LS_triangulation(u1,P1,u2,P2,tolerance,x,x_status)
 *                          triangulation.c:85-161   (Python fallback triangulation.py:100-195)
 *   trgl_linear_eigen   <- cv2.triangulatePoints + dehomogenise + finite mask, triangulation.py:6-25
 *   trgl_polynomial     <- cv2.invert / F=[t]xR / cv2.correctMatches / linear_eigen, triangulation.py:198-232
 *   trgl_reproj_error   <- cv2.projectPoints + RMS, Work/python_libs/calibration_tools.py:116-124 (and :89-113)
 *   trgl_pair_reproj    <- the harness' re-projection of x into both cameras and status>0 good mask,
 *                          Work/triangulation_comparison/triangulation_comparison.py:190-217,242-260
 *
 * Conventions (same as the reference's weave boundary, triangulation_c/__init__.py:18-86):
 *   - caller allocates every output, callee fills it in place, inputs are never modified;
 *   - u1,u2: (n,2) row-major normalised image coordinates; x: (n,3) row-major; P1,P2: 12 doubles =
 *     rows 0..2 of the row-major 3x4 / 4x4 camera matrix (row stride 4), always HOST memory;
 *   - all functions return 0 on success or a negative TRGL_E_* / positive cudaError_t code; they never
 *     abort and never throw.  NaN/Inf inputs propagate into x / status like the reference.
 *   - mem = TRGL_MEM_HOST: u1,u2,x,status are host buffers, the library stages them through the GPU
 *     (H2D + kernel + D2H, chunked and overlapped; pinned buffers from trgl_host_alloc go at full PCIe speed);
 *     mem = TRGL_MEM_DEVICE: they are device pointers on the current device, the kernel is enqueued on
 *     `stream` (a cudaStream_t, NULL = default stream) and the call returns without synchronising.
 *   - Re-entrancy: device-mode calls keep no shared mutable state (reduction scratch is owned per device and stream), so
 *     different host threads may drive different streams concurrently, like the stack-local reference code
 *     (triangulation.c:67,106); host-mode calls share the staging pipeline and are serialised by an internal mutex.
 *   - There is NO CPU fallback: without a CUDA device every compute entry point returns an error.
 *
 * Precision modes (`mode`):
 *   TRGL_F64       u,x float64, float64 arithmetic            (reference arithmetic, triangulation.c)
 *   TRGL_F32IO     u,x float32, float64 arithmetic            (the SLAM convention: slam2.py:19,551-555)
 *   TRGL_F32       u,x float32, float32 arithmetic            ("FP32 mode" of BASELINE.json).  Only linear_LS, the
 *                  HBM-bound solver, has float32 arithmetic; iterative_LS (absolute 3e-5 depth tolerance at depth ~40
 *                  is 6 float ulps), linear_eigen and polynomial (smallest singular vector, degree-6 coefficients)
 *                  keep float64 registers and only halve the bytes moved -- for them TRGL_F32 == TRGL_F32IO.
 *                  linear_LS solves a correspondence in float32 when its kappa^2 bound is below 300 (error <= 2e-5) and
 *                  in float64 otherwise; a batch in which more than 1/8 of the correspondences are beyond that bound is
 *                  redone in float64 as a whole.  Results are within 1e-4 of the float64 solution either way, but their
 *                  last bits may depend on the batch a correspondence is in (the float64 modes have no such dependence).
 *   TRGL_F64_OUT32 u float64, x float32, float64 arithmetic   (output_dtype=float32 on float64 inputs)
 *   TRGL_F32_OUT64 u float32, x float64, float64 arithmetic   (float32 inputs, default output dtype)
 */
#ifndef TRIANGL_CUDA_H
#define TRIANGL_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TRGL_VERSION 105

enum { TRGL_F64 = 0, TRGL_F32IO = 1, TRGL_F32 = 2, TRGL_F64_OUT32 = 3, TRGL_F32_OUT64 = 4 };
enum { TRGL_MEM_HOST = 0, TRGL_MEM_DEVICE = 1, TRGL_MEM_DEVICE_IN = 2 };
enum { TRGL_ITER_C = 0, TRGL_ITER_PY = 1 };          /* iterative_LS control flow: triangulation.c vs triangulation.py */

enum {
    TRGL_OK = 0,
    TRGL_E_BADARG = -1,        /* NULL pointer with n > 0, unknown mode / mem / rows */
    TRGL_E_NODEVICE = -2,      /* no usable CUDA device */
    TRGL_E_NOMEM = -3
};

/* ---- library / device management ---- */
int trgl_version(void);
const char* trgl_last_error_string(void);
int trgl_device_count(void);
int trgl_set_device(int device);
int trgl_device_synchronize(void);

/* ---- memory and stream helpers (so a ctypes host needs no other CUDA binding) ---- */
int trgl_device_alloc(void** ptr, size_t bytes);
int trgl_device_free(void* ptr);
int trgl_host_alloc(void** ptr, size_t bytes);           /* pinned */
int trgl_host_free(void* ptr);
int trgl_memcpy_h2d(void* dst, const void* src, size_t bytes, void* stream);
int trgl_memcpy_d2h(void* dst, const void* src, size_t bytes, void* stream);
int trgl_memcpy_d2d(void* dst, const void* src, size_t bytes, void* stream);
int trgl_memset_d(void* dst, int value, size_t bytes, void* stream);
int trgl_stream_create(void** stream);
int trgl_stream_destroy(void* stream);
int trgl_stream_synchronize(void* stream);
int trgl_event_create(void** event);
int trgl_event_destroy(void* event);
int trgl_event_record(void* event, void* stream);
int trgl_event_synchronize(void* event);
int trgl_event_elapsed_ms(void* start, void* stop, float* ms);   /* synchronises on `stop` */

/* ---- the four solvers ---- */

/* status: n bytes, always 1 (triangulation.c:65-83 writes none; the wrapper returns np.ones(bool)). */
int trgl_linear_ls(const void* u1, const void* u2, const double* P1, const double* P2,
                   void* x, uint8_t* status, int64_t n, int mode, int mem, void* stream);

/* status: n int32 in {1,0,-1,-2,-3} (triangulation.c:154-159).  semantics: TRGL_ITER_C / TRGL_ITER_PY. */
int trgl_iterative_ls(const void* u1, const void* u2, const double* P1, const double* P2,
                      void* x, int32_t* status, int64_t n, double tolerance, int semantics,
                      int mode, int mem, void* stream);

/* status: n bytes, max|x| <= max_coordinate_value (NaN/Inf -> 0).  rows: 4 (OpenCV >= 3) or 6 (OpenCV 2.4). */
int trgl_linear_eigen(const void* u1, const void* u2, const double* P1, const double* P2,
                      void* x, uint8_t* status, int64_t n, double max_coordinate_value, int rows,
                      int mode, int mem, void* stream);

/* Hartley-Sturm correction with F derived from P1,P2 (triangulation.py:211-216), then linear_eigen.
 * u1_corr/u2_corr (n,2, dtype of u) may be NULL; when given they receive the corrected matches.
 * all_nan (host int*, may be NULL): set to 1 when every corrected point is NaN, i.e. when the reference would
 * take its findFundamentalMat fallback (triangulation.py:227-229); the caller then re-runs with
 * trgl_polynomial_F.  Writing all_nan synchronises the stream. */
int trgl_polynomial(const void* u1, const void* u2, const double* P1, const double* P2,
                    void* x, uint8_t* status, void* u1_corr, void* u2_corr, int* all_nan, int64_t n,
                    double max_coordinate_value, int rows, int mode, int mem, void* stream);

/* Same with an explicit fundamental matrix F (9 doubles, row-major, x2^T F x1 = 0). */
int trgl_polynomial_F(const void* u1, const void* u2, const double* P1, const double* P2, const double* F,
                      void* x, uint8_t* status, void* u1_corr, void* u2_corr, int* all_nan, int64_t n,
                      double max_coordinate_value, int rows, int mode, int mem, void* stream);

/* Whole-batch normalised 8-point F (cv2.findFundamentalMat(u1,u2,FM_8POINT), triangulation.py:228): device
 * reductions of the moments and of the 9x9 normal matrix, tiny eigen-solve on the host.  F: 9 doubles out (host).
 * Only the element type of u is taken from `mode`.  Synchronises. */
int trgl_fundamental_8point(const void* u1, const void* u2, int64_t n, int mode, int mem, double* F, void* stream);

/* Asynchronous variant for device-resident pipelines: same kernel, but the grid-level sum is finished inside the kernel
 * (last-block reduction in block order, so the four sums are bit-identical to the synchronous variant) and written to
 * sums_device (4 doubles in DEVICE memory).  Nothing is copied back and the stream is not synchronised: the call only
 * enqueues, and the result is owned by whatever the caller enqueues next on `stream`.  Device buffers only. */
int trgl_pair_reproj_async(const void* x, const void* u1, const void* u2, const double* P1, const double* P2,
                           const void* status, int status_is_i32, int min_status, double max_sq_err,
                           void* err1, void* err2, uint8_t* good, double* sums_device,
                           int64_t n, int mode, void* stream);

/* The same evaluation FUSED into the solver kernel: request it with trgl_set_fused_eval, then make a device-mode solver
 * call (any of the four, or their _px twins) from the same host thread; the request applies to that one call.  The kernel
 * evaluates every point from the registers the solve just produced -- x rounded to its storage type exactly as a separate
 * pass would read it back, the normalised observations, the status just written, the call's own P1 / P2 -- so err1, err2
 * and good are bit-identical to trgl_pair_reproj, while the second pass over x, u1, u2 and status (57-60 B/point)
 * disappears.  sums_device: 4 doubles in DEVICE memory (same meaning as `sums` above), finished inside the kernel by a
 * last-block reduction; nothing is synchronised.  err1 / err2 (n elements of x's dtype) and good (n bytes) may be NULL.
 * Supported where u and x have the same storage type (TRGL_F64, TRGL_F32IO, TRGL_F32); device buffers only. */
int trgl_set_fused_eval(int min_status, double max_sq_err, void* err1, void* err2, uint8_t* good, double* sums_device);

/* ---- input normalisation in front of the solvers ---- */

/* Multi-view linear least-squares triangulation: the 2m x 3 generalisation of
 * trgl_linear_ls for a point seen by m >= 2 cameras (the multi-quadrotor scene).  The reference has no such call -- every
 * triangulation in Work/ is two-view (triangulation.py:31-94) -- so this extends its interface in its own conventions:
 *   u      : (m, n, 2) observations in normalised coordinates, view-major (each view a contiguous (n,2) array)
 *   valid  : (m, n) uint8, 1 = view v observes point i; NULL = every view observes every point
 *   P      : (m, 12) rows 0-2 of the camera matrices (host doubles)
 *   x      : (n,3) minimum-norm least-squares solution of the rows of the observing views (cvSolve(DECOMP_SVD) rule);
 *            a point nobody observes gets (0,0,0)
 *   status : (n,) uint8, 1 when at least min_views views observe the point
 * With m = 2, valid = NULL this is trgl_linear_ls.  Arithmetic is float64 (TRGL_F32 runs as TRGL_F32IO). */
int trgl_multiview_ls(const void* u, const uint8_t* valid, const double* P, int m, void* x, uint8_t* status, int64_t n,
                      int min_views, int mode, int mem, void* stream);

/* cv2.undistortPoints(src, K, dist) with the default criteria (5 fixed-point iterations) and no R / P:
 * call sites Work/SLAM/application/own/slam2.py:551-552, Work/triangulation_comparison/triangulation_comparison.py:164-173,
 * Work/calibration/calibrate.py:252-253.  src, dst: (n,2) pixel / normalised coordinates, both float32 (in_is_f32 = 1)
 * or both float64 -- cv2 returns its input dtype.  K: 9 doubles row-major (host); dist: (k1,k2,p1,p2,k3) (host) or NULL.
 * Arithmetic is float64 in OpenCV's evaluation order without FMA contraction: results are bit-identical to cv2 4.13. */
int trgl_undistort_points(const void* src, void* dst, const double* K, const double* dist, int64_t n, int in_is_f32,
                          int mem, void* stream);

/* The all-NaN test of the LAST device-mode polynomial call on `stream` (np.isnan(u_new).all(), triangulation.py:227),
 * without blocking: copies the two "some corrected point of view 1 / view 2 is not NaN" words to host_flags2 (page-locked
 * host memory, 2 x uint32) in stream order and returns; the caller reads them after synchronising with the stream or an
 * event (all_nan = either word == 0).  For callers that keep several batches in flight instead of passing `all_nan`,
 * which synchronises inside the call. */
int trgl_polynomial_flags_async(unsigned int* host_flags2, void* stream);

/* The four solvers on PIXEL coordinates: px1, px2 are what the callers hand to cv2.undistortPoints, (K1,dist1) and
 * (K2,dist2) the intrinsics of the two views (the reference uses one camera for both, slam2.py:551-552).  The
 * undistortion runs in registers in front of the solve (one HBM round trip of u1,u2 less); the normalised pair is
 * rounded to the dtype of px exactly as cv2.undistortPoints' output would be, so every *_px call returns bit for bit
 * what trgl_undistort_points followed by the plain solver returns.  All other arguments as in the plain solvers. */
int trgl_linear_ls_px(const void* px1, const void* px2, const double* K1, const double* dist1, const double* K2,
                      const double* dist2, const double* P1, const double* P2, void* x, uint8_t* status, int64_t n,
                      int mode, int mem, void* stream);
int trgl_iterative_ls_px(const void* px1, const void* px2, const double* K1, const double* dist1, const double* K2,
                         const double* dist2, const double* P1, const double* P2, void* x, int32_t* status, int64_t n,
                         double tolerance, int semantics, int mode, int mem, void* stream);
int trgl_linear_eigen_px(const void* px1, const void* px2, const double* K1, const double* dist1, const double* K2,
                         const double* dist2, const double* P1, const double* P2, void* x, uint8_t* status, int64_t n,
                         double max_coordinate_value, int rows, int mode, int mem, void* stream);
int trgl_polynomial_px(const void* px1, const void* px2, const double* K1, const double* dist1, const double* K2,
                       const double* dist2, const double* P1, const double* P2, void* x, uint8_t* status,
                       void* u1_corr, void* u2_corr, int* all_nan, int64_t n, double max_coordinate_value, int rows,
                       int mode, int mem, void* stream);

/* ---- fused reprojection error / good-point mask ---- */

/* cv2.projectPoints(x, rvec, tvec, K, dist) then sum of squared residuals against imgp (n,2).
 * K: 9 doubles row-major; dist: 5 doubles (k1,k2,p1,p2,k3); rvec,tvec: 3 doubles (all host).
 * x_is_f32 / img_is_f32 select the element type of x and of imgp/proj.  proj (n,2) may be NULL.
 * sums (host, 3 doubles): sum dx^2, sum dy^2, count of finite residuals; rms = sqrt((s0+s1)/n)
 * (calibration_tools.py:124 divides by the number of points).  abs_sums (host, 2 doubles, may be NULL):
 * sum |dx|, sum |dy| for reprojection_error_ext (calibration_tools.py:107).  Synchronises the stream. */
int trgl_reproj_error(const void* x, const void* imgp, const double* K, const double* dist,
                      const double* rvec, const double* tvec, void* proj, double* sums, double* abs_sums,
                      int64_t n, int x_is_f32, int img_is_f32, int mem, void* stream);

/* Two-view evaluation right after a solver call: re-project x through P1 and P2 (normalised cameras),
 * per-point squared reprojection errors against u1,u2 (err1, err2: n elements of x's dtype, may be NULL),
 * good[i] = (status[i] > min_status) && err1[i] <= max_sq_err && err2[i] <= max_sq_err && both depths > 0
 * (good: n bytes, may be NULL).  status may be uint8 (status_is_i32 = 0) or int32.
 * sums (host, 4 doubles): sum err1, sum err2 over good points, number of good points, number of points with
 * status > min_status.  Synchronises the stream. */
int trgl_pair_reproj(const void* x, const void* u1, const void* u2, const double* P1, const double* P2,
                     const void* status, int status_is_i32, int min_status, double max_sq_err,
                     void* err1, void* err2, uint8_t* good, double* sums,
                     int64_t n, int mode, int mem, void* stream);

/* ---- harness statistics on the device ---- */

/* error_vectors_3D + error_rms + robustness_stat of the comparison harness
 * (Work/triangulation_comparison/triangulation_comparison.py:179-188, 205-217, 242-260):
 * errors[i] = |x[i] - exact[i, 0:3]|^2 (n doubles, may be NULL; same memory space as x), exact: (n, exact_stride) doubles
 * (stride 4 for the harness' homogeneous cloud).  status (uint8 / int32, may be NULL) feeds robustness_stat with the
 * thresholds robustness_thresh_max / _min (:372-373).  stats (host, 4 doubles): sum of errors, number of NaN errors,
 * number of false positives (error > thresh_max and status > 0), number of false negatives (error <= thresh_min and
 * not status > 0); divide by n for the reference's ratios.  Synchronises the stream. */
int trgl_eval_errors_3d(const void* x, const double* exact, int exact_stride, const void* status, int status_is_i32,
                        double thresh_max, double thresh_min, double* errors, double* stats, int64_t n, int x_is_f32,
                        int mem, void* stream);
/* error_rms on 2-D residuals: errors[i] = |proj[i] - exact[i]|^2 with proj from trgl_reproj_error (pixels) and the exact
 * pixel positions (error_vectors_2D, :190-203).  stats (host, 4 doubles): sum of errors, number of NaN errors, 0, 0. */
int trgl_eval_errors_2d(const void* proj, const double* exact, double* errors, double* stats, int64_t n, int proj_is_f32,
                        int mem, void* stream);
/* vector_stat of the comparison harness (Work/triangulation_comparison/triangulation_comparison.py:219-240, used at
 * :483-487 for the last pose of every trajectory): x_trials is the (trials, n, 3) array of the solver results of all
 * repetitions (errors_partitioned + exact), exact the (n, exact_stride) cloud; means (n,3) and covars (n,3,3) doubles
 * receive, per point, the mean of its 3-D error vectors x[t,i,:] - exact[i,0:3] over the trials and their population
 * covariance (divisor = trials, like the reference).  mem = DEVICE: all five arrays on the device, enqueued on stream,
 * no synchronisation; mem = HOST: staged through the GPU. */
int trgl_vector_stat(const void* x_trials, const double* exact, int exact_stride, int trials, double* means, double* covars,
                     int64_t n, int x_is_f32, int mem, void* stream);

/* np.median of n non-negative doubles (the squared errors above), exact: most-significant-digit radix selection on the
 * IEEE bit patterns (8 histogram passes), mean of the two middle elements for even n, NaN if any element is NaN or
 * n == 0.  median: host double.  Synchronises the stream. */
int trgl_median(const double* values, int64_t n, int mem, double* median, void* stream);

/* ---- upload once, solve many (host results from device-resident observations) ---- */

/* The reference's callers run several solvers on the SAME (u1, u2) (the comparison harness:
 * Work/triangulation_comparison/triangulation_comparison.py:466-469 loops the four methods over one noisy observation set;
 * slam2.py:553-584 re-triangulates a subset).  In plain host mode every call uploads u1 / u2 again.  Two pieces avoid it:
 *   trgl_set_input_retention(d_u1, d_u2): the NEXT host-mode solver call of this thread (mem = TRGL_MEM_HOST) uploads its
 *     u1 / u2 chunk by chunk into these caller-owned device buffers (n x 2 elements of the input dtype each) instead of
 *     its pipeline scratch -- same overlap of H2D, kernel and D2H -- and leaves them there;
 *   mem = TRGL_MEM_DEVICE_IN on any solver entry: u1 / u2 are device pointers (e.g. the retained buffers), x / status
 *     (and the other outputs) are HOST buffers; kernels and D2H copies are pipelined per chunk, nothing is uploaded.
 * Four solvers on 10 M float64 correspondences then move 0.32 GB host-to-device instead of 1.28 GB. */
int trgl_set_input_retention(void* u1_device, void* u2_device);

/* ---- multi-GPU: the result gather fused into the solver's stores ---- */

/* One process per GPU.  Every rank allocates the gathered arrays x_all (N_total,3) and status_all (N_total,) with
 * trgl_device_alloc, exports them (trgl_ipc_export -> 64-byte CUDA IPC handle), exchanges the handles out of band
 * (torch.distributed / MPI / a pipe) and maps its peers' arrays with trgl_ipc_import (peer access over NVLink / NVSwitch
 * is enabled by the mapping).  Before a solver call on its shard [lo, lo+n) a rank then passes, for each peer,
 * the addresses of that shard inside the peer's arrays:
 *     x_mirrors[r] = peer_r.x_all + 3*lo (elements of x's dtype),  status_mirrors[r] = peer_r.status_all + lo
 * with trgl_set_result_mirrors (at most 8 entries: 7 peers + optionally a second local copy).  The table applies to the NEXT device-mode solver call of the calling
 * host thread (trgl_linear_ls / iterative_ls / linear_eigen / polynomial and their _px twins) and is cleared by it.
 * That kernel repeats every store of x and status to all mirrors, so when it has finished on every rank (stream
 * synchronise + a barrier of the caller's choice) each rank holds the complete result -- no all-gather pass, no second
 * read of x from HBM, and the NVLink traffic overlaps the solve.  The reference has no multi-process code (SURVEY.md
 * F10); this replaces the ncclAllGather of SURVEY.md 8e. */
int trgl_set_result_mirrors(void* const* x_mirrors, void* const* status_mirrors, int count);
/* Same, but the mirrors hold FLOAT32 rows whatever the dtype of the rank's own x (x_mirrors[r] = peer_r.x_all + 3*lo in
 * float32 elements): the gathered map in the dtype the reference's SLAM keeps it in
 * (Work/SLAM/application/own/slam2.py:19, set_triangl_output_dtype(np.float32)) at 12 + 1|4 instead of 24 + 1|4 bytes
 * per point over NVLink -- the gather is link-bound.  Each mirror value is the float64 result
 * rounded once, i.e. exactly what `.astype(np.float32)` (triangulation.py:242-243) makes of it. */
int trgl_set_result_mirrors_f32(void* const* x_mirrors, void* const* status_mirrors, int count);
int trgl_ipc_export(void* device_ptr, void* handle64);            /* device_ptr from trgl_device_alloc */
int trgl_ipc_import(const void* handle64, void** device_ptr);     /* in ANOTHER process than the exporter */
int trgl_ipc_close(void* device_ptr);

/* ---- bench / diagnostics ---- */
/* Number of kernels this library has launched since load (for bench.py's gpu_launches). */
int64_t trgl_launch_count(void);
/* Tuning knob: input path of linear_LS.  -1 = auto (default: 0 for float64 arithmetic, 8 for TRGL_F32), 0 = per-thread
 * vector loads; 1..6 = cp.async.bulk (TMA engine) shared-memory ring with (points per thread, stages) = (2,4) (4,3) (1,6)
 * (2,6) (4,4) (1,8); 7..12 = per-thread cp.async (LDGSTS) ring with (points per thread, stages, min CTAs/SM) = (1,8,2)
 * (2,4,2) (4,2,2) (4,3,2) (2,4,3) (2,6,2).  Measured on B200 in profiles/.  Returns the previous value. */
int trgl_set_stream_variant(int variant);
/* Tuning knob: points per thread of the per-thread-load linear_LS kernel (1, 2 or 4, default 4); returns the previous value. */
int trgl_set_points_per_thread(int ppt);
/* Tuning / test knob: the two-ray closed forms.  1 = on (default): iterative_LS runs the closed form of the re-weighted
 * solve for every correspondence it certifies (finite camera centres, kappa^2 bound below 1e10, weight ratio in range)
 * and polynomial takes the certified intersection of the two viewing rays of the corrected match; everything else goes
 * through the reference's arithmetic as written (triangulation.c:104-161 re-weighted normal equations / SVD tiers;
 * cv2.triangulatePoints' smallest singular vector).  0 = the reference's arithmetic for every correspondence.  Both give
 * the same status vectors and points within 1e-9.  Returns the previous value. */
int trgl_set_two_ray(int enabled);
/* Test knob: the hot kernels of iterative_LS, linear_eigen, polynomial and multiview_LS hand the correspondences they do
 * not certify to a follow-up kernel through a list of one slot per point of the batch (at most 2^26); when more points
 * are deferred than the list holds, the follow-up kernel redoes every point instead.  This limits the list to
 * max_points (1 .. 2^26) so that the overflow path can be exercised on small batches.  Returns the previous limit. */
int64_t trgl_set_deferred_capacity(int64_t max_points);
/* Diagnostics: running total of the correspondences the hot kernels have handed to their follow-up kernels on this
 * (current device, stream) since the library was loaded (bench.py reports the deferred fraction per solver with it).
 * Synchronises the stream. */
int trgl_deferred_total(void* stream, int64_t* total);
/* Diagnostics: how often the rare paths of polynomial's correction ran on the current device since load (or the last reset):
 * out5 = {points through the certified real-root isolation, dyadic intervals it visited, points it could not certify,
 * points through Durand-Kerner (cv::solvePoly's method, ~57 000 instructions each), most intervals visited for one point}.
 * Synchronises the device. */
int trgl_rare_path_counters(unsigned long long* out5, int reset);
/* Diagnostics: FP64 FMA rate of the current device, measured (SURVEY.md section 8d asks for the FP64 peak from an FMA
 * microbenchmark in the same run; the reference has no counterpart).  Every thread of ctas_per_sm x SMs CTAs of 256 threads
 * runs `chains` (1, 2 or 8) independent chains x = fma(y, z, x); operands = 2: z comes from the constant bank (two
 * register sources per instruction, the pipe's nominal 2 cycles per warp instruction), operands = 3: three distinct
 * 64-bit register sources (3 cycles: register-file banking, csrc/trgl_probe.cuh).  Result: warp-level FMA instructions
 * per second over the whole device (x 64 = flop/s).  Synchronises the device; ~10 ms. */
int trgl_fp64_fma_rate(int operands, int chains, int ctas_per_sm, double* warp_fma_per_second);
/* Diagnostics of the small-batch host path (n <= 32 Ki points: the SLAM keyframe sizes).  trgl_set_trace(1) makes every such
 * call add its host-side microseconds per phase to five accumulators -- staging memcpy in, kernel launches, stream
 * synchronise, memcpy out, number of calls -- which trgl_get_trace copies to out5 and clears.  Returns the previous setting. */
int trgl_set_trace(int enabled);
int trgl_get_trace(double* out5);

#ifdef __cplusplus
}
#endif
#endif
