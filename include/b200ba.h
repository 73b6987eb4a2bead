/* b200ba.h -- C ABI of the B200-native sparse bundle-adjustment engine (libb200ba.so).
 *
 * This is the drop-in boundary for SequenceSFM's BA entry point.  The reference interface it replaces
 * (paths relative to the reference tree):
 *
 *   int adjustBundle_ssba(vector<CloudPoint>&, Mat& K, map<int,vector<KeyPoint>>&, map<int,Matx34d>&)   src/sfm_ba.h:29-32
 *   int adjustBundle_pba (vector<CloudPoint>&, Mat& K, Mat& dist, map<...>&, map<...>&)                 src/sfm_ba.h:39-43
 *
 * whose bodies (src/sfm_ba_ssba.cpp:66-286, src/sfm_ba_pba.cpp:33-174) marshal the tracker state into flat
 * arrays and hand them to SSBA-3.0 (`CommonInternalsMetricBundleOptimizer(...).minimize()`,
 * Thirdparty/SSBA-3.0/Geometry/v3d_metricbundle.h:150-238) or PBA (`ParallelBA::SetCameraData / SetPointData /
 * SetProjection / RunBundleAdjustment`, Thirdparty/pba/pba/pba.h:103-108).  The functions below take the same
 * flat arrays.  adapter/sfm_ba_b200.cpp is the replacement body of the two C++ functions on top of this ABI;
 * INTEGRATION.md shows the build change.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success or a negative B200BA_E_* code and
 * never throws; host buffers are caller-owned and copied; one handle owns one CUDA device (and optionally one NCCL
 * rank); a handle is not re-entrant, distinct handles are independent.  There is NO CPU fallback: when no CUDA
 * device is usable b200ba_create fails with B200BA_E_CUDA.
 */
#ifndef B200BA_H
#define B200BA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200BA_VERSION 1

enum {
    B200BA_OK             =  0,
    B200BA_E_INVALID      = -1,   /* bad argument / problem (indices out of range, obs_pt decreasing, ...) */
    B200BA_E_CUDA         = -2,   /* CUDA runtime error, no device, out of memory */
    B200BA_E_STATE        = -3,   /* call order violated (e.g. solve before set_problem) */
    B200BA_E_COMM         = -4    /* NCCL unavailable or failed */
};

/* LM exit status, same numbering as SSBA (Thirdparty/SSBA-3.0/Math/v3d_optimization.h:16-21). */
enum {
    B200BA_STATUS_TIMEOUT      = 0,   /* max_iterations reached                          */
    B200BA_STATUS_SMALL_UPDATE = 1,   /* |delta| < 1e-8 (|params| + 1e-8) after >= min_iterations */
    B200BA_STATUS_CONVERGED    = 2,   /* |J^t e|_inf < gradient_threshold                */
    B200BA_STATUS_MSE          = 3    /* PBA variant only: mean squared error <= 0.25 px^2 (ConfigBA.cpp:47, return code 'M') */
};

typedef struct b200ba_options {
    int32_t struct_size;        /* = sizeof(b200ba_options); set by b200ba_default_options            */
    int32_t device;             /* CUDA device ordinal; -1 = current device                               */
    /* --- LM control, defaults = what src/sfm_ba_ssba.cpp:193-217 hard-codes --- */
    int32_t max_iterations;     /* 50                                                                     */
    int32_t min_iterations;     /* 10   (v3d_optimization.h:28)                                           */
    double  tau;                /* 1e-3 lambda0 = tau * max diag(J^t J)                                   */
    double  inlier_px;          /* 2.0  Huber threshold in pixels; <= 0 disables the robustifier          */
    double  gradient_threshold; /* 1e-8                                                                   */
    double  update_threshold;   /* 1e-8                                                                   */
    int32_t round_pixels;       /* 1: measurements rounded to whole pixels like `Point cvp = Point2f`     */
                                /*    (src/sfm_ba_ssba.cpp:171); 0: keep sub-pixel (src/sfm_ba_pba.cpp:122) */
    int32_t normalize;          /* 1: ScopedBundleExtrinsicNormalizer (v3d_metricbundle.h:327-375)        */
    int32_t fix_first_camera;   /* 0 (SSBA: all free); 1 = PBA's SetConstantCamera on camera 0            */
    int32_t discard_converged;  /* 1: replicate `good_adjustment = (status != 2)` (src/sfm_ba_ssba.cpp:221) */
    /* --- linear solver: block-Jacobi PCG on the reduced camera system --- */
    int32_t cg_max_steps;       /* 50                                                                     */
    double  cg_rel_tol;         /* 0: always run cg_max_steps; > 0: stop when sqrt(r.z / r0.z0) < tol     */
    int32_t precond;            /* 1: diagonal blocks of S; 0: (U + lambda I)^-1 like PBA                 */
    /* --- host interaction --- */
    int32_t poll_every;         /* copy the device status word back every n LM iterations (default 5)     */
    /* --- layout --- */
    int32_t lane_window;        /* 64: candidate window of the bank-conflict-aware grouping of points into     */
                                /*     quarter-warps for the shared-memory camera table; 0 = keep input order  */
    /* --- precision --- */
    int32_t pcg_fp32;           /* 0: everything fp64 (default).  1: optional mixed-precision mode - the two sweeps that apply  */
                                /*    W (V + lambda I)^-1 W^t inside the PCG loop run on fp32 copies of the per-observation / */
                                /*    per-point data; linearisation, preconditioner, PCG vectors and dot products, back-      */
                                /*    substitution, cost and the LM control stay fp64 (per-iteration cost within 1e-5)        */
    /* --- persistent handles --- */
    int32_t keep_problem;       /* 0 (default).  1: the handle keeps the marshalled observations so that b200ba_append can extend  */
                                /*    the problem by what a new keyframe adds                     */
    /* --- LM variant --- */
    int32_t lm_variant;         /* 0 (default): SSBA's LM (additive damping, x10 / :10, v3d_optimization.cpp:879-978).                */
                                /* 1: PBA's LM (NonlinearOptimizeLM, SparseBundleCPU.cpp:3772-3949): lambda0 = 1e-3, damping          */
                                /*    lambda diag(J^t J), accept iff the error drops, gain ratio reduction / (2 d.g - |J d|^2),       */
                                /*    lambda *= max(1/3, 1 - (2 rho - 1)^3) / *= 2, 4, 8 ..., stops on |delta| <= 1e-6 (column-scaled    */
                                /*    coordinates) and on MSE <= 0.25 px^2; PCG: cg_rel_tol 0.1 after >= cg_min_steps 10, <= 100, guard 1 */
                                /*    (b200ba_pba_lm_preset sets all of this); normalize = 2 = PBA's NormalizeData (depth scaling)    */
    int32_t cg_min_steps;       /* PCG steps before cg_rel_tol may stop it (PBA: 10; default 0)                                     */
    /* --- reduced camera system --- */
    int32_t reduced_system;     /* 0 (default): automatic - tracker-sized problems (up to ~300 cameras, single rank, fp64) assemble   */
                                /*    S = (U + damping) - W (V + damping)^-1 W^t explicitly once per LM iteration and run the whole  */
                                /*    PCG loop in ONE kernel with S resident in the shared memory of a thread-block cluster (or of   */
                                /*    the whole grid); larger problems apply S matrix-free in two sweeps per step.                    */
                                /* 1: always matrix-free.  2: explicit whenever it fits (same as 0 today)                             */
    int32_t reserved[1];
} b200ba_options;

#define B200BA_TRACE_COLS 8     /* iter, err, lambda, new_err, rho, accepted(1/0/-1), cg_steps, cg_rel_residual */

typedef struct b200ba_report {
    int32_t status;             /* B200BA_STATUS_*                                                        */
    int32_t iterations;         /* LM loop iterations executed (accepted + rejected)                      */
    int32_t ret;                /* what adjustBundle_ssba returns for this outcome: 0 or -1               */
    int32_t n_trace;            /* rows valid in trace                                                    */
    int32_t total_cg_steps;
    int32_t gpu_launches;       /* kernels of this library launched by the call                           */
    double  mean_px[2];         /* mean reprojection error in pixels before / after (sfm_ba_ssba.cpp:26-51) */
    double  initial_cost, final_cost;   /* sum of squared weighted residuals (normalised units)           */
    double  ms_solve;           /* device time of the LM loop (CUDA events)                               */
    double  ms_h2d, ms_d2h;     /* upload (set_problem) / download (get_result) wall time                 */
    double* trace;              /* optional caller buffer, trace_cap rows x B200BA_TRACE_COLS; may be NULL */
    int32_t trace_cap;
    int32_t iterations_total;   /* LM iterations / accepted steps summed over every b200ba_restart since b200ba_begin  */
    int32_t accepted_total;     /* (iterations counts the last run only)                                              */
    int32_t reserved[3];
} b200ba_report;

typedef struct b200ba_handle b200ba_handle;

/* Fill *opt with the SSBA-path defaults (src/sfm_ba_ssba.cpp:193-217). */
int b200ba_default_options(b200ba_options* opt);
/* PBA-path preset (src/sfm_ba_pba.cpp:60-62,94,136): camera 0 constant, no robustifier, sub-pixel measurements. */
int b200ba_pba_preset(b200ba_options* opt);
/* The same plus PBA's own LM variant, PCG stopping rule, U-block preconditioner and data normalisation (ConfigBA.cpp:42-118,
 * SparseBundleCPU.cpp:3117-3321,3323-3450,3772-3949): what a -DUSE_PBA build of the reference runs. */
int b200ba_pba_lm_preset(b200ba_options* opt);

int b200ba_create(const b200ba_options* opt, b200ba_handle** out);
void b200ba_destroy(b200ba_handle* h);
const char* b200ba_last_error(const b200ba_handle* h);   /* h may be NULL: error of the last failed create */

/* Upload one problem.  cam_RT: N x 12 row-major [R|T] world->camera (Pmats, src/sfm_ba_ssba.cpp:130-146);
 * K5 = {fx, skew, cx, fy, cy} (camIntrinsic, :90-94); pts: M x 3 (pointcloud[j].pt); obs_xy: K x 2 float pixels
 * (imgpts[frame][kp].pt); obs_cam / obs_pt: K indices, obs_pt non-decreasing (the order :161-182 produces).
 * In a multi-rank handle (b200ba_comm_init) cameras are replicated and pts / observations are this rank's shard
 * with obs_pt indexing the local points. */
int b200ba_set_problem(b200ba_handle* h, int32_t N, int32_t M, int32_t K,
                       const double* cam_RT, const double* K5, const double* pts,
                       const double* obs_xy, const int32_t* obs_cam, const int32_t* obs_pt);

/* ---- persistent state across keyframes (the reference re-marshals the whole map through std::map
 * look-ups on every keyframe, src/sfm_ba_ssba.cpp:124-182, src/sfm_tracker.cpp:192-195,389-399) ----
 * A handle may be reused: b200ba_set_problem on a handle that already holds a problem keeps the stream, the pinned staging pool
 * and the device arena (grow-only, 25 % headroom when it has to grow).  With options.keep_problem = 1 the problem can also be
 * EXTENDED: b200ba_append adds n_cam cameras (indices N .. N + n_cam - 1), n_pts points (M .. M + n_pts - 1) and n_obs
 * observations whose indices refer to the extended numbering (obs_pt non-decreasing; observations of old points are merged behind
 * the stored ones of the same point, i.e. the new frames have the largest ids).  The result is bit-identical to
 * b200ba_set_problem on the full arrays.  b200ba_set_state replaces the parameter values (either pointer may be NULL) without
 * touching the observation layout. */
int b200ba_append(b200ba_handle* h, int32_t n_cam, const double* cam_RT, int32_t n_pts, const double* pts,
                  int32_t n_obs, const double* obs_xy, const int32_t* obs_cam, const int32_t* obs_pt);
int b200ba_set_state(b200ba_handle* h, const double* cam_RT, const double* pts);

/* Whole BA: normalise, run LM until a stop criterion, de-normalise.  Equivalent of opt.minimize() inside the
 * normaliser scope of src/sfm_ba_ssba.cpp:200-222. */
int b200ba_solve(b200ba_handle* h, b200ba_report* report);

/* The same, in pieces (bench.py times b200ba_iterate): begin = (re)start from the uploaded parameters and
 * normalise; iterate = run up to n LM loop iterations; finish = de-normalise and fill the report. */
int b200ba_begin(b200ba_handle* h);
int b200ba_iterate(b200ba_handle* h, int32_t n, int32_t* done);
/* Put the device state back to what b200ba_begin produced (device-to-device copy of the normalised start point, LM
 * state reset) without touching the host: lets a benchmark time many full LM iterations from the same start. */
int b200ba_restart(b200ba_handle* h);
/* Device time (CUDA events on the engine's stream) of the last b200ba_iterate call, in milliseconds. */
int b200ba_last_iterate_ms(b200ba_handle* h, double* ms);
/* Device time of the last b200ba_restart (its device-to-device copies), in milliseconds. */
int b200ba_last_restart_ms(b200ba_handle* h, double* ms);
int b200ba_finish(b200ba_handle* h, b200ba_report* report);

/* Download the adjusted cameras / points (what src/sfm_ba_ssba.cpp:250-271 writes back).  When the outcome is
 * "discarded" (report.ret == -1) the inputs are returned unchanged, like the reference. */
int b200ba_get_result(b200ba_handle* h, double* cam_RT, double* pts);

/* The same with the points left on the device (state stays resident between the stages that
 * follow a BA): cam_RT (host, N x 12, may be NULL) as above; d_pts = DEVICE memory on the handle's device, M x 3 doubles in
 * the caller's point order, bit-identical to what b200ba_get_result writes.  b200post_ground_stage_ba (b200ba_post.h)
 * chains the post-BA stage onto it.  The reference has no counterpart (its BA result lives in host containers). */
int b200ba_get_result_device(b200ba_handle* h, double* cam_RT, double* d_pts);
/* Sizes of the uploaded problem (any pointer may be NULL). */
int b200ba_problem_size(const b200ba_handle* h, int32_t* N, int32_t* M, int32_t* K);

/* Execution plan b200ba_set_problem chose for the uploaded problem (the reference has no counterpart: SSBA / PBA have one
 * code path).  Tests use it to make sure a given plan is the one a parity case exercised. */
enum {
    B200BA_TABLE_GLOBAL  = 0,   /* packed camera table [quaternion | T | p] read through L2 by every observation        */
    B200BA_TABLE_SHARED  = 1,   /* whole table staged into one SM's shared memory by TMA bulk copies (table + row rings <= 227 KB: ~1 797 cameras) */
    B200BA_TABLE_CLUSTER = 2    /* table sharded over the distributed shared memory of a thread-block cluster           */
};
typedef struct b200ba_plan {
    int32_t struct_size;
    int32_t camera_table;             /* B200BA_TABLE_*                                                              */
    int32_t cluster_size;             /* CTAs per cluster of the by-point sweeps (1 unless camera_table == CLUSTER)   */
    int32_t cooperative_step;         /* camera-side PCG step as one cooperative kernel                               */
    int32_t cuda_graph;               /* one LM iteration replayed as a CUDA graph                                    */
    int32_t peer_windows;             /* NVLink peer-memory collectives mapped                                        */
    int32_t world, rank;
    int32_t launches_per_iteration;   /* kernels of one LM iteration (0 until the first iteration has been enqueued)  */
    int32_t n_slices, padded_entries; /* SELL-32 slices / entries of the point-order layout                           */
    int32_t n_rows, n_segments;       /* camera-order rows of 32 / partial-sum segments                               */
    int32_t fp32_pcg;
    int64_t device_bytes;             /* size of the problem's device arena                                           */
    int32_t explicit_schur;           /* 0: matrix-free PCG; else CTAs that hold the explicit reduced system in shared  */
                                      /* memory (8 / 16 = one thread-block cluster, more = the whole cooperative grid)    */
    int32_t explicit_pairs;           /* observation pairs summed into the reduced system per LM iteration               */
    int32_t reserved[6];
} b200ba_plan;
int b200ba_get_plan(const b200ba_handle* h, b200ba_plan* plan);

/* ---- multi-GPU: one handle per rank, problem partitioned by point ---- */
#define B200BA_COMM_ID_BYTES 128
int b200ba_comm_unique_id(void* id_bytes);    /* rank 0 creates, the launcher broadcasts (e.g. torch.distributed) */
int b200ba_comm_init(b200ba_handle* h, const void* id_bytes, int32_t rank, int32_t world);
/* Optional, after b200ba_comm_init and before b200ba_set_problem: NVLink peer-memory collectives.  Every rank allocates an
 * exchange window sized for problems of up to max_cameras cameras and returns its CUDA IPC handle; the launcher
 * all-gathers the world x B200BA_PEER_HANDLE_BYTES handles and every rank maps its peers.  From then on the per-PCG-step
 * all-reduce of the camera sums runs INSIDE the camera-side step kernel (publish, flag, peer loads over NVLink, summed in
 * rank order) and the once-per-iteration reductions as two small kernels, so a multi-rank LM iteration has no host-side
 * collective call and is replayed as one CUDA graph.  Without these two calls every collective is an NCCL all-reduce.
 * All ranks must have returned from b200ba_solve / b200ba_finish before any rank calls b200ba_destroy (launcher barrier).
 * The reference has no counterpart (SSBA and PBA are single-process). */
#define B200BA_PEER_HANDLE_BYTES 64
int b200ba_comm_peer_handle(b200ba_handle* h, int32_t max_cameras, void* handle_bytes);
int b200ba_comm_open_peers(b200ba_handle* h, const void* all_handles);

/* ---- stage-level access for parity tests and profiling (not needed by the adapter) ---- */
/* Linearise at the current (normalised) state after b200ba_begin: any output may be NULL.
 * cost, w[K], gc[6N], gp[3M], U[36N], V[9M] (full symmetric blocks), lambda0 = tau * max diag. */
int b200ba_debug_linearize(b200ba_handle* h, double* cost, double* w, double* gc, double* gp,
                           double* U, double* V, double* lambda0);
/* Sx = ((U + lambda I) - W (V + lambda I)^-1 W^t) x for the last linearisation; x, Sx: 6N. */
int b200ba_debug_schur_matvec(b200ba_handle* h, double lambda, const double* x, double* Sx);
/* Time `reps` launches of one kernel group with CUDA events; which: see b200ba_kernel_name. */
int b200ba_debug_time_kernel(b200ba_handle* h, int32_t which, int32_t reps, double* ms_per_launch);
const char* b200ba_kernel_name(int32_t which);   /* NULL past the last kernel */
/* Host-only, needs no device: conflict factor of the point-order layout b200ba_set_problem builds for these observations
 * (expected shared-memory wavefronts of one 16-byte camera-table read per quarter-warp phase; 1.0 = conflict-free).
 * out[4] = {factor with options.lane_window grouping + slot assignment, factor of the plain track-length order,
 * milliseconds of the layout build, padded SELL entries / K}.  B200BA_E_STATE if the slot assignment is not a permutation. */
int b200ba_debug_layout_stats(int32_t N, int32_t M, int32_t K, const int32_t* obs_cam, const int32_t* obs_pt,
                              int32_t lane_window, double* out);
/* Host-only, needs no device: the pair lists of the explicit reduced camera system (options.reduced_system) for these
 * observations, with observation k standing for its own layout entry.  n[4] = {pairs, chunks, blocks, 0}; when the arrays
 * are non-NULL and cap >= pairs: pair_a / pair_b = the two observations of every pair in summation order, pair_block = index of
 * the destination block in the row-major upper block triangle.  B200BA_E_INVALID when the problem has more pairs than the
 * explicit plan accepts.  shape[6] (may be NULL) = {kind, CTAs, rows per CTA, padded row length, threads, shared-memory bytes}
 * of the PCG kernel that would hold a problem of N cameras on a device with n_sm multiprocessors (kind -1: matrix-free). */
int b200ba_debug_dense_lists(int32_t N, int32_t M, int32_t K, const int32_t* obs_cam, const int32_t* obs_pt, int64_t* n,
                             int32_t* pair_a, int32_t* pair_b, int32_t* pair_block, int64_t cap, int32_t n_sm, int64_t* shape);
int b200ba_version(void);

#ifdef __cplusplus
}
#endif
#endif /* B200BA_H */
