/*
 * frx.h — C ABI of the B200-native MINCO / SE(3)-GCOPTER trajectory-optimisation back end.
 *
 * This is the drop-in boundary for ONE hot path of ZJU-FAST-Lab/Fast-Racing: the L-BFGS
 * cost-and-gradient evaluation of SE3GCOPTER (and the L-BFGS solve that drives it), batched over
 * independent candidate trajectories.  Plain pointers and sizes only; no C++/torch types.
 * "CPU.hpp" below = src/plan_manage/include/se3gcopter/se3gcopter_cpu.hpp of the reference,
 * "lbfgs.hpp" = src/plan_manage/include/se3gcopter/lbfgs.hpp, "cc.cuh" =
 * src/plan_manage/include/cuda_computer.cuh.
 *
 * Conventions
 *  - fp64 everywhere; integer tables are 32-bit int (the reference uses Eigen::VectorXi).
 *  - 3x3 states are column-major [p v a] (Eigen default), faces are (nx,ny,nz,px,py,pz) = one
 *    column of the reference's 6xF half-space matrix (MinCoPlan_CPU.cpp:101).
 *  - Every function returns an int status: 0 = ok, FRX_ERR_* < 0 otherwise (L-BFGS status codes
 *    keep the numeric values of lbfgs.hpp:148-205).  No exceptions or exit() cross this ABI
 *    (the reference's CUDA path calls exit(1), cc.cuh:35-45).
 *  - There is NO CPU fallback: if no CUDA device is usable every compute entry point returns
 *    FRX_ERR_CUDA.
 *  - A frx_batch is not thread-safe; distinct batches are independent (each owns its stream).
 *  - Every entry point runs on its handle's device and leaves the caller's current device as it found it.
 */
#ifndef FRX_H_
#define FRX_H_

#ifdef __cplusplus
extern "C" {
#endif

#define FRX_OK 0
#define FRX_ERR_INVALID_ARG (-1)
#define FRX_ERR_INFEASIBLE_GEOMETRY (-2) /* SE3GCOPTER::setup returning false, CPU.hpp:1118-1121 */
#define FRX_ERR_CUDA (-3)
#define FRX_ERR_CAPACITY (-4)            /* piece count / face count beyond a kernel limit */
#define FRX_ERR_NOMEM (-5)

/* Scalar arguments of SE3GCOPTER::setup, CPU.hpp:1076-1092 (same order, same meaning). */
typedef struct frx_params {
    double rho;            /* rh: time-regularisation weight; > 0 selects soft total time */
    double total_t;        /* st: fixed total time when rho <= 0 */
    double grid_res;       /* gridRes (INFINITY from the only caller, MinCoPlan_CPU.cpp:117) */
    int itg_spaces;        /* itgSpaces = QdIntervals: quadrature intervals per piece */
    int c2_diffeo;         /* c2diffeo */
    double horiz_half_len; /* horiHalfLen */
    double vert_half_len;  /* vertHalfLen */
    double safe_margin;    /* margin */
    double v_max;          /* vm */
    double thr_acc_min;    /* minThrAcc */
    double thr_acc_max;    /* maxThrAcc */
    double body_rate_max;  /* bodyRateMax */
    double grav_acc;       /* g */
    double chi[4];         /* w: penalty weights (P, V, T, B) */
} frx_params;

/* One candidate trajectory = the remaining arguments of SE3GCOPTER::setup. */
typedef struct frx_problem {
    double ini[9];       /* iniState 3x3 column-major [p v a] */
    double fin[9];       /* finState */
    int n_poly;          /* cfgPolyHs.size() */
    const int *face_off; /* n_poly + 1 offsets (in faces) */
    const double *faces; /* 6 * face_off[n_poly] doubles */
    /* Optional V-polytope override (2*n_poly - 1 polytopes in the reference's "origin + edge"
     * form, CPU.hpp:1049: column 0 = v0, column k = v_k - v0).  NULL = enumerate from the
     * half-spaces (replaces SE3GCOPTER::extractVs, CPU.hpp:1031-1074). */
    const int *vert_off;  /* 2*n_poly offsets (in vertices), or NULL */
    const double *verts;  /* 3 * vert_off[2*n_poly-1] doubles, or NULL */
    frx_params params;
} frx_problem;

typedef struct frx_batch frx_batch; /* opaque: device + pinned memory for a batch of problems */

/* --- construction: replaces SE3GCOPTER::setup (CPU.hpp:1076) and cuda_computer::setup
 *     (cc.cuh:120) for every problem of the batch.  Builds the integer tables intervals /
 *     idxVs / idxHs / cons / dimFreeT / dimFreeP (CPU.hpp:1123-1152) on the host, uploads
 *     everything once.  device_ordinal < 0 = current device. ------------------------------- */
int frx_batch_create(const frx_problem *problems, int count, int device_ordinal, frx_batch **out);
/* Replaces ~SE3GCOPTER / cuda_computer::kill_kernel (cc.cuh:138): nothing spins on the device,
 * so this only frees memory. */
void frx_batch_destroy(frx_batch *b);

/* --- queries ----------------------------------------------------------------------------- */
int frx_batch_count(const frx_batch *b);
/* n = dimFreeT + dimFreeP (the L-BFGS dimension), pieces = fineN, CPU.hpp:1125-1152 */
int frx_batch_dims(const frx_batch *b, int traj, int *n, int *pieces, int *dim_t, int *dim_p);
/* Packed layout of all per-trajectory x / gradient vectors: off[count+1]. */
int frx_batch_x_offsets(const frx_batch *b, int *off);
/* Same for per-piece outputs (durations): off[count+1] (in pieces). */
int frx_batch_piece_offsets(const frx_batch *b, int *off);
/* Integer tables of one trajectory, bit-exact with CPU.hpp:1123-1152: intervals[n_poly],
 * idx_hs[pieces], idx_vs[pieces-1], cons[pieces].  Any pointer may be NULL. */
int frx_batch_tables(const frx_batch *b, int traj, int *intervals, int *idx_hs, int *idx_vs, int *cons);
/* V-polytopes in use for one trajectory (enumerated or supplied): vert_off[2*n_poly],
 * verts[3*vert_off[2*n_poly-1]].  Call with verts == NULL to get only the offsets. */
int frx_batch_vpolys(const frx_batch *b, int traj, int *vert_off, double *verts);

/* --- the hot path ------------------------------------------------------------------------ */
/* Starting point of the reference's optimize(): setInitial + backwardT + backwardP
 * (CPU.hpp:1237-1240) for every trajectory; x is the packed host array (x_offsets). */
int frx_batch_initial_guess(frx_batch *b, double *x_host);

/* One cost+gradient evaluation of SE3GCOPTER::objectiveFunc (CPU.hpp:961-1000) for every
 * trajectory of the batch.  Host buffers: x (packed), cost[count], grad (packed).  Copies
 * host->device, runs the kernels, copies back, synchronises. */
int frx_batch_eval(frx_batch *b, const double *x_host, double *cost_host, double *grad_host);
/* The batch's own pinned (page-locked) host buffers: x[total], cost[count], grad[total].  Passing exactly these
 * pointers to frx_batch_eval skips the staging copies (the host<->device transfers then run directly from/to them). */
int frx_batch_host_buffers(frx_batch *b, double **x_pinned, double **cost_pinned, double **grad_pinned);
/* Same with device-resident buffers on the given CUDA stream (cudaStream_t passed as void*;
 * NULL = the CUDA default stream, as everywhere in CUDA).  Asynchronous: no synchronisation. */
int frx_batch_eval_device(frx_batch *b, const double *x_dev, double *cost_dev, double *grad_dev, void *stream);
/* The batch's own (non-blocking) stream, on which frx_batch_eval / frx_batch_optimize run. */
void *frx_batch_stream(frx_batch *b);
/* Number of kernel launches issued by one frx_batch_eval_device call. */
int frx_batch_eval_launches(const frx_batch *b);

/* Signature-identical to lbfgs_evaluate_t (lbfgs.hpp:223-226) so that the reference's own
 * lbfgs_optimize can drive the CUDA evaluator: instance = a frx_batch of ONE trajectory. */
double frx_eval(void *instance, const double *x, double *g, const int n);

/* Full solve: replaces SE3GCOPTER::optimize (CPU.hpp:1230-1268) for every trajectory: batched
 * device-side L-BFGS (mem_size 128, past 3, g_epsilon 1e-16, min_step 1e-32, delta = rel_tol;
 * CPU.hpp:1243-1247) from frx_batch_initial_guess.  Outputs (any may be NULL):
 *   durations[sum pieces]          piece durations (piece_offsets)
 *   coeffs[18 * sum pieces]        per piece 3x6 column-major, highest order first (CPU.hpp:561)
 *   jerk_cost[count]               return value of optimize() (jerk energy only, CPU.hpp:1267)
 *   final_cost[count]              penalised objective at the final iterate
 *   lbfgs_status[count]            lbfgs_optimize return code (discarded by the reference)
 *   iterations[count], evaluations[count]
 *   x_final (packed)               final (tau, xi) */
int frx_batch_optimize(frx_batch *b, double rel_tol, double *durations, double *coeffs, double *jerk_cost,
                       double *final_cost, int *lbfgs_status, int *iterations, int *evaluations,
                       double *x_final);

/* Queue-fed solver slots: with 0 < slots < count, frx_batch_optimize keeps `slots` trajectories in flight; a trajectory that
 * finishes hands its slot to the next one of the batch, so a solver tick costs what the slots cost and the slowest
 * trajectories of an early wave overlap the bulk of a later one.  Results per trajectory do not depend on the slot count
 * (every trajectory's solve is independent of what else is in flight).  0 (default) = one slot per trajectory. */
int frx_batch_set_solver_slots(frx_batch *b, int slots);

/* Arg-min of final_cost over the trajectories of this batch (ties -> lowest index).
 * Device-side reduction; the cross-GPU gather is the caller's (NCCL all-gather of 16 B/rank). */
int frx_batch_argmin(frx_batch *b, int *best_traj, double *best_cost);

/* Multi-GPU, one rank per GPU (the batch is sharded by the caller): arg-min over ALL ranks after
 * frx_batch_optimize.  `nccl_comm` = the caller's ncclComm_t for these n_ranks; global_index_offset = index of this rank's
 * first trajectory in the global batch.  One ncclAllGather of 16 B per rank on the batch stream; every rank gets the same
 * answer (ties -> lowest global index).  NCCL is taken from the process (dlopen), libfrx.so does not link it. */
int frx_batch_argmin_allgather(frx_batch *b, void *nccl_comm, int n_ranks, long long global_index_offset,
                               long long *best_global_index, double *best_cost);

/* Reference-order mode (on != 0): evaluation and solver run in the reference's exact operation order — no FMA
 * contraction, divisions as written, sequential reductions (the kernels are compiled a second time with --fmad=false,
 * the penalty integral and the L-BFGS dot products run sequentially).  Results are bit-identical to the reference CPU
 * code for the C2 time map; slow by construction — for validation and bit-reproducible regression runs. */
int frx_batch_set_reference_order(frx_batch *b, int on);

/* Device-time (ms, CUDA events on the batch stream) of the last eval / optimize call. */
double frx_batch_last_ms(const frx_batch *b);

/* ---- corridor generation: the step BEFORE setup() in MavGlobalPlanner::plan (MinCoPlan_CPU.cpp:37-105) ----
 * obstacle cloud: 3 doubles per point (the planner's *obs_pointer); path: 3 doubles per point (the front end's sample path);
 * bbox: config.polyhedronBox; a polytope comes back as (n, p) rows = the columns of the reference's 6 x F hPolys matrices. */
#define FRX_CORRIDOR_MAX_FACES 96 /* faces of one polytope (obstacle-tangent planes + 6 box planes + ceiling/floor) */
/* map_util->isBlocked(a, b) of the front end (out of scope here): non-zero when the straight line a -> b is blocked. */
typedef int (*frx_is_blocked_t)(void *ctx, const double *a, const double *b);
/* LineSegment3D::dilate(0) (decomp_util/line_segment.h:33-37) for one segment, on the host: obstacle-tangent planes in the
 * order they are found, then the six local bounding-box planes.  Returns the number of faces (< 0: FRX_ERR_*).
 * ellipsoid12 (may be NULL): C row-major 3x3, then the centre d. */
int frx_corridor_dilate(const double *obs, int n_obs, const double *p1, const double *p2, const double *bbox, double *faces,
                        int cap_faces, double *ellipsoid12);
/* The whole walk along the path with the ceiling / floor planes added (MinCoPlan_CPU.cpp:44-91), on the host.  face_off
 * receives *n_poly + 1 offsets.  `blocked` may be NULL (nothing is blocked: only the 4 m rule cuts the segments). */
int frx_corridor_generate(const double *obs, int n_obs, const double *path, int n_path, const double *bbox, double map_height,
                          frx_is_blocked_t blocked, void *ctx, int *face_off, double *faces, int cap_poly, int cap_faces,
                          int *n_poly);
/* The obstacle cloud resident on one GPU, and the segment step batched over many segments (one CTA per segment): what a
 * planner that grows the corridors of many candidate paths at once calls per walk step. */
typedef struct frx_cloud frx_cloud;
int frx_cloud_create(const double *obs, int n_obs, int device_ordinal, frx_cloud **out);
void frx_cloud_destroy(frx_cloud *c);
/* segments: 6 doubles each (p1, p2); face_cnt[n_seg]; faces[n_seg][FRX_CORRIDOR_MAX_FACES][6]; ellipsoids (may be NULL):
 * 12 per segment.  face_cnt < 0 for a segment whose polytope would exceed FRX_CORRIDOR_MAX_FACES. */
int frx_cloud_dilate_batch(frx_cloud *c, const double *segments, int n_seg, const double *bbox, int *face_cnt, double *faces,
                           double *ellipsoids);

/* The walk of MavGlobalPlanner::plan for MANY candidate paths at once (e.g. the distinct corridor topologies of one planning
 * cycle): all walks advance in lock-step, every step is ONE frx_cloud_dilate_batch over the paths still walking.
 * path_off[n_paths + 1] offsets (in points) into paths[3 * total]; poly_off[n_paths + 1] receives offsets (in polytopes) into
 * face_off[total_polys + 1] (offsets in faces) into faces.  `blocked` (may be NULL) is called as blocked(ctx, a, b). */
int frx_cloud_generate_batch(frx_cloud *c, const double *paths, const int *path_off, int n_paths, const double *bbox, double map_height,
                             frx_is_blocked_t blocked, void *ctx, int *poly_off, int *face_off, double *faces, int cap_poly, int cap_faces);

/* One process (or thread) per GPU: NUMA node of the GPU (-1 = unknown), and "move the calling thread and the threads it
 * creates later to that node's CPUs and prefer its memory" — call it first in each rank of a multi-GPU job.  The batch's
 * pinned host buffers are placed on the GPU's node regardless (FRX_NUMA=0 disables both).  The reference is a single
 * process on device 0 (cuda_computer::setup, cuda_computer.cu:414) and has no counterpart. */
int frx_device_numa_node(int device_ordinal);
int frx_bind_thread_to_device(int device_ordinal);

const char *frx_strerror(int status);
const char *frx_version(void);

#ifdef __cplusplus
}
#endif
#endif /* FRX_H_ */
