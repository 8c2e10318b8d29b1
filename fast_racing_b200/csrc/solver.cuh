// Batched device-side L-BFGS: replaces lbfgs::lbfgs_optimize (lbfgs.hpp:1103-1444) as called by
// SE3GCOPTER::optimize (CPU.hpp:1243-1256) for every trajectory of a batch at once.
//
// Design ("lockstep ticks"): the evaluator cannot be called back from inside a line search on the
// device, so every trajectory carries a resumable line-search state (lbfgs_core.cuh).  One tick =
//   (1) the three evaluation kernels on the current trial points of all unfinished trajectories,
//   (2) lbfgs_update_kernel: one CTA per trajectory consumes (f, g): directional derivative,
//       More-Thuente / backtracking transition, and — when a line search finishes — convergence
//       tests, (s, y) history update, two-loop recursion and the first trial point of the next
//       iteration.  The four kernels run over solver slots (SolverView::slot_list): a finished trajectory hands its
//       slot to the next one of the batch, or leaves it empty.
// The host only polls a device counter of unfinished trajectories every few ticks; a poll group of ticks is one CUDA graph.
#pragma once
#include <cuda_runtime.h>

#include <vector>

#include "../../include/frx.h"
#include "device_batch.cuh"
#include "lbfgs_core.cuh"

namespace frx {

enum : int { PH_INIT = 0, PH_MT = 1, PH_BT = 2, PH_DONE = 3 };

struct TrajSolve {  // per-trajectory scalar state, lives in global memory
    LineSearch ls;
    int phase, status, k, end, n_evals, pad;
    double fx, fp, step, stepp;
    double pf[8];  // ring of past objective values (past <= 8)
};

struct SolverView {
    int B;
    LbfgsConfig cfg;
    const int* x_off;   // B+1
    TrajSolve* st;      // B
    int* active;        // B: 1 while the trajectory still wants evaluations (read by the eval kernels)
    int* n_active;      // 1: trajectories not finished yet (of all B)
    // Queue-fed slots: the kernels of a tick run over n_slots entries; slot_list[s] = trajectory being solved in slot s (-1:
    // empty).  A trajectory that finishes hands its slot to the next pending one (next_pending), so a tick costs what the
    // slots cost, not what the whole batch costs, and stragglers of an early wave overlap the bulk of a later one.
    // n_slots == B: every trajectory has its own slot from the start (no queue).
    int* slot_list;     // n_slots
    int* next_pending;  // 1
    int n_slots;
    int upd_nst;        // ring stages of the streamed history passes
    int upd_ch;         // doubles per ring stage of the streamed history passes
    int upd_smem;       // dynamic shared memory of the update kernel (streamed history passes), 0: plain loads
    double *x, *g, *cost;          // current trial point / gradient / objective (packed)
    double *xp, *gp, *d;           // previous iterate, its gradient, search direction
    double *S, *Y;                 // history: trajectory b, slot j at mem_size*x_off[b] + j*n_b
    double *ys, *alpha;            // mem_size per trajectory
    double *SY, *SYT, *YY;         // mem_size^2 per trajectory: Gram matrices of the vector-free direction update
    int vector_free;               // 1: vector-free direction update (default); 0: sequential two-loop in the reference's order
    unsigned long long* hist_bytes;  // measurement: algorithmic history bytes streamed by the sampled launches (see solver_run)
    int count_bytes;                 // 1 on sampled launches
    double* dchk;                    // self-check (FRX_LBFGS_CHECK=1): scratch direction of the two-loop recursion
    unsigned long long* dir_err_bits;  // max relative deviation |d_vf - d_two_loop|_inf / |d_two_loop|_inf, as double bits
    int check_dir;
};

struct SolverIO {
    const double* x0_host = nullptr;
    double rel_tol = 1e-5;
    cudaStream_t stream = nullptr;
    int max_ticks = 400000;  // safety net only; the reference has no iteration limit (max_iterations = 0)
    int slots = 0;           // solver slots (0 or >= B: one per trajectory)
    int poll_every = 8;
};

}  // namespace frx

#include <functional>

// The solver is compiled twice (see build.py): namespace frx_fast (default) and frx_exact (reference operation order).
#define FRX_DECLARE_SOLVER_API(NS)                                                                                          \
    namespace NS {                                                                                                          \
    struct SolverState;                                                                                                     \
    typedef std::function<cudaError_t(const double*, double*, double*, cudaStream_t)> EvalFn;                               \
    int solver_create(const frx::BatchView& bv, const std::vector<int>& x_off, SolverState** out);                          \
    void solver_destroy(SolverState* s);                                                                                    \
    /* Runs the solve: eval(x_dev, cost_dev, grad_dev, stream) launches one batched evaluation. */                          \
    int solver_run(SolverState* s, frx::BatchView& bv, const frx::SolverIO& io, const EvalFn& eval);                        \
    int solver_fetch(SolverState* s, double* final_cost, int* status, int* iters, int* evals, double* x_final);             \
    const double* solver_final_cost_dev(SolverState* s);                                                                    \
    int solver_argmin(SolverState* s, const double* cost_dev, int B, cudaStream_t st, int* idx, double* val);               \
    int solver_ticks(SolverState* s);                                                                                       \
    cudaError_t fetch_update_clocks(long long* out16, int reset);                                                           \
    void debug_update_ring(int max_n, int n, int batch, int* out8);                                                         \
    /* sampled lbfgs_update_kernel launches of the last run: device ms, algorithmic bytes streamed, number of samples */   \
    int solver_update_stats(SolverState* s, double* ms, double* bytes, int* samples);                                       \
    int solver_direction_check(SolverState* s, double* max_rel);                                                            \
    }
FRX_DECLARE_SOLVER_API(frx_fast)
FRX_DECLARE_SOLVER_API(frx_exact)
