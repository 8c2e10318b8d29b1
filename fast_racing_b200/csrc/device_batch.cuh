// Device-resident layout of a batch of candidate trajectories (see DESIGN.md "Data layout in HBM").
//
// Everything is packed CSR-style so that ragged batches (different piece counts M, face counts F,
// vertex counts V per trajectory) need no padding:
//   trajectory b owns pieces      [piece_off[b], piece_off[b+1])      ("global piece index" gp)
//                  variables x    [x_off[b],     x_off[b+1])           (tau block then xi block)
//                  polytopes      [coarse_off[b], coarse_off[b+1])
//   piece gp      reads faces     faces[6*piece_face_off[gp] ...] , piece_face_cnt[gp] faces
//   waypoint i of trajectory b lives in slot gp = piece_off[b]+i (the slot of its LEFT piece):
//                  vertices       verts[3*wp_vert_off[gp] ...], wp_vert_cnt[gp] columns (origin+edge form)
//                  xi segment     x[x_off[b] + dimT[b] + wp_xi_off[gp] ...], wp_vert_cnt[gp]-1 entries
#pragma once
#include <cuda_runtime.h>

namespace frx {

struct DevParams {  // per trajectory; derived from frx_params exactly as SE3GCOPTER::setup does (CPU.hpp:1095-1163)
    double rho, sumT;
    double ell[3];
    double margin;
    double vmax_sqr, thrmin_sqr, thrmax_sqr, bdrmax_sqr, gacc;
    double chi[4];
    int softT, c2;
};

struct BatchView {
    int B;
    int total_pieces;
    int total_x;
    int max_faces;    // max faces of any piece's polytope
    int max_samples;  // max cons+1
    // sub-batch window of this launch (the evaluation is pipelined over sub-batches on two streams)
    int sub_traj0;    // first trajectory handled by K1/K3 of this launch
    int sub_traj1;    // one past the last trajectory of this launch
    int sub_piece0;   // first global piece handled by K2 of this launch
    int sub_piece1;   // one past the last global piece handled by K2 of this launch
    // per trajectory
    const int* piece_off;   // B+1
    const int* x_off;       // B+1
    const int* coarse_off;  // B+1
    const int* dimT;        // B
    const DevParams* par;   // B
    const double* headtail; // 18*B : head PVA (col-major 3x3) then tail PVA
    const int* active;      // B, or nullptr: trajectories with active[b]==0 are skipped (batched solver mask)
    // Slot mode of the batched solver (solver.cuh "queue-fed slots"): when slot_list != nullptr the evaluation kernels run over
    // n_slots entries, entry s naming the trajectory currently solved in slot s (or -1: empty slot)
    const int* slot_list;
    int n_slots;
    int max_pieces;         // max pieces of any trajectory (grid of the penalty kernel in slot mode)
    // per coarse polytope
    const int* intervals;   // sum coarseN
    // per piece (global piece index)
    const int* piece_traj;      // owning trajectory
    const int* piece_poly;      // idxHs: owning polytope, local to the trajectory (CPU.hpp:1149)
    const int* cons;            // quadrature intervals (CPU.hpp:1126-1127)
    const int* piece_face_off;  // in faces
    const int* piece_face_cnt;
    const int* wp_vert_off;     // in vertices (columns)
    const int* wp_vert_cnt;     // columns of the waypoint's V-polytope (k+1); 0 for the last slot
    const int* wp_xi_off;       // offset inside the xi block
    const double* faces;        // 6 * total faces
    const float* faces_f32;     // 4 * total faces: (n, d_lo) in single precision for the conservative near-face prefilter of K2
    const double* verts;        // 3 * total vertex columns
    // work arrays (per piece unless noted)
    double* T;        // piece durations
    double* q;        // 3 per slot: inner waypoints
    double* wpw;      // 4 per slot: (sum_a V_{a+1} xi_a^2, |xi|^2) of the waypoint map, reused by its gradient in K3
    double* c;        // 18 per piece: coefficient block rows 6i..6i+5, row-major (6x3)
    double* lu;       // 16 per row, 6 rows per piece: band LU factors (13 diagonals) + 3 unused
    double* gdC;      // 18 per piece: dJ/dc, becomes lambda after the adjoint solve
    double* gdT;      // dJ/dT partial (energy + penalty)
    double* pen;      // per piece penalty integral
    double* energy;   // per piece jerk energy
};

// Launch configuration of the penalty kernel, chosen on the host from max_samples.
struct PenaltyLaunch {
    int spt;      // consecutive samples per thread
    int tpp;      // threads per piece
    int ppc;      // pieces per CTA
    int threads;  // CTA size
    int grid;
    size_t smem;
};

}  // namespace frx
