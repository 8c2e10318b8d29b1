// Host-side model of one candidate trajectory: everything SE3GCOPTER::setup derives before the
// hot loop starts (reference: se3gcopter_cpu.hpp "CPU.hpp" :1076-1186) plus the starting point of
// optimize() (CPU.hpp:1188-1240).  Pure host C++ (no CUDA), compiled into libfrx.so.
#pragma once
#include <vector>

#include "../../include/frx.h"

namespace frx {

struct VPolytope {
    // "origin + edge" form of CPU.hpp:1049: col 0 = v0, col k>0 = v_k - v0.  3*V doubles.
    std::vector<double> v;
    int cols() const { return (int)v.size() / 3; }
};

struct TrajectoryModel {
    frx_params par;
    bool softT = true, c2 = true;
    double rho = 0.0, sumT = 1.0;
    double ini[9], fin[9];          // unclamped boundary states (what MINCO_S3::reset gets, CPU.hpp:1180)
    double ini_clamped_p[3], fin_clamped_p[3];  // iState/fState column 0 (positions are untouched by the clamp)
    int coarseN = 0, fineN = 0, dimT = 0, dimP = 0;
    std::vector<int> face_off;       // coarseN+1
    std::vector<double> faces;       // 6*F total, normals normalised (CPU.hpp:1116)
    std::vector<VPolytope> vpolys;   // 2*coarseN-1
    std::vector<int> intervals;      // coarseN
    std::vector<int> idxVs;          // fineN-1
    std::vector<int> idxHs;          // fineN
    std::vector<int> cons;           // fineN
    std::vector<int> xi_off;         // fineN: offset of waypoint i's xi segment inside the P block (running j += k, CPU.hpp:744)
    int n() const { return dimT + dimP; }
};

// Builds the model.  Returns FRX_OK or FRX_ERR_*.
int build_model(const frx_problem& p, TrajectoryModel& out);

// setInitial + backwardT + backwardP (CPU.hpp:1237-1240): writes the n() entries of x.
void initial_guess(const TrajectoryModel& m, double* x);
// The sequential part alone (setInitial + backwardT): writes the tau block of x and the target waypoints innerP (3 per inner
// waypoint); the per-waypoint solves of backwardP then run on the device (geometry_kernels.cu: waypoint_backward_kernel).
void initial_guess_prepare(const TrajectoryModel& m, double* x, std::vector<double>& innerP);

// --- geometry (replaces geoutils::findInterior / enumerateVs, geoutils.hpp:43-149; algorithm in geometry_core.hpp) ---------
// faces: 6*F with unit normals.  out = 3*V vertex coordinates (absolute).  Returns false if the polytope has no interior, is
// unbounded, or has more than geo::MAX_FACES faces.
bool enumerate_vertices(const double* faces, int F, std::vector<double>& out);
struct GeoCandidate {
    double x[3];
    int key;  // geo::triple_key of the face triple that produced it
};
void pack_planes(const double* faces, int F, double* nd);  // (n, p) rows -> (n, d = n.p) rows
bool finish_vertices(const double* nd, int F, std::vector<GeoCandidate>& cand, std::vector<double>& out);

}  // namespace frx
