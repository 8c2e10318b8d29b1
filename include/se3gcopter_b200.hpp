// se3gcopter_b200.hpp — header-only C++ façade over the C ABI (frx.h) with the method set of the
// reference's optimiser, so that its caller compiles against it unchanged in shape:
//
//   reference                                                        this header
//   ---------------------------------------------------------------  ------------------------------------
//   class SE3GCOPTER            se3gcopter_cpu.hpp:567                frx::SE3GCOPTER
//     bool   setup(rho, totalT, iniState, finState, hPolys, gridRes,   same argument order and meaning
//                  itgSpaces, horizHalfLen, vertHalfLen, margin, vMax,
//                  minThrAcc, maxThrAcc, bodyRateMax, g, w, c2diffeo)   :1076-1092
//     double optimize(Trajectory& traj, relCostTol)                    :1230-1268 (returns the jerk cost)
//     void   kill_kernel()                 se3gcopter_gpu.hpp:907-909   no-op (nothing spins on the device)
//   class Piece / Trajectory    trajectory.hpp:43, :317               frx::Piece / frx::Trajectory
//     getPieceNum, getTotalDuration, operator[], getDuration, getOrder, getCoeffMat (3x6, highest order
//     first), normalizePosCoeffMat, getPos/getVel/getAcc/getJer, getRotation, getTiltRate
//   caller: MavGlobalPlanner::plan   MinCoPlan_CPU.cpp:114-124, MinCoPlan_GPU.cpp:127
//
// Matrices cross this façade as plain column-major arrays (Mat3 = double[9] [p v a]; an H-polytope is a
// std::vector<double> of 6*F values = the columns of the reference's 6xF matrix).  When Eigen is present at
// the integrator's site (`__has_include(<Eigen/Eigen>)`), overloads taking Eigen::MatrixXd /
// std::vector<Eigen::MatrixXd> / Eigen::Vector4d with the reference's exact signature are enabled.
//
// Batch extension (not in the reference, which optimises one trajectory per call): frx::SE3GCOPTERBatch
// solves many candidate trajectories at once and returns the arg-min.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <stdexcept>
#include <string>
#include <vector>

#include "frx.h"

#if defined(__has_include)
#if __has_include(<Eigen/Eigen>)
#include <Eigen/Eigen>
#define FRX_HAVE_EIGEN 1
#endif
#endif

namespace frx {

namespace detail {
// Polynomials with coefficients in ASCENDING powers.  real_roots: every real root of p in [lo, hi] (simple roots and roots at
// the ends; used for the stationary points of |vel|^2 / |acc|^2), by recursion on the derivative - the roots of p' split
// [lo, hi] into monotone pieces - and bisection to `tol` on each piece with a sign change.  Replaces what the reference gets
// from root_finder.hpp (closed forms / Sturm sequences / companion eigenvalues, trajectory.hpp:177-273) for these queries.
inline double poly_val(const std::vector<double>& p, double t) {
    double v = 0.0;
    for (int i = (int)p.size() - 1; i >= 0; i--) v = v * t + p[i];
    return v;
}
inline std::vector<double> poly_der(const std::vector<double>& p) {
    std::vector<double> d;
    for (size_t i = 1; i < p.size(); i++) d.push_back(p[i] * (double)i);
    return d;
}
inline std::vector<double> poly_sqr_sum3(const std::vector<double> (&a)[3]) {  // sum_d a_d(t)^2
    std::vector<double> q(2 * a[0].size() - 1, 0.0);
    for (int d = 0; d < 3; d++)
        for (size_t i = 0; i < a[d].size(); i++)
            for (size_t j = 0; j < a[d].size(); j++) q[i + j] += a[d][i] * a[d][j];
    return q;
}
inline void real_roots(std::vector<double> p, double lo, double hi, double tol, std::vector<double>& out) {
    double scale = 0.0;
    for (double c : p) scale = std::max(scale, std::fabs(c));
    while (!p.empty() && std::fabs(p.back()) <= 1e-300 + 1e-15 * scale) p.pop_back();  // drop vanishing leading terms
    if (p.size() < 2) return;
    if (p.size() == 2) {
        const double r = -p[0] / p[1];
        if (r >= lo && r <= hi) out.push_back(r);
        return;
    }
    std::vector<double> knots{lo};
    real_roots(poly_der(p), lo, hi, tol, knots);
    knots.push_back(hi);
    std::sort(knots.begin(), knots.end());
    for (size_t i = 0; i + 1 < knots.size(); i++) {
        double a = knots[i], b = knots[i + 1];
        double fa = poly_val(p, a), fb = poly_val(p, b);
        if (fa == 0.0) out.push_back(a);
        if (i + 2 == knots.size() && fb == 0.0) out.push_back(b);
        if (fa * fb >= 0.0) continue;  // monotone on (a, b): no interior root without a sign change
        while (b - a > tol) {
            const double m = 0.5 * (a + b), fm = poly_val(p, m);
            if (fm == 0.0) {
                a = b = m;
                break;
            }
            if ((fm < 0.0) == (fa < 0.0))
                a = m, fa = fm;
            else
                b = m;
        }
        out.push_back(0.5 * (a + b));
    }
}
}  // namespace detail

// ---- trajectory containers (trajectory.hpp:43-145, 317-593) -------------------------------------
class Piece {
    double duration = 0.0;
    std::array<double, 18> c{};  // 3x6 column-major, highest order first (trajectory.hpp:39)

public:
    Piece() = default;
    Piece(double dur, const double* coeff18) : duration(dur) {
        for (int i = 0; i < 18; i++) c[i] = coeff18[i];
    }
    int getDim() const { return 3; }
    int getOrder() const { return 5; }
    double getDuration() const { return duration; }
    const std::array<double, 18>& getCoeffMat() const { return c; }
    double coeff(int dim, int col) const { return c[3 * col + dim]; }
    // trajectory.hpp:75-133
    std::array<double, 3> getPos(double t) const { return eval(t, 0); }
    std::array<double, 3> getVel(double t) const { return eval(t, 1); }
    std::array<double, 3> getAcc(double t) const { return eval(t, 2); }
    std::array<double, 3> getJer(double t) const { return eval(t, 3); }
    // trajectory.hpp:135-145: coefficients of the duration-normalised polynomial (wire format of traj2msg)
    std::array<double, 18> normalizePosCoeffMat() const {
        std::array<double, 18> n{};
        double t = 1.0;
        for (int i = 5; i >= 0; i--) {
            for (int d = 0; d < 3; d++) n[3 * i + d] = c[3 * i + d] * t;
            t *= duration;
        }
        return n;
    }
    // trajectory.hpp:177-273: exact maximum of |vel| / |acc| over the piece - the stationary points of the squared norm
    // (roots of its derivative in [0, T], isolated to FLT_EPSILON * T as the reference) and the two ends.
    double getMaxVelRate() const { return maxRate(1); }
    double getMaxAccRate() const { return maxRate(2); }
    // trajectory.hpp:275-313: true iff the rate stays below the limit on the whole piece
    bool checkMaxVelRate(double maxVelRate) const { return maxRate(1) < maxVelRate; }
    bool checkMaxAccRate(double maxAccRate) const { return maxRate(2) < maxAccRate; }

private:
    double maxRate(int deriv) const {
        std::vector<double> a[3];
        for (int d = 0; d < 3; d++)
            for (int p = deriv; p <= 5; p++) {  // power p of t has column 5 - p
                double f = 1.0;
                for (int q = 0; q < deriv; q++) f *= (double)(p - q);
                a[d].push_back(f * c[3 * (5 - p) + d]);
            }
        const std::vector<double> q = detail::poly_sqr_sum3(a);
        std::vector<double> cand{0.0, duration};
        detail::real_roots(detail::poly_der(q), 0.0, duration, 1.1920929e-07 * duration, cand);
        double m = 0.0;
        for (double t : cand) m = std::max(m, detail::poly_val(q, t));
        return std::sqrt(std::max(m, 0.0));
    }
    std::array<double, 3> eval(double t, int deriv) const {
        std::array<double, 3> out{0.0, 0.0, 0.0};
        double tn = 1.0;
        for (int i = 5 - deriv; i >= 0; i--) {  // column i has power p = 5 - i
            const int p = 5 - i;
            double f = 1.0;
            for (int q = 0; q < deriv; q++) f *= (double)(p - q);
            for (int d = 0; d < 3; d++) out[d] += f * tn * c[3 * i + d];
            tn *= t;
        }
        return out;
    }
};

class Trajectory {
    std::vector<Piece> pieces;

public:
    int getPieceNum() const { return (int)pieces.size(); }
    double getTotalDuration() const {
        double s = 0.0;
        for (const auto& p : pieces) s += p.getDuration();
        return s;
    }
    std::vector<double> getDurations() const {
        std::vector<double> d;
        for (const auto& p : pieces) d.push_back(p.getDuration());
        return d;
    }
    const Piece& operator[](int i) const { return pieces[i]; }
    Piece& operator[](int i) { return pieces[i]; }
    void clear() { pieces.clear(); }
    void reserve(int n) { pieces.reserve(n); }
    void emplace_back(double dur, const double* coeff18) { pieces.emplace_back(dur, coeff18); }
    std::vector<Piece>::const_iterator begin() const { return pieces.begin(); }
    std::vector<Piece>::const_iterator end() const { return pieces.end(); }
    // trajectory.hpp:437-455
    int locatePieceIdx(double& t) const {
        const int N = getPieceNum();
        int idx = 0;
        double dur = 0.0;
        for (; idx < N && t > (dur = pieces[idx].getDuration()); idx++) t -= dur;
        if (idx == N) {
            idx--;
            t += pieces[idx].getDuration();
        }
        return idx;
    }
    std::array<double, 3> getPos(double t) const { const int i = locatePieceIdx(t); return pieces[i].getPos(t); }
    std::array<double, 3> getVel(double t) const { const int i = locatePieceIdx(t); return pieces[i].getVel(t); }
    std::array<double, 3> getAcc(double t) const { const int i = locatePieceIdx(t); return pieces[i].getAcc(t); }
    std::array<double, 3> getJer(double t) const { const int i = locatePieceIdx(t); return pieces[i].getJer(t); }
    // trajectory.hpp:579-592
    std::array<double, 3> getTiltRate(double t, double gAcc) const {
        auto a = getAcc(t);
        a[2] += gAcc;
        const double thr = std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
        const double z[3] = {a[0] / thr, a[1] / thr, a[2] / thr};
        double y[3] = {0.0, z[2], -z[1]};  // z x e_x
        const double yn = std::sqrt(y[1] * y[1] + y[2] * y[2]);
        y[1] /= yn, y[2] /= yn;
        const double xb[3] = {y[1] * z[2] - y[2] * z[1], y[2] * z[0], -y[1] * z[0]};
        const auto j = getJer(t);
        const double b0 = (xb[0] * j[0] + xb[1] * j[1] + xb[2] * j[2]) / thr;
        const double b1 = (y[1] * j[1] + y[2] * j[2]) / thr;
        return {-b1, b0, 0.0};
    }
    // trajectory.hpp:565-577: body frame for a yaw angle; rotM is 3x3 column-major (columns xB, yB, zB)
    void getRotation(double t, double yaw, double gAcc, double* rotM) const {
        auto z = getAcc(t);
        z[2] += gAcc;
        const double zn = std::sqrt(z[0] * z[0] + z[1] * z[1] + z[2] * z[2]);
        for (int d = 0; d < 3; d++) z[d] /= zn;
        const double hx = std::cos(yaw), hy = std::sin(yaw);
        double y[3] = {z[1] * 0.0 - z[2] * hy, z[2] * hx - z[0] * 0.0, z[0] * hy - z[1] * hx};  // z x (cos yaw, sin yaw, 0)
        const double yn = std::sqrt(y[0] * y[0] + y[1] * y[1] + y[2] * y[2]);
        for (int d = 0; d < 3; d++) y[d] /= yn;
        const double x[3] = {y[1] * z[2] - y[2] * z[1], y[2] * z[0] - y[0] * z[2], y[0] * z[1] - y[1] * z[0]};  // y x z
        for (int d = 0; d < 3; d++) rotM[d] = x[d], rotM[3 + d] = y[d], rotM[6 + d] = z[d];
    }
    // trajectory.hpp:517-559: maxima over the pieces (exact per piece, see Piece::getMaxVelRate) and the feasibility checks
    double getMaxVelRate() const {
        double m = 0.0;
        for (const auto& p : pieces) m = std::max(m, p.getMaxVelRate());
        return m;
    }
    double getMaxAccRate() const {
        double m = 0.0;
        for (const auto& p : pieces) m = std::max(m, p.getMaxAccRate());
        return m;
    }
    bool checkMaxVelRate(double maxVelRate) const {
        for (const auto& p : pieces)
            if (!p.checkMaxVelRate(maxVelRate)) return false;
        return true;
    }
    bool checkMaxAccRate(double maxAccRate) const {
        for (const auto& p : pieces)
            if (!p.checkMaxAccRate(maxAccRate)) return false;
        return true;
    }
    // junction states (trajectory.hpp:457-515): juncIdx 0..N, the start of piece juncIdx (or the end of the last piece)
    std::array<double, 3> getJuncPos(int j) const { return j < getPieceNum() ? pieces[j].getPos(0.0) : pieces[j - 1].getPos(pieces[j - 1].getDuration()); }
    std::array<double, 3> getJuncVel(int j) const { return j < getPieceNum() ? pieces[j].getVel(0.0) : pieces[j - 1].getVel(pieces[j - 1].getDuration()); }
    std::array<double, 3> getJuncAcc(int j) const { return j < getPieceNum() ? pieces[j].getAcc(0.0) : pieces[j - 1].getAcc(pieces[j - 1].getDuration()); }
};

// ---- wire format (se3_planner.cpp:31-58 traj2msg -> quadrotor_msgs/PolynomialTrajectory.msg) -----------
// The payload fields of the message the planner publishes to traj_server, without the ROS header: per piece the
// duration-NORMALISED position coefficients, highest order first, one array per axis.  A ROS caller copies these fields
// into quadrotor_msgs::PolynomialTrajectory one to one.
struct PolynomialTrajectoryFields {
    enum : unsigned { ACTION_ADD = 1 };
    unsigned action = ACTION_ADD;
    unsigned num_order = 5, num_segment = 0;
    double start_yaw = 0.0, final_yaw = 0.0, mag_coeff = 1.0;
    std::vector<double> coef_x, coef_y, coef_z, time;
    std::vector<unsigned> order;
};
inline PolynomialTrajectoryFields traj2fields(const Trajectory& traj) {
    PolynomialTrajectoryFields m;
    m.num_segment = (unsigned)traj.getPieceNum();
    m.num_order = m.num_segment ? (unsigned)traj[0].getOrder() : 5u;
    for (unsigned i = 0; i < m.num_segment; i++) {
        const auto n = traj[(int)i].normalizePosCoeffMat();
        for (int j = 0; j <= traj[(int)i].getOrder(); j++) {
            m.coef_x.push_back(n[3 * j + 0]);
            m.coef_y.push_back(n[3 * j + 1]);
            m.coef_z.push_back(n[3 * j + 2]);
        }
        m.time.push_back(traj[(int)i].getDuration());
        m.order.push_back((unsigned)traj[(int)i].getOrder());
    }
    return m;
}

// ---- the optimiser --------------------------------------------------------------------------------
class SE3GCOPTER {
    frx_batch* h = nullptr;
    int last_status_ = 0;

public:
    SE3GCOPTER() = default;
    SE3GCOPTER(const SE3GCOPTER&) = delete;
    SE3GCOPTER& operator=(const SE3GCOPTER&) = delete;
    ~SE3GCOPTER() { frx_batch_destroy(h); }

    // se3gcopter_cpu.hpp:1076-1092.  iniState / finState: 3x3 column-major [p v a]; cfgPolyHs[i]: 6*F_i values
    // (column k = (outer normal, point) of face k); w = penalty weights (P, V, T, B).
    // Returns false exactly where the reference does: a polytope or a consecutive intersection with empty
    // interior (se3gcopter_cpu.hpp:1118-1121).  Any other failure (no CUDA device, ...) throws.
    bool setup(const double& rh, const double& st, const double* iniState, const double* finState,
               const std::vector<std::vector<double>>& cfgPolyHs, const double& gridRes, const int& itgSpaces,
               const double& horiHalfLen, const double& vertHalfLen, const double& margin, const double& vm,
               const double& minThrAcc, const double& maxThrAcc, const double& bodyRateMax, const double& g,
               const double* w, bool c2diffeo) {
        frx_batch_destroy(h);
        h = nullptr;
        frx_problem p{};
        for (int i = 0; i < 9; i++) p.ini[i] = iniState[i], p.fin[i] = finState[i];
        std::vector<int> off(1, 0);
        std::vector<double> faces;
        for (const auto& hp : cfgPolyHs) {
            faces.insert(faces.end(), hp.begin(), hp.end());
            off.push_back((int)(faces.size() / 6));
        }
        p.n_poly = (int)cfgPolyHs.size();
        p.face_off = off.data();
        p.faces = faces.data();
        p.params = frx_params{rh, st, gridRes, itgSpaces, c2diffeo ? 1 : 0, horiHalfLen, vertHalfLen, margin, vm,
                              minThrAcc, maxThrAcc, bodyRateMax, g, {w[0], w[1], w[2], w[3]}};
        const int s = frx_batch_create(&p, 1, -1, &h);
        if (s == FRX_ERR_INFEASIBLE_GEOMETRY) return false;
        if (s != FRX_OK) throw std::runtime_error(std::string("SE3GCOPTER::setup: ") + frx_strerror(s));
        return true;
    }

    // se3gcopter_cpu.hpp:1230-1268: returns the jerk cost of the final trajectory.
    double optimize(Trajectory& traj, const double& relCostTol) {
        if (!h) throw std::runtime_error("SE3GCOPTER::optimize before a successful setup");
        int n = 0, M = 0;
        frx_batch_dims(h, 0, &n, &M, nullptr, nullptr);
        std::vector<double> dur(M), co(18 * (size_t)M);
        double jerk = 0.0;
        const int s = frx_batch_optimize(h, relCostTol, dur.data(), co.data(), &jerk, nullptr, &last_status_, nullptr,
                                         nullptr, nullptr);
        if (s != FRX_OK) throw std::runtime_error(std::string("SE3GCOPTER::optimize: ") + frx_strerror(s));
        traj.clear();
        traj.reserve(M);
        for (int i = 0; i < M; i++) traj.emplace_back(dur[i], &co[18 * (size_t)i]);
        return jerk;
    }
    // The reference discards lbfgs_optimize's return code (se3gcopter_cpu.hpp:1249); it is kept here.
    int lbfgs_status() const { return last_status_; }
    void kill_kernel() {}  // se3gcopter_gpu.hpp:907-909: nothing to kill, kept for source compatibility

#ifdef FRX_HAVE_EIGEN
    bool setup(const double& rh, const double& st, const Eigen::MatrixXd& iniState, const Eigen::MatrixXd& finState,
               const std::vector<Eigen::MatrixXd>& cfgPolyHs, const double& gridRes, const int& itgSpaces,
               const double& horiHalfLen, const double& vertHalfLen, const double& margin, const double& vm,
               const double& minThrAcc, const double& maxThrAcc, const double& bodyRateMax, const double& g,
               const Eigen::Vector4d& w, bool c2diffeo) {
        const Eigen::Matrix3d i3 = iniState, f3 = finState;  // column-major
        std::vector<std::vector<double>> hp;
        for (const auto& m : cfgPolyHs) hp.emplace_back(m.data(), m.data() + m.size());  // 6 x F column-major
        return setup(rh, st, i3.data(), f3.data(), hp, gridRes, itgSpaces, horiHalfLen, vertHalfLen, margin, vm,
                     minThrAcc, maxThrAcc, bodyRateMax, g, w.data(), c2diffeo);
    }
#endif
};

// ---- corridor generation: the EllipsoidDecomp3D surface MavGlobalPlanner::plan uses (MinCoPlan_CPU.cpp:37-63) -----------
// decomp_util/ellipsoid_decomp.h:32-36,58-80 (set_obs, set_local_bbox, dilate, get_polyhedrons, get_ellipsoids) and
// decomp_geometry/polyhedron.h (Hyperplane3D, Polyhedron3D::add / inside / hyperplanes) on plain arrays.
typedef std::array<double, 3> Vec3f;
struct Hyperplane3D {
    Vec3f p_, n_;
    Hyperplane3D() {}
    Hyperplane3D(const Vec3f& p, const Vec3f& n) : p_(p), n_(n) {}
    double signed_dist(const Vec3f& pt) const { return n_[0] * (pt[0] - p_[0]) + n_[1] * (pt[1] - p_[1]) + n_[2] * (pt[2] - p_[2]); }
};
struct Polyhedron3D {
    std::vector<Hyperplane3D> vs_;
    void add(const Hyperplane3D& v) { vs_.push_back(v); }
    bool inside(const Vec3f& pt) const {
        for (const auto& v : vs_)
            if (v.signed_dist(pt) > 1e-10) return false;  // epsilon_ of decomp_basis/data_type.h:127
        return true;
    }
    const std::vector<Hyperplane3D>& hyperplanes() const { return vs_; }
};
struct Ellipsoid3D {
    double C_[9];  // row-major
    Vec3f d_;
};
class EllipsoidDecomp3D {
    std::vector<double> obs_;
    double bbox_[3] = {0, 0, 0};
    std::vector<Polyhedron3D> polyhedrons_;
    std::vector<Ellipsoid3D> ellipsoids_;

public:
    void set_obs(const std::vector<Vec3f>& obs) {
        obs_.clear();
        for (const auto& o : obs) obs_.insert(obs_.end(), o.begin(), o.end());
    }
    void set_local_bbox(const Vec3f& bbox) { bbox_[0] = bbox[0], bbox_[1] = bbox[1], bbox_[2] = bbox[2]; }
    // one polytope + ellipsoid per path segment, as ellipsoid_decomp.h:64-80
    void dilate(const std::vector<Vec3f>& path, double /*offset_x*/ = 0) {
        polyhedrons_.clear(), ellipsoids_.clear();
        std::vector<double> faces(6 * 512);
        for (size_t i = 0; i + 1 < path.size(); i++) {
            Ellipsoid3D e;
            double ell[12];
            const int F = frx_corridor_dilate(obs_.data(), (int)obs_.size() / 3, path[i].data(), path[i + 1].data(), bbox_, faces.data(), 512, ell);
            if (F < 0) throw std::runtime_error(std::string("EllipsoidDecomp3D::dilate: ") + frx_strerror(F));
            Polyhedron3D poly;
            for (int k = 0; k < F; k++)
                poly.add(Hyperplane3D({faces[6 * k + 3], faces[6 * k + 4], faces[6 * k + 5]}, {faces[6 * k], faces[6 * k + 1], faces[6 * k + 2]}));
            for (int k = 0; k < 9; k++) e.C_[k] = ell[k];
            e.d_ = {ell[9], ell[10], ell[11]};
            polyhedrons_.push_back(poly), ellipsoids_.push_back(e);
        }
    }
    const std::vector<Polyhedron3D>& get_polyhedrons() const { return polyhedrons_; }
    const std::vector<Ellipsoid3D>& get_ellipsoids() const { return ellipsoids_; }
};

// ---- batch of candidate trajectories (extension) ----------------------------------------------------
class SE3GCOPTERBatch {
    frx_batch* h = nullptr;

public:
    ~SE3GCOPTERBatch() { frx_batch_destroy(h); }
    // problems: one frx_problem per candidate; device < 0 = current device
    int setup(const std::vector<frx_problem>& problems, int device = -1) {
        frx_batch_destroy(h);
        h = nullptr;
        return frx_batch_create(problems.data(), (int)problems.size(), device, &h);
    }
    // Solves every candidate; returns the index of the cheapest one (penalised objective) and its trajectory.
    int optimize(double relCostTol, Trajectory& best, double* best_cost = nullptr) {
        const int B = frx_batch_count(h);
        std::vector<int> poff(B + 1);
        frx_batch_piece_offsets(h, poff.data());
        std::vector<double> dur(poff[B]), co(18 * (size_t)poff[B]);
        const int s = frx_batch_optimize(h, relCostTol, dur.data(), co.data(), nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
        if (s != FRX_OK) return s;
        int idx = -1;
        double c = 0.0;
        const int s2 = frx_batch_argmin(h, &idx, &c);
        if (s2 != FRX_OK) return s2;
        best.clear();
        for (int i = poff[idx]; i < poff[idx + 1]; i++) best.emplace_back(dur[i], &co[18 * (size_t)i]);
        if (best_cost) *best_cost = c;
        return idx;
    }
    // Keep `slots` candidates in flight and feed the others through as slots free up (0: all at once); results do not change.
    int setSolverSlots(int slots) { return frx_batch_set_solver_slots(h, slots); }
    frx_batch* handle() { return h; }
};

}  // namespace frx
