// H-polytope -> interior point -> V-polytope, host side.
//
// Replaces the reference's geoutils::findInterior (geoutils.hpp:43-60: a d=4 Seidel LP through
// sdlp::linprog) and geoutils::enumerateVs (geoutils.hpp:102-131: polar dual + quickhull + filterVs).
// Same mathematical pipeline, independent implementation:
//   1. interior point: maximise the slack t of  n_i . x + t <= n_i . p_i   (randomised-incremental
//      LP in 4 variables, recursing on the violated hyperplane),
//   2. polar dual about that point, incremental 3-D convex hull of the dual points,
//   3. every hull triangle is a primal vertex  normal / (normal . point)  (geoutils.hpp:117-126),
//   4. first-seen de-duplication on a grid of pitch max(epsilon, mag*DBL_EPSILON) (geoutils.hpp:75-97).
// The vertex ORDER (hence the coordinates of xi) follows this hull's face order, not quickhull's;
// callers that need the reference's order pass V-polytopes in through frx_problem::verts.
#include <algorithm>
#include <array>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <set>
#include <vector>

#include "host_model.hpp"

namespace frx {
namespace {

struct Row {
    double a[4];
    double b;
};

// minimise c.x subject to rows[0..m) (a.x <= b) in d variables.  Returns false if infeasible.
bool lp_solve(int d, const std::vector<Row>& rows, int m, const double* c, double* x) {
    if (d == 1) {
        double lo = -INFINITY, hi = INFINITY;
        for (int i = 0; i < m; i++) {
            const double a = rows[i].a[0], b = rows[i].b;
            if (std::fabs(a) < 1e-14) {
                if (b < -1e-9) return false;
                continue;
            }
            const double v = b / a;
            if (a > 0)
                hi = std::min(hi, v);
            else
                lo = std::max(lo, v);
        }
        if (lo > hi + 1e-9 * (1.0 + std::fabs(lo))) return false;
        if (lo > hi) lo = hi = 0.5 * (lo + hi);
        if (c[0] > 0)
            x[0] = lo;
        else if (c[0] < 0)
            x[0] = hi;
        else
            x[0] = std::isfinite(lo) ? (std::isfinite(hi) ? 0.5 * (lo + hi) : lo) : (std::isfinite(hi) ? hi : 0.0);
        return std::isfinite(x[0]);
    }
    // The first 2d rows are the bounding box: start at its optimal corner.
    for (int j = 0; j < d; j++) {
        double lo = -INFINITY, hi = INFINITY;
        for (int i = 0; i < 2 * d && i < m; i++) {
            bool single = true;
            for (int q = 0; q < d; q++)
                if (q != j && rows[i].a[q] != 0.0) single = false;
            if (!single || rows[i].a[j] == 0.0) continue;
            const double v = rows[i].b / rows[i].a[j];
            if (rows[i].a[j] > 0)
                hi = std::min(hi, v);
            else
                lo = std::max(lo, v);
        }
        x[j] = c[j] > 0 ? lo : (c[j] < 0 ? hi : 0.5 * (lo + hi));
    }
    std::vector<Row> sub(m);
    for (int i = 2 * d; i < m; i++) {
        const Row& r = rows[i];
        double lhs = 0.0;
        for (int j = 0; j < d; j++) lhs += r.a[j] * x[j];
        if (lhs <= r.b + 1e-12 * (1.0 + std::fabs(r.b))) continue;
        // optimum moves onto the hyperplane r.a . x = r.b : eliminate variable k
        int k = 0;
        for (int j = 1; j < d; j++)
            if (std::fabs(r.a[j]) > std::fabs(r.a[k])) k = j;
        if (std::fabs(r.a[k]) < 1e-14) return false;
        const double inv = 1.0 / r.a[k];
        for (int h = 0; h < i; h++) {
            const Row& s = rows[h];
            const double f = s.a[k] * inv;
            int q = 0;
            for (int j = 0; j < d; j++)
                if (j != k) sub[h].a[q++] = s.a[j] - f * r.a[j];
            sub[h].b = s.b - f * r.b;
        }
        double cs[4];
        {
            const double f = c[k] * inv;
            int q = 0;
            for (int j = 0; j < d; j++)
                if (j != k) cs[q++] = c[j] - f * r.a[j];
        }
        // In the sub-problem the box rows are no longer axis aligned in general; prepend a fresh,
        // slightly larger box so that the recursion always has one.
        std::vector<Row> subrows;
        subrows.reserve(i + 2 * (d - 1));
        double big = 0.0;
        for (int h = 0; h < 2 * d; h++) big = std::max(big, std::fabs(rows[h].b));
        big *= 4.0;
        for (int j = 0; j < d - 1; j++) {
            Row p{}, n{};
            p.a[j] = 1.0, p.b = big;
            n.a[j] = -1.0, n.b = big;
            subrows.push_back(p);
            subrows.push_back(n);
        }
        for (int h = 0; h < i; h++) subrows.push_back(sub[h]);
        double y[4];
        if (!lp_solve(d - 1, subrows, (int)subrows.size(), cs, y)) return false;
        double acc = r.b;
        int q = 0;
        for (int j = 0; j < d; j++)
            if (j != k) {
                x[j] = y[q++];
                acc -= r.a[j] * x[j];
            }
        x[k] = acc * inv;
    }
    return true;
}

struct Tri {
    int a, b, c;
    double n[3], off;  // outward plane: n.x = off
    bool alive;
};

inline void sub3(const double* a, const double* b, double* o) { o[0] = a[0] - b[0], o[1] = a[1] - b[1], o[2] = a[2] - b[2]; }
inline void cross(const double* a, const double* b, double* o) {
    o[0] = a[1] * b[2] - a[2] * b[1], o[1] = a[2] * b[0] - a[0] * b[2], o[2] = a[0] * b[1] - a[1] * b[0];
}
inline double dot(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// Incremental convex hull of n points (3*n).  Triangles are oriented with outward normals.
bool convex_hull(const std::vector<double>& P, int n, std::vector<Tri>& tris) {
    if (n < 4) return false;
    auto pt = [&](int i) { return &P[3 * (size_t)i]; };
    double scale = 0.0;
    for (double v : P) scale = std::max(scale, std::fabs(v));
    if (!(scale > 0.0)) return false;
    const double eps = 1e-11 * scale;
    // initial simplex
    int i0 = 0, i1 = 0;
    for (int i = 1; i < n; i++) {
        if (pt(i)[0] < pt(i0)[0]) i0 = i;
        if (pt(i)[0] > pt(i1)[0]) i1 = i;
    }
    if (i0 == i1) return false;
    double e01[3];
    sub3(pt(i1), pt(i0), e01);
    int i2 = -1;
    double best = 0.0;
    for (int i = 0; i < n; i++) {
        double e[3], cr[3];
        sub3(pt(i), pt(i0), e);
        cross(e01, e, cr);
        const double v = dot(cr, cr);
        if (v > best) best = v, i2 = i;
    }
    if (i2 < 0 || best < eps * eps) return false;
    double e02[3], nrm[3];
    sub3(pt(i2), pt(i0), e02);
    cross(e01, e02, nrm);
    int i3 = -1;
    best = 0.0;
    for (int i = 0; i < n; i++) {
        double e[3];
        sub3(pt(i), pt(i0), e);
        const double v = std::fabs(dot(nrm, e));
        if (v > best) best = v, i3 = i;
    }
    if (i3 < 0 || best < eps * std::sqrt(dot(nrm, nrm))) return false;
    double centroid[3];
    for (int d = 0; d < 3; d++) centroid[d] = 0.25 * (pt(i0)[d] + pt(i1)[d] + pt(i2)[d] + pt(i3)[d]);
    auto add_tri = [&](int a, int b, int c) {
        Tri t;
        t.a = a, t.b = b, t.c = c, t.alive = true;
        double u[3], v[3];
        sub3(pt(b), pt(a), u);
        sub3(pt(c), pt(a), v);
        cross(u, v, t.n);
        const double len = std::sqrt(dot(t.n, t.n));
        if (len > 0) t.n[0] /= len, t.n[1] /= len, t.n[2] /= len;
        t.off = dot(t.n, pt(a));
        if (dot(t.n, centroid) > t.off) {  // flip to face outward
            std::swap(t.b, t.c);
            t.n[0] = -t.n[0], t.n[1] = -t.n[1], t.n[2] = -t.n[2];
            t.off = -t.off;
        }
        tris.push_back(t);
    };
    tris.clear();
    add_tri(i0, i1, i2);
    add_tri(i0, i1, i3);
    add_tri(i0, i2, i3);
    add_tri(i1, i2, i3);
    std::vector<char> used(n, 0);
    used[i0] = used[i1] = used[i2] = used[i3] = 1;
    std::vector<int> visible;
    std::vector<std::pair<int, int>> horizon;
    for (int p = 0; p < n; p++) {
        if (used[p]) continue;
        visible.clear();
        for (int t = 0; t < (int)tris.size(); t++)
            if (tris[t].alive && dot(tris[t].n, pt(p)) - tris[t].off > eps) visible.push_back(t);
        if (visible.empty()) continue;
        // horizon = directed edges of visible faces whose reverse edge is not on a visible face
        horizon.clear();
        std::set<std::pair<int, int>> vis_edges;
        for (int t : visible) {
            vis_edges.insert({tris[t].a, tris[t].b});
            vis_edges.insert({tris[t].b, tris[t].c});
            vis_edges.insert({tris[t].c, tris[t].a});
        }
        for (int t : visible) {
            const int e[3][2] = {{tris[t].a, tris[t].b}, {tris[t].b, tris[t].c}, {tris[t].c, tris[t].a}};
            for (auto& ed : e)
                if (!vis_edges.count({ed[1], ed[0]})) horizon.push_back({ed[0], ed[1]});
        }
        for (int t : visible) tris[t].alive = false;
        for (auto& ed : horizon) add_tri(ed.first, ed.second, p);
    }
    std::vector<Tri> keep;
    for (auto& t : tris)
        if (t.alive) keep.push_back(t);
    tris.swap(keep);
    return tris.size() >= 4;
}

}  // namespace

bool find_interior(const double* faces, int F, double interior[3]) {
    if (F < 4) return false;
    double shift[3] = {0, 0, 0};
    for (int i = 0; i < F; i++)
        for (int d = 0; d < 3; d++) shift[d] += faces[6 * i + 3 + d] / F;
    double extent = 1.0;
    for (int i = 0; i < F; i++)
        for (int d = 0; d < 3; d++) extent = std::max(extent, std::fabs(faces[6 * i + 3 + d] - shift[d]));
    const double BIG = 1.0e3 * extent;
    std::vector<Row> rows;
    rows.reserve(F + 8);
    for (int j = 0; j < 4; j++) {
        Row p{}, n{};
        p.a[j] = 1.0, p.b = BIG;
        n.a[j] = -1.0, n.b = BIG;
        rows.push_back(p);
        rows.push_back(n);
    }
    // deterministic shuffle of the face order (expected-linear behaviour of the incremental LP)
    std::vector<int> order(F);
    for (int i = 0; i < F; i++) order[i] = i;
    uint64_t s = 0x9E3779B97F4A7C15ull ^ (uint64_t)F;
    for (int i = F - 1; i > 0; i--) {
        s = s * 6364136223846793005ull + 1442695040888963407ull;
        std::swap(order[i], order[(int)((s >> 33) % (uint64_t)(i + 1))]);
    }
    for (int oi = 0; oi < F; oi++) {
        const double* f = faces + 6 * order[oi];
        Row r{};
        r.a[0] = f[0], r.a[1] = f[1], r.a[2] = f[2], r.a[3] = 1.0;
        r.b = f[0] * (f[3] - shift[0]) + f[1] * (f[4] - shift[1]) + f[2] * (f[5] - shift[2]);
        rows.push_back(r);
    }
    const double c[4] = {0, 0, 0, -1.0};
    double x[4];
    if (!lp_solve(4, rows, (int)rows.size(), c, x)) return false;
    const double t = x[3];
    if (!(t > 1e-9) || !(t < 0.25 * BIG)) return false;  // empty interior / unbounded (geoutils.hpp:59)
    for (int d = 0; d < 3; d++) interior[d] = x[d] + shift[d];
    return true;
}

bool enumerate_vertices(const double* faces, int F, std::vector<double>& out, double epsilon) {
    double inner[3];
    if (!find_interior(faces, F, inner)) return false;
    // polar dual, geoutils.hpp:108-110:  A.col(i) = n_i / (n_i.p_i - n_i.inner)
    std::vector<double> dual(3 * (size_t)F);
    for (int i = 0; i < F; i++) {
        const double* f = faces + 6 * i;
        const double b = (f[0] * f[3] + f[1] * f[4] + f[2] * f[5]) - (inner[0] * f[0] + inner[1] * f[1] + inner[2] * f[2]);
        if (!(b > 0)) return false;
        for (int d = 0; d < 3; d++) dual[3 * (size_t)i + d] = f[d] / b;
    }
    std::vector<Tri> tris;
    if (!convex_hull(dual, F, tris)) return false;
    // geoutils.hpp:117-126
    std::vector<double> raw;
    raw.reserve(3 * tris.size());
    for (const Tri& t : tris) {
        const double *p0 = &dual[3 * (size_t)t.a], *p1 = &dual[3 * (size_t)t.b], *p2 = &dual[3 * (size_t)t.c];
        double e0[3], e1[3], nrm[3];
        sub3(p1, p0, e0);
        sub3(p2, p1, e1);
        cross(e0, e1, nrm);
        const double den = dot(nrm, p1);
        if (!(std::fabs(den) > 0)) return false;
        for (int d = 0; d < 3; d++) raw.push_back(nrm[d] / den);
    }
    // geoutils.hpp:75-97 first-seen de-duplication on a grid
    double mx = -INFINITY, mn = INFINITY;
    for (double v : raw) mx = std::max(mx, v), mn = std::min(mn, v);
    const double mag = std::max(std::fabs(mx), std::fabs(mn));
    const double res = mag * std::max(std::fabs(epsilon) / mag, DBL_EPSILON);
    std::set<std::array<double, 3>> seen;
    out.clear();
    for (size_t i = 0; i < raw.size() / 3; i++) {
        std::array<double, 3> q = {std::round(raw[3 * i] / res), std::round(raw[3 * i + 1] / res), std::round(raw[3 * i + 2] / res)};
        if (seen.insert(q).second)
            for (int d = 0; d < 3; d++) out.push_back(raw[3 * i + d] + inner[d]);
    }
    return out.size() >= 12;  // at least a tetrahedron
}

}  // namespace frx
