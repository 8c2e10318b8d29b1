// H-polytope -> V-polytope, host side: the sequential form of the enumeration in geometry_core.hpp (the batched device form
// is geometry_kernels.cu; both give the same vertices in the same order).  Replaces the reference's
// geoutils::findInterior (geoutils.hpp:43-60) and geoutils::enumerateVs (geoutils.hpp:102-131).
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <vector>

#include "geometry_core.hpp"
#include "host_model.hpp"

namespace frx {

// Candidates (x, y, z, key) sorted by key -> accepted vertices: first-seen merge, then the interior test.  Shared by the host
// enumeration below and by the post-processing of the device kernel's output (frx_capi.cu).
bool finish_vertices(const double* nd, int F, std::vector<GeoCandidate>& cand, std::vector<double>& out) {
    std::sort(cand.begin(), cand.end(), [](const GeoCandidate& a, const GeoCandidate& b) { return a.key < b.key; });
    out.clear();
    for (const GeoCandidate& c : cand) {
        bool dup = false;
        for (size_t v = 0; v < out.size() / 3 && !dup; v++)
            dup = std::fabs(out[3 * v] - c.x[0]) < geo::MERGE_TOL && std::fabs(out[3 * v + 1] - c.x[1]) < geo::MERGE_TOL &&
                  std::fabs(out[3 * v + 2] - c.x[2]) < geo::MERGE_TOL;
        if (!dup) out.insert(out.end(), c.x, c.x + 3);
    }
    const int V = (int)out.size() / 3;
    if (V < 4) return false;  // at least a tetrahedron
    double cen[3] = {0, 0, 0};
    for (int v = 0; v < V; v++)
        for (int d = 0; d < 3; d++) cen[d] += out[3 * v + d] / V;
    for (int k = 0; k < F; k++)  // strictly interior centroid <=> non-empty interior (geoutils.hpp:59: minmaxsd < 0)
        if (!(nd[4 * k] * cen[0] + nd[4 * k + 1] * cen[1] + nd[4 * k + 2] * cen[2] - nd[4 * k + 3] < -geo::FEAS_TOL)) return false;
    return true;
}

void pack_planes(const double* faces, int F, double* nd) {
    for (int k = 0; k < F; k++) {
        const double* f = faces + 6 * (size_t)k;
        nd[4 * k] = f[0], nd[4 * k + 1] = f[1], nd[4 * k + 2] = f[2];
        nd[4 * k + 3] = f[0] * f[3] + f[1] * f[4] + f[2] * f[5];
    }
}

bool enumerate_vertices(const double* faces, int F, std::vector<double>& out) {
    if (F < 4 || F > geo::MAX_FACES) return false;
    std::vector<double> nd(4 * (size_t)F);
    pack_planes(faces, F, nd.data());
    for (int a = 0; a < F; a++)
        for (int b = a + 1; b < F; b++)
            if (geo::pair_ray(nd.data(), F, a, b, 1.0) || geo::pair_ray(nd.data(), F, a, b, -1.0)) return false;  // unbounded
    std::vector<GeoCandidate> cand;
    for (int a = 0; a < F; a++)
        for (int b = a + 1; b < F; b++)
            for (int c = b + 1; c < F; c++) {
                GeoCandidate g;
                if (geo::triple_vertex(nd.data(), F, a, b, c, g.x)) {
                    g.key = geo::triple_key(a, b, c);
                    cand.push_back(g);
                }
            }
    return finish_vertices(nd.data(), F, cand, out);
}

}  // namespace frx
