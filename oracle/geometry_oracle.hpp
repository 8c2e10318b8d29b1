// ORACLE — TEST INFRASTRUCTURE ONLY (see lbfgs_rc.hpp header).
//
// Independent, deliberately naive vertex enumeration of an H-polytope used to CHECK the
// product's V-polytope extraction (the reference's geoutils::enumerateVs, geoutils.hpp:102-149,
// is LP interior point + dual quickhull + first-seen dedup).  Here: every triple of faces is
// intersected (Cramer), kept if it satisfies all half-spaces within tol, and deduplicated.
// O(F^4) — fine for the test sizes.  Vertex ORDER is not reproduced (it is an artefact of
// quickhull's triangle order); tests compare vertex SETS.
#pragma once
#include <cmath>
#include <vector>

namespace frx_oracle {

// faces: 6*F (unit normal n, point p on the plane).  Returns 3*V vertex coordinates.
inline std::vector<double> brute_force_vertices(const double* faces, int F, double tol = 1e-9, double dedup = 1e-6) {
    std::vector<double> out;
    auto offs = [&](int k) { return faces[6 * k] * faces[6 * k + 3] + faces[6 * k + 1] * faces[6 * k + 4] + faces[6 * k + 2] * faces[6 * k + 5]; };
    for (int a = 0; a < F; a++)
        for (int b = a + 1; b < F; b++)
            for (int c = b + 1; c < F; c++) {
                const double *A = faces + 6 * a, *B = faces + 6 * b, *C = faces + 6 * c;
                const double det = A[0] * (B[1] * C[2] - B[2] * C[1]) - A[1] * (B[0] * C[2] - B[2] * C[0]) + A[2] * (B[0] * C[1] - B[1] * C[0]);
                if (std::fabs(det) < 1e-10) continue;
                const double da = offs(a), db = offs(b), dc = offs(c);
                const double x = (da * (B[1] * C[2] - B[2] * C[1]) - A[1] * (db * C[2] - B[2] * dc) + A[2] * (db * C[1] - B[1] * dc)) / det;
                const double y = (A[0] * (db * C[2] - B[2] * dc) - da * (B[0] * C[2] - B[2] * C[0]) + A[2] * (B[0] * dc - db * C[0])) / det;
                const double z = (A[0] * (B[1] * dc - db * C[1]) - A[1] * (B[0] * dc - db * C[0]) + da * (B[0] * C[1] - B[1] * C[0])) / det;
                bool ok = true;
                for (int k = 0; k < F && ok; k++) {
                    const double* f = faces + 6 * k;
                    if (f[0] * (x - f[3]) + f[1] * (y - f[4]) + f[2] * (z - f[5]) > tol) ok = false;
                }
                if (!ok) continue;
                bool dup = false;
                for (size_t v = 0; v < out.size() / 3 && !dup; v++)
                    if (std::fabs(out[3 * v] - x) < dedup && std::fabs(out[3 * v + 1] - y) < dedup && std::fabs(out[3 * v + 2] - z) < dedup) dup = true;
                if (!dup) {
                    out.push_back(x);
                    out.push_back(y);
                    out.push_back(z);
                }
            }
    return out;
}

}  // namespace frx_oracle
