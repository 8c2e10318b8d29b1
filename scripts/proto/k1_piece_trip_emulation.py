"""Index-level emulation (numpy, CPU) of the NEXT K1 elimination scheme: NP pivots per trip (NP = 2 is today's kernel,
NP = 3 and NP = 6 = a whole piece are the candidates; csrc/eval_kernels.cu carries the CUDA version under FRX_K1_NP).

Every lane (r, cg) of the pivot warp owns row i = k0+6+r and the columns k0+6+cg, k0+9+cg and right-hand side cg (as today),
and recomputes - redundantly, in registers - the LU of the 6 x 6 diagonal block of the pivot rows k0..k0+5 and those rows'
updates at ITS columns, so that a trip needs ONE shared-memory load round and ONE store round (today: three two-pivot trips).
Per element the operations are exactly the sequential elimination's (a_ij -= l_ip u_pj for p ascending), so the CUDA version
must be bit-identical to the current kernel; this script checks the slot algebra against a dense no-pivot LU.

Band storage (as csrc/eval_kernels.cu): row i, 16 slots: slot(c) = c - i + 6 for columns i-6..i+6, slots 13..15 = right-hand sides.
After elimination: slots 0..5 multipliers, slot 6 the RECIPROCAL of u_ii, 7..12 upper band, 13..15 forward-substituted rhs."""
import numpy as np


def band_from_dense(A, B):
    N = A.shape[0]
    rows = np.zeros((N + 12, 16))
    for i in range(N):
        for c in range(max(0, i - 6), min(N, i + 7)):
            rows[i, c - i + 6] = A[i, c]
        rows[i, 13:16] = B[i]
    return rows


def trip(rows, k0, NP):
    """One trip: pivots k0..k0+NP-1; lane (r, cg) owns row k0+NP+r, columns k0+NP+cg and k0+NP+3+cg, right-hand side cg.
    `rows` is the (unbounded) ring; loads first, then stores, as the warp would."""
    stores = []
    for r in range(6):
        for cg in range(3):
            i = k0 + NP + r
            # ---- loads (every lane) ----
            D = np.zeros((NP, NP))
            for a in range(NP):
                for b in range(NP):
                    D[a, b] = rows[k0 + a, b - a + 6]
            rcp = np.zeros(NP)
            rcp[0] = D[0, 0]  # slot 6 of the first pivot row already holds 1/u (stored by the previous trip / the prologue)
            Cx = np.array([rows[k0 + a, NP + cg - a + 6] if a >= NP + cg - 6 else 0.0 for a in range(NP)])
            Cy = np.array([rows[k0 + a, NP + cg - a + 9] if a >= NP + cg - 3 else 0.0 for a in range(NP)])
            Rr = np.array([rows[k0 + a, 13 + cg] for a in range(NP)])
            Mm = np.array([rows[i, b - NP - r + 6] if b >= NP + r - 6 else 0.0 for b in range(NP)])
            Ex, Ey, Er = rows[i, 6 + cg - r], rows[i, 9 + cg - r], rows[i, 13 + cg]
            # ---- arithmetic: right-looking, pivot p ascending ----
            for p in range(NP):
                if p > 0:
                    rcp[p] = 1.0 / D[p, p]
                for a in range(p + 1, NP):
                    l = D[a, p] * rcp[p]
                    D[a, p] = l
                    for b in range(p + 1, NP):
                        D[a, b] -= l * D[p, b]
                    Cx[a] -= l * Cx[p]
                    Cy[a] -= l * Cy[p]
                    Rr[a] -= l * Rr[p]
                l = Mm[p] * rcp[p]
                Mm[p] = l
                for b in range(p + 1, NP):
                    Mm[b] -= l * D[p, b]
                Ex -= l * Cx[p]
                Ey -= l * Cy[p]
                Er -= l * Rr[p]
            # ---- stores ----
            first = (r == 0 and cg == 0)
            stores.append((i, 6 + cg - r, 1.0 / Ex if first else Ex))  # the next trip's first pivot: store 1/u
            stores.append((i, 9 + cg - r, Ey))
            stores.append((i, 13 + cg, Er))
            if cg == 0:
                for b in range(max(0, NP + r - 6), NP):
                    stores.append((i, b - NP - r + 6, Mm[b]))  # multipliers of the own row
            if r < NP:  # this lane also publishes pivot row a = r at its columns (rows a >= 6 do not exist: NP <= 6)
                a = r
                if a >= NP + cg - 6:
                    stores.append((k0 + a, NP + cg - a + 6, Cx[a]))
                if a >= NP + cg - 3:
                    stores.append((k0 + a, NP + cg - a + 9, Cy[a]))
                stores.append((k0 + a, 13 + cg, Rr[a]))
                if cg == 0:
                    for b in range(NP):  # row a of the diagonal block: multipliers | 1/u | upper part
                        stores.append((k0 + a, b - a + 6, rcp[a] if b == a else D[a, b]))
    for (i, s, v) in stores:
        rows[i, s] = v


def check(M, seed, NP):
    rng = np.random.default_rng(seed)
    N = 6 * M
    A = np.zeros((N, N))
    for i in range(N):
        for c in range(max(0, i - 6), min(N, i + 7)):
            A[i, c] = rng.normal()
        A[i, i] += 8.0  # keeps the no-pivot LU well behaved
    B = rng.normal(size=(N, 3))
    rows = band_from_dense(A, B)
    rows[0, 6] = 1.0 / rows[0, 6]
    for k0 in range(0, N, NP):
        trip(rows, k0, NP)
    # dense reference: no-pivot LU with forward substitution
    L, U, Y = np.eye(N), A.copy(), B.copy()
    for k in range(N - 1):
        for i in range(k + 1, min(N, k + 7)):
            l = U[i, k] / U[k, k]
            L[i, k] = l
            U[i, k:] -= l * U[k, k:]
            Y[i] -= l * Y[k]
    err = 0.0
    for i in range(N):
        err = max(err, abs(rows[i, 6] - 1.0 / U[i, i]) * abs(U[i, i]))
        for c in range(max(0, i - 6), i):
            err = max(err, abs(rows[i, c - i + 6] - L[i, c]))
        for c in range(i + 1, min(N, i + 7)):
            err = max(err, abs(rows[i, c - i + 6] - U[i, c]))
        err = max(err, np.abs(rows[i, 13:16] - Y[i]).max())
    return err


if __name__ == "__main__":
    for NP in (2, 3, 6):
        for M, seed in ((1, 0), (2, 1), (5, 2), (16, 3)):
            e = check(M, seed, NP)
            print(f"NP={NP} M={M}: max deviation from the dense no-pivot LU {e:.2e}")
            assert e < 1e-10
    print("OK")
