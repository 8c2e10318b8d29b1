"""Prototype (numpy) of the reduced MINCO formulation used by the fast-mode kernels.

The reference solves the 6M x 6M banded system A(T) c = b(q) (MINCO_S3::generate, se3gcopter_cpu.hpp:425-505) by a
no-pivot banded LU (trajectory.hpp:655-719) and back-propagates with A^T lambda = dJ/dc (:724-751) plus addPropCtoT/P
(se3gcopter_cpu.hpp:104-161).  The same c and the same gradients follow from a much smaller system: a quintic piece is
fixed by its end states (p, v, a), the waypoints fix p at the nodes, and the remaining rows of A (jerk and snap continuity)
are the stationarity conditions of the jerk energy w.r.t. the node velocities/accelerations y_j = (v_j, a_j), j = 1..M-1:

    S(T) y = r(q, T),   S symmetric positive definite, block tridiagonal with 2 x 2 blocks.

This script checks forward (c) and backward (dJ/dT, dJ/dq) of that formulation against a dense restatement of the
reference's system.  Run:  python scripts/proto/reduced_minco.py
"""
import numpy as np


# ---------- dense restatement of the reference system ----------
def dense_A(T, M):
    N = 6 * M
    A = np.zeros((N, N))
    A[0, 0] = 1.0
    A[1, 1] = 1.0
    A[2, 2] = 2.0
    for i in range(M - 1):
        t = T[i]
        r, c = 6 * i + 3, 6 * i
        A[r, c + 3:c + 6] = [6.0, 24.0 * t, 60.0 * t**2]
        A[r, c + 9] = -6.0
        A[r + 1, c + 4:c + 6] = [24.0, 120.0 * t]
        A[r + 1, c + 10] = -24.0
        A[r + 2, c:c + 6] = [1, t, t**2, t**3, t**4, t**5]
        A[r + 3, c:c + 6] = [1, t, t**2, t**3, t**4, t**5]
        A[r + 3, c + 6] = -1.0
        A[r + 4, c + 1:c + 6] = [1, 2 * t, 3 * t**2, 4 * t**3, 5 * t**4]
        A[r + 4, c + 7] = -1.0
        A[r + 5, c + 2:c + 6] = [2, 6 * t, 12 * t**2, 20 * t**3]
        A[r + 5, c + 8] = -2.0
    t = T[M - 1]
    c = 6 * (M - 1)
    A[N - 3, c:c + 6] = [1, t, t**2, t**3, t**4, t**5]
    A[N - 2, c + 1:c + 6] = [1, 2 * t, 3 * t**2, 4 * t**3, 5 * t**4]
    A[N - 1, c + 2:c + 6] = [2, 6 * t, 12 * t**2, 20 * t**3]
    return A


def dense_forward(T, q, head, tail):
    M = len(T)
    A = dense_A(T, M)
    b = np.zeros((6 * M, 3))
    b[0:3] = head  # rows p, v, a
    for i in range(M - 1):
        b[6 * i + 5] = q[i]
    b[6 * M - 3:] = tail
    return np.linalg.solve(A, b), A


def dense_backward(T, c, A, gdC):
    M = len(T)
    lam = np.linalg.solve(A.T, gdC)
    gT = np.zeros(M)
    gq = np.zeros((M - 1, 3))
    for i in range(M):
        t = T[i]
        ci = c[6 * i:6 * i + 6]
        nvel = -(ci[1] + 2 * t * ci[2] + 3 * t**2 * ci[3] + 4 * t**3 * ci[4] + 5 * t**4 * ci[5])
        nacc = -(2 * ci[2] + 6 * t * ci[3] + 12 * t**2 * ci[4] + 20 * t**3 * ci[5])
        njer = -(6 * ci[3] + 24 * t * ci[4] + 60 * t**2 * ci[5])
        if i < M - 1:
            nsnp = -(24 * ci[4] + 120 * t * ci[5])
            ncrk = -120 * ci[5]
            B1 = np.stack([nsnp, ncrk, nvel, nvel, nacc, njer])
            gT[i] = (B1 * lam[6 * i + 3:6 * i + 9]).sum()
            gq[i] = lam[6 * i + 5]
        else:
            B2 = np.stack([nvel, nacc, njer])
            gT[i] = (B2 * lam[6 * M - 3:]).sum()
    return gT, gq


# ---------- reduced formulation ----------
def hermite(P0, V0, A0, P1, V1, A1, t):
    """Quintic through end states: rows c0..c5 (each a 3-vector)."""
    d = P1 - P0
    i1, i2, i3, i4, i5 = 1 / t, 1 / t**2, 1 / t**3, 1 / t**4, 1 / t**5
    c3 = 10 * d * i3 - (4 * V1 + 6 * V0) * i2 - (3 * A0 - A1) * (0.5 * i1)
    c4 = -15 * d * i4 + (7 * V1 + 8 * V0) * i3 + (3 * A0 - 2 * A1) * (0.5 * i2)
    c5 = 6 * d * i5 - 3 * (V1 + V0) * i4 + (A1 - A0) * (0.5 * i3)
    return np.stack([P0, V0, 0.5 * A0, c3, c4, c5])


def hermite_T(g, t):
    """Adjoint of `hermite` w.r.t. the end states: g = 6 x 3 -> (P0b, V0b, A0b, P1b, V1b, A1b)."""
    i1, i2, i3, i4, i5 = 1 / t, 1 / t**2, 1 / t**3, 1 / t**4, 1 / t**5
    dP = 10 * i3 * g[3] - 15 * i4 * g[4] + 6 * i5 * g[5]
    P0b = g[0] - dP
    P1b = dP
    V0b = g[1] - 6 * i2 * g[3] + 8 * i3 * g[4] - 3 * i4 * g[5]
    V1b = -4 * i2 * g[3] + 7 * i3 * g[4] - 3 * i4 * g[5]
    A0b = 0.5 * g[2] - 1.5 * i1 * g[3] + 1.5 * i2 * g[4] - 0.5 * i3 * g[5]
    A1b = 0.5 * i1 * g[3] - i2 * g[4] + 0.5 * i3 * g[5]
    return P0b, V0b, A0b, P1b, V1b, A1b


def blocks(T):
    """Block tridiagonal S: D[j] (2x2 sym), L[j] (couples node j to node j-1), nodes j = 1..M-1 stored at index j."""
    M = len(T)
    D = np.zeros((M, 2, 2))
    L = np.zeros((M, 2, 2))
    for j in range(1, M):
        ta, tb = T[j - 1], T[j]
        D[j] = [[192 / ta**3 + 192 / tb**3, 36 / tb**2 - 36 / ta**2], [36 / tb**2 - 36 / ta**2, 9 / ta + 9 / tb]]
        L[j] = [[168 / ta**3, 24 / ta**2], [-24 / ta**2, -3 / ta]]  # rows (V_j, A_j), cols (V_{j-1}, A_{j-1})
    return D, L


def rhs(T, P, head, tail):
    """r_j: everything of the node-j rows (-snap jump; jerk jump) that does not multiply an unknown."""
    M = len(T)
    r = np.zeros((M, 2, 3))
    for j in range(1, M):
        ta, tb = T[j - 1], T[j]
        da, db = P[j] - P[j - 1], P[j + 1] - P[j]
        r[j, 0] = 360 * da / ta**4 + 360 * db / tb**4
        r[j, 1] = -60 * da / ta**3 + 60 * db / tb**3
    # known neighbours: node 0 (head v, a) and node M (tail v, a)
    D, L = blocks(T)
    if M > 1:
        r[1] -= L[1] @ np.stack([head[1], head[2]])
        tb = T[M - 1]
        U = np.array([[168 / tb**3, -24 / tb**2], [24 / tb**2, -3 / tb]])
        r[M - 1] -= U @ np.stack([tail[1], tail[2]])
    return r


def thomas(D, L, r):
    """Block LDL^T (no pivoting: S is SPD) for nodes 1..M-1; returns y and the inverse pivot blocks."""
    M = D.shape[0]
    Di = np.zeros_like(D)
    z = r.copy()
    Dp = D.copy()
    for j in range(1, M):
        if j > 1:
            G = L[j] @ Di[j - 1]
            Dp[j] = D[j] - G @ L[j].T
            z[j] = r[j] - G @ z[j - 1]
        Di[j] = np.linalg.inv(Dp[j])
    y = np.zeros_like(r)
    for j in range(M - 1, 0, -1):
        t = z[j].copy()
        if j < M - 1:
            t -= L[j + 1].T @ y[j + 1]
        y[j] = Di[j] @ t
    return y


def reduced_forward(T, q, head, tail):
    M = len(T)
    P = np.vstack([head[0][None], q.reshape(M - 1, 3), tail[0][None]])
    D, L = blocks(T)
    y = thomas(D, L, rhs(T, P, head, tail))
    V = np.vstack([head[1][None], y[1:, 0], tail[1][None]])
    Ac = np.vstack([head[2][None], y[1:, 1], tail[2][None]])
    c = np.zeros((6 * M, 3))
    for i in range(M):
        c[6 * i:6 * i + 6] = hermite(P[i], V[i], Ac[i], P[i + 1], V[i + 1], Ac[i + 1], T[i])
    return c, (P, V, Ac)


def reduced_backward(T, c, nodes, gdC):
    M = len(T)
    P, V, Ac = nodes
    Pb = np.zeros((M + 1, 3))
    yb = np.zeros((M, 2, 3))
    gT = np.zeros(M)
    for i in range(M):
        t = T[i]
        ci = c[6 * i:6 * i + 6]
        P0b, V0b, A0b, P1b, V1b, A1b = hermite_T(gdC[6 * i:6 * i + 6], t)
        Pb[i] += P0b
        Pb[i + 1] += P1b
        if i >= 1:
            yb[i, 0] += V0b
            yb[i, 1] += A0b
        if i + 1 <= M - 1:
            yb[i + 1, 0] += V1b
            yb[i + 1, 1] += A1b
        J1 = 6 * ci[3] + 24 * t * ci[4] + 60 * t**2 * ci[5]
        gT[i] = -((P1b * V[i + 1]).sum() + (V1b * Ac[i + 1]).sum() + (A1b * J1).sum())  # dc/dT at fixed end states
    D, L = blocks(T)
    mu = thomas(D, L, yb)
    muV = np.vstack([np.zeros((1, 3)), mu[1:, 0], np.zeros((1, 3))])
    muA = np.vstack([np.zeros((1, 3)), mu[1:, 1], np.zeros((1, 3))])
    z3 = np.zeros(3)
    gP = Pb.copy()
    for i in range(M):
        t = T[i]
        ci = c[6 * i:6 * i + 6]
        ct = hermite(z3, muV[i], muA[i], z3, muV[i + 1], muA[i + 1], t)  # the adjoint quintic of piece i
        # d/dP of  s^T Q dels : crackle of the adjoint quintic at the ends
        gP[i] -= -120 * ct[5]
        gP[i + 1] -= 120 * ct[5]
        jc = 6 * ci[3] + 24 * t * ci[4] + 60 * t**2 * ci[5]
        sc = 24 * ci[4] + 120 * t * ci[5]
        kc = 120 * ci[5]
        jt = 6 * ct[3] + 24 * t * ct[4] + 60 * t**2 * ct[5]
        st = 24 * ct[4] + 120 * t * ct[5]
        kt = 120 * ct[5]
        dB = (-jc * jt + st * Ac[i + 1] + sc * muA[i + 1] - kt * V[i + 1] - kc * muV[i + 1]).sum()
        gT[i] -= dB
    return gT, gP[1:M]


def main():
    rng = np.random.default_rng(3)
    worst = 0.0
    for M in (1, 2, 3, 8, 33, 64, 150):
        T = rng.uniform(0.3, 2.5, size=M)
        q = np.cumsum(rng.normal(size=(max(M - 1, 0), 3)) * 2.0, axis=0)
        head = rng.normal(size=(3, 3))
        tail = rng.normal(size=(3, 3)) + (q[-1] if M > 1 else 0)
        c_ref, A = dense_forward(T, q, head, tail)
        c_new, nodes = reduced_forward(T, q, head, tail)
        e_c = np.abs(c_new - c_ref).max() / np.abs(c_ref).max()
        gdC = rng.normal(size=(6 * M, 3))
        gT_ref, gq_ref = dense_backward(T, c_ref, A, gdC)
        gT_new, gq_new = reduced_backward(T, c_new, nodes, gdC)
        e_t = np.abs(gT_new - gT_ref).max() / np.abs(gT_ref).max()
        e_q = np.abs(gq_new - gq_ref).max() / max(np.abs(gq_ref).max(), 1e-300) if M > 1 else 0.0
        print(f"M={M:4d}  c {e_c:.2e}  dJ/dT {e_t:.2e}  dJ/dq {e_q:.2e}   cond(A) {np.linalg.cond(A):.1e}")
        worst = max(worst, e_c, e_t, e_q)
    return worst


if __name__ == "__main__":
    w = main()
    assert w < 1e-9, w
