"""Parallel cyclic reduction on the reduced MINCO system (2x2 blocks, SPD block tridiagonal): numpy emulation of the slot algebra of
pcr_solve() in eval_kernels.cu (symmetric storage: D (3), A = coupling to the node s below (4), C implied = A_{j+s}^T), checked
against a dense solve."""
import numpy as np

def build(T):
    M = len(T); N = M - 1
    iT = 1.0 / T
    D = np.zeros((N + 2, 3)); A = np.zeros((N + 2, 4))
    for j in range(1, M):
        ia, ib = iT[j - 1], iT[j]
        D[j] = [192 * (ia**3 + ib**3), 36 * (ib**2 - ia**2), 9 * (ia + ib)]
        A[j] = [168 * ia**3, 24 * ia**2, -24 * ia**2, -3 * ia]
    A[1] = 0.0
    return D, A

def dense(D, A, N):
    S = np.zeros((2 * N, 2 * N))
    for j in range(1, N + 1):
        d11, d12, d22 = D[j]
        S[2*(j-1):2*j, 2*(j-1):2*j] = [[d11, d12], [d12, d22]]
        if j >= 2:
            L = A[j].reshape(2, 2)
            S[2*(j-1):2*j, 2*(j-2):2*(j-1)] = L
            S[2*(j-2):2*(j-1), 2*(j-1):2*j] = L.T
    return S

def inv3(d):
    det = d[0] * d[2] - d[1] * d[1]
    return np.array([d[2] / det, -d[1] / det, d[0] / det])

def pcr(D, A, R, N):
    D = D.copy(); A = A.copy(); R = R.copy()
    I = np.zeros_like(D)
    for j in range(1, N + 1): I[j] = inv3(D[j])
    s = 1
    while s < N:
        nD, nA, nR = D.copy(), np.zeros_like(A), R.copy()
        for j in range(1, N + 1):
            a11, a12, a21, a22 = A[j]
            lo, hi = j - s, j + s
            if lo >= 1:
                i11, i12, i22 = I[lo]
                al = -np.array([a11*i11 + a12*i12, a11*i12 + a12*i22, a21*i11 + a22*i12, a21*i12 + a22*i22])
                nD[j] += [al[0]*a11 + al[1]*a12, al[0]*a21 + al[1]*a22, al[2]*a21 + al[3]*a22]
                b = A[lo]
                nA[j] = [al[0]*b[0] + al[1]*b[2], al[0]*b[1] + al[1]*b[3], al[2]*b[0] + al[3]*b[2], al[2]*b[1] + al[3]*b[3]]
                for d in range(3):
                    nR[j, d] += al[0]*R[lo, d] + al[1]*R[lo, 3+d]
                    nR[j, 3+d] += al[2]*R[lo, d] + al[3]*R[lo, 3+d]
            if hi <= N:
                i11, i12, i22 = I[hi]
                h11, h12, h21, h22 = A[hi]
                ga = -np.array([h11*i11 + h21*i12, h11*i12 + h21*i22, h12*i11 + h22*i12, h12*i12 + h22*i22])  # -A_hi^T I_hi
                nD[j] += [ga[0]*h11 + ga[1]*h21, ga[0]*h12 + ga[1]*h22, ga[2]*h12 + ga[3]*h22]
                for d in range(3):
                    nR[j, d] += ga[0]*R[hi, d] + ga[1]*R[hi, 3+d]
                    nR[j, 3+d] += ga[2]*R[hi, d] + ga[3]*R[hi, 3+d]
        D, A, R = nD, nA, nR
        for j in range(1, N + 1): I[j] = inv3(D[j])
        s *= 2
    Y = np.zeros_like(R)
    for j in range(1, N + 1):
        i11, i12, i22 = I[j]
        for d in range(3):
            Y[j, d] = i11 * R[j, d] + i12 * R[j, 3+d]
            Y[j, 3+d] = i12 * R[j, d] + i22 * R[j, 3+d]
    return Y

if __name__ == "__main__":
    rng = np.random.default_rng(0)
    worst = 0.0
    for M in (2, 3, 4, 5, 8, 17, 33, 64, 65):
        for trial in range(5):
            T = np.exp(rng.normal(0, 1.0, size=M))  # durations spread over two decades
            N = M - 1
            D, A = build(T)
            R = rng.normal(size=(N + 2, 6)) * 100
            Y = pcr(D, A, R, N)
            S = dense(D, A, N)
            for d in range(3):
                rhs = np.zeros(2 * N)
                rhs[0::2] = R[1:N+1, d]; rhs[1::2] = R[1:N+1, 3+d]
                y = np.linalg.solve(S, rhs)
                got = np.zeros(2 * N); got[0::2] = Y[1:N+1, d]; got[1::2] = Y[1:N+1, 3+d]
                err = np.abs(got - y).max() / np.abs(y).max()
                worst = max(worst, err)
    print("worst relative error vs dense solve:", worst)
