"""ORACLE — TEST INFRASTRUCTURE ONLY.  Reference solutions of the equality-constrained QP that SoftSQPOptimizer hands to
OSQP every iteration (include/ungar/optimization/soft_sqp.hpp:141-158, :193-233):

    min_d  1/2 d^T P d + q^T d    s.t.  A d = -g          (l = u = -g, soft_sqp.hpp:156-157)

  * kkt_solve           — the oracle: one sparse LU solve of the quasi-definite KKT system [P A^T; A -delta I] assembled from
                          the monolithic pieces exactly as the reference assembles them (any exact QP solver returns this
                          minimiser; OSQP v0.6.3 itself is absent, external/config/osqp/CMakeLists.txt.in:16)
  * schur_stagewise     — numpy statement of the stage-wise algorithm the CUDA kernel implements (block-tridiagonal Schur
                          complement on the multipliers), used to test the kernel's intermediate blocks
Both take one trajectory's KKT block record (ungar_b200_kkt_layout) of the quadruped problem.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

NX, NU, NZ, TRI, LEGS = 13, 24, 37, 703, 4
DELTA = 1e-9  # quasi-definite regularisation of the multiplier block (same value in the CUDA kernel)


def unpack_sym(packed: np.ndarray, n: int) -> np.ndarray:
    m = np.zeros((n, n))
    m[np.triu_indices(n)] = packed
    return m + np.triu(m, 1).T


def stage_blocks(rec: np.ndarray, L: dict):
    """Per-stage dense pieces of the quadruped QP: H_j, q_j, U_j, V_j, b_j in the grouping of DESIGN.md §9."""
    N = L["horizon"]
    g = rec[L["g"]:L["g"] + L["m_eq"]]
    A = rec[L["A"]:L["A"] + N * NX * NZ].reshape(N, NX, NZ)
    C = rec[L["C"]:L["C"] + N * LEGS * 80].reshape(N, LEGS, 4, 20)
    grad = rec[L["grad"]:L["grad"] + L["n_dec"]]
    H = [unpack_sym(rec[L["H"] + k * TRI:L["H"] + (k + 1) * TRI], NZ) for k in range(N)]
    H.append(unpack_sym(rec[L["HN"]:L["HN"] + 91], NX))
    q = [np.concatenate([grad[NX * k:NX * k + NX], grad[NX * (N + 1) + NU * k:NX * (N + 1) + NU * k + NU]]) for k in range(N)]
    q.append(grad[NX * N:NX * N + NX])

    def contact(k, prev):
        M = np.zeros((16, NZ))
        for leg in range(LEGS):
            blk = C[k, leg]  # [4][20]
            off = 10 if prev else 0
            M[4 * leg:4 * leg + 4, 0:7] = blk[:, off:off + 7]
            M[4 * leg:4 * leg + 4, NX + 6 * leg + 3:NX + 6 * leg + 6] = blk[:, off + 7:off + 10]
        return M

    U, V, b = [], [], []
    for j in range(N + 1):
        nzj = NZ if j < N else NX
        Ix = np.zeros((NX, nzj))
        Ix[:, :NX] = np.eye(NX)
        if j < N:
            U.append(np.vstack([Ix, contact(j, False)]))                       # nu_j = [dyn_{j-1} | c0 ; con_j] on w_j
            V.append(np.vstack([A[j], contact(j + 1, True) if j + 1 < N else np.zeros((0, NZ))]))  # nu_{j+1} on w_j
            bj = -np.concatenate([g[NX * j:NX * j + NX], g[NX * (N + 1) + 16 * j:NX * (N + 1) + 16 * j + 16]])
        else:
            U.append(Ix)
            bj = -g[NX * N:NX * N + NX]
        b.append(bj)
    return H, q, U, V, b


def to_reference_rows(nu: np.ndarray, N: int) -> np.ndarray:
    """Multipliers from group order (nu_0 .. nu_N) to the reference's row order [x_0 - x_m | defects | contact rows]."""
    lam = np.zeros(NX * (N + 1) + 16 * N)
    r0 = 0
    for j in range(N + 1):
        lam[NX * j:NX * j + NX] = nu[r0:r0 + NX]
        r0 += NX
        if j < N:
            lam[NX * (N + 1) + 16 * j:NX * (N + 1) + 16 * j + 16] = nu[r0:r0 + 16]
            r0 += 16
    return lam


def kkt_solve(rec: np.ndarray, L: dict, delta: float = DELTA):
    """(d, multipliers) from one sparse LU of the assembled KKT system; d in the reference's [X | U] order."""
    N, n, m = L["horizon"], L["n_dec"], L["m_eq"]
    H, q, U, V, b = stage_blocks(rec, L)
    xi = lambda j: np.arange(NX * j, NX * j + NX)  # noqa: E731
    wi = lambda j: np.concatenate([xi(j), NX * (N + 1) + NU * j + np.arange(NU)]) if j < N else xi(j)  # noqa: E731
    P = sp.lil_matrix((n, n))
    qq = np.zeros(n)
    for j in range(N + 1):
        P[np.ix_(wi(j), wi(j))] = H[j]
        qq[wi(j)] = q[j]
    rows, rhs, r0 = [], [], 0
    Amat = sp.lil_matrix((m, n))
    for j in range(N + 1):
        nr = U[j].shape[0]
        Amat[r0:r0 + nr, wi(j)] = U[j]
        if j > 0:
            Amat[r0:r0 + nr, wi(j - 1)] = V[j - 1]
        rhs.append(b[j])
        r0 += nr
    assert r0 == m
    K = sp.bmat([[P.tocsr(), Amat.T.tocsr()], [Amat.tocsr(), -delta * sp.identity(m)]], format="csc")
    sol = spla.spsolve(K, np.concatenate([-qq, np.concatenate(rhs)]))
    return sol[:n], to_reference_rows(sol[n:], N)


def schur_stagewise(rec: np.ndarray, L: dict, delta: float = DELTA):
    """The algorithm of the CUDA kernel: S nu = -(b + A P^-1 q) with S block tridiagonal over the groups nu_j."""
    N = L["horizon"]
    H, q, U, V, b = stage_blocks(rec, L)
    Pinv = [np.linalg.inv(h) for h in H]
    t = [Pinv[j] @ q[j] for j in range(N + 1)]
    D, E, r = [], [], []
    for j in range(N + 1):
        Sjj = U[j] @ Pinv[j] @ U[j].T + delta * np.eye(U[j].shape[0])
        rj = b[j] + U[j] @ t[j]
        if j > 0:
            Sjj += V[j - 1] @ Pinv[j - 1] @ V[j - 1].T
            rj += V[j - 1] @ t[j - 1]
            E.append(V[j - 1] @ Pinv[j - 1] @ U[j - 1].T)   # S_{j, j-1}
        D.append(Sjj)
        r.append(-rj)
    # block Cholesky, forward and backward substitution
    Ld, Lo, y = [], [], []
    for j in range(N + 1):
        Sjj = D[j].copy()
        if j > 0:
            Lo.append(np.linalg.solve(Ld[j - 1], E[j - 1].T).T)  # L_{j,j-1} = S_{j,j-1} L_{j-1,j-1}^-T
            Sjj -= Lo[j - 1] @ Lo[j - 1].T
        Ld.append(np.linalg.cholesky(Sjj))
        rhs = r[j] - (Lo[j - 1] @ y[j - 1] if j > 0 else 0.0)
        y.append(np.linalg.solve(Ld[j], rhs))
    nu = [None] * (N + 1)
    for j in range(N, -1, -1):
        rhs = y[j] - (Lo[j].T @ nu[j + 1] if j < N else 0.0)
        nu[j] = np.linalg.solve(Ld[j].T, rhs)
    d = np.zeros(L["n_dec"])
    for j in range(N + 1):
        v = q[j] + U[j].T @ nu[j] + (V[j].T @ nu[j + 1] if j < N else 0.0)
        dw = -Pinv[j] @ v
        d[NX * j:NX * j + NX] = dw[:NX]
        if j < N:
            d[NX * (N + 1) + NU * j:NX * (N + 1) + NU * j + NU] = dw[NX:]
    return d, to_reference_rows(np.concatenate(nu), N), dict(D=D, E=E, r=r)


def schur_twisted(rec: np.ndarray, L: dict, delta: float = DELTA, middle: int | None = None):
    """The round-2 CUDA kernel's algorithm: the same block-tridiagonal Schur complement, eliminated from BOTH ends at once
    (a "twisted" block Cholesky).  Groups 0 .. m-1 are eliminated top-down, groups N .. m+1 bottom-up — two independent
    dependent chains of half the length (one warp each on the device) — and meet at group m; the substitutions then run
    outward from m in both directions.  Exact; only the elimination order differs from `schur_stagewise`."""
    N = L["horizon"]
    m = N // 2 if middle is None else middle
    H, q, U, V, b = stage_blocks(rec, L)
    Pinv = [np.linalg.inv(h) for h in H]
    t = [Pinv[j] @ q[j] for j in range(N + 1)]
    D, r = [], []
    for j in range(N + 1):
        Sjj = U[j] @ Pinv[j] @ U[j].T + delta * np.eye(U[j].shape[0])
        rj = b[j] + U[j] @ t[j]
        if j > 0:
            Sjj += V[j - 1] @ Pinv[j - 1] @ V[j - 1].T
            rj += V[j - 1] @ t[j - 1]
        D.append(Sjj)
        r.append(-rj)
    E = [V[j] @ Pinv[j] @ U[j].T for j in range(N)]  # E[j] = S_{j+1, j}
    Ld, y = [None] * (N + 1), [None] * (N + 1)
    Lo = [None] * (N + 1)  # Lo[j] = S_{j, j-1} L_{j-1}^-T   (top chain, j = 1 .. m)
    Uo = [None] * (N + 1)  # Uo[j] = S_{j, j+1} M_{j+1}^-T   (bottom chain, j = m .. N-1)
    for j in range(m):  # top-down
        S = D[j].copy()
        rhs = r[j].copy()
        if j > 0:
            S -= Lo[j] @ Lo[j].T
            rhs -= Lo[j] @ y[j - 1]
        Ld[j] = np.linalg.cholesky(S)
        y[j] = np.linalg.solve(Ld[j], rhs)
        Lo[j + 1] = np.linalg.solve(Ld[j], E[j].T).T
    for j in range(N, m, -1):  # bottom-up
        S = D[j].copy()
        rhs = r[j].copy()
        if j < N:
            S -= Uo[j] @ Uo[j].T
            rhs -= Uo[j] @ y[j + 1]
        Ld[j] = np.linalg.cholesky(S)
        y[j] = np.linalg.solve(Ld[j], rhs)
        Uo[j - 1] = np.linalg.solve(Ld[j], E[j - 1]).T  # S_{j-1, j} M_j^-T = E[j-1]^T M_j^-T
    S = D[m].copy()
    rhs = r[m].copy()
    if m > 0:
        S -= Lo[m] @ Lo[m].T
        rhs -= Lo[m] @ y[m - 1]
    if m < N:
        S -= Uo[m] @ Uo[m].T
        rhs -= Uo[m] @ y[m + 1]
    Ld[m] = np.linalg.cholesky(S)
    y[m] = np.linalg.solve(Ld[m], rhs)
    nu = [None] * (N + 1)
    nu[m] = np.linalg.solve(Ld[m].T, y[m])
    for j in range(m - 1, -1, -1):
        nu[j] = np.linalg.solve(Ld[j].T, y[j] - Lo[j + 1].T @ nu[j + 1])
    for j in range(m + 1, N + 1):
        nu[j] = np.linalg.solve(Ld[j].T, y[j] - Uo[j - 1].T @ nu[j - 1])
    d = np.zeros(L["n_dec"])
    for j in range(N + 1):
        v = q[j] + U[j].T @ nu[j] + (V[j].T @ nu[j + 1] if j < N else 0.0)
        dw = -Pinv[j] @ v
        d[NX * j:NX * j + NX] = dw[:NX]
        if j < N:
            d[NX * (N + 1) + NU * j:NX * (N + 1) + NU * j + NU] = dw[NX:]
    return d, to_reference_rows(np.concatenate(nu), N)
