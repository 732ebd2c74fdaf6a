import sys, os, ctypes
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import ungar_b200
from ungar_b200 import workloads as W
from oracle import qp_reference as Q
np.set_printoptions(linewidth=220, precision=3)
N = 2
model = ungar_b200.Model("quadruped", N, dtype="f64", barrier=(1.0, 1.0))
xp = W.synthetic_batch(W.QUADRUPED, N, 2, seed=41)
rec = model.kkt_blocks(torch.from_numpy(xp).cuda(), torch.zeros((2, model.layout["size"]), dtype=torch.float64, device="cuda"))
steps, mult = model.qp_solve(rec)
torch.cuda.synchronize()
WSG = 510
ws = np.zeros((N + 1) * WSG)
model._lib.ungar_b200_debug_qp_workspace.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
assert model._lib.ungar_b200_debug_qp_workspace(model._handle, ws.ctypes.data, ws.size) == 0
L = dict(model.layout)
H, q, U, V, b = Q.stage_blocks(rec[0].cpu().numpy(), L)
Pinv = [np.linalg.inv(h) for h in H]
t = [Pinv[j] @ q[j] for j in range(N + 1)]
bc = lambda c: 28 * c - (c * (c - 1)) // 2 + c // 2
def factor(j):
    g = ws[j * WSG:(j + 1) * WSG]
    Lm = np.zeros((29, 29))
    for c in range(29):
        for r in range(c, 29):
            Lm[r, c] = g[bc(c) + r] * g[450 + c]
    return Lm, g[480:509].copy()
# group 0 (top): S_00 = U0 P0^-1 U0^T + delta
S00 = U[0] @ Pinv[0] @ U[0].T + 1e-9 * np.eye(29)
r0 = -(b[0] + U[0] @ t[0])
L0 = np.linalg.cholesky(S00)
y0 = np.linalg.solve(L0, r0)
Lg, yg = factor(0)
print("group 0: |L - Lref| %.2e  |y - yref| %.2e   (|L| %.2e |y| %.2e)" % (np.max(np.abs(Lg - L0)), np.max(np.abs(yg - y0)), np.max(np.abs(L0)), np.max(np.abs(y0))))
print("col errs", np.max(np.abs(Lg - L0), axis=0))
# group N (bottom first): S_NN = U_N P_N^-1 U_N^T + V_{N-1} P^-1 V^T + delta (13 rows real)
SNN = U[N] @ Pinv[N] @ U[N].T + V[N - 1][:13] @ Pinv[N - 1] @ V[N - 1][:13].T + 1e-9 * np.eye(13)
rN = -(b[N] + U[N] @ t[N] + V[N - 1][:13] @ t[N - 1])
LN = np.linalg.cholesky(SNN)
yN = np.linalg.solve(LN, rN)
Lg, yg = factor(N)
print("group N: |L - Lref| %.2e  |y - yref| %.2e   (|L| %.2e |y| %.2e)" % (np.max(np.abs(Lg[:13, :13] - LN)), np.max(np.abs(yg[:13] - yN)), np.max(np.abs(LN)), np.max(np.abs(yN))))
print("col errs", np.max(np.abs(Lg[:13, :13] - LN), axis=0))
