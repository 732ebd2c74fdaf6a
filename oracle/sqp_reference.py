"""ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement of the outer loop that consumes the derivative path:

  * ``line_search``  — BacktrackingLineSearch::Do (include/ungar/optimization/backtracking_line_search.hpp:81-165) with the
                       two merit lambdas SoftSQPOptimizer::Optimize hands it (include/ungar/optimization/soft_sqp.hpp:85-99):
                       phi(x) = f(x) + Zsoft(h(x)),  theta(x) = multiplier * sqrt(|g(x)|^2);
  * ``monolithic_qp``— the QP data exactly as AssembleOSQPInstance builds it from the three Functions (soft_sqp.hpp:141-158,
                       :236-264), any model, solved by one sparse LU of the KKT system (OSQP v0.6.3 is absent from the tree,
                       external/config/osqp/CMakeLists.txt.in:16; any exact QP solver returns this minimiser);
  * ``soft_sqp``     — SoftSQPOptimizer::Optimize (soft_sqp.hpp:63-109): objective, local QP, line search, convergence test.

Function values and derivatives come from the fp64 oracle (oracle.cpp / models.hpp).  Pure numpy / scipy, one trajectory
per call: sized for tests, never on a product path.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

OBJECTIVE, EQUALITIES, INEQUALITIES = 0, 1, 2
RUNNING, CONVERGED, LINE_SEARCH_FAILED = 0, 1, 2  # same codes as ungar_b200_sqp_status


@dataclass
class LineSearchParameters:
    """BacktrackingLineSearch::Parameters defaults (backtracking_line_search.hpp:70-76)."""
    alphaMin: float = 1e-4
    thetaMin: float = 1e-6
    thetaMax: float = 1e-2
    eta: float = 1e-4
    gammaPhi: float = 1e-6
    gammaTheta: float = 1e-6
    gammaAlpha: float = 0.5


@dataclass
class LineSearchResult:
    accepted: bool
    alpha: float
    trials: int
    theta0: float
    phi0: float
    theta: float
    phi: float
    projection: float
    history: list = field(default_factory=list)


class Merit:
    """The cost-function and constraint-violation lambdas of soft_sqp.hpp:85-99 for one trajectory."""

    def __init__(self, oracle, model: int, N: int, xp: np.ndarray, stiffness: float, epsilon: float, multiplier: float):
        self.o, self.model, self.N = oracle, model, N
        self.xp = np.array(xp, dtype=np.float64)
        self.k, self.eps, self.mult = stiffness, epsilon, multiplier

    def _full(self, x):
        self.xp[:x.size] = x
        return self.xp

    def objective(self, x) -> float:
        return float(self.o.evaluate(self.model, OBJECTIVE, self.N, self._full(x))[0])

    def phi(self, x) -> float:
        h = self.o.evaluate(self.model, INEQUALITIES, self.N, self._full(x))
        return self.objective(x) + self.o.barrier(self.k, self.eps, h)[0]

    def theta(self, x) -> float:
        g = self.o.evaluate(self.model, EQUALITIES, self.N, self._full(x))
        return self.mult * float(np.sqrt(np.dot(g, g)))


def line_search(gradient: np.ndarray, dw: np.ndarray, phi_fn, theta_fn, w: np.ndarray,
                p: LineSearchParameters = LineSearchParameters()):
    """BacktrackingLineSearch::Do.  Returns (LineSearchResult, w_next) — w_next is w itself when no step is accepted."""
    projection = float(np.sum(gradient * dw))
    alpha = 1.0
    theta, phi = theta_fn(w), phi_fn(w)
    accepted, trials, hist = False, 0, []
    theta_next, phi_next = theta, phi
    while not accepted and alpha >= p.alphaMin:
        w_next = w + alpha * dw
        theta_next, phi_next = theta_fn(w_next), phi_fn(w_next)
        trials += 1
        hist.append((alpha, theta_next, phi_next))
        if theta_next > p.thetaMax:
            if theta_next < (1.0 - p.gammaTheta) * theta:
                accepted = True
        elif max(theta, theta_next) < p.thetaMin and projection < 0.0:
            if phi_next < phi + p.eta * alpha * projection:
                accepted = True
        else:
            if phi_next < (1.0 - p.gammaPhi) * phi or theta_next < (1.0 - p.gammaTheta) * theta:
                accepted = True
        if not accepted:
            alpha *= p.gammaAlpha
    res = LineSearchResult(accepted, alpha if accepted else 0.0, trials, theta, phi, theta_next, phi_next, projection, hist)
    return res, (w + alpha * dw if accepted else w.copy())


def _csr(triplets, shape):
    rows, cols, vals = triplets
    return sp.csr_matrix((vals, (rows, cols)), shape=shape)


def monolithic_qp(oracle, model: int, N: int, xp: np.ndarray, stiffness: float, epsilon: float):
    """(P, q, A, g) of AssembleOSQPInstance: P = triu(H_f) + J_h^T diag(b'') J_h + 1e-6 I (full symmetric here),
    q = grad f + J_h^T b', A = J_g, l = u = -g."""
    s = oracle.sizes(model, N)
    n = s["n_dec"]
    Hu = _csr(oracle.hessian(model, N, xp), (n, n))
    Hf = Hu + sp.triu(Hu, 1).T
    Jf = _csr(oracle.jacobian(model, OBJECTIVE, N, xp), (1, n))
    Jg = _csr(oracle.jacobian(model, EQUALITIES, N, xp), (s["m_eq"], n))
    Jh = _csr(oracle.jacobian(model, INEQUALITIES, N, xp), (s["m_ineq"], n))
    h = oracle.evaluate(model, INEQUALITIES, N, xp)
    g = oracle.evaluate(model, EQUALITIES, N, xp)
    _, dz, d2z = oracle.barrier(stiffness, epsilon, h)
    P = Hf + Jh.T @ sp.diags(d2z) @ Jh + 1e-6 * sp.identity(n)
    q = np.asarray(Jf.todense()).ravel() + Jh.T @ dz
    return P.tocsc(), q, Jg.tocsc(), g, np.asarray(Jf.todense()).ravel()


def solve_qp(P, q, A, g, delta: float = 1e-9):
    """argmin 1/2 d^T P d + q^T d  s.t.  A d = -g, via the quasi-definite KKT system (same delta as the CUDA kernel)."""
    n, m = P.shape[0], A.shape[0]
    if m == 0:
        return spla.spsolve(sp.csc_matrix(P), -q), np.zeros(0)
    K = sp.bmat([[P, A.T], [A, -delta * sp.identity(m)]], format="csc")
    sol = spla.spsolve(K, np.concatenate([-q, -g]))
    return sol[:n], sol[n:]


class OracleProblem:
    """One trajectory of an example NLP problem (objective / equalities / inequalities Functions + relaxed barrier), evaluated by
    the fp64 oracle.  ``x`` is the decision part; the parameters stay fixed."""

    def __init__(self, oracle, model: int, N: int, xp: np.ndarray, stiffness: float, epsilon: float):
        self.o, self.model, self.N = oracle, model, N
        self.xp = np.array(xp, dtype=np.float64)
        self.k, self.eps = stiffness, epsilon
        self.n = oracle.sizes(model, N)["n_dec"]

    def _full(self, x):
        self.xp[:self.n] = x
        return self.xp

    def objective(self, x) -> float:
        return float(self.o.evaluate(self.model, OBJECTIVE, self.N, self._full(x))[0])

    def equalities(self, x):
        return self.o.evaluate(self.model, EQUALITIES, self.N, self._full(x))

    def inequalities(self, x):
        return self.o.evaluate(self.model, INEQUALITIES, self.N, self._full(x))

    def barrier(self, h):
        return self.o.barrier(self.k, self.eps, h)

    def qp_data(self, x):
        """(P, q, A, g, grad f) of AssembleOSQPInstance at x."""
        return monolithic_qp(self.o, self.model, self.N, self._full(x).copy(), self.k, self.eps)


class DenseProblem:
    """An NLP problem given by numpy callables (tests of the loop itself against test/optimization/soft_sqp.test.cpp):
    f, grad f, hess f; g, J_g (or None: `hana::nothing`); h, J_h; the barrier comes from the oracle's restatement."""

    def __init__(self, oracle, n, f, grad, hess, g=None, Jg=None, h=None, Jh=None, stiffness=100.0, epsilon=2e-5):
        self.o, self.n, self.k, self.eps = oracle, n, stiffness, epsilon
        self.f, self.grad, self.hess, self.g, self.Jg, self.h, self.Jh = f, grad, hess, g, Jg, h, Jh

    def objective(self, x) -> float:
        return float(self.f(x))

    def equalities(self, x):
        return np.atleast_1d(self.g(x)).astype(float) if self.g else np.zeros(0)

    def inequalities(self, x):
        return np.atleast_1d(self.h(x)).astype(float) if self.h else np.zeros(0)

    def barrier(self, h):
        return self.o.barrier(self.k, self.eps, h) if h.size else (0.0, np.zeros(0), np.zeros(0))

    def qp_data(self, x):
        n = self.n
        Jh = np.atleast_2d(self.Jh(x)) if self.h else np.zeros((0, n))
        Jg = np.atleast_2d(self.Jg(x)) if self.g else np.zeros((0, n))
        _, dz, d2z = self.barrier(self.inequalities(x))
        gf = np.asarray(self.grad(x), dtype=float)
        P = np.triu(self.hess(x))
        P = P + np.triu(P, 1).T + Jh.T @ np.diag(d2z) @ Jh + 1e-6 * np.eye(n)
        return sp.csc_matrix(P), gf + Jh.T @ dz, sp.csc_matrix(Jg), self.equalities(x), gf


def soft_sqp_loop(problem, x0: np.ndarray, multiplier: float = 1.0, max_iterations: int = 10,
                  params: LineSearchParameters = LineSearchParameters(), qp=None):
    """SoftSQPOptimizer::Optimize (soft_sqp.hpp:63-109) on a problem object.  Returns (x, status, iterations, log)."""
    x = np.array(x0, dtype=np.float64)

    def phi(w):  # soft_sqp.hpp:85-89
        return problem.objective(w) + problem.barrier(problem.inequalities(w))[0]

    def theta(w):  # soft_sqp.hpp:90-98
        g = problem.equalities(w)
        return multiplier * float(np.sqrt(np.dot(g, g)))

    status, iterations, log = RUNNING, 0, []
    for _ in range(max_iterations):
        objective = problem.objective(x)
        P, q, A, g, grad_f = problem.qp_data(x)
        d = qp(x) if qp is not None else solve_qp(P, q, A, g)[0]
        res, w_next = line_search(grad_f, d, phi, theta, x.copy(), params)
        iterations += 1
        log.append(dict(objective=objective, step=d, ls=res))
        if not res.accepted:
            status = LINE_SEARCH_FAILED
            break
        x = w_next
        diff = problem.objective(x) - objective
        if diff < 0.0 and abs(diff) < 1e-6:
            status = CONVERGED
            break
    return x, status, iterations, log


def soft_sqp(oracle, model: int, N: int, xp: np.ndarray, stiffness: float, epsilon: float, multiplier: float = 1.0,
             max_iterations: int = 10, params: LineSearchParameters = LineSearchParameters(), qp=None):
    """SoftSQPOptimizer::Optimize for one trajectory of an example problem.  ``qp(xp) -> d`` overrides the QP solve (tests pass
    the GPU step in to isolate the line search).  Returns (xp_final, status, iterations, log)."""
    prob = OracleProblem(oracle, model, N, xp, stiffness, epsilon)
    qp_x = (lambda x: qp(prob._full(x).copy())) if qp is not None else None
    x, status, iterations, log = soft_sqp_loop(prob, np.array(xp[:prob.n], dtype=np.float64), multiplier, max_iterations, params, qp_x)
    out = np.array(xp, dtype=np.float64)
    out[:prob.n] = x
    return out, status, iterations, log
