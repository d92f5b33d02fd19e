"""Oracle (test infrastructure): restatement of scipy 1.4.1 ``sparse.linalg.cg`` as patched by
the reference's author to also return the iteration count (ref README.md:20-22).

Call sites: ref main/example1/utils.py:115,157,252,305 (``sol, info, iter = krylov(A, b, ...)``).
Semantics restated per SURVEY.md B.5: x0 = 0, ``maxiter = 10 n`` by default, legacy ``atol``
(absolute early exit ``||b - A x0|| <= tol``, then threshold ``tol * ||b||``), preconditioned CG
with a true-residual re-check on convergence when it > 1.
"""
import numpy as np


def cg141(A, b, M=None, tol=1e-5, maxiter=None, callback=None):
    """Returns (x, info, iters) like the patched scipy cg; ``M`` is a callable r -> z or None."""
    n = len(b)
    if maxiter is None:
        maxiter = 10 * n
    x = np.zeros(n)
    bnrm = np.linalg.norm(b)
    r = b - A @ x
    if np.linalg.norm(r) <= tol:
        return x, 0, 0
    atol = tol * bnrm if bnrm != 0 else tol
    rho_prev = 1.0
    p = None
    for it in range(1, maxiter + 1):
        z = M(r) if M is not None else r.copy()
        rho = r @ z
        p = z.copy() if it == 1 else z + (rho / rho_prev) * p
        q = A @ p
        alpha = rho / (p @ q)
        x += alpha * p
        r -= alpha * q
        rho_prev = rho
        if callback is not None:
            callback(x)
        if np.linalg.norm(r) <= atol:
            if it == 1:
                return x, 0, it
            r = b - A @ x
            if np.linalg.norm(r) <= atol:
                return x, 0, it
    return x, maxiter, maxiter


def build_pc_diag(A):
    """``build_pc_diag`` (ref main/example1/utils.py:48-50) as a callable."""
    dinv = 1.0 / A.diagonal()
    return lambda r: dinv * r


def solver_iter_mgcg(tol=1e-8, maxiter=None, record=None):
    """``solver_iter_mgcg`` (ref utils.py:127-166): RS-AMG V-cycle preconditioned CG."""
    from .pyamg_rs import RugeStuben

    def solver(A, b):
        ml = RugeStuben(A)
        x, info, it = cg141(A, b, M=ml.precondition, tol=tol, maxiter=maxiter)
        if record is not None:
            record.append(it)
        return x
    return solver


def solver_iter_pcg(tol=1e-8, maxiter=None, record=None):
    """``solver_iter_krylov(Precondition=True)`` (ref utils.py:265-320): Jacobi-PCG."""
    def solver(A, b):
        x, info, it = cg141(A, b, M=build_pc_diag(A), tol=tol, maxiter=maxiter)
        if record is not None:
            record.append(it)
        return x
    return solver
