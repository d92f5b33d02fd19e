"""Oracle (test infrastructure): main/example2 on the CPU - errors of the level-i solutions against the level-N one.

Restates ref main/example2/run.py:78-179.  The reference has no exact solution for this problem (eps = 1e-6,
f = 2 pi^2 sin(pi x) sin(pi y)); it loads the level-1..7 solutions from ``solutions/*.npy`` (absent from the
repository; ref main/example2/utils.py:835-841) and measures u^N - u_h, N = base_order = 7, at the quadrature
points of the level-N mesh:

  * L2   : ``sqrt(sum(dx_N * (u^N - u_h(x_N))^2))``                                         run.py:112
  * D1   : u_h's gradient is first L2-projected onto DG-P1 of the coarse mesh (``project(..., diff=0/1)``,
           run.py:114-116) and that projection is evaluated at x_N (run.py:121-137).  A Morley function is
           quadratic on every element, its derivatives are linear, mass and derivative forms are integrated
           exactly by the intorder-3 rule, so the projection IS the elementwise derivative;
           ``dg_projection_check`` below does the projection literally to confirm it.
  * D2   : the same one level further, DG-P0 projections of the derivatives of the DG-P1 fields (run.py:142-172),
           exact for the same reason; with penalty the boundary term of the COARSE solution is added:
           ``sqrt(D2^2 + L2pnvError.assemble(fbasis, w=fbasis.interpolate(uh0)))``            run.py:174-177
  * table: H1 = L2 + D1, H2 = H1 + D2, energy = sqrt(eps^2 D2^2 + D1^2)                       run.py:179-191

Point location: the reference calls ``basis.interpolator(u)(x)`` for one point at a time (hours of Python: the
reference's own log reports 43 285 s).  Uniform refinement appends the children of triangle k at k, k + nt,
k + 2 nt, k + 3 nt (oracle.skfem_mesh.MeshTri.refine, as skfem 2.1.1), so the level-i ancestor of level-N triangle
T is T mod nt_i; ``ancestors`` checks that every point really lies in that triangle.
"""
import numpy as np

from .skfem_mesh import MeshTri
from .skfem_basis import InteriorBasis, morley_tabulate, affine, quadrature_tri
from . import pipeline as pl


def solve_level(level, wkind='P1', penalty=True, intorder=3, eps=1e-6, solver=pl.direct_solver):
    """``solve_problem2(m, element_type, 'mgcg', intorder=3, tol=1e-8, epsilon=1e-6)`` (penalty) or
    ``solve_problem1`` (no penalty) on the level-``level`` mesh (ref main/example2/run.py:88-91, utils.py:625-721)."""
    m = MeshTri().refine(level)
    uh, ub, fb = pl.solve_problem(m, pl.Ex3(eps), wkind=wkind, intorder=intorder, penalty=penalty,
                                  solver_w=solver, solver_u=solver)
    return m, uh, ub, fb


def ancestors(fine, coarse, x, tol=1e-12):
    """Coarse triangle holding each fine triangle (and its quadrature points x (2, nt_f, nq))."""
    anc = np.arange(fine.nt) % coarse.nt
    A, b, _, invA = affine(coarse)
    d = x - b[:, anc][:, :, None]
    loc = np.einsum('ijk,jkl->ikl', invA[:, :, anc], d)
    if not (loc.min() >= -tol and (loc[0] + loc[1]).max() <= 1 + tol):
        raise ValueError('meshes are not nested by uniform refinement')
    return anc


def evaluate_on(fine_basis, coarse_mesh, coarse_basis, uh):
    """u_h, its elementwise gradient and Hessian at the fine quadrature points (each (.., nt_f, nq))."""
    x = fine_basis.x
    anc = ancestors(fine_basis.mesh, coarse_mesh, x)
    phi = morley_tabulate(coarse_mesh, x, tind=anc)
    ed = coarse_basis.element_dofs[:, anc]
    val = sum(uh[ed[i]][:, None] * phi[i][0] for i in range(6))
    grad = sum(uh[ed[i]][None, :, None] * phi[i][1] for i in range(6))
    hess = sum(uh[ed[i]][None, None, :, None] * phi[i][2] for i in range(6))
    return val, grad, hess


def nested_errors(base, coarse, eps=1e-6, penalty=True):
    """(L2, D1, D2) of one 'Testing order' pass; base/coarse = (mesh, uh, ubasis, fbasis) from ``solve_level``."""
    _, uN, ubN, _ = base
    mc, uh, ubc, fbc = coarse
    dx = ubN.dx
    vN, gN, hN = ubN.interpolate(uN)
    vc, gc, hc = evaluate_on(ubN, mc, ubc, uh)
    L2 = np.sqrt(np.sum(dx * (vN - vc) ** 2))
    D1 = np.sqrt(np.sum(((gN[0] - gc[0]) ** 2 + (gN[1] - gc[1]) ** 2) * dx))
    D2 = np.sqrt(np.sum(((hN[0, 0] - hc[0, 0]) ** 2 + (hN[0, 1] - hc[0, 1]) ** 2 + (hN[1, 1] - hc[1, 1]) ** 2
                         + (hN[1, 0] - hc[1, 0]) ** 2) * dx))
    if penalty:
        _, gf, _ = fbc.interpolate(uh)
        pn = np.sum((fbc.h * np.einsum('i...,i...', fbc.n, gf)) ** 2 * fbc.dx)
        D2 = np.sqrt(D2 ** 2 + pn)
    return L2, D1, D2


def table(base_order=7, refine_time=6, wkind='P1', penalty=True, intorder=3, eps=1e-6, solver=pl.direct_solver,
          solutions=None):
    """Rows (L2, H1, H2, Energy) for levels 1..refine_time, as saved by ref main/example2/run.py:186-191.
    ``solutions`` = {level: uh} replaces the solves (the reference's ``solutions/*.npy``)."""
    def level(i):
        if solutions is None:
            return solve_level(i, wkind, penalty, intorder, eps, solver)
        m = MeshTri().refine(i)
        from .skfem_basis import FacetBasis
        return m, solutions[i], InteriorBasis(m, 'Morley', intorder), (FacetBasis(m, 'Morley') if penalty else None)
    base = level(base_order)
    rows = []
    for i in range(1, refine_time + 1):
        L2, D1, D2 = nested_errors(base, level(i), eps, penalty)
        rows.append((L2, L2 + D1, L2 + D1 + D2, np.sqrt(eps ** 2 * D2 ** 2 + D1 ** 2)))
    return np.array(rows)


def dg_projection_check(level=2, intorder=3, wkind='P1', penalty=True):
    """The literal ``project(uh, basis_from=Morley, basis_to=DG-P1, diff=d)`` of run.py:114-116, element by element
    (DG mass matrices are block diagonal), against the elementwise derivative; returns the largest difference."""
    m, uh, ub, _ = solve_level(level, wkind, penalty, intorder)
    X, W = quadrature_tri(intorder)
    P1 = np.array([1 - X[0] - X[1], X[0], X[1]])                      # (3, nq) on the reference triangle
    _, g, _ = ub.interpolate(uh)                                      # (2, nt, nq)
    M = np.einsum('iq,jq,q->ij', P1, P1, W)                           # |det| cancels between M and the load
    worst = 0.0
    for d in range(2):
        rhs = np.einsum('iq,tq,q->ti', P1, g[d], W)
        coef = np.linalg.solve(M, rhs.T).T                            # DG-P1 coefficients per element
        worst = max(worst, np.abs(coef @ P1 - g[d]).max())
    return worst
