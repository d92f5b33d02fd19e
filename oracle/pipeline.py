"""Oracle (test infrastructure): the reference's problem drivers and error norms on the CPU.

Restates ``solve_problem1`` / ``solve_problem2`` (ref main/example1/run.py:182-272), the
example-3 driver with inhomogeneous data (ref main/example3/run.py:39-77), the problem data
(ref main/example1/run.py:275-493, ref main/example3/utils.py:668-700) and the error norms
(ref main/example1/run.py:106-108,157-179,519-543).  ``solver``s are the closures of
``oracle.scipy_cg`` or a direct solve; hooks let the tests swap in the GPU product for the
K2 assembly / solve while everything around it stays the oracle.
"""
import numpy as np
import scipy.sparse.linalg as spl

from .skfem_basis import InteriorBasis, FacetBasis
from . import skfem_asm as oa

pi, sin, cos = np.pi, np.sin, np.cos


# ------------------------------------------------------------------ problem data
class Ex1:
    """ref main/example1/run.py:275-309 (u = sin^2(pi x) sin^2(pi y))."""
    name = 'ex1'

    def __init__(self, eps):
        self.eps = eps

    def f(self, x, y):
        pix, piy = pi * x, pi * y
        lu = 2 * pi ** 2 * (cos(2 * pix) * sin(piy) ** 2 + cos(2 * piy) * sin(pix) ** 2)
        llu = -8 * pi ** 4 * (cos(2 * pix) * sin(piy) ** 2 + cos(2 * piy) * sin(pix) ** 2
                              - cos(2 * pix) * cos(2 * piy))
        return self.eps ** 2 * llu - lu

    def u(self, x, y):
        return (sin(pi * x) * sin(pi * y)) ** 2

    def du(self, x, y):
        return (2 * pi * cos(pi * x) * sin(pi * x) * sin(pi * y) ** 2,
                2 * pi * cos(pi * y) * sin(pi * x) ** 2 * sin(pi * y))

    def ddu(self, x, y):
        uxx = 2 * pi ** 2 * cos(pi * x) ** 2 * sin(pi * y) ** 2 - 2 * pi ** 2 * sin(pi * x) ** 2 * sin(pi * y) ** 2
        uxy = 2 * pi * cos(pi * x) * sin(pi * x) * 2 * pi * cos(pi * y) * sin(pi * y)
        uyy = 2 * pi ** 2 * cos(pi * y) ** 2 * sin(pi * x) ** 2 - 2 * pi ** 2 * sin(pi * y) ** 2 * sin(pi * x) ** 2
        return uxx, uxy, uxy, uyy


class Ex3:
    """ref main/example1/run.py:466-493 (u0 = sin(pi x) sin(pi y), f = 2 pi^2 u0)."""
    name = 'ex3'

    def __init__(self, eps):
        self.eps = eps

    def f(self, x, y):
        return 2 * pi ** 2 * sin(pi * x) * sin(pi * y)

    def u(self, x, y):
        return sin(pi * x) * sin(pi * y)

    def du(self, x, y):
        return pi * cos(pi * x) * sin(pi * y), pi * cos(pi * y) * sin(pi * x)

    def ddu(self, x, y):
        uxx = -pi ** 2 * sin(pi * x) * sin(pi * y)
        uxy = pi * cos(pi * x) * pi * cos(pi * y)
        return uxx, uxy, uxy, uxx


def _arctan3(y, x):
    th = np.arctan2(y, x)
    return np.where(th <= 0, th + 2 * pi, th)


class Ex4:
    """ref main/example3/utils.py:668-700 (L-shape, u = r^{5/3} sin(5 theta / 3), f = 0)."""
    name = 'ex4'

    def __init__(self, eps):
        self.eps = eps

    def f(self, x, y):
        return 0.0 * x

    def u(self, x, y):
        return (x ** 2 + y ** 2) ** (5 / 6) * sin(5 * _arctan3(y, x) / 3)

    def du(self, x, y):
        th = _arctan3(y, x)
        r2 = x ** 2 + y ** 2
        return ((5 * x * sin(5 * th / 3)) / (3 * r2 ** (1 / 6)) - (5 * y * cos(5 * th / 3) * r2 ** (5 / 6)) / (3 * r2),
                (5 * y * sin(5 * th / 3)) / (3 * r2 ** (1 / 6)) + (5 * x * cos(5 * th / 3) * r2 ** (5 / 6)) / (3 * r2))

    def ddu(self, x, y):
        th = _arctan3(y, x)
        den = 9 * (x ** 2 + y ** 2) ** (7 / 6)
        uxx = -(10 * (y ** 2 * sin(5 * th / 3) - x ** 2 * sin(5 * th / 3) + 2 * x * y * cos(5 * th / 3))) / den
        uxy = (10 * (x ** 2 * cos(5 * th / 3) - y ** 2 * cos(5 * th / 3) + 2 * x * y * sin(5 * th / 3))) / den
        return uxx, uxy, uxy, -uxx


# ------------------------------------------------------------------ solvers
def direct_solver(A, b):
    return spl.spsolve(A.tocsc(), b)


# ------------------------------------------------------------------ drivers
def w_problem(mesh, prob, wkind, intorder, solver, inhomogeneous=False):
    """First half of solve_problem1/2: the Poisson problem for w_h (ref run.py:198-206)."""
    wb = InteriorBasis(mesh, wkind, intorder)
    K1 = oa.asm_bilinear(oa.laplace, wb)
    f1 = oa.asm_linear(prob.f, wb)
    D = wb.find_dofs()
    wh0 = np.zeros(wb.N)
    if inhomogeneous:                                   # ref main/example3/run.py:57-60
        loc = wb.doflocs()
        wh0[D] = prob.u(loc[0][D], loc[1][D])
    wh = oa.solve(*oa.condense(K1, f1, wh0, D=D), solver=solver)
    return wb, wh


def assemble_K2(ub, eps, fbasis=None, sigma=5.0):
    """``K2 = eps^2 asm(a_load) [+ eps^2 P] + asm(b_load)`` (ref run.py:210 / 253-260)."""
    A = oa.asm_bilinear(oa.a_load, ub)
    B = oa.asm_bilinear(oa.b_load, ub)
    if fbasis is None:
        return eps ** 2 * A + B
    p1, p2, p3 = (oa.asm_bilinear(f, fbasis) for f in oa.make_penalty_forms(fbasis, sigma))
    P = p1 + p2 + p3
    return eps ** 2 * A + eps ** 2 * P + B


def solve_problem(mesh, prob, wkind='P1', intorder=5, penalty=False, sigma=5.0,
                  solver_w=direct_solver, solver_u=direct_solver,
                  K2_hook=None, inhomogeneous=False, return_system=False):
    """solve_problem1 (penalty=False) / solve_problem2 (penalty=True) / example-3 driver."""
    wb, wh = w_problem(mesh, prob, wkind, intorder, solver_w, inhomogeneous)
    ub = InteriorBasis(mesh, 'Morley', intorder)
    fb = FacetBasis(mesh, 'Morley') if penalty else None
    if K2_hook is not None:
        K2 = K2_hook(mesh, prob.eps, penalty, sigma)
    else:
        K2 = assemble_K2(ub, prob.eps, fb, sigma)
    f2 = oa.asm_bilinear(oa.wv_load, wb, ub) @ wh
    if inhomogeneous:                                   # ref main/example3/run.py:63-67
        D = ub.find_dofs()
        M = oa.asm_bilinear(oa.mass, ub)
        uh0 = spl.spsolve(M.tocsc(), oa.asm_linear(prob.u, ub))      # project(exact_u, basis_to=...)
    else:
        D = oa.easy_boundary_penalty(ub) if penalty else oa.easy_boundary(ub)
        uh0 = None
    Kc, bc, x, I = oa.condense(K2, f2, uh0, D=D)
    if return_system:
        return dict(K2=K2, f2=f2, Kc=Kc, bc=bc, I=I, D=D, x=x, ub=ub, fb=fb, wb=wb, wh=wh)
    uh = oa.solve(Kc, bc, x, I, solver=solver_u)
    return uh, ub, fb


def error_norms(uh, ub, prob, fb=None):
    """(L2, H1, H2, Energy) exactly as tabulated by ref main/example1/run.py:519-548."""
    eps = prob.eps
    x, dx = ub.x, ub.dx
    val, g, h = ub.interpolate(uh)
    L2 = np.sqrt(np.sum((val - prob.u(x[0], x[1])) ** 2 * dx))
    ux, uy = prob.du(x[0], x[1])
    D1 = np.sqrt(np.sum(((g[0] - ux) ** 2 + (g[1] - uy) ** 2) * dx))
    uxx, uxy, uyx, uyy = prob.ddu(x[0], x[1])
    D2 = np.sqrt(np.sum(((h[0, 0] - uxx) ** 2 + (h[0, 1] - uxy) ** 2
                         + (h[1, 1] - uyy) ** 2 + (h[1, 0] - uyx) ** 2) * dx))
    if fb is not None:                                  # L2pnvError, ref run.py:106-108,527
        _, gf, _ = fb.interpolate(uh)
        pn = np.sum((fb.h * np.einsum('i...,i...', fb.n, gf)) ** 2 * fb.dx)
        D2 = np.sqrt(D2 ** 2 + pn)
    energy = np.sqrt(eps ** 2 * D2 ** 2 + D1 ** 2)
    return L2, L2 + D1, L2 + D1 + D2, energy
