#!/usr/bin/env python
"""main/example3/run.py on the B200 path: L-shaped domain, u = r^(5/3) sin(5 theta / 3), inhomogeneous boundary data.

Mirrors ref main/example3/run.py:12-21 (tol 1e-11, intorder 8, P1, no penalty), :39-68 (driver: boundary values of
w_h by point evaluation, boundary values of u_h from a global L2 projection, condense with x != 0), :108-159.
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fem_forth_perturb_b200 as F
from fem_forth_perturb_b200 import (MeshTri, ElementTriMorley, ElementTriP1, ElementTriP2, InteriorBasis, asm_gpu as asm,
                                    condense, solve, solver_iter_mgcg, a_load, b_load, laplace, wv_load, LinearForm,
                                    morley_error_norms, project)

pi, sin, cos = np.pi, np.sin, np.cos


def arctan3(y, x):
    theta = np.arctan2(y, x)
    return np.where(theta <= 0, theta + 2 * pi, theta)          # ref main/example3/utils.py:668-671


def exact_u(x, y):
    return (x ** 2 + y ** 2) ** (5 / 6) * sin(5 * arctan3(y, x) / 3)


def dexact_u(x, y):
    th = arctan3(y, x)
    r2 = x ** 2 + y ** 2
    return ((5 * x * sin(5 * th / 3)) / (3 * r2 ** (1 / 6)) - (5 * y * cos(5 * th / 3) * r2 ** (5 / 6)) / (3 * r2),
            (5 * y * sin(5 * th / 3)) / (3 * r2 ** (1 / 6)) + (5 * x * cos(5 * th / 3) * r2 ** (5 / 6)) / (3 * r2))


def ddexact(x, y):
    th = arctan3(y, x)
    den = 9 * (x ** 2 + y ** 2) ** (7 / 6)
    duxx = -(10 * (y ** 2 * sin(5 * th / 3) - x ** 2 * sin(5 * th / 3) + 2 * x * y * cos(5 * th / 3))) / den
    duxy = (10 * (x ** 2 * cos(5 * th / 3) - y ** 2 * cos(5 * th / 3) + 2 * x * y * sin(5 * th / 3))) / den
    return duxx, duxy, duxy, -duxx


@LinearForm
def f_load(v, w):
    return 0.0 * w.x[0] * v                                       # ref main/example3/run.py:70-77 (lu = llu = 0)


def solve_problem1(m, element_type='P1', intorder=8, tol=1e-11, epsilon=1e-6, mode='parity'):
    """ref main/example3/run.py:39-68."""
    element = {'w': ElementTriP1() if element_type == 'P1' else ElementTriP2(), 'u': ElementTriMorley()}
    basis = {v: InteriorBasis(m, e, intorder=intorder) for v, e in element.items()}
    K1 = asm(laplace, basis['w'])
    f1 = asm(f_load, basis['w'])
    wh = np.zeros(basis['w'].N)
    boundary_dofs = basis['w'].find_dofs()['all'].all()
    loc = basis['w'].doflocs
    wh[boundary_dofs] = exact_u(loc[0][boundary_dofs], loc[1][boundary_dofs])
    wh = solve(*condense(K1, f1, wh, D=boundary_dofs), solver=solver_iter_mgcg(tol=tol, mode=mode))
    K2 = asm(epsilon ** 2 * a_load + b_load, basis['u'])
    f2 = asm(wv_load, basis['w'], basis['u']) * wh
    boundary_dofs = basis['u'].find_dofs()['all'].all()
    uh0 = project(exact_u, basis_to=basis['u'])
    uh0 = solve(*condense(K2, f2, uh0, D=boundary_dofs), solver=solver_iter_mgcg(tol=tol, mode=mode))
    return uh0, basis


def run(refine_time=6, epsilons=(1e-5, 1e-6), element_type='P1', intorder=8, tol=1e-11, mode='parity', out=print):
    results = {}
    t0 = time.time()
    for epsilon in epsilons:
        m = MeshTri.init_lshaped()
        rows = []
        for i in range(1, refine_time + 1):
            m.refine()
            uh0, basis = solve_problem1(m, element_type, intorder, tol, epsilon, mode)
            L2u, Du, D2u = morley_error_norms(basis['u'], uh0, exact_u, dexact_u, ddexact)
            rows.append((L2u, L2u + Du, L2u + Du + D2u, np.sqrt(epsilon ** 2 * D2u ** 2 + Du ** 2)))
        tab = np.array(rows)
        results[epsilon] = tab
        L2s, H1s, H2s, epus = tab.T
        out('epsilion:', epsilon)
        out('  h    L2u   H1u   H2u   epu')
        for i in range(tab.shape[0] - 1):
            out('2^-' + str(i + 2), ' {:.2f}  {:.2f}  {:.2f}  {:.2f}'.format(
                -np.log2(L2s[i + 1] / L2s[i]), -np.log2(H1s[i + 1] / H1s[i]), -np.log2(H2s[i + 1] / H2s[i]),
                -np.log2(epus[i + 1] / epus[i])))
            out('2^-' + str(i + 2), ' {:.3e}  {:.3e}  {:.3e}  {:.3e}'.format(L2s[i + 1], H1s[i + 1], H2s[i + 1], epus[i + 1]))
    out('Total Time Cost {:.2f} s'.format(time.time() - t0))
    return results


if __name__ == '__main__':
    run()
