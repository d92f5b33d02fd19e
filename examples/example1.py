#!/usr/bin/env python
"""main/example1/run.py on the B200 path - same parameters, same drivers, same printed tables.

Mirrors ref main/example1/run.py:18-31 (parameters), :182-222 (solve_problem1), :497-569 (loops, norms, table).
Everything numerical runs on the device (mesh, K1/K2 assembly, condense, AMG-PCG, w-to-u coupling, error
integrals); the problem data (f, exact solution and its derivatives) stay the reference's Python functions and
are sampled at the quadrature points on the host.

    python examples/example1.py [--refine 6] [--eps-range 6] [--element P2] [--mode parity|chebyshev]
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fem_forth_perturb_b200 as F
from fem_forth_perturb_b200 import (MeshTri, ElementTriMorley, ElementTriP1, ElementTriP2, InteriorBasis, asm_gpu as asm,
                                    condense, solve, solver_iter_mgcg, solver_iter_mgcg_iter, solver_iter_krylov,
                                    a_load, b_load, laplace, wv_load, LinearForm, morley_error_norms)

pi, sin, cos = np.pi, np.sin, np.cos


def make_problem_ex3(epsilon):
    """'ex3' of ref main/example1/run.py:466-493 (used with penalty=True)."""
    @LinearForm
    def f_load(v, w):
        return (2 * pi ** 2 * sin(pi * w.x[0]) * sin(pi * w.x[1])) * v

    def exact_u(x, y):
        return sin(pi * x) * sin(pi * y)

    def dexact_u(x, y):
        return pi * cos(pi * x) * sin(pi * y), pi * cos(pi * y) * sin(pi * x)

    def ddexact(x, y):
        duxx = -pi ** 2 * sin(pi * x) * sin(pi * y)
        duxy = pi * cos(pi * x) * pi * cos(pi * y)
        return duxx, duxy, duxy, duxx
    return f_load, exact_u, dexact_u, ddexact


def make_problem(epsilon):
    """'ex1' of ref main/example1/run.py:275-309."""
    @LinearForm
    def f_load(v, w):
        pix, piy = pi * w.x[0], pi * w.x[1]
        lu = 2 * pi ** 2 * (cos(2 * pix) * sin(piy) ** 2 + cos(2 * piy) * sin(pix) ** 2)
        llu = -8 * pi ** 4 * (cos(2 * pix) * sin(piy) ** 2 + cos(2 * piy) * sin(pix) ** 2 - cos(2 * pix) * cos(2 * piy))
        return (epsilon ** 2 * llu - lu) * v

    def exact_u(x, y):
        return (sin(pi * x) * sin(pi * y)) ** 2

    def dexact_u(x, y):
        return (2 * pi * cos(pi * x) * sin(pi * x) * sin(pi * y) ** 2, 2 * pi * cos(pi * y) * sin(pi * x) ** 2 * sin(pi * y))

    def ddexact(x, y):
        duxx = 2 * pi ** 2 * cos(pi * x) ** 2 * sin(pi * y) ** 2 - 2 * pi ** 2 * sin(pi * x) ** 2 * sin(pi * y) ** 2
        duxy = 2 * pi * cos(pi * x) * sin(pi * x) * 2 * pi * cos(pi * y) * sin(pi * y)
        duyy = 2 * pi ** 2 * cos(pi * y) ** 2 * sin(pi * x) ** 2 - 2 * pi ** 2 * sin(pi * y) ** 2 * sin(pi * x) ** 2
        return duxx, duxy, duxy, duyy
    return f_load, exact_u, dexact_u, ddexact


def easy_boundary(m, basis):
    """ref main/example1/run.py:86-104."""
    dofs = basis.find_dofs({
        'left': m.facets_satisfying(lambda x: x[0] == 0, boundaries_only=True),
        'right': m.facets_satisfying(lambda x: x[0] == 1, boundaries_only=True),
        'top': m.facets_satisfying(lambda x: x[1] == 1, boundaries_only=True),
        'buttom': m.facets_satisfying(lambda x: x[1] == 0, boundaries_only=True)})
    return np.concatenate([dofs[k].nodal['u'] for k in ('left', 'right', 'top', 'buttom')]
                          + [dofs[k].facet['u_n'] for k in ('left', 'right', 'top', 'buttom')])


def easy_boundary_penalty(m, basis):
    """ref main/example1/run.py:67-83: nodal DOFs only - the normal derivative is imposed weakly."""
    dofs = basis.find_dofs({
        'left': m.facets_satisfying(lambda x: x[0] == 0, boundaries_only=True),
        'right': m.facets_satisfying(lambda x: x[0] == 1, boundaries_only=True),
        'top': m.facets_satisfying(lambda x: x[1] == 1, boundaries_only=True),
        'buttom': m.facets_satisfying(lambda x: x[1] == 0, boundaries_only=True)})
    return np.concatenate([dofs[k].nodal['u'] for k in ('left', 'right', 'top', 'buttom')])


def solve_problem2(m, epsilon, f_load, element_type='P1', intorder=6, tol=1e-8, sigma=5, mode='parity', iters=None):
    """ref main/example1/run.py:225-272 (penalty): K2 = eps^2 A + eps^2 P + B, nodal-only Dirichlet set."""
    element = {'w': ElementTriP1() if element_type == 'P1' else ElementTriP2(), 'u': ElementTriMorley()}
    basis = {v: InteriorBasis(m, e, intorder=intorder) for v, e in element.items()}
    K1 = asm(laplace, basis['w'])
    f1 = asm(f_load, basis['w'])
    sw = solver_iter_mgcg(tol=tol, mode=mode)
    wh = solve(*condense(K1, f1, D=basis['w'].find_dofs()), solver=sw)
    fbasis = F.FacetBasis(m, element['u'])
    fbasis.interior = basis['u']                                   # one CSR pattern for A, B and P
    P = asm(F.penalty_1 + F.penalty_2 + F.penalty_3, fbasis, sigma=sigma)
    K2 = asm(epsilon ** 2 * a_load + b_load, basis['u']) + epsilon ** 2 * P          # run.py:260
    f2 = asm(wv_load, basis['w'], basis['u']) * wh
    su = solver_iter_mgcg(tol=tol, mode=mode)
    uh0 = solve(*condense(K2, f2, D=easy_boundary_penalty(m, basis['u'])), solver=su)
    if iters is not None:
        iters.append((sw.last['iters'], su.last['iters']))
    return uh0, basis, fbasis


def run_penalty(refine_time=7, epsilons=(1, 1e-2, 1e-4, 1e-6), element_type='P1', intorder=6, tol=1e-8, sigma=5,
                mode='parity', out=print):
    """The reference's 'ex3 + penalty' runs (log/ex3_range_2/ex3_P*_pen_*.csv)."""
    results = {}
    for epsilon in epsilons:
        f_load, exact_u, dexact_u, ddexact = make_problem_ex3(epsilon)
        m = MeshTri()
        rows = []
        for i in range(1, refine_time + 1):
            m.refine()
            uh0, basis, fbasis = solve_problem2(m, epsilon, f_load, element_type, intorder, tol, sigma, mode)
            L2u, Du, D2i = morley_error_norms(basis['u'], uh0, exact_u, dexact_u, ddexact)
            D2u = np.sqrt(D2i ** 2 + F.morley_facet_pnv(fbasis, uh0))             # run.py:527
            rows.append((L2u, L2u + Du, L2u + Du + D2u, np.sqrt(epsilon ** 2 * D2u ** 2 + Du ** 2)))
        results[epsilon] = np.array(rows)
        out('epsilon =', epsilon)
        for r in rows:
            out(' {:.3e}  {:.3e}  {:.3e}  {:.3e}'.format(*r))
    return results


def solve_problem1(m, epsilon, f_load, element_type='P1', solver_type='mgcg', intorder=5, tol=1e-8, mode='parity',
                   iters=None, maxiter=None):
    """ref main/example1/run.py:182-222 with the B200 package."""
    element = {'w': ElementTriP1() if element_type == 'P1' else ElementTriP2(), 'u': ElementTriMorley()}
    basis = {v: InteriorBasis(m, e, intorder=intorder) for v, e in element.items()}

    def make_solver():
        kw = {} if maxiter is None else {'maxiter': maxiter}       # ref main/example1/test/test_solver.py: maxiter=1000
        if solver_type == 'mgcg':
            return solver_iter_mgcg(tol=tol, mode=mode, **kw)
        if solver_type == 'pcg':
            return solver_iter_krylov(Precondition=True, tol=tol, **kw)
        raise Exception("Solver not supported")

    K1 = asm(laplace, basis['w'])
    f1 = asm(f_load, basis['w'])
    sw = make_solver()
    wh = solve(*condense(K1, f1, D=basis['w'].find_dofs()), solver=sw)
    K2 = asm(epsilon ** 2 * a_load + b_load, basis['u'])         # run.py:210, one fused pass
    f2 = asm(wv_load, basis['w'], basis['u']) * wh
    su = make_solver()
    uh0 = solve(*condense(K2, f2, D=easy_boundary(m, basis['u'])), solver=su)
    if iters is not None:
        iters.append((sw.last['iters'], su.last['iters']))
    return uh0, basis


def run(refine_time=6, epsilon_range=6, element_type='P2', solver_type='mgcg', intorder=5, tol=1e-8, mode='parity',
        out=print, maxiter=None):
    """Returns {epsilon: array (refine_time, 4) of L2, H1, H2, Energy} and prints the reference's tables."""
    results, iteration_log = {}, {}
    t0 = time.time()
    for j in range(epsilon_range):
        epsilon = 1 * 10 ** (-j)
        f_load, exact_u, dexact_u, ddexact = make_problem(epsilon)
        m = MeshTri()
        rows, its = [], []
        for i in range(1, refine_time + 1):
            m.refine()
            uh0, basis = solve_problem1(m, epsilon, f_load, element_type, solver_type, intorder, tol, mode, its, maxiter)
            L2u, Du, D2u = morley_error_norms(basis['u'], uh0, exact_u, dexact_u, ddexact)
            epu = np.sqrt(epsilon ** 2 * D2u ** 2 + Du ** 2)
            rows.append((L2u, L2u + Du, L2u + Du + D2u, epu))
        tab = np.array(rows)
        results[epsilon], iteration_log[epsilon] = tab, its
        L2s, H1s, H2s, epus = tab.T
        out('epsilon = {}'.format(epsilon))
        out('  h    L2u   H1u   H2u   epu')
        for i in range(tab.shape[0] - 1):
            out('2^-' + str(i + 2), ' {:.2f}  {:.2f}  {:.2f}  {:.2f}'.format(
                -np.log2(L2s[i + 1] / L2s[i]), -np.log2(H1s[i + 1] / H1s[i]), -np.log2(H2s[i + 1] / H2s[i]),
                -np.log2(epus[i + 1] / epus[i])))
            out('2^-' + str(i + 2), ' {:.3e}  {:.3e}  {:.3e}  {:.3e}'.format(L2s[i + 1], H1s[i + 1], H2s[i + 1], epus[i + 1]))
    out('Total Time Cost {:.2f} s'.format(time.time() - t0))
    return results, iteration_log


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--refine', type=int, default=6)
    ap.add_argument('--eps-range', type=int, default=6)
    ap.add_argument('--element', default='P2', choices=['P1', 'P2'])
    ap.add_argument('--solver', default='mgcg', choices=['mgcg', 'pcg'])
    ap.add_argument('--mode', default='parity', choices=['parity', 'chebyshev', 'jacobi'])
    a = ap.parse_args()
    print('=======Arguments=======')
    print('example:\tex1\npenalty:\tFalse\nelement_type:\t{}\nsolver_type:\t{}\ntol:\t1e-08\nintorder:\t5'.format(a.element, a.solver))
    print('refine_time:\t{}\nepsilon_range:\t{}'.format(a.refine, a.eps_range))
    print('=======Results=======')
    run(a.refine, a.eps_range, a.element, a.solver, mode=a.mode)
