#!/usr/bin/env python
"""main/example2/run.py on the B200 path - errors of the level-1..6 solutions against the level-7 solution.

Mirrors ref main/example2/run.py:13-24 (parameters), :78-179 (loop), :181-196 (table) and
ref main/example2/utils.py:625-721 (solve_problem1/2), :819-841 (show_result, load_solution).  The reference loads
its solutions from ``solutions/{P1,P2}_{pen,nopen}/uh0_{1..7}.npy``, which the repository does not ship;
``make_solutions`` produces them with the same call the files were made with
(``solve_problem2(m, element_type, 'mgcg', intorder=3, tol=1e-8, epsilon=1e-6)``, or ``solve_problem1`` without
penalty) and ``--save DIR`` writes them in the reference's layout.  The comparison itself (per-point
``interpolator`` loops + DG projections in the reference: 43 285 s in its own log) is two kernels here.

    python examples/example2.py [--element P1|P2] [--nopen] [--base 7] [--refine 6] [--mode parity] [--save DIR]
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import fem_forth_perturb_b200 as F
from fem_forth_perturb_b200 import MeshTri, morley_nested_error_norms, morley_facet_pnv
from example1 import make_problem_ex3, solve_problem1, solve_problem2

tol = 1e-8
intorder = 3
epsilon = 1e-6
sigma = 5


def show_result(L2s, H1s, H2s, epus, out=print):
    """ref main/example2/utils.py:819-832."""
    out('  h    L2u   H1u   H2u   epu')
    for i in range(H2s.shape[0] - 1):
        out('2^-' + str(i + 2), ' {:.2f}  {:.2f}  {:.2f}  {:.2f}'.format(
            -np.log2(L2s[i + 1] / L2s[i]), -np.log2(H1s[i + 1] / H1s[i]), -np.log2(H2s[i + 1] / H2s[i]),
            -np.log2(epus[i + 1] / epus[i])))
        out('2^-' + str(i + 2), ' {:.3e}  {:.3e}  {:.3e}  {:.3e}'.format(L2s[i + 1], H1s[i + 1], H2s[i + 1], epus[i + 1]))


def make_solutions(levels, element_type='P1', penalty=True, mode='parity', iters=None):
    """{level: (uh0, basis, fbasis, mesh)} - the contents of the reference's ``solutions/`` directory and the bases
    ``load_solution`` rebuilds beside them (ref main/example2/utils.py:835-841)."""
    f_load = make_problem_ex3(epsilon)[0]
    out = {}
    for i in levels:
        m = MeshTri().refine(i)
        if penalty:
            uh0, basis, fbasis = solve_problem2(m, epsilon, f_load, element_type, intorder, tol, sigma, mode, iters)
        else:
            uh0, basis = solve_problem1(m, epsilon, f_load, element_type, 'mgcg', intorder, tol, mode, iters)
            fbasis = None
        out[i] = (uh0, basis, fbasis, m)
    return out


def run(refine_time=6, base_order=7, element_type='P1', penalty=True, mode='parity', out=print, save=None, solutions=None):
    """Returns the (refine_time, 4) array of L2, H1, H2, Energy and prints the reference's table."""
    t0 = time.time()
    iters = []
    sol = solutions or make_solutions(list(range(1, refine_time + 1)) + [base_order], element_type, penalty, mode, iters)
    if save:
        d = os.path.join(save, '{}_{}'.format(element_type, 'pen' if penalty else 'nopen'))
        os.makedirs(d, exist_ok=True)
        for i, s in sol.items():
            np.save(os.path.join(d, 'uh0_{}.npy'.format(i)), np.asarray(s[0]))
    base_uh0, base_basis = sol[base_order][0], sol[base_order][1]
    rows = []
    for i in range(1, refine_time + 1):
        out('Testing order: ', i)
        uh0, basis, fbasis, _ = sol[i]
        L2u, Du, D2u = morley_nested_error_norms(base_basis['u'], base_uh0, basis['u'], uh0)
        if penalty:
            D2u = np.sqrt(D2u ** 2 + morley_facet_pnv(fbasis, uh0))                 # run.py:174-175
        rows.append((L2u, L2u + Du, L2u + Du + D2u, np.sqrt(epsilon ** 2 * D2u ** 2 + Du ** 2)))
    tab = np.array(rows)
    show_result(*tab.T, out=out)
    out('Total Time Cost {:.2f} s'.format(time.time() - t0))
    return tab, iters


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--element', default='P1', choices=['P1', 'P2'])
    ap.add_argument('--nopen', action='store_true')
    ap.add_argument('--base', type=int, default=7)
    ap.add_argument('--refine', type=int, default=6)
    ap.add_argument('--mode', default='parity', choices=['parity', 'chebyshev', 'jacobi', 'sa'])
    ap.add_argument('--save', default=None)
    a = ap.parse_args()
    print('=======Arguments=======')
    print('penalty:\t{}\nelement_type:\t{}\nbase_order:\t{}\nsolver_type:\tmgcg\ntol:\t{}\nintorder:\t{}'.format(
        not a.nopen, a.element, a.base, tol, intorder))
    print('refine_time:\t{}\nepsilon:\t{}\nsigma:\t{}'.format(a.refine, epsilon, sigma))
    print('=======Results=======')
    run(a.refine, a.base, a.element, not a.nopen, a.mode, save=a.save)
