"""The oracle against the reference's own golden artefacts (tests/golden/*.json, extracted from the
reference's committed logs by tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle.skfem_mesh import MeshTri
from oracle import pipeline as pl
from oracle.scipy_cg import solver_iter_mgcg, solver_iter_pcg, cg141

PROB = {'ex1': pl.Ex1, 'ex3': pl.Ex3, 'ex4': pl.Ex4}


def cfg_pen(g):
    return bool(g['config']['penalty'])


def _run_ladder(cfg, eps, levels, **solvers):
    mesh = MeshTri.init_lshaped() if cfg['example'] == 'ex4' else MeshTri()
    out = []
    for _ in range(levels):
        mesh.refine()
        prob = PROB[cfg['example']](eps)
        uh, ub, fb = pl.solve_problem(mesh, prob, wkind=cfg['w'], intorder=cfg['intorder'],
                                      penalty=cfg['penalty'], sigma=cfg.get('sigma', 5),
                                      inhomogeneous=(cfg['example'] == 'ex4'), **solvers)
        out.append(pl.error_norms(uh, ub, prob, fb))
    return np.array(out)


@pytest.mark.parametrize('key,levels', [('ex1_P1_nopen', 4), ('ex1_P2_nopen', 3), ('ex3_P1_pen', 3),
                                        ('ex3_P2_pen', 3), ('ex4_P1_nopen', 3)])
def test_error_norms_match_reference_logs(golden_norms, key, levels):
    g = golden_norms[key]
    for eps_s, rows in g['rows'].items():
        mine = _run_ladder(g['config'], float(eps_s), levels)
        gold = np.array(rows[:levels])
        rel = np.abs(mine - gold) / np.abs(gold)
        # level 1: one CG step is exact -> pins assembly + BCs + norms to round-off
        # (penalty path: 17 unknowns at level 1, so the reference's CG tolerance shows; its P1 log was
        #  also printed with 9-10 digits only)
        tol1 = 5e-9 if cfg_pen(g) else 1e-12
        assert rel[0].max() < tol1, (key, eps_s, rel[0])
        # finer levels: the reference stopped CG at tol 1e-8 (1e-11 for ex4), the oracle solves directly
        assert rel.max() < 2e-7, (key, eps_s, rel.max())


def test_level1_rows_to_16_digits(golden_norms):
    g = golden_norms['ex1_P1_nopen']
    mine = _run_ladder(g['config'], 1.0, 1)[0]
    gold = np.array(g['rows']['1.0'][0])
    assert np.max(np.abs(mine - gold) / gold) < 1e-14


def test_amg_pcg_iteration_table(golden_iterations):
    """pyamg RS hierarchy + symmetric GS + scipy-1.4.1 cg restated: the reference's printed counts."""
    g = golden_iterations
    levels = 6
    for eps_s in ['1.0', '0.01', '0.0001']:
        eps = float(eps_s)
        mesh = MeshTri()
        rw, ru = [], []
        for _ in range(levels):
            mesh.refine()
            pl.solve_problem(mesh, pl.Ex1(eps), wkind='P1', intorder=5,
                             solver_w=solver_iter_mgcg(tol=1e-8, maxiter=1000, record=rw),
                             solver_u=solver_iter_mgcg(tol=1e-8, maxiter=1000, record=ru))
        assert rw == g['w'][eps_s][:levels]
        assert ru == g['u'][eps_s][:levels]
        assert [(2 ** (k + 1) + 1) ** 2 for k in range(1, levels + 1)] == g['dofs'][eps_s][:levels]


def test_amg_pcg_iteration_table_ill_conditioned(golden_iterations):
    """eps = 0.1: counts are round-off sensitive - within 3 % of the log."""
    mesh = MeshTri()
    ru = []
    for _ in range(6):
        mesh.refine()
        pl.solve_problem(mesh, pl.Ex1(0.1), wkind='P1', intorder=5,
                         solver_u=solver_iter_mgcg(tol=1e-8, maxiter=1000, record=ru))
    gold = golden_iterations['u']['0.1'][:6]
    assert ru[:4] == gold[:4]
    assert all(abs(a - b) <= max(1, 0.03 * b) for a, b in zip(ru, gold))


def test_mesh_counts_and_pattern_size():
    """SURVEY section 8 header: nt = 2*4^k, N_u = (2^(k+1)+1)^2, nnz(K2) = 46*4^k + 16*2^k + 1."""
    from oracle.skfem_basis import InteriorBasis
    mesh = MeshTri()
    for k in range(1, 5):
        mesh.refine()
        assert mesh.nt == 2 * 4 ** k and mesh.nv == (2 ** k + 1) ** 2
        assert mesh.nf == 3 * 4 ** k + 2 * 2 ** k
        ub = InteriorBasis(mesh, 'Morley', 5)
        K2 = pl.assemble_K2(ub, 1.0)
        assert K2.shape[0] == (2 ** (k + 1) + 1) ** 2
        assert K2.nnz == 46 * 4 ** k + 16 * 2 ** k + 1


def test_cg141_legacy_semantics():
    import scipy.sparse as sp
    A = sp.diags([2.0] * 5).tocsr()
    b = np.ones(5)
    x, info, it = cg141(A, b, tol=1e-8)
    assert info == 0 and it == 1 and np.allclose(x, 0.5)
    x, info, it = cg141(A, 1e-9 * b, tol=1e-8)          # ||b|| <= tol: legacy absolute early exit
    assert it == 0 and np.all(x == 0)
    T = sp.diags([-np.ones(19), 2 * np.ones(20), -np.ones(19)], [-1, 0, 1]).tocsr()
    x, info, it = cg141(T, np.arange(20.0), M=lambda r: r, tol=1e-30, maxiter=3)
    assert info == 3 and it == 3                          # not converged: info = maxiter
    # Jacobi factory
    rec = []
    solver_iter_pcg(tol=1e-10, record=rec)(A, b)
    assert rec == [1]


# Tolerances of the example-2 comparison.  The reference's level-1..7 solutions came from mgcg at tol 1e-8 on an
# operator with eps = 1e-6; the oracle solves directly.  |u^7 - u_h|_1,h, |.|_2,h and the energy norm are far above
# that solver error on every level; ||u^7 - u_h||_0 falls to 1e-6..1e-4 on levels 5-6, where the reference's
# own CG error is a visible part of it (8 % for P2 with penalty), so that column is held to
# 2e-3 up to level 4 (the level-7 CG error alone moves it by 1.2e-3 at level 1 of P2/nopen) and looser below.
EX2_TOL = np.array([[2e-3, 1e-3, 1e-3, 1e-3]] * 4 + [[1e-2, 1e-3, 1e-3, 1e-3], [1e-1, 1e-3, 1e-3, 1e-3]])


@pytest.mark.parametrize('key', ['P1_pen', 'P2_nopen'])
def test_example2_tables_match_reference_logs(golden_example2, key):
    """main/example2 as shipped (base_order 7, levels 1-6, intorder 3, eps 1e-6) vs main/example2/log/ex3_*.csv."""
    from oracle import example2 as e2
    wkind, pen = key.split('_')
    mine = e2.table(base_order=7, refine_time=6, wkind=wkind, penalty=(pen == 'pen'))
    gold = np.array(golden_example2['tables'][key]['rows'])
    rel = np.abs(mine - gold) / np.abs(gold)
    assert (rel < EX2_TOL).all(), (key, rel)


def test_example2_dg_projections_are_the_elementwise_derivatives():
    """``project(uh, basis_from=Morley, basis_to=DG-P1, diff=d)`` (ref main/example2/run.py:114-116) done literally
    gives the elementwise derivative to round-off - the step the oracle and the device path skip."""
    from oracle import example2 as e2
    assert e2.dg_projection_check(level=3) < 1e-12
