"""SURVEY 8(f) rows 1-2 on the device, and the reference's example-1 tables end to end through the product path
(no oracle inside the run; the oracle and the reference's golden CSVs are only the checkers)."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT
from oracle.skfem_mesh import MeshTri as OMesh
from oracle.skfem_basis import InteriorBasis as OInterior
from oracle import skfem_asm as oa
from oracle import pipeline as pl

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('kind', ['P1', 'P2'])
@pytest.mark.parametrize('level', [1, 3, 4])
def test_w_path_matches_oracle(cuda, kind, level):
    import fem_forth_perturb_b200 as F
    om = OMesh().refine(level)
    gm = F.MeshTri().refine(level)
    elem = F.ElementTriP1() if kind == 'P1' else F.ElementTriP2()
    wb_o = OInterior(om, kind, 5)
    wb = F.InteriorBasis(gm, elem, intorder=5)
    # K1 = asm(laplace): same pattern after eliminate_zeros (5-point stencil for P1), values to round-off
    K1_o = oa.asm_bilinear(oa.laplace, wb_o)
    K1 = F.asm_gpu(F.laplace, wb).tocsr()
    if kind == 'P1':
        assert np.array_equal(K1.indptr, K1_o.indptr) and np.array_equal(K1.indices, K1_o.indices)
    else:
        # P2: entries that are mathematically zero come out of skfem's quadrature as ~1e-17 noise and stay in its
        # pattern ("round-off luck"); the closed form gives exact zeros, which eliminate_zeros drops.
        pat, ref = K1.copy(), K1_o.copy()
        pat.data[:] = 1.0; ref.data[:] = 1.0
        assert (pat - pat.multiply(ref)).nnz == 0                 # our pattern is a subset of the reference's
    assert abs(K1 - K1_o).max() <= 1e-12 * abs(K1_o).max()
    # f1 = asm(f_load)
    prob = pl.Ex1(0.1)
    f1_o = oa.asm_linear(prob.f, wb_o)

    @F.LinearForm
    def f_load(v, w):
        return prob.f(w.x[0], w.x[1]) * v
    f1 = F.asm_gpu(f_load, wb)
    assert np.max(np.abs(f1 - f1_o)) <= 1e-12 * np.max(np.abs(f1_o))
    # f2 = asm(wv_load, w, u) * wh
    ub_o = OInterior(om, 'Morley', 5, shift=True)
    ub = F.InteriorBasis(gm, F.ElementTriMorley(), intorder=5)
    wh = np.random.default_rng(0).standard_normal(wb.N)
    f2_o = oa.asm_bilinear(oa.wv_load, wb_o, ub_o) @ wh
    f2 = F.asm_gpu(F.wv_load, wb, ub) * wh
    assert np.max(np.abs(f2 - f2_o)) <= 1e-11 * np.max(np.abs(f2_o))
    # error functionals of a random Morley function
    uh = np.random.default_rng(1).standard_normal(ub.N)
    ex = pl.Ex1(1.0)
    L2, D1, D2 = F.morley_error_norms(ub, uh, ex.u, ex.du, ex.ddu)
    ub_o5 = OInterior(om, 'Morley', 5, shift=True)
    r = pl.error_norms(uh, ub_o5, ex)
    assert abs(L2 - r[0]) <= 1e-11 * r[0] and abs(L2 + D1 - r[1]) <= 1e-11 * r[1]
    assert abs(L2 + D1 + D2 - r[2]) <= 1e-11 * r[2]


@pytest.mark.parametrize('element,key', [('P1', 'ex1_P1_nopen'), ('P2', 'ex1_P2_nopen')])
def test_example1_tables_match_reference_logs(cuda, golden_norms, golden_iterations, element, key):
    """examples/example1.py = main/example1/run.py on the device: printed numbers agree with the reference's log
    to 3 significant digits (north_star), in fact to the CG tolerance; iteration counts within +-1 (P1 config)."""
    sys.path.insert(0, os.path.join(ROOT, 'examples'))
    import example1
    lines = []
    results, its = example1.run(refine_time=6, epsilon_range=3, element_type=element, out=lambda *a: lines.append(a))
    g = golden_norms[key]
    for eps, tab in results.items():
        gold = np.array(g['rows'][repr(float(eps))])
        rel = np.abs(tab - gold) / np.abs(gold)
        assert rel.max() < 5e-4, (eps, rel.max())                 # 3 significant digits
        assert rel[0].max() < 1e-11                               # level 1 is exact after one CG step
        for a, b in zip(tab[1:].ravel(), gold[1:].ravel()):
            assert '{:.3e}'.format(a)[:4] == '{:.3e}'.format(b)[:4] or abs(a - b) <= 6e-4 * abs(b)
    if element == 'P1':                                           # the golden iteration log is the P1 configuration
        for eps, pairs in its.items():
            gw, gu = golden_iterations['w'][repr(float(eps))][:6], golden_iterations['u'][repr(float(eps))][:6]
            assert all(abs(p[0] - a) <= 1 for p, a in zip(pairs, gw)), (eps, pairs, gw)
            assert all(abs(p[1] - a) <= (1 if eps != 0.1 else 2) for p, a in zip(pairs, gu)), (eps, pairs, gu)
    assert any('epsilon = 1' in ' '.join(map(str, l)) for l in lines)


def test_example1_in_sa_mode_reproduces_the_table(cuda, golden_norms):
    """The throughput mode (smoothed aggregation; constant candidate for the w-problem, Morley interpolants of 1, x, y
    for the u-problem) solves the same systems: the printed table still matches the reference's log to 3 digits."""
    sys.path.insert(0, os.path.join(ROOT, 'examples'))
    import example1
    results, its = example1.run(refine_time=6, epsilon_range=2, element_type='P1', mode='sa', out=lambda *a: None)
    g = golden_norms['ex1_P1_nopen']
    for eps, tab in results.items():
        gold = np.array(g['rows'][repr(float(eps))])
        assert (np.abs(tab - gold) / np.abs(gold)).max() < 5e-4, eps
    assert max(p[1] for pairs in its.values() for p in pairs) < 40


def test_penalty_example_tables_match_reference_logs(cuda, golden_norms):
    """solve_problem2 (Nitsche penalty, nodal-only BCs, L2pnvError) on the device vs log/ex3_range_2/ex3_P1_pen."""
    sys.path.insert(0, os.path.join(ROOT, 'examples'))
    import example1
    results = example1.run_penalty(refine_time=5, epsilons=(1, 1e-2, 1e-4), element_type='P1', out=lambda *a: None)
    g = golden_norms['ex3_P1_pen']
    for eps, tab in results.items():
        gold = np.array(g['rows'][repr(float(eps))][:5])
        rel = np.abs(tab - gold) / np.abs(gold)
        assert rel.max() < 5e-4, (eps, rel)
        assert rel[0].max() < 1e-7                                # level 1: 17 unknowns, CG tolerance 1e-8


def test_example3_lshape_tables_match_reference_log(cuda, golden_norms):
    """examples/example3.py = main/example3/run.py as shipped (L-shape, inhomogeneous data through a global L2
    projection, tol 1e-11, intorder 8) vs main/example3/log/ex4_P1_nopen_2021-02-22_12-51-55.csv."""
    sys.path.insert(0, os.path.join(ROOT, 'examples'))
    import example3
    results = example3.run(refine_time=5, out=lambda *a: None)
    g = golden_norms['ex4_P1_nopen']
    for eps, tab in results.items():
        gold = np.array(g['rows'][repr(float(eps))][:5])
        rel = np.abs(tab - gold) / np.abs(gold)
        assert rel.max() < 5e-4, (eps, rel)                       # 3 significant digits
        assert rel[0].max() < 1e-9                                # level 1


def test_project_is_the_l2_projection(cuda):
    import fem_forth_perturb_b200 as F
    gm = F.MeshTri.init_lshaped().refine(3)
    ub = F.InteriorBasis(gm, F.ElementTriMorley(), intorder=8)
    om = OMesh.init_lshaped().refine(3)
    ub_o = OInterior(om, 'Morley', 8, shift=True)
    prob = pl.Ex4(1e-5)
    M_o = oa.asm_bilinear(oa.mass, ub_o)
    x_o = pl.spl.spsolve(M_o.tocsc(), oa.asm_linear(prob.u, ub_o))
    M = F.asm_gpu(F.mass, ub).tocsr()
    assert abs(M - M_o).max() <= 1e-12 * abs(M_o).max()
    x = F.project(prob.u, ub)
    assert np.linalg.norm(x - x_o) <= 1e-10 * np.linalg.norm(x_o)


@pytest.mark.parametrize('name,eps', [('ex1', 0.1), ('ex3', 1.0)])
def test_device_problem_data_matches_host_samples(cuda, name, eps):
    """problems.Problem: the load and the exact solution evaluated by the kernels (mwx_load_vector_problem,
    mwx_morley_error_sums_problem) against the same Python functions sampled on the host."""
    import fem_forth_perturb_b200 as F
    gm = F.MeshTri().refine(4)
    prob = F.Problem(name, eps)
    uh = np.random.default_rng(5).standard_normal(gm.nv + gm.nf)
    ub = F.InteriorBasis(gm, F.ElementTriMorley(), intorder=5)
    for elem in (F.ElementTriP1(), F.ElementTriP2()):
        wb = F.InteriorBasis(gm, elem, intorder=5)
        f_host = F.asm_gpu(prob.f_load, wb)
        f_dev = F.asm_gpu(prob.f_load, wb, on_device=True).cpu().numpy()
        assert np.max(np.abs(f_dev - f_host)) <= 1e-12 * np.max(np.abs(f_host))
    host = F.morley_error_norms(ub, uh, prob.exact_u, prob.dexact_u, prob.ddexact)
    dev = F.morley_error_norms(ub, uh, problem=prob)
    assert np.allclose(dev, host, rtol=1e-12, atol=0)
    with pytest.raises(NotImplementedError):
        F.asm_gpu(F.LinearForm(lambda v, w: v), wb, on_device=True)


def test_example2_nested_errors_match_oracle(cuda):
    """mwx_morley_sample_nested + mwx_morley_error_sums vs oracle.example2.nested_errors on the same vectors
    (ref main/example2/run.py:101-172), penalty term included."""
    import fem_forth_perturb_b200 as F
    from oracle import example2 as e2
    base = e2.solve_level(4)
    gmN = F.MeshTri().refine(4)
    ubN = F.InteriorBasis(gmN, F.ElementTriMorley(), intorder=3)
    for lev in (1, 2, 3):
        coarse = e2.solve_level(lev)
        gm = F.MeshTri().refine(lev)
        ub = F.InteriorBasis(gm, F.ElementTriMorley(), intorder=3)
        fb = F.FacetBasis(gm, F.ElementTriMorley())
        fb.interior = ub
        L2, D1, D2 = F.morley_nested_error_norms(ubN, base[1], ub, coarse[1])
        D2 = np.sqrt(D2 ** 2 + F.morley_facet_pnv(fb, coarse[1]))
        ref = e2.nested_errors(base, coarse)
        assert np.allclose([L2, D1, D2], ref, rtol=1e-10, atol=0), (lev, (L2, D1, D2), ref)
    with pytest.raises(F.MwxError):                                  # not a refinement of the same mesh
        other = F.MeshTri.init_lshaped().refine(2)
        F.morley_nested_error_norms(ubN, base[1], F.InteriorBasis(other, F.ElementTriMorley(), intorder=3),
                                    np.zeros(other.nv + other.nf))


@pytest.mark.parametrize('element,penalty', [('P1', True), ('P1', False), ('P2', True), ('P2', False)])
def test_example2_tables_match_reference_logs(cuda, golden_example2, element, penalty):
    """examples/example2.py = main/example2/run.py as shipped (all four configurations of its log directory): levels
    1-6 against level 7, solutions produced by the device mgcg path (parity mode, tol 1e-8, intorder 3)."""
    sys.path.insert(0, os.path.join(ROOT, 'examples'))
    import example2
    tab, its = example2.run(refine_time=6, base_order=7, element_type=element, penalty=penalty, out=lambda *a: None)
    gold = np.array(golden_example2['tables']['{}_{}'.format(element, 'pen' if penalty else 'nopen')]['rows'])
    rel = np.abs(tab - gold) / np.abs(gold)
    # see test_oracle_golden.EX2_TOL; level 1 of P2/pen sits 1.4e-3..1.8e-3 from the log for the direct-solve oracle too
    # (the reference's own level-7 CG error), hence 2e-3 on that row
    tol = np.array([[2e-3] * 4] + [[2e-3, 1e-3, 1e-3, 1e-3]] * 3 + [[1e-2, 1e-3, 1e-3, 1e-3], [1e-1, 1e-3, 1e-3, 1e-3]])
    assert (rel < tol).all(), (element, penalty, rel)


def test_example1_full_epsilon_sweep_levels_1_to_8(cuda, golden_norms, golden_iterations):
    """BASELINE config 2: the whole eps sweep 1 -> 1e-6 on levels 1-8 through the device path in parity mode, against
    ref main/example1/test/log/test_solver_original_amg1_ex1_8_2020-11-20_12-58-38.{txt,csv} (maxiter = 1000).
    w-solve counts are exact.  u-solve counts are exact or +-1 below 100 iterations; runs of several hundred
    iterations are round-off sensitive (the CPU oracle, same arithmetic order as pyamg, also drifts there)
    and are held to 4 %.  eps = 1 at level 8 does not converge in 1000 iterations in the reference either (its log
    prints 1000 and the warning): the same warning and count are required here."""
    sys.path.insert(0, os.path.join(ROOT, 'examples'))
    import example1
    with pytest.warns(UserWarning, match='Convergence not achieved'):
        results, its = example1.run(refine_time=8, epsilon_range=7, element_type='P1', maxiter=1000, out=lambda *a: None)
    g = golden_norms['amg1_ex1_P1_8']
    for eps, pairs in its.items():
        key = repr(float(eps)) if eps >= 1e-5 else '1e-05'          # the log stops at 1e-5; 1e-6 behaves the same
        gw, gu = golden_iterations['w'][key], golden_iterations['u'][key]
        assert [p[0] for p in pairs] == gw, (eps, pairs, gw)
        for lev, (p, gold) in enumerate(zip(pairs, gu), start=1):
            slack = 1 if gold < 100 else 0.04 * gold
            assert abs(p[1] - gold) <= slack, (eps, lev, p[1], gold)
        if eps >= 1e-5:
            gold = np.array(g['rows'][key])
            rel = np.abs(results[eps] - gold) / np.abs(gold)
            assert rel.max() < 5e-4, (eps, rel.max(axis=1))           # 3 significant digits on every level
            assert rel[:6].max() < 1e-7
    assert its[1][7][1] == 1000
    # where the log is missed by more than one iteration the device count IS the CPU oracle's (recorded by
    # tests/golden/make_oracle_counts.py): the drift is the restatement's summation order, not the device path
    oracle_counts = {(1, 7): 249, (0.1, 7): 186, (0.1, 8): 475, (0.01, 8): 63}
    for (eps, lev), n in oracle_counts.items():
        assert its[eps][lev - 1][1] == n, (eps, lev, its[eps][lev - 1][1], n)
    # eps = 1e-6 has no log: levels 1-5 against the oracle with a direct solve
    from oracle.skfem_mesh import MeshTri as OM
    om = OM()
    for lev in range(5):
        om.refine()
        prob = pl.Ex1(1e-6)
        uh, ub, fb = pl.solve_problem(om, prob, wkind='P1', intorder=5)
        ref = np.array(pl.error_norms(uh, ub, prob))
        assert np.allclose(results[1e-6][lev], ref, rtol=1e-6), (lev, results[1e-6][lev], ref)
