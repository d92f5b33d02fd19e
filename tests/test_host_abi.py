"""CPU-only checks of the product's host side: the C-ABI library loads and exports every symbol the
header declares, the C++ AMG setup reproduces the oracle hierarchy, and nothing falls back to the CPU."""
import ctypes
import os
import re

import numpy as np
import pytest
import scipy.sparse as sp

import fem_forth_perturb_b200 as F
from fem_forth_perturb_b200 import _lib as L
from conftest import ROOT


def _declared():
    txt = open(os.path.join(ROOT, 'include', 'mwx_b200.h')).read()
    return sorted(set(re.findall(r'^(?:int|void)\s+(mwx_\w+)\s*\(', txt, flags=re.M)))


def test_library_exports_every_declared_symbol():
    lib = L.lib()
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f'{n} declared in include/mwx_b200.h but not exported'
        assert n in L._SIGNATURES, f'{n} has no ctypes signature'
    assert lib.mwx_version() >= 100


def test_ctypes_signatures_have_the_header_arity():
    """Every declaration of include/mwx_b200.h has as many parameters as its ctypes signature (a changed C signature
    that the Python side missed would otherwise only show up as garbage arguments on the GPU)."""
    txt = open(os.path.join(ROOT, 'include', 'mwx_b200.h')).read()
    txt = re.sub(r'/\*.*?\*/', '', txt, flags=re.S)                      # comments contain commas and parentheses
    decls = re.findall(r'^(?:int|void)\s+(mwx_\w+)\s*\(([^;]*?)\)\s*;', txt, flags=re.M | re.S)
    assert len(decls) >= 100
    for name, args in decls:
        args = args.strip()
        n = 0 if args in ('', 'void') else args.count(',') + 1
        assert n == len(L._SIGNATURES[name]), (name, n, len(L._SIGNATURES[name]))


def test_bench_launch_count_matches_the_profiled_cycle():
    """bench.py's `gpu_launches` claim: the cycle-shape formula reproduces the ncu launch list of one iteration
    (profiles/r02_launches_pcg_level12_sa.txt: 138 launches = 134 of the cycle + 4 CG kernels)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location('bench_mod', os.path.join(ROOT, 'bench.py'))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert bench.vcycle_launches(6, [2, 3, 4]) == 134
    assert bench.vcycle_launches(6, []) == 5 * 7 + 1
    assert bench.vcycle_launches(2, []) == 8
    listing = open(os.path.join(ROOT, 'profiles', 'r02_launches_pcg_level12_sa.txt')).read()
    assert 'launches 138' in listing


def test_coarse_inverse_small_is_pinv2():
    """AMGHierarchy._pinv_coarse up to 300 rows = scipy 1.4.1 pinv2 semantics (SVD with the 1e6 * eps cutoff)."""
    rng = np.random.default_rng(0)
    B = rng.standard_normal((40, 40))
    A = B @ B.T + 40 * np.eye(40)
    P = F.AMGHierarchy._pinv_coarse(A)
    assert abs(P @ A - np.eye(40)).max() <= 1e-11
    B = rng.standard_normal((60, 55))
    A = B @ B.T                                                     # rank deficient: singular values below the cutoff are dropped
    P = F.AMGHierarchy._pinv_coarse(A)
    assert abs(A @ P @ A - A).max() <= 1e-10 * abs(A).max()


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    with pytest.raises(F.MwxError):
        F.MeshTri()
    with pytest.raises(F.MwxError):
        F.solver_iter_pcg(tol=1e-8)(sp.identity(4, format='csr'), np.ones(4))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'fem_forth_perturb_b200')
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith(('.py', '.cu', '.cuh', '.cpp', '.h')):
                src = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle', src, flags=re.M), fn
                assert 'oracle/' not in src.replace('oracle/scipy_cg.py', ''), fn


def test_form_dispatch_is_by_name_and_rejects_unknown():
    f = 1e-4 * F.a_load + F.b_load
    assert f.terms == {'a_load': 1e-4, 'b_load': 1.0} and f.kind == 'interior'

    def a_load(u, v, w):          # a function named like the reference's form
        return None
    assert F.as_form(a_load).terms == {'a_load': 1.0}

    class SkfemLike:              # skfem BilinearForm keeps the python function in .form
        def __init__(self, fn):
            self.form = fn
    assert F.as_form(SkfemLike(a_load)).terms == {'a_load': 1.0}

    def linear_elasticity(u, v, w):
        return None
    with pytest.raises(NotImplementedError):
        F.as_form(linear_elasticity)
    with pytest.raises(NotImplementedError):
        (F.a_load + F.penalty_1).kind


def _k2c(level, eps):
    from oracle.skfem_mesh import MeshTri
    from oracle import pipeline as pl
    m = MeshTri()
    m.refine(level)
    S = pl.solve_problem(m, pl.Ex1(eps), wkind='P1', intorder=5, return_system=True)
    K = S['Kc']
    K.sort_indices()
    return K, S['bc']


def _native_levels(K):
    lib = L.lib()
    ip, ix, va = K.indptr.astype(np.int64), K.indices.astype(np.int32), K.data.astype(np.float64)
    h = ctypes.c_void_p()
    L.check(lib.mwx_amg_setup_host(K.shape[0], L.ptr(ip), L.ptr(ix), L.ptr(va), 0.25, 10, 10, ctypes.byref(h)))
    out = []
    nl = lib.mwx_amg_num_levels(h)
    for l in range(nl):
        mats = []
        for which in (0, 1, 2):
            if which and l == nl - 1:
                mats.append(None)
                continue
            dims = np.zeros(3, np.int64)
            L.check(lib.mwx_amg_level_dims(h, l, which, L.ptr(dims)))
            r, c, nnz = (int(v) for v in dims)
            p = np.empty(r + 1, np.int64); j = np.empty(nnz, np.int32); v = np.empty(nnz)
            L.check(lib.mwx_amg_level_copy(h, l, which, L.ptr(p), L.ptr(j), L.ptr(v)))
            mats.append(sp.csr_matrix((v, j, p), shape=(r, c)))
        split = None
        if l < nl - 1:
            split = np.empty(mats[0].shape[0], np.int32)
            L.check(lib.mwx_amg_level_splitting(h, l, L.ptr(split)))
        out.append((mats, split))
    lib.mwx_amg_destroy_host(h)
    return out


@pytest.mark.parametrize('level,eps', [(2, 1.0), (4, 1.0), (5, 1e-2), (6, 1e-4)])
def test_native_amg_setup_matches_oracle_hierarchy(level, eps):
    """mwx_amg_setup_host (C++) vs the oracle's pyamg restatement: identical C/F splittings and
    patterns on every level, operators equal to round-off."""
    from oracle.pyamg_rs import RugeStuben
    K, _ = _k2c(level, eps)
    ml = RugeStuben(K)
    nat = _native_levels(K)
    assert [m[0][0].shape[0] for m in nat] == [a.shape[0] for a in ml.A]
    for l, ((A, P, R), split) in enumerate(nat):
        Ao = ml.A[l]
        assert np.array_equal(A.indptr, Ao.indptr) and np.array_equal(A.indices, Ao.indices)
        assert abs(A - Ao).max() <= 1e-13 * abs(Ao).max()
        if split is not None:
            assert np.array_equal(split, ml.split[l])
            assert np.array_equal(P.indices, ml.P[l].indices)
            assert abs(P - ml.P[l]).max() <= 1e-14
            assert abs(R - ml.P[l].T).max() <= 1e-14


def _morley_system_with_candidates(level, eps):
    """Condensed K2 of the oracle plus the Morley interpolants of 1, x, y restricted to the free DOFs."""
    from oracle.skfem_mesh import MeshTri
    from oracle.skfem_basis import InteriorBasis
    from oracle import pipeline as pl
    m = MeshTri().refine(level)
    K = pl.assemble_K2(InteriorBasis(m, 'Morley', 5, shift=True), eps)
    bf = m.boundary_facets()
    D = np.concatenate((m.boundary_nodes(), bf + m.nv))
    I = np.setdiff1d(np.arange(K.shape[0]), D)
    A = K[I][:, I].tocsr()
    A.sort_indices()
    d = m.p[:, m.facets[1]] - m.p[:, m.facets[0]]
    ln = np.sqrt((d ** 2).sum(0))
    B = np.zeros((K.shape[0], 3))
    B[:m.nv, 0], B[:m.nv, 1], B[:m.nv, 2] = 1.0, m.p[0], m.p[1]
    B[m.nv:, 1], B[m.nv:, 2] = d[1] / ln, -d[0] / ln
    # the fourth-order part annihilates the candidates (sign convention of the facet normals)
    Kh = pl.assemble_K2(InteriorBasis(m, 'Morley', 5, shift=True), 1.0) - pl.assemble_K2(InteriorBasis(m, 'Morley', 5, shift=True), 0.0)
    assert abs(Kh @ B).max() <= 1e-12 * abs(Kh).max()
    return A, np.ascontiguousarray(B[I])


def test_smoothed_aggregation_host_setup():
    """mwx_amg_setup_sa_host: Galerkin consistency, seed-ordered aggregates, and a V-cycle that beats the
    Ruge-Stueben hierarchy on the fourth-order-dominated problem (the reason the mode exists)."""
    lib = L.lib()
    A, B = _morley_system_with_candidates(5, 0.5)
    n = A.shape[0]
    ip, ix, va = A.indptr.astype(np.int64), A.indices.astype(np.int32), A.data.astype(np.float64)
    h = ctypes.c_void_p()
    L.check(lib.mwx_amg_setup_sa_host(n, L.ptr(ip), L.ptr(ix), L.ptr(va), L.ptr(B), 3, 1, 0, 0.1, 10, 100, ctypes.byref(h)))
    nl = lib.mwx_amg_num_levels(h)
    assert nl >= 3

    def mat(l, which):
        dims = np.zeros(3, np.int64)
        L.check(lib.mwx_amg_level_dims(h, l, which, L.ptr(dims)))
        r, c, nnz = (int(v) for v in dims)
        p = np.empty(r + 1, np.int64); j = np.empty(nnz, np.int32); v = np.empty(nnz)
        L.check(lib.mwx_amg_level_copy(h, l, which, L.ptr(p), L.ptr(j), L.ptr(v)))
        return sp.csr_matrix((v, j, p), shape=(r, c))
    As = [mat(l, 0) for l in range(nl)]
    Ps = [mat(l, 1) for l in range(nl - 1)]
    assert abs(As[0] - A).max() == 0.0
    for l in range(nl - 1):
        assert As[l + 1].shape[0] % 3 == 0 and As[l + 1].shape[0] < As[l].shape[0] / 3
        assert abs(mat(l, 2) - Ps[l].T).max() == 0.0
        G = (Ps[l].T @ As[l] @ Ps[l]).tocsr()
        dead = G.diagonal() == 0.0               # a candidate that vanishes on its aggregate: unit diagonal instead
        assert dead.sum() < 0.2 * len(dead)
        assert abs(G + sp.diags(dead.astype(float)) - As[l + 1]).max() <= 1e-12 * abs(G).max()
        assert np.all(As[l + 1].diagonal() > 0)
        # aggregates are numbered by their seed DOF: roots ascending, 3 coarse DOFs each
        k = ctypes.c_int32(0)
        roots = np.empty(Ps[l].shape[1] // 3, np.int64)
        L.check(lib.mwx_amg_level_roots(h, l, L.ptr(roots), ctypes.byref(k)))
        assert k.value == 3 and np.all(np.diff(roots) > 0) and roots[-1] < As[l].shape[0]
        with pytest.raises(L.MwxError):
            L.check(lib.mwx_amg_level_splitting(h, l, L.ptr(np.empty(As[l].shape[0], np.int32))))
    # continuing from a coarse level of the same hierarchy reproduces the remaining levels (replicated tail)
    k = ctypes.c_int32(0)
    B1 = np.empty((As[1].shape[0], 3))
    L.check(lib.mwx_amg_level_candidates(h, 1, L.ptr(B1), ctypes.byref(k)))
    assert k.value == 3
    A1 = As[1]
    ip1, ix1, va1 = A1.indptr.astype(np.int64), A1.indices.astype(np.int32), A1.data.astype(np.float64)
    h1 = ctypes.c_void_p()
    L.check(lib.mwx_amg_setup_sa_host(A1.shape[0], L.ptr(ip1), L.ptr(ix1), L.ptr(va1), L.ptr(B1), 3, 3, 1, 0.1, 9, 100,
                                      ctypes.byref(h1)))
    assert lib.mwx_amg_num_levels(h1) == nl - 1
    for l in range(1, nl):
        dims = np.zeros(3, np.int64)
        L.check(lib.mwx_amg_level_dims(h1, l - 1, 0, L.ptr(dims)))
        assert (int(dims[0]), int(dims[2])) == (As[l].shape[0], As[l].nnz)
        v = np.empty(int(dims[2])); pp = np.empty(int(dims[0]) + 1, np.int64); jj = np.empty(int(dims[2]), np.int32)
        L.check(lib.mwx_amg_level_copy(h1, l - 1, 0, L.ptr(pp), L.ptr(jj), L.ptr(v)))
        assert np.array_equal(v, As[l].data) and np.array_equal(jj, As[l].indices)
    lib.mwx_amg_destroy_host(h1)
    lib.mwx_amg_destroy_host(h)

    def vcycle_pcg(As, Ps):
        dinv = [1.0 / a.diagonal() for a in As]
        coarse = np.linalg.pinv(As[-1].toarray())

        def cyc(l, b):
            if l == len(As) - 1:
                return coarse @ b
            x = 0.6 * dinv[l] * b
            x = x + 0.6 * dinv[l] * (b - As[l] @ x)
            x = x + Ps[l] @ cyc(l + 1, Ps[l].T @ (b - As[l] @ x))
            x = x + 0.6 * dinv[l] * (b - As[l] @ x)
            return x + 0.6 * dinv[l] * (b - As[l] @ x)
        rng = np.random.default_rng(0)
        b = A @ rng.uniform(-1, 1, n)
        x = np.zeros(n); r = b.copy(); z = cyc(0, r); p = z.copy(); rz = r @ z
        for it in range(1, 500):
            q = A @ p; a = rz / (p @ q); x += a * p; r -= a * q
            if np.linalg.norm(r) <= 1e-8 * np.linalg.norm(b):
                return it
            z = cyc(0, r); rzn = r @ z; p = z + (rzn / rz) * p; rz = rzn
        return 500
    from oracle.pyamg_rs import RugeStuben
    rs = RugeStuben(A)
    its_sa, its_rs = vcycle_pcg(As, Ps), vcycle_pcg(list(rs.A), list(rs.P))
    assert its_sa < 0.6 * its_rs, (its_sa, its_rs)
    # bad arguments
    assert lib.mwx_amg_setup_sa_host(n, L.ptr(ip), L.ptr(ix), L.ptr(va), None, 3, 1, 0, 0.1, 10, 100, ctypes.byref(h)) == -1
    assert lib.mwx_amg_setup_sa_host(n, L.ptr(ip), L.ptr(ix), L.ptr(va), L.ptr(B), 9, 1, 0, 0.1, 10, 100, ctypes.byref(h)) == -1


def test_host_setups_do_not_depend_on_the_thread_count():
    """Both hierarchies come out bit-identical with 1 and with many host threads (the partitioned solver rebuilds the
    coarse tail on every rank and relies on it)."""
    lib = L.lib()
    A, B = _morley_system_with_candidates(5, 0.3)
    n = A.shape[0]
    ip, ix, va = A.indptr.astype(np.int64), A.indices.astype(np.int32), A.data.astype(np.float64)

    def levels(sa, threads):
        lib.mwx_set_host_threads(threads)
        h = ctypes.c_void_p()
        if sa:
            L.check(lib.mwx_amg_setup_sa_host(n, L.ptr(ip), L.ptr(ix), L.ptr(va), L.ptr(B), 3, 1, 0, 0.1, 10, 50, ctypes.byref(h)))
        else:
            L.check(lib.mwx_amg_setup_host(n, L.ptr(ip), L.ptr(ix), L.ptr(va), 0.25, 10, 10, ctypes.byref(h)))
        out = []
        for l in range(lib.mwx_amg_num_levels(h)):
            for which in (0, 1):
                dims = np.zeros(3, np.int64)
                if lib.mwx_amg_level_dims(h, l, which, L.ptr(dims)) != 0:
                    continue
                p = np.empty(int(dims[0]) + 1, np.int64); j = np.empty(int(dims[2]), np.int32); v = np.empty(int(dims[2]))
                L.check(lib.mwx_amg_level_copy(h, l, which, L.ptr(p), L.ptr(j), L.ptr(v)))
                out.append((p, j, v))
        lib.mwx_amg_destroy_host(h)
        return out
    for sa in (False, True):
        one, many = levels(sa, 1), levels(sa, 7)
        assert len(one) == len(many) and len(one) >= 4
        for (p1, j1, v1), (p2, j2, v2) in zip(one, many):
            assert np.array_equal(p1, p2) and np.array_equal(j1, j2) and np.array_equal(v1, v2)
    lib.mwx_set_host_threads(os.cpu_count() or 1)


def test_amg_setup_tiny_and_bad_arguments():
    lib = L.lib()
    K = sp.identity(5, format='csr')
    nat = _native_levels(K)                       # n <= max_coarse: single level, no P
    assert len(nat) == 1 and nat[0][0][1] is None
    h = ctypes.c_void_p()
    assert lib.mwx_amg_setup_host(0, None, None, None, 0.25, 10, 10, ctypes.byref(h)) == -1
    assert lib.mwx_spmv(0, -1, None, None, None, None, None, None) == -1
    assert lib.mwx_pcg_jacobi(4, None, None, None, None, None, 1e-8, 0, 1, None, None, None, None, None) == -1


def test_gs_wavefronts_respect_dependencies():
    lib = L.lib()
    K, _ = _k2c(4, 1.0)
    n = K.shape[0]
    ip, ix = K.indptr.astype(np.int64), K.indices.astype(np.int32)
    for fn, backward in ((lib.mwx_gs_wavefronts_host, False), (lib.mwx_gs_wavefronts_backward_host, True)):
        order = np.empty(n, np.int32); wp = np.zeros(n + 1, np.int64); nw = ctypes.c_int64(0)
        L.check(fn(n, L.ptr(ip), L.ptr(ix), L.ptr(order), L.ptr(wp), ctypes.byref(nw)))
        nw = nw.value
        assert sorted(order.tolist()) == list(range(n)) and wp[nw] == n
        wave = np.empty(n, np.int64)
        for w in range(nw):
            wave[order[wp[w]:wp[w + 1]]] = w
        for i in range(n):
            nb = ix[ip[i]:ip[i + 1]]
            dep = nb[nb > i] if backward else nb[nb < i]
            assert np.all(wave[dep] < wave[i])          # every dependency sits in an earlier wavefront
            assert wave[i] == (wave[dep].max() + 1 if len(dep) else 0)
