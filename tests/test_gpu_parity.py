"""GPU parity tests: the CUDA path (through the C ABI / ctypes layer) against the CPU oracle.

Bars (BASELINE.json north_star): CSR indptr/indices bit-exact; matrix values within 1e-12 relative
(normwise, fp64); solutions within 1e-10 relative L2; CG iteration counts within +-1 of the reference.
"""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle.skfem_mesh import MeshTri as OMesh
from oracle.skfem_basis import InteriorBasis as OInterior, FacetBasis as OFacet
from oracle import skfem_asm as oa
from oracle import pipeline as pl
from oracle.scipy_cg import cg141, build_pc_diag as o_pc_diag
from oracle.pyamg_rs import RugeStuben

pytestmark = pytest.mark.gpu


def _meshes(level, lshaped=False):
    import fem_forth_perturb_b200 as F
    om = OMesh.init_lshaped() if lshaped else OMesh()
    gm = F.MeshTri.init_lshaped() if lshaped else F.MeshTri()
    om.refine(level)
    gm.refine(level)
    return om, gm


def _relmax(A, B):
    return abs(A - B).max() / abs(B).max()


# ------------------------------------------------------------------ mesh numbering (bit-exact)
@pytest.mark.parametrize('lshaped', [False, True])
def test_mesh_numbering_bit_exact(cuda, lshaped):
    import fem_forth_perturb_b200 as F
    om = OMesh.init_lshaped() if lshaped else OMesh()
    gm = F.MeshTri.init_lshaped() if lshaped else F.MeshTri()
    for level in range(0, 7):
        if level:
            om.refine(); gm.refine()
        assert (gm.nv, gm.nt, gm.nf) == (om.nv, om.nt, om.nf)
        assert np.array_equal(gm.p, om.p)                       # midpoints are exact in fp64
        assert np.array_equal(gm.t, om.t)
        assert np.array_equal(gm.facets, om.facets)
        assert np.array_equal(gm.t2f, om.t2f)
        assert np.array_equal(gm.boundary_facets(), om.boundary_facets())
        assert abs(gm.param() - om.param()) < 1e-15


def test_mesh_ragged_input_and_bad_shapes(cuda):
    import fem_forth_perturb_b200 as F
    # unsorted columns + a valence-7 fan: columns get sorted, numbering still matches the oracle
    rng = np.random.default_rng(3)
    ang = np.sort(rng.uniform(0, 2 * np.pi, 7))
    p = np.hstack((np.zeros((2, 1)), np.vstack((np.cos(ang), np.sin(ang)))))
    t = np.array([[0, 1 + k, 1 + (k + 1) % 7] for k in range(7)]).T[::-1]       # deliberately unsorted
    om, gm = OMesh(p, t), F.MeshTri(p, t)
    for _ in range(3):
        assert np.array_equal(gm.t, om.t) and np.array_equal(gm.facets, om.facets) and np.array_equal(gm.t2f, om.t2f)
        om.refine(); gm.refine()
    with pytest.raises(ValueError):
        F.MeshTri(np.zeros((3, 4)), np.zeros((3, 2), dtype=int))


# ------------------------------------------------------------------ CSR pattern + values
@pytest.mark.parametrize('level', [1, 2, 3, 4, 5])
@pytest.mark.parametrize('eps', [1.0, 1e-3])
def test_K2_pattern_bit_exact_and_values(cuda, level, eps):
    import fem_forth_perturb_b200 as F
    om, gm = _meshes(level)
    K_or = pl.assemble_K2(OInterior(om, 'Morley', 5), eps)
    basis = F.InteriorBasis(gm, F.ElementTriMorley(), intorder=5)
    K = F.asm_gpu(eps ** 2 * F.a_load + F.b_load, basis).tocsr()
    assert K.shape == K_or.shape
    assert K.nnz == 46 * 4 ** level + 16 * 2 ** level + 1
    assert np.array_equal(K.indptr, K_or.indptr)
    assert np.array_equal(K.indices, K_or.indices)
    assert K.indptr.dtype == np.int32 and K.indices.dtype == np.int32 and K.data.dtype == np.float64
    assert _relmax(K, K_or) < 1e-12                               # tolerance stated by north_star
    # two-call form of the reference (eps**2 * asm(a_load) + asm(b_load)) gives the same matrix
    K_two = (eps ** 2 * F.asm_gpu(F.a_load, basis) + F.asm_gpu(F.b_load, basis)).tocsr()
    assert _relmax(K_two, K_or) < 1e-12


@pytest.mark.parametrize('level', [3, 6, 7])
def test_values_against_well_conditioned_oracle(cuda, level):
    """skfem's global-monomial Vandermonde loses digits as h -> 0; the centroid-shifted
    oracle does not.  The closed-form kernel must sit at round-off from the well-conditioned one."""
    import fem_forth_perturb_b200 as F
    om, gm = _meshes(level)
    ub = OInterior(om, 'Morley', 5, shift=True)
    A_or = oa.asm_bilinear(oa.a_load, ub, eliminate=False)
    B_or = oa.asm_bilinear(oa.b_load, ub, eliminate=False)
    basis = F.InteriorBasis(gm, F.ElementTriMorley(), intorder=5)
    A = F.asm_gpu(F.a_load, basis).tocsr()
    B = F.asm_gpu(F.b_load, basis).tocsr()
    assert _relmax(A, A_or) < 2e-13 and _relmax(B, B_or) < 2e-13
    # individual forms: the reference drops entries that happen to be exactly 0.0; ours is a superset
    A_ref = oa.asm_bilinear(oa.a_load, OInterior(om, 'Morley', 5))
    assert A.nnz >= A_ref.nnz and A.nnz == 46 * 4 ** level + 16 * 2 ** level + 1
    pat = A.copy(); pat.data[:] = 1.0
    ref_pat = A_ref.copy(); ref_pat.data[:] = 1.0
    assert (ref_pat - ref_pat.multiply(pat)).nnz == 0              # reference pattern is a subset of ours
    assert abs(A - A_ref).max() <= 1e-10 * abs(A_ref).max()


def test_renumbered_mesh_is_the_same_mesh_and_assembles_bit_compatibly(cuda):
    """MeshTri.renumbered(): a vertex/element permutation of the same mesh; skfem's numbering rules applied to the
    permuted arrays (oracle built from the same p, t) give the same facets, pattern and values."""
    import fem_forth_perturb_b200 as F
    om0 = OMesh.init_lshaped().refine(4)
    rng = np.random.default_rng(3)
    inner = np.setdiff1d(np.arange(om0.nv), om0.boundary_nodes())
    p0 = om0.p.copy()
    p0[:, inner] += 0.004 * rng.standard_normal((2, len(inner)))   # generic triangles: no exactly-zero entries that
    gm0 = F.MeshTri(p0, om0.t)                                      # skfem's eliminate_zeros() would drop
    gm = gm0.renumbered()
    vo, eo = gm.vertex_order.cpu().numpy(), gm.element_order.cpu().numpy()
    assert np.array_equal(np.sort(vo), np.arange(gm0.nv)) and np.array_equal(np.sort(eo), np.arange(gm0.nt))
    assert np.array_equal(gm.p, gm0.p[:, vo])
    assert np.array_equal(np.sort(vo[gm.t], axis=0), np.sort(gm0.t[:, eo], axis=0))      # same triangles
    om = OMesh(gm.p, gm.t)
    assert np.array_equal(gm.t, om.t) and np.array_equal(gm.facets, om.facets) and np.array_equal(gm.t2f, om.t2f)
    K_or = pl.assemble_K2(OInterior(om, 'Morley', 5, shift=True), 0.1)
    K = F.asm_gpu(0.01 * F.a_load + F.b_load, F.InteriorBasis(gm, F.ElementTriMorley())).tocsr()
    assert np.array_equal(K.indptr, K_or.indptr) and np.array_equal(K.indices, K_or.indices)
    assert _relmax(K, K_or) < 1e-12
    # locality: neighbouring vertices of an element are close in the new numbering
    assert np.median(np.ptp(gm.t, axis=0)) < np.median(np.ptp(gm0.t, axis=0))


def test_assembly_accumulate_and_repeatability(cuda):
    """values += (accumulate path) and bit-reproducibility of the persistent row kernel."""
    import fem_forth_perturb_b200 as F
    gm = F.MeshTri().refine(6)
    basis = F.InteriorBasis(gm, F.ElementTriMorley())
    K1 = F.asm_gpu(0.25 * F.a_load + F.b_load, basis)
    K2 = F.asm_gpu(0.25 * F.a_load + F.b_load, basis)
    assert bool((K1.d_data == K2.d_data).all())                    # no atomics: identical bits run to run
    Ksum = (0.25 * F.asm_gpu(F.a_load, basis) + F.asm_gpu(F.b_load, basis)).tocsr()
    assert _relmax(Ksum, K1.tocsr()) < 1e-14
    Kacc = F.asm_gpu(0.25 * F.a_load, basis)
    F.asm_gpu(F.b_load, basis, out=Kacc.d_data, accumulate=True)
    assert _relmax(Kacc.tocsr(), K1.tocsr()) < 1e-14


def _fan_mesh(valence, seed=3):
    rng = np.random.default_rng(seed)
    ang = np.sort(rng.uniform(0, 2 * np.pi, valence))
    p = np.hstack((np.zeros((2, 1)), np.vstack((np.cos(ang), np.sin(ang)))))
    t = np.array([[0, 1 + k, 1 + (k + 1) % valence] for k in range(valence)]).T
    return p, t


@pytest.mark.parametrize('case', ['square3', 'square7', 'lshape5', 'fan7', 'perturbed'])
def test_tiled_kernel_equals_row_kernel_and_oracle(cuda, case):
    """The default (tiled, every element once) kernel against the row-gather kernel of round 1 and the oracle:
    same CSR, values to round-off, bit-identical run to run, accumulate path, partial tiles, irregular valence."""
    import fem_forth_perturb_b200 as F
    if case.startswith('square'):
        lvl = int(case[-1])
        om, gm = _meshes(lvl)
    elif case == 'lshape5':
        om, gm = _meshes(5, lshaped=True)
    elif case == 'fan7':
        p, t = _fan_mesh(7)
        om, gm = OMesh(p, t).refine(3), F.MeshTri(p, t).refine(3)
    else:
        om0 = OMesh().refine(5)
        rng = np.random.default_rng(11)
        inner = np.setdiff1d(np.arange(om0.nv), om0.boundary_nodes())
        p = om0.p.copy()
        p[:, inner] += 0.004 * rng.standard_normal((2, len(inner)))
        om, gm = OMesh(p, om0.t), F.MeshTri(p, om0.t)
    basis = F.InteriorBasis(gm, F.ElementTriMorley(), intorder=5)
    form = 0.04 * F.a_load + F.b_load
    Kt = F.asm_gpu(form, basis, kernel='tiles')
    Kr = F.asm_gpu(form, basis, kernel='rows')
    assert F.assembly_kernel_name(basis).startswith('k_assemble_morley_tiles')
    info = basis.pattern()['tile_plan'].info
    assert info['tiles'] == -(-basis.N // info['rows_per_tile']) and info['rows_per_tile'] in (16, 32, 64, 128, 256)
    scale = float(Kr.d_data.abs().max())
    # two evaluation orders of the same closed form; the random fan has needle triangles (cancellation in both)
    assert float((Kt.d_data - Kr.d_data).abs().max()) <= (5e-13 if case == 'fan7' else 2e-14) * scale
    Kt2 = F.asm_gpu(form, basis, kernel='tiles')
    assert bool((Kt.d_data == Kt2.d_data).all())                   # fixed summation order: identical bits
    Ksym = Kt.tocsr()
    assert abs(Ksym - Ksym.T).max() == 0.0                         # one stored value per symmetric pair of the local matrix
    Kacc = F.asm_gpu(0.04 * F.a_load, basis, kernel='tiles')
    F.asm_gpu(F.b_load, basis, out=Kacc.d_data, accumulate=True, kernel='tiles')
    assert float((Kacc.d_data - Kt.d_data).abs().max()) <= 2e-14 * scale
    if case != 'square7':
        K_or = pl.assemble_K2(OInterior(om, 'Morley', 5, shift=True), 0.2)
        K = Kt.tocsr()
        if case != 'lshape5':      # on the L-shape a few hundred entries are EXACT zeros that skfem's eliminate_zeros()
            assert np.array_equal(K.indptr, K_or.indptr) and np.array_equal(K.indices, K_or.indices)   # drops
        assert _relmax(K, K_or) < 1e-12


def test_tile_plan_falls_back_on_high_valence(cuda):
    """A vertex shared by 20 triangles exceeds the plan's limits (16 elements / 64 entries per row): asm_gpu then uses
    the row-gather kernel - same matrix, no error."""
    import fem_forth_perturb_b200 as F
    p, t = _fan_mesh(20)
    om, gm = OMesh(p, t), F.MeshTri(p, t)
    basis = F.InteriorBasis(gm, F.ElementTriMorley(), intorder=5)
    K = F.asm_gpu(0.3 * F.a_load + F.b_load, basis)
    assert basis.pattern()['tile_plan'] is None
    assert F.assembly_kernel_name(basis).startswith('k_element_geometry')
    with pytest.raises(F.MwxError):
        F.asm_gpu(F.a_load, basis, kernel='tiles')
    K_or = pl.assemble_K2(OInterior(om, 'Morley', 5, shift=True), np.sqrt(0.3))
    assert np.array_equal(K.tocsr().indices, K_or.indices) and _relmax(K.tocsr(), K_or) < 1e-12


def test_generic_triangles_lshape(cuda):
    """Non-right, differently oriented triangles: perturbed L-shape mesh."""
    import fem_forth_perturb_b200 as F
    om = OMesh.init_lshaped().refine(3)
    rng = np.random.default_rng(0)
    inner = np.setdiff1d(np.arange(om.nv), om.boundary_nodes())
    p = om.p.copy()
    p[:, inner] += 0.02 * rng.standard_normal((2, len(inner)))
    om2, gm = OMesh(p, om.t), F.MeshTri(p, om.t)
    K_or = pl.assemble_K2(OInterior(om2, 'Morley', 5, shift=True), 0.3)
    K = F.asm_gpu(0.09 * F.a_load + F.b_load, F.InteriorBasis(gm, F.ElementTriMorley())).tocsr()
    assert np.array_equal(K.indptr, K_or.indptr) and np.array_equal(K.indices, K_or.indices)
    assert _relmax(K, K_or) < 1e-12


@pytest.mark.parametrize('level', [1, 2, 4])
def test_penalty_forms(cuda, level):
    import fem_forth_perturb_b200 as F
    om, gm = _meshes(level)
    fb_o = OFacet(om, 'Morley', shift=True)
    p1, p2, p3 = (oa.asm_bilinear(f, fb_o, eliminate=False) for f in oa.make_penalty_forms(fb_o, 5.0))
    fb = F.FacetBasis(gm, F.ElementTriMorley())
    assert np.array_equal(fb.find, fb_o.find) and np.array_equal(fb.tind, fb_o.tind)
    for form, ref in ((F.penalty_1, p1), (F.penalty_2, p2), (F.penalty_3, p3),
                      (F.penalty_1 + F.penalty_2 + F.penalty_3, p1 + p2 + p3)):
        P = F.asm_gpu(form, fb, sigma=5.0).tocsr()
        assert abs(P - ref).max() <= 1e-12 * abs(ref).max()
    # K2 with penalty, as ref main/example1/run.py:260
    eps = 0.1
    ib = F.InteriorBasis(gm, F.ElementTriMorley())
    fb2 = F.FacetBasis(gm, F.ElementTriMorley())
    fb2.interior = ib                                            # share the pattern object
    K2 = eps ** 2 * F.asm_gpu(F.a_load, ib) + eps ** 2 * F.asm_gpu(F.penalty_1 + F.penalty_2 + F.penalty_3, fb2) \
        + F.asm_gpu(F.b_load, ib)
    K_or = pl.assemble_K2(OInterior(om, 'Morley', 6), eps, OFacet(om, 'Morley'), 5.0)
    assert _relmax(K2.tocsr(), K_or) < 1e-12


# ------------------------------------------------------------------ condense / SpMV
@pytest.mark.parametrize('level,penalty', [(1, False), (3, False), (4, True)])
def test_condense_matches_reference_semantics(cuda, level, penalty):
    import fem_forth_perturb_b200 as F
    om, gm = _meshes(level)
    ub = OInterior(om, 'Morley', 5)
    K_or = pl.assemble_K2(ub, 0.5)
    rng = np.random.default_rng(level)
    b = rng.standard_normal(K_or.shape[0])
    x0 = rng.standard_normal(K_or.shape[0])
    D = oa.easy_boundary_penalty(ub) if penalty else oa.easy_boundary(ub)
    Ko, bo, _, Io = oa.condense(K_or, b, x0, D=np.unique(D))
    K = F.asm_gpu(0.25 * F.a_load + F.b_load, F.InteriorBasis(gm, F.ElementTriMorley()))
    Kc, bc, x, I = F.condense(K, b, x0, D=D)
    Kc = Kc.tocsr(); Ko.sort_indices()
    assert np.array_equal(I, Io)
    assert np.array_equal(Kc.indptr, Ko.indptr) and np.array_equal(Kc.indices, Ko.indices)
    assert _relmax(Kc, Ko) < 1e-12
    assert np.max(np.abs(bc - bo)) <= 1e-12 * np.max(np.abs(bo))
    n_expect = K_or.shape[0] - (4 if penalty else 8) * 2 ** level
    assert Kc.shape == (n_expect, n_expect)


def test_spmv_matches_scipy_and_is_linear(cuda):
    import fem_forth_perturb_b200 as F
    om, gm = _meshes(6)
    K = F.asm_gpu(1e-2 * F.a_load + F.b_load, F.InteriorBasis(gm, F.ElementTriMorley()))
    Ks = K.tocsr()
    rng = np.random.default_rng(1)
    x, y = rng.standard_normal(Ks.shape[0]), rng.standard_normal(Ks.shape[0])
    ref = Ks @ x
    assert np.max(np.abs(K.dot(x) - ref)) <= 1e-13 * np.max(np.abs(ref))
    lin = K.dot(2.0 * x - 3.0 * y) - (2.0 * K.dot(x) - 3.0 * K.dot(y))
    assert np.max(np.abs(lin)) <= 1e-12 * np.max(np.abs(ref))
    # empty rows, a very long row (exceeds the shared-memory tile) and a 1x1 matrix
    n = 700
    R = sp.random(n, n, density=0.01, random_state=2, format='lil')
    R[5, :] = 1.0
    R[7, :] = 0.0
    R = R.tocsr()
    Rd = F.DeviceCSR.from_scipy(R)
    v = rng.standard_normal(n)
    assert np.allclose(Rd.dot(v), R @ v, rtol=1e-13, atol=1e-13)
    one = F.DeviceCSR.from_scipy(sp.csr_matrix(np.array([[3.0]])))
    assert np.allclose(one.dot(np.array([2.0])), [6.0])


# ------------------------------------------------------------------ size-independent properties at scale
def test_properties_at_4M_dofs(cuda):
    """Level 10 (4.2 M DOFs): polynomial exactness of the Morley space gives closed-form energies."""
    import fem_forth_perturb_b200 as F
    torch = cuda
    gm = F.MeshTri().refine(10)
    basis = F.InteriorBasis(gm, F.ElementTriMorley())
    eps2 = 1e-4
    K = F.asm_gpu(eps2 * F.a_load + F.b_load, basis)
    assert K.shape[0] == (2 ** 11 + 1) ** 2 and K.nnz == 46 * 4 ** 10 + 16 * 2 ** 10 + 1
    dev = gm.device
    pxy = gm.d_pxy.reshape(gm.nv, 2)
    fac = gm.d_facets.reshape(2, gm.nf).long()
    mid = 0.5 * (pxy[fac[0]] + pxy[fac[1]])
    d = pxy[fac[0]] - pxy[fac[1]]
    nrm = torch.stack((-d[:, 1], d[:, 0]), dim=1) / torch.sqrt((d * d).sum(dim=1, keepdim=True))

    def interp(f, grad):            # Morley DOFs: vertex values, then normal derivatives at edge midpoints
        g = grad(mid[:, 0], mid[:, 1])
        return torch.cat((f(pxy[:, 0], pxy[:, 1]), g[0] * nrm[:, 0] + g[1] * nrm[:, 1]))

    one = interp(lambda x, y: torch.ones_like(x), lambda x, y: (torch.zeros_like(x), torch.zeros_like(x)))
    lin = interp(lambda x, y: x, lambda x, y: (torch.ones_like(x), torch.zeros_like(x)))
    quad = interp(lambda x, y: x * x, lambda x, y: (2 * x, torch.zeros_like(x)))
    k1 = K.dot(one)
    assert float(k1.abs().max()) < 1e-9                       # constants are in the kernel
    # round-off budget of u.Ku: n * eps_mach * max|K| * max|u|^2 = 4.2e6 * 1.1e-16 * ~100 ~ 5e-8
    assert abs(float(torch.dot(lin, K.dot(lin))) - 1.0) < 2e-7                         # |grad x|^2 = 1
    assert abs(float(torch.dot(quad, K.dot(quad))) - (4 * eps2 + 4.0 / 3.0)) < 2e-7   # 4 eps^2 + int 4x^2
    Kb = F.asm_gpu(F.b_load, basis)                            # max|K| ~ 4 -> budget ~ 2e-9
    assert abs(float(torch.dot(lin, Kb.dot(lin))) - 1.0) < 1e-8
    assert abs(float(torch.dot(quad, Kb.dot(quad))) - 4.0 / 3.0) < 1e-8
    Ka = F.asm_gpu(F.a_load, basis)                            # hess(x^2) = diag(2, 0): |hess|^2 = 4
    amax = float(Ka.d_data.abs().max())                        # ~ h^-2 = 4e6: budget n * eps_mach * max|A|
    assert abs(float(torch.dot(quad, Ka.dot(quad))) - 4.0) < 1.1e-16 * K.shape[0] * amax
    assert float(Ka.dot(lin).abs().max()) < 1e-13 * amax
    # symmetry: x.Ky == y.Kx
    g = torch.Generator(device=dev).manual_seed(7)
    x = torch.rand(K.shape[0], dtype=torch.float64, device=dev, generator=g)
    y = torch.rand(K.shape[0], dtype=torch.float64, device=dev, generator=g)
    a, b = float(torch.dot(x, K.dot(y))), float(torch.dot(y, K.dot(x)))
    assert abs(a - b) <= 1e-12 * abs(a)
    # assembly is deterministic (no atomics): bit-identical values on a second pass
    K_again = F.asm_gpu(eps2 * F.a_load + F.b_load, basis)
    assert torch.equal(K.d_data, K_again.d_data)


def test_8k_mesh_needs_int64_indptr(cuda):
    """Level 13 (8192^2 squares): 268 M DOFs, nnz = 3 087 138 817 > 2^31 - the 8k x 8k microbench config."""
    import fem_forth_perturb_b200 as F
    torch = cuda
    free, _ = torch.cuda.mem_get_info()
    if free < 100e9:
        pytest.skip('needs ~80 GB of free HBM')
    gm = F.MeshTri().refine(13)
    basis = F.InteriorBasis(gm, F.ElementTriMorley())
    K = F.asm_gpu(1e-4 * F.a_load + F.b_load, basis)
    assert K.shape[0] == (2 ** 14 + 1) ** 2 == 268468225
    assert K.nnz == 46 * 4 ** 13 + 16 * 2 ** 13 + 1 == 3087138817 and K.nnz > 2 ** 31
    assert int(K.d_indptr[-1]) == K.nnz and K.d_indptr.dtype == torch.int64
    one = torch.cat((torch.ones(gm.nv, dtype=torch.float64, device=gm.device),
                     torch.zeros(gm.nf, dtype=torch.float64, device=gm.device)))
    assert float(K.dot(one).abs().max()) < 1e-9                      # constants are in the kernel of K2
    g = torch.Generator(device=gm.device).manual_seed(1)
    x = torch.rand(K.shape[0], dtype=torch.float64, device=gm.device, generator=g)
    y = torch.rand(K.shape[0], dtype=torch.float64, device=gm.device, generator=g)
    a, b = float(torch.dot(x, K.dot(y))), float(torch.dot(y, K.dot(x)))
    assert abs(a - b) <= 1e-12 * abs(a)
    del K, basis, gm, x, y, one
    torch.cuda.empty_cache()


# ------------------------------------------------------------------ PCG (Jacobi) parity
@pytest.mark.parametrize('level,eps', [(2, 1.0), (4, 1.0), (5, 1e-2), (6, 1e-4)])
def test_pcg_jacobi_matches_oracle_cg(cuda, level, eps):
    import fem_forth_perturb_b200 as F
    om = OMesh().refine(level)
    S = pl.solve_problem(om, pl.Ex1(eps), wkind='P1', intorder=5, return_system=True)
    Kc, bc = S['Kc'], S['bc']
    x_or, info_or, it_or = cg141(Kc, bc, M=o_pc_diag(Kc), tol=1e-8)
    solver = F.solver_iter_krylov(Precondition=True, tol=1e-8)
    x, it = solver.solve_with_info(Kc, bc)
    assert solver.last['info'] == 0 and info_or == 0
    assert abs(it - it_or) <= 1                                  # +-1 iteration bar
    # same algorithm, same stopping rule: the two iterates differ by round-off only
    assert np.linalg.norm(x - x_or) <= 1e-9 * np.linalg.norm(x_or)
    x_d = pl.direct_solver(Kc, bc)
    assert np.linalg.norm(x - x_d) <= 1e-5 * np.linalg.norm(x_d)          # tol 1e-8 times conditioning
    if eps <= 1e-2:
        # 1e-10 solution parity needs a system conditioned well enough to converge below it
        x_t, _ = F.solver_iter_krylov(Precondition=True, tol=1e-12).solve_with_info(Kc, bc)
        assert np.linalg.norm(x_t - x_d) <= 1e-10 * np.linalg.norm(x_d)
    # unpreconditioned flavour (Precondition omitted -> plain CG), as solver_iter_krylov allows
    x_p, it_p = F.solver_iter_krylov(tol=1e-8).solve_with_info(Kc, bc)
    _, _, it_po = cg141(Kc, bc, M=None, tol=1e-8)
    assert abs(it_p - it_po) <= max(1, it_po // 20)        # hundreds of its: round-off sensitive


def test_verbose_prints_the_iterate_norms_like_the_reference_callback(cuda, capsys):
    """``verbose=True``: scipy calls ``callback(x)`` every iteration and the reference prints ``np.linalg.norm(x)``
    (ref main/example1/utils.py:104-110, 229-235); here the norms are recorded on the device and printed after the solve."""
    import fem_forth_perturb_b200 as F
    om = OMesh().refine(4)
    S = pl.solve_problem(om, pl.Ex1(1e-2), wkind='P1', intorder=5, return_system=True)
    Kc, bc = S['Kc'], S['bc']
    norms = []
    x_or, info_or, it_or = cg141(Kc, bc, M=o_pc_diag(Kc), tol=1e-8, callback=lambda x: norms.append(np.linalg.norm(x)))
    for factory, kw in ((F.solver_iter_krylov, dict(Precondition=True)), (F.solver_iter_mgcg, dict())):
        capsys.readouterr()
        solver = factory(verbose=True, tol=1e-8, **kw)
        x, it = solver.solve_with_info(Kc, bc)
        lines = capsys.readouterr().out.strip().splitlines()
        assert lines[-1] == 'cg converged to tol=1e-08 and atol=default'
        printed = np.array([float(v) for v in lines[:-1]])
        assert len(printed) == it
        assert abs(printed[-1] - np.linalg.norm(x)) <= 1e-12 * printed[-1]
        if factory is F.solver_iter_krylov:                       # same algorithm as the oracle: same sequence of norms
            assert it == it_or and np.allclose(printed, norms[:it], rtol=1e-9)


def test_pcg_edge_cases(cuda):
    import fem_forth_perturb_b200 as F
    A = sp.diags([2.0] * 6).tocsr()
    s = F.solver_iter_pcg(Precondition=True, tol=1e-8)
    x, it = s.solve_with_info(A, np.ones(6))
    assert it == 1 and np.allclose(x, 0.5)
    x, it = s.solve_with_info(A, 1e-9 * np.ones(6))             # ||b|| <= tol: scipy-1.4.1 legacy exit
    assert it == 0 and np.all(x == 0)
    T = sp.diags([-np.ones(199), 2 * np.ones(200), -np.ones(199)], [-1, 0, 1]).tocsr()
    with pytest.warns(UserWarning, match='Convergence not achieved'):
        s2 = F.solver_iter_krylov(Precondition=True, tol=1e-30, maxiter=3)
        s2(T, np.arange(200.0))
    assert s2.last['iters'] == 3 and s2.last['info'] == 3
    with pytest.raises(NotImplementedError):
        F.solver_iter_krylov(M=lambda r: r)(A, np.ones(6))


# ------------------------------------------------------------------ AMG: V-cycle + iteration parity
@pytest.mark.parametrize('level,eps', [(3, 1.0), (5, 1e-2), (6, 1.0)])
def test_vcycle_parity_mode_equals_pyamg_restatement(cuda, level, eps):
    """One V(1,1) cycle with exact lexicographic symmetric Gauss-Seidel: z must equal the oracle's."""
    import fem_forth_perturb_b200 as F
    om = OMesh().refine(level)
    S = pl.solve_problem(om, pl.Ex1(eps), wkind='P1', intorder=5, return_system=True)
    Kc, bc = S['Kc'], S['bc']
    Kc.sort_indices()
    ml_o = RugeStuben(Kc)
    ml = F.AMGHierarchy(Kc, mode='parity')
    assert ml.level_sizes() == [a.shape[0] for a in ml_o.A]
    z_o = ml_o.precondition(bc)
    z = ml.precondition(bc)
    assert np.linalg.norm(z - z_o) <= 1e-11 * np.linalg.norm(z_o)


def test_amg_pcg_iteration_counts_match_reference_log(cuda, golden_iterations):
    """solver_iter_mgcg in parity mode vs the reference's printed counts (levels 1-6)."""
    import fem_forth_perturb_b200 as F
    for eps_s, slack in (('1.0', 1), ('0.01', 1), ('0.0001', 1), ('0.1', 2)):
        eps = float(eps_s)
        om = OMesh()
        mine = []
        for level in range(1, 7):
            om.refine()
            S = pl.solve_problem(om, pl.Ex1(eps), wkind='P1', intorder=5, return_system=True)
            solver = F.solver_iter_mgcg(tol=1e-8, maxiter=1000)
            x, it = solver.solve_with_info(S['Kc'], S['bc'])
            mine.append(it)
        gold = golden_iterations['u'][eps_s][:6]
        assert all(abs(a - b) <= slack for a, b in zip(mine, gold)), (eps_s, mine, gold)


@pytest.mark.parametrize('level,eps', [(7, 1e-2), (7, 1e-4), (8, 1e-2), (8, 1e-4)])
def test_amg_pcg_iteration_counts_levels_7_8(cuda, golden_iterations, level, eps):
    """The larger rows of the reference's table (66 049 and 263 169 DOFs): parity mode within +-1 of the log."""
    import fem_forth_perturb_b200 as F
    om = OMesh().refine(level)
    S = pl.solve_problem(om, pl.Ex1(eps), wkind='P1', intorder=5, return_system=True)
    solver = F.solver_iter_mgcg(tol=1e-8, maxiter=1000)
    x, it = solver.solve_with_info(S['Kc'], S['bc'])
    gold = golden_iterations['u'][repr(float(eps))][level - 1]
    assert solver.last['info'] == 0 and abs(it - gold) <= 1, (level, eps, it, gold)


@pytest.mark.parametrize('mode', ['chebyshev', 'jacobi'])
def test_amg_fast_modes_converge_to_the_same_solution(cuda, mode):
    import fem_forth_perturb_b200 as F
    om = OMesh().refine(6)
    S = pl.solve_problem(om, pl.Ex1(1e-2), wkind='P1', intorder=5, return_system=True)
    Kc, bc = S['Kc'], S['bc']
    solver = F.solver_iter_mgcg(tol=1e-12, mode=mode)
    x, it = solver.solve_with_info(Kc, bc)
    assert solver.last['info'] == 0 and it < 80
    x_d = pl.direct_solver(Kc, bc)
    assert np.linalg.norm(x - x_d) <= 1e-10 * np.linalg.norm(x_d)


def test_smoothed_aggregation_mode(cuda, monkeypatch):
    """mode='sa' (throughput): same solution as the reference hierarchy, far fewer iterations where the fourth-order
    term dominates; candidates travel with the matrix through condense and the solver's renumbering."""
    import fem_forth_perturb_b200 as F
    monkeypatch.setenv('MWX_SA_MAX_COARSE', '300')          # a deep hierarchy at this small size (default: stop at <= 4000 rows)
    gm = F.MeshTri().refine(6)
    basis = F.InteriorBasis(gm, F.ElementTriMorley())
    K = F.asm_gpu(0.25 * F.a_load + F.b_load, basis)
    B = basis.nullspace_candidates().cpu().numpy()
    Ah = F.asm_gpu(F.a_load, basis).tocsr()
    assert abs(Ah @ B).max() <= 1e-12 * abs(Ah).max()             # hess:hess annihilates the interpolants of 1, x, y
    bf = gm.boundary_facets()
    D = np.concatenate((np.unique(gm.facets[:, bf]), bf + gm.nv))
    rng = np.random.default_rng(11)
    f = rng.standard_normal(K.shape[0])
    Kc, bc, _, I = F.condense(K, f, D=D)
    assert np.array_equal(Kc.candidates().cpu().numpy(), B[I])
    x_d = pl.direct_solver(Kc.tocsr(), bc)
    its = {}
    for mode in ('chebyshev', 'sa'):
        solver = F.solver_iter_mgcg(tol=1e-10, mode=mode)
        x, it = solver.solve_with_info(Kc, bc)
        assert solver.last['info'] == 0
        assert np.linalg.norm(x - x_d) <= 1e-8 * np.linalg.norm(x_d)
        its[mode] = it
    assert its['sa'] < 0.6 * its['chebyshev'], its
    ml = F.AMGHierarchy(Kc, mode='sa')
    assert ml.levels >= 3 and 1.0 < ml.operator_complexity() < 2.0
    assert all(n % 3 == 0 for n in ml.level_sizes()[1:])
    with pytest.raises(ValueError):
        F.AMGHierarchy(F.DeviceCSR.from_scipy(Kc.tocsr()), mode='sa')          # no candidates attached
    ml2 = F.AMGHierarchy(F.DeviceCSR.from_scipy(Kc.tocsr()), mode='sa', candidates=B[I])
    assert ml2.level_sizes() == F.AMGHierarchy(Kc, mode='sa', reorder=False).level_sizes()
    # default coarsest level: up to 4000 rows, solved by one dense GEMV with the Cholesky inverse (built on the device)
    monkeypatch.delenv('MWX_SA_MAX_COARSE')
    ml3 = F.AMGHierarchy(Kc, mode='sa')
    assert ml3.levels < ml.levels and 300 < ml3.level_sizes()[-1] <= 4000
    x3, it3, info3, _ = F.pcg(Kc, bc, tol=1e-10, amg=ml3)
    assert info3 == 0 and it3 <= its['sa'] + 1 and np.linalg.norm(x3.cpu().numpy() - x_d) <= 1e-8 * np.linalg.norm(x_d)


def test_smoothed_aggregation_device_setup(cuda, monkeypatch):
    """mode='sa' builds its hierarchy on the device by default (csrc/amg_setup_dev.cu).  Checked against sparse algebra
    on the host: R = P^T exactly, A_c = R A P, the smoothed prolongator reproduces the candidates (P B_c = (I - w D^-1 A) B),
    aggregates cover every node; the setup is bit-reproducible; the solve matches the host-built hierarchy's."""
    import fem_forth_perturb_b200 as F
    monkeypatch.setenv('MWX_SA_MAX_COARSE', '300')
    gm = F.MeshTri().refine(6)
    basis = F.InteriorBasis(gm, F.ElementTriMorley())
    K = F.asm_gpu(1e-4 * F.a_load + F.b_load, basis)
    bf = gm.boundary_facets()
    D = np.concatenate((np.unique(gm.facets[:, bf]), bf + gm.nv))
    rng = np.random.default_rng(3)
    f = rng.standard_normal(K.shape[0])
    Kc, bc, _, I = F.condense(K, f, D=D)
    ml = F.AMGHierarchy(Kc, mode='sa')
    assert ml.setup == 'device' and ml.levels >= 3
    sizes = ml.level_sizes()
    assert all(n % 3 == 0 for n in sizes[1:]) and sizes[1] * 4 <= sizes[0]
    assert 1.0 < ml.operator_complexity() < 2.2
    for l in range(ml.levels - 1):
        A, P, R = ml.level_matrix(l, 0), ml.level_matrix(l, 1), ml.level_matrix(l, 2)
        An = ml.level_matrix(l + 1, 0)
        assert P.has_sorted_indices and R.has_sorted_indices and An.has_sorted_indices
        assert abs(R - P.T).max() == 0.0
        G = (R @ A @ P).tocsr()
        dead = np.flatnonzero(G.diagonal() == 0.0)                       # vanished candidates get a unit diagonal
        G = G + sp.csr_matrix((np.ones(dead.size), (dead, dead)), shape=G.shape)
        assert abs(An - G).max() <= 1e-12 * abs(G).max()
        B, Bc = ml.level_candidates(l), ml.level_candidates(l + 1)
        dinv = 1.0 / A.diagonal()
        # P = (I - w D^-1 A) T and T B_c = B, whatever w is: fit w from one entry, then compare everything
        lhs, AB = P @ Bc - B, dinv[:, None] * (A @ B)
        k = np.unravel_index(np.argmax(np.abs(AB)), AB.shape)
        w = -lhs[k] / AB[k]
        assert 0.3 < w < 1.4
        assert np.abs(lhs + w * AB).max() <= 1e-9 * max(1.0, np.abs(B).max())
        assert np.all(np.diff(P.indptr) % 3 == 0) and np.diff(P.indptr).min() >= 3      # every row belongs to an aggregate
    ml_again = F.AMGHierarchy(Kc, mode='sa')
    for l in range(ml.levels):
        a, b = ml.level_matrix(l, 0), ml_again.level_matrix(l, 0)
        assert np.array_equal(a.indptr, b.indptr) and np.array_equal(a.indices, b.indices) and np.array_equal(a.data, b.data)
    x_d = pl.direct_solver(Kc.tocsr(), bc)
    its = {}
    for setup in ('device', 'host'):
        mls = F.AMGHierarchy(Kc, mode='sa', setup=setup)
        x, it, info, _ = F.pcg(Kc, bc, tol=1e-10, amg=mls)
        assert info == 0 and np.linalg.norm(x.cpu().numpy() - x_d) <= 1e-8 * np.linalg.norm(x_d)
        its[setup] = it
    assert its['device'] <= 1.3 * its['host'] + 2, its
    with pytest.raises(ValueError):
        F.AMGHierarchy(Kc, mode='chebyshev', setup='device')


def test_block_csr_kernels_match_the_scalar_spmv(cuda, monkeypatch):
    """csrc/bsr.cu: the V-cycle of a device-built 'sa' hierarchy applies A_l, P_l, R_l through block-CSR kernels (3x3,
    1x3, 3x1 blocks).  Same hierarchy with MWX_BSR=0 (scalar k_spmv everywhere): the preconditioned vectors agree to
    round-off, with a plain V-cycle and with the gamma = 2 cycle."""
    import fem_forth_perturb_b200 as F
    torch = cuda
    gm = F.MeshTri().refine(7)
    basis = F.InteriorBasis(gm, F.ElementTriMorley())
    K = F.asm_gpu(1e-4 * F.a_load + F.b_load, basis)
    bf = gm.boundary_facets()
    D = np.concatenate((np.unique(gm.facets[:, bf]), bf + gm.nv))
    Kc = F.condense(K, D=D, expand=False)
    r = torch.from_numpy(np.random.default_rng(8).standard_normal(Kc.shape[0])).cuda()
    monkeypatch.setenv('MWX_SA_MAX_COARSE', '300')
    monkeypatch.setenv('MWX_PACKED_CSR', '1')      # fp32 storage: the scalar kernels read the same rounded values (packed records)
    for fp32 in ('0', '1'):                        # fp64-stored, then fp32-stored operators
        monkeypatch.setenv('MWX_SA_FP32', fp32)
        monkeypatch.setenv('MWX_BSR', '1')
        ml_b = F.AMGHierarchy(Kc, mode='sa')
        monkeypatch.setenv('MWX_BSR', '0')
        ml_s = F.AMGHierarchy(Kc, mode='sa')
        assert ml_b.levels >= 4 and ml_b.level_sizes() == ml_s.level_sizes()
        assert ml_b.storage_bits == ml_s.storage_bits == (32 if fp32 == '1' else 64)
        for gamma, k in ((1, 0), (2, 9)):
            ml_b.set_cycle(gamma, k); ml_s.set_cycle(gamma, k)
            zb, zs = ml_b.precondition(r), ml_s.precondition(r)
            assert float(torch.linalg.norm(zb - zs) / torch.linalg.norm(zs)) < 1e-12
    # the gamma = 2 cycle is still a symmetric operator (CG needs that)
    s2 = torch.from_numpy(np.random.default_rng(9).standard_normal(Kc.shape[0])).cuda()
    a, b = float(torch.dot(s2, ml_b.precondition(r))), float(torch.dot(r, ml_b.precondition(s2)))
    assert abs(a - b) <= 1e-10 * abs(a)
    b_rhs = Kc.dot(r)
    x2, it2, info2, _ = F.pcg(Kc, b_rhs, tol=1e-10, amg=ml_b)
    ml_b.set_cycle(1, 0)
    x1, it1, info1, _ = F.pcg(Kc, b_rhs, tol=1e-10, amg=ml_b)
    assert info1 == 0 and info2 == 0 and it2 < it1
    assert float(torch.linalg.norm(x1 - x2) / torch.linalg.norm(x1)) < 1e-7


def test_fp32_stored_hierarchy(cuda, monkeypatch):
    """mode='sa' stores the operators of its hierarchy in fp32 (values only; vectors and sums fp64).  The packed-record
    SpMV equals the fp64 SpMV of the ROUNDED matrix to round-off; the V-cycle stays a symmetric operator, differs from the
    fp64-stored one at the 1e-7 level only, and CG converges to the same solution in (nearly) the same iterations."""
    import ctypes
    import fem_forth_perturb_b200 as F
    from fem_forth_perturb_b200 import _lib as L
    torch = cuda
    gm = F.MeshTri().refine(7)
    basis = F.InteriorBasis(gm, F.ElementTriMorley())
    K = F.asm_gpu(1e-4 * F.a_load + F.b_load, basis)
    bf = gm.boundary_facets()
    D = np.concatenate((np.unique(gm.facets[:, bf]), bf + gm.nv))
    Kc = F.condense(K, D=D, expand=False)
    n = Kc.shape[0]
    rng = np.random.default_rng(21)
    x = torch.from_numpy(rng.standard_normal(n)).cuda()
    # ---- kernels: y = A32 x, c -+ A32 x, smoother step on packed (fp32 value, column) records
    pk = torch.empty(Kc.nnz, dtype=torch.int64, device='cuda')
    L.check(L.lib().mwx_csr_pack(Kc.nnz, L.ptr(Kc.d_indices), L.ptr(Kc.d_data), L.ptr(pk), L.stream()), 'mwx_csr_pack')
    A32 = sp.csr_matrix((Kc.d_data.cpu().numpy().astype(np.float32).astype(np.float64), Kc.d_indices.cpu().numpy(),
                         Kc.d_indptr.cpu().numpy()), shape=Kc.shape)
    y = torch.empty(n, dtype=torch.float64, device='cuda')
    L.check(L.lib().mwx_spmv_pk(n, Kc.spmv_hint(), L.ptr(Kc.d_indptr), L.ptr(pk), L.ptr(x), L.ptr(y), L.stream()), 'mwx_spmv_pk')
    ref = A32 @ x.cpu().numpy()
    assert np.abs(y.cpu().numpy() - ref).max() <= 1e-13 * np.abs(ref).max()
    assert np.abs(ref - Kc.tocsr() @ x.cpu().numpy()).max() > 1e-9 * np.abs(ref).max()        # it IS the rounded operator
    c = torch.from_numpy(rng.standard_normal(n)).cuda()
    for sign in (-1, 1):
        L.check(L.lib().mwx_spmv_axpy_pk(n, Kc.spmv_hint(), L.ptr(Kc.d_indptr), L.ptr(pk), L.ptr(x), L.ptr(c), sign, L.ptr(y),
                                         L.stream()), 'mwx_spmv_axpy_pk')
        assert np.abs(y.cpu().numpy() - (c.cpu().numpy() + sign * ref)).max() <= 1e-13 * np.abs(ref).max()
    dinv = torch.from_numpy(1.0 / Kc.tocsr().diagonal()).cuda()
    d = torch.from_numpy(rng.standard_normal(n)).cuda()
    d0 = d.cpu().numpy().copy()
    L.check(L.lib().mwx_smoother_step_pk(n, Kc.spmv_hint(), L.ptr(Kc.d_indptr), L.ptr(pk), L.ptr(x), L.ptr(y), L.ptr(c),
                                         L.ptr(dinv), L.ptr(d), 0.3, 0.7, L.stream()), 'mwx_smoother_step_pk')
    dn = 0.3 * d0 + 0.7 * dinv.cpu().numpy() * (c.cpu().numpy() - ref)
    assert np.abs(d.cpu().numpy() - dn).max() <= 1e-12 * np.abs(dn).max()
    assert np.abs(y.cpu().numpy() - (x.cpu().numpy() + dn)).max() <= 1e-12 * np.abs(dn).max()
    # ---- hierarchy: fp32-stored (default: block-CSR planes; here the packed twins of the scalar operators too) against fp64
    monkeypatch.setenv('MWX_PACKED_CSR', '1')
    ml32 = F.AMGHierarchy(Kc, mode='sa')
    monkeypatch.setenv('MWX_SA_FP32', '0')
    ml64 = F.AMGHierarchy(Kc, mode='sa')
    assert ml32.storage_bits == 32 and ml64.storage_bits == 64 and ml32.level_sizes() == ml64.level_sizes()
    r = torch.from_numpy(rng.standard_normal(n)).cuda()
    s2 = torch.from_numpy(rng.standard_normal(n)).cuda()
    z32, z64 = ml32.precondition(r), ml64.precondition(r)
    rel = float(torch.linalg.norm(z32 - z64) / torch.linalg.norm(z64))
    assert 0.0 < rel < 1e-5, rel
    a, b = float(torch.dot(s2, ml32.precondition(r))), float(torch.dot(r, ml32.precondition(s2)))
    assert abs(a - b) <= 1e-10 * abs(a)                                   # still symmetric: a valid CG preconditioner
    b_rhs = Kc.dot(x)
    x32, it32, info32, _ = F.pcg(Kc, b_rhs, tol=1e-10, amg=ml32)
    x64, it64, info64, _ = F.pcg(Kc, b_rhs, tol=1e-10, amg=ml64)
    assert info32 == 0 and info64 == 0 and abs(it32 - it64) <= 2, (it32, it64)
    assert float(torch.linalg.norm(x32 - x64) / torch.linalg.norm(x64)) < 1e-7
    assert float(torch.linalg.norm(x32 - x) / torch.linalg.norm(x)) < 1e-6


def test_tile_local_column_spmv(cuda):
    """csrc/spmv_lx.cu: per tile the distinct columns once + 2-byte local indices; the kernel gathers each distinct x entry
    once into shared memory.  Same results as the plain kernel (same products, same summation order: bit for bit), in
    every epilogue flavour, on the solver-numbered and on the skfem-numbered matrix."""
    import ctypes
    import fem_forth_perturb_b200 as F
    from fem_forth_perturb_b200 import _lib as L
    from fem_forth_perturb_b200.solvers import Reordering
    torch = cuda
    gm = F.MeshTri().refine(7)
    basis = F.InteriorBasis(gm, F.ElementTriMorley())
    K = F.asm_gpu(1e-4 * F.a_load + F.b_load, basis)
    bf = gm.boundary_facets()
    D = np.concatenate((np.unique(gm.facets[:, bf]), bf + gm.nv))
    Kc = F.condense(K, D=D, expand=False)
    rng = np.random.default_rng(31)
    lib, st = L.lib(), L.stream()
    built = 0
    for A in (Reordering.for_matrix(Kc).matrix(Kc), Kc):
        n, plan = A.shape[0], A.spmv_hint()
        ip = A.d_indptr if A.d_indptr.dtype == torch.int64 else A.d_indptr.to(torch.int64)
        h = ctypes.c_void_p()
        L.check(lib.mwx_spmv_lx_create(n, L.ptr(ip), L.ptr(A.d_indices), plan, ctypes.byref(h), st), 'mwx_spmv_lx_create')
        if not h.value:
            continue                                   # pattern does not localise (allowed for the skfem numbering)
        built += 1
        x = torch.from_numpy(rng.standard_normal(n)).cuda()
        c = torch.from_numpy(rng.standard_normal(n)).cuda()
        y0, y1 = torch.empty_like(x), torch.empty_like(x)
        args = (n, plan, L.ptr(ip), L.ptr(A.d_indices), L.ptr(A.d_data))
        L.check(lib.mwx_spmv(*args, L.ptr(x), L.ptr(y0), st), 'mwx_spmv')
        L.check(lib.mwx_spmv_lx(*args, h, L.ptr(x), L.ptr(y1), st), 'mwx_spmv_lx')
        ref = A.tocsr() @ x.cpu().numpy()
        assert np.abs(y0.cpu().numpy() - ref).max() <= 1e-13 * np.abs(ref).max()
        assert torch.equal(y0, y1)
        for sign in (-1, 1):
            L.check(lib.mwx_spmv_axpy(*args, L.ptr(x), L.ptr(c), sign, L.ptr(y0), st), 'mwx_spmv_axpy')
            L.check(lib.mwx_spmv_axpy_lx(*args, h, L.ptr(x), L.ptr(c), sign, L.ptr(y1), st), 'mwx_spmv_axpy_lx')
            assert torch.equal(y0, y1)
        dinv = torch.from_numpy(1.0 / A.tocsr().diagonal()).cuda()
        d0 = torch.from_numpy(rng.standard_normal(n)).cuda()
        d1 = d0.clone()
        L.check(lib.mwx_smoother_step(*args, L.ptr(x), L.ptr(y0), L.ptr(c), L.ptr(dinv), L.ptr(d0), 0.3, 0.7, st), 'step')
        L.check(lib.mwx_smoother_step_lx(*args, h, L.ptr(x), L.ptr(y1), L.ptr(c), L.ptr(dinv), L.ptr(d1), 0.3, 0.7, st), 'step_lx')
        assert torch.equal(y0, y1) and torch.equal(d0, d1)
        ws0 = torch.zeros(lib.mwx_dcg_workspace_doubles(), dtype=torch.float64, device='cuda')
        ws1 = torch.zeros_like(ws0)
        L.check(lib.mwx_dcg_spmv_dot(*args, L.ptr(x), L.ptr(y0), L.ptr(ws0), st), 'spmv_dot')
        L.check(lib.mwx_dcg_spmv_dot_lx(*args, h, L.ptr(x), L.ptr(y1), L.ptr(ws1), st), 'spmv_dot_lx')
        assert torch.equal(y0, y1) and float(ws0[2]) == float(ws1[2])
        assert abs(float(ws0[2]) - float(x.cpu().numpy() @ ref)) <= 1e-11 * abs(float(ws0[2]))
        lib.mwx_spmv_lx_destroy(h)
    assert built >= 1                                  # the Hilbert-ordered matrix must localise
    # through the solver: mode 'sa' with and without the plans gives the same iterates
    import os
    b = Kc.dot(torch.from_numpy(rng.standard_normal(Kc.shape[0])).cuda())
    x0, it0, info0, _ = F.pcg(Kc, b, tol=1e-10, amg=F.AMGHierarchy(Kc, mode='sa'))
    os.environ['MWX_SPMV_LX'] = '1'                    # off by default (measured slower than the plain kernel, amg.cu)
    try:
        x1, it1, info1, _ = F.pcg(Kc, b, tol=1e-10, amg=F.AMGHierarchy(Kc, mode='sa'))
    finally:
        del os.environ['MWX_SPMV_LX']
    assert info0 == 0 and info1 == 0 and it0 == it1 and torch.equal(x0, x1)


def test_double_double_residual_and_refinement(cuda):
    """``mwx_residual_dd``: b - A x with double-double row accumulation - checked against exact rational arithmetic on a few
    rows and against the fp64 residual's error; ``refine``: on an ill-conditioned system the forward error of a tol-1e-8
    PCG solution drops by orders of magnitude (the plain solve stalls at cond * tol)."""
    from fractions import Fraction
    import fem_forth_perturb_b200 as F
    torch = cuda
    gm = F.MeshTri().refine(6)
    basis = F.InteriorBasis(gm, F.ElementTriMorley())
    K = F.asm_gpu(1.0 * F.a_load + F.b_load, basis)               # eps = 1: the fourth-order term dominates, cond ~ h^-4
    bf = gm.boundary_facets()
    D = np.concatenate((np.unique(gm.facets[:, bf]), bf + gm.nv))
    Kc = F.condense(K, D=D, expand=False)
    A = Kc.tocsr()
    n = A.shape[0]
    rng = np.random.default_rng(41)
    xs = rng.standard_normal(n)
    b = A @ xs
    x_t, b_t = torch.from_numpy(xs).cuda(), torch.from_numpy(b).cuda()
    r_dd = F.residual_dd(Kc, x_t, b_t).cpu().numpy()
    rows = rng.choice(n, 40, replace=False)
    exact = np.array([float(Fraction(b[i]) - sum(Fraction(A.data[k]) * Fraction(xs[A.indices[k]])
                                                   for k in range(A.indptr[i], A.indptr[i + 1]))) for i in rows])
    scale = np.abs(A[rows]).dot(np.abs(xs)) + np.abs(b[rows])
    assert np.abs(r_dd[rows] - exact).max() <= 1e-30 * scale.max() + 2.3e-16 * np.abs(exact).max()      # one final rounding
    r64 = b - A @ xs
    assert np.abs(r64[rows] - exact).max() > 10 * np.abs(r_dd[rows] - exact).max()                       # fp64 is visibly worse
    # refinement: b is exactly representable data; the system's true solution is x* up to cond * eps
    ml = F.AMGHierarchy(Kc, mode='sa')
    x0, it0, info0, _ = F.pcg(Kc, b_t, tol=1e-8, amg=ml)
    log = []
    x1 = F.refine(Kc, b_t, x0, steps=2, tol=1e-8, amg=ml, log=log)
    e0 = float(torch.linalg.norm(x0 - x_t) / torch.linalg.norm(x_t))
    e1 = float(torch.linalg.norm(x1 - x_t) / torch.linalg.norm(x_t))
    assert info0 == 0 and len(log) == 3 and log[-1]['rel_residual_dd'] < 1e-3 * log[0]['rel_residual_dd'], log
    assert e1 < 1e-2 * e0, (e0, e1)


def test_coarse_inverse_of_the_sa_hierarchy(cuda):
    """AMGHierarchy._pinv_coarse above 300 rows (device, setup only): Cholesky inverse of an SPD coarse operator,
    pseudo-inverse with pinv2's cutoff when the factorisation breaks down."""
    import fem_forth_perturb_b200 as F
    rng = np.random.default_rng(0)
    B = rng.standard_normal((700, 700))
    A = B @ B.T + 700 * np.eye(700)
    P = F.AMGHierarchy._pinv_coarse(A)
    assert abs(P @ A - np.eye(700)).max() <= 1e-11 and abs(P - P.T).max() <= 1e-13 * abs(P).max()
    B = rng.standard_normal((400, 395))
    A = B @ B.T
    P = F.AMGHierarchy._pinv_coarse(A)
    assert abs(A @ P @ A - A).max() <= 1e-9 * abs(A).max()


def test_locality_reordering_is_a_similarity_transform(cuda):
    """Fast-mode solvers run on P A P^T (Hilbert order of the DOF locations) and permute back."""
    import fem_forth_perturb_b200 as F
    from fem_forth_perturb_b200.solvers import Reordering
    gm = F.MeshTri().refine(5)
    basis = F.InteriorBasis(gm, F.ElementTriMorley())
    K = F.asm_gpu(1e-4 * F.a_load + F.b_load, basis)
    bf = gm.boundary_facets()
    D = np.concatenate((np.unique(gm.facets[:, bf]), bf + gm.nv))
    rng = np.random.default_rng(5)
    f = rng.standard_normal(K.shape[0])
    Kc, bc, _, I = F.condense(K, f, D=D)
    assert Kc.coords is not None
    ro = Reordering.for_matrix(Kc)
    perm = ro.perm.cpu().numpy()
    assert sorted(perm.tolist()) == list(range(Kc.shape[0]))
    assert np.array_equal(ro.iperm.cpu().numpy()[perm], np.arange(Kc.shape[0]))
    Ks = Kc.tocsr()
    Kp = ro.matrix(Kc).tocsr()
    ref = Ks[perm][:, perm].tocsr(); ref.sort_indices()
    assert np.array_equal(Kp.indptr, ref.indptr) and np.array_equal(Kp.indices, ref.indices)
    assert np.array_equal(Kp.data, ref.data)                      # pure data movement: bit-exact
    # later calls on the same pattern reuse the permuted pattern and only gather the values (cached source map)
    Kc2 = 2.0 * Kc
    for scratch in (False, True, False):
        Kp2 = ro.matrix(Kc2, scratch=scratch).tocsr()
        assert np.array_equal(Kp2.indptr, ref.indptr) and np.array_equal(Kp2.indices, ref.indices)
        assert np.array_equal(Kp2.data, 2.0 * ref.data)
    # ... and a CondensePlan applied repeatedly to re-assembled matrices does the same
    plan = F.CondensePlan(K, D=D)
    for scale in (1.0, 3.0, 0.5):
        Kci, none = plan.apply(scale * K)
        Kci = Kci.tocsr()
        assert none is None and np.array_equal(Kci.indptr, Ks.indptr) and np.array_equal(Kci.indices, Ks.indices)
        assert np.array_equal(Kci.data, scale * Ks.data)
    Kcb, bcb = plan.apply(K, f, np.zeros(K.shape[0]))             # with a load vector: the full pass
    assert np.array_equal(Kcb.tocsr().data, Ks.data) and np.allclose(bcb.cpu().numpy(), bc, rtol=0, atol=0)
    # locality: consecutive DOFs along the curve are geometric neighbours (mean jump ~ h, not ~ 1)
    xy = Kc.coords().cpu().numpy()[perm]
    jump = np.linalg.norm(np.diff(xy, axis=0), axis=1)
    assert np.median(jump) < 4 * 2.0 ** -5 and jump.mean() < np.linalg.norm(np.diff(Kc.coords().cpu().numpy(), axis=0), axis=1).mean()
    # solves agree with the un-reordered ones
    x_d = pl.direct_solver(Ks, bc)
    for solver in (F.solver_iter_mgcg(tol=1e-12, mode='chebyshev'), F.solver_iter_mgcg(tol=1e-12, mode='jacobi'),
                   F.solver_iter_krylov(Precondition=True, tol=1e-12, reorder=True)):
        x, it = solver.solve_with_info(Kc, bc)
        assert solver.last['info'] == 0
        assert np.linalg.norm(x - x_d) <= 1e-10 * np.linalg.norm(x_d)
    with pytest.raises(ValueError):
        F.AMGHierarchy(Kc, mode='parity', reorder=True)


def test_solver_iter_pyamg_standalone_vcycles(cuda):
    """solver_type='amg' (ref utils.py:168-210): same number of V-cycles as the pyamg restatement's ml.solve."""
    import fem_forth_perturb_b200 as F
    om = OMesh().refine(4)
    S = pl.solve_problem(om, pl.Ex1(1e-2), wkind='P1', intorder=5, return_system=True)
    Kc, bc = S['Kc'], S['bc']
    Kc.sort_indices()
    x_o, it_o = RugeStuben(Kc).solve(bc, tol=1e-10, maxiter=10000)
    solver = F.solver_iter_pyamg()
    x, it = solver.solve_with_info(Kc, bc)
    assert abs(it - it_o) <= 1, (it, it_o)
    assert np.linalg.norm(x - x_o) <= 1e-8 * np.linalg.norm(x_o)


def test_mgcg_iter_prints_like_the_reference(cuda, capsys):
    import fem_forth_perturb_b200 as F
    om = OMesh().refine(2)
    S = pl.solve_problem(om, pl.Ex1(1.0), wkind='P1', intorder=5, return_system=True)
    F.solver_iter_mgcg_iter(tol=1e-8, maxiter=1000)(S['Kc'], S['bc'])
    assert 'mgcg total interation steps: 6' in capsys.readouterr().out


# ------------------------------------------------------------------ end to end against the reference's tables
@pytest.mark.parametrize('key,levels', [('ex1_P1_nopen', 5), ('ex3_P1_pen', 4)])
def test_golden_error_tables_through_the_gpu_path(cuda, golden_norms, key, levels):
    """The reference's pipeline with K2 assembled and both systems solved on the GPU; everything else
    (w-problem forms, norms) is the oracle.  Printed tables must agree to 3 significant digits."""
    import fem_forth_perturb_b200 as F
    g = golden_norms[key]
    cfg = g['config']
    prob_cls = {'ex1': pl.Ex1, 'ex3': pl.Ex3}[cfg['example']]

    def k2_hook(om, eps, penalty, sigma):
        gm = F.MeshTri(om.p, om.t)
        ib = F.InteriorBasis(gm, F.ElementTriMorley(), intorder=cfg['intorder'])
        K = F.asm_gpu(eps ** 2 * F.a_load + F.b_load, ib)
        if penalty:
            fb = F.FacetBasis(gm, F.ElementTriMorley()); fb.interior = ib
            K = K + eps ** 2 * F.asm_gpu(F.penalty_1 + F.penalty_2 + F.penalty_3, fb, sigma=sigma)
        return K.tocsr()

    for eps_s in list(g['rows'])[:3]:
        eps = float(eps_s)
        om = OMesh()
        for level in range(1, levels + 1):
            om.refine()
            prob = prob_cls(eps)
            uh, ub, fb = pl.solve_problem(om, prob, wkind=cfg['w'], intorder=cfg['intorder'],
                                          penalty=cfg['penalty'], sigma=cfg.get('sigma', 5), K2_hook=k2_hook,
                                          solver_w=F.solver_iter_mgcg(tol=cfg['tol']),
                                          solver_u=F.solver_iter_mgcg(tol=cfg['tol']))
            mine = np.array(pl.error_norms(uh, ub, prob, fb))
            gold = np.array(g['rows'][eps_s][level - 1])
            assert np.max(np.abs(mine - gold) / gold) < 5e-4, (key, eps_s, level, mine, gold)
