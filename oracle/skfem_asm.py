"""Oracle (test infrastructure): restatement of scikit-fem 2.1.1 ``asm`` + the reference's forms.

``asm_bilinear`` follows skfem's ``BilinearForm.assemble``: block (j, i)
of the COO triplets holds ``sum_q form(u_j, v_i) dx`` with rows = test DOFs, cols = trial
DOFs; then ``eliminate_zeros()`` on the COO and ``tocsr()`` (canonical, duplicates summed).
Call sites: ref main/example1/run.py:198-199,210-211,255-257,260-261.

The forms restate ref main/example1/run.py:110-154 (a_load, b_load, wv_load, penalty_1/2/3,
laplace) on the (value, grad, hess) triples of ``oracle.skfem_basis``.
``condense``/``solve`` restate ref main/example1/utils.py:326-368,384-460.
"""
import numpy as np
import scipy.sparse as sp


# ------------------------------------------------------------------ forms (ref run.py:110-154)
def a_load(u, v, w=None):
    return np.einsum('ij...,ij...', u[2], v[2])          # ddot(dd(u), dd(v))


def b_load(u, v, w=None):
    return np.einsum('i...,i...', u[1], v[1])            # dot(grad(u), grad(v))


laplace = b_load
wv_load = b_load


def mass(u, v, w=None):
    return u[0] * v[0]


def make_penalty_forms(fbasis, sigma):
    """penalty_1/2/3 of ref main/example1/run.py:134-146 bound to a FacetBasis (w.n, w.h)."""
    n, h = fbasis.n, fbasis.h
    nn = np.einsum('i...,j...->ij...', n, n)             # prod(w.n, w.n)

    def penalty_1(u, v, w=None):
        return np.einsum('ij...,ij...', -u[2], nn) * np.einsum('i...,i...', v[1], n)

    def penalty_2(u, v, w=None):
        return np.einsum('ij...,ij...', -v[2], nn) * np.einsum('i...,i...', u[1], n)

    def penalty_3(u, v, w=None):
        return (sigma / h) * np.einsum('i...,i...', u[1], n) * np.einsum('i...,i...', v[1], n)

    return penalty_1, penalty_2, penalty_3


# ------------------------------------------------------------------ asm (A.4)
def asm_bilinear(form, ubasis, vbasis=None, eliminate=True, return_coo=False):
    vbasis = ubasis if vbasis is None else vbasis
    nt = ubasis.dx.shape[0]
    nu, nv = ubasis.Nbfun, vbasis.Nbfun
    data = np.zeros(nu * nv * nt)
    rows = np.zeros(nu * nv * nt, dtype=np.int64)
    cols = np.zeros(nu * nv * nt, dtype=np.int64)
    for j in range(nu):
        for i in range(nv):
            ixs = slice(nt * (nv * j + i), nt * (nv * j + i + 1))
            rows[ixs] = vbasis.element_dofs[i]
            cols[ixs] = ubasis.element_dofs[j]
            data[ixs] = np.sum(form(ubasis.phi[j], vbasis.phi[i]) * ubasis.dx, axis=1)
    if return_coo:
        return data, rows, cols
    K = sp.coo_matrix((data, (rows, cols)), shape=(vbasis.N, ubasis.N))
    if eliminate:
        K.eliminate_zeros()
    return K.tocsr()


def asm_linear(fun, basis):
    """``asm(LinearForm, basis)`` with integrand ``fun(x, y) * v`` (ref run.py:277-288,199)."""
    x = basis.x
    fx = fun(x[0], x[1]) * np.ones_like(x[0])
    out = np.zeros(basis.N)
    for i in range(basis.Nbfun):
        np.add.at(out, basis.element_dofs[i], np.sum(fx * basis.phi[i][0] * basis.dx, axis=1))
    return out


def local_matrices(form, basis):
    """[i, j, e] local matrices of ``form`` (rows = test i, cols = trial j)."""
    return np.array([[np.sum(form(basis.phi[j], basis.phi[i]) * basis.dx, axis=1)
                      for j in range(basis.Nbfun)] for i in range(basis.Nbfun)])


# ------------------------------------------------------------------ boundary DOFs (ref run.py:67-104)
def _four_sides(mesh):
    sides = [lambda x: x[0] == 0, lambda x: x[0] == 1, lambda x: x[1] == 1, lambda x: x[1] == 0]
    return [mesh.facets_satisfying(s) for s in sides]


def easy_boundary(basis):
    """nodal 'u' + facet 'u_n' DOFs of the four sides, concatenated WITH corner duplicates."""
    m = basis.mesh
    fs = _four_sides(m)
    nodal = [np.unique(m.facets[:, f]) for f in fs]
    facet = [f + m.nv for f in fs]
    return np.concatenate(nodal + facet)


def easy_boundary_penalty(basis):
    m = basis.mesh
    return np.concatenate([np.unique(m.facets[:, f]) for f in _four_sides(m)])


# ------------------------------------------------------------------ condense / solve (ref utils.py:326-460)
def condense(A, b, x=None, D=None):
    if x is None:
        x = np.zeros(A.shape[0])
    I = np.setdiff1d(np.arange(A.shape[0]), D)
    Aout = A[I].T[I].T
    bout = b[I] - A[I].T[D].T @ x[D]
    return Aout.tocsr(), bout, x, I


def solve(A, b, x, I, solver):
    y = x.copy()
    y[I] = solver(A, b)
    return y
