"""Oracle (test infrastructure): restatement of scikit-fem 2.1.1 bases.

Stands in for ``InteriorBasis(m, ElementTriMorley()/P1/P2, intorder=...)`` and
``FacetBasis(m, ElementTriMorley())`` as called at ref main/example1/run.py:193-196,253
(scikit-fem is un-vendored, pin ref README.md:17; behaviour per SURVEY.md A.2-A.6).

The Morley tabulation deliberately follows skfem's ``ElementGlobal`` route - a
per-element 6x6 Vandermonde in GLOBAL monomials 1,y,y^2,x,xy,x^2, inverted with
LAPACK - because that is what the reference's numbers (incl. their round-off) come
from.  ``shift=True`` gives the well-conditioned centroid-shifted variant used to
measure skfem's own round-off.
"""
import itertools
import numpy as np


# ------------------------------------------------------------------ quadrature (A.6)
def quadrature_tri(intorder):
    """Reference-triangle rule skfem picks for ``intorder`` -> (X (2,nq), W (nq,)), sum W = 1/2."""
    if intorder <= 1:
        pts, w = [(1 / 3, 1 / 3)], [1.0]
    elif intorder == 2:
        pts, w = [(1 / 6, 1 / 6), (2 / 3, 1 / 6), (1 / 6, 2 / 3)], [1 / 3] * 3
    elif intorder in (3, 4):
        pts, w = [], []
        for wt, a, b in [(0.223381589678011, 0.108103018168070, 0.445948490915965),
                         (0.109951743655322, 0.816847572980459, 0.091576213509771)]:
            for P in [(a, b), (b, a), (b, b)]:
                pts.append(P); w.append(wt)
    elif intorder == 5:
        s = np.sqrt(15.0)
        a1, a2 = (6 - s) / 21, (6 + s) / 21
        w1, w2 = (155 - s) / 1200, (155 + s) / 1200
        pts = [(1 / 3, 1 / 3), (a1, a1), (1 - 2 * a1, a1), (a1, 1 - 2 * a1),
               (a2, a2), (1 - 2 * a2, a2), (a2, 1 - 2 * a2)]
        w = [9 / 40, w1, w1, w1, w2, w2, w2]
    elif intorder == 6:
        pts, w = [], []
        for wt, a, b in [(0.116786275726379, 0.501426509658179, 0.249286745170910),
                         (0.050844906370207, 0.873821971016996, 0.063089014491502)]:
            for P in [(a, b), (b, a), (b, b)]:
                pts.append(P); w.append(wt)
        for P in itertools.permutations((0.053145049844817, 0.310352451033784, 0.636502499121399)):
            pts.append(P[:2]); w.append(0.082851075618374)
    elif intorder in (7, 8):
        pts, w = [(1 / 3, 1 / 3)], [0.144315607677787]
        for wt, a, b in [(0.095091634267285, 0.081414823414554, 0.459292588292723),
                         (0.103217370534718, 0.658861384496480, 0.170569307751760),
                         (0.032458497623198, 0.898905543365938, 0.050547228317031)]:
            for P in [(a, b), (b, a), (b, b)]:
                pts.append(P); w.append(wt)
        for P in itertools.permutations((0.008394777409958, 0.263112829634638, 0.728492392955404)):
            pts.append(P[:2]); w.append(0.027230314174435)
    else:
        raise NotImplementedError(f"intorder {intorder}")
    return np.array(pts, dtype=np.float64).T, 0.5 * np.array(w, dtype=np.float64)


def quadrature_line(intorder):
    n = int(np.ceil((intorder + 1) / 2.0))
    x, w = np.polynomial.legendre.leggauss(n)
    return 0.5 * x + 0.5, 0.5 * w


# ------------------------------------------------------------------ affine map
def affine(mesh):
    p, t = mesh.p, mesh.t
    A = np.empty((2, 2, t.shape[1]))
    A[:, 0, :] = p[:, t[1]] - p[:, t[0]]
    A[:, 1, :] = p[:, t[2]] - p[:, t[0]]
    det = A[0, 0] * A[1, 1] - A[0, 1] * A[1, 0]
    invA = np.empty_like(A)
    invA[0, 0] = A[1, 1] / det
    invA[0, 1] = -A[0, 1] / det
    invA[1, 0] = -A[1, 0] / det
    invA[1, 1] = A[0, 0] / det
    return A, p[:, t[0]], det, invA


def global_points(mesh, X):
    A, b, _, _ = affine(mesh)
    return np.einsum('ijk,jl->ikl', A, X) + b[:, :, None]       # (2, nt, nq)


# ------------------------------------------------------------------ Morley, skfem-style (A.3)
_POWERS = [(0, 0), (0, 1), (0, 2), (1, 0), (1, 1), (2, 0)]       # 1, y, y^2, x, xy, x^2


def _mono(i, j, dx, dy):
    cx = cy = 1
    for l in range(dx):
        cx *= i - l
    for l in range(dy):
        cy *= j - l
    c = cx * cy
    ex, ey = max(i - dx, 0), max(j - dy, 0)
    return lambda x, y: c * x ** ex * y ** ey


def morley_vandermonde(mesh, shift=False):
    """V[e, dof, mono] and the three global edge normals n_k = R(x_a-x_b)/|.|, a<b."""
    p, t = mesh.p, mesh.t
    nt = t.shape[1]
    v = np.array([p[:, t[i]] for i in range(3)])                  # (3, 2, nt)
    if shift:
        v = v - v.mean(axis=0)[None]
    ends = [(0, 1), (1, 2), (0, 2)]
    mid = np.array([0.5 * (v[a] + v[b]) for a, b in ends])
    nrm = np.empty((3, 2, nt))
    for k, (a, b) in enumerate(ends):
        d = v[a] - v[b]
        r = np.array([-d[1], d[0]])
        nrm[k] = r / np.sqrt(np.sum(r * r, axis=0))
    V = np.zeros((nt, 6, 6))
    for m_, (i, j) in enumerate(_POWERS):
        f, fx, fy = _mono(i, j, 0, 0), _mono(i, j, 1, 0), _mono(i, j, 0, 1)
        for d in range(3):
            V[:, d, m_] = f(*v[d])
            V[:, 3 + d, m_] = fx(*mid[d]) * nrm[d, 0] + fy(*mid[d]) * nrm[d, 1]
    return V, nrm


def morley_tabulate(mesh, x, tind=None, shift=False):
    """Six (value, grad, hess) triples at global points x (2, n, nq) of elements ``tind``."""
    V, _ = morley_vandermonde(mesh, shift)
    Vi = np.linalg.inv(V)
    if shift:
        c = np.array([mesh.p[:, mesh.t[i]] for i in range(3)]).mean(axis=0)
        if tind is not None:
            c = c[:, tind]
        x = x - c[:, :, None]
    if tind is not None:
        Vi = Vi[tind]
    out = []
    for jb in range(6):
        val = np.zeros(x[0].shape)
        g = np.zeros((2,) + x[0].shape)
        h = np.zeros((2, 2) + x[0].shape)
        for m_, (i, j) in enumerate(_POWERS):
            c = Vi[:, m_, jb][:, None]
            val += c * _mono(i, j, 0, 0)(*x)
            g[0] += c * _mono(i, j, 1, 0)(*x)
            g[1] += c * _mono(i, j, 0, 1)(*x)
            h[0, 0] += c * _mono(i, j, 2, 0)(*x)
            h[0, 1] += c * _mono(i, j, 1, 1)(*x)
            h[1, 0] += c * _mono(i, j, 1, 1)(*x)
            h[1, 1] += c * _mono(i, j, 0, 2)(*x)
        out.append((val, g, h))
    return out


# ------------------------------------------------------------------ Lagrange P1 / P2 (A.3c)
def _lagrange_tabulate(mesh, X, kind):
    _, _, _, invA = affine(mesh)
    nt = mesh.t.shape[1]
    x, y = X
    o, z = np.ones_like(x), np.zeros_like(x)
    if kind == 'P1':
        ref = [(1 - x - y, np.array([-o, -o])), (x, np.array([o, z])), (y, np.array([z, o]))]
    else:
        ref = [(1 - 3 * x - 3 * y + 2 * x ** 2 + 4 * x * y + 2 * y ** 2,
                np.array([-3 + 4 * x + 4 * y, -3 + 4 * x + 4 * y])),
               (2 * x ** 2 - x, np.array([4 * x - 1, z])),
               (2 * y ** 2 - y, np.array([z, 4 * y - 1])),
               (4 * x - 4 * x ** 2 - 4 * x * y, np.array([4 - 8 * x - 4 * y, -4 * x])),
               (4 * x * y, np.array([4 * y, 4 * x])),
               (4 * y - 4 * x * y - 4 * y ** 2, np.array([-4 * y, 4 - 4 * x - 8 * y]))]
    return [(np.tile(phi, (nt, 1)), np.einsum('ijk,il->jkl', invA, dphi), None) for phi, dphi in ref]


def element_dofs(mesh, kind):
    """A.2: nodal DOFs = vertex ids, facet DOFs = nv + facet id (same for P2 and Morley)."""
    if kind == 'P1':
        return mesh.t.copy()
    return np.vstack((mesh.t, mesh.t2f + mesh.nv))


def num_dofs(mesh, kind):
    return mesh.nv + (0 if kind == 'P1' else mesh.nf)


class InteriorBasis:
    """``InteriorBasis(m, element, intorder)`` (ref main/example1/run.py:193-196)."""

    def __init__(self, mesh, kind, intorder, shift=False):
        assert kind in ('P1', 'P2', 'Morley')
        self.mesh, self.kind, self.intorder = mesh, kind, intorder
        self.X, self.W = quadrature_tri(intorder)
        _, _, det, _ = affine(mesh)
        self.dx = np.abs(det)[:, None] * self.W[None, :]
        self.x = global_points(mesh, self.X)
        self.element_dofs = element_dofs(mesh, kind)
        self.N = num_dofs(mesh, kind)
        self.nelems = mesh.nt
        if kind == 'Morley':
            self.phi = morley_tabulate(mesh, self.x, shift=shift)
        else:
            self.phi = _lagrange_tabulate(mesh, self.X, kind)
        self.Nbfun = len(self.phi)

    def find_dofs(self):
        """``basis.find_dofs()`` with no argument: all boundary nodal (+ facet) DOFs (A.2)."""
        m = self.mesh
        bf = m.boundary_facets()
        bv = np.unique(m.facets[:, bf])
        if self.kind == 'P1':
            return bv
        return np.concatenate((bv, bf + m.nv))

    def doflocs(self):
        m = self.mesh
        if self.kind == 'P1':
            return m.p
        return np.hstack((m.p, 0.5 * (m.p[:, m.facets[0]] + m.p[:, m.facets[1]])))

    def interpolate(self, u):
        ed = self.element_dofs
        val = sum(u[ed[i]][:, None] * self.phi[i][0] for i in range(self.Nbfun))
        grad = sum(u[ed[i]][None, :, None] * self.phi[i][1] for i in range(self.Nbfun))
        hess = None
        if self.phi[0][2] is not None:
            hess = sum(u[ed[i]][None, None, :, None] * self.phi[i][2] for i in range(self.Nbfun))
        return val, grad, hess


class FacetBasis:
    """``FacetBasis(m, ElementTriMorley())`` on the boundary facets (A.3b; ref run.py:253)."""

    def __init__(self, mesh, kind='Morley', intorder=4, shift=False):
        assert kind == 'Morley'
        self.mesh = mesh
        gx, gw = quadrature_line(intorder)
        self.find = mesh.boundary_facets()
        self.tind = mesh.f2t()[0, self.find]
        p0 = mesh.p[:, mesh.facets[0, self.find]]
        p1 = mesh.p[:, mesh.facets[1, self.find]]
        tan = p1 - p0
        length = np.sqrt(np.sum(tan * tan, axis=0))
        self.x = p0[:, :, None] + tan[:, :, None] * gx[None, None, :]
        nrm = np.array([tan[1], -tan[0]]) / length
        cen = mesh.p[:, mesh.t[:, self.tind]].mean(axis=1)
        sgn = np.sign(np.sum(nrm * (0.5 * (p0 + p1) - cen), axis=0))
        nq = len(gx)
        self.n = np.repeat((nrm * sgn)[:, :, None], nq, axis=2)      # outward unit normal
        self.h = np.repeat(length[:, None], nq, axis=1)             # facet length
        self.dx = length[:, None] * gw[None, :]
        self.element_dofs = element_dofs(mesh, kind)[:, self.tind]
        self.N = num_dofs(mesh, kind)
        self.phi = morley_tabulate(mesh, self.x, tind=self.tind, shift=shift)
        self.Nbfun = 6

    def interpolate(self, u):
        ed = self.element_dofs
        val = sum(u[ed[i]][:, None] * self.phi[i][0] for i in range(6))
        grad = sum(u[ed[i]][None, :, None] * self.phi[i][1] for i in range(6))
        hess = sum(u[ed[i]][None, None, :, None] * self.phi[i][2] for i in range(6))
        return val, grad, hess
