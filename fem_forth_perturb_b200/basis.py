"""Host mirrors of the skfem objects the MWX path touches: elements and bases.

``InteriorBasis(m, ElementTriMorley(), intorder=...)`` / ``FacetBasis(m, ElementTriMorley())``
(ref main/example1/run.py:193-196,253).  For the Morley forms the device kernel integrates in closed
form, so a basis here carries no tabulated values: it owns the DOF numbering (skfem ``Dofs``,
SURVEY A.2), the DOF->element incidence, and the cached CSR pattern + slot map.
"""
import ctypes

import numpy as np

from . import _lib as L


class ElementTriMorley:
    nodal_dofs, facet_dofs, dofnames = 1, 1, ['u', 'u_n']


class ElementTriP1:
    nodal_dofs, facet_dofs, dofnames = 1, 0, ['u']


class ElementTriP2:
    nodal_dofs, facet_dofs, dofnames = 1, 1, ['u', 'u']


class DofsView:
    """The slice of ``basis.find_dofs(...)`` the reference reads: ``.nodal['u']``, ``.facet['u_n']``."""

    def __init__(self, nodal, facet, names):
        self.nodal = {names[0]: nodal}
        self.facet = {names[-1]: facet} if facet is not None else {}
        self._all = np.concatenate([nodal] + ([facet] if facet is not None else []))

    def all(self):
        return self._all

    def flatten(self):
        return self._all


class InteriorBasis:
    def __init__(self, mesh, elem, intorder=None):
        self.mesh, self.elem, self.intorder = mesh, elem, intorder
        self.nelems = mesh.nt
        self.N = mesh.nv * elem.nodal_dofs + mesh.nf * elem.facet_dofs
        self.Nbfun = 3 * elem.nodal_dofs + 3 * elem.facet_dofs
        self._edofs = None
        self._pattern = None
        self._host_edofs = None

    # -- DOF numbering (device) --------------------------------------------------------------------
    @property
    def d_element_dofs(self):
        if self._edofs is None:
            torch = L.require_cuda()
            m = self.mesh
            if self.elem.facet_dofs:
                ed = torch.empty(6 * m.nt, dtype=torch.int32, device=m.device)
                L.check(L.lib().mwx_element_dofs(m.nv, m.nt, L.ptr(m.d_t), L.ptr(m.d_t2f), L.ptr(ed), L.stream()),
                        'mwx_element_dofs')
            else:
                ed = m.d_t
            self._edofs = ed
        return self._edofs

    @property
    def element_dofs(self):
        if self._host_edofs is None:
            self._host_edofs = self.d_element_dofs.reshape(self.Nbfun, self.nelems).cpu().numpy().astype(np.int64)
        return self._host_edofs

    # -- CSR pattern + slot map (cached; setup cost, timed separately by bench.py) ------------------
    def pattern(self):
        if self._pattern is None:
            torch = L.require_cuda()
            lib, st, m = L.lib(), L.stream(), self.mesh
            ed, nb = self.d_element_dofs, self.Nbfun
            r2e_ptr = torch.empty(self.N + 1, dtype=torch.int64, device=m.device)
            r2e_idx = torch.empty(nb * m.nt, dtype=torch.int32, device=m.device)
            L.check(lib.mwx_incidence(self.N, m.nt, nb, L.ptr(ed), L.ptr(r2e_ptr), L.ptr(r2e_idx), st), 'mwx_incidence')
            indptr = torch.empty(self.N + 1, dtype=torch.int64, device=m.device)
            nnz, maxrow = ctypes.c_int64(0), ctypes.c_int32(0)
            L.check(lib.mwx_pattern_count(self.N, m.nt, nb, L.ptr(ed), L.ptr(r2e_ptr), L.ptr(r2e_idx), L.ptr(indptr),
                                          ctypes.byref(nnz), ctypes.byref(maxrow), st), 'mwx_pattern_count')
            indices = torch.empty(nnz.value, dtype=torch.int32, device=m.device)
            slot8 = torch.empty(nb * m.nt, dtype=torch.int64, device=m.device)         # 8 bytes per (row, element)
            L.check(lib.mwx_pattern_fill(self.N, m.nt, nb, L.ptr(ed), L.ptr(r2e_ptr), L.ptr(r2e_idx), L.ptr(indptr),
                                         L.ptr(indices), L.ptr(slot8), st), 'mwx_pattern_fill')
            # 64-byte element geometry records (Morley kernel only)
            exy = torch.empty(8 * m.nt if isinstance(self.elem, ElementTriMorley) else 8, dtype=torch.float64,
                              device=m.device)
            from .assembly import new_pattern_id
            self._pattern = dict(key=new_pattern_id(), r2e_ptr=r2e_ptr, r2e_idx=r2e_idx, indptr=indptr, indices=indices, slot8=slot8, exy=exy,
                                 nnz=int(nnz.value), max_row=int(maxrow.value))
        return self._pattern

    # -- quadrature (host side: problem data - loads, exact solutions - are Python functions of these points)
    def quadrature(self):
        from .quadrature import triangle_rule
        if self.intorder is None:
            raise ValueError("this basis was built without intorder")
        return triangle_rule(self.intorder)

    def d_quadrature(self):
        """[xi | eta | w] as one device vector (kernel argument)."""
        if getattr(self, '_xw', None) is None:
            torch = L.require_cuda()
            X, W = self.quadrature()
            self._xw = torch.from_numpy(np.concatenate((X[0], X[1], W))).to(self.mesh.device)
        return self._xw

    def global_coordinates(self):
        """(2, nt, nq) quadrature points in global coordinates (skfem ``basis.global_coordinates().value``)."""
        m = self.mesh
        X, _ = self.quadrature()
        p, t = m.p, m.t
        x0 = p[:, t[0]]
        return (x0[:, :, None] + (p[:, t[1]] - x0)[:, :, None] * X[0][None, None, :]
                + (p[:, t[2]] - x0)[:, :, None] * X[1][None, None, :])

    def dof_coords(self):
        """Device (N, 2) fp64 DOF locations (vertices, then facet midpoints): the locality key of the
        fast-mode solver's internal renumbering."""
        if getattr(self, '_coords', None) is None:
            torch = L.require_cuda()
            m = self.mesh
            pxy = m.d_pxy.reshape(m.nv, 2)
            if self.elem.facet_dofs:
                f = m.d_facets.reshape(2, m.nf).long()
                self._coords = torch.cat((pxy, 0.5 * (pxy[f[0]] + pxy[f[1]])), dim=0).contiguous()
            else:
                self._coords = pxy.contiguous()
        return self._coords

    def constant_candidates(self):
        """Device (N, 1) of ones: the near-nullspace of a second-order operator (`laplace` on P1 / P2 - every DOF of
        those elements is a point value)."""
        torch = L.require_cuda()
        return torch.ones(self.N, 1, dtype=torch.float64, device=self.mesh.device)

    def nullspace_candidates(self):
        """Device (N, 3) fp64: the interpolants of 1, x, y in this basis - the near-nullspace of the stiffness
        eps^2 (hess:hess) + (grad.grad) that smoothed aggregation is built on.  Morley: vertex DOFs (1, x, y),
        facet DOFs (0, n_x, n_y) with the facet normal n = (d_y, -d_x)/|d|, d = p[f1] - p[f0] (skfem's
        orientation: hess:hess annihilates these vectors to round-off).  Lagrange bases: nodal values."""
        if getattr(self, '_candidates', None) is None:
            torch = L.require_cuda()
            m = self.mesh
            pxy = m.d_pxy.reshape(m.nv, 2)
            ones = torch.ones(m.nv, 1, dtype=torch.float64, device=m.device)
            B = torch.cat((ones, pxy), dim=1)
            if self.elem.facet_dofs:
                f = m.d_facets.reshape(2, m.nf).long()
                if isinstance(self.elem, ElementTriMorley):
                    d = pxy[f[1]] - pxy[f[0]]
                    ln = torch.sqrt((d * d).sum(dim=1))
                    Bf = torch.stack((torch.zeros_like(ln), d[:, 1] / ln, -d[:, 0] / ln), dim=1)
                else:                                    # P2: facet DOF = value at the midpoint
                    mid = 0.5 * (pxy[f[0]] + pxy[f[1]])
                    Bf = torch.cat((torch.ones(m.nf, 1, dtype=torch.float64, device=m.device), mid), dim=1)
                B = torch.cat((B, Bf), dim=0)
            self._candidates = B.contiguous()
        return self._candidates

    # -- boundary DOFs -----------------------------------------------------------------------------
    def find_dofs(self, facets=None):
        """``basis.find_dofs()`` -> {'all': DofsView of the boundary}; with a dict of facet sets,
        one DofsView per key (ref main/example1/run.py:74-79, 206)."""
        m = self.mesh
        if facets is None:
            facets = {'all': m.boundary_facets()}
        out = {}
        for key, fs in facets.items():
            fs = np.asarray(fs, dtype=np.int64)
            nodal = np.unique(m.facets[:, fs])
            facet = (fs + m.nv) if self.elem.facet_dofs else None
            out[key] = DofsView(nodal, facet, self.elem.dofnames)
        return out

    @property
    def doflocs(self):
        m = self.mesh
        if not self.elem.facet_dofs:
            return m.p
        return np.hstack((m.p, 0.5 * (m.p[:, m.facets[0]] + m.p[:, m.facets[1]])))


class FacetBasis:
    """Boundary-facet basis for the penalty forms (ref main/example1/run.py:253)."""

    def __init__(self, mesh, elem, intorder=None):
        if not isinstance(elem, ElementTriMorley):
            raise NotImplementedError("FacetBasis is implemented for ElementTriMorley")
        self.mesh, self.elem = mesh, elem
        self.interior = InteriorBasis(mesh, elem)
        self.N = self.interior.N
        self.find = mesh.boundary_facets()
        # owning triangle + local edge of every boundary facet (O(sqrt N) host work)
        torch = L.require_cuda()
        t2f = mesh.d_t2f.reshape(3, mesh.nt).long()
        f2t = torch.full((mesh.nf,), -1, dtype=torch.int64, device=mesh.device)
        ids = torch.arange(mesh.nt, dtype=torch.int64, device=mesh.device) * 4
        for k in range(3):
            f2t[t2f[k]] = ids + k            # a boundary facet has exactly one writer -> deterministic
        code = f2t[torch.from_numpy(self.find).to(mesh.device)].cpu().numpy()
        tind, loc = code >> 2, code & 3
        self.tind, self.local_edge = tind, loc
        self._plan = None

    def plan(self):
        """Row-wise gather plan for mwx_assemble_penalty: touched rows and their (facet, local) pairs."""
        if self._plan is None:
            torch = L.require_cuda()
            dev = self.mesh.device
            ed = self.interior.element_dofs[:, self.tind]                 # (6, nbf)
            nbf = ed.shape[1]
            rows = ed.reshape(-1)
            code = (np.tile(np.arange(nbf), 6) << 3) | np.repeat(np.arange(6), nbf)
            order = np.lexsort((code, rows))
            rows_s, code_s = rows[order], code[order]
            urows, start = np.unique(rows_s, return_index=True)
            ptr = np.r_[start, len(rows_s)].astype(np.int64)
            self._plan = dict(
                nbf=nbf, nrows=len(urows),
                rows=torch.from_numpy(urows.astype(np.int32)).to(dev),
                ptr=torch.from_numpy(ptr).to(dev),
                code=torch.from_numpy(code_s.astype(np.int32)).to(dev),
                btri=torch.from_numpy(self.tind.astype(np.int32)).to(dev),
                blocal=torch.from_numpy(self.local_edge.astype(np.int32)).to(dev))
        return self._plan
