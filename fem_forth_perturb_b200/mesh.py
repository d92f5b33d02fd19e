"""Device-resident triangle mesh with scikit-fem's numbering.

Host-side mirror of ``skfem.MeshTri`` as the reference uses it
(``MeshTri()``, ``m.refine()``, ``MeshTri().init_lshaped()``, ``m.facets_satisfying``, ``m.param()``;
ref main/example1/run.py:508,512,74-79,534; ref main/example3/run.py:117).  All arrays live in HBM
(torch tensors are only the allocation); ``p``, ``t``, ``facets``, ``t2f`` are host copies made on
demand with skfem's shapes and dtypes.
"""
import ctypes

import numpy as np

from . import _lib as L


class MeshTri:
    def __init__(self, p=None, t=None):
        torch = L.require_cuda()
        if p is None:
            p = np.array([[0., 1., 0., 1.], [0., 0., 1., 1.]])
            t = np.array([[0, 1, 2], [1, 2, 3]]).T
        p = np.ascontiguousarray(np.asarray(p, dtype=np.float64))
        t = np.ascontiguousarray(np.asarray(t, dtype=np.int32))
        if p.ndim != 2 or p.shape[0] != 2 or t.ndim != 2 or t.shape[0] != 3:
            raise ValueError("MeshTri expects p of shape (2, nv) and t of shape (3, nt)")
        dev = torch.device('cuda', torch.cuda.current_device())
        self.device = dev
        self.d_pxy = torch.from_numpy(np.ascontiguousarray(p.T)).to(dev).reshape(-1)       # (x0,y0,x1,y1,...)
        self.d_t = torch.from_numpy(t).to(dev).reshape(-1)                                # [t0 | t1 | t2]
        self.nv, self.nt = p.shape[1], t.shape[1]
        L.check(L.lib().mwx_mesh_sort_columns(self.nt, L.ptr(self.d_t), L.stream()), 'mwx_mesh_sort_columns')
        self._build_facets()

    # -- construction -------------------------------------------------------------------------
    @classmethod
    def init_lshaped(cls):
        """L-shaped domain with the first quadrant missing (ref main/example3/run.py:117)."""
        p = np.array([[0., 1., 0., -1., 0., -1., -1., 1.],
                      [0., 0., 1., 0., -1., -1., 1., -1.]])
        t = np.array([[0, 1, 7], [0, 2, 6], [0, 6, 3], [0, 7, 4], [0, 4, 5], [0, 5, 3]]).T
        return cls(p, t)

    def _build_facets(self):
        torch = L.require_cuda()
        lib, st = L.lib(), L.stream()
        dev = self.device
        self._host = {}
        self.d_v2t_ptr = torch.empty(self.nv + 1, dtype=torch.int64, device=dev)
        self.d_v2t_idx = torch.empty(3 * self.nt, dtype=torch.int32, device=dev)
        L.check(lib.mwx_incidence(self.nv, self.nt, 3, L.ptr(self.d_t), L.ptr(self.d_v2t_ptr),
                                  L.ptr(self.d_v2t_idx), st), 'mwx_incidence')
        fstart = torch.empty(self.nv + 1, dtype=torch.int64, device=dev)
        nf = ctypes.c_int64(0)
        L.check(lib.mwx_mesh_facets_count(self.nv, self.nt, L.ptr(self.d_t), L.ptr(self.d_v2t_ptr),
                                          L.ptr(self.d_v2t_idx), L.ptr(fstart), ctypes.byref(nf), st),
                'mwx_mesh_facets_count')
        self.nf = int(nf.value)
        self.d_facets = torch.empty(2 * self.nf, dtype=torch.int32, device=dev)
        self.d_t2f = torch.empty(3 * self.nt, dtype=torch.int32, device=dev)
        L.check(lib.mwx_mesh_facets_fill(self.nv, self.nt, L.ptr(self.d_t), L.ptr(self.d_v2t_ptr),
                                         L.ptr(self.d_v2t_idx), L.ptr(fstart), self.nf,
                                         L.ptr(self.d_facets), L.ptr(self.d_t2f), st), 'mwx_mesh_facets_fill')

    @classmethod
    def _from_device(cls, d_pxy, d_t, nv, nt):
        """Mesh from device arrays (pxy interleaved, t as [t0 | t1 | t2] int32); columns of t get sorted like the
        public constructor does."""
        self = cls.__new__(cls)
        self.device = d_pxy.device
        self.d_pxy, self.d_t, self.nv, self.nt = d_pxy.contiguous(), d_t.contiguous(), int(nv), int(nt)
        L.check(L.lib().mwx_mesh_sort_columns(self.nt, L.ptr(self.d_t), L.stream()), 'mwx_mesh_sort_columns')
        self._build_facets()
        return self

    def renumbered(self):
        """The same mesh with vertices and elements renumbered along a Hilbert curve (no reference counterpart; the
        reference's meshes keep the order red refinement produces).  Facet and DOF numbers then follow skfem's rules
        on the renumbered arrays, so everything downstream is unchanged - but an element's vertices, edges and
        neighbours now sit close together in memory, which is what the row-gather assembly and the SpMV's x-gathers
        want.  ``vertex_order[i]`` / ``element_order[j]`` give the old number of new vertex i / new element j."""
        torch = L.require_cuda()
        lib, st = L.lib(), L.stream()
        pxy = self.d_pxy.reshape(self.nv, 2)
        lo = pxy.min(dim=0).values
        ext = float((pxy.max(dim=0).values - lo).max().item()) or 1.0
        x0, y0 = float(lo[0].item()), float(lo[1].item())
        keys = torch.empty(self.nv, dtype=torch.int64, device=self.device)
        L.check(lib.mwx_hilbert_keys(self.nv, L.ptr(self.d_pxy), x0, y0, ext, L.ptr(keys), st), 'mwx_hilbert_keys')
        vorder = torch.argsort(keys, stable=True)
        inv = torch.empty_like(vorder)
        inv[vorder] = torch.arange(self.nv, dtype=torch.int64, device=self.device)
        new_pxy = pxy[vorder].contiguous().reshape(-1)
        t = self.d_t.reshape(3, self.nt).long()
        cxy = (pxy[t[0]] + pxy[t[1]] + pxy[t[2]]) / 3.0
        ckeys = torch.empty(self.nt, dtype=torch.int64, device=self.device)
        cflat = cxy.contiguous().reshape(-1)
        L.check(lib.mwx_hilbert_keys(self.nt, L.ptr(cflat), x0, y0, ext, L.ptr(ckeys), st), 'mwx_hilbert_keys')
        eorder = torch.argsort(ckeys, stable=True)
        new_t = inv[t[:, eorder]].to(torch.int32).contiguous().reshape(-1)
        m = MeshTri._from_device(new_pxy, new_t, self.nv, self.nt)
        m.vertex_order, m.element_order = vorder, eorder
        return m

    def refine(self, times=1):
        """Uniform red refinement, in place like skfem's ``m.refine()``."""
        torch = L.require_cuda()
        for _ in range(times):
            nvn, ntn = self.nv + self.nf, 4 * self.nt
            pxy = torch.empty(2 * nvn, dtype=torch.float64, device=self.device)
            tn = torch.empty(3 * ntn, dtype=torch.int32, device=self.device)
            L.check(L.lib().mwx_mesh_refine(self.nv, self.nt, self.nf, L.ptr(self.d_pxy), L.ptr(self.d_t),
                                            L.ptr(self.d_facets), L.ptr(self.d_t2f), L.ptr(pxy), L.ptr(tn),
                                            L.stream()), 'mwx_mesh_refine')
            self.d_pxy, self.d_t, self.nv, self.nt = pxy, tn, nvn, ntn
            self._build_facets()
        return self

    # -- skfem-shaped host views -----------------------------------------------------------------
    def _cached(self, key, make):
        if key not in self._host:
            self._host[key] = make()
        return self._host[key]

    @property
    def p(self):
        return self._cached('p', lambda: np.ascontiguousarray(self.d_pxy.reshape(self.nv, 2).cpu().numpy().T))

    @property
    def t(self):
        return self._cached('t', lambda: self.d_t.reshape(3, self.nt).cpu().numpy().astype(np.int64))

    @property
    def facets(self):
        return self._cached('facets', lambda: self.d_facets.reshape(2, self.nf).cpu().numpy().astype(np.int64))

    @property
    def t2f(self):
        return self._cached('t2f', lambda: self.d_t2f.reshape(3, self.nt).cpu().numpy().astype(np.int64))

    # -- queries ---------------------------------------------------------------------------------
    def facet_element_count(self):
        torch = L.require_cuda()
        cnt = torch.empty(self.nf, dtype=torch.int32, device=self.device)
        L.check(L.lib().mwx_facet_element_count(self.nf, self.nt, L.ptr(self.d_t2f), L.ptr(cnt), L.stream()),
                'mwx_facet_element_count')
        return cnt

    def boundary_facets(self):
        def make():
            torch = L.require_cuda()
            return torch.nonzero(self.facet_element_count() == 1).reshape(-1).cpu().numpy()
        return self._cached('bfacets', make)

    def boundary_nodes(self):
        return np.unique(self.facets[:, self.boundary_facets()])

    def facets_satisfying(self, test, boundaries_only=False):
        """Facets whose midpoint satisfies ``test`` (skfem semantics, ref main/example1/run.py:74-79).
        ``boundaries_only=True`` tests only boundary facets (O(sqrt N) host work) - what the
        reference's side selectors x == 0 / x == 1 / y == 0 / y == 1 can match anyway."""
        if boundaries_only:
            bf = self.boundary_facets()
            fac = self.facets[:, bf]
            mid = 0.5 * (self.p[:, fac[0]] + self.p[:, fac[1]])
            return bf[np.nonzero(test(mid))[0]]
        mid = 0.5 * (self.p[:, self.facets[0]] + self.p[:, self.facets[1]])
        return np.nonzero(test(mid))[0]

    def param(self):
        torch = L.require_cuda()
        pxy = self.d_pxy.reshape(self.nv, 2)
        f = self.d_facets.reshape(2, self.nf).long()
        d = pxy[f[0]] - pxy[f[1]]
        return float(torch.sqrt((d * d).sum(dim=1)).max().item())
