"""Oracle (test infrastructure): restatement of scikit-fem 2.1.1 ``MeshTri``.

Stands in for ``MeshTri()``, ``m.refine()``, ``MeshTri().init_lshaped()``,
``m.facets_satisfying`` and ``m.param()`` as called at
ref main/example1/run.py:508,512,74-79,534 and ref main/example3/run.py:117.
scikit-fem itself is an un-vendored dependency (pin: ref README.md:17); its
published behaviour is:

* every column of ``t`` is sorted ascending;
* ``facets`` = lexicographically sorted unique (min,max) vertex pairs, collected in
  the order local edge (0,1) for all triangles, then (1,2), then (0,2);
* ``t2f[k, e]`` = facet id of local edge k of triangle e;
* uniform (red) refinement: midpoint of facet j becomes vertex ``nv + j``; children
  are emitted in four blocks (t0,m01,m02), (t1,m01,m12), (t2,m02,m12), (m01,m12,m02).
"""
import numpy as np


def _build_mappings(t):
    t = np.sort(t, axis=0)
    pairs = np.hstack((t[[0, 1]], t[[1, 2]], t[[0, 2]]))          # (2, 3 nt), already (min,max)
    uniq, inverse = np.unique(pairs.T, axis=0, return_inverse=True)
    facets = np.ascontiguousarray(uniq.T)
    t2f = inverse.reshape(3, t.shape[1])
    return t, facets, t2f


class MeshTri:
    def __init__(self, p=None, t=None):
        if p is None:
            p = np.array([[0., 1., 0., 1.],
                          [0., 0., 1., 1.]])
            t = np.array([[0, 1, 2],
                          [1, 2, 3]]).T
        self.p = np.asarray(p, dtype=np.float64)
        self.t, self.facets, self.t2f = _build_mappings(np.asarray(t, dtype=np.int64))

    @classmethod
    def init_lshaped(cls):
        # first quadrant missing (matches arctan3's branch, ref main/example3/utils.py:668-671)
        p = np.array([[0., 1., 0., -1., 0., -1., -1., 1.],
                      [0., 0., 1., 0., -1., -1., 1., -1.]])
        t = np.array([[0, 1, 7], [0, 2, 6], [0, 6, 3],
                      [0, 7, 4], [0, 4, 5], [0, 5, 3]]).T
        return cls(p, t)

    @property
    def nv(self):
        return self.p.shape[1]

    @property
    def nt(self):
        return self.t.shape[1]

    @property
    def nf(self):
        return self.facets.shape[1]

    def refine(self, times=1):
        for _ in range(times):
            p, t, fac = self.p, self.t, self.facets
            mid = self.t2f + p.shape[1]
            newp = 0.5 * (p[:, fac[0]] + p[:, fac[1]])
            self.p = np.hstack((p, newp))
            t = np.hstack((np.vstack((t[0], mid[0], mid[2])),
                           np.vstack((t[1], mid[0], mid[1])),
                           np.vstack((t[2], mid[2], mid[1])),
                           np.vstack((mid[0], mid[1], mid[2]))))
            self.t, self.facets, self.t2f = _build_mappings(t)
        return self

    def f2t(self):
        """(2, nf): the (up to) two triangles of each facet, -1 if none (skfem ``f2t``)."""
        nf, nt = self.nf, self.nt
        out = np.full((2, nf), -1, dtype=np.int64)
        flat = self.t2f.ravel()
        tri = np.tile(np.arange(nt), 3)
        order = np.argsort(flat, kind='stable')
        fs, ts = flat[order], tri[order]
        first = np.r_[True, fs[1:] != fs[:-1]]
        out[0, fs[first]] = ts[first]
        second = ~first
        out[1, fs[second]] = ts[second]
        return out

    def boundary_facets(self):
        cnt = np.bincount(self.t2f.ravel(), minlength=self.nf)
        return np.nonzero(cnt == 1)[0]

    def boundary_nodes(self):
        return np.unique(self.facets[:, self.boundary_facets()])

    def facets_satisfying(self, test):
        mid = 0.5 * (self.p[:, self.facets[0]] + self.p[:, self.facets[1]])
        return np.nonzero(test(mid))[0]

    def param(self):
        d = self.p[:, self.facets[0]] - self.p[:, self.facets[1]]
        return float(np.max(np.sqrt(np.sum(d * d, axis=0))))
