"""``asm_gpu(form, basis)``: device assembly behind skfem's ``asm`` call shape.

Replaces ``asm(a_load, basis['u'])``, ``asm(b_load, basis['u'])`` and ``asm(penalty_k, fbasis)``
(ref main/example1/run.py:210,255-257,260).  The result is a :class:`DeviceCSR`: it keeps its arrays
in HBM so that ``eps**2 * A + B``, ``condense`` and the solvers stay on the device, and converts to a
canonical ``scipy.sparse.csr_matrix`` (skfem's DOF numbering, sorted indices) on demand.
"""
import ctypes

import numpy as np

from . import _lib as L
from .basis import FacetBasis, InteriorBasis
from .forms import as_form


class DeviceCSR:
    """CSR matrix resident in HBM (int64 indptr, int32 indices, fp64 data); scipy-compatible view."""

    def __init__(self, indptr, indices, data, shape, pattern_key=None, coords=None):
        self.d_indptr, self.d_indices, self.d_data = indptr, indices, data
        self.shape = tuple(int(s) for s in shape)
        self.pattern_key = pattern_key if pattern_key is not None else (indptr.data_ptr(), indices.data_ptr())
        self.coords = coords          # optional callable -> device (n, 2) DOF locations (locality key)
        self._spmv_plan = None

    def spmv_hint(self):
        """-(tile rows) planned once per matrix for mwx_spmv (exact fit of the TMA stage)."""
        if self._spmv_plan is None:
            rows = ctypes.c_int32(0)
            L.check(L.lib().mwx_spmv_plan(self.shape[0], L.ptr(self.d_indptr), ctypes.byref(rows), L.stream()),
                    'mwx_spmv_plan')
            self._spmv_plan = -int(rows.value)
        return self._spmv_plan

    # -- scipy-compatible surface ---------------------------------------------------------------
    @property
    def nnz(self):
        return int(self.d_data.numel())

    @property
    def dtype(self):
        return np.dtype(np.float64)

    @property
    def indptr(self):
        ip = self.d_indptr.cpu().numpy()
        return ip.astype(np.int32) if self.nnz < 2 ** 31 else ip

    @property
    def indices(self):
        return self.d_indices.cpu().numpy()

    @property
    def data(self):
        return self.d_data.cpu().numpy()

    def tocsr(self):
        import scipy.sparse as sp
        return sp.csr_matrix((self.data, self.indices, self.indptr), shape=self.shape)

    to_scipy = tocsr

    def toarray(self):
        return self.tocsr().toarray()

    def diagonal(self):
        torch = L.require_cuda()
        d = torch.empty(self.shape[0], dtype=torch.float64, device=self.d_data.device)
        L.check(L.lib().mwx_csr_diagonal(self.shape[0], L.ptr(self.d_indptr), L.ptr(self.d_indices),
                                         L.ptr(self.d_data), L.ptr(d), L.stream()), 'mwx_csr_diagonal')
        return d.cpu().numpy()

    # -- arithmetic that stays on the device -------------------------------------------------------
    def _same_pattern(self, other):
        return isinstance(other, DeviceCSR) and other.pattern_key == self.pattern_key

    def _axpby(self, alpha, other, beta):
        torch = L.require_cuda()
        out = torch.empty_like(self.d_data)
        b = other.d_data if other is not None else self.d_data
        L.check(L.lib().mwx_axpby(self.nnz, float(alpha), L.ptr(self.d_data), float(beta), L.ptr(b), L.ptr(out),
                                  L.stream()), 'mwx_axpby')
        return DeviceCSR(self.d_indptr, self.d_indices, out, self.shape, self.pattern_key, self.coords)

    def __mul__(self, c):
        if np.isscalar(c):
            return self._axpby(float(c), None, 0.0)
        return self.dot(c)

    def __rmul__(self, c):
        if np.isscalar(c):
            return self._axpby(float(c), None, 0.0)
        return NotImplemented

    def __neg__(self):
        return self._axpby(-1.0, None, 0.0)

    def __add__(self, other):
        if self._same_pattern(other):
            return self._axpby(1.0, other, 1.0)
        raise NotImplementedError("DeviceCSR + operand with a different sparsity pattern: convert with "
                                  ".tocsr() first (no silent host fallback)")

    __radd__ = __add__

    def __sub__(self, other):
        if self._same_pattern(other):
            return self._axpby(1.0, other, -1.0)
        raise NotImplementedError("DeviceCSR - operand with a different sparsity pattern")

    def dot(self, x):
        """y = A x on the device; numpy in -> numpy out, torch (cuda) in -> torch out."""
        torch = L.require_cuda()
        is_np = isinstance(x, np.ndarray)
        xd = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).to(self.d_data.device) if is_np else x
        if xd.numel() != self.shape[1]:
            raise ValueError("dimension mismatch")
        y = torch.empty(self.shape[0], dtype=torch.float64, device=self.d_data.device)
        L.check(L.lib().mwx_spmv(self.shape[0], self.spmv_hint(), L.ptr(self.d_indptr), L.ptr(self.d_indices), L.ptr(self.d_data),
                                 L.ptr(xd), L.ptr(y), L.stream()), 'mwx_spmv')
        return y.cpu().numpy() if is_np else y

    __matmul__ = dot

    @classmethod
    def from_scipy(cls, A, device=None):
        torch = L.require_cuda()
        import scipy.sparse as sp
        A = sp.csr_matrix(A)
        if not A.has_sorted_indices:
            A = A.sorted_indices()
        dev = device or torch.device('cuda', torch.cuda.current_device())
        return cls(torch.from_numpy(A.indptr.astype(np.int64)).to(dev),
                   torch.from_numpy(A.indices.astype(np.int32)).to(dev),
                   torch.from_numpy(A.data.astype(np.float64)).to(dev), A.shape)


def asm_gpu(form, ubasis, vbasis=None, sigma=5.0, out=None):
    """Assemble ``form`` on ``ubasis`` -> :class:`DeviceCSR` (``.tocsr()`` gives scipy's matrix).

    ``form``: a descriptor from :mod:`forms` (or a linear combination such as
    ``eps**2 * a_load + b_load`` - one fused kernel pass), or an object named like the reference's forms.
    ``sigma``: the module-level ``sigma`` that ``penalty_3`` closes over (ref main/example1/run.py:27,146).
    Facet forms are returned on the INTERIOR pattern so that ``eps**2*A + eps**2*P + B`` is a device axpby.
    """
    torch = L.require_cuda()
    f = as_form(form)
    if vbasis is not None and vbasis is not ubasis:
        raise NotImplementedError("asm_gpu: mixed trial/test bases are not on the device path yet")
    lib, st = L.lib(), L.stream()
    if isinstance(ubasis, FacetBasis):
        if f.kind != 'facet':
            raise NotImplementedError("interior forms need an InteriorBasis")
        ib, m = ubasis.interior, ubasis.mesh
        pat = ib.pattern()
        vals = out if out is not None else torch.zeros(pat['nnz'], dtype=torch.float64, device=m.device)
        pl = ubasis.plan()
        t = f.terms
        L.check(lib.mwx_assemble_penalty(m.nt, pl['nbf'], L.ptr(m.d_pxy), L.ptr(m.d_t), L.ptr(ib.d_element_dofs),
                                         L.ptr(pl['btri']), L.ptr(pl['blocal']), pl['nrows'], L.ptr(pl['rows']),
                                         L.ptr(pl['ptr']), L.ptr(pl['code']), L.ptr(pat['indptr']),
                                         L.ptr(pat['indices']), t.get('penalty_1', 0.0), t.get('penalty_2', 0.0),
                                         t.get('penalty_3', 0.0), float(sigma), L.ptr(vals), st),
                'mwx_assemble_penalty')
        return DeviceCSR(pat['indptr'], pat['indices'], vals, (ib.N, ib.N), coords=ib.dof_coords)
    if not isinstance(ubasis, InteriorBasis):
        raise TypeError("asm_gpu expects an InteriorBasis or FacetBasis of this package")
    if f.kind != 'interior':
        raise NotImplementedError("facet forms need a FacetBasis")
    m = ubasis.mesh
    pat = ubasis.pattern()
    vals = out if out is not None else torch.empty(pat['nnz'], dtype=torch.float64, device=m.device)
    L.check(lib.mwx_assemble_morley(ubasis.N, m.nt, L.ptr(m.d_pxy), L.ptr(m.d_t), L.ptr(pat['exy']), 1,
                                    L.ptr(pat['r2e_ptr']), L.ptr(pat['r2e_idx']), L.ptr(pat['slot8']),
                                    L.ptr(pat['indptr']), f.terms.get('a_load', 0.0), f.terms.get('b_load', 0.0), 0,
                                    L.ptr(vals), st),
            'mwx_assemble_morley')
    return DeviceCSR(pat['indptr'], pat['indices'], vals, (ubasis.N, ubasis.N), coords=ubasis.dof_coords)


asm = asm_gpu
