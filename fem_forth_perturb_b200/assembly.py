"""``asm_gpu(form, basis)``: device assembly behind skfem's ``asm`` call shape.

Replaces ``asm(a_load, basis['u'])``, ``asm(b_load, basis['u'])`` and ``asm(penalty_k, fbasis)``
(ref main/example1/run.py:210,255-257,260).  The result is a :class:`DeviceCSR`: it keeps its arrays
in HBM so that ``eps**2 * A + B``, ``condense`` and the solvers stay on the device, and converts to a
canonical ``scipy.sparse.csr_matrix`` (skfem's DOF numbering, sorted indices) on demand.
"""
import ctypes
import itertools

import numpy as np

from . import _lib as L
from .basis import FacetBasis, InteriorBasis
from .forms import as_form

_PATTERN_IDS = itertools.count(1)


def new_pattern_id():
    """Identity of a sparsity pattern: handed out once where a pattern is built and carried by every matrix that
    shares it.  (Device addresses are not identities: the caching allocator hands freed addresses out again.)"""
    return next(_PATTERN_IDS)


class DeviceCSR:
    """CSR matrix resident in HBM (int64 indptr, int32 indices, fp64 data); scipy-compatible view."""

    def __init__(self, indptr, indices, data, shape, pattern_key=None, coords=None, candidates=None):
        self.d_indptr, self.d_indices, self.d_data = indptr, indices, data
        self.shape = tuple(int(s) for s in shape)
        self.pattern_key = pattern_key if pattern_key is not None else new_pattern_id()
        self.coords = coords          # optional callable -> device (n, 2) DOF locations (locality key)
        self.candidates = candidates  # optional callable -> device (n, k) near-nullspace vectors (smoothed aggregation)
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
        return (isinstance(other, DeviceCSR) and other.pattern_key == self.pattern_key and other.shape == self.shape
                and other.nnz == self.nnz)

    def _axpby(self, alpha, other, beta):
        torch = L.require_cuda()
        out = torch.empty_like(self.d_data)
        b = other.d_data if other is not None else self.d_data
        L.check(L.lib().mwx_axpby(self.nnz, float(alpha), L.ptr(self.d_data), float(beta), L.ptr(b), L.ptr(out),
                                  L.stream()), 'mwx_axpby')
        return DeviceCSR(self.d_indptr, self.d_indices, out, self.shape, self.pattern_key, self.coords, self.candidates)

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

    def eliminate_zeros(self):
        """Drop stored entries that are exactly 0.0 (skfem calls ``K.eliminate_zeros()`` before ``tocsr()``)."""
        torch = L.require_cuda()
        mask = self.d_data != 0.0
        csum = torch.zeros(self.nnz + 1, dtype=torch.int64, device=self.d_data.device)
        csum[1:] = torch.cumsum(mask.to(torch.int64), 0)
        return DeviceCSR(csum[self.d_indptr].contiguous(), self.d_indices[mask].contiguous(),
                         self.d_data[mask].contiguous(), self.shape, coords=self.coords, candidates=self.candidates)

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


class MixedAction:
    """``asm(wv_load, basis['w'], basis['u'])`` (ref main/example1/run.py:211) as an operator: the reference only
    ever multiplies this (N_u x N_w) matrix by ``wh``, so it is applied element-wise instead of being stored."""

    def __init__(self, wbasis, ubasis):
        self.wbasis, self.ubasis = wbasis, ubasis
        self.shape = (ubasis.N, wbasis.N)

    def dot(self, wh):
        torch = L.require_cuda()
        is_np = isinstance(wh, np.ndarray)
        m, ub, wb = self.ubasis.mesh, self.ubasis, self.wbasis
        whd = torch.from_numpy(np.ascontiguousarray(wh, dtype=np.float64)).to(m.device) if is_np else wh
        pat = ub.pattern()
        out = torch.empty(ub.N, dtype=torch.float64, device=m.device)
        L.check(L.lib().mwx_wv_action(ub.N, m.nt, wb.Nbfun, L.ptr(m.d_pxy), L.ptr(m.d_t), L.ptr(wb.d_element_dofs),
                                      L.ptr(pat['r2e_ptr']), L.ptr(pat['r2e_idx']), L.ptr(whd), L.ptr(out), L.stream()),
                'mwx_wv_action')
        return out.cpu().numpy() if is_np else out

    __mul__ = dot
    __matmul__ = dot

    def tocsr(self):
        raise NotImplementedError("the mixed (Morley x Lagrange) matrix is applied, not stored; use `A * wh`")


def asm_load(form, basis, on_device=False):
    """``asm(LinearForm, basis)``: host-sampled integrand, device integration -> numpy vector (like skfem).
    ``on_device=True`` (forms of :mod:`problems` on a P1 / P2 basis): the load is evaluated by the kernel and the
    result stays a device tensor - no nt*nq host samples."""
    torch = L.require_cuda()
    m, pat = basis.mesh, basis.pattern()
    out = torch.empty(basis.N, dtype=torch.float64, device=m.device)
    X, W = basis.quadrature()
    from .basis import ElementTriMorley
    if on_device:
        prob = getattr(form, 'device_problem', None)
        if prob is None or isinstance(basis.elem, ElementTriMorley):
            raise NotImplementedError("on_device=True needs a load from fem_forth_perturb_b200.problems on a P1 / P2 "
                                      "basis; arbitrary Python integrands are sampled on the host")
        L.check(L.lib().mwx_load_vector_problem(prob[0], prob[1], basis.N, m.nt, basis.Nbfun, len(W),
                                                L.ptr(basis.d_quadrature()), L.ptr(m.d_pxy), L.ptr(m.d_t),
                                                L.ptr(pat['r2e_ptr']), L.ptr(pat['r2e_idx']), L.ptr(out), L.stream()),
                'mwx_load_vector_problem')
        return out
    fq = torch.from_numpy(form.sample(basis)).to(m.device)
    if isinstance(basis.elem, ElementTriMorley):
        L.check(L.lib().mwx_morley_load_vector(basis.N, m.nt, len(W), L.ptr(basis.d_quadrature()), L.ptr(m.d_pxy),
                                               L.ptr(m.d_t), L.ptr(pat['r2e_ptr']), L.ptr(pat['r2e_idx']), L.ptr(fq),
                                               L.ptr(out), L.stream()), 'mwx_morley_load_vector')
    else:
        L.check(L.lib().mwx_load_vector(basis.N, m.nt, basis.Nbfun, len(W), L.ptr(basis.d_quadrature()), L.ptr(m.d_pxy),
                                        L.ptr(m.d_t), L.ptr(pat['r2e_ptr']), L.ptr(pat['r2e_idx']), L.ptr(fq), L.ptr(out),
                                        L.stream()), 'mwx_load_vector')
    return out.cpu().numpy()


def asm_gpu(form, ubasis, vbasis=None, sigma=5.0, out=None, accumulate=False, on_device=False, kernel=None):
    """Assemble ``form`` on ``ubasis`` -> :class:`DeviceCSR` (``.tocsr()`` gives scipy's matrix).

    ``form``: a descriptor from :mod:`forms` (or a linear combination such as
    ``eps**2 * a_load + b_load`` - one fused kernel pass), or an object named like the reference's forms.
    ``sigma``: the module-level ``sigma`` that ``penalty_3`` closes over (ref main/example1/run.py:27,146).
    Facet forms are returned on the INTERIOR pattern so that ``eps**2*A + eps**2*P + B`` is a device axpby.
    ``out`` / ``accumulate``: write the values into an existing nnz-sized device vector, or add to it (Morley
    interior forms; the penalty forms always add into ``out``).
    ``kernel``: None = the tiled element-once kernel when the mesh fits its plan, else the row-gather kernel;
    'tiles' / 'rows' force one of them (A/B comparisons, tests).
    """
    torch = L.require_cuda()
    from .forms import LinearFormGPU
    from .basis import ElementTriMorley
    if isinstance(form, LinearFormGPU):
        return asm_load(form, ubasis, on_device=on_device)
    f = as_form(form)
    if f.kind == 'mixed':
        if vbasis is None or not isinstance(vbasis.elem, ElementTriMorley) or isinstance(ubasis.elem, ElementTriMorley):
            raise NotImplementedError("wv_load: trial basis P1/P2, test basis Morley (ref run.py:211)")
        return MixedAction(ubasis, vbasis)
    if vbasis is not None and vbasis is not ubasis:
        raise NotImplementedError("asm_gpu: mixed trial/test bases only for wv_load")
    lib, st = L.lib(), L.stream()
    if f.kind == 'mass':
        if not isinstance(ubasis.elem, ElementTriMorley):
            raise NotImplementedError("`mass` is implemented on the Morley basis (example 3's L2 projection)")
        m, pat = ubasis.mesh, ubasis.pattern()
        vals = torch.empty(pat['nnz'], dtype=torch.float64, device=m.device)
        _, W = ubasis.quadrature()
        L.check(lib.mwx_assemble_morley_mass(ubasis.N, m.nt, len(W), L.ptr(ubasis.d_quadrature()), L.ptr(m.d_pxy),
                                             L.ptr(m.d_t), L.ptr(pat['r2e_ptr']), L.ptr(pat['r2e_idx']),
                                             L.ptr(pat['slot8']), L.ptr(pat['indptr']), L.ptr(vals), st),
                'mwx_assemble_morley_mass')
        K = DeviceCSR(pat['indptr'], pat['indices'], vals, (ubasis.N, ubasis.N), pat['key'], coords=ubasis.dof_coords,
                      candidates=ubasis.nullspace_candidates)
        return K if f.terms['mass'] == 1.0 else f.terms['mass'] * K
    if f.kind == 'lagrange':
        if isinstance(ubasis.elem, ElementTriMorley):
            raise NotImplementedError("`laplace` is the w-problem form (P1 / P2); use b_load on the Morley basis")
        m, pat = ubasis.mesh, ubasis.pattern()
        vals = torch.empty(pat['nnz'], dtype=torch.float64, device=m.device)
        L.check(lib.mwx_assemble_lagrange_laplace(ubasis.N, m.nt, ubasis.Nbfun, L.ptr(m.d_pxy), L.ptr(m.d_t),
                                                  L.ptr(pat['r2e_ptr']), L.ptr(pat['r2e_idx']), L.ptr(pat['slot8']),
                                                  L.ptr(pat['indptr']), L.ptr(vals), st), 'mwx_assemble_lagrange_laplace')
        K = DeviceCSR(pat['indptr'], pat['indices'], vals, (ubasis.N, ubasis.N), pat['key'], coords=ubasis.dof_coords,
                      candidates=ubasis.constant_candidates)
        scale = f.terms['laplace']
        K = K if scale == 1.0 else scale * K
        return K.eliminate_zeros()              # skfem: COO.eliminate_zeros() -> 5-point pattern on this mesh
    if not isinstance(ubasis, FacetBasis) and not isinstance(getattr(ubasis, 'elem', None), ElementTriMorley):
        raise NotImplementedError("a_load / b_load are the Morley forms; use `laplace` on P1 / P2")
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
        return DeviceCSR(pat['indptr'], pat['indices'], vals, (ib.N, ib.N), pat['key'], coords=ib.dof_coords,
                         candidates=ib.nullspace_candidates)
    if not isinstance(ubasis, InteriorBasis):
        raise TypeError("asm_gpu expects an InteriorBasis or FacetBasis of this package")
    if f.kind != 'interior':
        raise NotImplementedError("facet forms need a FacetBasis")
    m = ubasis.mesh
    pat = ubasis.pattern()
    vals = out if out is not None else torch.empty(pat['nnz'], dtype=torch.float64, device=m.device)
    tp = _tile_plan(ubasis) if kernel != 'rows' else None
    if kernel == 'tiles' and tp is None:
        raise L.MwxError("asm_gpu(kernel='tiles'): this mesh does not fit the tile plan (see mwx_asm_plan_create)")
    if tp is not None:
        tp.assemble(m, f.terms.get('a_load', 0.0), f.terms.get('b_load', 0.0), vals, accumulate and out is not None)
        return DeviceCSR(pat['indptr'], pat['indices'], vals, (ubasis.N, ubasis.N), pat['key'], coords=ubasis.dof_coords,
                         candidates=ubasis.nullspace_candidates)
    L.check(lib.mwx_assemble_morley(ubasis.N, m.nt, L.ptr(m.d_pxy), L.ptr(m.d_t), L.ptr(pat['exy']), 1,
                                    L.ptr(pat['r2e_ptr']), L.ptr(pat['r2e_idx']), L.ptr(pat['slot8']),
                                    L.ptr(pat['indptr']), f.terms.get('a_load', 0.0), f.terms.get('b_load', 0.0),
                                    1 if (accumulate and out is not None) else 0, L.ptr(vals), st),
            'mwx_assemble_morley')
    return DeviceCSR(pat['indptr'], pat['indices'], vals, (ubasis.N, ubasis.N), pat['key'], coords=ubasis.dof_coords,
                     candidates=ubasis.nullspace_candidates)


asm = asm_gpu


class TilePlan:
    """Plan of the tiled Morley kernel (``mwx_asm_plan_create``) for one destination layout; owns the device blobs.

    rows: device int32 skfem DOF per destination row in locality order; dst_off / dst_len / dst_col: where each row
    lives in the destination value array and which skfem DOF each of its entries is."""

    def __init__(self, basis, rows, dst_off, dst_len, dst_col):
        m, pat = basis.mesh, basis.pattern()
        self._h = ctypes.c_void_p()
        self._keep = (rows, dst_off, dst_len, dst_col)
        rc = L.lib().mwx_asm_plan_create(rows.numel(), L.ptr(rows), L.ptr(dst_off), L.ptr(dst_len), L.ptr(dst_col), m.nt,
                                         L.ptr(m.d_t), L.ptr(basis.d_element_dofs), L.ptr(pat['r2e_ptr']),
                                         L.ptr(pat['r2e_idx']), ctypes.byref(self._h), L.stream())
        self.ok = rc == 0
        if rc not in (0, -2):                       # -2 = MWX_E_CAPACITY: the mesh needs the row-gather kernel
            L.check(rc, 'mwx_asm_plan_create')
        self._keep = None                           # the plan holds its own copies of everything it needs
        self.info = None
        if self.ok:
            info = np.zeros(4, dtype=np.int64)
            L.check(L.lib().mwx_asm_plan_info(self._h, L.ptr(info)), 'mwx_asm_plan_info')
            self.info = dict(tiles=int(info[0]), rows_per_tile=int(info[1]), entry_bytes=int(info[2]), geometry_bytes=int(info[3]))

    @classmethod
    def for_matrix(cls, basis):
        """Plan for the global K2 of ``basis`` (skfem CSR layout): rows in Hilbert order of the DOF locations."""
        torch = L.require_cuda()
        pat = basis.pattern()
        order = hilbert_order(basis.dof_coords())
        indptr = pat['indptr']
        rows = order.to(torch.int32)
        dst_off = indptr[order].contiguous()
        dst_len = (indptr[1:] - indptr[:-1])[order].to(torch.int32).contiguous()
        return cls(basis, rows, dst_off, dst_len, pat['indices'])

    def assemble(self, mesh, ca, cb, values, accumulate=False):
        L.check(L.lib().mwx_assemble_morley_plan(self._h, L.ptr(mesh.d_pxy), float(ca), float(cb), 1 if accumulate else 0,
                                                 L.ptr(values), L.stream()), 'mwx_assemble_morley_plan')

    def __del__(self):
        try:
            if getattr(self, '_h', None):
                L.lib().mwx_asm_plan_destroy(self._h)
        except Exception:
            pass


def hilbert_order(coords):
    """argsort of the Hilbert-curve keys of (n, 2) device points (new -> old), int64."""
    torch = L.require_cuda()
    xy = coords.contiguous()
    n = xy.shape[0]
    lo, hi = xy.min(dim=0).values, xy.max(dim=0).values
    extent = float((hi - lo).max().item()) or 1.0
    keys = torch.empty(n, dtype=torch.int64, device=xy.device)
    L.check(L.lib().mwx_hilbert_keys(n, L.ptr(xy), float(lo[0].item()), float(lo[1].item()), extent, L.ptr(keys), L.stream()),
            'mwx_hilbert_keys')
    return torch.sort(keys, stable=True).indices               # library sort: setup only


def _tile_plan(basis):
    """The cached tile plan of a Morley basis, or None when the row-gather kernel has to be used."""
    import os
    if os.environ.get('MWX_ASM_KERNEL', 'tiles') != 'tiles':
        return None
    pat = basis.pattern()
    if 'tile_plan' not in pat:
        plan = TilePlan.for_matrix(basis)
        pat['tile_plan'] = plan if plan.ok else None
    return pat['tile_plan']


def assembly_kernel_name(basis=None):
    """Name of the kernel(s) behind ``asm_gpu(eps**2 * a_load + b_load, basis)`` (bench.py's roofline label)."""
    if basis is None or _tile_plan(basis) is not None:
        return 'k_assemble_morley_tiles (tiled, every element once; plan built at setup)'
    return 'k_element_geometry + k_assemble_morley_persistent (row-gather fallback)'


def morley_error_norms(basis, uh, exact_u=None, dexact_u=None, ddexact=None, problem=None):
    """(L2, D1, D2) = (||u_h - u||_0, |u_h - u|_{1,h}, |u_h - u|_{2,h}) as ref main/example1/run.py:157-179,519-531
    (``L2uError``, ``get_DuError``, ``get_D2uError``): the exact solution is sampled at the quadrature points on
    the host (it is a Python function), the element integrals and their sum run on the device.
    ``problem`` = a :class:`problems.Problem`: its exact solution is evaluated by the kernel instead (no host samples)."""
    torch = L.require_cuda()
    m = basis.mesh
    if problem is not None:
        uhd = torch.from_numpy(np.ascontiguousarray(uh, dtype=np.float64)).to(m.device) if isinstance(uh, np.ndarray) else uh
        sums = torch.empty(3, dtype=torch.float64, device=m.device)
        _, W = basis.quadrature()
        L.check(L.lib().mwx_morley_error_sums_problem(problem.device_problem[0], m.nt, len(W), L.ptr(basis.d_quadrature()),
                                                      L.ptr(m.d_pxy), L.ptr(m.d_t), L.ptr(basis.d_element_dofs), L.ptr(uhd),
                                                      L.ptr(sums), L.stream()), 'mwx_morley_error_sums_problem')
        s = np.sqrt(sums.cpu().numpy())
        return float(s[0]), float(s[1]), float(s[2])
    x = basis.global_coordinates()
    u = exact_u(x[0], x[1])
    ux, uy = dexact_u(x[0], x[1])
    uxx, uxy, _, uyy = ddexact(x[0], x[1])
    ex = np.ascontiguousarray(np.stack([np.broadcast_to(a, x[0].shape) for a in (u, ux, uy, uxx, uxy, uyy)]),
                              dtype=np.float64)
    exd = torch.from_numpy(ex).to(m.device)
    uhd = torch.from_numpy(np.ascontiguousarray(uh, dtype=np.float64)).to(m.device) if isinstance(uh, np.ndarray) else uh
    sums = torch.empty(3, dtype=torch.float64, device=m.device)
    _, W = basis.quadrature()
    L.check(L.lib().mwx_morley_error_sums(m.nt, len(W), L.ptr(basis.d_quadrature()), L.ptr(m.d_pxy), L.ptr(m.d_t),
                                          L.ptr(basis.d_element_dofs), L.ptr(uhd), L.ptr(exd), L.ptr(sums), L.stream()),
            'mwx_morley_error_sums')
    s = np.sqrt(sums.cpu().numpy())
    return float(s[0]), float(s[1]), float(s[2])


def morley_nested_error_norms(base_basis, u_base, coarse_basis, u_coarse):
    """(L2, D1, D2) of u^N - u_h for a level-N solution and a solution of a coarser mesh of the same uniform
    refinement family, at the quadrature points of the level-N mesh (ref main/example2/run.py:101-172: the
    per-point ``interpolator`` loops and the DG-P1 / DG-P0 ``project`` chain - exact for a piecewise quadratic, so
    they reduce to sampling u_h and its elementwise derivatives).  One sampling kernel + the error-sum kernel."""
    torch = L.require_cuda()
    mf, mc = base_basis.mesh, coarse_basis.mesh
    dev = mf.device
    as_dev = lambda v: torch.from_numpy(np.ascontiguousarray(v, dtype=np.float64)).to(dev) if isinstance(v, np.ndarray) else v
    uN, uc = as_dev(u_base), as_dev(u_coarse)
    _, W = base_basis.quadrature()
    nq = len(W)
    samples = torch.empty(6 * mf.nt * nq, dtype=torch.float64, device=dev)
    L.check(L.lib().mwx_morley_sample_nested(mf.nt, nq, L.ptr(base_basis.d_quadrature()), L.ptr(mf.d_pxy), L.ptr(mf.d_t),
                                             mc.nt, L.ptr(mc.d_pxy), L.ptr(mc.d_t), L.ptr(coarse_basis.d_element_dofs),
                                             L.ptr(uc), L.ptr(samples), L.stream()), 'mwx_morley_sample_nested')
    sums = torch.empty(3, dtype=torch.float64, device=dev)
    L.check(L.lib().mwx_morley_error_sums(mf.nt, nq, L.ptr(base_basis.d_quadrature()), L.ptr(mf.d_pxy), L.ptr(mf.d_t),
                                          L.ptr(base_basis.d_element_dofs), L.ptr(uN), L.ptr(samples), L.ptr(sums),
                                          L.stream()), 'mwx_morley_error_sums')
    s = np.sqrt(sums.cpu().numpy())
    return float(s[0]), float(s[1]), float(s[2])


def morley_facet_pnv(fbasis, uh):
    """``L2pnvError.assemble(fbasis, w=fbasis.interpolate(uh0))`` (ref main/example1/run.py:106-108,527):
    sum over the boundary facets of int_F (h_F n . grad u_h)^2."""
    torch = L.require_cuda()
    m, ib, pl = fbasis.mesh, fbasis.interior, fbasis.plan()
    uhd = torch.from_numpy(np.ascontiguousarray(uh, dtype=np.float64)).to(m.device) if isinstance(uh, np.ndarray) else uh
    terms = torch.empty(pl['nbf'], dtype=torch.float64, device=m.device)
    L.check(L.lib().mwx_morley_facet_pnv(m.nt, pl['nbf'], L.ptr(m.d_pxy), L.ptr(m.d_t), L.ptr(ib.d_element_dofs),
                                         L.ptr(pl['btri']), L.ptr(pl['blocal']), L.ptr(uhd), L.ptr(terms), L.stream()),
            'mwx_morley_facet_pnv')
    return float(np.sum(terms.cpu().numpy()))          # O(sqrt N) terms, summed in index order


def project(fun, basis_to):
    """``project(exact_u, basis_to=basis['u'])`` (ref main/example3/run.py:66; main/example3/utils.py:494-565):
    global L2 projection M x = int fun phi.  The reference calls a sparse direct solver; the Morley mass matrix is
    well conditioned, so Jacobi-PCG to 1e-14 on the device gives the same vector to round-off."""
    from .forms import LinearFormGPU, mass
    from .solvers import pcg
    M = asm_gpu(mass, basis_to)
    f = asm_gpu(LinearFormGPU(lambda v, w: fun(w.x[0], w.x[1]) * v), basis_to)
    x, its, info, resid = pcg(M, f, tol=1e-14, maxiter=20000, precondition=True)
    if info != 0:
        raise L.MwxError(f"project: the mass-matrix solve did not converge (info={info}, residual {resid:.3e})")
    return x.cpu().numpy()
