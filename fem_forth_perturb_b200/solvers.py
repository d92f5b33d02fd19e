"""Drop-in solver layer: the reference's ``utils.py`` call shapes on top of the device kernels.

Mirrors ref main/example1/utils.py:
  ``solver_iter_krylov`` / ``solver_iter_pcg`` (:265-320), ``solver_iter_krylov_iter`` (:212-263),
  ``solver_iter_mgcg`` (:127-166), ``solver_iter_mgcg_iter`` (:87-125), ``build_pc_diag`` (:48-50),
  ``solve`` (:326-368), ``condense`` (:384-460).
Every factory returns a closure ``solver(A, b, **kw) -> x`` (numpy) that the reference's ``solve()``
accepts unchanged; the closures additionally expose ``solver.solve_with_info(A, b) -> (x, iters)`` and
``solver.last`` (iterations, info, residual, setup/solve seconds).  Same kwargs (``tol``, ``maxiter``,
``Precondition``, ``verbose``), same ``"... total interation steps: N"`` print in the ``_iter`` flavours,
same ``warnings.warn("Convergence not achieved!")``.  A user-supplied ``M`` / ``krylov`` callable cannot
run on the device and raises (no CPU fallback).
"""
import ctypes
import time
import warnings

import numpy as np

from . import _lib as L
from .assembly import DeviceCSR, new_pattern_id


# ------------------------------------------------------------------ helpers
def _to_device_csr(A):
    return A if isinstance(A, DeviceCSR) else DeviceCSR.from_scipy(A)


def _to_device_vec(b, device):
    torch = L.require_cuda()
    if isinstance(b, np.ndarray):
        return torch.from_numpy(np.ascontiguousarray(b, dtype=np.float64)).to(device)
    return b.to(device=device, dtype=torch.float64).contiguous()


def _gather_values(src, idx, out):
    """out[k] = src[idx[k]] on the device (idx int32 or int64)."""
    fn = L.lib().mwx_gather_f64 if idx.dtype == L.require_cuda().int32 else L.lib().mwx_gather_f64_i64
    L.check(fn(idx.numel(), L.ptr(src), L.ptr(idx), L.ptr(out), L.stream()), 'mwx_gather_f64')
    return out


def _index_dtype(nnz):
    torch = L.require_cuda()
    return torch.int32 if nnz < 2 ** 31 else torch.int64


class Reordering:
    """Hilbert-curve renumbering of a system (internal to the throughput solvers; results are permuted back).

    perm[i] = old index of new DOF i.  Built from the DOF locations that ``asm_gpu`` attaches to its
    matrices; a matrix without locations (e.g. uploaded from scipy) cannot be reordered and is solved as is."""

    def __init__(self, coords):
        torch = L.require_cuda()
        xy = coords.contiguous()
        n = xy.shape[0]
        lo, hi = xy.min(dim=0).values, xy.max(dim=0).values
        extent = float((hi - lo).max().item()) or 1.0
        keys = torch.empty(n, dtype=torch.int64, device=xy.device)
        L.check(L.lib().mwx_hilbert_keys(n, L.ptr(xy), float(lo[0].item()), float(lo[1].item()), extent, L.ptr(keys),
                                         L.stream()), 'mwx_hilbert_keys')
        order = torch.sort(keys, stable=True).indices               # library sort: setup only
        self.perm = order.to(torch.int32)
        self.iperm = torch.empty(n, dtype=torch.int32, device=xy.device)
        self.iperm[order] = torch.arange(n, dtype=torch.int32, device=xy.device)
        self.n = n
        self._patterns = {}                 # (pattern key, nnz) -> permuted pattern + value map, see matrix()

    @classmethod
    def for_matrix(cls, A):
        if getattr(A, 'coords', None) is None:
            return None
        return cls(A.coords())

    def matrix(self, A, out=None, scratch=False):
        """P A P^T as a DeviceCSR.  The first call for a sparsity pattern runs the full permutation
        (``mwx_csr_permute``: rows gathered, columns renumbered and re-sorted); from the second call on the
        permuted pattern is reused and only the values move, through a cached source-index map (one gather
        pass, ~20x cheaper).  ``out`` = (indptr, indices, values) buffers to fill; ``scratch=True`` returns the
        values in a buffer owned by this object (overwritten by the next scratch call)."""
        torch = L.require_cuda()
        dev = A.d_data.device
        key = (A.pattern_key, A.nnz)
        patterns = self.__dict__.setdefault('_patterns', {})
        ent = patterns.get(key)
        meta = {}
        for attr in ('coords', 'candidates'):                       # per-DOF metadata follows the rows
            src = getattr(A, attr, None)
            meta[attr] = None if src is None else (lambda src=src: src()[self.perm.long()])
        if ent is not None:
            if ent['vmap'] is None:                                     # second use: pay for the map once
                ids = torch.arange(A.nnz, dtype=torch.float64, device=dev)
                tmp = torch.empty(A.nnz, dtype=torch.float64, device=dev)
                L.check(L.lib().mwx_csr_permute(self.n, L.ptr(A.d_indptr), L.ptr(A.d_indices), L.ptr(ids),
                                                L.ptr(self.perm), L.ptr(self.iperm), L.ptr(ent['indptr']),
                                                L.ptr(ent['indices']), L.ptr(tmp), L.stream()), 'mwx_csr_permute')
                ent['vmap'] = tmp.to(_index_dtype(A.nnz))
                del ids, tmp
            if scratch:
                if ent['scratch'] is None:
                    ent['scratch'] = torch.empty(A.nnz, dtype=torch.float64, device=dev)
                vals = ent['scratch']
            else:
                vals = out[2] if out is not None else torch.empty(A.nnz, dtype=torch.float64, device=dev)
            _gather_values(A.d_data, ent['vmap'], vals)
            return DeviceCSR(ent['indptr'], ent['indices'], vals, A.shape, ent['key'], **meta)
        own = out is None                   # buffers allocated here: the cached pattern IS the returned one (later calls hand
        if out is None:                     # out the same indptr / indices tensors - plans keyed to them stay valid)
            out = (torch.empty(self.n + 1, dtype=torch.int64, device=dev),
                   torch.empty(A.nnz, dtype=torch.int32, device=dev),
                   torch.empty(A.nnz, dtype=torch.float64, device=dev))
        L.check(L.lib().mwx_csr_permute(self.n, L.ptr(A.d_indptr), L.ptr(A.d_indices), L.ptr(A.d_data),
                                        L.ptr(self.perm), L.ptr(self.iperm), L.ptr(out[0]), L.ptr(out[1]),
                                        L.ptr(out[2]), L.stream()), 'mwx_csr_permute')
        ent = patterns[key] = dict(indptr=out[0] if own else out[0].clone(), indices=out[1] if own else out[1].clone(),
                                   vmap=None, scratch=None, key=new_pattern_id())
        return DeviceCSR(out[0], out[1], out[2], A.shape, ent['key'], **meta)

    def _gather(self, v, idx):
        torch = L.require_cuda()
        out = torch.empty_like(v)
        L.check(L.lib().mwx_gather_f64(self.n, L.ptr(v), L.ptr(idx), L.ptr(out), L.stream()), 'mwx_gather_f64')
        return out

    def forward(self, v):          # old numbering -> new
        return self._gather(v, self.perm)

    def backward(self, v):         # new numbering -> old
        return self._gather(v, self.iperm)


class AMGHierarchy:
    """``pyamg.ruge_stuben_solver(A)`` (host C++ setup, timed) uploaded once to the device.

    mode: 'parity' = symmetric lexicographic Gauss-Seidel (pyamg's default; iteration-count parity),
          'chebyshev' = Chebyshev(degree) smoothing (fast), 'jacobi' = 2 damped Jacobi steps (fast).
    """
    MODES = {'parity': 0, 'chebyshev': 1, 'jacobi': 2, 'sa': 1}

    def __init__(self, A, mode='parity', degree=2, theta=None, max_levels=10, max_coarse=None, reorder=None,
                 upload=True, candidates=None, block_size=1, first_level=0, setup=None):
        """mode 'sa' (throughput, no reference counterpart): smoothed-aggregation hierarchy on the near-nullspace
        candidates of A (``A.candidates`` - attached by ``asm_gpu`` on a Morley basis: the interpolants of 1, x, y -
        or the ``candidates`` argument, (n, k)), Chebyshev smoothing.  The other modes use pyamg's Ruge-Stueben
        hierarchy (theta 0.25, max_coarse 10).

        setup: 'host' = C++/OpenMP setup + upload (the only choice for the Ruge-Stueben modes: its C/F splitting is a
        serial sweep whose tie-breaks the reference's iteration counts depend on); 'device' = every pass of the
        smoothed-aggregation setup as CUDA kernels, nothing leaves HBM (default for mode 'sa' when ``upload``)."""
        torch = L.require_cuda()
        if mode not in self.MODES:
            raise ValueError(f"mode must be one of {list(self.MODES)}")
        sa = mode == 'sa'
        if setup is None:
            setup = 'device' if (sa and upload) else 'host'
        if setup not in ('host', 'device') or (setup == 'device' and not sa):
            raise ValueError("setup='device' is the smoothed-aggregation setup (mode='sa')")
        theta = (0.1 if sa else 0.25) if theta is None else theta
        if max_coarse is None:
            # 'sa': the last levels of an aggregation hierarchy are small but nearly dense - a few thousand rows cost
            # 30-100 us per kernel (latency-bound) and a gamma = 2 cycle visits them 8 times per iteration, so the
            # hierarchy stops at <= 4000 rows and that level is solved by one dense GEMV with its inverse
            import os
            max_coarse = int(os.environ.get('MWX_SA_MAX_COARSE', 4000)) if sa else 10
        lib = L.lib()
        A = _to_device_csr(A)
        self.n, self.mode, self.setup = A.shape[0], mode, setup
        self._host, self._dev, self._keep = None, None, None
        # fast modes run on a locality-renumbered copy; parity mode must keep skfem's order (GS depends on it)
        if reorder is None:
            reorder = mode != 'parity'
        if reorder and mode == 'parity':
            raise ValueError("parity mode reproduces lexicographic Gauss-Seidel in skfem's numbering; reorder=False")
        cand = None
        if sa:
            cand = candidates if candidates is not None else (A.candidates() if getattr(A, 'candidates', None) else None)
            if cand is None:
                raise ValueError("mode='sa' needs near-nullspace candidates: matrices assembled by asm_gpu on a Morley "
                                 "basis carry them (also through condense); otherwise pass candidates=(n, k)")
            if isinstance(cand, np.ndarray):
                cand = torch.from_numpy(np.ascontiguousarray(cand, dtype=np.float64)).to(A.d_data.device)
            if cand.ndim != 2 or cand.shape[0] != self.n or not 1 <= cand.shape[1] <= 8:
                raise ValueError("candidates must be an (n, k) array with 1 <= k <= 8")
        t0 = time.perf_counter()
        self.reordering = Reordering.for_matrix(A) if reorder else None
        if self.reordering is not None:
            A = self.reordering.matrix(A)
            if cand is not None:
                cand = cand[self.reordering.perm.long()]
            torch.cuda.synchronize()
        self.reorder_seconds = time.perf_counter() - t0
        self.upload_seconds = 0.0
        t0 = time.perf_counter()
        if setup == 'device':
            ip = A.d_indptr if A.d_indptr.dtype == torch.int64 else A.d_indptr.to(torch.int64)
            cand = cand.contiguous().to(torch.float64)
            self._keep = (A, ip)                    # the device hierarchy borrows the fine-level arrays (candidates are copied)
            self._dev = ctypes.c_void_p()
            L.check(lib.mwx_amg_setup_sa_dev(self.n, L.ptr(ip), L.ptr(A.d_indices), L.ptr(A.d_data), L.ptr(cand),
                                             cand.shape[1], int(block_size), int(first_level), float(theta),
                                             int(max_levels), int(max_coarse),
                                             self.MODES[mode], int(degree), ctypes.byref(self._dev), L.stream()),
                    'mwx_amg_setup_sa_dev')
            self.levels = lib.mwx_amg_dev_num_levels(self._dev)
            if upload:                      # upload=False: operators only (the partitioned solver slices them itself)
                self._check_coarsest()
                pinv = self._pinv_coarse(self.level_matrix(self.levels - 1, 0).toarray())
                L.check(lib.mwx_amg_dev_finalize(self._dev, L.ptr(pinv), L.stream()), 'mwx_amg_dev_finalize')
                self._default_cycle()
                self._default_storage()
            torch.cuda.synchronize()
            self.setup_host_seconds = 0.0
            self.setup_device_seconds = time.perf_counter() - t0
            return
        ip = np.ascontiguousarray(A.d_indptr.cpu().numpy(), dtype=np.int64)
        ix = np.ascontiguousarray(A.d_indices.cpu().numpy(), dtype=np.int32)
        va = np.ascontiguousarray(A.d_data.cpu().numpy(), dtype=np.float64)
        self._host = ctypes.c_void_p()
        import os
        lib.mwx_set_host_threads(int(os.environ.get('MWX_HOST_THREADS', os.cpu_count() or 1)))
        if sa:
            Bh = np.ascontiguousarray(cand.cpu().numpy(), dtype=np.float64)
            L.check(lib.mwx_amg_setup_sa_host(self.n, L.ptr(ip), L.ptr(ix), L.ptr(va), L.ptr(Bh), Bh.shape[1],
                                              int(block_size), int(first_level), float(theta), int(max_levels),
                                              int(max_coarse), ctypes.byref(self._host)), 'mwx_amg_setup_sa_host')
        else:
            L.check(lib.mwx_amg_setup_host(self.n, L.ptr(ip), L.ptr(ix), L.ptr(va), float(theta), int(max_levels),
                                           int(max_coarse), ctypes.byref(self._host)), 'mwx_amg_setup_host')
        self.levels = lib.mwx_amg_num_levels(self._host)
        self.setup_host_seconds = time.perf_counter() - t0
        self.setup_device_seconds = 0.0
        if not upload:                      # host hierarchy only (the partitioned solver distributes it itself)
            return
        t1 = time.perf_counter()
        # coarsest level: dense pseudo-inverse (scipy 1.4.1 pinv2 semantics) through LAPACK
        self._check_coarsest()
        Ac = self.level_matrix(self.levels - 1, 0).toarray()
        pinv = self._pinv_coarse(Ac) if sa else self._pinv2(Ac)
        self._dev = ctypes.c_void_p()
        L.check(lib.mwx_amg_upload(self._host, self.MODES[mode], int(degree), L.ptr(pinv),
                                   ctypes.byref(self._dev), L.stream()), 'mwx_amg_upload')
        if sa:
            self._default_storage()
        torch.cuda.synchronize()
        self.upload_seconds = time.perf_counter() - t1

    storage_bits = 64

    def _default_storage(self):
        """Mode 'sa': the block-CSR operators of the hierarchy (every A_l, P_l, R_l below the fine matrix) are STORED in
        fp32 - values only; vectors, products and sums stay fp64 - which halves the bytes those levels read (level-1
        smoother step 1.11 -> 0.69 ms at 67 M DOFs).  The preconditioner is the exact V-cycle of the rounded hierarchy,
        still a fixed symmetric operator; iteration counts are unchanged.  MWX_SA_FP32=0 keeps fp64."""
        import os
        if os.environ.get('MWX_SA_FP32', '1') != '0':
            L.check(L.lib().mwx_amg_set_storage(self._dev, 32, L.stream()), 'mwx_amg_set_storage')
            self.storage_bits = 32

    def _default_cycle(self):
        """Mode 'sa': levels 2.. with at least 10 000 rows are visited twice per visit of their parent (gamma = 2 there, a
        V-cycle above and below).  Measured at 67 M DOFs (profiles/r02_gamma_cycle_*.json): 25 instead of 39 iterations for
        17 % more time per iteration.  Level 1 is too expensive to revisit, the tiny levels are launch-latency bound.
        MWX_SA_CYCLE=v keeps the plain V-cycle."""
        import os
        if os.environ.get('MWX_SA_CYCLE', 'auto') == 'v':
            return
        sizes = self.level_sizes()
        min_rows = int(os.environ.get('MWX_SA_CYCLE_MIN_ROWS', 10000))
        last = max([l for l in range(2, self.levels - 1) if sizes[l] >= min_rows], default=0)
        if last >= 2:
            self.set_cycle(2, last, first=2)

    def _check_coarsest(self):
        if self.level_dims(self.levels - 1, 0)[0] > 8192:
            raise RuntimeError(f"AMG coarsening stopped at {self.level_dims(self.levels - 1, 0)[0]} rows (levels "
                               f"{self.level_sizes()}): too large for the dense coarse solve; raise max_levels or "
                               f"lower theta")

    @staticmethod
    def _pinv2(Ac):
        """scipy 1.4.1 ``pinv2`` (pyamg's coarse solver): SVD with the cutoff 1e6 * eps * s_max."""
        u, s, vh = np.linalg.svd(Ac, full_matrices=False)
        rank = int(np.sum(s > 1e6 * np.finfo(np.float64).eps * s.max())) if s.size else 0
        return np.ascontiguousarray(((u[:, :rank] / s[:rank]) @ vh[:rank]).T)

    @classmethod
    def _pinv_coarse(cls, Ac):
        """Inverse of the coarsest operator of a smoothed-aggregation hierarchy (symmetric positive definite: Galerkin
        products of an SPD matrix with full-rank prolongators, unit diagonal for dead coarse DOFs).  Up to 300 rows: the
        SVD of ``_pinv2``; larger: Cholesky on the device (setup only), symmetric eigendecomposition with ``_pinv2``'s
        relative cutoff if the factorisation breaks down.  (Host LAPACK dpotrf + dpotri was measured instead: 0.6 s at
        3400 rows on 16 cores against 0.1 s here - but the first cuSOLVER call of a process costs 1-2 s of library load.)"""
        n = Ac.shape[0]
        if n <= 300:
            return cls._pinv2(Ac)
        torch = L.require_cuda()
        Ad = torch.from_numpy(np.ascontiguousarray(Ac, dtype=np.float64)).cuda()
        Ad = 0.5 * (Ad + Ad.T)
        Lc, info = torch.linalg.cholesky_ex(Ad)
        if int(info.item()) == 0:
            inv = torch.cholesky_inverse(Lc)
            if bool(torch.isfinite(inv).all()):
                return np.ascontiguousarray((0.5 * (inv + inv.T)).cpu().numpy())
        w, V = torch.linalg.eigh(Ad)
        keep = w.abs() > 1e6 * np.finfo(np.float64).eps * w.abs().max()
        winv = torch.where(keep, 1.0 / torch.where(keep, w, torch.ones_like(w)), torch.zeros_like(w))
        return np.ascontiguousarray(((V * winv) @ V.T).cpu().numpy())

    def set_chebyshev(self, degree=2, ratio=10.0):
        L.check(L.lib().mwx_amg_set_chebyshev(self._dev, int(degree), float(ratio)), 'mwx_amg_set_chebyshev')

    def set_cycle(self, gamma=1, last=0, first=1):
        """gamma = 2: levels ``first``..``last`` are visited twice per visit of their parent (still an SPD preconditioner)."""
        L.check(L.lib().mwx_amg_set_cycle(self._dev, int(gamma), int(first), int(last)), 'mwx_amg_set_cycle')

    def level_dims(self, level, which=0):
        dims = np.zeros(3, dtype=np.int64)
        if self._host is None:
            L.check(L.lib().mwx_amg_dev_level_dims(self._dev, level, which, L.ptr(dims)), 'mwx_amg_dev_level_dims')
        else:
            L.check(L.lib().mwx_amg_level_dims(self._host, level, which, L.ptr(dims)), 'mwx_amg_level_dims')
        return tuple(int(v) for v in dims)

    def level_matrix(self, level, which=0):
        """scipy copy of A_l (which=0), P_l (1) or R_l (2) - for tests and diagnostics."""
        import scipy.sparse as sp
        rows, cols, nnz = self.level_dims(level, which)
        ip = np.empty(rows + 1, np.int64); ix = np.empty(nnz, np.int32); va = np.empty(nnz, np.float64)
        if self._host is None:
            L.check(L.lib().mwx_amg_dev_level_copy(self._dev, level, which, L.ptr(ip), L.ptr(ix), L.ptr(va), L.stream()),
                    'mwx_amg_dev_level_copy')
        else:
            L.check(L.lib().mwx_amg_level_copy(self._host, level, which, L.ptr(ip), L.ptr(ix), L.ptr(va)),
                    'mwx_amg_level_copy')
        return sp.csr_matrix((va, ix, ip), shape=(rows, cols))

    def level_tensors(self, level, which=0):
        """(indptr, indices, data) of A_l / P_l / R_l as device tensors (copies)."""
        torch = L.require_cuda()
        rows, cols, nnz = self.level_dims(level, which)
        dev = torch.device('cuda', torch.cuda.current_device())
        if self._host is None:
            ip = torch.empty(rows + 1, dtype=torch.int64, device=dev)
            ix = torch.empty(nnz, dtype=torch.int32, device=dev)
            va = torch.empty(nnz, dtype=torch.float64, device=dev)
            L.check(L.lib().mwx_amg_dev_level_copy(self._dev, level, which, L.ptr(ip), L.ptr(ix), L.ptr(va), L.stream()),
                    'mwx_amg_dev_level_copy')
            return ip, ix, va
        M = self.level_matrix(level, which)
        return (torch.from_numpy(M.indptr.astype(np.int64)).to(dev), torch.from_numpy(M.indices.astype(np.int32)).to(dev),
                torch.from_numpy(M.data.astype(np.float64)).to(dev))

    def splitting(self, level):
        rows = self.level_dims(level, 0)[0]
        s = np.empty(rows, np.int32)
        L.check(L.lib().mwx_amg_level_splitting(self._host, level, L.ptr(s)), 'mwx_amg_level_splitting')
        return s

    def level_candidates(self, level):
        """(rows, k) near-nullspace candidates of a smoothed-aggregation level."""
        k = ctypes.c_int32(0)
        if self._host is None:
            L.check(L.lib().mwx_amg_dev_level_aux(self._dev, level, None, None, None, None, ctypes.byref(k), L.stream()),
                    'mwx_amg_dev_level_aux')
            B = np.empty((self.level_dims(level, 0)[0], k.value), dtype=np.float64)
            L.check(L.lib().mwx_amg_dev_level_aux(self._dev, level, None, None, None, L.ptr(B), None, L.stream()),
                    'mwx_amg_dev_level_aux')
            return B
        L.check(L.lib().mwx_amg_level_candidates(self._host, level, None, ctypes.byref(k)), 'mwx_amg_level_candidates')
        B = np.empty((self.level_dims(level, 0)[0], k.value), dtype=np.float64)
        L.check(L.lib().mwx_amg_level_candidates(self._host, level, L.ptr(B), ctypes.byref(k)), 'mwx_amg_level_candidates')
        return B

    def coarse_offsets(self, level, offsets):
        """Row-block boundaries of level+1 induced by the boundaries ``offsets`` of ``level``: a coarse DOF goes with
        the fine DOF it comes from (its C point, or the seed of its aggregate), so the blocks stay contiguous."""
        k = ctypes.c_int32(0)
        if self._host is None:
            na = ctypes.c_int64(0)
            L.check(L.lib().mwx_amg_dev_level_aux(self._dev, level, None, ctypes.byref(na), ctypes.byref(k), None, None,
                                                  L.stream()), 'mwx_amg_dev_level_aux')
            roots = np.empty(na.value, dtype=np.int64)
            L.check(L.lib().mwx_amg_dev_level_aux(self._dev, level, L.ptr(roots), None, None, None, None, L.stream()),
                    'mwx_amg_dev_level_aux')
            return [int(k.value * np.searchsorted(roots, o, side='left')) for o in offsets]
        L.check(L.lib().mwx_amg_level_roots(self._host, level, None, ctypes.byref(k)), 'mwx_amg_level_roots')
        if k.value == 0:                                   # Ruge-Stueben level
            csum = np.concatenate(([0], np.cumsum(self.splitting(level))))
            return [int(csum[o]) for o in offsets]
        na = self.level_dims(level, 1)[1] // k.value
        roots = np.empty(na, dtype=np.int64)
        L.check(L.lib().mwx_amg_level_roots(self._host, level, L.ptr(roots), ctypes.byref(k)), 'mwx_amg_level_roots')
        return [int(k.value * np.searchsorted(roots, o, side='left')) for o in offsets]

    def level_sizes(self):
        return [self.level_dims(l, 0)[0] for l in range(self.levels)]

    def operator_complexity(self):
        return sum(self.level_dims(l, 0)[2] for l in range(self.levels)) / self.level_dims(0, 0)[2]

    def precondition(self, r):
        """z = M^-1 r (one V(1,1) cycle) - ``ml.aspreconditioner()``."""
        torch = L.require_cuda()
        is_np = isinstance(r, np.ndarray)
        rd = _to_device_vec(r, torch.device('cuda', torch.cuda.current_device()))
        z = torch.empty_like(rd)
        L.check(L.lib().mwx_amg_vcycle(self._dev, L.ptr(rd), L.ptr(z), L.stream()), 'mwx_amg_vcycle')
        return z.cpu().numpy() if is_np else z

    def __del__(self):
        try:
            lib = L.lib()
            if getattr(self, '_dev', None):
                lib.mwx_amg_destroy_dev(self._dev)
            if getattr(self, '_host', None):
                lib.mwx_amg_destroy_host(self._host)
        except Exception:
            pass


def pcg(A, b, tol=1e-5, maxiter=None, precondition=True, amg=None, reordering=None, max_rechecks=0, monitor=None,
        history=None):
    """Device PCG with scipy-1.4.1 semantics on a DeviceCSR / scipy matrix.  ``amg`` = a prebuilt
    :class:`AMGHierarchy` (hierarchy reuse across solves: north_star's "built once as timed setup"), else Jacobi /
    plain CG.  ``b`` numpy or device tensor.  Returns (x as a device tensor, iters, info, resid).
    ``history`` = a list: receives ||x_k|| of every iteration (what the reference's ``verbose=True`` callback prints)."""
    torch = L.require_cuda()
    A = _to_device_csr(A)
    n = A.shape[0]
    dev = A.d_data.device
    bd = _to_device_vec(b, dev)
    ro = reordering if reordering is not None else (amg.reordering if amg is not None else None)
    if ro is not None:
        A = ro.matrix(A, scratch=True)
        bd = ro.forward(bd)
    x = torch.empty(n, dtype=torch.float64, device=dev)
    work = torch.empty(4 * n, dtype=torch.float64, device=dev)
    iters, info, resid = ctypes.c_int64(0), ctypes.c_int32(0), ctypes.c_double(0.0)
    mi = 0 if maxiter is None else int(maxiter)
    hist = None
    if history is not None:
        hist = np.zeros(max(1, min(mi if mi > 0 else 10 * n, 1000000)), dtype=np.float64)
    if hist is not None and amg is None:
        rc = L.lib().mwx_pcg_jacobi_history(n, L.ptr(A.d_indptr), L.ptr(A.d_indices), L.ptr(A.d_data), L.ptr(bd), L.ptr(x),
                                            float(tol), mi, 1 if precondition else 0, L.ptr(work), ctypes.byref(iters),
                                            ctypes.byref(info), ctypes.byref(resid), L.ptr(hist), hist.size, L.stream())
    elif hist is not None:
        rc = L.lib().mwx_pcg_amg_history(amg._dev, n, L.ptr(A.d_indptr), L.ptr(A.d_indices), L.ptr(A.d_data), L.ptr(bd),
                                         L.ptr(x), float(tol), mi, L.ptr(work), ctypes.byref(iters), ctypes.byref(info),
                                         ctypes.byref(resid), L.ptr(hist), hist.size, L.stream())
    elif amg is None:
        rc = L.lib().mwx_pcg_jacobi(n, L.ptr(A.d_indptr), L.ptr(A.d_indices), L.ptr(A.d_data), L.ptr(bd), L.ptr(x),
                                    float(tol), mi, 1 if precondition else 0, L.ptr(work), ctypes.byref(iters),
                                    ctypes.byref(info), ctypes.byref(resid), L.stream())
    elif max_rechecks or monitor is not None:
        # diagnostics beyond scipy's cg (see mwx_pcg_amg_monitored): first iteration that met the tolerance on the
        # recurrence residual, optional stop after `max_rechecks` failed true-residual re-checks
        first = ctypes.c_int64(0)
        rc = L.lib().mwx_pcg_amg_monitored(amg._dev, n, L.ptr(A.d_indptr), L.ptr(A.d_indices), L.ptr(A.d_data), L.ptr(bd),
                                           L.ptr(x), float(tol), mi, int(max_rechecks), L.ptr(work), ctypes.byref(iters),
                                           ctypes.byref(info), ctypes.byref(resid), ctypes.byref(first), L.stream())
        if monitor is not None:
            monitor['first_hit'] = int(first.value)
    else:
        rc = L.lib().mwx_pcg_amg(amg._dev, n, L.ptr(A.d_indptr), L.ptr(A.d_indices), L.ptr(A.d_data), L.ptr(bd),
                                 L.ptr(x), float(tol), mi, L.ptr(work), ctypes.byref(iters), ctypes.byref(info),
                                 ctypes.byref(resid), L.stream())
    if rc not in (0, -3):                       # -3 = MWX_E_NOCONV (breakdown): reported through info
        L.check(rc, 'mwx_pcg')
    if ro is not None:
        x = ro.backward(x)
    if hist is not None:
        history.extend(hist[:min(int(iters.value), hist.size)].tolist())
    return x, int(iters.value), int(info.value), float(resid.value)


_run_pcg = pcg


def residual_dd(A, x, b):
    """r = b - A x on the device with double-double row accumulation, rounded once (``mwx_residual_dd``)."""
    torch = L.require_cuda()
    A = _to_device_csr(A)
    ip = A.d_indptr if A.d_indptr.dtype == torch.int64 else A.d_indptr.to(torch.int64)
    r = torch.empty_like(x)
    L.check(L.lib().mwx_residual_dd(A.shape[0], L.ptr(ip), L.ptr(A.d_indices), L.ptr(A.d_data), L.ptr(x), L.ptr(b), L.ptr(r),
                                    L.stream()), 'mwx_residual_dd')
    return r


def refine(A, b, x, steps=2, tol=1e-8, maxiter=None, amg=None, log=None):
    """Mixed-precision iterative refinement of a PCG solution (an extension; the reference has none):
    repeat ``steps`` times  r = b - A x in double-double,  d = PCG(A, r / ||r||) * ||r||,  x += d.
    With the residual at twice the working precision the forward error goes to fp64 round-off level instead of
    stalling at cond(A) * eps (1e-6 at 67 M DOFs), as long as cond(A) * eps < 1.  x, b: device tensors.
    ``log`` = a list receiving one dict per step (exact relative residual before the step, inner iterations)."""
    torch = L.require_cuda()
    A = _to_device_csr(A)
    bn = float(torch.linalg.norm(b))
    x = x.clone()
    for _ in range(int(steps)):
        r = residual_dd(A, x, b)
        rn = float(torch.linalg.norm(r))
        entry = {'rel_residual_dd': rn / bn if bn > 0 else rn}
        if log is not None:
            log.append(entry)
        if not rn > 0.0:
            break
        d, its, info, _ = pcg(A, r / rn, tol, maxiter, amg=amg, max_rechecks=2)    # unit right-hand side: scipy's absolute
        x += rn * d                                                                 # early exit (||b|| <= tol) must not fire
        entry.update(inner_iters=its, inner_info=info)
    if log is not None:
        rn = float(torch.linalg.norm(residual_dd(A, x, b)))
        log.append({'rel_residual_dd': rn / bn if bn > 0 else rn})
    return x


def _reject_host_callables(kwargs):
    if kwargs.get('M') is not None:
        raise NotImplementedError("a user-supplied preconditioner M is a host callable; the device path "
                                  "offers Precondition=True (Jacobi) or solver_iter_mgcg (AMG). No CPU fallback.")


def _finish(x, b, info, verbose, name, kwargs):
    if info > 0:
        warnings.warn("Convergence not achieved!")
    elif info == 0 and verbose:
        print(f"{name} converged to tol={kwargs.get('tol', 'default')} and atol={kwargs.get('atol', 'default')}")
    return x.cpu().numpy() if isinstance(b, np.ndarray) else x


# ------------------------------------------------------------------ factories (reference call shapes)
def build_pc_diag(A):
    """Diagonal preconditioner (ref utils.py:48-50): returns 1/diag(A) as a vector."""
    return 1.0 / _to_device_csr(A).diagonal()


def _make_krylov(print_iters, krylov=None, verbose=False, **kwargs):
    if krylov is not None and getattr(krylov, '__name__', 'cg') != 'cg':
        raise NotImplementedError("only conjugate gradients (scipy's cg) has a device implementation")

    def run(A, b, **solve_time_kwargs):
        kwargs.update(solve_time_kwargs)
        kw = dict(kwargs)
        pre = kw.pop('Precondition', False) == True   # noqa: E712 - the reference's own test (utils.py:283)
        reorder = bool(kw.pop('reorder', False))
        _reject_host_callables(kw)
        t0 = time.perf_counter()
        A = _to_device_csr(A)
        ro = Reordering.for_matrix(A) if reorder else None
        hist = [] if verbose else None             # verbose: the reference's callback prints ||x_k|| every iteration
        x, it, info, resid = _run_pcg(A, b, kw.get('tol', 1e-5), kw.get('maxiter'), precondition=pre, reordering=ro,
                                      history=hist)
        for v in hist or ():
            print(v)
        solver.last = dict(iters=it, info=info, resid=resid, solve_seconds=time.perf_counter() - t0)
        if print_iters:
            print('cg', 'total interation steps:', it)
        return _finish(x, b, info, verbose, 'cg', kw), it

    def solver(A, b, **kw):
        return run(A, b, **kw)[0]

    solver.solve_with_info = run
    solver.last = None
    return solver


def solver_iter_krylov(krylov=None, verbose=False, **kwargs):
    """ref utils.py:265-316.  ``Precondition=True`` -> Jacobi (``build_pc_diag``) preconditioned CG."""
    return _make_krylov(False, krylov, verbose, **kwargs)


def solver_iter_krylov_iter(krylov=None, verbose=False, **kwargs):
    """ref utils.py:212-263: same, and prints ``cg total interation steps: N``."""
    return _make_krylov(True, krylov, verbose, **kwargs)


def solver_iter_pcg(**kwargs):
    """ref utils.py:318-320."""
    return solver_iter_krylov(**kwargs)


def _make_mgcg(print_iters, krylov=None, verbose=False, mode='parity', degree=2, **kwargs):
    if krylov is not None and getattr(krylov, '__name__', 'cg') != 'cg':
        raise NotImplementedError("only conjugate gradients (scipy's cg) has a device implementation")

    def run(A, b, **solve_time_kwargs):
        kwargs.update(solve_time_kwargs)
        kw = dict(kwargs)
        _reject_host_callables(kw)
        A_d = _to_device_csr(A)
        t0 = time.perf_counter()
        ml = AMGHierarchy(A_d, mode=mode, degree=degree)          # rebuilt per solve, like the reference
        t1 = time.perf_counter()
        hist = [] if verbose else None             # verbose: the reference's callback prints ||x_k|| every iteration
        x, it, info, resid = _run_pcg(A_d, b, kw.get('tol', 1e-5), kw.get('maxiter'), amg=ml, history=hist)
        for v in hist or ():
            print(v)
        solver.last = dict(iters=it, info=info, resid=resid, setup_seconds=t1 - t0,
                           solve_seconds=time.perf_counter() - t1, levels=ml.level_sizes(), mode=mode)
        if print_iters:
            print('mgcg total interation steps:', it)
        return _finish(x, b, info, verbose, 'cg', kw), it

    def solver(A, b, **kw):
        return run(A, b, **kw)[0]

    solver.solve_with_info = run
    solver.last = None
    return solver


def solver_iter_mgcg(krylov=None, verbose=False, mode='parity', degree=2, **kwargs):
    """ref utils.py:127-166: RS-AMG V-cycle preconditioned CG.  ``mode`` selects the smoother
    ('parity' = pyamg's symmetric Gauss-Seidel, 'chebyshev' / 'jacobi' = fast device smoothers)."""
    return _make_mgcg(False, krylov, verbose, mode, degree, **kwargs)


def solver_iter_mgcg_iter(krylov=None, verbose=False, mode='parity', degree=2, **kwargs):
    """ref utils.py:87-125: same, and prints ``mgcg total interation steps: N``."""
    return _make_mgcg(True, krylov, verbose, mode, degree, **kwargs)


def solver_iter_pyamg(verbose=False, mode='parity', **kwargs):
    """ref utils.py:168-210: stand-alone AMG iteration ``ml.solve(b, tol=1e-10, maxiter=10000)``.
    A V-cycle with initial guess x equals x + M(b - A x) with M the cycle from zero, so the iteration is
    r = b - A x;  x += V(r);  stop when ||r|| / ||b|| < tol (pyamg's test, after each cycle)."""
    def run(A, b, **solve_time_kwargs):
        torch = L.require_cuda()
        kwargs.update(solve_time_kwargs)
        A_d = _to_device_csr(A)
        dev = A_d.d_data.device
        ml = AMGHierarchy(A_d, mode=mode)
        Ap = ml.reordering.matrix(A_d) if ml.reordering is not None else A_d
        bd = _to_device_vec(b, dev)
        if ml.reordering is not None:
            bd = ml.reordering.forward(bd)
        x = torch.zeros_like(bd)
        normb = float(torch.linalg.norm(bd)) or 1.0
        it, resid = 0, normb
        while it < 10000:
            r = bd - Ap.dot(x)
            x = x + ml.precondition(r)
            it += 1
            resid = float(torch.linalg.norm(bd - Ap.dot(x)))
            if verbose:
                print(float(torch.linalg.norm(x)))
            if resid / normb < 1e-10:
                break
        if ml.reordering is not None:
            x = ml.reordering.backward(x)
        solver.last = dict(iters=it, resid=resid)
        if resid > 1 and verbose:
            print('Warning: residual norm =', resid)
        return (x.cpu().numpy() if isinstance(b, np.ndarray) else x), it

    def solver(A, b, **kw):
        return run(A, b, **kw)[0]

    solver.solve_with_info = run
    solver.last = None
    return solver


# ------------------------------------------------------------------ solve / condense
def solve(A, b, x=None, I=None, solver=None, **kwargs):
    """ref utils.py:326-368 (linear systems only)."""
    if solver is None:
        raise NotImplementedError("the device path has no direct solver; pass solver_iter_pcg/mgcg (no CPU fallback)")
    if x is not None and I is not None:
        y = x.copy()
        y[I] = solver(A, b, **kwargs)
        return y
    return solver(A, b, **kwargs)


def _flatten_dofs(S):
    if S is None:
        return None
    if isinstance(S, np.ndarray):
        return S
    if hasattr(S, 'flatten') and not isinstance(S, dict):
        return S.flatten()
    if isinstance(S, dict):
        return np.unique(np.concatenate([S[k].flatten() for k in S]))
    raise NotImplementedError("Unable to flatten the given set of DOFs.")


class CondensePlan:
    """Reusable elimination of a fixed Dirichlet set from matrices sharing one sparsity pattern.

    The count pass of ``condense`` (kept-row numbering, condensed indptr) depends only on the pattern and
    on D, so an epsilon sweep or a re-assembly on the same mesh pays for it once."""

    def __init__(self, A, I=None, D=None):
        torch = L.require_cuda()
        A = _to_device_csr(A)
        n = A.shape[0]
        if I is None and D is None:
            raise Exception("Either I or D must be given!")
        if I is not None and D is not None:
            raise Exception("Give only I or only D!")
        keep_h = np.zeros(n, dtype=np.int32)
        if I is None:
            keep_h[:] = 1
            keep_h[D] = 0
        else:
            # The kernels keep rows in ascending DOF order.  The reference's A[I].T[I].T follows the order of I; every
            # call site passes sorted, duplicate-free sets, so I is normalised here and the normalised set is what
            # condense() returns (an unsorted I would otherwise scatter the solution to the wrong DOFs in solve()).
            keep_h[np.asarray(I, dtype=np.int64)] = 1
        I = np.nonzero(keep_h)[0]
        self.I, self.n, self.pattern_key, self.out_key = I, n, A.pattern_key, new_pattern_id()
        self._I_dev = None
        dev = A.d_data.device
        self.keep = torch.from_numpy(keep_h).to(dev)
        self.newid = torch.empty(n + 1, dtype=torch.int32, device=dev)
        indptr_out = torch.empty(n + 1, dtype=torch.int64, device=dev)
        n_out, nnz_out = ctypes.c_int64(0), ctypes.c_int64(0)
        L.check(L.lib().mwx_condense_count(n, L.ptr(A.d_indptr), L.ptr(A.d_indices), L.ptr(self.keep),
                                           L.ptr(self.newid), L.ptr(indptr_out), ctypes.byref(n_out),
                                           ctypes.byref(nnz_out), L.stream()), 'mwx_condense_count')
        self.n_out, self.nnz_out = int(n_out.value), int(nnz_out.value)
        self.indptr_out = indptr_out[:self.n_out + 1].clone()
        self.indices_out = torch.empty(self.nnz_out, dtype=torch.int32, device=dev)
        self._matrix_applies, self._vmap = 0, None      # matrix-only applies: values move through a cached map

    def apply(self, A, b=None, x=None, values_out=None):
        """-> (Aout DeviceCSR, bout or None); b/x may be numpy or device tensors."""
        torch = L.require_cuda()
        A = _to_device_csr(A)
        if A.pattern_key != self.pattern_key:
            raise ValueError("CondensePlan was built for a different sparsity pattern")
        dev = A.d_data.device
        vals = values_out if values_out is not None else torch.empty(self.nnz_out, dtype=torch.float64, device=dev)
        have_b = b is not None
        if not have_b:
            # Matrix only (re-assembly on the same mesh, eps sweep): from the second time on just gather the kept
            # values through a source-index map instead of re-walking the pattern.
            self._matrix_applies += 1
            if self._matrix_applies >= 2 and self._vmap is None:
                ids = torch.arange(A.nnz, dtype=torch.float64, device=dev)
                tmp = torch.empty(self.nnz_out, dtype=torch.float64, device=dev)
                L.check(L.lib().mwx_condense_fill(self.n, L.ptr(A.d_indptr), L.ptr(A.d_indices), L.ptr(ids),
                                                  L.ptr(self.keep), L.ptr(self.newid), L.ptr(self.indptr_out),
                                                  L.ptr(self.indices_out), L.ptr(tmp), None, None, None,
                                                  L.stream()), 'mwx_condense_fill')
                self._vmap = tmp.to(_index_dtype(A.nnz))
                del ids, tmp
            if self._vmap is not None:
                _gather_values(A.d_data, self._vmap, vals)
                return DeviceCSR(self.indptr_out, self.indices_out, vals, (self.n_out, self.n_out), self.out_key,
                                 coords=self._rows_of(A, 'coords', dev),
                                 candidates=self._rows_of(A, 'candidates', dev)), None
        bd = _to_device_vec(b, dev) if have_b else None
        xd = _to_device_vec(x, dev) if (have_b and x is not None) else None
        b_out = torch.empty(self.n_out, dtype=torch.float64, device=dev) if have_b else None
        L.check(L.lib().mwx_condense_fill(self.n, L.ptr(A.d_indptr), L.ptr(A.d_indices), L.ptr(A.d_data),
                                          L.ptr(self.keep), L.ptr(self.newid), L.ptr(self.indptr_out),
                                          L.ptr(self.indices_out), L.ptr(vals), L.ptr(bd), L.ptr(xd), L.ptr(b_out),
                                          L.stream()), 'mwx_condense_fill')
        Aout = DeviceCSR(self.indptr_out, self.indices_out, vals, (self.n_out, self.n_out), self.out_key,
                         coords=self._rows_of(A, 'coords', dev), candidates=self._rows_of(A, 'candidates', dev))
        return Aout, b_out

    def _rows_of(self, A, attr, dev):
        """Per-DOF metadata of A (DOF locations, near-nullspace candidates) restricted to the kept rows."""
        if getattr(A, attr, None) is None:
            return None
        torch = L.require_cuda()
        src, plan = getattr(A, attr), self

        def rows():
            if plan._I_dev is None:
                plan._I_dev = torch.from_numpy(np.ascontiguousarray(plan.I, dtype=np.int64)).to(dev)
            return src()[plan._I_dev]
        return rows


def condense(A, b=None, x=None, I=None, D=None, expand=True):
    """ref utils.py:384-460 on the device: ``Aout = A[I].T[I].T``, ``bout = b[I] - A[I].T[D].T @ x[D]``.
    ``A`` may be a DeviceCSR (stays in HBM) or a scipy matrix (uploaded).  Divergence from the reference: a given
    ``I`` is normalised to ascending, duplicate-free order (``np.unique``) and that normalised set is returned."""
    D, I = _flatten_dofs(D), _flatten_dofs(I)
    A = _to_device_csr(A)
    if b is not None and not isinstance(b, np.ndarray) and not hasattr(b, 'data_ptr'):
        raise NotImplementedError("condense: generalized eigenproblems (matrix b) are not on the device path")
    if x is None:
        x = np.zeros(A.shape[0])
    plan = CondensePlan(A, I=I, D=D)
    Aout, b_out = plan.apply(A, b, x if b is not None else None)
    ret = (Aout,) if b is None else (Aout, b_out.cpu().numpy() if isinstance(b, np.ndarray) else b_out)
    if expand:
        ret += (x, plan.I)
    return ret if len(ret) > 1 else ret[0]
