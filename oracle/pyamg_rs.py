"""Oracle (test infrastructure): restatement of pyamg 4.0.0 ``ruge_stuben_solver`` + V-cycle.

Stands in for ``ml = pyamg.ruge_stuben_solver(A); M = ml.aspreconditioner()``
(ref main/example1/utils.py:153-155) and ``ml.solve(b, tol=1e-10, maxiter=10000)``
(ref main/example1/utils.py:193-196).  pyamg is un-vendored (pin ref README.md:18); defaults
restated per SURVEY.md B.1-B.4: strength ('classical', theta=0.25), CF='RS' (first pass),
direct interpolation, Galerkin RAP, max_levels=10, max_coarse=10, symmetric Gauss-Seidel pre/post
smoothing (1 iteration each), dense pseudo-inverse (scipy 1.4.1 ``pinv2``) on the coarsest level.
The serial loops are in ``amg_core_oracle.c`` (loaded through ctypes).
"""
import ctypes
import os
import subprocess

import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, '_build', 'liboracle_amg.so')
_lib = None

_I = ctypes.c_int64
_pI = np.ctypeslib.ndpointer(dtype=np.int64, flags='C_CONTIGUOUS')
_pD = np.ctypeslib.ndpointer(dtype=np.float64, flags='C_CONTIGUOUS')


def build():
    src = os.path.join(_HERE, 'amg_core_oracle.c')
    if (not os.path.exists(_SO)) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(['make', '-C', _HERE, '-s'])
    return _SO


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        L.oracle_strength_abs.restype = _I
        L.oracle_strength_abs.argtypes = [_I, ctypes.c_double, _pI, _pI, _pD, _pI, _pI, _pD]
        L.oracle_rs_cf_splitting.restype = None
        L.oracle_rs_cf_splitting.argtypes = [_I, _pI, _pI, _pI, _pI, _pI]
        L.oracle_direct_interp_pass1.restype = None
        L.oracle_direct_interp_pass1.argtypes = [_I, _pI, _pI, _pI, _pI]
        L.oracle_direct_interp_pass2.restype = _I
        L.oracle_direct_interp_pass2.argtypes = [_I, _pI, _pI, _pD, _pI, _pI, _pD, _pI, _pI, _pI, _pD]
        L.oracle_gauss_seidel.restype = None
        L.oracle_gauss_seidel.argtypes = [_pI, _pI, _pD, _pD, _pD, _I, _I, _I]
        L.oracle_csr_matvec.restype = None
        L.oracle_csr_matvec.argtypes = [_I, _pI, _pI, _pD, _pD, _pD]
        _lib = L
    return _lib


def _csr64(A):
    A = sp.csr_matrix(A)
    if not A.has_sorted_indices:
        A = A.sorted_indices()
    return (np.ascontiguousarray(A.indptr, dtype=np.int64),
            np.ascontiguousarray(A.indices, dtype=np.int64),
            np.ascontiguousarray(A.data, dtype=np.float64))


def pinv2(a):
    """scipy 1.4.1 ``linalg.pinv2``: SVD pseudo-inverse, cutoff 1e6*eps*max(s) for float64."""
    u, s, vh = np.linalg.svd(a, full_matrices=False)
    cond = 1e6 * np.finfo(np.float64).eps
    rank = int(np.sum(s > cond * np.max(s))) if s.size else 0
    u = u[:, :rank] / s[:rank]
    return np.conjugate(u @ vh[:rank]).T


def strength_abs(A, theta=0.25):
    Ap, Aj, Ax = _csr64(A)
    n = A.shape[0]
    Sp = np.empty(n + 1, np.int64); Sj = np.empty(len(Aj), np.int64); Sx = np.empty(len(Ax))
    nnz = lib().oracle_strength_abs(n, theta, Ap, Aj, Ax, Sp, Sj, Sx)
    return sp.csr_matrix((Sx[:nnz].copy(), Sj[:nnz].copy(), Sp), shape=A.shape)


def rs_splitting(S):
    """``pyamg.classical.split.RS(S)``: 1 = C point, 0 = F point."""
    C = S.tocoo()
    keep = C.row != C.col                                   # remove_diagonal(S)
    S0 = sp.csr_matrix((C.data[keep], (C.row[keep], C.col[keep])), shape=S.shape)
    T = S0.T.tocsr()
    Sp, Sj, _ = _csr64(S0); Tp, Tj, _ = _csr64(T)
    split = np.empty(S.shape[0], np.int64)
    lib().oracle_rs_cf_splitting(S.shape[0], Sp, Sj, Tp, Tj, split)
    return split


def direct_interpolation(A, S, split):
    Ap, Aj, Ax = _csr64(A); Sp, Sj, Sx = _csr64(S)
    n = A.shape[0]
    Bp = np.empty(n + 1, np.int64)
    lib().oracle_direct_interp_pass1(n, Sp, Sj, split, Bp)
    Bj = np.empty(Bp[n], np.int64); Bx = np.empty(Bp[n])
    nc = lib().oracle_direct_interp_pass2(n, Ap, Aj, Ax, Sp, Sj, Sx, split, Bp, Bj, Bx)
    return sp.csr_matrix((Bx, Bj, Bp), shape=(n, nc))


class RugeStuben:
    """``pyamg.ruge_stuben_solver(A)`` with its 4.0.0 defaults."""

    def __init__(self, A, max_levels=10, max_coarse=10, theta=0.25):
        A = sp.csr_matrix(A).astype(np.float64)
        A.sort_indices()
        self.A, self.P, self.R, self.split = [A], [], [], []
        while len(self.A) < max_levels and self.A[-1].shape[0] > max_coarse:
            A = self.A[-1]
            S = strength_abs(A, theta)
            split = rs_splitting(S)
            P = direct_interpolation(A, S, split)
            R = P.T.tocsr()
            Ac = (R @ A @ P).tocsr()
            Ac.sort_indices()
            self.split.append(split); self.P.append(P); self.R.append(R); self.A.append(Ac)
        self.coarse_pinv = pinv2(self.A[-1].toarray())
        self._csr = [_csr64(a) for a in self.A]

    def _gs_symmetric(self, l, x, b):
        Ap, Aj, Ax = self._csr[l]
        n = len(x)
        lib().oracle_gauss_seidel(Ap, Aj, Ax, x, b, 0, n, 1)
        lib().oracle_gauss_seidel(Ap, Aj, Ax, x, b, n - 1, -1, -1)

    def _cycle(self, l, x, b):
        A = self.A[l]
        self._gs_symmetric(l, x, b)
        r = b - A @ x
        cb = self.R[l] @ r
        cx = np.zeros_like(cb)
        if l == len(self.A) - 2:
            cx[:] = self.coarse_pinv @ cb
        else:
            self._cycle(l + 1, cx, cb)
        x += self.P[l] @ cx
        self._gs_symmetric(l, x, b)

    def precondition(self, b):
        """``ml.aspreconditioner()``: one V(1,1) cycle from x0 = 0."""
        b = np.ascontiguousarray(b, dtype=np.float64)
        if len(self.A) == 1:
            return self.coarse_pinv @ b
        x = np.zeros_like(b)
        self._cycle(0, x, b)
        return x

    def solve(self, b, tol=1e-10, maxiter=10000):
        """``ml.solve(b, tol, maxiter)``: stand-alone V-cycle iteration (ref utils.py:196)."""
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.zeros_like(b)
        A = self.A[0]
        normb = np.linalg.norm(b)
        if normb == 0.0:
            normb = 1.0
        it = 0
        while True:
            if len(self.A) == 1:
                x = self.coarse_pinv @ b
            else:
                self._cycle(0, x, b)
            it += 1
            if np.linalg.norm(b - A @ x) / normb < tol or it == maxiter:
                return x, it

    def operator_complexity(self):
        return sum(a.nnz for a in self.A) / self.A[0].nnz
