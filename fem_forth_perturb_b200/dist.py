"""Row-block partitioned hot path for N GPUs (the reference is a single CPU process).

* partition: recursive coordinate bisection (RCB) of the DOF locations into P parts; inside a part the DOFs
  follow a Hilbert curve.  The solver numbering is (part, curve position), so every rank owns a contiguous
  row block [lo, hi) and the user-visible skfem numbering is only touched at the boundary.
* assembly: every rank assembles the matrix rows it owns (``mwx_assemble_morley_rows``) - no communication.
* PCG: per iteration one halo exchange of p (neighbour point-to-point) and three small all-reduces
  (p.Ap; r.r [+ r.z]); the preconditioner is Jacobi, or block-Jacobi AMG: each rank runs the single-GPU V-cycle
  on its own diagonal block (iteration counts then depend on P - bench.py reports them).

The driver code is written once over a list of ``parts``: with ``torch.distributed`` the list holds this
rank's part (one process per GPU, NCCL); with :class:`InProcessComm` it holds all P parts of one process,
which is how the partitioned algorithm is checked on a single GPU (and, with a CPU ``ops`` backend supplied by
the tests, under gloo on CPU).  The kernels behind ``CudaOps`` are the same ones the single-GPU solver uses.
"""
import ctypes
import math

import numpy as np

from . import _lib as L

SLOT_RHO, SLOT_RHO_PREV, SLOT_PQ, SLOT_RR, SLOT_AUX = 0, 1, 2, 3, 4


# ------------------------------------------------------------------------------------------ partition
def rcb_parts(coords, nparts):
    """RCB part id (int64 tensor, same device) of every point; parts are balanced to +-1 point."""
    import torch
    n = coords.shape[0]
    part = torch.zeros(n, dtype=torch.int64, device=coords.device)

    def split(idx, first, count):
        if count == 1:
            part[idx] = first
            return
        left = count // 2
        xy = coords[idx]
        ext = xy.max(dim=0).values - xy.min(dim=0).values
        axis = 0 if float(ext[0]) >= float(ext[1]) else 1
        # stable order along the axis, ties broken by the other coordinate then by index
        other = torch.argsort(xy[:, 1 - axis], stable=True)
        order = other[torch.argsort(xy[other, axis], stable=True)]
        k = (idx.numel() * left) // count
        split(idx[order[:k]], first, left)
        split(idx[order[k:]], first + left, count - left)

    split(torch.arange(n, device=coords.device), 0, int(nparts))
    return part


class Partition:
    """Solver numbering for P ranks: perm[new] = old (condensed) id, offsets[p] .. offsets[p+1] = rows of rank p."""

    def __init__(self, coords, nparts):
        import torch
        n = coords.shape[0]
        self.n, self.nparts = n, int(nparts)
        xy = coords.contiguous()
        lo, hi = xy.min(dim=0).values, xy.max(dim=0).values
        extent = float((hi - lo).max().item()) or 1.0
        keys = torch.empty(n, dtype=torch.int64, device=xy.device)
        L.check(L.lib().mwx_hilbert_keys(n, L.ptr(xy), float(lo[0].item()), float(lo[1].item()), extent, L.ptr(keys),
                                         L.stream()), 'mwx_hilbert_keys')
        part = rcb_parts(xy, self.nparts) if self.nparts > 1 else torch.zeros(n, dtype=torch.int64, device=xy.device)
        order = torch.sort((part << 48) | keys, stable=True).indices
        self.perm = order.to(torch.int32)
        self.iperm = torch.empty(n, dtype=torch.int32, device=xy.device)
        self.iperm[order] = torch.arange(n, dtype=torch.int32, device=xy.device)
        counts = torch.bincount(part, minlength=self.nparts)
        self.offsets = [0] + [int(v) for v in torch.cumsum(counts, 0).cpu()]


# ------------------------------------------------------------------------------------------ one rank's block
class RankSystem:
    """Local operator of one rank: rows [lo, hi) of the solver-numbered K2c, columns = [owned | ghost]."""

    def __init__(self, basis, cplan, partition, rank):
        torch = L.require_cuda()
        lib, st = L.lib(), L.stream()
        m, pat = basis.mesh, basis.pattern()
        dev = m.device
        self.basis, self.rank, self.partition = basis, rank, partition
        self.lo, self.hi = partition.offsets[rank], partition.offsets[rank + 1]
        self.nloc = self.hi - self.lo
        I_dev = torch.from_numpy(np.ascontiguousarray(cplan.I, dtype=np.int64)).to(dev)
        gid = torch.full((basis.N,), -1, dtype=torch.int32, device=dev)
        gid[I_dev] = partition.iperm                                        # skfem DOF -> solver id
        self.rows = I_dev[partition.perm[self.lo:self.hi].long()].to(torch.int32).contiguous()
        self.raw_indptr = torch.empty(self.nloc + 1, dtype=torch.int64, device=dev)
        self.indptr = torch.empty(self.nloc + 1, dtype=torch.int64, device=dev)
        L.check(lib.mwx_dist_rows_count(self.nloc, L.ptr(self.rows), L.ptr(pat['indptr']), L.ptr(pat['indices']),
                                        L.ptr(gid), L.ptr(self.raw_indptr), L.ptr(self.indptr), st), 'mwx_dist_rows_count')
        self.raw_nnz, self.nnz = int(self.raw_indptr[-1].item()), int(self.indptr[-1].item())
        cols_g = torch.empty(self.nnz, dtype=torch.int32, device=dev)
        self.src = torch.empty(self.nnz, dtype=torch.int64, device=dev)
        L.check(lib.mwx_dist_rows_fill(self.nloc, L.ptr(self.rows), L.ptr(pat['indptr']), L.ptr(pat['indices']),
                                       L.ptr(gid), L.ptr(self.raw_indptr), L.ptr(self.indptr), L.ptr(cols_g),
                                       L.ptr(self.src), st), 'mwx_dist_rows_fill')
        off = (cols_g < self.lo) | (cols_g >= self.hi)
        self.ghosts = torch.unique(cols_g[off]).to(torch.int32).contiguous()         # sorted global ids
        self.nghost = int(self.ghosts.numel())
        self.cols = torch.empty(self.nnz, dtype=torch.int32, device=dev)
        L.check(lib.mwx_dist_localize(self.nnz, L.ptr(cols_g), self.lo, self.hi, L.ptr(self.ghosts), self.nghost,
                                      self.nloc, L.ptr(self.cols), st), 'mwx_dist_localize')
        self.raw_vals = torch.empty(self.raw_nnz, dtype=torch.float64, device=dev)
        self.vals = torch.empty(self.nnz, dtype=torch.float64, device=dev)
        rows_plan = ctypes.c_int32(0)
        L.check(lib.mwx_spmv_plan(self.nloc, L.ptr(self.indptr), ctypes.byref(rows_plan), st), 'mwx_spmv_plan')
        self.spmv_hint = -int(rows_plan.value)
        # ghost ownership: sorted ghosts are grouped by owner because ownership ranges are contiguous
        offs = torch.tensor(partition.offsets[1:], dtype=torch.int32, device=dev)
        owner = torch.searchsorted(offs, self.ghosts, right=True).cpu().numpy()
        self.recv = {}                                                       # peer -> (start, count) in the ghost segment
        for peer in np.unique(owner):
            sel = np.nonzero(owner == peer)[0]
            self.recv[int(peer)] = (int(sel[0]), int(len(sel)))
        self.send = {}                                                       # peer -> local indices (device int32)

    def requests(self):
        """{owner: global ids this rank needs from it}."""
        return {peer: self.ghosts[s:s + c] for peer, (s, c) in self.recv.items()}

    def set_send_lists(self, asked):
        """asked: {peer: global ids (device) that peer needs from this rank}."""
        for peer, ids in asked.items():
            self.send[peer] = (ids - self.lo).to(L.require_cuda().int32).contiguous()

    def assemble(self, form):
        """Hot path A for this rank: its own rows of K2 (closed-form kernel), then the local value layout."""
        from .forms import as_form
        f = as_form(form)
        m, pat, lib, st = self.basis.mesh, self.basis.pattern(), L.lib(), L.stream()
        L.check(lib.mwx_assemble_morley_rows(self.nloc, L.ptr(self.rows), L.ptr(self.raw_indptr), m.nt, L.ptr(m.d_pxy),
                                             L.ptr(m.d_t), L.ptr(pat['r2e_ptr']), L.ptr(pat['r2e_idx']),
                                             L.ptr(pat['slot8']), L.ptr(pat['indptr']), f.terms.get('a_load', 0.0),
                                             f.terms.get('b_load', 0.0), L.ptr(self.raw_vals), st),
                'mwx_assemble_morley_rows')
        L.check(lib.mwx_gather_f64_i64(self.nnz, L.ptr(self.raw_vals), L.ptr(self.src), L.ptr(self.vals), st),
                'mwx_gather_f64_i64')
        return self

    def diagonal_block(self):
        """Owned-columns-only square block as a DeviceCSR (block-Jacobi AMG is built on it)."""
        torch = L.require_cuda()
        from .assembly import DeviceCSR
        mask = self.cols < self.nloc
        csum = torch.zeros(self.nnz + 1, dtype=torch.int64, device=self.cols.device)
        csum[1:] = torch.cumsum(mask.to(torch.int64), 0)
        indptr = csum[self.indptr]
        return DeviceCSR(indptr.contiguous(), self.cols[mask].contiguous(), self.vals[mask].contiguous(),
                         (self.nloc, self.nloc))

    def to_scipy_global(self):
        """Rows of this rank with GLOBAL solver column ids (tests)."""
        import scipy.sparse as sp
        ghosts = self.ghosts.cpu().numpy()
        cols = self.cols.cpu().numpy().astype(np.int64)
        gcols = np.where(cols < self.nloc, cols + self.lo, ghosts[np.maximum(cols - self.nloc, 0)] if len(ghosts) else 0)
        return sp.csr_matrix((self.vals.cpu().numpy(), gcols, self.indptr.cpu().numpy()),
                             shape=(self.nloc, self.partition.n))


# ------------------------------------------------------------------------------------------ communication
class InProcessComm:
    """All P parts live in this process (single GPU): exchanges are direct copies.  Used to validate the
    partitioned algorithm without a second GPU."""

    def __init__(self, nparts):
        self.world = nparts

    def ranks(self):
        return list(range(self.world))

    def exchange_requests(self, parts):
        asked = {p.rank: {} for p in parts}
        for p in parts:
            for owner, ids in p.requests().items():
                asked[owner][p.rank] = ids
        for p in parts:
            p.set_send_lists(asked[p.rank])

    def halo(self, parts, vecs):
        """vecs[i] = [owned | ghost] vector of parts[i]; fills the ghost segments from the owners."""
        by_rank = {p.rank: (p, v) for p, v in zip(parts, vecs)}
        for p, v in zip(parts, vecs):
            for owner, (s, c) in p.recv.items():
                po, vo = by_rank[owner]
                v[p.nloc + s:p.nloc + s + c] = vo[:po.nloc][po.send[p.rank].long()]

    def allreduce_sum(self, tensors):
        total = tensors[0].clone()
        for t in tensors[1:]:
            total += t.to(total.device)
        for t in tensors:
            t.copy_(total)


class TorchDistComm:
    """One part per process; NCCL (or gloo) point-to-point halo + all-reduce."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)

    def ranks(self):
        return [self.rank]

    def exchange_requests(self, parts):
        import torch
        (p,) = parts
        dev = p.ghosts.device
        counts = torch.zeros(self.world, dtype=torch.int64, device=dev)
        for owner, (s, c) in p.recv.items():
            counts[owner] = c
        all_counts = [torch.zeros_like(counts) for _ in range(self.world)]
        self.dist.all_gather(all_counts, counts, group=self.group)
        ops, bufs = [], {}
        for peer in range(self.world):
            need = int(all_counts[peer][self.rank].item())            # what peer asks of me
            if need and peer != self.rank:
                bufs[peer] = torch.empty(need, dtype=torch.int32, device=dev)
                ops.append(self.dist.P2POp(self.dist.irecv, bufs[peer], peer, group=self.group))
        for owner, ids in p.requests().items():
            if owner != self.rank:
                ops.append(self.dist.P2POp(self.dist.isend, ids.contiguous(), owner, group=self.group))
        if ops:
            for w in self.dist.batch_isend_irecv(ops):
                w.wait()
        p.set_send_lists(bufs)
        p._sendbuf = {peer: torch.empty(idx.numel(), dtype=torch.float64, device=dev) for peer, idx in p.send.items()}

    def halo(self, parts, vecs):
        (p,), (v,) = parts, vecs
        ops = []
        for owner, (s, c) in p.recv.items():
            ops.append(self.dist.P2POp(self.dist.irecv, v[p.nloc + s:p.nloc + s + c], owner, group=self.group))
        for peer, idx in p.send.items():
            buf = p._sendbuf[peer]
            p.ops.gather(v, idx, buf)
            ops.append(self.dist.P2POp(self.dist.isend, buf, peer, group=self.group))
        if ops:
            for w in self.dist.batch_isend_irecv(ops):
                w.wait()

    def allreduce_sum(self, tensors):
        (t,) = tensors
        self.dist.all_reduce(t, group=self.group)


# ------------------------------------------------------------------------------------------ kernels behind the driver
class CudaOps:
    """The device kernels (C ABI) the partitioned CG chains; tests substitute a CPU stand-in under gloo."""

    def workspace(self, device):
        torch = L.require_cuda()
        return torch.zeros(L.lib().mwx_dcg_workspace_doubles(), dtype=torch.float64, device=device)

    def zeros(self, n, device):
        return L.require_cuda().zeros(n, dtype=L.require_cuda().float64, device=device)

    def gather(self, v, idx, out):
        L.check(L.lib().mwx_gather_f64(idx.numel(), L.ptr(v), L.ptr(idx), L.ptr(out), L.stream()), 'mwx_gather_f64')

    def diag_inv(self, p):
        torch = L.require_cuda()
        d = torch.empty(p.nloc, dtype=torch.float64, device=p.vals.device)
        L.check(L.lib().mwx_csr_diagonal(p.nloc, L.ptr(p.indptr), L.ptr(p.cols), L.ptr(p.vals), L.ptr(d), L.stream()),
                'mwx_csr_diagonal')
        return 1.0 / d

    def dot(self, n, a, b, ws, slot):
        L.check(L.lib().mwx_dcg_dot(n, L.ptr(a), L.ptr(b), L.ptr(ws), slot, L.stream()), 'mwx_dcg_dot')

    def p_update(self, n, r, dinv, z, p_ext, ws, first):
        L.check(L.lib().mwx_dcg_p_update(n, L.ptr(r), L.ptr(dinv), L.ptr(z), L.ptr(p_ext), L.ptr(ws), int(first),
                                         L.stream()), 'mwx_dcg_p_update')

    def spmv_dot(self, p, x_ext, q, ws):
        L.check(L.lib().mwx_dcg_spmv_dot(p.nloc, p.spmv_hint, L.ptr(p.indptr), L.ptr(p.cols), L.ptr(p.vals),
                                         L.ptr(x_ext), L.ptr(q), L.ptr(ws), L.stream()), 'mwx_dcg_spmv_dot')

    def xr_update(self, n, p_ext, q, x_ext, r, dinv, ws, jacobi):
        L.check(L.lib().mwx_dcg_xr_update(n, L.ptr(p_ext), L.ptr(q), L.ptr(x_ext), L.ptr(r), L.ptr(dinv), L.ptr(ws),
                                          int(jacobi), L.stream()), 'mwx_dcg_xr_update')

    def residual(self, p, x_ext, b, r, dinv, ws):
        L.check(L.lib().mwx_dcg_residual(p.nloc, L.ptr(p.indptr), L.ptr(p.cols), L.ptr(p.vals), L.ptr(x_ext), L.ptr(b),
                                         L.ptr(r), L.ptr(dinv), L.ptr(ws), L.stream()), 'mwx_dcg_residual')

    def spmv(self, p, x_ext, y):
        L.check(L.lib().mwx_spmv(p.nloc, p.spmv_hint, L.ptr(p.indptr), L.ptr(p.cols), L.ptr(p.vals), L.ptr(x_ext),
                                 L.ptr(y), L.stream()), 'mwx_spmv')


# ------------------------------------------------------------------------------------------ the partitioned PCG
def dist_pcg(parts, comm, b_list, tol=1e-8, maxiter=None, precond='jacobi', amg=None, ops=None):
    """scipy-1.4.1-semantics PCG on the row-block partitioned system.

    parts: this process's RankSystem-like blocks; b_list: their right-hand sides (owned entries, solver order);
    precond: 'jacobi' | 'none' | 'amg' (block-Jacobi: amg[i].precondition applied to parts[i]'s residual).
    Returns (x_list, iters, info, resid)."""
    ops = ops or CudaOps()
    n_global = parts[0].partition.n
    if maxiter is None or maxiter <= 0:
        maxiter = 10 * n_global
    S = []
    for p, b in zip(parts, b_list):
        dev = b.device
        p.ops = ops
        st = dict(b=b, ws=ops.workspace(dev), x=ops.zeros(p.nloc + p.nghost, dev), p=ops.zeros(p.nloc + p.nghost, dev),
                  r=b.clone(), q=ops.zeros(p.nloc, dev), z=None,
                  dinv=ops.diag_inv(p) if precond == 'jacobi' else None)
        S.append(st)
    jacobi = precond in ('jacobi', 'none')

    def reduce_slots(slots):
        packs = [st['ws'][slots].clone() for st in S]
        comm.allreduce_sum(packs)
        for st, pk in zip(S, packs):
            st['ws'][slots] = pk
        return packs[0]

    for p, st in zip(parts, S):
        ops.dot(p.nloc, st['b'], st['b'], st['ws'], SLOT_AUX)
    bnrm = math.sqrt(float(reduce_slots([SLOT_AUX])[0]))
    xs = lambda: [st['x'][:p.nloc] for p, st in zip(parts, S)]
    if bnrm <= tol:
        return xs(), 0, 0, bnrm
    atol = tol * bnrm if bnrm != 0.0 else tol
    if jacobi:
        for p, st in zip(parts, S):                              # x = 0: r = b, rr, rho = sum dinv r^2
            ops.residual(p, st['x'], st['b'], st['r'], st['dinv'], st['ws'])
        reduce_slots([SLOT_RHO, SLOT_RR])
    resid = bnrm
    for it in range(1, maxiter + 1):
        if not jacobi:
            for i, (p, st) in enumerate(zip(parts, S)):
                st['z'] = amg[i].precondition(st['r'])
                ops.dot(p.nloc, st['r'], st['z'], st['ws'], SLOT_RHO)
            reduce_slots([SLOT_RHO])
        for p, st in zip(parts, S):
            ops.p_update(p.nloc, st['r'], st['dinv'], st['z'] if not jacobi else None, st['p'], st['ws'], it == 1)
        comm.halo(parts, [st['p'] for st in S])
        for p, st in zip(parts, S):
            ops.spmv_dot(p, st['p'], st['q'], st['ws'])
        reduce_slots([SLOT_PQ])
        for p, st in zip(parts, S):
            ops.xr_update(p.nloc, st['p'], st['q'], st['x'], st['r'], st['dinv'], st['ws'], jacobi)
        red = reduce_slots([SLOT_RHO, SLOT_RR] if jacobi else [SLOT_RR])
        resid = math.sqrt(float(red[-1]))
        if resid != resid:
            return xs(), it, it, resid
        if resid <= atol:
            if it == 1:
                return xs(), it, 0, resid
            comm.halo(parts, [st['x'] for st in S])              # true-residual re-check (scipy 1.4.1)
            for p, st in zip(parts, S):
                ops.residual(p, st['x'], st['b'], st['r'], st['dinv'], st['ws'])
            red = reduce_slots([SLOT_RHO, SLOT_RR] if jacobi else [SLOT_RR])
            resid = math.sqrt(float(red[-1]))
            if resid <= atol:
                return xs(), it, 0, resid
    return xs(), maxiter, maxiter, resid


class PartitionedProblem:
    """Convenience wrapper: mesh/basis/Dirichlet set -> partition -> this process's blocks."""

    def __init__(self, basis, D, comm):
        from .solvers import CondensePlan
        from .assembly import DeviceCSR
        torch = L.require_cuda()
        pat = basis.pattern()
        proto = DeviceCSR(pat['indptr'], pat['indices'], torch.empty(0, dtype=torch.float64, device=basis.mesh.device),
                          (basis.N, basis.N), coords=basis.dof_coords)
        self.cplan = CondensePlan(proto, D=D)
        I_dev = torch.from_numpy(np.ascontiguousarray(self.cplan.I, dtype=np.int64)).to(basis.mesh.device)
        self.partition = Partition(basis.dof_coords()[I_dev], comm.world)
        self.comm, self.basis = comm, basis
        self.parts = [RankSystem(basis, self.cplan, self.partition, r) for r in comm.ranks()]
        comm.exchange_requests(self.parts)

    def assemble(self, form):
        for p in self.parts:
            p.assemble(form)
        return self
