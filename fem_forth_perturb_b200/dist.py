"""Row-block partitioned hot path for N GPUs (the reference is a single CPU process).

* partition: recursive coordinate bisection (RCB) of the DOF locations into P parts; inside a part the DOFs
  follow a Hilbert curve.  The solver numbering is (part, curve position), so every rank owns a contiguous
  row block [lo, hi) and the user-visible skfem numbering is only touched at the boundary.
* assembly: every rank assembles the matrix rows it owns (``mwx_assemble_morley_rows``) - no communication.
* PCG: per iteration one halo exchange of p and three small all-reduces (p.Ap; r.r [+ r.z]); the preconditioner is
  Jacobi, the GLOBAL AMG hierarchy applied as a collective V-cycle (:class:`DistAMG`: row-blocked upper levels,
  replicated tail), or block-Jacobi AMG (each rank's own diagonal block; iteration counts then depend on P).
* communication: :class:`PeerComm` (default in bench.py) makes every exchange of an iteration - halo push fused with
  an epoch-flag barrier, all-reduce, all-gather - a kernel storing into peer memory over NVLink, so iterations 2.. are
  ONE CUDA-graph launch each; :class:`SymmMemComm` (push kernel + symmetric-memory barrier + NCCL all-reduce) and
  :class:`TorchDistComm` (NCCL / gloo send-recv) are the fallbacks.

The driver code is written once over a list of ``parts``: with ``torch.distributed`` the list holds this
rank's part (one process per GPU, NCCL); with :class:`InProcessComm` it holds all P parts of one process,
which is how the partitioned algorithm is checked on a single GPU (and, with a CPU ``ops`` backend supplied by
the tests, under gloo on CPU).  The kernels behind ``CudaOps`` are the same ones the single-GPU solver uses.
"""
import ctypes
import os
import math
import time

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
class HaloPlan:
    """Column side of a row-block operator: owned columns [clo, clo + ncol) then `ghosts` (sorted global ids)."""

    def _plan_ghosts(self, col_offsets, device):
        torch = L.require_cuda()
        offs = torch.tensor(col_offsets[1:], dtype=torch.int32, device=device)
        owner = torch.searchsorted(offs, self.ghosts, right=True).cpu().numpy()
        self.recv = {}                                   # owner -> (start, count) inside the ghost segment
        for peer in np.unique(owner):
            sel = np.nonzero(owner == peer)[0]
            self.recv[int(peer)] = (int(sel[0]), int(len(sel)))
        self.send = {}                                   # peer -> local owned indices it needs (device int32)

    def requests(self):
        """{owner: global ids this rank needs from it}."""
        return {peer: self.ghosts[s:s + c] for peer, (s, c) in self.recv.items()}

    def set_send_lists(self, asked):
        """asked: {peer: global ids (device) that peer needs from this rank}."""
        for peer, ids in asked.items():
            self.send[peer] = (ids - self.clo).to(L.require_cuda().int32).contiguous()


def pack_fp32(cols, vals):
    """fp32-stored twin of a CSR operator: one 8-byte (fp32 value, int32 column) record per nonzero (krylov.cu)."""
    torch = L.require_cuda()
    out = torch.empty(max(int(vals.numel()), 2), dtype=torch.int64, device=vals.device)        # 8 B per record, 16-B aligned
    L.check(L.lib().mwx_csr_pack(vals.numel(), L.ptr(cols), L.ptr(vals), L.ptr(out), L.stream()), 'mwx_csr_pack')
    return out


def make_lx(op):
    """Tile-local column plan (csrc/spmv_lx.cu) of a local operator, or None (pattern does not localise / not asked for)."""
    if os.environ.get('MWX_SPMV_LX', '0') == '0' or op.nloc <= 0 or op.spmv_hint > -8:      # off by default: measured slower (amg.cu)
        return None
    h = ctypes.c_void_p()
    L.check(L.lib().mwx_spmv_lx_create(op.nloc, L.ptr(op.indptr), L.ptr(op.cols), int(op.spmv_hint), ctypes.byref(h),
                                       L.stream()), 'mwx_spmv_lx_create')
    return h if h.value else None


class LocalOp(HaloPlan):
    """Rows [rlo, rhi) of a global CSR operator (device tensors), columns renumbered to [owned | ghost]."""

    def __init__(self, indptr, indices, values, rlo, rhi, col_offsets, rank, block=None, fp32=False):
        """block = (br, bc): also build the block-CSR twin (csrc/bsr.cu) when the local operator has that structure -
        owned columns start at a block boundary and ghosts come in whole blocks for the operators of a
        smoothed-aggregation hierarchy; the kernels behind CudaOps then use it (``self.bsr``)."""
        torch = L.require_cuda()
        dev = values.device
        self.rank, self.nloc = rank, rhi - rlo
        self.bsr, self.pk, self.lx = None, None, None
        self.clo, self.ncol = col_offsets[rank], col_offsets[rank + 1] - col_offsets[rank]
        s0, s1 = int(indptr[rlo].item()), int(indptr[rhi].item())
        self.indptr = (indptr[rlo:rhi + 1] - s0).contiguous()
        cols_g = indices[s0:s1].contiguous()
        self.vals = values[s0:s1].clone()
        self.nnz = s1 - s0
        off = (cols_g < self.clo) | (cols_g >= self.clo + self.ncol)
        self.ghosts = torch.unique(cols_g[off]).to(torch.int32).contiguous()
        self.nghost = int(self.ghosts.numel())
        self.cols = torch.empty(self.nnz, dtype=torch.int32, device=dev)
        L.check(L.lib().mwx_dist_localize(self.nnz, L.ptr(cols_g), self.clo, self.clo + self.ncol, L.ptr(self.ghosts),
                                          self.nghost, self.ncol, L.ptr(self.cols), L.stream()), 'mwx_dist_localize')
        rows_plan = ctypes.c_int32(0)
        if self.nloc > 0:
            L.check(L.lib().mwx_spmv_plan(self.nloc, L.ptr(self.indptr), ctypes.byref(rows_plan), L.stream()), 'mwx_spmv_plan')
        self.spmv_hint = -int(rows_plan.value) if rows_plan.value else -1
        self._plan_ghosts(col_offsets, dev)
        if block is not None and self.nnz > 0 and os.environ.get('MWX_BSR', '1') != '0':
            h = ctypes.c_void_p()
            L.check(L.lib().mwx_bsr_create(self.nloc, self.ncol + self.nghost, self.nnz, L.ptr(self.indptr), L.ptr(self.cols),
                                           L.ptr(self.vals), int(block[0]), int(block[1]), 32 if fp32 else 64, ctypes.byref(h),
                                           L.stream()), 'mwx_bsr_create')
            self.bsr = h if h.value else None
        if self.bsr is None and self.nnz > 0:
            self.lx = make_lx(self)
        if fp32 and self.bsr is None and self.nnz > 0 and os.environ.get('MWX_PACKED_CSR', '0') != '0':
            self.pk = pack_fp32(self.cols, self.vals)      # scalar operators: measured slower than fp64 CSR (amg.cu), A/B only

    def __del__(self):
        try:
            if getattr(self, 'bsr', None):
                L.lib().mwx_bsr_destroy(self.bsr)
            if getattr(self, 'lx', None):
                L.lib().mwx_spmv_lx_destroy(self.lx)
        except Exception:
            pass


class RankSystem(HaloPlan):
    """Local operator of one rank: rows [lo, hi) of the solver-numbered K2c, columns = [owned | ghost]."""

    def __init__(self, basis, cplan, partition, rank):
        torch = L.require_cuda()
        lib, st = L.lib(), L.stream()
        m, pat = basis.mesh, basis.pattern()
        dev = m.device
        self.basis, self.rank, self.partition = basis, rank, partition
        self.lo, self.hi = partition.offsets[rank], partition.offsets[rank + 1]
        self.nloc = self.hi - self.lo
        self.clo, self.ncol = self.lo, self.nloc
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
        self.vals = torch.empty(self.nnz, dtype=torch.float64, device=dev)
        # Hot path A on this rank's rows: the tiled element-once kernel writing STRAIGHT into the local layout (the
        # plan matches an element's columns against the skfem DOF of every kept entry; Dirichlet columns have no
        # destination and are skipped).  The row-gather kernel + value gather stays as the fallback for meshes the
        # plan refuses (MWX_E_CAPACITY) and under MWX_ASM_KERNEL=rows.
        self.tile_plan, self.raw_vals = None, None
        if os.environ.get('MWX_ASM_KERNEL', 'tiles') == 'tiles' and self.nloc > 0:
            from .assembly import TilePlan
            dst_col = I_dev[partition.perm[cols_g.long()].long()].to(torch.int32).contiguous()
            dst_len = (self.indptr[1:] - self.indptr[:-1]).to(torch.int32).contiguous()
            plan = TilePlan(basis, self.rows, self.indptr[:-1].contiguous(), dst_len, dst_col)
            self.tile_plan = plan if plan.ok else None
        if self.tile_plan is None:
            self.raw_vals = torch.empty(self.raw_nnz, dtype=torch.float64, device=dev)
        rows_plan = ctypes.c_int32(0)
        L.check(lib.mwx_spmv_plan(self.nloc, L.ptr(self.indptr), ctypes.byref(rows_plan), st), 'mwx_spmv_plan')
        self.spmv_hint = -int(rows_plan.value)
        self.lx = make_lx(self)           # never freed explicitly: lives as long as the process's partitioned problem
        # ghost ownership: sorted ghosts are grouped by owner because ownership ranges are contiguous
        self._plan_ghosts(partition.offsets, dev)

    def assemble(self, form):
        """Hot path A for this rank: its own rows of K2 (closed-form kernel), then the local value layout."""
        from .forms import as_form
        f = as_form(form)
        m, pat, lib, st = self.basis.mesh, self.basis.pattern(), L.lib(), L.stream()
        if self.tile_plan is not None:
            self.tile_plan.assemble(m, f.terms.get('a_load', 0.0), f.terms.get('b_load', 0.0), self.vals)
            return self
        L.check(lib.mwx_assemble_morley_rows(self.nloc, L.ptr(self.rows), L.ptr(self.raw_indptr), m.nt, L.ptr(m.d_pxy),
                                             L.ptr(m.d_t), L.ptr(pat['exy']), 1, L.ptr(pat['r2e_ptr']), L.ptr(pat['r2e_idx']),
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
                v[p.ncol + s:p.ncol + s + c] = vo[:po.ncol][po.send[p.rank].long()]

    def allreduce_sum(self, tensors):
        total = tensors[0].clone()
        for t in tensors[1:]:
            total += t.to(total.device)
        for t in tensors:
            t.copy_(total)

    def ext_zeros(self, op, device):
        """[owned | ghost] work vector for operator `op`."""
        import torch
        return torch.zeros(op.ncol + op.nghost, dtype=torch.float64, device=device)

    def allgather(self, vecs, offsets):
        """Every part gets the concatenation of all owned segments."""
        import torch
        full = torch.cat(vecs)
        return [full for _ in vecs]

    def is_root(self):
        return True

    def broadcast_tensors(self, tensors):
        """Root's list of device tensors (None elsewhere) -> the same list on every rank."""
        return tensors


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
            ops.append(self.dist.P2POp(self.dist.irecv, v[p.ncol + s:p.ncol + s + c], owner, group=self.group))
        for peer, idx in p.send.items():
            buf = p._sendbuf[peer]
            (getattr(p, 'ops', None) or _default_ops()).gather(v, idx, buf)
            ops.append(self.dist.P2POp(self.dist.isend, buf, peer, group=self.group))
        if ops:
            for w in self.dist.batch_isend_irecv(ops):
                w.wait()

    def allreduce_sum(self, tensors):
        (t,) = tensors
        self.dist.all_reduce(t, group=self.group)

    def ext_zeros(self, op, device):
        import torch
        return torch.zeros(op.ncol + op.nghost, dtype=torch.float64, device=device)

    def allgather(self, vecs, offsets):
        import torch
        (v,) = vecs
        sizes = [offsets[r + 1] - offsets[r] for r in range(self.world)]
        m = max(sizes)
        pad = torch.zeros(m, dtype=v.dtype, device=v.device)
        pad[:v.numel()] = v
        out = torch.empty(self.world * m, dtype=v.dtype, device=v.device)
        self.dist.all_gather_into_tensor(out, pad, group=self.group)
        return [torch.cat([out[r * m:r * m + sizes[r]] for r in range(self.world)])]

    def is_root(self):
        return self.rank == 0

    def broadcast_tensors(self, tensors):
        import torch
        dev = torch.device('cuda', torch.cuda.current_device()) if torch.cuda.is_available() else torch.device('cpu')
        meta = [[(tuple(t.shape), str(t.dtype)) for t in tensors]] if self.is_root() else [None]
        self.dist.broadcast_object_list(meta, src=0, group=self.group)
        out = []
        for i, (shape, dt) in enumerate(meta[0]):
            t = tensors[i].contiguous() if self.is_root() else torch.empty(shape, dtype=getattr(torch, dt.split('.')[-1]), device=dev)
            self.dist.broadcast(t, src=0, group=self.group)
            out.append(t)
        return out


class SymmMemComm(TorchDistComm):
    """Halo exchange over NVLink peer memory instead of NCCL send/recv.

    Every [owned | ghost] work vector lives in symmetric memory (``torch.distributed._symmetric_memory``:
    same allocation on every rank, peer pointers mapped into this process).  A halo exchange is then one
    pack-and-push kernel per neighbour - ``mwx_gather_f64`` reading this rank's boundary values and storing
    them STRAIGHT INTO the neighbour's ghost segment through its peer pointer - followed by one device-side
    barrier on the allocation's signal pads (~12 us per exchange on B200/NVSwitch, measured, against ~100 us
    for a batched NCCL send/recv plus its host bookkeeping).  All-reduces / all-gathers stay on NCCL.

    Write-after-read safety: a ghost segment may only be overwritten after its consumers finished reading the
    previous halo.  Every barrier and every NCCL collective is stream-ordered after those consumers on each
    rank, so one synchronisation event between two halos of the same vector is enough; the comm counts such
    events and inserts an extra barrier when a vector is exchanged twice with none in between."""

    def __init__(self, group=None):
        super().__init__(group)
        import torch.distributed._symmetric_memory as symm_mem
        self.symm_mem = symm_mem
        self.group_name = (group or self.dist.group.WORLD).group_name
        self.handles = {}                 # data_ptr of the local tensor -> (handle, tensor)
        self.epoch = 0
        self.last_halo = {}
        self._plans = {}
        self._push = L.lib().mwx_halo_push

    def ext_zeros(self, op, device):
        import torch
        n = torch.tensor([op.ncol + op.nghost], dtype=torch.int64, device=device)
        self.dist.all_reduce(n, op=self.dist.ReduceOp.MAX, group=self.group)
        t = self.symm_mem.empty(max(int(n.item()), 1), dtype=torch.float64, device=device)
        t.zero_()
        hdl = self.symm_mem.rendezvous(t, self.group_name)
        self.handles[t.data_ptr()] = (hdl, t)
        return t[:op.ncol + op.nghost]

    def release(self, vecs):
        """Forget symmetric-memory vectors handed out by ext_zeros (their handles and cached push plans)."""
        for v in vecs:
            ptr_ = v.data_ptr()
            self.handles.pop(ptr_, None)
            self.last_halo.pop(ptr_, None)
            for k in [k for k in self._plans if k[1] == ptr_]:
                del self._plans[k]

    def exchange_requests(self, parts):
        import torch
        super().exchange_requests(parts)
        (p,) = parts
        dev = p.ghosts.device
        # tell every owner where its values land in MY ext vectors: element offset ncol + start
        mine = torch.full((self.world,), -1, dtype=torch.int64, device=dev)
        for owner, (s, c) in p.recv.items():
            mine[owner] = p.ncol + s
        table = [torch.empty_like(mine) for _ in range(self.world)]
        self.dist.all_gather(table, mine, group=self.group)
        p.push = {peer: int(table[peer][self.rank].item()) for peer in p.send}     # peer -> dst element offset

    def allreduce_sum(self, tensors):
        super().allreduce_sum(tensors)
        self.epoch += 1

    def allgather(self, vecs, offsets):
        out = super().allgather(vecs, offsets)
        self.epoch += 1
        return out

    def _push_plan(self, p, v, hdl):
        """Pre-marshalled arguments of the one-launch halo push of vector v through operator p (cached)."""
        key = (id(p), v.data_ptr())
        plan = self._plans.get(key)
        if plan is None:
            import torch
            dev = v.device
            peers = [peer for peer in p.send if p.send[peer].numel()]
            counts = [p.send[peer].numel() for peer in peers]
            seg = torch.tensor(np.concatenate([[0], np.cumsum(counts)]).astype(np.int64), device=dev)
            idx = torch.cat([p.send[peer] for peer in peers]) if peers else torch.empty(0, dtype=torch.int32, device=dev)
            dst = torch.tensor([hdl.buffer_ptrs[peer] + 8 * p.push[peer] for peer in peers] or [0], dtype=torch.int64,
                               device=dev)
            args = (int(sum(counts)), len(peers), L.ptr(seg), L.ptr(idx), L.ptr(v), L.ptr(dst), L.stream())
            plan = self._plans[key] = (args, (seg, idx, dst, p, v))      # keep the device arrays (and the key's objects) alive
        return plan[0]

    def halo(self, parts, vecs):
        (p,), (v,) = parts, vecs
        vp_ = v.data_ptr()
        hdl, base = self.handles[vp_]
        if self.last_halo.get(vp_) == self.epoch:                 # no sync event since this vector's last halo
            hdl.barrier(channel=0)
            self.epoch += 1
        args = self._push_plan(p, v, hdl)
        if args[0]:
            rc = self._push(*args)
            if rc:
                L.check(rc, 'mwx_halo_push')
        hdl.barrier(channel=0)
        self.epoch += 1
        self.last_halo[vp_] = self.epoch


class _PeerKernels:
    """Launchers of the peer-memory collectives (csrc/dist.cu: mwx_peer_push / mwx_peer_allreduce / mwx_peer_barrier).
    They are plain kernel launches on the current stream with a device-resident barrier epoch, so a whole PCG
    iteration built from them can be captured in a CUDA graph and replayed (``graphable``)."""
    graphable = True

    def _ctl_alloc(self, make):
        """make(nbytes) -> zeroed device tensor (float64) of the control block."""
        nbytes = L.lib().mwx_peer_ctl_bytes()
        return make((nbytes + 7) // 8)

    def _launch_push(self, rank, plan, v, wait):
        total, nseg, seg, idx, dst = plan
        rc = L.lib().mwx_peer_push(total, nseg, L.ptr(seg), L.ptr(idx), L.ptr(v), L.ptr(dst), L.ptr(self._ctl[rank]),
                                   L.ptr(self._peer_table), rank, self.world, self._wait, L.stream())
        if rc:
            L.check(rc, 'mwx_peer_push')

    def _launch_allreduce(self, rank, t, phase):
        rc = L.lib().mwx_peer_allreduce(L.ptr(t), t.numel(), L.ptr(self._ctl[rank]), L.ptr(self._peer_table), rank,
                                        self.world, phase, L.stream())
        if rc:
            L.check(rc, 'mwx_peer_allreduce')

    def _launch_barrier(self, rank):
        rc = L.lib().mwx_peer_barrier(L.ptr(self._ctl[rank]), L.ptr(self._peer_table), rank, self.world, self._wait,
                                      L.stream())
        if rc:
            L.check(rc, 'mwx_peer_barrier')

    @staticmethod
    def _push_arrays(counts, idx_list, dst_ptrs, dev):
        """(total, nseg, seg, idx, dst) device arrays of one push: segment s has counts[s] values for dst_ptrs[s]."""
        import torch
        seg = torch.tensor(np.concatenate([[0], np.cumsum(counts)]).astype(np.int64), device=dev)
        idx = None
        if idx_list is not None:
            idx = torch.cat(idx_list) if idx_list else torch.empty(0, dtype=torch.int32, device=dev)
        dst = torch.tensor(list(dst_ptrs) or [0], dtype=torch.int64, device=dev)
        return (int(sum(counts)), len(counts), seg, idx, dst)


class PeerComm(_PeerKernels, SymmMemComm):
    """One process per GPU; EVERY exchange of the iteration is a kernel storing into the peers' memory over NVLink:

    * halo = pack-and-push fused with the barrier (``mwx_peer_push``: the last block to finish signals all peers'
      epoch flags and waits for theirs) - one launch per exchange instead of push + symmetric-memory barrier;
    * all-reduce of the CG scalars = ``mwx_peer_allreduce`` (every rank stores its partials into all peers' slots,
      barrier, sum in rank order: deterministic, ~one NVLink round trip instead of an NCCL launch);
    * all-gather of the coarse residual for the replicated tail = the same push kernel with one segment per rank.

    No NCCL call and no host-side state is left inside an iteration, which is what lets ``dist_pcg`` capture it in
    a CUDA graph.  NCCL (the parent classes) still does the setup-time exchanges."""

    def __init__(self, group=None):
        super().__init__(group)
        import torch
        dev = torch.device('cuda', torch.cuda.current_device())
        self._wait = 1
        ctl = self._ctl_alloc(lambda n: self.symm_mem.empty(n, dtype=torch.float64, device=dev))
        ctl.zero_()
        hdl = self.symm_mem.rendezvous(ctl, self.group_name)
        self._ctl = {self.rank: ctl}
        self._ctl_hdl = hdl
        self._peer_table = torch.tensor([int(hdl.buffer_ptrs[r]) for r in range(self.world)], dtype=torch.int64, device=dev)
        self._gather_bufs = {}
        torch.cuda.synchronize()
        self.dist.barrier(group=self.group)           # every control block is zeroed before anyone signals

    def _push_plan(self, p, v, hdl):
        key = (id(p), v.data_ptr())
        plan = self._plans.get(key)
        if plan is None:
            peers = [peer for peer in p.send if p.send[peer].numel()]
            arrays = self._push_arrays([p.send[peer].numel() for peer in peers], [p.send[peer] for peer in peers],
                                       [hdl.buffer_ptrs[peer] + 8 * p.push[peer] for peer in peers], v.device)
            plan = self._plans[key] = (arrays, (p, v))
        return plan[0]

    def _guard(self, ptr_):
        """A buffer the peers write into may only be overwritten after a synchronisation that follows its readers."""
        if self.last_halo.get(ptr_) == self.epoch:
            self._launch_barrier(self.rank)
            self.epoch += 1

    def halo(self, parts, vecs):
        (p,), (v,) = parts, vecs
        vp_ = v.data_ptr()
        hdl, _base = self.handles[vp_]
        self._guard(vp_)
        self._launch_push(self.rank, self._push_plan(p, v, hdl), v, self._wait)
        self.epoch += 1
        self.last_halo[vp_] = self.epoch

    def allreduce_sum(self, tensors):
        (t,) = tensors
        import torch
        if t.is_cuda and t.dtype == torch.float64 and t.is_contiguous() and 1 <= t.numel() <= 8:
            self._launch_allreduce(self.rank, t, 0)
            self.epoch += 1
        else:
            super().allreduce_sum(tensors)

    def allgather(self, vecs, offsets):
        import torch
        (v,) = vecs
        key = tuple(offsets)
        ent = self._gather_bufs.get(key)
        if ent is None:
            full = self.symm_mem.empty(max(int(offsets[-1]), 1), dtype=torch.float64, device=v.device)
            full.zero_()
            hdl = self.symm_mem.rendezvous(full, self.group_name)
            n = offsets[self.rank + 1] - offsets[self.rank]
            arrays = self._push_arrays([n] * self.world, None,
                                       [hdl.buffer_ptrs[r] + 8 * offsets[self.rank] for r in range(self.world)], v.device)
            torch.cuda.synchronize()
            self.dist.barrier(group=self.group)
            ent = self._gather_bufs[key] = (full, hdl, arrays)
        full, _hdl, arrays = ent
        self._guard(full.data_ptr())
        self._launch_push(self.rank, arrays, v, self._wait)
        self.epoch += 1
        self.last_halo[full.data_ptr()] = self.epoch
        return [full[:offsets[-1]]]


class InProcessPeerComm(_PeerKernels, InProcessComm):
    """All P parts in this process on one GPU, exchanged with the SAME peer kernels as :class:`PeerComm` (signal-only
    barriers: one stream already orders the parts).  Checks the push / all-reduce / all-gather kernels and the
    graph-captured iteration without a second GPU."""

    def __init__(self, nparts):
        super().__init__(nparts)
        torch = L.require_cuda()
        dev = torch.device('cuda', torch.cuda.current_device())
        self._wait = 0
        self._ctl = {r: self._ctl_alloc(lambda n: torch.zeros(n, dtype=torch.float64, device=dev)) for r in range(nparts)}
        self._peer_table = torch.tensor([self._ctl[r].data_ptr() for r in range(nparts)], dtype=torch.int64, device=dev)
        self._plans, self._gather_bufs = {}, {}

    def halo(self, parts, vecs):
        by_rank = {p.rank: (p, v) for p, v in zip(parts, vecs)}
        for p, v in zip(parts, vecs):
            key = (id(p), v.data_ptr()) + tuple(by_rank[q][1].data_ptr() for q in sorted(p.send))
            plan = self._plans.get(key)
            if plan is None:
                peers = [q for q in sorted(p.send) if p.send[q].numel()]
                dst = [by_rank[q][1].data_ptr() + 8 * (by_rank[q][0].ncol + by_rank[q][0].recv[p.rank][0]) for q in peers]
                plan = self._plans[key] = (self._push_arrays([p.send[q].numel() for q in peers],
                                                             [p.send[q] for q in peers], dst, v.device), (p, v))
            self._launch_push(p.rank, plan[0], v, 0)

    def allreduce_sum(self, tensors):
        import torch
        ok = len(tensors) == self.world and all(t.is_cuda and t.dtype == torch.float64 and t.is_contiguous() and
                                                1 <= t.numel() <= 8 for t in tensors)
        if not ok:
            return super().allreduce_sum(tensors)
        for r, t in enumerate(tensors):
            self._launch_allreduce(r, t, 1)
        for r, t in enumerate(tensors):
            self._launch_allreduce(r, t, 2)

    def allgather(self, vecs, offsets):
        import torch
        key = tuple(offsets)
        ent = self._gather_bufs.get(key)
        if ent is None:
            fulls = [torch.zeros(max(int(offsets[-1]), 1), dtype=torch.float64, device=v.device) for v in vecs]
            plans = [self._push_arrays([offsets[r + 1] - offsets[r]] * self.world, None,
                                       [f.data_ptr() + 8 * offsets[r] for f in fulls], vecs[0].device)
                     for r in range(self.world)]
            ent = self._gather_bufs[key] = (fulls, plans)
        fulls, plans = ent
        for r, v in enumerate(vecs):
            self._launch_push(r, plans[r], v, 0)
        return [f[:offsets[-1]] for f in fulls]


# ------------------------------------------------------------------------------------------ kernels behind the driver
class CudaOps:
    """The device kernels (C ABI) the partitioned CG chains; tests substitute a CPU stand-in under gloo.

    Every kernel launch goes through ``_emit``.  With ``record`` set to a list the launches are not executed but
    stored as ``(fn, ctypes-ready argument tuple)`` - :class:`DistAMG` traces its V-cycle once that way and replays
    the tuples afterwards, which removes the per-launch argument marshalling (~15 us x ~70 launches per iteration)
    from the partitioned solver's critical path."""

    def __init__(self):
        self.record = None

    def _emit(self, what, fn, *args):
        if self.record is not None:
            self.record.append((fn, args, what))
        else:
            L.check(fn(*args), what)

    def workspace(self, device):
        torch = L.require_cuda()
        return torch.zeros(L.lib().mwx_dcg_workspace_doubles(), dtype=torch.float64, device=device)

    def zeros(self, n, device):
        return L.require_cuda().zeros(n, dtype=L.require_cuda().float64, device=device)

    def gather(self, v, idx, out):
        self._emit('mwx_gather_f64', L.lib().mwx_gather_f64, idx.numel(), L.ptr(v), L.ptr(idx), L.ptr(out), L.stream())

    def diag_inv(self, p):
        torch = L.require_cuda()
        d = torch.empty(p.nloc, dtype=torch.float64, device=p.vals.device)
        L.check(L.lib().mwx_csr_diagonal(p.nloc, L.ptr(p.indptr), L.ptr(p.cols), L.ptr(p.vals), L.ptr(d), L.stream()),
                'mwx_csr_diagonal')
        return 1.0 / d

    def dot(self, n, a, b, ws, slot):
        self._emit('mwx_dcg_dot', L.lib().mwx_dcg_dot, n, L.ptr(a), L.ptr(b), L.ptr(ws), slot, L.stream())

    def p_update(self, n, r, dinv, z, p_ext, ws, first):
        self._emit('mwx_dcg_p_update', L.lib().mwx_dcg_p_update, n, L.ptr(r), L.ptr(dinv), L.ptr(z), L.ptr(p_ext),
                   L.ptr(ws), int(first), L.stream())

    def spmv_dot(self, p, x_ext, q, ws):
        if getattr(p, 'lx', None):
            return self._emit('mwx_dcg_spmv_dot_lx', L.lib().mwx_dcg_spmv_dot_lx, p.nloc, p.spmv_hint, L.ptr(p.indptr),
                              L.ptr(p.cols), L.ptr(p.vals), p.lx, L.ptr(x_ext), L.ptr(q), L.ptr(ws), L.stream())
        self._emit('mwx_dcg_spmv_dot', L.lib().mwx_dcg_spmv_dot, p.nloc, p.spmv_hint, L.ptr(p.indptr), L.ptr(p.cols),
                   L.ptr(p.vals), L.ptr(x_ext), L.ptr(q), L.ptr(ws), L.stream())

    def xr_update(self, n, p_ext, q, x_ext, r, dinv, ws, jacobi):
        self._emit('mwx_dcg_xr_update', L.lib().mwx_dcg_xr_update, n, L.ptr(p_ext), L.ptr(q), L.ptr(x_ext), L.ptr(r),
                   L.ptr(dinv), L.ptr(ws), int(jacobi), L.stream())

    def residual(self, p, x_ext, b, r, dinv, ws):
        self._emit('mwx_dcg_residual', L.lib().mwx_dcg_residual, p.nloc, L.ptr(p.indptr), L.ptr(p.cols), L.ptr(p.vals),
                   L.ptr(x_ext), L.ptr(b), L.ptr(r), L.ptr(dinv), L.ptr(ws), L.stream())

    def spmv(self, p, x_ext, y):
        if getattr(p, 'bsr', None):
            return self._emit('mwx_bsr_spmv', L.lib().mwx_bsr_spmv, p.bsr, L.ptr(x_ext), L.ptr(y), L.stream())
        if getattr(p, 'pk', None) is not None:
            return self._emit('mwx_spmv_pk', L.lib().mwx_spmv_pk, p.nloc, p.spmv_hint, L.ptr(p.indptr), L.ptr(p.pk),
                              L.ptr(x_ext), L.ptr(y), L.stream())
        if getattr(p, 'lx', None):
            return self._emit('mwx_spmv_lx', L.lib().mwx_spmv_lx, p.nloc, p.spmv_hint, L.ptr(p.indptr), L.ptr(p.cols),
                              L.ptr(p.vals), p.lx, L.ptr(x_ext), L.ptr(y), L.stream())
        self._emit('mwx_spmv', L.lib().mwx_spmv, p.nloc, p.spmv_hint, L.ptr(p.indptr), L.ptr(p.cols), L.ptr(p.vals),
                   L.ptr(x_ext), L.ptr(y), L.stream())

    def spmv_axpy(self, p, x_ext, c, sign, y):
        if getattr(p, 'bsr', None):
            return self._emit('mwx_bsr_spmv_axpy', L.lib().mwx_bsr_spmv_axpy, p.bsr, L.ptr(x_ext), L.ptr(c), int(sign),
                              L.ptr(y), L.stream())
        if getattr(p, 'pk', None) is not None:
            return self._emit('mwx_spmv_axpy_pk', L.lib().mwx_spmv_axpy_pk, p.nloc, p.spmv_hint, L.ptr(p.indptr), L.ptr(p.pk),
                              L.ptr(x_ext), L.ptr(c), int(sign), L.ptr(y), L.stream())
        if getattr(p, 'lx', None):
            return self._emit('mwx_spmv_axpy_lx', L.lib().mwx_spmv_axpy_lx, p.nloc, p.spmv_hint, L.ptr(p.indptr),
                              L.ptr(p.cols), L.ptr(p.vals), p.lx, L.ptr(x_ext), L.ptr(c), int(sign), L.ptr(y), L.stream())
        self._emit('mwx_spmv_axpy', L.lib().mwx_spmv_axpy, p.nloc, p.spmv_hint, L.ptr(p.indptr), L.ptr(p.cols),
                   L.ptr(p.vals), L.ptr(x_ext), L.ptr(c), int(sign), L.ptr(y), L.stream())

    def smoother_first(self, n, b, dinv, c2, x_ext, d):
        self._emit('mwx_smoother_first', L.lib().mwx_smoother_first, n, L.ptr(b), L.ptr(dinv), float(c2), L.ptr(x_ext),
                   L.ptr(d), L.stream())

    def smoother_step(self, p, x_in, x_out, b, dinv, d, c1, c2):
        if getattr(p, 'bsr', None):
            return self._emit('mwx_bsr_smoother_step', L.lib().mwx_bsr_smoother_step, p.bsr, L.ptr(x_in), L.ptr(x_out),
                              L.ptr(b), L.ptr(dinv), L.ptr(d), float(c1), float(c2), L.stream())
        if getattr(p, 'pk', None) is not None:
            return self._emit('mwx_smoother_step_pk', L.lib().mwx_smoother_step_pk, p.nloc, p.spmv_hint, L.ptr(p.indptr),
                              L.ptr(p.pk), L.ptr(x_in), L.ptr(x_out), L.ptr(b), L.ptr(dinv), L.ptr(d), float(c1), float(c2),
                              L.stream())
        if getattr(p, 'lx', None):
            return self._emit('mwx_smoother_step_lx', L.lib().mwx_smoother_step_lx, p.nloc, p.spmv_hint, L.ptr(p.indptr),
                              L.ptr(p.cols), L.ptr(p.vals), p.lx, L.ptr(x_in), L.ptr(x_out), L.ptr(b), L.ptr(dinv), L.ptr(d),
                              float(c1), float(c2), L.stream())
        self._emit('mwx_smoother_step', L.lib().mwx_smoother_step, p.nloc, p.spmv_hint, L.ptr(p.indptr), L.ptr(p.cols),
                   L.ptr(p.vals), L.ptr(x_in), L.ptr(x_out), L.ptr(b), L.ptr(dinv), L.ptr(d), float(c1), float(c2),
                   L.stream())


_DEFAULT_OPS = None


def _default_ops():
    global _DEFAULT_OPS
    if _DEFAULT_OPS is None:
        _DEFAULT_OPS = CudaOps()
    return _DEFAULT_OPS


# ------------------------------------------------------------------------------------------ the partitioned PCG
def dist_pcg(parts, comm, b_list, tol=1e-8, maxiter=None, precond='jacobi', amg=None, ops=None):
    """scipy-1.4.1-semantics PCG on the row-block partitioned system.

    parts: this process's RankSystem-like blocks; b_list: their right-hand sides (owned entries, solver order);
    precond: 'jacobi' | 'none' | 'amg'; amg = a :class:`DistAMG` (global hierarchy, collective V-cycle) or a list of
    per-part AMGHierarchy objects (block-Jacobi: each rank preconditions with its own diagonal block).
    Returns (x_list, iters, info, resid)."""
    ops = ops or _default_ops()
    n_global = parts[0].partition.n
    if maxiter is None or maxiter <= 0:
        maxiter = 10 * n_global
    # Per-part CG state.  It is cached on the parts so that repeated solves reuse the symmetric-memory vectors (a miss
    # allocates and rendezvouses new ones).  The entry holds STRONG references to the communicator / preconditioner /
    # kernel set it was built for and is matched by identity against them, so a recycled id() can never alias it.
    key = (comm, precond, amg, ops)
    cache = getattr(parts[0], '_cg_state', None)
    same = cache is not None and len(cache['S']) == len(parts) and cache['key'][1] == precond and all(
        a is b for a, b in zip((cache['key'][0], cache['key'][2], cache['key'][3]), (comm, amg, ops)))
    if not same:
        if cache is not None and hasattr(cache['key'][0], 'release'):
            for st in cache['S']:
                cache['key'][0].release([st['x'], st['p'], st['xp']])
        S = []
        for p, b in zip(parts, b_list):
            dev = b.device
            S.append(dict(b=ops.zeros(p.nloc, dev), ws=ops.workspace(dev), x=comm.ext_zeros(p, dev),
                          p=comm.ext_zeros(p, dev), xp=comm.ext_zeros(p, dev), r=ops.zeros(p.nloc, dev),
                          q=ops.zeros(p.nloc, dev), z=None, dinv=None))
        cache = dict(key=key, S=S)
        parts[0]._cg_state = cache
    S = cache['S']
    for p, st, b in zip(parts, S, b_list):
        p.ops = ops
        st['b'].copy_(b)
        st['r'].copy_(b)
        st['x'].zero_(); st['p'].zero_(); st['ws'].zero_()
        st['dinv'] = ops.diag_inv(p) if precond == 'jacobi' else None
    jacobi = precond in ('jacobi', 'none')

    slot_index = {}

    def reduce_slots(slots):
        """All-reduce the listed scalar slots of every part's workspace.  A contiguous run of slots is reduced in
        place through a view; scattered slots go through a device-side gather/scatter (no host tensors either way)."""
        key = tuple(slots)
        if key[-1] - key[0] + 1 == len(key):
            views = [st['ws'][key[0]:key[-1] + 1] for st in S]
            comm.allreduce_sum(views)
            return views[0]
        packs = []
        for st in S:
            dev_key = (key, st['ws'].device)
            if dev_key not in slot_index:
                import torch
                slot_index[dev_key] = torch.tensor(list(key), dtype=torch.int64, device=st['ws'].device)
            packs.append(st['ws'].index_select(0, slot_index[dev_key]))
        comm.allreduce_sum(packs)
        for st, pk in zip(S, packs):
            st['ws'].index_copy_(0, slot_index[(key, st['ws'].device)], pk)
        return packs[0]

    for p, st in zip(parts, S):
        ops.dot(p.nloc, st['b'], st['b'], st['ws'], SLOT_AUX)
    bnrm = math.sqrt(float(reduce_slots([SLOT_AUX])[0]))
    # the iterate lives in the cached state: hand out copies, or the next solve would overwrite a returned solution
    xs = lambda: [st['x'][:p.nloc].clone() for p, st in zip(parts, S)]
    if bnrm <= tol:
        return xs(), 0, 0, bnrm
    atol = tol * bnrm if bnrm != 0.0 else tol
    if jacobi:
        for p, st in zip(parts, S):                              # x = 0: r = b, rr, rho = sum dinv r^2
            ops.residual(p, st['x'], st['b'], st['r'], st['dinv'], st['ws'])
        reduce_slots([SLOT_RHO, SLOT_RR])
    resid = bnrm

    def iteration(first):
        """One CG iteration: kernels, halo exchanges and all-reduces only - no host read-back (graph-capturable)."""
        if not jacobi:
            zs = amg.apply([st['r'] for st in S]) if hasattr(amg, 'apply') else [h.precondition(st['r']) for h, st in zip(amg, S)]
            for p, st, z in zip(parts, S, zs):
                st['z'] = z
                ops.dot(p.nloc, st['r'], st['z'], st['ws'], SLOT_RHO)
            reduce_slots([SLOT_RHO])
        for p, st in zip(parts, S):
            ops.p_update(p.nloc, st['r'], st['dinv'], st['z'] if not jacobi else None, st['p'], st['ws'], first)
        comm.halo(parts, [st['p'] for st in S])
        for p, st in zip(parts, S):
            ops.spmv_dot(p, st['p'], st['q'], st['ws'])
        reduce_slots([SLOT_PQ])
        for p, st in zip(parts, S):
            ops.xr_update(p.nloc, st['p'], st['q'], st['x'], st['r'], st['dinv'], st['ws'], jacobi)
        reduce_slots([SLOT_RHO, SLOT_RR] if jacobi else [SLOT_RR])

    # (A CUDA-graph replay of `iteration` was tried: capture works, but replayed symmetric-memory barriers mixed
    #  with eager ones deadlock, and a per-solve capture costs more than the ~0.5 ms of launches it saves.)
    #
    # The stopping test lags ONE iteration behind the launches: ||r_k|| travels to pinned host memory behind
    # iteration k while the host is already queueing iteration k + 1, so the device never waits for the ~100 ctypes
    # launches of an iteration (the blocking read-back used to expose them on every iteration - at 8 GPUs a third
    # of the iteration time).  The iterate of iteration k is kept in `xp` (one device copy per iteration), so the
    # returned x, the iteration count and the true-residual re-check are exactly those of the unlagged loop.
    import torch
    on_gpu = S[0]['ws'].is_cuda
    rr_host = [torch.empty(1, dtype=torch.float64, pin_memory=on_gpu) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)] if on_gpu else None
    xs_prev = lambda: [st['xp'][:p.nloc].clone() for p, st in zip(parts, S)]

    def post(k):                              # ||r_k||^2 -> host, asynchronously
        rr_host[k & 1].copy_(S[0]['ws'][SLOT_RR:SLOT_RR + 1], non_blocking=True)
        if on_gpu:
            ready[k & 1].record()

    def verdict(k, x_vecs, x_out):
        """Stopping rule of iteration k (whose iterate is x_vecs): None = go on, else the return tuple."""
        if on_gpu:
            ready[k & 1].synchronize()
        resid_k = math.sqrt(float(rr_host[k & 1][0])) if rr_host[k & 1][0] >= 0 else float('nan')
        if resid_k != resid_k:
            return x_out(), k, k, resid_k
        if resid_k <= atol:
            if k == 1:
                return x_out(), k, 0, resid_k
            saved = [st['ws'][:5].clone() for st in S]           # the scalars of the iteration already in flight
            comm.halo(parts, x_vecs)                             # true-residual re-check (scipy 1.4.1)
            for p, st, xv in zip(parts, S, x_vecs):
                ops.residual(p, xv, st['b'], st['q'], st['dinv'], st['ws'])      # q is free between iterations
            red = reduce_slots([SLOT_RHO, SLOT_RR] if jacobi else [SLOT_RR])
            true_resid = math.sqrt(float(red[-1]))
            if true_resid <= atol:
                return x_out(), k, 0, true_resid
            # Failed re-check (recurrence and true residual have drifted apart: fp64 floor).  scipy restarts from
            # the true residual here; the lagged loop keeps the recurrence of the iteration already queued and only
            # puts its scalars back - the one place where this driver is not scipy 1.4.1 step for step.
            for st, sv in zip(S, saved):
                st['ws'][:5].copy_(sv)
            verdict.last = true_resid
        else:
            verdict.last = resid_k
        return None
    verdict.last = resid

    # With a communicator whose exchanges are all kernels (PeerComm / InProcessPeerComm: `graphable`) iterations 2..
    # are ONE CUDA-graph launch each: the iteration (xp copy, V-cycle with its halos, CG kernels, all-reduces) is
    # captured once per cached solver state and replayed - no per-launch host work is left on the critical path.
    use_graph = on_gpu and getattr(comm, 'graphable', False) and os.environ.get('MWX_DIST_GRAPH', '1') != '0'

    def later_iteration():
        for p, st in zip(parts, S):
            st['xp'][:p.nloc].copy_(st['x'][:p.nloc])            # iterate of the previous iteration
        iteration(False)

    for it in range(1, maxiter + 1):
        if it == 1:
            iteration(True)
        elif use_graph:
            if cache.get('graph') is None:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, capture_error_mode='thread_local'):
                    later_iteration()
                cache['graph'] = g
            cache['graph'].replay()
        else:
            later_iteration()
        post(it)
        if it > 1:
            # NOTE: the re-check's residual kernel and all-reduce run AFTER iteration `it` was queued; they only
            # read xp / b and write q and the scalar slots, which iteration it + 1 recomputes before using them.
            out = verdict(it - 1, [st['xp'] for st in S], xs_prev)
            if out is not None:
                return out
    out = verdict(maxiter, [st['x'] for st in S], xs)
    if out is not None:
        return out
    return xs(), maxiter, maxiter, verdict.last


class PartitionedProblem:
    """Convenience wrapper: mesh/basis/Dirichlet set -> partition -> this process's blocks."""

    def __init__(self, basis, D, comm):
        from .solvers import CondensePlan
        from .assembly import DeviceCSR
        torch = L.require_cuda()
        pat = basis.pattern()
        proto = DeviceCSR(pat['indptr'], pat['indices'], torch.empty(0, dtype=torch.float64, device=basis.mesh.device),
                          (basis.N, basis.N), pat['key'], coords=basis.dof_coords)
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


# ------------------------------------------------------------------------------------------ distributed AMG V-cycle
class BlockJacobiAMG:
    """Each rank preconditions with the single-GPU V-cycle of its own diagonal block (no communication; the
    iteration count grows with the number of blocks because the coarse correction stops at the cuts)."""

    def __init__(self, hierarchies):
        self.h = hierarchies

    def apply(self, r_list):
        return [h.precondition(r) for h, r in zip(self.h, r_list)]


class _DeferredComm:
    """Routes halo exchanges through DistAMG._host_step so that they are deferred while a cycle is traced."""

    def __init__(self, comm, host_step):
        self.comm, self.host_step = comm, host_step

    def halo(self, parts, vecs):
        parts, vecs = list(parts), list(vecs)
        self.host_step(lambda: self.comm.halo(parts, vecs))


class DistAMG:
    """The GLOBAL Ruge-Stueben hierarchy (same C++ host setup as the single-GPU solver, built once on the root
    rank from the solver-numbered matrix) applied as a collective V(1,1) cycle.

    Levels with more than ``switch_rows`` rows are row-blocked like the fine matrix (coarse DOFs inherit the owner
    of their C point, so every level's ownership ranges stay contiguous); A_l, P_l, R_l each get their own
    [owned | ghost] column numbering and halo plan.  From the first level at or below ``switch_rows`` on, the
    residual is all-gathered and every rank runs the remaining levels redundantly with the single-GPU V-cycle.
    Smoother = Chebyshev (fast mode); iteration counts match the single-GPU fast mode to round-off."""

    def __init__(self, parts, comm, global_matrix, mode='chebyshev', degree=2, ratio=10.0, switch_rows=None,
                 max_levels=10, ops=None, setup='device'):
        """mode 'chebyshev': the reference's Ruge-Stueben hierarchy; 'sa': smoothed aggregation (the global matrix
        must carry near-nullspace candidates, see AMGHierarchy).  Chebyshev smoothing either way."""
        torch = L.require_cuda()
        from .solvers import AMGHierarchy
        from .assembly import DeviceCSR
        if mode not in ('chebyshev', 'sa'):
            raise NotImplementedError("the partitioned V-cycle implements the Chebyshev smoother (modes 'chebyshev', 'sa')")
        if switch_rows is None:
            switch_rows = int(os.environ.get('MWX_DIST_SWITCH_ROWS', 500000))
        self.parts, self.comm, self.ops = parts, comm, ops or CudaOps()
        self.degree, self.ratio = int(degree), float(ratio)
        self._programs, self._tail_out = {}, None
        # mode 'sa': operators of the hierarchy stored in fp32 as on one GPU (AMGHierarchy._default_storage)
        self.storage_bits = 32 if (mode == 'sa' and os.environ.get('MWX_SA_FP32', '1') != '0') else 64
        dev = parts[0].vals.device
        world = comm.world
        offsets0 = parts[0].partition.offsets
        # ---- global hierarchy.  mode 'sa' (default setup 'device'): EVERY rank builds it on its own GPU with the
        # deterministic device setup (bit-identical on all ranks, 1-2 s at 67 M DOFs, no broadcast, no idle GPUs).
        # Ruge-Stueben, or setup='host': the root builds it on its host and ships it level by level.
        replicated = mode == 'sa' and setup == 'device'
        payload, meta = None, None
        if replicated or comm.is_root():
            t_h = time.perf_counter()
            host = AMGHierarchy(global_matrix() if callable(global_matrix) else global_matrix, mode=mode,
                                max_levels=max_levels, reorder=False, upload=False, setup=('device' if replicated else 'host'))
            self.setup_host_seconds = host.setup_host_seconds
            self.setup_device_seconds = host.setup_device_seconds
            sizes = host.level_sizes()
            nl = len(sizes)
            lsw = nl - 1
            for l in range(1, nl):
                if sizes[l] <= switch_rows:
                    lsw = l
                    break
            offs = [list(offsets0)]
            for l in range(lsw):
                offs.append(host.coarse_offsets(l, offs[l]))
            payload = [torch.tensor([nl, lsw] + [o for lv in offs for o in lv], dtype=torch.int64, device=dev)]

            def put(l, which):
                payload.extend(host.level_tensors(l, which))
            for l in range(lsw):
                if l > 0:
                    put(l, 0)
                put(l, 1)
                put(l, 2)
            put(lsw, 0)
            if mode == 'sa':                   # the replicated tail continues the aggregation from level lsw
                payload.append(torch.from_numpy(host.level_candidates(lsw)).to(dev))
            del host
            self.setup_global_seconds = time.perf_counter() - t_h
        if not replicated:
            payload = comm.broadcast_tensors(payload)
        head = payload[0].cpu().numpy()
        nl, lsw = int(head[0]), int(head[1])
        offs = [list(map(int, head[2 + l * (world + 1):2 + (l + 1) * (world + 1)])) for l in range(lsw + 1)]
        self.nlevels, self.lsw, self.offsets = nl, lsw, offs
        it = iter(payload[1:])
        trip = lambda: (next(it), next(it), next(it))
        self.levels = []
        for l in range(lsw):
            A_full = trip() if l > 0 else None
            P_full, R_full = trip(), trip()
            lv = dict(A=[], P=[], R=[])
            for p in parts:
                r = p.rank
                # mode 'sa' on three candidates: every operator below the fine matrix comes in 3-wide blocks (bsr.cu)
                wide = 3 if mode == 'sa' else None
                blk = (lambda br, bc: (br, bc)) if wide else (lambda br, bc: None)
                fine = 1 if l == 0 else wide
                f32 = self.storage_bits == 32
                if l == 0 and f32 and os.environ.get('MWX_PACKED_CSR', '0') != '0':
                    # (A/B switch) the V-cycle smooths with the fp32-stored twin of the rank's fine rows; CG keeps multiplying
                    # with the fp64 values of `p` itself (a shallow copy: same halo plan, own kernel arguments)
                    import copy
                    a0 = copy.copy(p)
                    a0.pk = pack_fp32(p.cols, p.vals)
                    lv['A'].append(a0)
                else:
                    lv['A'].append(p if l == 0 else LocalOp(*A_full, offs[l][r], offs[l][r + 1], offs[l], r,
                                                            block=blk(wide, wide), fp32=f32))
                lv['P'].append(LocalOp(*P_full, offs[l][r], offs[l][r + 1], offs[l + 1], r, block=blk(fine, wide), fp32=f32))
                lv['R'].append(LocalOp(*R_full, offs[l + 1][r], offs[l + 1][r + 1], offs[l], r, block=blk(wide, fine), fp32=f32))
            for key in ('A', 'P', 'R'):
                if not (key == 'A' and l == 0):
                    comm.exchange_requests(lv[key])
            self.levels.append(lv)
            del A_full, P_full, R_full
        A_sw = trip()
        n_sw = A_sw[0].numel() - 1
        tail = DeviceCSR(A_sw[0], A_sw[1], A_sw[2], (n_sw, n_sw))
        if mode == 'sa':
            cand_sw = next(it)
            self.tail = AMGHierarchy(tail, mode='sa', degree=self.degree, max_levels=max(1, max_levels - lsw),
                                     reorder=False, candidates=cand_sw, block_size=(cand_sw.shape[1] if lsw > 0 else 1),
                                     first_level=lsw)
        else:
            self.tail = AMGHierarchy(tail, mode='chebyshev', degree=self.degree, max_levels=max(1, max_levels - lsw),
                                     reorder=False)
        self.tail.set_chebyshev(self.degree, self.ratio)
        self._tail_matrix = tail
        # Cycle shape of mode 'sa' as on one GPU (AMGHierarchy._default_cycle): GLOBAL levels 2 .. `last` (the last one
        # with at least 10 000 rows) are visited twice per visit of their parent.  Distributed levels are revisited by
        # _cycle, the first replicated level inside the tail step, the rest by the tail hierarchy's own cycle setting.
        self.revisit = ()
        # Only where a rank's share of the fine level is large (>= 2^24 rows): the revisited levels are small, their cost
        # does not shrink with the number of ranks (halo latencies, the replicated tail), and at 8 GPUs / 67 M DOFs the
        # revisits cost more (8.9 ms x 25 iterations) than the plain V-cycle (4.9 ms x 40) - MWX_SA_CYCLE=k forces them.
        # Smaller shares revisit only the REPLICATED levels (inside the tail step: no extra halo or all-gather).
        cycle_env = os.environ.get('MWX_SA_CYCLE', 'auto')
        everywhere = cycle_env == 'k' or 'MWX_SA_CYCLE_MIN_ROWS' in os.environ or \
            offs[0][-1] // max(world, len(parts)) >= (1 << 24)
        if mode == 'sa' and cycle_env != 'v':
            sizes = [o[-1] for o in offs[:lsw]] + self.tail.level_sizes()
            min_rows = int(os.environ.get('MWX_SA_CYCLE_MIN_ROWS', 10000))
            last = max([g for g in range(2, len(sizes) - 1) if sizes[g] >= min_rows], default=0)
            first_small = int(os.environ.get('MWX_SA_CYCLE_FIRST_SMALL', max(2, lsw)))
            self.revisit = tuple(range(2 if everywhere else max(2, first_small), last + 1))
            self.tail.set_cycle(1, 0)
            if last - lsw >= 1 and self.revisit:
                self.tail.set_cycle(2, last - lsw, first=max(1, 2 - lsw))
        self._cur_ext = {}
        # ---- per level work vectors, inverse diagonals, spectral radius estimates
        ops = self.ops
        for l, lv in enumerate(self.levels):
            lv['dinv'] = [ops.diag_inv(a) for a in lv['A']]
            lv['xa'] = [comm.ext_zeros(a, dev) for a in lv['A']]
            lv['xb'] = [comm.ext_zeros(a, dev) for a in lv['A']]
            lv['d'] = [ops.zeros(a.nloc, dev) for a in lv['A']]
            lv['r'] = [comm.ext_zeros(r, dev) for r in lv['R']]                      # fine residual, R's column layout
            lv['bc'] = [ops.zeros(r.nloc, dev) for r in lv['R']]
            lv['xc'] = [comm.ext_zeros(p_, dev) for p_ in lv['P']]                  # coarse iterate, P's column layout
            lv['rho'] = self._estimate_rho(lv)
            if l in self.revisit:
                lv['g_b'] = [ops.zeros(a.nloc, dev) for a in lv['A']]
                lv['g_x'] = [ops.zeros(a.nloc, dev) for a in lv['A']]

    # power iteration on D^-1 A (collective; deterministic start)
    def _estimate_rho(self, lv):
        torch = L.require_cuda()
        ops, comm = self.ops, self.comm
        A, dinv = lv['A'], lv['dinv']
        v = lv['xa']
        w = [ops.zeros(a.nloc, v[0].device) for a in A]
        for a, vi in zip(A, v):                                  # same start vector as the single-GPU estimate
            h = np.arange(a.clo, a.clo + a.ncol, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)
            h ^= h >> np.uint64(29)
            h *= np.uint64(0xBF58476D1CE4E5B9)
            h ^= h >> np.uint64(32)
            seed = 0.5 + (h & np.uint64(0xFFFFF)).astype(np.float64) / 1048576.0
            vi[:a.ncol] = torch.from_numpy(seed).to(vi.device)
        lam = 1.0
        for it in range(20):
            sq = [torch.sum(vi[:a.ncol] ** 2).reshape(1) for a, vi in zip(A, v)]
            comm.allreduce_sum(sq)
            nrm = math.sqrt(float(sq[0]))
            if not nrm > 0:
                break
            if it > 0:
                lam = nrm
            for a, vi in zip(A, v):
                vi[:a.ncol] /= nrm
            comm.halo(A, v)
            for a, vi, wi, di in zip(A, v, w, dinv):
                ops.spmv(a, vi, wi)
                vi[:a.ncol] = di * wi
        sq = [torch.sum(vi[:a.ncol] ** 2).reshape(1) for a, vi in zip(A, v)]
        comm.allreduce_sum(sq)
        if float(sq[0]) > 0:
            lam = math.sqrt(float(sq[0]))
        return lam

    def _smooth(self, lv, cur, alt, b, zero_start, comm=None):
        """Chebyshev(degree) on D^-1 A; returns the list of buffers holding the result (cur or alt)."""
        ops = self.ops
        comm = comm or self.comm
        lmax = 1.1 * lv['rho']
        lmin = lmax / self.ratio
        theta, delta = 0.5 * (lmax + lmin), 0.5 * (lmax - lmin)
        sigma = theta / delta
        rho_k = 1.0 / sigma
        for s in range(self.degree):
            if s == 0:
                c1, c2 = 0.0, 1.0 / theta
            else:
                rho_n = 1.0 / (2.0 * sigma - rho_k)
                c1, c2 = rho_n * rho_k, 2.0 * rho_n / delta
                rho_k = rho_n
            if s == 0 and zero_start:
                for a, x, bi, di, d in zip(lv['A'], cur, b, lv['dinv'], lv['d']):
                    ops.smoother_first(a.nloc, bi, di, c2, x, d)
            else:
                comm.halo(lv['A'], cur)
                for a, xi, xo, bi, di, d in zip(lv['A'], cur, alt, b, lv['dinv'], lv['d']):
                    ops.smoother_step(a, xi, xo, bi, di, d, c1, c2)
                cur, alt = alt, cur
        return cur, alt

    def _host_step(self, fn):
        """A step that is not a single kernel launch (halo exchange, all-gather, tail cycle): runs now, or is
        deferred as a thunk while the cycle is being traced."""
        rec = getattr(self.ops, 'record', None)
        if rec is not None:
            rec.append((fn, None, 'host step'))
        else:
            fn()

    def _cycle(self, l, b):
        ops = self.ops
        comm = _DeferredComm(self.comm, self._host_step)
        if l == self.lsw:
            torch = L.require_cuda()
            if self._tail_out is None:
                self._tail_out = [torch.empty(self.offsets[l][p.rank + 1] - self.offsets[l][p.rank], dtype=torch.float64,
                                              device=p.vals.device) for p in self.parts]

            def tail_step(b=b, l=l):
                full = self.comm.allgather(b, self.offsets[l])
                for part_idx, f in enumerate(full):
                    z = self.tail.precondition(f)
                    if l in self.revisit:                       # second visit of the first replicated level
                        z = z + self.tail.precondition(f - self._tail_matrix.dot(z))
                    r = self.parts[part_idx].rank
                    self._tail_out[part_idx].copy_(z[self.offsets[l][r]:self.offsets[l][r + 1]])
            self._host_step(tail_step)
            return self._tail_out
        lv = self.levels[l]
        cur, alt = self._smooth(lv, lv['xa'], lv['xb'], b, True, comm)
        comm.halo(lv['A'], cur)
        for a, x, bi, r in zip(lv['A'], cur, b, lv['r']):
            ops.spmv_axpy(a, x, bi, -1, r)                      # r[:nloc] = b - A x   (r is laid out for R's halo)
        comm.halo(lv['R'], lv['r'])
        for R, r, bc in zip(lv['R'], lv['r'], lv['bc']):
            ops.spmv(R, r, bc)
        xc = self._cycle(l + 1, lv['bc'])
        if l + 1 < self.lsw and (l + 1) in self.revisit:        # second visit: x_c += M^-1 (b_c - A x_c), still SPD
            lvm = self.levels[l + 1]
            cur_m = self._cur_ext[l + 1]
            comm.halo(lvm['A'], cur_m)
            for a, x, bi, gb in zip(lvm['A'], cur_m, lv['bc'], lvm['g_b']):
                ops.spmv_axpy(a, x, bi, -1, gb)

            def save(xc=xc, lvm=lvm):                           # the nested cycle reuses the level's work vectors
                for gx, x_c in zip(lvm['g_x'], xc):
                    gx.copy_(x_c)
            self._host_step(save)
            x2 = self._cycle(l + 1, lvm['g_b'])

            def accumulate(x2=x2, lvm=lvm):
                for gx, d in zip(lvm['g_x'], x2):
                    gx.add_(d)
            self._host_step(accumulate)
            xc = lvm['g_x']

        def fill_coarse(lv=lv, xc=xc):
            for P, xce, x_c in zip(lv['P'], lv['xc'], xc):
                xce[:P.ncol] = x_c
        self._host_step(fill_coarse)
        comm.halo(lv['P'], lv['xc'])
        for P, xce, x in zip(lv['P'], lv['xc'], cur):
            ops.spmv_axpy(P, xce, x, +1, x)                     # x += P x_c (owned rows)
        cur, alt = self._smooth(lv, cur, alt, b, False, comm)
        self._cur_ext[l] = cur
        return [x[:a.nloc] for a, x in zip(lv['A'], cur)]

    def apply(self, r_list):
        """z = M^-1 r (collective).  The cycle is traced once per set of input buffers into a list of
        (kernel, prebuilt arguments) / host-step entries and replayed afterwards."""
        if not hasattr(self.ops, 'record') or os.environ.get('MWX_DIST_REPLAY', '1') == '0':   # stand-in ops / A-B switch
            return self._cycle(0, r_list)
        torch = L.require_cuda()
        if torch.cuda.is_current_stream_capturing():            # graph capture: launch on the capturing stream
            return self._cycle(0, r_list)
        key = tuple(r.data_ptr() for r in r_list)
        if key not in self._programs:
            self.ops.record = []
            try:
                out = self._cycle(0, r_list)
                self._programs[key] = (self.ops.record, out, r_list)
            finally:
                self.ops.record = None
        prog, out, _keep = self._programs[key]
        for fn, args, what in prog:
            if args is None:
                fn()
            else:
                rc = fn(*args)
                if rc:
                    L.check(rc, what)
        return out
