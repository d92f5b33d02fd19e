"""Partitioned (multi-GPU) path.

* CPU, world_size 2, gloo: the driver's communication + orchestration (request exchange, halo, all-reduces,
  scipy-1.4.1 stopping logic) with a NumPy stand-in for the device kernels (test infrastructure only).
* GPU, one device: all P parts in one process (InProcessComm) - the partitioned operator is exactly the
  permuted global one and the iteration counts do not depend on P.
* GPU, >= 2 devices: the same through torchrun + NCCL (skipped on a single-GPU box).
"""
import math
import os
import subprocess
import sys

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import ROOT

CPU_WORKER = r'''
import os, sys, math
sys.path.insert(0, os.environ["MWX_ROOT"])
import numpy as np, scipy.sparse as sp, torch, torch.distributed as dist
from fem_forth_perturb_b200 import dist as D

class Part:                                     # CPU stand-in for RankSystem (same attributes)
    def __init__(self, A, offsets, rank):
        self.rank, self.lo, self.hi = rank, offsets[rank], offsets[rank + 1]
        self.nloc = self.hi - self.lo
        self.clo, self.ncol = self.lo, self.nloc
        blk = A[self.lo:self.hi].tocsr(); blk.sort_indices()
        cols = blk.indices.astype(np.int64)
        off = (cols < self.lo) | (cols >= self.hi)
        gh = np.unique(cols[off])
        self.ghosts = torch.from_numpy(gh.astype(np.int32)); self.nghost = len(gh)
        loc = np.where(off, self.nloc + np.searchsorted(gh, cols), cols - self.lo)
        self.indptr, self.cols, self.vals = blk.indptr, loc, blk.data
        owner = np.searchsorted(np.array(offsets[1:]), gh, side="right")
        self.recv = {int(o): (int(np.nonzero(owner == o)[0][0]), int((owner == o).sum())) for o in np.unique(owner)}
        self.send = {}
        class Pt: pass
        self.partition = Pt(); self.partition.n = A.shape[0]
    def requests(self): return {o: self.ghosts[s:s + c] for o, (s, c) in self.recv.items()}
    def set_send_lists(self, asked): self.send = {o: (ids - self.lo).to(torch.int32) for o, ids in asked.items()}
    def matrix(self): return sp.csr_matrix((self.vals, self.cols, self.indptr), shape=(self.nloc, self.nloc + self.nghost))

class NumpyOps:                                 # the kernels' contract, restated on the CPU for this test
    def workspace(self, device): return torch.zeros(16, dtype=torch.float64)
    def zeros(self, n, device): return torch.zeros(n, dtype=torch.float64)
    def gather(self, v, idx, out): out.copy_(v[idx.long()])
    def diag_inv(self, p): return torch.from_numpy(1.0 / p.matrix()[:, :p.nloc].diagonal())
    def dot(self, n, a, b, ws, slot): ws[slot] = float(a[:n] @ b[:n])
    def p_update(self, n, r, dinv, z, p_ext, ws, first):
        zz = z if z is not None else (dinv * r if dinv is not None else r)
        p_ext[:n] = zz if first else zz + (ws[0] / ws[1]) * p_ext[:n]
    def spmv_dot(self, p, x_ext, q, ws):
        q.copy_(torch.from_numpy(p.matrix() @ x_ext.numpy())); ws[2] = float(x_ext[:p.nloc] @ q)
    def xr_update(self, n, p_ext, q, x_ext, r, dinv, ws, jacobi):
        alpha = ws[0] / ws[2]
        x_ext[:n] += alpha * p_ext[:n]; r -= alpha * q
        ws[3] = float(r @ r)
        rho_new = float(((dinv if dinv is not None else 1.0) * r) @ r)
        ws[1] = ws[0].clone()
        if jacobi: ws[0] = rho_new
    def residual(self, p, x_ext, b, r, dinv, ws):
        r.copy_(b - torch.from_numpy(p.matrix() @ x_ext.numpy()))
        ws[3] = float(r @ r); ws[0] = float(((dinv if dinv is not None else 1.0) * r) @ r)

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n = 60
T = sp.diags([-np.ones(n - 1), 2.2 * np.ones(n), -np.ones(n - 1)], [-1, 0, 1])
A = (sp.kron(sp.identity(n), T) + sp.kron(T, sp.identity(n))).tocsr()          # 2-D Laplacian-like, 3600 rows
b = np.random.default_rng(0).standard_normal(A.shape[0])
offsets = [0, 1700, A.shape[0]]                                                 # uneven split
part = Part(A, offsets, rank)
comm = D.TorchDistComm()
comm.exchange_requests([part])
part.ops = NumpyOps()
v = torch.zeros(part.nloc + part.nghost, dtype=torch.float64)
v[:part.nloc] = torch.arange(part.lo, part.hi, dtype=torch.float64)
comm.halo([part], [v])
assert np.array_equal(v[part.nloc:].numpy(), part.ghosts.numpy().astype(float)), "halo delivers the owners' values"
xs, its, info, resid = D.dist_pcg([part], comm, [torch.from_numpy(b[part.lo:part.hi].copy())], tol=1e-10,
                                  precond="jacobi", ops=NumpyOps())
sys.path.insert(0, os.environ["MWX_ROOT"])
from oracle.scipy_cg import cg141, build_pc_diag
x_ref, info_ref, it_ref = cg141(A, b, M=build_pc_diag(A), tol=1e-10)
err = np.linalg.norm(xs[0].numpy() - x_ref[part.lo:part.hi]) / np.linalg.norm(x_ref)
assert info == 0 and abs(its - it_ref) <= 1 and err < 1e-9, (its, it_ref, err)
xs2, its2, info2, _ = D.dist_pcg([part], comm, [torch.from_numpy(b[part.lo:part.hi].copy())], tol=1e-30, maxiter=5,
                                 precond="none", ops=NumpyOps())
assert its2 == 5 and info2 == 5
sys.stdout.write(f"RANK{rank}-OK its={its} ref={it_ref}\n"); sys.stdout.flush()
dist.destroy_process_group()
'''


def _torchrun(script, nproc, extra_env=None, timeout=300):
    env = dict(os.environ, MWX_ROOT=ROOT, **(extra_env or {}))
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={nproc}',
           '--master-addr', '127.0.0.1', '--master-port', str(29500 + os.getpid() % 400), script]
    return subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=timeout)


def test_partitioned_pcg_driver_gloo_world2(tmp_path):
    script = tmp_path / 'worker.py'
    script.write_text(CPU_WORKER)
    res = _torchrun(str(script), 2)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert res.stdout.count('-OK') == 2


def test_rcb_partition_is_balanced_and_compact():
    import torch
    from fem_forth_perturb_b200.dist import rcb_parts
    g = torch.Generator().manual_seed(0)
    xy = torch.rand(10007, 2, dtype=torch.float64, generator=g)
    for P in (2, 3, 4, 8):
        part = rcb_parts(xy, P)
        counts = torch.bincount(part, minlength=P)
        assert int(counts.max() - counts.min()) <= 1 + P // 2
        # boxes of different parts do not overlap (RCB cells are axis-aligned rectangles)
        boxes = [(xy[part == q].min(0).values, xy[part == q].max(0).values) for q in range(P)]
        area = sum(float((hi - lo).prod()) for lo, hi in boxes)
        assert area <= 1.0 + 1e-9


# ------------------------------------------------------------------ GPU
def _setup(level, nparts, eps2=1e-4, peer=False):
    import fem_forth_perturb_b200 as F
    from fem_forth_perturb_b200 import dist as D
    gm = F.MeshTri().refine(level)
    basis = F.InteriorBasis(gm, F.ElementTriMorley())
    bf = gm.boundary_facets()
    Dset = np.concatenate((np.unique(gm.facets[:, bf]), bf + gm.nv))
    form = eps2 * F.a_load + F.b_load
    prob = D.PartitionedProblem(basis, Dset, (D.InProcessPeerComm if peer else D.InProcessComm)(nparts)).assemble(form)
    K2 = F.asm_gpu(form, basis)
    Kc = F.condense(K2, D=Dset, expand=False)
    return F, D, prob, Kc


@pytest.mark.gpu
@pytest.mark.parametrize('nparts', [1, 2, 3, 4])
def test_partitioned_operator_equals_permuted_global(cuda, nparts):
    F, D, prob, Kc = _setup(5, nparts)
    perm = prob.partition.perm.cpu().numpy()
    ref = Kc.tocsr()[perm][:, perm].tocsr()
    stacked = sp.vstack([p.to_scipy_global() for p in prob.parts]).tocsr()
    stacked.sort_indices(); ref.sort_indices()
    assert np.array_equal(stacked.indptr, ref.indptr) and np.array_equal(stacked.indices, ref.indices)
    assert abs(stacked - ref).max() <= 1e-13 * abs(ref).max()     # rows assembled by the row-list kernel
    assert prob.partition.offsets[-1] == Kc.shape[0]
    sizes = np.diff(prob.partition.offsets)
    assert sizes.max() - sizes.min() <= 1 + nparts
    if nparts > 1:
        ghosts = sum(p.nghost for p in prob.parts)
        assert 0 < ghosts < 0.25 * Kc.shape[0]                      # compact parts: the halo is a thin layer


@pytest.mark.gpu
@pytest.mark.parametrize('peer', [False, True])
def test_partitioned_pcg_independent_of_partition_count(cuda, peer):
    torch = cuda
    its_seen = {}
    for nparts in (1, 2, 4):
        F, D, prob, Kc = _setup(5, nparts, eps2=1e-4, peer=peer)
        rng = np.random.default_rng(3)
        b_old = rng.standard_normal(Kc.shape[0])
        perm = prob.partition.perm.cpu().numpy()
        b_new = torch.from_numpy(b_old[perm]).cuda()
        b_list = [b_new[p.lo:p.hi].clone() for p in prob.parts]
        xs, its, info, resid = D.dist_pcg(prob.parts, prob.comm, b_list, tol=1e-8, precond='jacobi')
        assert info == 0
        x_new = torch.cat(xs).cpu().numpy()
        x_old = np.empty_like(x_new); x_old[perm] = x_new
        x_ref, it_ref = F.solver_iter_krylov(Precondition=True, tol=1e-8).solve_with_info(Kc, b_old)
        assert abs(its - it_ref) <= 1, (nparts, its, it_ref)
        assert np.linalg.norm(x_old - x_ref) <= 1e-7 * np.linalg.norm(x_ref)
        its_seen[nparts] = its
        # block-Jacobi AMG: converges, to the same solution
        amg = [F.AMGHierarchy(p.diagonal_block(), mode='chebyshev') for p in prob.parts]
        xs, its_a, info, _ = D.dist_pcg(prob.parts, prob.comm, b_list, tol=1e-10, precond='amg', amg=amg)
        assert info == 0 and its_a < 200
        x_new = torch.cat(xs).cpu().numpy()
        x_old = np.empty_like(x_new); x_old[perm] = x_new
        x_d = sp.linalg.spsolve(Kc.tocsr().tocsc(), b_old)
        assert np.linalg.norm(x_old - x_d) <= 1e-7 * np.linalg.norm(x_d)
    assert max(its_seen.values()) - min(its_seen.values()) <= 1


@pytest.mark.gpu
@pytest.mark.parametrize('peer', [False, True])
@pytest.mark.parametrize('min_rows,switch_rows', [(10000, 600), (50, 100), (50, 600), (50, 3000)])
def test_distributed_smoothed_aggregation_matches_single_gpu(cuda, monkeypatch, min_rows, switch_rows, peer):
    """mode='sa' through the partitioned V-cycle: aggregates are numbered by seed, so coarse levels stay row-blocked.
    min_rows = 50 turns the gamma = 2 revisit of level 2 on at this problem size: with that level row-blocked
    (switch_rows 100), as the first replicated level (600: revisited inside the tail step) and as a level inside the
    replicated tail hierarchy (3000)."""
    torch = cuda
    monkeypatch.setenv('MWX_SA_CYCLE_MIN_ROWS', str(min_rows))
    monkeypatch.setenv('MWX_SA_MAX_COARSE', '300')          # keep a deep hierarchy at this small size (default: stop at <= 4000 rows)
    F, D, prob, Kc = _setup(6, 3, eps2=0.25, peer=peer)
    from fem_forth_perturb_b200.solvers import Reordering, _run_pcg
    ro = Reordering.for_matrix(Kc)
    ro.perm, ro.iperm = prob.partition.perm, prob.partition.iperm
    Kp = ro.matrix(Kc)
    assert Kp.candidates is not None
    amg = D.DistAMG(prob.parts, prob.comm, Kp, mode='sa', switch_rows=switch_rows)
    assert amg.lsw >= 1 and (len(amg.revisit) > 0) == (min_rows == 50)
    rng = np.random.default_rng(4)
    b_new = torch.from_numpy(rng.standard_normal(Kc.shape[0])).cuda()
    single = F.AMGHierarchy(Kp, mode='sa', reorder=False)
    z_ref = single.precondition(b_new)
    z = torch.cat(amg.apply([b_new[p.lo:p.hi].clone() for p in prob.parts]))
    # both hierarchies are STORED in fp32: the ranks round the rows they assembled themselves, which agree with the
    # globally assembled matrix to 1e-13 only - an entry on a rounding boundary differs by one fp32 ulp
    assert float(torch.linalg.norm(z - z_ref) / torch.linalg.norm(z_ref)) < 2e-6
    xs, its, info, _ = D.dist_pcg(prob.parts, prob.comm, [b_new[p.lo:p.hi].clone() for p in prob.parts], tol=1e-8,
                                  precond='amg', amg=amg)
    x1, its1, info1, _ = _run_pcg(Kp, b_new, 1e-8, None, amg=single, reordering=None)
    assert info == 0 and info1 == 0 and abs(its - its1) <= 1
    assert float(torch.linalg.norm(torch.cat(xs) - x1) / torch.linalg.norm(x1)) < 1e-6
    if peer:
        # peer kernels: iterations 2.. are replays of ONE captured CUDA graph; a second solve (other right-hand side)
        # reuses it, and the eager path (MWX_DIST_GRAPH=0) gives the same iterates
        assert prob.parts[0]._cg_state.get('graph') is not None
        b2 = torch.from_numpy(rng.standard_normal(Kc.shape[0])).cuda()
        xs2, its2, info2, _ = D.dist_pcg(prob.parts, prob.comm, [b2[p.lo:p.hi].clone() for p in prob.parts], tol=1e-8,
                                         precond='amg', amg=amg)
        monkeypatch.setenv('MWX_DIST_GRAPH', '0')
        xs3, its3, info3, _ = D.dist_pcg(prob.parts, prob.comm, [b2[p.lo:p.hi].clone() for p in prob.parts], tol=1e-8,
                                         precond='amg', amg=amg)
        assert info2 == 0 and info3 == 0 and its2 == its3
        assert float(torch.linalg.norm(torch.cat(xs2) - torch.cat(xs3)) / torch.linalg.norm(torch.cat(xs3))) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize('peer', [False, True])
@pytest.mark.parametrize('nparts', [1, 2, 4])
def test_distributed_vcycle_matches_single_gpu_fast_mode(cuda, nparts, peer):
    """Global hierarchy applied as a collective V-cycle: same iteration count as the single-GPU fast mode."""
    torch = cuda
    F, D, prob, Kc = _setup(6, nparts, eps2=1e-4, peer=peer)
    from fem_forth_perturb_b200.solvers import Reordering
    ro = Reordering.for_matrix(Kc)
    ro.perm, ro.iperm = prob.partition.perm, prob.partition.iperm          # the partition's numbering
    Kp = ro.matrix(Kc)
    amg = D.DistAMG(prob.parts, prob.comm, Kp, switch_rows=600)
    assert amg.lsw >= 2                                                     # at least two distributed levels
    rng = np.random.default_rng(4)
    b_old = rng.standard_normal(Kc.shape[0])
    perm = prob.partition.perm.cpu().numpy()
    b_new = torch.from_numpy(b_old[perm]).cuda()
    # the collective V-cycle equals the single-GPU one on the same (permuted) matrix
    single = F.AMGHierarchy(Kp, mode='chebyshev', reorder=False)
    z_ref = single.precondition(b_new)
    z = torch.cat(amg.apply([b_new[p.lo:p.hi].clone() for p in prob.parts]))
    assert float(torch.linalg.norm(z - z_ref) / torch.linalg.norm(z_ref)) < 1e-8
    xs, its, info, _ = D.dist_pcg(prob.parts, prob.comm, [b_new[p.lo:p.hi].clone() for p in prob.parts], tol=1e-8,
                                  precond='amg', amg=amg)
    from fem_forth_perturb_b200.solvers import _run_pcg
    x_ref, it_ref, info_ref, _ = _run_pcg(Kp, b_new, 1e-8, None, amg=single)
    assert info == 0 and info_ref == 0 and abs(its - it_ref) <= 1, (its, it_ref)
    x = torch.cat(xs)
    assert float(torch.linalg.norm(x - x_ref) / torch.linalg.norm(x_ref)) < 1e-7


GPU_WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["MWX_ROOT"])
import numpy as np, torch, torch.distributed as dist
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import fem_forth_perturb_b200 as F
from fem_forth_perturb_b200 import dist as D
gm = F.MeshTri().refine(6)
basis = F.InteriorBasis(gm, F.ElementTriMorley())
bf = gm.boundary_facets()
Dset = np.concatenate((np.unique(gm.facets[:, bf]), bf + gm.nv))
form = 1e-4 * F.a_load + F.b_load
comm = {"peer": D.PeerComm, "symm": D.SymmMemComm, "nccl": D.TorchDistComm}[os.environ["MWX_TEST_HALO"]]()
prob = D.PartitionedProblem(basis, Dset, comm).assemble(form)
(p,) = prob.parts
Kc = F.condense(F.asm_gpu(form, basis), D=Dset, expand=False)
b_old = np.random.default_rng(3).standard_normal(Kc.shape[0])
perm = prob.partition.perm.cpu().numpy()
b_new = torch.from_numpy(b_old[perm]).cuda()
xs, its, info, resid = D.dist_pcg(prob.parts, comm, [b_new[p.lo:p.hi].clone()], tol=1e-8, precond="jacobi")
x_ref, it_ref = F.solver_iter_krylov(Precondition=True, tol=1e-8).solve_with_info(Kc, b_old)
mine = xs[0].cpu().numpy(); ref = x_ref[perm][p.lo:p.hi]
assert info == 0 and abs(its - it_ref) <= 1, (its, it_ref)
assert np.linalg.norm(mine - ref) <= 1e-7 * np.linalg.norm(x_ref)
from fem_forth_perturb_b200.solvers import Reordering, _run_pcg
ro = Reordering.for_matrix(Kc); ro.perm, ro.iperm = prob.partition.perm, prob.partition.iperm
Kp = ro.matrix(Kc)
amg = D.DistAMG(prob.parts, comm, (lambda: Kp), switch_rows=600)
xs, its_a, info, _ = D.dist_pcg(prob.parts, comm, [b_new[p.lo:p.hi].clone()], tol=1e-8, precond="amg", amg=amg)
single = F.AMGHierarchy(Kp, mode="chebyshev", reorder=False)
x1, it1, info1, _ = _run_pcg(Kp, b_new, 1e-8, None, amg=single)
assert info == 0 and abs(its_a - it1) <= 1, (its_a, it1)
assert float(torch.linalg.norm(xs[0] - x1[p.lo:p.hi]) / torch.linalg.norm(x1)) < 1e-7
if os.environ["MWX_TEST_HALO"] == "peer":          # graph-captured iteration: replayed by a second solve, equal to the eager path
    assert prob.parts[0]._cg_state.get("graph") is not None
    xs2, its2, info2, _ = D.dist_pcg(prob.parts, comm, [b_new[p.lo:p.hi].clone()], tol=1e-8, precond="amg", amg=amg)
    os.environ["MWX_DIST_GRAPH"] = "0"
    xs3, its3, info3, _ = D.dist_pcg(prob.parts, comm, [b_new[p.lo:p.hi].clone()], tol=1e-8, precond="amg", amg=amg)
    assert info2 == 0 and info3 == 0 and its2 == its_a and its3 == its_a, (its_a, its2, its3)
    assert float(torch.linalg.norm(xs2[0] - xs3[0])) <= 1e-12 * float(torch.linalg.norm(xs3[0]))
sys.stdout.write(f"RANK{dist.get_rank()}-OK its={its} ref={it_ref} amg={its_a}\n"); sys.stdout.flush()
dist.destroy_process_group()
'''


@pytest.mark.gpu
@pytest.mark.parametrize('halo', ['nccl', 'symm', 'peer'])
def test_partitioned_pcg_two_gpus(cuda, tmp_path, halo):
    """halo = NCCL send/recv; pack-and-push kernels into the neighbour's peer memory + symmetric-memory barrier ('symm');
    or every exchange (halo, all-reduce, all-gather) as peer kernels with the iteration captured in a CUDA graph ('peer')."""
    if cuda.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs (run with gpurun --gpus 2)')
    script = tmp_path / 'worker_gpu.py'
    script.write_text(GPU_WORKER)
    res = _torchrun(str(script), 2, extra_env={'MWX_TEST_HALO': halo}, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert res.stdout.count('-OK') == 2
