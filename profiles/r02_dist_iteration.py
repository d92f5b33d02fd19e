"""Partitioned PCG iteration with all P parts in ONE process on one GPU (InProcessPeerComm: the same peer kernels, signal-only
barriers): per-iteration time eager / graph-replayed, and - under `ncu --profile-from-start off` - the launch list of the
last iterations.   python profiles/r02_dist_iteration.py [level] [parts] [maxiter]"""
import json, os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import fem_forth_perturb_b200 as F
from fem_forth_perturb_b200 import dist as D
from fem_forth_perturb_b200.solvers import Reordering

level = int(sys.argv[1]) if len(sys.argv) > 1 else 11
nparts = int(sys.argv[2]) if len(sys.argv) > 2 else 2
maxiter = int(sys.argv[3]) if len(sys.argv) > 3 else 12
gm = F.MeshTri().refine(level)
basis = F.InteriorBasis(gm, F.ElementTriMorley(), intorder=5)
bf = gm.boundary_facets()
Dset = np.concatenate((np.unique(gm.facets[:, bf]), bf + gm.nv))
form = 0.01 ** 2 * F.a_load + F.b_load
comm = D.InProcessPeerComm(nparts)
prob = D.PartitionedProblem(basis, Dset, comm).assemble(form)

def global_matrix():
    Kc, _ = prob.cplan.apply(F.asm_gpu(form, basis))
    ro = Reordering.__new__(Reordering)
    ro.perm, ro.iperm, ro.n = prob.partition.perm, prob.partition.iperm, prob.partition.n
    return ro.matrix(Kc)

ml = D.DistAMG(prob.parts, comm, global_matrix, mode='sa')
torch.cuda.empty_cache()
g = torch.Generator(device='cuda').manual_seed(1234)
b_list = [torch.rand(p.nloc, dtype=torch.float64, device='cuda', generator=g) * 2 - 1 for p in prob.parts]
out = dict(level=level, parts=nparts, n=prob.partition.n, levels=[o[-1] for o in ml.offsets] + ml.tail.level_sizes()[1:],
           distributed_levels=ml.lsw, revisit=list(ml.revisit),
           bsr_operators=sum(1 for lv in ml.levels for k in 'APR' for op in lv[k] if getattr(op, 'bsr', None)))
for graph in ('1', '0'):
    os.environ['MWX_DIST_GRAPH'] = graph
    D.dist_pcg(prob.parts, comm, b_list, tol=1e-30, maxiter=4, precond='amg', amg=ml)      # warm (captures the graph)
    torch.cuda.synchronize(); t = time.perf_counter()
    if graph == '0':
        torch.cuda.profiler.start()
    xs, its, info, _ = D.dist_pcg(prob.parts, comm, b_list, tol=1e-30, maxiter=maxiter, precond='amg', amg=ml)
    torch.cuda.synchronize()
    if graph == '0':
        torch.cuda.profiler.stop()
    out['ms_per_iter_graph' if graph == '1' else 'ms_per_iter_eager'] = 1e3 * (time.perf_counter() - t) / its
print(json.dumps(out), flush=True)
