import os, sys, time, torch, numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import fem_forth_perturb_b200 as F
from fem_forth_perturb_b200 import _lib as L
lvl = int(sys.argv[1]) if len(sys.argv) > 1 else 11
def timeit(fn, n=5, w=3):
    for _ in range(w): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for name in ('refine-order', 'hilbert-renumbered'):
    gm = F.MeshTri().refine(lvl)
    if name != 'refine-order':
        t0 = time.perf_counter(); gm = gm.renumbered(); torch.cuda.synchronize(); print('renumber s', time.perf_counter() - t0)
    basis = F.InteriorBasis(gm, F.ElementTriMorley())
    form = 1e-4 * F.a_load + F.b_load
    K = F.asm_gpu(form, basis)
    ms = timeit(lambda: F.asm_gpu(form, basis, out=K.d_data))
    x = torch.rand(basis.N, dtype=torch.float64, device='cuda'); y = torch.empty_like(x)
    ms2 = timeit(lambda: K.dot(x))
    print(name, 'level', lvl, 'assembly ms', round(ms, 3), 'spmv ms', round(ms2, 3), 'N', basis.N, flush=True)
    if lvl <= 6:
        import scipy.sparse as sp
        A = K.tocsr(); print('sym err', abs(A - A.T).max(), 'rowsum', abs(A @ np.ones(A.shape[0])).max())
    del K, basis, gm, x, y; torch.cuda.empty_cache()
