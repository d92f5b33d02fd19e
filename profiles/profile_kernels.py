"""Short driver for ncu: level-L K2 assembly, condense, SpMV and a few PCG iterations.
    python profiles/profile_kernels.py [level] [pcg_maxiter] [amg mode: sa | chebyshev]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import fem_forth_perturb_b200 as F
from fem_forth_perturb_b200 import _lib as L
from fem_forth_perturb_b200.solvers import _run_pcg

level = int(sys.argv[1]) if len(sys.argv) > 1 else 12
maxit = int(sys.argv[2]) if len(sys.argv) > 2 else 0
amg_mode = sys.argv[3] if len(sys.argv) > 3 else 'sa'
gm = F.MeshTri().refine(level)
basis = F.InteriorBasis(gm, F.ElementTriMorley())
form = 1e-4 * F.a_load + F.b_load
K2 = F.asm_gpu(form, basis)
bf = gm.boundary_facets()
D = np.concatenate((np.unique(gm.facets[:, bf]), bf + gm.nv))
plan = F.CondensePlan(K2, D=D)
Kc, _ = plan.apply(K2)
if os.environ.get('MWX_PROFILE_REORDER', '1') == '1':
    from fem_forth_perturb_b200.solvers import Reordering
    Kc = Reordering.for_matrix(Kc).matrix(Kc)          # the numbering the fast-mode solver runs on
n = Kc.shape[0]
x = torch.rand(n, dtype=torch.float64, device='cuda')
y = torch.empty_like(x)
for _ in range(3):
    K2 = F.asm_gpu(form, basis, out=K2.d_data)
    L.check(L.lib().mwx_spmv(n, Kc.spmv_hint(), L.ptr(Kc.d_indptr), L.ptr(Kc.d_indices), L.ptr(Kc.d_data), L.ptr(x), L.ptr(y), L.stream()))
torch.cuda.synchronize()
if maxit:
    b = Kc.dot(x)
    ml = F.AMGHierarchy(Kc, mode=amg_mode, reorder=False)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        _run_pcg(Kc, b, 1e-8, maxit, amg=ml, reordering=None)
torch.cuda.synchronize()
print('done', n, Kc.nnz)
