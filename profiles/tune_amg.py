"""Smoother tuning on the level-L system: iterations, ms/iteration and time to solution per variant."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fem_forth_perturb_b200 as F
from fem_forth_perturb_b200.solvers import _run_pcg

level = int(sys.argv[1]) if len(sys.argv) > 1 else 11
gm = F.MeshTri().refine(level)
basis = F.InteriorBasis(gm, F.ElementTriMorley())
K2 = F.asm_gpu(1e-4 * F.a_load + F.b_load, basis)
bf = gm.boundary_facets()
D = np.concatenate((np.unique(gm.facets[:, bf]), bf + gm.nv))
Kc, _ = F.CondensePlan(K2, D=D).apply(K2)
n = Kc.shape[0]
g = torch.Generator(device='cuda').manual_seed(1234)
xs = torch.rand(n, dtype=torch.float64, device='cuda', generator=g) * 2 - 1
b = Kc.dot(xs)
hier = {}
for mode, deg, ratio in [('chebyshev', 3, 30.0), ('chebyshev', 2, 30.0), ('chebyshev', 2, 10.0), ('chebyshev', 3, 10.0),
                         ('chebyshev', 4, 30.0), ('chebyshev', 1, 4.0), ('jacobi', 0, 0)]:
    if mode not in hier:
        hier[mode] = F.AMGHierarchy(Kc, mode=mode)
    ml = hier[mode]
    if mode == 'chebyshev':
        ml.set_chebyshev(deg, ratio)
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        x, its, info, resid = _run_pcg(Kc, b, 1e-8, 3000, amg=ml)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f'level {level} n={n} {mode} degree={deg} ratio={ratio}: its={its} info={info} solve={dt:.3f}s ms/it={1e3*dt/max(its,1):.2f}', flush=True)
