"""Smoothed-aggregation tuning on the level-L system: strength threshold x Chebyshev degree/ratio."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fem_forth_perturb_b200 as F
from fem_forth_perturb_b200.solvers import _run_pcg

level = int(sys.argv[1]) if len(sys.argv) > 1 else 11
thetas = [float(t) for t in sys.argv[2].split(',')] if len(sys.argv) > 2 else [0.1]
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
for theta in thetas:
    t0 = time.perf_counter(); ml = F.AMGHierarchy(Kc, mode='sa', theta=theta); ts = time.perf_counter() - t0
    print(f'theta {theta}: levels {ml.level_sizes()} cx {ml.operator_complexity():.2f} setup {ts:.1f} s', flush=True)
    for deg, ratio in [(2, 10.0), (1, 4.0), (2, 4.0), (2, 30.0), (3, 10.0), (3, 30.0)]:
        ml.set_chebyshev(deg, ratio)
        for rep in range(2):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            x, its, info, resid = _run_pcg(Kc, b, 1e-8, 3000, amg=ml)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f'  level {level} sa theta={theta} degree={deg} ratio={ratio}: its={its} info={info} solve={dt:.3f}s ms/it={1e3*dt/max(its,1):.2f}', flush=True)
    del ml; torch.cuda.empty_cache()
