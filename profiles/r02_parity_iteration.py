"""Parity mode (RS hierarchy + exact symmetric Gauss-Seidel) per-iteration time on levels 5-10, beside the fast modes.
    python profiles/r02_parity_iteration.py [levels...]"""
import json, os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import fem_forth_perturb_b200 as F

def system(level, eps=1.0):
    gm = F.MeshTri().refine(level)
    basis = F.InteriorBasis(gm, F.ElementTriMorley())
    K = F.asm_gpu(eps ** 2 * F.a_load + F.b_load, basis)
    bf = gm.boundary_facets()
    D = np.concatenate((np.unique(gm.facets[:, bf]), bf + gm.nv))
    Kc = F.condense(K, D=D, expand=False)
    x = torch.from_numpy(np.random.default_rng(1234).uniform(-1, 1, Kc.shape[0])).cuda()
    return Kc, Kc.dot(x)

for level in [int(a) for a in sys.argv[1:]] or [5, 6, 7, 8, 9, 10]:
    Kc, b = system(level)
    for mode in ('parity', 'chebyshev'):
        ml = F.AMGHierarchy(Kc, mode=mode)
        F.pcg(Kc, b, 1e-8, 20, amg=ml)                       # warm
        torch.cuda.synchronize(); t = time.perf_counter()
        x, it, info, res = F.pcg(Kc, b, 1e-8, 60, amg=ml)
        torch.cuda.synchronize(); dt = time.perf_counter() - t
        print(json.dumps(dict(level=level, n=Kc.shape[0], mode=mode, levels=ml.level_sizes(), iters=it, ms_per_iter=1e3 * dt / max(it, 1))), flush=True)
        del ml
