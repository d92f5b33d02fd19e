"""RS (reference hierarchy, Chebyshev smoothing) vs smoothed aggregation on the bench problem: iterations, solve and setup time."""
import os, sys, time, torch, numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import fem_forth_perturb_b200 as F
lvl = int(sys.argv[1]) if len(sys.argv) > 1 else 11
modes = sys.argv[2].split(',') if len(sys.argv) > 2 else ['chebyshev', 'sa']
gm = F.MeshTri().refine(lvl)
basis = F.InteriorBasis(gm, F.ElementTriMorley())
K = F.asm_gpu(1e-4 * F.a_load + F.b_load, basis)
bf = gm.boundary_facets()
D = np.concatenate((np.unique(gm.facets[:, bf]), bf + gm.nv))
Kc = F.condense(K, D=D, expand=False)
n = Kc.shape[0]
g = torch.Generator(device='cuda').manual_seed(1234)
xs = torch.rand(n, dtype=torch.float64, device='cuda', generator=g) * 2 - 1
b = Kc.dot(xs)
for mode in modes:
    t0 = time.perf_counter(); ml = F.AMGHierarchy(Kc, mode=mode); torch.cuda.synchronize(); t_setup = time.perf_counter() - t0
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        x, its, info, resid = F.pcg(Kc, b, 1e-8, 2000, amg=ml)
        torch.cuda.synchronize(); t_solve = time.perf_counter() - t0
    err = float(torch.linalg.norm(x - xs) / torch.linalg.norm(xs))
    print(f'level {lvl} n {n} mode {mode}: its {its} info {info} solve {t_solve:.3f} s ({t_solve / its * 1e3:.2f} ms/it) setup {t_setup:.1f} s '
          f'(host {ml.setup_host_seconds:.1f}, upload {ml.upload_seconds:.1f}) levels {ml.level_sizes()} cx {ml.operator_complexity():.2f} err {err:.2e}', flush=True)
    del ml; torch.cuda.empty_cache()
