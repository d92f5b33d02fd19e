"""Does a smoothed-aggregation hierarchy whose operator entries are rounded to fp32 still precondition the fp64 CG as
well?  (Probe before storing the hierarchy in fp32.)  The fine matrix handed to the SA setup is rounded to fp32 - every
level below is then built from perturbed data - and PCG runs on the exact fp64 matrix.
    python profiles/r02_fp32_hierarchy_probe.py [level]"""
import json, os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import fem_forth_perturb_b200 as F
from profiles.r02_sa_setup_probe_lib import system

level = int(sys.argv[1]) if len(sys.argv) > 1 else 11
Kc, bc = system(level)
Kr = F.DeviceCSR(Kc.d_indptr, Kc.d_indices, Kc.d_data.to(torch.float32).to(torch.float64), Kc.shape, coords=Kc.coords,
                 candidates=Kc.candidates)
for name, A in (('fp64 hierarchy', Kc), ('hierarchy from the fp32-rounded matrix', Kr)):
    ml = F.AMGHierarchy(A, mode='sa')
    F.pcg(Kc, bc, tol=1e-8, amg=ml, max_rechecks=3)
    torch.cuda.synchronize(); t = time.perf_counter()
    x, it, info, res = F.pcg(Kc, bc, tol=1e-8, amg=ml, max_rechecks=3)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(json.dumps(dict(level=level, hierarchy=name, iters=it, info=info, resid=res, pcg_s=dt)), flush=True)
    del ml
