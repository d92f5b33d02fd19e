"""Single-GPU PCG iterations in mode 'sa' for an ncu launch list (--profile-from-start off):
    python profiles/r02_sa_iteration.py [level] [maxiter]"""
import json, os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import fem_forth_perturb_b200 as F
from profiles.r02_sa_setup_probe_lib import system

level = int(sys.argv[1]) if len(sys.argv) > 1 else 12
maxiter = int(sys.argv[2]) if len(sys.argv) > 2 else 6
Kc, bc = system(level)
ml = F.AMGHierarchy(Kc, mode='sa')
F.pcg(Kc, bc, tol=1e-30, maxiter=3, amg=ml)
torch.cuda.synchronize(); t = time.perf_counter()
torch.cuda.profiler.start()
x, it, info, res = F.pcg(Kc, bc, tol=1e-30, maxiter=maxiter, amg=ml)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(json.dumps(dict(level=level, levels=ml.level_sizes(), storage_bits=ml.storage_bits, iters=it,
                      ms_per_iter=1e3 * (time.perf_counter() - t) / it)), flush=True)
