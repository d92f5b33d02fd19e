"""Device vs host smoothed-aggregation setup on the example-1 system: levels, iterations, seconds."""
import json, os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import fem_forth_perturb_b200 as F

from profiles.r02_sa_setup_probe_lib import system

for level in [int(a) for a in sys.argv[1:]] or [6, 9]:
    Kc, bc = system(level)
    for setup in (['device', 'host'] if level <= 11 else ['device']):
        torch.cuda.synchronize(); t = time.perf_counter()
        ml = F.AMGHierarchy(Kc, mode='sa', setup=setup)
        torch.cuda.synchronize(); ts = time.perf_counter() - t
        t = time.perf_counter()
        x, it, info, res = F.pcg(Kc, bc, tol=1e-8, amg=ml, max_rechecks=3)
        torch.cuda.synchronize(); tp = time.perf_counter() - t
        print(json.dumps(dict(level=level, n=Kc.shape[0], setup=setup, setup_s=ts, reorder_s=ml.reorder_seconds, levels=ml.level_sizes(),
                              opc=ml.operator_complexity(), iters=it, info=info, pcg_s=tp)), flush=True)
        del ml
        torch.cuda.empty_cache()
