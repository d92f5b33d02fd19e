"""gamma = 2 cycle (levels 1..k visited twice) against the V-cycle, mode 'sa': iterations and seconds per solve."""
import json, os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import fem_forth_perturb_b200 as F
from profiles.r02_sa_setup_probe_lib import system

level = int(sys.argv[1]) if len(sys.argv) > 1 else 11
Kc, bc = system(level)
ml = F.AMGHierarchy(Kc, mode='sa')
print(json.dumps(dict(levels=ml.level_sizes(), max_coarse=os.environ.get('MWX_SA_MAX_COARSE'), setup_s=ml.setup_device_seconds)), flush=True)
for gamma, k0, k in ((1, 1, 0), (2, 2, 2), (2, 2, 3), (2, 2, 4), (2, 2, 9), (2, 3, 4), (2, 3, 3)):
    ml.set_cycle(gamma, k, first=k0)
    F.pcg(Kc, bc, tol=1e-8, amg=ml, max_rechecks=3)
    torch.cuda.synchronize(); t = time.perf_counter()
    x, it, info, res = F.pcg(Kc, bc, tol=1e-8, amg=ml, max_rechecks=3)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(json.dumps(dict(level=level, gamma=gamma, revisited_levels=[k0, k], iters=it, info=info, pcg_s=dt, ms_per_iter=1e3 * dt / it)), flush=True)
