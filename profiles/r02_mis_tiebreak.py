"""Aggregate-root tie-break of the device SA setup (MWX_SA_MIS_HASH_BITS): setup seconds and PCG iterations."""
import json, os, sys, time
sys.path.insert(0, os.getcwd())
import torch
import fem_forth_perturb_b200 as F
from profiles.r02_sa_setup_probe_lib import system
level = int(sys.argv[1]) if len(sys.argv) > 1 else 11
Kc, bc = system(level)
os.environ['MWX_SA_CYCLE'] = 'v'
torch.cuda.synchronize(); t = time.perf_counter()
ml = F.AMGHierarchy(Kc, mode='sa')
torch.cuda.synchronize(); ts = time.perf_counter() - t
x, it, info, res = F.pcg(Kc, bc, tol=1e-8, amg=ml, max_rechecks=3)
print(json.dumps(dict(level=level, hash_bits=os.environ.get('MWX_SA_MIS_HASH_BITS', '24'), setup_s=ts, levels=ml.level_sizes(), opc=ml.operator_complexity(), iters=it)))
