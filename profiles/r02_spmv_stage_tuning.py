"""Fine-level SpMV / smoother-step time for library variants built with other stage sizes and CTA residencies
(-DMWX_SPMV_STAGE_ROWS=R -DMWX_SPMV_MINB=B, see csrc/krylov.cu); one process per variant (MWX_LIB_PATH).
    python profiles/r02_spmv_stage_tuning.py [level]"""
import json, os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import fem_forth_perturb_b200 as F
from fem_forth_perturb_b200 import _lib as L
from fem_forth_perturb_b200.solvers import Reordering
from profiles.r02_sa_setup_probe_lib import system

level = int(sys.argv[1]) if len(sys.argv) > 1 else 12
Kc, bc = system(level)
A = Reordering.for_matrix(Kc).matrix(Kc)
n = A.shape[0]
x = torch.rand(n, dtype=torch.float64, device='cuda'); y = torch.empty_like(x); c = torch.rand_like(x)
dinv = torch.rand_like(x); d = torch.rand_like(x)
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device='cuda')
lib, st = L.lib(), L.stream()
plan = A.spmv_hint()
def timed(fn, reps=10):
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))
args = (n, plan, L.ptr(A.d_indptr), L.ptr(A.d_indices), L.ptr(A.d_data))
t_spmv = timed(lambda: L.check(lib.mwx_spmv(*args, L.ptr(x), L.ptr(y), st)))
t_step = timed(lambda: L.check(lib.mwx_smoother_step(*args, L.ptr(x), L.ptr(y), L.ptr(c), L.ptr(dinv), L.ptr(d), 0.3, 0.7, st)))
print(json.dumps(dict(lib=os.path.basename(L.LIB_PATH), level=level, plan=-plan, spmv_ms=t_spmv, smoother_step_ms=t_step,
                      spmv_gbs=(12 * A.nnz + 20 * n) / t_spmv / 1e6)), flush=True)
