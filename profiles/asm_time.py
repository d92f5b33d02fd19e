"""Assembly-only timing / ncu driver: K2 = eps^2 A + B on the level-L unit square with the tiled and the row-gather kernel.
    python profiles/asm_time.py [level] [kernels: tiles,rows] [reps]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import fem_forth_perturb_b200 as F

level = int(sys.argv[1]) if len(sys.argv) > 1 else 12
kernels = (sys.argv[2] if len(sys.argv) > 2 else 'tiles,rows').split(',')
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
gm = F.MeshTri().refine(level)
basis = F.InteriorBasis(gm, F.ElementTriMorley())
pat = basis.pattern()
form = 1e-4 * F.a_load + F.b_load
out = {'level': level, 'dofs': basis.N, 'nt': gm.nt, 'nnz': pat['nnz']}
torch.cuda.synchronize()
t0 = time.perf_counter()
name = F.assembly_kernel_name(basis)
torch.cuda.synchronize()
out['tile_plan_s'] = time.perf_counter() - t0
out['tile_plan'] = pat['tile_plan'].info if pat.get('tile_plan') is not None else None
vals = {}
for k in kernels:
    K = F.asm_gpu(form, basis, kernel=k)
    for _ in range(3):
        F.asm_gpu(form, basis, out=K.d_data, kernel=k)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        F.asm_gpu(form, basis, out=K.d_data, kernel=k)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    out[k] = {'ms': ms, 'gdof_s': basis.N / ms / 1e6, 'gbs_308B': 308 * gm.nt / ms / 1e6}
    vals[k] = K.d_data
if len(vals) == 2:
    a, b = vals['tiles'], vals['rows']
    out['max_rel_diff'] = float((a - b).abs().max() / b.abs().max())
print(json.dumps(out))
