"""AMG-PCG iteration counts of the CPU oracle (pyamg RS + symmetric GS + scipy-1.4.1 cg restated) on the rows of the
reference's table that need hundreds of iterations - minutes of CPU, so the counts are recorded here instead of
being recomputed in the GPU test:   python tests/golden/make_oracle_counts.py
    (eps, level): oracle count  [reference log]
    (1.0, 7): 249 [241]   (0.1, 7): 186 [183]   (0.1, 8): 475 [484]   (0.01, 8): 63 [62]
"""
import sys; import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.skfem_mesh import MeshTri
from oracle import pipeline as pl
from oracle.scipy_cg import solver_iter_mgcg
import time
for eps, lev in ((1.0,7),(0.1,7),(0.1,8),(0.01,8)):
    t=time.time(); ru=[]
    m=MeshTri().refine(lev)
    pl.solve_problem(m, pl.Ex1(eps), wkind='P1', intorder=5, solver_u=solver_iter_mgcg(tol=1e-8, maxiter=1000, record=ru))
    print(eps, lev, ru, time.time()-t, flush=True)
