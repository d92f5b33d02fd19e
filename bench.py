#!/usr/bin/env python
"""bench.py - the reference's hot path (MWX assembly -> condense -> SpMV -> AMG-PCG) on B200.

One "step" = one pass of the hot path over the level-L unit-square mesh (default L = 12: 67 125 249
Morley DOFs, nnz(K2) = 771 817 473 - BASELINE.json config "strong/weak scaling: >= 50M-DOF unit-square
mesh"; config "assembly-only microbench" is the same mesh family, see --level):
    assemble K2 = eps^2 A + B      (mwx_assemble_morley, one launch)
    condense the Dirichlet DOFs    (mwx_condense_*)
    one SpMV y = K2c x             (mwx_spmv; the roofline kernel)
    one AMG-PCG solve K2c u = b    (mwx_pcg_amg, Chebyshev-smoothed V-cycle, tol 1e-8)
Mesh refinement, CSR pattern / slot map and the AMG hierarchy are SETUP: built once, timed and
reported separately (north_star: "hierarchy built once on the host as timed setup").
N > 1 (torchrun): STRONG scaling on the same level-L mesh - RCB row blocks, every rank assembles its own rows,
partitioned PCG with the global AMG hierarchy applied as a collective V-cycle (run_partitioned).
--microbench: the assembly-only 2k/4k/8k microbench (levels 11-13).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--level L] [--impl reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...      (N > 1)
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

EPS = 1e-2            # eps^2 = 1e-4 (SURVEY 8d microbench setting)
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture of the SAME
# kernels on the level-12 workload (profiles/r01_ncu_final_spmv_assembly.txt); None for other levels.
NCU_TRAFFIC_BYTES = {12: {'k_spmv': 10.374417e9 + 0.536002e9, 'k_assemble_morley': 12.228423e9 + 6.146362e9,
                          'k_assemble_morley_renumbered': 7.890940e9 + 6.137720e9}}
TOL = 1e-8            # the reference's tol (ref main/example1/run.py:20)


def measured_peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, gpu_index=0):
        self.samples, self.reasons, self.proc = [], set(), None
        self.gpu_index = gpu_index

    def start(self):
        if os.environ.get('MWX_BENCH_NO_SAMPLER'):          # A/B switch: is a stall in the timed region the sampler's doing?
            return
        q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu_index), f'--query-gpu={q}',
                                          '--format=csv,noheader,nounits', '-lms', '500'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def wait_first(self, timeout=8.0):
        """Block until the first sample has arrived (or the sampler could not start): the sampling then runs DURING the
        timed region, but nvidia-smi's start-up - seconds of NVML initialisation that can stall CUDA calls - does not."""
        t0 = time.perf_counter()
        while self.proc is not None and not self.samples and time.perf_counter() - t0 < timeout:
            time.sleep(0.05)

    def _read(self):
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(',')]
            try:
                self.samples.append((float(parts[0]), float(parts[1])))
                for nm, v in zip(names, parts[2:6]):
                    if v.lower().startswith('active'):
                        self.reasons.add(nm)
            except Exception:
                pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': sorted(self.reasons), 'samples': 0}
        sm = sorted(s[0] for s in self.samples)
        return {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': max(s[1] for s in self.samples),
                'reasons': sorted(self.reasons), 'samples': len(sm)}


# ------------------------------------------------------------------------------------------ reference arm
def cpu_reference_pass(level, eps=EPS, tol=TOL):
    """The reference's own CPU path, restated (oracle port: skfem asm + condense + pyamg RS + scipy-1.4.1 cg)
    on a bounded level.  Returns phase seconds.  Only this function and the parity tests touch oracle/."""
    from oracle.skfem_mesh import MeshTri as OMesh
    from oracle.skfem_basis import InteriorBasis as OInterior
    from oracle import skfem_asm as oa
    from oracle.pyamg_rs import RugeStuben
    from oracle.scipy_cg import cg141
    m = OMesh().refine(level)
    t0 = time.perf_counter()
    ub = OInterior(m, 'Morley', 5)                                   # basis tabulation (part of skfem assembly)
    K2 = eps ** 2 * oa.asm_bilinear(oa.a_load, ub) + oa.asm_bilinear(oa.b_load, ub)
    t1 = time.perf_counter()
    rng = np.random.default_rng(1234)
    D = oa.easy_boundary(ub)
    Kc, _, _, I = oa.condense(K2, np.zeros(K2.shape[0]), D=D)
    xs = rng.uniform(-1, 1, Kc.shape[0])
    b = Kc @ xs
    t2 = time.perf_counter()
    y = Kc @ xs
    t3 = time.perf_counter()
    ml = RugeStuben(Kc)
    t4 = time.perf_counter()
    x, info, its = cg141(Kc, b, M=ml.precondition, tol=tol, maxiter=1000)
    t5 = time.perf_counter()
    return dict(ndofs=K2.shape[0], n=Kc.shape[0], nnz=Kc.nnz, asm_s=t1 - t0, condense_s=t2 - t1, spmv_s=t3 - t2,
                amg_setup_s=t4 - t3, pcg_s=t5 - t4, iters=its, info=info)


def run_reference(args):
    """`--impl reference`: the CPU restatement of the reference path on the host cores (rank 0 only)."""
    if int(os.environ.get('RANK', '0')) != 0:
        return
    level = args.ref_level
    res = []
    for i in range(args.warmup + args.steps):
        r = cpu_reference_pass(level)
        if i >= args.warmup:
            res.append(r)
    step_s = float(np.mean([r['asm_s'] + r['condense_s'] + r['spmv_s'] + r['pcg_s'] for r in res]))
    step_setup_s = step_s + float(np.mean([r['amg_setup_s'] for r in res]))
    r0 = res[0]
    value = r0['ndofs'] / step_s / 1e6
    sample = (f"level {level} of the same mesh family ({r0['ndofs']} DOFs): asm(a_load)+asm(b_load) incl. basis "
              f"tabulation, condense, 1 SpMV, RS-AMG(sym. GS)-PCG tol {TOL} ({r0['iters']} its); AMG setup "
              f"({np.mean([r['amg_setup_s'] for r in res]):.2f} s) reported separately like the GPU arm")
    line = {
        'impl': 'reference', 'metric': 'hot_path_MDOF_per_s', 'value': value, 'unit': 'MDOF/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': step_s * 1e3,
        'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': f'MWX assemble+condense+SpMV+AMG-PCG, unit square level {level} (bounded CPU sample '
                               f'of the level-{args.level} workload), eps={EPS}, tol={TOL}'},
        'assembly_mdof_s': r0['ndofs'] / float(np.mean([r['asm_s'] for r in res])) / 1e6,
        'pcg': {'s_per_solve': float(np.mean([r['pcg_s'] for r in res])), 'iters': r0['iters'],
                'setup_s': float(np.mean([r['amg_setup_s'] for r in res])), 'mode': 'parity (symmetric GS)'},
        'spmv_gbs': (12 * r0['nnz'] + 20 * r0['n']) / float(np.mean([r['spmv_s'] for r in res])) / 1e9,
        'value_incl_amg_setup': r0['ndofs'] / step_setup_s / 1e6,
        'cpu_baseline': {'value': value, 'unit': 'MDOF/s', 'cores': 1, 'kind': 'port', 'sample': sample,
                         'host_cores_available': os.cpu_count()},
        'e2e': {'value': value, 'unit': 'MDOF/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--level', type=int, default=12, help='refinements of the unit square (12 -> 67.1 M DOFs)')
    ap.add_argument('--ref-level', type=int, default=8, help='bounded level for the CPU reference arm')
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--mode', default='sa', choices=['sa', 'chebyshev', 'jacobi', 'parity'],
                    help="AMG: 'sa' = smoothed aggregation on the Morley interpolants of 1, x, y + Chebyshev (throughput mode); "
                         "'chebyshev' / 'jacobi' = the reference's Ruge-Stueben hierarchy with device smoothers; 'parity' = RS + "
                         "symmetric Gauss-Seidel (pyamg's defaults, iteration-count parity)")
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-rs-compare', action='store_true',
                    help="skip the one untimed solve on the reference's Ruge-Stueben hierarchy that mode 'sa' reports beside itself")
    ap.add_argument('--no-ex1', action='store_true', help='skip the real example-1 problem (error norms vs the exact solution)')
    ap.add_argument('--no-same-config', action='store_true', help="skip the level-8 pass in the reference's own mode")
    ap.add_argument('--ex1-rs-top', action='store_true',
                    help="also run the reference's RS hierarchy on the real ex1 right-hand side at the top level (thousands of iterations)")
    ap.add_argument('--asm-levels', type=int, nargs='*', default=[11, 13],
                    help='extra levels of the assembly-only microbench reported in the line (the main level is timed in the step)')
    ap.add_argument('--microbench', action='store_true',
                    help='assembly-only microbench on the 2k/4k/8k meshes (levels 11, 12, 13), 1 GPU')
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == 'reference':
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    import fem_forth_perturb_b200 as F
    from fem_forth_perturb_b200 import _lib as L
    dev = torch.device('cuda', local)
    if world > 1:
        run_partitioned(args, torch, dist, F, L, dev, rank, world, local)
        dist.destroy_process_group()
        return
    if args.microbench:
        run_microbench(args, torch, F)
        return

    run_single(args, torch, F, L, dev, local)


def vcycle_launches(nlev, revisit=()):
    """Kernel launches of one AMG cycle on `nlev` levels: 7 per visit of a smoothed level (first smoother step, one
    smoother SpMV, residual, restriction, prolongation, two smoother SpMVs), one dense GEMV per visit of the coarsest
    level, and residual + accumulate for every second visit of a level in `revisit` (gamma = 2 there).  Matches the ncu
    launch list (profiles/r02_launches_pcg_level12_sa.txt: 134 + the 4 CG kernels)."""
    visits, total = 1, 0
    for l in range(nlev - 1):
        if l in revisit:
            total += 2 * visits           # per visit of the parent: residual for the second visit + accumulate
            visits *= 2
        total += 7 * visits
    return total + visits


def kernel_traffic(kernel, level):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture
    (profiles/ncu_traffic.json, written by profiles/ncu_summary.py).  An entry is keyed to the sha256 of the kernel's
    source file at capture time: if the source has changed since, the number is stale and None is reported."""
    import hashlib
    try:
        with open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')) as f:
            ent = json.load(f).get(f'{kernel}@level{level}')
        if not ent:
            return None
        with open(os.path.join(ROOT, ent['source']), 'rb') as f:
            sha = hashlib.sha256(f.read()).hexdigest()
        return ent['bytes'] if sha == ent['source_sha256'] else None
    except Exception:
        return None


def ex1_pipeline(torch, F, gm, ubasis, plan, K2, eps, solve_u, w_mode, tol, refine_u=None):
    """The reference's solve_problem1 (ref main/example1/run.py:182-222, w in P1 as in the golden 8-level log) with
    every array on the device: K1 = asm(laplace), f1 = asm(f_load) -> w_h; f2 = asm(wv_load) * w_h; condense; u_h.
    Loads and the exact solution are the 'ex1' data of problems.py evaluated by the kernels.
    solve_u: {name: callable(Kc, bc) -> (x, iters, info)}.  Returns per-solver error tables and iteration counts."""
    prob = F.Problem('ex1', eps)
    wb = F.InteriorBasis(gm, F.ElementTriP1(), intorder=5)
    K1 = F.asm_gpu(F.laplace, wb)
    f1 = F.asm_gpu(prob.f_load, wb, on_device=True)
    planw = F.CondensePlan(K1, D=np.unique(wb.find_dofs()['all'].all()))
    K1c, f1c = planw.apply(K1, f1)
    t0 = time.perf_counter()
    mlw = F.AMGHierarchy(K1c, mode=w_mode)
    t_wsetup = time.perf_counter() - t0
    xw, its_w, info_w, _ = F.pcg(K1c, f1c, tol, 2000, amg=mlw)
    del mlw
    wh = torch.zeros(wb.N, dtype=torch.float64, device=gm.device)
    wh[torch.from_numpy(planw.I).to(gm.device)] = xw
    f2 = F.asm_gpu(F.wv_load, wb, ubasis).dot(wh)
    Kc, bc = plan.apply(K2, f2)
    I_dev = torch.from_numpy(np.ascontiguousarray(plan.I, dtype=np.int64)).to(gm.device)
    out = {'w_solve': {'iters': its_w, 'info': info_w, 'mode': w_mode, 'n': K1c.shape[0], 'amg_setup_s': t_wsetup}}
    sols = {}
    bn = float(torch.linalg.norm(bc))
    for name, fn in solve_u.items():
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        x, its, info, first_hit = fn(Kc, bc)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        uh = torch.zeros(ubasis.N, dtype=torch.float64, device=gm.device)
        uh[I_dev] = x
        L2, D1, D2 = F.morley_error_norms(ubasis, uh, problem=prob)
        out[name] = {'iters': its, 'info': info, 'first_iter_with_recurrence_residual_below_tol': first_hit,
                     'true_rel_residual': float(torch.linalg.norm(bc - Kc.dot(x))) / bn, 's_per_solve': dt,
                     'L2': L2, 'H1': L2 + D1, 'H2': L2 + D1 + D2, 'energy': float(np.sqrt(eps ** 2 * D2 ** 2 + D1 ** 2))}
        sols[name] = x
        if refine_u is not None and its < 500:
            # mixed-precision iterative refinement (residual in double-double, solvers.refine): the forward error leaves the
            # cond(A) * eps floor of the plain fp64 solve, so the discretisation error - and its order - shows on the finest levels
            log = []
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            xr = refine_u(name, Kc, bc, x, log)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            uh.zero_()
            uh[I_dev] = xr
            L2, D1, D2 = F.morley_error_norms(ubasis, uh, problem=prob)
            out[name]['refined'] = {'steps': log, 's': dt, 'L2': L2, 'H1': L2 + D1, 'H2': L2 + D1 + D2,
                                    'energy': float(np.sqrt(eps ** 2 * D2 ** 2 + D1 ** 2)),
                                    'rel_change_of_solution': float(torch.linalg.norm(xr - x) / torch.linalg.norm(xr))}
            sols[name + '_refined'] = xr
    return out, sols


def ex1_level(torch, F, level, eps, tol, modes, maxiter, setup=None):
    """ex1 on one level with every hierarchy in `modes`; setup = (gm, basis, plan, K2, {mode: hierarchy}) to reuse the
    bench's own objects.  Stops after 2 failed true-residual re-checks (mwx_pcg_amg_monitored): on the finest levels
    tol * ||b|| lies below the fp64 floor eps_mach ||A|| ||x|| of the residual and scipy's rule can never be met."""
    if setup is None:
        gm = F.MeshTri().refine(level)
        ub = F.InteriorBasis(gm, F.ElementTriMorley(), intorder=5)
        bf = gm.boundary_facets()
        D = np.concatenate((np.unique(gm.facets[:, bf]), bf + gm.nv))
        K2 = F.asm_gpu(eps ** 2 * F.a_load + F.b_load, ub)
        plan = F.CondensePlan(K2, D=D)
        Kc, _ = plan.apply(K2)
        mls = {mode: F.AMGHierarchy(Kc, mode=mode) for mode in modes}
        del Kc
    else:
        gm, ub, plan, K2, mls = setup

    def solver(mode):
        def fn(Kc_, bc_):
            mon = {}
            x, its, info, _ = F.pcg(Kc_, bc_, tol, maxiter[mode], amg=mls[mode], max_rechecks=2, monitor=mon)
            return x, its, info, mon.get('first_hit')
        return fn
    def refine_u(mode, Kc_, bc_, x, log):
        return F.refine(Kc_, bc_, x, steps=2, tol=tol, maxiter=maxiter[mode], amg=mls[mode], log=log)
    res, sols = ex1_pipeline(torch, F, gm, ub, plan, K2, eps, {m: solver(m) for m in modes if m in mls},
                             modes[0] if modes[0] in ('sa', 'chebyshev') else 'chebyshev', tol, refine_u=refine_u)
    res['dofs'] = ub.N
    names = [m for m in modes if m in sols]
    if len(names) == 2:
        a, b2 = sols[names[0]], sols[names[1]]
        res[f'rel_l2_diff_{names[0]}_vs_{names[1]}'] = float(torch.linalg.norm(a - b2) / torch.linalg.norm(b2))
        ra, rb = sols.get(names[0] + '_refined'), sols.get(names[1] + '_refined')
        if ra is not None and rb is not None:
            res[f'rel_l2_diff_refined_{names[0]}_vs_{names[1]}'] = float(torch.linalg.norm(ra - rb) / torch.linalg.norm(rb))
    return res


def same_config_level8(torch, F, L, level, tol):
    """ONE pass timed exactly like cpu_reference_pass (same level, same phases, the reference's own preconditioner:
    Ruge-Stueben hierarchy + symmetric Gauss-Seidel = mode 'parity', maxiter 1000, rhs = K2c @ U(-1,1) seed 1234)
    on the device, so that the record holds one same-config GPU-vs-CPU pair."""
    dev = torch.device('cuda', torch.cuda.current_device())
    gm = F.MeshTri().refine(level)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ub = F.InteriorBasis(gm, F.ElementTriMorley(), intorder=5)
    K2 = F.asm_gpu(EPS ** 2 * F.a_load + F.b_load, ub)             # includes pattern + slot map (skfem: basis tabulation)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    bf = gm.boundary_facets()
    D = np.concatenate((np.unique(gm.facets[:, bf]), bf + gm.nv))
    Kc = F.condense(K2, D=D, expand=False)
    xs = torch.from_numpy(np.random.default_rng(1234).uniform(-1, 1, Kc.shape[0])).to(dev)
    b = Kc.dot(xs)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    y = Kc.dot(xs)
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    ml = F.AMGHierarchy(Kc, mode='parity')
    torch.cuda.synchronize()
    t4 = time.perf_counter()
    x, its, info, resid = F.pcg(Kc, b, tol, 1000, amg=ml)
    torch.cuda.synchronize()
    t5 = time.perf_counter()
    step = (t1 - t0) + (t2 - t1) + (t3 - t2) + (t5 - t4)
    return dict(level=level, ndofs=ub.N, n=Kc.shape[0], mode='parity (RS hierarchy + symmetric Gauss-Seidel, the reference\'s own)',
                asm_s=t1 - t0, condense_s=t2 - t1, spmv_s=t3 - t2, amg_setup_s=t4 - t3, pcg_s=t5 - t4, iters=its, info=info,
                value_mdof_s=ub.N / step / 1e6, value_incl_amg_setup_mdof_s=ub.N / (step + t4 - t3) / 1e6,
                pcg_ms_per_iter=(t5 - t4) / max(its, 1) * 1e3,
                rel_err_vs_manufactured=float(torch.linalg.norm(x - xs) / torch.linalg.norm(xs)),
                timing='host wall clock around synchronised phases, one cold pass - exactly as the CPU arm is timed')


def run_single(args, torch, F, L, dev, local):
    rank = 0
    ev = lambda: torch.cuda.Event(enable_timing=True)
    form = EPS ** 2 * F.a_load + F.b_load
    spec_mode = 'chebyshev'                          # north_star (3): pyamg hierarchy built on the host, Chebyshev smoothing
    modes = [args.mode] + ([spec_mode] if args.mode != spec_mode and not args.no_rs_compare else [])

    # ---------------------------------------------------------------- setup (timed, not part of a step)
    t0 = time.perf_counter()
    gm = F.MeshTri().refine(args.level)
    torch.cuda.synchronize()
    t_mesh = time.perf_counter() - t0
    basis = F.InteriorBasis(gm, F.ElementTriMorley(), intorder=5)
    t0 = time.perf_counter()
    pat = basis.pattern()
    torch.cuda.synchronize()
    t_pattern = time.perf_counter() - t0
    N, nt, nnz = basis.N, gm.nt, pat['nnz']
    t0 = time.perf_counter()
    asm_kernel = F.assembly_kernel_name(basis)     # builds the tile plan of the element-once kernel (setup)
    torch.cuda.synchronize()
    t_tileplan = time.perf_counter() - t0
    # Dirichlet set of easy_boundary (ref main/example1/run.py:86-104): boundary vertices + boundary facets
    bf = gm.boundary_facets()
    D = np.concatenate((np.unique(gm.facets[:, bf]), bf + gm.nv))
    K2 = F.asm_gpu(form, basis)
    t0 = time.perf_counter()
    plan = F.CondensePlan(K2, D=D)                 # kept-row numbering + condensed indptr: pattern-only, once
    torch.cuda.synchronize()
    t_cplan = time.perf_counter() - t0
    Kc, _ = plan.apply(K2)
    n, nnzc = Kc.shape[0], Kc.nnz
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    xstar = torch.rand(n, dtype=torch.float64, device=dev, generator=g) * 2 - 1     # manufactured solution
    b = Kc.dot(xstar)
    mls, t_amg = {}, {}

    def build(mode):                         # one hierarchy at a time lives on the device (40 GB for 'sa' at level 12)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        mls[mode] = F.AMGHierarchy(Kc, mode=mode)
        torch.cuda.synchronize()
        t_amg[mode] = time.perf_counter() - t0
    build(args.mode)
    # pinned host inputs / outputs for the end-to-end leg
    p_host = torch.empty(2 * gm.nv, dtype=torch.float64).pin_memory()
    t_host = torch.empty(3 * gm.nt, dtype=torch.int32).pin_memory()
    b_host = torch.empty(n, dtype=torch.float64).pin_memory()
    x_host = torch.empty(n, dtype=torch.float64).pin_memory()
    p_host.copy_(gm.d_pxy); t_host.copy_(gm.d_t); b_host.copy_(b)
    torch.cuda.synchronize()
    vals = torch.empty(nnz, dtype=torch.float64, device=dev)
    vals_c = torch.empty(nnzc, dtype=torch.float64, device=dev)
    ysp = torch.empty(n, dtype=torch.float64, device=dev)
    perm_bufs = (torch.empty(n + 1, dtype=torch.int64, device=dev), torch.empty(nnzc, dtype=torch.int32, device=dev),
                 torch.empty(nnzc, dtype=torch.float64, device=dev))           # solver-side copy P K2c P^T, refreshed every step
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)     # 256 MiB > 126 MB L2
    hints = {}

    # Events are created (and recorded once) up front and reused: nothing that takes the driver's context lock
    # (cudaEventCreate / Destroy, cudaMalloc of a fresh output block) is left at the step boundaries, where the
    # nvidia-smi sampler polling the same driver was seen to stall such calls for 0.3-0.5 s on some boxes.
    step_events, region_events = [ev() for _ in range(8)], [ev(), ev()]      # a step synchronises and reads its events before it returns
    for e in step_events + region_events:
        e.record()
    torch.cuda.synchronize()

    def step(e2e, mode):
        """One pass of the hot path.  Returns per-phase device milliseconds and the iteration count."""
        ml = mls[mode]
        t_step = time.perf_counter()
        marks = step_events
        if e2e:
            gm.d_pxy.copy_(p_host, non_blocking=True)
            gm.d_t.copy_(t_host, non_blocking=True)
            bd = b_host.to(dev, non_blocking=True)
        else:
            bd = b
        host_t = [time.perf_counter()]
        marks[0].record()
        K = F.asm_gpu(form, basis, out=vals)
        marks[1].record()
        host_t.append(time.perf_counter())
        Kcs, _ = plan.apply(K, values_out=vals_c)
        marks[2].record()
        host_t.append(time.perf_counter())
        if 'kc' not in hints:                           # tile plans depend on the pattern only: once, outside the marks
            hints['kc'] = Kcs.spmv_hint()
        flush.zero_()                                   # evict the operator from L2 before the SpMV sample
        marks[3].record()
        L.check(L.lib().mwx_spmv(n, hints['kc'], L.ptr(Kcs.d_indptr), L.ptr(Kcs.d_indices), L.ptr(Kcs.d_data), L.ptr(xstar),
                                 L.ptr(ysp), L.stream()), 'mwx_spmv')
        marks[4].record()
        Ks = ml.reordering.matrix(Kcs, out=perm_bufs) if ml.reordering is not None else Kcs
        if 'ks' not in hints:
            hints['ks'] = Ks.spmv_hint()
        flush.zero_()
        marks[5].record()
        L.check(L.lib().mwx_spmv(n, hints['ks'], L.ptr(Ks.d_indptr), L.ptr(Ks.d_indices), L.ptr(Ks.d_data), L.ptr(xstar),
                                 L.ptr(ysp), L.stream()), 'mwx_spmv')
        marks[6].record()
        host_t.append(time.perf_counter())
        x, its, info, resid = F.pcg(Kcs, bd, TOL, 2000, amg=ml)
        marks[7].record()
        host_t.append(time.perf_counter())
        if e2e:
            x_host.copy_(x, non_blocking=True)
        torch.cuda.synchronize()
        host_t.append(time.perf_counter())
        ms = [marks[i].elapsed_time(marks[i + 1]) for i in range(7)]
        return dict(asm=ms[0], condense=ms[1], spmv_skfem=ms[3], reorder=ms[4], spmv=ms[5], pcg=ms[6], its=its,
                    info=info, resid=resid, x=x, wall_ms=(time.perf_counter() - t_step) * 1e3,
                    host_ms=[round((host_t[0] - t_step) * 1e3, 1)] + [round((b_ - a_) * 1e3, 1) for a_, b_ in zip(host_t[:-1], host_t[1:])])

    def timed(e2e, mode, steps, warmup):
        for _ in range(warmup):
            step(e2e, mode)
        torch.cuda.synchronize()
        e0, e1 = region_events
        e0.record()
        rs = []
        for i in range(steps):
            r = step(e2e, mode)
            if i < steps - 1:
                r['x'] = None                      # only the last iterate is looked at: its block is reused, no new cudaMalloc per step
            rs.append(r)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), rs

    def mean(key, rr):
        return float(np.mean([r[key] for r in rr]))

    sampler = ClockSampler(local)
    sampler.start()
    sampler.wait_first()          # nvidia-smi's start-up (NVML initialisation) must not land inside the timed steps
    total_ms, rs = timed(False, args.mode, args.steps, args.warmup)
    clocks = sampler.stop()
    e2e_ms, rs_e2e = timed(True, args.mode, args.steps, 1)
    ms_step = total_ms / args.steps
    value = N / (ms_step * 1e-3) / 1e6
    xerr = float(torch.linalg.norm(rs[-1]['x'] - xstar) / torch.linalg.norm(xstar))

    def mode_block(mode, ms_step_, e2e_ms_step, rr):
        ml = mls[mode]
        its = rr[-1]['its']
        setup = t_amg[mode]
        return {'hierarchy': ('smoothed aggregation on the Morley interpolants of 1, x, y (no reference counterpart)' if mode == 'sa'
                              else "pyamg ruge_stuben_solver defaults (the reference's hierarchy; C++ host restatement)"),
                'smoother': 'symmetric Gauss-Seidel (pyamg default)' if mode == 'parity' else ('2x damped Jacobi' if mode == 'jacobi' else 'Chebyshev(2)'),
                'value_mdof_s': N / (ms_step_ * 1e-3) / 1e6, 'ms_per_step': ms_step_,
                'e2e_mdof_s': None if e2e_ms_step is None else N / (e2e_ms_step * 1e-3) / 1e6,
                'value_incl_amg_setup_mdof_s': N / (ms_step_ * 1e-3 + setup) / 1e6,
                'pcg_iters': its, 'pcg_info': rr[-1]['info'], 'pcg_s_per_solve': mean('pcg', rr) * 1e-3,
                'pcg_ms_per_iter': mean('pcg', rr) / max(its, 1),
                'step_wall_ms': [round(r['wall_ms'], 1) for r in rr], 'step_pcg_ms': [round(r['pcg'], 1) for r in rr],
                'step_host_ms_[pre,asm,condense,spmv+reorder,pcg,sync]': [r['host_ms'] for r in rr],
                'e2e_step_wall_ms': [round(r['wall_ms'], 1) for r in rs_e2e] if mode == args.mode else None,
                'amg_setup_s': setup, 'amg_setup_where': ml.setup, 'amg_host_setup_s': ml.setup_host_seconds,
                'amg_device_setup_s': ml.setup_device_seconds, 'amg_upload_s': ml.upload_seconds,
                'amg_reorder_s': ml.reorder_seconds, 'levels': ml.level_sizes(),
                'operator_complexity': ml.operator_complexity(),
                'hierarchy_storage': ('operator values in fp32 (8-byte packed CSR records / fp32 block-CSR planes), vectors, products and sums '
                                      'in fp64; the system matrix CG multiplies with stays fp64' if getattr(ml, 'storage_bits', 64) == 32 else 'fp64'),
                'rel_err_vs_manufactured': float(torch.linalg.norm(rr[-1]['x'] - xstar) / torch.linalg.norm(xstar))}

    by_mode = {args.mode: mode_block(args.mode, ms_step, e2e_ms / args.steps, rs)}
    x_by_mode = {args.mode: rs[-1]['x']}

    # ---------------------------------------------------------------- the real example-1 problem on this level and the two below
    ex1 = None
    ex1_maxiter = {m: (200 if m == 'sa' else 3000) for m in modes}
    if not args.no_ex1:
        ex1 = {'what': "ref main/example1/run.py solve_problem1 (w in P1, eps = %g, tol = %g) on the device; errors vs the exact "
                       "solution u = (sin pi x sin pi y)^2; H1 = L2 + |.|_1h, H2 = H1 + |.|_2h, energy = sqrt(eps^2 |.|_2h^2 + |.|_1h^2) "
                       "as the reference prints them (run.py:519-543).  A solve stops on scipy's rule or after 2 failed true-residual "
                       "re-checks (fp64 floor of the residual, see true_rel_residual); the reference's RS hierarchy is only run "
                       "where its iteration count stays bounded (levels below the top one)" % (EPS, TOL), 'levels': {}}
        try:
            top_modes = [m for m in modes if m == 'sa' or args.ex1_rs_top] or modes[:1]
            ex1['levels'][str(args.level)] = ex1_level(torch, F, args.level, EPS, TOL, top_modes, ex1_maxiter,
                                                       setup=(gm, basis, plan, K2, mls))
        except Exception as exc:
            ex1['levels'][str(args.level)] = {'error': repr(exc)[:300]}

    # one hierarchy at a time: release the headline hierarchy, time the reference-shaped call, then the spec mode
    ml_levels, ml_reordered = mls[args.mode].levels, mls[args.mode].reordering is not None
    ml_sizes = mls[args.mode].level_sizes()
    mls.pop(args.mode)
    torch.cuda.empty_cache()
    # the reference-shaped call: solver_iter_mgcg(tol)(A, b) rebuilds the hierarchy inside every call
    # (ref main/example1/utils.py:150-164) - timed end to end, setup included, on the step's own system
    closure = None
    try:
        Kcl, _ = plan.apply(K2)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        sol = F.solver_iter_mgcg(tol=TOL, maxiter=2000, mode=args.mode)
        xcl, itcl = sol.solve_with_info(Kcl, b)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        closure = {'call': "solver_iter_mgcg(tol=%g, mode=%r)(K2c, b) with K2c and b on the device" % (TOL, args.mode),
                   's_per_call': dt, 'amg_setup_s': sol.last['setup_seconds'], 'solve_s': sol.last['solve_seconds'],
                   'iters': itcl, 'info': sol.last['info'], 'mdof_s': N / dt / 1e6,
                   'rel_err_vs_manufactured': float(torch.linalg.norm(xcl - xstar) / torch.linalg.norm(xstar))}
        del Kcl, xcl, sol
    except Exception as exc:
        closure = {'error': repr(exc)[:300]}
    torch.cuda.empty_cache()                 # the library allocates with cudaMalloc: hand torch's cached blocks back
    for mode in modes[1:]:                             # the spec mode beside it: same step function, fewer steps
        try:
            build(mode)
            k = max(1, min(2, args.steps))
            tms, rr = timed(False, mode, k, 1)
            by_mode[mode] = mode_block(mode, tms / k, None, rr)
            by_mode[mode]['steps'] = k
            x_by_mode[mode] = rr[-1]['x']
        except Exception as exc:                        # never let the side measurement take the bench line down
            by_mode[mode] = {'error': repr(exc)[:300]}
    if len(x_by_mode) == 2:
        xa, xb = x_by_mode[modes[0]], x_by_mode[modes[1]]
        by_mode['rel_diff_between_modes_tol1e-8_manufactured_rhs'] = float(torch.linalg.norm(xa - xb) / torch.linalg.norm(xb))
    del x_by_mode

    # free the big level before the side measurements
    same_cfg = None
    asm_renum_ms = None
    peak, peak_src = measured_peaks()
    spmv_bytes = 12 * nnzc + 20 * n                            # values 8 + col 4 per nnz; 20 B/row
    asm_bytes = 308 * nt                                       # DESIGN.md section 4 (gather formulation)
    spmv_gbs = spmv_bytes / (mean('spmv', rs) * 1e-3) / 1e9
    asm_gbs = asm_bytes / (mean('asm', rs) * 1e-3) / 1e9
    iters = rs[-1]['its']
    nlev = ml_levels
    sizes_main = ml_sizes
    revisit_main = [l for l in range(2, nlev - 1) if sizes_main and sizes_main[l] >= 10000] if (args.mode == 'sa' and os.environ.get('MWX_SA_CYCLE', 'auto') != 'v') else []
    per_vcycle = vcycle_launches(nlev, revisit_main)
    gpu_launches = int(args.steps * (2 + 1 + 1 + 1 + 3 + iters * (4 + per_vcycle)))
    level_main = args.level
    numbering = 'hilbert (solver-internal)' if ml_reordered else 'skfem'

    # release the level-L objects, then the smaller side blocks
    del mls, K2, Kc, plan, basis, pat, gm, vals, vals_c, perm_bufs, ysp, xstar, b, rs_e2e
    for r in rs:
        r.pop('x', None)
    torch.cuda.empty_cache()
    if args.level <= 12:                                       # same triangles, Hilbert-renumbered mesh (MeshTri.renumbered())
        try:
            gm2 = F.MeshTri().refine(args.level).renumbered()
            basis2 = F.InteriorBasis(gm2, F.ElementTriMorley(), intorder=5)
            K2r = F.asm_gpu(form, basis2)
            for _ in range(3):
                F.asm_gpu(form, basis2, out=K2r.d_data)
            e0, e1 = ev(), ev()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(5):
                F.asm_gpu(form, basis2, out=K2r.d_data)
            e1.record()
            torch.cuda.synchronize()
            asm_renum_ms = e0.elapsed_time(e1) / 5
            del K2r, basis2, gm2
        except Exception:
            asm_renum_ms = None
        torch.cuda.empty_cache()
    if ex1 is not None:
        for lvl in (args.level - 1, args.level - 2):
            if lvl < 3:
                continue
            try:
                ex1['levels'][str(lvl)] = ex1_level(torch, F, lvl, EPS, TOL, modes, ex1_maxiter)
            except Exception as exc:
                ex1['levels'][str(lvl)] = {'error': repr(exc)[:300]}
            torch.cuda.empty_cache()
        for lvl in (args.level, args.level - 1):         # convergence orders between consecutive levels (reference table: 2 / 2 / 1 / 1..2)
            hi, lo = ex1['levels'].get(str(lvl), {}), ex1['levels'].get(str(lvl - 1), {})
            for name in modes:
                if isinstance(hi.get(name), dict) and isinstance(lo.get(name), dict):
                    hi[name]['orders_vs_level_below'] = {k: float(-np.log2(hi[name][k] / lo[name][k]))
                                                         for k in ('L2', 'H1', 'H2', 'energy')}
                    if isinstance(hi[name].get('refined'), dict) and isinstance(lo[name].get('refined'), dict):
                        hi[name]['refined']['orders_vs_level_below'] = {
                            k: float(-np.log2(hi[name]['refined'][k] / lo[name]['refined'][k])) for k in ('L2', 'H1', 'H2', 'energy')}
    if not args.no_same_config:
        try:
            same_cfg = {'gpu': same_config_level8(torch, F, L, args.ref_level, TOL)}
        except Exception as exc:
            same_cfg = {'gpu': {'error': repr(exc)[:300]}}
        torch.cuda.empty_cache()
    micro = None
    if args.asm_levels:
        micro = []
        for lvl in args.asm_levels:
            try:
                micro.append(assembly_only(torch, F, lvl, peak))
            except Exception as exc:
                micro.append({'level': lvl, 'error': repr(exc)[:200]})
            torch.cuda.empty_cache()

    main_mode = by_mode[args.mode]
    line = {
        'metric': 'hot_path_MDOF_per_s', 'value': value, 'unit': 'MDOF/s', 'n_gpus': 1, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'strong',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'value_mode': args.mode,
        'value_incl_amg_setup': main_mode['value_incl_amg_setup_mdof_s'],
        'value_spec_mode': by_mode.get(spec_mode, {}).get('value_mdof_s'),
        'value_spec_mode_incl_amg_setup': by_mode.get(spec_mode, {}).get('value_incl_amg_setup_mdof_s'),
        'config': {'workload': f'MWX assemble+condense+SpMV+AMG-PCG, unit square level {level_main}: {N} Morley DOFs, '
                               f'{nt} triangles, nnz(K2)={nnz}, condensed n={n}; eps={EPS}, tol={TOL}, rhs=K2c@x* '
                               f'(x*~U(-1,1), seed 1234); `value` = AMG mode {args.mode!r}; the spec mode '
                               f'(reference RS hierarchy + Chebyshev) is timed beside it in `modes`/`value_spec_mode`',
                   'l2': 'operator (9.3 GB) and vectors (0.5 GB each) exceed the 126 MB L2; a 256 MiB buffer is '
                         'written before the SpMV sample'},
        'modes': by_mode,
        'reference_shaped_call': closure,
        'assembly_mdof_s': N / (mean('asm', rs) * 1e-3) / 1e6,
        'assembly_ms': mean('asm', rs), 'condense_ms': mean('condense', rs),
        'pcg': {'s_per_solve': mean('pcg', rs) * 1e-3, 'iters': iters, 'info': rs[-1]['info'], 'mode': args.mode,
                'ms_per_iter': mean('pcg', rs) / max(iters, 1), 'rel_err_vs_manufactured': xerr},
        'spmv': {'ms': mean('spmv', rs), 'gbs': spmv_gbs, 'frac_of_peak': spmv_gbs / peak,
                 'numbering': numbering,
                 'skfem_numbering_ms': mean('spmv_skfem', rs),
                 'skfem_numbering_gbs': spmv_bytes / (mean('spmv_skfem', rs) * 1e-3) / 1e9},
        'reorder_ms': mean('reorder', rs),
        'setup_s': {'mesh_refine': t_mesh, 'pattern_slotmap': t_pattern, 'assembly_tile_plan': t_tileplan, 'condense_plan': t_cplan,
                    'amg_total': t_amg[args.mode], 'amg_total_by_mode': t_amg},
        'roofline': {'kernel': 'k_spmv (CSR SpMV, fine level, solver numbering)', 'bound': 'hbm', 'achieved': spmv_gbs, 'peak': peak,
                     'unit': 'GB/s', 'frac': spmv_gbs / peak,
                     'traffic': kernel_traffic('k_spmv', level_main), 'peak_source': peak_src,
                     'algorithmic_bytes': spmv_bytes},
        'roofline_assembly': {'kernel': asm_kernel, 'bound': 'hbm', 'achieved': asm_gbs,
                              'peak': peak, 'unit': 'GB/s', 'frac': asm_gbs / peak,
                              'traffic': kernel_traffic('assembly', level_main),
                              'algorithmic_bytes': asm_bytes,
                              'renumbered_mesh': None if asm_renum_ms is None else {
                                  'ms': asm_renum_ms, 'mdof_s': N / (asm_renum_ms * 1e-3) / 1e6,
                                  'achieved': asm_bytes / (asm_renum_ms * 1e-3) / 1e9,
                                  'frac': asm_bytes / (asm_renum_ms * 1e-3) / 1e9 / peak}},
        'assembly_microbench': micro,
        'ex1_check': ex1,
        'same_config': same_cfg,
        'e2e': {'value': N / (e2e_ms / args.steps * 1e-3) / 1e6, 'unit': 'MDOF/s',
                'h2d_bytes_per_step': int(p_host.numel() * 8 + t_host.numel() * 4 + b_host.numel() * 8),
                'd2h_bytes_per_step': int(x_host.numel() * 8), 'ms_per_step': e2e_ms / args.steps},
        'gpu_launches': gpu_launches,
        'clocks': clocks,
    }
    if not args.no_cpu_baseline:
        t0 = time.perf_counter()
        r = cpu_reference_pass(args.ref_level)
        cpu_s = r['asm_s'] + r['condense_s'] + r['spmv_s'] + r['pcg_s']
        line['cpu_baseline'] = {
            'value': r['ndofs'] / cpu_s / 1e6, 'unit': 'MDOF/s', 'cores': 1, 'kind': 'port',
            'sample': f"level {args.ref_level} ({r['ndofs']} DOFs) of the same mesh family, one pass: asm {r['asm_s']:.2f} s, "
                      f"condense {r['condense_s']:.2f} s, SpMV {r['spmv_s'] * 1e3:.2f} ms, RS-AMG(sym. GS)-PCG "
                      f"{r['pcg_s']:.2f} s / {r['iters']} its (AMG setup {r['amg_setup_s']:.2f} s separate)",
            'assembly_mdof_s': r['ndofs'] / r['asm_s'] / 1e6,
            'spmv_gbs': (12 * r['nnz'] + 20 * r['n']) / r['spmv_s'] / 1e9,
            'pcg_ms_per_iter': r['pcg_s'] / max(r['iters'], 1) * 1e3,
            'host_cores_available': os.cpu_count(), 'wall_s': time.perf_counter() - t0}
        if same_cfg is not None:
            same_cfg['cpu'] = dict(level=args.ref_level, ndofs=r['ndofs'], n=r['n'], asm_s=r['asm_s'], condense_s=r['condense_s'],
                                   spmv_s=r['spmv_s'], amg_setup_s=r['amg_setup_s'], pcg_s=r['pcg_s'], iters=r['iters'],
                                   info=r['info'], value_mdof_s=r['ndofs'] / cpu_s / 1e6,
                                   value_incl_amg_setup_mdof_s=r['ndofs'] / (cpu_s + r['amg_setup_s']) / 1e6,
                                   pcg_ms_per_iter=r['pcg_s'] / max(r['iters'], 1) * 1e3, kind='oracle port, 1 core')
    print(json.dumps(line))


def assembly_only(torch, F, level, peak):
    """Assembly-only sample of one level (BASELINE config 'assembly-only microbench 2k x 2k - 8k x 8k')."""
    gm = F.MeshTri().refine(level)
    basis = F.InteriorBasis(gm, F.ElementTriMorley())
    form = EPS ** 2 * F.a_load + F.b_load
    K = F.asm_gpu(form, basis)
    for _ in range(3):
        F.asm_gpu(form, basis, out=K.d_data)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        F.asm_gpu(form, basis, out=K.d_data)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    return {'level': level, 'squares': f'{2 ** level}x{2 ** level}', 'dofs': basis.N, 'nnz': K.nnz, 'assembly_ms': ms,
            'gdof_s': basis.N / ms / 1e6, 'gbs_308B_per_element': 308 * gm.nt / ms / 1e6,
            'frac_of_peak': 308 * gm.nt / ms / 1e6 / peak}


def run_microbench(args, torch, F):
    """BASELINE.json config 'assembly-only microbench: MWX stiffness on 2k x 2k - 8k x 8k refined meshes, 1 B200'."""
    peak, peak_src = measured_peaks()
    rows = []
    for level in (11, 12, 13):
        t0 = time.perf_counter()
        gm = F.MeshTri().refine(level)
        basis = F.InteriorBasis(gm, F.ElementTriMorley())
        pat = basis.pattern()
        torch.cuda.synchronize()
        t_setup = time.perf_counter() - t0
        form = EPS ** 2 * F.a_load + F.b_load
        K = F.asm_gpu(form, basis)
        for _ in range(max(args.warmup, 3)):
            F.asm_gpu(form, basis, out=K.d_data)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(args.steps):
            F.asm_gpu(form, basis, out=K.d_data)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        row = {'level': level, 'squares': f'{2 ** level}x{2 ** level}', 'dofs': basis.N, 'triangles': gm.nt,
               'nnz': pat['nnz'], 'assembly_ms': ms, 'mdof_s': basis.N / ms / 1e3,
               'gbs_algorithmic_308B_per_element': 308 * gm.nt / ms / 1e6,
               'frac_of_peak': 308 * gm.nt / ms / 1e6 / peak, 'setup_s_mesh_pattern': t_setup}
        del K, basis, pat
        torch.cuda.empty_cache()
        # same triangles, Hilbert-renumbered vertices / elements (MeshTri.renumbered())
        t0 = time.perf_counter()
        gm2 = gm.renumbered()
        del gm
        basis = F.InteriorBasis(gm2, F.ElementTriMorley())
        basis.pattern()
        torch.cuda.synchronize()
        t_setup2 = time.perf_counter() - t0
        K = F.asm_gpu(form, basis)
        for _ in range(max(args.warmup, 3)):
            F.asm_gpu(form, basis, out=K.d_data)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(args.steps):
            F.asm_gpu(form, basis, out=K.d_data)
        e1.record()
        torch.cuda.synchronize()
        ms2 = e0.elapsed_time(e1) / args.steps
        row['renumbered_mesh'] = {'assembly_ms': ms2, 'mdof_s': basis.N / ms2 / 1e3,
                                  'gbs_algorithmic_308B_per_element': 308 * gm2.nt / ms2 / 1e6,
                                  'frac_of_peak': 308 * gm2.nt / ms2 / 1e6 / peak, 'setup_s_renumber_pattern': t_setup2}
        rows.append(row)
        del K, basis, gm2
        torch.cuda.empty_cache()
    print(json.dumps({'metric': 'assembly_MDOF_per_s', 'unit': 'MDOF/s', 'value': rows[1]['mdof_s'], 'n_gpus': 1,
                      'steps': args.steps, 'warmup': max(args.warmup, 3), 'dtype': 'f64', 'data': 'synthetic',
                      'config': {'workload': 'assembly-only microbench, K2 = eps^2 A + B on the 2k/4k/8k unit-square meshes; '
                                             'every pass rewrites 1.5-25 GB of CSR values (>> 126 MB L2)'},
                      'peak_gbs': peak, 'peak_source': peak_src, 'levels': rows}))


def run_partitioned(args, torch, dist, F, L, dev, rank, world, local):
    """N > 1: STRONG scaling on the fixed level-L mesh.  RCB row blocks, every rank assembles the rows it owns
    (no communication), PCG with one halo exchange + 3 all-reduces per iteration, block-Jacobi AMG."""
    from fem_forth_perturb_b200 import dist as D

    def sync_all():
        dist.barrier()
        torch.cuda.synchronize()

    t0 = time.perf_counter()
    gm = F.MeshTri().refine(args.level)
    basis = F.InteriorBasis(gm, F.ElementTriMorley(), intorder=5)
    pat = basis.pattern()
    torch.cuda.synchronize()
    t_mesh = time.perf_counter() - t0
    N, nt, nnz = basis.N, gm.nt, pat['nnz']
    form = EPS ** 2 * F.a_load + F.b_load
    bf = gm.boundary_facets()
    Dset = np.concatenate((np.unique(gm.facets[:, bf]), bf + gm.nv))
    t0 = time.perf_counter()
    # halo exchange: pack-and-push kernels into the neighbours' peer memory over NVLink (MWX_HALO=nccl: send/recv)
    halo_kind = os.environ.get('MWX_HALO', 'peer')
    try:
        comm = {'nccl': D.TorchDistComm, 'symm': D.SymmMemComm, 'peer': D.PeerComm}[halo_kind]()
    except Exception as exc:               # no symmetric memory / peer mapping on this box: NCCL send/recv + all-reduce (stated in `config`)
        if halo_kind == 'nccl':
            raise
        sys.stderr.write(f'[bench] {halo_kind} communicator unavailable ({exc!r}); falling back to NCCL\n')
        halo_kind, comm = 'nccl', D.TorchDistComm()
    prob = D.PartitionedProblem(basis, Dset, comm).assemble(form)
    (p,) = prob.parts
    torch.cuda.synchronize()
    t_part = time.perf_counter() - t0
    ops = D.CudaOps()
    p.ops = ops
    n_glob = prob.partition.n
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    xs_ext = comm.ext_zeros(p, dev)
    xs_ext[:p.nloc] = torch.rand(p.nloc, dtype=torch.float64, device=dev, generator=g) * 2 - 1
    comm.halo([p], [xs_ext])
    b = torch.empty(p.nloc, dtype=torch.float64, device=dev)
    ops.spmv(p, xs_ext, b)
    t0 = time.perf_counter()

    def global_matrix():                       # root only: the solver-numbered K2c the RS setup runs on (host, C++)
        from fem_forth_perturb_b200.solvers import Reordering
        Kc_full, _ = prob.cplan.apply(F.asm_gpu(form, basis))
        ro = Reordering.__new__(Reordering)
        ro.perm, ro.iperm, ro.n = prob.partition.perm, prob.partition.iperm, n_glob
        return ro.matrix(Kc_full)
    amg_mode = args.mode if args.mode in ('sa', 'chebyshev') else 'chebyshev'
    ml = D.DistAMG(prob.parts, comm, global_matrix, mode=amg_mode)       # global hierarchy, collective V-cycle
    torch.cuda.empty_cache()
    t_amg = time.perf_counter() - t0
    p_host = torch.empty(2 * gm.nv, dtype=torch.float64).pin_memory()
    t_host = torch.empty(3 * gm.nt, dtype=torch.int32).pin_memory()
    b_host = torch.empty(p.nloc, dtype=torch.float64).pin_memory()
    x_host = torch.empty(p.nloc, dtype=torch.float64).pin_memory()
    p_host.copy_(gm.d_pxy); t_host.copy_(gm.d_t); b_host.copy_(b)
    ysp = torch.empty(p.nloc, dtype=torch.float64, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)
    ev = lambda: torch.cuda.Event(enable_timing=True)

    step_events, region_events = [ev() for _ in range(7)], [ev(), ev()]      # created once, reused (see run_single)
    for e in step_events + region_events:
        e.record()
    torch.cuda.synchronize()

    def step(e2e):
        marks = step_events[:5]
        if e2e:
            gm.d_pxy.copy_(p_host, non_blocking=True)
            if p.tile_plan is None:                   # the tiled kernel reads the connectivity from its plan, not from t
                gm.d_t.copy_(t_host, non_blocking=True)
            bd = b_host.to(dev, non_blocking=True)
        else:
            bd = b
        marks[0].record()
        p.assemble(form)
        marks[1].record()
        flush.zero_()
        marks[2].record()
        ops.spmv(p, xs_ext, ysp)
        marks[3].record()
        flush.zero_()                      # a second SpMV sample per step (see `spmv.note`: one outlier is dropped)
        extra = step_events[5:7]
        extra[0].record()
        ops.spmv(p, xs_ext, ysp)
        extra[1].record()
        xs, its, info, resid = D.dist_pcg([p], comm, [bd], tol=TOL, maxiter=3000, precond='amg', amg=ml, ops=ops)
        marks[4].record()
        if e2e:
            x_host.copy_(xs[0], non_blocking=True)
        torch.cuda.synchronize()
        return dict(asm=marks[0].elapsed_time(marks[1]), spmv=marks[2].elapsed_time(marks[3]),
                    spmv2=extra[0].elapsed_time(extra[1]), pcg=extra[1].elapsed_time(marks[4]), its=its, info=info, x=xs[0])

    def timed(e2e, steps, warmup):
        for _ in range(warmup):
            step(e2e)
        sync_all()
        e0, e1 = region_events
        e0.record()
        rs = []
        for i in range(steps):
            r = step(e2e)
            if i < steps - 1:
                r['x'] = None                      # only the last iterate is looked at afterwards
            rs.append(r)
        e1.record()
        sync_all()
        tmax = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        return float(tmax.item()), rs

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        sampler.wait_first()
    total_ms, rs = timed(False, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    e2e_ms, _ = timed(True, args.steps, 1)
    err2 = torch.stack((torch.sum((rs[-1]['x'] - xs_ext[:p.nloc]) ** 2), torch.sum(xs_ext[:p.nloc] ** 2)))
    dist.all_reduce(err2)
    # per-rank maxima of the phases (device events), reduced with MAX
    sp_samples = sorted([r['spmv'] for r in rs] + [r['spmv2'] for r in rs])
    sp_mean = float(np.mean(sp_samples[:-1]))         # 2 samples per step; the largest one is dropped (clock-sampler hiccups)
    ph = torch.tensor([np.mean([r['asm'] for r in rs]), sp_mean, np.mean([r['pcg'] for r in rs]),
                       float(p.nnz), float(p.nloc), float(p.nghost), float(nt) / world], dtype=torch.float64, device=dev)
    phmax = ph.clone(); dist.all_reduce(phmax, op=dist.ReduceOp.MAX)
    phsum = ph.clone(); dist.all_reduce(phsum)
    if rank != 0:
        return
    peak, peak_src = measured_peaks()
    ms_step = total_ms / args.steps
    value = N / (ms_step * 1e-3) / 1e6
    iters = rs[-1]['its']
    nnzc_glob, nloc_max = float(phsum[3]), float(phmax[4])
    spmv_bytes_rank = 12 * float(phmax[3]) + 20 * nloc_max
    spmv_gbs_rank = spmv_bytes_rank / (float(phmax[1]) * 1e-3) / 1e9
    nlev = ml.nlevels
    # cycle kernels + per distributed level 5 halo pushes (2 more per revisit) + all-gather / copies around the tail; a lower bound
    per_vcycle = vcycle_launches(nlev, getattr(ml, 'revisit', ())) + 5 * ml.lsw + 3
    line = {
        'metric': 'hot_path_MDOF_per_s', 'value': value, 'unit': 'MDOF/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'strong',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic', 'value_mode': amg_mode,
        'value_incl_amg_setup': N / (ms_step * 1e-3 + t_amg) / 1e6,
        'config': {'workload': f'MWX assemble+SpMV+AMG-PCG, unit square level {args.level}: {N} Morley DOFs, {nt} triangles, '
                               f'condensed n={n_glob}, RCB-partitioned into {world} row blocks (Hilbert order inside a block); '
                               f'eps={EPS}, tol={TOL}, rhs=K2c@x*; preconditioner = global ' + ('smoothed-aggregation' if amg_mode == 'sa' else 'RS-AMG') + ' hierarchy applied as a '
                               f'collective Chebyshev V-cycle ({ml.lsw} row-blocked levels, tail of {nlev - ml.lsw} levels '
                               f'replicated); halo = '
                               + {'nccl': 'NCCL send/recv, all-reduces over NCCL', 'symm': 'pack-and-push kernels into peer memory over NVLink + signal-pad barrier, all-reduces over NCCL',
                                  'peer': 'every exchange of the iteration (halo push fused with the barrier, all-reduce of the CG scalars, all-gather for the replicated tail) is a kernel storing into peer memory over NVLink; iterations 2.. are one CUDA-graph launch each'}[halo_kind],
                   'l2': 'per-rank operator exceeds the 126 MB L2 for level >= 11; 256 MiB written before the SpMV sample'},
        'assembly_mdof_s': N / (float(phmax[0]) * 1e-3) / 1e6, 'assembly_ms': float(phmax[0]),
        'pcg': {'s_per_solve': float(phmax[2]) * 1e-3, 'iters': iters, 'info': rs[-1]['info'], 'mode': f'distributed {amg_mode} hierarchy, Chebyshev V-cycle',
                'ms_per_iter': float(phmax[2]) / max(iters, 1),
                'rel_err_vs_manufactured': float(torch.sqrt(err2[0] / err2[1])),
                'level_offsets': [o[-1] for o in ml.offsets], 'distributed_levels': ml.lsw, 'levels': nlev,
                'revisited_levels': list(getattr(ml, 'revisit', ())),
                'hierarchy_storage_bits': getattr(ml, 'storage_bits', 64),
                'block_csr_operators': sum(1 for lv in ml.levels for k in 'APR' for op in lv[k] if getattr(op, 'bsr', None)),
                'cuda_graph': bool(getattr(comm, 'graphable', False)) and os.environ.get('MWX_DIST_GRAPH', '1') != '0'},
        'spmv': {'ms': float(phmax[1]), 'gbs_per_gpu': spmv_gbs_rank, 'frac_of_peak': spmv_gbs_rank / peak,
                 'note': 'local block of the slowest rank, halo not included; two samples per step (256 MiB written before each), '
                         'mean of all but the largest one'},
        'partition': {'rows_max': int(nloc_max), 'ghosts_max': int(phmax[5]), 'nnz_total': int(nnzc_glob)},
        'setup_s': {'mesh_refine_pattern': t_mesh, 'partition_plan': t_part, 'amg_global_setup_and_distribution': t_amg,
                    'amg_setup_where': ('every rank builds the global hierarchy on its own GPU (device setup, bit-identical)' if amg_mode == 'sa' else 'root host, broadcast'),
                    'amg_device_setup_s': getattr(ml, 'setup_device_seconds', None)},
        'roofline': {'kernel': 'k_spmv (local row block, solver numbering)', 'bound': 'hbm', 'achieved': spmv_gbs_rank,
                     'peak': peak, 'unit': 'GB/s', 'frac': spmv_gbs_rank / peak, 'traffic': None, 'peak_source': peak_src,
                     'algorithmic_bytes': spmv_bytes_rank},
        'e2e': {'value': N / (e2e_ms / args.steps * 1e-3) / 1e6, 'unit': 'MDOF/s',
                'h2d_bytes_per_step': int((p_host.numel() * 8 + (t_host.numel() * 4 if p.tile_plan is None else 0)) * world + n_glob * 8),
                'd2h_bytes_per_step': int(n_glob * 8), 'ms_per_step': e2e_ms / args.steps},
        'gpu_launches': int(args.steps * world * (2 + 1 + iters * (4 + per_vcycle))),
        'clocks': clocks,
    }
    print(json.dumps(line))


if __name__ == '__main__':
    main()
