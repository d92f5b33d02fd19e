"""Write profiles/ncu_traffic.json (read by bench.py's `roofline.traffic`) from an `ncu --set full` report:
    python profiles/ncu_traffic.py REPORT.ncu-rep LEVEL
For every kernel of interest the LAST captured launch gives dram__bytes_read.sum + dram__bytes_write.sum; the entry
records the sha256 of the kernel's source file, so that bench.py reports null once the source has changed."""
import csv, hashlib, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KERNELS = {'k_spmv<0, 0, 0>': ('k_spmv', 'fem_forth_perturb_b200/csrc/krylov.cu'), 'k_spmv<false, 0, 0>': ('k_spmv', 'fem_forth_perturb_b200/csrc/krylov.cu'),
           'k_assemble_morley_tiles': ('assembly', 'fem_forth_perturb_b200/csrc/asm_tiles.cu')}
UNIT = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}
rep, level = sys.argv[1], int(sys.argv[2])
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
path = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
table = json.load(open(path)) if os.path.exists(path) else {}
for r in rows[2:]:
    d, u = dict(zip(hdr, r)), dict(zip(hdr, units))
    for pat, (key, src) in KERNELS.items():
        if pat in d['Kernel Name']:
            b = sum(float(d[m].replace(',', '')) * UNIT[u[m]] for m in ('dram__bytes_read.sum', 'dram__bytes_write.sum'))
            with open(os.path.join(ROOT, src), 'rb') as f:
                sha = hashlib.sha256(f.read()).hexdigest()
            table[f'{key}@level{level}'] = {'bytes': b, 'ms': float(d['gpu__time_duration.sum'].replace(',', '')) * {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 'msecond': 1.0, 'usecond': 1e-3, 'nsecond': 1e-6}.get(u['gpu__time_duration.sum'], 1.0),
                                                'kernel': d['Kernel Name'][:80], 'report': os.path.basename(rep), 'source': src, 'source_sha256': sha}
json.dump(table, open(path, 'w'), indent=1)
print(json.dumps(table, indent=1))
