"""Per-kernel counts of the SASS mnemonics that show how data moves (TMA bulk copies, mbarrier waits, cp.async, wide
loads) in the built library:   python profiles/sass_summary.py > profiles/r02_sass_summary.txt
(cuobjdump -sass fem_forth_perturb_b200/libmwx_b200.so; no GPU needed.)"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, 'fem_forth_perturb_b200', 'libmwx_b200.so')
out = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
pats = collections.OrderedDict([('UBLKCP (TMA bulk copy)', r'\bUBLKCP'), ('SYNCS (mbarrier)', r'\bSYNCS'), ('LDGSTS (cp.async)', r'\bLDGSTS'),
                                ('LDG.*128/256', r'\bLDG\.E\.[A-Z0-9.]*(128|256)'), ('STG.*128', r'\bSTG\.E\.[A-Z0-9.]*128'),
                                ('ATOMS/ATOMG', r'\bATOM[SG]'), ('DFMA', r'\bDFMA'), ('HMMA/UTCMMA (tensor)', r'\b(HMMA|UTC.MMA|UTCHMMA)')])
kern, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        d = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        d = d.replace('(anonymous namespace)::', '').replace('void ', '')
        kern = re.match(r'[\w:]+(<[^(]*>)?', d).group(0)
        counts.setdefault(kern, collections.Counter())
        continue
    if kern:
        counts[kern]['instructions'] += bool(re.search(r'^\s+/\*[0-9a-f]{4}\*/', line))
        for name, pat in pats.items():
            if re.search(pat, line):
                counts[kern][name] += 1
print('# cuobjdump -sass libmwx_b200.so (sm_100a), mnemonic counts per kernel; kernels without any of them are listed by size only')
cols = list(pats)
print('kernel'.ljust(58), 'instr'.rjust(6), ' '.join(c.split(' ')[0].rjust(8) for c in cols))
for k, c in sorted(counts.items(), key=lambda kv: -kv[1]['instructions']):
    print(k[-58:].ljust(58), str(c['instructions']).rjust(6), ' '.join(str(c[n]).rjust(8) for n in cols))
