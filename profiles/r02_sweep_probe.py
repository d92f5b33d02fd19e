import sys, os, time, warnings, json
sys.path.insert(0, '.'); sys.path.insert(0, './examples')
import numpy as np
import example1, example2
t=time.time()
with warnings.catch_warnings(record=True) as w:
    warnings.simplefilter('always')
    res, its = example1.run(refine_time=8, epsilon_range=7, element_type='P1', maxiter=1000, out=lambda *a: None)
print('sweep s', time.time()-t, 'warnings', [str(x.message) for x in w])
gi=json.load(open('./tests/golden/amg1_iterations.json'))
gn=json.load(open('./tests/golden/error_norms.json'))['amg1_ex1_P1_8']
for eps, pairs in its.items():
    k=repr(float(eps))
    print(eps, 'w', [p[0] for p in pairs], gi['w'].get(k)); print(eps, 'u', [p[1] for p in pairs], gi['u'].get(k))
    if k in gn['rows']:
        gold=np.array(gn['rows'][k]); print('  relerr per level', (np.abs(res[eps]-gold)/np.abs(gold)).max(axis=1))
    else:
        print(res[eps])
for el in ('P1','P2'):
    for pen in (True, False):
        t=time.time()
        tab, it = example2.run(6, 7, el, pen, out=lambda *a: None)
        gold=np.array(json.load(open('./tests/golden/example2_tables.json'))['tables']['{}_{}'.format(el,'pen' if pen else 'nopen')]['rows'])
        print(el, pen, 'time', time.time()-t, 'iters', it); print(np.abs(tab-gold)/np.abs(gold))
