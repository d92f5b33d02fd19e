"""Extract the reference's own golden artefacts into small JSON fixtures (run in the build container;
/root/reference does not exist on the GPU box).  Nothing here is reference SOURCE - only the numbers
its committed logs hold for the hot path.

    python tests/golden/make_golden.py
"""
import csv
import json
import os
import re

REF = '/root/reference/main'
OUT = os.path.dirname(os.path.abspath(__file__))


def read_csv(path):
    with open(path, newline='') as f:
        rows = list(csv.reader(f))
    return [[float(v) for v in r[1:]] for r in rows[1:]]


def norms():
    src = {
        # key: (file, levels per epsilon block, epsilons, description)
        'ex1_P1_nopen': ('example1/log/ex1_P1_nopen_2021-02-22_21-42-42.csv', 6,
                         [1, 1e-1, 1e-2, 1e-3, 1e-4, 1e-5], dict(example='ex1', w='P1', penalty=False, intorder=5, tol=1e-8)),
        'ex1_P2_nopen': ('example1/log/ex1_P2_nopen_2021-02-22_21-52-49.csv', 6,
                         [1, 1e-1, 1e-2, 1e-3, 1e-4, 1e-5], dict(example='ex1', w='P2', penalty=False, intorder=5, tol=1e-8)),
        'amg1_ex1_P1_8': ('example1/test/log/test_solver_original_amg1_ex1_8_2020-11-20_12-58-38.csv', 8,
                          [1, 1e-1, 1e-2, 1e-3, 1e-4, 1e-5], dict(example='ex1', w='P1', penalty=False, intorder=5, tol=1e-8, maxiter=1000)),
        'ex3_P1_pen': ('example1/log/ex3_range_2/ex3_P1_pen_2020-11-22_16-21-26.csv', 7,
                       [1, 1e-2, 1e-4, 1e-6], dict(example='ex3', w='P1', penalty=True, intorder=6, tol=1e-8, sigma=5)),
        'ex3_P2_pen': ('example1/log/ex3_range_2/ex3_P2_pen_2020-11-22_16-21-10.csv', 7,
                       [1, 1e-2, 1e-4, 1e-6], dict(example='ex3', w='P2', penalty=True, intorder=6, tol=1e-8, sigma=5)),
        'ex4_P1_nopen': ('example3/log/ex4_P1_nopen_2021-02-22_12-51-55.csv', 6,
                         [1e-5, 1e-6], dict(example='ex4', w='P1', penalty=False, intorder=8, tol=1e-11)),
    }
    out = {}
    for key, (rel, nlev, epss, cfg) in src.items():
        rows = read_csv(os.path.join(REF, rel))
        assert len(rows) == nlev * len(epss), (key, len(rows))
        out[key] = dict(source='main/' + rel, columns=['L2', 'H1', 'H2', 'Energy'], levels=nlev,
                        epsilons=epss, config=cfg,
                        rows={repr(float(e)): rows[i * nlev:(i + 1) * nlev] for i, e in enumerate(epss)})
    return out


def example2():
    """main/example2/log/ex3_*_2021-02-06_*.csv: rows 0-5 = levels 1-6 against the level-7 solution (the files were
    extended by hand below those rows with a transposed copy and a LaTeX table; only the first block is data)."""
    src = {'P1_pen': 'ex3_P1_pen_2021-02-06_10-55-09.csv', 'P1_nopen': 'ex3_P1_nopen_2021-02-06_10-55-24.csv',
           'P2_pen': 'ex3_P2_pen_2021-02-06_10-55-13.csv', 'P2_nopen': 'ex3_P2_nopen_2021-02-06_10-55-20.csv'}
    out = dict(config=dict(example='ex3', base_order=7, refine_time=6, intorder=3, tol=1e-8, epsilon=1e-6, sigma=5,
                           solver_type='mgcg'), columns=['L2', 'H1', 'H2', 'Energy'], tables={})
    for key, name in src.items():
        with open(os.path.join(REF, 'example2/log', name), newline='') as f:
            rows = list(csv.reader(f))[1:7]
        assert [r[0] for r in rows] == [str(i) for i in range(6)], key
        out['tables'][key] = dict(source='main/example2/log/' + name, rows=[[float(v) for v in r[1:5]] for r in rows])
    return out


def iterations():
    rel = 'example1/test/log/test_solver_original_amg1_ex1_8_2020-11-20_12-58-38.txt'
    txt = open(os.path.join(REF, rel)).read().replace('\r', '')
    blocks = re.split(r'^epsilon = ', txt, flags=re.M)[1:]
    out = dict(source='main/' + rel, config=dict(example='ex1', w='P1', tol=1e-8, maxiter=1000, intorder=5),
               w={}, u={}, dofs={}, seconds={})
    for blk in blocks:
        eps = repr(float(blk.split('\n', 1)[0]))
        its = [int(v) for v in re.findall(r'mgcg total interation steps: (\d+)', blk)]
        out['w'][eps] = its[0::2]
        out['u'][eps] = its[1::2]
        out['dofs'][eps] = [int(v) for v in re.findall(r'dofs: (\d+)', blk)]
        out['seconds'][eps] = [float(v) for v in re.findall(r'MGCG Time Cost ([0-9.e+-]+) s', blk)]
    return out


if __name__ == '__main__':
    with open(os.path.join(OUT, 'error_norms.json'), 'w') as f:
        json.dump(norms(), f, indent=1)
    with open(os.path.join(OUT, 'amg1_iterations.json'), 'w') as f:
        json.dump(iterations(), f, indent=1)
    with open(os.path.join(OUT, 'example2_tables.json'), 'w') as f:
        json.dump(example2(), f, indent=1)
    print('wrote', os.listdir(OUT))
