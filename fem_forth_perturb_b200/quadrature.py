"""Triangle quadrature rules as scikit-fem 2.1.1 selects them by ``intorder``:
5 -> 7-point Radon rule, 6 -> 12-point Dunavant, 8 -> 16-point Dunavant, 3/4 -> 6-point Dunavant, 2 -> 3-point,
1 -> centroid.  Points on the reference triangle (0,0),(1,0),(0,1); weights sum to 1/2."""
import itertools

import numpy as np


def _sym3(pts, w, wt, a, b):
    for P in ((a, b), (b, a), (b, b)):
        pts.append(P)
        w.append(wt)


def triangle_rule(intorder):
    pts, w = [], []
    if intorder <= 1:
        pts, w = [(1 / 3, 1 / 3)], [1.0]
    elif intorder == 2:
        pts, w = [(1 / 6, 1 / 6), (2 / 3, 1 / 6), (1 / 6, 2 / 3)], [1 / 3] * 3
    elif intorder in (3, 4):
        _sym3(pts, w, 0.223381589678011, 0.108103018168070, 0.445948490915965)
        _sym3(pts, w, 0.109951743655322, 0.816847572980459, 0.091576213509771)
    elif intorder == 5:
        s = np.sqrt(15.0)
        a1, a2 = (6 - s) / 21, (6 + s) / 21
        w1, w2 = (155 - s) / 1200, (155 + s) / 1200
        pts = [(1 / 3, 1 / 3), (a1, a1), (1 - 2 * a1, a1), (a1, 1 - 2 * a1), (a2, a2), (1 - 2 * a2, a2), (a2, 1 - 2 * a2)]
        w = [9 / 40, w1, w1, w1, w2, w2, w2]
    elif intorder == 6:
        _sym3(pts, w, 0.116786275726379, 0.501426509658179, 0.249286745170910)
        _sym3(pts, w, 0.050844906370207, 0.873821971016996, 0.063089014491502)
        for P in itertools.permutations((0.053145049844817, 0.310352451033784, 0.636502499121399)):
            pts.append(P[:2])
            w.append(0.082851075618374)
    elif intorder in (7, 8):
        pts, w = [(1 / 3, 1 / 3)], [0.144315607677787]
        _sym3(pts, w, 0.095091634267285, 0.081414823414554, 0.459292588292723)
        _sym3(pts, w, 0.103217370534718, 0.658861384496480, 0.170569307751760)
        _sym3(pts, w, 0.032458497623198, 0.898905543365938, 0.050547228317031)
        for P in itertools.permutations((0.008394777409958, 0.263112829634638, 0.728492392955404)):
            pts.append(P[:2])
            w.append(0.027230314174435)
    else:
        raise NotImplementedError(f"intorder {intorder}")
    return np.array(pts, dtype=np.float64).T, 0.5 * np.array(w, dtype=np.float64)
