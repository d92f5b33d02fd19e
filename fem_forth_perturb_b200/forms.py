"""Form descriptors for ``asm_gpu``.

In the reference the forms are opaque Python closures over skfem ``DiscreteField``s
(``@BilinearForm def a_load(u, v, w): return ddot(dd(u), dd(v))`` ..., ref main/example1/run.py:110-146).
The device kernels implement exactly those integrands, so ``asm_gpu`` dispatches on the form's NAME:
it accepts the descriptors below, or any object whose ``__name__`` (or skfem's ``.form.__name__``) is
one of them.  Anything else raises - there is no CPU fallback that would evaluate arbitrary closures.

Descriptors form a small algebra so that ``eps**2 * a_load + b_load`` assembles in ONE kernel pass.
"""

INTERIOR = ('a_load', 'b_load')
LAGRANGE = ('laplace',)          # grad.grad on P1 / P2 (w-problem)
MASS = ('mass',)                 # u v on the Morley basis (L2 projection of example 3)
MIXED = ('wv_load',)             # grad(w_h) . grad(v) with w in P1/P2, v Morley (applied, not stored)
FACET = ('penalty_1', 'penalty_2', 'penalty_3')
KNOWN = INTERIOR + FACET + LAGRANGE + MIXED + MASS


class BilinearFormGPU:
    """Linear combination of the known MWX integrands: {name: coefficient}."""

    def __init__(self, terms, name=None):
        self.terms = {k: float(v) for k, v in terms.items() if v != 0.0}
        for k in self.terms:
            if k not in KNOWN:
                raise NotImplementedError(f"form '{k}' has no device kernel (known: {KNOWN})")
        self.__name__ = name or '+'.join(sorted(self.terms))

    def __mul__(self, c):
        return BilinearFormGPU({k: v * float(c) for k, v in self.terms.items()})

    __rmul__ = __mul__

    def __add__(self, other):
        other = as_form(other)
        out = dict(self.terms)
        for k, v in other.terms.items():
            out[k] = out.get(k, 0.0) + v
        return BilinearFormGPU(out)

    def __neg__(self):
        return self * -1.0

    def __sub__(self, other):
        return self + (-as_form(other))

    @property
    def kind(self):
        names = set(self.terms)
        if names <= set(INTERIOR):
            return 'interior'
        if names <= set(FACET):
            return 'facet'
        if names <= set(LAGRANGE):
            return 'lagrange'
        if names <= set(MIXED):
            return 'mixed'
        if names <= set(MASS):
            return 'mass'
        raise NotImplementedError("interior and facet integrands need different bases; assemble them "
                                  "separately and add the matrices (as ref main/example1/run.py:260 does)")

    def __repr__(self):
        return 'BilinearFormGPU(' + ' + '.join(f'{v:g}*{k}' for k, v in self.terms.items()) + ')'


a_load = BilinearFormGPU({'a_load': 1.0}, 'a_load')          # ddot(dd(u), dd(v))      ref run.py:110-115
b_load = BilinearFormGPU({'b_load': 1.0}, 'b_load')          # dot(grad(u), grad(v))   ref run.py:118-123
penalty_1 = BilinearFormGPU({'penalty_1': 1.0}, 'penalty_1')  # ref run.py:134-136
penalty_2 = BilinearFormGPU({'penalty_2': 1.0}, 'penalty_2')  # ref run.py:139-141
penalty_3 = BilinearFormGPU({'penalty_3': 1.0}, 'penalty_3')  # ref run.py:144-146 (sigma passed to asm_gpu)
laplace = BilinearFormGPU({'laplace': 1.0}, 'laplace')        # ref run.py:149-154 (P1 / P2 bases)
mass = BilinearFormGPU({'mass': 1.0}, 'mass')                 # skfem.models.poisson.mass on the Morley basis
wv_load = BilinearFormGPU({'wv_load': 1.0}, 'wv_load')        # ref run.py:126-131 (trial P1/P2, test Morley)


class LinearFormGPU:
    """``@LinearForm def f_load(v, w): return f(w.x) * v`` (ref run.py:277-288).  The integrand is sampled on the
    host - it is the reference's own Python closure, called with ``v = 1`` and ``w.x`` = the quadrature points -
    and integrated against the basis on the device."""

    def __init__(self, fn):
        self.fn = fn
        self.__name__ = getattr(fn, '__name__', 'linear_form')

    def sample(self, basis):
        import numpy as np

        class _W:
            pass
        w = _W()
        w.x = basis.global_coordinates()
        vals = self.fn(np.ones(w.x.shape[1:]), w)
        return np.array(np.broadcast_to(np.asarray(vals, dtype=np.float64), w.x.shape[1:]), dtype=np.float64, order='C')


LinearForm = LinearFormGPU


def as_form(form):
    """Resolve a descriptor, a name, or a (skfem-)decorated function called like the reference's."""
    if isinstance(form, BilinearFormGPU):
        return form
    name = form if isinstance(form, str) else None
    if name is None:
        inner = getattr(form, 'form', None)                  # skfem BilinearForm keeps the function in .form
        name = getattr(inner, '__name__', None) or getattr(form, '__name__', None)
    if name in KNOWN:
        return BilinearFormGPU({name: 1.0}, name)
    raise NotImplementedError(
        f"asm_gpu: no device kernel for form {name!r}; supported: {KNOWN}. "
        "Arbitrary Python integrands are not evaluated on the CPU as a fallback.")
