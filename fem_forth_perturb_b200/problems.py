"""The reference's problem data (loads and exact solutions) - ref main/example1/run.py:275-309 ('ex1') and
:466-493 ('ex3').

They are Python functions in the reference and stay Python functions here: ``asm_gpu(f_load, basis)`` and
``morley_error_norms`` sample them at the quadrature points on the host.  For meshes where nt*nq host samples are
not an option (level 12: 235 M points; 11 GB of samples for the error norms) the same two problems can be
evaluated by the device kernels: the objects below carry ``device_problem = (id, eps)``, which ``asm_gpu`` /
``morley_error_norms`` hand to ``mwx_load_vector_problem`` / ``mwx_morley_error_sums_problem`` when asked to
(``on_device=True``).  tests/test_gpu_examples.py checks host-sampled against device-evaluated.
"""
import numpy as np

from .forms import LinearFormGPU

pi, sin, cos = np.pi, np.sin, np.cos
PROBLEM_IDS = {'ex1': 1, 'ex3': 3}


class Problem:
    """f_load (a LinearForm), exact_u, dexact_u, ddexact - the four callables the reference's run.py defines."""

    def __init__(self, name, epsilon):
        if name not in PROBLEM_IDS:
            raise NotImplementedError(f"problem {name!r}: known {sorted(PROBLEM_IDS)}")
        self.name, self.epsilon = name, float(epsilon)
        self.device_problem = (PROBLEM_IDS[name], self.epsilon)
        eps = self.epsilon
        if name == 'ex1':
            def f(v, w):
                pix, piy = pi * w.x[0], pi * w.x[1]
                lu = 2 * pi ** 2 * (cos(2 * pix) * sin(piy) ** 2 + cos(2 * piy) * sin(pix) ** 2)
                llu = -8 * pi ** 4 * (cos(2 * pix) * sin(piy) ** 2 + cos(2 * piy) * sin(pix) ** 2
                                      - cos(2 * pix) * cos(2 * piy))
                return (eps ** 2 * llu - lu) * v
        else:
            def f(v, w):
                return (2 * pi ** 2 * sin(pi * w.x[0]) * sin(pi * w.x[1])) * v
        self.f_load = LinearFormGPU(f)
        self.f_load.device_problem = self.device_problem

    def exact_u(self, x, y):
        s = sin(pi * x) * sin(pi * y)
        return s ** 2 if self.name == 'ex1' else s

    def dexact_u(self, x, y):
        if self.name == 'ex1':
            return (2 * pi * cos(pi * x) * sin(pi * x) * sin(pi * y) ** 2,
                    2 * pi * cos(pi * y) * sin(pi * x) ** 2 * sin(pi * y))
        return pi * cos(pi * x) * sin(pi * y), pi * cos(pi * y) * sin(pi * x)

    def ddexact(self, x, y):
        if self.name == 'ex1':
            duxx = 2 * pi ** 2 * cos(pi * x) ** 2 * sin(pi * y) ** 2 - 2 * pi ** 2 * sin(pi * x) ** 2 * sin(pi * y) ** 2
            duxy = 2 * pi * cos(pi * x) * sin(pi * x) * 2 * pi * cos(pi * y) * sin(pi * y)
            duyy = 2 * pi ** 2 * cos(pi * y) ** 2 * sin(pi * x) ** 2 - 2 * pi ** 2 * sin(pi * y) ** 2 * sin(pi * x) ** 2
            return duxx, duxy, duxy, duyy
        duxx = -pi ** 2 * sin(pi * x) * sin(pi * y)
        duxy = pi * cos(pi * x) * pi * cos(pi * y)
        return duxx, duxy, duxy, duxx
