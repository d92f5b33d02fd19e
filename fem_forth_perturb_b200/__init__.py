"""fem_forth_perturb_b200 - B200 (sm_100a) drop-in for the two data-parallel hot paths of
YerbaPage/FEM_forth_perturb: Morley-Wang-Xu stiffness assembly (``asm_gpu``) and the
Jacobi-/AMG-preconditioned CG solve (``solver_iter_pcg`` / ``solver_iter_mgcg``).

Everything numerical runs in hand-written CUDA kernels behind the C ABI of ``include/mwx_b200.h``
(``libmwx_b200.so``, loaded through ctypes).  PyTorch only allocates device buffers.  There is no CPU
fallback: without the library or without a GPU the entry points raise.
"""
from ._lib import MwxError, LIB_PATH                                     # noqa: F401
from .mesh import MeshTri                                                # noqa: F401
from .basis import (ElementTriMorley, ElementTriP1, ElementTriP2,        # noqa: F401
                    InteriorBasis, FacetBasis)
from .forms import (BilinearFormGPU, a_load, b_load, penalty_1,          # noqa: F401
                    penalty_2, penalty_3, laplace, wv_load, mass, LinearForm, LinearFormGPU, as_form)
from .assembly import (asm_gpu, asm, DeviceCSR, MixedAction, morley_error_norms,   # noqa: F401
                       morley_facet_pnv, morley_nested_error_norms, project, assembly_kernel_name, TilePlan, hilbert_order)
from .problems import Problem                                            # noqa: F401
from .solvers import (solver_iter_krylov, solver_iter_krylov_iter,       # noqa: F401
                      solver_iter_pcg, solver_iter_mgcg, solver_iter_mgcg_iter, solver_iter_pyamg,
                      build_pc_diag, solve, condense, CondensePlan, AMGHierarchy, pcg, refine, residual_dd)

__all__ = ['MeshTri', 'ElementTriMorley', 'ElementTriP1', 'ElementTriP2', 'InteriorBasis', 'FacetBasis',
           'a_load', 'b_load', 'penalty_1', 'penalty_2', 'penalty_3', 'laplace', 'wv_load', 'LinearForm', 'asm_gpu', 'asm',
           'DeviceCSR', 'MixedAction', 'morley_error_norms', 'morley_facet_pnv', 'morley_nested_error_norms', 'project', 'mass',
           'solver_iter_krylov', 'solver_iter_krylov_iter', 'solver_iter_pcg', 'solver_iter_mgcg',
           'solver_iter_mgcg_iter', 'solver_iter_pyamg', 'build_pc_diag', 'solve', 'condense', 'CondensePlan', 'AMGHierarchy', 'pcg', 'refine', 'residual_dd', 'MwxError', 'Problem']
