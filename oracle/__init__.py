"""CPU oracle for the MWX assembly + PCG hot path.  TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (NumPy + a little C) of what the reference
stack computes on the two hot paths of YerbaPage/FEM_forth_perturb:

* path A: ``asm(a_load/b_load/penalty_k, basis)``           ref: main/example1/run.py:110-146,210,253-260
* path B: ``solver_iter_mgcg`` / ``solver_iter_pcg`` + ``solve``/``condense``
                                                             ref: main/example1/utils.py:127-166,265-460

The arithmetic of both paths lives in three third-party packages that are NOT
vendored in the reference and cannot be installed here (no network):
scikit-fem 2.1.1, pyamg 4.0.0, scipy 1.4.1 (pins: ref README.md:14-23).  Their
published algorithms are restated here; every function cites the reference call
site it stands in for.

Parity status ("pinned" = checked against artefacts the reference ships):
  * error norms of the whole pipeline ............ PINNED  (golden CSVs, tests/golden/)
  * CG iteration counts (AMG-PCG, w and u solves). PINNED  (golden log, tests/golden/)
  * CSR indptr/indices, matrix values, hierarchy . parity unpinned directly - the
    reference stores no matrix anywhere; they are pinned only through the 16-digit
    level-1 norms (direct-solve exact) and the order-sensitive iteration table.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package.  The product package
``fem_forth_perturb_b200`` never does, and raises if its CUDA library is missing.
"""
