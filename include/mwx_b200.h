/* mwx_b200.h - C ABI of libmwx_b200.so: the B200 (sm_100a) drop-in for the two data-parallel hot
 * paths of YerbaPage/FEM_forth_perturb.
 *
 * The reference is pure Python over un-vendored scikit-fem 2.1.1 / scipy 1.4.1 / pyamg 4.0.0 and
 * has no FFI of its own.  Each entry point below names the reference call
 * (file:line under /root/reference) whose arithmetic it replaces; INTEGRATION.md shows the ctypes
 * stub a maintainer would add to main/example1/utils.py / run.py.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless the name ends in _host; the caller owns all buffers;
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *  - return value: 0 = ok, <0 = MWX_E_* (bad argument / capacity), >0 = cudaError_t;
 *  - no global state, no exceptions, re-entrant per set of buffers;
 *  - mesh arrays are structure-of-arrays: t = [t0 | t1 | t2] (3*nt int32), facets = [f0 | f1],
 *    t2f = [e0 | e1 | e2], coordinates pxy = interleaved (x,y) pairs (2*nv doubles);
 *  - CSR: indptr int64 (N+1), indices int32, values fp64.  (scipy-compatible after a dtype cast;
 *    int64 indptr is what makes the 8k x 8k mesh, nnz = 3.09e9, addressable.)
 */
#ifndef MWX_B200_H
#define MWX_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MWX_E_BADARG   (-1)
#define MWX_E_CAPACITY (-2)   /* a per-row / per-vertex local capacity was exceeded        */
#define MWX_E_NOCONV   (-3)   /* iteration limit reached (result still written)           */
#define MWX_E_NOMEM    (-4)

int mwx_version(void);

/* ------------------------------------------------------------------ mesh  (hot path A, setup)
 * Replaces skfem MeshTri._build_mappings / refine reached from
 *   ref main/example1/run.py:508 (MeshTri()), :512 (m.refine()).                              */

/* Sort every column of t ascending (in place). */
int mwx_mesh_sort_columns(int64_t nt, int32_t *t, void *stream);

/* Entity -> element incidence (CSR) of a (ndof_per_elem x nt) SoA connectivity:
 * inc_ptr (nent+1), inc_idx (ndof_per_elem*nt) holding codes (element << 3 | local index),
 * ascending inside each entity.  Used for vertex->triangle and DOF->triangle maps. */
int mwx_incidence(int64_t nent, int64_t nt, int32_t ndof_per_elem, const int32_t *conn,
                  int64_t *inc_ptr, int32_t *inc_idx, void *stream);

/* Facets = lexicographically sorted unique (min,max) vertex pairs.  Pass 1 fills fstart (nv+1):
 * facets whose smaller vertex is a are numbered fstart[a] .. fstart[a+1]-1, and returns nf. */
int mwx_mesh_facets_count(int64_t nv, int64_t nt, const int32_t *t, const int64_t *v2t_ptr,
                          const int32_t *v2t_idx, int64_t *fstart, int64_t *nf_host, void *stream);
/* Pass 2 writes facets (2*nf) and t2f (3*nt; local edges (0,1),(1,2),(0,2)). */
int mwx_mesh_facets_fill(int64_t nv, int64_t nt, const int32_t *t, const int64_t *v2t_ptr,
                         const int32_t *v2t_idx, const int64_t *fstart, int64_t nf,
                         int32_t *facets, int32_t *t2f, void *stream);

/* Uniform (red) refinement: vertex nv+j = midpoint of facet j; children in skfem's block order
 * (t0,m01,m02),(t1,m01,m12),(t2,m02,m12),(m01,m12,m02), columns sorted.  Outputs: pxy_new
 * (2*(nv+nf)), t_new (3*4*nt). */
int mwx_mesh_refine(int64_t nv, int64_t nt, int64_t nf, const double *pxy, const int32_t *t,
                    const int32_t *facets, const int32_t *t2f, double *pxy_new, int32_t *t_new,
                    void *stream);

/* Morley / P2 element_dofs (6*nt): rows 0-2 = t, rows 3-5 = nv + t2f  (skfem Dofs). */
int mwx_element_dofs(int64_t nv, int64_t nt, const int32_t *t, const int32_t *t2f,
                     int32_t *edofs, void *stream);

/* Boundary facets: count of elements per facet == 1.  facet_count (nf) int32 output. */
int mwx_facet_element_count(int64_t nf, int64_t nt, const int32_t *t2f, int32_t *facet_count,
                            void *stream);

/* ------------------------------------------------------------------ CSR pattern + slot map
 * Replaces scipy coo_matrix(...).tocsr() reached from skfem BilinearForm.assemble
 *   ref main/example1/run.py:210 (asm(a_load, basis['u']) + asm(b_load, basis['u'])).
 * Pass 1: row lengths -> indptr (N+1), returns nnz and the longest row. */
int mwx_pattern_count(int64_t ndofs, int64_t nt, int32_t ndpe, const int32_t *edofs, const int64_t *r2e_ptr,
                      const int32_t *r2e_idx, int64_t *indptr, int64_t *nnz_host,
                      int32_t *max_row_host, void *stream);
/* Pass 2: indices (nnz, ascending per row) and the slot map: for pair p of r2e_idx (row r, element e)
 * slot8[8*p + j] = position of column edofs[j][e] inside row r (j < ndpe <= 6).  Byte 7 is an accumulate mask:
 * bit j set = an earlier pair of the same row already wrote that slot (this pair adds), clear = first touch (plain
 * store) - mwx_assemble_morley relies on it to skip the zero-fill of its shared-memory row image; byte 6 unused.
 * ndpe = DOFs per element: 6 (Morley, P2) or 3 (P1). */
int mwx_pattern_fill(int64_t ndofs, int64_t nt, int32_t ndpe, const int32_t *edofs, const int64_t *r2e_ptr,
                     const int32_t *r2e_idx, const int64_t *indptr, int32_t *indices,
                     uint8_t *slot8, void *stream);

/* ------------------------------------------------------------------ Morley element kernel (hot)
 * values[k] = ca * (hess u : hess v) + cb * (grad u . grad v), summed over elements in ascending
 * element order per CSR slot (deterministic, no atomics).  Replaces
 *   K2 = epsilon**2 * asm(a_load, basis['u']) + asm(b_load, basis['u'])   ref main/example1/run.py:210
 * (a_load/b_load: ref main/example1/run.py:110-123).  accumulate != 0 adds into values. */
int mwx_assemble_morley(int64_t ndofs, int64_t nt, const double *pxy, const int32_t *t, double *exy,
                        int32_t refresh_exy, const int64_t *r2e_ptr, const int32_t *r2e_idx,
                        const uint8_t *slot8, const int64_t *indptr, double ca, double cb,
                        int32_t accumulate, double *values, void *stream);
/* exy: workspace of 8*nt doubles (64-byte aligned): per-element geometry (gam_0..2, mu_01, mu_02, mu_12, |K|, -), one
 * 64-byte record per element so that a (row, element) pair costs two sectors instead of six scattered ones;
 * refresh_exy != 0 regathers it from (pxy, t) first (needed whenever the coordinates changed). */

/* Boundary-facet (Nitsche penalty) terms, ref main/example1/run.py:134-146,253-260:
 * values += c1*penalty_1 + c2*penalty_2 + c3*penalty_3(sigma) over nbf boundary facets (FacetBasis:
 * 3-point Gauss rule, w.n = outward unit normal, w.h = facet length).  btri[b] = owning triangle,
 * blocal[b] = local edge (0,1,2) of the facet in it.  Gather form, like the interior kernel: one
 * thread per touched matrix row rows[r] walks pair_code[pair_ptr[r] .. pair_ptr[r+1]) = (b << 3 | i)
 * in ascending order and adds row i of facet b's 6x6 matrix - fixed order, no atomics, O(sqrt N). */
int mwx_assemble_penalty(int64_t nt, int64_t nbf, const double *pxy, const int32_t *t,
                         const int32_t *edofs, const int32_t *btri, const int32_t *blocal,
                         int64_t nrows, const int32_t *rows, const int64_t *pair_ptr,
                         const int32_t *pair_code, const int64_t *indptr, const int32_t *indices,
                         double c1, double c2, double c3, double sigma, double *values, void *stream);

/* values_out = alpha * a + beta * b  (same pattern; scipy's `eps**2 * A + B`). */
int mwx_axpby(int64_t n, double alpha, const double *a, double beta, const double *b,
              double *out, void *stream);

/* Same kernel for a row subset (partitioned assembly, SURVEY 8e: every rank assembles the rows it owns, no
 * communication): row_ids (nrows) lists matrix rows, the rows are written back to back through out_indptr
 * (nrows+1; out_indptr[i+1]-out_indptr[i] = length of row row_ids[i] in the global pattern). */
int mwx_assemble_morley_rows(int64_t nrows, const int32_t *row_ids, const int64_t *out_indptr, int64_t nt,
                             const double *pxy, const int32_t *t, double *exy, int32_t refresh_exy,
                             const int64_t *r2e_ptr, const int32_t *r2e_idx, const uint8_t *slot8,
                             const int64_t *indptr, double ca, double cb, double *values, void *stream);

/* ------------------------------------------------------------------ Morley element kernel, tiled form (the default)
 * Same matrix as mwx_assemble_morley, organised by a PLAN built once per (mesh, destination layout):
 *   row_dof[i]  (nrows)  skfem DOF of the i-th destination row, in a LOCALITY order (Hilbert curve over the DOF
 *                        locations); the rows are cut into tiles of <= 256 consecutive entries of this list;
 *   dst_off[i], dst_len[i]  where row i lives in the destination value array and how many entries it has;
 *   dst_col[dst_off[i] + k] skfem DOF of the k-th destination entry of row i (any order; columns of the element
 *                        matrices that do not occur are dropped - a condensed or rank-local layout).
 * For the global K2: row_dof = Hilbert order of all DOFs, dst_off = indptr[row_dof], dst_col = indices.
 * t (3*nt), edofs (6*nt), r2e_* as for mwx_pattern_*.  Per tile the plan stores the unique vertices / elements of
 * the patch and, per destination entry, the <= 2 contributing (element, local pair) codes; the kernel evaluates
 * every unique element of a tile ONCE (21 entries of the symmetric local matrix, shared memory) and writes each
 * value with one coalesced store - no atomics, fixed summation order (ascending element id).
 * Returns MWX_E_CAPACITY when the mesh does not fit the tile limits (vertex valence > 16, rows longer than 64
 * entries, an edge with three triangles): callers then use mwx_assemble_morley. */
typedef struct mwx_asm_plan mwx_asm_plan_t;
int mwx_asm_plan_create(int64_t nrows, const int32_t *row_dof, const int64_t *dst_off, const int32_t *dst_len,
                        const int32_t *dst_col, int64_t nt, const int32_t *t, const int32_t *edofs,
                        const int64_t *r2e_ptr, const int32_t *r2e_idx, mwx_asm_plan_t **out, void *stream);
void mwx_asm_plan_destroy(mwx_asm_plan_t *plan);
/* info4 = {tiles, rows per tile, bytes of the entry blobs, bytes of the geometry blobs}. */
int mwx_asm_plan_info(const mwx_asm_plan_t *plan, int64_t *info4);
/* values[dst] = ca * (hess u : hess v) + cb * (grad u . grad v) for every destination entry of the plan
 * (ref main/example1/run.py:110-123, 210); pxy = CURRENT vertex coordinates (interleaved x, y; skfem vertex order).
 * accumulate != 0 adds into values. */
int mwx_assemble_morley_plan(const mwx_asm_plan_t *plan, const double *pxy, double ca, double cb,
                             int32_t accumulate, double *values, void *stream);

/* ------------------------------------------------------------------ condense (glue A -> B)
 * Replaces condense(K2, f2, D=...) ref main/example1/utils.py:384-460:
 *   Aout = A[I].T[I].T,  bout = b[I] - A[I].T[D].T @ x[D].
 * keep (N) int32 0/1 mask of I; newid (N+1) int32 = exclusive scan of keep (output of pass 1);
 *   indptr_out must hold N+1 entries (n_out is only known after pass 1). */
int mwx_condense_count(int64_t n, const int64_t *indptr, const int32_t *indices,
                       const int32_t *keep, int32_t *newid, int64_t *indptr_out,
                       int64_t *n_out_host, int64_t *nnz_out_host, void *stream);
int mwx_condense_fill(int64_t n, const int64_t *indptr, const int32_t *indices,
                      const double *values, const int32_t *keep, const int32_t *newid,
                      const int64_t *indptr_out, int32_t *indices_out, double *values_out,
                      const double *b, const double *x, double *b_out, void *stream);

/* ------------------------------------------------------------------ locality renumbering (internal to the fast solver)
 * No reference counterpart.  keys[i] = position of DOF location xy[i] (interleaved x,y) along a Hilbert
 * curve over the square [x0, x0+extent] x [y0, y0+extent]; sorting the keys gives the permutation. */
int mwx_hilbert_keys(int64_t n, const double *xy, double x0, double y0, double extent, int64_t *keys,
                     void *stream);
/* B = P A P^T: row i of B is row perm[i] of A, columns renumbered through iperm and re-sorted.
 * indptr_out (n+1), indices_out / values_out (nnz). */
int mwx_csr_permute(int64_t n, const int64_t *indptr, const int32_t *indices, const double *values,
                    const int32_t *perm, const int32_t *iperm, int64_t *indptr_out, int32_t *indices_out,
                    double *values_out, void *stream);
/* dst[i] = src[idx[i]]  (b -> P b, x -> P^T x). */
int mwx_gather_f64(int64_t n, const double *src, const int32_t *idx, double *dst, void *stream);

/* ------------------------------------------------------------------ SpMV / vector kernels (hot path B)
 * y = A x.  Replaces scipy csr_matvec inside spl.cg  ref main/example1/utils.py:157. */
int mwx_spmv(int64_t n, int64_t nnz, const int64_t *indptr, const int32_t *indices,
             const double *values, const double *x, double *y, void *stream);   /* nnz >= 0: tiling hint, -1 unknown, <= -8: -(tile rows) from mwx_spmv_plan */
/* y = c - A x (sign < 0) or c + A x (sign > 0); c may alias y.  Residual / prolongation of the V-cycle. */
int mwx_spmv_axpy(int64_t n, int64_t nnz, const int64_t *indptr, const int32_t *indices, const double *values,
                  const double *x, const double *c, int32_t sign, double *y, void *stream);
/* One polynomial-smoother step (Chebyshev / damped Jacobi): d = c1 d + c2 dinv (b - A x_in); x_out = x_in + d.
 * mwx_smoother_first is the step from a zero iterate: d = x = c2 dinv b (no SpMV). */
int mwx_smoother_step(int64_t n, int64_t nnz, const int64_t *indptr, const int32_t *indices,
                      const double *values, const double *x_in, double *x_out, const double *b,
                      const double *dinv, double *d, double c1, double c2, void *stream);
int mwx_smoother_first(int64_t n, const double *b, const double *dinv, double c2, double *x, double *d,
                       void *stream);
/* out = M v for a dense row-major n x n M (coarsest-level pseudo-inverse). */
int mwx_dense_gemv(int64_t n, const double *M, const double *v, double *out, void *stream);
/* SpMV plan of an operator: the tallest row tile whose every tile fits one TMA stage, plus (for dense operators with
 * 16-64 rows per tile) how many lanes share a row, packed as rows + 4096*log2(lanes).  Opaque: pass -(plan) as `nnz`. */
int mwx_spmv_plan(int64_t n, const int64_t *indptr, int32_t *tile_rows_host, void *stream);
/* diag[i] = A[i,i] (0 if absent): build_pc_diag, ref main/example1/utils.py:48-50. */
int mwx_csr_diagonal(int64_t n, const int64_t *indptr, const int32_t *indices,
                     const double *values, double *diag, void *stream);

/* Jacobi-preconditioned CG with scipy-1.4.1 `cg` stopping semantics:
 *   solver_iter_krylov(Precondition=True) / solver_iter_pcg   ref main/example1/utils.py:265-320.
 * work: 5*n doubles.  info_host[0] = 0 converged / maxiter not converged; iters_host = iterations;
 * resid_host = final ||r||_2.  precondition = 0 runs plain CG (M = I). */
int mwx_pcg_jacobi(int64_t n, const int64_t *indptr, const int32_t *indices, const double *values,
                   const double *b, double *x, double tol, int64_t maxiter, int32_t precondition,
                   double *work, int64_t *iters_host, int32_t *info_host, double *resid_host,
                   void *stream);

/* ------------------------------------------------------------------ AMG (hot path B)
 * Host setup = pyamg.ruge_stuben_solver(A) ref main/example1/utils.py:154 (classical strength 0.25,
 * first-pass RS splitting, direct interpolation, Galerkin RAP, <=10 levels, coarse <=10 rows).
 * All arrays here are HOST pointers.  Opaque hierarchy; query level sizes then copy out. */
typedef struct mwx_amg_host mwx_amg_host_t;
int mwx_amg_setup_host(int64_t n, const int64_t *indptr_host, const int32_t *indices_host,
                       const double *values_host, double theta, int32_t max_levels,
                       int32_t max_coarse, mwx_amg_host_t **out);
int mwx_amg_num_levels(const mwx_amg_host_t *h);
/* OpenMP threads of the host setup (launchers such as torchrun pin OMP_NUM_THREADS=1); returns the setting. */
/* Throughput-mode hierarchy (no reference counterpart): smoothed aggregation with `ncand` near-nullspace candidates
 * (row-major n x ncand; for the Morley stiffness the interpolants of 1, x, y).  Produces the same opaque hierarchy as
 * mwx_amg_setup_host, so upload / V-cycle / PCG are shared; theta = symmetric strength threshold (0.1, the same on every level). */
int mwx_amg_setup_sa_host(int64_t n, const int64_t *indptr, const int32_t *indices, const double *values,
                          const double *candidates, int32_t ncand, int32_t block_size, int32_t first_level,
                          double theta, int32_t max_levels, int32_t max_coarse, mwx_amg_host_t **out);
/* block_size / first_level: 1 / 0 for a fine-level matrix.  To continue a hierarchy from one of its own coarse levels
 * l (the partitioned solver replicates the small levels on every rank) pass that level's matrix, its candidates
 * (mwx_amg_level_candidates), block_size = ncand and first_level = l: the remaining levels come out identical. */
int mwx_amg_level_candidates(const mwx_amg_host_t *h, int32_t level, double *candidates, int32_t *ncand);
int mwx_set_host_threads(int32_t n);
/* which: 0 = A_l, 1 = P_l (n_l x n_{l+1}), 2 = R_l (n_{l+1} x n_l).  dims_host = {rows, cols, nnz}. */
int mwx_amg_level_dims(const mwx_amg_host_t *h, int32_t level, int32_t which, int64_t *dims_host);
int mwx_amg_level_copy(const mwx_amg_host_t *h, int32_t level, int32_t which, int64_t *indptr_host,
                       int32_t *indices_host, double *values_host);
/* Borrowed pointers into the host hierarchy (valid until mwx_amg_destroy_host): no staging copy. */
int mwx_amg_level_view(const mwx_amg_host_t *h, int32_t level, int32_t which, const int64_t **indptr,
                       const int32_t **indices, const double **values);
/* C/F splitting of level `level` (1 = C) for tests. */
/* Smoothed-aggregation levels: coarse DOFs coarse_per_root*a .. +coarse_per_root-1 belong to aggregate a, whose seed
 * DOF is roots[a] (ascending; P.cols / coarse_per_root entries).  coarse_per_root = 0 on a Ruge-Stueben level. */
int mwx_amg_level_roots(const mwx_amg_host_t *h, int32_t level, int64_t *roots, int32_t *coarse_per_root);
int mwx_amg_level_splitting(const mwx_amg_host_t *h, int32_t level, int32_t *split_host);
void mwx_amg_destroy_host(mwx_amg_host_t *h);

/* Dependency levels of the lexicographic Gauss-Seidel sweep of a CSR matrix (host): row i is in
 * wavefront 1 + max(wavefront(j) : j < i, a_ij != 0).  order_host (n) lists rows grouped by
 * wavefront, wave_ptr_host (nwaves+1).  Returns nwaves through nwaves_host. */
int mwx_gs_wavefronts_host(int64_t n, const int64_t *indptr_host, const int32_t *indices_host,
                           int32_t *order_host, int64_t *wave_ptr_host, int64_t *nwaves_host);
/* Same for the backward sweep (row i waits for its neighbours j > i). */
int mwx_gs_wavefronts_backward_host(int64_t n, const int64_t *indptr_host, const int32_t *indices_host,
                                    int32_t *order_host, int64_t *wave_ptr_host, int64_t *nwaves_host);

/* Device V-cycle preconditioner (ml.aspreconditioner(), ref main/example1/utils.py:155). */
typedef struct mwx_amg_dev mwx_amg_dev_t;
/* smoother: 0 = exact lexicographic symmetric Gauss-Seidel (parity mode, wavefront-scheduled),
 *           1 = Chebyshev(degree) on D^-1 A (fast mode), 2 = damped Jacobi x2 (fast mode). */
/* coarse_pinv_host: dense row-major pseudo-inverse of the coarsest A (scipy pinv2 in the reference);
 * NULL = computed inside by a symmetric Jacobi eigen-solver with pinv2's cutoff. */
int mwx_amg_upload(const mwx_amg_host_t *h, int32_t smoother, int32_t cheby_degree,
                   const double *coarse_pinv_host, mwx_amg_dev_t **out, void *stream);
/* Smoothed-aggregation setup ON THE DEVICE (throughput mode; no reference counterpart - the reference's
 * ``pyamg.ruge_stuben_solver(A)``, ref main/example1/utils.py:154, runs per solve and is what mode 'sa' replaces).
 * Same scheme and parameters as mwx_amg_setup_sa_host; aggregate roots are a maximal independent set of the squared
 * strength graph instead of the host's serial greedy sweep, every sparse product is a deterministic hash SpGEMM
 * (csrc/amg_setup_dev.cu).  All arrays are DEVICE pointers; indptr/indices/values are BORROWED as the fine-level
 * operator and must outlive the hierarchy.  block_size / first_level as in mwx_amg_setup_sa_host (continuing from
 * one of the scheme's own coarse levels reproduces the remaining levels bit for bit).  smoother 1 (Chebyshev) or
 * 2 (Jacobi).  The result needs
 * mwx_amg_dev_finalize (work vectors, SpMV plans, coarse pseudo-inverse) before mwx_amg_vcycle / mwx_pcg_amg.
 * MWX_E_CAPACITY: a coarse operator lost its ncand x ncand block structure or a product row exceeds 8192 entries. */
int mwx_amg_setup_sa_dev(int64_t n, const int64_t *indptr, const int32_t *indices, const double *values,
                         const double *candidates, int32_t ncand, int32_t block_size, int32_t first_level,
                         double theta, int32_t max_levels, int32_t max_coarse, int32_t smoother,
                         int32_t cheby_degree, mwx_amg_dev_t **out, void *stream);
/* coarse_pinv_host as in mwx_amg_upload (NULL = computed inside, coarsest level of at most 4096 rows). */
int mwx_amg_dev_finalize(mwx_amg_dev_t *d, const double *coarse_pinv_host, void *stream);
int mwx_amg_dev_num_levels(const mwx_amg_dev_t *d);
/* which / dims as mwx_amg_level_dims; the copy goes to host OR device buffers (cudaMemcpyDefault). */
int mwx_amg_dev_level_dims(const mwx_amg_dev_t *d, int32_t level, int32_t which, int64_t *dims_host);
int mwx_amg_dev_level_copy(const mwx_amg_dev_t *d, int32_t level, int32_t which, int64_t *indptr, int32_t *indices,
                           double *values, void *stream);
/* Aggregate roots (ascending first DOF of every aggregate's root node), coarse DOFs per root and the near-nullspace
 * candidates (rows x ncand) of a device-built level; any output pointer may be NULL; roots/candidates host or device. */
int mwx_amg_dev_level_aux(const mwx_amg_dev_t *d, int32_t level, int64_t *roots, int64_t *nroots_host,
                          int32_t *coarse_per_root_host, double *candidates, int32_t *ncand_host, void *stream);
/* Chebyshev smoother: polynomial degree and eigenvalue interval [1.1 rho / ratio, 1.1 rho] (default 2, 10; tuned on the
 * level-11/12 meshes, profiles/r01_tune_amg.txt). */
int mwx_amg_set_chebyshev(mwx_amg_dev_t *d, int32_t degree, double ratio);
/* Cycle shape.  gamma = 1: V(1,1) (pyamg's default, what the reference uses).  gamma = 2 (fast modes only; no
 * reference counterpart): levels first_level .. last_level (>= 1; the coarsest level is a direct solve and never
 * revisited) are visited twice per visit of their parent - the second visit solves for the residual of the first, so
 * the cycle is still a symmetric positive definite preconditioner for CG. */
int mwx_amg_set_cycle(mwx_amg_dev_t *d, int32_t gamma, int32_t first_level, int32_t last_level);
/* Estimated spectral radius of D^-1 A on a level (fast-mode smoothers; 0 in parity mode). */
int mwx_amg_level_rho(const mwx_amg_dev_t *d, int32_t level, double *rho_host);
void mwx_amg_destroy_dev(mwx_amg_dev_t *d);
/* z = M^-1 r : one V(1,1) cycle from a zero initial guess. */
int mwx_amg_vcycle(mwx_amg_dev_t *d, const double *r, double *z, void *stream);
/* AMG-preconditioned CG, scipy-1.4.1 stopping semantics: solver_iter_mgcg ref utils.py:127-166. */
int mwx_pcg_amg(mwx_amg_dev_t *d, int64_t n, const int64_t *indptr, const int32_t *indices,
                const double *values, const double *b, double *x, double tol, int64_t maxiter,
                double *work, int64_t *iters_host, int32_t *info_host, double *resid_host,
                void *stream);

/* Same solve with two diagnostics that scipy's cg does not have (both off = mwx_pcg_amg): *first_hit_host = the first
 * iteration whose recurrence residual met the tolerance; max_rechecks > 0 stops after that many FAILED true-residual
 * re-checks, or 10 iterations after the first one (info = iteration count, *resid_host = the true residual), instead
 * of iterating to maxiter - on very fine
 * meshes eps_mach ||A|| ||x|| exceeds tol ||b|| and the scipy 1.4.1 rule can never be met in fp64. */
int mwx_pcg_amg_monitored(mwx_amg_dev_t *d, int64_t n, const int64_t *indptr, const int32_t *indices,
                          const double *values, const double *b, double *x, double tol, int64_t maxiter,
                          int32_t max_rechecks, double *work, int64_t *iters_host, int32_t *info_host,
                          double *resid_host, int64_t *first_hit_host, void *stream);

/* ``verbose=True`` of the reference's solver factories (ref main/example1/utils.py:104-110, 146-152, 229-235, 282-288):
 * scipy calls ``callback(x)`` after every iteration and the reference prints ``np.linalg.norm(x)``.  Same solves as
 * mwx_pcg_amg / mwx_pcg_jacobi; xnorm_hist_host[k-1] = ||x_k||_2 for the first xnorm_cap iterations (one extra dot
 * per iteration, read back with the residual), printed by the Python closure. */
int mwx_pcg_amg_history(mwx_amg_dev_t *d, int64_t n, const int64_t *indptr, const int32_t *indices,
                        const double *values, const double *b, double *x, double tol, int64_t maxiter,
                        double *work, int64_t *iters_host, int32_t *info_host, double *resid_host,
                        double *xnorm_hist_host, int64_t xnorm_cap, void *stream);
int mwx_pcg_jacobi_history(int64_t n, const int64_t *indptr, const int32_t *indices, const double *values,
                           const double *b, double *x, double tol, int64_t maxiter, int32_t precondition,
                           double *work, int64_t *iters_host, int32_t *info_host, double *resid_host,
                           double *xnorm_hist_host, int64_t xnorm_cap, void *stream);

/* ------------------------------------------------------------------ callers either side of the hot path
 * K1 = asm(laplace, basis['w']) for P1 (ndpe 3) / P2 (ndpe 6), ref main/example1/run.py:149-154,198 (closed form). */
int mwx_assemble_lagrange_laplace(int64_t nrows, int64_t nt, int32_t ndpe, const double *pxy, const int32_t *t,
                                  const int64_t *r2e_ptr, const int32_t *r2e_idx, const uint8_t *slot8,
                                  const int64_t *indptr, double *values, void *stream);
/* f1 = asm(f_load, basis['w']), ref run.py:277-288,199: fq (nt*nq) = the load sampled at the quadrature points
 * (evaluated by the caller: problem data stays host Python); xw = [xi (nq) | eta (nq) | weights (nq)]. */
int mwx_load_vector(int64_t nrows, int64_t nt, int32_t ndpe, int32_t nq, const double *xw, const double *pxy,
                    const int32_t *t, const int64_t *r2e_ptr, const int32_t *r2e_idx, const double *fq,
                    double *out, void *stream);
/* Same, with the load of a NAMED problem evaluated on the device instead of host samples: problem 1 = 'ex1'
 * (f = eps^2 lap^2 u - lap u, u = (sin pi x sin pi y)^2, ref main/example1/run.py:275-288), problem 3 = 'ex3'
 * (f = 2 pi^2 sin pi x sin pi y, ref run.py:466-472).  For meshes where nt*nq host samples are not an option. */
int mwx_load_vector_problem(int32_t problem, double eps, int64_t nrows, int64_t nt, int32_t ndpe, int32_t nq,
                            const double *xw, const double *pxy, const int32_t *t, const int64_t *r2e_ptr,
                            const int32_t *r2e_idx, double *out, void *stream);
/* f2 = asm(wv_load, basis['w'], basis['u']) * wh, ref run.py:126-131,211, applied without storing the rectangular
 * matrix: out[r] = sum_K int grad(w_h) . grad(phi_r) over the Morley rows (r2e = Morley incidence). */
int mwx_wv_action(int64_t nrows, int64_t nt, int32_t ndpe_w, const double *pxy, const int32_t *t,
                  const int32_t *wdofs, const int64_t *r2e_ptr, const int32_t *r2e_idx, const double *wh,
                  double *out, void *stream);
/* L2uError / get_DuError / get_D2uError, ref run.py:157-179: sums3 = {int (u_h-u)^2, int |grad(u_h-u)|^2,
 * int |hess(u_h-u)|_F^2} of a Morley function; exact6 = u, ux, uy, uxx, uxy, uyy sampled at the quadrature points. */
int mwx_morley_error_sums(int64_t nt, int32_t nq, const double *xw, const double *pxy, const int32_t *t,
                          const int32_t *edofs, const double *uh, const double *exact6, double *sums3,
                          void *stream);
/* Same against the exact solution of a named problem (1 = 'ex1', 3 = 'ex3'; ref run.py:290-309, 474-493) evaluated
 * on the device at the quadrature points. */
int mwx_morley_error_sums_problem(int32_t problem, int64_t nt, int32_t nq, const double *xw, const double *pxy,
                                  const int32_t *t, const int32_t *edofs, const double *uh, double *sums3,
                                  void *stream);

/* main/example2 (ref main/example2/run.py:101-172): samples (u, ux, uy, uxx, uxy, uyy) of a Morley function of a
 * coarse mesh at the quadrature points of a mesh refined uniformly from it (coarse triangle of fine triangle T is
 * T mod nt_c; replaces the per-point ``test_basis['u'].interpolator(test_uh0)(x)`` loops and the DG-P1/DG-P0
 * ``project(..., diff=)`` chain, which is exact for a piecewise quadratic).  out6 = 6 arrays of nt_f*nq, the
 * ``exact6`` layout of mwx_morley_error_sums.  MWX_E_BADARG if a point falls outside its assumed triangle. */
int mwx_morley_sample_nested(int64_t nt_f, int32_t nq, const double *xw, const double *pxy_f, const int32_t *t_f,
                             int64_t nt_c, const double *pxy_c, const int32_t *t_c, const int32_t *edofs_c,
                             const double *uh_c, double *out6, void *stream);

/* project(exact_u, basis_to=basis['u']) of example 3 (ref main/example3/run.py:66, main/example3/utils.py:494-565):
 * Morley mass matrix (quadrature, same pattern as K2) and the load vector int f phi_r (f sampled at the points). */
int mwx_assemble_morley_mass(int64_t nrows, int64_t nt, int32_t nq, const double *xw, const double *pxy,
                             const int32_t *t, const int64_t *r2e_ptr, const int32_t *r2e_idx,
                             const uint8_t *slot8, const int64_t *indptr, double *values, void *stream);
int mwx_morley_load_vector(int64_t nrows, int64_t nt, int32_t nq, const double *xw, const double *pxy,
                           const int32_t *t, const int64_t *r2e_ptr, const int32_t *r2e_idx, const double *fq,
                           double *out, void *stream);
/* L2pnvError, ref run.py:106-108,527: terms[b] = int_F (h_F n . grad u_h)^2 for each listed boundary facet. */
int mwx_morley_facet_pnv(int64_t nt, int64_t nbf, const double *pxy, const int32_t *t, const int32_t *edofs,
                         const int32_t *btri, const int32_t *blocal, const double *uh, double *terms,
                         void *stream);

/* ------------------------------------------------------------------ partitioned solve (no reference counterpart)
 * A rank owns the contiguous range [lo, hi) of the solver numbering.  gid (ndofs) maps a skfem DOF to its
 * solver id, or -1 for an eliminated (Dirichlet) DOF.  rows (nloc) = skfem ids of the owned DOFs in solver order.
 * Pass 1: raw_indptr (rows as assembled by mwx_assemble_morley_rows) and loc_indptr (Dirichlet columns dropped). */
int mwx_dist_rows_count(int64_t nloc, const int32_t *rows, const int64_t *indptr, const int32_t *indices,
                        const int32_t *gid, int64_t *raw_indptr, int64_t *loc_indptr, void *stream);
/* Pass 2: cols_g = solver ids (ascending per row), src = position of each kept value in the raw rows. */
int mwx_dist_rows_fill(int64_t nloc, const int32_t *rows, const int64_t *indptr, const int32_t *indices,
                       const int32_t *gid, const int64_t *raw_indptr, const int64_t *loc_indptr,
                       int32_t *cols_g, int64_t *src, void *stream);
/* Local column ids: owned g -> g - lo, otherwise nloc + rank of g in the sorted ghost list. */
int mwx_dist_localize(int64_t nnz, const int32_t *cols_g, int32_t lo, int32_t hi, const int32_t *ghosts,
                      int32_t nghost, int32_t nloc, int32_t *cols_l, void *stream);
int mwx_gather_f64_i64(int64_t n, const double *src, const int64_t *idx, double *dst, void *stream);
/* Halo push over peer memory: dst[s][i - seg[s]] = v[idx[i]] for seg[s] <= i < seg[s+1]; dst[s] = the neighbour's
 * ghost segment (a peer pointer), all arrays device-resident.  One launch per exchange whatever the neighbour count. */
int mwx_halo_push(int64_t total, int32_t nseg, const int64_t *seg, const int32_t *idx, const double *v,
                  double *const *dst, void *stream);

/* fp32-STORED hierarchy (mode 'sa', throughput; no reference counterpart).  The operators a V-cycle applies are read
 * once per application and dominate its HBM traffic; storing their VALUES in single precision - vectors, products and
 * row sums stay fp64 - halves those bytes for the block-CSR operators (8.4 -> 4.4 B per nonzero).  Scalar CSR operators
 * (the fine matrix) keep fp64: their kernel is bound by the x gathers, and the packed (fp32 value, int32 column) record
 * layout below (12 -> 8 B per nonzero) measured slower - it is built only under MWX_PACKED_CSR=1.  The preconditioner
 * is exactly the V-cycle of the rounded hierarchy: a fixed symmetric linear operator, so CG's theory holds; iteration
 * counts are unchanged at 67 M DOFs.  The system matrix CG multiplies with is never touched.  bits: 32 or 64 (no-op);
 * parity mode ignores the call. */
int mwx_amg_set_storage(mwx_amg_dev_t *d, int32_t bits, void *stream);
/* The same for operators held outside a hierarchy (row blocks of the partitioned V-cycle): packed = nnz 8-byte records
 * (16-byte aligned); the _pk kernels are mwx_spmv / mwx_spmv_axpy / mwx_smoother_step reading those records. */
int mwx_csr_pack(int64_t nnz, const int32_t *indices, const double *values, void *packed, void *stream);
int mwx_spmv_pk(int64_t n, int64_t nnz_hint, const int64_t *indptr, const void *packed, const double *x, double *y,
                void *stream);
int mwx_spmv_axpy_pk(int64_t n, int64_t nnz_hint, const int64_t *indptr, const void *packed, const double *x,
                     const double *c, int32_t sign, double *y, void *stream);
int mwx_smoother_step_pk(int64_t n, int64_t nnz_hint, const int64_t *indptr, const void *packed, const double *x_in,
                         double *x_out, const double *b, const double *dinv, double *d, double c1, double c2,
                         void *stream);

/* r = b - A x, every row accumulated in double-double and rounded once (no reference counterpart: the reference never ran
 * meshes where it matters).  Used by the optional iterative refinement of `pcg(..., refine=k)`: with condition numbers
 * ~1e10 the fp64 residual's own rounding error exceeds tol * ||b||, and the forward error stalls at ~1e-6. */
int mwx_residual_dd(int64_t n, const int64_t *indptr, const int32_t *indices, const double *values, const double *x,
                    const double *b, double *r, void *stream);

/* Tile-local column plan of a CSR pattern (csrc/spmv_lx.cu; no reference counterpart).  The SpMV kernel is bound by the
 * L1 wavefronts of its x gathers (11.5 scattered columns per Morley row), not by DRAM; the rows of one tile share their
 * columns, so the plan lists each tile's DISTINCT columns once (sorted) and gives every nonzero a 2-byte index into that
 * list: the kernel gathers each distinct x entry once into shared memory.  MEASURED SLOWER than the plain kernel on
 * B200 (3.15 against 2.73 ms per smoother step at 67 M rows: fewer wavefronts and bytes, but one more barrier per tile),
 * so nothing uses it unless MWX_SPMV_LX=1; results are bit-identical.  plan = the (negative) value mwx_spmv_plan
 * returned for this matrix.  *out = NULL (status 0) when some tile has more than 768 distinct columns: keep the plain
 * entry points.  The _lx entry points are mwx_spmv / mwx_spmv_axpy / mwx_smoother_step / mwx_dcg_spmv_dot with the plan. */
int mwx_spmv_lx_create(int64_t n, const int64_t *indptr, const int32_t *indices, int64_t plan, void **out, void *stream);
void mwx_spmv_lx_destroy(void *lx);
int mwx_spmv_lx(int64_t n, int64_t plan, const int64_t *indptr, const int32_t *indices, const double *values,
                const void *lx, const double *x, double *y, void *stream);
int mwx_spmv_axpy_lx(int64_t n, int64_t plan, const int64_t *indptr, const int32_t *indices, const double *values,
                     const void *lx, const double *x, const double *c, int32_t sign, double *y, void *stream);
int mwx_smoother_step_lx(int64_t n, int64_t plan, const int64_t *indptr, const int32_t *indices, const double *values,
                         const void *lx, const double *x_in, double *x_out, const double *b, const double *dinv,
                         double *d, double c1, double c2, void *stream);
int mwx_dcg_spmv_dot_lx(int64_t n, int64_t plan, const int64_t *indptr, const int32_t *indices, const double *values,
                        const void *lx, const double *x_ext, double *q, double *ws, void *stream);

/* Block-CSR twin of a CSR operator whose nonzeros come in br x bc blocks (3x3, 1x3, 3x1: the coarse operators of a
 * smoothed-aggregation hierarchy on three candidates); *out = NULL (status 0) when the operator is not blocked that way.
 * Used for the row blocks of the partitioned V-cycle; inside a single-GPU hierarchy the library builds them itself.
 * storage_bits = 32 keeps the value planes in fp32 (see mwx_amg_set_storage), 64 in fp64. */
int mwx_bsr_create(int64_t rows, int64_t cols, int64_t nnz, const int64_t *indptr, const int32_t *indices,
                   const double *values, int32_t br, int32_t bc, int32_t storage_bits, void **out, void *stream);
void mwx_bsr_destroy(void *bsr);
int mwx_bsr_spmv(const void *bsr, const double *x, double *y, void *stream);                       /* y = A x */
int mwx_bsr_spmv_axpy(const void *bsr, const double *x, const double *c, int32_t sign, double *y,
                      void *stream);                                                               /* y = c +- A x */
int mwx_bsr_smoother_step(const void *bsr, const double *x_in, double *x_out, const double *rhs, const double *dinv,
                          double *d, double c1, double c2, void *stream);   /* as mwx_smoother_step */

/* Peer-memory collectives of the partitioned iteration (one process per GPU; NVLink / NVSwitch peer mappings).
 * ctl: this rank's control block (mwx_peer_ctl_bytes() bytes, zero-initialised, in memory every peer has mapped);
 * peers: device array of `world` pointers to the ranks' control blocks.  The barrier epoch is device-resident, so
 * these launches can be captured in a CUDA graph and replayed (the whole PCG iteration is one graph launch).
 * wait = 1: full barrier; 0: signal only (all parts in one process on one stream - tests). */
int mwx_peer_ctl_bytes(void);
int mwx_peer_barrier(void *ctl, const void *peers, int32_t rank, int32_t world, int32_t wait, void *stream);
/* Halo push / all-gather fused with the barrier: dst[s][i - seg[s]] = v[idx ? idx[i] : i - seg[s]], then the last
 * block to finish signals the peers and waits for them; when the launch retires every rank's values have landed. */
int mwx_peer_push(int64_t total, int32_t nseg, const int64_t *seg, const int32_t *idx, const double *v,
                  double *const *dst, void *ctl, const void *peers, int32_t rank, int32_t world, int32_t wait,
                  void *stream);
/* vals[0..count) (count <= 8) <- sum over the ranks, added in rank order (bitwise identical on every rank).
 * phase 0: the whole all-reduce; 1 / 2: its write / sum half (in-process parts). */
int mwx_peer_allreduce(double *vals, int32_t count, void *ctl, const void *peers, int32_t rank, int32_t world,
                       int32_t phase, void *stream);

/* CG building blocks chained by the multi-GPU driver between halo exchanges and all-reduces.  ws: zero-initialised
 * workspace of mwx_dcg_workspace_doubles() doubles; ws[0..4] = {rho, rho_prev, pq, rr, aux}.  Kernels leave LOCAL
 * partial sums there (deterministic per rank); the driver all-reduces them.  x_ext = [owned | ghost] vector. */
int mwx_dcg_workspace_doubles(void);
int mwx_dcg_dot(int64_t n, const double *a, const double *b, double *ws, int32_t slot, void *stream);
int mwx_dcg_p_update(int64_t n, const double *r, const double *dinv, const double *z, double *p, double *ws,
                     int32_t first, void *stream);               /* p = z + (rho/rho_prev) p;  z NULL -> dinv .* r */
int mwx_dcg_spmv_dot(int64_t n, int64_t nnz_hint, const int64_t *indptr, const int32_t *indices,
                     const double *values, const double *x_ext, double *q, double *ws, void *stream);  /* pq */
int mwx_dcg_xr_update(int64_t n, const double *p, const double *q, double *x, double *r, const double *dinv,
                      double *ws, int32_t jacobi, void *stream);   /* alpha = rho/pq; rr (and rho if jacobi) */
int mwx_dcg_residual(int64_t n, const int64_t *indptr, const int32_t *indices, const double *values,
                     const double *x_ext, const double *b, double *r, const double *dinv, double *ws,
                     void *stream);                                /* r = b - A x; rr, rho */

#ifdef __cplusplus
}
#endif
#endif /* MWX_B200_H */
