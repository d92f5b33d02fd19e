// Device building blocks of hot path B shared by the Jacobi- and AMG-preconditioned CG drivers.
#pragma once
#include "common.cuh"

namespace mwx {

// Scalars of one CG solve, resident in device memory so that kernels chain without host round trips.
struct CgScalars {
    double rho;        // r . z of the current iteration
    double rho_prev;   // r . z of the previous iteration
    double pq;         // p . A p
    double rr;         // r . r
    double aux;        // scratch dot
    unsigned int ticket[4];   // "last block done" counters, one per reduction site
};

constexpr int RED_MAX_BLOCKS = 4096;   // partial-sum slots per reduction site

// Tile-local column plan of a CSR pattern (spmv_lx.cu): per tile of `rows_per_tile` rows the sorted distinct columns
// (ucols[uptr[t] .. uptr[t+1]), padded to a multiple of 4) and, per nonzero, its 2-byte position in that list.
constexpr int LX_MAX_U = 768;
struct SpmvLX {
    int rows_per_tile = 0, plan = 0;
    int64_t ntiles = 0, n = 0, nnz = 0, nucols = 0;
    int32_t *uptr = nullptr;
    int32_t *ucols = nullptr;
    uint16_t *lidx = nullptr;
};
// plan = the value spmv_plan_rows returned for the matrix.  *out stays nullptr (status 0) when the pattern does not
// localise (a tile with more than LX_MAX_U distinct columns).
int spmv_lx_create(int64_t n, const int64_t *indptr, const int32_t *indices, int plan, SpmvLX **out, cudaStream_t st);
void spmv_lx_destroy(SpmvLX *p);

// `nnz` in the SpMV launchers sizes the row tiles: >= 0 = nonzero count (heuristic), -1 = unknown,
// <= -8 = -(tile height) from spmv_plan_rows (exact: the tallest tile that always fits a TMA stage).
int spmv_plan_rows(int64_t n, const int64_t *indptr, int *rows_out, cudaStream_t st);
// y = A x (CSR, int64 indptr, int32 cols, fp64).  If dot_with != nullptr also accumulates
// sum_i dot_with[i] * y[i] into *dot_out deterministically (per-CTA partials, last CTA folds them in
// index order).  partials: RED_MAX_BLOCKS doubles, ticket: zero-initialised counter.
int launch_spmv(int64_t n, int64_t nnz, const int64_t *indptr, const int32_t *indices, const double *values,
                const double *x, double *y, const double *dot_with, double *dot_out, double *partials,
                unsigned int *ticket, cudaStream_t st, const SpmvLX *lx = nullptr);

// y = c - A x (sign < 0) or y = c + A x (sign > 0); c may alias y.
int launch_spmv_axpy(int64_t n, int64_t nnz, const int64_t *indptr, const int32_t *indices, const double *values,
                     const double *x, double *y, const double *c, int sign, cudaStream_t st, const SpmvLX *lx = nullptr);

// One polynomial-smoother step: d = c1 d + c2 dinv (b - A x_in);  x_out = x_in + d  (x_out != x_in).
int launch_smoother_step(int64_t n, int64_t nnz, const int64_t *indptr, const int32_t *indices, const double *values,
                         const double *x_in, double *x_out, const double *b, const double *dinv, double *d,
                         double c1, double c2, cudaStream_t st, const SpmvLX *lx = nullptr);

// r = b - A x, and (deterministic) rr = r.r, rho = sum dinv_i r_i^2 (dinv may be nullptr -> rho = rr).
int launch_residual(int64_t n, const int64_t *indptr, const int32_t *indices, const double *values,
                    const double *x, const double *b, double *r, const double *dinv, CgScalars *sc,
                    double *partials, cudaStream_t st);

// fp32-STORED operator (the AMG hierarchy of mode 'sa'): one 8-byte record per nonzero instead of 8 + 4 bytes in two
// arrays.  Only the storage is single precision - x, y, the products and the row sums are fp64 - so the operator applied
// is exactly the linear operator of the rounded matrix (symmetric where the fp64 one is).
struct PackedNz { float v; int32_t c; };
int csr_pack(int64_t nnz, const int32_t *indices, const double *values, PackedNz *out, cudaStream_t st);
int launch_spmv_pk(int64_t n, int64_t nnz, const int64_t *indptr, const PackedNz *pk, const double *x, double *y,
                   cudaStream_t st);
int launch_spmv_axpy_pk(int64_t n, int64_t nnz, const int64_t *indptr, const PackedNz *pk, const double *x, double *y,
                        const double *c, int sign, cudaStream_t st);
int launch_smoother_step_pk(int64_t n, int64_t nnz, const int64_t *indptr, const PackedNz *pk, const double *x_in,
                            double *x_out, const double *b, const double *dinv, double *d, double c1, double c2,
                            cudaStream_t st);

// diag[i] = A[i,i] (invert != 0: its reciprocal).
int launch_csr_diagonal(int64_t n, const int64_t *indptr, const int32_t *indices, const double *values,
                        double *diag, int invert, cudaStream_t st);

// out = sum a_i b_i (deterministic).
int launch_dot(int64_t n, const double *a, const double *b, double *out, double *partials,
               unsigned int *ticket, cudaStream_t st);

// z = M^-1 r launcher used by cg_solve (nullptr = Jacobi / identity handled inside the vector kernels).
typedef int (*PrecondFn)(void *ctx, const double *r, double *z, cudaStream_t st);

// Preconditioned CG with scipy 1.4.1 `cg` semantics.  r, z, p, q: n-vectors of workspace.
int cg_solve(int64_t n, const int64_t *indptr, const int32_t *indices, const double *values,
             const double *b, double *x, double tol, int64_t maxiter, const double *dinv,
             PrecondFn precond, void *ctx, double *r, double *z, double *p, double *q,
             int64_t *iters_host, int32_t *info_host, double *resid_host, cudaStream_t st, int max_rechecks = 0,
             int64_t *first_hit_host = nullptr, double *xnorm_hist_host = nullptr, int64_t xnorm_cap = 0,
             const SpmvLX *lx = nullptr);
// xnorm_hist_host[k - 1] = ||x_k|| for k <= xnorm_cap: what scipy's `callback(x)` of the reference's verbose solvers
// prints (ref main/example1/utils.py:104-106).

}  // namespace mwx
