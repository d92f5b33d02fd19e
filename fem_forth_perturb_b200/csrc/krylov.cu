// Hot path B: CSR SpMV, fused vector kernels and the Jacobi-preconditioned CG driver.
//
// Replaces scipy.sparse.linalg.cg (+ scipy csr_matvec, BLAS ddot/daxpy/dnrm2) as called by
//   solver_iter_krylov / solver_iter_pcg   ref main/example1/utils.py:265-320   (M = build_pc_diag, :48-50)
// with scipy 1.4.1's stopping semantics (restated in oracle/scipy_cg.py).
//
// SpMV: the Morley stiffness has short rows (9 for a facet row, 19 for an interior vertex row, 11.5
// on average), so a warp per row would idle 2/3 of its lanes.  Instead a CTA streams the contiguous
// values/indices of a 128-row tile with fully coalesced loads, multiplies by gathered x into shared
// memory, and then one thread per row adds its products in CSR order (bank-conflict free because the
// row lengths are odd).  CTAs are persistent (grid = SMs x residency) and stride over the tiles, which
// also makes the fused p.Ap reduction deterministic: one partial per CTA, folded in index order by
// the last CTA to finish.
//
// Algorithmic bytes per SpMV: 12*nnz + 20*n  (values 8 + column 4 per nonzero;
// indptr 4, x 8, y 8 per row - SURVEY 8d; this build reads an int64 indptr, i.e. 4 B/row more).
#include "krylov.cuh"

namespace mwx {
namespace {

constexpr int SPMV_ROWS = 128;
constexpr int SPMV_CAP = SPMV_ROWS * 24;        // 24 KB of fp64 products per CTA
constexpr int VEC_THREADS = 256;

// Fold per-CTA partial sums: the last CTA to arrive adds them in index order.  Returns true in the
// last CTA (all its threads), with the total valid in thread 0.
__device__ __forceinline__ bool fold_partials(double v_thread0, double *partials, unsigned int *ticket,
                                              double *sh, double &total) {
    __shared__ bool is_last;
    __syncthreads();                      // protects is_last / sh when several folds share a kernel
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = v_thread0;
        __threadfence();
        const unsigned int tk = atomicInc(ticket, gridDim.x - 1);   // wraps back to 0: self-resetting
        is_last = (tk == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return false;
    __threadfence();
    double s = 0.0;
    for (unsigned int i = threadIdx.x; i < gridDim.x; i += blockDim.x) s += __ldcg(partials + i);
    total = block_sum(s, sh);
    return true;
}

// MODE 0: y = A x   1: y = c - A x   2: y = c + A x      (c = `cvec`, may alias y for MODE 2)
// MODE 3 (smoother step): d = c1 d + c2 dinv (c - A x);  y = x + d     (y must not alias x)
struct SmootherArgs { const double *dinv; double *d; double c1, c2; };
template <bool DOT, int MODE>
__global__ void __launch_bounds__(SPMV_ROWS) k_spmv(int64_t n, const int64_t *__restrict__ indptr,
                                                    const int32_t *__restrict__ indices,
                                                    const double *__restrict__ values,
                                                    const double *__restrict__ x, double *y,
                                                    const double *cvec,
                                                    const double *__restrict__ dot_with,
                                                    double *__restrict__ dot_out, double *partials,
                                                    unsigned int *ticket, SmootherArgs sm) {
    __shared__ double prod[SPMV_CAP];
    __shared__ double red[32];
    const int64_t ntiles = (n + SPMV_ROWS - 1) / SPMV_ROWS;
    double dacc = 0.0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t r0 = tile * SPMV_ROWS;
        const int64_t r1 = r0 + SPMV_ROWS < n ? r0 + SPMV_ROWS : n;
        const int64_t r = r0 + threadIdx.x;
        const int64_t s0 = indptr[r0];
        const int len = (int)min((int64_t)(SPMV_CAP + 1), indptr[r1] - s0);
        int64_t a = 0, b = 0;
        if (r < r1) { a = indptr[r]; b = indptr[r + 1]; }
        double sum = 0.0;
        if (len <= SPMV_CAP) {
            const int32_t *ci = indices + s0;
            const double *va = values + s0;
#pragma unroll 4
            for (int k = threadIdx.x; k < len; k += SPMV_ROWS) prod[k] = va[k] * __ldg(x + ci[k]);
            __syncthreads();
            for (int k = (int)(a - s0); k < (int)(b - s0); ++k) sum += prod[k];
            __syncthreads();
        } else {
            for (int64_t k = a; k < b; ++k) sum += values[k] * __ldg(x + indices[k]);
        }
        if (r < r1) {
            if (MODE == 1) sum = cvec[r] - sum;
            if (MODE == 2) sum = cvec[r] + sum;
            if (MODE == 3) {
                const double dn = sm.c1 * sm.d[r] + sm.c2 * sm.dinv[r] * (cvec[r] - sum);
                sm.d[r] = dn;
                sum = x[r] + dn;
            }
            y[r] = sum;
            if (DOT) dacc += dot_with[r] * sum;
        }
    }
    if (DOT) {
        const double bs = block_sum(dacc, red);
        double total;
        if (fold_partials(bs, partials, ticket, red, total) && threadIdx.x == 0) *dot_out = total;
    }
}

// r = b - A x with fused rr = r.r and rho = sum dinv r^2.
__global__ void __launch_bounds__(SPMV_ROWS) k_residual(int64_t n, const int64_t *__restrict__ indptr,
                                                        const int32_t *__restrict__ indices,
                                                        const double *__restrict__ values,
                                                        const double *__restrict__ x,
                                                        const double *__restrict__ b, double *__restrict__ r_out,
                                                        const double *__restrict__ dinv, CgScalars *sc,
                                                        double *partials) {
    __shared__ double red[32];
    double rr = 0.0, rho = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride) {
        double sum = 0.0;
        for (int64_t k = indptr[r]; k < indptr[r + 1]; ++k) sum += values[k] * __ldg(x + indices[k]);
        const double res = b[r] - sum;
        r_out[r] = res;
        rr += res * res;
        rho += (dinv ? dinv[r] : 1.0) * res * res;
    }
    const double s1 = block_sum(rr, red);
    const double s2 = block_sum(rho, red);
    double t1, t2;
    // two reduction sites share the launch: fold them one after the other
    const bool last1 = fold_partials(s1, partials, &sc->ticket[2], red, t1);
    const bool last2 = fold_partials(s2, partials + RED_MAX_BLOCKS, &sc->ticket[3], red, t2);
    if (threadIdx.x == 0) {
        if (last1) sc->rr = t1;
        if (last2) sc->rho = t2;
    }
}

__global__ void __launch_bounds__(VEC_THREADS) k_dot(int64_t n, const double *__restrict__ a,
                                                     const double *__restrict__ b, double *out,
                                                     double *partials, unsigned int *ticket) {
    __shared__ double red[32];
    double s = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) s += a[i] * b[i];
    const double bs = block_sum(s, red);
    double total;
    if (fold_partials(bs, partials, ticket, red, total) && threadIdx.x == 0) *out = total;
}

// p = z + beta p with z = dinv .* r (Jacobi, z == nullptr) or an explicit z; beta = rho / rho_prev.
__global__ void __launch_bounds__(VEC_THREADS) k_p_update(int64_t n, const double *__restrict__ r,
                                                          const double *__restrict__ dinv,
                                                          const double *__restrict__ z, double *__restrict__ p,
                                                          const CgScalars *sc, int first) {
    const double beta = first ? 0.0 : sc->rho / sc->rho_prev;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double zi = z ? z[i] : (dinv ? dinv[i] * r[i] : r[i]);
        p[i] = first ? zi : zi + beta * p[i];
    }
}

// alpha = rho / pq;  x += alpha p;  r -= alpha q;  rr = r.r;  (Jacobi) rho_next = sum dinv r^2.
// The last CTA rotates the scalars: rho_prev = rho, and rho = rho_next when jacobi != 0.
__global__ void __launch_bounds__(VEC_THREADS) k_xr_update(int64_t n, const double *__restrict__ p,
                                                           const double *__restrict__ q, double *__restrict__ x,
                                                           double *__restrict__ r, const double *__restrict__ dinv,
                                                           CgScalars *sc, double *partials, int jacobi) {
    __shared__ double red[32];
    const double alpha = sc->rho / sc->pq;
    double rr = 0.0, rho = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        x[i] += alpha * p[i];
        const double ri = r[i] - alpha * q[i];
        r[i] = ri;
        rr += ri * ri;
        if (jacobi) rho += (dinv ? dinv[i] : 1.0) * ri * ri;
    }
    const double s1 = block_sum(rr, red);
    const double s2 = block_sum(rho, red);
    double t1 = 0.0, t2 = 0.0;
    const bool last1 = fold_partials(s1, partials, &sc->ticket[1], red, t1);
    bool last2 = false;
    if (jacobi) last2 = fold_partials(s2, partials + RED_MAX_BLOCKS, &sc->ticket[2], red, t2);
    if (threadIdx.x == 0) {
        if (last1) sc->rr = t1;
        if (jacobi) {
            if (last2) { sc->rho_prev = sc->rho; sc->rho = t2; }
        } else if (last1) {
            sc->rho_prev = sc->rho;
        }
    }
}

__global__ void k_csr_diagonal(int64_t n, const int64_t *__restrict__ indptr,
                               const int32_t *__restrict__ indices, const double *__restrict__ values,
                               double *__restrict__ diag, int invert) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    double d = 0.0;
    for (int64_t k = indptr[r]; k < indptr[r + 1]; ++k)
        if (indices[k] == r) d += values[k];
    diag[r] = invert ? 1.0 / d : d;
}

inline unsigned vec_grid(int64_t n) {
    int64_t g = (n + VEC_THREADS - 1) / VEC_THREADS;
    const int64_t cap = (int64_t)sm_count() * 8;          // 8 x 256 threads = full residency
    if (g > cap) g = cap;
    if (g > RED_MAX_BLOCKS) g = RED_MAX_BLOCKS;
    return (unsigned)(g < 1 ? 1 : g);
}

inline unsigned spmv_grid(int64_t n) {
    int64_t g = (n + SPMV_ROWS - 1) / SPMV_ROWS;
    const int64_t cap = (int64_t)sm_count() * 9;          // 9 CTAs x 24 KB smem per SM
    if (g > cap) g = cap;
    if (g > RED_MAX_BLOCKS) g = RED_MAX_BLOCKS;
    return (unsigned)(g < 1 ? 1 : g);
}

}  // namespace

int launch_spmv(int64_t n, const int64_t *indptr, const int32_t *indices, const double *values,
                const double *x, double *y, const double *dot_with, double *dot_out, double *partials,
                unsigned int *ticket, cudaStream_t st) {
    const unsigned g = spmv_grid(n);
    if (dot_with)
        k_spmv<true, 0><<<g, SPMV_ROWS, 0, st>>>(n, indptr, indices, values, x, y, nullptr, dot_with, dot_out, partials, ticket, SmootherArgs{});
    else
        k_spmv<false, 0><<<g, SPMV_ROWS, 0, st>>>(n, indptr, indices, values, x, y, nullptr, nullptr, nullptr, nullptr, nullptr, SmootherArgs{});
    MWX_LAUNCH_CHECK();
    return 0;
}

int launch_spmv_axpy(int64_t n, const int64_t *indptr, const int32_t *indices, const double *values,
                     const double *x, double *y, const double *c, int sign, cudaStream_t st) {
    const unsigned g = spmv_grid(n);
    if (sign < 0)
        k_spmv<false, 1><<<g, SPMV_ROWS, 0, st>>>(n, indptr, indices, values, x, y, c, nullptr, nullptr, nullptr, nullptr, SmootherArgs{});
    else
        k_spmv<false, 2><<<g, SPMV_ROWS, 0, st>>>(n, indptr, indices, values, x, y, c, nullptr, nullptr, nullptr, nullptr, SmootherArgs{});
    MWX_LAUNCH_CHECK();
    return 0;
}

int launch_smoother_step(int64_t n, const int64_t *indptr, const int32_t *indices, const double *values,
                         const double *x_in, double *x_out, const double *b, const double *dinv, double *d,
                         double c1, double c2, cudaStream_t st) {
    k_spmv<false, 3><<<spmv_grid(n), SPMV_ROWS, 0, st>>>(n, indptr, indices, values, x_in, x_out, b, nullptr,
                                                         nullptr, nullptr, nullptr, SmootherArgs{dinv, d, c1, c2});
    MWX_LAUNCH_CHECK();
    return 0;
}

int launch_residual(int64_t n, const int64_t *indptr, const int32_t *indices, const double *values,
                    const double *x, const double *b, double *r, const double *dinv, CgScalars *sc,
                    double *partials, cudaStream_t st) {
    int64_t g = (n + SPMV_ROWS - 1) / SPMV_ROWS;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (g > cap) g = cap;
    k_residual<<<(unsigned)g, SPMV_ROWS, 0, st>>>(n, indptr, indices, values, x, b, r, dinv, sc, partials);
    MWX_LAUNCH_CHECK();
    return 0;
}

int launch_dot(int64_t n, const double *a, const double *b, double *out, double *partials,
               unsigned int *ticket, cudaStream_t st) {
    k_dot<<<vec_grid(n), VEC_THREADS, 0, st>>>(n, a, b, out, partials, ticket);
    MWX_LAUNCH_CHECK();
    return 0;
}

int launch_csr_diagonal(int64_t n, const int64_t *indptr, const int32_t *indices, const double *values,
                        double *diag, int invert, cudaStream_t st) {
    k_csr_diagonal<<<grid_for(n, 256), 256, 0, st>>>(n, indptr, indices, values, diag, invert);
    MWX_LAUNCH_CHECK();
    return 0;
}

// Shared CG driver: `precond(ctx, r, z, st)` launches z = M^-1 r (nullptr -> Jacobi / identity fused
// into the vector kernels).  Follows scipy 1.4.1's cg step for step.
int cg_solve(int64_t n, const int64_t *indptr, const int32_t *indices, const double *values,
             const double *b, double *x, double tol, int64_t maxiter, const double *dinv,
             PrecondFn precond, void *ctx, double *r, double *z, double *p, double *q,
             int64_t *iters_host, int32_t *info_host, double *resid_host, cudaStream_t st) {
    Scratch<CgScalars> sc(st);
    Scratch<double> partials(st);
    MWX_CUDA(sc.alloc(1));
    MWX_CUDA(partials.alloc(2 * RED_MAX_BLOCKS));
    MWX_CUDA(cudaMemsetAsync(sc.p, 0, sizeof(CgScalars), st));
    MWX_CUDA(cudaMemsetAsync(x, 0, sizeof(double) * (size_t)n, st));
    if (maxiter <= 0) maxiter = 10 * n;
    double h[2];
    // ||b|| and r = b - A*0 = b
    int rc = launch_dot(n, b, b, &sc.p->aux, partials.p, &sc.p->ticket[0], st);
    if (rc) return rc;
    MWX_CUDA(cudaMemcpyAsync(r, b, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, st));
    MWX_CUDA(cudaMemcpyAsync(h, &sc.p->aux, sizeof(double), cudaMemcpyDeviceToHost, st));
    MWX_CUDA(cudaStreamSynchronize(st));
    const double bnrm = sqrt(h[0]);
    *iters_host = 0; *info_host = 0; *resid_host = bnrm;
    if (bnrm <= tol) return 0;                                   // legacy-atol early exit (x0 = 0)
    const double atol = bnrm != 0.0 ? tol * bnrm : tol;
    const bool jacobi = (precond == nullptr);
    if (jacobi) {
        // rho_1 = sum dinv r^2 through the residual kernel's reduction (x = 0 -> r = b)
        rc = launch_residual(n, indptr, indices, values, x, b, r, dinv, sc.p, partials.p, st);
        if (rc) return rc;
    }
    for (int64_t it = 1; it <= maxiter; ++it) {
        if (!jacobi) {
            rc = precond(ctx, r, z, st);
            if (rc) return rc;
            rc = launch_dot(n, r, z, &sc.p->rho, partials.p, &sc.p->ticket[0], st);
            if (rc) return rc;
        }
        k_p_update<<<vec_grid(n), VEC_THREADS, 0, st>>>(n, r, dinv, jacobi ? nullptr : z, p, sc.p, it == 1);
        MWX_LAUNCH_CHECK();
        rc = launch_spmv(n, indptr, indices, values, p, q, p, &sc.p->pq, partials.p, &sc.p->ticket[0], st);
        if (rc) return rc;
        k_xr_update<<<vec_grid(n), VEC_THREADS, 0, st>>>(n, p, q, x, r, dinv, sc.p, partials.p, jacobi ? 1 : 0);
        MWX_LAUNCH_CHECK();
        MWX_CUDA(cudaMemcpyAsync(h, &sc.p->rr, sizeof(double), cudaMemcpyDeviceToHost, st));
        MWX_CUDA(cudaStreamSynchronize(st));
        double rn = sqrt(h[0]);
        *iters_host = it; *resid_host = rn;
        if (!(rn == rn)) { *info_host = (int32_t)(it < 0x7fffffff ? it : 0x7fffffff); return MWX_E_NOCONV; }
        if (rn <= atol) {
            if (it == 1) return 0;
            // true-residual re-check; on failure CG continues from the recomputed r
            rc = launch_residual(n, indptr, indices, values, x, b, r, jacobi ? dinv : nullptr, sc.p, partials.p, st);
            if (rc) return rc;
            if (!jacobi) {
                // launch_residual overwrote rho with r.r; the AMG path recomputes rho after the V-cycle
            }
            MWX_CUDA(cudaMemcpyAsync(h, &sc.p->rr, sizeof(double), cudaMemcpyDeviceToHost, st));
            MWX_CUDA(cudaStreamSynchronize(st));
            rn = sqrt(h[0]);
            *resid_host = rn;
            if (rn <= atol) return 0;
        }
    }
    *info_host = (int32_t)(maxiter < 0x7fffffff ? maxiter : 0x7fffffff);
    return 0;
}

}  // namespace mwx

using namespace mwx;

extern "C" int mwx_spmv(int64_t n, const int64_t *indptr, const int32_t *indices, const double *values,
                        const double *x, double *y, void *stream) {
    if (n <= 0 || !indptr || !indices || !values || !x || !y) return MWX_E_BADARG;
    return launch_spmv(n, indptr, indices, values, x, y, nullptr, nullptr, nullptr, nullptr, as_stream(stream));
}

extern "C" int mwx_csr_diagonal(int64_t n, const int64_t *indptr, const int32_t *indices,
                                const double *values, double *diag, void *stream) {
    if (n <= 0 || !indptr || !indices || !values || !diag) return MWX_E_BADARG;
    k_csr_diagonal<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(n, indptr, indices, values, diag, 0);
    MWX_LAUNCH_CHECK();
    return 0;
}

extern "C" int mwx_pcg_jacobi(int64_t n, const int64_t *indptr, const int32_t *indices, const double *values,
                              const double *b, double *x, double tol, int64_t maxiter, int32_t precondition,
                              double *work, int64_t *iters_host, int32_t *info_host, double *resid_host,
                              void *stream) {
    if (n <= 0 || !indptr || !indices || !values || !b || !x || !work || !iters_host || !info_host || !resid_host)
        return MWX_E_BADARG;
    cudaStream_t st = as_stream(stream);
    double *r = work, *p = work + n, *q = work + 2 * n, *dinv = work + 3 * n;
    if (precondition) {
        k_csr_diagonal<<<grid_for(n, 256), 256, 0, st>>>(n, indptr, indices, values, dinv, 1);
        MWX_LAUNCH_CHECK();
    }
    return cg_solve(n, indptr, indices, values, b, x, tol, maxiter, precondition ? dinv : nullptr,
                    nullptr, nullptr, r, nullptr, p, q, iters_host, info_host, resid_host, st);
}
