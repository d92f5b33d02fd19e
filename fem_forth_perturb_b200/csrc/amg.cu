// Hot path B: device-resident AMG V(1,1) cycle and the AMG-preconditioned CG.
//
// Replaces  M = ml.aspreconditioner();  spl.cg(A, b, M=M, tol=...)   ref main/example1/utils.py:155-157
// (pyamg multilevel_solver cycle): per level pre-smooth, r = b - A x, b_c = R r, recurse or
// dense pseudo-inverse on the coarsest level, x += P x_c, post-smooth.  The hierarchy comes from
// mwx_amg_setup_host and is uploaded once; every operator application is one of the SpMV kernels of
// krylov.cu.
//
// Smoothers
//   0  parity mode : pyamg's default, one symmetric lexicographic Gauss-Seidel sweep.  The sequential
//      sweep is executed wavefront by wavefront (rows whose lower-numbered neighbours are all done), by
//      one 1024-thread CTA with a block barrier between wavefronts - exact GS order, bit-reproducible.
//      Meant for parity runs (iteration counts within +-1 of the reference), not for throughput.
//   1  fast mode   : Chebyshev polynomial of D^-1 A (degree cheby_degree) - embarrassingly parallel,
//      symmetric, every step is one fused SpMV-epilogue kernel.
//   2  fast mode   : two damped-Jacobi steps (omega = 4 / (3 rho)).
#include <cooperative_groups.h>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "amg_dev.cuh"

namespace mwx {
namespace {

// Exact lexicographic Gauss-Seidel, level-scheduled (GsSweep in amg_dev.cuh).  A wavefront = rows whose lower-numbered
// (forward sweep) neighbours are all done; its rows are independent, wavefronts run in order.  Each row is summed by one
// thread in CSR order with the diagonal skipped - the same sequence of FMAs as a sequential sweep, whatever the launch
// shape.  Before the barrier that ends a wavefront every thread prefetches the records and entries it will read in
// the next one (they do not depend on x), so after the barrier only the x gathers are exposed latency.
__device__ __forceinline__ void gs_prefetch(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// One row: rsum = sum of a_ij x_j over the off-diagonal entries IN CSR ORDER (one FMA chain - the arithmetic of a
// sequential sweep), the x gathers issued eight at a time so that their latencies overlap.
template <bool COHERENT_L1>
__device__ __forceinline__ void gs_row(const GsRow r, double diag, double bi, const int32_t *__restrict__ col,
                                       const double *__restrict__ val, double *x) {
    const int32_t *__restrict__ c = col + r.start;
    const double *__restrict__ v = val + r.start;
    double rsum = 0.0;
    for (int k = 0; k < r.len; k += 8) {
        int32_t cc[8];
        double vv[8], xx[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (k + u < r.len) { cc[u] = c[k + u]; vv[u] = v[k + u]; }
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (k + u < r.len) xx[u] = COHERENT_L1 ? x[cc[u]] : __ldcg(x + cc[u]);
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (k + u < r.len) rsum += vv[u] * xx[u];
    }
    if (diag != 0.0) x[r.row] = (bi - rsum) / diag;
}

// The sweep, shared by the two launch shapes.  Everything a thread needs for its first row of the NEXT wavefront that
// does not depend on x (row record, diagonal, right-hand side, the wavefront bounds; its entries are prefetched into
// L1) is fetched BEFORE the barrier that ends the current wavefront, so after the barrier only the x gathers are
// exposed latency.  SYNC() = block barrier (one CTA) or cooperative grid barrier.
template <bool COHERENT_L1, typename Sync>
__device__ __forceinline__ void gs_sweep(const GsSweep &g, double *x, const double *__restrict__ b, int64_t tid,
                                         int64_t nthreads, Sync sync) {
    if (g.nwaves <= 0) return;
    int64_t s = g.wave_ptr[0], e = g.wave_ptr[1];
    int64_t e_next = g.nwaves > 1 ? g.wave_ptr[2] : e;
    GsRow r = {0, 0, 0};
    double dg = 0.0, bi = 0.0;
    bool have = s + tid < e;
    if (have) { r = g.rows[s + tid]; dg = g.diag[s + tid]; bi = b[r.row]; }
    for (int64_t w = 0; w < g.nwaves; ++w) {
        if (have) gs_row<COHERENT_L1>(r, dg, bi, g.col, g.val, x);
        for (int64_t q = s + tid + nthreads; q < e; q += nthreads) {          // wavefront wider than the launch
            const GsRow r2 = g.rows[q];
            gs_row<COHERENT_L1>(r2, g.diag[q], b[r2.row], g.col, g.val, x);
        }
        const int64_t e_after = w + 3 <= g.nwaves ? g.wave_ptr[w + 3] : e_next;
        have = w + 1 < g.nwaves && e + tid < e_next;
        if (have) {
            r = g.rows[e + tid]; dg = g.diag[e + tid]; bi = b[r.row];
            gs_prefetch(g.col + r.start);
            gs_prefetch(g.val + r.start);
            gs_prefetch(g.val + r.start + 15);
        }
        sync();
        s = e; e = e_next; e_next = e_after;
    }
}

// One CTA: wavefronts narrower than a few CTAs (coarse levels; a block barrier costs less than a grid barrier).
__global__ void __launch_bounds__(512) k_gs_wavefront_sweep(GsSweep g, double *x, const double *__restrict__ b) {
    gs_sweep<true>(g, x, b, threadIdx.x, blockDim.x, [] { __syncthreads(); });      // one SM: its L1 sees its own stores
}

// Small levels (up to 25 600 rows): the whole symmetric sweep - forward, then backward - in one launch with the iterate
// held in shared memory, so that the x gathers between two barriers cost a shared-memory access instead of an L2 round
// trip (a coarse level is a chain of 20-40 narrow wavefronts: pure latency).
constexpr int64_t GS_SMEM_ROWS = 25600;
__global__ void __launch_bounds__(512) k_gs_symmetric_sweep_smem(GsSweep gf, GsSweep gb, double *x, const double *__restrict__ b,
                                                                 int x_is_zero) {
    extern __shared__ double xs[];
    const int64_t n = gf.nrows;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) xs[i] = x_is_zero ? 0.0 : x[i];
    __syncthreads();
    gs_sweep<true>(gf, xs, b, threadIdx.x, blockDim.x, [] { __syncthreads(); });
    gs_sweep<true>(gb, xs, b, threadIdx.x, blockDim.x, [] { __syncthreads(); });
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) x[i] = xs[i];
}

// Cooperative grid: wide wavefronts (skfem numbers the DOFs refinement level by refinement level, so the dependency
// DAG of a 261 k-row Morley operator is only 12 wavefronts deep, ~20 000 rows each).  x is read with ld.global.cg:
// values written by other SMs in the previous wavefront must not come from a stale L1 line.
__global__ void __launch_bounds__(256) k_gs_wavefront_sweep_grid(GsSweep g, double *x, const double *__restrict__ b) {
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    gs_sweep<false>(g, x, b, (int64_t)blockIdx.x * blockDim.x + threadIdx.x, (int64_t)gridDim.x * blockDim.x,
                    [&] { grid.sync(); });
}

int launch_gs_sweep(const GsSweep &g, double *x, const double *b, cudaStream_t st) {
    static int max_blocks = -1;                    // co-resident CTAs of the cooperative kernel on this device
    if (max_blocks < 0) {
        int per_sm = 0, coop = 0, dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
        if (coop && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_gs_wavefront_sweep_grid, 256, 0) == cudaSuccess)
            max_blocks = per_sm * sm_count();
        else
            max_blocks = 0;
    }
    const int64_t avg_wave = g.nwaves > 0 ? g.nrows / g.nwaves : 0;
    if (max_blocks <= 0 || avg_wave < 768) {            // narrow wavefronts: one CTA, block barriers
        k_gs_wavefront_sweep<<<1, 512, 0, st>>>(g, x, b);
        MWX_LAUNCH_CHECK();
        return 0;
    }
    int64_t blocks = (2 * avg_wave + 255) / 256;   // enough threads for a wavefront of twice the average width
    if (blocks > max_blocks) blocks = max_blocks;
    GsSweep gg = g;
    void *args[] = {&gg, &x, &b};
    MWX_CUDA(cudaLaunchCooperativeKernel((const void *)k_gs_wavefront_sweep_grid, dim3((unsigned)blocks), dim3(256), args, 0, st));
    return 0;
}

__global__ void k_dense_gemv(int64_t n, const double *__restrict__ M, const double *__restrict__ v,
                             double *__restrict__ out) {
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    double s = 0.0;
    for (int64_t j = lane; j < n; j += 32) s += M[row * n + j] * v[j];
    s = warp_sum(s);
    if (lane == 0) out[row] = s;
}

__global__ void k_scale_by(int64_t n, const double *__restrict__ a, const double *__restrict__ s, double c,
                           double *__restrict__ out) {
    const int64_t st = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += st)
        out[i] = c * (s ? s[i] : 1.0) * a[i];
}

__global__ void k_add_into(int64_t n, const double *__restrict__ a, double *__restrict__ x) {
    const int64_t st = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += st) x[i] += a[i];
}

__global__ void k_cheby_first_zero(int64_t n, const double *__restrict__ b, const double *__restrict__ dinv,
                                   double c2, double *__restrict__ x, double *__restrict__ d) {
    const int64_t st = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += st) {
        const double v = c2 * dinv[i] * b[i];
        d[i] = v;
        x[i] = v;
    }
}

__global__ void k_seed_vector(int64_t n, double *__restrict__ v) {
    const int64_t st = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += st) {
        uint64_t h = (uint64_t)i * 0x9E3779B97F4A7C15ull;
        h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
        v[i] = 0.5 + (double)(h & 0xFFFFF) / 1048576.0;
    }
}

template <typename T>
int upload(const std::vector<T> &h, T **d, cudaStream_t st) {
    MWX_CUDA(cudaMalloc((void **)d, sizeof(T) * (h.size() ? h.size() : 1)));
    if (!h.empty()) MWX_CUDA(cudaMemcpyAsync(*d, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice, st));
    return 0;
}

int upload_csr(const mwx_amg_host_t *h, int level, int which, DCsr &out, cudaStream_t st) {
    int64_t dims[3];
    int rc = mwx_amg_level_dims(h, level, which, dims);
    if (rc) return rc;
    const int64_t *ptr = nullptr;
    const int32_t *idx = nullptr;
    const double *val = nullptr;
    if ((rc = mwx_amg_level_view(h, level, which, &ptr, &idx, &val))) return rc;      // straight from the hierarchy
    out.rows = dims[0]; out.cols = dims[1]; out.nnz = dims[2];
    const size_t nnz = (size_t)dims[2];
    if (cudaMalloc((void **)&out.ptr, sizeof(int64_t) * ((size_t)dims[0] + 1)) != cudaSuccess ||
        cudaMalloc((void **)&out.idx, sizeof(int32_t) * (nnz ? nnz : 1)) != cudaSuccess ||
        cudaMalloc((void **)&out.val, sizeof(double) * (nnz ? nnz : 1)) != cudaSuccess)
        return MWX_E_NOMEM;
    MWX_CUDA(cudaMemcpyAsync(out.ptr, ptr, sizeof(int64_t) * ((size_t)dims[0] + 1), cudaMemcpyHostToDevice, st));
    MWX_CUDA(cudaMemcpyAsync(out.idx, idx, sizeof(int32_t) * nnz, cudaMemcpyHostToDevice, st));
    MWX_CUDA(cudaMemcpyAsync(out.val, val, sizeof(double) * nnz, cudaMemcpyHostToDevice, st));
    int tile_rows = 0;
    if ((rc = spmv_plan_rows(out.rows, out.ptr, &tile_rows, st))) return rc;
    out.plan = -(int64_t)tile_rows;
    return 0;
}


// Symmetric pseudo-inverse by cyclic Jacobi rotations (coarsest level, a handful of rows); cutoff as in
// scipy 1.4.1 pinv2: singular values below 1e6*eps*max are dropped.
void dense_pinv_sym(int64_t n, std::vector<double> &a, std::vector<double> &out) {
    std::vector<double> v((size_t)(n * n), 0.0);
    for (int64_t i = 0; i < n; ++i) v[(size_t)(i * n + i)] = 1.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0;
        for (int64_t p = 0; p < n; ++p)
            for (int64_t q = p + 1; q < n; ++q) off += a[(size_t)(p * n + q)] * a[(size_t)(p * n + q)];
        if (off < 1e-300) break;
        for (int64_t p = 0; p < n; ++p)
            for (int64_t q = p + 1; q < n; ++q) {
                const double apq = a[(size_t)(p * n + q)];
                if (apq == 0.0) continue;
                const double app = a[(size_t)(p * n + p)], aqq = a[(size_t)(q * n + q)];
                const double tau = (aqq - app) / (2.0 * apq);
                const double t = (tau >= 0 ? 1.0 : -1.0) / (std::fabs(tau) + std::sqrt(1.0 + tau * tau));
                const double c = 1.0 / std::sqrt(1.0 + t * t), s = t * c;
                for (int64_t k = 0; k < n; ++k) {
                    const double akp = a[(size_t)(k * n + p)], akq = a[(size_t)(k * n + q)];
                    a[(size_t)(k * n + p)] = c * akp - s * akq;
                    a[(size_t)(k * n + q)] = s * akp + c * akq;
                }
                for (int64_t k = 0; k < n; ++k) {
                    const double apk = a[(size_t)(p * n + k)], aqk = a[(size_t)(q * n + k)];
                    a[(size_t)(p * n + k)] = c * apk - s * aqk;
                    a[(size_t)(q * n + k)] = s * apk + c * aqk;
                }
                for (int64_t k = 0; k < n; ++k) {
                    const double vkp = v[(size_t)(k * n + p)], vkq = v[(size_t)(k * n + q)];
                    v[(size_t)(k * n + p)] = c * vkp - s * vkq;
                    v[(size_t)(k * n + q)] = s * vkp + c * vkq;
                }
            }
    }
    double smax = 0.0;
    for (int64_t i = 0; i < n; ++i) smax = std::max(smax, std::fabs(a[(size_t)(i * n + i)]));
    const double cut = 1e6 * 2.220446049250313e-16 * smax;
    out.assign((size_t)(n * n), 0.0);
    for (int64_t k = 0; k < n; ++k) {
        const double lam = a[(size_t)(k * n + k)];
        if (std::fabs(lam) <= cut) continue;
        const double inv = 1.0 / lam;
        for (int64_t i = 0; i < n; ++i)
            for (int64_t j = 0; j < n; ++j) out[(size_t)(i * n + j)] += v[(size_t)(i * n + k)] * inv * v[(size_t)(j * n + k)];
    }
}

inline unsigned vgrid(int64_t n) {
    int64_t g = (n + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (g > cap) g = cap;
    return (unsigned)(g < 1 ? 1 : g);
}

}  // namespace

void free_csr(DCsr &m) {
    free_bsr(m.bsr);
    cudaFree(m.pk);
    spmv_lx_destroy(m.lx);
    if (!m.borrowed) { cudaFree(m.ptr); cudaFree(m.idx); cudaFree(m.val); }
    m = DCsr();
}

}  // namespace mwx

using namespace mwx;


extern "C" int mwx_gs_wavefronts_backward_host(int64_t n, const int64_t *indptr, const int32_t *indices,
                                               int32_t *order, int64_t *wave_ptr, int64_t *nwaves);

extern "C" void mwx_amg_destroy_dev(mwx_amg_dev_t *d) {
    if (!d) return;
    for (auto &L : d->lv) {
        free_csr(L.A); free_csr(L.P); free_csr(L.R);
        cudaFree(L.dinv); cudaFree(L.x); cudaFree(L.x2); cudaFree(L.b); cudaFree(L.r); cudaFree(L.d);
        for (GsSweep *g : {&L.gs_f, &L.gs_b}) {
            cudaFree(g->wave_ptr); cudaFree(g->rows); cudaFree(g->col); cudaFree(g->val); cudaFree(g->diag);
        }
        cudaFree(L.roots); cudaFree(L.cand); cudaFree(L.g_b); cudaFree(L.g_x);
    }
    cudaFree(d->pinv); cudaFree(d->sc); cudaFree(d->partials);
    delete d;
}

int mwx::amg_estimate_rho(mwx_amg_dev *d, DLevel &L, cudaStream_t st) {
    // power iteration on D^-1 A with a fixed pseudo-random start (deterministic)
    const int64_t n = L.A.rows;
    double *v = L.x2, *w = L.r;               // scratch: both are free outside a cycle
    k_seed_vector<<<vgrid(n), 256, 0, st>>>(n, v);
    MWX_LAUNCH_CHECK();
    double lam = 1.0, h = 0.0;
    for (int it = 0; it < 20; ++it) {
        int rc = launch_dot(n, v, v, &d->sc->aux, d->partials, &d->sc->ticket[0], st);
        if (rc) return rc;
        MWX_CUDA(cudaMemcpyAsync(&h, &d->sc->aux, sizeof(double), cudaMemcpyDeviceToHost, st));
        MWX_CUDA(cudaStreamSynchronize(st));
        const double nrm = std::sqrt(h);
        if (!(nrm > 0.0)) break;
        if (it > 0) lam = nrm;                      // ||D^-1 A v_prev|| with ||v_prev|| = 1
        k_scale_by<<<vgrid(n), 256, 0, st>>>(n, v, nullptr, 1.0 / nrm, v);
        MWX_LAUNCH_CHECK();
        rc = launch_spmv(n, L.A.plan, L.A.ptr, L.A.idx, L.A.val, v, w, nullptr, nullptr, nullptr, nullptr, st);
        if (rc) return rc;
        k_scale_by<<<vgrid(n), 256, 0, st>>>(n, w, L.dinv, 1.0, v);
        MWX_LAUNCH_CHECK();
    }
    int rc = launch_dot(n, v, v, &d->sc->aux, d->partials, &d->sc->ticket[0], st);
    if (rc) return rc;
    MWX_CUDA(cudaMemcpyAsync(&h, &d->sc->aux, sizeof(double), cudaMemcpyDeviceToHost, st));
    MWX_CUDA(cudaStreamSynchronize(st));
    if (std::sqrt(h) > 0.0) lam = std::sqrt(h);
    L.rho = lam;
    return 0;
}

// Everything a hierarchy needs besides its operators (which are already on the device): SpMV plans, work vectors,
// inverse diagonals, the dense pseudo-inverse of the coarsest operator and the smoother data.
int mwx::amg_finalize_device_levels(mwx_amg_dev *d, const double *coarse_pinv_host, cudaStream_t st) {
    const int nl = (int)d->lv.size();
    int rc = 0;
    if (!d->sc) {
        if (cudaMalloc((void **)&d->sc, sizeof(CgScalars)) != cudaSuccess) return MWX_E_NOMEM;
        if (cudaMalloc((void **)&d->partials, sizeof(double) * 2 * RED_MAX_BLOCKS) != cudaSuccess) return MWX_E_NOMEM;
        cudaMemsetAsync(d->sc, 0, sizeof(CgScalars), st);
    }
    auto plan = [&](DCsr &M) {
        if (M.plan != -1 || !M.ptr) return 0;
        int tile_rows = 0;
        const int prc = spmv_plan_rows(M.rows, M.ptr, &tile_rows, st);
        if (!prc) M.plan = -(int64_t)tile_rows;
        return prc;
    };
    for (int l = 0; l < nl; ++l) {
        DLevel &L = d->lv[(size_t)l];
        if ((rc = plan(L.A))) return rc;
        const int64_t n = L.A.rows;
        if (l + 1 < nl) {
            if ((rc = plan(L.P)) || (rc = plan(L.R))) return rc;
            const size_t bytes = sizeof(double) * (size_t)n;
            const bool have_dinv = L.dinv != nullptr;       // a device setup has already needed it
            auto need = [&](double **p) { return *p || cudaMalloc((void **)p, bytes) == cudaSuccess; };
            if (!need(&L.dinv) || !need(&L.r) || !need(&L.x2) || !need(&L.d)) return MWX_E_NOMEM;
            if (!have_dinv && (rc = launch_csr_diagonal(n, L.A.ptr, L.A.idx, L.A.val, L.dinv, 1, st))) return rc;
        }
        if (l > 0) {
            const size_t bytes = sizeof(double) * (size_t)n;
            if ((!L.x && cudaMalloc((void **)&L.x, bytes) != cudaSuccess) || (!L.b && cudaMalloc((void **)&L.b, bytes) != cudaSuccess))
                return MWX_E_NOMEM;
        }
    }
    // block-CSR twins (bsr.cu) for the operators of a smoothed-aggregation hierarchy built on 3 candidates; an operator
    // that fails the structure check simply keeps the scalar kernel.  MWX_BSR=0 switches them off (A/B runs).
    const char *bsr_env = std::getenv("MWX_BSR");
    const bool use_bsr = !bsr_env || std::atoi(bsr_env) != 0;
    if (use_bsr && d->smoother != 0)
        for (int l = 0; l + 1 < nl; ++l) {
            DLevel &L = d->lv[(size_t)l];
            if (L.coarse_per_root != 3) continue;
            int br = l == 0 ? 1 : 3;
            if (l == 0 && L.A.rows % 3 == 0 && L.A.nnz % 9 == 0 && !L.P.bsr.bptr) {
                // the fine level of a TAIL hierarchy (replicated levels of the partitioned solver) is itself a coarse
                // level of the scheme: its P has 3x3 blocks - the structure check tells (a scalar fine level fails it)
                if ((rc = bsr_from_csr(L.P, 3, 3, L.P.bsr, st))) return rc;
                if (L.P.bsr.bptr) {
                    br = 3;
                    if (!L.A.bsr.bptr && (rc = bsr_from_csr(L.A, 3, 3, L.A.bsr, st))) return rc;
                }
            }
            if (!L.P.bsr.bptr && (rc = bsr_from_csr(L.P, br, 3, L.P.bsr, st))) return rc;
            if (!L.R.bsr.bptr && (rc = bsr_from_csr(L.R, 3, br, L.R.bsr, st))) return rc;
            DLevel &C = d->lv[(size_t)l + 1];
            if (l + 2 < nl && !C.A.bsr.bptr && (rc = bsr_from_csr(C.A, 3, 3, C.A.bsr, st))) return rc;
        }
    // tile-local column plans (spmv_lx.cu) for the scalar operators A_l the smoother and the residual stream - the fine
    // matrix above all.  OFF by default (MWX_SPMV_LX=1 builds them): measured at 67 M rows the plan cuts the L1 wavefronts
    // (496 M per smoother step) and the DRAM bytes (12.4 -> 11.9 GB), but the kernel gets SLOWER, 3.15 against 2.73 ms -
    // its extra barrier per tile (gather distinct columns -> products) costs more than the gathers did (stall reasons:
    // barrier 26 %, long scoreboard 42 % at 3 CTAs per SM).  Kept as an A/B switch with its bit-for-bit test.
    {
        const char *lx_env = std::getenv("MWX_SPMV_LX");
        if (lx_env && std::atoi(lx_env) != 0 && d->smoother != 0)
            for (int l = 0; l + 1 < nl; ++l) {
                DCsr &A = d->lv[(size_t)l].A;
                if (A.bsr.bptr || A.lx || !A.ptr || A.plan > -8) continue;
                if ((rc = spmv_lx_create(A.rows, A.ptr, A.idx, (int)(-A.plan), &A.lx, st))) return rc;
            }
    }
    auto download = [&](const DCsr &M, std::vector<int64_t> &ptr, std::vector<int32_t> &idx, std::vector<double> &val) {
        ptr.resize((size_t)M.rows + 1); idx.resize((size_t)M.nnz); val.resize((size_t)M.nnz);
        MWX_CUDA(cudaMemcpyAsync(ptr.data(), M.ptr, sizeof(int64_t) * ptr.size(), cudaMemcpyDeviceToHost, st));
        MWX_CUDA(cudaMemcpyAsync(idx.data(), M.idx, sizeof(int32_t) * idx.size(), cudaMemcpyDeviceToHost, st));
        MWX_CUDA(cudaMemcpyAsync(val.data(), M.val, sizeof(double) * val.size(), cudaMemcpyDeviceToHost, st));
        MWX_CUDA(cudaStreamSynchronize(st));
        return 0;
    };
    // coarsest level: dense pseudo-inverse
    {
        const DCsr &Ac = d->lv[(size_t)nl - 1].A;
        const int64_t nc = Ac.rows;
        d->n_last = nc;
        std::vector<double> pinv;
        if (coarse_pinv_host) {
            pinv.assign(coarse_pinv_host, coarse_pinv_host + nc * nc);
        } else {
            if (nc > 4096) return MWX_E_CAPACITY;
            std::vector<int64_t> ptr;
            std::vector<int32_t> idx;
            std::vector<double> val, dense((size_t)(nc * nc), 0.0);
            if ((rc = download(Ac, ptr, idx, val))) return rc;
            for (int64_t i = 0; i < nc; ++i)
                for (int64_t k = ptr[(size_t)i]; k < ptr[(size_t)i + 1]; ++k) dense[(size_t)(i * nc + idx[(size_t)k])] += val[(size_t)k];
            for (int64_t i = 0; i < nc; ++i)                 // symmetrise round-off
                for (int64_t j = i + 1; j < nc; ++j) {
                    const double m = 0.5 * (dense[(size_t)(i * nc + j)] + dense[(size_t)(j * nc + i)]);
                    dense[(size_t)(i * nc + j)] = dense[(size_t)(j * nc + i)] = m;
                }
            dense_pinv_sym(nc, dense, pinv);
        }
        if ((rc = upload(pinv, &d->pinv, st))) return rc;
        cudaStreamSynchronize(st);
    }
    // smoother data
    for (int l = 0; l + 1 < nl; ++l) {
        DLevel &L = d->lv[(size_t)l];
        const int64_t n = L.A.rows;
        if (d->smoother == 0) {
            std::vector<int64_t> ptr;
            std::vector<int32_t> idx;
            std::vector<double> val;
            if ((rc = download(L.A, ptr, idx, val))) return rc;
            std::vector<int32_t> ord((size_t)n);
            std::vector<int64_t> wp((size_t)n + 1);
            for (int backward = 0; backward < 2; ++backward) {
                int64_t nw = 0;
                wp.assign((size_t)n + 1, 0);
                if (backward) mwx_gs_wavefronts_backward_host(n, ptr.data(), idx.data(), ord.data(), wp.data(), &nw);
                else mwx_gs_wavefronts_host(n, ptr.data(), idx.data(), ord.data(), wp.data(), &nw);
                wp.resize((size_t)nw + 1);
                // pack the operator in wavefront order: one record per row, off-diagonal entries end to end
                std::vector<GsRow> rows((size_t)n);
                std::vector<int32_t> col;
                std::vector<double> vals, diag((size_t)n, 0.0);
                col.reserve(idx.size()); vals.reserve(val.size());
                for (int64_t q = 0; q < n; ++q) {
                    const int32_t i = ord[(size_t)q];
                    GsRow r;
                    r.row = i; r.start = (int64_t)col.size();
                    for (int64_t k = ptr[(size_t)i]; k < ptr[(size_t)i + 1]; ++k) {
                        if (idx[(size_t)k] == i) diag[(size_t)q] = val[(size_t)k];       // pyamg: the last stored diagonal entry
                        else { col.push_back(idx[(size_t)k]); vals.push_back(val[(size_t)k]); }
                    }
                    r.len = (int32_t)((int64_t)col.size() - r.start);
                    rows[(size_t)q] = r;
                }
                GsSweep &g = backward ? L.gs_b : L.gs_f;
                g.nwaves = nw; g.nrows = n;
                if ((rc = upload(wp, &g.wave_ptr, st)) || (rc = upload(rows, &g.rows, st)) || (rc = upload(col, &g.col, st)) ||
                    (rc = upload(vals, &g.val, st)) || (rc = upload(diag, &g.diag, st)))
                    return rc;
                cudaStreamSynchronize(st);
            }
        } else {
            if (L.rho <= 0.0 && (rc = amg_estimate_rho(d, L, st))) return rc;
        }
    }
    cudaStreamSynchronize(st);
    return 0;
}

extern "C" int mwx_amg_upload(const mwx_amg_host_t *h, int32_t smoother, int32_t cheby_degree,
                              const double *coarse_pinv_host, mwx_amg_dev_t **out, void *stream) {
    if (!h || !out || smoother < 0 || smoother > 2) return MWX_E_BADARG;
    cudaStream_t st = as_stream(stream);
    const int nl = mwx_amg_num_levels(h);
    if (nl < 1) return MWX_E_BADARG;
    auto *d = new (std::nothrow) mwx_amg_dev();
    if (!d) return MWX_E_NOMEM;
    d->smoother = smoother;
    d->degree = cheby_degree > 0 ? cheby_degree : 2;
    d->lv.resize((size_t)nl);
    int rc = 0;
    auto fail = [&](int code) { mwx_amg_destroy_dev(d); return code; };
    for (int l = 0; l < nl; ++l) {
        DLevel &L = d->lv[(size_t)l];
        if ((rc = upload_csr(h, l, 0, L.A, st))) return fail(rc);
        if (l + 1 < nl) {
            if ((rc = upload_csr(h, l, 1, L.P, st))) return fail(rc);
            if ((rc = upload_csr(h, l, 2, L.R, st))) return fail(rc);
        }
    }
    if ((rc = amg_finalize_device_levels(d, coarse_pinv_host, st))) return fail(rc);
    *out = d;
    return 0;
}

// Operator applications of the V-cycle: the block kernels where the operator has a block-CSR twin, else k_spmv.
static int op_spmv(const DCsr &M, const double *x, double *y, cudaStream_t st) {
    if (M.bsr.bptr) return bsr_spmv(M.bsr, x, y, st);
    if (M.pk) return launch_spmv_pk(M.rows, M.plan, M.ptr, M.pk, x, y, st);
    return launch_spmv(M.rows, M.plan, M.ptr, M.idx, M.val, x, y, nullptr, nullptr, nullptr, nullptr, st, M.lx);
}
static int op_spmv_axpy(const DCsr &M, const double *x, double *y, const double *c, int sign, cudaStream_t st) {
    if (M.bsr.bptr) return bsr_spmv_axpy(M.bsr, x, y, c, sign, st);
    if (M.pk) return launch_spmv_axpy_pk(M.rows, M.plan, M.ptr, M.pk, x, y, c, sign, st);
    return launch_spmv_axpy(M.rows, M.plan, M.ptr, M.idx, M.val, x, y, c, sign, st, M.lx);
}
static int op_smoother_step(const DCsr &M, const double *x_in, double *x_out, const double *b, const double *dinv,
                            double *d, double c1, double c2, cudaStream_t st) {
    if (M.bsr.bptr) return bsr_smoother_step(M.bsr, x_in, x_out, b, dinv, d, c1, c2, st);
    if (M.pk) return launch_smoother_step_pk(M.rows, M.plan, M.ptr, M.pk, x_in, x_out, b, dinv, d, c1, c2, st);
    return launch_smoother_step(M.rows, M.plan, M.ptr, M.idx, M.val, x_in, x_out, b, dinv, d, c1, c2, st, M.lx);
}

// One smoothing pass on level l.  `x_is_zero`: the iterate is known to be zero (pre-smoothing).
// Returns through *xp the buffer that holds the smoothed iterate (x or the level's ping-pong buffer).
static int smooth(mwx_amg_dev *d, DLevel &L, double *x, const double *b, bool x_is_zero, double **xp,
                  cudaStream_t st) {
    const int64_t n = L.A.rows;
    *xp = x;
    if (d->smoother == 0) {
        if (n <= GS_SMEM_ROWS && L.gs_f.nwaves > 0 && n / L.gs_f.nwaves < 768) {
            static bool configured = false;
            if (!configured) {
                MWX_CUDA(cudaFuncSetAttribute(k_gs_symmetric_sweep_smem, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)(GS_SMEM_ROWS * sizeof(double))));
                configured = true;
            }
            k_gs_symmetric_sweep_smem<<<1, 512, sizeof(double) * (size_t)n, st>>>(L.gs_f, L.gs_b, x, b, x_is_zero ? 1 : 0);
            MWX_LAUNCH_CHECK();
            return 0;
        }
        if (x_is_zero) MWX_CUDA(cudaMemsetAsync(x, 0, sizeof(double) * (size_t)n, st));
        int rc;
        if ((rc = launch_gs_sweep(L.gs_f, x, b, st))) return rc;
        return launch_gs_sweep(L.gs_b, x, b, st);
    }
    double *cur = x, *alt = L.x2;
    int rc;
    if (d->smoother == 2) {
        const double omega = 4.0 / (3.0 * L.rho);
        for (int s = 0; s < 2; ++s) {
            if (s == 0 && x_is_zero) {
                k_cheby_first_zero<<<vgrid(n), 256, 0, st>>>(n, b, L.dinv, omega, cur, L.d);
                MWX_LAUNCH_CHECK();
            } else {
                if ((rc = op_smoother_step(L.A, cur, alt, b, L.dinv, L.d, 0.0, omega, st))) return rc;
                std::swap(cur, alt);
            }
        }
    } else {
        const double lmax = 1.1 * L.rho, lmin = lmax / d->cheb_ratio;
        const double theta = 0.5 * (lmax + lmin), delta = 0.5 * (lmax - lmin);
        const double sigma = theta / delta;
        double rho_k = 1.0 / sigma;
        for (int s = 0; s < d->degree; ++s) {
            double c1, c2;
            if (s == 0) { c1 = 0.0; c2 = 1.0 / theta; }
            else {
                const double rho_n = 1.0 / (2.0 * sigma - rho_k);
                c1 = rho_n * rho_k; c2 = 2.0 * rho_n / delta;
                rho_k = rho_n;
            }
            if (s == 0 && x_is_zero) {
                k_cheby_first_zero<<<vgrid(n), 256, 0, st>>>(n, b, L.dinv, c2, cur, L.d);
                MWX_LAUNCH_CHECK();
            } else {
                if ((rc = op_smoother_step(L.A, cur, alt, b, L.dinv, L.d, c1, c2, st))) return rc;
                std::swap(cur, alt);
            }
        }
    }
    *xp = cur;
    return 0;
}

static int cycle(mwx_amg_dev *d, int l, double *x, const double *b, cudaStream_t st) {
    DLevel &L = d->lv[(size_t)l];
    const int64_t n = L.A.rows;
    const int nl = (int)d->lv.size();
    int rc;
    double *xs = x;
    if ((rc = smooth(d, L, x, b, true, &xs, st))) return rc;
    if ((rc = op_spmv_axpy(L.A, xs, L.r, b, -1, st))) return rc;
    DLevel &C = d->lv[(size_t)l + 1];
    const int64_t nc = C.A.rows;
    if ((rc = op_spmv(L.R, L.r, C.b, st))) return rc;
    if (l + 1 == nl - 1) {
        k_dense_gemv<<<grid_for(nc * 32, 256), 256, 0, st>>>(nc, d->pinv, C.b, C.x);
        MWX_LAUNCH_CHECK();
    } else {
        if ((rc = cycle(d, l + 1, C.x, C.b, st))) return rc;
        if (d->gamma == 2 && l + 1 >= d->gamma_first && l + 1 <= d->gamma_last && C.g_b && C.g_x) {
            // second visit (W-cycle flavour): x_c += M_c^-1 (b_c - A_c x_c).  Two steps of the symmetric stationary
            // iteration from zero = (2 M^-1 - M^-1 A M^-1) b: the preconditioner stays symmetric positive definite.
            if ((rc = op_spmv_axpy(C.A, C.x, C.g_b, C.b, -1, st))) return rc;
            if ((rc = cycle(d, l + 1, C.g_x, C.g_b, st))) return rc;
            k_add_into<<<vgrid(nc), 256, 0, st>>>(nc, C.g_x, C.x);
            MWX_LAUNCH_CHECK();
        }
    }
    // x += P x_c   (in place on whichever buffer holds the iterate)
    if ((rc = op_spmv_axpy(L.P, C.x, xs, xs, +1, st))) return rc;
    double *xo = xs;
    if (d->smoother == 0) {
        if ((rc = smooth(d, L, xs, b, false, &xo, st))) return rc;
    } else {
        // make the ping-pong land in the caller's buffer: smooth from xs, alternate is the other one
        double *other = (xs == x) ? L.x2 : x;
        double *keep = L.x2;
        L.x2 = other;
        rc = smooth(d, L, xs, b, false, &xo, st);
        L.x2 = keep;
        if (rc) return rc;
    }
    if (xo != x) MWX_CUDA(cudaMemcpyAsync(x, xo, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, st));
    return 0;
}

extern "C" int mwx_amg_vcycle(mwx_amg_dev_t *d, const double *r, double *z, void *stream) {
    if (!d || !r || !z) return MWX_E_BADARG;
    cudaStream_t st = as_stream(stream);
    if (d->lv.size() == 1) {
        const int64_t n = d->n_last;
        k_dense_gemv<<<grid_for(n * 32, 256), 256, 0, st>>>(n, d->pinv, r, z);
        MWX_LAUNCH_CHECK();
        return 0;
    }
    return cycle(d, 0, z, r, st);
}

// CG multiplies with the caller's matrix; the fine level's column plan describes its PATTERN, so it applies when the
// caller's pattern arrays are the ones the hierarchy was built on (a device setup borrows them).
static const SpmvLX *system_lx(const mwx_amg_dev *d, int64_t n, const int64_t *indptr, const int32_t *indices) {
    const DCsr &A = d->lv[0].A;
    return (A.lx && A.rows == n && A.ptr == indptr && A.idx == indices) ? A.lx : nullptr;
}

static int amg_precond(void *ctx, const double *r, double *z, cudaStream_t st) {
    return mwx_amg_vcycle((mwx_amg_dev_t *)ctx, r, z, (void *)st);
}

extern "C" int mwx_pcg_amg(mwx_amg_dev_t *d, int64_t n, const int64_t *indptr, const int32_t *indices,
                           const double *values, const double *b, double *x, double tol, int64_t maxiter,
                           double *work, int64_t *iters_host, int32_t *info_host, double *resid_host,
                           void *stream) {
    if (!d || n <= 0 || !indptr || !indices || !values || !b || !x || !work || !iters_host || !info_host || !resid_host)
        return MWX_E_BADARG;
    if (d->lv[0].A.rows != n) return MWX_E_BADARG;
    double *r = work, *z = work + n, *p = work + 2 * n, *q = work + 3 * n;
    return cg_solve(n, indptr, indices, values, b, x, tol, maxiter, nullptr, amg_precond, d, r, z, p, q,
                    iters_host, info_host, resid_host, as_stream(stream), 0, nullptr, nullptr, 0, system_lx(d, n, indptr, indices));
}

extern "C" int mwx_pcg_amg_monitored(mwx_amg_dev_t *d, int64_t n, const int64_t *indptr, const int32_t *indices,
                                     const double *values, const double *b, double *x, double tol, int64_t maxiter,
                                     int32_t max_rechecks, double *work, int64_t *iters_host, int32_t *info_host,
                                     double *resid_host, int64_t *first_hit_host, void *stream) {
    if (!d || n <= 0 || !indptr || !indices || !values || !b || !x || !work || !iters_host || !info_host || !resid_host ||
        max_rechecks < 0)
        return MWX_E_BADARG;
    if (d->lv[0].A.rows != n) return MWX_E_BADARG;
    double *r = work, *z = work + n, *p = work + 2 * n, *q = work + 3 * n;
    return cg_solve(n, indptr, indices, values, b, x, tol, maxiter, nullptr, amg_precond, d, r, z, p, q,
                    iters_host, info_host, resid_host, as_stream(stream), max_rechecks, first_hit_host, nullptr, 0,
                    system_lx(d, n, indptr, indices));
}

extern "C" int mwx_pcg_amg_history(mwx_amg_dev_t *d, int64_t n, const int64_t *indptr, const int32_t *indices,
                                   const double *values, const double *b, double *x, double tol, int64_t maxiter,
                                   double *work, int64_t *iters_host, int32_t *info_host, double *resid_host,
                                   double *xnorm_hist_host, int64_t xnorm_cap, void *stream) {
    if (!d || n <= 0 || !indptr || !indices || !values || !b || !x || !work || !iters_host || !info_host || !resid_host ||
        !xnorm_hist_host || xnorm_cap < 0)
        return MWX_E_BADARG;
    if (d->lv[0].A.rows != n) return MWX_E_BADARG;
    double *r = work, *z = work + n, *p = work + 2 * n, *q = work + 3 * n;
    return cg_solve(n, indptr, indices, values, b, x, tol, maxiter, nullptr, amg_precond, d, r, z, p, q,
                    iters_host, info_host, resid_host, as_stream(stream), 0, nullptr, xnorm_hist_host, xnorm_cap,
                    system_lx(d, n, indptr, indices));
}

extern "C" int mwx_smoother_first(int64_t n, const double *b, const double *dinv, double c2, double *x, double *d,
                                  void *stream) {
    if (n <= 0 || !b || !dinv || !x || !d) return MWX_E_BADARG;
    k_cheby_first_zero<<<vgrid(n), 256, 0, as_stream(stream)>>>(n, b, dinv, c2, x, d);
    MWX_LAUNCH_CHECK();
    return 0;
}

extern "C" int mwx_dense_gemv(int64_t n, const double *M, const double *v, double *out, void *stream) {
    if (n <= 0 || !M || !v || !out) return MWX_E_BADARG;
    k_dense_gemv<<<grid_for(n * 32, 256), 256, 0, as_stream(stream)>>>(n, M, v, out);
    MWX_LAUNCH_CHECK();
    return 0;
}

extern "C" int mwx_amg_set_chebyshev(mwx_amg_dev_t *d, int32_t degree, double ratio) {
    if (!d || degree < 1 || degree > 16 || !(ratio > 1.0)) return MWX_E_BADARG;
    d->degree = degree;
    d->cheb_ratio = ratio;
    return 0;
}

// Keep the operators the V-cycle applies in fp32 (values only): block-CSR planes are converted in place.  Scalar
// operators (the fine matrix) stay fp64: their kernel is bound by the L1 wavefronts of the x gathers, not by DRAM, and
// the packed (fp32 value, column) twin measured SLOWER (2.95 against 2.73 ms per smoother step at 67 M rows: one more
// shared-memory read per record pair) - MWX_PACKED_CSR=1 builds those twins anyway (A/B runs, tests).  The fp64 CSR
// arrays stay for the accessors; vectors, products and sums stay fp64.  Fast smoothers only (parity mode is the
// reference's arithmetic).  bits = 64 is a no-op (the twins are not rebuilt in fp64).
extern "C" int mwx_amg_set_storage(mwx_amg_dev_t *d, int32_t bits, void *stream) {
    if (!d || (bits != 32 && bits != 64)) return MWX_E_BADARG;
    if (bits == 64 || d->smoother == 0) return 0;
    cudaStream_t st = as_stream(stream);
    const int nl = (int)d->lv.size();
    const char *pk_env = std::getenv("MWX_PACKED_CSR");
    const bool packed_too = pk_env && std::atoi(pk_env) != 0;
    for (int l = 0; l + 1 < nl; ++l) {
        DLevel &L = d->lv[(size_t)l];
        DCsr *ops[3] = {&L.A, &L.P, &L.R};
        for (DCsr *M : ops) {
            if (!M->ptr || M->nnz <= 0) continue;
            int rc;
            if (M->bsr.bptr) {
                if ((rc = bsr_store_fp32(M->bsr, st))) return rc;
            } else if (packed_too && !M->pk) {
                if (cudaMalloc((void **)&M->pk, sizeof(PackedNz) * (size_t)M->nnz) != cudaSuccess) { M->pk = nullptr; return MWX_E_NOMEM; }
                if ((rc = csr_pack(M->nnz, M->idx, M->val, M->pk, st))) return rc;
            }
        }
    }
    MWX_CUDA(cudaStreamSynchronize(st));
    d->storage_bits = 32;
    return 0;
}

extern "C" int mwx_amg_set_cycle(mwx_amg_dev_t *d, int32_t gamma, int32_t first_level, int32_t last_level) {
    if (!d || (gamma != 1 && gamma != 2) || first_level < 1 || last_level < 0) return MWX_E_BADARG;
    const int nl = (int)d->lv.size();
    if (gamma == 2)
        for (int l = first_level; l < nl - 1 && l <= last_level; ++l) {           // the coarsest level is a direct solve: never revisited
            DLevel &L = d->lv[(size_t)l];
            const size_t bytes = sizeof(double) * (size_t)L.A.rows;
            if ((!L.g_b && cudaMalloc((void **)&L.g_b, bytes) != cudaSuccess) ||
                (!L.g_x && cudaMalloc((void **)&L.g_x, bytes) != cudaSuccess))
                return MWX_E_NOMEM;
        }
    d->gamma = gamma;
    d->gamma_first = first_level;
    d->gamma_last = last_level;
    return 0;
}

extern "C" int mwx_amg_level_rho(const mwx_amg_dev_t *d, int32_t level, double *rho_host) {
    if (!d || !rho_host || level < 0 || level >= (int32_t)d->lv.size()) return MWX_E_BADARG;
    *rho_host = d->lv[(size_t)level].rho;
    return 0;
}
