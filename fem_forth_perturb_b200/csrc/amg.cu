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
#include <cmath>
#include <cstring>
#include <new>
#include <vector>

#include "amg_dev.cuh"

namespace mwx {
namespace {

__global__ void __launch_bounds__(1024) k_gs_wavefront_sweep(int64_t nwaves, const int64_t *__restrict__ wave_ptr,
                                                             const int32_t *__restrict__ order,
                                                             const int64_t *__restrict__ Ap,
                                                             const int32_t *__restrict__ Aj,
                                                             const double *__restrict__ Ax, double *x,
                                                             const double *__restrict__ b) {
    for (int64_t w = 0; w < nwaves; ++w) {
        const int64_t s = wave_ptr[w], e = wave_ptr[w + 1];
        for (int64_t q = s + threadIdx.x; q < e; q += blockDim.x) {
            const int32_t i = order[q];
            double rsum = 0.0, diag = 0.0;
            for (int64_t k = Ap[i]; k < Ap[i + 1]; ++k) {
                const int32_t j = Aj[k];
                if (j == i) diag = Ax[k];
                else rsum += Ax[k] * x[j];          // plain (coherent) loads: x changes between wavefronts
            }
            if (diag != 0.0) x[i] = (b[i] - rsum) / diag;
        }
        __syncthreads();
    }
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
        cudaFree(L.ord_f); cudaFree(L.ord_b); cudaFree(L.wp_f); cudaFree(L.wp_b);
        cudaFree(L.roots); cudaFree(L.cand);
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
            int64_t nw = 0;
            mwx_gs_wavefronts_host(n, ptr.data(), idx.data(), ord.data(), wp.data(), &nw);
            wp.resize((size_t)nw + 1);
            L.nw_f = nw;
            if ((rc = upload(ord, &L.ord_f, st)) || (rc = upload(wp, &L.wp_f, st))) return rc;
            cudaStreamSynchronize(st);
            wp.assign((size_t)n + 1, 0);
            mwx_gs_wavefronts_backward_host(n, ptr.data(), idx.data(), ord.data(), wp.data(), &nw);
            wp.resize((size_t)nw + 1);
            L.nw_b = nw;
            if ((rc = upload(ord, &L.ord_b, st)) || (rc = upload(wp, &L.wp_b, st))) return rc;
            cudaStreamSynchronize(st);
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

// One smoothing pass on level l.  `x_is_zero`: the iterate is known to be zero (pre-smoothing).
// Returns through *xp the buffer that holds the smoothed iterate (x or the level's ping-pong buffer).
static int smooth(mwx_amg_dev *d, DLevel &L, double *x, const double *b, bool x_is_zero, double **xp,
                  cudaStream_t st) {
    const int64_t n = L.A.rows;
    *xp = x;
    if (d->smoother == 0) {
        if (x_is_zero) MWX_CUDA(cudaMemsetAsync(x, 0, sizeof(double) * (size_t)n, st));
        k_gs_wavefront_sweep<<<1, 1024, 0, st>>>(L.nw_f, L.wp_f, L.ord_f, L.A.ptr, L.A.idx, L.A.val, x, b);
        MWX_LAUNCH_CHECK();
        k_gs_wavefront_sweep<<<1, 1024, 0, st>>>(L.nw_b, L.wp_b, L.ord_b, L.A.ptr, L.A.idx, L.A.val, x, b);
        MWX_LAUNCH_CHECK();
        return 0;
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
                if ((rc = launch_smoother_step(n, L.A.plan, L.A.ptr, L.A.idx, L.A.val, cur, alt, b, L.dinv, L.d, 0.0, omega, st))) return rc;
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
                if ((rc = launch_smoother_step(n, L.A.plan, L.A.ptr, L.A.idx, L.A.val, cur, alt, b, L.dinv, L.d, c1, c2, st))) return rc;
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
    if ((rc = launch_spmv_axpy(n, L.A.plan, L.A.ptr, L.A.idx, L.A.val, xs, L.r, b, -1, st))) return rc;
    DLevel &C = d->lv[(size_t)l + 1];
    const int64_t nc = C.A.rows;
    if ((rc = launch_spmv(nc, L.R.plan, L.R.ptr, L.R.idx, L.R.val, L.r, C.b, nullptr, nullptr, nullptr, nullptr, st))) return rc;
    if (l + 1 == nl - 1) {
        k_dense_gemv<<<grid_for(nc * 32, 256), 256, 0, st>>>(nc, d->pinv, C.b, C.x);
        MWX_LAUNCH_CHECK();
    } else {
        if ((rc = cycle(d, l + 1, C.x, C.b, st))) return rc;
    }
    // x += P x_c   (in place on whichever buffer holds the iterate)
    if ((rc = launch_spmv_axpy(n, L.P.plan, L.P.ptr, L.P.idx, L.P.val, C.x, xs, xs, +1, st))) return rc;
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
                    iters_host, info_host, resid_host, as_stream(stream));
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
                    iters_host, info_host, resid_host, as_stream(stream), max_rechecks, first_hit_host);
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

extern "C" int mwx_amg_level_rho(const mwx_amg_dev_t *d, int32_t level, double *rho_host) {
    if (!d || !rho_host || level < 0 || level >= (int32_t)d->lv.size()) return MWX_E_BADARG;
    *rho_host = d->lv[(size_t)level].rho;
    return 0;
}
