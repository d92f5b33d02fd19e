// Block-CSR SpMV for the coarse operators of the smoothed-aggregation hierarchy (mode 'sa'; no reference counterpart).
//
// With ncand = 3 near-nullspace candidates (the Morley interpolants of 1, x, y) every operator below the fine matrix
// comes in 3-wide blocks: A_l, P_l, R_l (l >= 1) are 3x3-blocked, P_0 has 1x3 blocks (a fine row times the three
// coarse DOFs of an aggregate), R_0 = P_0^T has 3x1 blocks.  In scalar CSR those operators pay 4 index bytes and one
// x gather per nonzero - and the gathers (one L1 wavefront each) are what limits the scalar kernel on 50-nonzero rows
// (level 1 runs at 0.47 of the HBM peak against 0.76 on the fine level).  Here an operator is stored as
//     bptr (block rows + 1), bidx (one block column per block), bval = BR*BC planes of nblocks doubles (SoA)
// so that lanes working on consecutive blocks read every plane fully coalesced, one index and BC consecutive x values
// per block: 8 + 4/(BR*BC) bytes and 1/BR gathers per nonzero (3x3: 8.44 B, a third of the gathers).
//   k_bsr_rows3<BC, MODE> : BR = 3.  LPR lanes (4 / 8 / 16, by the average blocks per block row) share a block row, each
//                           walks its blocks with stride LPR keeping three row sums, a shuffle tree folds them in a
//                           fixed order, lane 0 applies the epilogue to the three rows.
//   k_bsr_rows1<MODE>     : BR = 1, BC = 3 (prolongation from level 1 to the fine level): one thread per row.
// Epilogues as k_spmv (krylov.cu): MODE 0 y = A x, 1 y = c - A x, 2 y = c + A x, 3 Chebyshev / Jacobi step.
// Deterministic (fixed lane assignment and fold order); the sums are associated differently from the scalar kernel,
// so results agree with it to round-off, not bit for bit.
#include "amg_dev.cuh"

namespace mwx {
namespace {

struct BsrSmoother { const double *dinv; double *d; double c1, c2; };

template <int MODE>
__device__ __forceinline__ void bsr_finish(int64_t r, double sum, const double *__restrict__ x, double *y,
                                           const double *cvec, const BsrSmoother &sm) {
    if (MODE == 1) sum = cvec[r] - sum;
    if (MODE == 2) sum = cvec[r] + sum;
    if (MODE == 3) {
        const double dn = sm.c1 * sm.d[r] + sm.c2 * sm.dinv[r] * (cvec[r] - sum);
        sm.d[r] = dn;
        sum = x[r] + dn;
    }
    y[r] = sum;
}

// CTAs per SM the kernels are compiled for (register cap): 4 x 256 threads at 64 registers.  Measured on the level-12
// SA iteration (ms per iteration): unconstrained (rows3 64 / rows1 72 registers, 3 CTAs for rows1) 22.96; 4 / 4: 21.64;
// 5 / 4 (rows3 at 48 registers) 23.41; 6 / 5 spills.
#ifndef MWX_BSR3_MINB
#define MWX_BSR3_MINB 4
#endif
#ifndef MWX_BSR1_MINB
#define MWX_BSR1_MINB 4
#endif
template <typename VT, int BC, int MODE>
__global__ void __launch_bounds__(256, MWX_BSR3_MINB) k_bsr_rows3(int64_t nbrows, int64_t nblocks, const int64_t *__restrict__ bptr,
                                                   const int32_t *__restrict__ bidx, const VT *__restrict__ bval,
                                                   const double *__restrict__ x, double *y, const double *cvec,
                                                   BsrSmoother sm, int lpr) {
    const int sub = threadIdx.x & (lpr - 1);
    const int64_t groups = ((int64_t)gridDim.x * blockDim.x) / lpr;
    const int64_t g0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / lpr;
    const int64_t rounds = (nbrows + groups - 1) / groups;          // uniform trip count: full-mask shuffles are safe
    for (int64_t it = 0; it < rounds; ++it) {
        const int64_t I = g0 + it * groups;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
        if (I < nbrows) {
            const int64_t b1 = bptr[I + 1];
            // two blocks per trip: both index loads, then all value loads, then all x gathers are in flight together
            // (the kernel is latency-bound: ncu long_scoreboard 42 of 47 stall cycles per issue with one block per trip)
            for (int64_t b = bptr[I] + sub; b < b1; b += 2 * lpr) {
                const int64_t bb = b + lpr;
                const bool two = bb < b1;
                const int64_t j0 = (int64_t)bidx[b] * BC, j1 = two ? (int64_t)bidx[bb] * BC : j0;
                double v[2][3 * BC], xv[2][BC];
#pragma unroll
                for (int k = 0; k < 3 * BC; ++k) {
                    v[0][k] = (double)bval[(int64_t)k * nblocks + b];
                    v[1][k] = two ? (double)bval[(int64_t)k * nblocks + bb] : 0.0;
                }
#pragma unroll
                for (int c = 0; c < BC; ++c) { xv[0][c] = __ldg(x + j0 + c); xv[1][c] = __ldg(x + j1 + c); }
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    if (BC == 3) {
                        s0 += v[u][0] * xv[u][0] + v[u][1] * xv[u][1] + v[u][2] * xv[u][2];
                        s1 += v[u][3] * xv[u][0] + v[u][4] * xv[u][1] + v[u][5] * xv[u][2];
                        s2 += v[u][6] * xv[u][0] + v[u][7] * xv[u][1] + v[u][8] * xv[u][2];
                    } else {
                        s0 += v[u][0] * xv[u][0];
                        s1 += v[u][1] * xv[u][0];
                        s2 += v[u][2] * xv[u][0];
                    }
                }
            }
        }
        for (int o = lpr >> 1; o > 0; o >>= 1) {
            s0 += __shfl_xor_sync(0xffffffffu, s0, o);
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        if (I < nbrows && sub == 0) {
            bsr_finish<MODE>(3 * I, s0, x, y, cvec, sm);
            bsr_finish<MODE>(3 * I + 1, s1, x, y, cvec, sm);
            bsr_finish<MODE>(3 * I + 2, s2, x, y, cvec, sm);
        }
    }
}

template <typename VT, int MODE>
__global__ void __launch_bounds__(256, MWX_BSR1_MINB) k_bsr_rows1(int64_t nrows, int64_t nblocks, const int64_t *__restrict__ bptr,
                                                   const int32_t *__restrict__ bidx, const VT *__restrict__ bval,
                                                   const double *__restrict__ x, double *y, const double *cvec,
                                                   BsrSmoother sm) {
    const int64_t st = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nrows; i += st) {
        double s = 0.0;
        const int64_t b1 = bptr[i + 1];
        for (int64_t b = bptr[i]; b < b1; b += 4) {               // four blocks per trip (a fine row meets 2-4 aggregates):
            int64_t j[4];                                         // index loads, value loads and gathers overlap
            double v[4][3], xv[4][3];
#pragma unroll
            for (int u = 0; u < 4; ++u) j[u] = b + u < b1 ? (int64_t)bidx[b + u] * 3 : -1;
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int k = 0; k < 3; ++k) v[u][k] = j[u] >= 0 ? (double)bval[(int64_t)k * nblocks + b + u] : 0.0;
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int k = 0; k < 3; ++k) xv[u][k] = j[u] >= 0 ? __ldg(x + j[u] + k) : 0.0;
#pragma unroll
            for (int u = 0; u < 4; ++u) s += v[u][0] * xv[u][0] + v[u][1] * xv[u][1] + v[u][2] * xv[u][2];
        }
        bsr_finish<MODE>(i, s, x, y, cvec, sm);
    }
}

// ---- conversion from the scalar CSR (block structure checked first) ----
__global__ void k_bsr_check(int64_t nbrows, int br, int bc, const int64_t *__restrict__ Ap, const int32_t *__restrict__ Aj,
                            int *__restrict__ bad) {
    const int64_t I = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (I >= nbrows) return;
    const int64_t s = Ap[br * I], len = Ap[br * I + 1] - s;
    int wrong = (len % bc) != 0 || (s % (br * bc)) != 0;
    for (int r = 1; r < br && !wrong; ++r) wrong |= (Ap[br * I + r + 1] - Ap[br * I + r]) != len;
    for (int64_t p = 0; p < len / bc && !wrong; ++p) {
        const int32_t c0 = Aj[s + bc * p];
        wrong |= (c0 % bc) != 0;
        for (int r = 0; r < br; ++r)
            for (int c = 0; c < bc; ++c) wrong |= Aj[Ap[br * I + r] + bc * p + c] != c0 + c;
    }
    if (wrong) atomicAdd(bad, 1);
}

__global__ void k_bsr_ptr(int64_t nbrows, int br, int bc, const int64_t *__restrict__ Ap, int64_t *__restrict__ bptr) {
    const int64_t I = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (I <= nbrows) bptr[I] = Ap[br * I] / (br * bc);
}

__global__ void k_bsr_fill(int64_t nbrows, int64_t nblocks, int br, int bc, const int64_t *__restrict__ Ap,
                           const int32_t *__restrict__ Aj, const double *__restrict__ Ax,
                           const int64_t *__restrict__ bptr, int32_t *__restrict__ bidx, double *__restrict__ bval) {
    const int64_t I = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (I >= nbrows) return;
    const int64_t b0 = bptr[I], nb = bptr[I + 1] - b0;
    for (int64_t p = 0; p < nb; ++p) {
        bidx[b0 + p] = Aj[Ap[br * I] + bc * p] / bc;
        for (int r = 0; r < br; ++r)
            for (int c = 0; c < bc; ++c) bval[(int64_t)(r * bc + c) * nblocks + b0 + p] = Ax[Ap[br * I + r] + bc * p + c];
    }
}

inline unsigned bsr_grid(int64_t threads_wanted) {
    int64_t g = (threads_wanted + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (g > cap) g = cap;
    return (unsigned)(g < 1 ? 1 : g);
}

template <typename VT, int MODE>
int bsr_launch_t(const DBsr &B, const VT *bval, const double *x, double *y, const double *c, BsrSmoother sm, cudaStream_t st) {
    if (B.br == 3) {
        const int lpr = B.lanes;
        const unsigned grid = bsr_grid(B.nbrows * lpr);
        if (B.bc == 3)
            k_bsr_rows3<VT, 3, MODE><<<grid, 256, 0, st>>>(B.nbrows, B.nblocks, B.bptr, B.bidx, bval, x, y, c, sm, lpr);
        else
            k_bsr_rows3<VT, 1, MODE><<<grid, 256, 0, st>>>(B.nbrows, B.nblocks, B.bptr, B.bidx, bval, x, y, c, sm, lpr);
    } else {
        k_bsr_rows1<VT, MODE><<<bsr_grid(B.nbrows), 256, 0, st>>>(B.nbrows, B.nblocks, B.bptr, B.bidx, bval, x, y, c, sm);
    }
    MWX_LAUNCH_CHECK();
    return 0;
}

template <int MODE>
int bsr_launch(const DBsr &B, const double *x, double *y, const double *c, BsrSmoother sm, cudaStream_t st) {
    if (B.bval32) return bsr_launch_t<float, MODE>(B, B.bval32, x, y, c, sm, st);       // fp32-stored planes
    return bsr_launch_t<double, MODE>(B, B.bval, x, y, c, sm, st);
}

__global__ void k_f64_to_f32(int64_t n, const double *__restrict__ in, float *__restrict__ out) {
    const int64_t st = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += st) out[i] = (float)in[i];
}

}  // namespace

void free_bsr(DBsr &B) {
    cudaFree(B.bptr); cudaFree(B.bidx); cudaFree(B.bval); cudaFree(B.bval32);
    B = DBsr();
}

// Keep the value planes in fp32 only (half the bytes of the operator); index arrays, x, y and the sums are unchanged.
int bsr_store_fp32(DBsr &B, cudaStream_t st) {
    if (!B.bptr || B.bval32 || !B.bval) return 0;
    const int64_t n = B.nblocks * B.br * B.bc;
    if (cudaMalloc((void **)&B.bval32, sizeof(float) * (size_t)(n ? n : 1)) != cudaSuccess) { B.bval32 = nullptr; return MWX_E_NOMEM; }
    k_f64_to_f32<<<bsr_grid(n), 256, 0, st>>>(n, B.bval, B.bval32);
    MWX_LAUNCH_CHECK();
    MWX_CUDA(cudaStreamSynchronize(st));
    cudaFree(B.bval);
    B.bval = nullptr;
    return 0;
}

// Returns 0 and leaves out.bptr == nullptr when the operator does not have the block structure (the caller keeps the
// scalar kernel for it).
int bsr_from_csr(const DCsr &A, int br, int bc, DBsr &out, cudaStream_t st) {
    out = DBsr();
    if (!A.ptr || A.rows % br || A.cols % bc || A.nnz % (br * bc) || !((br == 3 && (bc == 3 || bc == 1)) || (br == 1 && bc == 3)))
        return 0;
    const int64_t nbrows = A.rows / br, nblocks = A.nnz / (br * bc);
    Scratch<int> bad(st);
    MWX_CUDA(bad.alloc(1));
    MWX_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), st));
    k_bsr_check<<<grid_for(nbrows, 256), 256, 0, st>>>(nbrows, br, bc, A.ptr, A.idx, bad.p);
    MWX_LAUNCH_CHECK();
    int h = 0;
    MWX_CUDA(cudaMemcpyAsync(&h, bad.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    MWX_CUDA(cudaStreamSynchronize(st));
    if (h) return 0;
    DBsr B;
    B.br = br; B.bc = bc; B.nbrows = nbrows; B.nblocks = nblocks;
    if (cudaMalloc((void **)&B.bptr, sizeof(int64_t) * ((size_t)nbrows + 1)) != cudaSuccess ||
        cudaMalloc((void **)&B.bidx, sizeof(int32_t) * (size_t)(nblocks ? nblocks : 1)) != cudaSuccess ||
        cudaMalloc((void **)&B.bval, sizeof(double) * (size_t)(A.nnz ? A.nnz : 1)) != cudaSuccess) {
        free_bsr(B);
        return MWX_E_NOMEM;
    }
    k_bsr_ptr<<<grid_for(nbrows + 1, 256), 256, 0, st>>>(nbrows, br, bc, A.ptr, B.bptr);
    k_bsr_fill<<<grid_for(nbrows, 256), 256, 0, st>>>(nbrows, nblocks, br, bc, A.ptr, A.idx, A.val, B.bptr, B.bidx, B.bval);
    MWX_LAUNCH_CHECK();
    const double avg = nbrows > 0 ? (double)nblocks / (double)nbrows : 0.0;
    B.lanes = avg <= 6.0 ? 4 : avg <= 24.0 ? 8 : avg <= 96.0 ? 16 : 32;   // nearly dense tail levels: a whole warp per block row
    out = B;
    return 0;
}

int bsr_spmv(const DBsr &B, const double *x, double *y, cudaStream_t st) {
    return bsr_launch<0>(B, x, y, nullptr, BsrSmoother{}, st);
}

int bsr_spmv_axpy(const DBsr &B, const double *x, double *y, const double *c, int sign, cudaStream_t st) {
    return sign < 0 ? bsr_launch<1>(B, x, y, c, BsrSmoother{}, st) : bsr_launch<2>(B, x, y, c, BsrSmoother{}, st);
}

int bsr_smoother_step(const DBsr &B, const double *x_in, double *x_out, const double *b, const double *dinv, double *d,
                      double c1, double c2, cudaStream_t st) {
    return bsr_launch<3>(B, x_in, x_out, b, BsrSmoother{dinv, d, c1, c2}, st);
}

}  // namespace mwx

// ---- C ABI: block-CSR twins for operators held outside an AMG hierarchy (the row blocks of the partitioned V-cycle) ----
extern "C" int mwx_bsr_create(int64_t rows, int64_t cols, int64_t nnz, const int64_t *indptr, const int32_t *indices,
                              const double *values, int32_t br, int32_t bc, int32_t storage_bits, void **out,
                              void *stream) {
    if (!out || (storage_bits != 64 && storage_bits != 32)) return MWX_E_BADARG;
    *out = nullptr;
    if (rows <= 0 || cols <= 0 || nnz <= 0 || !indptr || !indices || !values) return 0;
    mwx::DCsr A;
    A.rows = rows; A.cols = cols; A.nnz = nnz; A.borrowed = true;
    A.ptr = const_cast<int64_t *>(indptr); A.idx = const_cast<int32_t *>(indices); A.val = const_cast<double *>(values);
    mwx::DBsr B;
    const int rc = mwx::bsr_from_csr(A, br, bc, B, mwx::as_stream(stream));
    if (rc) return rc;
    if (B.bptr && storage_bits == 32) {
        const int rc2 = mwx::bsr_store_fp32(B, mwx::as_stream(stream));
        if (rc2) { mwx::free_bsr(B); return rc2; }
    }
    if (B.bptr) *out = new mwx::DBsr(B);
    return 0;
}

extern "C" void mwx_bsr_destroy(void *b) {
    if (!b) return;
    mwx::DBsr *B = (mwx::DBsr *)b;
    mwx::free_bsr(*B);
    delete B;
}

extern "C" int mwx_bsr_spmv(const void *b, const double *x, double *y, void *stream) {
    if (!b || !x || !y) return MWX_E_BADARG;
    return mwx::bsr_spmv(*(const mwx::DBsr *)b, x, y, mwx::as_stream(stream));
}

extern "C" int mwx_bsr_spmv_axpy(const void *b, const double *x, const double *c, int32_t sign, double *y, void *stream) {
    if (!b || !x || !y || !c || sign == 0) return MWX_E_BADARG;
    return mwx::bsr_spmv_axpy(*(const mwx::DBsr *)b, x, y, c, sign, mwx::as_stream(stream));
}

extern "C" int mwx_bsr_smoother_step(const void *b, const double *x_in, double *x_out, const double *rhs,
                                     const double *dinv, double *d, double c1, double c2, void *stream) {
    if (!b || !x_in || !x_out || !rhs || !dinv || !d || x_in == x_out) return MWX_E_BADARG;
    return mwx::bsr_smoother_step(*(const mwx::DBsr *)b, x_in, x_out, rhs, dinv, d, c1, c2, mwx::as_stream(stream));
}
