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
#include "tma.cuh"

namespace mwx {
namespace {

constexpr int SPMV_ROWS = 128;                 // rows per tile
constexpr int SPMV_THREADS = 256;              // all threads gather; rows of a tile are dealt round-robin
constexpr int SPMV_MAX_TILE_ROWS = 1024;       // very sparse operators (prolongation): long tiles
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

// Shared-memory plan of the SpMV CTA: two stages of (values, column indices) filled by TMA.
#ifndef MWX_SPMV_STAGE_ROWS
// Stage capacity (in 20-nonzero rows) and CTAs per SM, tuned together (profiles/r02_spmv_stage_tuning.txt, 67 M rows):
// 128 rows x 3 CTAs (80 registers) 2.16 ms; 112 x 4 (64 registers, the same 192-row tiles) 2.02 ms; 96 x 4 2.06; 80 x 4 2.27;
// 64 x 5 2.42; 160 x 2 2.34 - the kernel is latency-bound, one more resident CTA pays as long as the tile height stays.
#define MWX_SPMV_STAGE_ROWS 112
#endif
#ifndef MWX_SPMV_MINB
#define MWX_SPMV_MINB 4
#endif
constexpr int SPMV_STAGE_ELEMS = MWX_SPMV_STAGE_ROWS * 20 + 8;       // staged nonzeros per stage (<= 20 per row)
constexpr int SPMV_STAGE_BYTES = SPMV_STAGE_ELEMS * 12;             // 8 B value + 4 B column
constexpr int SPMV_SMEM_BYTES = 2 * SPMV_STAGE_BYTES + 64 + 32 * 8; // + barriers/meta + reduction scratch
// PK (fp32-stored operator, see PackedNz in krylov.cuh): one 8-byte (value, column) record per nonzero; the fp64
// product overwrites the record it came from, so a stage is 8 B per nonzero and ONE bulk copy per tile.
constexpr int SPMV_STAGE_BYTES_PK = SPMV_STAGE_ELEMS * 8;
constexpr int SPMV_SMEM_BYTES_PK = 2 * SPMV_STAGE_BYTES_PK + 64 + 32 * 8;
// LX (tile-local columns, spmv_lx.cu): a stage = values (8 B) + 2-byte local indices + the tile's distinct columns;
// one buffer of gathered x entries per CTA.
constexpr int SPMV_LX_LIDX_OFF = SPMV_STAGE_ELEMS * 8;
constexpr int SPMV_LX_UCOL_OFF = SPMV_LX_LIDX_OFF + SPMV_STAGE_ELEMS * 2;
constexpr int SPMV_STAGE_BYTES_LX = SPMV_LX_UCOL_OFF + LX_MAX_U * 4;
constexpr int SPMV_SMEM_BYTES_LX = 2 * SPMV_STAGE_BYTES_LX + 2 * LX_MAX_U * 8 + 64 + 32 * 8;
static_assert(SPMV_LX_LIDX_OFF % 16 == 0 && SPMV_LX_UCOL_OFF % 16 == 0 && SPMV_STAGE_BYTES_LX % 16 == 0, "TMA alignment");
enum { FMT_CSR = 0, FMT_PK = 1, FMT_LX = 2 };
struct LxArgs { const int32_t *uptr; const int32_t *ucols; const uint16_t *lidx; };

// MODE 0: y = A x   1: y = c - A x   2: y = c + A x      (c = `cvec`, may alias y for MODE 2)
// MODE 3 (smoother step): d = c1 d + c2 dinv (c - A x);  y = x + d     (y must not alias x)
struct SmootherArgs { const double *dinv; double *d; double c1, c2; };

// Persistent CTAs stride over 128-row tiles.  One elected thread streams the tile's contiguous
// values / column indices into shared memory with two TMA bulk copies (cp.async.bulk + mbarrier
// complete_tx), one tile ahead of the consumers; the 128 threads then only issue the x gathers (all of
// a thread's gathers are in flight together), write the products back in place, and one thread per
// row adds its products in CSR order.
template <bool DOT, int MODE, int FMT = FMT_CSR>
__global__ void __launch_bounds__(SPMV_THREADS, MWX_SPMV_MINB) k_spmv(int64_t n, const int64_t *__restrict__ indptr,
                                                    const int32_t *__restrict__ indices,
                                                    const double *__restrict__ values,
                                                    const double *__restrict__ x, double *y,
                                                    const double *cvec,
                                                    const double *__restrict__ dot_with,
                                                    double *__restrict__ dot_out, double *partials,
                                                    unsigned int *ticket, SmootherArgs sm, int tma_ok,
                                                    int rows_per_tile, int lanes_per_row, LxArgs lx) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr bool PK = FMT == FMT_PK, LX = FMT == FMT_LX;
    constexpr int STAGE_BYTES = PK ? SPMV_STAGE_BYTES_PK : (LX ? SPMV_STAGE_BYTES_LX : SPMV_STAGE_BYTES);
    const PackedNz *__restrict__ packed = reinterpret_cast<const PackedNz *>(values);     // PK: `values` is the packed array
    auto stage_vals = [&](int stage) { return reinterpret_cast<double *>(smem_raw + stage * STAGE_BYTES); };
    auto stage_cols = [&](int stage) {
        return reinterpret_cast<int32_t *>(smem_raw + stage * STAGE_BYTES + SPMV_STAGE_ELEMS * 8);          // !PK only
    };
    auto stage_lidx = [&](int stage) { return reinterpret_cast<uint16_t *>(smem_raw + stage * STAGE_BYTES + SPMV_LX_LIDX_OFF); };
    auto stage_ucol = [&](int stage) { return reinterpret_cast<int32_t *>(smem_raw + stage * STAGE_BYTES + SPMV_LX_UCOL_OFF); };
    // LX: gathered x entries of the staged tile, one buffer per stage (the next tile's are fetched during this tile's row sums)
    auto stage_xs = [&](int stage) { return reinterpret_cast<double *>(smem_raw + 2 * STAGE_BYTES + stage * LX_MAX_U * 8); };
    unsigned char *tail = smem_raw + 2 * STAGE_BYTES + (LX ? 2 * LX_MAX_U * 8 : 0);
    uint64_t *bar = reinterpret_cast<uint64_t *>(tail);                 // [2]
    int64_t *s_base = reinterpret_cast<int64_t *>(tail + 16);           // [2] first staged nonzero
    int32_t *s_nst = reinterpret_cast<int32_t *>(tail + 32);            // [2] staged count, -1 = direct path
    int32_t *s_ucnt = reinterpret_cast<int32_t *>(tail + 40);           // [2] LX: distinct columns of the staged tile
    double *red = reinterpret_cast<double *>(tail + 64);                // [32]

    const int64_t ntiles = (n + rows_per_tile - 1) / rows_per_tile;
    const int tid = threadIdx.x;
    int64_t nnz_total = 0;
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_fence_init();
        nnz_total = indptr[n];
    }
    // producer (thread 0): describe tile -> stage and start its TMA copies
    auto tile_bounds = [&](int64_t tile, int64_t &s0, int64_t &s1) {
        const int64_t r0 = tile * rows_per_tile;
        const int64_t r1 = r0 + rows_per_tile < n ? r0 + rows_per_tile : n;
        s0 = indptr[r0];
        s1 = indptr[r1];
    };
    auto issue = [&](int64_t s0, int64_t s1, int stage, int64_t tl) {
        constexpr int64_t AL = LX ? 7 : 3;                          // 16-byte aligned in every staged array (LX: 2-byte indices)
        const int64_t base = s0 & ~AL, end = (s1 + AL) & ~AL;
        const int64_t nst = end - base;
        int32_t u0 = 0, ucnt = 0;
        if (LX) { u0 = lx.uptr[tl]; ucnt = lx.uptr[tl + 1] - u0; }
        if (tma_ok && nst <= SPMV_STAGE_ELEMS && end <= nnz_total && nst > 0 && (!LX || ucnt > 0)) {
            s_base[stage] = base;
            s_nst[stage] = (int32_t)nst;
            fence_proxy_async();                                   // consumers' generic accesses -> before the refill
            if (LX) {
                s_ucnt[stage] = ucnt;
                mbar_expect_tx(&bar[stage], (uint32_t)(nst * 10 + (int64_t)ucnt * 4));
                bulk_g2s(stage_vals(stage), values + base, (uint32_t)(nst * 8), &bar[stage]);
                bulk_g2s(stage_lidx(stage), lx.lidx + base, (uint32_t)(nst * 2), &bar[stage]);
                bulk_g2s(stage_ucol(stage), lx.ucols + u0, (uint32_t)(ucnt * 4), &bar[stage]);
            } else if (PK) {
                mbar_expect_tx(&bar[stage], (uint32_t)(nst * 8));
                bulk_g2s(stage_vals(stage), packed + base, (uint32_t)(nst * 8), &bar[stage]);
            } else {
                mbar_expect_tx(&bar[stage], (uint32_t)(nst * 12));
                bulk_g2s(stage_vals(stage), values + base, (uint32_t)(nst * 8), &bar[stage]);
                bulk_g2s(stage_cols(stage), indices + base, (uint32_t)(nst * 4), &bar[stage]);
            }
        } else {
            s_base[stage] = s0;
            s_nst[stage] = -1;
        }
    };
    int64_t tile = blockIdx.x;
    int64_t ns0 = 0, ns1 = 0;                 // thread 0: bounds of the NEXT tile, loaded one iteration early
    if (tid == 0 && tile < ntiles) {
        tile_bounds(tile, ns0, ns1);
        issue(ns0, ns1, 0, tile);
        if (tile + gridDim.x < ntiles) tile_bounds(tile + gridDim.x, ns0, ns1);
    }
    __syncthreads();
    uint32_t phase0 = 0, phase1 = 0;
    double dacc = 0.0;
    bool xs_ready = false;                   // LX: this tile's x entries were already gathered (CTA-uniform)
    for (int k = 0; tile < ntiles; tile += gridDim.x, ++k) {
        const int stage = k & 1;
        const int64_t next = tile + gridDim.x;
        if (tid == 0 && next < ntiles) {
            issue(ns0, ns1, stage ^ 1, next);
            if (next + gridDim.x < ntiles) tile_bounds(next + gridDim.x, ns0, ns1);   // lands during this tile
        }
        const int64_t r0 = tile * rows_per_tile;
        const int64_t r1 = r0 + rows_per_tile < n ? r0 + rows_per_tile : n;
        const int nst = s_nst[stage];
        const int64_t base = s_base[stage];
        double *vs = stage_vals(stage);
        if (nst >= 0) {
            const int32_t *cs = stage_cols(stage);
            if (!(LX && xs_ready)) {
                if (stage == 0) { mbar_wait(&bar[0], phase0); phase0 ^= 1; }
                else            { mbar_wait(&bar[1], phase1); phase1 ^= 1; }
            }
            constexpr int U = (SPMV_STAGE_ELEMS + 2 * SPMV_THREADS - 1) / (2 * SPMV_THREADS);
            if (LX) {
                // every distinct column of the tile is gathered ONCE (sorted addresses), the products read x from shared
                const uint16_t *ls = stage_lidx(stage);
                const double *xs = stage_xs(stage);
                if (!xs_ready) {                     // first tile of the CTA (or the one before took the direct path)
                    const int32_t *us = stage_ucol(stage);
                    const int ucnt = s_ucnt[stage];
                    for (int j = tid; j < ucnt; j += SPMV_THREADS) stage_xs(stage)[j] = __ldg(x + us[j]);
                    __syncthreads();
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int e = tid * 2 + u * 2 * SPMV_THREADS;
                    if (e < nst) {
                        const uint32_t li = *reinterpret_cast<const uint32_t *>(ls + e);
                        double2 v = *reinterpret_cast<double2 *>(vs + e);
                        v.x *= xs[li & 0xffffu];
                        v.y *= xs[li >> 16];
                        *reinterpret_cast<double2 *>(vs + e) = v;
                    }
                }
                __syncthreads();
            } else {
            // all of this thread's gathers are issued before any is consumed (U x 2 in flight)
            int2 c[U];
            double xa[U], xb[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int e = tid * 2 + u * 2 * SPMV_THREADS;
                if (PK) {
                    // two (value, column) records = 16 bytes; staged counts are multiples of 4, so the pair is in range
                    const int4 rec = e < nst ? *reinterpret_cast<const int4 *>(vs + e) : make_int4(0, 0, 0, 0);
                    c[u] = make_int2(rec.y, rec.w);             // the values are re-read from the stage below (registers)
                } else {
                    c[u] = e < nst ? *reinterpret_cast<const int2 *>(cs + e) : make_int2(0, 0);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int e = tid * 2 + u * 2 * SPMV_THREADS;
                if (e < nst) { xa[u] = __ldg(x + c[u].x); xb[u] = __ldg(x + c[u].y); }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int e = tid * 2 + u * 2 * SPMV_THREADS;
                if (e < nst) {
                    double2 v;
                    if (PK) {
                        const float4 rec = *reinterpret_cast<const float4 *>(vs + e);
                        v = make_double2((double)rec.x, (double)rec.z);
                    } else {
                        v = *reinterpret_cast<double2 *>(vs + e);
                    }
                    v.x *= xa[u];
                    v.y *= xb[u];
                    *reinterpret_cast<double2 *>(vs + e) = v;
                }
            }
            __syncthreads();
            }
        }
        // LX: the next tile's stage has usually landed by now - start the gathers of its distinct columns, let them fly
        // during this tile's row sums, store them afterwards.  (s_nst / s_ucnt of the other stage were written by thread 0
        // before the barrier that ended the product phase.)
        constexpr int XG = (LX_MAX_U + SPMV_THREADS - 1) / SPMV_THREADS;
        double xg[XG];
        bool pre = false;
        int pre_cnt = 0;
        if (LX && nst >= 0 && next < ntiles && s_nst[stage ^ 1] >= 0) {
            if (stage == 0) { mbar_wait(&bar[1], phase1); phase1 ^= 1; }
            else            { mbar_wait(&bar[0], phase0); phase0 ^= 1; }
            const int32_t *us = stage_ucol(stage ^ 1);
            pre_cnt = s_ucnt[stage ^ 1];
#pragma unroll
            for (int q = 0; q < XG; ++q) {
                const int j = tid + q * SPMV_THREADS;
                if (j < pre_cnt) xg[q] = __ldg(x + us[j]);
            }
            pre = true;
        }
        auto finish = [&](int64_t r, double sum) {
            if (MODE == 1) sum = cvec[r] - sum;
            if (MODE == 2) sum = cvec[r] + sum;
            if (MODE == 3) {
                const double dn = sm.c1 * sm.d[r] + sm.c2 * sm.dinv[r] * (cvec[r] - sum);
                sm.d[r] = dn;
                sum = x[r] + dn;
            }
            y[r] = sum;
            if (DOT) dacc += dot_with[r] * sum;
        };
        if (lanes_per_row == 1) {
            // one thread per row (a thread owns several rows when the tile is longer than the CTA)
            for (int64_t r = r0 + tid; r < r1; r += SPMV_THREADS) {
                const int64_t a = indptr[r], b = indptr[r + 1];
                double sum = 0.0;
                if (nst >= 0) {
                    for (int e = (int)(a - base); e < (int)(b - base); ++e) sum += vs[e];
                } else if (PK) {
                    for (int64_t e = a; e < b; ++e) { const PackedNz q = packed[e]; sum += (double)q.v * __ldg(x + q.c); }
                } else {
                    for (int64_t e = a; e < b; ++e) sum += values[e] * __ldg(x + indices[e]);
                }
                finish(r, sum);
            }
        } else {
            // dense operators (AMG coarse levels, restriction: 50-80 nonzeros per row, 16-64 rows per tile): a group of
            // 2-16 lanes shares a row, strided partial sums folded by shuffles in a fixed order
            const int lpr = lanes_per_row, sub = tid & (lpr - 1), grp = tid / lpr, ngrp = SPMV_THREADS / lpr;
            for (int64_t rb = r0; rb < r1; rb += ngrp) {           // CTA-uniform trip count: full-mask shuffles are safe
                const int64_t r = rb + grp;
                const bool valid = r < r1;
                double sum = 0.0;
                if (valid) {
                    const int64_t a = indptr[r], b = indptr[r + 1];
                    if (nst >= 0) {
                        for (int e = (int)(a - base) + sub; e < (int)(b - base); e += lpr) sum += vs[e];
                    } else if (PK) {
                        for (int64_t e = a + sub; e < b; e += lpr) { const PackedNz q = packed[e]; sum += (double)q.v * __ldg(x + q.c); }
                    } else {
                        for (int64_t e = a + sub; e < b; e += lpr) sum += values[e] * __ldg(x + indices[e]);
                    }
                }
                for (int o = lpr >> 1; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
                if (valid && sub == 0) finish(r, sum);
            }
        }
        // The products were written back INTO the TMA stage through the generic proxy: every writing thread orders
        // its own stores before the async-proxy refill (PTX memory model: the fence must come from the writers, a
        // barrier plus thread 0's fence alone is not enough), then the barrier hands the stage back to the producer.
        if (LX) {
            if (pre) {
                double *xn = stage_xs(stage ^ 1);
#pragma unroll
                for (int q = 0; q < XG; ++q) {
                    const int j = tid + q * SPMV_THREADS;
                    if (j < pre_cnt) xn[j] = xg[q];
                }
            }
            xs_ready = pre;
        }
        if (nst >= 0) fence_proxy_async();
        __syncthreads();                 // stage fully consumed before thread 0 refills it next iteration
    }
    if (DOT) {
        const double bs = block_sum(dacc, red);
        double total;
        if (fold_partials(bs, partials, ticket, red, total) && tid == 0) *dot_out = total;
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

// r = b - A x with every row accumulated in double-double (error-free products by FMA, two-sum accumulation) and
// rounded once at the end: the residual a mixed-precision iterative refinement needs.  On the finest meshes the fp64
// residual carries a rounding error eps ||A|| ||x|| that is LARGER than tol ||b|| (condition number ~1e10), so neither
// the stopping rule nor the forward error can go below ~1e-6 without it.  One thread per row; not a hot kernel.
__global__ void k_residual_dd(int64_t n, const int64_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                              const double *__restrict__ values, const double *__restrict__ x,
                              const double *__restrict__ b, double *__restrict__ r) {
    const int64_t st = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += st) {
        double hi = b[i], lo = 0.0;
        for (int64_t k = indptr[i]; k < indptr[i + 1]; ++k) {
            const double a = values[k], xv = x[indices[k]];
            // round-to-nearest intrinsics: the compiler must not contract these into FMAs or reassociate them
            const double p = -__dmul_rn(a, xv);
            const double pe = -__fma_rn(a, xv, p);         // a * xv = -(p + pe) exactly
            const double s = __dadd_rn(hi, p);             // two-sum of hi and p: hi + p = s + e exactly
            const double bb = __dadd_rn(s, -hi);
            const double e = __dadd_rn(__dadd_rn(hi, -__dadd_rn(s, -bb)), __dadd_rn(p, -bb));
            hi = s;
            lo = __dadd_rn(lo, __dadd_rn(e, pe));
        }
        r[i] = __dadd_rn(hi, lo);
    }
}

inline unsigned vec_grid(int64_t n) {
    int64_t g = (n + VEC_THREADS - 1) / VEC_THREADS;
    const int64_t cap = (int64_t)sm_count() * 8;          // 8 x 256 threads = full residency
    if (g > cap) g = cap;
    if (g > RED_MAX_BLOCKS) g = RED_MAX_BLOCKS;
    return (unsigned)(g < 1 ? 1 : g);
}

// Rows per tile such that an average tile (+25 %) fits one TMA stage; denser operators (AMG coarse
// levels, restriction) get shorter tiles instead of falling off the staged path.
inline int spmv_rows_per_tile(int64_t n, int64_t nnz) {
    int rows = SPMV_ROWS;
    if (nnz > 0 && n > 0) {
        const double avg = (double)nnz / (double)n;
        const double fit = (SPMV_STAGE_ELEMS - 8) / (avg * 1.2 + 0.5);      // 20 % headroom for long-row clusters
        rows = fit > SPMV_MAX_TILE_ROWS ? SPMV_MAX_TILE_ROWS : (int)fit;
        rows = rows >= 32 ? (rows / 32) * 32 : (rows < 4 ? 4 : rows);
    }
    return rows;
}

constexpr int SPMV_NCAND = 12;
__constant__ int c_spmv_cand[SPMV_NCAND] = {1024, 768, 512, 384, 256, 192, 160, 128, 96, 64, 32, 16};

// max over tiles of the (16-byte padded) staged element count, for every candidate tile height
__global__ void k_spmv_plan(int64_t n, const int64_t *__restrict__ indptr, int *__restrict__ max_staged) {
    const int cand = blockIdx.y;
    const int rows = c_spmv_cand[cand];
    const int64_t ntiles = (n + rows - 1) / rows;
    int64_t m = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < ntiles; t += stride) {
        const int64_t r0 = t * rows, r1 = r0 + rows < n ? r0 + rows : n;
        const int64_t base = indptr[r0] & ~(int64_t)3, end = (indptr[r1] + 3) & ~(int64_t)3;
        m = end - base > m ? end - base : m;
    }
    for (int o = 16; o > 0; o >>= 1) {
        const int64_t v = __shfl_xor_sync(0xffffffffu, m, o);
        m = v > m ? v : m;
    }
    if ((threadIdx.x & 31) == 0) atomicMax(&max_staged[cand], (int)(m > 0x7fffffff ? 0x7fffffff : m));
}

inline unsigned spmv_grid(int64_t n, int ctas_per_sm, int rows_per_tile) {
    int64_t g = (n + rows_per_tile - 1) / rows_per_tile;
    const int64_t cap = (int64_t)sm_count() * ctas_per_sm;   // persistent: exactly one resident wave
    if (g > cap) g = cap;
    if (g > RED_MAX_BLOCKS) g = RED_MAX_BLOCKS;
    return (unsigned)(g < 1 ? 1 : g);
}

}  // namespace

// Exact tile height for a matrix: the tallest candidate whose largest tile still fits one TMA stage
// (one tiny kernel + one 48-byte read-back; done once per operator, not per SpMV).
int spmv_plan_rows(int64_t n, const int64_t *indptr, int *rows_out, cudaStream_t st) {
    Scratch<int> mx(st);
    MWX_CUDA(mx.alloc(SPMV_NCAND));
    MWX_CUDA(cudaMemsetAsync(mx.p, 0, sizeof(int) * SPMV_NCAND, st));
    int64_t gx = (n / 16 + 255) / 256;
    if (gx > 512) gx = 512;
    if (gx < 1) gx = 1;
    k_spmv_plan<<<dim3((unsigned)gx, SPMV_NCAND), 256, 0, st>>>(n, indptr, mx.p);
    MWX_LAUNCH_CHECK();
    int h[SPMV_NCAND];
    MWX_CUDA(cudaMemcpyAsync(h, mx.p, sizeof(h), cudaMemcpyDeviceToHost, st));
    MWX_CUDA(cudaStreamSynchronize(st));
    static const int cand[SPMV_NCAND] = {1024, 768, 512, 384, 256, 192, 160, 128, 96, 64, 32, 16};
    int rows = 8;
    for (int c = 0; c < SPMV_NCAND; ++c)
        if (h[c] <= SPMV_STAGE_ELEMS) { rows = cand[c]; break; }
    // short tiles mean long rows: let 4 / 8 / 16 lanes share a row (encoded above the tile height, see spmv_launch)
    const int lpr_log2 = rows <= 16 ? 4 : (rows <= 32 ? 3 : (rows <= 64 ? 2 : 0));
    *rows_out = rows + 4096 * lpr_log2;
    return 0;
}

namespace {

// PK: `values` points at the packed (fp32 value, column) records and `indices` is unused.
template <bool DOT, int MODE, int FMT = FMT_CSR>
int spmv_launch_fmt(int64_t n, int64_t nnz, const int64_t *indptr, const int32_t *indices, const double *values, const double *x,
                    double *y, const double *c, const double *dot_with, double *dot_out, double *partials,
                    unsigned int *ticket, SmootherArgs sm, LxArgs lx, cudaStream_t st) {
    static int ctas_per_sm = 0;                   // per instantiation; the attributes are per device function
    constexpr int SMEM = FMT == FMT_PK ? SPMV_SMEM_BYTES_PK : (FMT == FMT_LX ? SPMV_SMEM_BYTES_LX : SPMV_SMEM_BYTES);
    if (!ctas_per_sm) {
        MWX_CUDA(cudaFuncSetAttribute(k_spmv<DOT, MODE, FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        MWX_CUDA(cudaFuncSetAttribute(k_spmv<DOT, MODE, FMT>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        int occ = 0;
        MWX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_spmv<DOT, MODE, FMT>, SPMV_THREADS, SMEM));
        ctas_per_sm = occ > 0 ? occ : 1;
    }
    // TMA bulk copies need 16-byte aligned global addresses; the kernel aligns offsets, the bases must be too
    const int tma_ok = (((uintptr_t)values & 15) == 0 && (FMT != FMT_CSR || ((uintptr_t)indices & 15) == 0)) ? 1 : 0;
    // nnz <= -8: a plan from spmv_plan_rows = -(tile height + 4096 * log2(lanes per row))
    const int plan = nnz <= -8 ? (int)(-nnz) : spmv_rows_per_tile(n, nnz);
    const int rows = plan & 4095, lanes = 1 << (plan >> 12);
    k_spmv<DOT, MODE, FMT><<<spmv_grid(n, ctas_per_sm, rows), SPMV_THREADS, SMEM, st>>>(
        n, indptr, indices, values, x, y, c, dot_with, dot_out, partials, ticket, sm, tma_ok, rows, lanes, lx);
    MWX_LAUNCH_CHECK();
    return 0;
}

// PK: `values` points at the packed records.  lx != nullptr (a plan built for exactly this tile height): the
// tile-local-column kernel; otherwise the plain one.
template <bool DOT, int MODE, bool PK = false>
int spmv_launch(int64_t n, int64_t nnz, const int64_t *indptr, const int32_t *indices, const double *values, const double *x,
                double *y, const double *c, const double *dot_with, double *dot_out, double *partials,
                unsigned int *ticket, SmootherArgs sm, cudaStream_t st, const SpmvLX *lx = nullptr) {
    if (PK)
        return spmv_launch_fmt<DOT, MODE, FMT_PK>(n, nnz, indptr, indices, values, x, y, c, dot_with, dot_out, partials, ticket,
                                                  sm, LxArgs{}, st);
    if (lx && nnz <= -8 && (int)(-nnz) == lx->plan && lx->n == n)
        return spmv_launch_fmt<DOT, MODE, FMT_LX>(n, nnz, indptr, indices, values, x, y, c, dot_with, dot_out, partials, ticket,
                                                  sm, LxArgs{lx->uptr, lx->ucols, lx->lidx}, st);
    return spmv_launch_fmt<DOT, MODE, FMT_CSR>(n, nnz, indptr, indices, values, x, y, c, dot_with, dot_out, partials, ticket,
                                               sm, LxArgs{}, st);
}

__global__ void k_csr_pack(int64_t nnz, const int32_t *__restrict__ indices, const double *__restrict__ values,
                           PackedNz *__restrict__ out) {
    const int64_t st = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += st) {
        PackedNz q;
        q.v = (float)values[k];
        q.c = indices[k];
        out[k] = q;
    }
}

}  // namespace

int launch_spmv(int64_t n, int64_t nnz, const int64_t *indptr, const int32_t *indices, const double *values,
                const double *x, double *y, const double *dot_with, double *dot_out, double *partials,
                unsigned int *ticket, cudaStream_t st, const SpmvLX *lx) {
    if (dot_with)
        return spmv_launch<true, 0>(n, nnz, indptr, indices, values, x, y, nullptr, dot_with, dot_out, partials, ticket,
                                    SmootherArgs{}, st, lx);
    return spmv_launch<false, 0>(n, nnz, indptr, indices, values, x, y, nullptr, nullptr, nullptr, nullptr, nullptr,
                                 SmootherArgs{}, st, lx);
}

int launch_spmv_axpy(int64_t n, int64_t nnz, const int64_t *indptr, const int32_t *indices, const double *values,
                     const double *x, double *y, const double *c, int sign, cudaStream_t st, const SpmvLX *lx) {
    if (sign < 0)
        return spmv_launch<false, 1>(n, nnz, indptr, indices, values, x, y, c, nullptr, nullptr, nullptr, nullptr,
                                     SmootherArgs{}, st, lx);
    return spmv_launch<false, 2>(n, nnz, indptr, indices, values, x, y, c, nullptr, nullptr, nullptr, nullptr,
                                 SmootherArgs{}, st, lx);
}

int launch_smoother_step(int64_t n, int64_t nnz, const int64_t *indptr, const int32_t *indices, const double *values,
                         const double *x_in, double *x_out, const double *b, const double *dinv, double *d,
                         double c1, double c2, cudaStream_t st, const SpmvLX *lx) {
    return spmv_launch<false, 3>(n, nnz, indptr, indices, values, x_in, x_out, b, nullptr, nullptr, nullptr, nullptr,
                                 SmootherArgs{dinv, d, c1, c2}, st, lx);
}

// fp32-stored twins: the same kernels reading 8-byte (fp32 value, column) records; vectors and sums stay fp64.
int csr_pack(int64_t nnz, const int32_t *indices, const double *values, PackedNz *out, cudaStream_t st) {
    if (nnz <= 0) return 0;
    k_csr_pack<<<vec_grid(nnz), VEC_THREADS, 0, st>>>(nnz, indices, values, out);
    MWX_LAUNCH_CHECK();
    return 0;
}

int launch_spmv_pk(int64_t n, int64_t nnz, const int64_t *indptr, const PackedNz *pk, const double *x, double *y,
                   cudaStream_t st) {
    return spmv_launch<false, 0, true>(n, nnz, indptr, nullptr, reinterpret_cast<const double *>(pk), x, y, nullptr, nullptr,
                                       nullptr, nullptr, nullptr, SmootherArgs{}, st);
}

int launch_spmv_axpy_pk(int64_t n, int64_t nnz, const int64_t *indptr, const PackedNz *pk, const double *x, double *y,
                        const double *c, int sign, cudaStream_t st) {
    if (sign < 0)
        return spmv_launch<false, 1, true>(n, nnz, indptr, nullptr, reinterpret_cast<const double *>(pk), x, y, c, nullptr,
                                           nullptr, nullptr, nullptr, SmootherArgs{}, st);
    return spmv_launch<false, 2, true>(n, nnz, indptr, nullptr, reinterpret_cast<const double *>(pk), x, y, c, nullptr,
                                       nullptr, nullptr, nullptr, SmootherArgs{}, st);
}

int launch_smoother_step_pk(int64_t n, int64_t nnz, const int64_t *indptr, const PackedNz *pk, const double *x_in,
                            double *x_out, const double *b, const double *dinv, double *d, double c1, double c2,
                            cudaStream_t st) {
    return spmv_launch<false, 3, true>(n, nnz, indptr, nullptr, reinterpret_cast<const double *>(pk), x_in, x_out, b, nullptr,
                                       nullptr, nullptr, nullptr, SmootherArgs{dinv, d, c1, c2}, st);
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
             int64_t *iters_host, int32_t *info_host, double *resid_host, cudaStream_t st, int max_rechecks,
             int64_t *first_hit_host, double *xnorm_hist_host, int64_t xnorm_cap, const SpmvLX *lx) {
    Scratch<CgScalars> sc(st);
    int rechecks = 0;
    int64_t floor_it = 0;
    if (first_hit_host) *first_hit_host = 0;
    Scratch<double> partials(st);
    MWX_CUDA(sc.alloc(1));
    MWX_CUDA(partials.alloc(2 * RED_MAX_BLOCKS));
    MWX_CUDA(cudaMemsetAsync(sc.p, 0, sizeof(CgScalars), st));
    MWX_CUDA(cudaMemsetAsync(x, 0, sizeof(double) * (size_t)n, st));
    if (maxiter <= 0) maxiter = 10 * n;
    double h[2];
    int64_t nnz = -1;                              // SpMV tile height, planned once per solve (encoded as -rows)
    {
        int rows = 0;
        int prc = spmv_plan_rows(n, indptr, &rows, st);
        if (prc) return prc;
        nnz = -(int64_t)rows;
    }
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
        rc = launch_spmv(n, nnz, indptr, indices, values, p, q, p, &sc.p->pq, partials.p, &sc.p->ticket[0], st, lx);
        if (rc) return rc;
        k_xr_update<<<vec_grid(n), VEC_THREADS, 0, st>>>(n, p, q, x, r, dinv, sc.p, partials.p, jacobi ? 1 : 0);
        MWX_LAUNCH_CHECK();
        const bool want_xnorm = xnorm_hist_host && it <= xnorm_cap;
        if (want_xnorm) {                                        // rr and aux are adjacent: one 16-byte read-back
            rc = launch_dot(n, x, x, &sc.p->aux, partials.p, &sc.p->ticket[0], st);
            if (rc) return rc;
        }
        MWX_CUDA(cudaMemcpyAsync(h, &sc.p->rr, sizeof(double) * (want_xnorm ? 2 : 1), cudaMemcpyDeviceToHost, st));
        MWX_CUDA(cudaStreamSynchronize(st));
        if (want_xnorm) xnorm_hist_host[it - 1] = sqrt(h[1]);
        double rn = sqrt(h[0]);
        *iters_host = it; *resid_host = rn;
        if (!(rn == rn)) { *info_host = (int32_t)(it < 0x7fffffff ? it : 0x7fffffff); return MWX_E_NOCONV; }
        if (rn <= atol) {
            if (first_hit_host && *first_hit_host == 0) *first_hit_host = it;
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
            // scipy 1.4.1 carries on from the recomputed residual until maxiter.  max_rechecks > 0 (an extension, off
            // by default): give up after that many failed re-checks - the recurrence has converged but the TRUE
            // residual sits on its fp64 floor (eps_mach ||A|| ||x|| > tol ||b|| on very fine meshes).
            if (max_rechecks > 0 && ++rechecks >= max_rechecks) {
                *info_host = (int32_t)(it < 0x7fffffff ? it : 0x7fffffff);
                return 0;
            }
            if (max_rechecks > 0 && floor_it == 0) floor_it = it;
        }
        // ... or 10 iterations after the first failed re-check: from the recomputed residual the recurrence often
        // hovers just above the threshold and never asks for another re-check
        if (max_rechecks > 0 && floor_it > 0 && it >= floor_it + 10) {
            rc = launch_residual(n, indptr, indices, values, x, b, r, jacobi ? dinv : nullptr, sc.p, partials.p, st);
            if (rc) return rc;
            MWX_CUDA(cudaMemcpyAsync(h, &sc.p->rr, sizeof(double), cudaMemcpyDeviceToHost, st));
            MWX_CUDA(cudaStreamSynchronize(st));
            *resid_host = sqrt(h[0]);
            *info_host = *resid_host <= atol ? 0 : (int32_t)(it < 0x7fffffff ? it : 0x7fffffff);
            return 0;
        }
    }
    *info_host = (int32_t)(maxiter < 0x7fffffff ? maxiter : 0x7fffffff);
    return 0;
}

}  // namespace mwx

using namespace mwx;

extern "C" int mwx_spmv(int64_t n, int64_t nnz, const int64_t *indptr, const int32_t *indices,
                        const double *values, const double *x, double *y, void *stream) {
    if (n <= 0 || !indptr || !indices || !values || !x || !y) return MWX_E_BADARG;
    return launch_spmv(n, nnz, indptr, indices, values, x, y, nullptr, nullptr, nullptr, nullptr, as_stream(stream));
}

extern "C" int mwx_spmv_axpy(int64_t n, int64_t nnz, const int64_t *indptr, const int32_t *indices,
                             const double *values, const double *x, const double *c, int32_t sign, double *y,
                             void *stream) {
    if (n <= 0 || !indptr || !indices || !values || !x || !c || !y || sign == 0) return MWX_E_BADARG;
    return launch_spmv_axpy(n, nnz, indptr, indices, values, x, y, c, sign, as_stream(stream));
}

extern "C" int mwx_smoother_step(int64_t n, int64_t nnz, const int64_t *indptr, const int32_t *indices,
                                 const double *values, const double *x_in, double *x_out, const double *b,
                                 const double *dinv, double *d, double c1, double c2, void *stream) {
    if (n <= 0 || !indptr || !indices || !values || !x_in || !x_out || !b || !dinv || !d || x_in == x_out)
        return MWX_E_BADARG;
    return launch_smoother_step(n, nnz, indptr, indices, values, x_in, x_out, b, dinv, d, c1, c2, as_stream(stream));
}

extern "C" int mwx_spmv_plan(int64_t n, const int64_t *indptr, int32_t *tile_rows_host, void *stream) {
    if (n <= 0 || !indptr || !tile_rows_host) return MWX_E_BADARG;
    int rows = 0;
    int rc = spmv_plan_rows(n, indptr, &rows, as_stream(stream));
    *tile_rows_host = rows;
    return rc;
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
                    nullptr, nullptr, r, nullptr, p, q, iters_host, info_host, resid_host, st, 0, nullptr);
}

extern "C" int mwx_pcg_jacobi_history(int64_t n, const int64_t *indptr, const int32_t *indices, const double *values,
                                      const double *b, double *x, double tol, int64_t maxiter, int32_t precondition,
                                      double *work, int64_t *iters_host, int32_t *info_host, double *resid_host,
                                      double *xnorm_hist_host, int64_t xnorm_cap, void *stream) {
    if (n <= 0 || !indptr || !indices || !values || !b || !x || !work || !iters_host || !info_host || !resid_host ||
        !xnorm_hist_host || xnorm_cap < 0)
        return MWX_E_BADARG;
    cudaStream_t st = as_stream(stream);
    double *r = work, *p = work + n, *q = work + 2 * n, *dinv = work + 3 * n;
    if (precondition) {
        k_csr_diagonal<<<grid_for(n, 256), 256, 0, st>>>(n, indptr, indices, values, dinv, 1);
        MWX_LAUNCH_CHECK();
    }
    return cg_solve(n, indptr, indices, values, b, x, tol, maxiter, precondition ? dinv : nullptr,
                    nullptr, nullptr, r, nullptr, p, q, iters_host, info_host, resid_host, st, 0, nullptr,
                    xnorm_hist_host, xnorm_cap);
}

// ------------------------------------------------------------------ building blocks for the partitioned CG
// The multi-GPU driver keeps one workspace per rank: ws[0..4] = {rho, rho_prev, pq, rr, aux} (the
// CgScalars doubles; ws[5..7] hold its counters and must not be touched), ws[8..] = reduction partials.
// Every kernel below leaves LOCAL partial sums in the scalar slots; the driver all-reduces them.
static inline CgScalars *ws_scalars(double *ws) { return reinterpret_cast<CgScalars *>(ws); }
static inline double *ws_partials(double *ws) { return ws + 8; }

extern "C" int mwx_dcg_workspace_doubles(void) { return 8 + 2 * RED_MAX_BLOCKS; }

extern "C" int mwx_dcg_dot(int64_t n, const double *a, const double *b, double *ws, int32_t slot, void *stream) {
    if (n <= 0 || !a || !b || !ws || slot < 0 || slot > 4) return MWX_E_BADARG;
    CgScalars *sc = ws_scalars(ws);
    return launch_dot(n, a, b, reinterpret_cast<double *>(sc) + slot, ws_partials(ws), &sc->ticket[0], as_stream(stream));
}

extern "C" int mwx_dcg_p_update(int64_t n, const double *r, const double *dinv, const double *z, double *p,
                                double *ws, int32_t first, void *stream) {
    if (n <= 0 || !p || !ws || (!z && !r)) return MWX_E_BADARG;
    k_p_update<<<vec_grid(n), VEC_THREADS, 0, as_stream(stream)>>>(n, r, dinv, z, p, ws_scalars(ws), first);
    MWX_LAUNCH_CHECK();
    return 0;
}

extern "C" int mwx_dcg_spmv_dot(int64_t n, int64_t nnz_hint, const int64_t *indptr, const int32_t *indices,
                                const double *values, const double *x_ext, double *q, double *ws, void *stream) {
    if (n <= 0 || !indptr || !indices || !values || !x_ext || !q || !ws) return MWX_E_BADARG;
    CgScalars *sc = ws_scalars(ws);
    // x_ext = [owned | ghosts]; the first n entries are the owned p used in p.Ap
    return launch_spmv(n, nnz_hint, indptr, indices, values, x_ext, q, x_ext, &sc->pq, ws_partials(ws), &sc->ticket[0],
                       as_stream(stream));
}

extern "C" int mwx_dcg_xr_update(int64_t n, const double *p, const double *q, double *x, double *r,
                                 const double *dinv, double *ws, int32_t jacobi, void *stream) {
    if (n <= 0 || !p || !q || !x || !r || !ws) return MWX_E_BADARG;
    k_xr_update<<<vec_grid(n), VEC_THREADS, 0, as_stream(stream)>>>(n, p, q, x, r, dinv, ws_scalars(ws), ws_partials(ws),
                                                                    jacobi ? 1 : 0);
    MWX_LAUNCH_CHECK();
    return 0;
}

extern "C" int mwx_dcg_residual(int64_t n, const int64_t *indptr, const int32_t *indices, const double *values,
                                const double *x_ext, const double *b, double *r, const double *dinv, double *ws,
                                void *stream) {
    if (n <= 0 || !indptr || !indices || !values || !x_ext || !b || !r || !ws) return MWX_E_BADARG;
    return launch_residual(n, indptr, indices, values, x_ext, b, r, dinv, ws_scalars(ws), ws_partials(ws), as_stream(stream));
}

// fp32-stored twin of a CSR operator for the kernels above (row blocks of the partitioned V-cycle): packed must hold
// nnz 8-byte records, 16-byte aligned.
extern "C" int mwx_csr_pack(int64_t nnz, const int32_t *indices, const double *values, void *packed, void *stream) {
    if (nnz < 0 || (nnz && (!indices || !values || !packed))) return MWX_E_BADARG;
    return csr_pack(nnz, indices, values, (PackedNz *)packed, as_stream(stream));
}

extern "C" int mwx_spmv_pk(int64_t n, int64_t nnz, const int64_t *indptr, const void *packed, const double *x, double *y,
                           void *stream) {
    if (n <= 0 || !indptr || !packed || !x || !y) return MWX_E_BADARG;
    return launch_spmv_pk(n, nnz, indptr, (const PackedNz *)packed, x, y, as_stream(stream));
}

extern "C" int mwx_spmv_axpy_pk(int64_t n, int64_t nnz, const int64_t *indptr, const void *packed, const double *x,
                                const double *c, int32_t sign, double *y, void *stream) {
    if (n <= 0 || !indptr || !packed || !x || !c || !y || sign == 0) return MWX_E_BADARG;
    return launch_spmv_axpy_pk(n, nnz, indptr, (const PackedNz *)packed, x, y, c, sign, as_stream(stream));
}

extern "C" int mwx_smoother_step_pk(int64_t n, int64_t nnz, const int64_t *indptr, const void *packed, const double *x_in,
                                    double *x_out, const double *b, const double *dinv, double *d, double c1, double c2,
                                    void *stream) {
    if (n <= 0 || !indptr || !packed || !x_in || !x_out || !b || !dinv || !d || x_in == x_out) return MWX_E_BADARG;
    return launch_smoother_step_pk(n, nnz, indptr, (const PackedNz *)packed, x_in, x_out, b, dinv, d, c1, c2, as_stream(stream));
}

// Tile-local column plan (spmv_lx.cu) for operators held outside a hierarchy - the rank's fine rows in the partitioned
// solver.  *out = NULL (status 0) when the pattern does not localise; the _lx entry points then must not be used.
extern "C" int mwx_spmv_lx_create(int64_t n, const int64_t *indptr, const int32_t *indices, int64_t plan, void **out,
                                  void *stream) {
    if (!out) return MWX_E_BADARG;
    *out = nullptr;
    if (n <= 0 || !indptr || !indices || plan > -8) return MWX_E_BADARG;
    SpmvLX *p = nullptr;
    const int rc = spmv_lx_create(n, indptr, indices, (int)(-plan), &p, as_stream(stream));
    if (rc) return rc;
    *out = p;
    return 0;
}

extern "C" void mwx_spmv_lx_destroy(void *lx) { spmv_lx_destroy((SpmvLX *)lx); }

extern "C" int mwx_spmv_lx(int64_t n, int64_t plan, const int64_t *indptr, const int32_t *indices, const double *values,
                           const void *lx, const double *x, double *y, void *stream) {
    if (n <= 0 || !indptr || !indices || !values || !x || !y) return MWX_E_BADARG;
    return launch_spmv(n, plan, indptr, indices, values, x, y, nullptr, nullptr, nullptr, nullptr, as_stream(stream),
                       (const SpmvLX *)lx);
}

extern "C" int mwx_spmv_axpy_lx(int64_t n, int64_t plan, const int64_t *indptr, const int32_t *indices,
                                const double *values, const void *lx, const double *x, const double *c, int32_t sign,
                                double *y, void *stream) {
    if (n <= 0 || !indptr || !indices || !values || !x || !c || !y || sign == 0) return MWX_E_BADARG;
    return launch_spmv_axpy(n, plan, indptr, indices, values, x, y, c, sign, as_stream(stream), (const SpmvLX *)lx);
}

extern "C" int mwx_smoother_step_lx(int64_t n, int64_t plan, const int64_t *indptr, const int32_t *indices,
                                    const double *values, const void *lx, const double *x_in, double *x_out,
                                    const double *b, const double *dinv, double *d, double c1, double c2, void *stream) {
    if (n <= 0 || !indptr || !indices || !values || !x_in || !x_out || !b || !dinv || !d || x_in == x_out)
        return MWX_E_BADARG;
    return launch_smoother_step(n, plan, indptr, indices, values, x_in, x_out, b, dinv, d, c1, c2, as_stream(stream),
                                (const SpmvLX *)lx);
}

extern "C" int mwx_dcg_spmv_dot_lx(int64_t n, int64_t plan, const int64_t *indptr, const int32_t *indices,
                                   const double *values, const void *lx, const double *x_ext, double *q, double *ws,
                                   void *stream) {
    if (n <= 0 || !indptr || !indices || !values || !x_ext || !q || !ws) return MWX_E_BADARG;
    CgScalars *sc = ws_scalars(ws);
    return launch_spmv(n, plan, indptr, indices, values, x_ext, q, x_ext, &sc->pq, ws_partials(ws), &sc->ticket[0],
                       as_stream(stream), (const SpmvLX *)lx);
}

extern "C" int mwx_residual_dd(int64_t n, const int64_t *indptr, const int32_t *indices, const double *values,
                               const double *x, const double *b, double *r, void *stream) {
    if (n <= 0 || !indptr || !indices || !values || !x || !b || !r) return MWX_E_BADARG;
    k_residual_dd<<<vec_grid(n), VEC_THREADS, 0, as_stream(stream)>>>(n, indptr, indices, values, x, b, r);
    MWX_LAUNCH_CHECK();
    return 0;
}
