// Smoothed-aggregation AMG setup on the device (throughput mode 'sa'; no reference counterpart - the reference's
// pyamg.ruge_stuben_solver hierarchy stays a host build, amg_setup.cpp).
//
// Same scheme as mwx_amg_setup_sa_host (Vanek/Mandel/Brezina: symmetric strength |a_ij| >= theta sqrt(a_ii a_jj) on
// node blocks, aggregates = a root and its strong neighbours plus the leftovers attached to an adjacent aggregate,
// per-aggregate QR of the candidates, one damped-Jacobi smoothing of the tentative prolongator, Galerkin product),
// but every pass is a kernel and nothing leaves HBM:
//   * roots: instead of the host's serial greedy sweep, a maximal independent set of the SQUARE of the strength graph
//     (distance >= 3 between roots - exactly the greedy pass-1 condition) found in O(log n) rounds of "largest key in
//     the 2-neighbourhood wins"; key = (strong degree, hash, index), so vertex DOFs of the Morley element seed the
//     aggregates as on the host.  Deterministic: no atomics decide anything.
//   * sparse products (A'T, AP, R(AP)) : one generic row-wise SpGEMM.  A group of G lanes owns a row of C and a hash
//     table in shared memory; it walks the entries k of A's row IN ORDER and, for each, its lanes add a_ik * B(k,:)
//     into the table (distinct columns -> distinct slots, no atomics on values).  The sum order of every entry is the
//     order of k: bit-reproducible.  Count pass, scan, fill pass; the row is emitted sorted by a rank count.
//     Rows whose table overflows are redone by a one-warp-per-row launch with a 8192-slot table.
//   * transpose: column histogram, scan, scatter with a cursor, then every row is rank-sorted (the scatter order is
//     the only nondeterministic thing and the sort removes it).
// Coarse levels keep the ncand x ncand block structure (rows/columns come in triples for the Morley candidates 1, x, y);
// the strength pass reads blocks straight out of the scalar CSR and checks that structure.
#include <cooperative_groups.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <new>
#include <vector>

#include "amg_dev.cuh"

namespace cg = cooperative_groups;

namespace mwx {
namespace {

#define SA_TRY(expr)            \
    do {                        \
        int _rc = (expr);       \
        if (_rc) return _rc;    \
    } while (0)

template <typename T>
int dev_alloc(T **p, size_t n) {
    return cudaMalloc((void **)p, sizeof(T) * (n ? n : 1)) == cudaSuccess ? 0 : MWX_E_NOMEM;
}

template <typename T>
int read_back(T *host, const T *dev, cudaStream_t st) {
    MWX_CUDA(cudaMemcpyAsync(host, dev, sizeof(T), cudaMemcpyDeviceToHost, st));
    MWX_CUDA(cudaStreamSynchronize(st));
    return 0;
}

// ------------------------------------------------------------------ node graph of strong connections
// Node I = rows bs*I .. bs*I + bs - 1.  Entry p of node row I: column block J, weight |a| (bs = 1) or the Frobenius
// norm of the bs x bs block, read from the scalar CSR whose block structure k_check_blocks has verified.
__device__ __forceinline__ int64_t node_row_len(int bs, const int64_t *Ap, int64_t I) {
    return (Ap[bs * I + 1] - Ap[bs * I]) / bs;
}

__device__ __forceinline__ double node_weight(int bs, const int64_t *Ap, const double *Ax, int64_t I, int64_t p) {
    if (bs == 1) return fabs(Ax[Ap[I] + p]);
    double s = 0.0;
    for (int r = 0; r < bs; ++r) {
        const int64_t o = Ap[bs * I + r] + bs * p;
        for (int c = 0; c < bs; ++c) s += Ax[o + c] * Ax[o + c];
    }
    return sqrt(s);
}

__global__ void k_check_blocks(int64_t nn, int bs, const int64_t *__restrict__ Ap, const int32_t *__restrict__ Aj,
                               int *__restrict__ bad) {
    const int64_t I = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (I >= nn) return;
    const int64_t len = Ap[bs * I + 1] - Ap[bs * I];
    int wrong = (len % bs) != 0;
    for (int r = 1; r < bs && !wrong; ++r) wrong |= (Ap[bs * I + r + 1] - Ap[bs * I + r]) != len;
    if (!wrong) {
        const int64_t nb = len / bs;
        for (int64_t p = 0; p < nb && !wrong; ++p) {
            const int32_t c0 = Aj[Ap[bs * I] + bs * p];
            wrong |= (c0 % bs) != 0;
            for (int r = 0; r < bs; ++r)
                for (int c = 0; c < bs; ++c) wrong |= Aj[Ap[bs * I + r] + bs * p + c] != c0 + c;
        }
    }
    if (wrong) atomicAdd(bad, 1);
}

__global__ void k_node_diag(int64_t nn, int bs, const int64_t *__restrict__ Ap, const int32_t *__restrict__ Aj,
                            const double *__restrict__ Ax, double *__restrict__ dn) {
    const int64_t I = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (I >= nn) return;
    const int64_t nb = node_row_len(bs, Ap, I);
    double d = 0.0;
    for (int64_t p = 0; p < nb; ++p)
        if (Aj[Ap[bs * I] + bs * p] / bs == I) d = node_weight(bs, Ap, Ax, I, p);
    dn[I] = d;
}

// FILL = false: cnt[I] = number of strong off-diagonal neighbours; FILL = true: Sj[Sp[I] ..] = their ids (ascending).
template <bool FILL>
__global__ void k_strength(int64_t nn, int bs, double theta, const int64_t *__restrict__ Ap,
                           const int32_t *__restrict__ Aj, const double *__restrict__ Ax,
                           const double *__restrict__ dn, int32_t *__restrict__ cnt,
                           const int64_t *__restrict__ Sp, int32_t *__restrict__ Sj) {
    const int64_t I = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (I >= nn) return;
    const int64_t nb = node_row_len(bs, Ap, I);
    const double di = dn[I];
    int32_t c = 0;
    int64_t o = FILL ? Sp[I] : 0;
    for (int64_t p = 0; p < nb; ++p) {
        const int32_t J = Aj[Ap[bs * I] + bs * p] / bs;
        if (J == I) continue;
        if (node_weight(bs, Ap, Ax, I, p) >= theta * sqrt(di * dn[J])) {
            if (FILL) Sj[o++] = J;
            ++c;
        }
    }
    if (!FILL) cnt[I] = c;
}

// ------------------------------------------------------------------ roots: maximal independent set of S^2
constexpr unsigned long long KEY_ROOT = ~0ull;

__global__ void k_mis_keys(int64_t nn, const int64_t *__restrict__ Sp, unsigned long long *__restrict__ key,
                           int8_t *__restrict__ state, unsigned long long hash_mask) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nn) return;
    const int64_t deg = Sp[i + 1] - Sp[i];
    uint64_t h = (uint64_t)i * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull;
    h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
    key[i] = ((unsigned long long)(deg < 254 ? deg + 1 : 255) << 56) | ((unsigned long long)(h & hash_mask) << 32) |
             (unsigned long long)(uint32_t)i;
    state[i] = 0;
}

__global__ void k_neighbour_max(int64_t nn, const int64_t *__restrict__ Sp, const int32_t *__restrict__ Sj,
                                const unsigned long long *__restrict__ in, unsigned long long *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nn) return;
    unsigned long long m = in[i];
    for (int64_t k = Sp[i]; k < Sp[i + 1]; ++k) {
        const unsigned long long v = in[Sj[k]];
        m = v > m ? v : m;
    }
    out[i] = m;
}

// state: 0 undecided, 1 root, -1 within distance 2 of a root.  key of a decided node becomes KEY_ROOT / 0.
__global__ void k_mis_decide(int64_t nn, const unsigned long long *__restrict__ m2, unsigned long long *__restrict__ key,
                             int8_t *__restrict__ state, int *__restrict__ undecided) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nn || state[i] != 0) return;
    if (m2[i] == key[i]) { state[i] = 1; key[i] = KEY_ROOT; }
    else if (m2[i] == KEY_ROOT) { state[i] = -1; key[i] = 0ull; }
    else *undecided = 1;
}

__global__ void k_root_flags(int64_t nn, const int8_t *__restrict__ state, int32_t *__restrict__ flag) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nn) flag[i] = state[i] == 1;
}

// pass 1: a root and its strong neighbours (a node cannot touch two roots: they are >= 3 apart)
__global__ void k_agg_pass1(int64_t nn, int bs, const int64_t *__restrict__ Sp, const int32_t *__restrict__ Sj,
                            const int8_t *__restrict__ state, const int32_t *__restrict__ rank,
                            int32_t *__restrict__ agg1, int64_t *__restrict__ roots) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nn) return;
    int32_t a = -1;
    if (state[i] == 1) {
        a = rank[i];
        roots[a] = i * bs;
    } else {
        for (int64_t k = Sp[i]; k < Sp[i + 1]; ++k)
            if (state[Sj[k]] == 1) { a = rank[Sj[k]]; break; }
    }
    agg1[i] = a;
}

// pass 2: the rest joins the pass-1 aggregate of its first (lowest-numbered) strong neighbour that has one
__global__ void k_agg_pass2(int64_t nn, const int64_t *__restrict__ Sp, const int32_t *__restrict__ Sj,
                            const int32_t *__restrict__ agg1, int32_t *__restrict__ agg, int32_t *__restrict__ members_cnt,
                            int *__restrict__ orphan) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nn) return;
    int32_t a = agg1[i];
    for (int64_t k = Sp[i]; a < 0 && k < Sp[i + 1]; ++k) a = agg1[Sj[k]];
    agg[i] = a;
    if (a < 0) atomicAdd(orphan, 1);
    else atomicAdd(&members_cnt[a], 1);
}

__global__ void k_fill_members(int64_t nn, const int32_t *__restrict__ agg, const int64_t *__restrict__ mstart,
                               int32_t *__restrict__ cursor, int32_t *__restrict__ members) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nn) return;
    const int32_t a = agg[i];
    members[mstart[a] + atomicAdd(&cursor[a], 1)] = (int32_t)i;
}

// ------------------------------------------------------------------ tentative prolongator
// One thread per aggregate: sort the member list (the cursor order above is arbitrary), copy the candidate rows into
// T's values, QR by modified Gram-Schmidt applied twice, in place; R goes to the coarse candidates.  T row i holds
// nc entries at columns agg*nc .. agg*nc + nc - 1 (explicit zeros for a candidate that vanished on the aggregate).
__global__ void k_tentative(int64_t na, int bs, int nc, const int64_t *__restrict__ mstart, int32_t *__restrict__ members,
                            const double *__restrict__ B, double *__restrict__ Tx, int32_t *__restrict__ Tj,
                            double *__restrict__ Bc) {
    const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= na) return;
    const int64_t s = mstart[a], e = mstart[a + 1];
    for (int64_t k = s + 1; k < e; ++k) {                     // insertion sort, a few dozen members at most
        const int32_t v = members[k];
        int64_t q = k;
        while (q > s && members[q - 1] > v) { members[q] = members[q - 1]; --q; }
        members[q] = v;
    }
    const int64_t m = (e - s) * bs;
    auto dof = [&](int64_t r) { return (int64_t)members[s + r / bs] * bs + r % bs; };
    for (int64_t r = 0; r < m; ++r) {
        const int64_t g = dof(r);
        for (int c = 0; c < nc; ++c) {
            Tx[g * nc + c] = B[g * nc + c];
            Tj[g * nc + c] = (int32_t)(a * nc + c);
        }
    }
    double *R = Bc + (a * nc) * nc;
    for (int c = 0; c < nc * nc; ++c) R[c] = 0.0;
    for (int c = 0; c < nc; ++c) {
        double norm0 = 0.0;
        for (int64_t r = 0; r < m; ++r) { const double v = Tx[dof(r) * nc + c]; norm0 += v * v; }
        norm0 = sqrt(norm0);
        for (int rep = 0; rep < 2; ++rep)
            for (int p = 0; p < c; ++p) {
                double dot = 0.0;
                for (int64_t r = 0; r < m; ++r) { const int64_t g = dof(r) * nc; dot += Tx[g + p] * Tx[g + c]; }
                for (int64_t r = 0; r < m; ++r) { const int64_t g = dof(r) * nc; Tx[g + c] -= dot * Tx[g + p]; }
                R[p * nc + c] += dot;
            }
        double nrm = 0.0;
        for (int64_t r = 0; r < m; ++r) { const double v = Tx[dof(r) * nc + c]; nrm += v * v; }
        nrm = sqrt(nrm);
        if (nrm > 1e-10 * norm0 && nrm > 0.0) {
            for (int64_t r = 0; r < m; ++r) Tx[dof(r) * nc + c] /= nrm;
            R[c * nc + c] = nrm;
        } else {                                             // dependent on this aggregate: dead coarse DOF
            for (int64_t r = 0; r < m; ++r) Tx[dof(r) * nc + c] = 0.0;
            R[c * nc + c] = 0.0;
        }
    }
}

__global__ void k_iota_ptr(int64_t n, int64_t stride, int64_t *__restrict__ ptr) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n) ptr[i] = i * stride;
}

// ------------------------------------------------------------------ row-wise SpGEMM  C = f(A) B
// mode 0: f(A) = A.   mode 1: f(A) = I - omega D^-1 A (prolongator smoothing; A has its diagonal stored).
struct SpgemmArgs {
    int64_t nrows;
    const int64_t *rows_list;        // nullptr: rows 0..nrows-1; else the overflow rows of the first launch
    const int64_t *Ap; const int32_t *Aj; const double *Ax;
    const int64_t *Bp; const int32_t *Bj; const double *Bx;
    int mode; double omega; const double *dinv;
    int32_t *row_nnz;                // count pass output
    uint8_t *redo;                   // rows left to the big-table launch
    int64_t *redo_rows; int *redo_count;
    const int64_t *Cp; int32_t *Cj; double *Cx;
    int log2_slots;
    int *failed;                     // the big table overflowed too
};

template <int G, bool FILL>
__global__ void __launch_bounds__(256) k_spgemm(SpgemmArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int groups = blockDim.x / G, gid = threadIdx.x / G, lane = threadIdx.x % G;
    const int HS = 1 << a.log2_slots;
    double *vals = reinterpret_cast<double *>(smem) + (size_t)gid * HS;
    int32_t *keys = reinterpret_cast<int32_t *>(smem + sizeof(double) * (size_t)groups * HS) + (size_t)gid * HS;
    auto tile = cg::tiled_partition<G>(cg::this_thread_block());
    for (int64_t w = (int64_t)blockIdx.x * groups + gid; w < a.nrows; w += (int64_t)gridDim.x * groups) {
        const int64_t i = a.rows_list ? a.rows_list[w] : w;
        if (!a.rows_list && FILL && a.redo[i]) continue;
        for (int s = lane; s < HS; s += G) { keys[s] = -1; if (FILL) vals[s] = 0.0; }
        tile.sync();
        int fresh = 0;
        bool full = false;
        const double sc = a.mode == 1 ? -a.omega * a.dinv[i] : 1.0;
        for (int64_t k = a.Ap[i]; k < a.Ap[i + 1]; ++k) {
            const int32_t kk = a.Aj[k];
            double av = 0.0;
            if (FILL) av = sc * a.Ax[k] + ((a.mode == 1 && kk == i) ? 1.0 : 0.0);
            for (int64_t t = a.Bp[kk] + lane; t < a.Bp[kk + 1]; t += G) {
                const int32_t col = a.Bj[t];
                unsigned slot = ((unsigned)col * 0x9E3779B1u) >> (32 - a.log2_slots);
                int probe = 0;
                for (; probe < HS; ++probe) {
                    int32_t cur = keys[slot];
                    if (cur == -1) {
                        cur = atomicCAS(&keys[slot], -1, col);
                        if (cur == -1) { ++fresh; break; }
                    }
                    if (cur == col) break;
                    slot = (slot + 1) & (HS - 1);
                }
                if (probe == HS) full = true;
                else if (FILL) vals[slot] += av * a.Bx[t];
            }
            if (FILL) tile.sync();                         // entries of one k never share a slot; k advances in order
        }
        full = tile.any(full);
        if (!FILL) {
            for (int o = G / 2; o > 0; o >>= 1) fresh += tile.shfl_xor(fresh, o);
            if (lane == 0) {
                if (full) {
                    if (a.rows_list) atomicAdd(a.failed, 1);
                    else {
                        a.redo[i] = 1;
                        a.redo_rows[atomicAdd(a.redo_count, 1)] = i;
                    }
                    a.row_nnz[i] = 0;
                } else {
                    a.row_nnz[i] = fresh;
                }
            }
            tile.sync();
            continue;
        }
        tile.sync();
        int n = 0;                                          // compact the table to its front, in slot order
        for (int base = 0; base < HS; base += G) {
            const int32_t key = keys[base + lane];
            const double v = vals[base + lane];
            const unsigned present = tile.ballot(key >= 0);
            const int pos = n + __popc(present & ((1u << lane) - 1u));
            tile.sync();
            if (key >= 0) { keys[pos] = key; vals[pos] = v; }
            n += __popc(present);
            tile.sync();
        }
        const int64_t o = a.Cp[i];
        for (int e = lane; e < n; e += G) {                 // rank = number of smaller columns
            const int32_t key = keys[e];
            int rank = 0;
            for (int t = 0; t < n; ++t) rank += keys[t] < key;
            a.Cj[o + rank] = key;
            a.Cx[o + rank] = vals[e];
        }
        tile.sync();
    }
}

template <int G>
int spgemm_launch_pair(bool fill, SpgemmArgs &args, int threads, int log2_slots, cudaStream_t st) {
    args.log2_slots = log2_slots;
    const int groups = threads / G;
    const size_t smem = (size_t)groups * ((size_t)1 << log2_slots) * (sizeof(double) + sizeof(int32_t));
    auto kern = fill ? k_spgemm<G, true> : k_spgemm<G, false>;
    MWX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t blocks = (args.nrows + groups - 1) / groups;
    const int64_t cap = (int64_t)sm_count() * 32;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    kern<<<(unsigned)blocks, threads, smem, st>>>(args);
    MWX_LAUNCH_CHECK();
    return 0;
}

int spgemm_launch(int G, bool fill, SpgemmArgs &args, int threads, int log2_slots, cudaStream_t st) {
    switch (G) {
        case 4: return spgemm_launch_pair<4>(fill, args, threads, log2_slots, st);
        case 8: return spgemm_launch_pair<8>(fill, args, threads, log2_slots, st);
        case 16: return spgemm_launch_pair<16>(fill, args, threads, log2_slots, st);
        default: return spgemm_launch_pair<32>(fill, args, threads, log2_slots, st);
    }
}

// C = f(A) B.  C's arrays are cudaMalloc'ed here (owned by the caller afterwards).
int spgemm(const DCsr &A, const DCsr &B, int mode, double omega, const double *dinv, DCsr &C, cudaStream_t st) {
    C = DCsr();
    C.rows = A.rows; C.cols = B.cols;
    const double avg_b = B.rows > 0 ? (double)B.nnz / (double)B.rows : 0.0;
    const int G = avg_b <= 4.0 ? 4 : avg_b <= 12.0 ? 8 : avg_b <= 28.0 ? 16 : 32;
    int log2_slots = 5;
    while ((1 << log2_slots) < 32 * G) ++log2_slots;       // 96 KB of tables per 256-thread CTA whatever G is
    Scratch<int32_t> row_nnz(st);
    Scratch<uint8_t> redo(st);
    Scratch<int64_t> redo_rows(st);
    Scratch<int> counters(st);
    MWX_CUDA(row_nnz.alloc((size_t)A.rows));
    MWX_CUDA(redo.alloc((size_t)A.rows));
    MWX_CUDA(redo_rows.alloc((size_t)A.rows));
    MWX_CUDA(counters.alloc(2));
    MWX_CUDA(cudaMemsetAsync(redo.p, 0, (size_t)A.rows, st));
    MWX_CUDA(cudaMemsetAsync(counters.p, 0, 2 * sizeof(int), st));
    SpgemmArgs args{};
    args.nrows = A.rows; args.rows_list = nullptr;
    args.Ap = A.ptr; args.Aj = A.idx; args.Ax = A.val;
    args.Bp = B.ptr; args.Bj = B.idx; args.Bx = B.val;
    args.mode = mode; args.omega = omega; args.dinv = dinv;
    args.row_nnz = row_nnz.p; args.redo = redo.p; args.redo_rows = redo_rows.p; args.redo_count = counters.p;
    args.failed = counters.p + 1;
    SA_TRY(spgemm_launch(G, false, args, 256, log2_slots, st));
    int host_counters[2] = {0, 0};
    MWX_CUDA(cudaMemcpyAsync(host_counters, counters.p, sizeof(host_counters), cudaMemcpyDeviceToHost, st));
    MWX_CUDA(cudaStreamSynchronize(st));
    const int nredo = host_counters[0];
    SpgemmArgs big = args;
    if (nredo > 0) {
        big.nrows = nredo; big.rows_list = redo_rows.p;
        SA_TRY(spgemm_launch(32, false, big, 32, 13, st));
        MWX_CUDA(cudaMemcpyAsync(host_counters, counters.p, sizeof(host_counters), cudaMemcpyDeviceToHost, st));
        MWX_CUDA(cudaStreamSynchronize(st));
        if (host_counters[1] > 0) return MWX_E_CAPACITY;     // a row of C with more than 8192 entries
    }
    SA_TRY(dev_alloc(&C.ptr, (size_t)A.rows + 1));
    SA_TRY(exclusive_scan_i32_to_i64(row_nnz.p, A.rows, C.ptr, st));
    SA_TRY(read_back(&C.nnz, C.ptr + A.rows, st));
    SA_TRY(dev_alloc(&C.idx, (size_t)C.nnz));
    SA_TRY(dev_alloc(&C.val, (size_t)C.nnz));
    args.Cp = C.ptr; args.Cj = C.idx; args.Cx = C.val;
    SA_TRY(spgemm_launch(G, true, args, 256, log2_slots, st));
    if (nredo > 0) {
        big.Cp = C.ptr; big.Cj = C.idx; big.Cx = C.val;
        SA_TRY(spgemm_launch(32, true, big, 32, 13, st));
    }
    return 0;
}

// ------------------------------------------------------------------ transpose
__global__ void k_column_histogram(int64_t nnz, const int32_t *__restrict__ Pj, int32_t *__restrict__ cnt) {
    const int64_t st = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += st) atomicAdd(&cnt[Pj[k]], 1);
}

__global__ void k_transpose_scatter(int64_t nrows, const int64_t *__restrict__ Pp, const int32_t *__restrict__ Pj,
                                    const double *__restrict__ Px, const int64_t *__restrict__ Rp,
                                    int32_t *__restrict__ cursor, int32_t *__restrict__ Tj, double *__restrict__ Tx) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrows) return;
    for (int64_t k = Pp[i]; k < Pp[i + 1]; ++k) {
        const int32_t c = Pj[k];
        const int64_t o = Rp[c] + atomicAdd(&cursor[c], 1);
        Tj[o] = (int32_t)i;
        Tx[o] = Px[k];
    }
}

// One warp per row: out = the row sorted by column (rank count; the keys of a row are distinct).
constexpr int SORT_STAGE = 1024;
__global__ void __launch_bounds__(256) k_sort_rows(int64_t nrows, const int64_t *__restrict__ Rp,
                                                   const int32_t *__restrict__ Tj, const double *__restrict__ Tx,
                                                   int32_t *__restrict__ Rj, double *__restrict__ Rx) {
    __shared__ int32_t stage[8][SORT_STAGE];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int64_t r = (int64_t)blockIdx.x * 8 + warp; r < nrows; r += (int64_t)gridDim.x * 8) {
        const int64_t s = Rp[r];
        const int len = (int)(Rp[r + 1] - s);
        const bool staged = len <= SORT_STAGE;
        if (staged)
            for (int e = lane; e < len; e += 32) stage[warp][e] = Tj[s + e];
        __syncwarp();
        for (int e = lane; e < len; e += 32) {
            const int32_t key = staged ? stage[warp][e] : Tj[s + e];
            int rank = 0;
            if (staged) for (int t = 0; t < len; ++t) rank += stage[warp][t] < key;
            else for (int t = 0; t < len; ++t) rank += Tj[s + t] < key;
            Rj[s + rank] = key;
            Rx[s + rank] = Tx[s + e];
        }
        __syncwarp();
    }
}

int transpose(const DCsr &P, DCsr &R, cudaStream_t st) {
    R = DCsr();
    R.rows = P.cols; R.cols = P.rows; R.nnz = P.nnz;
    Scratch<int32_t> cnt(st), tj(st);
    Scratch<double> tx(st);
    MWX_CUDA(cnt.alloc((size_t)P.cols));
    MWX_CUDA(tj.alloc((size_t)P.nnz));
    MWX_CUDA(tx.alloc((size_t)P.nnz));
    MWX_CUDA(cudaMemsetAsync(cnt.p, 0, sizeof(int32_t) * (size_t)P.cols, st));
    const int64_t cap = (int64_t)sm_count() * 32;
    k_column_histogram<<<(unsigned)std::min<int64_t>(cap, std::max<int64_t>(1, (P.nnz + 255) / 256)), 256, 0, st>>>(P.nnz, P.idx, cnt.p);
    MWX_LAUNCH_CHECK();
    SA_TRY(dev_alloc(&R.ptr, (size_t)R.rows + 1));
    SA_TRY(exclusive_scan_i32_to_i64(cnt.p, R.rows, R.ptr, st));
    MWX_CUDA(cudaMemsetAsync(cnt.p, 0, sizeof(int32_t) * (size_t)P.cols, st));
    k_transpose_scatter<<<grid_for(P.rows, 256), 256, 0, st>>>(P.rows, P.ptr, P.idx, P.val, R.ptr, cnt.p, tj.p, tx.p);
    MWX_LAUNCH_CHECK();
    SA_TRY(dev_alloc(&R.idx, (size_t)R.nnz));
    SA_TRY(dev_alloc(&R.val, (size_t)R.nnz));
    k_sort_rows<<<(unsigned)std::min<int64_t>(cap, std::max<int64_t>(1, (R.rows + 7) / 8)), 256, 0, st>>>(R.rows, R.ptr, tj.p, tx.p, R.idx, R.val);
    MWX_LAUNCH_CHECK();
    return 0;
}

// Dead coarse DOFs (a candidate that vanished on its aggregate) have an all-zero row and column, diagonal included
// (it is stored: the blocks are structurally full): a unit diagonal keeps D^-1 defined; their right-hand side
// P^T r is zero, so they stay zero.
__global__ void k_fix_dead_diagonal(int64_t n, const int64_t *__restrict__ Ap, const int32_t *__restrict__ Aj,
                                    double *__restrict__ Ax, int *__restrict__ missing) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool found = false;
    for (int64_t k = Ap[i]; k < Ap[i + 1]; ++k)
        if (Aj[k] == i) {
            found = true;
            if (Ax[k] == 0.0) Ax[k] = 1.0;
        }
    if (!found) atomicAdd(missing, 1);
}

}  // namespace
}  // namespace mwx

using namespace mwx;

namespace {

// strength graph + roots + both aggregation passes + member lists; returns na and the scratch arrays.
struct LevelAgg {
    int64_t na = 0;
    int32_t *agg = nullptr, *members = nullptr;
    int64_t *mstart = nullptr, *roots = nullptr;
    cudaStream_t st = nullptr;
    void release() {
        if (agg) cudaFreeAsync(agg, st);
        if (members) cudaFreeAsync(members, st);
        if (mstart) cudaFreeAsync(mstart, st);
        agg = members = nullptr; mstart = nullptr;
    }
};

int build_aggregates(const DCsr &A, int bs, double theta, const double *dn, LevelAgg &out, cudaStream_t st) {
    out.st = st;
    out.release();
    cudaFree(out.roots); out.roots = nullptr;
    const int64_t nn = A.rows / bs;
    const unsigned gn = grid_for(nn, 256);
    Scratch<int32_t> cnt(st), sj(st), rank(st), agg1(st), mcnt(st);
    Scratch<int64_t> sp(st);
    Scratch<unsigned long long> key(st), m1(st), m2(st);
    Scratch<int8_t> state(st);
    Scratch<int> flag(st);
    MWX_CUDA(cnt.alloc((size_t)nn));
    MWX_CUDA(sp.alloc((size_t)nn + 1));
    MWX_CUDA(flag.alloc(2));
    k_strength<false><<<gn, 256, 0, st>>>(nn, bs, theta, A.ptr, A.idx, A.val, dn, cnt.p, nullptr, nullptr);
    MWX_LAUNCH_CHECK();
    SA_TRY(exclusive_scan_i32_to_i64(cnt.p, nn, sp.p, st));
    int64_t snnz = 0;
    SA_TRY(read_back(&snnz, sp.p + nn, st));
    MWX_CUDA(sj.alloc((size_t)snnz));
    k_strength<true><<<gn, 256, 0, st>>>(nn, bs, theta, A.ptr, A.idx, A.val, dn, nullptr, sp.p, sj.p);
    MWX_LAUNCH_CHECK();
    MWX_CUDA(key.alloc((size_t)nn));
    MWX_CUDA(m1.alloc((size_t)nn));
    MWX_CUDA(m2.alloc((size_t)nn));
    MWX_CUDA(state.alloc((size_t)nn));
    static const unsigned long long hash_mask = [] {
        const char *e = std::getenv("MWX_SA_MIS_HASH_BITS");
        const int bits = e ? std::atoi(e) : 24;
        return bits <= 0 ? 0ull : (bits >= 24 ? 0xFFFFFFull : ((1ull << bits) - 1ull));
    }();
    k_mis_keys<<<gn, 256, 0, st>>>(nn, sp.p, key.p, state.p, hash_mask);
    MWX_LAUNCH_CHECK();
    for (int round = 0;; ++round) {
        if (round > 500) return MWX_E_NOCONV;
        MWX_CUDA(cudaMemsetAsync(flag.p, 0, sizeof(int), st));
        k_neighbour_max<<<gn, 256, 0, st>>>(nn, sp.p, sj.p, key.p, m1.p);
        k_neighbour_max<<<gn, 256, 0, st>>>(nn, sp.p, sj.p, m1.p, m2.p);
        k_mis_decide<<<gn, 256, 0, st>>>(nn, m2.p, key.p, state.p, flag.p);
        MWX_LAUNCH_CHECK();
        int undecided = 0;
        SA_TRY(read_back(&undecided, flag.p, st));
        if (!undecided) break;
    }
    MWX_CUDA(rank.alloc((size_t)nn + 1));
    k_root_flags<<<gn, 256, 0, st>>>(nn, state.p, cnt.p);
    MWX_LAUNCH_CHECK();
    SA_TRY(exclusive_scan_i32_to_i32(cnt.p, nn, rank.p, st));
    int32_t na32 = 0;
    SA_TRY(read_back(&na32, rank.p + nn, st));
    out.na = na32;
    SA_TRY(dev_alloc(&out.roots, (size_t)out.na));
    MWX_CUDA(agg1.alloc((size_t)nn));
    k_agg_pass1<<<gn, 256, 0, st>>>(nn, bs, sp.p, sj.p, state.p, rank.p, agg1.p, out.roots);
    MWX_LAUNCH_CHECK();
    MWX_CUDA(cudaMallocAsync((void **)&out.agg, sizeof(int32_t) * (size_t)nn, st));
    MWX_CUDA(mcnt.alloc((size_t)out.na));
    MWX_CUDA(cudaMemsetAsync(mcnt.p, 0, sizeof(int32_t) * (size_t)out.na, st));
    MWX_CUDA(cudaMemsetAsync(flag.p, 0, 2 * sizeof(int), st));
    k_agg_pass2<<<gn, 256, 0, st>>>(nn, sp.p, sj.p, agg1.p, out.agg, mcnt.p, flag.p + 1);
    MWX_LAUNCH_CHECK();
    int orphans = 0;
    SA_TRY(read_back(&orphans, flag.p + 1, st));
    if (orphans) return MWX_E_NOCONV;                       // cannot happen for a symmetric strength graph
    MWX_CUDA(cudaMallocAsync((void **)&out.mstart, sizeof(int64_t) * ((size_t)out.na + 1), st));
    SA_TRY(exclusive_scan_i32_to_i64(mcnt.p, out.na, out.mstart, st));
    MWX_CUDA(cudaMemsetAsync(mcnt.p, 0, sizeof(int32_t) * (size_t)out.na, st));
    MWX_CUDA(cudaMallocAsync((void **)&out.members, sizeof(int32_t) * (size_t)nn, st));
    k_fill_members<<<gn, 256, 0, st>>>(nn, out.agg, out.mstart, mcnt.p, out.members);
    MWX_LAUNCH_CHECK();
    return 0;
}

double seconds_since(cudaStream_t st, double &t_prev) {
    cudaStreamSynchronize(st);
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    const double now = (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
    const double d = now - t_prev;
    t_prev = now;
    return d;
}

}  // namespace

extern "C" int mwx_amg_setup_sa_dev(int64_t n, const int64_t *indptr, const int32_t *indices, const double *values,
                                    const double *candidates, int32_t ncand, int32_t block_size, int32_t first_level,
                                    double theta, int32_t max_levels, int32_t max_coarse, int32_t smoother,
                                    int32_t cheby_degree, mwx_amg_dev_t **out, void *stream) {
    if (n <= 0 || !indptr || !indices || !values || !candidates || !out || max_levels < 1 || ncand < 1 || ncand > 8 ||
        smoother < 1 || smoother > 2 || block_size < 1 || n % block_size || first_level < 0)
        return MWX_E_BADARG;
    cudaStream_t st = as_stream(stream);
    auto *d = new (std::nothrow) mwx_amg_dev();
    if (!d) return MWX_E_NOMEM;
    d->smoother = smoother;
    d->degree = cheby_degree > 0 ? cheby_degree : 2;
    const bool prof = std::getenv("MWX_AMG_PROFILE") != nullptr;
    int rc = 0;
    LevelAgg ag;
    double *B = nullptr;                                      // candidates of the current level (owned from level 1 on)
    auto fail = [&](int code) {
        ag.release();
        cudaFree(ag.roots);
        mwx_amg_destroy_dev(d);
        return code;
    };
    d->lv.emplace_back();
    {
        DCsr &A0 = d->lv.back().A;                            // the caller's arrays, borrowed: no 9 GB copy at level 12
        A0.rows = A0.cols = n;
        A0.ptr = const_cast<int64_t *>(indptr); A0.idx = const_cast<int32_t *>(indices); A0.val = const_cast<double *>(values);
        A0.borrowed = true;
        if ((rc = read_back(&A0.nnz, indptr + n, st))) return fail(rc);
        if ((rc = dev_alloc(&B, (size_t)n * ncand))) return fail(rc);
        if (cudaMemcpyAsync(B, candidates, sizeof(double) * (size_t)n * ncand, cudaMemcpyDeviceToDevice, st) != cudaSuccess)
            return fail(MWX_E_BADARG);
    }
    if (cudaMalloc((void **)&d->sc, sizeof(CgScalars)) != cudaSuccess ||
        cudaMalloc((void **)&d->partials, sizeof(double) * 2 * RED_MAX_BLOCKS) != cudaSuccess) {
        cudaFree(B);
        return fail(MWX_E_NOMEM);
    }
    cudaMemsetAsync(d->sc, 0, sizeof(CgScalars), st);
    int bs = block_size;                                     // > 1: the input is itself a coarse level of this scheme
    Scratch<int> flag(st);
    if (flag.alloc(1) != cudaSuccess) { cudaFree(B); return fail(MWX_E_NOMEM); }
    double tprev = 0.0;
    if (prof) seconds_since(st, tprev);
    while ((int32_t)d->lv.size() < max_levels && d->lv.back().A.rows > max_coarse) {
        const size_t li = d->lv.size() - 1;
        const int64_t nl = d->lv[li].A.rows, nn = nl / bs;
        double t_strength = 0, t_tent = 0, t_smooth = 0, t_rap = 0;
        {
            DCsr &A = d->lv[li].A;
            if (bs > 1) {
                cudaMemsetAsync(flag.p, 0, sizeof(int), st);
                k_check_blocks<<<grid_for(nn, 256), 256, 0, st>>>(nn, bs, A.ptr, A.idx, flag.p);
                int bad = 0;
                if ((rc = read_back(&bad, flag.p, st))) break;
                if (bad) { rc = MWX_E_CAPACITY; break; }
            }
            Scratch<double> dn(st);
            if (dn.alloc((size_t)nn) != cudaSuccess) { rc = MWX_E_NOMEM; break; }
            k_node_diag<<<grid_for(nn, 256), 256, 0, st>>>(nn, bs, A.ptr, A.idx, A.val, dn.p);
            if ((rc = build_aggregates(A, bs, theta, dn.p, ag, st))) break;
            bool stalled = false;
            if ((first_level + (int)li == 0 ? 4 : 2) * ag.na * ncand > nl) {    // too few strong connections: aggregate over the whole pattern
                if ((rc = build_aggregates(A, bs, 0.0, dn.p, ag, st))) break;
                stalled = 2 * ag.na * ncand > nl;
            }
            if (stalled) { ag.release(); cudaFree(ag.roots); ag.roots = nullptr; break; }   // this level is the coarsest
        }
        if (prof) t_strength = seconds_since(st, tprev);
        const int64_t na = ag.na;
        DLevel &L = d->lv[li];
        L.roots = ag.roots; ag.roots = nullptr;
        L.nroots = na; L.coarse_per_root = ncand;
        L.cand = B; L.ncand = ncand; B = nullptr;
        // tentative prolongator and the coarse candidates
        DCsr T;
        T.rows = nl; T.cols = na * ncand; T.nnz = nl * ncand;
        double *Bc = nullptr;
        if ((rc = dev_alloc(&T.ptr, (size_t)nl + 1)) || (rc = dev_alloc(&T.idx, (size_t)T.nnz)) ||
            (rc = dev_alloc(&T.val, (size_t)T.nnz)) || (rc = dev_alloc(&Bc, (size_t)(na * ncand) * ncand))) {
            free_csr(T); cudaFree(Bc);
            break;
        }
        k_iota_ptr<<<grid_for(nl + 1, 256), 256, 0, st>>>(nl, ncand, T.ptr);
        k_tentative<<<grid_for(na, 128), 128, 0, st>>>(na, bs, ncand, ag.mstart, ag.members, L.cand, T.val, T.idx, Bc);
        ag.release();
        B = Bc;
        // rho(D^-1 A) for the prolongator smoothing (and, later, for the Chebyshev bounds of this level)
        {
            const size_t bytes = sizeof(double) * (size_t)nl;
            if (cudaMalloc((void **)&L.dinv, bytes) != cudaSuccess || cudaMalloc((void **)&L.r, bytes) != cudaSuccess ||
                cudaMalloc((void **)&L.x2, bytes) != cudaSuccess) { free_csr(T); rc = MWX_E_NOMEM; break; }
            if ((rc = launch_csr_diagonal(nl, L.A.ptr, L.A.idx, L.A.val, L.dinv, 1, st))) { free_csr(T); break; }
            if (L.A.plan == -1) {
                int tile_rows = 0;
                if ((rc = spmv_plan_rows(nl, L.A.ptr, &tile_rows, st))) { free_csr(T); break; }
                L.A.plan = -(int64_t)tile_rows;
            }
            if ((rc = amg_estimate_rho(d, L, st))) { free_csr(T); break; }
        }
        if (prof) t_tent = seconds_since(st, tprev);
        rc = spgemm(L.A, T, 1, 4.0 / (3.0 * L.rho), L.dinv, L.P, st);
        free_csr(T);
        if (rc) break;
        if ((rc = transpose(L.P, L.R, st))) break;
        if (prof) t_smooth = seconds_since(st, tprev);
        DCsr AP;
        rc = spgemm(L.A, L.P, 0, 0.0, nullptr, AP, st);
        if (rc) { free_csr(AP); break; }
        DLevel next;
        rc = spgemm(L.R, AP, 0, 0.0, nullptr, next.A, st);
        free_csr(AP);
        if (rc) { free_csr(next.A); break; }
        cudaMemsetAsync(flag.p, 0, sizeof(int), st);
        k_fix_dead_diagonal<<<grid_for(next.A.rows, 256), 256, 0, st>>>(next.A.rows, next.A.ptr, next.A.idx, next.A.val, flag.p);
        int missing = 0;
        if ((rc = read_back(&missing, flag.p, st))) { free_csr(next.A); break; }
        if (missing) { free_csr(next.A); rc = MWX_E_CAPACITY; break; }
        if (prof) {
            t_rap = seconds_since(st, tprev);
            std::fprintf(stderr, "[mwx sa dev] level %zu n=%lld -> %lld (%lld aggregates, rho %.3f, nnz P %lld, A_c %lld): "
                                 "aggregate %.3f tentative+rho %.3f smooth+R %.3f galerkin %.3f s\n",
                         li, (long long)nl, (long long)(na * ncand), (long long)na, L.rho, (long long)L.P.nnz,
                         (long long)next.A.nnz, t_strength, t_tent, t_smooth, t_rap);
        }
        bs = ncand;
        d->lv.push_back(next);
    }
    ag.release();
    cudaFree(ag.roots);
    if (rc) { cudaFree(B); return fail(rc); }
    d->lv.back().cand = B;
    d->lv.back().ncand = ncand;
    *out = d;
    return 0;
}

extern "C" int mwx_amg_dev_finalize(mwx_amg_dev_t *d, const double *coarse_pinv_host, void *stream) {
    if (!d || d->lv.empty()) return MWX_E_BADARG;
    return amg_finalize_device_levels(d, coarse_pinv_host, as_stream(stream));
}

extern "C" int mwx_amg_dev_num_levels(const mwx_amg_dev_t *d) { return d ? (int)d->lv.size() : MWX_E_BADARG; }

static const DCsr *dev_matrix(const mwx_amg_dev_t *d, int32_t level, int32_t which) {
    if (!d || level < 0 || level >= (int32_t)d->lv.size() || which < 0 || which > 2) return nullptr;
    const DLevel &L = d->lv[(size_t)level];
    const DCsr &M = which == 0 ? L.A : which == 1 ? L.P : L.R;
    return M.ptr ? &M : nullptr;
}

extern "C" int mwx_amg_dev_level_dims(const mwx_amg_dev_t *d, int32_t level, int32_t which, int64_t *dims) {
    const DCsr *M = dev_matrix(d, level, which);
    if (!M || !dims) return MWX_E_BADARG;
    dims[0] = M->rows; dims[1] = M->cols; dims[2] = M->nnz;
    return 0;
}

extern "C" int mwx_amg_dev_level_copy(const mwx_amg_dev_t *d, int32_t level, int32_t which, int64_t *indptr,
                                      int32_t *indices, double *values, void *stream) {
    const DCsr *M = dev_matrix(d, level, which);
    if (!M || !indptr || !indices || !values) return MWX_E_BADARG;
    cudaStream_t st = as_stream(stream);
    MWX_CUDA(cudaMemcpyAsync(indptr, M->ptr, sizeof(int64_t) * ((size_t)M->rows + 1), cudaMemcpyDefault, st));
    MWX_CUDA(cudaMemcpyAsync(indices, M->idx, sizeof(int32_t) * (size_t)M->nnz, cudaMemcpyDefault, st));
    MWX_CUDA(cudaMemcpyAsync(values, M->val, sizeof(double) * (size_t)M->nnz, cudaMemcpyDefault, st));
    MWX_CUDA(cudaStreamSynchronize(st));
    return 0;
}

extern "C" int mwx_amg_dev_level_aux(const mwx_amg_dev_t *d, int32_t level, int64_t *roots, int64_t *nroots,
                                     int32_t *coarse_per_root, double *candidates, int32_t *ncand, void *stream) {
    if (!d || level < 0 || level >= (int32_t)d->lv.size()) return MWX_E_BADARG;
    const DLevel &L = d->lv[(size_t)level];
    cudaStream_t st = as_stream(stream);
    if (nroots) *nroots = L.nroots;
    if (coarse_per_root) *coarse_per_root = L.coarse_per_root;
    if (ncand) *ncand = L.ncand;
    if (roots && L.roots) MWX_CUDA(cudaMemcpyAsync(roots, L.roots, sizeof(int64_t) * (size_t)L.nroots, cudaMemcpyDefault, st));
    if (candidates && L.cand)
        MWX_CUDA(cudaMemcpyAsync(candidates, L.cand, sizeof(double) * (size_t)L.A.rows * L.ncand, cudaMemcpyDefault, st));
    MWX_CUDA(cudaStreamSynchronize(st));
    return 0;
}
