// Host-side classical Ruge-Stueben AMG setup (C++17 + OpenMP).  Part of libmwx_b200.so.
//
// Replaces  ml = pyamg.ruge_stuben_solver(A)   ref main/example1/utils.py:154
// with pyamg 4.0.0's defaults: classical strength (|a_ij| >= theta max_k|a_ik|,
// theta = 0.25), first-pass RS C/F splitting, direct interpolation, R = P^T, A_c = R A P, at most
// max_levels levels, stop when a level has <= max_coarse rows.
//
// Design notes (this is not pyamg's code layout):
//   * strength and its transpose are built as index-only graphs (no value copies) with a parallel
//     count/scan/fill; interpolation reads A's values through the strength mask directly;
//   * the C/F splitting is the inherently serial part (its tie-breaks define the hierarchy, and the
//     reference's CG iteration counts depend on them), kept as a single-threaded bucket queue with
//     O(1) promote/demote;
//   * the Galerkin product is one fused row-wise pass  A_c(I,:) = sum_i R(I,i) * (A(i,:) P)  with a
//     per-thread sparse accumulator, parallel over coarse rows, output columns sorted.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <numeric>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <utility>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../../include/mwx_b200.h"

namespace {

// std::vector whose resize() leaves new elements uninitialised: the big index/value arrays are always filled right
// after they are sized, in parallel - a serial zero-fill (and its first touch of every page from one thread) would
// cost seconds at 10^9 entries.
template <typename T>
struct DefaultInit : std::allocator<T> {
    template <typename U> struct rebind { using other = DefaultInit<U>; };
    using std::allocator<T>::allocator;
    template <typename U> void construct(U *p) noexcept { ::new (static_cast<void *>(p)) U; }
    template <typename U, typename... Args> void construct(U *p, Args &&...args) {
        ::new (static_cast<void *>(p)) U(std::forward<Args>(args)...);
    }
};
template <typename T> using Vec = std::vector<T, DefaultInit<T>>;

template <typename T>
void parallel_copy(const T *src, size_t n, Vec<T> &dst) {
    dst.resize(n);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i) dst[(size_t)i] = src[(size_t)i];
}

struct Csr {
    int64_t rows = 0, cols = 0;
    Vec<int64_t> ptr;
    Vec<int32_t> idx;
    Vec<double> val;
    int64_t nnz() const { return ptr.empty() ? 0 : ptr.back(); }
};

struct Graph {                       // index-only CSR
    std::vector<int64_t> ptr;
    std::vector<int32_t> idx;
};

struct Level {
    Csr A, P, R;
    std::vector<int32_t> split;      // 1 = C, 0 = F (empty on the coarsest level)
    std::vector<int64_t> roots;      // smoothed aggregation: first DOF of every aggregate's seed node, ascending
    std::vector<double> cand;        //   ... the near-nullspace candidates of this level (rows x ncand, row-major)
    int32_t ncand = 0;
    int32_t coarse_per_root = 0;     //   ... and the number of coarse DOFs per aggregate
};

// ---- strength of connection: off-diagonal graph S (no diagonal) and a per-entry mask on A ----------
void strength_graph(const Csr &A, double theta, Graph &S, std::vector<uint8_t> &strong) {
    const int64_t n = A.rows;
    strong.assign((size_t)A.nnz(), 0);
    std::vector<int64_t> cnt((size_t)n + 1, 0);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        double mx = std::numeric_limits<double>::min();
        for (int64_t k = A.ptr[i]; k < A.ptr[i + 1]; ++k)
            if (A.idx[k] != i) mx = std::max(mx, std::fabs(A.val[k]));
        const double thr = theta * mx;
        int64_t c = 0;
        for (int64_t k = A.ptr[i]; k < A.ptr[i + 1]; ++k)
            if (A.idx[k] != i && std::fabs(A.val[k]) >= thr) { strong[(size_t)k] = 1; ++c; }
        cnt[(size_t)i + 1] = c;
    }
    S.ptr.assign((size_t)n + 1, 0);
    for (int64_t i = 0; i < n; ++i) S.ptr[(size_t)i + 1] = S.ptr[(size_t)i] + cnt[(size_t)i + 1];
    S.idx.resize((size_t)S.ptr[(size_t)n]);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        int64_t o = S.ptr[(size_t)i];
        for (int64_t k = A.ptr[i]; k < A.ptr[i + 1]; ++k)
            if (strong[(size_t)k]) S.idx[(size_t)o++] = A.idx[k];
    }
}

void transpose_graph(const Graph &S, int64_t n, Graph &T) {
    T.ptr.assign((size_t)n + 1, 0);
    for (int32_t j : S.idx) ++T.ptr[(size_t)j + 1];
    for (int64_t i = 0; i < n; ++i) T.ptr[(size_t)i + 1] += T.ptr[(size_t)i];
    T.idx.resize(S.idx.size());
    std::vector<int64_t> cur(T.ptr.begin(), T.ptr.end() - 1);
    for (int64_t i = 0; i < n; ++i)
        for (int64_t k = S.ptr[(size_t)i]; k < S.ptr[(size_t)i + 1]; ++k)
            T.idx[(size_t)cur[(size_t)S.idx[(size_t)k]]++] = (int32_t)i;   // rows ascending inside each column
}

// ---- RS first pass ------------------------------------------------------------------------------
// Nodes sit in one array ordered by weight lambda (number of undecided nodes they influence); the
// node at the very end is the next C point.  promote/demote move a node to the end/front of its
// bucket and shift the bucket boundary, which reproduces pyamg's tie-breaking exactly.
class WeightQueue {
public:
    WeightQueue(int64_t n, const std::vector<int64_t> &lam0) : lam(lam0), start((size_t)n + 1, 0),
                                                               count((size_t)n + 1, 0), at((size_t)n),
                                                               where((size_t)n) {
        for (int64_t i = 0; i < n; ++i) ++count[(size_t)lam[(size_t)i]];
        int64_t c = 0;
        for (int64_t l = 0; l <= n; ++l) { start[(size_t)l] = c; c += count[(size_t)l]; count[(size_t)l] = 0; }
        for (int64_t i = 0; i < n; ++i) {
            const int64_t l = lam[(size_t)i], pos = start[(size_t)l] + count[(size_t)l]++;
            at[(size_t)pos] = i; where[(size_t)i] = pos;
        }
    }
    int64_t node_at(int64_t pos) const { return at[(size_t)pos]; }
    int64_t weight(int64_t i) const { return lam[(size_t)i]; }
    void retire(int64_t i) { --count[(size_t)lam[(size_t)i]]; }
    void promote(int64_t k) {
        const int64_t l = lam[(size_t)k];
        const int64_t dst = start[(size_t)l] + count[(size_t)l] - 1;
        swap_to(k, dst);
        --count[(size_t)l]; ++count[(size_t)l + 1];
        start[(size_t)l + 1] = dst;
        lam[(size_t)k] = l + 1;
    }
    void demote(int64_t j) {
        const int64_t l = lam[(size_t)j];
        const int64_t dst = start[(size_t)l];
        swap_to(j, dst);
        --count[(size_t)l]; ++count[(size_t)l - 1];
        ++start[(size_t)l];
        start[(size_t)l - 1] = start[(size_t)l] - count[(size_t)l - 1];
        lam[(size_t)j] = l - 1;
    }
private:
    void swap_to(int64_t node, int64_t dst) {
        const int64_t src = where[(size_t)node], other = at[(size_t)dst];
        at[(size_t)src] = other; where[(size_t)other] = src;
        at[(size_t)dst] = node; where[(size_t)node] = dst;
    }
    std::vector<int64_t> lam, start, count, at, where;
};

void rs_first_pass(int64_t n, const Graph &S, const Graph &T, std::vector<int32_t> &split) {
    constexpr int32_t F = 0, C = 1, U = 2;
    std::vector<int64_t> lam((size_t)n);
    for (int64_t i = 0; i < n; ++i) lam[(size_t)i] = T.ptr[(size_t)i + 1] - T.ptr[(size_t)i];
    split.assign((size_t)n, U);
    for (int64_t i = 0; i < n; ++i)
        if (lam[(size_t)i] == 0) split[(size_t)i] = F;            // influences nobody
    WeightQueue q(n, lam);
    for (int64_t top = n - 1; top >= 0; --top) {
        const int64_t i = q.node_at(top);
        q.retire(i);
        if (split[(size_t)i] == F) continue;
        split[(size_t)i] = C;
        for (int64_t a = T.ptr[(size_t)i]; a < T.ptr[(size_t)i + 1]; ++a) {
            const int64_t j = T.idx[(size_t)a];
            if (split[(size_t)j] != U) continue;
            split[(size_t)j] = F;
            for (int64_t b = S.ptr[(size_t)j]; b < S.ptr[(size_t)j + 1]; ++b) {
                const int64_t k = S.idx[(size_t)b];
                if (split[(size_t)k] == U && q.weight(k) < n - 1) q.promote(k);
            }
        }
        for (int64_t a = S.ptr[(size_t)i]; a < S.ptr[(size_t)i + 1]; ++a) {
            const int64_t j = S.idx[(size_t)a];
            if (split[(size_t)j] == U && q.weight(j) > 0) q.demote(j);
        }
    }
}

// ---- direct interpolation --------------------------------------------------------------------------
void direct_interpolation(const Csr &A, const std::vector<uint8_t> &strong,
                          const std::vector<int32_t> &split, Csr &P) {
    const int64_t n = A.rows;
    std::vector<int32_t> cid((size_t)n);
    int32_t nc = 0;
    for (int64_t i = 0; i < n; ++i) { cid[(size_t)i] = nc; nc += split[(size_t)i]; }
    P.rows = n; P.cols = nc;
    P.ptr.assign((size_t)n + 1, 0);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        int64_t c = 0;
        if (split[(size_t)i]) c = 1;
        else
            for (int64_t k = A.ptr[i]; k < A.ptr[i + 1]; ++k)
                if (strong[(size_t)k] && split[(size_t)A.idx[k]]) ++c;
        P.ptr[(size_t)i + 1] = c;
    }
    for (int64_t i = 0; i < n; ++i) P.ptr[(size_t)i + 1] += P.ptr[(size_t)i];
    P.idx.resize((size_t)P.ptr[(size_t)n]);
    P.val.resize((size_t)P.ptr[(size_t)n]);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        int64_t o = P.ptr[(size_t)i];
        if (split[(size_t)i]) { P.idx[(size_t)o] = cid[(size_t)i]; P.val[(size_t)o] = 1.0; continue; }
        double strong_neg = 0, strong_pos = 0, all_neg = 0, all_pos = 0, diag = 0;
        for (int64_t k = A.ptr[i]; k < A.ptr[i + 1]; ++k) {
            const double a = A.val[k];
            if (A.idx[k] == i) { diag += a; continue; }
            if (a < 0) all_neg += a; else all_pos += a;
            if (strong[(size_t)k] && split[(size_t)A.idx[k]]) { if (a < 0) strong_neg += a; else strong_pos += a; }
        }
        const double alpha = strong_neg != 0 ? all_neg / strong_neg : 0.0;
        double beta = strong_pos != 0 ? all_pos / strong_pos : 0.0;
        if (strong_pos == 0) { diag += all_pos; beta = 0.0; }
        const double wneg = -alpha / diag, wpos = -beta / diag;
        for (int64_t k = A.ptr[i]; k < A.ptr[i + 1]; ++k)
            if (strong[(size_t)k] && split[(size_t)A.idx[k]]) {
                P.idx[(size_t)o] = cid[(size_t)A.idx[k]];
                P.val[(size_t)o] = (A.val[k] < 0 ? wneg : wpos) * A.val[k];
                ++o;
            }
    }
}

// R = P^T.  Every thread owns a contiguous range of P's columns (= rows of R) and streams twice (count, fill) over the
// row blocks of P whose column span meets its range - sequential reads, no atomics, rows ascending inside every
// column: deterministic for any thread count.  (Prolongators are banded in the solver's numbering: coarse DOFs are
// numbered after the fine DOFs they come from, so a thread skips most blocks.)
void transpose_csr(const Csr &P, Csr &R) {
    R.rows = P.cols; R.cols = P.rows;
    R.ptr.assign((size_t)R.rows + 1, 0);
    R.idx.resize(P.idx.size()); R.val.resize(P.val.size());
    int nthreads = 1;
#ifdef _OPENMP
    nthreads = omp_get_max_threads();
#endif
    if ((int64_t)P.idx.size() < (int64_t)1 << 16) nthreads = 1;
    constexpr int64_t BLK = 4096;
    const int64_t nblk = (P.rows + BLK - 1) / BLK;
    std::vector<int32_t> bmin((size_t)nblk, std::numeric_limits<int32_t>::max()), bmax((size_t)nblk, -1);
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (int64_t b = 0; b < nblk; ++b) {
        const int64_t r1 = std::min(P.rows, (b + 1) * BLK);
        int32_t lo = std::numeric_limits<int32_t>::max(), hi = -1;
        for (int64_t k = P.ptr[(size_t)(b * BLK)]; k < P.ptr[(size_t)r1]; ++k) {
            lo = std::min(lo, P.idx[(size_t)k]);
            hi = std::max(hi, P.idx[(size_t)k]);
        }
        bmin[(size_t)b] = lo; bmax[(size_t)b] = hi;
    }
#pragma omp parallel num_threads(nthreads)
    {
        int t = 0;
#ifdef _OPENMP
        t = omp_get_thread_num();
#endif
        const int32_t c0 = (int32_t)(R.rows * t / nthreads), c1 = (int32_t)(R.rows * (t + 1) / nthreads);
        for (int64_t b = 0; b < nblk; ++b) {
            if (bmax[(size_t)b] < c0 || bmin[(size_t)b] >= c1) continue;
            const int64_t r1 = std::min(P.rows, (b + 1) * BLK);
            for (int64_t k = P.ptr[(size_t)(b * BLK)]; k < P.ptr[(size_t)r1]; ++k) {
                const int32_t j = P.idx[(size_t)k];
                if (j >= c0 && j < c1) ++R.ptr[(size_t)j + 1];
            }
        }
    }
    for (int64_t i = 0; i < R.rows; ++i) R.ptr[(size_t)i + 1] += R.ptr[(size_t)i];
    std::vector<int64_t> cur(R.ptr.begin(), R.ptr.end() - 1);
#pragma omp parallel num_threads(nthreads)
    {
        int t = 0;
#ifdef _OPENMP
        t = omp_get_thread_num();
#endif
        const int32_t c0 = (int32_t)(R.rows * t / nthreads), c1 = (int32_t)(R.rows * (t + 1) / nthreads);
        for (int64_t b = 0; b < nblk; ++b) {
            if (bmax[(size_t)b] < c0 || bmin[(size_t)b] >= c1) continue;
            const int64_t r1 = std::min(P.rows, (b + 1) * BLK);
            for (int64_t i = b * BLK; i < r1; ++i)
                for (int64_t k = P.ptr[(size_t)i]; k < P.ptr[(size_t)i + 1]; ++k) {
                    const int32_t j = P.idx[(size_t)k];
                    if (j >= c0 && j < c1) {
                        const int64_t o = cur[(size_t)j]++;
                        R.idx[(size_t)o] = (int32_t)i; R.val[(size_t)o] = P.val[(size_t)k];
                    }
                }
        }
    }
}

// ---- Galerkin product A_c = R A P, fused, row-parallel ----------------------------------------------
void galerkin(const Csr &R, const Csr &A, const Csr &P, Csr &Ac) {
    // Per coarse row I: w = R(I,:) A first (a small open-addressing table keyed by the fine column - neighbouring fine
    // rows share most of their columns), then A_c(I,:) = w P into the per-thread dense accumulator.  The sums run in a
    // fixed order (R's entries, A's entries, then the table in slot order), so the result is deterministic.
    const int64_t nc = R.rows;
    Ac.rows = Ac.cols = nc;
    Ac.ptr.assign((size_t)nc + 1, 0);
    std::vector<std::vector<int32_t>> row_idx((size_t)nc);
    std::vector<std::vector<double>> row_val((size_t)nc);
#pragma omp parallel
    {
        std::vector<double> acc((size_t)nc, 0.0);
        std::vector<int32_t> mark((size_t)nc, -1), touched;
        std::vector<int32_t> hkey;
        std::vector<double> hval;
        std::vector<int32_t> hlist;                          // occupied slots in insertion order
#pragma omp for schedule(dynamic, 256)
        for (int64_t I = 0; I < nc; ++I) {
            touched.clear();
            int64_t cap = 0;
            for (int64_t a = R.ptr[(size_t)I]; a < R.ptr[(size_t)I + 1]; ++a) {
                const int64_t i = R.idx[(size_t)a];
                cap += A.ptr[(size_t)i + 1] - A.ptr[(size_t)i];
            }
            size_t hs = 64;
            while ((int64_t)hs < 2 * cap) hs <<= 1;
            if (hkey.size() < hs) { hkey.assign(hs, -1); hval.assign(hs, 0.0); }
            const size_t mask = hs - 1;                      // table is all -1 on entry (cleared below after use)
            hlist.clear();
            for (int64_t a = R.ptr[(size_t)I]; a < R.ptr[(size_t)I + 1]; ++a) {
                const int64_t i = R.idx[(size_t)a];
                const double rv = R.val[(size_t)a];
                for (int64_t b = A.ptr[(size_t)i]; b < A.ptr[(size_t)i + 1]; ++b) {
                    const int32_t j = A.idx[(size_t)b];
                    size_t slot = ((size_t)(uint32_t)j * 2654435761u) & mask;
                    while (hkey[slot] != -1 && hkey[slot] != j) slot = (slot + 1) & mask;
                    if (hkey[slot] == -1) { hkey[slot] = j; hval[slot] = 0.0; hlist.push_back((int32_t)slot); }
                    hval[slot] += rv * A.val[(size_t)b];
                }
            }
            for (const int32_t slot : hlist) {
                const int64_t j = hkey[(size_t)slot];
                const double w = hval[(size_t)slot];
                hkey[(size_t)slot] = -1;
                for (int64_t c = P.ptr[(size_t)j]; c < P.ptr[(size_t)j + 1]; ++c) {
                    const int32_t J = P.idx[(size_t)c];
                    if (mark[(size_t)J] != (int32_t)I) { mark[(size_t)J] = (int32_t)I; acc[(size_t)J] = 0.0; touched.push_back(J); }
                    acc[(size_t)J] += w * P.val[(size_t)c];
                }
            }
            std::sort(touched.begin(), touched.end());
            auto &ri = row_idx[(size_t)I];
            auto &rv2 = row_val[(size_t)I];
            ri.assign(touched.begin(), touched.end());
            rv2.resize(touched.size());
            for (size_t k = 0; k < touched.size(); ++k) rv2[k] = acc[(size_t)touched[k]];
        }
    }
    for (int64_t I = 0; I < nc; ++I) Ac.ptr[(size_t)I + 1] = Ac.ptr[(size_t)I] + (int64_t)row_idx[(size_t)I].size();
    Ac.idx.resize((size_t)Ac.ptr[(size_t)nc]); Ac.val.resize((size_t)Ac.ptr[(size_t)nc]);
#pragma omp parallel for schedule(static)
    for (int64_t I = 0; I < nc; ++I) {
        std::copy(row_idx[(size_t)I].begin(), row_idx[(size_t)I].end(), Ac.idx.begin() + Ac.ptr[(size_t)I]);
        std::copy(row_val[(size_t)I].begin(), row_val[(size_t)I].end(), Ac.val.begin() + Ac.ptr[(size_t)I]);
    }
}

const Csr *pick(const mwx_amg_host *h, int32_t level, int32_t which);

}  // namespace

struct mwx_amg_host {
    std::vector<Level> levels;
};

namespace {
const Csr *pick(const mwx_amg_host *h, int32_t level, int32_t which) {
    if (!h || level < 0 || level >= (int32_t)h->levels.size()) return nullptr;
    const Level &L = h->levels[(size_t)level];
    if (which == 0) return &L.A;
    if (level + 1 >= (int32_t)h->levels.size()) return nullptr;
    return which == 1 ? &L.P : (which == 2 ? &L.R : nullptr);
}
}  // namespace

extern "C" int mwx_amg_setup_host(int64_t n, const int64_t *indptr, const int32_t *indices,
                                  const double *values, double theta, int32_t max_levels,
                                  int32_t max_coarse, mwx_amg_host_t **out) {
    if (n <= 0 || !indptr || !indices || !values || !out || max_levels < 1) return MWX_E_BADARG;
    auto *h = new (std::nothrow) mwx_amg_host();
    if (!h) return MWX_E_NOMEM;
    try {
        h->levels.emplace_back();
        Csr &A0 = h->levels.back().A;
        A0.rows = A0.cols = n;
        parallel_copy(indptr, (size_t)n + 1, A0.ptr);
        parallel_copy(indices, (size_t)indptr[n], A0.idx);
        parallel_copy(values, (size_t)indptr[n], A0.val);
        const bool prof = std::getenv("MWX_AMG_PROFILE") != nullptr;
        auto now = [] { return std::chrono::steady_clock::now(); };
        auto secs = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
            return std::chrono::duration<double>(b - a).count();
        };
        while ((int32_t)h->levels.size() < max_levels && h->levels.back().A.rows > max_coarse) {
            Level &L = h->levels.back();
            Graph S, T;
            std::vector<uint8_t> strong;
            const auto t0 = now();
            strength_graph(L.A, theta, S, strong);
            const auto t1 = now();
            transpose_graph(S, L.A.rows, T);
            const auto t2 = now();
            rs_first_pass(L.A.rows, S, T, L.split);
            const auto t3 = now();
            direct_interpolation(L.A, strong, L.split, L.P);
            if (L.P.cols == 0) break;                      // nothing left to coarsen
            transpose_csr(L.P, L.R);
            const auto t4 = now();
            Level next;
            galerkin(L.R, L.A, L.P, next.A);
            const auto t5 = now();
            if (prof)
                std::fprintf(stderr, "[mwx amg] level %zu n=%lld: strength %.3f transpose %.3f split %.3f interp+R %.3f galerkin %.3f s\n",
                             h->levels.size() - 1, (long long)L.A.rows, secs(t0, t1), secs(t1, t2), secs(t2, t3),
                             secs(t3, t4), secs(t4, t5));
            h->levels.push_back(std::move(next));
        }
    } catch (const std::bad_alloc &) {
        delete h;
        return MWX_E_NOMEM;
    }
    *out = h;
    return 0;
}

// ---- smoothed aggregation (throughput mode; no reference counterpart) ---------------------------------
// The reference's pyamg.ruge_stuben_solver hierarchy (above) interpolates with direct interpolation, which only
// reproduces constants; for the fourth-order part eps^2 (hess u : hess v) the near-nullspace is span{1, x, y}
// and CG on that hierarchy needs ~240 iterations at 67 M DOFs.  Smoothed aggregation with the Morley interpolants
// of 1, x, y as candidates (Vanek/Mandel/Brezina: strength |a_ij| >= theta sqrt(a_ii a_jj) on node blocks, greedy
// three-pass aggregation, per-aggregate QR of the candidates, one damped-Jacobi smoothing of the tentative
// prolongator, Galerkin coarse operator) cuts that about threefold at the same operator complexity.
namespace {

// Node graph of strong connections; nodes are blocks of `bs` consecutive DOFs (bs = 1 on the finest level).
void sa_strength(const Csr &A, int bs, double theta, Graph &S) {
    const int64_t nn = A.rows / bs;
    if (bs == 1) {                                           // finest level: entries are the node graph; count + fill
        std::vector<double> d((size_t)nn, 0.0);
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < nn; ++i)
            for (int64_t k = A.ptr[(size_t)i]; k < A.ptr[(size_t)i + 1]; ++k)
                if (A.idx[(size_t)k] == i) d[(size_t)i] = std::fabs(A.val[(size_t)k]);
        auto strong = [&](int64_t i, int64_t k) {
            const int64_t j = A.idx[(size_t)k];
            return j != i && std::fabs(A.val[(size_t)k]) >= theta * std::sqrt(d[(size_t)i] * d[(size_t)j]);
        };
        S.ptr.assign((size_t)nn + 1, 0);
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < nn; ++i) {
            int64_t c = 0;
            for (int64_t k = A.ptr[(size_t)i]; k < A.ptr[(size_t)i + 1]; ++k) c += strong(i, k);
            S.ptr[(size_t)i + 1] = c;
        }
        for (int64_t i = 0; i < nn; ++i) S.ptr[(size_t)i + 1] += S.ptr[(size_t)i];
        S.idx.resize((size_t)S.ptr[(size_t)nn]);
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < nn; ++i) {
            int64_t o = S.ptr[(size_t)i];
            for (int64_t k = A.ptr[(size_t)i]; k < A.ptr[(size_t)i + 1]; ++k)
                if (strong(i, k)) S.idx[(size_t)o++] = A.idx[(size_t)k];
        }
        return;
    }
    std::vector<double> diag((size_t)nn, 0.0);
    std::vector<std::vector<std::pair<int32_t, double>>> rows((size_t)nn);
#pragma omp parallel
    {
        std::vector<std::pair<int32_t, double>> ent;
#pragma omp for schedule(dynamic, 512)
        for (int64_t I = 0; I < nn; ++I) {
            ent.clear();
            for (int64_t i = I * bs; i < (I + 1) * bs; ++i)
                for (int64_t k = A.ptr[(size_t)i]; k < A.ptr[(size_t)i + 1]; ++k)
                    ent.emplace_back((int32_t)(A.idx[(size_t)k] / bs), bs == 1 ? std::fabs(A.val[(size_t)k])
                                                                               : A.val[(size_t)k] * A.val[(size_t)k]);
            std::sort(ent.begin(), ent.end(), [](const auto &a, const auto &b) { return a.first < b.first; });
            auto &out = rows[(size_t)I];
            for (size_t k = 0; k < ent.size();) {
                size_t e = k;
                double v = 0.0;
                while (e < ent.size() && ent[e].first == ent[k].first) v += ent[e++].second;
                if (bs > 1) v = std::sqrt(v);
                if (ent[k].first == (int32_t)I) diag[(size_t)I] = v; else out.emplace_back(ent[k].first, v);
                k = e;
            }
        }
    }
    S.ptr.assign((size_t)nn + 1, 0);
#pragma omp parallel for schedule(static)
    for (int64_t I = 0; I < nn; ++I) {
        auto &r = rows[(size_t)I];
        size_t w = 0;
        for (size_t k = 0; k < r.size(); ++k)
            if (r[k].second >= theta * std::sqrt(diag[(size_t)I] * diag[(size_t)r[k].first])) r[w++] = r[k];
        r.resize(w);
    }
    for (int64_t I = 0; I < nn; ++I) S.ptr[(size_t)I + 1] = S.ptr[(size_t)I] + (int64_t)rows[(size_t)I].size();
    S.idx.resize((size_t)S.ptr[(size_t)nn]);
#pragma omp parallel for schedule(static)
    for (int64_t I = 0; I < nn; ++I) {
        int64_t o = S.ptr[(size_t)I];
        for (const auto &e : rows[(size_t)I]) S.idx[(size_t)o++] = e.first;
    }
}

// Greedy aggregation (serial: the visiting order defines the aggregates).  Nodes are visited by descending number of
// strong neighbours (ties by index), so the result does not depend on how the DOFs happen to be numbered - vertex
// DOFs of the Morley element (18 neighbours) seed the aggregates whether they come first (skfem numbering) or
// interleaved with the facet DOFs (the solver's Hilbert numbering).  Returns the aggregate count.
int64_t sa_aggregate(const Graph &S, std::vector<int32_t> &agg) {
    const int64_t nn = (int64_t)S.ptr.size() - 1;
    agg.assign((size_t)nn, -1);
    int64_t maxdeg = 0;
    for (int64_t i = 0; i < nn; ++i) maxdeg = std::max(maxdeg, S.ptr[(size_t)i + 1] - S.ptr[(size_t)i]);
    std::vector<int64_t> bucket((size_t)maxdeg + 2, 0);
    for (int64_t i = 0; i < nn; ++i) ++bucket[(size_t)(maxdeg - (S.ptr[(size_t)i + 1] - S.ptr[(size_t)i])) + 1];
    for (int64_t d = 0; d <= maxdeg; ++d) bucket[(size_t)d + 1] += bucket[(size_t)d];
    std::vector<int32_t> order((size_t)nn);
    for (int64_t i = 0; i < nn; ++i) order[(size_t)bucket[(size_t)(maxdeg - (S.ptr[(size_t)i + 1] - S.ptr[(size_t)i]))]++] = (int32_t)i;
    int64_t na = 0;
    for (int64_t o = 0; o < nn; ++o) {                       // pass 1: a free node whose strong neighbours are all free
        const int64_t i = order[(size_t)o];
        if (agg[(size_t)i] >= 0 || S.ptr[(size_t)i] == S.ptr[(size_t)i + 1]) continue;
        bool free_nb = true;
        for (int64_t k = S.ptr[(size_t)i]; k < S.ptr[(size_t)i + 1]; ++k)
            if (agg[(size_t)S.idx[(size_t)k]] >= 0) { free_nb = false; break; }
        if (!free_nb) continue;
        agg[(size_t)i] = (int32_t)na;
        for (int64_t k = S.ptr[(size_t)i]; k < S.ptr[(size_t)i + 1]; ++k) agg[(size_t)S.idx[(size_t)k]] = (int32_t)na;
        ++na;
    }
    std::vector<int32_t> first(agg);                         // pass 2 attaches to pass-1 aggregates only
    for (int64_t i = 0; i < nn; ++i) {
        if (first[(size_t)i] >= 0) continue;
        for (int64_t k = S.ptr[(size_t)i]; k < S.ptr[(size_t)i + 1]; ++k)
            if (first[(size_t)S.idx[(size_t)k]] >= 0) { agg[(size_t)i] = first[(size_t)S.idx[(size_t)k]]; break; }
    }
    for (int64_t o = 0; o < nn; ++o) {                       // pass 3: leftovers seed aggregates of their own
        const int64_t i = order[(size_t)o];
        if (agg[(size_t)i] >= 0) continue;
        agg[(size_t)i] = (int32_t)na;
        for (int64_t k = S.ptr[(size_t)i]; k < S.ptr[(size_t)i + 1]; ++k)
            if (agg[(size_t)S.idx[(size_t)k]] < 0) agg[(size_t)S.idx[(size_t)k]] = (int32_t)na;
        ++na;
    }
    return na;
}

// Renumber the aggregates by their seed node, so that coarse numbering is monotone in the fine numbering (a row-block
// partition of the fine level then induces contiguous coarse blocks).  seeds[a] = seed node of (new) aggregate a.
void sa_order_by_seed(std::vector<int32_t> &agg, int64_t na, std::vector<int64_t> &seeds) {
    const int64_t nn = (int64_t)agg.size();
    std::vector<int64_t> seed((size_t)na, -1);
    for (int64_t i = 0; i < nn; ++i)
        if (seed[(size_t)agg[(size_t)i]] < 0) seed[(size_t)agg[(size_t)i]] = i;      // lowest member stands in for the seed
    std::vector<int32_t> order((size_t)na);
    std::iota(order.begin(), order.end(), 0);
    std::sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return seed[(size_t)a] < seed[(size_t)b]; });
    std::vector<int32_t> newid((size_t)na);
    seeds.resize((size_t)na);
    for (int64_t k = 0; k < na; ++k) { newid[(size_t)order[(size_t)k]] = (int32_t)k; seeds[(size_t)k] = seed[(size_t)order[(size_t)k]]; }
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nn; ++i) agg[(size_t)i] = newid[(size_t)agg[(size_t)i]];
}

// Tentative prolongator: per aggregate, QR (modified Gram-Schmidt, twice) of the candidate rows; T gets Q, the
// coarse candidates get R.  Row i of T holds nc entries at columns agg*nc .. agg*nc + nc - 1.
void sa_tentative(int64_t n, int bs, int nc, const std::vector<int32_t> &agg, int64_t na, const std::vector<double> &B,
                  Csr &T, std::vector<double> &Bc) {
    const int64_t nn = n / bs;
    std::vector<int64_t> start((size_t)na + 1, 0);
    for (int64_t i = 0; i < nn; ++i) ++start[(size_t)agg[(size_t)i] + 1];
    for (int64_t a = 0; a < na; ++a) start[(size_t)a + 1] += start[(size_t)a];
    std::vector<int32_t> members((size_t)nn);
    {
        std::vector<int64_t> cur(start.begin(), start.end() - 1);
        for (int64_t i = 0; i < nn; ++i) members[(size_t)cur[(size_t)agg[(size_t)i]]++] = (int32_t)i;
    }
    T.rows = n; T.cols = na * nc;
    T.ptr.resize((size_t)n + 1);
    T.idx.resize((size_t)n * nc);
    T.val.assign((size_t)n * nc, 0.0);
    Bc.assign((size_t)(na * nc) * nc, 0.0);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i <= n; ++i) T.ptr[(size_t)i] = i * nc;
#pragma omp parallel
    {
        std::vector<double> q;
        std::vector<int64_t> dofs;
#pragma omp for schedule(dynamic, 256)
        for (int64_t a = 0; a < na; ++a) {
            dofs.clear();
            for (int64_t m = start[(size_t)a]; m < start[(size_t)a + 1]; ++m)
                for (int d = 0; d < bs; ++d) dofs.push_back((int64_t)members[(size_t)m] * bs + d);
            const size_t m = dofs.size();
            q.assign(m * nc, 0.0);
            for (size_t r = 0; r < m; ++r)
                for (int c = 0; c < nc; ++c) q[r * nc + c] = B[(size_t)dofs[r] * nc + c];
            double *R = &Bc[(size_t)(a * nc) * nc];
            for (int c = 0; c < nc; ++c) {
                double norm0 = 0.0;
                for (size_t r = 0; r < m; ++r) norm0 += q[r * nc + c] * q[r * nc + c];
                norm0 = std::sqrt(norm0);
                for (int rep = 0; rep < 2; ++rep)
                    for (int p2 = 0; p2 < c; ++p2) {
                        double dot = 0.0;
                        for (size_t r = 0; r < m; ++r) dot += q[r * nc + p2] * q[r * nc + c];
                        for (size_t r = 0; r < m; ++r) q[r * nc + c] -= dot * q[r * nc + p2];
                        R[p2 * nc + c] += dot;
                    }
                double nrm = 0.0;
                for (size_t r = 0; r < m; ++r) nrm += q[r * nc + c] * q[r * nc + c];
                nrm = std::sqrt(nrm);
                if (nrm > 1e-10 * norm0 && nrm > 0.0) {
                    for (size_t r = 0; r < m; ++r) q[r * nc + c] /= nrm;
                    R[c * nc + c] = nrm;
                } else {                                    // candidate dependent on this aggregate: dead coarse DOF
                    for (size_t r = 0; r < m; ++r) q[r * nc + c] = 0.0;
                    R[c * nc + c] = 0.0;
                }
            }
            for (size_t r = 0; r < m; ++r)
                for (int c = 0; c < nc; ++c) {
                    T.idx[(size_t)dofs[r] * nc + c] = (int32_t)(a * nc + c);
                    T.val[(size_t)dofs[r] * nc + c] = q[r * nc + c];
                }
        }
    }
}

double sa_rho_dinv_a(const Csr &A, const std::vector<double> &dinv, int iters) {
    const int64_t n = A.rows;
    std::vector<double> v((size_t)n), w((size_t)n);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        uint64_t z = (uint64_t)i * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull;
        z ^= z >> 29; z *= 0xBF58476D1CE4E5B9ull; z ^= z >> 32;
        v[(size_t)i] = (double)(z >> 11) / 9007199254740992.0 - 0.5;
    }
    double lam = 1.0;
    constexpr int64_t CH = 8192;                             // fixed chunks: the sums do not depend on the thread count
    const int64_t nch = (n + CH - 1) / CH;
    std::vector<double> pv((size_t)nch), pw((size_t)nch);
    for (int it = 0; it < iters; ++it) {
#pragma omp parallel for schedule(static)
        for (int64_t c = 0; c < nch; ++c) {
            double nv = 0.0, nw = 0.0;
            const int64_t i1 = std::min(n, (c + 1) * CH);
            for (int64_t i = c * CH; i < i1; ++i) {
                double sum = 0.0;
                for (int64_t k = A.ptr[(size_t)i]; k < A.ptr[(size_t)i + 1]; ++k) sum += A.val[(size_t)k] * v[(size_t)A.idx[(size_t)k]];
                sum *= dinv[(size_t)i];
                w[(size_t)i] = sum;
                nw += sum * sum;
                nv += v[(size_t)i] * v[(size_t)i];
            }
            pv[(size_t)c] = nv; pw[(size_t)c] = nw;
        }
        double nv = 0.0, nw = 0.0;
        for (int64_t c = 0; c < nch; ++c) { nv += pv[(size_t)c]; nw += pw[(size_t)c]; }
        lam = std::sqrt(nw / nv);
        const double sc = 1.0 / std::sqrt(nw);
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; ++i) v[(size_t)i] = w[(size_t)i] * sc;
    }
    return lam;
}

// P = (I - omega D^-1 A) T, columns sorted.  Every thread takes one contiguous block of rows and appends to its own
// buffers (one pass, no per-row allocations); the blocks are then laid end to end.
void sa_smooth_prolongator(const Csr &A, const Csr &T, const std::vector<double> &dinv, double omega, Csr &P) {
    const auto tt0 = std::chrono::steady_clock::now();
    const int64_t n = A.rows;
    P.rows = n; P.cols = T.cols;
    P.ptr.assign((size_t)n + 1, 0);
    int nthreads = 1;
#ifdef _OPENMP
    nthreads = omp_get_max_threads();
#endif
    std::vector<std::vector<int32_t>> bidx((size_t)nthreads);
    std::vector<std::vector<double>> bval((size_t)nthreads);
    std::vector<int64_t> first((size_t)nthreads + 1, 0);
    for (int t = 0; t <= nthreads; ++t) first[(size_t)t] = n * t / nthreads;
#pragma omp parallel num_threads(nthreads)
    {
        int t = 0;
#ifdef _OPENMP
        t = omp_get_thread_num();
#endif
        // T has exactly nc entries per row, at the columns of the row's aggregate: accumulate per aggregate
        const int nc = (int)(T.ptr[1] - T.ptr[0]);
        std::vector<int32_t> ids, ord;
        std::vector<double> blk;
        std::vector<int32_t> oi;             // thread-local (the shared vector headers would false-share on every push)
        std::vector<double> ov;
        oi.reserve((size_t)(first[(size_t)t + 1] - first[(size_t)t]) * nc * 4);      // ~4 aggregates per row
        ov.reserve((size_t)(first[(size_t)t + 1] - first[(size_t)t]) * nc * 4);
        for (int64_t i = first[(size_t)t]; i < first[(size_t)t + 1]; ++i) {
            ids.clear(); blk.clear();
            auto add = [&](int64_t j, double a) {
                const int32_t g = T.idx[(size_t)j * nc] / nc;
                size_t pos = 0;
                while (pos < ids.size() && ids[pos] != g) ++pos;
                if (pos == ids.size()) { ids.push_back(g); blk.insert(blk.end(), (size_t)nc, 0.0); }
                for (int c = 0; c < nc; ++c) blk[pos * nc + c] += a * T.val[(size_t)j * nc + c];
            };
            const double sc = -omega * dinv[(size_t)i];
            for (int64_t k = A.ptr[(size_t)i]; k < A.ptr[(size_t)i + 1]; ++k) add(A.idx[(size_t)k], sc * A.val[(size_t)k]);
            add(i, 1.0);
            ord.resize(ids.size());
            for (size_t k = 0; k < ids.size(); ++k) ord[k] = (int32_t)k;
            std::sort(ord.begin(), ord.end(), [&](int32_t a, int32_t b) { return ids[(size_t)a] < ids[(size_t)b]; });
            for (const int32_t k : ord)
                for (int c = 0; c < nc; ++c) { oi.push_back(ids[(size_t)k] * nc + c); ov.push_back(blk[(size_t)k * nc + c]); }
            P.ptr[(size_t)i + 1] = (int64_t)ids.size() * nc;
        }
        bidx[(size_t)t].swap(oi);
        bval[(size_t)t].swap(ov);
    }
    const auto tt1 = std::chrono::steady_clock::now();
    for (int64_t i = 0; i < n; ++i) P.ptr[(size_t)i + 1] += P.ptr[(size_t)i];
    P.idx.resize((size_t)P.ptr[(size_t)n]); P.val.resize((size_t)P.ptr[(size_t)n]);
#pragma omp parallel for schedule(static, 1) num_threads(nthreads)
    for (int t = 0; t < nthreads; ++t) {
        const int64_t o = P.ptr[(size_t)first[(size_t)t]];
        std::copy(bidx[(size_t)t].begin(), bidx[(size_t)t].end(), P.idx.begin() + o);
        std::copy(bval[(size_t)t].begin(), bval[(size_t)t].end(), P.val.begin() + o);
    }
    const auto tt2 = std::chrono::steady_clock::now();
    if (std::getenv("MWX_AMG_PROFILE"))
        std::fprintf(stderr, "[mwx sa]     smooth: rows %.3f assemble %.3f (threads %d)\n", std::chrono::duration<double>(tt1 - tt0).count(),
                     std::chrono::duration<double>(tt2 - tt1).count(), nthreads);
}

// Dead coarse DOFs (a candidate that vanished on its aggregate) leave empty rows/columns: give them a unit
// diagonal so that the smoothers' D^-1 exists; their right-hand side P^T r is zero, so they stay zero.
void sa_fix_dead_rows(Csr &A) {
    const int64_t n = A.rows;
    std::vector<uint8_t> dead((size_t)n, 0);
    int64_t ndead = 0;
    for (int64_t i = 0; i < n; ++i) {
        double d = 0.0;
        for (int64_t k = A.ptr[(size_t)i]; k < A.ptr[(size_t)i + 1]; ++k)
            if (A.idx[(size_t)k] == i) d = A.val[(size_t)k];
        if (d == 0.0) { dead[(size_t)i] = 1; ++ndead; }
    }
    if (!ndead) return;
    Csr B;
    B.rows = B.cols = n;
    B.ptr.assign((size_t)n + 1, 0);
    for (int64_t i = 0; i < n; ++i) {
        bool has = false;
        for (int64_t k = A.ptr[(size_t)i]; k < A.ptr[(size_t)i + 1]; ++k) has |= A.idx[(size_t)k] == i;
        B.ptr[(size_t)i + 1] = B.ptr[(size_t)i] + (A.ptr[(size_t)i + 1] - A.ptr[(size_t)i]) + ((dead[(size_t)i] && !has) ? 1 : 0);
    }
    B.idx.resize((size_t)B.ptr[(size_t)n]); B.val.resize((size_t)B.ptr[(size_t)n]);
    for (int64_t i = 0; i < n; ++i) {
        int64_t o = B.ptr[(size_t)i];
        bool placed = !dead[(size_t)i];
        for (int64_t k = A.ptr[(size_t)i]; k < A.ptr[(size_t)i + 1]; ++k) {
            const int32_t j = A.idx[(size_t)k];
            if (!placed && j > i) { B.idx[(size_t)o] = (int32_t)i; B.val[(size_t)o] = 1.0; ++o; placed = true; }
            B.idx[(size_t)o] = j;
            B.val[(size_t)o] = (dead[(size_t)i] && j == i) ? 1.0 : A.val[(size_t)k];
            if (j == i) placed = true;
            ++o;
        }
        if (!placed) { B.idx[(size_t)o] = (int32_t)i; B.val[(size_t)o] = 1.0; ++o; }
    }
    A = std::move(B);
}

}  // namespace

extern "C" int mwx_amg_setup_sa_host(int64_t n, const int64_t *indptr, const int32_t *indices, const double *values,
                                     const double *candidates, int32_t ncand, int32_t block_size, int32_t first_level,
                                     double theta, int32_t max_levels, int32_t max_coarse, mwx_amg_host_t **out) {
    if (n <= 0 || !indptr || !indices || !values || !candidates || !out || max_levels < 1 || ncand < 1 || ncand > 8 ||
        block_size < 1 || n % block_size || first_level < 0)
        return MWX_E_BADARG;
    auto *h = new (std::nothrow) mwx_amg_host();
    if (!h) return MWX_E_NOMEM;
    try {
        h->levels.emplace_back();
        Csr &A0 = h->levels.back().A;
        A0.rows = A0.cols = n;
        parallel_copy(indptr, (size_t)n + 1, A0.ptr);
        parallel_copy(indices, (size_t)indptr[n], A0.idx);
        parallel_copy(values, (size_t)indptr[n], A0.val);
        std::vector<double> B(candidates, candidates + (size_t)n * ncand);
        int bs = block_size;                                 // > 1: the input is itself a coarse level of this scheme
        const bool prof = std::getenv("MWX_AMG_PROFILE") != nullptr;
        auto now = [] { return std::chrono::steady_clock::now(); };
        auto secs = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
            return std::chrono::duration<double>(b - a).count();
        };
        while ((int32_t)h->levels.size() < max_levels && h->levels.back().A.rows > max_coarse) {
            Level &L = h->levels.back();
            const int64_t nl = L.A.rows;
            const auto t0 = now();
            Graph S;
            // The same threshold on every level (MWX_SA_DECAY < 1 makes it shrink per level, Vanek-style): measured on
            // this operator family, a constant 0.1 keeps the coarse aggregates small enough to need 25 % fewer
            // iterations than halving it per level, for 6 % more operator complexity.
            static const double decay = [] { const char *e = std::getenv("MWX_SA_DECAY"); return e ? std::atof(e) : 1.0; }();
            const int lev = first_level + (int)h->levels.size() - 1;
            sa_strength(L.A, bs, theta * std::pow(decay, (double)lev), S);
            const auto t1 = now();
            std::vector<int32_t> agg;
            int64_t na = sa_aggregate(S, agg);
            if ((lev == 0 ? 4 : 2) * na * ncand > nl) {    // too few strong connections at this threshold (the fine level
                sa_strength(L.A, bs, 0.0, S);              // would shrink less than fourfold): aggregate over the whole pattern
                na = sa_aggregate(S, agg);
                if (2 * na * ncand > nl) break;            // coarsening has stalled: this level is the coarsest
            }
            std::vector<int64_t> seeds;
            sa_order_by_seed(agg, na, seeds);
            L.roots.resize((size_t)na);
            for (int64_t a = 0; a < na; ++a) L.roots[(size_t)a] = seeds[(size_t)a] * bs;
            L.coarse_per_root = ncand;
            const auto t2 = now();
            Csr T;
            std::vector<double> Bc;
            sa_tentative(nl, bs, ncand, agg, na, B, T, Bc);
            std::vector<double> dinv((size_t)nl, 1.0);
#pragma omp parallel for schedule(static)
            for (int64_t i = 0; i < nl; ++i)
                for (int64_t k = L.A.ptr[(size_t)i]; k < L.A.ptr[(size_t)i + 1]; ++k)
                    if (L.A.idx[(size_t)k] == i && L.A.val[(size_t)k] != 0.0) dinv[(size_t)i] = 1.0 / L.A.val[(size_t)k];
            const double rho = sa_rho_dinv_a(L.A, dinv, 15);
            const auto t3 = now();
            sa_smooth_prolongator(L.A, T, dinv, 4.0 / (3.0 * rho), L.P);
            const auto t3b = now();
            transpose_csr(L.P, L.R);
            const auto t4 = now();
            if (prof) std::fprintf(stderr, "[mwx sa]   smooth %.3f transpose %.3f\n", secs(t3, t3b), secs(t3b, t4));
            Level next;
            galerkin(L.R, L.A, L.P, next.A);
            sa_fix_dead_rows(next.A);
            const auto t5 = now();
            if (prof)
                std::fprintf(stderr, "[mwx sa] level %zu n=%lld -> %lld (%lld aggregates, rho %.3f): strength %.3f aggregate %.3f "
                                     "tentative+rho %.3f smooth+R %.3f galerkin %.3f s\n",
                             h->levels.size() - 1, (long long)nl, (long long)(na * ncand), (long long)na, rho, secs(t0, t1),
                             secs(t1, t2), secs(t2, t3), secs(t3, t4), secs(t4, t5));
            L.split.clear();
            L.cand.swap(B);
            L.ncand = ncand;
            B.swap(Bc);
            bs = ncand;
            h->levels.push_back(std::move(next));
        }
        h->levels.back().cand.swap(B);
        h->levels.back().ncand = ncand;
    } catch (const std::bad_alloc &) {
        delete h;
        return MWX_E_NOMEM;
    }
    *out = h;
    return 0;
}

extern "C" int mwx_set_host_threads(int32_t n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

extern "C" int mwx_amg_num_levels(const mwx_amg_host_t *h) { return h ? (int)h->levels.size() : MWX_E_BADARG; }

extern "C" int mwx_amg_level_dims(const mwx_amg_host_t *h, int32_t level, int32_t which, int64_t *dims) {
    const Csr *M = pick(h, level, which);
    if (!M || !dims) return MWX_E_BADARG;
    dims[0] = M->rows; dims[1] = M->cols; dims[2] = M->nnz();
    return 0;
}

extern "C" int mwx_amg_level_copy(const mwx_amg_host_t *h, int32_t level, int32_t which, int64_t *indptr,
                                  int32_t *indices, double *values) {
    const Csr *M = pick(h, level, which);
    if (!M || !indptr || !indices || !values) return MWX_E_BADARG;
    std::memcpy(indptr, M->ptr.data(), sizeof(int64_t) * M->ptr.size());
    std::memcpy(indices, M->idx.data(), sizeof(int32_t) * M->idx.size());
    std::memcpy(values, M->val.data(), sizeof(double) * M->val.size());
    return 0;
}

extern "C" int mwx_amg_level_view(const mwx_amg_host_t *h, int32_t level, int32_t which, const int64_t **indptr,
                                  const int32_t **indices, const double **values) {
    const Csr *M = pick(h, level, which);
    if (!M || !indptr || !indices || !values) return MWX_E_BADARG;
    *indptr = M->ptr.data();
    *indices = M->idx.data();
    *values = M->val.data();
    return 0;
}

extern "C" int mwx_amg_level_candidates(const mwx_amg_host_t *h, int32_t level, double *candidates, int32_t *ncand) {
    if (!h || !ncand || level < 0 || level >= (int32_t)h->levels.size()) return MWX_E_BADARG;
    const Level &L = h->levels[(size_t)level];
    *ncand = L.ncand;
    if (candidates && !L.cand.empty()) std::memcpy(candidates, L.cand.data(), sizeof(double) * L.cand.size());
    return 0;
}

extern "C" int mwx_amg_level_roots(const mwx_amg_host_t *h, int32_t level, int64_t *roots, int32_t *coarse_per_root) {
    if (!h || !coarse_per_root || level < 0 || level + 1 >= (int32_t)h->levels.size()) return MWX_E_BADARG;
    const Level &L = h->levels[(size_t)level];
    *coarse_per_root = L.coarse_per_root;
    if (roots && !L.roots.empty()) std::memcpy(roots, L.roots.data(), sizeof(int64_t) * L.roots.size());
    return 0;
}

extern "C" int mwx_amg_level_splitting(const mwx_amg_host_t *h, int32_t level, int32_t *split) {
    if (!h || !split || level < 0 || level + 1 >= (int32_t)h->levels.size()) return MWX_E_BADARG;
    const auto &s = h->levels[(size_t)level].split;
    if (s.empty()) return MWX_E_BADARG;              // smoothed-aggregation level: see mwx_amg_level_roots
    std::memcpy(split, s.data(), sizeof(int32_t) * s.size());
    return 0;
}

extern "C" void mwx_amg_destroy_host(mwx_amg_host_t *h) { delete h; }

// Wavefronts of the lexicographic Gauss-Seidel dependency DAG.  backward != 0 gives the DAG of the
// reverse sweep (row i waits for its neighbours j > i); both are needed for a symmetric sweep.
static int gs_wavefronts(int64_t n, const int64_t *ptr, const int32_t *idx, int backward,
                         int32_t *order, int64_t *wave_ptr, int64_t *nwaves) {
    std::vector<int32_t> lev((size_t)n, 0);
    int32_t maxlev = 0;
    if (!backward) {
        for (int64_t i = 0; i < n; ++i) {
            int32_t l = 0;
            for (int64_t k = ptr[i]; k < ptr[i + 1]; ++k)
                if (idx[k] < i) l = std::max(l, lev[(size_t)idx[k]] + 1);
            lev[(size_t)i] = l; maxlev = std::max(maxlev, l);
        }
    } else {
        for (int64_t i = n - 1; i >= 0; --i) {
            int32_t l = 0;
            for (int64_t k = ptr[i]; k < ptr[i + 1]; ++k)
                if (idx[k] > i) l = std::max(l, lev[(size_t)idx[k]] + 1);
            lev[(size_t)i] = l; maxlev = std::max(maxlev, l);
        }
    }
    const int64_t nw = (int64_t)maxlev + 1;
    std::vector<int64_t> cnt((size_t)nw + 1, 0);
    for (int64_t i = 0; i < n; ++i) ++cnt[(size_t)lev[(size_t)i] + 1];
    for (int64_t w = 0; w < nw; ++w) cnt[(size_t)w + 1] += cnt[(size_t)w];
    std::memcpy(wave_ptr, cnt.data(), sizeof(int64_t) * ((size_t)nw + 1));
    std::vector<int64_t> cur(cnt.begin(), cnt.end() - 1);
    for (int64_t i = 0; i < n; ++i) order[(size_t)cur[(size_t)lev[(size_t)i]]++] = (int32_t)i;
    *nwaves = nw;
    return 0;
}

extern "C" int mwx_gs_wavefronts_host(int64_t n, const int64_t *indptr, const int32_t *indices,
                                      int32_t *order, int64_t *wave_ptr, int64_t *nwaves) {
    if (n <= 0 || !indptr || !indices || !order || !wave_ptr || !nwaves) return MWX_E_BADARG;
    return gs_wavefronts(n, indptr, indices, 0, order, wave_ptr, nwaves);
}

extern "C" int mwx_gs_wavefronts_backward_host(int64_t n, const int64_t *indptr, const int32_t *indices,
                                               int32_t *order, int64_t *wave_ptr, int64_t *nwaves) {
    if (n <= 0 || !indptr || !indices || !order || !wave_ptr || !nwaves) return MWX_E_BADARG;
    return gs_wavefronts(n, indptr, indices, 1, order, wave_ptr, nwaves);
}
