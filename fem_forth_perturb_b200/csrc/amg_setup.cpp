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
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../../include/mwx_b200.h"

namespace {

struct Csr {
    int64_t rows = 0, cols = 0;
    std::vector<int64_t> ptr;
    std::vector<int32_t> idx;
    std::vector<double> val;
    int64_t nnz() const { return ptr.empty() ? 0 : ptr.back(); }
};

struct Graph {                       // index-only CSR
    std::vector<int64_t> ptr;
    std::vector<int32_t> idx;
};

struct Level {
    Csr A, P, R;
    std::vector<int32_t> split;      // 1 = C, 0 = F (empty on the coarsest level)
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

void transpose_csr(const Csr &P, Csr &R) {
    R.rows = P.cols; R.cols = P.rows;
    R.ptr.assign((size_t)R.rows + 1, 0);
    for (int32_t j : P.idx) ++R.ptr[(size_t)j + 1];
    for (int64_t i = 0; i < R.rows; ++i) R.ptr[(size_t)i + 1] += R.ptr[(size_t)i];
    R.idx.resize(P.idx.size()); R.val.resize(P.val.size());
    std::vector<int64_t> cur(R.ptr.begin(), R.ptr.end() - 1);
    for (int64_t i = 0; i < P.rows; ++i)
        for (int64_t k = P.ptr[(size_t)i]; k < P.ptr[(size_t)i + 1]; ++k) {
            const int64_t o = cur[(size_t)P.idx[(size_t)k]]++;
            R.idx[(size_t)o] = (int32_t)i; R.val[(size_t)o] = P.val[(size_t)k];
        }
}

// ---- Galerkin product A_c = R A P, fused, row-parallel ----------------------------------------------
void galerkin(const Csr &R, const Csr &A, const Csr &P, Csr &Ac) {
    const int64_t nc = R.rows;
    Ac.rows = Ac.cols = nc;
    Ac.ptr.assign((size_t)nc + 1, 0);
    std::vector<std::vector<int32_t>> row_idx((size_t)nc);
    std::vector<std::vector<double>> row_val((size_t)nc);
#pragma omp parallel
    {
        std::vector<double> acc((size_t)nc, 0.0);
        std::vector<int32_t> mark((size_t)nc, -1), touched;
#pragma omp for schedule(dynamic, 256)
        for (int64_t I = 0; I < nc; ++I) {
            touched.clear();
            for (int64_t a = R.ptr[(size_t)I]; a < R.ptr[(size_t)I + 1]; ++a) {
                const int64_t i = R.idx[(size_t)a];
                const double rv = R.val[(size_t)a];
                for (int64_t b = A.ptr[(size_t)i]; b < A.ptr[(size_t)i + 1]; ++b) {
                    const int64_t j = A.idx[(size_t)b];
                    const double rav = rv * A.val[(size_t)b];
                    for (int64_t c = P.ptr[(size_t)j]; c < P.ptr[(size_t)j + 1]; ++c) {
                        const int32_t J = P.idx[(size_t)c];
                        if (mark[(size_t)J] != (int32_t)I) { mark[(size_t)J] = (int32_t)I; acc[(size_t)J] = 0.0; touched.push_back(J); }
                        acc[(size_t)J] += rav * P.val[(size_t)c];
                    }
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
        A0.ptr.assign(indptr, indptr + n + 1);
        A0.idx.assign(indices, indices + indptr[n]);
        A0.val.assign(values, values + indptr[n]);
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

extern "C" int mwx_amg_level_splitting(const mwx_amg_host_t *h, int32_t level, int32_t *split) {
    if (!h || !split || level < 0 || level + 1 >= (int32_t)h->levels.size()) return MWX_E_BADARG;
    const auto &s = h->levels[(size_t)level].split;
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
