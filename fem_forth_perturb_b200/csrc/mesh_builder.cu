// Device mesh -> facets -> DOFs -> incidence builder (setup side of hot path A).
//
// Reproduces the NUMBERING of scikit-fem 2.1.1's MeshTri (reached from ref main/example1/run.py:508,512):
// facets are the lexicographically sorted unique (min,max) vertex pairs, t2f follows local edges
// (0,1),(1,2),(0,2), red refinement emits children in four blocks.  skfem gets there with a global
// np.unique over 3*nt rows; here nothing is sorted globally: a vertex->triangle incidence (count /
// scan / fill) makes every remaining step a small per-vertex problem that one thread solves in
// registers / local memory, so the whole builder is a handful of streaming passes.
#include "common.cuh"

namespace mwx {
namespace {

constexpr int TPB = 256;
constexpr int MAX_NB = 64;      // distinct higher-numbered neighbours per vertex (capacity)

__global__ void k_sort_columns(int64_t nt, int32_t *__restrict__ t) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nt) return;
    int32_t a = t[e], b = t[nt + e], c = t[2 * nt + e], x;
    if (a > b) { x = a; a = b; b = x; }
    if (b > c) { x = b; b = c; c = x; }
    if (a > b) { x = a; a = b; b = x; }
    t[e] = a; t[nt + e] = b; t[2 * nt + e] = c;
}

__global__ void k_inc_count(int64_t total, const int32_t *__restrict__ conn, int32_t *__restrict__ cnt) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total) atomicAdd(&cnt[conn[i]], 1);
}

__global__ void k_inc_fill(int64_t nt, int32_t ndpe, const int32_t *__restrict__ conn,
                           const int64_t *__restrict__ ptr, int32_t *__restrict__ cursor,
                           int32_t *__restrict__ idx) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nt * ndpe) return;
    const int32_t l = (int32_t)(i / nt);
    const int64_t e = i - (int64_t)l * nt;
    const int32_t v = conn[i];
    const int32_t pos = atomicAdd(&cursor[v], 1);
    idx[ptr[v] + pos] = (int32_t)((e << 3) | l);
}

// The atomic fill leaves each list in arbitrary order; sorting it restores determinism.
__global__ void k_inc_sort(int64_t nent, const int64_t *__restrict__ ptr, int32_t *__restrict__ idx) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nent) return;
    const int64_t s = ptr[v], e = ptr[v + 1];
    for (int64_t i = s + 1; i < e; ++i) {
        const int32_t key = idx[i];
        int64_t j = i - 1;
        while (j >= s && idx[j] > key) { idx[j + 1] = idx[j]; --j; }
        idx[j + 1] = key;
    }
}

// Sorted unique higher-numbered neighbours of vertex a.  Columns of t are ascending, so only local
// positions 0 and 1 of a triangle see higher neighbours.  Returns the count (or -1 on overflow).
__device__ __forceinline__ int higher_neighbours(int64_t a, int64_t nt, const int32_t *__restrict__ t,
                                                 const int64_t *__restrict__ ptr,
                                                 const int32_t *__restrict__ idx, int32_t *nb) {
    int n = 0;
    for (int64_t q = ptr[a]; q < ptr[a + 1]; ++q) {
        const int32_t code = idx[q];
        const int64_t e = code >> 3;
        const int l = code & 7;
        if (l == 2) continue;
        int32_t cand[2];
        int nc = 0;
        if (l == 0) cand[nc++] = t[nt + e];
        cand[nc++] = t[2 * nt + e];
        for (int c = 0; c < nc; ++c) {
            const int32_t b = cand[c];
            int pos = 0;
            while (pos < n && nb[pos] < b) ++pos;
            if (pos < n && nb[pos] == b) continue;
            if (n == MAX_NB) return -1;
            for (int k = n; k > pos; --k) nb[k] = nb[k - 1];
            nb[pos] = b;
            ++n;
        }
    }
    return n;
}

__global__ void k_facet_count(int64_t nv, int64_t nt, const int32_t *__restrict__ t,
                              const int64_t *__restrict__ ptr, const int32_t *__restrict__ idx,
                              int32_t *__restrict__ cnt, int32_t *__restrict__ err) {
    const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= nv) return;
    int32_t nb[MAX_NB];
    const int n = higher_neighbours(a, nt, t, ptr, idx, nb);
    if (n < 0) { atomicExch(err, 1); cnt[a] = 0; return; }
    cnt[a] = n;
}

__global__ void k_facet_fill(int64_t nv, int64_t nt, int64_t nf, const int32_t *__restrict__ t,
                             const int64_t *__restrict__ ptr, const int32_t *__restrict__ idx,
                             const int64_t *__restrict__ fstart, int32_t *__restrict__ facets) {
    const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= nv) return;
    int32_t nb[MAX_NB];
    const int n = higher_neighbours(a, nt, t, ptr, idx, nb);
    const int64_t f0 = fstart[a];
    for (int k = 0; k < n; ++k) {
        facets[f0 + k] = (int32_t)a;
        facets[nf + f0 + k] = nb[k];
    }
}

__device__ __forceinline__ int32_t facet_of(int32_t a, int32_t b, int64_t nf,
                                            const int64_t *__restrict__ fstart,
                                            const int32_t *__restrict__ facets) {
    int64_t f = fstart[a];
    const int64_t fe = fstart[a + 1];
    while (f < fe && facets[nf + f] != b) ++f;
    return (int32_t)f;
}

__global__ void k_t2f(int64_t nt, int64_t nf, const int32_t *__restrict__ t,
                      const int64_t *__restrict__ fstart, const int32_t *__restrict__ facets,
                      int32_t *__restrict__ t2f) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nt) return;
    const int32_t v0 = t[e], v1 = t[nt + e], v2 = t[2 * nt + e];
    t2f[e] = facet_of(v0, v1, nf, fstart, facets);
    t2f[nt + e] = facet_of(v1, v2, nf, fstart, facets);
    t2f[2 * nt + e] = facet_of(v0, v2, nf, fstart, facets);
}

__global__ void k_refine_points(int64_t nv, int64_t nf, const double2 *__restrict__ p,
                                const int32_t *__restrict__ facets, double2 *__restrict__ pn) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nv) {
        pn[i] = p[i];
    } else if (i < nv + nf) {
        const int64_t f = i - nv;
        const double2 a = p[facets[f]], b = p[facets[nf + f]];
        pn[i] = make_double2(0.5 * (a.x + b.x), 0.5 * (a.y + b.y));
    }
}

__device__ __forceinline__ void store_sorted(int32_t *__restrict__ tn, int64_t ntn, int64_t col,
                                             int32_t a, int32_t b, int32_t c) {
    int32_t x;
    if (a > b) { x = a; a = b; b = x; }
    if (b > c) { x = b; b = c; c = x; }
    if (a > b) { x = a; a = b; b = x; }
    tn[col] = a; tn[ntn + col] = b; tn[2 * ntn + col] = c;
}

__global__ void k_refine_tris(int64_t nv, int64_t nt, const int32_t *__restrict__ t,
                              const int32_t *__restrict__ t2f, int32_t *__restrict__ tn) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nt) return;
    const int64_t ntn = 4 * nt;
    const int32_t v0 = t[e], v1 = t[nt + e], v2 = t[2 * nt + e];
    const int32_t m01 = (int32_t)nv + t2f[e], m12 = (int32_t)nv + t2f[nt + e], m02 = (int32_t)nv + t2f[2 * nt + e];
    store_sorted(tn, ntn, e, v0, m01, m02);
    store_sorted(tn, ntn, nt + e, v1, m01, m12);
    store_sorted(tn, ntn, 2 * nt + e, v2, m02, m12);
    store_sorted(tn, ntn, 3 * nt + e, m01, m12, m02);
}

__global__ void k_element_dofs(int64_t nv, int64_t nt, const int32_t *__restrict__ t,
                               const int32_t *__restrict__ t2f, int32_t *__restrict__ ed) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 3 * nt) return;
    ed[i] = t[i];
    ed[3 * nt + i] = (int32_t)nv + t2f[i];
}

}  // namespace
}  // namespace mwx

using namespace mwx;

extern "C" int mwx_version(void) { return 100; }

extern "C" int mwx_mesh_sort_columns(int64_t nt, int32_t *t, void *stream) {
    if (nt < 0 || (nt && !t)) return MWX_E_BADARG;
    if (nt == 0) return 0;
    k_sort_columns<<<grid_for(nt, TPB), TPB, 0, as_stream(stream)>>>(nt, t);
    MWX_LAUNCH_CHECK();
    return 0;
}

extern "C" int mwx_incidence(int64_t nent, int64_t nt, int32_t ndpe, const int32_t *conn,
                             int64_t *inc_ptr, int32_t *inc_idx, void *stream) {
    if (nent < 0 || nt < 0 || ndpe < 1 || ndpe > 8 || !inc_ptr) return MWX_E_BADARG;
    if (nt >= ((int64_t)1 << 28)) return MWX_E_CAPACITY;     // (e << 3 | l) must fit in int32
    cudaStream_t st = as_stream(stream);
    Scratch<int32_t> cnt(st);
    MWX_CUDA(cnt.alloc((size_t)nent));
    MWX_CUDA(cudaMemsetAsync(cnt.p, 0, sizeof(int32_t) * (size_t)(nent ? nent : 1), st));
    const int64_t total = nt * ndpe;
    if (total) {
        k_inc_count<<<grid_for(total, TPB), TPB, 0, st>>>(total, conn, cnt.p);
        MWX_LAUNCH_CHECK();
    }
    int rc = exclusive_scan_i32_to_i64(cnt.p, nent, inc_ptr, st);
    if (rc) return rc;
    if (total) {
        MWX_CUDA(cudaMemsetAsync(cnt.p, 0, sizeof(int32_t) * (size_t)nent, st));
        k_inc_fill<<<grid_for(total, TPB), TPB, 0, st>>>(nt, ndpe, conn, inc_ptr, cnt.p, inc_idx);
        MWX_LAUNCH_CHECK();
        k_inc_sort<<<grid_for(nent, TPB), TPB, 0, st>>>(nent, inc_ptr, inc_idx);
        MWX_LAUNCH_CHECK();
    }
    return 0;
}

extern "C" int mwx_mesh_facets_count(int64_t nv, int64_t nt, const int32_t *t, const int64_t *v2t_ptr,
                                     const int32_t *v2t_idx, int64_t *fstart, int64_t *nf_host,
                                     void *stream) {
    if (nv <= 0 || nt <= 0 || !t || !v2t_ptr || !v2t_idx || !fstart || !nf_host) return MWX_E_BADARG;
    cudaStream_t st = as_stream(stream);
    Scratch<int32_t> cnt(st), err(st);
    MWX_CUDA(cnt.alloc((size_t)nv));
    MWX_CUDA(err.alloc(1));
    MWX_CUDA(cudaMemsetAsync(err.p, 0, sizeof(int32_t), st));
    k_facet_count<<<grid_for(nv, 128), 128, 0, st>>>(nv, nt, t, v2t_ptr, v2t_idx, cnt.p, err.p);
    MWX_LAUNCH_CHECK();
    int rc = exclusive_scan_i32_to_i64(cnt.p, nv, fstart, st);
    if (rc) return rc;
    int32_t herr = 0;
    MWX_CUDA(cudaMemcpyAsync(&herr, err.p, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    MWX_CUDA(cudaMemcpyAsync(nf_host, fstart + nv, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    MWX_CUDA(cudaStreamSynchronize(st));
    return herr ? MWX_E_CAPACITY : 0;
}

extern "C" int mwx_mesh_facets_fill(int64_t nv, int64_t nt, const int32_t *t, const int64_t *v2t_ptr,
                                    const int32_t *v2t_idx, const int64_t *fstart, int64_t nf,
                                    int32_t *facets, int32_t *t2f, void *stream) {
    if (nv <= 0 || nt <= 0 || nf <= 0 || !facets || !t2f) return MWX_E_BADARG;
    cudaStream_t st = as_stream(stream);
    k_facet_fill<<<grid_for(nv, 128), 128, 0, st>>>(nv, nt, nf, t, v2t_ptr, v2t_idx, fstart, facets);
    MWX_LAUNCH_CHECK();
    k_t2f<<<grid_for(nt, TPB), TPB, 0, st>>>(nt, nf, t, fstart, facets, t2f);
    MWX_LAUNCH_CHECK();
    return 0;
}

extern "C" int mwx_mesh_refine(int64_t nv, int64_t nt, int64_t nf, const double *pxy, const int32_t *t,
                               const int32_t *facets, const int32_t *t2f, double *pxy_new,
                               int32_t *t_new, void *stream) {
    if (nv <= 0 || nt <= 0 || nf <= 0 || !pxy || !t || !facets || !t2f || !pxy_new || !t_new)
        return MWX_E_BADARG;
    if (nv + nf >= ((int64_t)1 << 31) || 4 * nt >= ((int64_t)1 << 28)) return MWX_E_CAPACITY;
    cudaStream_t st = as_stream(stream);
    k_refine_points<<<grid_for(nv + nf, TPB), TPB, 0, st>>>(nv, nf, (const double2 *)pxy, facets,
                                                           (double2 *)pxy_new);
    MWX_LAUNCH_CHECK();
    k_refine_tris<<<grid_for(nt, TPB), TPB, 0, st>>>(nv, nt, t, t2f, t_new);
    MWX_LAUNCH_CHECK();
    return 0;
}

extern "C" int mwx_element_dofs(int64_t nv, int64_t nt, const int32_t *t, const int32_t *t2f,
                                int32_t *edofs, void *stream) {
    if (nt <= 0 || !t || !t2f || !edofs) return MWX_E_BADARG;
    k_element_dofs<<<grid_for(3 * nt, TPB), TPB, 0, as_stream(stream)>>>(nv, nt, t, t2f, edofs);
    MWX_LAUNCH_CHECK();
    return 0;
}

extern "C" int mwx_facet_element_count(int64_t nf, int64_t nt, const int32_t *t2f, int32_t *facet_count,
                                       void *stream) {
    if (nf <= 0 || nt <= 0 || !t2f || !facet_count) return MWX_E_BADARG;
    cudaStream_t st = as_stream(stream);
    MWX_CUDA(cudaMemsetAsync(facet_count, 0, sizeof(int32_t) * (size_t)nf, st));
    k_inc_count<<<grid_for(3 * nt, TPB), TPB, 0, st>>>(3 * nt, t2f, facet_count);
    MWX_LAUNCH_CHECK();
    return 0;
}
