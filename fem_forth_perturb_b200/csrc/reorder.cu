// Locality renumbering for the throughput (fast-mode) solver.
//
// skfem's DOF numbering (vertices level by level, then facets in lexicographic vertex order) is what the
// user-visible CSR must carry, but it is not cache-local: on the 67 M-DOF mesh the x-gathers of one SpMV
// pull 5 GB of extra sectors and every gather is its own L1 wavefront.  The fast-mode solver therefore runs
// on P A P^T with P = order of the DOF locations along a Hilbert curve, and permutes b / x at the boundary.
// (Parity mode keeps skfem's order: lexicographic Gauss-Seidel depends on it.)  Nothing here exists in the
// reference; it is internal to solver_iter_mgcg(mode='chebyshev'|'jacobi') / solver_iter_pcg(reorder=True).
#include "common.cuh"

namespace mwx {
namespace {

constexpr int HILBERT_BITS = 24;

__global__ void k_hilbert_keys(int64_t n, const double2 *__restrict__ xy, double x0, double y0, double scale,
                               int64_t *__restrict__ keys) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double2 p = xy[i];
    const double maxq = (double)((1u << HILBERT_BITS) - 1);
    double qx = (p.x - x0) * scale, qy = (p.y - y0) * scale;
    qx = qx < 0 ? 0 : (qx > maxq ? maxq : qx);
    qy = qy < 0 ? 0 : (qy > maxq ? maxq : qy);
    uint32_t x = (uint32_t)qx, y = (uint32_t)qy;
    // xy -> distance along the Hilbert curve of order HILBERT_BITS
    uint64_t d = 0;
    for (uint32_t s = 1u << (HILBERT_BITS - 1); s > 0; s >>= 1) {
        const uint32_t rx = (x & s) ? 1u : 0u, ry = (y & s) ? 1u : 0u;
        d += (uint64_t)s * (uint64_t)s * (uint64_t)((3u * rx) ^ ry);
        if (ry == 0) {
            if (rx == 1) { x = (1u << HILBERT_BITS) - 1 - x; y = (1u << HILBERT_BITS) - 1 - y; }
            const uint32_t t = x; x = y; y = t;
        }
    }
    keys[i] = (int64_t)d;
}

__global__ void k_permute_rowlen(int64_t n, const int64_t *__restrict__ indptr, const int32_t *__restrict__ perm,
                                 int32_t *__restrict__ len) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t o = perm[i];
    len[i] = (int32_t)(indptr[o + 1] - indptr[o]);
}

constexpr int PERM_MAX_ROW = 256;

__global__ void __launch_bounds__(128) k_permute_fill(int64_t n, const int64_t *__restrict__ indptr,
                                                      const int32_t *__restrict__ indices,
                                                      const double *__restrict__ values,
                                                      const int32_t *__restrict__ perm,
                                                      const int32_t *__restrict__ iperm,
                                                      const int64_t *__restrict__ indptr_out,
                                                      int32_t *__restrict__ indices_out,
                                                      double *__restrict__ values_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t o = perm[i];
    const int64_t s = indptr[o], len = indptr[o + 1] - s, d = indptr_out[i];
    if (len <= PERM_MAX_ROW) {
        int32_t c[PERM_MAX_ROW];
        double v[PERM_MAX_ROW];
        for (int k = 0; k < (int)len; ++k) {              // insertion sort by new column id
            const int32_t ck = iperm[indices[s + k]];
            const double vk = values[s + k];
            int j = k - 1;
            while (j >= 0 && c[j] > ck) { c[j + 1] = c[j]; v[j + 1] = v[j]; --j; }
            c[j + 1] = ck; v[j + 1] = vk;
        }
        for (int k = 0; k < (int)len; ++k) { indices_out[d + k] = c[k]; values_out[d + k] = v[k]; }
    } else {                                              // long row: keep the source order
        for (int64_t k = 0; k < len; ++k) { indices_out[d + k] = iperm[indices[s + k]]; values_out[d + k] = values[s + k]; }
    }
}

__global__ void k_gather_f64(int64_t n, const double *__restrict__ src, const int32_t *__restrict__ idx,
                             double *__restrict__ dst) {
    const int64_t st = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += st) dst[i] = src[idx[i]];
}

}  // namespace
}  // namespace mwx

using namespace mwx;

extern "C" int mwx_hilbert_keys(int64_t n, const double *xy, double x0, double y0, double extent, int64_t *keys,
                                void *stream) {
    if (n <= 0 || !xy || !keys || !(extent > 0)) return MWX_E_BADARG;
    const double scale = (double)((1u << HILBERT_BITS) - 1) / extent;
    k_hilbert_keys<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(n, (const double2 *)xy, x0, y0, scale, keys);
    MWX_LAUNCH_CHECK();
    return 0;
}

extern "C" int mwx_csr_permute(int64_t n, const int64_t *indptr, const int32_t *indices, const double *values,
                               const int32_t *perm, const int32_t *iperm, int64_t *indptr_out,
                               int32_t *indices_out, double *values_out, void *stream) {
    if (n <= 0 || !indptr || !indices || !values || !perm || !iperm || !indptr_out || !indices_out || !values_out)
        return MWX_E_BADARG;
    cudaStream_t st = as_stream(stream);
    Scratch<int32_t> len(st);
    MWX_CUDA(len.alloc((size_t)n));
    k_permute_rowlen<<<grid_for(n, 256), 256, 0, st>>>(n, indptr, perm, len.p);
    MWX_LAUNCH_CHECK();
    int rc = exclusive_scan_i32_to_i64(len.p, n, indptr_out, st);
    if (rc) return rc;
    k_permute_fill<<<grid_for(n, 128), 128, 0, st>>>(n, indptr, indices, values, perm, iperm, indptr_out,
                                                     indices_out, values_out);
    MWX_LAUNCH_CHECK();
    return 0;
}

extern "C" int mwx_gather_f64(int64_t n, const double *src, const int32_t *idx, double *dst, void *stream) {
    if (n < 0 || (n && (!src || !idx || !dst))) return MWX_E_BADARG;
    if (n == 0) return 0;
    int64_t g = (n + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (g > cap) g = cap;
    k_gather_f64<<<(unsigned)g, 256, 0, as_stream(stream)>>>(n, src, idx, dst);
    MWX_LAUNCH_CHECK();
    return 0;
}
