// Glue between hot paths A and B: eliminate the Dirichlet DOFs on device.
//
// Replaces condense(K2, f2, x, D=...) ref main/example1/utils.py:384-460:
//     Aout = A[I].T[I].T          bout = b[I] - A[I].T[D].T @ x[D]
// scipy does two CSR fancy-index passes and two transposes; here the kept rows/columns are compacted
// in one count pass and one fill pass, and the boundary contribution to the right-hand side is
// folded into the fill pass.  (The reference's D may list corner DOFs twice, which would count their
// x[D] twice; x[D] = 0 wherever the reference does that, so the plain set semantics are used.)
#include "common.cuh"

namespace mwx {
namespace {

__global__ void k_condense_rowlen(int64_t n, const int64_t *__restrict__ indptr,
                                  const int32_t *__restrict__ indices, const int32_t *__restrict__ keep,
                                  const int32_t *__restrict__ newid, int32_t *__restrict__ len) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n || !keep[r]) return;
    int c = 0;
    for (int64_t k = indptr[r]; k < indptr[r + 1]; ++k) c += keep[indices[k]] != 0;
    len[newid[r]] = c;
}

__global__ void k_condense_fill(int64_t n, const int64_t *__restrict__ indptr,
                                const int32_t *__restrict__ indices, const double *__restrict__ values,
                                const int32_t *__restrict__ keep, const int32_t *__restrict__ newid,
                                const int64_t *__restrict__ indptr_out, int32_t *__restrict__ indices_out,
                                double *__restrict__ values_out, const double *__restrict__ b,
                                const double *__restrict__ x, double *__restrict__ b_out) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n || !keep[r]) return;
    const int32_t ro = newid[r];
    int64_t o = indptr_out[ro];
    double sub = 0.0;
    for (int64_t k = indptr[r]; k < indptr[r + 1]; ++k) {
        const int32_t c = indices[k];
        if (keep[c]) {
            indices_out[o] = newid[c];
            values_out[o] = values[k];
            ++o;
        } else if (x) {
            sub += values[k] * x[c];
        }
    }
    if (b_out) b_out[ro] = (b ? b[r] : 0.0) - sub;
}

}  // namespace
}  // namespace mwx

using namespace mwx;

extern "C" int mwx_condense_count(int64_t n, const int64_t *indptr, const int32_t *indices,
                                  const int32_t *keep, int32_t *newid, int64_t *indptr_out,
                                  int64_t *n_out_host, int64_t *nnz_out_host, void *stream) {
    if (n <= 0 || !indptr || !indices || !keep || !newid || !indptr_out || !n_out_host || !nnz_out_host)
        return MWX_E_BADARG;
    if (n >= ((int64_t)1 << 31)) return MWX_E_CAPACITY;
    cudaStream_t st = as_stream(stream);
    int rc = exclusive_scan_i32_to_i32(keep, n, newid, st);     // newid has n+1 entries
    if (rc) return rc;
    int32_t nout32 = 0;
    MWX_CUDA(cudaMemcpyAsync(&nout32, newid + n, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    MWX_CUDA(cudaStreamSynchronize(st));
    const int64_t nout = nout32;
    *n_out_host = nout;
    Scratch<int32_t> len(st);
    MWX_CUDA(len.alloc((size_t)nout));
    if (nout) {
        k_condense_rowlen<<<grid_for(n, 256), 256, 0, st>>>(n, indptr, indices, keep, newid, len.p);
        MWX_LAUNCH_CHECK();
    }
    rc = exclusive_scan_i32_to_i64(len.p, nout, indptr_out, st);
    if (rc) return rc;
    MWX_CUDA(cudaMemcpyAsync(nnz_out_host, indptr_out + nout, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    MWX_CUDA(cudaStreamSynchronize(st));
    return 0;
}

extern "C" int mwx_condense_fill(int64_t n, const int64_t *indptr, const int32_t *indices,
                                 const double *values, const int32_t *keep, const int32_t *newid,
                                 const int64_t *indptr_out, int32_t *indices_out, double *values_out,
                                 const double *b, const double *x, double *b_out, void *stream) {
    if (n <= 0 || !indptr || !indices || !values || !keep || !newid || !indptr_out || !indices_out || !values_out)
        return MWX_E_BADARG;
    k_condense_fill<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(
        n, indptr, indices, values, keep, newid, indptr_out, indices_out, values_out, b, x, b_out);
    MWX_LAUNCH_CHECK();
    return 0;
}
