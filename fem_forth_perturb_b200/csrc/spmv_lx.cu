// Tile-local column plan for the CSR SpMV (krylov.cu, FMT = LX).
//
// k_spmv is bound by the L1 wavefronts of its x gathers, not by DRAM: a Morley row has 11.5 nonzeros whose columns are
// scattered, so nearly every gather is a wavefront of its own (771 M per product at 67 M rows ~ 2.6 ms at one wavefront
// per clock and SM) while the 10.6 GB it streams would take 1.6 ms.  But the rows of one tile (a compact patch of the
// Hilbert-ordered mesh) share their columns: ~190 rows x 11.5 nonzeros touch only ~330 distinct x entries.  The plan
// lists those per tile (`ucols`, sorted) and replaces every 4-byte column by a 2-byte index into that list (`lidx`);
// the kernel gathers each distinct entry ONCE into shared memory (sorted addresses: few wavefronts) and the products
// read x from there.  Per nonzero: 8 B value + 2 B index + ~0.6 B of the unique list instead of 12 B.
//
// Built once per operator pattern (setup): one CTA per tile sorts the tile's columns in shared memory (bitonic),
// compacts the distinct ones and binary-searches every nonzero in them.  Tiles with more than LX_MAX_U distinct
// columns, or too many nonzeros to sort, make the plan fail (*out = nullptr): the caller keeps the plain kernel.
#include <climits>

#include "krylov.cuh"

namespace mwx {
namespace {

constexpr int LX_SORT_MAX = 4096;       // nonzeros of a tile that can be sorted (a TMA stage holds 2568)
constexpr int LX_THREADS = 256;

__device__ __forceinline__ int next_pow2(int v) {
    int p = LX_THREADS;
    while (p < v) p <<= 1;
    return p;
}

// pass 0: ucount[tile] = number of distinct columns rounded up to a multiple of 4 (0: tile not localised)
// pass 1: fill ucols (padding repeats the last column) and lidx
__global__ void __launch_bounds__(LX_THREADS) k_lx_build(int64_t n, const int64_t *__restrict__ indptr,
                                                         const int32_t *__restrict__ indices, int rows, int64_t ntiles,
                                                         int32_t *__restrict__ ucount, const int32_t *__restrict__ uptr,
                                                         int32_t *__restrict__ ucols, uint16_t *__restrict__ lidx,
                                                         int *__restrict__ fail, int pass) {
    __shared__ int32_t keys[LX_SORT_MAX];
    __shared__ int32_t uniq[LX_MAX_U];
    __shared__ int32_t wsum[LX_THREADS / 32];
    __shared__ int32_t total_sh;
    const int tid = threadIdx.x;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t r0 = tile * rows, r1 = r0 + rows < n ? r0 + rows : n;
        const int64_t s0 = indptr[r0], s1 = indptr[r1];
        const int64_t cnt64 = s1 - s0;
        __syncthreads();                                     // previous tile's shared data fully consumed
        if (cnt64 > LX_SORT_MAX || cnt64 <= 0) {
            if (cnt64 > LX_SORT_MAX && tid == 0) atomicExch(fail, 1);
            if (pass == 0 && tid == 0) ucount[tile] = 0;
            continue;
        }
        const int cnt = (int)cnt64, P = next_pow2(cnt);
        for (int i = tid; i < P; i += LX_THREADS) keys[i] = i < cnt ? indices[s0 + i] : INT_MAX;
        __syncthreads();
        for (int k = 2; k <= P; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = tid; i < P; i += LX_THREADS) {
                    const int l = i ^ j;
                    if (l > i) {
                        const int32_t a = keys[i], b = keys[l];
                        const bool up = (i & k) == 0;
                        if ((a > b) == up) { keys[i] = b; keys[l] = a; }
                    }
                }
                __syncthreads();
            }
        // distinct keys: thread t owns the contiguous chunk [t*C, (t+1)*C)
        const int C = P / LX_THREADS;
        int mine = 0;
        for (int i = tid * C; i < (tid + 1) * C; ++i) mine += (i < cnt && (i == 0 || keys[i] != keys[i - 1])) ? 1 : 0;
        int inc = mine;
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, inc, o);
            if ((tid & 31) >= o) inc += v;
        }
        if ((tid & 31) == 31) wsum[tid >> 5] = inc;
        __syncthreads();
        if (tid == 0) {
            int run = 0;
            for (int w = 0; w < LX_THREADS / 32; ++w) { const int v = wsum[w]; wsum[w] = run; run += v; }
            total_sh = run;
        }
        __syncthreads();
        const int U = total_sh;
        if (U > LX_MAX_U) {
            if (tid == 0) { atomicExch(fail, 1); if (pass == 0) ucount[tile] = 0; }
            continue;
        }
        if (pass == 0) {
            if (tid == 0) ucount[tile] = (U + 3) & ~3;
            continue;
        }
        int pos = wsum[tid >> 5] + inc - mine;
        for (int i = tid * C; i < (tid + 1) * C; ++i)
            if (i < cnt && (i == 0 || keys[i] != keys[i - 1])) uniq[pos++] = keys[i];
        __syncthreads();
        const int32_t u0 = uptr[tile], upad = uptr[tile + 1] - u0;
        for (int j = tid; j < upad; j += LX_THREADS) ucols[u0 + j] = uniq[j < U ? j : U - 1];
        for (int i = tid; i < cnt; i += LX_THREADS) {
            const int32_t c = indices[s0 + i];
            int a = 0, b = U - 1;
            while (a < b) {
                const int m = (a + b) >> 1;
                if (uniq[m] < c) a = m + 1; else b = m;
            }
            lidx[s0 + i] = (uint16_t)a;
        }
    }
}

}  // namespace

void spmv_lx_destroy(SpmvLX *p) {
    if (!p) return;
    cudaFree(p->uptr); cudaFree(p->ucols); cudaFree(p->lidx);
    delete p;
}

int spmv_lx_create(int64_t n, const int64_t *indptr, const int32_t *indices, int plan, SpmvLX **out, cudaStream_t st) {
    *out = nullptr;
    const int rows = plan & 4095;
    if (n <= 0 || rows <= 0) return 0;
    const int64_t ntiles = (n + rows - 1) / rows;
    if (ntiles > INT_MAX / 2) return 0;
    int64_t nnz = 0;
    MWX_CUDA(cudaMemcpyAsync(&nnz, indptr + n, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    MWX_CUDA(cudaStreamSynchronize(st));
    if (nnz <= 0) return 0;
    SpmvLX *p = new SpmvLX();
    p->rows_per_tile = rows; p->plan = plan; p->ntiles = ntiles; p->n = n; p->nnz = nnz;
    Scratch<int32_t> ucount(st);
    Scratch<int> fail(st);
    auto bail = [&](int code) { spmv_lx_destroy(p); return code; };
    if (ucount.alloc((size_t)ntiles) != cudaSuccess || fail.alloc(1) != cudaSuccess) return bail(MWX_E_NOMEM);
    if (cudaMalloc((void **)&p->uptr, sizeof(int32_t) * ((size_t)ntiles + 1)) != cudaSuccess ||
        cudaMalloc((void **)&p->lidx, sizeof(uint16_t) * ((size_t)nnz + 16)) != cudaSuccess)
        return bail(MWX_E_NOMEM);
    if (cudaMemsetAsync(fail.p, 0, sizeof(int), st) != cudaSuccess ||
        cudaMemsetAsync(p->lidx, 0, sizeof(uint16_t) * ((size_t)nnz + 16), st) != cudaSuccess)
        return bail(MWX_E_NOMEM);
    int64_t g = ntiles;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (g > cap) g = cap;
    k_lx_build<<<(unsigned)g, LX_THREADS, 0, st>>>(n, indptr, indices, rows, ntiles, ucount.p, nullptr, nullptr, nullptr, fail.p, 0);
    if (cudaGetLastError() != cudaSuccess) return bail(MWX_E_BADARG);
    int h = 0;
    if (cudaMemcpyAsync(&h, fail.p, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess)
        return bail(MWX_E_BADARG);
    if (h) return bail(0);                                   // some tile is too wide: no plan, the plain kernel stays
    int rc = exclusive_scan_i32_to_i32(ucount.p, ntiles, p->uptr, st);
    if (rc) return bail(rc);
    int32_t total = 0;
    if (cudaMemcpyAsync(&total, p->uptr + ntiles, sizeof(int32_t), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess)
        return bail(MWX_E_BADARG);
    p->nucols = total;
    if (cudaMalloc((void **)&p->ucols, sizeof(int32_t) * ((size_t)total + 4)) != cudaSuccess) return bail(MWX_E_NOMEM);
    k_lx_build<<<(unsigned)g, LX_THREADS, 0, st>>>(n, indptr, indices, rows, ntiles, nullptr, p->uptr, p->ucols, p->lidx, fail.p, 1);
    if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) return bail(MWX_E_BADARG);
    *out = p;
    return 0;
}

}  // namespace mwx
