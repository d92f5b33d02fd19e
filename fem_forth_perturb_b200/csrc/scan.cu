// Exclusive prefix sums (setup-time utility for the mesh / pattern / condense builders).
// Three-phase tiled scan: tile sums -> (recursive) scan of the sums -> per-tile scan + offset.
#include "common.cuh"

namespace mwx {

namespace {
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_sums(const TIn *__restrict__ in, int64_t n,
                                                               TOut *__restrict__ tile_sum) {
    __shared__ TOut sh[SCAN_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    TOut s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        const int64_t i = base + (int64_t)k * SCAN_THREADS + threadIdx.x;   // coalesced
        if (i < n) s += (TOut)in[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        TOut t = 0;
        for (int w = 0; w < SCAN_THREADS / 32; ++w) t += sh[w];
        tile_sum[blockIdx.x] = t;
    }
}

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_apply(const TIn *__restrict__ in, int64_t n,
                                                                const TOut *__restrict__ tile_off,
                                                                TOut *__restrict__ out) {
    __shared__ TOut warp_tot[SCAN_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    TOut v[SCAN_ITEMS];
    TOut tsum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        const int64_t i = base + k;
        v[k] = i < n ? (TOut)in[i] : (TOut)0;
        tsum += v[k];
    }
    // inclusive warp scan of the per-thread totals
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    TOut inc = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        TOut y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += y;
    }
    if (lane == 31) warp_tot[w] = inc;
    __syncthreads();
    TOut woff = 0;
    for (int k = 0; k < w; ++k) woff += warp_tot[k];
    TOut run = tile_off[blockIdx.x] + woff + (inc - tsum);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        const int64_t i = base + k;
        if (i < n) {
            out[i] = run;
            run += v[k];
            if (i == n - 1) out[n] = run;
        }
    }
}

template <typename T>
__global__ void scan_set_zero(T *p) { p[0] = 0; }

template <typename TIn, typename TOut>
int exclusive_scan_impl(const TIn *in, int64_t n, TOut *out, cudaStream_t st) {
    if (n <= 0) {
        scan_set_zero<TOut><<<1, 1, 0, st>>>(out);
        MWX_LAUNCH_CHECK();
        return 0;
    }
    const int64_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
    Scratch<TOut> sums(st), offs(st);
    MWX_CUDA(sums.alloc((size_t)nb));
    MWX_CUDA(offs.alloc((size_t)nb + 1));
    scan_tile_sums<TIn, TOut><<<(unsigned)nb, SCAN_THREADS, 0, st>>>(in, n, sums.p);
    MWX_LAUNCH_CHECK();
    if (nb == 1) {
        scan_set_zero<TOut><<<1, 1, 0, st>>>(offs.p);
        MWX_LAUNCH_CHECK();
    } else {
        int rc = exclusive_scan_impl<TOut, TOut>(sums.p, nb, offs.p, st);
        if (rc) return rc;
    }
    scan_tile_apply<TIn, TOut><<<(unsigned)nb, SCAN_THREADS, 0, st>>>(in, n, offs.p, out);
    MWX_LAUNCH_CHECK();
    return 0;
}
}  // namespace

int exclusive_scan_i32_to_i64(const int32_t *in, int64_t n, int64_t *out, cudaStream_t st) {
    return exclusive_scan_impl<int32_t, int64_t>(in, n, out, st);
}
int exclusive_scan_i32_to_i32(const int32_t *in, int64_t n, int32_t *out, cudaStream_t st) {
    return exclusive_scan_impl<int32_t, int32_t>(in, n, out, st);
}

void keep_pool_memory() {
    static bool done = false;
    if (done) return;
    done = true;
    int dev = 0;
    cudaMemPool_t pool;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        uint64_t keep = ~(uint64_t)0;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
}

int sm_count() {
    static int cached = 0;
    if (!cached) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            cached = n;
        else
            cached = 148;
    }
    return cached;
}

}  // namespace mwx
