// Row-block partitioned systems: local-matrix planning for one rank, and the CG building
// blocks the multi-GPU driver (fem_forth_perturb_b200/dist.py) chains between its halo exchanges and
// all-reduces.  Nothing here has a reference counterpart - the reference is one CPU process.
//
// A rank owns a contiguous range [lo, hi) of the (partition, Hilbert)-ordered condensed DOFs.  Its local
// operator is built from the rows it assembled itself (mwx_assemble_morley_rows): Dirichlet columns are
// dropped, the rest renumbered to the global solver numbering and then to [owned | ghost] local indices.
#include "krylov.cuh"

namespace mwx {
namespace {

constexpr int DIST_MAX_ROW = 256;

__global__ void k_dist_rows_count(int64_t nloc, const int32_t *__restrict__ rows, const int64_t *__restrict__ indptr,
                                  const int32_t *__restrict__ indices, const int32_t *__restrict__ gid,
                                  int32_t *__restrict__ raw_len, int32_t *__restrict__ kept_len) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nloc) return;
    const int64_t r = rows[i];
    int c = 0;
    for (int64_t k = indptr[r]; k < indptr[r + 1]; ++k) c += gid[indices[k]] >= 0;
    raw_len[i] = (int32_t)(indptr[r + 1] - indptr[r]);
    kept_len[i] = c;
}

// cols_g: global solver ids, ascending per row; src: where the value sits in the raw assembled rows.
__global__ void __launch_bounds__(128) k_dist_rows_fill(int64_t nloc, const int32_t *__restrict__ rows,
                                                        const int64_t *__restrict__ indptr,
                                                        const int32_t *__restrict__ indices,
                                                        const int32_t *__restrict__ gid,
                                                        const int64_t *__restrict__ raw_indptr,
                                                        const int64_t *__restrict__ loc_indptr,
                                                        int32_t *__restrict__ cols_g, int64_t *__restrict__ src,
                                                        int32_t *__restrict__ err) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nloc) return;
    const int64_t r = rows[i];
    const int64_t s = indptr[r], len = indptr[r + 1] - s;
    if (len > DIST_MAX_ROW) { atomicExch(err, 1); return; }
    int32_t c[DIST_MAX_ROW];
    int32_t o[DIST_MAX_ROW];
    int n = 0;
    for (int k = 0; k < (int)len; ++k) {
        const int32_t g = gid[indices[s + k]];
        if (g < 0) continue;
        int j = n - 1;
        while (j >= 0 && c[j] > g) { c[j + 1] = c[j]; o[j + 1] = o[j]; --j; }
        c[j + 1] = g; o[j + 1] = k;
        ++n;
    }
    const int64_t d = loc_indptr[i], rb = raw_indptr[i];
    for (int k = 0; k < n; ++k) { cols_g[d + k] = c[k]; src[d + k] = rb + o[k]; }
}

__global__ void k_dist_localize(int64_t nnz, const int32_t *__restrict__ cols_g, int32_t lo, int32_t hi,
                                const int32_t *__restrict__ ghosts, int32_t nghost, int32_t nloc,
                                int32_t *__restrict__ cols_l) {
    const int64_t st = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += st) {
        const int32_t g = cols_g[k];
        if (g >= lo && g < hi) { cols_l[k] = g - lo; continue; }
        int32_t a = 0, b = nghost - 1;
        while (a < b) {
            const int32_t m = (a + b) >> 1;
            if (ghosts[m] < g) a = m + 1; else b = m;
        }
        cols_l[k] = nloc + a;
    }
}

__global__ void k_gather_f64_i64(int64_t n, const double *__restrict__ src, const int64_t *__restrict__ idx,
                                 double *__restrict__ dst) {
    const int64_t st = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += st) dst[i] = src[idx[i]];
}

inline unsigned sgrid(int64_t n) {
    int64_t g = (n + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (g > cap) g = cap;
    return (unsigned)(g < 1 ? 1 : g);
}


// ---------------------------------------------------------------- peer-memory collectives (one process per GPU)
// Everything the partitioned iteration exchanges goes through kernels that store straight into the peers' memory over
// NVLink: halo segments, the CG scalars (all-reduce) and the coarse residual (all-gather).  Synchronisation is a
// barrier on monotonically increasing epoch flags that live in a small symmetric control block per rank; the epoch
// counter itself is DEVICE-resident, so a captured CUDA graph of one iteration can be replayed any number of times and
// mixed with eager launches (the barrier of torch's symmetric memory keeps its state on the host and cannot).
constexpr int PEER_MAX_WORLD = 32;
constexpr int PEER_MAX_SLOTS = 8;

struct PeerCtl {
    uint32_t flags[PEER_MAX_WORLD];   // flags[r]: last barrier epoch rank r has entered (written by rank r)
    uint32_t epoch;                   // local: barriers entered so far
    uint32_t nred;                    // local: all-reduces finished so far (selects the buffer half)
    uint32_t done;                    // local: "last block" ticket of the fused push
    uint32_t pad[PEER_MAX_WORLD - 3];
    double red[2][PEER_MAX_WORLD][PEER_MAX_SLOTS];   // red[half][rank][slot], written by rank `rank`
};

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const double *p) {
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

// Called by ALL threads of one block (blockDim.x >= world).  wait = 0: signal only (all parts in one process on one
// stream: stream order already is the barrier and a spin would never be released).
__device__ void peer_barrier(PeerCtl *me, PeerCtl *const *peers, int rank, int world, int wait) {
    __shared__ uint32_t e_sh;
    __syncthreads();
    if (threadIdx.x == 0) { e_sh = me->epoch + 1; me->epoch = e_sh; }
    __syncthreads();
    const uint32_t e = e_sh;
    if ((int)threadIdx.x < world) {
        __threadfence_system();
        st_release_sys(&peers[threadIdx.x]->flags[rank], e);
        if (wait)
            while ((int32_t)(ld_acquire_sys(&me->flags[threadIdx.x]) - e) < 0) { }
    }
    __syncthreads();
}

__global__ void k_peer_barrier(PeerCtl *me, PeerCtl *const *peers, int rank, int world, int wait) {
    peer_barrier(me, peers, rank, world, wait);
}

// dst[s][i - seg[s]] = v[idx ? idx[i] : i - seg[s]]; the block that finishes last runs the barrier, so the exchange
// is ONE launch: when it retires, every rank's pushes have landed.
__global__ void __launch_bounds__(256) k_peer_push(int64_t total, int32_t nseg, const int64_t *__restrict__ seg,
                                                   const int32_t *__restrict__ idx, const double *__restrict__ v,
                                                   double *const *__restrict__ dst, PeerCtl *me, PeerCtl *const *peers,
                                                   int rank, int world, int wait) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int32_t s = 0;
        while (s + 1 < nseg && i >= seg[s + 1]) ++s;
        const int64_t k = i - seg[s];
        dst[s][k] = v[idx ? (int64_t)idx[i] : k];
    }
    __shared__ int last;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(&me->done, 1u);
        last = (t == gridDim.x - 1);
        if (last) me->done = 0;
    }
    __syncthreads();
    if (last) peer_barrier(me, peers, rank, world, wait);
}

// vals[0..count) <- sum over ranks, added in rank order on every rank (bitwise identical results everywhere).
// phase 0: write + barrier + sum; 1: write only; 2: sum only (the in-process driver runs 1 for all parts, then 2).
__global__ void k_peer_allreduce(double *vals, int count, PeerCtl *me, PeerCtl *const *peers, int rank, int world,
                                 int phase) {
    const int half = (int)(me->nred & 1u);
    if (phase != 2) {
        for (int k = threadIdx.x; k < world * count; k += blockDim.x)
            peers[k / count]->red[half][rank][k % count] = vals[k % count];
    }
    if (phase == 0) peer_barrier(me, peers, rank, world, 1);
    if (phase != 1) {
        __syncthreads();
        if ((int)threadIdx.x < count) {
            double s = 0.0;
            for (int r = 0; r < world; ++r) s += ld_relaxed_sys_f64(&me->red[half][r][threadIdx.x]);
            vals[threadIdx.x] = s;
        }
        __syncthreads();
        if (threadIdx.x == 0) me->nred = me->nred + 1;
    }
}

}  // namespace
}  // namespace mwx

using namespace mwx;

extern "C" int mwx_dist_rows_count(int64_t nloc, const int32_t *rows, const int64_t *indptr,
                                   const int32_t *indices, const int32_t *gid, int64_t *raw_indptr,
                                   int64_t *loc_indptr, void *stream) {
    if (nloc <= 0 || !rows || !indptr || !indices || !gid || !raw_indptr || !loc_indptr) return MWX_E_BADARG;
    cudaStream_t st = as_stream(stream);
    Scratch<int32_t> raw(st), kept(st);
    MWX_CUDA(raw.alloc((size_t)nloc));
    MWX_CUDA(kept.alloc((size_t)nloc));
    k_dist_rows_count<<<grid_for(nloc, 256), 256, 0, st>>>(nloc, rows, indptr, indices, gid, raw.p, kept.p);
    MWX_LAUNCH_CHECK();
    int rc = exclusive_scan_i32_to_i64(raw.p, nloc, raw_indptr, st);
    if (rc) return rc;
    return exclusive_scan_i32_to_i64(kept.p, nloc, loc_indptr, st);
}

extern "C" int mwx_dist_rows_fill(int64_t nloc, const int32_t *rows, const int64_t *indptr,
                                  const int32_t *indices, const int32_t *gid, const int64_t *raw_indptr,
                                  const int64_t *loc_indptr, int32_t *cols_g, int64_t *src, void *stream) {
    if (nloc <= 0 || !rows || !indptr || !indices || !gid || !raw_indptr || !loc_indptr || !cols_g || !src)
        return MWX_E_BADARG;
    cudaStream_t st = as_stream(stream);
    Scratch<int32_t> err(st);
    MWX_CUDA(err.alloc(1));
    MWX_CUDA(cudaMemsetAsync(err.p, 0, sizeof(int32_t), st));
    k_dist_rows_fill<<<grid_for(nloc, 128), 128, 0, st>>>(nloc, rows, indptr, indices, gid, raw_indptr, loc_indptr,
                                                         cols_g, src, err.p);
    MWX_LAUNCH_CHECK();
    int32_t h = 0;
    MWX_CUDA(cudaMemcpyAsync(&h, err.p, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    MWX_CUDA(cudaStreamSynchronize(st));
    return h ? MWX_E_CAPACITY : 0;
}

extern "C" int mwx_dist_localize(int64_t nnz, const int32_t *cols_g, int32_t lo, int32_t hi,
                                 const int32_t *ghosts, int32_t nghost, int32_t nloc, int32_t *cols_l,
                                 void *stream) {
    if (nnz < 0 || (nnz && (!cols_g || !cols_l)) || (nghost && !ghosts)) return MWX_E_BADARG;
    if (nnz == 0) return 0;
    k_dist_localize<<<sgrid(nnz), 256, 0, as_stream(stream)>>>(nnz, cols_g, lo, hi, ghosts, nghost, nloc, cols_l);
    MWX_LAUNCH_CHECK();
    return 0;
}

// Halo push: one launch packs this rank's boundary values for ALL neighbours and stores each segment straight into
// that neighbour's ghost range (dst[s] is a peer pointer mapped over NVLink).  seg[s]..seg[s+1] delimits segment s of idx.
__global__ void k_halo_push(int64_t total, int32_t nseg, const int64_t *__restrict__ seg, const int32_t *__restrict__ idx,
                            const double *__restrict__ v, double *const *__restrict__ dst) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int32_t s = 0;
        while (s + 1 < nseg && i >= seg[s + 1]) ++s;
        dst[s][i - seg[s]] = v[idx[i]];
    }
}

extern "C" int mwx_halo_push(int64_t total, int32_t nseg, const int64_t *seg, const int32_t *idx, const double *v,
                             double *const *dst, void *stream) {
    if (total < 0 || nseg < 0 || (total && (!seg || !idx || !v || !dst || nseg == 0))) return MWX_E_BADARG;
    if (total == 0) return 0;
    k_halo_push<<<sgrid(total), 256, 0, as_stream(stream)>>>(total, nseg, seg, idx, v, dst);
    MWX_LAUNCH_CHECK();
    return 0;
}

extern "C" int mwx_gather_f64_i64(int64_t n, const double *src, const int64_t *idx, double *dst, void *stream) {
    if (n < 0 || (n && (!src || !idx || !dst))) return MWX_E_BADARG;
    if (n == 0) return 0;
    k_gather_f64_i64<<<sgrid(n), 256, 0, as_stream(stream)>>>(n, src, idx, dst);
    MWX_LAUNCH_CHECK();
    return 0;
}

extern "C" int mwx_peer_ctl_bytes(void) { return (int)sizeof(PeerCtl); }

static bool peer_args_ok(const void *ctl, const void *peers, int32_t rank, int32_t world) {
    return ctl && peers && world >= 1 && world <= PEER_MAX_WORLD && rank >= 0 && rank < world;
}

extern "C" int mwx_peer_barrier(void *ctl, const void *peers, int32_t rank, int32_t world, int32_t wait, void *stream) {
    if (!peer_args_ok(ctl, peers, rank, world)) return MWX_E_BADARG;
    k_peer_barrier<<<1, PEER_MAX_WORLD, 0, as_stream(stream)>>>((PeerCtl *)ctl, (PeerCtl *const *)peers, rank, world, wait);
    MWX_LAUNCH_CHECK();
    return 0;
}

extern "C" int mwx_peer_push(int64_t total, int32_t nseg, const int64_t *seg, const int32_t *idx, const double *v,
                             double *const *dst, void *ctl, const void *peers, int32_t rank, int32_t world,
                             int32_t wait, void *stream) {
    if (!peer_args_ok(ctl, peers, rank, world) || total < 0 || nseg < 0 || (total && (!seg || !v || !dst || nseg == 0)))
        return MWX_E_BADARG;
    k_peer_push<<<sgrid(total), 256, 0, as_stream(stream)>>>(total, nseg, seg, idx, v, dst, (PeerCtl *)ctl,
                                                             (PeerCtl *const *)peers, rank, world, wait);
    MWX_LAUNCH_CHECK();
    return 0;
}

extern "C" int mwx_peer_allreduce(double *vals, int32_t count, void *ctl, const void *peers, int32_t rank,
                                  int32_t world, int32_t phase, void *stream) {
    if (!peer_args_ok(ctl, peers, rank, world) || !vals || count < 1 || count > PEER_MAX_SLOTS || phase < 0 || phase > 2)
        return MWX_E_BADARG;
    k_peer_allreduce<<<1, 256, 0, as_stream(stream)>>>(vals, count, (PeerCtl *)ctl, (PeerCtl *const *)peers, rank, world,
                                                       phase);
    MWX_LAUNCH_CHECK();
    return 0;
}
