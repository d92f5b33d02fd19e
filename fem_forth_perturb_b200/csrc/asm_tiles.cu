// Hot path A, second generation: every element evaluated ONCE per tile, values written straight into the CSR.
//
// Replaces  K2 = epsilon**2 * asm(a_load, basis['u']) + asm(b_load, basis['u'])   (ref main/example1/run.py:210;
// forms ref main/example1/run.py:110-123) like assembly.cu's row-gather kernel, which stays as the general fallback.
// That kernel owns one matrix ROW per thread, re-evaluates an element's local row for each of its six rows and
// gathers a scattered 64-byte geometry record per (row, element) pair; in skfem's numbering a tile's elements are
// spread over the whole mesh, so the records were re-read from DRAM ~4x (ncu: 1.78x the algorithmic traffic, L1 data
// pipe 61 % = the limiter, 0.25 of the HBM roofline).
//
// Here the work is organised by a PLAN built once per (mesh, destination layout):
//   * destination rows are taken in a locality order chosen by the caller (Hilbert curve over the DOF locations)
//     and cut into tiles of <= 256 rows: a tile is a compact patch of the mesh, whatever skfem's numbering is;
//   * per tile the plan lists the patch's unique vertices and unique elements (three 1-byte local vertex ids each),
//     and for every destination entry (row i, column j) of the tile the <= 2 elements that contribute to it, as
//     16-bit indices into the tile's table of local matrices.  (Two distinct DOFs share at most two triangles; the
//     diagonal of a vertex row has valence-many contributions and goes through a per-row list instead.)
// The kernel (persistent CTAs, 2 per SM, one tile at a time per CTA):
//   phase 1  one thread per unique element: 3 vertex coordinates from shared memory -> the 21 entries of the
//            symmetric 6x6 matrix ca*A_K + cb*B_K in closed form -> shared memory (SoA, conflict-free);
//   phase D  one thread per row: its diagonal = sum over the row's elements in ascending element id;
//   phase 2  one thread per destination ENTRY, entries in destination order row by row: value = first + second
//            contribution (fixed order) -> ONE coalesced 8-byte store per entry, straight into the CSR values
//            (a row's entries are consecutive lanes; the rows of a tile go to their own places in skfem's order).
// No atomics, no read-modify-write, a fixed summation order => bit-reproducible; each unique element of a tile is
// evaluated once (~1.5x redundancy between tiles instead of 6x), and everything a tile needs arrives by TMA bulk
// copies (cp.async.bulk + mbarrier) one tile ahead, its vertex coordinates by cp.async gathers issued while the
// previous tile is being stored.
//
// Algorithmic bytes per element as bench.py counts them: still 308 B (the round-1 model of the row-gather kernel), so
// that roofline fractions stay comparable; this kernel's own traffic per element is ~184 (values) + 92 (entry codes)
// + 12 (diagonal lists) + 16 (row offsets) + ~20 (vertex / element lists, coordinates) = ~325 B.
#include <new>
#include <vector>

#include "common.cuh"
#include "tma.cuh"

namespace mwx {
namespace {

constexpr int AT_THREADS = 256;   // threads per CTA = maximum rows per tile
constexpr int AT_ECAP = 288;      // unique elements per tile (a 256-row Hilbert window of the uniform mesh has ~200)
constexpr int AT_VCAP = 256;      // unique vertices per tile: a local vertex id is one byte
constexpr int AT_NCAP = 3584;     // destination entries per tile (uniform mesh: 64 x 19 + 192 x 9 = 2944)
constexpr int AT_DCAP = 1024;     // (row, element) pairs per tile (uniform mesh: 768)
constexpr int AT_MAXP = 16;       // elements per row (vertex valence)
constexpr int AT_MAXLEN = 64;     // destination entries per row
constexpr int AT_MAXRUNS = 16;    // runs of equal row length inside a tile
constexpr int AT_NSYM = 21;       // upper triangle of the symmetric 6x6 local matrix
constexpr int AT_LSTRIDE = AT_ECAP + 1;   // ODD stride (in doubles) of the local-matrix table loc[k * AT_LSTRIDE + e]: the lanes
                                          // of a row read different k of the same one or two elements - with an even stride
                                          // they all hit one bank (measured: 12 wavefronts per LDS.64 instead of 2)
// The table has two more regions so that phase 2 is branch-free (value = loc[first] + loc[second], always):
constexpr int AT_DIAG_BASE = AT_NSYM * AT_LSTRIDE;        // loc[AT_DIAG_BASE + row]: the row's diagonal (phase D)
constexpr int AT_ZERO = AT_DIAG_BASE + AT_THREADS;        // loc[AT_ZERO] = 0.0: the "no second contribution" slot
constexpr int AT_LOC_DOUBLES = AT_ZERO + 1;
static_assert(AT_LOC_DOUBLES < 65536, "entry codes are 16-bit indices into the table");

// Per-tile header at the start of the tile's "big" blob.  288 bytes (a multiple of 16: TMA granularity).  Rows of a
// tile are sorted by (element count, length) descending; a RUN is a maximal sequence of rows of equal length, so an
// entry index maps to (row, position) by one multiply-high with the run's magic number ceil(2^32 / length).
struct AsmTileHdr {
    int32_t nrows, nentries, npairs, nruns;
    int32_t npq, pad0, pad1, pad2;                // npq = largest number of elements of a row
    uint32_t run_magic[AT_MAXRUNS];               // ceil(2^32 / run_len) (0 for length 1)
    uint16_t run_es[AT_MAXRUNS + 2];              // first entry of run r; run_es[r >= nruns] = nentries
    uint16_t run_rs[AT_MAXRUNS + 2];              // first row of run r
    uint16_t jds_off[AT_MAXP + 2];                // diagonal lists, jagged-diagonal layout: offset of the q-th elements
    uint16_t jds_cnt[AT_MAXP];                    // rows with more than q elements
    uint8_t run_len[AT_MAXRUNS];                  // entries per row in run r
    uint8_t pad3[36];
};
static_assert(sizeof(AsmTileHdr) == 288, "header layout");

// big blob  = [AsmTileHdr][row_out: pad2(nrows) x int64][codes: pad4(nentries) x uint32][dcodes: pad8(npairs) x uint16]
// geo blob  = [nvert, nelem, 0, 0 : int32][vlist: pad4(nvert) x uint32][etri: pad4(nelem) x uchar4]
__host__ __device__ inline int pad_to(int n, int m) { return (n + m - 1) / m * m; }
__host__ __device__ inline int big_bytes(int nrows, int nentries, int npairs) {
    return (int)sizeof(AsmTileHdr) + pad_to(nrows, 2) * 8 + pad_to(nentries, 4) * 4 + pad_to(npairs, 8) * 2;
}
__host__ __device__ inline int geo_bytes(int nvert, int nelem) { return 16 + pad_to(nvert, 4) * 4 + pad_to(nelem, 4) * 4; }
constexpr int AT_BIG_MAX = 288 + AT_THREADS * 8 + AT_NCAP * 4 + AT_DCAP * 2;      // 18720
constexpr int AT_GEO_MAX = 16 + AT_VCAP * 4 + AT_ECAP * 4;                        // 2192

// Tile table entry (what the producer thread reads one tile ahead): blob offsets in bytes.
struct AsmTileRef { int64_t big_off, geo_off; int32_t big_len, geo_len; };
static_assert(sizeof(AsmTileRef) == 24, "tile table layout");

__device__ __forceinline__ int sym_index(int a, int b) {          // a, b in 0..5
    const int lo = a < b ? a : b, hi = a < b ? b : a;
    return lo * 6 - (lo * (lo - 1)) / 2 + (hi - lo);
}

// ------------------------------------------------------------------------------------------------ the hot kernel
constexpr int AT_SMEM_LOC = (AT_LOC_DOUBLES * 8 + 15) / 16 * 16;       // 50608
constexpr int AT_SMEM_COORD = 2 * AT_VCAP * 16;                    // 8192
constexpr int AT_SMEM_DIAG = 0;                                    // (the diagonals live in the loc table)
constexpr int AT_SMEM_BIG = 2 * AT_BIG_MAX;                        // 37440
constexpr int AT_SMEM_GEO = 3 * AT_GEO_MAX;                        // 6576
constexpr int AT_SMEM_TAIL = 64;                                   // barriers
constexpr int AT_SMEM_BYTES = AT_SMEM_LOC + AT_SMEM_COORD + AT_SMEM_DIAG + AT_SMEM_BIG + AT_SMEM_GEO + AT_SMEM_TAIL;
static_assert(AT_BIG_MAX % 16 == 0 && AT_GEO_MAX % 16 == 0, "stage alignment");

__device__ __forceinline__ void cp_async_16(void *dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}

// The 21 entries of ca*A_K + cb*B_K (upper triangle, local order v0 v1 v2 e0 e1 e2) of the triangle (x0, x1, x2)
// whose vertices are in ascending GLOBAL order (that fixes the edge-normal orientation).  Same closed form as
// assembly.cu::morley_local_row, evaluated once for the whole matrix.
__device__ __forceinline__ void morley_local_sym(double2 x0, double2 x1, double2 x2, double ca, double cb, double *loc,
                                                 int stride) {
    const double d0x = x0.x - x1.x, d0y = x0.y - x1.y;
    const double d1x = x1.x - x2.x, d1y = x1.y - x2.y;
    const double d2x = x0.x - x2.x, d2y = x0.y - x2.y;
    const double det = d0x * d2y - d0y * d2x;
    const double rdet = 1.0 / det;
    const double area = 0.5 * fabs(det);
    const double q0 = d0x * d0x + d0y * d0y, q1 = d1x * d1x + d1y * d1y, q2 = d2x * d2x + d2y * d2y;
    const double r0 = rsqrt(q0), r1 = rsqrt(q1), r2 = rsqrt(q2);
    const double g0 = q0 * r0 * rdet, g1 = q1 * r1 * rdet, g2 = -(q2 * r2) * rdet;
    const double m01 = (d0x * d1x + d0y * d1y) * (r0 * r1);
    const double m02 = (d0x * d2x + d0y * d2y) * (r0 * r2);
    const double m12 = (d1x * d2x + d1y * d2y) * (r1 * r2);
    const double a4 = 4.0 * area;
    const double E00 = a4 * g0 * g0, E11 = a4 * g1 * g1, E22 = a4 * g2 * g2;
    const double E01 = a4 * g0 * g1 * m01 * m01, E02 = a4 * g0 * g2 * m02 * m02, E12 = a4 * g1 * g2 * m12 * m12;
    // Dv[i][k] = gam_{p(i)} mu_{k,p(i)}, p = (1, 2, 0)
    const double D00 = g1 * m01, D01 = g1, D02 = g1 * m12;
    const double D10 = g2 * m02, D11 = g2 * m12, D12 = g2;
    const double D20 = g0, D21 = g0 * m01, D22 = g0 * m02;
    // W[i][m] = sum_k Dv[i][k] E[k][m]   (= the vertex-edge block of A)
    const double W00 = D00 * E00 + D01 * E01 + D02 * E02, W01 = D00 * E01 + D01 * E11 + D02 * E12,
                 W02 = D00 * E02 + D01 * E12 + D02 * E22;
    const double W10 = D10 * E00 + D11 * E01 + D12 * E02, W11 = D10 * E01 + D11 * E11 + D12 * E12,
                 W12 = D10 * E02 + D11 * E12 + D12 * E22;
    const double W20 = D20 * E00 + D21 * E01 + D22 * E02, W21 = D20 * E01 + D21 * E11 + D22 * E12,
                 W22 = D20 * E02 + D21 * E12 + D22 * E22;
    const double t3 = 1.0 / 3.0, cba = cb * area;
    // vertex-vertex: ca * W[i] . Dv[j] + cb |K| (g_i . g_j - 1/3 Dv[i] . Dv[j])
    loc[0 * stride] = ca * (W00 * D00 + W01 * D01 + W02 * D02) + cba * (g1 * g1 - t3 * (D00 * D00 + D01 * D01 + D02 * D02));
    loc[1 * stride] = ca * (W00 * D10 + W01 * D11 + W02 * D12) + cba * (g1 * g2 * m12 - t3 * (D00 * D10 + D01 * D11 + D02 * D12));
    loc[2 * stride] = ca * (W00 * D20 + W01 * D21 + W02 * D22) + cba * (g1 * g0 * m01 - t3 * (D00 * D20 + D01 * D21 + D02 * D22));
    loc[6 * stride] = ca * (W10 * D10 + W11 * D11 + W12 * D12) + cba * (g2 * g2 - t3 * (D10 * D10 + D11 * D11 + D12 * D12));
    loc[7 * stride] = ca * (W10 * D20 + W11 * D21 + W12 * D22) + cba * (g2 * g0 * m02 - t3 * (D10 * D20 + D11 * D21 + D12 * D22));
    loc[11 * stride] = ca * (W20 * D20 + W21 * D21 + W22 * D22) + cba * (g0 * g0 - t3 * (D20 * D20 + D21 * D21 + D22 * D22));
    // vertex-edge (B_ve = 0)
    loc[3 * stride] = ca * W00; loc[4 * stride] = ca * W01; loc[5 * stride] = ca * W02;
    loc[8 * stride] = ca * W10; loc[9 * stride] = ca * W11; loc[10 * stride] = ca * W12;
    loc[12 * stride] = ca * W20; loc[13 * stride] = ca * W21; loc[14 * stride] = ca * W22;
    // edge-edge: ca * E + cb |K| / 3 I
    const double third = cba * t3;
    loc[15 * stride] = ca * E00 + third; loc[16 * stride] = ca * E01; loc[17 * stride] = ca * E02;
    loc[18 * stride] = ca * E11 + third; loc[19 * stride] = ca * E12;
    loc[20 * stride] = ca * E22 + third;
}

template <bool ACC>
__global__ void __launch_bounds__(AT_THREADS, 2) k_assemble_morley_tiles(
    int64_t ntiles, const AsmTileRef *__restrict__ tiles, const unsigned char *__restrict__ big,
    const unsigned char *__restrict__ geo, const double2 *__restrict__ pxy, double ca, double cb,
    double *__restrict__ values) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *loc = reinterpret_cast<double *>(smem_raw);
    double2 *coords = reinterpret_cast<double2 *>(smem_raw + AT_SMEM_LOC);                 // [2][AT_VCAP]
    unsigned char *s_big = smem_raw + AT_SMEM_LOC + AT_SMEM_COORD + AT_SMEM_DIAG;          // [2][AT_BIG_MAX]
    unsigned char *s_geo = s_big + AT_SMEM_BIG;                                            // [3][AT_GEO_MAX]
    unsigned char *tail = s_geo + AT_SMEM_GEO;
    uint64_t *bar_big = reinterpret_cast<uint64_t *>(tail);            // [2]
    uint64_t *bar_geo = reinterpret_cast<uint64_t *>(tail + 16);       // [3]

    const int tid = threadIdx.x;
    const int64_t G = gridDim.x;
    const int64_t first = blockIdx.x;
    if (first >= ntiles) return;
    if (tid == 0) {
        mbar_init(&bar_big[0], 1); mbar_init(&bar_big[1], 1);
        mbar_init(&bar_geo[0], 1); mbar_init(&bar_geo[1], 1); mbar_init(&bar_geo[2], 1);
        mbar_fence_init();
        loc[AT_ZERO] = 0.0;
    }
    __syncthreads();

    auto issue_big = [&](const AsmTileRef &r, int stage) {             // thread 0
        fence_proxy_async();
        mbar_expect_tx(&bar_big[stage], (uint32_t)r.big_len);
        bulk_g2s(s_big + stage * AT_BIG_MAX, big + r.big_off, (uint32_t)r.big_len, &bar_big[stage]);
    };
    auto issue_geo = [&](const AsmTileRef &r, int stage) {             // thread 0
        fence_proxy_async();
        mbar_expect_tx(&bar_geo[stage], (uint32_t)r.geo_len);
        bulk_g2s(s_geo + stage * AT_GEO_MAX, geo + r.geo_off, (uint32_t)r.geo_len, &bar_geo[stage]);
    };
    // all threads: gather the vertex coordinates of the tile in geo stage `gs` into coordinate buffer `cs`
    auto gather_coords = [&](int gs, int cs) {
        const int32_t *ghdr = reinterpret_cast<const int32_t *>(s_geo + gs * AT_GEO_MAX);
        const int nvert = ghdr[0];
        const uint32_t *vlist = reinterpret_cast<const uint32_t *>(s_geo + gs * AT_GEO_MAX + 16);
        for (int v = tid; v < nvert; v += AT_THREADS) cp_async_16(&coords[cs * AT_VCAP + v], pxy + vlist[v]);
        cp_async_commit();
    };

    // ---- prologue: tile j = 0 fully staged, geo of j = 1 in flight, refs of big(1) / geo(2) in registers (thread 0)
    AsmTileRef ref_big_next = {0, 0, 0, 0}, ref_geo_next2 = {0, 0, 0, 0};
    if (tid == 0) {
        const AsmTileRef r0 = tiles[first];
        issue_big(r0, 0);
        issue_geo(r0, 0);
        if (first + G < ntiles) {
            ref_big_next = tiles[first + G];
            issue_geo(ref_big_next, 1);
        }
        if (first + 2 * G < ntiles) ref_geo_next2 = tiles[first + 2 * G];
    }
    mbar_wait(&bar_geo[0], 0);
    gather_coords(0, 0);
    cp_async_wait_all();
    __syncthreads();

    int64_t tile = first;
    for (int j = 0; tile < ntiles; tile += G, ++j) {
        const int bs = j & 1, gs = j % 3;
        const bool have_next = tile + G < ntiles, have_next2 = tile + 2 * G < ntiles;
        if (tid == 0) {
            if (have_next) issue_big(ref_big_next, bs ^ 1);                        // stage freed at the end of iteration j-1
            if (have_next2) {
                issue_geo(ref_geo_next2, (j + 2) % 3);                             // held tile j-1: consumed in iteration j-1
                ref_big_next = ref_geo_next2;                                      // big(j+2) is issued next iteration
            }
            if (tile + 3 * G < ntiles) ref_geo_next2 = tiles[tile + 3 * G];        // lands during this tile
        }
        // ---- phase 1: local matrices of the tile's unique elements
        mbar_wait(&bar_geo[gs], (uint32_t)((j / 3) & 1));
        {
            const int32_t *ghdr = reinterpret_cast<const int32_t *>(s_geo + gs * AT_GEO_MAX);
            const int nvert = ghdr[0], nelem = ghdr[1];
            const uchar4 *etri = reinterpret_cast<const uchar4 *>(s_geo + gs * AT_GEO_MAX + 16 + pad_to(nvert, 4) * 4);
            const double2 *cx = coords + bs * AT_VCAP;
            for (int e = tid; e < nelem; e += AT_THREADS) {
                const uchar4 tv = etri[e];
                morley_local_sym(cx[tv.x], cx[tv.y], cx[tv.z], ca, cb, loc + e, AT_LSTRIDE);
            }
        }
        mbar_wait(&bar_big[bs], (uint32_t)((j / 2) & 1));
        const unsigned char *B = s_big + bs * AT_BIG_MAX;
        const AsmTileHdr *H = reinterpret_cast<const AsmTileHdr *>(B);
        const int nrows = H->nrows, nentries = H->nentries, npq = H->npq;
        const int64_t *row_out = reinterpret_cast<const int64_t *>(B + sizeof(AsmTileHdr));
        const uint32_t *codes = reinterpret_cast<const uint32_t *>(B + sizeof(AsmTileHdr) + pad_to(nrows, 2) * 8);
        const uint16_t *dcodes = reinterpret_cast<const uint16_t *>(B + sizeof(AsmTileHdr) + pad_to(nrows, 2) * 8 +
                                                                     pad_to(nentries, 4) * 4);
        __syncthreads();                                                            // [A] the tile's local matrices are complete
        // ---- next tile's vertex coordinates: in flight during phases D and 2
        if (have_next) {
            mbar_wait(&bar_geo[(j + 1) % 3], (uint32_t)(((j + 1) / 3) & 1));
            gather_coords((j + 1) % 3, bs ^ 1);
        }
        // ---- phase D: the diagonal of every row (its elements in ascending element id).  All table reads of a row
        //      are issued before the first add; rows shorter than the tile's longest read the zero slot.
        if (tid < nrows) {
            double d = 0.0;
            for (int q0 = 0; q0 < npq; q0 += 8) {
                uint32_t dc[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int q = q0 + u;
                    dc[u] = (q < npq && tid < (int)H->jds_cnt[q]) ? (uint32_t)dcodes[H->jds_off[q] + tid] : (uint32_t)AT_ZERO;
                }
                double v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = loc[dc[u]];
#pragma unroll
                for (int u = 0; u < 8; ++u) d += v[u];
            }
            loc[AT_DIAG_BASE + tid] = d;
        }
        __syncthreads();                                                            // [B]
        // ---- phase 2: one destination entry per thread and step, written straight into the CSR.  Two entries per
        //      iteration so that two independent chains (code -> two table reads -> add -> store) are in flight.
        {
            int r = 0, es0 = 0, es1 = H->run_es[1], rs0 = 0;                        // the current run lives in registers
            uint32_t len = H->run_len[0], magic = H->run_magic[0];
            auto locate = [&](int i, int &row, int &k) {
                while (i >= es1) {                                                  // i only grows: a forward scan
                    ++r;
                    es0 = es1; es1 = H->run_es[r + 1]; rs0 = H->run_rs[r];
                    len = H->run_len[r]; magic = H->run_magic[r];
                }
                const int within = i - es0;
                const int rr = len > 1 ? (int)__umulhi((uint32_t)within, magic) : within;
                row = rs0 + rr;
                k = within - rr * (int)len;
            };
            int i = tid;
            for (; i + AT_THREADS < nentries; i += 2 * AT_THREADS) {
                int rowA, kA, rowB, kB;
                locate(i, rowA, kA);
                locate(i + AT_THREADS, rowB, kB);
                const uint32_t cA = codes[i], cB = codes[i + AT_THREADS];
                const double a0 = loc[cA & 0xFFFFu], a1 = loc[cA >> 16];            // second = AT_ZERO when there is none
                const double b0 = loc[cB & 0xFFFFu], b1 = loc[cB >> 16];
                double *dA = values + row_out[rowA] + kA;
                double *dB = values + row_out[rowB] + kB;
                double vA = a0 + a1, vB = b0 + b1;
                if (ACC) { vA += *dA; vB += *dB; }
                *dA = vA;
                *dB = vB;
            }
            if (i < nentries) {
                int row, k;
                locate(i, row, k);
                const uint32_t c = codes[i];
                double v = loc[c & 0xFFFFu] + loc[c >> 16];
                double *dst = values + row_out[row] + k;
                if (ACC) v += *dst;
                *dst = v;
            }
        }
        cp_async_wait_all();             // next tile's coordinates have landed
        __syncthreads();                 // [C] loc, this tile's stages and tables are free again
    }
}

// ------------------------------------------------------------------------------------------------ plan builder
// Block-wide bitonic sort of n <= cap (cap a power of two) 32-bit keys in shared memory, ascending; the slots
// n .. cap-1 must already hold 0xFFFFFFFF.
__device__ void bitonic_sort_u32(uint32_t *a, int cap) {
    for (int k = 2; k <= cap; k <<= 1) {
        for (int jj = k >> 1; jj > 0; jj >>= 1) {
            for (int i = threadIdx.x; i < cap; i += blockDim.x) {
                const int ixj = i ^ jj;
                if (ixj > i) {
                    const uint32_t x = a[i], y = a[ixj];
                    const bool up = (i & k) == 0;
                    if ((x > y) == up) { a[i] = y; a[ixj] = x; }
                }
            }
            __syncthreads();
        }
    }
}

__device__ __forceinline__ int next_pow2(int n) { int c = 1; while (c < n) c <<= 1; return c; }

// keep the first of every run of equal keys of the sorted array a[0..n); returns the count (all threads).
// scratch: int32 flags/positions of size >= n + 1 in shared memory.
__device__ int unique_sorted_u32(uint32_t *a, int n, int32_t *scratch, uint32_t *out) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) scratch[i] = (i == 0 || a[i] != a[i - 1]) ? 1 : 0;
    __syncthreads();
    if (threadIdx.x == 0) {                                   // n <= a few thousand: a serial scan is setup-time noise
        int run = 0;
        for (int i = 0; i < n; ++i) { const int f = scratch[i]; scratch[i] = run; run += f; }
        scratch[n] = run;
    }
    __syncthreads();
    const int cnt = scratch[n];
    for (int i = threadIdx.x; i < n; i += blockDim.x)
        if (i == 0 || a[i] != a[i - 1]) out[scratch[i]] = a[i];
    __syncthreads();
    return cnt;
}

__device__ __forceinline__ int bsearch_u32(const uint32_t *a, int n, uint32_t key) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

constexpr int PB_PAIRCAP = 4096;          // pairs a tile may hold while it is being measured (>= 256 rows x 16)
constexpr int PB_VERTCAP = 1024;          // 3 x AT_ECAP = 864 vertex slots
enum { PB_OK = 0, PB_SPLIT = 1, PB_FATAL = 2 };   // SPLIT: halve the tile height; FATAL: this mesh needs the fallback kernel

struct PlanCounts { int32_t nelem, nvert, nentries, npairs; };

// One CTA per tile.  pass 0: measure (counts[tile], status flags); pass 1: write the blobs.
__global__ void __launch_bounds__(AT_THREADS) k_plan_tiles(
    int pass, int T, int64_t nrows_total, const int32_t *__restrict__ row_dof, const int64_t *__restrict__ dst_off,
    const int32_t *__restrict__ dst_len, const int32_t *__restrict__ dst_col, int64_t nt,
    const int32_t *__restrict__ t, const int32_t *__restrict__ edofs, const int64_t *__restrict__ r2e_ptr,
    const int32_t *__restrict__ r2e_idx, PlanCounts *__restrict__ counts, int32_t *__restrict__ status,
    const AsmTileRef *__restrict__ tiles, unsigned char *__restrict__ big, unsigned char *__restrict__ geo) {
    __shared__ uint32_t s_key[AT_THREADS];                 // row sort keys
    __shared__ int32_t s_np[AT_THREADS], s_len[AT_THREADS];   // per SORTED row
    __shared__ int32_t s_poff[AT_THREADS + 1], s_eoff[AT_THREADS + 1];
    __shared__ uint32_t s_elem[PB_PAIRCAP];
    __shared__ uint32_t s_uelem[AT_ECAP];
    __shared__ uint32_t s_vert[PB_VERTCAP];
    __shared__ uint32_t s_uvert[AT_VCAP];
    __shared__ int32_t s_scr[PB_PAIRCAP + 1];
    __shared__ int32_t s_bad;
    __shared__ AsmTileHdr s_hdr;
    __shared__ int32_t s_run_rows[AT_MAXRUNS];

    const int tid = threadIdx.x;
    const int64_t tile = blockIdx.x;
    const int64_t r0 = tile * T;
    const int nr = (int)(r0 + T < nrows_total ? T : nrows_total - r0);
    if (tid == 0) s_bad = PB_OK;
    // ---- 1. sort the tile's rows by (element count desc, length desc, position)
    int my_np = 0, my_len = 0;
    if (tid < nr) {
        const int32_t dof = row_dof[r0 + tid];
        my_np = (int)(r2e_ptr[dof + 1] - r2e_ptr[dof]);
        my_len = dst_len[r0 + tid];
    }
    __syncthreads();
    if (tid < nr && (my_np > AT_MAXP || my_len > AT_MAXLEN || my_np < 1 || my_len < 1)) s_bad = PB_FATAL;
    s_key[tid] = tid < nr ? ((uint32_t)(255 - (my_np > 255 ? 255 : my_np)) << 16) | ((uint32_t)(255 - (my_len > 255 ? 255 : my_len)) << 8) | (uint32_t)tid
                          : 0xFFFFFFFFu;
    __syncthreads();
    if (s_bad != PB_OK) { if (tid == 0) atomicMax(status, s_bad); return; }
    bitonic_sort_u32(s_key, AT_THREADS);
    int orig = 0;                                           // original in-tile position of sorted row `tid`
    if (tid < nr) {
        const uint32_t k = s_key[tid];
        orig = (int)(k & 255u);
        s_np[tid] = 255 - (int)((k >> 16) & 255u);
        s_len[tid] = 255 - (int)((k >> 8) & 255u);
    }
    __syncthreads();
    if (tid == 0) {
        int po = 0, eo = 0;
        for (int i = 0; i < nr; ++i) { s_poff[i] = po; s_eoff[i] = eo; po += s_np[i]; eo += s_len[i]; }
        s_poff[nr] = po; s_eoff[nr] = eo;
        // runs of equal length and the jagged-diagonal counts
        int nruns = 0;
        for (int i = 0; i < nr; ++i) {
            if (i == 0 || s_len[i] != s_len[i - 1]) {
                if (nruns == AT_MAXRUNS) { s_bad = PB_SPLIT; break; }
                s_hdr.run_len[nruns] = (uint8_t)s_len[i];
                s_run_rows[nruns] = 0;
                ++nruns;
            }
            ++s_run_rows[nruns - 1];
        }
        for (int r = nruns; r < AT_MAXRUNS; ++r) { s_hdr.run_len[r] = 0; s_run_rows[r] = 0; }
        {
            int es = 0, rs = 0;
            for (int r = 0; r < AT_MAXRUNS + 2; ++r) {
                s_hdr.run_es[r] = (uint16_t)es; s_hdr.run_rs[r] = (uint16_t)rs;
                if (r < nruns) {
                    const uint32_t len = s_hdr.run_len[r];
                    s_hdr.run_magic[r] = len > 1 ? (uint32_t)(0xFFFFFFFFu / len) + 1u : 0u;
                    es += (int)len * s_run_rows[r];
                    rs += s_run_rows[r];
                } else if (r < AT_MAXRUNS) {
                    s_hdr.run_magic[r] = 0u;
                }
            }
        }
        const int npq = nr ? s_np[0] : 0;
        {
            int o = 0;
            for (int q = 0; q < AT_MAXP + 2; ++q) {
                int c = 0;
                if (q < AT_MAXP)
                    for (int i = 0; i < nr; ++i) c += s_np[i] > q;
                s_hdr.jds_off[q] = (uint16_t)o;
                if (q < AT_MAXP) s_hdr.jds_cnt[q] = (uint16_t)c;
                o += c;
            }
        }
        s_hdr.nrows = nr; s_hdr.nentries = eo; s_hdr.npairs = po; s_hdr.nruns = nruns; s_hdr.npq = npq;
        s_hdr.pad0 = s_hdr.pad1 = s_hdr.pad2 = 0;
        for (int i = 0; i < 36; ++i) s_hdr.pad3[i] = 0;
        if (po > AT_DCAP || eo > AT_NCAP || po > PB_PAIRCAP) s_bad = PB_SPLIT;
    }
    __syncthreads();
    if (s_bad != PB_OK) { if (tid == 0) atomicMax(status, s_bad); return; }
    const int npairs = s_poff[nr], nentries = s_eoff[nr];
    // ---- 2. unique elements of the tile
    const int pcap = next_pow2(npairs);
    for (int i = tid; i < pcap; i += AT_THREADS) s_elem[i] = 0xFFFFFFFFu;
    __syncthreads();
    int32_t my_dof = 0;
    int64_t my_ptr = 0;
    if (tid < nr) {
        my_dof = row_dof[r0 + orig];
        my_ptr = r2e_ptr[my_dof];
        for (int q = 0; q < s_np[tid]; ++q) s_elem[s_poff[tid] + q] = (uint32_t)(r2e_idx[my_ptr + q] >> 3);
    }
    __syncthreads();
    bitonic_sort_u32(s_elem, pcap);
    // count first (the unique list has only AT_ECAP slots)
    for (int i = tid; i < npairs; i += AT_THREADS) s_scr[i] = (i == 0 || s_elem[i] != s_elem[i - 1]) ? 1 : 0;
    __syncthreads();
    if (tid == 0) {
        int c = 0;
        for (int i = 0; i < npairs; ++i) c += s_scr[i];
        if (c > AT_ECAP) s_bad = PB_SPLIT;
    }
    __syncthreads();
    if (s_bad != PB_OK) { if (tid == 0) atomicMax(status, s_bad); return; }
    const int nelem = unique_sorted_u32(s_elem, npairs, s_scr, s_uelem);
    // ---- 3. unique vertices of those elements
    const int vcap = next_pow2(3 * nelem);
    for (int i = tid; i < vcap; i += AT_THREADS) s_vert[i] = 0xFFFFFFFFu;
    __syncthreads();
    for (int e = tid; e < nelem; e += AT_THREADS) {
        const int64_t el = s_uelem[e];
        s_vert[3 * e] = (uint32_t)t[el]; s_vert[3 * e + 1] = (uint32_t)t[nt + el]; s_vert[3 * e + 2] = (uint32_t)t[2 * nt + el];
    }
    __syncthreads();
    bitonic_sort_u32(s_vert, vcap);
    for (int i = tid; i < 3 * nelem; i += AT_THREADS) s_scr[i] = (i == 0 || s_vert[i] != s_vert[i - 1]) ? 1 : 0;
    __syncthreads();
    if (tid == 0) {
        int c = 0;
        for (int i = 0; i < 3 * nelem; ++i) c += s_scr[i];
        if (c > AT_VCAP) s_bad = PB_SPLIT;
    }
    __syncthreads();
    if (s_bad != PB_OK) { if (tid == 0) atomicMax(status, s_bad); return; }
    const int nvert = unique_sorted_u32(s_vert, 3 * nelem, s_scr, s_uvert);
    if (pass == 0) {
        if (tid == 0) { counts[tile].nelem = nelem; counts[tile].nvert = nvert; counts[tile].nentries = nentries; counts[tile].npairs = npairs; }
        return;
    }
    // ---- 4. write the blobs
    const AsmTileRef ref = tiles[tile];
    unsigned char *Bg = big + ref.big_off;
    unsigned char *Gg = geo + ref.geo_off;
    {   // header
        const uint32_t *src = reinterpret_cast<const uint32_t *>(&s_hdr);
        uint32_t *dst = reinterpret_cast<uint32_t *>(Bg);
        for (int i = tid; i < (int)(sizeof(AsmTileHdr) / 4); i += AT_THREADS) dst[i] = src[i];
    }
    int64_t *g_row_out = reinterpret_cast<int64_t *>(Bg + sizeof(AsmTileHdr));
    uint32_t *g_codes = reinterpret_cast<uint32_t *>(Bg + sizeof(AsmTileHdr) + pad_to(nr, 2) * 8);
    uint16_t *g_dcodes = reinterpret_cast<uint16_t *>(Bg + sizeof(AsmTileHdr) + pad_to(nr, 2) * 8 + pad_to(nentries, 4) * 4);
    int32_t *g_ghdr = reinterpret_cast<int32_t *>(Gg);
    uint32_t *g_vlist = reinterpret_cast<uint32_t *>(Gg + 16);
    uchar4 *g_etri = reinterpret_cast<uchar4 *>(Gg + 16 + pad_to(nvert, 4) * 4);
    if (tid == 0) { g_ghdr[0] = nvert; g_ghdr[1] = nelem; g_ghdr[2] = 0; g_ghdr[3] = 0; }
    for (int v = tid; v < pad_to(nvert, 4); v += AT_THREADS) g_vlist[v] = v < nvert ? s_uvert[v] : 0u;
    for (int e = tid; e < pad_to(nelem, 4); e += AT_THREADS) {
        uchar4 tv = make_uchar4(0, 0, 0, 0);
        if (e < nelem) {
            const int64_t el = s_uelem[e];
            tv.x = (unsigned char)bsearch_u32(s_uvert, nvert, (uint32_t)t[el]);
            tv.y = (unsigned char)bsearch_u32(s_uvert, nvert, (uint32_t)t[nt + el]);
            tv.z = (unsigned char)bsearch_u32(s_uvert, nvert, (uint32_t)t[2 * nt + el]);
        }
        g_etri[e] = tv;
    }
    // padding words (never consumed, but the TMA copies them: keep them defined)
    if (tid == 0) {
        if (nr & 1) g_row_out[nr] = 0;
        for (int i = nentries; i < pad_to(nentries, 4); ++i) g_codes[i] = ((uint32_t)AT_ZERO << 16) | (uint32_t)AT_ZERO;
        for (int i = npairs; i < pad_to(npairs, 8); ++i) g_dcodes[i] = 0;
    }
    __syncthreads();
    if (tid < nr) {
        const int64_t doff = dst_off[r0 + orig];
        const int len = s_len[tid], np = s_np[tid];
        g_row_out[tid] = doff;
        uint16_t lo[AT_MAXLEN], hi[AT_MAXLEN];
        for (int s = 0; s < len; ++s) { lo[s] = (uint16_t)AT_ZERO; hi[s] = (uint16_t)AT_ZERO; }
        bool bad = false;
        for (int q = 0; q < np; ++q) {
            const int32_t code = r2e_idx[my_ptr + q];
            const int64_t el = code >> 3;
            const int li = code & 7;
            const int eloc = bsearch_u32(s_uelem, nelem, (uint32_t)el);
            g_dcodes[s_hdr.jds_off[q] + tid] = (uint16_t)(sym_index(li, li) * AT_LSTRIDE + eloc);
            for (int lj = 0; lj < 6; ++lj) {
                const int32_t c = edofs[(int64_t)lj * nt + el];
                int s = -1;
                for (int k = 0; k < len; ++k)
                    if (dst_col[doff + k] == c) { s = k; break; }
                if (s < 0) continue;                                    // column not kept by the destination layout
                if (c == my_dof) { lo[s] = (uint16_t)(AT_DIAG_BASE + tid); continue; }
                const uint16_t w = (uint16_t)(sym_index(li, lj) * AT_LSTRIDE + eloc);
                if (lo[s] == (uint16_t)AT_ZERO) lo[s] = w;
                else if (hi[s] == (uint16_t)AT_ZERO) hi[s] = w;
                else bad = true;                                        // three triangles on one edge: not a manifold mesh
            }
        }
        if (bad) atomicMax(status, (int)PB_FATAL);
        for (int s = 0; s < len; ++s) g_codes[s_eoff[tid] + s] = (uint32_t)lo[s] | ((uint32_t)hi[s] << 16);
    }
}

__global__ void k_plan_layout(int64_t ntiles, const PlanCounts *__restrict__ counts, int T, int64_t nrows_total,
                              int32_t *__restrict__ big_len, int32_t *__restrict__ geo_len) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ntiles) return;
    const int nr = (int)(i * T + T < nrows_total ? T : nrows_total - i * T);
    big_len[i] = big_bytes(nr, counts[i].nentries, counts[i].npairs);
    geo_len[i] = geo_bytes(counts[i].nvert, counts[i].nelem);
}

__global__ void k_plan_refs(int64_t ntiles, const int64_t *__restrict__ big_off, const int64_t *__restrict__ geo_off,
                            const int32_t *__restrict__ big_len, const int32_t *__restrict__ geo_len,
                            AsmTileRef *__restrict__ tiles) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ntiles) return;
    AsmTileRef r;
    r.big_off = big_off[i]; r.geo_off = geo_off[i]; r.big_len = big_len[i]; r.geo_len = geo_len[i];
    tiles[i] = r;
}

}  // namespace
}  // namespace mwx

using namespace mwx;

struct mwx_asm_plan {
    int64_t ntiles = 0, nrows = 0, big_total = 0, geo_total = 0;
    int T = 0;
    AsmTileRef *tiles = nullptr;
    unsigned char *big = nullptr, *geo = nullptr;
};

extern "C" void mwx_asm_plan_destroy(mwx_asm_plan_t *p) {
    if (!p) return;
    cudaFree(p->tiles); cudaFree(p->big); cudaFree(p->geo);
    delete p;
}

extern "C" int mwx_asm_plan_create(int64_t nrows, const int32_t *row_dof, const int64_t *dst_off, const int32_t *dst_len,
                                   const int32_t *dst_col, int64_t nt, const int32_t *t, const int32_t *edofs,
                                   const int64_t *r2e_ptr, const int32_t *r2e_idx, mwx_asm_plan_t **out, void *stream) {
    if (nrows <= 0 || nt <= 0 || !row_dof || !dst_off || !dst_len || !dst_col || !t || !edofs || !r2e_ptr || !r2e_idx || !out)
        return MWX_E_BADARG;
    cudaStream_t st = as_stream(stream);
    *out = nullptr;
    Scratch<int32_t> status(st);
    MWX_CUDA(status.alloc(1));
    for (int T = AT_THREADS; T >= 16; T >>= 1) {
        const int64_t ntiles = (nrows + T - 1) / T;
        if (ntiles >= ((int64_t)1 << 31)) return MWX_E_CAPACITY;
        Scratch<PlanCounts> counts(st);
        Scratch<int32_t> blen(st), glen(st);
        Scratch<int64_t> boff(st), goff(st);
        MWX_CUDA(counts.alloc((size_t)ntiles));
        MWX_CUDA(blen.alloc((size_t)ntiles)); MWX_CUDA(glen.alloc((size_t)ntiles));
        MWX_CUDA(boff.alloc((size_t)ntiles + 1)); MWX_CUDA(goff.alloc((size_t)ntiles + 1));
        MWX_CUDA(cudaMemsetAsync(status.p, 0, sizeof(int32_t), st));
        k_plan_tiles<<<(unsigned)ntiles, AT_THREADS, 0, st>>>(0, T, nrows, row_dof, dst_off, dst_len, dst_col, nt, t, edofs,
                                                             r2e_ptr, r2e_idx, counts.p, status.p, nullptr, nullptr, nullptr);
        MWX_LAUNCH_CHECK();
        int32_t hs = 0;
        MWX_CUDA(cudaMemcpyAsync(&hs, status.p, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        MWX_CUDA(cudaStreamSynchronize(st));
        if (hs == PB_FATAL) return MWX_E_CAPACITY;
        if (hs == PB_SPLIT) continue;
        k_plan_layout<<<grid_for(ntiles, 256), 256, 0, st>>>(ntiles, counts.p, T, nrows, blen.p, glen.p);
        MWX_LAUNCH_CHECK();
        int rc = exclusive_scan_i32_to_i64(blen.p, ntiles, boff.p, st);
        if (rc) return rc;
        rc = exclusive_scan_i32_to_i64(glen.p, ntiles, goff.p, st);
        if (rc) return rc;
        int64_t totals[2] = {0, 0};
        MWX_CUDA(cudaMemcpyAsync(&totals[0], boff.p + ntiles, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        MWX_CUDA(cudaMemcpyAsync(&totals[1], goff.p + ntiles, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        MWX_CUDA(cudaStreamSynchronize(st));
        auto *p = new (std::nothrow) mwx_asm_plan();
        if (!p) return MWX_E_NOMEM;
        p->ntiles = ntiles; p->nrows = nrows; p->T = T; p->big_total = totals[0]; p->geo_total = totals[1];
        if (cudaMalloc((void **)&p->tiles, sizeof(AsmTileRef) * (size_t)ntiles) != cudaSuccess ||
            cudaMalloc((void **)&p->big, (size_t)totals[0] + 16) != cudaSuccess ||
            cudaMalloc((void **)&p->geo, (size_t)totals[1] + 16) != cudaSuccess) {
            mwx_asm_plan_destroy(p);
            return MWX_E_NOMEM;
        }
        k_plan_refs<<<grid_for(ntiles, 256), 256, 0, st>>>(ntiles, boff.p, goff.p, blen.p, glen.p, p->tiles);
        k_plan_tiles<<<(unsigned)ntiles, AT_THREADS, 0, st>>>(1, T, nrows, row_dof, dst_off, dst_len, dst_col, nt, t, edofs,
                                                             r2e_ptr, r2e_idx, counts.p, status.p, p->tiles, p->big, p->geo);
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpyAsync(&hs, status.p, sizeof(int32_t), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { mwx_asm_plan_destroy(p); return (int)e; }
        if (hs != PB_OK) { mwx_asm_plan_destroy(p); return MWX_E_CAPACITY; }
        *out = p;
        return 0;
    }
    return MWX_E_CAPACITY;
}

extern "C" int mwx_asm_plan_info(const mwx_asm_plan_t *p, int64_t *info4) {
    if (!p || !info4) return MWX_E_BADARG;
    info4[0] = p->ntiles; info4[1] = p->T; info4[2] = p->big_total; info4[3] = p->geo_total;
    return 0;
}

extern "C" int mwx_assemble_morley_plan(const mwx_asm_plan_t *p, const double *pxy, double ca, double cb,
                                        int32_t accumulate, double *values, void *stream) {
    if (!p || !pxy || !values || ((uintptr_t)pxy & 15)) return MWX_E_BADARG;
    cudaStream_t st = as_stream(stream);
    static int ctas_per_sm = 0;
    if (!ctas_per_sm) {
        MWX_CUDA(cudaFuncSetAttribute(k_assemble_morley_tiles<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM_BYTES));
        MWX_CUDA(cudaFuncSetAttribute(k_assemble_morley_tiles<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM_BYTES));
        MWX_CUDA(cudaFuncSetAttribute(k_assemble_morley_tiles<false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        MWX_CUDA(cudaFuncSetAttribute(k_assemble_morley_tiles<true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        int occ = 0;
        MWX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_assemble_morley_tiles<false>, AT_THREADS, AT_SMEM_BYTES));
        ctas_per_sm = occ > 0 ? occ : 1;
    }
    const int64_t cap = (int64_t)sm_count() * ctas_per_sm;
    const unsigned grid = (unsigned)(p->ntiles < cap ? p->ntiles : cap);
    if (accumulate)
        k_assemble_morley_tiles<true><<<grid, AT_THREADS, AT_SMEM_BYTES, st>>>(p->ntiles, p->tiles, p->big, p->geo,
                                                                               (const double2 *)pxy, ca, cb, values);
    else
        k_assemble_morley_tiles<false><<<grid, AT_THREADS, AT_SMEM_BYTES, st>>>(p->ntiles, p->tiles, p->big, p->geo,
                                                                                (const double2 *)pxy, ca, cb, values);
    MWX_LAUNCH_CHECK();
    return 0;
}
