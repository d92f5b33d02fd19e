// Hot path A: CSR pattern / slot map (setup) and the Morley element kernel (hot).
//
// Replaces  K2 = epsilon**2 * asm(a_load, basis['u']) + asm(b_load, basis['u'])
// (ref main/example1/run.py:210; forms ref main/example1/run.py:110-123).  skfem tabulates the
// Morley basis by inverting a 6x6 Vandermonde per element in global monomials and then evaluates
// 36 Python form calls over (nt x nq) arrays into COO triplets that scipy sorts.  Here:
//
//   * the CSR pattern is built row-wise from the DOF->element incidence (no COO, no global sort);
//   * the 6x6 local matrices come from a closed form in the edge normals (no inverse, no quadrature:
//     hess is constant per element and grad.grad integrates exactly);
//   * the scatter is a GATHER: one thread owns one matrix row, walks the row's elements in ascending
//     element id, and accumulates that element's local row into a shared-memory image of the row's
//     CSR slots through a 1-byte-per-entry slot map.  No atomics, fixed summation order, and the
//     values leave the SM as one coalesced stream.
//
// Algorithmic bytes per element (the figure bench.py's roofline uses):
//   read  incidence codes 6*4 = 24, slot map 6*8 = 48, t 12, coordinates ~8, indptr+r2e_ptr ~32
//   write CSR values 23*8 = 184                                                   => 308 B/element
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "tma.cuh"

namespace mwx {
namespace {

constexpr int MAX_ROW = 255;             // slot index must fit in one byte
constexpr int ASM_ROWS = 128;            // rows (= threads) per CTA
constexpr int ASM_CAP = ASM_ROWS * 24;   // shared-memory CSR slots per CTA (24 KB of fp64)

// ------------------------------------------------------------------ pattern (setup)
__device__ __forceinline__ int gather_row_cols(int64_t r, int64_t nt, int ndpe, const int32_t *__restrict__ edofs,
                                               const int64_t *__restrict__ r2e_ptr,
                                               const int32_t *__restrict__ r2e_idx, int32_t *cols) {
    int n = 0;
    for (int64_t p = r2e_ptr[r]; p < r2e_ptr[r + 1]; ++p) {
        const int64_t e = r2e_idx[p] >> 3;
        for (int j = 0; j < ndpe; ++j) {
            const int32_t c = edofs[(int64_t)j * nt + e];
            int pos = 0;
            while (pos < n && cols[pos] < c) ++pos;
            if (pos < n && cols[pos] == c) continue;
            if (n == MAX_ROW) return -1;
            for (int k = n; k > pos; --k) cols[k] = cols[k - 1];
            cols[pos] = c;
            ++n;
        }
    }
    return n;
}

__global__ void __launch_bounds__(128) k_pattern_count(int64_t ndofs, int64_t nt, int ndpe,
                                                       const int32_t *__restrict__ edofs,
                                                       const int64_t *__restrict__ r2e_ptr,
                                                       const int32_t *__restrict__ r2e_idx,
                                                       int32_t *__restrict__ row_len,
                                                       int32_t *__restrict__ stats) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= ndofs) return;
    int32_t cols[MAX_ROW];
    const int n = gather_row_cols(r, nt, ndpe, edofs, r2e_ptr, r2e_idx, cols);
    if (n < 0) { atomicExch(&stats[0], 1); row_len[r] = 0; return; }
    row_len[r] = n;
    atomicMax(&stats[1], n);
}

__global__ void __launch_bounds__(128) k_pattern_fill(int64_t ndofs, int64_t nt, int ndpe,
                                                      const int32_t *__restrict__ edofs,
                                                      const int64_t *__restrict__ r2e_ptr,
                                                      const int32_t *__restrict__ r2e_idx,
                                                      const int64_t *__restrict__ indptr,
                                                      int32_t *__restrict__ indices,
                                                      unsigned long long *__restrict__ slot8) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= ndofs) return;
    int32_t cols[MAX_ROW];
    const int n = gather_row_cols(r, nt, ndpe, edofs, r2e_ptr, r2e_idx, cols);
    const int64_t s = indptr[r];
    for (int k = 0; k < n; ++k) indices[s + k] = cols[k];
    // Byte j of a pair's word = CSR slot (within the row) of the element's j-th DOF.  With at most six DOFs per
    // element, byte 7 carries a mask: bit j set = an EARLIER pair of this row already wrote that slot, so this
    // pair adds to it; clear = first touch, a plain store (the assembly kernels then need no zero-fill pass).
    unsigned long long seen[4] = {0ull, 0ull, 0ull, 0ull};
    for (int64_t p = r2e_ptr[r]; p < r2e_ptr[r + 1]; ++p) {
        const int64_t e = r2e_idx[p] >> 3;
        unsigned long long w = 0;
        for (int j = 0; j < ndpe; ++j) {
            const int32_t c = edofs[(int64_t)j * nt + e];
            int lo = 0, hi = n - 1;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (cols[mid] < c) lo = mid + 1; else hi = mid;
            }
            w |= (unsigned long long)(unsigned)lo << (8 * j);
            if (ndpe <= 6) {
                const unsigned long long bit = 1ull << (lo & 63);
                if (seen[lo >> 6] & bit) w |= 1ull << (56 + j);
                seen[lo >> 6] |= bit;
            }
        }
        slot8[p] = w;
    }
}

// ------------------------------------------------------------------ Morley local row (hot)
__device__ __forceinline__ double sel3(int i, double a, double b, double c) {
    return i == 0 ? a : (i == 1 ? b : c);
}

// Row `l` (0-2 vertex, 3-5 edge) of ca*A_K + cb*B_K for the triangle (x0,x1,x2) whose vertices are
// in ascending GLOBAL order (that fixes the edge-normal orientation n_k = R(x_a - x_b)/|.|, a<b).
//   H(psi_k) = 2 gam_k n_k n_k^T,  H(phi_i) = sum_k Dv_ik H(psi_k),  Dv_ik = -n_k . grad(lambda_i)
//   A = |K| H:H,   B_vv = |K| (g_i.g_j - 1/3 sum_k c_ik c_jk),  B_ee = |K|/3 I,  B_ve = 0.
// The geometry of one triangle as the element kernel needs it: gam_k (3), mu_km (3), |K| - seven doubles,
// stored with one pad double as a 64-byte record so the six (row, element) pairs of an element do the
// rsqrt / reciprocal work once (in k_element_geometry) instead of six times.
struct MorleyGeom { double g0, g1, g2, m01, m02, m12, area, pad; };

__device__ __forceinline__ MorleyGeom morley_geometry(double2 x0, double2 x1, double2 x2) {
    // edge vectors d_k = x_a - x_b for local edges (0,1),(1,2),(0,2); t_k = R d_k = (-d.y, d.x)
    const double d0x = x0.x - x1.x, d0y = x0.y - x1.y;
    const double d1x = x1.x - x2.x, d1y = x1.y - x2.y;
    const double d2x = x0.x - x2.x, d2y = x0.y - x2.y;
    const double det = d0x * d2y - d0y * d2x;                 // = (x1-x0) x (x2-x0), signed 2|K|
    const double rdet = 1.0 / det;
    const double area = 0.5 * fabs(det);
    const double q0 = d0x * d0x + d0y * d0y, q1 = d1x * d1x + d1y * d1y, q2 = d2x * d2x + d2y * d2y;
    const double r0 = rsqrt(q0), r1 = rsqrt(q1), r2 = rsqrt(q2);
    // gam_k = s_k |e_k| / (2|K|) with s = orient * (+,+,-)
    const double g0 = q0 * r0 * rdet, g1 = q1 * r1 * rdet, g2 = -(q2 * r2) * rdet;
    // mu_km = n_k . n_m  (rotation invariant -> use d directly)
    const double m01 = (d0x * d1x + d0y * d1y) * (r0 * r1);
    const double m02 = (d0x * d2x + d0y * d2y) * (r0 * r2);
    const double m12 = (d1x * d2x + d1y * d2y) * (r1 * r2);
    MorleyGeom G;
    G.g0 = g0; G.g1 = g1; G.g2 = g2; G.m01 = m01; G.m02 = m02; G.m12 = m12; G.area = area; G.pad = 0.0;
    return G;
}

__device__ __forceinline__ void morley_local_row(const MorleyGeom &G, int l, double ca, double cb, double out[6]) {
    const double g0 = G.g0, g1 = G.g1, g2 = G.g2, m01 = G.m01, m02 = G.m02, m12 = G.m12, area = G.area;
    // E_km = |K| (2 gam_k)(2 gam_m) mu_km^2  (edge-edge block of A)
    const double a4 = 4.0 * area;
    const double E00 = a4 * g0 * g0, E11 = a4 * g1 * g1, E22 = a4 * g2 * g2;
    const double E01 = a4 * g0 * g1 * m01 * m01, E02 = a4 * g0 * g2 * m02 * m02, E12 = a4 * g1 * g2 * m12 * m12;
    // Dv[i][k] = gam_{p(i)} mu_{k,p(i)} with p(0)=1, p(1)=2, p(2)=0  (vertex i is opposite edge p(i))
    const double D00 = g1 * m01, D01 = g1, D02 = g1 * m12;         // vertex 0 (opposite edge 1)
    const double D10 = g2 * m02, D11 = g2 * m12, D12 = g2;         // vertex 1 (opposite edge 2)
    const double D20 = g0, D21 = g0 * m01, D22 = g0 * m02;         // vertex 2 (opposite edge 0)
    const double third = area * (1.0 / 3.0);
    if (l >= 3) {
        const int m = l - 3;
        const double e0 = sel3(m, E00, E01, E02), e1 = sel3(m, E01, E11, E12), e2 = sel3(m, E02, E12, E22);
        out[0] = ca * (D00 * e0 + D01 * e1 + D02 * e2);
        out[1] = ca * (D10 * e0 + D11 * e1 + D12 * e2);
        out[2] = ca * (D20 * e0 + D21 * e1 + D22 * e2);
        out[3] = ca * e0 + (m == 0 ? cb * third : 0.0);
        out[4] = ca * e1 + (m == 1 ? cb * third : 0.0);
        out[5] = ca * e2 + (m == 2 ? cb * third : 0.0);
    } else {
        const double a0 = sel3(l, D00, D10, D20), a1 = sel3(l, D01, D11, D21), a2 = sel3(l, D02, D12, D22);
        // W_m = sum_k Dv_lk E_km
        const double w0 = a0 * E00 + a1 * E01 + a2 * E02;
        const double w1 = a0 * E01 + a1 * E11 + a2 * E12;
        const double w2 = a0 * E02 + a1 * E12 + a2 * E22;
        // g_l . g_j = gam_p(l) gam_p(j) mu_{p(l) p(j)};  p = (1,2,0)
        const double gl = sel3(l, g1, g2, g0);
        const double gg0 = gl * g1 * sel3(l, 1.0, m12, m01);      // j = 0 -> p = 1
        const double gg1 = gl * g2 * sel3(l, m12, 1.0, m02);      // j = 1 -> p = 2
        const double gg2 = gl * g0 * sel3(l, m01, m02, 1.0);      // j = 2 -> p = 0
        const double t3 = 1.0 / 3.0;
        out[0] = ca * (w0 * D00 + w1 * D01 + w2 * D02) + cb * area * (gg0 - t3 * (a0 * D00 + a1 * D01 + a2 * D02));
        out[1] = ca * (w0 * D10 + w1 * D11 + w2 * D12) + cb * area * (gg1 - t3 * (a0 * D10 + a1 * D11 + a2 * D12));
        out[2] = ca * (w0 * D20 + w1 * D21 + w2 * D22) + cb * area * (gg2 - t3 * (a0 * D20 + a1 * D21 + a2 * D22));
        out[3] = ca * w0;
        out[4] = ca * w1;
        out[5] = ca * w2;
    }
}

// ROWLIST = false: tile position i is matrix row i and the output is the global CSR (values + indptr).
// ROWLIST = true : position i is row row_ids[i]; the rows are written back to back through out_indptr
//                  (nrows + 1 entries) - a rank of the partitioned solver assembles only the rows it owns.
// Per-element geometry record (MorleyGeom, 64 bytes = two sectors per (row, element) pair instead of the six
// scattered sectors of t[3] + pxy[3]); recomputed from the vertex coordinates before every assembly.
__global__ void k_element_geometry(int64_t nt, const double2 *__restrict__ pxy, const int32_t *__restrict__ t,
                                   double2 *__restrict__ exy) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nt) return;
    const MorleyGeom G = morley_geometry(__ldg(pxy + t[e]), __ldg(pxy + t[nt + e]), __ldg(pxy + t[2 * nt + e]));
    exy[4 * e] = make_double2(G.g0, G.g1);
    exy[4 * e + 1] = make_double2(G.g2, G.m01);
    exy[4 * e + 2] = make_double2(G.m02, G.m12);
    exy[4 * e + 3] = make_double2(G.area, 0.0);
}

constexpr int ASM_MAXP = 8;               // (row, element) pairs per row staged in shared memory (row-list mode)
constexpr int ASM_PAIRS = ASM_MAXP * ASM_ROWS;   // pair words staged per CTA

// The 64-byte record is fetched with two 256-bit loads (sm_100 LDG.E.256: one 32-byte sector per thread and
// instruction).  Every thread of a warp reads a different record, so L1 handles one tag per thread per load:
// two instead of the four a 128-bit load sequence needs - the kernel's busiest unit is L1tex.
__device__ __forceinline__ MorleyGeom load_geom(const double2 *__restrict__ exy, int64_t e) {
    const double *p = reinterpret_cast<const double *>(exy + 4 * e);
    MorleyGeom G;
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(G.g0), "=d"(G.g1), "=d"(G.g2), "=d"(G.m01) : "l"(p));
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(G.m02), "=d"(G.m12), "=d"(G.area), "=d"(G.pad) : "l"(p + 4));
    return G;
}

// One thread owns one matrix row.  The tile's incidence codes and slot words - one contiguous range of the
// row-to-element arrays - are copied coalesced into shared memory (row-list mode: each thread stages its own
// row's run); the 64-byte geometry record of pair q+1 is in flight while pair q is evaluated; the six slot
// updates of a pair (distinct columns) go to the shared-memory image of the row: a plain store where the slot
// map says "first touch", load-add-store otherwise, so the image needs no zero-fill.  Rows are summed in
// ascending element order - no atomics, bit-reproducible.
template <bool ROWLIST>
__global__ void __launch_bounds__(ASM_ROWS, 6) k_assemble_morley(
    int64_t nrows, int64_t nt, const double2 *__restrict__ exy,
    const int64_t *__restrict__ r2e_ptr, const int32_t *__restrict__ r2e_idx,
    const unsigned long long *__restrict__ slot8, const int64_t *__restrict__ indptr,
    const int32_t *__restrict__ row_ids, const int64_t *__restrict__ out_indptr,
    double ca, double cb, int accumulate, double *__restrict__ values) {
    __shared__ double acc[ASM_CAP];
    __shared__ unsigned long long s_slot[ASM_PAIRS];
    __shared__ int32_t s_code[ASM_PAIRS];
    const int64_t *optr = ROWLIST ? out_indptr : indptr;
    const int64_t i0 = (int64_t)blockIdx.x * ASM_ROWS;
    const int64_t i1 = i0 + ASM_ROWS < nrows ? i0 + ASM_ROWS : nrows;
    const int64_t s0 = optr[i0], s1 = optr[i1];
    const int64_t len = s1 - s0;
    const bool staged = len <= ASM_CAP;          // CTA-uniform
    const int64_t i = i0 + threadIdx.x;
    const int tid = threadIdx.x;

    int64_t pb = 0;
    int np = 0;
    int64_t rs = 0;
    int mbase = 0, mstride = 1, mcount = 0;      // this row's staged pair words: s_*[mbase + q * mstride], q < mcount
    if (ROWLIST) {
        if (i < i1) {
            const int64_t r = (int64_t)row_ids[i];
            pb = r2e_ptr[r];
            np = (int)(r2e_ptr[r + 1] - pb);
            rs = optr[i];
            mbase = tid; mstride = ASM_ROWS; mcount = np < ASM_MAXP ? np : ASM_MAXP;
            for (int q = 0; q < mcount; ++q) {
                s_code[q * ASM_ROWS + tid] = __ldg(r2e_idx + pb + q);
                s_slot[q * ASM_ROWS + tid] = __ldg(slot8 + pb + q);
            }
        }
    } else {
        const int64_t pb0 = r2e_ptr[i0], pb1 = r2e_ptr[i1];
        const bool flat = pb1 - pb0 <= ASM_PAIRS;             // CTA-uniform
        if (flat) {
            const int cnt = (int)(pb1 - pb0);
            for (int k = tid; k < cnt; k += ASM_ROWS) {
                s_code[k] = __ldg(r2e_idx + pb0 + k);
                s_slot[k] = __ldg(slot8 + pb0 + k);
            }
        }
        if (i < i1) {
            pb = r2e_ptr[i];
            np = (int)(r2e_ptr[i + 1] - pb);
            rs = optr[i];
            if (flat) { mbase = (int)(pb - pb0); mcount = np; }
        }
    }
    __syncthreads();
    if (i < i1) {
        if (!staged && !accumulate) {
            const int64_t re = optr[i + 1];
            for (int64_t k = rs; k < re; ++k) values[k] = 0.0;
        }
        double *arow = acc + (rs - s0);
        MorleyGeom G;
        int32_t code = 0;
        if (np > 0) {
            code = 0 < mcount ? s_code[mbase] : __ldg(r2e_idx + pb);
            G = load_geom(exy, (int64_t)(code >> 3));
        }
        for (int q = 0; q < np; ++q) {
            const unsigned long long sl = q < mcount ? s_slot[mbase + q * mstride] : __ldg(slot8 + pb + q);
            const MorleyGeom Gc = G;
            const int l = code & 7;
            if (q + 1 < np) {                      // next pair's record is in flight during this pair's math
                code = q + 1 < mcount ? s_code[mbase + (q + 1) * mstride] : __ldg(r2e_idx + pb + q + 1);
                G = load_geom(exy, (int64_t)(code >> 3));
            }
            double row[6];
            morley_local_row(Gc, l, ca, cb, row);
            if (staged) {
                const unsigned rm = (unsigned)(sl >> 56);
                double o[6];
#pragma unroll
                for (int j = 0; j < 6; ++j) o[j] = (rm >> j) & 1u ? arow[(sl >> (8 * j)) & 0xffull] : 0.0;
#pragma unroll
                for (int j = 0; j < 6; ++j) arow[(sl >> (8 * j)) & 0xffull] = o[j] + row[j];
            } else {
#pragma unroll
                for (int j = 0; j < 6; ++j) values[rs + ((sl >> (8 * j)) & 0xffull)] += row[j];
            }
        }
    }
    if (staged) {
        __syncthreads();
        if (accumulate)
            for (int k = tid; k < (int)len; k += ASM_ROWS) values[s0 + k] += acc[k];
        else
            for (int k = tid; k < (int)len; k += ASM_ROWS) values[s0 + k] = acc[k];
    }
}

// ---- whole-matrix assembly: persistent, software-pipelined form of the kernel above ---------------------------
// The row kernel above is bound by a CHAIN of dependent DRAM latencies per CTA (tile bounds -> pair words ->
// first record -> ... ; ncu: 55 % of warp time on long_scoreboard, DRAM 36 %) rather than by bytes.
// Here CTAs are persistent (occupancy x 148) and stride over the 128-row tiles with everything a tile needs
// fetched one tile ahead, asynchronously, into the other of two shared-memory stages:
//   * the tile's pair words (incidence codes + slot words: one contiguous range each) by two TMA bulk copies
//     issued by one thread (mbarrier complete_tx), their range bounds loaded one further tile ahead;
//   * its row pointers (r2e_ptr, indptr segments) by per-thread 8-byte cp.async;
//   * and inside a tile each thread keeps TWO geometry records in flight.
// After that the kernel sits on the L1/shared-memory data pipe (74 %): every scattered 32-byte record sector and
// every conflicting row-image update costs a full wavefront.  Measured and rejected on top of this form (level 12,
// all within -8 .. +1 %): a three-stage ring with L2 prefetch of the next tile's records, streaming the previous
// tile out while the first records arrive, LDS-only addressing of the pair words, 5 CTAs/SM at 96 registers.
constexpr int ASMP_PAIRS = 800;                  // pair words staged per tile (768 = a tile of interior vertices)
constexpr int ASMP_CAP = ASM_ROWS * 20;          // shared-memory row image (2432 = a tile of interior vertices)
struct AsmStage {
    unsigned long long slot[ASMP_PAIRS + 2];     // staged from an even pair index (16-byte aligned source)
    int32_t code[ASMP_PAIRS + 8];                // staged from a multiple-of-4 pair index
    int64_t rptr[ASM_ROWS + 2];                  // r2e_ptr[i0 .. i1]
    int64_t optr[ASM_ROWS + 2];                  // indptr[i0 .. i1]
};
static_assert(sizeof(AsmStage) % 16 == 0, "stage must keep 16-byte alignment");
static_assert(offsetof(AsmStage, code) % 16 == 0, "TMA destination alignment");
constexpr int ASMP_SMEM_BYTES = ASMP_CAP * 8 + 2 * (int)sizeof(AsmStage) + 64;

__global__ void __launch_bounds__(ASM_ROWS) k_assemble_morley_persistent(
    int64_t nrows, const double2 *__restrict__ exy, const int64_t *__restrict__ r2e_ptr,
    const int32_t *__restrict__ r2e_idx, const unsigned long long *__restrict__ slot8,
    const int64_t *__restrict__ indptr, double ca, double cb, int accumulate, double *__restrict__ values,
    int tma_ok) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *acc = reinterpret_cast<double *>(smem_raw);
    AsmStage *stg = reinterpret_cast<AsmStage *>(smem_raw + ASMP_CAP * 8);
    unsigned char *tail = smem_raw + ASMP_CAP * 8 + 2 * sizeof(AsmStage);
    uint64_t *bar = reinterpret_cast<uint64_t *>(tail);                  // [2]
    int64_t *s_cb0 = reinterpret_cast<int64_t *>(tail + 16);            // [2] first staged code index
    int64_t *s_sb0 = reinterpret_cast<int64_t *>(tail + 32);            // [2] first staged slot-word index
    int32_t *s_flat = reinterpret_cast<int32_t *>(tail + 48);           // [2] pair words staged by TMA?

    const int tid = threadIdx.x;
    const int64_t ntiles = (nrows + ASM_ROWS - 1) / ASM_ROWS;
    const int64_t G = gridDim.x;
    int64_t npairs = 0;
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_fence_init();
        npairs = r2e_ptr[nrows];
    }
    auto bounds = [&](int64_t tile, int64_t &a, int64_t &b) {
        const int64_t i0 = tile * ASM_ROWS;
        const int64_t i1 = i0 + ASM_ROWS < nrows ? i0 + ASM_ROWS : nrows;
        a = r2e_ptr[i0];
        b = r2e_ptr[i1];
    };
    auto issue = [&](int64_t pb0, int64_t pb1, int stage) {              // thread 0
        const int64_t cb0 = pb0 & ~(int64_t)3, ce = (pb1 + 3) & ~(int64_t)3;
        const int64_t sb0 = pb0 & ~(int64_t)1, se = (pb1 + 1) & ~(int64_t)1;
        if (tma_ok && pb1 > pb0 && pb1 - pb0 <= ASMP_PAIRS && ce <= npairs && se <= npairs) {
            s_cb0[stage] = cb0;
            s_sb0[stage] = sb0;
            s_flat[stage] = 1;
            fence_proxy_async();
            mbar_expect_tx(&bar[stage], (uint32_t)((ce - cb0) * 4 + (se - sb0) * 8));
            bulk_g2s(stg[stage].code, r2e_idx + cb0, (uint32_t)((ce - cb0) * 4), &bar[stage]);
            bulk_g2s(stg[stage].slot, slot8 + sb0, (uint32_t)((se - sb0) * 8), &bar[stage]);
        } else {
            s_flat[stage] = 0;
        }
    };
    auto fetch_ptrs = [&](int64_t tile, int stage) {                     // all threads
        const int64_t i0 = tile * ASM_ROWS;
        const int64_t cnt = (i0 + ASM_ROWS < nrows ? ASM_ROWS : nrows - i0) + 1;   // entries i0 .. i1
        if (tid < cnt) {
            cp_async_8(&stg[stage].rptr[tid], r2e_ptr + i0 + tid);
            cp_async_8(&stg[stage].optr[tid], indptr + i0 + tid);
        }
        if (tid == 0 && cnt == ASM_ROWS + 1) {
            cp_async_8(&stg[stage].rptr[ASM_ROWS], r2e_ptr + i0 + ASM_ROWS);
            cp_async_8(&stg[stage].optr[ASM_ROWS], indptr + i0 + ASM_ROWS);
        }
        cp_async_commit();
    };

    int64_t tile = blockIdx.x;
    int64_t nb0 = 0, nb1 = 0;                 // thread 0: pair range of the NEXT tile, loaded one iteration early
    if (tile < ntiles) {
        if (tid == 0) {
            bounds(tile, nb0, nb1);
            issue(nb0, nb1, 0);
            if (tile + G < ntiles) bounds(tile + G, nb0, nb1);
        }
        fetch_ptrs(tile, 0);
    }
    cp_async_wait_all();
    __syncthreads();
    uint32_t phase0 = 0, phase1 = 0;
    for (int k = 0; tile < ntiles; tile += G, ++k) {
        const int stage = k & 1;
        const int64_t next = tile + G;
        if (next < ntiles) {
            if (tid == 0) {
                issue(nb0, nb1, stage ^ 1);
                if (next + G < ntiles) bounds(next + G, nb0, nb1);       // lands during this tile
            }
            fetch_ptrs(next, stage ^ 1);
        }
        const AsmStage &S = stg[stage];
        const int64_t i0 = tile * ASM_ROWS;
        const int nr = (int)(i0 + ASM_ROWS < nrows ? ASM_ROWS : nrows - i0);
        const int64_t s0 = S.optr[0];
        const int64_t len = S.optr[nr] - s0;
        const bool staged = len <= ASMP_CAP;         // CTA-uniform
        const bool flat = s_flat[stage] != 0;        // CTA-uniform
        if (flat) {
            if (stage == 0) { mbar_wait(&bar[0], phase0); phase0 ^= 1; }
            else            { mbar_wait(&bar[1], phase1); phase1 ^= 1; }
        }
        if (tid < nr) {
            const int64_t pb = S.rptr[tid];
            const int np = (int)(S.rptr[tid + 1] - pb);
            const int64_t rs = S.optr[tid];
            // one (generic) pointer each: shared-memory stage, or the global arrays for an oversized tile
            const int32_t *codes = flat ? S.code + (pb - s_cb0[stage]) : r2e_idx + pb;
            const unsigned long long *slots = flat ? S.slot + (pb - s_sb0[stage]) : slot8 + pb;
            if (!staged && !accumulate) {
                const int64_t re = S.optr[tid + 1];
                for (int64_t kk = rs; kk < re; ++kk) values[kk] = 0.0;
            }
            double *arow = acc + (rs - s0);
            MorleyGeom Ga, Gb;
            int32_t c0 = 0, c1 = 0;
            if (np > 0) { c0 = codes[0]; Ga = load_geom(exy, (int64_t)(c0 >> 3)); }
            if (np > 1) { c1 = codes[1]; Gb = load_geom(exy, (int64_t)(c1 >> 3)); }
            for (int q = 0; q < np; ++q) {
                const MorleyGeom Gc = Ga;
                const int l = c0 & 7;
                Ga = Gb;
                c0 = c1;
                if (q + 2 < np) {                  // two records in flight behind the one being evaluated
                    c1 = codes[q + 2];
                    Gb = load_geom(exy, (int64_t)(c1 >> 3));
                }
                const unsigned long long sl = slots[q];
                double row[6];
                morley_local_row(Gc, l, ca, cb, row);
                if (staged) {
                    const unsigned rm = (unsigned)(sl >> 56);
                    double o[6];
#pragma unroll
                    for (int j = 0; j < 6; ++j) o[j] = (rm >> j) & 1u ? arow[(sl >> (8 * j)) & 0xffull] : 0.0;
#pragma unroll
                    for (int j = 0; j < 6; ++j) arow[(sl >> (8 * j)) & 0xffull] = o[j] + row[j];
                } else {
#pragma unroll
                    for (int j = 0; j < 6; ++j) values[rs + ((sl >> (8 * j)) & 0xffull)] += row[j];
                }
            }
        }
        if (staged) {
            __syncthreads();
            if (accumulate)
                for (int kk = tid; kk < (int)len; kk += ASM_ROWS) values[s0 + kk] += acc[kk];
            else
                for (int kk = tid; kk < (int)len; kk += ASM_ROWS) values[s0 + kk] = acc[kk];
        }
        cp_async_wait_all();             // next tile's row pointers have landed
        __syncthreads();                 // stage and row image fully consumed before they are refilled
    }
}

// ------------------------------------------------------------------ boundary-facet penalty terms
// Row i of  c1*penalty_1 + c2*penalty_2 + c3*penalty_3  (ref main/example1/run.py:134-146) for the
// boundary facet that is local edge m of triangle (x0,x1,x2):
//   P_ij = |F| sum_q w_q [ -c1 hnn_j gn_i(q) - c2 hnn_i gn_j(q) + c3 (sigma/|F|) gn_i(q) gn_j(q) ]
// with hnn_i = nu^T H_i nu (constant), gn_i = grad f_i . nu (linear along the facet), nu the outward
// unit normal, and the 3-point Gauss rule skfem's FacetBasis picks (intorder 4).
// Basis in the edge-normal frame:  grad f_i = a_i - sum_k B_ik (2 lambda_{o(k)} - 1) n_k,
// H_i = sum_k B_ik 2 gam_k n_k n_k^T;  B = Dv for vertex functions, identity rows for edge functions.
__constant__ double c_gauss3_x[3] = {0.5 - 0.38729833462074170, 0.5, 0.5 + 0.38729833462074170};
__constant__ double c_gauss3_w[3] = {5.0 / 18.0, 8.0 / 18.0, 5.0 / 18.0};

__device__ void penalty_local_row(double2 x0, double2 x1, double2 x2, int m, int i, double c1, double c2,
                                  double c3, double sigma, double out[6]) {
    const double dx[3] = {x0.x - x1.x, x1.x - x2.x, x0.x - x2.x};
    const double dy[3] = {x0.y - x1.y, x1.y - x2.y, x0.y - x2.y};
    const double det = dx[0] * dy[2] - dy[0] * dx[2];
    double len[3], gam[3], mu[3][3];
    for (int k = 0; k < 3; ++k) len[k] = sqrt(dx[k] * dx[k] + dy[k] * dy[k]);
    gam[0] = len[0] / det; gam[1] = len[1] / det; gam[2] = -len[2] / det;
    for (int k = 0; k < 3; ++k)
        for (int l = 0; l < 3; ++l) mu[k][l] = (dx[k] * dx[l] + dy[k] * dy[l]) / (len[k] * len[l]);
    const int pv[3] = {1, 2, 0};                 // vertex i is opposite edge pv[i]
    const int ov[3] = {2, 0, 1};                 // edge k is opposite vertex ov[k]
    double B[6][3], an[6];                       // an_i = a_i . nu
    const double sg = gam[m] > 0 ? 1.0 : -1.0;   // nu = sg * n_m
    for (int v = 0; v < 3; ++v) {
        for (int k = 0; k < 3; ++k) B[v][k] = gam[pv[v]] * mu[k][pv[v]];
        an[v] = -gam[pv[v]] * mu[pv[v]][m] * sg;
    }
    for (int e = 0; e < 3; ++e) {
        for (int k = 0; k < 3; ++k) B[3 + e][k] = (k == e) ? 1.0 : 0.0;
        an[3 + e] = 0.0;
    }
    double hnn[6];
    for (int f = 0; f < 6; ++f) {
        double s = 0.0;
        for (int k = 0; k < 3; ++k) s += B[f][k] * 2.0 * gam[k] * mu[k][m] * mu[k][m];
        hnn[f] = s;
    }
    // facet runs from local vertex ea (xi = 0) to eb (xi = 1); lambda_{o(m)} = 0 on it
    const int ea = (m == 1) ? 1 : 0, eb = (m == 0) ? 1 : 2;
    for (int j = 0; j < 6; ++j) out[j] = 0.0;
    const double h = len[m];
    for (int q = 0; q < 3; ++q) {
        const double xi = c_gauss3_x[q];
        double lam[3] = {0.0, 0.0, 0.0};
        lam[ea] = 1.0 - xi; lam[eb] = xi;
        double gn[6];
        for (int f = 0; f < 6; ++f) {
            double s = an[f];
            for (int k = 0; k < 3; ++k) s -= B[f][k] * (2.0 * lam[ov[k]] - 1.0) * mu[k][m] * sg;
            gn[f] = s;
        }
        const double w = h * c_gauss3_w[q];
        for (int j = 0; j < 6; ++j)
            out[j] += w * (-c1 * hnn[j] * gn[i] - c2 * hnn[i] * gn[j] + c3 * (sigma / h) * gn[i] * gn[j]);
    }
}

__global__ void k_assemble_penalty(int64_t nt, const double2 *__restrict__ pxy, const int32_t *__restrict__ t,
                                   const int32_t *__restrict__ edofs, const int32_t *__restrict__ btri,
                                   const int32_t *__restrict__ blocal, int64_t nrows,
                                   const int32_t *__restrict__ rows, const int64_t *__restrict__ pair_ptr,
                                   const int32_t *__restrict__ pair_code, const int64_t *__restrict__ indptr,
                                   const int32_t *__restrict__ indices, double c1, double c2, double c3,
                                   double sigma, double *__restrict__ values, int32_t *__restrict__ err) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nrows) return;
    const int64_t r = rows[q];
    const int64_t rs = indptr[r], re = indptr[r + 1];
    for (int64_t p = pair_ptr[q]; p < pair_ptr[q + 1]; ++p) {
        const int32_t code = pair_code[p];
        const int64_t b = code >> 3;
        const int i = code & 7;
        const int64_t e = btri[b];
        const double2 x0 = pxy[t[e]], x1 = pxy[t[nt + e]], x2 = pxy[t[2 * nt + e]];
        double row[6];
        penalty_local_row(x0, x1, x2, blocal[b], i, c1, c2, c3, sigma, row);
        for (int j = 0; j < 6; ++j) {
            const int32_t c = edofs[(int64_t)j * nt + e];
            int64_t lo = rs, hi = re - 1;
            while (lo < hi) {
                const int64_t mid = (lo + hi) >> 1;
                if (indices[mid] < c) lo = mid + 1; else hi = mid;
            }
            if (lo < re && indices[lo] == c) values[lo] += row[j];
            else atomicExch(err, 1);
        }
    }
}

// L2pnvError (ref main/example1/run.py:106-108,527): sum over boundary facets of int_F (h_F n . grad u_h)^2.
// One thread per boundary facet writes its own term; the (tiny) sum is folded in index order.
__global__ void k_facet_pnv(int64_t nt, int64_t nbf, const double2 *__restrict__ pxy, const int32_t *__restrict__ t,
                            const int32_t *__restrict__ edofs, const int32_t *__restrict__ btri,
                            const int32_t *__restrict__ blocal, const double *__restrict__ uh,
                            double *__restrict__ terms) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nbf) return;
    const int64_t e = btri[b];
    double acc = 0.0;
    // n . grad u_h = sum_i u_i gn_i; gn_i is linear along the facet: reuse the penalty row machinery with
    // c3 = 1, sigma = h (so that the weight is gn_i gn_j |F| w_q): sum_ij u_i u_j P3_ij = int (n.grad u_h)^2
    const double2 x0 = pxy[t[e]], x1 = pxy[t[nt + e]], x2 = pxy[t[2 * nt + e]];
    const int m = blocal[b];
    double u[6];
    for (int j = 0; j < 6; ++j) u[j] = uh[edofs[(int64_t)j * nt + e]];
    const double dx = (m == 0 ? x0.x - x1.x : (m == 1 ? x1.x - x2.x : x0.x - x2.x));
    const double dy = (m == 0 ? x0.y - x1.y : (m == 1 ? x1.y - x2.y : x0.y - x2.y));
    const double h = sqrt(dx * dx + dy * dy);
    for (int i = 0; i < 6; ++i) {
        double row[6];
        penalty_local_row(x0, x1, x2, m, i, 0.0, 0.0, 1.0, h, row);      // (sigma / h) = 1
        for (int j = 0; j < 6; ++j) acc += u[i] * row[j] * u[j];
    }
    terms[b] = h * h * acc;
}

__global__ void k_axpby(int64_t n, double alpha, const double *__restrict__ a, double beta,
                        const double *__restrict__ b, double *__restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = alpha * a[i] + beta * b[i];
}

}  // namespace
}  // namespace mwx

using namespace mwx;

extern "C" int mwx_pattern_count(int64_t ndofs, int64_t nt, int32_t ndpe, const int32_t *edofs, const int64_t *r2e_ptr,
                                 const int32_t *r2e_idx, int64_t *indptr, int64_t *nnz_host,
                                 int32_t *max_row_host, void *stream) {
    if (ndofs <= 0 || nt <= 0 || ndpe < 1 || ndpe > 6 || !edofs || !r2e_ptr || !r2e_idx || !indptr || !nnz_host)
        return MWX_E_BADARG;
    cudaStream_t st = as_stream(stream);
    Scratch<int32_t> len(st), stats(st);
    MWX_CUDA(len.alloc((size_t)ndofs));
    MWX_CUDA(stats.alloc(2));
    MWX_CUDA(cudaMemsetAsync(stats.p, 0, 2 * sizeof(int32_t), st));
    k_pattern_count<<<grid_for(ndofs, 128), 128, 0, st>>>(ndofs, nt, ndpe, edofs, r2e_ptr, r2e_idx, len.p, stats.p);
    MWX_LAUNCH_CHECK();
    int rc = exclusive_scan_i32_to_i64(len.p, ndofs, indptr, st);
    if (rc) return rc;
    int32_t hs[2] = {0, 0};
    MWX_CUDA(cudaMemcpyAsync(hs, stats.p, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    MWX_CUDA(cudaMemcpyAsync(nnz_host, indptr + ndofs, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    MWX_CUDA(cudaStreamSynchronize(st));
    if (max_row_host) *max_row_host = hs[1];
    return hs[0] ? MWX_E_CAPACITY : 0;
}

extern "C" int mwx_pattern_fill(int64_t ndofs, int64_t nt, int32_t ndpe, const int32_t *edofs, const int64_t *r2e_ptr,
                                const int32_t *r2e_idx, const int64_t *indptr, int32_t *indices,
                                uint8_t *slot8, void *stream) {
    if (ndofs <= 0 || nt <= 0 || !edofs || !r2e_ptr || !r2e_idx || !indptr || !indices || !slot8)
        return MWX_E_BADARG;
    if ((uintptr_t)slot8 & 7) return MWX_E_BADARG;
    if (ndpe < 1 || ndpe > 6) return MWX_E_BADARG;
    k_pattern_fill<<<grid_for(ndofs, 128), 128, 0, as_stream(stream)>>>(
        ndofs, nt, ndpe, edofs, r2e_ptr, r2e_idx, indptr, indices, (unsigned long long *)slot8);
    MWX_LAUNCH_CHECK();
    return 0;
}

static int refresh_element_coords(int64_t nt, const double *pxy, const int32_t *t, double *exy, cudaStream_t st) {
    k_element_geometry<<<grid_for(nt, 256), 256, 0, st>>>(nt, (const double2 *)pxy, t, (double2 *)exy);
    MWX_LAUNCH_CHECK();
    return 0;
}

extern "C" int mwx_assemble_morley(int64_t ndofs, int64_t nt, const double *pxy, const int32_t *t, double *exy,
                                   int32_t refresh_exy, const int64_t *r2e_ptr, const int32_t *r2e_idx,
                                   const uint8_t *slot8, const int64_t *indptr, double ca, double cb,
                                   int32_t accumulate, double *values, void *stream) {
    if (ndofs <= 0 || nt <= 0 || !pxy || !t || !exy || !r2e_ptr || !r2e_idx || !slot8 || !indptr || !values)
        return MWX_E_BADARG;
    if (((uintptr_t)slot8 & 7) || ((uintptr_t)pxy & 15) || ((uintptr_t)exy & 63)) return MWX_E_BADARG;
    cudaStream_t st = as_stream(stream);
    if (refresh_exy) {
        int rc = refresh_element_coords(nt, pxy, t, exy, st);
        if (rc) return rc;
    }
    static const bool tile_kernel = [] { const char *e = getenv("MWX_ASM_KERNEL"); return e && !strcmp(e, "tile"); }();
    if (tile_kernel) {                               // A/B switch: the one-CTA-per-tile kernel
        const unsigned g = (unsigned)((ndofs + ASM_ROWS - 1) / ASM_ROWS);
        k_assemble_morley<false><<<g, ASM_ROWS, 0, st>>>(
            ndofs, nt, (const double2 *)exy, r2e_ptr, r2e_idx, (const unsigned long long *)slot8, indptr, nullptr,
            nullptr, ca, cb, accumulate, values);
        MWX_LAUNCH_CHECK();
        return 0;
    }
    static int ctas_per_sm = 0;
    if (!ctas_per_sm) {
        MWX_CUDA(cudaFuncSetAttribute(k_assemble_morley_persistent, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      ASMP_SMEM_BYTES));
        MWX_CUDA(cudaFuncSetAttribute(k_assemble_morley_persistent, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        int occ = 0;
        MWX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_assemble_morley_persistent, ASM_ROWS,
                                                               ASMP_SMEM_BYTES));
        ctas_per_sm = occ > 0 ? occ : 1;
    }
    const int64_t ntiles = (ndofs + ASM_ROWS - 1) / ASM_ROWS;
    const int64_t cap = (int64_t)sm_count() * ctas_per_sm;
    const unsigned grid = (unsigned)(ntiles < cap ? ntiles : cap);
    const int tma_ok = !(((uintptr_t)r2e_idx | (uintptr_t)slot8) & 15);
    k_assemble_morley_persistent<<<grid, ASM_ROWS, ASMP_SMEM_BYTES, st>>>(
        ndofs, (const double2 *)exy, r2e_ptr, r2e_idx, (const unsigned long long *)slot8, indptr, ca, cb, accumulate,
        values, tma_ok);
    MWX_LAUNCH_CHECK();
    return 0;
}

extern "C" int mwx_assemble_morley_rows(int64_t nrows, const int32_t *row_ids, const int64_t *out_indptr,
                                        int64_t nt, const double *pxy, const int32_t *t, double *exy,
                                        int32_t refresh_exy, const int64_t *r2e_ptr, const int32_t *r2e_idx,
                                        const uint8_t *slot8, const int64_t *indptr, double ca, double cb,
                                        double *values, void *stream) {
    if (nrows < 0 || nt <= 0 || !pxy || !t || !exy || !r2e_ptr || !r2e_idx || !slot8 || !indptr || !values)
        return MWX_E_BADARG;
    if (nrows == 0) return 0;
    if (!row_ids || !out_indptr || ((uintptr_t)slot8 & 7) || ((uintptr_t)pxy & 15) || ((uintptr_t)exy & 63))
        return MWX_E_BADARG;
    cudaStream_t st = as_stream(stream);
    if (refresh_exy) {
        int rc = refresh_element_coords(nt, pxy, t, exy, st);
        if (rc) return rc;
    }
    const unsigned grid = (unsigned)((nrows + ASM_ROWS - 1) / ASM_ROWS);
    k_assemble_morley<true><<<grid, ASM_ROWS, 0, st>>>(
        nrows, nt, (const double2 *)exy, r2e_ptr, r2e_idx, (const unsigned long long *)slot8, indptr, row_ids,
        out_indptr, ca, cb, 0, values);
    MWX_LAUNCH_CHECK();
    return 0;
}

extern "C" int mwx_axpby(int64_t n, double alpha, const double *a, double beta, const double *b,
                         double *out, void *stream) {
    if (n < 0 || (n && (!a || !b || !out))) return MWX_E_BADARG;
    if (n == 0) return 0;
    int64_t g = (n + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (g > cap) g = cap;
    k_axpby<<<(unsigned)g, 256, 0, as_stream(stream)>>>(n, alpha, a, beta, b, out);
    MWX_LAUNCH_CHECK();
    return 0;
}

extern "C" int mwx_assemble_penalty(int64_t nt, int64_t nbf, const double *pxy, const int32_t *t,
                                    const int32_t *edofs, const int32_t *btri, const int32_t *blocal,
                                    int64_t nrows, const int32_t *rows, const int64_t *pair_ptr,
                                    const int32_t *pair_code, const int64_t *indptr, const int32_t *indices,
                                    double c1, double c2, double c3, double sigma, double *values,
                                    void *stream) {
    if (nt <= 0 || nbf < 0 || nrows < 0 || !pxy || !t || !edofs || !indptr || !indices || !values) return MWX_E_BADARG;
    if (nrows == 0) return 0;
    if (!btri || !blocal || !rows || !pair_ptr || !pair_code) return MWX_E_BADARG;
    cudaStream_t st = as_stream(stream);
    Scratch<int32_t> err(st);
    MWX_CUDA(err.alloc(1));
    MWX_CUDA(cudaMemsetAsync(err.p, 0, sizeof(int32_t), st));
    k_assemble_penalty<<<grid_for(nrows, 64), 64, 0, st>>>(nt, (const double2 *)pxy, t, edofs, btri, blocal, nrows,
                                                           rows, pair_ptr, pair_code, indptr, indices, c1, c2, c3,
                                                           sigma, values, err.p);
    MWX_LAUNCH_CHECK();
    int32_t herr = 0;
    MWX_CUDA(cudaMemcpyAsync(&herr, err.p, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    MWX_CUDA(cudaStreamSynchronize(st));
    return herr ? MWX_E_BADARG : 0;
}

extern "C" int mwx_morley_facet_pnv(int64_t nt, int64_t nbf, const double *pxy, const int32_t *t, const int32_t *edofs,
                                    const int32_t *btri, const int32_t *blocal, const double *uh, double *terms,
                                    void *stream) {
    if (nt <= 0 || nbf <= 0 || !pxy || !t || !edofs || !btri || !blocal || !uh || !terms) return MWX_E_BADARG;
    k_facet_pnv<<<grid_for(nbf, 64), 64, 0, as_stream(stream)>>>(nt, nbf, (const double2 *)pxy, t, edofs, btri, blocal, uh,
                                                                 terms);
    MWX_LAUNCH_CHECK();
    return 0;
}
