// Device-resident AMG hierarchy shared by amg.cu (upload of a host hierarchy, V-cycle, PCG) and amg_setup_dev.cu
// (smoothed-aggregation setup on the device).
#pragma once
#include <vector>

#include "krylov.cuh"

namespace mwx {

// Block-CSR twin of an operator whose nonzeros come in br x bc blocks (bsr.cu): planes of nblocks doubles.
struct DBsr {
    int br = 0, bc = 0, lanes = 8;
    int64_t nbrows = 0, nblocks = 0;
    int64_t *bptr = nullptr;
    int32_t *bidx = nullptr;
    double *bval = nullptr;
    float *bval32 = nullptr;                   // fp32-stored planes (then bval == nullptr), see bsr_store_fp32
};

struct DCsr {
    int64_t rows = 0, cols = 0, nnz = 0;
    int64_t plan = -1;                         // -(planned SpMV tile height), see spmv_plan_rows
    bool borrowed = false;                     // the arrays belong to the caller (fine-level operator of a device setup)
    int64_t *ptr = nullptr;
    int32_t *idx = nullptr;
    double *val = nullptr;
    DBsr bsr;                                  // set (bptr != nullptr) when the V-cycle should use the block kernels
    PackedNz *pk = nullptr;                    // fp32-stored twin of a scalar operator (krylov.cuh), used instead of idx/val
    SpmvLX *lx = nullptr;                      // tile-local column plan of a scalar operator (spmv_lx.cu)
};

// One direction of the exact lexicographic Gauss-Seidel sweep, level-scheduled: rows grouped by dependency wavefront,
// and for every row, in that order, one 16-byte record plus its off-diagonal entries (CSR order) laid end to end - a
// wavefront is three contiguous streams, so the next wavefront can be prefetched while the current one computes.
struct GsRow { int32_t row, len; int64_t start; };
struct GsSweep {
    int64_t nwaves = 0, nrows = 0;
    int64_t *wave_ptr = nullptr;        // (nwaves + 1) offsets into rows[]
    GsRow *rows = nullptr;              // row id, number of off-diagonal entries, offset of the first one
    int32_t *col = nullptr;
    double *val = nullptr;
    double *diag = nullptr;             // per row, wavefront order
};

struct DLevel {
    DCsr A, P, R;
    double *dinv = nullptr, *x = nullptr, *x2 = nullptr, *b = nullptr, *r = nullptr, *d = nullptr;
    // parity mode: the operator packed in Gauss-Seidel wavefront order, once per sweep direction (see GsSweep)
    GsSweep gs_f, gs_b;
    double rho = 0.0;                   // estimate of the spectral radius of D^-1 A
    double *g_b = nullptr, *g_x = nullptr;   // right-hand side / correction of the second visit of a gamma = 2 cycle
    // smoothed-aggregation levels built on the device keep what the partitioned solver asks for
    int64_t *roots = nullptr;           // first fine DOF of every aggregate (ascending)
    int64_t nroots = 0;
    int32_t coarse_per_root = 0;
    double *cand = nullptr;             // near-nullspace candidates of this level (rows x ncand)
    int32_t ncand = 0;
};

void free_csr(DCsr &m);
void free_bsr(DBsr &B);
int bsr_from_csr(const DCsr &A, int br, int bc, DBsr &out, cudaStream_t st);
int bsr_store_fp32(DBsr &B, cudaStream_t st);
int bsr_spmv(const DBsr &B, const double *x, double *y, cudaStream_t st);
int bsr_spmv_axpy(const DBsr &B, const double *x, double *y, const double *c, int sign, cudaStream_t st);
int bsr_smoother_step(const DBsr &B, const double *x_in, double *x_out, const double *b, const double *dinv, double *d,
                      double c1, double c2, cudaStream_t st);

// Work vectors, inverse diagonals, SpMV plans and (fast smoothers) spectral-radius estimates of every level whose
// operators are already on the device; dense pseudo-inverse of the coarsest operator (coarse_pinv_host, or computed
// here by cyclic Jacobi when null - coarsest level of at most a few thousand rows).
}  // namespace mwx

struct mwx_amg_dev {
    std::vector<mwx::DLevel> lv;
    double *pinv = nullptr;              // dense (n_last x n_last), row-major
    int64_t n_last = 0;
    int smoother = 0, degree = 2;
    double cheb_ratio = 10.0;
    int storage_bits = 64;               // 32: the V-cycle's operators are stored in fp32 (mwx_amg_set_storage)
    int gamma = 1, gamma_first = 1, gamma_last = 0;    // gamma = 2: levels gamma_first .. gamma_last are visited twice per visit of their parent
    mwx::CgScalars *sc = nullptr;
    double *partials = nullptr;
};

namespace mwx {
int amg_finalize_device_levels(mwx_amg_dev *d, const double *coarse_pinv_host, cudaStream_t st);
// L.rho = spectral radius of D^-1 A by power iteration (needs L.dinv, L.x2, L.r, L.A.plan and d->sc / d->partials).
int amg_estimate_rho(mwx_amg_dev *d, DLevel &L, cudaStream_t st);
}
