// the callers either side of the MWX hot path, on device.
//
//   * K1 = asm(laplace, basis['w'])                 ref main/example1/run.py:149-154,198   (P1 / P2 Lagrange)
//   * f1 = asm(f_load, basis['w'])                  ref main/example1/run.py:277-288,199   (load sampled at the
//                                                    quadrature points on the host: problem data stays Python)
//   * f2 = asm(wv_load, basis['w'], basis['u']) * wh ref main/example1/run.py:126-131,211   (applied, not stored)
//   * L2uError / get_DuError / get_D2uError         ref main/example1/run.py:157-179,519-531
//
// All of them reuse the row-gather skeleton of the Morley kernel: one thread per output row walks the row's
// (element, local index) incidence in ascending element order - deterministic, no atomics.
#include "common.cuh"

namespace mwx {
namespace {

struct Tri {
    double2 g[3];      // grad lambda_i
    double area, det;
};

__device__ __forceinline__ Tri tri_geometry(double2 x0, double2 x1, double2 x2) {
    Tri T;
    T.det = (x1.x - x0.x) * (x2.y - x0.y) - (x1.y - x0.y) * (x2.x - x0.x);
    const double rd = 1.0 / T.det;
    T.area = 0.5 * fabs(T.det);
    // grad lambda_i = R(x_{i+2} - x_{i+1}) / det,  R(x, y) = (-y, x)
    T.g[0] = make_double2(-(x2.y - x1.y) * rd, (x2.x - x1.x) * rd);
    T.g[1] = make_double2(-(x0.y - x2.y) * rd, (x0.x - x2.x) * rd);
    T.g[2] = make_double2(-(x1.y - x0.y) * rd, (x1.x - x0.x) * rd);
    return T;
}

__device__ __forceinline__ double dot2(double2 a, double2 b) { return a.x * b.x + a.y * b.y; }

// local edges of P2 / Morley: k -> (a, b) = (0,1), (1,2), (0,2)
__device__ __forceinline__ void edge_ends(int k, int &a, int &b) { a = (k == 1) ? 1 : 0; b = (k == 0) ? 1 : 2; }

// Row l of the P1 (ndpe = 3) or P2 (ndpe = 6) stiffness of one triangle (closed form; exact for affine maps).
__device__ void lagrange_laplace_row(const Tri &T, int ndpe, int l, double out[6]) {
    double G[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) G[i][j] = dot2(T.g[i], T.g[j]);
    if (ndpe == 3) {
        for (int j = 0; j < 3; ++j) out[j] = T.area * G[l][j];
        return;
    }
    const double A = T.area;
    if (l < 3) {
        for (int j = 0; j < 3; ++j) out[j] = A * G[l][j] * (j == l ? 1.0 : -1.0 / 3.0);
        for (int k = 0; k < 3; ++k) {
            int a, b;
            edge_ends(k, a, b);
            out[3 + k] = (4.0 * A / 3.0) * ((l == a ? G[l][b] : 0.0) + (l == b ? G[l][a] : 0.0));
        }
    } else {
        int a, b;
        edge_ends(l - 3, a, b);
        for (int j = 0; j < 3; ++j) out[j] = (4.0 * A / 3.0) * ((j == a ? G[j][b] : 0.0) + (j == b ? G[j][a] : 0.0));
        for (int k = 0; k < 3; ++k) {
            int c, d;
            edge_ends(k, c, d);
            out[3 + k] = (16.0 * A / 12.0) * ((1 + (a == c)) * G[b][d] + (1 + (a == d)) * G[b][c] +
                                              (1 + (b == c)) * G[a][d] + (1 + (b == d)) * G[a][c]);
        }
    }
}

__global__ void __launch_bounds__(128) k_assemble_lagrange(int64_t nrows, int64_t nt, int ndpe,
                                                           const double2 *__restrict__ pxy,
                                                           const int32_t *__restrict__ t,
                                                           const int64_t *__restrict__ r2e_ptr,
                                                           const int32_t *__restrict__ r2e_idx,
                                                           const unsigned long long *__restrict__ slot8,
                                                           const int64_t *__restrict__ indptr,
                                                           double *__restrict__ values) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    const int64_t rs = indptr[r], re = indptr[r + 1];
    for (int64_t k = rs; k < re; ++k) values[k] = 0.0;
    for (int64_t p = r2e_ptr[r]; p < r2e_ptr[r + 1]; ++p) {
        const int32_t code = r2e_idx[p];
        const unsigned long long sl = slot8[p];
        const int64_t e = code >> 3;
        const Tri T = tri_geometry(pxy[t[e]], pxy[t[nt + e]], pxy[t[2 * nt + e]]);
        double row[6];
        lagrange_laplace_row(T, ndpe, code & 7, row);
        for (int j = 0; j < ndpe; ++j) values[rs + ((sl >> (8 * j)) & 0xffull)] += row[j];
    }
}

// ---- problem data evaluated on the device (optional: `problem` > 0 and no host samples).  The reference's loads and
// exact solutions are Python closures (ref main/example1/run.py:275-309 'ex1', :466-493 'ex3'); sampling them on the
// host is fine for the shipped ladders (<= 263 k DOFs) but not for a 33.5 M-triangle mesh (235 M points, 11 GB of
// samples for the error norms), so the two problems the examples use are also available as device functions.
//   problem 1 ('ex1'): u = (sin pi x sin pi y)^2,  f = eps^2 lap^2 u - lap u
//   problem 3 ('ex3'): u = sin pi x sin pi y,      f = 2 pi^2 u   (the reference's 'ex3' load has no eps term)
constexpr double kPi = 3.14159265358979323846;

__device__ __forceinline__ double problem_load(int problem, double eps, double x, double y) {
    double sx, cx, sy, cy;
    sincospi(x, &sx, &cx);
    sincospi(y, &sy, &cy);
    if (problem == 1) {
        const double c2x = cx * cx - sx * sx, c2y = cy * cy - sy * sy;             // cos(2 pi x), cos(2 pi y)
        const double g = c2x * sy * sy + c2y * sx * sx;
        const double lu = 2.0 * kPi * kPi * g;
        const double llu = -8.0 * kPi * kPi * kPi * kPi * (g - c2x * c2y);
        return eps * eps * llu - lu;
    }
    return 2.0 * kPi * kPi * sx * sy;
}

// out = {u, ux, uy, uxx, uxy, uyy}
__device__ __forceinline__ void problem_exact(int problem, double x, double y, double out[6]) {
    double sx, cx, sy, cy;
    sincospi(x, &sx, &cx);
    sincospi(y, &sy, &cy);
    if (problem == 1) {
        out[0] = sx * sx * sy * sy;
        out[1] = 2.0 * kPi * cx * sx * sy * sy;
        out[2] = 2.0 * kPi * cy * sy * sx * sx;
        out[3] = 2.0 * kPi * kPi * (cx * cx - sx * sx) * sy * sy;
        out[4] = 4.0 * kPi * kPi * cx * sx * cy * sy;
        out[5] = 2.0 * kPi * kPi * (cy * cy - sy * sy) * sx * sx;
    } else {
        out[0] = sx * sy;
        out[1] = kPi * cx * sy;
        out[2] = kPi * cy * sx;
        out[3] = -kPi * kPi * sx * sy;
        out[4] = kPi * kPi * cx * cy;
        out[5] = -kPi * kPi * sx * sy;
    }
}

__device__ __forceinline__ double problem_exact_value(int problem, double x, double y) {
    double u6[6];
    problem_exact(problem, x, y, u6);
    return u6[0];
}

// reference basis values at X = (xi, eta): lambda = (1 - xi - eta, xi, eta)
__device__ __forceinline__ double lagrange_value(int ndpe, int l, double l0, double l1, double l2) {
    const double lam[3] = {l0, l1, l2};
    if (ndpe == 3) return lam[l];
    if (l < 3) return lam[l] * (2.0 * lam[l] - 1.0);
    int a, b;
    edge_ends(l - 3, a, b);
    return 4.0 * lam[a] * lam[b];
}

// f1[r] = sum over incident elements, sum_q fq[e, q] phi_l(X_q) |det| W_q
__global__ void k_load_vector(int64_t nrows, int64_t nt, int ndpe, int nq, const double *__restrict__ XW,
                              const double2 *__restrict__ pxy, const int32_t *__restrict__ t,
                              const int64_t *__restrict__ r2e_ptr, const int32_t *__restrict__ r2e_idx,
                              const double *__restrict__ fq, int problem, double eps, double *__restrict__ out) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    double acc = 0.0;
    for (int64_t p = r2e_ptr[r]; p < r2e_ptr[r + 1]; ++p) {
        const int32_t code = r2e_idx[p];
        const int64_t e = code >> 3;
        const int l = code & 7;
        const double2 x0 = pxy[t[e]], x1 = pxy[t[nt + e]], x2 = pxy[t[2 * nt + e]];
        const double adet = fabs((x1.x - x0.x) * (x2.y - x0.y) - (x1.y - x0.y) * (x2.x - x0.x));
        double s = 0.0;
        for (int q = 0; q < nq; ++q) {
            const double xi = XW[q], eta = XW[nq + q], w = XW[2 * nq + q];
            const double f = fq ? fq[e * nq + q]
                                : problem_load(problem, eps, x0.x + (x1.x - x0.x) * xi + (x2.x - x0.x) * eta,
                                               x0.y + (x1.y - x0.y) * xi + (x2.y - x0.y) * eta);
            s += f * lagrange_value(ndpe, l, 1.0 - xi - eta, xi, eta) * (adet * w);
        }
        acc += s;
    }
    out[r] = acc;
}

// Morley basis in the edge-normal frame (see assembly.cu): per triangle gam_k, unit normals n_k, B[6][3], a[6].
struct MorleyFrame {
    double2 n[3], a[6];
    double gam[3], B[6][3], area;
};

__device__ void morley_frame(double2 x0, double2 x1, double2 x2, MorleyFrame &F) {
    const double2 d[3] = {make_double2(x0.x - x1.x, x0.y - x1.y), make_double2(x1.x - x2.x, x1.y - x2.y),
                          make_double2(x0.x - x2.x, x0.y - x2.y)};
    const double det = d[0].x * d[2].y - d[0].y * d[2].x;
    F.area = 0.5 * fabs(det);
    double len[3];
    for (int k = 0; k < 3; ++k) {
        len[k] = sqrt(d[k].x * d[k].x + d[k].y * d[k].y);
        F.n[k] = make_double2(-d[k].y / len[k], d[k].x / len[k]);
    }
    F.gam[0] = len[0] / det; F.gam[1] = len[1] / det; F.gam[2] = -len[2] / det;
    const int pv[3] = {1, 2, 0};                     // vertex i is opposite edge pv[i]: grad lambda_i = -gam n
    for (int v = 0; v < 3; ++v) {
        F.a[v] = make_double2(-F.gam[pv[v]] * F.n[pv[v]].x, -F.gam[pv[v]] * F.n[pv[v]].y);
        for (int k = 0; k < 3; ++k) F.B[v][k] = F.gam[pv[v]] * dot2(F.n[k], F.n[pv[v]]);
    }
    for (int e = 0; e < 3; ++e) {
        F.a[3 + e] = make_double2(0.0, 0.0);
        for (int k = 0; k < 3; ++k) F.B[3 + e][k] = (k == e) ? 1.0 : 0.0;
    }
}

// grad f_i at barycentrics lam:  a_i - sum_k B_ik (2 lam_{o(k)} - 1) n_k,  o = (2, 0, 1)
__device__ __forceinline__ double2 morley_grad(const MorleyFrame &F, int i, const double lam[3]) {
    const int ov[3] = {2, 0, 1};
    double2 g = F.a[i];
    for (int k = 0; k < 3; ++k) {
        const double c = F.B[i][k] * (2.0 * lam[ov[k]] - 1.0);
        g.x -= c * F.n[k].x;
        g.y -= c * F.n[k].y;
    }
    return g;
}

// value of f_i: vertex: lam_i + sum_k B_ik psi_k, edge m: psi_m;  psi_k = lam_o (lam_o - 1) / gam_k
__device__ __forceinline__ double morley_value(const MorleyFrame &F, int i, const double lam[3]) {
    const int ov[3] = {2, 0, 1};
    double v = i < 3 ? lam[i] : 0.0;
    for (int k = 0; k < 3; ++k) v += F.B[i][k] * lam[ov[k]] * (lam[ov[k]] - 1.0) / F.gam[k];
    return v;
}

// f2[r] = sum over incident elements of  int grad(w_h) . grad(phi_l)   (phi_l Morley, w_h P1 or P2):
// degree <= 2 integrand, integrated exactly by the three edge-midpoint points.
__global__ void k_wv_action(int64_t nrows, int64_t nt, int ndpe_w, const double2 *__restrict__ pxy,
                            const int32_t *__restrict__ t, const int32_t *__restrict__ wdofs,
                            const int64_t *__restrict__ r2e_ptr, const int32_t *__restrict__ r2e_idx,
                            const double *__restrict__ wh, double *__restrict__ out) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    double acc = 0.0;
    for (int64_t p = r2e_ptr[r]; p < r2e_ptr[r + 1]; ++p) {
        const int32_t code = r2e_idx[p];
        const int64_t e = code >> 3;
        const int l = code & 7;
        const double2 x0 = pxy[t[e]], x1 = pxy[t[nt + e]], x2 = pxy[t[2 * nt + e]];
        MorleyFrame F;
        morley_frame(x0, x1, x2, F);
        const Tri T = tri_geometry(x0, x1, x2);
        double w[6];
        for (int j = 0; j < ndpe_w; ++j) w[j] = wh[wdofs[(int64_t)j * nt + e]];
        double s = 0.0;
        for (int q = 0; q < 3; ++q) {
            double lam[3] = {0.5, 0.5, 0.5};
            lam[(q + 2) % 3] = 0.0;                      // (1/2,1/2,0), (0,1/2,1/2), (1/2,0,1/2)
            double2 gw = make_double2(0.0, 0.0);
            if (ndpe_w == 3) {
                for (int i = 0; i < 3; ++i) { gw.x += w[i] * T.g[i].x; gw.y += w[i] * T.g[i].y; }
            } else {
                for (int i = 0; i < 3; ++i) {
                    const double c = w[i] * (4.0 * lam[i] - 1.0);
                    gw.x += c * T.g[i].x; gw.y += c * T.g[i].y;
                }
                for (int k = 0; k < 3; ++k) {
                    int a, b;
                    edge_ends(k, a, b);
                    gw.x += 4.0 * w[3 + k] * (lam[a] * T.g[b].x + lam[b] * T.g[a].x);
                    gw.y += 4.0 * w[3 + k] * (lam[a] * T.g[b].y + lam[b] * T.g[a].y);
                }
            }
            s += dot2(gw, morley_grad(F, l, lam));
        }
        acc += s * (T.area / 3.0);
    }
    out[r] = acc;
}

// Per-element squared errors of a Morley function against samples of the exact solution at the quadrature
// points: sums[0] = int (u_h - u)^2, sums[1] = int |grad u_h - grad u|^2, sums[2] = int |hess u_h - hess u|^2_F.
// ex: 6 arrays of nt*nq (u, ux, uy, uxx, uxy, uyy).  Deterministic: per-element values, block partials in order.
__global__ void __launch_bounds__(128) k_morley_errors(int64_t nt, int nq, const double *__restrict__ XW,
                                                       const double2 *__restrict__ pxy,
                                                       const int32_t *__restrict__ t,
                                                       const int32_t *__restrict__ edofs,
                                                       const double *__restrict__ uh,
                                                       const double *__restrict__ ex, int problem,
                                                       double *__restrict__ partials) {
    __shared__ double sh[32];
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    if (e < nt) {
        const double2 x0 = pxy[t[e]], x1 = pxy[t[nt + e]], x2 = pxy[t[2 * nt + e]];
        MorleyFrame F;
        morley_frame(x0, x1, x2, F);
        double u[6];
        for (int j = 0; j < 6; ++j) u[j] = uh[edofs[(int64_t)j * nt + e]];
        // Hessian (constant): sum_i u_i sum_k B_ik 2 gam_k n_k n_k^T
        double hxx = 0.0, hxy = 0.0, hyy = 0.0;
        for (int k = 0; k < 3; ++k) {
            double c = 0.0;
            for (int i = 0; i < 6; ++i) c += u[i] * F.B[i][k];
            c *= 2.0 * F.gam[k];
            hxx += c * F.n[k].x * F.n[k].x; hxy += c * F.n[k].x * F.n[k].y; hyy += c * F.n[k].y * F.n[k].y;
        }
        const double adet = 2.0 * F.area;
        for (int q = 0; q < nq; ++q) {
            const double xi = XW[q], eta = XW[nq + q], w = XW[2 * nq + q] * adet;
            const double lam[3] = {1.0 - xi - eta, xi, eta};
            double val = 0.0, gx = 0.0, gy = 0.0;
            for (int i = 0; i < 6; ++i) {
                val += u[i] * morley_value(F, i, lam);
                const double2 g = morley_grad(F, i, lam);
                gx += u[i] * g.x; gy += u[i] * g.y;
            }
            const int64_t o = e * nq + q, N = nt * (int64_t)nq;
            double u6[6];
            if (ex) {
                for (int c = 0; c < 6; ++c) u6[c] = ex[c * N + o];
            } else {
                problem_exact(problem, x0.x + (x1.x - x0.x) * xi + (x2.x - x0.x) * eta,
                              x0.y + (x1.y - x0.y) * xi + (x2.y - x0.y) * eta, u6);
            }
            const double dv = val - u6[0], dgx = gx - u6[1], dgy = gy - u6[2];
            const double dxx = hxx - u6[3], dxy = hxy - u6[4], dyy = hyy - u6[5];
            s0 += dv * dv * w;
            s1 += (dgx * dgx + dgy * dgy) * w;
            s2 += (dxx * dxx + 2.0 * dxy * dxy + dyy * dyy) * w;
        }
    }
    const double b0 = block_sum(s0, sh), b1 = block_sum(s1, sh), b2 = block_sum(s2, sh);
    if (threadIdx.x == 0) {
        partials[3 * (int64_t)blockIdx.x] = b0;
        partials[3 * (int64_t)blockIdx.x + 1] = b1;
        partials[3 * (int64_t)blockIdx.x + 2] = b2;
    }
}

// main/example2 (ref main/example2/run.py:101-172): a Morley function of a COARSE mesh sampled at the quadrature points
// of a FINE mesh obtained from it by uniform refinement.  The children of triangle k sit at k, k + nt, k + 2 nt, k + 3 nt
// after every refinement (mesh_builder.cu, as skfem), so the coarse triangle holding fine triangle T is T mod nt_c - the
// reference's one-point-at-a-time ``basis.interpolator(u)(x)`` becomes index arithmetic.  out6 = (u, ux, uy, uxx, uxy, uyy),
// six arrays of nt_f * nq - the layout k_morley_errors takes for its comparison samples.  *bad counts points that do not
// lie in the assumed triangle (meshes not nested): the caller turns that into an error.
__global__ void __launch_bounds__(128) k_morley_sample_nested(int64_t nt_f, int nq, const double *__restrict__ XW,
                                                              const double2 *__restrict__ pxy_f,
                                                              const int32_t *__restrict__ t_f, int64_t nt_c,
                                                              const double2 *__restrict__ pxy_c,
                                                              const int32_t *__restrict__ t_c,
                                                              const int32_t *__restrict__ edofs_c,
                                                              const double *__restrict__ uh_c,
                                                              double *__restrict__ out6, int *__restrict__ bad) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nt_f) return;
    const int64_t a = e % nt_c;
    const double2 y0 = pxy_f[t_f[e]], y1 = pxy_f[t_f[nt_f + e]], y2 = pxy_f[t_f[2 * nt_f + e]];
    const double2 x0 = pxy_c[t_c[a]], x1 = pxy_c[t_c[nt_c + a]], x2 = pxy_c[t_c[2 * nt_c + a]];
    MorleyFrame F;
    morley_frame(x0, x1, x2, F);
    double u[6];
    for (int j = 0; j < 6; ++j) u[j] = uh_c[edofs_c[(int64_t)j * nt_c + a]];
    double hxx = 0.0, hxy = 0.0, hyy = 0.0;
    for (int k = 0; k < 3; ++k) {
        double c = 0.0;
        for (int i = 0; i < 6; ++i) c += u[i] * F.B[i][k];
        c *= 2.0 * F.gam[k];
        hxx += c * F.n[k].x * F.n[k].x; hxy += c * F.n[k].x * F.n[k].y; hyy += c * F.n[k].y * F.n[k].y;
    }
    const double det = (x1.x - x0.x) * (x2.y - x0.y) - (x1.y - x0.y) * (x2.x - x0.x);
    const int64_t N = nt_f * (int64_t)nq;
    int nbad = 0;
    for (int q = 0; q < nq; ++q) {
        const double xi = XW[q], eta = XW[nq + q];
        const double px = y0.x + (y1.x - y0.x) * xi + (y2.x - y0.x) * eta - x0.x;
        const double py = y0.y + (y1.y - y0.y) * xi + (y2.y - y0.y) * eta - x0.y;
        const double l1 = (px * (x2.y - x0.y) - py * (x2.x - x0.x)) / det;
        const double l2 = (py * (x1.x - x0.x) - px * (x1.y - x0.y)) / det;
        const double lam[3] = {1.0 - l1 - l2, l1, l2};
        nbad += (lam[0] < -1e-9) | (lam[1] < -1e-9) | (lam[2] < -1e-9);
        double val = 0.0, gx = 0.0, gy = 0.0;
        for (int i = 0; i < 6; ++i) {
            val += u[i] * morley_value(F, i, lam);
            const double2 g = morley_grad(F, i, lam);
            gx += u[i] * g.x; gy += u[i] * g.y;
        }
        const int64_t o = e * nq + q;
        out6[o] = val; out6[N + o] = gx; out6[2 * N + o] = gy;
        out6[3 * N + o] = hxx; out6[4 * N + o] = hxy; out6[5 * N + o] = hyy;
    }
    if (nbad) atomicAdd(bad, nbad);
}

// Morley mass matrix rows (project(exact_u, basis_to=basis['u']), ref main/example3/run.py:66 +
// main/example3/utils.py:494-565): quartic integrand, integrated with the basis' quadrature rule.
__global__ void __launch_bounds__(128) k_morley_mass(int64_t nrows, int64_t nt, int nq, const double *__restrict__ XW,
                                                     const double2 *__restrict__ pxy, const int32_t *__restrict__ t,
                                                     const int64_t *__restrict__ r2e_ptr,
                                                     const int32_t *__restrict__ r2e_idx,
                                                     const unsigned long long *__restrict__ slot8,
                                                     const int64_t *__restrict__ indptr, double *__restrict__ values) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    const int64_t rs = indptr[r], re = indptr[r + 1];
    for (int64_t k = rs; k < re; ++k) values[k] = 0.0;
    for (int64_t p = r2e_ptr[r]; p < r2e_ptr[r + 1]; ++p) {
        const int32_t code = r2e_idx[p];
        const unsigned long long sl = slot8[p];
        const int64_t e = code >> 3;
        const int l = code & 7;
        MorleyFrame F;
        morley_frame(pxy[t[e]], pxy[t[nt + e]], pxy[t[2 * nt + e]], F);
        double row[6] = {0, 0, 0, 0, 0, 0};
        for (int q = 0; q < nq; ++q) {
            const double xi = XW[q], eta = XW[nq + q], w = XW[2 * nq + q] * 2.0 * F.area;
            const double lam[3] = {1.0 - xi - eta, xi, eta};
            const double vl = morley_value(F, l, lam) * w;
            for (int j = 0; j < 6; ++j) row[j] += vl * morley_value(F, j, lam);
        }
        for (int j = 0; j < 6; ++j) values[rs + ((sl >> (8 * j)) & 0xffull)] += row[j];
    }
}

__global__ void k_morley_load_vector(int64_t nrows, int64_t nt, int nq, const double *__restrict__ XW,
                                     const double2 *__restrict__ pxy, const int32_t *__restrict__ t,
                                     const int64_t *__restrict__ r2e_ptr, const int32_t *__restrict__ r2e_idx,
                                     const double *__restrict__ fq, int problem, double eps,
                                     double *__restrict__ out) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    double acc = 0.0;
    for (int64_t p = r2e_ptr[r]; p < r2e_ptr[r + 1]; ++p) {
        const int32_t code = r2e_idx[p];
        const int64_t e = code >> 3;
        const double2 x0 = pxy[t[e]], x1 = pxy[t[nt + e]], x2 = pxy[t[2 * nt + e]];
        MorleyFrame F;
        morley_frame(x0, x1, x2, F);
        double s = 0.0;
        for (int q = 0; q < nq; ++q) {
            const double xi = XW[q], eta = XW[nq + q], w = XW[2 * nq + q] * 2.0 * F.area;
            const double lam[3] = {1.0 - xi - eta, xi, eta};
            const double f = fq ? fq[e * nq + q]
                                : problem_exact_value(problem, x0.x + (x1.x - x0.x) * xi + (x2.x - x0.x) * eta,
                                                      x0.y + (x1.y - x0.y) * xi + (x2.y - x0.y) * eta);
            s += f * morley_value(F, code & 7, lam) * w;
        }
        acc += s;
    }
    out[r] = acc;
}

__global__ void k_fold3(int64_t nb, const double *__restrict__ partials, double *__restrict__ out) {
    __shared__ double sh[32];
    for (int c = 0; c < 3; ++c) {
        double s = 0.0;
        for (int64_t i = threadIdx.x; i < nb; i += blockDim.x) s += partials[3 * i + c];
        const double tot = block_sum(s, sh);
        if (threadIdx.x == 0) out[c] = tot;
    }
}

}  // namespace
}  // namespace mwx

using namespace mwx;

extern "C" int mwx_assemble_lagrange_laplace(int64_t nrows, int64_t nt, int32_t ndpe, const double *pxy,
                                             const int32_t *t, const int64_t *r2e_ptr, const int32_t *r2e_idx,
                                             const uint8_t *slot8, const int64_t *indptr, double *values,
                                             void *stream) {
    if (nrows <= 0 || nt <= 0 || (ndpe != 3 && ndpe != 6) || !pxy || !t || !r2e_ptr || !r2e_idx || !slot8 || !indptr || !values)
        return MWX_E_BADARG;
    k_assemble_lagrange<<<grid_for(nrows, 128), 128, 0, as_stream(stream)>>>(
        nrows, nt, ndpe, (const double2 *)pxy, t, r2e_ptr, r2e_idx, (const unsigned long long *)slot8, indptr, values);
    MWX_LAUNCH_CHECK();
    return 0;
}

extern "C" int mwx_load_vector(int64_t nrows, int64_t nt, int32_t ndpe, int32_t nq, const double *xw,
                               const double *pxy, const int32_t *t, const int64_t *r2e_ptr,
                               const int32_t *r2e_idx, const double *fq, double *out, void *stream) {
    if (nrows <= 0 || nt <= 0 || (ndpe != 3 && ndpe != 6) || nq <= 0 || !xw || !pxy || !t || !r2e_ptr || !r2e_idx || !fq || !out)
        return MWX_E_BADARG;
    k_load_vector<<<grid_for(nrows, 128), 128, 0, as_stream(stream)>>>(nrows, nt, ndpe, nq, xw, (const double2 *)pxy, t,
                                                                     r2e_ptr, r2e_idx, fq, 0, 0.0, out);
    MWX_LAUNCH_CHECK();
    return 0;
}

extern "C" int mwx_load_vector_problem(int32_t problem, double eps, int64_t nrows, int64_t nt, int32_t ndpe, int32_t nq,
                                       const double *xw, const double *pxy, const int32_t *t, const int64_t *r2e_ptr,
                                       const int32_t *r2e_idx, double *out, void *stream) {
    if ((problem != 1 && problem != 3) || nrows <= 0 || nt <= 0 || (ndpe != 3 && ndpe != 6) || nq <= 0 || !xw || !pxy || !t ||
        !r2e_ptr || !r2e_idx || !out)
        return MWX_E_BADARG;
    k_load_vector<<<grid_for(nrows, 128), 128, 0, as_stream(stream)>>>(nrows, nt, ndpe, nq, xw, (const double2 *)pxy, t,
                                                                     r2e_ptr, r2e_idx, nullptr, problem, eps, out);
    MWX_LAUNCH_CHECK();
    return 0;
}

extern "C" int mwx_wv_action(int64_t nrows, int64_t nt, int32_t ndpe_w, const double *pxy, const int32_t *t,
                             const int32_t *wdofs, const int64_t *r2e_ptr, const int32_t *r2e_idx,
                             const double *wh, double *out, void *stream) {
    if (nrows <= 0 || nt <= 0 || (ndpe_w != 3 && ndpe_w != 6) || !pxy || !t || !wdofs || !r2e_ptr || !r2e_idx || !wh || !out)
        return MWX_E_BADARG;
    k_wv_action<<<grid_for(nrows, 128), 128, 0, as_stream(stream)>>>(nrows, nt, ndpe_w, (const double2 *)pxy, t, wdofs,
                                                                   r2e_ptr, r2e_idx, wh, out);
    MWX_LAUNCH_CHECK();
    return 0;
}

extern "C" int mwx_morley_error_sums(int64_t nt, int32_t nq, const double *xw, const double *pxy, const int32_t *t,
                                     const int32_t *edofs, const double *uh, const double *exact6, double *sums3,
                                     void *stream) {
    if (nt <= 0 || nq <= 0 || !xw || !pxy || !t || !edofs || !uh || !exact6 || !sums3) return MWX_E_BADARG;
    cudaStream_t st = as_stream(stream);
    const int64_t nb = (nt + 127) / 128;
    Scratch<double> partials(st);
    MWX_CUDA(partials.alloc((size_t)(3 * nb)));
    k_morley_errors<<<(unsigned)nb, 128, 0, st>>>(nt, nq, xw, (const double2 *)pxy, t, edofs, uh, exact6, 0, partials.p);
    MWX_LAUNCH_CHECK();
    k_fold3<<<1, 1024, 0, st>>>(nb, partials.p, sums3);
    MWX_LAUNCH_CHECK();
    return 0;
}

extern "C" int mwx_morley_error_sums_problem(int32_t problem, int64_t nt, int32_t nq, const double *xw, const double *pxy,
                                             const int32_t *t, const int32_t *edofs, const double *uh, double *sums3,
                                             void *stream) {
    if ((problem != 1 && problem != 3) || nt <= 0 || nq <= 0 || !xw || !pxy || !t || !edofs || !uh || !sums3) return MWX_E_BADARG;
    cudaStream_t st = as_stream(stream);
    const int64_t nb = (nt + 127) / 128;
    Scratch<double> partials(st);
    MWX_CUDA(partials.alloc((size_t)(3 * nb)));
    k_morley_errors<<<(unsigned)nb, 128, 0, st>>>(nt, nq, xw, (const double2 *)pxy, t, edofs, uh, nullptr, problem, partials.p);
    MWX_LAUNCH_CHECK();
    k_fold3<<<1, 1024, 0, st>>>(nb, partials.p, sums3);
    MWX_LAUNCH_CHECK();
    return 0;
}

extern "C" int mwx_morley_sample_nested(int64_t nt_f, int32_t nq, const double *xw, const double *pxy_f, const int32_t *t_f,
                                        int64_t nt_c, const double *pxy_c, const int32_t *t_c, const int32_t *edofs_c,
                                        const double *uh_c, double *out6, void *stream) {
    if (nt_f <= 0 || nt_c <= 0 || nt_f % nt_c || nq <= 0 || !xw || !pxy_f || !t_f || !pxy_c || !t_c || !edofs_c || !uh_c || !out6)
        return MWX_E_BADARG;
    cudaStream_t st = as_stream(stream);
    Scratch<int> bad(st);
    MWX_CUDA(bad.alloc(1));
    MWX_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), st));
    k_morley_sample_nested<<<grid_for(nt_f, 128), 128, 0, st>>>(nt_f, nq, xw, (const double2 *)pxy_f, t_f, nt_c,
                                                                 (const double2 *)pxy_c, t_c, edofs_c, uh_c, out6, bad.p);
    MWX_LAUNCH_CHECK();
    int h = 0;
    MWX_CUDA(cudaMemcpyAsync(&h, bad.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    MWX_CUDA(cudaStreamSynchronize(st));
    return h ? MWX_E_BADARG : 0;
}

extern "C" int mwx_assemble_morley_mass(int64_t nrows, int64_t nt, int32_t nq, const double *xw, const double *pxy,
                                        const int32_t *t, const int64_t *r2e_ptr, const int32_t *r2e_idx,
                                        const uint8_t *slot8, const int64_t *indptr, double *values, void *stream) {
    if (nrows <= 0 || nt <= 0 || nq <= 0 || !xw || !pxy || !t || !r2e_ptr || !r2e_idx || !slot8 || !indptr || !values)
        return MWX_E_BADARG;
    k_morley_mass<<<grid_for(nrows, 128), 128, 0, as_stream(stream)>>>(nrows, nt, nq, xw, (const double2 *)pxy, t, r2e_ptr,
                                                                     r2e_idx, (const unsigned long long *)slot8, indptr,
                                                                     values);
    MWX_LAUNCH_CHECK();
    return 0;
}

extern "C" int mwx_morley_load_vector(int64_t nrows, int64_t nt, int32_t nq, const double *xw, const double *pxy,
                                      const int32_t *t, const int64_t *r2e_ptr, const int32_t *r2e_idx,
                                      const double *fq, double *out, void *stream) {
    if (nrows <= 0 || nt <= 0 || nq <= 0 || !xw || !pxy || !t || !r2e_ptr || !r2e_idx || !fq || !out) return MWX_E_BADARG;
    k_morley_load_vector<<<grid_for(nrows, 128), 128, 0, as_stream(stream)>>>(nrows, nt, nq, xw, (const double2 *)pxy, t,
                                                                            r2e_ptr, r2e_idx, fq, 0, 0.0, out);
    MWX_LAUNCH_CHECK();
    return 0;
}
