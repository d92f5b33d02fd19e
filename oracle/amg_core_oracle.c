/* Oracle (TEST INFRASTRUCTURE, never linked into the product): plain-C restatement of the
 * serial kernels of pyamg 4.0.0's amg_core that the reference reaches through
 * pyamg.ruge_stuben_solver(A) / ml.aspreconditioner()  (ref main/example1/utils.py:154-155).
 * pyamg is an un-vendored dependency (pin: ref README.md:18); the algorithm restated here is
 * its published classical Ruge-Stueben setup:
 *   strength (abs, theta) -> first-pass RS C/F splitting -> direct interpolation -> GS sweeps.
 * Pinned by the reference's golden CG-iteration table (tests/golden/amg1_iterations.json).
 *
 * Build: make -C oracle   (gcc -O2 -shared -fPIC)  ->  oracle/_build/liboracle_amg.so
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

typedef int64_t I;

/* classical_strength_of_connection_abs: keep a_ij (j != i) iff |a_ij| >= theta * max_{k != i} |a_ik|;
 * the diagonal is always kept.  Returns nnz(S).  Sj/Sx must hold nnz(A) entries. */
I oracle_strength_abs(I n, double theta, const I *Ap, const I *Aj, const double *Ax,
                      I *Sp, I *Sj, double *Sx)
{
    I nnz = 0;
    Sp[0] = 0;
    for (I i = 0; i < n; i++) {
        double mx = 2.2250738585072014e-308;
        for (I jj = Ap[i]; jj < Ap[i + 1]; jj++)
            if (Aj[jj] != i && fabs(Ax[jj]) > mx) mx = fabs(Ax[jj]);
        const double thr = theta * mx;
        for (I jj = Ap[i]; jj < Ap[i + 1]; jj++) {
            if (Aj[jj] == i || fabs(Ax[jj]) >= thr) {
                Sj[nnz] = Aj[jj];
                Sx[nnz] = Ax[jj];
                nnz++;
            }
        }
        Sp[i + 1] = nnz;
    }
    return nnz;
}

/* rs_cf_splitting, first pass only (pyamg 4.0.0 has no second pass by default).
 * S, T = S^T are CSR WITHOUT the diagonal.  split: 1 = C, 0 = F.  Bucket bookkeeping as in
 * SURVEY.md B.2 (tie-breaks depend on it). */
void oracle_rs_cf_splitting(I n, const I *Sp, const I *Sj, const I *Tp, const I *Tj, I *split)
{
    enum { F = 0, C = 1, U = 2 };
    I *lam = (I *)malloc(sizeof(I) * (size_t)n);
    I *ptr = (I *)calloc((size_t)n + 1, sizeof(I));
    I *cnt = (I *)calloc((size_t)n + 1, sizeof(I));
    I *i2n = (I *)malloc(sizeof(I) * (size_t)n);
    I *n2i = (I *)malloc(sizeof(I) * (size_t)n);

    for (I i = 0; i < n; i++) { lam[i] = Tp[i + 1] - Tp[i]; cnt[lam[i]]++; }
    for (I l = 0, c = 0; l <= n; l++) { ptr[l] = c; c += cnt[l]; cnt[l] = 0; }
    for (I i = 0; i < n; i++) {
        I l = lam[i], idx = ptr[l] + cnt[l];
        i2n[idx] = i; n2i[i] = idx; cnt[l]++;
    }
    for (I i = 0; i < n; i++) {
        split[i] = U;
        if (lam[i] == 0 || (lam[i] == 1 && Tj[Tp[i]] == i)) split[i] = F;
    }
    for (I top = n - 1; top >= 0; top--) {
        I i = i2n[top];
        cnt[lam[i]]--;
        if (split[i] == F) continue;
        split[i] = C;
        for (I jj = Tp[i]; jj < Tp[i + 1]; jj++) {
            I j = Tj[jj];
            if (split[j] != U) continue;
            split[j] = F;
            for (I kk = Sp[j]; kk < Sp[j + 1]; kk++) {
                I k = Sj[kk];
                if (split[k] != U || lam[k] >= n - 1) continue;
                I lk = lam[k], old = n2i[k], nw = ptr[lk] + cnt[lk] - 1;
                n2i[i2n[old]] = nw; n2i[i2n[nw]] = old;
                I tmp = i2n[old]; i2n[old] = i2n[nw]; i2n[nw] = tmp;
                cnt[lk]--; cnt[lk + 1]++; ptr[lk + 1] = nw;
                lam[k]++;
            }
        }
        for (I jj = Sp[i]; jj < Sp[i + 1]; jj++) {
            I j = Sj[jj];
            if (split[j] != U || lam[j] == 0) continue;
            I lj = lam[j], old = n2i[j], nw = ptr[lj];
            n2i[i2n[old]] = nw; n2i[i2n[nw]] = old;
            I tmp = i2n[old]; i2n[old] = i2n[nw]; i2n[nw] = tmp;
            cnt[lj]--; cnt[lj - 1]++; ptr[lj]++;
            ptr[lj - 1] = ptr[lj] - cnt[lj - 1];
            lam[j]--;
        }
    }
    free(lam); free(ptr); free(cnt); free(i2n); free(n2i);
}

/* rs_direct_interpolation pass 1: row pointer of P. */
void oracle_direct_interp_pass1(I n, const I *Sp, const I *Sj, const I *split, I *Bp)
{
    Bp[0] = 0;
    for (I i = 0; i < n; i++) {
        I c = 0;
        if (split[i] == 1) c = 1;
        else
            for (I jj = Sp[i]; jj < Sp[i + 1]; jj++)
                if (split[Sj[jj]] == 1 && Sj[jj] != i) c++;
        Bp[i + 1] = Bp[i] + c;
    }
}

/* pass 2: weights; S carries A's values on S's pattern (diagonal included). Returns #coarse. */
I oracle_direct_interp_pass2(I n, const I *Ap, const I *Aj, const double *Ax,
                             const I *Sp, const I *Sj, const double *Sx,
                             const I *split, const I *Bp, I *Bj, double *Bx)
{
    for (I i = 0; i < n; i++) {
        if (split[i] == 1) { Bj[Bp[i]] = i; Bx[Bp[i]] = 1.0; continue; }
        double ssp = 0, ssn = 0, sap = 0, san = 0, diag = 0;
        for (I jj = Sp[i]; jj < Sp[i + 1]; jj++)
            if (split[Sj[jj]] == 1 && Sj[jj] != i) { if (Sx[jj] < 0) ssn += Sx[jj]; else ssp += Sx[jj]; }
        for (I jj = Ap[i]; jj < Ap[i + 1]; jj++) {
            if (Aj[jj] == i) diag += Ax[jj];
            else if (Ax[jj] < 0) san += Ax[jj];
            else sap += Ax[jj];
        }
        double alpha = ssn != 0 ? san / ssn : 0.0;
        double beta = ssp != 0 ? sap / ssp : 0.0;
        if (ssp == 0) { diag += sap; beta = 0; }
        const double nc = -alpha / diag, pc = -beta / diag;
        I nnz = Bp[i];
        for (I jj = Sp[i]; jj < Sp[i + 1]; jj++)
            if (split[Sj[jj]] == 1 && Sj[jj] != i) {
                Bj[nnz] = Sj[jj];
                Bx[nnz] = (Sx[jj] < 0 ? nc : pc) * Sx[jj];
                nnz++;
            }
    }
    I *map = (I *)malloc(sizeof(I) * (size_t)n);
    I s = 0;
    for (I i = 0; i < n; i++) { map[i] = s; s += split[i]; }
    for (I k = 0; k < Bp[n]; k++) Bj[k] = map[Bj[k]];
    free(map);
    return s;
}

/* amg_core.gauss_seidel: in-place sweep over rows start, start+step, ... (stop exclusive). */
void oracle_gauss_seidel(const I *Ap, const I *Aj, const double *Ax, double *x, const double *b,
                         I start, I stop, I step)
{
    for (I i = start; i != stop; i += step) {
        double r = 0, d = 0;
        for (I jj = Ap[i]; jj < Ap[i + 1]; jj++) {
            I j = Aj[jj];
            if (j == i) d = Ax[jj];
            else r += Ax[jj] * x[j];
        }
        if (d != 0) x[i] = (b[i] - r) / d;
    }
}

/* scipy sparsetools csr_matvec (y = A x), single thread - the SpMV the reference's cg runs. */
void oracle_csr_matvec(I n, const I *Ap, const I *Aj, const double *Ax, const double *x, double *y)
{
    for (I i = 0; i < n; i++) {
        double s = 0;
        for (I jj = Ap[i]; jj < Ap[i + 1]; jj++) s += Ax[jj] * x[Aj[jj]];
        y[i] = s;
    }
}
