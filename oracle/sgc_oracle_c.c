/* sgc_oracle_c.c — C restatement of the reference's CPU algorithms.  TEST INFRASTRUCTURE ONLY.
 *
 * Purpose: (1) the CPU baseline timed by bench.py (`cpu_baseline`, `--impl reference`) on sizes
 * the numpy oracle cannot reach, (2) a second, independent restatement that tests/ cross-check
 * bit-for-bit against oracle/sgc_oracle.py.  Nothing under staggeredgridcalculus.jl_b200/ links
 * or loads this file.
 *
 * It keeps the reference's ALGORITHM AND PASS STRUCTURE, not an optimised rewrite:
 *   - apply_∂!: one sweep per call, `@threads` over the outermost index of the interior loop
 *     only (OpenMP static schedule), serial boundary planes
 *     (src/matrixfree/differential/partial.jl:33-237);
 *   - apply_curl!: six apply_∂! sweeps + one `Gv .= 0` fill in the reference's order
 *     (src/matrixfree/differential/curl.jl:66-91);
 *   - create_curl: per block the M-sized temporaries I, Jₛ, V₀, Vₛ and the three length-2M
 *     concatenations of create_∂info (src/matrix/differential/partial.jl:118-311), block
 *     concatenation into Itot/Jtot/Vtot (curl.jl:63-98), then a single-threaded
 *     COO → CSR → CSC double counting sort with duplicate summation followed by a zero-dropping
 *     pass: the contract of Julia's stdlib `dropzeros!(sparse(I,J,V,m,n))` (curl.jl:103).
 *
 * Parity unpinned at the ulp level (no Julia in this image); compiled with -ffp-contract=off.
 * Complex numbers are (re,im) pairs of doubles; complex arithmetic is written out explicitly.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { double re, im; } cx;

static inline cx cx_mul(cx a, cx b) { cx r = {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; return r; }
static inline cx cx_sub(cx a, cx b) { cx r = {a.re - b.re, a.im - b.im}; return r; }
static inline cx cx_add(cx a, cx b) { cx r = {a.re + b.re, a.im + b.im}; return r; }

/* Julia Base `/(::ComplexF64, ::ComplexF64)` (robust, scaled) */
static double rcdiv2(double a, double b, double c, double d, double r, double t) {
    if (r != 0.0) { double br = b * r; return br != 0.0 ? (a + br) * t : a * t + (b * t) * r; }
    return (a + d * (b / c)) * t;
}
static void rcdiv1(double a, double b, double c, double d, double *p, double *q) {
    double r = d / c, t = 1.0 / (c + d * r);
    *p = rcdiv2(a, b, c, d, r, t);
    *q = rcdiv2(b, -a, c, d, r, t);
}
static cx cx_div(cx z, cx w) {
    double a = z.re, b = z.im, c = w.re, d = w.im;
    const double ov = 1.7976931348623157e308, un = 2.2250738585072014e-308, eps = 2.220446049250313e-16;
    const double bs = 2.0 / (eps * eps);
    double ab = fmax(fabs(a), fabs(b)), cd = fmax(fabs(c), fabs(d)), s = 1.0, p, q;
    if (ab >= 0.5 * ov) { a *= 0.5; b *= 0.5; s *= 2.0; }
    if (cd >= 0.5 * ov) { c *= 0.5; d *= 0.5; s *= 0.5; }
    if (ab <= un * 2.0 / eps) { a *= bs; b *= bs; s /= bs; }
    if (cd <= un * 2.0 / eps) { c *= bs; d *= bs; s *= bs; }
    if (fabs(d) <= fabs(c)) rcdiv1(a, b, c, d, &p, &q);
    else { rcdiv1(b, a, d, c, &p, &q); q = -q; }
    cx r = {p * s, q * s};
    return r;
}
static cx cx_inv(cx w) { /* Julia inv(::ComplexF64) */
    double c = w.re, d = w.im, absc = fabs(c), absd = fabs(d), cd = absc > absd ? absc : absd;
    const double eps = 2.220446049250313e-16, bs = 2.0 / (eps * eps);
    double s = 1.0, p, q;
    if (cd >= 1.7976931348623157e308 / 2) { c *= 0.5; d *= 0.5; s = 0.5; }
    else if (cd <= 2 * 2.2250738585072014e-308 / eps) { c *= bs; d *= bs; s = bs; }
    if (absd <= absc) { double r = d / c; p = 1.0 / (d * r + c); q = -r * p; }
    else { double cc = -d, dd = -c, r = dd / cc, pp = 1.0 / (dd * r + cc), qq = -r * pp; q = pp; p = qq; }
    cx r2 = {p * s, q * s};
    return r2;
}

static inline void set_or_add(cx *g, cx v, int add) { if (add) *g = cx_add(*g, v); else *g = v; }

/* apply_∂! for ComplexF64 fields, complex ∆w⁻¹, complex e and α; 3-D (partial.jl:33-237).
 * nw is 1-based. */
void sgc_ref_apply_partial_c128(cx *Gv, const cx *Fu, int64_t Nx, int64_t Ny, int64_t Nz, int nw, int isfwd,
                                const cx *dwinv, int isbloch, cx e, cx alpha, int add) {
    const int64_t sx = 1, sy = Nx, sz = Nx * Ny;
    const int64_t N[3] = {Nx, Ny, Nz}, st[3] = {sx, sy, sz};
    const int64_t Nw = N[nw - 1], sw = st[nw - 1];
    /* interior range along w */
    int64_t w0, w1; /* 0-based half-open */
    if (isfwd) { w0 = isbloch ? 0 : 1; w1 = Nw - 1; }
    else { w0 = 1; w1 = Nw; }
    const int64_t k0 = nw == 3 ? w0 : 0, k1 = nw == 3 ? w1 : Nz;
    const int64_t j0 = nw == 2 ? w0 : 0, j1 = nw == 2 ? w1 : Ny;
    const int64_t i0 = nw == 1 ? w0 : 0, i1 = nw == 1 ? w1 : Nx;
    const int64_t nbo = isfwd ? sw : -sw;
#pragma omp parallel for schedule(static)
    for (int64_t k = k0; k < k1; ++k)
        for (int64_t j = j0; j < j1; ++j)
            for (int64_t i = i0; i < i1; ++i) {
                const int64_t o = i + sy * j + sz * k;
                const int64_t iw = nw == 1 ? i : (nw == 2 ? j : k);
                const cx c = cx_mul(alpha, dwinv[iw]); /* (α * ∆w⁻¹[i]) */
                const cx d = isfwd ? cx_sub(Fu[o + nbo], Fu[o]) : cx_sub(Fu[o], Fu[o + nbo]);
                set_or_add(&Gv[o], cx_mul(c, d), add);
            }
    /* boundary planes: serial (partial.jl:62-65, 77-87, 179-188) */
    const int64_t a1 = nw == 1 ? Ny : Nx, a2 = nw == 3 ? Ny : Nz;
    const int64_t s1 = nw == 1 ? sy : sx, s2 = nw == 3 ? sy : sz;
    for (int64_t q2 = 0; q2 < a2; ++q2)
        for (int64_t q1 = 0; q1 < a1; ++q1) {
            const int64_t b = q1 * s1 + q2 * s2;
            if (isfwd) {
                if (isbloch) {
                    const cx beta = cx_mul(alpha, dwinv[Nw - 1]);
                    const int64_t o = b + (Nw - 1) * sw;
                    set_or_add(&Gv[o], cx_mul(beta, cx_sub(cx_mul(e, Fu[b]), Fu[o])), add);
                } else {
                    cx beta = cx_mul(alpha, dwinv[0]);
                    set_or_add(&Gv[b], cx_mul(beta, Fu[b + sw]), add);
                }
            } else {
                if (isbloch) {
                    const cx beta = cx_mul(alpha, dwinv[0]);
                    set_or_add(&Gv[b], cx_mul(beta, cx_sub(Fu[b], cx_div(Fu[b + (Nw - 1) * sw], e))), add);
                } else {
                    cx zero = {0.0, 0.0};
                    set_or_add(&Gv[b], zero, add);
                }
            }
        }
    if (isfwd && !isbloch) {
        cx na = {-alpha.re, -alpha.im};
        const cx beta = cx_mul(na, dwinv[Nw - 1]);
        for (int64_t q2 = 0; q2 < a2; ++q2)
            for (int64_t q1 = 0; q1 < a1; ++q1) {
                const int64_t o = q1 * s1 + q2 * s2 + (Nw - 1) * sw;
                set_or_add(&Gv[o], cx_mul(beta, Fu[o]), add);
            }
    }
}

static const int CURL_BLK[3][3] = {{0, -1, 1}, {1, 0, -1}, {-1, 1, 0}};

/* apply_curl! (curl.jl:52-137), full 3-D, ComplexF64, complex ∆l⁻¹ vectors. */
void sgc_ref_apply_curl_c128(cx *G, const cx *F, int64_t Nx, int64_t Ny, int64_t Nz, const int *isfwd,
                             const cx *dxinv, const cx *dyinv, const cx *dzinv, const int *isbloch, const cx *e,
                             cx alpha, int add) {
    const int64_t M = Nx * Ny * Nz;
    const cx *dl[3] = {dxinv, dyinv, dzinv};
    for (int nv = 1; nv <= 3; ++nv) {
        cx *Gv = G + (nv - 1) * M;
        int op_add = add;
        for (int nu = 1; nu <= 3; ++nu) {
            const int parity = CURL_BLK[nv - 1][nu - 1];
            if (parity == 0) {
                if (!op_add) {
#pragma omp parallel for schedule(static)
                    for (int64_t q = 0; q < M; ++q) { Gv[q].re = 0.0; Gv[q].im = 0.0; } /* Gv .= 0 */
                    op_add = 1;
                }
                continue;
            }
            const int nw = 6 - nv - nu;
            cx a = {parity * alpha.re, parity * alpha.im}; /* α = parity*α */
            sgc_ref_apply_partial_c128(Gv, F + (nu - 1) * M, Nx, Ny, Nz, nw, isfwd[nw - 1], dl[nw - 1],
                                       isbloch[nw - 1], e[nw - 1], a, op_add);
            op_add = 1;
        }
    }
}

/* ---------------------------------------------------------------------------------------------
 * create_curl restated: temporaries of create_∂info, block concatenation, sparse + dropzeros!.
 * Float64 values (real ∆l⁻¹, real e).  order_cmpfirst as in curl.jl:74,81.
 * Outputs are malloc'ed; returns nnz, or -1 on allocation failure.
 * ------------------------------------------------------------------------------------------- */
static void create_partial_info_f64(int nw, int isfwd, const int64_t N[3], const double *dwinv, int isbloch, double e,
                                    int64_t *Is, int64_t *Js, double *Vs_out) {
    const int64_t M = N[0] * N[1] * N[2], Nw = N[nw - 1];
    const double ns = isfwd ? 1.0 : -1.0;
    int64_t *I = (int64_t *)malloc(sizeof(int64_t) * M);   /* I  = reshape(collect(1:M), N) */
    int64_t *Jc = (int64_t *)malloc(sizeof(int64_t) * M);  /* Jₛ = collect(1:M) */
    int64_t *Jsft = (int64_t *)malloc(sizeof(int64_t) * M); /* circshift(Jₛ, -ns ŵ) */
    double *V0 = (double *)malloc(sizeof(double) * M), *Vs = (double *)malloc(sizeof(double) * M);
    for (int64_t q = 0; q < M; ++q) { I[q] = q + 1; Jc[q] = q + 1; }
    const int64_t st[3] = {1, N[0], N[0] * N[1]};
    for (int64_t k = 0; k < N[2]; ++k)
        for (int64_t j = 0; j < N[1]; ++j)
            for (int64_t i = 0; i < N[0]; ++i) {
                const int64_t sub[3] = {i, j, k};
                const int64_t q = i + st[1] * j + st[2] * k;
                int64_t w = sub[nw - 1] + (isfwd ? 1 : -1);
                if (w < 0) w += Nw;
                if (w >= Nw) w -= Nw;
                Jsft[q] = Jc[q + (w - sub[nw - 1]) * st[nw - 1]];
                V0[q] = (-ns * 1.0) * dwinv[sub[nw - 1]];
                Vs[q] = (ns * 1.0) * dwinv[sub[nw - 1]];
            }
    /* boundary conditions on the slabs (partial.jl:216-217, 290-299) */
    const int64_t iw = isfwd ? Nw - 1 : 0;
    const double en = isfwd ? e : 1.0 / e; /* e^ns */
    for (int64_t k = 0; k < N[2]; ++k)
        for (int64_t j = 0; j < N[1]; ++j)
            for (int64_t i = 0; i < N[0]; ++i) {
                const int64_t sub[3] = {i, j, k};
                const int64_t q = i + st[1] * j + st[2] * k;
                if (isbloch) { if (sub[nw - 1] == iw) Vs[q] *= en; }
                else {
                    if (sub[nw - 1] == 0) V0[q] = 0.0;
                    if (sub[nw - 1] == iw) Vs[q] = 0.0;
                }
            }
    memcpy(Is, I, sizeof(int64_t) * M); memcpy(Is + M, I, sizeof(int64_t) * M);      /* [i; i]  */
    memcpy(Js, I, sizeof(int64_t) * M); memcpy(Js + M, Jsft, sizeof(int64_t) * M);   /* [i; jₛ] */
    memcpy(Vs_out, V0, sizeof(double) * M); memcpy(Vs_out + M, Vs, sizeof(double) * M);
    free(I); free(Jc); free(Jsft); free(V0); free(Vs);
}

/* dropzeros!(sparse(I,J,V,m,n)): COO → CSR (counting sort on rows) → CSC (counting sort on columns,
 * summing duplicates, rows come out ascending), then drop stored zeros.  Single thread. */
static int64_t sparse_dropzeros_f64(int64_t nz, const int64_t *I, const int64_t *J, const double *V, int64_t m, int64_t n,
                                    int64_t **colptr_o, int64_t **rowval_o, double **nzval_o) {
    int64_t *rp = (int64_t *)calloc(m + 1, sizeof(int64_t));
    int64_t *cj = (int64_t *)malloc(sizeof(int64_t) * (nz ? nz : 1));
    double *cv = (double *)malloc(sizeof(double) * (nz ? nz : 1));
    if (!rp || !cj || !cv) return -1;
    for (int64_t q = 0; q < nz; ++q) rp[I[q]]++;
    for (int64_t r = 0; r < m; ++r) rp[r + 1] += rp[r];
    {
        int64_t *pos = (int64_t *)malloc(sizeof(int64_t) * (m + 1));
        memcpy(pos, rp, sizeof(int64_t) * (m + 1));
        for (int64_t q = 0; q < nz; ++q) { const int64_t p = pos[I[q] - 1]++; cj[p] = J[q]; cv[p] = V[q]; }
        free(pos);
    }
    int64_t *colptr = (int64_t *)calloc(n + 2, sizeof(int64_t));
    int64_t *rowval = (int64_t *)malloc(sizeof(int64_t) * (nz ? nz : 1));
    double *nzval = (double *)malloc(sizeof(double) * (nz ? nz : 1));
    if (!colptr || !rowval || !nzval) return -1;
    for (int64_t q = 0; q < nz; ++q) colptr[cj[q] + 1]++;
    colptr[0] = 0;
    for (int64_t c = 0; c <= n; ++c) colptr[c + 1] += colptr[c];
    /* colptr[c] (1-based c) = start of column c (0-based offsets) */
    int64_t *pos = (int64_t *)malloc(sizeof(int64_t) * (n + 2));
    memcpy(pos, colptr, sizeof(int64_t) * (n + 2));
    int64_t *last = (int64_t *)malloc(sizeof(int64_t) * (n + 1)); /* last row stored in each column */
    for (int64_t c = 0; c <= n; ++c) last[c] = -1;
    for (int64_t r = 0; r < m; ++r)
        for (int64_t p = rp[r]; p < rp[r + 1]; ++p) {
            const int64_t c = cj[p];
            if (last[c] == r) nzval[pos[c] - 1] += cv[p]; /* duplicate (r,c): sum in input order */
            else { rowval[pos[c]] = r + 1; nzval[pos[c]] = cv[p]; pos[c]++; last[c] = r; }
        }
    /* compact (duplicates left holes at column ends) + dropzeros! */
    int64_t w = 0;
    int64_t *cp = (int64_t *)malloc(sizeof(int64_t) * (n + 1));
    for (int64_t c = 1; c <= n; ++c) {
        cp[c - 1] = w + 1;
        for (int64_t p = colptr[c]; p < pos[c]; ++p)
            if (nzval[p] != 0.0) { rowval[w] = rowval[p]; nzval[w] = nzval[p]; ++w; }
    }
    cp[n] = w + 1;
    free(rp); free(cj); free(cv); free(colptr); free(pos); free(last);
    *colptr_o = cp; *rowval_o = rowval; *nzval_o = nzval;
    return w;
}

int64_t sgc_ref_create_curl_f64(int64_t Nx, int64_t Ny, int64_t Nz, const int *isfwd, const double *dxinv,
                                const double *dyinv, const double *dzinv, const int *isbloch, const double *e,
                                int order_cmpfirst, int64_t **colptr, int64_t **rowval, double **nzval) {
    const int64_t N[3] = {Nx, Ny, Nz}, M = Nx * Ny * Nz;
    const double *dl[3] = {dxinv, dyinv, dzinv};
    const int num_blk = 6;
    int64_t *Itot = (int64_t *)malloc(sizeof(int64_t) * num_blk * 2 * M);
    int64_t *Jtot = (int64_t *)malloc(sizeof(int64_t) * num_blk * 2 * M);
    double *Vtot = (double *)malloc(sizeof(double) * num_blk * 2 * M);
    int64_t *I = (int64_t *)malloc(sizeof(int64_t) * 2 * M), *J = (int64_t *)malloc(sizeof(int64_t) * 2 * M);
    double *V = (double *)malloc(sizeof(double) * 2 * M);
    if (!Itot || !Jtot || !Vtot || !I || !J || !V) return -1;
    int ind_blk = 0;
    for (int nv = 1; nv <= 3; ++nv) {
        const int64_t istr = order_cmpfirst ? 3 : 1, ioff = order_cmpfirst ? nv - 3 : M * (nv - 1);
        for (int nu = 1; nu <= 3; ++nu) {
            const int parity = CURL_BLK[nv - 1][nu - 1];
            if (!parity) continue;
            const int64_t jstr = order_cmpfirst ? 3 : 1, joff = order_cmpfirst ? nu - 3 : M * (nu - 1);
            const int nw = 6 - nv - nu;
            create_partial_info_f64(nw, isfwd[nw - 1], N, dl[nw - 1], isbloch[nw - 1], e[nw - 1], I, J, V);
            for (int64_t q = 0; q < 2 * M; ++q) { I[q] = istr * I[q] + ioff; J[q] = jstr * J[q] + joff; V[q] *= parity; }
            memcpy(Itot + (int64_t)ind_blk * 2 * M, I, sizeof(int64_t) * 2 * M);
            memcpy(Jtot + (int64_t)ind_blk * 2 * M, J, sizeof(int64_t) * 2 * M);
            memcpy(Vtot + (int64_t)ind_blk * 2 * M, V, sizeof(double) * 2 * M);
            ++ind_blk;
        }
    }
    free(I); free(J); free(V);
    const int64_t nnz = sparse_dropzeros_f64((int64_t)num_blk * 2 * M, Itot, Jtot, Vtot, 3 * M, 3 * M, colptr, rowval, nzval);
    free(Itot); free(Jtot); free(Vtot);
    return nnz;
}

void sgc_ref_free(void *p) { free(p); }
int sgc_ref_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void sgc_ref_set_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}
