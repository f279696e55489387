// sgc_curlcurl.cuh — fused curl-of-curl  G OP C₂(C₁ F)  in ONE pass over HBM (sm_100a).
//
// The step either side of the curl in a Maxwell solve (reference usage
// test/curl.jl:123-153: `Cu = create_curl(isfwd…)`, `Cv = create_curl(.!isfwd…)`, `Cv*Cu`) is two
// apply_curl! calls (matrixfree/differential/curl.jl:52-137) with the intermediate field H
// written to and re-read from memory: 2 × 96 = 192 B per ComplexF64 cell.  This kernel keeps H in
// shared memory (temporal blocking along the z-march), so a cell costs 96 B: read E, write E′.
//
//   CTA = TX×TY output cells, marching `march` planes up in z.
//   per z-step m:  * plane (TX+2)×(TY+2) of Ex, Ey, Ez arrives in an SE-stage cp.async ring
//                  * B(m): H plane m on the (TX+1)×(TY+1) tile the outputs need (C₁ applied to the
//                    staged E planes; the extra row/column is recomputed, not communicated)
//                    → 3-stage H ring in shared memory
//                  * one __syncthreads
//                  * D(m−1): E′ plane from the H planes (C₂), written to HBM.
//
// Every value is produced by exactly the arithmetic of the two-pass path (same `dterm` formulas,
// same accumulation order, H rounded to double exactly as when it is stored), so the result is
// bit-identical to apply_curl! twice; the tests require that.
//
// Directions and boundary conditions of C₁ and C₂ are independent per axis (any combination).
// Tile positions are mapped to grid indices modulo N, so Bloch wrap-around needs no special
// loader; positions whose value a boundary rule discards are computed from valid memory and
// dropped by a select.  z-slabs (multi-GPU): the planes −1 and Nz of E are read through ghost
// pointers (peer memory) at slab interfaces and at physical Bloch faces whose wrap partner is
// another rank; this needs opposite z-directions in C₁ and C₂ (reach ±1).
#pragma once
#include "sgc_kernels.cuh"

namespace sgc {

template <typename T, typename C>
struct CurlCurlParams {
    const T *F[3];  // E
    T *G[3];        // E′
    const C *c1[3], *c2[3];  // α·∆l⁻¹ tables of C₁ and C₂
    cplx e1[3], e2[3];
    int e1_real[3], e2_real[3];
    int fwd1[3], fwd2[3];
    int bc1_lo[3], bc1_hi[3], bc2_lo[3], bc2_hi[3];
    int Nx, Ny, Nz;
    int k_begin, k_end, march;
    const T *ghost_lo[3], *ghost_hi[3];  // E planes −1 / Nz (slab interfaces); unused otherwise
    C c1z_lo, c1z_hi;                    // c1[2] of the H planes −1 / Nz (the neighbour's entries)
};

// one ∂ term with the boundary rules folded in (the general path of k_curl3p):
//   forward  c·(nb′ − cur′): Bloch edge nb′ = e·nb, symmetric edge nb′ = 0, symmetric i = 0: cur′ = 0
//   backward c·(cur − nb′):  Bloch edge nb′ = nb/e, symmetric edge: the whole term is 0
// kind: 0 plain, 1 Bloch edge, 2 symmetric edge
template <typename T, typename C>
SGC_D T dterm(bool fwd, int kind, bool lo_zero, C c, T cur, T nb, cplx e, int e_real) {
    const T Z = zero_of(T{});
    T hi, lo;
    bool zero_term = false;
    if (fwd) {
        hi = nb; lo = cur;
        if (kind == 1) hi = mul_e(e, hi, e_real);
        else if (kind == 2) hi = Z;
        if (lo_zero) lo = Z;
    } else {
        hi = cur; lo = nb;
        if (kind == 1) lo = div_e(lo, e, e_real);
        else if (kind == 2) zero_term = true;
    }
    return zero_term ? Z : mul(c, sub(hi, lo));
}

// boundary description of one operator at one (i, j)
template <typename C>
struct XYEdge {
    int xk, yk;
    bool xl0, yl0;
    C cxs, cys;  // coefficients with the direction folded into the sign: c (forward) or −c (backward);
                 // the general path recovers c by an exact negation
};

template <typename C>
SGC_D XYEdge<C> xy_edge(const int *fwd, const int *bc_lo, const int *bc_hi, const C *const *c, int gi, int gj, int Nx, int Ny) {
    XYEdge<C> b;
    b.xk = 0; b.yk = 0;
    if (fwd[0]) { if (gi == Nx - 1) b.xk = (bc_hi[0] == BC_BLOCH) ? 1 : 2; }
    else        { if (gi == 0) b.xk = (bc_lo[0] == BC_BLOCH) ? 1 : 2; }
    if (fwd[1]) { if (gj == Ny - 1) b.yk = (bc_hi[1] == BC_BLOCH) ? 1 : 2; }
    else        { if (gj == 0) b.yk = (bc_lo[1] == BC_BLOCH) ? 1 : 2; }
    b.xl0 = fwd[0] && (gi == 0) && (bc_lo[0] == BC_SYM);
    b.yl0 = fwd[1] && (gj == 0) && (bc_lo[1] == BC_SYM);
    b.cxs = fwd[0] ? c[0][gi] : neg(c[0][gi]);
    b.cys = fwd[1] ? c[1][gj] : neg(c[1][gj]);
    return b;
}

SGC_D int wrap_index(int v, int N) {
    v %= N;
    return v < 0 ? v + N : v;
}

template <typename T, typename C, bool ADD, int TX, int TY, int SE, int MINB>
__global__ void __launch_bounds__(TX *TY, MINB) k_curlcurl3(const __grid_constant__ CurlCurlParams<T, C> p) {
    static_assert(SE >= 3, "need at least 3 stages");
    constexpr int NT = TX * TY;
    constexpr int HX = TX + 1, HY = TY + 1, EX = TX + 2, EY = TY + 2;
    constexpr int HCOMP = HX * HY, ECOMP = EX * EY;
    constexpr int ESTAGE = 3 * ECOMP, HSTAGE = 3 * HCOMP;
    constexpr int NPASS = (HCOMP + NT - 1) / NT;
    constexpr int NLD = (ECOMP + NT - 1) / NT;
    extern __shared__ __align__(16) unsigned char sgc_smem_raw[];
    T *const sE = reinterpret_cast<T *>(sgc_smem_raw);
    T *const sH = sE + SE * ESTAGE;

    const int tx = threadIdx.x, ty = threadIdx.y;
    const int tid = ty * TX + tx;
    const int Nx = p.Nx, Ny = p.Ny, Nz = p.Nz;
    const int64_t sy = Nx, sz = (int64_t)Nx * Ny;
    const int fx1 = p.fwd1[0], fy1 = p.fwd1[1], fz1 = p.fwd1[2];
    const int fx2 = p.fwd2[0], fy2 = p.fwd2[1], fz2 = p.fwd2[2];
    const int s1x = fx1 ? 1 : -1, s1y = fy1 ? 1 : -1, s1z = fz1 ? 1 : -1;
    const int s2x = fx2 ? 1 : -1, s2y = fy2 ? 1 : -1, s2z = fz2 ? 1 : -1;
    const int i0 = blockIdx.x * TX, j0 = blockIdx.y * TY;
    const int hx0 = i0 + min(0, s2x), hy0 = j0 + min(0, s2y);  // grid origin of the H tile
    const int ex0 = hx0 + min(0, s1x), ey0 = hy0 + min(0, s1y);  // grid origin of the E tile
    const T Z = zero_of(T{});
    // no H position (hence no output cell) of this tile lies on an x/y boundary
    const bool tile_plain = (hx0 >= 1) && (hx0 + HX - 1 <= Nx - 2) && (hy0 >= 1) && (hy0 + HY - 1 <= Ny - 2);
    // planes beyond the slab come through ghost pointers: slab interfaces, and physical Bloch faces
    // whose wrap-around partner lives on another rank
    const bool lo_ext = (p.ghost_lo[0] != nullptr), hi_ext = (p.ghost_hi[0] != nullptr);

    // ---- loader duty: E-tile elements tid, tid + NT, … -------------------------------------------
    int64_t ld_g[NLD];
    bool ld_on[NLD];
#pragma unroll
    for (int u = 0; u < NLD; ++u) {
        const int q = tid + u * NT;
        ld_on[u] = q < ECOMP;
        const int py = q / EX, px = q - py * EX;
        ld_g[u] = (int64_t)wrap_index(ey0 + py, Ny) * sy + wrap_index(ex0 + px, Nx);
    }
    // ---- H positions of this thread -----------------------------------------------------------------
    bool h_on[NPASS];
    int es[NPASS];
    XYEdge<C> b1[NPASS];
#pragma unroll
    for (int u = 0; u < NPASS; ++u) {
        const int pos = tid + u * NT;
        h_on[u] = pos < HCOMP;
        const int py = pos / HX, px = pos - py * HX;
        es[u] = (py + (s1y < 0 ? 1 : 0)) * EX + (px + (s1x < 0 ? 1 : 0));
        b1[u] = xy_edge<C>(p.fwd1, p.bc1_lo, p.bc1_hi, p.c1, wrap_index(hx0 + px, Nx), wrap_index(hy0 + py, Ny), Nx, Ny);
    }
    const int e_dx = s1x, e_dy = s1y * EX;
    // ---- output cell ----------------------------------------------------------------------------------
    const int i = i0 + tx, j = j0 + ty;
    const bool valid = (i < Nx) && (j < Ny);
    const int hs = (ty + (s2y < 0 ? 1 : 0)) * HX + (tx + (s2x < 0 ? 1 : 0));
    const int h_dx = s2x, h_dy = s2y * HX;
    const XYEdge<C> b2 = xy_edge<C>(p.fwd2, p.bc2_lo, p.bc2_hi, p.c2, valid ? i : 0, valid ? j : 0, Nx, Ny);
    const int64_t ij = (int64_t)j * sy + i;

    const int k0 = p.k_begin + blockIdx.z * p.march;
    const int k1 = min(k0 + p.march, p.k_end);
    if (k0 >= k1) return;
    const int n = k1 - k0;              // output planes; H slots 0..n; E slots 0..n+1
    const int hb = k0 + min(0, s2z);    // grid plane of H slot 0
    const int eb = hb + min(0, s1z);    // grid plane of E slot 0

    // ---- producer: E slot jj → ring stage jj % SE ------------------------------------------------------
    int jj_issue = 0, stage_issue = 0;
    auto issue = [&]() {
        if (jj_issue <= n + 1) {
            const int ke = eb + jj_issue;
            const T *s0, *s1, *s2;
            if (ke < 0 && lo_ext) { s0 = p.ghost_lo[0]; s1 = p.ghost_lo[1]; s2 = p.ghost_lo[2]; }
            else if (ke >= Nz && hi_ext) { s0 = p.ghost_hi[0]; s1 = p.ghost_hi[1]; s2 = p.ghost_hi[2]; }
            else {
                const int64_t off = (int64_t)wrap_index(ke, Nz) * sz;
                s0 = p.F[0] + off; s1 = p.F[1] + off; s2 = p.F[2] + off;
            }
            const bool centre = fz1 ? (jj_issue <= n) : (jj_issue >= 1);  // Ez only where an H plane is centred
            T *st = sE + stage_issue * ESTAGE;
#pragma unroll
            for (int u = 0; u < NLD; ++u) {
                if (ld_on[u]) {
                    const int q = tid + u * NT;
                    cp_async_elem(st + q, s0 + ld_g[u]);
                    cp_async_elem(st + ECOMP + q, s1 + ld_g[u]);
                    if (centre) cp_async_elem(st + 2 * ECOMP + q, s2 + ld_g[u]);
                }
            }
        }
        cp_async_commit();
        ++jj_issue;
        stage_issue = (stage_issue + 1 == SE) ? 0 : stage_issue + 1;
    };
#pragma unroll
    for (int q = 0; q < SE; ++q) issue();
    cp_async_wait<SE - 2>();  // slots 0 and 1 have landed
    __syncthreads();

    T exlo[NPASS], eylo[NPASS];  // E "lo" plane values of my H positions, carried between steps
    T hxlo = Z, hylo = Z;        // H "lo" plane values of my output cell
    int est_lo = 0;              // ring stage of E slot m
    int hst = 0;                 // H stage written in step m  (m % 3)
    int64_t off_out = (int64_t)k0 * sz + ij;

    for (int m = 0; m <= n; ++m) {
        // ================= B(m): H plane hb + m from E slots m (lo) and m + 1 (hi) ====================
        {
            const int hk = hb + m;
            // an H plane of the neighbour rank: plane −1 (reached by a backward C₂, so C₁ is forward and
            // the plane is interior — or, at a physical Bloch face, the global LAST plane, whose forward
            // difference multiplies its upper neighbour by e) or plane Nz (C₁ backward: interior, or the
            // global plane 0, whose backward difference divides its lower neighbour by e)
            const bool ghostH = (hk < 0 && lo_ext) || (hk >= Nz && hi_ext);
            const int kk = ghostH ? hk : wrap_index(hk, Nz);
            const C cz = ghostH ? (hk < 0 ? p.c1z_lo : p.c1z_hi) : p.c1[2][kk];
            int zk = 0;
            bool zl0 = false;
            bool ghost_plain = false;
            if (ghostH) {
                const int side = (hk < 0) ? p.bc1_lo[2] : p.bc1_hi[2];
                if (side == BC_BLOCH) zk = 1;
                else ghost_plain = true;
            } else {
                if (fz1) {
                    if (kk == Nz - 1) zk = (p.bc1_hi[2] == BC_BLOCH) ? 1 : (p.bc1_hi[2] == BC_SYM ? 2 : 0);
                    zl0 = (kk == 0) && (p.bc1_lo[2] == BC_SYM);
                } else if (kk == 0) {
                    zk = (p.bc1_lo[2] == BC_BLOCH) ? 1 : (p.bc1_lo[2] == BC_SYM ? 2 : 0);
                }
            }
            const bool fast = tile_plain && (ghostH ? ghost_plain : (kk > 0 && kk < Nz - 1));
            const int est_hi = (est_lo + 1 == SE) ? 0 : est_lo + 1;
            const T *slo = sE + est_lo * ESTAGE;
            const T *shi = sE + est_hi * ESTAGE;
            const T *sc = fz1 ? slo : shi;
            T *hout = sH + hst * HSTAGE;
#pragma unroll
            for (int u = 0; u < NPASS; ++u) {
                if (!h_on[u]) continue;
                const int q = es[u];
                if (m == 0) { exlo[u] = slo[q]; eylo[u] = slo[ECOMP + q]; }
                const T xlo = exlo[u], ylo = eylo[u];
                const T xhi = shi[q], yhi = shi[ECOMP + q];
                const T xc = fz1 ? xlo : xhi, yc = fz1 ? ylo : yhi;
                const T zc = sc[2 * ECOMP + q];
                const T x_yn = sc[q + e_dy], z_yn = sc[2 * ECOMP + q + e_dy];
                const T y_xn = sc[ECOMP + q + e_dx], z_xn = sc[2 * ECOMP + q + e_dx];
                T hx, hy, hz;
                if (fast) {
                    const T dzx = mul(cz, sub(xhi, xlo)), dzy = mul(cz, sub(yhi, ylo));
                    const T dyx = mul(b1[u].cys, sub(x_yn, xc)), dyz = mul(b1[u].cys, sub(z_yn, zc));
                    const T dxy = mul(b1[u].cxs, sub(y_xn, yc)), dxz = mul(b1[u].cxs, sub(z_xn, zc));
                    hx = add(add(Z, neg(dzy)), dyz);
                    hy = add(dzx, neg(dxz));
                    hz = add(neg(dyx), dxy);
                } else {
                    const T dzx = dterm<T, C>(fz1, zk, zl0, cz, xc, fz1 ? xhi : xlo, p.e1[2], p.e1_real[2]);
                    const T dzy = dterm<T, C>(fz1, zk, zl0, cz, yc, fz1 ? yhi : ylo, p.e1[2], p.e1_real[2]);
                    const C cy = fy1 ? b1[u].cys : neg(b1[u].cys), cx = fx1 ? b1[u].cxs : neg(b1[u].cxs);
                    const T dyx = dterm<T, C>(fy1, b1[u].yk, b1[u].yl0, cy, xc, x_yn, p.e1[1], p.e1_real[1]);
                    const T dyz = dterm<T, C>(fy1, b1[u].yk, b1[u].yl0, cy, zc, z_yn, p.e1[1], p.e1_real[1]);
                    const T dxy = dterm<T, C>(fx1, b1[u].xk, b1[u].xl0, cx, yc, y_xn, p.e1[0], p.e1_real[0]);
                    const T dxz = dterm<T, C>(fx1, b1[u].xk, b1[u].xl0, cx, zc, z_xn, p.e1[0], p.e1_real[0]);
                    hx = add(add(Z, neg(dzy)), dyz);  // Gv .= 0 ; += (−∂zFy) ; += ∂yFz   (curl.jl:66-91)
                    hy = add(dzx, neg(dxz));
                    hz = add(neg(dyx), dxy);
                }
                const int pos = tid + u * NT;
                hout[pos] = hx;
                hout[HCOMP + pos] = hy;
                hout[2 * HCOMP + pos] = hz;
                exlo[u] = xhi;
                eylo[u] = yhi;
            }
            est_lo = est_hi;
        }
        // the read-modify-write loads of D(m−1) go out before the barrier
        T g0 = Z, g1 = Z, g2 = Z;
        if (ADD && valid && m >= 1) {
            g0 = ldg_plain(p.G[0] + off_out); g1 = ldg_plain(p.G[1] + off_out); g2 = ldg_plain(p.G[2] + off_out);
        }
        cp_async_wait<SE - 3>();  // E slot m + 2 has landed (for this thread) …
        __syncthreads();          // … and for everybody; H plane m is visible; E slot m is free
        issue();                  // E slot m + SE → the stage of slot m

        // ================= D(m − 1): output plane k0 + m − 1 from H slots m − 1 (lo) and m (hi) ========
        if (m >= 1) {
            const int k = k0 + m - 1;
            const int hst_lo = (hst == 0) ? 2 : hst - 1;
            const T *hlo = sH + hst_lo * HSTAGE;
            const T *hhi = sH + hst * HSTAGE;
            const T *hc = fz2 ? hlo : hhi;
            if (m == 1) { hxlo = hlo[hs]; hylo = hlo[HCOMP + hs]; }
            const T xlo = hxlo, ylo = hylo;
            const T xhi = hhi[hs], yhi = hhi[HCOMP + hs];
            const T xc = fz2 ? xlo : xhi, yc = fz2 ? ylo : yhi;
            const T zc = hc[2 * HCOMP + hs];
            const T x_yn = hc[hs + h_dy], z_yn = hc[2 * HCOMP + hs + h_dy];
            const T y_xn = hc[HCOMP + hs + h_dx], z_xn = hc[2 * HCOMP + hs + h_dx];
            const C cz = p.c2[2][k];
            T gx, gy, gz;
            if (tile_plain && k > 0 && k < Nz - 1) {
                const T dzx = mul(cz, sub(xhi, xlo)), dzy = mul(cz, sub(yhi, ylo));
                const T dyx = mul(b2.cys, sub(x_yn, xc)), dyz = mul(b2.cys, sub(z_yn, zc));
                const T dxy = mul(b2.cxs, sub(y_xn, yc)), dxz = mul(b2.cxs, sub(z_xn, zc));
                if (ADD) { gx = add(add(g0, neg(dzy)), dyz); gy = add(add(g1, dzx), neg(dxz)); gz = add(add(g2, neg(dyx)), dxy); }
                else     { gx = add(add(Z, neg(dzy)), dyz);  gy = add(dzx, neg(dxz));          gz = add(neg(dyx), dxy); }
            } else {
                int zk = 0;
                bool zl0 = false;
                if (fz2) {
                    if (k == Nz - 1) zk = (p.bc2_hi[2] == BC_BLOCH) ? 1 : (p.bc2_hi[2] == BC_SYM ? 2 : 0);
                    zl0 = (k == 0) && (p.bc2_lo[2] == BC_SYM);
                } else if (k == 0) {
                    zk = (p.bc2_lo[2] == BC_BLOCH) ? 1 : (p.bc2_lo[2] == BC_SYM ? 2 : 0);
                }
                const T dzx = dterm<T, C>(fz2, zk, zl0, cz, xc, fz2 ? xhi : xlo, p.e2[2], p.e2_real[2]);
                const T dzy = dterm<T, C>(fz2, zk, zl0, cz, yc, fz2 ? yhi : ylo, p.e2[2], p.e2_real[2]);
                const C cy = fy2 ? b2.cys : neg(b2.cys), cx = fx2 ? b2.cxs : neg(b2.cxs);
                const T dyx = dterm<T, C>(fy2, b2.yk, b2.yl0, cy, xc, x_yn, p.e2[1], p.e2_real[1]);
                const T dyz = dterm<T, C>(fy2, b2.yk, b2.yl0, cy, zc, z_yn, p.e2[1], p.e2_real[1]);
                const T dxy = dterm<T, C>(fx2, b2.xk, b2.xl0, cx, yc, y_xn, p.e2[0], p.e2_real[0]);
                const T dxz = dterm<T, C>(fx2, b2.xk, b2.xl0, cx, zc, z_xn, p.e2[0], p.e2_real[0]);
                if (ADD) { gx = add(add(g0, neg(dzy)), dyz); gy = add(add(g1, dzx), neg(dxz)); gz = add(add(g2, neg(dyx)), dxy); }
                else     { gx = add(add(Z, neg(dzy)), dyz);  gy = add(dzx, neg(dxz));          gz = add(neg(dyx), dxy); }
            }
            if (valid) {
                stg_stream(p.G[0] + off_out, gx);
                stg_stream(p.G[1] + off_out, gy);
                stg_stream(p.G[2] + off_out, gz);
            }
            hxlo = xhi;
            hylo = yhi;
            off_out += sz;
        }
        hst = (hst == 2) ? 0 : hst + 1;
    }
    cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------------------
// Version 2: warp roles.  Block = 32 × (TY+1) threads.
//   every warp w   : B   — H-tile row w, columns 0..31                     (one position per lane)
//   warps w < TY   : D   — output row w of the plane two steps behind      (one cell per lane)
//   warp  w = TY   : B′  — the extra H column 32 of rows 0..TY             (lanes 0..TY)
// so every warp carries two units of work per z-step and the single barrier per step is balanced.
// The z-directions of C₁ and C₂ are compile-time (FZ1, FZ2); x/y directions only change shared-
// memory offsets and the sign folded into the coefficients.
//   phase m:  issue E slot m−1+SE | B(m), B′(m) | D(m−2) | cp.async.wait, __syncthreads
// ---------------------------------------------------------------------------------------------
template <typename C>
struct PosConst {
    C cxs, cys;  // ±c (direction folded in)
    int flags;   // bits 0-1 xk, 2-3 yk, 4 xl0, 5 yl0
};
template <typename C>
SGC_D PosConst<C> pos_const(const int *fwd, const int *bc_lo, const int *bc_hi, const C *const *c, int gi, int gj, int Nx, int Ny) {
    const XYEdge<C> b = xy_edge<C>(fwd, bc_lo, bc_hi, c, gi, gj, Nx, Ny);
    PosConst<C> r;
    r.cxs = b.cxs; r.cys = b.cys;
    r.flags = b.xk | (b.yk << 2) | (b.xl0 ? 16 : 0) | (b.yl0 ? 32 : 0);
    return r;
}

// ---- shared-memory access with 32-bit addresses and immediate offsets ------------------------------
template <int OFF>
SGC_D cplx lds_at(unsigned a, cplx) {
    double x, y;
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2+%3];" : "=d"(x), "=d"(y) : "r"(a), "n"(OFF));
    return cplx{x, y};
}
template <int OFF>
SGC_D double lds_at(unsigned a, double) {
    double x;
    asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(x) : "r"(a), "n"(OFF));
    return x;
}
template <int OFF>
SGC_D void sts_at(unsigned a, cplx v) {
    asm volatile("st.shared.v2.f64 [%0+%1], {%2,%3};" ::"r"(a), "n"(OFF), "d"(v.re), "d"(v.im) : "memory");
}
template <int OFF>
SGC_D void sts_at(unsigned a, double v) {
    asm volatile("st.shared.f64 [%0+%1], %2;" ::"r"(a), "n"(OFF), "d"(v) : "memory");
}
SGC_D void cp_async_to(unsigned dst, const cplx *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
SGC_D void cp_async_to(unsigned dst, const double *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}

// boundary rules as operand fix-ups for the two terms of one axis (the `dterm` rules, applied to
// the operands so that the arithmetic afterwards is the straight-line  (±c)·(nb − cur)):
//   forward : Bloch edge nb ← e·nb, symmetric edge nb ← 0, symmetric first cell cur ← 0
//   backward: Bloch edge nb ← nb/e, symmetric edge: the terms are zero
template <typename T>
SGC_D void fix_axis(bool fwd, int kind, bool lo_zero, cplx e, int e_real, T &curA, T &nbA, T &curB, T &nbB, bool &zero) {
    const T Z = zero_of(T{});
    if (fwd) {
        if (kind == 1) { nbA = mul_e(e, nbA, e_real); nbB = mul_e(e, nbB, e_real); }
        else if (kind == 2) { nbA = Z; nbB = Z; }
        if (lo_zero) { curA = Z; curB = Z; }
    } else {
        if (kind == 1) { nbA = div_e(nbA, e, e_real); nbB = div_e(nbB, e, e_real); }
        else if (kind == 2) zero = true;
    }
}

// The six ∂ terms of a curl at one point.  flags: bits 0-1 xk, 2-3 yk, 4 xl0, 5 yl0, 8-9 zk, 10 zl0
// (0 for every point that no boundary rule touches: one predictable branch).
template <typename T, typename C, bool FZ>
SGC_D void curl_point(int flags, C cxs, C cys, bool fx, bool fy, C cz, const cplx *e, const int *e_real, T xlo, T xhi, T ylo, T yhi,
                      T zc, T x_yn, T z_yn, T y_xn, T z_xn, T &d0, T &d1, T &d2, T &d3, T &d4, T &d5) {
    const T Z = zero_of(T{});
    T xc_y = FZ ? xlo : xhi, yc_x = FZ ? ylo : yhi, zc_y = zc, zc_x = zc;
    bool z0 = false, y0 = false, x0 = false;
    if (flags) {
        const int zk = (flags >> 8) & 3, yk = (flags >> 2) & 3, xk = flags & 3;
        const bool zl0 = (flags & 1024) != 0, yl0 = (flags & 32) != 0, xl0 = (flags & 16) != 0;
        if (zk | (int)zl0) {
            if (FZ) fix_axis<T>(true, zk, zl0, e[2], e_real[2], xlo, xhi, ylo, yhi, z0);
            else fix_axis<T>(false, zk, false, e[2], e_real[2], xhi, xlo, yhi, ylo, z0);
        }
        if (yk | (int)yl0) fix_axis<T>(fy, yk, yl0, e[1], e_real[1], xc_y, x_yn, zc_y, z_yn, y0);
        if (xk | (int)xl0) fix_axis<T>(fx, xk, xl0, e[0], e_real[0], yc_x, y_xn, zc_x, z_xn, x0);
    }
    d0 = mul(cz, sub(xhi, xlo)); d1 = mul(cz, sub(yhi, ylo));    // ∂zFx, ∂zFy
    d2 = mul(cys, sub(x_yn, xc_y)); d3 = mul(cys, sub(z_yn, zc_y));  // ∂yFx, ∂yFz
    d4 = mul(cxs, sub(y_xn, yc_x)); d5 = mul(cxs, sub(z_xn, zc_x));  // ∂xFy, ∂xFz
    if (flags) {
        if (z0) { d0 = Z; d1 = Z; }
        if (y0) { d2 = Z; d3 = Z; }
        if (x0) { d4 = Z; d5 = Z; }
    }
}

// resident CTAs per SM: 2 for ComplexF64 (128 registers), 3 for Float64 fields (half the registers and shared memory)
template <typename T, typename C, bool ADD, int TY, int SE, bool FZ1, bool FZ2>
__global__ void __launch_bounds__(32 * (TY + 1), (sizeof(T) == 8 ? 3 : 2)) k_curlcurl3w(const __grid_constant__ CurlCurlParams<T, C> p) {
    static_assert(SE >= 3, "need at least 3 stages");
    constexpr int TX = 32;
    constexpr int NT = 32 * (TY + 1);
    constexpr int HX = TX + 1, HY = TY + 1, EX = TX + 2, EY = TY + 2;
    constexpr int HCOMP = HX * HY, ECOMP = EX * EY;
    constexpr int ES = (int)sizeof(T);
    constexpr int ECB = ECOMP * ES, HCB = HCOMP * ES;        // bytes per component tile
    constexpr int ESTAGE = 3 * ECB, HSTAGE = 3 * HCB;         // bytes per ring stage
    constexpr int NLD = (ECOMP + NT - 1) / NT;
    static_assert(HY <= 32, "the extra column is handled by one warp");
    extern __shared__ __align__(16) unsigned char sgc_smem_raw[];
    const unsigned sE = (unsigned)__cvta_generic_to_shared(sgc_smem_raw);
    const unsigned sH = sE + SE * ESTAGE;

    const int lane = threadIdx.x, w = threadIdx.y;
    const int tid = w * 32 + lane;
    const int Nx = p.Nx, Ny = p.Ny, Nz = p.Nz;
    const int64_t sy = Nx, sz = (int64_t)Nx * Ny;
    const bool fx1 = p.fwd1[0], fy1 = p.fwd1[1], fx2 = p.fwd2[0], fy2 = p.fwd2[1];
    const int s1x = fx1 ? 1 : -1, s1y = fy1 ? 1 : -1, s2x = fx2 ? 1 : -1, s2y = fy2 ? 1 : -1;
    constexpr int s1z = FZ1 ? 1 : -1, s2z = FZ2 ? 1 : -1;
    const int i0 = blockIdx.x * TX, j0 = blockIdx.y * TY;
    const int hx0 = i0 + min(0, s2x), hy0 = j0 + min(0, s2y);
    const int ex0 = hx0 + min(0, s1x), ey0 = hy0 + min(0, s1y);
    const T Z = zero_of(T{});
    const bool lo_ext = (p.ghost_lo[0] != nullptr), hi_ext = (p.ghost_hi[0] != nullptr);
    const int ox1 = s1x < 0 ? 1 : 0, oy1 = s1y < 0 ? 1 : 0;

    // ---- loader duty ------------------------------------------------------------------------------
    int ld_g[NLD];  // element offset inside a plane (Nx·Ny < 2^31 is checked on the host), −1: no duty
#pragma unroll
    for (int u = 0; u < NLD; ++u) {
        const int q = tid + u * NT;
        const int py = q / EX, px = q - py * EX;
        ld_g[u] = (q < ECOMP) ? wrap_index(ey0 + py, Ny) * Nx + wrap_index(ex0 + px, Nx) : -1;
    }
    // ---- role B: H position (row w, column lane) -----------------------------------------------------
    const PosConst<C> ca = pos_const<C>(p.fwd1, p.bc1_lo, p.bc1_hi, p.c1, wrap_index(hx0 + lane, Nx), wrap_index(hy0 + w, Ny), Nx, Ny);
    const unsigned a_q = (unsigned)(((w + oy1) * EX + (lane + ox1)) * ES);  // byte offset of the position in an E tile
    const unsigned a_out = sH + (unsigned)((w * HX + lane) * ES);           // its slot in H stage 0
    const int e_dxb = s1x * ES, e_dyb = s1y * EX * ES;
    // ---- second role: D (warps < TY) or B′ (warp TY, lanes < HY) ----------------------------------------
    const bool roleD = (w < TY);
    const bool roleB2 = (w == TY) && (lane < HY);
    const int i = i0 + lane, j = j0 + w;
    const bool valid = roleD && (i < Nx) && (j < Ny);
    PosConst<C> cb;
    unsigned b_q, b_out = 0;
    if (roleD) {
        b_q = (unsigned)(((w + (s2y < 0 ? 1 : 0)) * HX + (lane + (s2x < 0 ? 1 : 0))) * ES);  // the cell in an H tile
        cb = pos_const<C>(p.fwd2, p.bc2_lo, p.bc2_hi, p.c2, valid ? i : 0, valid ? j : 0, Nx, Ny);
    } else {
        const int py = roleB2 ? lane : 0;
        b_q = (unsigned)(((py + oy1) * EX + (TX + ox1)) * ES);
        b_out = sH + (unsigned)((py * HX + TX) * ES);
        cb = pos_const<C>(p.fwd1, p.bc1_lo, p.bc1_hi, p.c1, wrap_index(hx0 + TX, Nx), wrap_index(hy0 + py, Ny), Nx, Ny);
    }
    const int h_dxb = s2x * ES, h_dyb = s2y * HX * ES;

    const int k0 = p.k_begin + blockIdx.z * p.march;
    const int k1 = min(k0 + p.march, p.k_end);
    if (k0 >= k1) return;
    const int n = k1 - k0;
    const int hb = k0 + min(0, s2z);
    const int eb = hb + min(0, s1z);

    // ---- producer ------------------------------------------------------------------------------------------
    int jj_issue = 0;
    unsigned st_issue = sE + (unsigned)(tid * ES);
    auto issue = [&]() {
        if (jj_issue <= n + 1) {
            const int ke = eb + jj_issue;
            const T *s0, *s1, *s2;
            if (ke >= 0 && ke < Nz) {
                const int64_t off = (int64_t)ke * sz;
                s0 = p.F[0] + off; s1 = p.F[1] + off; s2 = p.F[2] + off;
            } else if (ke < 0 && lo_ext) { s0 = p.ghost_lo[0]; s1 = p.ghost_lo[1]; s2 = p.ghost_lo[2]; }
            else if (ke >= Nz && hi_ext) { s0 = p.ghost_hi[0]; s1 = p.ghost_hi[1]; s2 = p.ghost_hi[2]; }
            else {
                const int64_t off = (int64_t)wrap_index(ke, Nz) * sz;
                s0 = p.F[0] + off; s1 = p.F[1] + off; s2 = p.F[2] + off;
            }
            const bool centre = FZ1 ? (jj_issue <= n) : (jj_issue >= 1);
#pragma unroll
            for (int u = 0; u < NLD; ++u) {
                if (ld_g[u] >= 0) {
                    cp_async_to(st_issue + u * NT * ES, s0 + ld_g[u]);
                    cp_async_to(st_issue + u * NT * ES + ECB, s1 + ld_g[u]);
                    if (centre) cp_async_to(st_issue + u * NT * ES + 2 * ECB, s2 + ld_g[u]);
                }
            }
        }
        cp_async_commit();
        ++jj_issue;
        st_issue += ESTAGE;
        if (st_issue >= sE + (unsigned)(SE * ESTAGE)) st_issue -= SE * ESTAGE;
    };
#pragma unroll
    for (int q = 0; q < SE; ++q) issue();
    cp_async_wait<SE - 2>();
    __syncthreads();

    unsigned e_lo = sE;      // ring stage of E slot m
    unsigned h_cur = 0;      // byte offset of the H stage of slot m
    int64_t off_out = (int64_t)k0 * sz + (int64_t)j * sy + i;
    // the carried "lo" plane values: E slot 0 for the H positions, H slot 0 for the output cell (set at m = 2)
    T axlo = lds_at<0>(e_lo + a_q, T{}), aylo = lds_at<ECB>(e_lo + a_q, T{});
    T bxlo = Z, bylo = Z;
    if (!roleD) { bxlo = lds_at<0>(e_lo + b_q, T{}); bylo = lds_at<ECB>(e_lo + b_q, T{}); }

    // one H value from the staged E planes (role B and role B′ share this)
    auto h_point = [&](unsigned q, const PosConst<C> &c, T &xlo_c, T &ylo_c, unsigned out, int zflags, C cz, unsigned e_hi) {
        const unsigned ahi = e_hi + q;
        const unsigned ac = (FZ1 ? e_lo : e_hi) + q;
        const T xhi = lds_at<0>(ahi, T{}), yhi = lds_at<ECB>(ahi, T{});
        const T zc = lds_at<2 * ECB>(ac, T{});
        const T x_yn = lds_at<0>(ac + e_dyb, T{}), z_yn = lds_at<2 * ECB>(ac + e_dyb, T{});
        const T y_xn = lds_at<ECB>(ac + e_dxb, T{}), z_xn = lds_at<2 * ECB>(ac + e_dxb, T{});
        T d0, d1, d2, d3, d4, d5;
        curl_point<T, C, FZ1>(c.flags | zflags, c.cxs, c.cys, fx1, fy1, cz, p.e1, p.e1_real, xlo_c, xhi, ylo_c, yhi, zc, x_yn, z_yn, y_xn,
                              z_xn, d0, d1, d2, d3, d4, d5);
        sts_at<0>(out + h_cur, add(add(Z, neg(d1)), d3));  // Gv .= 0 ; += (−∂zFy) ; += ∂yFz   (curl.jl:66-91)
        sts_at<HCB>(out + h_cur, add(d0, neg(d5)));
        sts_at<2 * HCB>(out + h_cur, add(neg(d2), d4));
        xlo_c = xhi;
        ylo_c = yhi;
    };

    int hk = hb;  // grid plane of H slot m (un-wrapped)
    for (int m = 0; m <= n + 1; ++m, ++hk) {
        if (m >= 1) issue();  // E slot m − 1 + SE → the stage slot m − 1 used (free since the last barrier)
        if (m <= n) {
            // ================= B(m), B′(m): H plane hb + m =============================================
            int zflags = 0;
            C cz;
            if (hk > 0 && hk < Nz - 1) {
                cz = p.c1[2][hk];
            } else {
                const bool ghostH = (hk < 0 && lo_ext) || (hk >= Nz && hi_ext);
                const int kk = ghostH ? hk : wrap_index(hk, Nz);
                cz = ghostH ? (hk < 0 ? p.c1z_lo : p.c1z_hi) : p.c1[2][kk];
                if (ghostH) {
                    const int side = (hk < 0) ? p.bc1_lo[2] : p.bc1_hi[2];
                    if (side == BC_BLOCH) zflags = 1 << 8;
                } else if (FZ1) {
                    if (kk == Nz - 1) zflags = ((p.bc1_hi[2] == BC_BLOCH) ? 1 : (p.bc1_hi[2] == BC_SYM ? 2 : 0)) << 8;
                    if (kk == 0 && p.bc1_lo[2] == BC_SYM) zflags |= 1024;
                } else if (kk == 0) {
                    zflags = ((p.bc1_lo[2] == BC_BLOCH) ? 1 : (p.bc1_lo[2] == BC_SYM ? 2 : 0)) << 8;
                }
            }
            unsigned e_hi = e_lo + ESTAGE;
            if (e_hi >= sE + (unsigned)(SE * ESTAGE)) e_hi = sE;
            h_point(a_q, ca, axlo, aylo, a_out, zflags, cz, e_hi);
            if (roleB2) h_point(b_q, cb, bxlo, bylo, b_out, zflags, cz, e_hi);
            e_lo = e_hi;
        }
        if (roleD && m >= 2) {
            // ================= D(m − 2): output plane k0 + m − 2 from H slots m − 2 (lo), m − 1 (hi) =======
            const int k = k0 + m - 2;
            T g0 = Z, g1 = Z, g2 = Z;
            if (ADD && valid) { g0 = ldg_plain(p.G[0] + off_out); g1 = ldg_plain(p.G[1] + off_out); g2 = ldg_plain(p.G[2] + off_out); }
            // h_cur is the stage of slot m (also at m = n + 1); slots m − 1 and m − 2 precede it in the ring
            const unsigned st_hi = (h_cur == 0) ? 2u * HSTAGE : h_cur - HSTAGE;            // slot m − 1
            const unsigned st_lo = (st_hi == 0) ? 2u * HSTAGE : st_hi - HSTAGE;           // slot m − 2
            const unsigned ahi = sH + st_hi + b_q, alo = sH + st_lo + b_q;
            const unsigned ac = FZ2 ? alo : ahi;
            if (m == 2) { bxlo = lds_at<0>(alo, T{}); bylo = lds_at<HCB>(alo, T{}); }
            const T xhi = lds_at<0>(ahi, T{}), yhi = lds_at<HCB>(ahi, T{});
            const T zc = lds_at<2 * HCB>(ac, T{});
            const T x_yn = lds_at<0>(ac + h_dyb, T{}), z_yn = lds_at<2 * HCB>(ac + h_dyb, T{});
            const T y_xn = lds_at<HCB>(ac + h_dxb, T{}), z_xn = lds_at<2 * HCB>(ac + h_dxb, T{});
            const C cz = p.c2[2][k];
            int zflags = 0;
            if (k == 0 || k == Nz - 1) {
                if (FZ2) {
                    if (k == Nz - 1) zflags = ((p.bc2_hi[2] == BC_BLOCH) ? 1 : (p.bc2_hi[2] == BC_SYM ? 2 : 0)) << 8;
                    if (k == 0 && p.bc2_lo[2] == BC_SYM) zflags |= 1024;
                } else if (k == 0) {
                    zflags = ((p.bc2_lo[2] == BC_BLOCH) ? 1 : (p.bc2_lo[2] == BC_SYM ? 2 : 0)) << 8;
                }
            }
            T d0, d1, d2, d3, d4, d5;
            curl_point<T, C, FZ2>(cb.flags | zflags, cb.cxs, cb.cys, fx2, fy2, cz, p.e2, p.e2_real, bxlo, xhi, bylo, yhi, zc, x_yn, z_yn,
                                  y_xn, z_xn, d0, d1, d2, d3, d4, d5);
            T gx, gy, gz;
            if (ADD) { gx = add(add(g0, neg(d1)), d3); gy = add(add(g1, d0), neg(d5)); gz = add(add(g2, neg(d2)), d4); }
            else     { gx = add(add(Z, neg(d1)), d3);  gy = add(d0, neg(d5));          gz = add(neg(d2), d4); }
            if (valid) {
                stg_stream(p.G[0] + off_out, gx);
                stg_stream(p.G[1] + off_out, gy);
                stg_stream(p.G[2] + off_out, gz);
            }
            bxlo = xhi;
            bylo = yhi;
            off_out += sz;
        }
        if (m <= n) h_cur = (h_cur == 2u * HSTAGE) ? 0u : h_cur + HSTAGE;
        cp_async_wait<SE - 3>();  // E slot m + 2 has landed (for this thread) …
        __syncthreads();          // … and for everybody; H plane m is complete; E slot m is free
    }
    cp_async_wait<0>();
}

}  // namespace sgc
