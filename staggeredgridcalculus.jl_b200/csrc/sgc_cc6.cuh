// sgc_cc6.cuh — fused curl-of-curl, sixth generation: TWO z-planes per barrier interval (sm_100a).
//
// Same mathematics and the same thread mapping as k_cc5 (sgc_cc4.cuh): G OP C₂(C₁ F) with H = C₁ F in shared memory
// only, every value produced by the arithmetic of two apply_curl! calls (matrixfree/differential/curl.jl:52-137), so
// the result is bit-identical to the two-pass path.  What the ncu source page of k_cc5 said (334 thread-instructions
// per cell of which 130 FP64; issue slots 45 % busy at 12 warps/SM; stall reasons: fixed-latency dependencies 27 %,
// barrier 16 %, shared-memory scoreboard 14 %): a warp holds ONE H point and ONE output point of independent work
// between two CTA-wide barriers.  Here every thread computes the H points of two consecutive planes, then the output
// points of two consecutive planes, between one pair of barriers:
//
//   * (steady state of boundary-free tiles: ONE straight-line block per iteration, the two output points first — they only
//     need the H pair of the previous iteration — so that the wait for the arriving E pair sits behind their arithmetic)
//   * twice the independent arithmetic per thread in one straight-line block (the boundary rules of both points sit
//     behind ONE branch), half the barriers, loop control, ring bookkeeping and mbarrier waits per cell;
//   * the plane in the middle of a pair is the upper plane of the first point and the lower plane of the second: its
//     own-position values are loaded once and never copied between registers;
//   * E planes arrive in PAIRS (six TMA box copies, one mbarrier) in an NP-stage ring of pairs; the H ring holds two
//     pairs (the one being written, the one being read);
//   * the per-plane z coefficients of the chunk are staged in shared memory once (uniform LDS instead of global loads).
//
// Applicability: ComplexF64 fields, C₁ forward and C₂ backward in z (the Maxwell pair E → H → E′), symmetric boundaries
// on all three axes (the PEC / PMC box of a PML-terminated domain: out-of-range box cells are zero-filled, which is what
// those rules discard, and every rule is a select — the kernel makes no call and has no stack frame), no slab ghost
// planes, z-chunks of at most CC6_MAXMARCH planes.  Everything else runs the third generation.
//
// Plane bookkeeping for a chunk of n output planes k0 … k0+n−1 (hb = k0 − 1):
//   E slot s ↔ plane hb + s,  s = 0 … n+1;  pair P_u = slots (2u, 2u+1)  → ring stage u mod NP
//   H slot s ↔ plane hb + s,  s = 0 … n;    = C₁ applied to E slots s (centre) and s+1;  stage (s+1) mod 4
//   output kk ↔ plane k0 + kk = C₂ applied to H slots kk (lower) and kk+1 (centre)
//   iteration t = 0 … T+1, T = ⌈n/2⌉:   t ≤ T: H slots 2t−1, 2t (slot −1 of t = 0 is computed from whatever the
//   ring holds and never read);  t ≥ 2: outputs 2t−4, 2t−3 from the H pair written in iteration t−1;  one barrier.
#pragma once
#include "sgc_cc4.cuh"

namespace sgc {

constexpr int CC6_MAXMARCH = 64;
constexpr int CC6_CZ = CC6_MAXMARCH + 4;  // entries of one staged coefficient table
constexpr int cc6_epair(int TY) { return (6 * 34 * (TY + 2) * 16 + 127) / 128 * 128; }  // one pair of E planes, three components, whole lines
constexpr size_t cc6_smem_bytes(int TY, int NP) {
    return (size_t)NP * cc6_epair(TY) + (size_t)4 * 3 * 33 * (TY + 1) * 16 + (size_t)2 * CC6_CZ * 16 + 8 * NP + 128;
}
SGC_D void tma_load_4d(unsigned dst, const CUtensorMap *tm, unsigned bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// The six ∂ terms at a point that a SYMMETRIC boundary rule touches (the rules of curl_point, sgc_curlcurl.cuh, without
// the Bloch cases: this kernel only runs on all-symmetric boxes): operands or terms replaced by zero behind selects —
// no call, no division, so the kernel needs no stack frame.  flags: bits 0-1 xk, 2-3 yk, 4 xl0, 5 yl0, 8-9 zk, 10 zl0.
template <typename T, typename C, bool FZ>
SGC_D void curl_point_sym(int flags, C cxs, C cys, bool fx, bool fy, C cz, T xlo, T xhi, T ylo, T yhi, T zc, T x_yn, T z_yn, T y_xn, T z_xn,
                          T &d0, T &d1, T &d2, T &d3, T &d4, T &d5) {
    const T Z = zero_of(T{});
    T xc_y = FZ ? xlo : xhi, yc_x = FZ ? ylo : yhi, zc_y = zc, zc_x = zc;
    const bool zedge = ((flags >> 8) & 3) == 2, yedge = ((flags >> 2) & 3) == 2, xedge = (flags & 3) == 2;
    const bool zl0 = (flags & 1024) != 0, yl0 = (flags & 32) != 0, xl0 = (flags & 16) != 0;
    if (FZ) {
        if (zedge) { xhi = Z; yhi = Z; }
        if (zl0) { xlo = Z; ylo = Z; }
    }
    if (fy) {
        if (yedge) { x_yn = Z; z_yn = Z; }
        if (yl0) { xc_y = Z; zc_y = Z; }
    }
    if (fx) {
        if (xedge) { y_xn = Z; z_xn = Z; }
        if (xl0) { yc_x = Z; zc_x = Z; }
    }
    d0 = mul(cz, sub(xhi, xlo)); d1 = mul(cz, sub(yhi, ylo));
    d2 = mul(cys, sub(x_yn, xc_y)); d3 = mul(cys, sub(z_yn, zc_y));
    d4 = mul(cxs, sub(y_xn, yc_x)); d5 = mul(cxs, sub(z_xn, zc_x));
    if (!FZ && zedge) { d0 = Z; d1 = Z; }
    if (!fy && yedge) { d2 = Z; d3 = Z; }
    if (!fx && xedge) { d4 = Z; d5 = Z; }
}

template <typename C, bool ADD, int TY, int NP, int MINB, int SP>
__global__ void __launch_bounds__(32 * (TY + 1), MINB) k_cc6(const __grid_constant__ CurlCurl4Params<C> P) {
    using T = cplx;
    static_assert(NP >= 3, "the E ring needs at least 3 pairs");
    const CurlCurlParams<T, C> &p = P.q;
    constexpr int TX = 32;
    constexpr int HX = TX + 1, HY = TY + 1, EX = TX + 2, EY = TY + 2;
    constexpr int ES = 16;
    constexpr int ECOMP = EX * EY;
    // one TMA box = (34 columns, TY+2 rows, 2 planes, 3 components) of the field array seen as a 4-D tensor: in shared
    // memory component c of plane q of a pair sits at (2c + q)·EPLANE
    constexpr int EPLANE = ECOMP * ES, ECB = 2 * EPLANE, EPAIR = cc6_epair(TY);
    constexpr int HCB = HX * HY * ES, HSTAGE = 3 * HCB;
    constexpr int CS = (int)sizeof(C);
    static_assert(HY <= 32, "the extra column is handled by one warp");
    extern __shared__ __align__(128) unsigned char sgc_smem_raw[];
    const unsigned sE = ((unsigned)__cvta_generic_to_shared(sgc_smem_raw) + 127u) & ~127u;
    const unsigned sH = sE + NP * EPAIR;
    const unsigned sC1 = sH + 4 * HSTAGE, sC2 = sC1 + CC6_CZ * 16;
    const unsigned bar_efull = sC2 + CC6_CZ * 16;

    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int Nx = p.Nx, Ny = p.Ny, Nz = p.Nz;
    const int64_t sy = Nx, sz = (int64_t)Nx * Ny;
    const bool fx1 = p.fwd1[0], fy1 = p.fwd1[1], fx2 = p.fwd2[0], fy2 = p.fwd2[1];
    const int s1x = fx1 ? 1 : -1, s1y = fy1 ? 1 : -1, s2x = fx2 ? 1 : -1, s2y = fy2 ? 1 : -1;
    const int i0 = blockIdx.x * TX, j0 = blockIdx.y * TY;
    const int hx0 = i0 + min(0, s2x), hy0 = j0 + min(0, s2y);
    const int ex0 = hx0 + min(0, s1x), ey0 = hy0 + min(0, s1y);
    const T Z = zero_of(T{});
    const int ox1 = s1x < 0 ? 1 : 0, oy1 = s1y < 0 ? 1 : 0;

    const int k0 = p.k_begin + blockIdx.z * p.march;
    const int k1 = min(k0 + p.march, p.k_end);
    if (k0 >= k1) return;
    const int n = k1 - k0;
    const int hb = k0 - 1;            // grid plane of H slot 0 and of E slot 0 (C₁ forward, C₂ backward)
    const int Th = (n + 1) >> 1;      // last iteration with an H pair; pairs P_0 … P_Th of E exist

    if (threadIdx.x == 0) {
        for (int s = 0; s < NP; ++s) mbar_init(bar_efull + 8 * s, 1u);
        mbar_fence_init();
    }
    // z coefficients of the chunk: c1z entry e ↔ H slot e−1 (plane hb−1+e), c2z entry e ↔ output plane k0+e
    for (int e = threadIdx.x; e < CC6_CZ; e += 32 * (TY + 1)) {
        const C a = p.c1[2][wrap_index(hb - 1 + e, Nz)];
        const C b = p.c2[2][min(k0 + e, Nz - 1)];
        sts_at<0>(sC1 + e * CS, a);
        sts_at<0>(sC2 + e * CS, b);
    }
    __syncthreads();

    // ---- producer duty of thread 0: pair P_u → ring stage u mod NP ---------------------------------------------------
    // ONE 4-D box copy per pair.  (Six 3-D copies from this thread put ≈ 150 instructions on warp 0's path to every
    // barrier; spreading those six over the warps measured 6 % slower, and an additional L2 prefetch of the pair after
    // next, cp.async.bulk.prefetch.tensor, changed nothing: profiles/r2_sweep_cc6_*.jsonl.)
    int issue_u = 0, issue_stage = 0;
    auto issue = [&]() {
        if (issue_u <= Th) {  // planes hb+2u, hb+2u+1: a plane outside the grid is zero-filled (and discarded by the symmetric rules)
            const unsigned st = sE + issue_stage * EPAIR, bar = bar_efull + 8 * issue_stage;
            mbar_arrive_expect_tx(bar, 6 * ECOMP * ES);
            tma_load_4d(st, &P.tm[0], bar, 2 * ex0, ey0, hb + 2 * issue_u, 0);
        }
        ++issue_u;
        issue_stage = (issue_stage + 1 == NP) ? 0 : issue_stage + 1;
    };
    if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < NP; ++q) issue();
    }

    // ---- role A: H position (row w, column lane) ---------------------------------------------------------------------
    const PosConst<C> ca = pos_const<C>(p.fwd1, p.bc1_lo, p.bc1_hi, p.c1, wrap_index(hx0 + lane, Nx), wrap_index(hy0 + w, Ny), Nx, Ny);
    const unsigned a_q = (unsigned)(((w + oy1) * EX + (lane + ox1)) * ES);
    const unsigned a_out = sH + (unsigned)((w * HX + lane) * ES);
    const int e_dxb = s1x * ES, e_dyb = s1y * EX * ES;
    // ---- second role: D (warps < TY) or B′ (warp TY, lanes < HY: the extra H column) ---------------------------------
    const bool roleD = (w < TY);
    const bool roleB2 = (w == TY) && (lane < HY);
    const int i = i0 + lane, j = j0 + w;
    const bool valid = roleD && (i < Nx) && (j < Ny);
    PosConst<C> cb;
    unsigned b_q, b_out = 0;
    if (roleD) {
        b_q = (unsigned)(((w + (s2y < 0 ? 1 : 0)) * HX + (lane + (s2x < 0 ? 1 : 0))) * ES);
        cb = pos_const<C>(p.fwd2, p.bc2_lo, p.bc2_hi, p.c2, valid ? i : 0, valid ? j : 0, Nx, Ny);
    } else {
        const int py = roleB2 ? lane : 0;
        b_q = (unsigned)(((py + oy1) * EX + (TX + ox1)) * ES);
        b_out = sH + (unsigned)((py * HX + TX) * ES);
        cb = pos_const<C>(p.fwd1, p.bc1_lo, p.bc1_hi, p.c1, wrap_index(hx0 + TX, Nx), wrap_index(hy0 + py, Ny), Nx, Ny);
    }
    const int h_dxb = s2x * ES, h_dyb = s2y * HX * ES;

    // boundary flags of an H plane (C₁ forward) and of an output plane (C₂ backward); 0 for interior planes
    auto zflags_h = [&](int hk) -> int {
        if (hk > 0 && hk < Nz - 1) return 0;
        const int kk = wrap_index(hk, Nz);
        int f = 0;
        if (kk == Nz - 1) f = ((p.bc1_hi[2] == BC_BLOCH) ? 1 : (p.bc1_hi[2] == BC_SYM ? 2 : 0)) << 8;
        if (kk == 0 && p.bc1_lo[2] == BC_SYM) f |= 1024;
        return f;
    };
    auto zflags_d = [&](int k) -> int {
        return (k == 0) ? (((p.bc2_lo[2] == BC_BLOCH) ? 1 : (p.bc2_lo[2] == BC_SYM ? 2 : 0)) << 8) : 0;
    };

    // two H values (slots 2t−1 and 2t) at one tile position from the staged E planes eA (slot 2t−1), eB, eC
    auto h_pair = [&](unsigned q, const PosConst<C> &c, T &xA, T &yA, unsigned out, int zfA, int zfB, C czA, C czB, unsigned eA, unsigned eB,
                      unsigned eC) {
        const unsigned aA = eA + q, aB = eB + q, aC = eC + q;
        const T xB = lds_at<0>(aB, T{}), yB = lds_at<ECB>(aB, T{});
        const T zA = lds_at<2 * ECB>(aA, T{});
        const T xynA = lds_at<0>(aA + e_dyb, T{}), zynA = lds_at<2 * ECB>(aA + e_dyb, T{});
        const T yxnA = lds_at<ECB>(aA + e_dxb, T{}), zxnA = lds_at<2 * ECB>(aA + e_dxb, T{});
        const T xC = lds_at<0>(aC, T{}), yC = lds_at<ECB>(aC, T{});
        const T zB = lds_at<2 * ECB>(aB, T{});
        const T xynB = lds_at<0>(aB + e_dyb, T{}), zynB = lds_at<2 * ECB>(aB + e_dyb, T{});
        const T yxnB = lds_at<ECB>(aB + e_dxb, T{}), zxnB = lds_at<2 * ECB>(aB + e_dxb, T{});
        T a0, a1, a2, a3, a4, a5, b0, b1, b2, b3, b4, b5;
        if ((c.flags | zfA | zfB) == 0) {
            a0 = mul(czA, sub(xB, xA)); a1 = mul(czA, sub(yB, yA));
            a2 = mul(c.cys, sub(xynA, xA)); a3 = mul(c.cys, sub(zynA, zA));
            a4 = mul(c.cxs, sub(yxnA, yA)); a5 = mul(c.cxs, sub(zxnA, zA));
            b0 = mul(czB, sub(xC, xB)); b1 = mul(czB, sub(yC, yB));
            b2 = mul(c.cys, sub(xynB, xB)); b3 = mul(c.cys, sub(zynB, zB));
            b4 = mul(c.cxs, sub(yxnB, yB)); b5 = mul(c.cxs, sub(zxnB, zB));
        } else {
            curl_point_sym<T, C, true>(c.flags | zfA, c.cxs, c.cys, fx1, fy1, czA, xA, xB, yA, yB, zA, xynA, zynA, yxnA, zxnA, a0, a1, a2, a3, a4, a5);
            curl_point_sym<T, C, true>(c.flags | zfB, c.cxs, c.cys, fx1, fy1, czB, xB, xC, yB, yC, zB, xynB, zynB, yxnB, zxnB, b0, b1, b2, b3, b4, b5);
        }
        sts_at<0>(out, add(add(Z, neg(a1)), a3));  // Gv .= 0 ; += (−∂zFy) ; += ∂yFz   (curl.jl:66-91)
        sts_at<HCB>(out, add(a0, neg(a5)));
        sts_at<2 * HCB>(out, add(neg(a2), a4));
        sts_at<HSTAGE>(out, add(add(Z, neg(b1)), b3));
        sts_at<HSTAGE + HCB>(out, add(b0, neg(b5)));
        sts_at<HSTAGE + 2 * HCB>(out, add(neg(b2), b4));
        xA = xC;
        yA = yC;
    };

    T axlo = Z, aylo = Z, bxlo = Z, bylo = Z;
    int st_cur = 0, st_prev = NP - 1;  // ring stages of P_t and P_{t−1}
    unsigned e_phase = 0;
    unsigned h_wr = 0;                 // byte offset of the H pair written in iteration t (the other pair is being read)
    int64_t off_out = (int64_t)k0 * sz + (int64_t)j * sy + i;

    // no H position (hence no output cell) of this tile lies on an x / y boundary, and the tile is whole
    const bool tile_plain = (hx0 >= 1) && (hx0 + HX - 1 <= Nx - 2) && (hy0 >= 1) && (hy0 + HY - 1 <= Ny - 2);

    for (int t = 0; t <= Th + 1; ++t) {
        if (t >= 2 && threadIdx.x == 0) issue();  // P_{t−2+NP} → the stage P_{t−2} used (free since the last barrier)
        // ---- steady state of a boundary-free tile (SP): both phases in ONE straight-line block, the loads of the output
        // points issued between the arithmetic of the H points, so that shared-memory latency and FP64 work overlap
        // inside a warp as well as between warps.  Warp TY (the extra column) keeps the two-phase code below.
        if (SP == 2 && tile_plain && roleD && t >= 2 && t <= Th && k0 + 2 * t - 4 >= 1 && k0 + 2 * t - 1 <= Nz - 2 && 2 * t - 3 < n) {
            // outputs FIRST: they only need the H pair of the previous iteration, which the barrier has already made
            // visible, so the wait for the arriving E pair moves behind the arithmetic of the first output point
            const C czA = lds_at<0>(sC1 + (unsigned)(2 * t) * CS, C{}), czB = lds_at<CS>(sC1 + (unsigned)(2 * t) * CS, C{});
            const C dzA = lds_at<0>(sC2 + (unsigned)(2 * t - 4) * CS, C{}), dzB = lds_at<CS>(sC2 + (unsigned)(2 * t - 4) * CS, C{});
            T gA0 = Z, gA1 = Z, gA2 = Z, gB0 = Z, gB1 = Z, gB2 = Z;
            if (ADD) {
                gA0 = ldg_plain(p.G[0] + off_out); gA1 = ldg_plain(p.G[1] + off_out); gA2 = ldg_plain(p.G[2] + off_out);
                gB0 = ldg_plain(p.G[0] + off_out + sz); gB1 = ldg_plain(p.G[1] + off_out + sz); gB2 = ldg_plain(p.G[2] + off_out + sz);
            }
            const unsigned out = a_out + h_wr;
            const unsigned hA = sH + (h_wr ^ (2u * HSTAGE)) + b_q, hB = hA + HSTAGE;
            const T hxA = lds_at<0>(hA, T{}), hyA = lds_at<HCB>(hA, T{}), hzA = lds_at<2 * HCB>(hA, T{});
            const T hxynA = lds_at<0>(hA + h_dyb, T{}), hzynA = lds_at<2 * HCB>(hA + h_dyb, T{});
            const T hyxnA = lds_at<HCB>(hA + h_dxb, T{}), hzxnA = lds_at<2 * HCB>(hA + h_dxb, T{});
            const T hxB = lds_at<0>(hB, T{}), hyB = lds_at<HCB>(hB, T{}), hzB = lds_at<2 * HCB>(hB, T{});
            const T hxynB = lds_at<0>(hB + h_dyb, T{}), hzynB = lds_at<2 * HCB>(hB + h_dyb, T{});
            const T hyxnB = lds_at<HCB>(hB + h_dxb, T{}), hzxnB = lds_at<2 * HCB>(hB + h_dxb, T{});
            {   // output 2t−4
                const T a0 = mul(dzA, sub(hxA, bxlo)), a1 = mul(dzA, sub(hyA, bylo));
                const T a2 = mul(cb.cys, sub(hxynA, hxA)), a3 = mul(cb.cys, sub(hzynA, hzA));
                const T a4 = mul(cb.cxs, sub(hyxnA, hyA)), a5 = mul(cb.cxs, sub(hzxnA, hzA));
                stg_stream(p.G[0] + off_out, ADD ? add(add(gA0, neg(a1)), a3) : add(add(Z, neg(a1)), a3));
                stg_stream(p.G[1] + off_out, ADD ? add(add(gA1, a0), neg(a5)) : add(a0, neg(a5)));
                stg_stream(p.G[2] + off_out, ADD ? add(add(gA2, neg(a2)), a4) : add(neg(a2), a4));
            }
            mbar_wait(bar_efull + 8 * st_cur, e_phase);  // pair P_t has landed
            const unsigned aA = sE + st_prev * EPAIR + EPLANE + a_q, aB = sE + st_cur * EPAIR + a_q, aC = aB + EPLANE;
            const T xB = lds_at<0>(aB, T{}), yB = lds_at<ECB>(aB, T{});
            const T zA = lds_at<2 * ECB>(aA, T{});
            const T xynA = lds_at<0>(aA + e_dyb, T{}), zynA = lds_at<2 * ECB>(aA + e_dyb, T{});
            const T yxnA = lds_at<ECB>(aA + e_dxb, T{}), zxnA = lds_at<2 * ECB>(aA + e_dxb, T{});
            const T xC = lds_at<0>(aC, T{}), yC = lds_at<ECB>(aC, T{});
            const T zB = lds_at<2 * ECB>(aB, T{});
            const T xynB = lds_at<0>(aB + e_dyb, T{}), zynB = lds_at<2 * ECB>(aB + e_dyb, T{});
            const T yxnB = lds_at<ECB>(aB + e_dxb, T{}), zxnB = lds_at<2 * ECB>(aB + e_dxb, T{});
            {   // output 2t−3
                const T b0 = mul(dzB, sub(hxB, hxA)), b1 = mul(dzB, sub(hyB, hyA));
                const T b2 = mul(cb.cys, sub(hxynB, hxB)), b3 = mul(cb.cys, sub(hzynB, hzB));
                const T b4 = mul(cb.cxs, sub(hyxnB, hyB)), b5 = mul(cb.cxs, sub(hzxnB, hzB));
                stg_stream(p.G[0] + off_out + sz, ADD ? add(add(gB0, neg(b1)), b3) : add(add(Z, neg(b1)), b3));
                stg_stream(p.G[1] + off_out + sz, ADD ? add(add(gB1, b0), neg(b5)) : add(b0, neg(b5)));
                stg_stream(p.G[2] + off_out + sz, ADD ? add(add(gB2, neg(b2)), b4) : add(neg(b2), b4));
            }
            bxlo = hxB;
            bylo = hyB;
            {   // H slot 2t−1
                const T a0 = mul(czA, sub(xB, axlo)), a1 = mul(czA, sub(yB, aylo));
                const T a2 = mul(ca.cys, sub(xynA, axlo)), a3 = mul(ca.cys, sub(zynA, zA));
                const T a4 = mul(ca.cxs, sub(yxnA, aylo)), a5 = mul(ca.cxs, sub(zxnA, zA));
                sts_at<0>(out, add(add(Z, neg(a1)), a3));
                sts_at<HCB>(out, add(a0, neg(a5)));
                sts_at<2 * HCB>(out, add(neg(a2), a4));
            }
            {   // H slot 2t
                const T b0 = mul(czB, sub(xC, xB)), b1 = mul(czB, sub(yC, yB));
                const T b2 = mul(ca.cys, sub(xynB, xB)), b3 = mul(ca.cys, sub(zynB, zB));
                const T b4 = mul(ca.cxs, sub(yxnB, yB)), b5 = mul(ca.cxs, sub(zxnB, zB));
                sts_at<HSTAGE>(out, add(add(Z, neg(b1)), b3));
                sts_at<HSTAGE + HCB>(out, add(b0, neg(b5)));
                sts_at<HSTAGE + 2 * HCB>(out, add(neg(b2), b4));
            }
            axlo = xC;
            aylo = yC;
            off_out += 2 * sz;
            st_prev = st_cur;
            if (++st_cur == NP) { st_cur = 0; e_phase ^= 1u; }
            h_wr ^= 2u * HSTAGE;
            __syncthreads();
            continue;
        }
        if (SP == 1 && tile_plain && roleD && t >= 2 && t <= Th && k0 + 2 * t - 4 >= 1 && k0 + 2 * t - 1 <= Nz - 2 && 2 * t - 3 < n) {
            const C czA = lds_at<0>(sC1 + (unsigned)(2 * t) * CS, C{}), czB = lds_at<CS>(sC1 + (unsigned)(2 * t) * CS, C{});
            const C dzA = lds_at<0>(sC2 + (unsigned)(2 * t - 4) * CS, C{}), dzB = lds_at<CS>(sC2 + (unsigned)(2 * t - 4) * CS, C{});
            T gA0 = Z, gA1 = Z, gA2 = Z, gB0 = Z, gB1 = Z, gB2 = Z;
            if (ADD) {
                gA0 = ldg_plain(p.G[0] + off_out); gA1 = ldg_plain(p.G[1] + off_out); gA2 = ldg_plain(p.G[2] + off_out);
                gB0 = ldg_plain(p.G[0] + off_out + sz); gB1 = ldg_plain(p.G[1] + off_out + sz); gB2 = ldg_plain(p.G[2] + off_out + sz);
            }
            mbar_wait(bar_efull + 8 * st_cur, e_phase);  // pair P_t has landed
            const unsigned aA = sE + st_prev * EPAIR + EPLANE + a_q, aB = sE + st_cur * EPAIR + a_q, aC = aB + EPLANE;
            const unsigned out = a_out + h_wr;
            const unsigned hA = sH + (h_wr ^ (2u * HSTAGE)) + b_q, hB = hA + HSTAGE;
            // E operands of the two H points
            const T xB = lds_at<0>(aB, T{}), yB = lds_at<ECB>(aB, T{});
            const T zA = lds_at<2 * ECB>(aA, T{});
            const T xynA = lds_at<0>(aA + e_dyb, T{}), zynA = lds_at<2 * ECB>(aA + e_dyb, T{});
            const T yxnA = lds_at<ECB>(aA + e_dxb, T{}), zxnA = lds_at<2 * ECB>(aA + e_dxb, T{});
            const T xC = lds_at<0>(aC, T{}), yC = lds_at<ECB>(aC, T{});
            const T zB = lds_at<2 * ECB>(aB, T{});
            const T xynB = lds_at<0>(aB + e_dyb, T{}), zynB = lds_at<2 * ECB>(aB + e_dyb, T{});
            const T yxnB = lds_at<ECB>(aB + e_dxb, T{}), zxnB = lds_at<2 * ECB>(aB + e_dxb, T{});
            {   // H slot 2t−1
                const T a0 = mul(czA, sub(xB, axlo)), a1 = mul(czA, sub(yB, aylo));
                const T a2 = mul(ca.cys, sub(xynA, axlo)), a3 = mul(ca.cys, sub(zynA, zA));
                const T a4 = mul(ca.cxs, sub(yxnA, aylo)), a5 = mul(ca.cxs, sub(zxnA, zA));
                sts_at<0>(out, add(add(Z, neg(a1)), a3));
                sts_at<HCB>(out, add(a0, neg(a5)));
                sts_at<2 * HCB>(out, add(neg(a2), a4));
            }
            // H operands of output 2t−4 (pair written in iteration t−1)
            const T hxA = lds_at<0>(hA, T{}), hyA = lds_at<HCB>(hA, T{}), hzA = lds_at<2 * HCB>(hA, T{});
            const T hxynA = lds_at<0>(hA + h_dyb, T{}), hzynA = lds_at<2 * HCB>(hA + h_dyb, T{});
            const T hyxnA = lds_at<HCB>(hA + h_dxb, T{}), hzxnA = lds_at<2 * HCB>(hA + h_dxb, T{});
            {   // H slot 2t
                const T b0 = mul(czB, sub(xC, xB)), b1 = mul(czB, sub(yC, yB));
                const T b2 = mul(ca.cys, sub(xynB, xB)), b3 = mul(ca.cys, sub(zynB, zB));
                const T b4 = mul(ca.cxs, sub(yxnB, yB)), b5 = mul(ca.cxs, sub(zxnB, zB));
                sts_at<HSTAGE>(out, add(add(Z, neg(b1)), b3));
                sts_at<HSTAGE + HCB>(out, add(b0, neg(b5)));
                sts_at<HSTAGE + 2 * HCB>(out, add(neg(b2), b4));
            }
            axlo = xC;
            aylo = yC;
            // H operands of output 2t−3
            const T hxB = lds_at<0>(hB, T{}), hyB = lds_at<HCB>(hB, T{}), hzB = lds_at<2 * HCB>(hB, T{});
            const T hxynB = lds_at<0>(hB + h_dyb, T{}), hzynB = lds_at<2 * HCB>(hB + h_dyb, T{});
            const T hyxnB = lds_at<HCB>(hB + h_dxb, T{}), hzxnB = lds_at<2 * HCB>(hB + h_dxb, T{});
            {   // output 2t−4
                const T a0 = mul(dzA, sub(hxA, bxlo)), a1 = mul(dzA, sub(hyA, bylo));
                const T a2 = mul(cb.cys, sub(hxynA, hxA)), a3 = mul(cb.cys, sub(hzynA, hzA));
                const T a4 = mul(cb.cxs, sub(hyxnA, hyA)), a5 = mul(cb.cxs, sub(hzxnA, hzA));
                stg_stream(p.G[0] + off_out, ADD ? add(add(gA0, neg(a1)), a3) : add(add(Z, neg(a1)), a3));
                stg_stream(p.G[1] + off_out, ADD ? add(add(gA1, a0), neg(a5)) : add(a0, neg(a5)));
                stg_stream(p.G[2] + off_out, ADD ? add(add(gA2, neg(a2)), a4) : add(neg(a2), a4));
            }
            {   // output 2t−3
                const T b0 = mul(dzB, sub(hxB, hxA)), b1 = mul(dzB, sub(hyB, hyA));
                const T b2 = mul(cb.cys, sub(hxynB, hxB)), b3 = mul(cb.cys, sub(hzynB, hzB));
                const T b4 = mul(cb.cxs, sub(hyxnB, hyB)), b5 = mul(cb.cxs, sub(hzxnB, hzB));
                stg_stream(p.G[0] + off_out + sz, ADD ? add(add(gB0, neg(b1)), b3) : add(add(Z, neg(b1)), b3));
                stg_stream(p.G[1] + off_out + sz, ADD ? add(add(gB1, b0), neg(b5)) : add(b0, neg(b5)));
                stg_stream(p.G[2] + off_out + sz, ADD ? add(add(gB2, neg(b2)), b4) : add(neg(b2), b4));
            }
            bxlo = hxB;
            bylo = hyB;
            off_out += 2 * sz;
            st_prev = st_cur;
            if (++st_cur == NP) { st_cur = 0; e_phase ^= 1u; }
            h_wr ^= 2u * HSTAGE;
            __syncthreads();
            continue;
        }
        if (t <= Th) {
            const int hkA = hb + 2 * t - 1;
            const int zfA = zflags_h(hkA), zfB = zflags_h(hkA + 1);
            const C czA = lds_at<0>(sC1 + (unsigned)(2 * t) * CS, C{}), czB = lds_at<CS>(sC1 + (unsigned)(2 * t) * CS, C{});
            mbar_wait(bar_efull + 8 * st_cur, e_phase);  // pair P_t has landed
            const unsigned eA = sE + st_prev * EPAIR + EPLANE, eB = sE + st_cur * EPAIR, eC = eB + EPLANE;
            h_pair(a_q, ca, axlo, aylo, a_out + h_wr, zfA, zfB, czA, czB, eA, eB, eC);
            if (roleB2) h_pair(b_q, cb, bxlo, bylo, b_out + h_wr, zfA, zfB, czA, czB, eA, eB, eC);
            st_prev = st_cur;
            if (++st_cur == NP) { st_cur = 0; e_phase ^= 1u; }
        }
        if (roleD) {
            const unsigned hA = sH + (h_wr ^ (2u * HSTAGE)) + b_q, hB = hA + HSTAGE;  // the pair written in iteration t−1
            if (t == 1) {
                bxlo = lds_at<0>(hB, T{});  // H slot 0: the lower plane of output 0
                bylo = lds_at<HCB>(hB, T{});
            } else if (t >= 2) {
                const int kk = 2 * t - 4, k = k0 + kk;
                const bool vA = valid && kk < n, vB = valid && kk + 1 < n;
                T gA0 = Z, gA1 = Z, gA2 = Z, gB0 = Z, gB1 = Z, gB2 = Z;
                if (ADD) {
                    if (vA) { gA0 = ldg_plain(p.G[0] + off_out); gA1 = ldg_plain(p.G[1] + off_out); gA2 = ldg_plain(p.G[2] + off_out); }
                    if (vB) { gB0 = ldg_plain(p.G[0] + off_out + sz); gB1 = ldg_plain(p.G[1] + off_out + sz); gB2 = ldg_plain(p.G[2] + off_out + sz); }
                }
                const T xA = lds_at<0>(hA, T{}), yA = lds_at<HCB>(hA, T{}), zA = lds_at<2 * HCB>(hA, T{});
                const T xynA = lds_at<0>(hA + h_dyb, T{}), zynA = lds_at<2 * HCB>(hA + h_dyb, T{});
                const T yxnA = lds_at<HCB>(hA + h_dxb, T{}), zxnA = lds_at<2 * HCB>(hA + h_dxb, T{});
                const T xB = lds_at<0>(hB, T{}), yB = lds_at<HCB>(hB, T{}), zB = lds_at<2 * HCB>(hB, T{});
                const T xynB = lds_at<0>(hB + h_dyb, T{}), zynB = lds_at<2 * HCB>(hB + h_dyb, T{});
                const T yxnB = lds_at<HCB>(hB + h_dxb, T{}), zxnB = lds_at<2 * HCB>(hB + h_dxb, T{});
                const C czA = lds_at<0>(sC2 + (unsigned)kk * CS, C{}), czB = lds_at<CS>(sC2 + (unsigned)kk * CS, C{});
                const int zfA = zflags_d(k), zfB = zflags_d(k + 1);
                T a0, a1, a2, a3, a4, a5, b0, b1, b2, b3, b4, b5;
                if ((cb.flags | zfA | zfB) == 0) {
                    a0 = mul(czA, sub(xA, bxlo)); a1 = mul(czA, sub(yA, bylo));
                    a2 = mul(cb.cys, sub(xynA, xA)); a3 = mul(cb.cys, sub(zynA, zA));
                    a4 = mul(cb.cxs, sub(yxnA, yA)); a5 = mul(cb.cxs, sub(zxnA, zA));
                    b0 = mul(czB, sub(xB, xA)); b1 = mul(czB, sub(yB, yA));
                    b2 = mul(cb.cys, sub(xynB, xB)); b3 = mul(cb.cys, sub(zynB, zB));
                    b4 = mul(cb.cxs, sub(yxnB, yB)); b5 = mul(cb.cxs, sub(zxnB, zB));
                } else {
                    curl_point_sym<T, C, false>(cb.flags | zfA, cb.cxs, cb.cys, fx2, fy2, czA, bxlo, xA, bylo, yA, zA, xynA, zynA, yxnA, zxnA, a0, a1, a2, a3,
                                                a4, a5);
                    curl_point_sym<T, C, false>(cb.flags | zfB, cb.cxs, cb.cys, fx2, fy2, czB, xA, xB, yA, yB, zB, xynB, zynB, yxnB, zxnB, b0, b1, b2, b3,
                                                b4, b5);
                }
                if (vA) {
                    stg_stream(p.G[0] + off_out, ADD ? add(add(gA0, neg(a1)), a3) : add(add(Z, neg(a1)), a3));
                    stg_stream(p.G[1] + off_out, ADD ? add(add(gA1, a0), neg(a5)) : add(a0, neg(a5)));
                    stg_stream(p.G[2] + off_out, ADD ? add(add(gA2, neg(a2)), a4) : add(neg(a2), a4));
                }
                if (vB) {
                    stg_stream(p.G[0] + off_out + sz, ADD ? add(add(gB0, neg(b1)), b3) : add(add(Z, neg(b1)), b3));
                    stg_stream(p.G[1] + off_out + sz, ADD ? add(add(gB1, b0), neg(b5)) : add(b0, neg(b5)));
                    stg_stream(p.G[2] + off_out + sz, ADD ? add(add(gB2, neg(b2)), b4) : add(neg(b2), b4));
                }
                bxlo = xB;
                bylo = yB;
                off_out += 2 * sz;
            }
        }
        h_wr ^= 2u * HSTAGE;
        if (t <= Th) __syncthreads();  // the H pair of iteration t is complete; E pair P_{t−1} is free
    }
}

}  // namespace sgc
