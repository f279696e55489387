// sgc_cc4.cuh — fused curl-of-curl, fourth generation (EXPERIMENT, opt-in): TMA-fed, warp-specialised,
// mbarrier-pipelined (sm_100a).  Selected with kernel version 4 (sgc_op_set_tuning); not the default.
//
// Same mathematics as k_curlcurl3w (sgc_curlcurl.cuh): G OP C₂(C₁ F) with H = C₁ F living only in shared memory,
// every value produced by the arithmetic of two apply_curl! calls (matrixfree/differential/curl.jl:52-137), so the
// result is bit-identical to the two-pass path (tests/test_gpu_matrixfree.py).  The ncu source page of the third
// generation (profiles/r2_cc3w_source_summary.txt) showed 497 thread-instructions per cell of which 132 are the FP64
// arithmetic, issue slots 46 % busy at 16 warps/SM behind one __syncthreads per plane, and the LDGSTS (cp.async)
// copies costing 15 shared-memory wavefronts per warp request instead of 4.  This kernel attacks exactly that:
//
//   * one PRODUCER warp feeds an SE-stage ring of E planes with `cp.async.bulk.tensor` (TMA): one elected lane issues
//     three 3-D box copies per plane — (TX+2)×(TY+2) elements of Ex, Ey, Ez — that land in shared memory as whole
//     lines and complete on an mbarrier (`complete_tx`); out-of-range box cells are zero-filled by the copy engine,
//     which is exactly what a symmetric (PEC/PMC) boundary discards anyway.  No thread computes a load address.
//   * TY+2 H-WARPS turn E planes into H planes (one tile row each, the last one the extra tile column) and TY
//     D-WARPS turn H planes into output planes; the two groups are coupled only by mbarriers per ring stage
//     (E full/empty, H full/empty): no CTA-wide barrier at all.
//   * one CTA per SM (2·TY+3 warps).
//
// Measured (B200, 256³ ComplexF64, PEC box, profiles/r2_sweep_cc4*.jsonl): 0.484 ms at TY = 6 (15 warps, 128
// registers), 0.543 ms at TY = 8 (19 warps, 96 registers + spills) — against 0.472 ms for the third generation and
// 0.530 ms for two k_curl3p launches.  The ncu source page (profiles/r2_cc4_source_summary.txt) says why it does not
// win: H- and D-warps are released by the same barrier events, so all warps of an SM run their FP64 phase at the same
// time (the FP64 pipe is saturated for ≈ 30 % of a plane period and idle for the rest) and their barrier / shared-
// memory latency phases at the same time; 27 % of all stall samples sit on mbarrier try-waits.  Fetching the next
// plane's operands ahead of the arithmetic (software pipelining inside each warp) made it slower (0.67 ms: the
// doubled operand set spills), as did running the two curls per z-chunk so that H is re-read from L2
// (tools/time_l2_chunked.py: 0.57–0.97 ms).  Kept as a working, bit-identical TMA / mbarrier reference for the next
// attempt (two CTAs per SM, or H- and D-groups a half period apart); Bloch x / y boundaries and z-slab ghost planes
// are not handled here — the launcher routes those to the third generation.
#pragma once
#include <cuda.h>

#include "sgc_curlcurl.cuh"

namespace sgc {

template <typename C>
struct alignas(64) CurlCurl4Params {
    CUtensorMap tm[3];             // Ex, Ey, Ez as 3-D tensors of Float64: (2·Nx, Ny, Nz), box (2·34, TY+2, 1)
    CurlCurlParams<cplx, C> q;
};

// ---- mbarrier / TMA primitives (PTX ISA 8.x; shared::cta addresses are 32-bit) ------------------------------------
SGC_D void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
SGC_D void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
SGC_D void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
SGC_D void mbar_arrive_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
SGC_D bool mbar_try_wait(unsigned bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
SGC_D void mbar_wait(unsigned bar, unsigned parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
SGC_D void tma_load_3d(unsigned dst, const CUtensorMap *tm, unsigned bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
SGC_D void tma_prefetch_desc(const CUtensorMap *tm) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}
// ring position: stage index and the parity of the phase a waiter expects
struct RingPos {
    int stage;
    unsigned phase;
    template <int S>
    SGC_D void advance() {
        if (++stage == S) { stage = 0; phase ^= 1u; }
    }
};

// The six ∂ terms at one point: straight-line arithmetic when no boundary rule touches the point (flags == 0, the
// overwhelming majority), the general rules of curl_point (sgc_curlcurl.cuh) out of line otherwise, so that their
// register needs (selects, the robust complex division) stay out of the hot loop.
template <typename T>
struct SixTerms { T d0, d1, d2, d3, d4, d5; };
template <typename T, typename C, bool FZ>
__device__ __noinline__ SixTerms<T> curl_point_general(int flags, C cxs, C cys, bool fx, bool fy, C cz, const cplx *e, const int *e_real, T xlo,
                                                       T xhi, T ylo, T yhi, T zc, T x_yn, T z_yn, T y_xn, T z_xn) {
    SixTerms<T> r;
    curl_point<T, C, FZ>(flags, cxs, cys, fx, fy, cz, e, e_real, xlo, xhi, ylo, yhi, zc, x_yn, z_yn, y_xn, z_xn, r.d0, r.d1, r.d2, r.d3, r.d4, r.d5);
    return r;
}
template <typename T, typename C, bool FZ>
SGC_D void curl_point4(int flags, C cxs, C cys, bool fx, bool fy, C cz, const cplx *e, const int *e_real, T xlo, T xhi, T ylo, T yhi,
                       T zc, T x_yn, T z_yn, T y_xn, T z_xn, T &d0, T &d1, T &d2, T &d3, T &d4, T &d5) {
    if (flags == 0) {
        const T xc_y = FZ ? xlo : xhi, yc_x = FZ ? ylo : yhi;
        d0 = mul(cz, sub(xhi, xlo)); d1 = mul(cz, sub(yhi, ylo));
        d2 = mul(cys, sub(x_yn, xc_y)); d3 = mul(cys, sub(z_yn, zc));
        d4 = mul(cxs, sub(y_xn, yc_x)); d5 = mul(cxs, sub(z_xn, zc));
    } else {
        const SixTerms<T> r = curl_point_general<T, C, FZ>(flags, cxs, cys, fx, fy, cz, e, e_real, xlo, xhi, ylo, yhi, zc, x_yn, z_yn, y_xn, z_xn);
        d0 = r.d0; d1 = r.d1; d2 = r.d2; d3 = r.d3; d4 = r.d4; d5 = r.d5;
    }
}

// bytes of one E component tile in shared memory, rounded up to whole 128-byte lines; total dynamic shared memory
constexpr int cc4_ecb(int TY) { return (34 * (TY + 2) * 16 + 127) / 128 * 128; }
constexpr size_t cc4_smem_bytes(int TY, int SE, int SH) {
    return (size_t)SE * 3 * cc4_ecb(TY) + (size_t)SH * 3 * 33 * (TY + 1) * 16 + 8 * (3 * SE + 2 * SH) + 128;
}

// warps: 0 = producer, 1 … TY+2 = H (tile rows 0 … TY, then the extra column), TY+3 … 2·TY+2 = D (output rows)
template <typename C, bool ADD, int TY, int SE, int SH, bool FZ1, bool FZ2>
__global__ void __launch_bounds__(32 * (2 * TY + 3), 1) k_cc4(const __grid_constant__ CurlCurl4Params<C> P) {
    using T = cplx;
    static_assert(SE >= 3 && SH >= 3, "rings need at least 3 stages");
    const CurlCurlParams<T, C> &p = P.q;
    constexpr int TX = 32;
    constexpr int HX = TX + 1, HY = TY + 1, EX = TX + 2, EY = TY + 2;
    constexpr int ES = 16;
    constexpr int ECOMP = EX * EY;
    constexpr int ECB = cc4_ecb(TY), HCB = HX * HY * ES;  // (E component tiles start on 128-byte lines: TMA destinations)
    constexpr int ESTAGE = 3 * ECB, HSTAGE = 3 * HCB;
    constexpr int W_H = HY + 1, W_D = TY;
    static_assert(HY <= 32, "the extra column is handled by one warp");
    extern __shared__ __align__(128) unsigned char sgc_smem_raw[];
    const unsigned sE = ((unsigned)__cvta_generic_to_shared(sgc_smem_raw) + 127u) & ~127u;
    const unsigned sH = sE + SE * ESTAGE;
    const unsigned bars = sH + SH * HSTAGE;
    const unsigned bar_efull = bars, bar_eempty = bars + 8 * SE;
    const unsigned bar_hfull = bars + 16 * SE, bar_hempty = bar_hfull + 8 * SH;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
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

    const int k0 = p.k_begin + blockIdx.z * p.march;
    const int k1 = min(k0 + p.march, p.k_end);
    if (k0 >= k1) return;
    const int n = k1 - k0;
    const int hb = k0 + min(0, s2z);
    const int eb = hb + min(0, s1z);
    const int nslots = n + 2;  // E slots 0 … n+1 ↔ planes eb … eb+n+1

    if (threadIdx.x == 0) {
        for (int s = 0; s < SE; ++s) {
            mbar_init(bar_efull + 8 * s, 1u);
            mbar_init(bar_eempty + 8 * s, W_H);
        }
        for (int s = 0; s < SH; ++s) {
            mbar_init(bar_hfull + 8 * s, W_H);
            mbar_init(bar_hempty + 8 * s, W_D);
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == 0) {
        // ======================================= PRODUCER ===========================================================
        if (lane == 0) { tma_prefetch_desc(&P.tm[0]); tma_prefetch_desc(&P.tm[1]); tma_prefetch_desc(&P.tm[2]); }
        if (lane == 0) {
            RingPos issue{0, 0};
            for (int j = 0; j < nslots; ++j) {
                const int kp = wrap_index(eb + j, Nz);  // (a plane beyond a Bloch z face wraps around; beyond a symmetric face it is never used)
                mbar_wait(bar_eempty + 8 * issue.stage, issue.phase ^ 1u);  // the H-warps are done with this stage
                const unsigned st = sE + issue.stage * ESTAGE;
                const unsigned bar = bar_efull + 8 * issue.stage;
                mbar_arrive_expect_tx(bar, 3 * ECOMP * ES);
                tma_load_3d(st, &P.tm[0], bar, 2 * ex0, ey0, kp);
                tma_load_3d(st + ECB, &P.tm[1], bar, 2 * ex0, ey0, kp);
                tma_load_3d(st + 2 * ECB, &P.tm[2], bar, 2 * ex0, ey0, kp);
                issue.advance<SE>();
            }
        }
        return;
    }

    if (warp <= W_H) {
        // ======================================= H-WARPS ============================================================
        const int hw = warp - 1;                 // 0 … HY−1: tile row; HY: the extra column (lanes = rows)
        const bool extra = (hw == HY);
        const int row = extra ? min(lane, HY - 1) : hw, col = extra ? TX : lane;
        const bool active = !extra || lane < HY;
        const int ox1 = s1x < 0 ? 1 : 0, oy1 = s1y < 0 ? 1 : 0;
        const PosConst<C> ca = pos_const<C>(p.fwd1, p.bc1_lo, p.bc1_hi, p.c1, wrap_index(hx0 + col, Nx), wrap_index(hy0 + row, Ny), Nx, Ny);
        const unsigned a_q = (unsigned)(((row + oy1) * EX + (col + ox1)) * ES);
        const unsigned a_out = sH + (unsigned)((row * HX + col) * ES);
        const int e_dxb = s1x * ES, e_dyb = s1y * EX * ES;
        RingPos e_lo{0, 0}, h{0, 0};
        mbar_wait(bar_efull, 0);  // E slot 0
        T xlo = lds_at<0>(sE + a_q, T{}), ylo = lds_at<ECB>(sE + a_q, T{});
        int hk = hb;
        for (int m = 0; m <= n; ++m, ++hk) {
            RingPos e_hi = e_lo;
            e_hi.advance<SE>();
            // z coefficient and boundary flags of H plane hk (uniform)
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
            mbar_wait(bar_efull + 8 * e_hi.stage, e_hi.phase);          // E slot m + 1 has landed
            mbar_wait(bar_hempty + 8 * h.stage, h.phase ^ 1u);          // the D-warps are done with this H stage
            const unsigned alo = sE + e_lo.stage * ESTAGE + a_q, ahi = sE + e_hi.stage * ESTAGE + a_q;
            const unsigned ac = FZ1 ? alo : ahi;
            const T xhi = lds_at<0>(ahi, T{}), yhi = lds_at<ECB>(ahi, T{});
            const T zc = lds_at<2 * ECB>(ac, T{});
            const T x_yn = lds_at<0>(ac + e_dyb, T{}), z_yn = lds_at<2 * ECB>(ac + e_dyb, T{});
            const T y_xn = lds_at<ECB>(ac + e_dxb, T{}), z_xn = lds_at<2 * ECB>(ac + e_dxb, T{});
            T d0, d1, d2, d3, d4, d5;
            curl_point4<T, C, FZ1>(ca.flags | zflags, ca.cxs, ca.cys, fx1, fy1, cz, p.e1, p.e1_real, xlo, xhi, ylo, yhi, zc, x_yn, z_yn,
                                  y_xn, z_xn, d0, d1, d2, d3, d4, d5);
            if (active) {
                const unsigned out = a_out + h.stage * HSTAGE;
                sts_at<0>(out, add(add(Z, neg(d1)), d3));  // Gv .= 0 ; += (−∂zFy) ; += ∂yFz   (curl.jl:66-91)
                sts_at<HCB>(out, add(d0, neg(d5)));
                sts_at<2 * HCB>(out, add(neg(d2), d4));
            }
            xlo = xhi;
            ylo = yhi;
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(bar_hfull + 8 * h.stage);       // H plane m: this warp's part is in shared memory
                mbar_arrive(bar_eempty + 8 * e_lo.stage);   // E slot m is no longer needed by this warp
            }
            e_lo = e_hi;
            h.advance<SH>();
        }
        return;
    }

    {
        // ======================================= D-WARPS ============================================================
        const int dw = warp - 1 - W_H;  // output row
        const int i = i0 + lane, j = j0 + dw;
        const bool valid = (i < Nx) && (j < Ny);
        const unsigned b_q = (unsigned)(((dw + (s2y < 0 ? 1 : 0)) * HX + (lane + (s2x < 0 ? 1 : 0))) * ES);
        const PosConst<C> cb = pos_const<C>(p.fwd2, p.bc2_lo, p.bc2_hi, p.c2, valid ? i : 0, valid ? j : 0, Nx, Ny);
        const int h_dxb = s2x * ES, h_dyb = s2y * HX * ES;
        RingPos h_lo{0, 0};
        mbar_wait(bar_hfull, 0);  // H slot 0
        T xlo = lds_at<0>(sH + b_q, T{}), ylo = lds_at<HCB>(sH + b_q, T{});
        int64_t off_out = (int64_t)k0 * sz + (int64_t)j * sy + i;
        for (int kk = 0; kk < n; ++kk, off_out += sz) {
            const int k = k0 + kk;
            RingPos h_hi = h_lo;
            h_hi.advance<SH>();
            T g0 = Z, g1 = Z, g2 = Z;
            if (ADD && valid) { g0 = ldg_plain(p.G[0] + off_out); g1 = ldg_plain(p.G[1] + off_out); g2 = ldg_plain(p.G[2] + off_out); }
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
            mbar_wait(bar_hfull + 8 * h_hi.stage, h_hi.phase);  // H slot kk + 1 is complete
            const unsigned alo = sH + h_lo.stage * HSTAGE + b_q, ahi = sH + h_hi.stage * HSTAGE + b_q;
            const unsigned ac = FZ2 ? alo : ahi;
            const T xhi = lds_at<0>(ahi, T{}), yhi = lds_at<HCB>(ahi, T{});
            const T zc = lds_at<2 * HCB>(ac, T{});
            const T x_yn = lds_at<0>(ac + h_dyb, T{}), z_yn = lds_at<2 * HCB>(ac + h_dyb, T{});
            const T y_xn = lds_at<HCB>(ac + h_dxb, T{}), z_xn = lds_at<2 * HCB>(ac + h_dxb, T{});
            T d0, d1, d2, d3, d4, d5;
            curl_point4<T, C, FZ2>(cb.flags | zflags, cb.cxs, cb.cys, fx2, fy2, cz, p.e2, p.e2_real, xlo, xhi, ylo, yhi, zc, x_yn, z_yn,
                                  y_xn, z_xn, d0, d1, d2, d3, d4, d5);
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_hempty + 8 * h_lo.stage);  // H slot kk has been read by this warp
            T gx, gy, gz;
            if (ADD) { gx = add(add(g0, neg(d1)), d3); gy = add(add(g1, d0), neg(d5)); gz = add(add(g2, neg(d2)), d4); }
            else     { gx = add(add(Z, neg(d1)), d3);  gy = add(d0, neg(d5));          gz = add(neg(d2), d4); }
            if (valid) {
                stg_stream(p.G[0] + off_out, gx);
                stg_stream(p.G[1] + off_out, gy);
                stg_stream(p.G[2] + off_out, gz);
            }
            xlo = xhi;
            ylo = yhi;
            h_lo = h_hi;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// k_cc5 — fifth generation: the thread mapping of the third (every warp owns one H row AND one output row, two CTAs per
// SM, one __syncthreads per plane), fed by TMA like the fourth, with the lean loops of the fourth.  What the profiles
// said: v3 spends 365 of its 497 thread-instructions per cell on anything but FP64 — per-thread address arithmetic and
// issue logic of the cp.async copies, branches around the boundary rules — and its copies cost 15 shared-memory
// wavefronts per request; v4 removed both but gave up the two independent CTAs per SM that keep an SM busy while one
// tile sits in its barrier.  Here one elected thread issues three box copies per plane (no address arithmetic anywhere
// else), completion is an mbarrier try-wait right before first use, the boundary rules are out of line
// (curl_point4), and the per-plane __syncthreads only orders the H ring.
// Same applicability as k_cc4: ComplexF64, symmetric x / y boundaries (out-of-range box cells are zero-filled, which is
// what those rules discard), no slab ghost planes.
// ---------------------------------------------------------------------------------------------------------------------
constexpr size_t cc5_smem_bytes(int TY, int SE) {
    return (size_t)SE * 3 * cc4_ecb(TY) + (size_t)3 * 3 * 33 * (TY + 1) * 16 + 8 * SE + 128;
}

template <typename C, bool ADD, int TY, int SE, bool FZ1, bool FZ2>
__global__ void __launch_bounds__(32 * (TY + 1), 2) k_cc5(const __grid_constant__ CurlCurl4Params<C> P) {
    using T = cplx;
    static_assert(SE >= 3, "need at least 3 stages");
    const CurlCurlParams<T, C> &p = P.q;
    constexpr int TX = 32;
    constexpr int HX = TX + 1, HY = TY + 1, EX = TX + 2, EY = TY + 2;
    constexpr int ES = 16;
    constexpr int ECOMP = EX * EY;
    constexpr int ECB = cc4_ecb(TY), HCB = HX * HY * ES;
    constexpr int ESTAGE = 3 * ECB, HSTAGE = 3 * HCB;
    static_assert(HY <= 32, "the extra column is handled by one warp");
    extern __shared__ __align__(128) unsigned char sgc_smem_raw[];
    const unsigned sE = ((unsigned)__cvta_generic_to_shared(sgc_smem_raw) + 127u) & ~127u;
    const unsigned sH = sE + SE * ESTAGE;
    const unsigned bar_efull = sH + 3 * HSTAGE;

    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int Nx = p.Nx, Ny = p.Ny, Nz = p.Nz;
    const int64_t sy = Nx, sz = (int64_t)Nx * Ny;
    const bool fx1 = p.fwd1[0], fy1 = p.fwd1[1], fx2 = p.fwd2[0], fy2 = p.fwd2[1];
    const int s1x = fx1 ? 1 : -1, s1y = fy1 ? 1 : -1, s2x = fx2 ? 1 : -1, s2y = fy2 ? 1 : -1;
    constexpr int s1z = FZ1 ? 1 : -1, s2z = FZ2 ? 1 : -1;
    const int i0 = blockIdx.x * TX, j0 = blockIdx.y * TY;
    const int hx0 = i0 + min(0, s2x), hy0 = j0 + min(0, s2y);
    const int ex0 = hx0 + min(0, s1x), ey0 = hy0 + min(0, s1y);
    const T Z = zero_of(T{});
    const int ox1 = s1x < 0 ? 1 : 0, oy1 = s1y < 0 ? 1 : 0;

    const int k0 = p.k_begin + blockIdx.z * p.march;
    const int k1 = min(k0 + p.march, p.k_end);
    if (k0 >= k1) return;
    const int n = k1 - k0;
    const int hb = k0 + min(0, s2z);
    const int eb = hb + min(0, s1z);
    const int nslots = n + 2;

    if (threadIdx.x == 0) {
        for (int s = 0; s < SE; ++s) mbar_init(bar_efull + 8 * s, 1u);
        mbar_fence_init();
    }
    __syncthreads();

    // ---- producer duty of thread 0: E slot j → ring stage j mod SE ------------------------------------------------
    int issue_slot = 0, issue_stage = 0;
    auto issue = [&]() {
        if (issue_slot < nslots) {
            const int kp = wrap_index(eb + issue_slot, Nz);
            const unsigned st = sE + issue_stage * ESTAGE, bar = bar_efull + 8 * issue_stage;
            mbar_arrive_expect_tx(bar, 3 * ECOMP * ES);
            tma_load_3d(st, &P.tm[0], bar, 2 * ex0, ey0, kp);
            tma_load_3d(st + ECB, &P.tm[1], bar, 2 * ex0, ey0, kp);
            tma_load_3d(st + 2 * ECB, &P.tm[2], bar, 2 * ex0, ey0, kp);
        }
        ++issue_slot;
        issue_stage = (issue_stage + 1 == SE) ? 0 : issue_stage + 1;
    };
    if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < SE; ++q) issue();
    }

    // ---- role A: H position (row w, column lane) ---------------------------------------------------------------------
    const PosConst<C> ca = pos_const<C>(p.fwd1, p.bc1_lo, p.bc1_hi, p.c1, wrap_index(hx0 + lane, Nx), wrap_index(hy0 + w, Ny), Nx, Ny);
    const unsigned a_q = (unsigned)(((w + oy1) * EX + (lane + ox1)) * ES);
    const unsigned a_out = sH + (unsigned)((w * HX + lane) * ES);
    const int e_dxb = s1x * ES, e_dyb = s1y * EX * ES;
    // ---- second role: D (warps < TY) or B′ (warp TY, lanes < HY) -----------------------------------------------------
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

    RingPos e_lo{0, 0};
    mbar_wait(bar_efull, 0);  // E slot 0
    T axlo = lds_at<0>(sE + a_q, T{}), aylo = lds_at<ECB>(sE + a_q, T{});
    T bxlo = Z, bylo = Z;
    if (!roleD) { bxlo = lds_at<0>(sE + b_q, T{}); bylo = lds_at<ECB>(sE + b_q, T{}); }
    unsigned h_cur = 0;
    int64_t off_out = (int64_t)k0 * sz + (int64_t)j * sy + i;

    auto h_point = [&](unsigned q, const PosConst<C> &c, T &xlo_c, T &ylo_c, unsigned out, int zflags, C cz, unsigned e_lo_a, unsigned e_hi_a) {
        const unsigned ahi = e_hi_a + q;
        const unsigned ac = (FZ1 ? e_lo_a : e_hi_a) + q;
        const T xhi = lds_at<0>(ahi, T{}), yhi = lds_at<ECB>(ahi, T{});
        const T zc = lds_at<2 * ECB>(ac, T{});
        const T x_yn = lds_at<0>(ac + e_dyb, T{}), z_yn = lds_at<2 * ECB>(ac + e_dyb, T{});
        const T y_xn = lds_at<ECB>(ac + e_dxb, T{}), z_xn = lds_at<2 * ECB>(ac + e_dxb, T{});
        T d0, d1, d2, d3, d4, d5;
        curl_point4<T, C, FZ1>(c.flags | zflags, c.cxs, c.cys, fx1, fy1, cz, p.e1, p.e1_real, xlo_c, xhi, ylo_c, yhi, zc, x_yn, z_yn, y_xn,
                               z_xn, d0, d1, d2, d3, d4, d5);
        sts_at<0>(out + h_cur, add(add(Z, neg(d1)), d3));  // Gv .= 0 ; += (−∂zFy) ; += ∂yFz   (curl.jl:66-91)
        sts_at<HCB>(out + h_cur, add(d0, neg(d5)));
        sts_at<2 * HCB>(out + h_cur, add(neg(d2), d4));
        xlo_c = xhi;
        ylo_c = yhi;
    };

    int hk = hb;
    for (int m = 0; m <= n + 1; ++m, ++hk) {
        if (m >= 1 && threadIdx.x == 0) issue();  // E slot m − 1 + SE → the stage slot m − 1 used (free since the last barrier)
        if (m <= n) {
            int zflags = 0;
            C cz;
            if (hk > 0 && hk < Nz - 1) {
                cz = p.c1[2][hk];
            } else {
                const int kk = wrap_index(hk, Nz);
                cz = p.c1[2][kk];
                if (FZ1) {
                    if (kk == Nz - 1) zflags = ((p.bc1_hi[2] == BC_BLOCH) ? 1 : (p.bc1_hi[2] == BC_SYM ? 2 : 0)) << 8;
                    if (kk == 0 && p.bc1_lo[2] == BC_SYM) zflags |= 1024;
                } else if (kk == 0) {
                    zflags = ((p.bc1_lo[2] == BC_BLOCH) ? 1 : (p.bc1_lo[2] == BC_SYM ? 2 : 0)) << 8;
                }
            }
            RingPos e_hi = e_lo;
            e_hi.advance<SE>();
            mbar_wait(bar_efull + 8 * e_hi.stage, e_hi.phase);  // E slot m + 1 has landed
            const unsigned e_lo_a = sE + e_lo.stage * ESTAGE, e_hi_a = sE + e_hi.stage * ESTAGE;
            h_point(a_q, ca, axlo, aylo, a_out, zflags, cz, e_lo_a, e_hi_a);
            if (roleB2) h_point(b_q, cb, bxlo, bylo, b_out, zflags, cz, e_lo_a, e_hi_a);
            e_lo = e_hi;
        }
        if (roleD && m >= 2) {
            const int k = k0 + m - 2;
            T g0 = Z, g1 = Z, g2 = Z;
            if (ADD && valid) { g0 = ldg_plain(p.G[0] + off_out); g1 = ldg_plain(p.G[1] + off_out); g2 = ldg_plain(p.G[2] + off_out); }
            const unsigned st_hi = (h_cur == 0) ? 2u * HSTAGE : h_cur - HSTAGE;   // slot m − 1
            const unsigned st_lo = (st_hi == 0) ? 2u * HSTAGE : st_hi - HSTAGE;  // slot m − 2
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
            curl_point4<T, C, FZ2>(cb.flags | zflags, cb.cxs, cb.cys, fx2, fy2, cz, p.e2, p.e2_real, bxlo, xhi, bylo, yhi, zc, x_yn, z_yn,
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
        __syncthreads();  // H plane m is complete; E slot m is free
    }
}

}  // namespace sgc
