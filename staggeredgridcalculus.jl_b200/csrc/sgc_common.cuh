// sgc_common.cuh — shared device/host helpers for libsgc_b200 (sm_100a only).
//
// Arithmetic contract: IEEE double, no FMA contraction (the library
// is compiled with -fmad=false and the host side with -ffp-contract=off), complex multiply
// in the 4-multiply/2-add form, complex division by Julia Base's scaled robust algorithm.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define SGC_HD __host__ __device__ __forceinline__
#define SGC_D __device__ __forceinline__

namespace sgc {

// ComplexF64, 16-byte aligned so that one element is one 128-bit memory transaction.
struct __align__(16) cplx {
    double re, im;
};

// ---- element arithmetic, overloaded for double and cplx ------------------------------------
SGC_HD double zero_of(double) { return 0.0; }
SGC_HD cplx zero_of(cplx) { return cplx{0.0, 0.0}; }

SGC_HD double add(double a, double b) { return a + b; }
SGC_HD double sub(double a, double b) { return a - b; }
SGC_HD double neg(double a) { return -a; }
SGC_HD cplx add(cplx a, cplx b) { return cplx{a.re + b.re, a.im + b.im}; }
SGC_HD cplx sub(cplx a, cplx b) { return cplx{a.re - b.re, a.im - b.im}; }
SGC_HD cplx neg(cplx a) { return cplx{-a.re, -a.im}; }

// coefficient * value.  Real*Complex scales both parts (Julia *(x::Real, z::Complex)).
SGC_HD double mul(double c, double v) { return c * v; }
SGC_HD cplx mul(double c, cplx v) { return cplx{c * v.re, c * v.im}; }
SGC_HD cplx mul(cplx c, cplx v) {
    return cplx{c.re * v.re - c.im * v.im, c.re * v.im + c.im * v.re};
}
SGC_HD cplx mul(cplx c, double v) { return cplx{v * c.re, v * c.im}; }

// Julia Base robust_cdiv2 / robust_cdiv1 / `/(::ComplexF64, ::ComplexF64)` (complex.jl).
SGC_HD double robust_cdiv2(double a, double b, double c, double d, double r, double t) {
    if (r != 0.0) {
        double br = b * r;
        return (br != 0.0) ? (a + br) * t : a * t + (b * t) * r;
    }
    return (a + d * (b / c)) * t;
}
SGC_HD void robust_cdiv1(double a, double b, double c, double d, double &p, double &q) {
    double r = d / c;
    double t = 1.0 / (c + d * r);
    p = robust_cdiv2(a, b, c, d, r, t);
    q = robust_cdiv2(b, -a, c, d, r, t);
}
__host__ __device__ __noinline__ inline cplx cdiv(cplx z, cplx w) {
    double a = z.re, b = z.im, c = w.re, d = w.im;
    const double half = 0.5, two = 2.0;
    double ab = fmax(fabs(a), fabs(b));
    double cd = fmax(fabs(c), fabs(d));
    const double ov = 1.7976931348623157e308;  // floatmax(Float64)
    const double un = 2.2250738585072014e-308;  // floatmin(Float64)
    const double eps = 2.220446049250313e-16;
    const double bs = two / (eps * eps);
    double s = 1.0;
    if (ab >= half * ov) { a = half * a; b = half * b; s = two * s; }
    if (cd >= half * ov) { c = half * c; d = half * d; s = s * half; }
    if (ab <= un * two / eps) { a = a * bs; b = b * bs; s = s / bs; }
    if (cd <= un * two / eps) { c = c * bs; d = d * bs; s = s * bs; }
    double p, q;
    if (fabs(d) <= fabs(c)) {
        robust_cdiv1(a, b, c, d, p, q);
    } else {
        robust_cdiv1(b, a, d, c, p, q);
        q = -q;
    }
    return cplx{p * s, q * s};
}
// value / e  with e carried as a complex number and a flag saying whether it is a Julia Real.
SGC_HD double div_e(double v, cplx e, bool) { return v / e.re; }
SGC_HD cplx div_e(cplx v, cplx e, bool e_is_real) {
    if (e_is_real) return cplx{v.re / e.re, v.im / e.re};  // /(z::Complex, x::Real)
    return cdiv(v, e);
}
// e * value
SGC_HD double mul_e(cplx e, double v, bool) { return e.re * v; }
SGC_HD cplx mul_e(cplx e, cplx v, bool e_is_real) {
    if (e_is_real) return cplx{e.re * v.re, e.re * v.im};
    return mul(e, v);
}

// Julia inv(::ComplexF64) (robust_cinv with scaling) — host only, used for e^(-1.0).
inline cplx cinv_host(cplx w) {
    double c = w.re, d = w.im;
    double absc = fabs(c), absd = fabs(d);
    double cd = absc > absd ? absc : absd;
    const double eps = 2.220446049250313e-16;
    const double bs = 2.0 / (eps * eps);
    double s = 1.0;
    if (cd >= 1.7976931348623157e308 / 2) { c *= 0.5; d *= 0.5; s = 0.5; }
    else if (cd <= 2 * 2.2250738585072014e-308 / eps) { c *= bs; d *= bs; s = bs; }
    double p, q;
    if (absd <= absc) {
        double r = d / c;
        p = 1.0 / (d * r + c);
        q = -r * p;
    } else {
        double cc = -d, dd = -c;
        double r = dd / cc;
        double pp = 1.0 / (dd * r + cc);
        double qq = -r * pp;
        q = pp; p = qq;
    }
    return cplx{p * s, q * s};
}

// Julia inv(::ComplexF64) for device code (the same scaled robust algorithm as cinv_host, multiply-add not fused).
SGC_HD cplx cinv(cplx w) {
    double c = w.re, d = w.im;
    const double absc = fabs(c), absd = fabs(d);
    const double cd = absc > absd ? absc : absd;
    const double eps = 2.220446049250313e-16;
    const double bs = 2.0 / (eps * eps);
    double s = 1.0;
    if (cd >= 1.7976931348623157e308 / 2) { c *= 0.5; d *= 0.5; s = 0.5; }
    else if (cd <= 2 * 2.2250738585072014e-308 / eps) { c *= bs; d *= bs; s = bs; }
    double p, q;
    if (absd <= absc) {
        const double r = d / c;
        p = 1.0 / (d * r + c);
        q = -r * p;
    } else {
        const double cc = -d, dd = -c;
        const double r = dd / cc;
        const double pp = 1.0 / (dd * r + cc);
        q = pp; p = -r * pp;
    }
    return cplx{p * s, q * s};
}

// ---- streaming memory access ------------------------------------------------------------------
// Fields are read once per kernel and written once: keep them out of L1 (no_allocate) and let
// the 126 MB L2 provide the halo reuse between neighbouring CTAs.
SGC_D cplx ldg_stream(const cplx *p) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return cplx{v.x, v.y};
}
SGC_D double ldg_stream(const double *p) {
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
// plain (coherent) load: used for G when accumulating and for peer/ghost planes
SGC_D cplx ldg_plain(const cplx *p) {
    double2 v = *reinterpret_cast<const double2 *>(p);
    return cplx{v.x, v.y};
}
SGC_D double ldg_plain(const double *p) { return *p; }
SGC_D void stg_stream(cplx *p, cplx v) {
    asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(v.re), "d"(v.im) : "memory");
}
SGC_D void stg_stream(double *p, double v) {
    asm volatile("st.global.L1::no_allocate.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}

// warp shuffles of one element
SGC_D double shfl_down1(double v) { return __shfl_down_sync(0xffffffffu, v, 1); }
SGC_D double shfl_up1(double v) { return __shfl_up_sync(0xffffffffu, v, 1); }
SGC_D cplx shfl_down1(cplx v) {
    return cplx{__shfl_down_sync(0xffffffffu, v.re, 1), __shfl_down_sync(0xffffffffu, v.im, 1)};
}
SGC_D cplx shfl_up1(cplx v) {
    return cplx{__shfl_up_sync(0xffffffffu, v.re, 1), __shfl_up_sync(0xffffffffu, v.im, 1)};
}

}  // namespace sgc
