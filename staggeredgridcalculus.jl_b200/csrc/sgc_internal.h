// sgc_internal.h — declarations shared by the translation units of libsgc_b200.so.
#pragma once
#include "../../include/sgc_b200.h"

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "sgc_common.cuh"

int sgc_fail(int code, const char *fmt, ...);
int sgc_check_launch(const char *what);
void sgc_count_launches(int n);

#define CUDA_TRY(expr)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess)                                                                  \
            return sgc_fail(SGC_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                            __FILE__, __LINE__);                                                \
    } while (0)
#define REQUIRE(cond, ...)                                                                      \
    do {                                                                                        \
        if (!(cond)) return sgc_fail(SGC_ERR_INVALID, __VA_ARGS__);                             \
    } while (0)

// =============================================================================================
// host scalars with Julia's promotion rules
// =============================================================================================
struct HV {
    double re, im;
    bool cx;
};
static inline HV hv(double x) { return HV{x, 0.0, false}; }
static inline HV hvc(double re, double im) { return HV{re, im, true}; }
static inline HV hmul(HV a, HV b) {
    if (!a.cx && !b.cx) return HV{a.re * b.re, 0.0, false};
    if (!a.cx) return HV{a.re * b.re, a.re * b.im, true};   // *(x::Real, z::Complex)
    if (!b.cx) return HV{b.re * a.re, b.re * a.im, true};   // *(z::Complex, x::Real)
    return HV{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re, true};
}
static inline HV hadd(HV a, HV b) {
    if (!a.cx && !b.cx) return HV{a.re + b.re, 0.0, false};
    if (!a.cx) return HV{a.re + b.re, b.im, true};
    if (!b.cx) return HV{a.re + b.re, a.im, true};
    return HV{a.re + b.re, a.im + b.im, true};
}
static inline HV hneg(HV a) { return HV{-a.re, -a.im, a.cx}; }
static inline HV hcomplex(HV a) { return a.cx ? a : HV{a.re, 0.0, true}; }
static inline HV hinv(HV a) {  // e^(-1.0): inv(e)
    if (!a.cx) return HV{1.0 / a.re, 0.0, false};
    sgc::cplx r = sgc::cinv_host(sgc::cplx{a.re, a.im});
    return HV{r.re, r.im, true};
}
static inline HV vec_at(const void *v, int is_complex, int64_t i) {
    if (!v) return hv(1.0);
    const double *d = (const double *)v;
    return is_complex ? hvc(d[2 * i], d[2 * i + 1]) : hv(d[i]);
}
static inline sgc::cplx to_cplx(HV a) { return sgc::cplx{a.re, a.cx ? a.im : 0.0}; }


// ---- sparse assembly handle ------------------------------------------------------------------
struct AsmBlockHost {
    int nv, nu, axis, ns;
    std::vector<HV> V0, Vs;
};

struct sgc_asm {
    int ndim = 0;
    int64_t N[3] = {1, 1, 1};
    int64_t M = 1;
    int Kout = 1, Kin = 1;
    int cmpfirst = 1;
    bool cx = false;  // value type is ComplexF64
    std::vector<AsmBlockHost> blocks;
    int64_t row_begin = 0, row_end = 0;
    void *arena = nullptr;
    std::vector<size_t> off0, offs;
    void *scan_tmp = nullptr;
    size_t scan_tmp_bytes = 0;
    int64_t nnz = -1;
    int64_t *cta_total = nullptr, *cta_offset = nullptr;  // per-CTA entry counts and their exclusive scan
    int64_t ncta = 0;
    bool counted = false;
    int *count_tables = nullptr;              // device: A[nu][axis][N_axis] of the separable count
    size_t count_off[3][3] = {};              // element offsets into count_tables
    // host convenience
    int64_t *d_colptr = nullptr;
};

