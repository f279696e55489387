// sgc_lib.cu — libsgc_b200.so: C-ABI entry points (include/sgc_b200.h), the operator runtime
// (device-resident coefficient tables, launch planning) and the kernel dispatch.
//
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -lineinfo -O3 ...
//         (see staggeredgridcalculus.jl_b200/csrc/Makefile)
// There is deliberately no CPU implementation in this library.
#include "sgc_internal.h"

#include <algorithm>
#include <mutex>

#include "sgc_kernels.cuh"
#include "sgc_fused.cuh"

using namespace sgc;

// =============================================================================================
// errors, launch counter
// =============================================================================================
static thread_local std::string g_err;
static std::atomic<int64_t> g_launches{0};

int sgc_fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

int sgc_check_launch(const char *what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) cudaGetLastError();
    if (e != cudaSuccess) return sgc_fail(SGC_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
    return SGC_OK;
}
void sgc_count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

extern "C" int sgc_abi_version(void) { return SGC_ABI_VERSION; }
extern "C" int sgc_build_flags(void) {
#ifdef SGC_SWEEP_BUILD
    return 1;
#else
    return 0;
#endif
}
extern "C" const char *sgc_last_error(void) { return g_err.c_str(); }
extern "C" int64_t sgc_launch_count(void) { return g_launches.load(); }

extern "C" int sgc_device_count(int *count) {
    REQUIRE(count, "count is NULL");
    CUDA_TRY(cudaGetDeviceCount(count));
    return SGC_OK;
}
extern "C" int sgc_set_device(int device) {
    CUDA_TRY(cudaSetDevice(device));
    return SGC_OK;
}
extern "C" int sgc_malloc(void **dptr, size_t bytes) {
    REQUIRE(dptr, "dptr is NULL");
    CUDA_TRY(cudaMalloc(dptr, bytes ? bytes : 1));
    return SGC_OK;
}
extern "C" int sgc_free(void *dptr) {
    CUDA_TRY(cudaFree(dptr));
    return SGC_OK;
}
extern "C" int sgc_malloc_host(void **hptr, size_t bytes) {
    REQUIRE(hptr, "hptr is NULL");
    CUDA_TRY(cudaMallocHost(hptr, bytes ? bytes : 1));
    return SGC_OK;
}
// Host evaluation of the library's complex division (the same source the kernels compile): lets the CPU tests pin
// the restated Julia Base algorithm against published known-answer vectors.  n pairs (re, im) each.
extern "C" int sgc_debug_cdiv(const double *z, const double *w, double *out, int64_t n) {
    REQUIRE(z && w && out && n >= 0, "bad arguments");
    for (int64_t q = 0; q < n; ++q) {
        const cplx r = cdiv(cplx{z[2 * q], z[2 * q + 1]}, cplx{w[2 * q], w[2 * q + 1]});
        out[2 * q] = r.re;
        out[2 * q + 1] = r.im;
    }
    return SGC_OK;
}
extern "C" int sgc_debug_cinv(const double *w, double *out, int64_t n) {
    REQUIRE(w && out && n >= 0, "bad arguments");
    for (int64_t q = 0; q < n; ++q) {
        const cplx r = cinv_host(cplx{w[2 * q], w[2 * q + 1]});
        out[2 * q] = r.re;
        out[2 * q + 1] = r.im;
    }
    return SGC_OK;
}
extern "C" int sgc_malloc_host_wc(void **hptr, size_t bytes) {
    REQUIRE(hptr, "hptr is NULL");
    CUDA_TRY(cudaHostAlloc(hptr, bytes ? bytes : 1, cudaHostAllocWriteCombined));
    return SGC_OK;
}
extern "C" int sgc_free_host(void *hptr) {
    CUDA_TRY(cudaFreeHost(hptr));
    return SGC_OK;
}
extern "C" int sgc_memcpy_h2d(void *dst, const void *src, size_t bytes, void *stream) {
    CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return SGC_OK;
}
extern "C" int sgc_memcpy_d2h(void *dst, const void *src, size_t bytes, void *stream) {
    CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return SGC_OK;
}
extern "C" int sgc_memset_zero(void *dptr, size_t bytes, void *stream) {
    CUDA_TRY(cudaMemsetAsync(dptr, 0, bytes, (cudaStream_t)stream));
    return SGC_OK;
}
extern "C" int sgc_stream_create(void **stream) {
    REQUIRE(stream, "stream is NULL");
    cudaStream_t s;
    CUDA_TRY(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    *stream = (void *)s;
    return SGC_OK;
}
extern "C" int sgc_stream_destroy(void *stream) {
    CUDA_TRY(cudaStreamDestroy((cudaStream_t)stream));
    return SGC_OK;
}
extern "C" int sgc_stream_sync(void *stream) {
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    return SGC_OK;
}

// =============================================================================================
// matrix-free operator handle
// =============================================================================================
// Does the term list match one of the compiled one-launch compositions (sgc_fused.cuh)?
static void detect_fused_pattern(sgc_op *op) {
    op->fused_pat = -1;
    const size_t nt = op->terms.size();
    if (op->ndim < 2 || nt < 1 || nt > 3) return;
    // a single ∂ along y: the row-register kernel wins (5.8 vs 5.0 TB/s at 8192² Float64); along x and z
    // the single-term kernels k_term_x / k_term_s stay ahead (measured), so those keep their path
    if (nt == 1 && op->terms[0].axis != 1) return;
    for (const Term &t : op->terms)
        if (t.kind != KIND_PARTIAL || t.cc != op->terms[0].cc) return;
    static const PatDesc *const pats[PAT_COUNT] = {&Pat<0>::D, &Pat<1>::D, &Pat<2>::D, &Pat<3>::D, &Pat<4>::D, &Pat<5>::D,
                                                   &Pat<6>::D, &Pat<7>::D, &Pat<8>::D};
    for (int P = 0; P < PAT_COUNT; ++P) {
        const PatDesc &D = *pats[P];
        if ((size_t)D.nt != nt || D.nin != op->n_in || D.nout != op->n_out) continue;
        bool same = true;
        for (size_t q = 0; q < nt && same; ++q)
            same = op->terms[q].out == D.out[q] && op->terms[q].in == D.in[q] && op->terms[q].axis == D.axis[q];
        if (same) { op->fused_pat = P; return; }
    }
}

static int finish_op(sgc_op *op, OpBuilder &b) {
    detect_fused_pattern(op);
    size_t bytes = b.host.size() ? b.host.size() : 16;
    CUDA_TRY(cudaMalloc(&op->arena, bytes));
    if (b.host.size()) CUDA_TRY(cudaMemcpy(op->arena, b.host.data(), b.host.size(), cudaMemcpyHostToDevice));
    return SGC_OK;
}

int sgc_check_common(int dtype, int ndim, const int64_t *N) {
    REQUIRE(dtype == SGC_F64 || dtype == SGC_C128, "dtype must be SGC_F64 or SGC_C128");
    REQUIRE(ndim >= 1 && ndim <= 3, "ndim must be 1, 2 or 3");
    REQUIRE(N, "N is NULL");
    for (int d = 0; d < ndim; ++d) REQUIRE(N[d] >= 1, "N[%d] must be positive", d);
    return SGC_OK;
}

static void init_op(sgc_op *op, int dtype, int ndim, const int64_t *N) {
    op->dtype = dtype;
    op->ndim = ndim;
    op->M = 1;
    for (int d = 0; d < 3; ++d) {
        op->N[d] = d < ndim ? N[d] : 1;
        op->M *= op->N[d];
    }
}

// Build one ∂ / m̂ term.  `sign` is the curl parity (±1) multiplied into α exactly as
// `α=parity*α` (matrixfree/differential/curl.jl:89).
static int add_term(sgc_op *op, OpBuilder &b, int kind, int out, int in, int axis, int isfwd,
                    const void *v1, const void *v2, int coef_complex, int isbloch, const double *e,
                    HV alpha, int sign) {
    const int64_t Nw = op->N[axis];
    HV ev = (e && e[1] != 0.0) ? hvc(e[0], e[1]) : hv(e ? e[0] : 1.0);
    if (!isbloch) ev = hv(1.0);  // unused
    HV al = hmul(hv((double)sign), alpha);
    if (sign == 1) al = alpha;  // 1*α is exact; keep the bits (incl. signed zeros) of α
    bool cc = al.cx || (coef_complex && (v1 || v2));
    if (kind == KIND_MEAN_WEIGHTED && ev.cx) cc = true;
    if (op->dtype == SGC_F64) {
        if (cc || ev.cx)
            return fail(SGC_ERR_INEXACT, "complex α, e⁻ⁱᵏᴸ or ∆ with a Float64 field (InexactError)");
    }
    if (isfwd && !isbloch && Nw < 2)
        return fail(SGC_ERR_INVALID, "forward operator with symmetric boundary needs N[nw] >= 2 (reads Fu[2])");
    Term t{};
    t.out = out; t.in = in; t.kind = kind; t.axis = axis; t.fwd = isfwd ? 1 : 0; t.bloch = isbloch ? 1 : 0;
    t.e = to_cplx(ev); t.e_real = ev.cx ? 0 : 1; t.cc = cc;
    t.al = al;
    std::vector<HV> c(Nw), w;
    if (kind == KIND_PARTIAL) {
        for (int64_t i = 0; i < Nw; ++i) c[i] = hmul(al, vec_at(v1, coef_complex, i));  // (α * ∆w⁻¹[i])
        t.off_c = b.push(c, cc);
    } else if (kind == KIND_MEAN_ARITH) {
        t.a2 = to_cplx(hmul(hv(0.5), al));  // α2 = 0.5α
        t.a1 = to_cplx(al);
    } else {
        HV a2 = hmul(hv(0.5), al);
        w.resize(Nw);
        for (int64_t i = 0; i < Nw; ++i) {
            c[i] = hmul(a2, vec_at(v2, coef_complex, i));  // (α2 * ∆w′⁻¹[i])
            w[i] = vec_at(v1, coef_complex, i);
        }
        t.off_c = b.push(c, cc);
        t.off_w = b.push(w, cc);
        t.b0 = to_cplx(hmul(al, vec_at(v2, coef_complex, 0)));  // β = α * ∆w′⁻¹[1]
        t.w_hi = to_cplx(isbloch ? hmul(w[0], ev) : w[0]);      // ∆w[1]*e⁻ⁱᵏᴸ
        t.w_lo = to_cplx(w[Nw - 1]);
    }
    op->terms.push_back(t);
    return SGC_OK;
}

static HV alpha_of(const double *alpha) {
    if (!alpha) return hv(1.0);
    return alpha[1] != 0.0 ? hvc(alpha[0], alpha[1]) : hv(alpha[0]);
}

extern "C" int sgc_op_destroy(sgc_op *op) {
    if (!op) return SGC_OK;
    if (op->arena) cudaFree(op->arena);
    if (op->hp) sgc_hostpipe_destroy(op->hp);
    delete op;
    return SGC_OK;
}

#define OP_PROLOGUE()                                         \
    REQUIRE(op_out, "op is NULL");                            \
    *op_out = nullptr;                                        \
    { int _s = sgc_check_common(dtype, ndim, N); if (_s) return _s; } \
    sgc_op *op = new sgc_op();                                \
    init_op(op, dtype, ndim, N);                              \
    OpBuilder b;
#define OP_EPILOGUE()                                         \
    { int _s = finish_op(op, b); if (_s) { sgc_op_destroy(op); return _s; } } \
    *op_out = op;                                             \
    return SGC_OK;
#define OP_TRY(expr) { int _s = (expr); if (_s) { sgc_op_destroy(op); return _s; } }
#define OP_REQUIRE(cond, ...) if (!(cond)) { sgc_op_destroy(op); return fail(SGC_ERR_INVALID, __VA_ARGS__); }

extern "C" int sgc_op_create_partial(sgc_op **op_out, int dtype, int ndim, const int64_t *N, int nw,
                                     int isfwd, const void *dwinv, int coef_complex, int isbloch,
                                     const double *e, const double *alpha) {
    OP_PROLOGUE();
    OP_REQUIRE(nw >= 1 && nw <= ndim, "nw = %d out of range 1..%d", nw, ndim);  // @assert 1≤nw≤3
    op->n_out = op->n_in = 1;
    OP_TRY(add_term(op, b, KIND_PARTIAL, 0, 0, nw - 1, isfwd, dwinv, nullptr, coef_complex, isbloch, e,
                    alpha_of(alpha), 1));
    OP_EPILOGUE();
}

extern "C" int sgc_op_create_mhat(sgc_op **op_out, int dtype, int ndim, const int64_t *N, int nw,
                                  int isfwd, const void *dw, const void *dwpinv, int coef_complex,
                                  int isbloch, const double *e, const double *alpha) {
    OP_PROLOGUE();
    OP_REQUIRE(nw >= 1 && nw <= ndim, "nw = %d out of range 1..%d", nw, ndim);
    OP_REQUIRE((dw == nullptr) == (dwpinv == nullptr), "dw and dwpinv must both be given or both be NULL");
    op->n_out = op->n_in = 1;
    OP_TRY(add_term(op, b, dw ? KIND_MEAN_WEIGHTED : KIND_MEAN_ARITH, 0, 0, nw - 1, isfwd, dw, dwpinv,
                    coef_complex, isbloch, e, alpha_of(alpha), 1));
    OP_EPILOGUE();
}

static const int CURL_PARITY[3][3] = {{0, -1, 1}, {1, 0, -1}, {-1, 1, 0}};  // [nv][nu], curl.jl:46-48

extern "C" int sgc_op_create_curl(sgc_op **op_out, int dtype, int ndim, const int64_t *N,
                                  const int *isfwd, const void *const *dlinv, int coef_complex,
                                  const int *isbloch, const double *e, const int *cmp_shp, int Kout,
                                  const int *cmp_out, int Kin, const int *cmp_in, const double *alpha) {
    OP_PROLOGUE();
    OP_REQUIRE(isfwd && isbloch && cmp_shp && cmp_out && cmp_in, "NULL argument");
    OP_REQUIRE(Kout >= 1 && Kout <= 3 && Kin >= 1 && Kin <= 3, "Kout, Kin must be in 1..3");
    for (int d = 0; d < ndim; ++d) OP_REQUIRE(cmp_shp[d] >= 1 && cmp_shp[d] <= 3, "cmp_shp entries must be 1..3");
    for (int q = 0; q < Kout; ++q) OP_REQUIRE(cmp_out[q] >= 1 && cmp_out[q] <= 3, "cmp_out entries must be 1..3");
    for (int q = 0; q < Kin; ++q) OP_REQUIRE(cmp_in[q] >= 1 && cmp_in[q] <= 3, "cmp_in entries must be 1..3");
    op->n_out = Kout;
    op->n_in = Kin;
    HV al = alpha_of(alpha);
    for (int iv = 0; iv < Kout; ++iv) {
        const int nv = cmp_out[iv];
        for (int iu = 0; iu < Kin; ++iu) {
            const int nu = cmp_in[iu];
            const int parity = CURL_PARITY[nv - 1][nu - 1];
            if (parity == 0) {  // curl.jl:75-81
                Term t{};
                t.out = iv; t.in = iu; t.kind = -1;
                op->terms.push_back(t);
                continue;
            }
            const int nw = 6 - nv - nu;
            int cnt = 0, ind = -1;
            for (int d = 0; d < ndim; ++d)
                if (cmp_shp[d] == nw) { ++cnt; ind = d; }
            OP_REQUIRE(cnt == 1, "cmp_shp does not have one and only one nw = %d", nw);  // curl.jl:85
            OP_TRY(add_term(op, b, KIND_PARTIAL, iv, iu, ind, isfwd[ind], dlinv ? dlinv[ind] : nullptr, nullptr,
                            coef_complex, isbloch[ind], e ? e + 2 * ind : nullptr, al, parity));
        }
    }
    // the full 3-D curl runs fused
    bool full = ndim == 3 && Kout == 3 && Kin == 3;
    for (int d = 0; d < 3 && full; ++d) full = cmp_shp[d] == d + 1 && cmp_out[d] == d + 1 && cmp_in[d] == d + 1;
    if (full) {
        op->fused_curl = true;
        op->alpha = al;
        op->fused_cc = al.cx || coef_complex;
        for (int d = 0; d < 3; ++d) {
            std::vector<HV> c(op->N[d]);
            for (int64_t i = 0; i < op->N[d]; ++i) c[i] = hmul(al, vec_at(dlinv ? dlinv[d] : nullptr, coef_complex, i));
            op->fused_off[d] = b.push(c, op->fused_cc);
        }
    }
    OP_EPILOGUE();
}

extern "C" int sgc_op_create_divg(sgc_op **op_out, int dtype, int ndim, const int64_t *N, const int *isfwd,
                                  const void *const *dlinv, int coef_complex, const int *isbloch,
                                  const double *e, const double *alpha) {
    OP_PROLOGUE();
    OP_REQUIRE(isfwd && isbloch, "NULL argument");
    op->n_out = 1;
    op->n_in = ndim;
    for (int d = 0; d < ndim; ++d)  // divergence.jl:52-63
        OP_TRY(add_term(op, b, KIND_PARTIAL, 0, d, d, isfwd[d], dlinv ? dlinv[d] : nullptr, nullptr, coef_complex,
                        isbloch[d], e ? e + 2 * d : nullptr, alpha_of(alpha), 1));
    OP_EPILOGUE();
}

extern "C" int sgc_op_create_grad(sgc_op **op_out, int dtype, int ndim, const int64_t *N, const int *isfwd,
                                  const void *const *dlinv, int coef_complex, const int *isbloch,
                                  const double *e, const double *alpha) {
    OP_PROLOGUE();
    OP_REQUIRE(isfwd && isbloch, "NULL argument");
    op->n_out = ndim;
    op->n_in = 1;
    for (int d = 0; d < ndim; ++d)  // gradient.jl:51-56
        OP_TRY(add_term(op, b, KIND_PARTIAL, d, 0, d, isfwd[d], dlinv ? dlinv[d] : nullptr, nullptr, coef_complex,
                        isbloch[d], e ? e + 2 * d : nullptr, alpha_of(alpha), 1));
    OP_EPILOGUE();
}

extern "C" int sgc_op_create_mean(sgc_op **op_out, int dtype, int ndim, const int64_t *N, const int *isfwd,
                                  const void *const *dl, const void *const *dlpinv, int coef_complex,
                                  const int *isbloch, const double *e, const double *alpha) {
    OP_PROLOGUE();
    OP_REQUIRE(isfwd && isbloch, "NULL argument");
    OP_REQUIRE((dl == nullptr) == (dlpinv == nullptr), "dl and dlpinv must both be given or both be NULL");
    op->n_out = op->n_in = ndim;
    for (int d = 0; d < ndim; ++d)  // mean.jl:56-64, 80-88
        OP_TRY(add_term(op, b, dl ? KIND_MEAN_WEIGHTED : KIND_MEAN_ARITH, d, d, d, isfwd[d], dl ? dl[d] : nullptr,
                        dlpinv ? dlpinv[d] : nullptr, coef_complex, isbloch[d], e ? e + 2 * d : nullptr,
                        alpha_of(alpha), 1));
    if (ndim == 3) {  // one launch when the three terms share the coefficient type
        op->fused_mean = true;
        for (const Term &t : op->terms) op->fused_mean = op->fused_mean && (t.cc == op->terms[0].cc);
    }
    OP_EPILOGUE();
}

extern "C" int sgc_op_set_slab(sgc_op *op, int lo_interface, int hi_interface) {
    REQUIRE(op, "op is NULL");
    op->lo_iface = lo_interface ? 1 : 0;
    op->hi_iface = hi_interface ? 1 : 0;
    return SGC_OK;
}

extern "C" int sgc_op_set_slab_weights(sgc_op *op, const double *dw_below, const double *dw_above, int coef_complex) {
    REQUIRE(op, "op is NULL");
    op->slab_w[0] = dw_below ? vec_at(dw_below, coef_complex, 0) : hv(1.0);
    op->slab_w[1] = dw_above ? vec_at(dw_above, coef_complex, 0) : hv(1.0);
    if (op->dtype == SGC_F64 && (op->slab_w[0].cx || op->slab_w[1].cx))
        return fail(SGC_ERR_INEXACT, "complex ∆w with a Float64 field (InexactError)");
    op->has_slab_w = true;
    return SGC_OK;
}

extern "C" int sgc_op_set_tuning(sgc_op *op, int tile_x, int tile_y, int march_z, int variant) {
    REQUIRE(op, "op is NULL");
    op->tile_x = tile_x; op->tile_y = tile_y; op->march = march_z; op->variant = variant;
    return SGC_OK;
}

// ---------------------------------------------------------------------------------------------
// launches
// ---------------------------------------------------------------------------------------------
template <typename C> static inline C narrow(cplx v);
template <> inline double narrow<double>(cplx v) { return v.re; }
template <> inline cplx narrow<cplx>(cplx v) { return v; }

template <typename T, typename C>
static int launch_term(const sgc_op *op, const Term &t, T *G, const T *F, const T *ghost_user, bool add_mode,
                       int64_t k_begin, int64_t k_end, cudaStream_t st) {
    TermParams<T, C> p{};
    const int a = t.axis, last = op->ndim - 1;
    p.G = G; p.F = F;
    p.c = (const C *)((const unsigned char *)op->arena + t.off_c);
    p.w = (const C *)((const unsigned char *)op->arena + t.off_w);
    p.a2 = narrow<C>(t.a2); p.a1 = narrow<C>(t.a1); p.b0 = narrow<C>(t.b0);
    p.w_hi = narrow<C>(t.w_hi); p.w_lo = narrow<C>(t.w_lo);
    if (op->has_slab_w && t.kind == KIND_MEAN_WEIGHTED && t.axis == op->ndim - 1) {
        // the plane beyond the slab belongs to a neighbour rank: its ∆w replaces the local wrap-around entry;
        // at a physical Bloch face the phase stays folded in (∆w[1]·e, mean.jl:524-527), at an interface it does not
        HV ev = t.e_real ? hv(t.e.re) : hvc(t.e.re, t.e.im);
        const HV whi = (t.bloch && !op->hi_iface) ? hmul(op->slab_w[1], ev) : op->slab_w[1];
        p.w_hi = narrow<C>(to_cplx(whi));
        p.w_lo = narrow<C>(to_cplx(op->slab_w[0]));
    }
    p.e = t.e; p.e_real = t.e_real; p.fwd = t.fwd;
    p.inner = 1;
    for (int d = 0; d < a; ++d) p.inner *= op->N[d];
    p.Nw = op->N[a];
    p.outer = 1;
    for (int d = a + 1; d < 3; ++d) p.outer *= op->N[d];
    p.bc_lo = p.bc_hi = t.bloch ? BC_BLOCH : BC_SYM;
    if (a == last) {
        if (op->lo_iface) p.bc_lo = BC_IFACE;
        if (op->hi_iface) p.bc_hi = BC_IFACE;
        p.i_begin = k_begin; p.i_end = k_end; p.o_begin = 0; p.o_end = p.outer;  // outer == 1
    } else {
        const int64_t mid = p.outer / op->N[last];
        p.i_begin = 0; p.i_end = p.Nw; p.o_begin = k_begin * mid; p.o_end = k_end * mid;
    }
    if (a == last && ghost_user) {
        p.ghost = ghost_user;
        p.ghost_stride_o = 0;
    } else {
        p.ghost = t.fwd ? F : F + (p.Nw - 1) * p.inner;
        p.ghost_stride_o = p.Nw * p.inner;
    }
    if (p.i_end <= p.i_begin || p.o_end <= p.o_begin) return SGC_OK;

    // ---- vectorised fast path ------------------------------------------------------------------
    {
        constexpr int VBIG = 32 / (int)sizeof(T);    // elements in 32 bytes
        constexpr int VSMALL = 16 / (int)sizeof(T);  // elements in 16 bytes
        constexpr int R = 4;
        const bool xax = (p.inner == 1);
        // (the contiguous-axis kernel reads its ghost element with a scalar load)
        const bool aligned = ((uintptr_t)p.G % 16 == 0) && ((uintptr_t)p.F % 16 == 0) && (xax || (uintptr_t)p.ghost % 16 == 0);
        const int64_t len = xax ? p.Nw : p.inner;
        int vec = 0;
        // 32-byte vectors, except where 16-byte vectors measured faster on B200 (8192² Float64, tools/time_2d_v4.py):
        // the non-accumulating contiguous-axis kernels (∂x 5.9 vs 5.6, m̂x 5.7 vs 5.1, weighted m̂x 5.3 vs 4.4 TB/s)
        const int vmode = op->variant & 0xF;  // 4: force 16-byte, 5: force 32-byte (experiments)
        const bool prefer_small = (vmode == 4) || (vmode != 5 && !add_mode && (xax || t.kind == KIND_MEAN_ARITH));  // (m̂y: 5.6 vs 5.3)
        if (aligned && len % VBIG == 0 && !prefer_small) vec = VBIG;
        else if (aligned && len % VSMALL == 0) vec = VSMALL;
        const bool full_range = (p.i_begin == 0 && p.i_end == p.Nw);
        if (vec && op->variant != 3 && (!xax || (full_range && p.Nw >= 64 * vec))) {
            dim3 vgrid, vblock;
            if (xax) {
                vblock = dim3(256, 1, 1);
                vgrid.x = (unsigned)((p.Nw / vec + 255) / 256);
                // a thread keeps the table entries of its VEC elements in registers and walks down the rows: enough CTAs to
                // fill the GPU a few times over (op->march: CTAs per SM, default 24; 8 for accumulating weighted means), not one CTA per R rows, so that the
                // table loads are shared by many rows (8192² Float64 weighted m̂x: see profiles/r2_time_2d_vec.jsonl)
                const int64_t rowblocks = (p.o_end - p.o_begin + R - 1) / R;
                const int per_sm = op->march > 0 ? op->march : ((add_mode && t.kind == KIND_MEAN_WEIGHTED) ? 8 : 24);
                const int64_t want = ((int64_t)148 * per_sm + vgrid.x - 1) / vgrid.x;
                vgrid.y = (unsigned)std::min<int64_t>(std::min<int64_t>(rowblocks, std::max<int64_t>(want, 1)), 65535);
                vgrid.z = 1;
            } else {
                int bx = 32;
                while (bx < 256 && bx < p.inner / vec) bx *= 2;
                const int by = 256 / bx;
                vblock = dim3(bx, by, 1);
                vgrid.x = (unsigned)((p.inner / vec + bx - 1) / bx);
                const int64_t rs = add_mode ? 2 : R;  // rows per thread of the strided kernel
                vgrid.y = (unsigned)std::min<int64_t>((p.i_end - p.i_begin + (int64_t)by * rs - 1) / ((int64_t)by * rs), 65535);
                vgrid.z = (unsigned)std::min<int64_t>(p.o_end - p.o_begin, 65535);
            }
#define LAUNCH_VEC(KERN, KIND, VEC)                                                            \
            do {                                                                                \
                if (add_mode) KERN<T, C, KIND, true, VEC, R><<<vgrid, vblock, 0, st>>>(p);       \
                else KERN<T, C, KIND, false, VEC, R><<<vgrid, vblock, 0, st>>>(p);               \
            } while (0)
#define LAUNCH_VEC_KIND(KERN, VEC)                                                              \
            switch (t.kind) {                                                                   \
                case KIND_PARTIAL: LAUNCH_VEC(KERN, KIND_PARTIAL, VEC); break;                  \
                case KIND_MEAN_ARITH: LAUNCH_VEC(KERN, KIND_MEAN_ARITH, VEC); break;            \
                default: LAUNCH_VEC(KERN, KIND_MEAN_WEIGHTED, VEC); break;                      \
            }
#define LAUNCH_STRIDED(KIND, VEC)                                                               \
            do {                                                                                \
                if (add_mode) k_term_s<T, C, KIND, true, VEC, 2><<<vgrid, vblock, 0, st>>>(p);   \
                else k_term_s<T, C, KIND, false, VEC, R><<<vgrid, vblock, 0, st>>>(p);           \
            } while (0)
#define LAUNCH_STRIDED_KIND(VEC)                                                                \
            switch (t.kind) {                                                                   \
                case KIND_PARTIAL: LAUNCH_STRIDED(KIND_PARTIAL, VEC); break;                    \
                case KIND_MEAN_ARITH: LAUNCH_STRIDED(KIND_MEAN_ARITH, VEC); break;              \
                default: LAUNCH_STRIDED(KIND_MEAN_WEIGHTED, VEC); break;                        \
            }
            if (xax) { if (vec == VBIG) { LAUNCH_VEC_KIND(k_term_x, VBIG) } else { LAUNCH_VEC_KIND(k_term_x, VSMALL) } }
            else     { if (vec == VBIG) { LAUNCH_STRIDED_KIND(VBIG) } else { LAUNCH_STRIDED_KIND(VSMALL) } }
#undef LAUNCH_STRIDED_KIND
#undef LAUNCH_STRIDED
#undef LAUNCH_VEC_KIND
#undef LAUNCH_VEC
            return check_launch(xax ? "k_term_x" : "k_term_s");
        }
    }

    const bool xaxis = (p.inner == 1);
    const int64_t lenx = xaxis ? (p.i_end - p.i_begin) : p.inner;
    int bdx = 32;
    while (bdx < 256 && bdx < lenx) bdx *= 2;
    const int bdy = 256 / bdx;
    dim3 block(bdx, bdy, 1), grid;
    grid.x = (unsigned)((lenx + bdx - 1) / bdx);
    if (xaxis) {
        const int64_t rows = p.o_end - p.o_begin;
        grid.y = (unsigned)std::min<int64_t>((rows + bdy - 1) / bdy, 65535);
        grid.z = 1;
    } else {
        const int64_t rows = p.i_end - p.i_begin;
        grid.y = (unsigned)std::min<int64_t>((rows + bdy - 1) / bdy, 65535);
        grid.z = (unsigned)std::min<int64_t>(p.o_end - p.o_begin, 65535);
    }
#define LAUNCH_TERM(KIND)                                                                     \
    do {                                                                                      \
        if (xaxis) {                                                                          \
            if (add_mode) k_term<T, C, KIND, true, true><<<grid, block, 0, st>>>(p);          \
            else k_term<T, C, KIND, false, true><<<grid, block, 0, st>>>(p);                  \
        } else {                                                                              \
            if (add_mode) k_term<T, C, KIND, true, false><<<grid, block, 0, st>>>(p);         \
            else k_term<T, C, KIND, false, false><<<grid, block, 0, st>>>(p);                 \
        }                                                                                     \
    } while (0)
    switch (t.kind) {
        case KIND_PARTIAL: LAUNCH_TERM(KIND_PARTIAL); break;
        case KIND_MEAN_ARITH: LAUNCH_TERM(KIND_MEAN_ARITH); break;
        default: LAUNCH_TERM(KIND_MEAN_WEIGHTED); break;
    }
#undef LAUNCH_TERM
    return check_launch("k_term");
}

template <typename T>
static int launch_zero(const sgc_op *op, T *G, int64_t k_begin, int64_t k_end, cudaStream_t st) {
    const int64_t plane = op->M / op->N[op->ndim - 1];
    const int64_t n = (k_end - k_begin) * plane;
    if (n <= 0) return SGC_OK;
    T *g = G + k_begin * plane;
    const int64_t blocks = std::min<int64_t>((n + 255) / 256, 148 * 32);
    k_fill_zero<T><<<(unsigned)blocks, 256, 0, st>>>(g, n);
    return check_launch("k_fill_zero");
}

// ---- one-launch compositions (sgc_fused.cuh): divg, grad, 2-D TE/TM sub-curls ---------------------
template <typename T, typename C, int P, bool ADD, int VEC, int R, int MINB = 1, int MODE = FUSED_ALL>
static void launch_fused_one(const FusedParams<T, C> &p, cudaStream_t st) {
    int bx = 32;
    while (bx < 256 && (int64_t)bx * VEC < p.Nx) bx *= 2;
    const int by = 256 / bx;
    dim3 block(bx, by, 1), grid;
    grid.x = (unsigned)((p.Nx / VEC + bx - 1) / bx);
    grid.y = (unsigned)((p.Ny + (int64_t)by * R - 1) / ((int64_t)by * R));
    grid.z = (unsigned)std::min<int64_t>(p.k_end - p.k_begin, 65535);
    k_fused<T, C, P, ADD, VEC, R, MINB, MODE><<<grid, block, 0, st>>>(p);
}

#ifdef SGC_FUSED_SWEEP
// development sweep (tools/sweep_fused.py): vector width × rows per thread × CTAs/SM × split launch,
// Float64 non-accumulating TE / TM / divg-3D / grad-3D only
template <typename T, typename C, int P, bool ADD>
static bool launch_fused_sweep(const FusedParams<T, C> &p, int vecbytes, int rows, int minb, int split, cudaStream_t st) {
    (void)split;
    if constexpr (std::is_same<T, double>::value || std::is_same<C, cplx>::value) {
#define SW(VB, RR, MB)                                                                                      \
        if (vecbytes == VB && rows == RR && minb == MB) {                                                   \
            launch_fused_one<T, C, P, ADD, VB / (int)sizeof(T), RR, MB, FUSED_ALL>(p, st);                   \
            return true; }
        SW(32, 4, 1) SW(32, 4, 2) SW(32, 2, 2) SW(32, 2, 3) SW(32, 2, 4) SW(16, 4, 2) SW(16, 4, 3) SW(16, 4, 4) SW(16, 2, 3) SW(16, 2, 4)
        SW(32, 1, 3) SW(32, 1, 4) SW(16, 1, 4)
#undef SW
    }
    return false;
}
#endif

// (vector bytes, rows per thread, CTAs/SM) per composition from the B200 sweep
// (profiles/r1_sweep_fused.jsonl): 16-byte vectors and 3-4 resident CTAs win almost everywhere; the
// Float64 two-input compositions prefer 32-byte vectors of a single row.
template <typename T, typename C, int P, int VB, int RR, int MB>
static void launch_fused_cfg(const FusedParams<T, C> &p, bool add_mode, cudaStream_t st) {
    if (add_mode) launch_fused_one<T, C, P, true, VB / (int)sizeof(T), RR, MB>(p, st);
    else launch_fused_one<T, C, P, false, VB / (int)sizeof(T), RR, MB>(p, st);
}
template <typename T, typename C, int P>
static void launch_fused_pat(const FusedParams<T, C> &p, bool add_mode, cudaStream_t st) {
    if constexpr (P == PAT_GATHER_XYZ || P == PAT_SCATTER_XYZ || P == PAT_ONE_Z) {
        launch_fused_cfg<T, C, P, 16, 2, 4>(p, add_mode, st);
    } else if constexpr (sizeof(T) == 16) {
        launch_fused_cfg<T, C, P, 16, 4, 2>(p, add_mode, st);
    } else if constexpr (P == PAT_GATHER_XY || P == PAT_GATHER_YX) {
        if (p.Nx % 4 == 0) launch_fused_cfg<T, C, P, 32, 1, 4>(p, add_mode, st);
        else launch_fused_cfg<T, C, P, 16, 4, 3>(p, add_mode, st);
    } else if constexpr (P == PAT_ONE_Y) {
        // accumulating ∂y (read F, read-modify-write G: 24 B/cell): 32-byte vectors of 2 rows, 4 CTAs/SM — 8192² Float64
        // 0.311 → 0.249 ms (profiles/r2_sweep_fused_y.jsonl); the non-accumulating form keeps 16-byte vectors of 4 rows
        if (add_mode && p.Nx % 4 == 0) launch_fused_one<T, C, P, true, 32 / (int)sizeof(T), 2, 4>(p, st);
        else launch_fused_cfg<T, C, P, 16, 4, 3>(p, add_mode, st);
    } else {
        launch_fused_cfg<T, C, P, 16, 4, 3>(p, add_mode, st);
    }
}

// *done = false when the call cannot take the fused path (caller falls back to the sweeps)
template <typename T, typename C>
static int launch_fused(const sgc_op *op, T *G, const T *F, const void *const *ghost, bool add_mode,
                        int64_t k_begin, int64_t k_end, cudaStream_t st, bool *done) {
    *done = false;
    constexpr int VSMALL = 16 / (int)sizeof(T);
    const int last = op->ndim - 1;
    // 3-D divergence / gradient: the staged z-march of the fused curl (every plane read from HBM once, z-neighbour
    // in registers) — 256³ ComplexF64: 0.19 ms against 0.26–0.28 ms for the row-register kernel (variant 6 selects that)
    if ((op->fused_pat == PAT_GATHER_XYZ || op->fused_pat == PAT_SCATTER_XYZ) && (op->variant & 0xF) != 6 &&
        op->N[0] * op->N[1] < (int64_t)0x7fffffff && op->N[2] < (int64_t)0x7fffffff) {
        int s = sgc_launch_divgrad3<T, C>(op, op->fused_pat == PAT_GATHER_XYZ, G, F, ghost, add_mode, k_begin, k_end, st);
        *done = (s == SGC_OK);
        return s;
    }
    if (op->ndim == 2 && (k_begin != 0 || k_end != op->N[1] || op->lo_iface || op->hi_iface)) return SGC_OK;
    if (((uintptr_t)G | (uintptr_t)F) % 16 != 0 || (op->M * sizeof(T)) % 16 != 0) return SGC_OK;
    if (op->N[0] % VSMALL != 0) return SGC_OK;  // rows must be whole 16-byte vectors
    FusedParams<T, C> p{};
    p.Nx = op->N[0]; p.Ny = op->N[1]; p.Nz = op->N[2];
    if (op->ndim == 3) { p.k_begin = k_begin; p.k_end = k_end; }
    else { p.k_begin = 0; p.k_end = 1; }
    for (int d = 0; d < 3; ++d) {
        p.F[d] = F + (int64_t)std::min(d, op->n_in - 1) * op->M;
        p.G[d] = G + (int64_t)std::min(d, op->n_out - 1) * op->M;
        p.ghost[d] = p.F[d];
    }
    for (size_t q = 0; q < op->terms.size(); ++q) {
        const Term &t = op->terms[q];
        FTerm<C> &f = p.t[q];
        f.c = (const C *)((const unsigned char *)op->arena + t.off_c);
        f.e = t.e; f.e_real = t.e_real; f.fwd = t.fwd;
        f.bc_lo = f.bc_hi = t.bloch ? BC_BLOCH : BC_SYM;
        if (t.axis == 2 && last == 2) {
            if (op->lo_iface) f.bc_lo = BC_IFACE;
            if (op->hi_iface) f.bc_hi = BC_IFACE;
            const T *g = ghost ? (const T *)ghost[t.in] : nullptr;
            if (g && (uintptr_t)g % 16 != 0) return SGC_OK;
            p.ghost[t.in] = g ? g : (t.fwd ? p.F[t.in] : p.F[t.in] + (op->N[2] - 1) * op->N[0] * op->N[1]);
        }
    }
    if (p.k_end <= p.k_begin) { *done = true; return SGC_OK; }
    if (p.Ny > 65535LL) return SGC_OK;  // grid.y limit (rows per thread may be 1)
    // (vector bytes, rows per thread, CTAs/SM) per composition from the B200 sweep
    // (profiles/r1_sweep_fused.jsonl): 16-byte vectors and 3-4 resident CTAs win almost everywhere;
    // the Float64 two-input compositions prefer 32-byte vectors of a single row.
#ifdef SGC_FUSED_SWEEP
    if (op->tile_x && (!add_mode || op->fused_pat == 7)) {
        bool ok = false;
        switch (op->fused_pat) {
            case 1: ok = launch_fused_sweep<T, C, 1, false>(p, op->tile_x, op->tile_y, op->march, 0, st); break;
            case 2: ok = launch_fused_sweep<T, C, 2, false>(p, op->tile_x, op->tile_y, op->march, 0, st); break;
            case 4: ok = launch_fused_sweep<T, C, 4, false>(p, op->tile_x, op->tile_y, op->march, 0, st); break;
            case 5: ok = launch_fused_sweep<T, C, 5, false>(p, op->tile_x, op->tile_y, op->march, 0, st); break;
            case 7: ok = add_mode ? launch_fused_sweep<T, C, 7, true>(p, op->tile_x, op->tile_y, op->march, 0, st)
                                  : launch_fused_sweep<T, C, 7, false>(p, op->tile_x, op->tile_y, op->march, 0, st); break;
            default: break;
        }
        if (!ok) return fail(SGC_ERR_INVALID, "fused sweep: configuration not built");
        *done = true;
        return check_launch("k_fused");
    }
#endif
    switch (op->fused_pat) {
        case 0: launch_fused_pat<T, C, 0>(p, add_mode, st); break;
        case 1: launch_fused_pat<T, C, 1>(p, add_mode, st); break;
        case 2: launch_fused_pat<T, C, 2>(p, add_mode, st); break;
        case 3: launch_fused_pat<T, C, 3>(p, add_mode, st); break;
        case 4: launch_fused_pat<T, C, 4>(p, add_mode, st); break;
        case 5: launch_fused_pat<T, C, 5>(p, add_mode, st); break;
        case 6: launch_fused_pat<T, C, 6>(p, add_mode, st); break;
        case 7: launch_fused_pat<T, C, 7>(p, add_mode, st); break;
        case 8: launch_fused_pat<T, C, 8>(p, add_mode, st); break;
        default: return SGC_OK;
    }
    *done = true;
    return check_launch("k_fused");
}

template <typename T>
static int apply_typed(const sgc_op *op, T *G, const T *F, int opmode, const void *const *ghost,
                       int64_t k_begin, int64_t k_end, cudaStream_t st) {
    const bool user_add = (opmode == SGC_OP_ADD);
    if (op->fused_curl && (op->variant & 0xF) != 2) {
        if constexpr (std::is_same<T, double>::value) {
            return sgc_launch_curl3<double, double>(op, G, F, ghost, user_add, k_begin, k_end, st);
        } else {
            if (op->fused_cc) return sgc_launch_curl3<cplx, cplx>(op, G, F, ghost, user_add, k_begin, k_end, st);
            return sgc_launch_curl3<cplx, double>(op, G, F, ghost, user_add, k_begin, k_end, st);
        }
    }
    if (op->fused_pat >= 0 && (op->variant & 0xF) != 2 && (op->variant & 0xF) != 3) {
        bool done = false;
        int s;
        if constexpr (std::is_same<T, double>::value) {
            s = launch_fused<double, double>(op, G, F, ghost, user_add, k_begin, k_end, st, &done);
        } else {
            s = op->terms[0].cc ? launch_fused<cplx, cplx>(op, G, F, ghost, user_add, k_begin, k_end, st, &done)
                                : launch_fused<cplx, double>(op, G, F, ghost, user_add, k_begin, k_end, st, &done);
        }
        if (s || done) return s;
    }
    // apply_mean! on a 3-D grid: one launch (variant 2 / 3: the reference's three sweeps); slabs keep the sweeps
    if (op->fused_mean && (op->variant & 0xF) != 2 && (op->variant & 0xF) != 3 && !ghost && !op->lo_iface && !op->hi_iface && !op->has_slab_w) {
        if constexpr (std::is_same<T, double>::value) {
            return sgc_launch_mean3<double, double>(op, G, F, user_add, k_begin, k_end, st);
        } else {
            return op->terms[0].cc ? sgc_launch_mean3<cplx, cplx>(op, G, F, user_add, k_begin, k_end, st)
                                   : sgc_launch_mean3<cplx, double>(op, G, F, user_add, k_begin, k_end, st);
        }
    }
    // per-block sweeps in the reference's order; the first executed step of each output
    // component uses the caller's OP, the following ones `+=`.
    int cur_out = -1;
    bool add_mode = user_add;
    for (const Term &t : op->terms) {
        if (t.out != cur_out) { cur_out = t.out; add_mode = user_add; }
        T *Gv = G + (int64_t)t.out * op->M;
        const T *Fu = F + (int64_t)t.in * op->M;
        if (t.kind < 0) {
            if (!add_mode) {
                int s = launch_zero<T>(op, Gv, k_begin, k_end, st);
                if (s) return s;
                add_mode = true;
            }
            continue;
        }
        const T *gh = ghost ? (const T *)ghost[t.in] : nullptr;
        int s;
        if constexpr (std::is_same<T, double>::value) {
            s = launch_term<double, double>(op, t, Gv, Fu, gh, add_mode, k_begin, k_end, st);
        } else {
            s = t.cc ? launch_term<cplx, cplx>(op, t, Gv, Fu, gh, add_mode, k_begin, k_end, st)
                     : launch_term<cplx, double>(op, t, Gv, Fu, gh, add_mode, k_begin, k_end, st);
        }
        if (s) return s;
        add_mode = true;
    }
    return SGC_OK;
}

extern "C" int sgc_op_apply_range(const sgc_op *op, void *G, const void *F, int opmode,
                                  const void *const *ghost, int64_t k_begin, int64_t k_end, void *stream) {
    REQUIRE(op && G && F, "NULL argument");
    REQUIRE(opmode == SGC_OP_SET || opmode == SGC_OP_ADD, "opmode must be SGC_OP_SET or SGC_OP_ADD");
    const int64_t Nl = op->N[op->ndim - 1];
    REQUIRE(0 <= k_begin && k_begin <= k_end && k_end <= Nl, "plane range [%lld,%lld) outside 0..%lld",
            (long long)k_begin, (long long)k_end, (long long)Nl);
    REQUIRE(G != F, "G and F must not alias");
    cudaStream_t st = (cudaStream_t)stream;
    if (op->dtype == SGC_F64) return apply_typed<double>(op, (double *)G, (const double *)F, opmode, ghost, k_begin, k_end, st);
    return apply_typed<cplx>(op, (cplx *)G, (const cplx *)F, opmode, ghost, k_begin, k_end, st);
}

extern "C" int sgc_op_apply(const sgc_op *op, void *G, const void *F, int opmode, void *stream) {
    REQUIRE(op, "op is NULL");
    return sgc_op_apply_range(op, G, F, opmode, nullptr, 0, op->N[op->ndim - 1], stream);
}

extern "C" int sgc_op_launches(const sgc_op *op, int opmode) {
    if (!op) return 0;
    if (op->fused_curl && (op->variant & 0xF) != 2) return 1;
    if (op->fused_pat >= 0 && (op->variant & 0xF) != 2 && (op->variant & 0xF) != 3 &&
        op->N[0] % (op->dtype == SGC_C128 ? 1 : 2) == 0 && (op->M * (op->dtype == SGC_C128 ? 16 : 8)) % 16 == 0)
        return 1;  // (aligned, full-range calls; others fall back to the sweeps)
    if (op->fused_mean && (op->variant & 0xF) != 2 && (op->variant & 0xF) != 3 && !op->lo_iface && !op->hi_iface && !op->has_slab_w) return 1;
    int n = 0, cur_out = -1;
    bool add_mode = opmode == SGC_OP_ADD;
    for (const Term &t : op->terms) {
        if (t.out != cur_out) { cur_out = t.out; add_mode = opmode == SGC_OP_ADD; }
        if (t.kind < 0) { if (!add_mode) { ++n; add_mode = true; } continue; }
        ++n;
        add_mode = true;
    }
    return n;
}

// ---- one-shot calls -------------------------------------------------------------------------------
// The reference's calls carry every parameter each time (apply_∂!(Gv, Fu, Val(OP), nw, isfwd, ∆w⁻¹, isbloch, e; α)).  A
// transient handle per call would cost a table upload, two allocations and a stream synchronisation against a kernel
// of a few hundred microseconds, so the handles live in a small cache keyed by EVERYTHING that defines the operator
// (the coefficient vectors by content): a repeated call — a solver loop written against the reference's API — finds
// its tables resident and only queues the kernel; the call returns without synchronising.
namespace {
struct OneShotKey {
    std::string k;
    void raw(const void *p, size_t n) { k.append((const char *)p, n); }
    template <typename V> void val(V v) { raw(&v, sizeof v); }
    void ints(const int *p, int n) { val(n); if (p) raw(p, sizeof(int) * (size_t)n); else val((char)0); }
    void vec(const void *p, int64_t n, int cc) { val((char)(p ? 1 : 0)); if (p) raw(p, (size_t)n * (cc ? 16 : 8)); }
    void pair(const double *p) { if (p) raw(p, 16); else val((char)0); }
    void pairs(const double *p, int n) { if (p) raw(p, 16 * (size_t)n); else val((char)0); }
};
struct OneShotEntry { std::string key; sgc_op *op; uint64_t used; };
std::mutex g_oneshot_mu;
std::vector<OneShotEntry> g_oneshot;
uint64_t g_oneshot_clock = 0;
constexpr size_t ONESHOT_CAPACITY = 32;

sgc_op *oneshot_find(const std::string &key) {
    std::lock_guard<std::mutex> lock(g_oneshot_mu);
    for (auto &e : g_oneshot)
        if (e.key == key) { e.used = ++g_oneshot_clock; return e.op; }
    return nullptr;
}
void oneshot_keep(std::string key, sgc_op *op) {
    sgc_op *victim = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_oneshot_mu);
        if (g_oneshot.size() >= ONESHOT_CAPACITY) {
            size_t lru = 0;
            for (size_t q = 1; q < g_oneshot.size(); ++q)
                if (g_oneshot[q].used < g_oneshot[lru].used) lru = q;
            victim = g_oneshot[lru].op;
            g_oneshot.erase(g_oneshot.begin() + lru);
        }
        g_oneshot.push_back(OneShotEntry{std::move(key), op, ++g_oneshot_clock});
    }
    if (victim) sgc_op_destroy(victim);  // (cudaFree waits for the kernels that still read its tables)
}
void key_head(OneShotKey &k, int tag, int dtype, int ndim, const int64_t *N) {
    int dev = -1;
    cudaGetDevice(&dev);
    k.val(tag); k.val(dev); k.val(dtype); k.val(ndim);
    for (int d = 0; d < ndim && d < 3; ++d) k.val(N ? N[d] : (int64_t)0);
}
}  // namespace

// drops every cached one-shot operator (tests; before a device reset)
extern "C" int sgc_oneshot_cache_clear(void) {
    std::vector<OneShotEntry> old;
    {
        std::lock_guard<std::mutex> lock(g_oneshot_mu);
        old.swap(g_oneshot);
    }
    for (auto &e : old) sgc_op_destroy(e.op);
    return SGC_OK;
}

#define ONE_SHOT(keyexpr, create_call, G, F)                 \
    if (sgc_check_common(dtype, ndim, N)) return SGC_ERR_INVALID; \
    OneShotKey key;                                          \
    keyexpr;                                                 \
    sgc_op *op = oneshot_find(key.k);                        \
    if (!op) {                                               \
        int s = create_call;                                 \
        if (s) return s;                                     \
        oneshot_keep(key.k, op);                             \
    }                                                        \
    return sgc_op_apply(op, G, F, opmode, stream);

extern "C" int sgc_apply_partial(void *Gv, const void *Fu, int opmode, int dtype, int ndim, const int64_t *N,
                                 int nw, int isfwd, const void *dwinv, int coef_complex, int isbloch,
                                 const double *e, const double *alpha, void *stream) {
    REQUIRE(nw >= 1 && nw <= ndim && ndim <= 3, "nw = %d out of range 1..%d", nw, ndim);
    ONE_SHOT((key_head(key, 1, dtype, ndim, N), key.val(nw), key.val(isfwd), key.val(coef_complex), key.val(isbloch), key.pair(e), key.pair(alpha),
              key.vec(dwinv, N[nw - 1], coef_complex)),
             sgc_op_create_partial(&op, dtype, ndim, N, nw, isfwd, dwinv, coef_complex, isbloch, e, alpha), Gv, Fu)
}
extern "C" int sgc_apply_mhat(void *Gv, const void *Fu, int opmode, int dtype, int ndim, const int64_t *N,
                              int nw, int isfwd, const void *dw, const void *dwpinv, int coef_complex,
                              int isbloch, const double *e, const double *alpha, void *stream) {
    REQUIRE(nw >= 1 && nw <= ndim && ndim <= 3, "nw = %d out of range 1..%d", nw, ndim);
    ONE_SHOT((key_head(key, 2, dtype, ndim, N), key.val(nw), key.val(isfwd), key.val(coef_complex), key.val(isbloch), key.pair(e), key.pair(alpha),
              key.vec(dw, N[nw - 1], coef_complex), key.vec(dwpinv, N[nw - 1], coef_complex)),
             sgc_op_create_mhat(&op, dtype, ndim, N, nw, isfwd, dw, dwpinv, coef_complex, isbloch, e, alpha), Gv, Fu)
}
extern "C" int sgc_apply_curl(void *G, const void *F, int opmode, int dtype, int ndim, const int64_t *N,
                              const int *isfwd, const void *const *dlinv, int coef_complex, const int *isbloch,
                              const double *e, const int *cmp_shp, int Kout, const int *cmp_out, int Kin,
                              const int *cmp_in, const double *alpha, void *stream) {
    REQUIRE(isfwd && isbloch && cmp_shp && cmp_out && cmp_in && Kout >= 1 && Kout <= 3 && Kin >= 1 && Kin <= 3, "bad arguments");
    ONE_SHOT((key_head(key, 3, dtype, ndim, N), key.ints(isfwd, ndim), key.ints(isbloch, ndim), key.ints(cmp_shp, ndim), key.ints(cmp_out, Kout),
              key.ints(cmp_in, Kin), key.val(coef_complex), key.pairs(e, ndim), key.pair(alpha),
              [&] { for (int d = 0; d < ndim; ++d) key.vec(dlinv ? dlinv[d] : nullptr, N[d], coef_complex); }()),
             sgc_op_create_curl(&op, dtype, ndim, N, isfwd, dlinv, coef_complex, isbloch, e, cmp_shp, Kout, cmp_out, Kin, cmp_in, alpha), G, F)
}
extern "C" int sgc_apply_divg(void *g, const void *F, int opmode, int dtype, int ndim, const int64_t *N,
                              const int *isfwd, const void *const *dlinv, int coef_complex, const int *isbloch,
                              const double *e, const double *alpha, void *stream) {
    REQUIRE(isfwd && isbloch, "NULL argument");
    ONE_SHOT((key_head(key, 4, dtype, ndim, N), key.ints(isfwd, ndim), key.ints(isbloch, ndim), key.val(coef_complex), key.pairs(e, ndim), key.pair(alpha),
              [&] { for (int d = 0; d < ndim; ++d) key.vec(dlinv ? dlinv[d] : nullptr, N[d], coef_complex); }()),
             sgc_op_create_divg(&op, dtype, ndim, N, isfwd, dlinv, coef_complex, isbloch, e, alpha), g, F)
}
extern "C" int sgc_apply_grad(void *G, const void *f, int opmode, int dtype, int ndim, const int64_t *N,
                              const int *isfwd, const void *const *dlinv, int coef_complex, const int *isbloch,
                              const double *e, const double *alpha, void *stream) {
    REQUIRE(isfwd && isbloch, "NULL argument");
    ONE_SHOT((key_head(key, 5, dtype, ndim, N), key.ints(isfwd, ndim), key.ints(isbloch, ndim), key.val(coef_complex), key.pairs(e, ndim), key.pair(alpha),
              [&] { for (int d = 0; d < ndim; ++d) key.vec(dlinv ? dlinv[d] : nullptr, N[d], coef_complex); }()),
             sgc_op_create_grad(&op, dtype, ndim, N, isfwd, dlinv, coef_complex, isbloch, e, alpha), G, f)
}
extern "C" int sgc_apply_mean(void *G, const void *F, int opmode, int dtype, int ndim, const int64_t *N,
                              const int *isfwd, const void *const *dl, const void *const *dlpinv,
                              int coef_complex, const int *isbloch, const double *e, const double *alpha,
                              void *stream) {
    REQUIRE(isfwd && isbloch, "NULL argument");
    ONE_SHOT((key_head(key, 6, dtype, ndim, N), key.ints(isfwd, ndim), key.ints(isbloch, ndim), key.val(coef_complex), key.pairs(e, ndim), key.pair(alpha),
              [&] { for (int d = 0; d < ndim; ++d) { key.vec(dl ? dl[d] : nullptr, N[d], coef_complex); key.vec(dlpinv ? dlpinv[d] : nullptr, N[d], coef_complex); } }()),
             sgc_op_create_mean(&op, dtype, ndim, N, isfwd, dl, dlpinv, coef_complex, isbloch, e, alpha), G, F)
}
