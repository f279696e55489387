// sgc_lib.cu — libsgc_b200.so: C-ABI entry points (include/sgc_b200.h), the operator runtime
// (device-resident coefficient tables, launch planning) and the kernel dispatch.
//
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -lineinfo -O3 ...
//         (see staggeredgridcalculus.jl_b200/csrc/Makefile)
// There is deliberately no CPU implementation in this library.
#include "sgc_internal.h"

#include <cub/device/device_scan.cuh>

#include "sgc_assembly.cuh"
#include "sgc_kernels.cuh"

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
#define fail sgc_fail

int sgc_check_launch(const char *what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) cudaGetLastError();
    if (e != cudaSuccess) return sgc_fail(SGC_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
    return SGC_OK;
}
#define check_launch sgc_check_launch
void sgc_count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

extern "C" int sgc_abi_version(void) { return SGC_ABI_VERSION; }
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
struct Term {
    int out, in;      // component positions in the last array dimension of G / F
    int kind;         // KIND_*  or -1 for "Gv .= 0"
    int axis;         // array dimension (0-based)
    int fwd, bloch;
    cplx e;
    int e_real;
    bool cc;          // coefficient tables are complex
    size_t off_c, off_w;  // offsets of the tables in the arena (bytes); w unused unless weighted
    cplx a2, a1, b0, w_hi, w_lo;  // scalars (stored as complex; narrowed when !cc)
};

struct sgc_op {
    int dtype = 0, ndim = 0;
    int64_t N[3] = {1, 1, 1};
    int64_t M = 1;
    int n_out = 0, n_in = 0;
    std::vector<Term> terms;      // in the reference's execution order, grouped by `out`
    bool fused_curl = false;
    bool fused_cc = false;
    size_t fused_off[3] = {0, 0, 0};
    int lo_iface = 0, hi_iface = 0;
    int tile_x = 0, tile_y = 0, march = 0, variant = 0;
    void *arena = nullptr;        // device
    // host-array convenience buffers
    mutable void *dF = nullptr, *dG = nullptr;
    mutable cudaStream_t hstream = nullptr;
};

struct OpBuilder {
    std::vector<unsigned char> host;
    size_t push(const std::vector<HV> &v, bool as_complex) {
        size_t off = (host.size() + 15) & ~size_t(15);
        size_t bytes = v.size() * (as_complex ? 16 : 8);
        host.resize(off + bytes);
        double *d = (double *)(host.data() + off);
        for (size_t i = 0; i < v.size(); ++i) {
            if (as_complex) { d[2 * i] = v[i].re; d[2 * i + 1] = v[i].cx ? v[i].im : 0.0; }
            else d[i] = v[i].re;
        }
        return off;
    }
};

static int finish_op(sgc_op *op, OpBuilder &b) {
    size_t bytes = b.host.size() ? b.host.size() : 16;
    CUDA_TRY(cudaMalloc(&op->arena, bytes));
    if (b.host.size()) CUDA_TRY(cudaMemcpy(op->arena, b.host.data(), b.host.size(), cudaMemcpyHostToDevice));
    return SGC_OK;
}

static int check_common(int dtype, int ndim, const int64_t *N) {
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
    if (op->dF) cudaFree(op->dF);
    if (op->dG) cudaFree(op->dG);
    if (op->hstream) cudaStreamDestroy(op->hstream);
    delete op;
    return SGC_OK;
}

#define OP_PROLOGUE()                                         \
    REQUIRE(op_out, "op is NULL");                            \
    *op_out = nullptr;                                        \
    { int _s = check_common(dtype, ndim, N); if (_s) return _s; } \
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
    OP_EPILOGUE();
}

extern "C" int sgc_op_set_slab(sgc_op *op, int lo_interface, int hi_interface) {
    REQUIRE(op, "op is NULL");
    op->lo_iface = lo_interface ? 1 : 0;
    op->hi_iface = hi_interface ? 1 : 0;
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
        if (aligned && len % VBIG == 0) vec = VBIG;
        else if (aligned && len % VSMALL == 0) vec = VSMALL;
        const bool full_range = (p.i_begin == 0 && p.i_end == p.Nw);
        if (vec && op->variant != 3 && (!xax || (full_range && p.Nw >= 64 * vec))) {
            dim3 vgrid, vblock;
            if (xax) {
                vblock = dim3(256, 1, 1);
                vgrid.x = (unsigned)((p.Nw / vec + 255) / 256);
                vgrid.y = (unsigned)std::min<int64_t>((p.o_end - p.o_begin + R - 1) / R, 65535);
                vgrid.z = 1;
            } else {
                int bx = 32;
                while (bx < 256 && bx < p.inner / vec) bx *= 2;
                const int by = 256 / bx;
                vblock = dim3(bx, by, 1);
                vgrid.x = (unsigned)((p.inner / vec + bx - 1) / bx);
                vgrid.y = (unsigned)std::min<int64_t>((p.i_end - p.i_begin + (int64_t)by * R - 1) / ((int64_t)by * R), 65535);
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
            if (xax) { if (vec == VBIG) { LAUNCH_VEC_KIND(k_term_x, VBIG) } else { LAUNCH_VEC_KIND(k_term_x, VSMALL) } }
            else     { if (vec == VBIG) { LAUNCH_VEC_KIND(k_term_s, VBIG) } else { LAUNCH_VEC_KIND(k_term_s, VSMALL) } }
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

template <typename T, typename C, bool ADD>
static void launch_curl3_tiles(const CurlParams<T, C> &p, int tx, int ty, dim3 grid_xy_z, cudaStream_t st);
template <typename T, typename C, bool ADD>
static int launch_curl3p_tiles(const CurlParams<T, C> &p, int tx, int ty, int stages, int minb, bool xshfl, dim3 grid, cudaStream_t st);

template <typename T, typename C>
static int launch_curl3(const sgc_op *op, T *G, const T *F, const void *const *ghost, bool add_mode,
                        int64_t k_begin, int64_t k_end, cudaStream_t st) {
    if (k_end <= k_begin) return SGC_OK;
    CurlParams<T, C> p{};
    const int64_t M = op->M, sz = op->N[0] * op->N[1];
    // per-axis BCs come from the ∂ terms (the term for axis d with the same flags exists twice)
    int fwd[3] = {0, 0, 0}, bloch[3] = {1, 1, 1};
    cplx e[3] = {{1, 0}, {1, 0}, {1, 0}};
    int e_real[3] = {1, 1, 1};
    for (const Term &t : op->terms)
        if (t.kind == KIND_PARTIAL) { fwd[t.axis] = t.fwd; bloch[t.axis] = t.bloch; e[t.axis] = t.e; e_real[t.axis] = t.e_real; }
    for (int d = 0; d < 3; ++d) {
        p.F[d] = F + d * M;
        p.G[d] = G + d * M;
        p.c[d] = (const C *)((const unsigned char *)op->arena + op->fused_off[d]);
        p.e[d] = e[d]; p.e_real[d] = e_real[d]; p.fwd[d] = fwd[d];
        p.bc_lo[d] = p.bc_hi[d] = bloch[d] ? BC_BLOCH : BC_SYM;
    }
    if (op->lo_iface) p.bc_lo[2] = BC_IFACE;
    if (op->hi_iface) p.bc_hi[2] = BC_IFACE;
    for (int c = 0; c < 2; ++c) {
        const T *g = ghost ? (const T *)ghost[c] : nullptr;
        p.ghost[c] = g ? g : (fwd[2] ? p.F[c] : p.F[c] + (op->N[2] - 1) * sz);
    }
    p.Nx = (int)op->N[0]; p.Ny = (int)op->N[1]; p.Nz = (int)op->N[2];
    p.k_begin = (int)k_begin; p.k_end = (int)k_end;

    const int kernel_variant = op->variant & 0xF;
    // defaults from the 256^3 sweeps on B200 (profiles/r1_sweep_curl_256*.jsonl): 32x4 tiles, 6-stage
    // ring, 4 CTAs/SM (128 registers, no spills), 12-plane z-march
    int tx = op->tile_x ? op->tile_x : 32;
    int ty = op->tile_y ? op->tile_y : (kernel_variant == 1 ? 8 : 4);
    const int64_t gx = (p.Nx + tx - 1) / tx, gy = (p.Ny + ty - 1) / ty;
    const int64_t nk = k_end - k_begin;
    int march = op->march;
    if (march <= 0) {
        if (kernel_variant == 1) {
            const int64_t want = 148 * 12;  // ≥ ~4 waves of 3 CTAs/SM
            const int64_t zch = std::max<int64_t>(1, (want + gx * gy - 1) / (gx * gy));
            march = (int)std::max<int64_t>(8, std::min<int64_t>(64, nk / zch));
        } else {
            march = 12;
        }
    }
    p.march = march;
    dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)((nk + march - 1) / march));
    if (grid.y > 65535 || grid.z > 65535) return fail(SGC_ERR_INVALID, "grid too large for the fused curl tiling");
    if (kernel_variant == 1) {  // v1: register-prefetch z-march
        if (add_mode) launch_curl3_tiles<T, C, true>(p, tx, ty, grid, st);
        else launch_curl3_tiles<T, C, false>(p, tx, ty, grid, st);
        return check_launch("k_curl3");
    }
    int stages = (op->variant >> 4) & 0xF;
    if (!stages) stages = (tx == 32 && ty == 4) ? 6 : 4;  // 6-stage ring: 4 planes in flight per CTA, 47.5 KB, 4 CTAs/SM
    int minb = (op->variant >> 8) & 0xF;
    if (!minb) minb = (tx * ty <= 128) ? 4 : ((tx * ty <= 256) ? 3 : 2);
    const bool xshfl = ((op->variant >> 12) & 1) == 0;  // bit 12: read x-neighbours from the shared tile instead
    int s = add_mode ? launch_curl3p_tiles<T, C, true>(p, tx, ty, stages, minb, xshfl, grid, st)
                     : launch_curl3p_tiles<T, C, false>(p, tx, ty, stages, minb, xshfl, grid, st);
    if (s) return s;
    return check_launch("k_curl3p");
}

// v2: cp.async shared-memory ring.
template <typename T, typename C, bool ADD, int TX, int TY, int S, int MINB, bool XSHFL>
static int launch_curl3p_one(const CurlParams<T, C> &p, dim3 grid, cudaStream_t st) {
    constexpr size_t smem = (size_t)S * 3 * (TY + 1) * (TX + 1) * sizeof(T);
    auto kern = k_curl3p<T, C, ADD, TX, TY, S, MINB, XSHFL>;
    static bool attr_set = false;  // per instantiation
    if (!attr_set) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    kern<<<grid, dim3(TX, TY, 1), smem, st>>>(p);
    return SGC_OK;
}

template <typename T, typename C, bool ADD>
static int launch_curl3p_tiles(const CurlParams<T, C> &p, int tx, int ty, int stages, int minb, bool xshfl, dim3 grid, cudaStream_t st) {
#define CURLP_CASE(TX, TY, S, MINB) \
    if (tx == TX && ty == TY && stages == S && minb == MINB)                                                   \
        return xshfl ? launch_curl3p_one<T, C, ADD, TX, TY, S, MINB, true>(p, grid, st)                        \
                     : launch_curl3p_one<T, C, ADD, TX, TY, S, MINB, false>(p, grid, st);
    CURLP_CASE(32, 4, 4, 4)
    CURLP_CASE(32, 4, 5, 4)
    CURLP_CASE(32, 4, 6, 4)
    CURLP_CASE(32, 4, 7, 4)
    CURLP_CASE(32, 4, 8, 3)
    CURLP_CASE(32, 4, 3, 4)
    CURLP_CASE(32, 8, 3, 3)
    CURLP_CASE(32, 8, 4, 3)
    CURLP_CASE(32, 8, 4, 2)
    CURLP_CASE(32, 8, 6, 2)
    CURLP_CASE(32, 16, 3, 2)
    CURLP_CASE(32, 16, 4, 1)
    CURLP_CASE(64, 4, 4, 3)
    CURLP_CASE(64, 4, 4, 2)
    CURLP_CASE(64, 8, 3, 2)
#undef CURLP_CASE
    return sgc_fail(SGC_ERR_INVALID, "fused curl: tile %dx%d, %d stages, %d CTAs/SM is not built", tx, ty, stages, minb);
}

template <typename T, typename C, bool ADD>
static void launch_curl3_tiles(const CurlParams<T, C> &p, int tx, int ty, dim3 grid, cudaStream_t st) {
#define CURL_CASE(TX, TY) \
    if (tx == TX && ty == TY) { k_curl3<T, C, ADD, TX, TY><<<grid, dim3(TX, TY, 1), 0, st>>>(p); return; }
    CURL_CASE(32, 4)
    CURL_CASE(32, 8)
    CURL_CASE(32, 16)
    CURL_CASE(64, 4)
    CURL_CASE(64, 8)
#undef CURL_CASE
    k_curl3<T, C, ADD, 32, 8><<<dim3((p.Nx + 31) / 32, (p.Ny + 7) / 8, grid.z), dim3(32, 8, 1), 0, st>>>(p);
}

template <typename T>
static int apply_typed(const sgc_op *op, T *G, const T *F, int opmode, const void *const *ghost,
                       int64_t k_begin, int64_t k_end, cudaStream_t st) {
    const bool user_add = (opmode == SGC_OP_ADD);
    if (op->fused_curl && (op->variant & 0xF) != 2) {
        if constexpr (std::is_same<T, double>::value) {
            return launch_curl3<double, double>(op, G, F, ghost, user_add, k_begin, k_end, st);
        } else {
            if (op->fused_cc) return launch_curl3<cplx, cplx>(op, G, F, ghost, user_add, k_begin, k_end, st);
            return launch_curl3<cplx, double>(op, G, F, ghost, user_add, k_begin, k_end, st);
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

extern "C" int sgc_op_apply_host(const sgc_op *op, void *G_host, const void *F_host, int opmode) {
    REQUIRE(op && G_host && F_host, "NULL argument");
    const size_t es = op->dtype == SGC_C128 ? 16 : 8;
    const size_t bF = es * (size_t)op->M * op->n_in, bG = es * (size_t)op->M * op->n_out;
    if (!op->hstream) CUDA_TRY(cudaStreamCreateWithFlags(&op->hstream, cudaStreamNonBlocking));
    if (!op->dF) CUDA_TRY(cudaMalloc(&op->dF, bF));
    if (!op->dG) CUDA_TRY(cudaMalloc(&op->dG, bG));
    CUDA_TRY(cudaMemcpyAsync(op->dF, F_host, bF, cudaMemcpyHostToDevice, op->hstream));
    if (opmode == SGC_OP_ADD) CUDA_TRY(cudaMemcpyAsync(op->dG, G_host, bG, cudaMemcpyHostToDevice, op->hstream));
    int s = sgc_op_apply(op, op->dG, op->dF, opmode, op->hstream);
    if (s) return s;
    CUDA_TRY(cudaMemcpyAsync(G_host, op->dG, bG, cudaMemcpyDeviceToHost, op->hstream));
    CUDA_TRY(cudaStreamSynchronize(op->hstream));
    return SGC_OK;
}

// ---- one-shot calls -------------------------------------------------------------------------------
#define ONE_SHOT(create_call, G, F)                          \
    sgc_op *op = nullptr;                                    \
    int s = create_call;                                     \
    if (s) return s;                                         \
    s = sgc_op_apply(op, G, F, opmode, stream);              \
    /* tables must outlive the queued kernels */             \
    if (!s) { cudaError_t ce = cudaStreamSynchronize((cudaStream_t)stream); \
              if (ce != cudaSuccess) s = fail(SGC_ERR_CUDA, "sync failed: %s", cudaGetErrorString(ce)); } \
    sgc_op_destroy(op);                                      \
    return s;

extern "C" int sgc_apply_partial(void *Gv, const void *Fu, int opmode, int dtype, int ndim, const int64_t *N,
                                 int nw, int isfwd, const void *dwinv, int coef_complex, int isbloch,
                                 const double *e, const double *alpha, void *stream) {
    ONE_SHOT(sgc_op_create_partial(&op, dtype, ndim, N, nw, isfwd, dwinv, coef_complex, isbloch, e, alpha), Gv, Fu)
}
extern "C" int sgc_apply_mhat(void *Gv, const void *Fu, int opmode, int dtype, int ndim, const int64_t *N,
                              int nw, int isfwd, const void *dw, const void *dwpinv, int coef_complex,
                              int isbloch, const double *e, const double *alpha, void *stream) {
    ONE_SHOT(sgc_op_create_mhat(&op, dtype, ndim, N, nw, isfwd, dw, dwpinv, coef_complex, isbloch, e, alpha), Gv, Fu)
}
extern "C" int sgc_apply_curl(void *G, const void *F, int opmode, int dtype, int ndim, const int64_t *N,
                              const int *isfwd, const void *const *dlinv, int coef_complex, const int *isbloch,
                              const double *e, const int *cmp_shp, int Kout, const int *cmp_out, int Kin,
                              const int *cmp_in, const double *alpha, void *stream) {
    ONE_SHOT(sgc_op_create_curl(&op, dtype, ndim, N, isfwd, dlinv, coef_complex, isbloch, e, cmp_shp, Kout,
                                cmp_out, Kin, cmp_in, alpha), G, F)
}
extern "C" int sgc_apply_divg(void *g, const void *F, int opmode, int dtype, int ndim, const int64_t *N,
                              const int *isfwd, const void *const *dlinv, int coef_complex, const int *isbloch,
                              const double *e, const double *alpha, void *stream) {
    ONE_SHOT(sgc_op_create_divg(&op, dtype, ndim, N, isfwd, dlinv, coef_complex, isbloch, e, alpha), g, F)
}
extern "C" int sgc_apply_grad(void *G, const void *f, int opmode, int dtype, int ndim, const int64_t *N,
                              const int *isfwd, const void *const *dlinv, int coef_complex, const int *isbloch,
                              const double *e, const double *alpha, void *stream) {
    ONE_SHOT(sgc_op_create_grad(&op, dtype, ndim, N, isfwd, dlinv, coef_complex, isbloch, e, alpha), G, f)
}
extern "C" int sgc_apply_mean(void *G, const void *F, int opmode, int dtype, int ndim, const int64_t *N,
                              const int *isfwd, const void *const *dl, const void *const *dlpinv,
                              int coef_complex, const int *isbloch, const double *e, const double *alpha,
                              void *stream) {
    ONE_SHOT(sgc_op_create_mean(&op, dtype, ndim, N, isfwd, dl, dlpinv, coef_complex, isbloch, e, alpha), G, F)
}

// =============================================================================================
// sparse assembly
// =============================================================================================
static void free_checked(void *p, const char *what) {
    if (!p) return;
    cudaError_t e = cudaFree(p);
    if (e != cudaSuccess) {
        fprintf(stderr, "libsgc_b200: cudaFree(%s) failed: %s\n", what, cudaGetErrorString(e));
        cudaGetLastError();  // do not let a failed free poison the next launch check
    }
}

extern "C" int sgc_asm_destroy(sgc_asm *a) {
    if (!a) return SGC_OK;
    free_checked(a->arena, "asm arena");
    free_checked(a->scan_tmp, "scan scratch");
    free_checked(a->d_colptr, "host-path colptr");
    free_checked(a->cta_total, "cta totals");
    free_checked(a->cta_offset, "cta offsets");
    free_checked(a->count_tables, "count tables");
    delete a;
    return SGC_OK;
}

static void init_asm(sgc_asm *a, int ndim, const int64_t *N) {
    a->ndim = ndim;
    a->M = 1;
    for (int d = 0; d < 3; ++d) {
        a->N[d] = d < ndim ? N[d] : 1;
        a->M *= a->N[d];
    }
}

static HV e_of(const double *e, int e_complex) {
    if (!e) return hv(1.0);
    return e_complex ? hvc(e[0], e[1]) : hv(e[0]);
}

// create_∂info (matrix/differential/partial.jl:87-314): per-axis value tables.
static AsmBlockHost partial_block(const sgc_asm *a, int nv, int nu, int axis, int isfwd, const void *dwinv,
                                  int coef_complex, int isbloch, HV e, bool Tcx, int parity) {
    AsmBlockHost B;
    B.nv = nv; B.nu = nu; B.axis = axis; B.ns = isfwd ? 1 : -1;
    const int64_t Nw = a->N[axis];
    const double ns = isfwd ? 1.0 : -1.0;
    const HV one = Tcx ? hvc(1.0, 0.0) : hv(1.0);
    B.V0.resize(Nw); B.Vs.resize(Nw);
    for (int64_t i = 0; i < Nw; ++i) {
        const HV d = vec_at(dwinv, coef_complex, i);
        B.V0[i] = hmul(hmul(hv(-ns), one), d);  // -ns .* ones(T,N) .* ∆W⁻¹          :137
        B.Vs[i] = hmul(hmul(hv(ns), one), d);   //  ns .* ones(T,N) .* ∆W⁻¹          :138
    }
    const int64_t slab = isfwd ? Nw - 1 : 0;
    if (isbloch) {
        const HV en = isfwd ? e : hinv(e);      // e⁻ⁱᵏᴸ^ns                           :217
        B.Vs[slab] = hmul(B.Vs[slab], en);
    } else {
        B.V0[0] = Tcx ? hvc(0.0, 0.0) : hv(0.0);     // :291
        B.Vs[slab] = Tcx ? hvc(0.0, 0.0) : hv(0.0);  // :299
    }
    if (parity != 1)
        for (int64_t i = 0; i < Nw; ++i) {      // V .*= parity                        curl.jl:91
            B.V0[i] = hmul(B.V0[i], hv((double)parity));
            B.Vs[i] = hmul(B.Vs[i], hv((double)parity));
        }
    return B;
}

// create_minfo (matrix/mean.jl:141-377)
static AsmBlockHost mean_block(const sgc_asm *a, int nv, int nu, int axis, int isfwd, const void *dw,
                               const void *dwpinv, int coef_complex, int isbloch, HV e, bool Tcx) {
    AsmBlockHost B;
    B.nv = nv; B.nu = nu; B.axis = axis; B.ns = isfwd ? 1 : -1;
    const int64_t Nw = a->N[axis];
    const HV half = Tcx ? hvc(0.5, 0.0) : hv(0.5);
    B.V0.resize(Nw); B.Vs.resize(Nw);
    for (int64_t i = 0; i < Nw; ++i) {
        const HV w = vec_at(dw, coef_complex, i), wp = vec_at(dwpinv, coef_complex, i);
        B.V0[i] = hmul(hmul(half, w), wp);                     // fill(T(0.5)) .* ∆W .* ∆W′⁻¹     :193
        int64_t is = i + B.ns;
        if (is < 0) is += Nw;
        if (is >= Nw) is -= Nw;
        B.Vs[i] = hmul(hmul(half, vec_at(dw, coef_complex, is)), wp);  // circshift, then .*= ∆W′⁻¹  :195-197
    }
    const int64_t slab = isfwd ? Nw - 1 : 0;
    if (isbloch) {
        const HV en = isfwd ? e : hinv(e);
        B.Vs[slab] = hmul(B.Vs[slab], en);                     // :274
    } else {
        B.V0[0] = hmul(B.V0[0], hv(isfwd ? 0.0 : 2.0));        // :351-352
        B.Vs[slab] = Tcx ? hvc(0.0, 0.0) : hv(0.0);            // :356
    }
    return B;
}

static int finish_asm(sgc_asm *a) {
    OpBuilder b;
    for (auto &B : a->blocks) {
        a->off0.push_back(b.push(B.V0, a->cx));
        a->offs.push_back(b.push(B.Vs, a->cx));
    }
    size_t bytes = b.host.size() ? b.host.size() : 16;
    CUDA_TRY(cudaMalloc(&a->arena, bytes));
    if (b.host.size()) CUDA_TRY(cudaMemcpy(a->arena, b.host.data(), b.host.size(), cudaMemcpyHostToDevice));
    a->row_begin = 0;
    a->row_end = (int64_t)a->Kout * a->M;
    // separable count tables: A[nu][axis][i] = number of nonzero values the blocks of block column
    // nu acting along `axis` store in a column whose subscript along that axis is i
    std::vector<int> tab;
    for (int nu = 0; nu < 3; ++nu)
        for (int ax = 0; ax < 3; ++ax) {
            a->count_off[nu][ax] = tab.size();
            tab.resize(tab.size() + (size_t)a->N[ax], 0);
        }
    auto nonzero = [](const HV &v) { return v.re != 0.0 || (v.cx && v.im != 0.0); };
    for (auto &B : a->blocks) {
        int *A = tab.data() + a->count_off[B.nu][B.axis];
        const int64_t Nw = a->N[B.axis];
        if (B.ns == 0) {  // diagonal-only block: the same count for every cell
            for (int64_t i = 0; i < Nw; ++i) A[i] += nonzero(B.V0[0]) ? 1 : 0;
        } else if (Nw == 1) {
            A[0] += nonzero(hadd(B.V0[0], B.Vs[0])) ? 1 : 0;
        } else {
            for (int64_t i = 0; i < Nw; ++i) {
                int64_t rw = i - B.ns;
                if (rw < 0) rw += Nw;
                if (rw >= Nw) rw -= Nw;
                A[i] += (nonzero(B.V0[i]) ? 1 : 0) + (nonzero(B.Vs[rw]) ? 1 : 0);
            }
        }
    }
    CUDA_TRY(cudaMalloc((void **)&a->count_tables, sizeof(int) * tab.size()));
    CUDA_TRY(cudaMemcpy(a->count_tables, tab.data(), sizeof(int) * tab.size(), cudaMemcpyHostToDevice));
    return SGC_OK;
}

#define ASM_PROLOGUE()                                          \
    REQUIRE(a_out, "handle pointer is NULL");                   \
    *a_out = nullptr;                                           \
    { int _s = check_common(SGC_F64, ndim, N); if (_s) return _s; } \
    sgc_asm *a = new sgc_asm();                                 \
    init_asm(a, ndim, N);
#define ASM_EPILOGUE()                                          \
    { int _s = finish_asm(a); if (_s) { sgc_asm_destroy(a); return _s; } } \
    *a_out = a;                                                 \
    return SGC_OK;
#define ASM_REQUIRE(cond, ...) if (!(cond)) { sgc_asm_destroy(a); return fail(SGC_ERR_INVALID, __VA_ARGS__); }

extern "C" int sgc_asm_create_partial(sgc_asm **a_out, int ndim, const int64_t *N, int nw, int isfwd,
                                      const void *dwinv, int coef_complex, int isbloch, const double *e,
                                      int e_complex) {
    ASM_PROLOGUE();
    ASM_REQUIRE(nw >= 1 && nw <= ndim, "nw = %d out of range 1..%d", nw, ndim);
    a->Kout = a->Kin = 1;
    a->cx = (coef_complex && dwinv) || e_complex;  // partial.jl:136
    a->blocks.push_back(partial_block(a, 0, 0, nw - 1, isfwd, dwinv, coef_complex, isbloch, e_of(e, e_complex), a->cx, 1));
    ASM_EPILOGUE();
}

extern "C" int sgc_asm_create_mhat(sgc_asm **a_out, int ndim, const int64_t *N, int nw, int isfwd, const void *dw,
                                   const void *dwpinv, int coef_complex, int isbloch, const double *e,
                                   int e_complex) {
    ASM_PROLOGUE();
    ASM_REQUIRE(nw >= 1 && nw <= ndim, "nw = %d out of range 1..%d", nw, ndim);
    ASM_REQUIRE((dw == nullptr) == (dwpinv == nullptr), "dw and dwpinv must both be given or both be NULL");
    a->Kout = a->Kin = 1;
    a->cx = (coef_complex && dw) || e_complex;  // mean.jl:192
    a->blocks.push_back(mean_block(a, 0, 0, nw - 1, isfwd, dw, dwpinv, coef_complex, isbloch, e_of(e, e_complex), a->cx));
    ASM_EPILOGUE();
}

extern "C" int sgc_asm_create_curl(sgc_asm **a_out, int ndim, const int64_t *N, const int *isfwd,
                                   const void *const *dlinv, int coef_complex, const int *isbloch,
                                   const double *e, int e_complex, const int *cmp_shp, int Kout,
                                   const int *cmp_out, int Kin, const int *cmp_in, int order_cmpfirst) {
    ASM_PROLOGUE();
    ASM_REQUIRE(isfwd && isbloch && cmp_shp && cmp_out && cmp_in, "NULL argument");
    ASM_REQUIRE(Kout >= 1 && Kout <= 3 && Kin >= 1 && Kin <= 3, "Kout, Kin must be in 1..3");
    a->Kout = Kout; a->Kin = Kin; a->cmpfirst = order_cmpfirst ? 1 : 0;
    a->cx = (coef_complex && dlinv) || e_complex;  // curl.jl:59
    for (int iv = 0; iv < Kout; ++iv)
        for (int iu = 0; iu < Kin; ++iu) {
            ASM_REQUIRE(cmp_out[iv] >= 1 && cmp_out[iv] <= 3 && cmp_in[iu] >= 1 && cmp_in[iu] <= 3, "components must be 1..3");
            const int nv = cmp_out[iv], nu = cmp_in[iu];
            const int parity = CURL_PARITY[nv - 1][nu - 1];
            if (!parity) continue;
            const int nw = 6 - nv - nu;
            int cnt = 0, ind = -1;
            for (int d = 0; d < ndim; ++d)
                if (cmp_shp[d] == nw) { ++cnt; ind = d; }
            ASM_REQUIRE(cnt == 1, "cmp_shp does not have one and only one nw = %d", nw);
            // each create_∂info call promotes with its own ∆ and e (partial.jl:136); values are then
            // converted to the common T
            a->blocks.push_back(partial_block(a, iv, iu, ind, isfwd[ind], dlinv ? dlinv[ind] : nullptr, coef_complex,
                                              isbloch[ind], e_of(e ? e + 2 * ind : nullptr, e_complex), a->cx, parity));
        }
    ASM_EPILOGUE();
}

extern "C" int sgc_asm_create_divg(sgc_asm **a_out, int ndim, const int64_t *N, const int *isfwd,
                                   const void *const *dlinv, int coef_complex, const int *isbloch,
                                   const double *e, int e_complex, int order_cmpfirst) {
    ASM_PROLOGUE();
    ASM_REQUIRE(isfwd && isbloch, "NULL argument");
    a->Kout = 1; a->Kin = ndim; a->cmpfirst = order_cmpfirst ? 1 : 0;
    a->cx = (coef_complex && dlinv) || e_complex;
    for (int d = 0; d < ndim; ++d)
        a->blocks.push_back(partial_block(a, 0, d, d, isfwd[d], dlinv ? dlinv[d] : nullptr, coef_complex, isbloch[d],
                                          e_of(e ? e + 2 * d : nullptr, e_complex), a->cx, 1));
    ASM_EPILOGUE();
}

extern "C" int sgc_asm_create_grad(sgc_asm **a_out, int ndim, const int64_t *N, const int *isfwd,
                                   const void *const *dlinv, int coef_complex, const int *isbloch,
                                   const double *e, int e_complex, int order_cmpfirst) {
    ASM_PROLOGUE();
    ASM_REQUIRE(isfwd && isbloch, "NULL argument");
    a->Kout = ndim; a->Kin = 1; a->cmpfirst = order_cmpfirst ? 1 : 0;
    a->cx = (coef_complex && dlinv) || e_complex;
    for (int d = 0; d < ndim; ++d)
        a->blocks.push_back(partial_block(a, d, 0, d, isfwd[d], dlinv ? dlinv[d] : nullptr, coef_complex, isbloch[d],
                                          e_of(e ? e + 2 * d : nullptr, e_complex), a->cx, 1));
    ASM_EPILOGUE();
}

extern "C" int sgc_asm_create_mean(sgc_asm **a_out, int ndim, const int64_t *N, const int *isfwd,
                                   const void *const *dl, const void *const *dlpinv, int coef_complex,
                                   const int *isbloch, const double *e, int e_complex, int order_cmpfirst) {
    ASM_PROLOGUE();
    ASM_REQUIRE(isfwd && isbloch, "NULL argument");
    ASM_REQUIRE((dl == nullptr) == (dlpinv == nullptr), "dl and dlpinv must both be given or both be NULL");
    a->Kout = a->Kin = ndim; a->cmpfirst = order_cmpfirst ? 1 : 0;
    a->cx = (coef_complex && dl) || e_complex;
    for (int d = 0; d < ndim; ++d)
        a->blocks.push_back(mean_block(a, d, d, d, isfwd[d], dl ? dl[d] : nullptr, dlpinv ? dlpinv[d] : nullptr,
                                       coef_complex, isbloch[d], e_of(e ? e + 2 * d : nullptr, e_complex), a->cx));
    ASM_EPILOGUE();
}

extern "C" int sgc_asm_create_picmp(sgc_asm **a_out, int ndim, const int64_t *N, int Kf, const int *permute,
                                    const void *scale, int scale_complex, int order_cmpfirst) {
    ASM_PROLOGUE();
    ASM_REQUIRE(Kf >= 1 && Kf <= 3 && permute, "Kf must be 1..3");
    a->Kout = a->Kin = Kf; a->cmpfirst = order_cmpfirst ? 1 : 0;
    a->cx = scale_complex && scale;
    for (int w = 0; w < Kf; ++w) {
        ASM_REQUIRE(permute[w] >= 1 && permute[w] <= Kf, "permute entries must be 1..Kf");
        AsmBlockHost B;
        B.nv = permute[w] - 1; B.nu = w; B.axis = 0; B.ns = 0;
        B.V0.assign(1, vec_at(scale, scale_complex, w));
        B.Vs.assign(1, hv(0.0));
        a->blocks.push_back(B);
    }
    ASM_EPILOGUE();
}

extern "C" int sgc_asm_set_row_range(sgc_asm *a, int64_t row_begin, int64_t row_end) {
    REQUIRE(a, "handle is NULL");
    REQUIRE(0 <= row_begin && row_begin <= row_end && row_end <= (int64_t)a->Kout * a->M, "row range out of bounds");
    a->row_begin = row_begin;
    a->row_end = row_end;
    a->nnz = -1;
    a->counted = false;
    return SGC_OK;
}

extern "C" int sgc_asm_shape(const sgc_asm *a, int64_t *m, int64_t *n, int *val_dtype) {
    REQUIRE(a, "handle is NULL");
    if (m) *m = a->row_end - a->row_begin;
    if (n) *n = (int64_t)a->Kin * a->M;
    if (val_dtype) *val_dtype = a->cx ? SGC_C128 : SGC_F64;
    return SGC_OK;
}

template <typename T>
static AsmParams<T> asm_params(const sgc_asm *a, int index_base) {
    AsmParams<T> p{};
    p.nblk = (int)a->blocks.size();
    for (int b = 0; b < p.nblk; ++b) {
        p.blk[b].V0 = (const T *)((const unsigned char *)a->arena + a->off0[b]);
        p.blk[b].Vs = (const T *)((const unsigned char *)a->arena + a->offs[b]);
        p.blk[b].nv = a->blocks[b].nv; p.blk[b].nu = a->blocks[b].nu;
        p.blk[b].axis = a->blocks[b].axis; p.blk[b].ns = a->blocks[b].ns;
    }
    p.Kout = a->Kout; p.Kin = a->Kin; p.cmpfirst = a->cmpfirst;
    for (int d = 0; d < 3; ++d) p.N[d] = a->N[d];
    p.M = a->M;
    p.ncol = (int64_t)a->Kin * a->M;
    p.row_begin = a->row_begin; p.row_end = a->row_end;
    p.index_base = index_base;
    for (int nu = 0; nu < 3; ++nu) p.nblk_of[nu] = 0;
    for (int b = 0; b < p.nblk; ++b) {
        const int nu = a->blocks[b].nu;
        if (p.nblk_of[nu] < 3) p.blk_of[nu][p.nblk_of[nu]] = b;
        p.nblk_of[nu]++;
    }
    p.cols_per_group = a->cmpfirst ? a->Kin : 1;
    p.ngroups = a->cmpfirst ? a->M : (int64_t)a->Kin * a->M;
    return p;
}

constexpr int ASM_BLOCK = 128;

// entry slots per column needed by this operator: 2, 4 or 6 (2 per block in the block column)
static int asm_slots_per_col(const sgc_asm *a) {
    int nb[3] = {0, 0, 0};
    for (auto &B : a->blocks) nb[B.nu]++;
    const int m = std::max(nb[0], std::max(nb[1], nb[2]));
    if (m > 3) return -1;
    return 2 * std::max(1, m);
}

template <typename T, typename IDX, typename R, int MODE, int NCOLS, int E>
static int asm_launch(const sgc_asm *a, int index_base, const int64_t *colptr_in, int64_t *colptr_out, int64_t *rowval,
                      T *nzval, cudaStream_t st) {
    AsmParams<T> p = asm_params<T>(a, index_base);
    if (a->ncta == 0) return SGC_OK;
    const bool needs_smem = (MODE == ASM_FILL || MODE == ASM_FILL_ALL);
    const size_t smem = needs_smem ? (sizeof(T) + sizeof(R)) * (size_t)ASM_BLOCK * NCOLS * E : 0;
    auto kern = k_asm_group<T, IDX, R, ASM_BLOCK, MODE, NCOLS, E>;
    if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)a->ncta, ASM_BLOCK, smem, st>>>(p, a->cta_total, a->cta_offset, colptr_in, colptr_out, rowval, nzval);
    return check_launch("k_asm_group");
}

template <typename T, typename IDX, typename R, int MODE>
static int asm_shape_dispatch(const sgc_asm *a, int index_base, const int64_t *ci, int64_t *co, int64_t *rv, T *nz, cudaStream_t st) {
    const int ncols = a->cmpfirst ? a->Kin : 1;
    const int e = asm_slots_per_col(a);
    if (e < 0 || ncols * e > 12)
        return fail(SGC_ERR_INVALID, "operator needs %d columns x %d entry slots per group (limit 12)", ncols, e);
#define ASM_CASE(NC, EE) if (ncols == NC && e == EE) return asm_launch<T, IDX, R, MODE, NC, EE>(a, index_base, ci, co, rv, nz, st);
    ASM_CASE(1, 2) ASM_CASE(1, 4) ASM_CASE(1, 6) ASM_CASE(2, 2) ASM_CASE(2, 4) ASM_CASE(2, 6) ASM_CASE(3, 2) ASM_CASE(3, 4)
#undef ASM_CASE
    return fail(SGC_ERR_INVALID, "unsupported assembly shape");
}

template <int MODE>
static int asm_dispatch(const sgc_asm *a, int index_base, const int64_t *colptr_in, int64_t *colptr_out, int64_t *rowval,
                        void *nzval, cudaStream_t st) {
    const int64_t ncol = (int64_t)a->Kin * a->M, nrow = (int64_t)a->Kout * a->M;
    const bool small = ncol < (int64_t)0x7fffffffLL && nrow < (int64_t)0x7fffffffLL;  // 32-bit cell / row arithmetic
    if (a->cx) return small ? asm_shape_dispatch<cplx, uint32_t, int32_t, MODE>(a, index_base, colptr_in, colptr_out, rowval, (cplx *)nzval, st)
                            : asm_shape_dispatch<cplx, uint64_t, int64_t, MODE>(a, index_base, colptr_in, colptr_out, rowval, (cplx *)nzval, st);
    return small ? asm_shape_dispatch<double, uint32_t, int32_t, MODE>(a, index_base, colptr_in, colptr_out, rowval, (double *)nzval, st)
                 : asm_shape_dispatch<double, uint64_t, int64_t, MODE>(a, index_base, colptr_in, colptr_out, rowval, (double *)nzval, st);
}

// pass 1: per-CTA totals + their exclusive scan; leaves nnz in the handle (synchronises)
static int asm_count(sgc_asm *a, cudaStream_t st) {
    if (a->counted) return SGC_OK;
    REQUIRE(a->blocks.size() <= (size_t)ASM_MAX_BLOCKS, "too many blocks");
    const int64_t ngroups = a->cmpfirst ? a->M : (int64_t)a->Kin * a->M;
    const int64_t ncta = (ngroups + ASM_BLOCK - 1) / ASM_BLOCK;
    REQUIRE(ncta < (int64_t)0x7fffffff, "matrix too large for one launch");
    if (ncta != a->ncta || !a->cta_total) {
        if (a->cta_total) cudaFree(a->cta_total);
        if (a->cta_offset) cudaFree(a->cta_offset);
        a->cta_total = a->cta_offset = nullptr;
        CUDA_TRY(cudaMalloc((void **)&a->cta_total, sizeof(int64_t) * (size_t)(ncta + 1)));
        CUDA_TRY(cudaMalloc((void **)&a->cta_offset, sizeof(int64_t) * (size_t)(ncta + 1)));
        a->ncta = ncta;
    }
    int s;
    if (a->row_begin == 0 && a->row_end == (int64_t)a->Kout * a->M) {  // all rows kept: separable count
        AsmCountTables t;
        for (int nu = 0; nu < 3; ++nu)
            for (int ax = 0; ax < 3; ++ax) t.A[nu][ax] = a->count_tables + a->count_off[nu][ax];
        const int cpg = a->cmpfirst ? a->Kin : 1;
        const bool small = (int64_t)a->Kin * a->M < (int64_t)0x7fffffffLL;
        if (small)
            k_asm_count_fast<uint32_t, ASM_BLOCK><<<(unsigned)ncta, ASM_BLOCK, 0, st>>>(t, cpg, a->Kin, a->M, ngroups, (uint32_t)a->N[0], (uint32_t)a->N[1], a->cta_total);
        else
            k_asm_count_fast<uint64_t, ASM_BLOCK><<<(unsigned)ncta, ASM_BLOCK, 0, st>>>(t, cpg, a->Kin, a->M, ngroups, (uint64_t)a->N[0], (uint64_t)a->N[1], a->cta_total);
        s = check_launch("k_asm_count_fast");
    } else {
        s = asm_dispatch<ASM_COUNT>(a, 0, nullptr, nullptr, nullptr, nullptr, st);
    }
    if (s) return s;
    CUDA_TRY(cudaMemsetAsync(a->cta_total + ncta, 0, sizeof(int64_t), st));
    size_t need = 0;
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, need, a->cta_total, a->cta_offset, (int)(ncta + 1), st));
    if (need > a->scan_tmp_bytes) {
        if (a->scan_tmp) cudaFree(a->scan_tmp);
        a->scan_tmp = nullptr;
        CUDA_TRY(cudaMalloc(&a->scan_tmp, need));
        a->scan_tmp_bytes = need;
    }
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(a->scan_tmp, need, a->cta_total, a->cta_offset, (int)(ncta + 1), st));
    sgc_count_launches(2);
    int64_t total = 0;
    CUDA_TRY(cudaMemcpyAsync(&total, a->cta_offset + ncta, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    a->nnz = total;
    a->counted = true;
    return SGC_OK;
}

extern "C" int sgc_asm_count(sgc_asm *a, int64_t *nnz, void *stream) {
    REQUIRE(a, "handle is NULL");
    int s = asm_count(a, (cudaStream_t)stream);
    if (s) return s;
    if (nnz) *nnz = a->nnz;
    return SGC_OK;
}

extern "C" int sgc_asm_fill_all(sgc_asm *a, int64_t *colptr, int index_base, int64_t *rowval, void *nzval, void *stream) {
    REQUIRE(a && colptr, "NULL argument");
    REQUIRE(index_base == 0 || index_base == 1, "index_base must be 0 or 1");
    int s = asm_count(a, (cudaStream_t)stream);
    if (s) return s;
    REQUIRE(a->nnz == 0 || (rowval && nzval), "rowval / nzval are NULL");
    if (a->ncta == 0) {  // empty matrix: colptr[0] only
        int64_t b = index_base;
        CUDA_TRY(cudaMemcpyAsync(colptr, &b, sizeof b, cudaMemcpyHostToDevice, (cudaStream_t)stream));
        return SGC_OK;
    }
    return asm_dispatch<ASM_FILL_ALL>(a, index_base, nullptr, colptr, rowval, nzval, (cudaStream_t)stream);
}

extern "C" int sgc_asm_colptr(sgc_asm *a, int64_t *colptr, int index_base, int64_t *nnz, void *stream) {
    REQUIRE(a && colptr, "NULL argument");
    REQUIRE(index_base == 0 || index_base == 1, "index_base must be 0 or 1");
    int s = asm_count(a, (cudaStream_t)stream);
    if (s) return s;
    if (nnz) *nnz = a->nnz;
    return asm_dispatch<ASM_COLPTR>(a, index_base, nullptr, colptr, nullptr, nullptr, (cudaStream_t)stream);
}

extern "C" int sgc_asm_fill(const sgc_asm *a, const int64_t *colptr, int index_base, int64_t *rowval,
                            void *nzval, void *stream) {
    REQUIRE(a && colptr, "NULL argument");
    REQUIRE(index_base == 0 || index_base == 1, "index_base must be 0 or 1");
    REQUIRE(a->counted, "call sgc_asm_colptr (or sgc_asm_count) before sgc_asm_fill");
    return asm_dispatch<ASM_FILL>(a, index_base, colptr, nullptr, rowval, nzval, (cudaStream_t)stream);
}

extern "C" int sgc_asm_nnz_host(sgc_asm *a, int64_t *nnz) {
    REQUIRE(a && nnz, "NULL argument");
    return sgc_asm_count(a, nnz, nullptr);
}

extern "C" int sgc_asm_fill_host(sgc_asm *a, int64_t *colptr, int index_base, int64_t *rowval, void *nzval) {
    REQUIRE(a && colptr, "NULL argument");
    const int64_t ncol = (int64_t)a->Kin * a->M;
    int64_t nnz = 0;
    int s = sgc_asm_count(a, &nnz, nullptr);
    if (s) return s;
    if (!a->d_colptr) CUDA_TRY(cudaMalloc((void **)&a->d_colptr, sizeof(int64_t) * (size_t)(ncol + 1)));
    const size_t es = a->cx ? 16 : 8;
    int64_t *d_row = nullptr;
    void *d_val = nullptr;
    CUDA_TRY(cudaMalloc((void **)&d_row, sizeof(int64_t) * (size_t)std::max<int64_t>(nnz, 1)));
    if (cudaMalloc(&d_val, es * (size_t)std::max<int64_t>(nnz, 1)) != cudaSuccess) {
        cudaFree(d_row);
        return fail(SGC_ERR_NOMEM, "out of device memory for nzval");
    }
    s = sgc_asm_fill_all(a, a->d_colptr, index_base, d_row, d_val, nullptr);
    if (!s) {
        cudaError_t e1 = cudaMemcpy(colptr, a->d_colptr, sizeof(int64_t) * (size_t)(ncol + 1), cudaMemcpyDeviceToHost);
        cudaError_t e2 = nnz ? cudaMemcpy(rowval, d_row, sizeof(int64_t) * (size_t)nnz, cudaMemcpyDeviceToHost) : cudaSuccess;
        cudaError_t e3 = nnz ? cudaMemcpy(nzval, d_val, es * (size_t)nnz, cudaMemcpyDeviceToHost) : cudaSuccess;
        if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) s = fail(SGC_ERR_CUDA, "device→host copy of the CSC arrays failed");
    }
    cudaFree(d_row);
    cudaFree(d_val);
    return s;
}

extern "C" int sgc_csc_matvec(int64_t m, int64_t n, const int64_t *colptr, const int64_t *rowval,
                              const void *nzval, int index_base, int val_dtype, const void *x, void *y,
                              void *stream) {
    REQUIRE(colptr && rowval && nzval && x && y, "NULL argument");
    (void)m;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t es = val_dtype == SGC_C128 ? 16 : 8;
    CUDA_TRY(cudaMemsetAsync(y, 0, es * (size_t)m, st));
    const unsigned blocks = (unsigned)std::min<int64_t>((n + 255) / 256, 148 * 16);
    if (n == 0) return SGC_OK;
    if (val_dtype == SGC_C128)
        k_csc_matvec<cplx><<<blocks, 256, 0, st>>>(n, colptr, rowval, (const cplx *)nzval, index_base, (const cplx *)x, (cplx *)y);
    else
        k_csc_matvec<double><<<blocks, 256, 0, st>>>(n, colptr, rowval, (const double *)nzval, index_base, (const double *)x, (double *)y);
    return check_launch("k_csc_matvec");
}
