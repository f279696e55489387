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
    void *arena = nullptr;            // device: value tables, count tables, prefix tables (one allocation)
    mutable cudaStream_t last_stream = nullptr;  // stream of the last fill (the arena is freed in its order)
    std::vector<size_t> off0, offs;
    void *scan_tmp = nullptr;
    size_t scan_tmp_bytes = 0;
    int64_t nnz = -1;
    int64_t *cta_total = nullptr, *cta_offset = nullptr;  // per-CTA entry counts and their exclusive scan
    int64_t ncta = 0;
    bool counted = false;
    int *count_tables = nullptr;              // device: A[nu][axis][N_axis] of the separable count
    size_t count_off[3][3] = {};              // element offsets into count_tables
    int64_t *prefix_tables = nullptr;         // device: exclusive prefix sums of the count tables, sets 0..2 = nu, 3 = all
    size_t prefix_off[4][3] = {};             // element offsets into prefix_tables
    int64_t set_total[4] = {0, 0, 0, 0};      // entries of each set when all rows are kept
    bool force_count_pass = false;            // testing: use the count kernel + device scan even when the closed form applies
    bool no_fast = false;                     // testing: interior cells take the general entry path too
    bool no_cells = false;                    // testing: one column per thread in the warp-granular kernel even where several cells fit
    bool no_warp = false;                     // testing: the CTA-granular kernel (block scan) even where the closed form applies
    // column partition (sgc_asm_set_col_range): the block of columns [col_begin, col_end) as a stand-alone CSC matrix
    int64_t col_begin = 0, col_end = -1;      // -1: all columns
    int64_t cta_first = 0;                    // first CTA (of the whole-matrix launch) of the column range
    int64_t entry_base = 0;                   // entries stored before col_begin in the whole matrix
    std::vector<int64_t> h_prefix;            // host copy of prefix_tables (closed-form offsets on the host)
    // host convenience
    int64_t *d_colptr = nullptr;
};

// ---- matrix-free operator handle ----------------------------------------------------------------
struct Term {
    int out, in;      // component positions in the last array dimension of G / F
    int kind;         // KIND_*  or -1 for "Gv .= 0"
    int axis;         // array dimension (0-based)
    int fwd, bloch;
    sgc::cplx e;
    int e_real;
    bool cc;          // coefficient tables are complex
    size_t off_c, off_w;  // offsets of the tables in the arena (bytes); w unused unless weighted
    sgc::cplx a2, a1, b0, w_hi, w_lo;  // scalars (stored as complex; narrowed when !cc)
    HV al;            // parity·α of a ∂ term (sgc_op_set_axes rebuilds the table c[i] = al·∆w⁻¹[i] on the device)
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
    int fused_pat = -1;           // PAT_* when the term list is one of the one-launch compositions
    bool fused_mean = false;      // apply_mean! on a 3-D grid: one launch of the staged z-march (k_curl3p MODE = MEAN)
    int lo_iface = 0, hi_iface = 0;
    // neighbour synchronisation folded into the fused curl kernel (sgc_op_attach_comm, sgc_comm.cu)
    long long *sync_ctrl = nullptr;                  // this rank's control block
    long long *sync_post[2] = {nullptr, nullptr};    // the neighbours' flag slots (peer addresses)
    int sync_side[2] = {0, 0};
    mutable bool sync_on = false;                    // set around the launch by sgc_op_apply_synced
    // weighted means on a slab: ∆w of the neighbour's plane below / above (sgc_op_set_slab_weights)
    bool has_slab_w = false;
    HV slab_w[2] = {{1.0, 0.0, false}, {1.0, 0.0, false}};
    HV alpha = {1.0, 0.0, false};                    // α of a curl operator (for the ghost-plane coefficients)
    sgc::cplx cz_ghost[2] = {{1.0, 0.0}, {1.0, 0.0}};  // α·∆z⁻¹ of the planes −1 / Nz (slab neighbours)
    int tile_x = 0, tile_y = 0, march = 0, variant = 0;
    void *arena = nullptr;        // device
    // host-array entry points (sgc_host.cu): device staging buffers, three streams, events — created on first use
    mutable struct sgc_hostpipe *hp = nullptr;
};
void sgc_hostpipe_destroy(struct sgc_hostpipe *hp);

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


// translation-unit boundaries: the fused 3-D curl lives in sgc_curl.cu (explicitly instantiated
// for the three (field, coefficient) type pairs), the assembly in sgc_asm.cu
template <typename T, typename C>
int sgc_launch_curl3(const sgc_op *op, T *G, const T *F, const void *const *ghost, bool add_mode,
                     int64_t k_begin, int64_t k_end, cudaStream_t st);
template <typename T, typename C>
int sgc_launch_mean3(const sgc_op *op, T *G, const T *F, bool add_mode, int64_t k_begin, int64_t k_end, cudaStream_t st);
template <typename T, typename C>
int sgc_launch_divgrad3(const sgc_op *op, bool is_divg, T *G, const T *F, const void *const *ghost, bool add_mode,
                        int64_t k_begin, int64_t k_end, cudaStream_t st);
int sgc_check_common(int dtype, int ndim, const int64_t *N);

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute: remember, per kernel instantiation, on
// which devices it has been set (bit d of `done`).  Lock-free; a racing second call is harmless.
template <typename K>
static inline cudaError_t sgc_set_dynamic_smem(K kern, size_t smem, std::atomic<uint64_t> &done) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const uint64_t bit = 1ull << (dev & 63);
    if (done.load(std::memory_order_acquire) & bit) return cudaSuccess;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) done.fetch_or(bit, std::memory_order_release);
    return e;
}
#define fail sgc_fail
#define check_launch sgc_check_launch
