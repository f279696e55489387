// sgc_asm.cu — C-ABI entry points and launch planning of the sparse-operator assembly
// (create_∂, create_m̂, create_mean, create_curl, create_divg, create_grad, create_πcmp →
// SparseMatrixCSC; reference src/matrix/).  Kernels: sgc_assembly.cuh.
#include "sgc_internal.h"

#include <algorithm>
#include <cub/device/device_scan.cuh>

#include "sgc_assembly.cuh"

using namespace sgc;

static const int CURL_PARITY[3][3] = {{0, -1, 1}, {1, 0, -1}, {-1, 1, 0}};  // [nv][nu], curl.jl:46-48
static int check_common(int dtype, int ndim, const int64_t *N) { return sgc_check_common(dtype, ndim, N); }

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
    if (a->arena) {  // stream-ordered: after the last kernel that reads the tables, without a device-wide synchronisation
        if (cudaFreeAsync(a->arena, a->last_stream) != cudaSuccess) cudaGetLastError();
    }
    free_checked(a->scan_tmp, "scan scratch");
    free_checked(a->d_colptr, "host-path colptr");
    free_checked(a->cta_total, "cta totals");
    free_checked(a->cta_offset, "cta offsets");
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

// One stream-ordered allocation and one upload per handle: [V₀/Vₛ value tables | count tables | prefix tables].
// (cudaMalloc / cudaFree each synchronise the device; against a sub-millisecond fill kernel they dominated create_*.)
static int finish_asm(sgc_asm *a) {
    OpBuilder b;
    for (auto &B : a->blocks) {
        a->off0.push_back(b.push(B.V0, a->cx));
        a->offs.push_back(b.push(B.Vs, a->cx));
    }
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
    // exclusive prefix sums of the count tables (sets 0..2: one block column; set 3: all of them) for
    // the closed-form entry offsets of the fill kernel
    std::vector<int64_t> pre;
    for (int set = 0; set < 4; ++set) {
        int64_t S[3];
        for (int ax = 0; ax < 3; ++ax) {
            a->prefix_off[set][ax] = pre.size();
            int64_t run = 0;
            pre.push_back(0);
            for (int64_t i = 0; i < a->N[ax]; ++i) {
                for (int nu = 0; nu < 3; ++nu)
                    if (set == 3 || set == nu) run += tab[a->count_off[nu][ax] + (size_t)i];
                pre.push_back(run);
            }
            S[ax] = run;
        }
        a->set_total[set] = a->N[1] * a->N[2] * S[0] + a->N[0] * a->N[2] * S[1] + a->N[0] * a->N[1] * S[2];
    }
    a->h_prefix = pre;
    const size_t off_tab = (b.host.size() + 15) & ~size_t(15);
    const size_t off_pre = (off_tab + sizeof(int) * tab.size() + 15) & ~size_t(15);
    const size_t bytes = off_pre + sizeof(int64_t) * pre.size();
    b.host.resize(bytes);
    memcpy(b.host.data() + off_tab, tab.data(), sizeof(int) * tab.size());
    memcpy(b.host.data() + off_pre, pre.data(), sizeof(int64_t) * pre.size());
    CUDA_TRY(cudaMallocAsync(&a->arena, bytes, (cudaStream_t)0));
    CUDA_TRY(cudaMemcpyAsync(a->arena, b.host.data(), bytes, cudaMemcpyHostToDevice, (cudaStream_t)0));
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)0));  // (the staging vector dies with this call)
    a->count_tables = (int *)((unsigned char *)a->arena + off_tab);
    a->prefix_tables = (int64_t *)((unsigned char *)a->arena + off_pre);
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
    REQUIRE(a->col_end < 0 || (row_begin == 0 && row_end == (int64_t)a->Kout * a->M), "a row range cannot be combined with a column range");
    a->row_begin = row_begin;
    a->row_end = row_end;
    a->nnz = -1;
    a->counted = false;
    return SGC_OK;
}

// entries stored before group g of the whole matrix (host evaluation of prefix_entries_before)
static int64_t host_entries_before(const sgc_asm *a, int64_t g) {
    const int64_t ngroups = a->cmpfirst ? a->M : (int64_t)a->Kin * a->M;
    if (g >= ngroups) return a->set_total[3];
    int set = 3;
    int64_t c = g, base = 0;
    if (!(a->cmpfirst && a->Kin > 1) && a->Kin != 1) {
        set = (int)(g / a->M);
        c = g - (int64_t)set * a->M;
        for (int q = 0; q < set; ++q) base += a->set_total[q];
    }
    const int64_t Nx = a->N[0], Ny = a->N[1];
    const int64_t t = c / Nx, i = c - t * Nx, k = t / Ny, j = t - k * Ny;
    const int64_t *PA = a->h_prefix.data() + a->prefix_off[set][0], *PB = a->h_prefix.data() + a->prefix_off[set][1],
                  *PC = a->h_prefix.data() + a->prefix_off[set][2];
    const int64_t SA = PA[Nx], SB = PB[Ny];
    const int64_t bj = PB[j + 1] - PB[j], ck = PC[k + 1] - PC[k];
    return base + k * (Ny * SA + Nx * SB) + Nx * Ny * PC[k] + j * SA + Nx * PB[j] + j * Nx * ck + PA[i] + i * (bj + ck);
}

static int64_t asm_col_align(const sgc_asm *a) { return (int64_t)128 * (a->cmpfirst ? a->Kin : 1); }  // columns per CTA

extern "C" int sgc_asm_col_partition(const sgc_asm *a, int part, int nparts, int64_t *col_begin, int64_t *col_end) {
    REQUIRE(a && col_begin && col_end, "NULL argument");
    REQUIRE(nparts >= 1 && part >= 0 && part < nparts, "part %d outside 0..%d", part, nparts - 1);
    const int64_t ncol = (int64_t)a->Kin * a->M, al = asm_col_align(a);
    const int64_t ncta = (ncol + al - 1) / al;
    *col_begin = std::min(ncol, (ncta * part / nparts) * al);
    *col_end = std::min(ncol, (ncta * (part + 1) / nparts) * al);
    return SGC_OK;
}

extern "C" int sgc_asm_set_col_range(sgc_asm *a, int64_t col_begin, int64_t col_end) {
    REQUIRE(a, "handle is NULL");
    const int64_t ncol = (int64_t)a->Kin * a->M, al = asm_col_align(a);
    REQUIRE(0 <= col_begin && col_begin <= col_end && col_end <= ncol, "column range out of bounds");
    REQUIRE(a->row_begin == 0 && a->row_end == (int64_t)a->Kout * a->M, "a column range cannot be combined with a row range");
    REQUIRE(col_begin % al == 0 || col_begin == ncol, "col_begin must be a multiple of %lld columns (use sgc_asm_col_partition)", (long long)al);
    REQUIRE(col_end % al == 0 || col_end == ncol, "col_end must be a multiple of %lld columns or the last column", (long long)al);
    a->col_begin = col_begin;
    a->col_end = (col_begin == 0 && col_end == ncol) ? -1 : col_end;
    a->nnz = -1;
    a->counted = false;
    return SGC_OK;
}

extern "C" int sgc_asm_set_tuning(sgc_asm *a, int force_count_pass) {
    REQUIRE(a, "handle is NULL");
    a->force_count_pass = (force_count_pass & 1) != 0;
    a->no_fast = (force_count_pass & 2) != 0;
    a->no_warp = (force_count_pass & 4) != 0;
    a->no_cells = (force_count_pass & 8) != 0;
    a->nnz = -1;
    a->counted = false;
    return SGC_OK;
}

extern "C" int sgc_asm_shape(const sgc_asm *a, int64_t *m, int64_t *n, int *val_dtype) {
    REQUIRE(a, "handle is NULL");
    if (m) *m = a->row_end - a->row_begin;
    if (n) *n = a->col_end < 0 ? (int64_t)a->Kin * a->M : a->col_end - a->col_begin;
    if (val_dtype) *val_dtype = a->cx ? SGC_C128 : SGC_F64;
    return SGC_OK;
}

static int asm_slots_per_col(const sgc_asm *a);

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
    for (int nu = 0; nu < 3; ++nu)
        for (int ax = 0; ax < 3; ++ax) p.cnt_tab[nu][ax] = a->count_tables + a->count_off[nu][ax];
    p.skip_enabled = 0;
    if (a->cmpfirst && !(a->row_begin == 0 && a->row_end == (int64_t)a->Kout * a->M)) {
        const int64_t span = a->M / a->N[a->ndim - 1];  // cells per hyper-plane of the slowest axis = longest reach
        const int64_t c0 = a->row_begin / a->Kout, c1 = (a->row_end + a->Kout - 1) / a->Kout;
        p.skip_enabled = 1;
        p.near_lo = std::max<int64_t>(c0 - span, 0);
        p.near_hi = std::min<int64_t>(c1 + span, a->M);
        p.wrap_span = span;
    }
    p.prefix_enabled = (a->row_begin == 0 && a->row_end == (int64_t)a->Kout * a->M && a->prefix_tables && !a->force_count_pass) ? 1 : 0;
    p.cta_first = 0; p.g_end = p.ngroups; p.col_base = 0; p.entry_base = 0;
    if (a->col_end >= 0) {
        p.cta_first = a->cta_first;
        p.g_end = a->col_end / p.cols_per_group;
        p.col_base = a->col_begin;
        p.entry_base = a->entry_base;
    }
    // interior fast path: per block column the entries in ascending row order (valid where nothing wraps or drops)
    p.fast_enabled = 0;
    {
        const int E = asm_slots_per_col(a);
        bool ok = p.prefix_enabled && E > 0 && !a->no_fast;
        for (int nu = 0; nu < a->Kin && ok; ++nu) ok = (2 * p.nblk_of[nu] == E);
        for (int b = 0; b < p.nblk && ok; ++b) ok = (a->blocks[b].ns != 0) && (a->N[a->blocks[b].axis] >= 3);
        if (ok) {
            const int64_t stride[3] = {1, a->N[0], a->N[0] * a->N[1]};
            for (int nu = 0; nu < a->Kin; ++nu) {
                struct Ent { int64_t key; const T *V; int axis, dsub; };
                std::vector<Ent> ents;
                for (int q = 0; q < p.nblk_of[nu]; ++q) {
                    const AsmBlock<T> &B = p.blk[p.blk_of[nu][q]];
                    const int64_t delta = -(int64_t)B.ns * stride[B.axis];
                    auto key = [&](int64_t d) { return a->cmpfirst ? (int64_t)a->Kout * d + B.nv : d + a->M * (int64_t)B.nv; };
                    ents.push_back({key(0), B.V0, B.axis, 0});
                    ents.push_back({key(delta), B.Vs, B.axis, -B.ns});
                }
                std::stable_sort(ents.begin(), ents.end(), [](const Ent &x, const Ent &y) { return x.key < y.key; });
                for (size_t e = 0; e + 1 < ents.size(); ++e) ok = ok && (ents[e].key != ents[e + 1].key);
                for (size_t e = 0; e < ents.size() && e < (size_t)ASM_MAX_ENTRIES; ++e)
                    p.fast[nu][e] = AsmFastEntry<T>{ents[e].V, ents[e].key, ents[e].axis, ents[e].dsub};
            }
            for (int d = 0; d < 3; ++d) {
                p.in_lo[d] = a->N[d] >= 3 ? 1u : 0u;
                p.in_span[d] = a->N[d] >= 3 ? (unsigned)(a->N[d] - 3) : (unsigned)(a->N[d] - 1);
            }
            p.fast_enabled = ok ? 1 : 0;
        }
    }
    int64_t before = 0;
    for (int set = 0; set < 4; ++set) {
        for (int ax = 0; ax < 3; ++ax) p.prefix[set][ax] = a->prefix_tables + a->prefix_off[set][ax];
        p.set_base[set] = (set < 3) ? before : 0;
        if (set < 3) before += a->set_total[set];
    }
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
    if (MODE == ASM_FILL_ALL && p.prefix_enabled && !a->no_warp) {
        // closed-form offsets: warp-granular kernel (no CTA-wide barrier, no block scan)
        if constexpr (NCOLS == 1) {
            // one column per group (block ordering, create_∂, create_m̂, create_grad): a thread takes CPT consecutive cells of
            // the row so that the offset arithmetic is shared by CPT·E entries (whole matrix only, Nx a multiple of CPT)
            constexpr int CPT = (E == 2) ? 4 : 2;
            if (a->col_end < 0 && a->N[0] % CPT == 0 && !a->no_cells) {
                const size_t smem = (sizeof(T) + sizeof(R)) * (size_t)ASM_BLOCK * (CPT * E + 1);
                auto kern = k_asm_warp<T, IDX, R, ASM_BLOCK, CPT, E, true>;
                if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                const int64_t nthreads = p.ngroups / CPT;
                kern<<<(unsigned)((nthreads + ASM_BLOCK - 1) / ASM_BLOCK), ASM_BLOCK, smem, st>>>(p, colptr_out, rowval, nzval);
                return check_launch("k_asm_warp<cells>");
            }
        }
        const size_t smem = (sizeof(T) + sizeof(R)) * (size_t)ASM_BLOCK * (NCOLS * E + 1);
        auto kern = k_asm_warp<T, IDX, R, ASM_BLOCK, NCOLS, E, false>;
        if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<(unsigned)a->ncta, ASM_BLOCK, smem, st>>>(p, colptr_out, rowval, nzval);
        return check_launch("k_asm_warp");
    }
    const bool needs_smem = (MODE == ASM_FILL || MODE == ASM_FILL_ALL);
    const size_t smem = needs_smem ? (sizeof(T) + sizeof(R)) * (size_t)ASM_BLOCK * (NCOLS * E + 1) : 0;
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
    if (a->col_end >= 0) {
        // a block of columns of the whole matrix: same closed form, evaluated at both ends of the range
        REQUIRE(a->prefix_tables && !a->force_count_pass, "a column range needs the closed-form offsets");
        const int cpg = a->cmpfirst ? a->Kin : 1;
        const int64_t gb = a->col_begin / cpg, ge = a->col_end / cpg;
        a->cta_first = gb / ASM_BLOCK;
        a->ncta = (ge - gb + ASM_BLOCK - 1) / ASM_BLOCK;
        a->entry_base = host_entries_before(a, gb);
        a->nnz = host_entries_before(a, ge) - a->entry_base;
        a->counted = true;
        return SGC_OK;
    }
    if (a->row_begin == 0 && a->row_end == (int64_t)a->Kout * a->M && a->prefix_tables && !a->force_count_pass) {
        // all rows kept: the count is separable and the fill kernel derives its offsets in closed form
        // (AsmParams::prefix) — no count kernel, no scan, no read-back, no per-CTA arrays
        a->ncta = ncta;
        a->nnz = a->set_total[3];
        a->counted = true;
        return SGC_OK;
    }
    if (ncta != a->ncta || !a->cta_total) {
        if (a->cta_total) cudaFree(a->cta_total);
        if (a->cta_offset) cudaFree(a->cta_offset);
        a->cta_total = a->cta_offset = nullptr;
        CUDA_TRY(cudaMalloc((void **)&a->cta_total, sizeof(int64_t) * (size_t)(ncta + 1)));
        CUDA_TRY(cudaMalloc((void **)&a->cta_offset, sizeof(int64_t) * (size_t)(ncta + 1)));
        a->ncta = ncta;
    }
    int s;
    if (a->row_begin == 0 && a->row_end == (int64_t)a->Kout * a->M && a->prefix_tables && !a->force_count_pass) {
        // all rows kept: the count is separable and the fill kernel derives its offsets in closed form
        // (AsmParams::prefix) — no count kernel, no scan, no read-back
        a->nnz = a->set_total[3];
        a->counted = true;
        return SGC_OK;
    }
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
    a->last_stream = (cudaStream_t)stream;
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
    a->last_stream = (cudaStream_t)stream;
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
    a->last_stream = (cudaStream_t)stream;
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
