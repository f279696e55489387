// sgc_grid.cu — device-side producers of the per-axis vectors the hot path consumes: the PML-stretched
// spacings s∆l = s(l)·∆l of create_stretched_∆l / create_sfactor / calc_sfactor (src/pml.jl:94-153) and their
// inverses (invert_∆l, src/util.jl:42-48), and the refresh of an operator handle's coefficient tables from them —
// so that a frequency sweep (ω changes ⇒ new s∆l ⇒ new α·∆l⁻¹ tables) never round-trips through the host.
//
// calc_sfactor is  s = κ(d) + σ(d) / (a(d) + iω)  with the polynomial profiles σ = σmax (d/L)^m, κ = 1 + (κmax−1)(d/L)^m,
// a = amax (1 − d/L)^ma.  Only the last expression depends on ω.  The profiles are tabulated ONCE per axis on the host
// (glibc log / pow, correctly rounded to within 1 ulp like Julia's) and kept on the device; the per-ω kernel then needs
// +, ×, / and Julia's complex inverse only, all of which are exactly reproducible (sgc_common.cuh).
#include "sgc_internal.h"

#include <cmath>

#include "sgc_kernels.cuh"

using namespace sgc;

struct sgc_axis {
    int64_t n = 0;
    int device = 0;
    double *prof = nullptr;  // device: σ[n], κ[n], a[n], ∆l[n], inside-PML flag[n]
    cplx *scratch = nullptr; // device: (s∆l)⁻¹ of the last sgc_op_set_axes
};

namespace {

__global__ void k_axis_stretch(int64_t n, const double *__restrict__ prof, double omega, cplx *__restrict__ sdl, cplx *__restrict__ inv) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double sig = prof[i], kap = prof[n + i], a = prof[2 * n + i], dl = prof[3 * n + i];
    cplx s{1.0, 0.0};                                    // ones(ComplexF, N)                            pml.jl:123
    if (prof[4 * n + i] != 0.0) {
        const cplx iz = cinv(cplx{a + 0.0, omega});      // σ / (a + im*ω) = σ * inv(a + im*ω)           pml.jl:152, Base `/(::Real, ::Complex)`
        s = cplx{kap + sig * iz.re, sig * iz.im};        // κ + …
    }
    const cplx v{s.re * dl, s.im * dl};                  // sfactor .* ∆l                                pml.jl:108-109
    if (sdl) sdl[i] = v;
    if (inv) {                                           // 1 ./ v = 1 * inv(v)                          util.jl:44-45
        const cplx w = cinv(v);
        inv[i] = cplx{1.0 * w.re, 1.0 * w.im};
    }
}

// c[i] = al · v[i]   (the tables of a ∂ term: (α * ∆w⁻¹[i]), partial.jl:56; real·complex scales both parts)
__global__ void k_scale_table(int64_t n, cplx *__restrict__ dst, const cplx *__restrict__ v, cplx al, int al_real) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const cplx x = v[i];
    dst[i] = al_real ? cplx{al.re * x.re, al.re * x.im} : mul(al, x);
}

}  // namespace

extern "C" int sgc_axis_create(sgc_axis **ax_out, int64_t n, const double *l, const double *dl, double lpml_neg, double lpml_pos,
                               double Lpml_neg, double Lpml_pos, const double *pml) {
    REQUIRE(ax_out && l && dl && n >= 1, "bad arguments");
    *ax_out = nullptr;
    const double m = pml ? pml[0] : 4.0, R = pml ? pml[1] : std::exp(-16.0), kmax = pml ? pml[2] : 1.0, amax = pml ? pml[3] : 0.0,
                 ma = pml ? pml[4] : 4.0;  // PMLParam(), pml.jl:66-73
    std::vector<double> h(5 * (size_t)n, 0.0);
    for (int64_t i = 0; i < n; ++i) {
        h[n + i] = 1.0;
        h[3 * n + i] = dl[i];
        // r < lpml₋ first, then r > lpml₊ (the later assignment wins, pml.jl:130-131)
        for (int side = 0; side < 2; ++side) {
            const bool in = side == 0 ? (l[i] < lpml_neg) : (l[i] > lpml_pos);
            if (!in) continue;
            const double d = side == 0 ? lpml_neg - l[i] : l[i] - lpml_pos;
            const double L = side == 0 ? Lpml_neg : Lpml_pos;
            const double smax = -(m + 1) * std::log(R) / (2 * L);   // pml.jl:147
            h[i] = smax * std::pow(d / L, m);                        // σ
            h[n + i] = 1 + (kmax - 1) * std::pow(d / L, m);          // κ
            h[2 * n + i] = amax * std::pow(1 - d / L, ma);           // a
            h[4 * n + i] = 1.0;
        }
    }
    sgc_axis *ax = new sgc_axis();
    ax->n = n;
    cudaError_t e = cudaGetDevice(&ax->device);
    if (e == cudaSuccess) e = cudaMalloc((void **)&ax->prof, sizeof(double) * h.size());
    if (e == cudaSuccess) e = cudaMalloc((void **)&ax->scratch, sizeof(cplx) * (size_t)n);
    if (e == cudaSuccess) e = cudaMemcpy(ax->prof, h.data(), sizeof(double) * h.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        if (ax->prof) cudaFree(ax->prof);
        if (ax->scratch) cudaFree(ax->scratch);
        delete ax;
        return sgc_fail(SGC_ERR_CUDA, "sgc_axis_create: %s", cudaGetErrorString(e));
    }
    *ax_out = ax;
    return SGC_OK;
}

extern "C" int sgc_axis_destroy(sgc_axis *ax) {
    if (!ax) return SGC_OK;
    cudaFree(ax->prof);
    cudaFree(ax->scratch);
    delete ax;
    return SGC_OK;
}

extern "C" int sgc_axis_stretch(const sgc_axis *ax, double omega, void *sdl_device, void *sdl_inv_device, void *stream) {
    REQUIRE(ax, "axis is NULL");
    REQUIRE(sdl_device || sdl_inv_device, "no output requested");
    const unsigned blocks = (unsigned)((ax->n + 255) / 256);
    k_axis_stretch<<<blocks, 256, 0, (cudaStream_t)stream>>>(ax->n, ax->prof, omega, (cplx *)sdl_device, (cplx *)sdl_inv_device);
    return sgc_check_launch("k_axis_stretch");
}

// New coefficient tables for a new ω, built on the device: for every ∂ term of the operator along array dimension d,
// c[i] = (parity·α) · (s∆l_d[i])⁻¹, and the tables of the fused kernels likewise.  The operator must have been created
// with ComplexF64 coefficient tables of the same lengths (PML-stretched ∆l is complex).
extern "C" int sgc_op_set_axes(sgc_op *op, const sgc_axis *const *axes, double omega, void *stream) {
    REQUIRE(op && axes, "NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    for (int d = 0; d < op->ndim; ++d) {
        REQUIRE(axes[d], "axis %d is NULL", d);
        REQUIRE(axes[d]->n == op->N[d], "axis %d has %lld points, the operator's grid %lld", d, (long long)axes[d]->n, (long long)op->N[d]);
    }
    for (const Term &t : op->terms)
        if (t.kind >= 0) REQUIRE(t.kind == KIND_PARTIAL && t.cc, "sgc_op_set_axes refreshes ∂ / curl / divergence / gradient operators with ComplexF64 coefficient tables");
    for (int d = 0; d < op->ndim; ++d) {
        const sgc_axis *ax = axes[d];
        const unsigned blocks = (unsigned)((ax->n + 255) / 256);
        k_axis_stretch<<<blocks, 256, 0, st>>>(ax->n, ax->prof, omega, nullptr, ax->scratch);
        { int s = sgc_check_launch("k_axis_stretch"); if (s) return s; }
        for (const Term &t : op->terms) {
            if (t.kind != KIND_PARTIAL || t.axis != d) continue;
            k_scale_table<<<blocks, 256, 0, st>>>(ax->n, (cplx *)((unsigned char *)op->arena + t.off_c), ax->scratch, to_cplx(t.al), t.al.cx ? 0 : 1);
            { int s = sgc_check_launch("k_scale_table"); if (s) return s; }
        }
        if (op->fused_curl) {
            REQUIRE(op->fused_cc, "the fused curl tables of this operator are real");
            k_scale_table<<<blocks, 256, 0, st>>>(ax->n, (cplx *)((unsigned char *)op->arena + op->fused_off[d]), ax->scratch, to_cplx(op->alpha),
                                                  op->alpha.cx ? 0 : 1);
            { int s = sgc_check_launch("k_scale_table"); if (s) return s; }
        }
    }
    return SGC_OK;
}
