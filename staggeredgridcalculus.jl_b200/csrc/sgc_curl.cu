// sgc_curl.cu — launch planning of the fused full 3-D curl kernels (sgc_kernels.cuh: k_curl3p,
// k_curl3) for apply_curl! (matrixfree/differential/curl.jl:52-137).  Own translation unit so
// that the many tile / ring-depth instantiations compile in parallel with the rest.
#include "sgc_internal.h"

#include <algorithm>

#include "sgc_kernels.cuh"

using namespace sgc;

template <typename T, typename C, bool ADD>
static void launch_curl3_tiles(const CurlParams<T, C> &p, int tx, int ty, dim3 grid_xy_z, cudaStream_t st);
template <typename T, typename C, bool ADD>
static int launch_curl3p_tiles(const CurlParams<T, C> &p, int tx, int ty, int stages, int minb, bool xshfl, dim3 grid, cudaStream_t st);

template <typename T, typename C>
int sgc_launch_curl3(const sgc_op *op, T *G, const T *F, const void *const *ghost, bool add_mode,
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
    if (op->sync_on) {
        if ((op->variant & 0xF) != 0 || k_begin != 0 || k_end != op->N[2] || !op->sync_ctrl)
            return sgc_fail(SGC_ERR_INVALID, "synchronised apply needs the default kernel, the whole slab and sgc_op_attach_comm");
        if (op->sync_side[0] || op->sync_side[1]) {
            p.sync_ctrl = op->sync_ctrl;
            for (int sd = 0; sd < 2; ++sd) { p.sync_post[sd] = op->sync_post[sd]; p.sync_side[sd] = op->sync_side[sd]; }
        }
    }

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
    // x-neighbours from the shared tile (default) or by warp shuffle (bit 12).  Cold the two are within 1 %; in the sustained,
    // power-capped state the benchmark boxes settle into (≈ 1.65 of 1.97 GHz) the shared-tile form, with fewer instructions per
    // plane, is ≈ 3 % faster (tools/time_curl_xshfl.py, profiles/r2_time_curl_xshfl.jsonl)
    const bool xshfl = ((op->variant >> 12) & 1) != 0;
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
    static std::atomic<uint64_t> attr_devices{0};  // per instantiation: the devices the attribute has been set on
    CUDA_TRY(sgc_set_dynamic_smem(kern, smem, attr_devices));
    kern<<<grid, dim3(TX, TY, 1), smem, st>>>(p);
    return SGC_OK;
}

template <typename T, typename C, bool ADD>
static int launch_curl3p_tiles(const CurlParams<T, C> &p, int tx, int ty, int stages, int minb, bool xshfl, dim3 grid, cudaStream_t st) {
#define CURLP_CASE(TX, TY, S, MINB) \
    if (tx == TX && ty == TY && stages == S && minb == MINB)                                                   \
        return xshfl ? launch_curl3p_one<T, C, ADD, TX, TY, S, MINB, true>(p, grid, st)                        \
                     : launch_curl3p_one<T, C, ADD, TX, TY, S, MINB, false>(p, grid, st);
    // the product library carries the default (32x4 tiles, 6-stage ring) and three alternatives; the full tile /
    // ring-depth / occupancy matrix of the round-1 sweeps (profiles/r1_sweep_curl_256*.jsonl) builds with
    // `make EXTRA=-DSGC_SWEEP_BUILD`
    CURLP_CASE(32, 4, 6, 4)
    CURLP_CASE(32, 4, 4, 4)
    CURLP_CASE(32, 8, 4, 3)
    CURLP_CASE(64, 4, 4, 3)
#ifdef SGC_SWEEP_BUILD
    CURLP_CASE(32, 4, 5, 4)
    CURLP_CASE(32, 4, 7, 4)
    CURLP_CASE(32, 4, 8, 3)
    CURLP_CASE(32, 4, 3, 4)
    CURLP_CASE(32, 8, 3, 3)
    CURLP_CASE(32, 8, 4, 2)
    CURLP_CASE(32, 8, 6, 2)
    CURLP_CASE(32, 16, 3, 2)
    CURLP_CASE(32, 16, 4, 1)
    CURLP_CASE(64, 4, 4, 2)
    CURLP_CASE(64, 8, 3, 2)
#endif
#undef CURLP_CASE
    return sgc_fail(SGC_ERR_INVALID, "fused curl: tile %dx%d, %d stages, %d CTAs/SM is not built", tx, ty, stages, minb);
}

template <typename T, typename C, bool ADD>
static void launch_curl3_tiles(const CurlParams<T, C> &p, int tx, int ty, dim3 grid, cudaStream_t st) {
#define CURL_CASE(TX, TY) \
    if (tx == TX && ty == TY) { k_curl3<T, C, ADD, TX, TY><<<grid, dim3(TX, TY, 1), 0, st>>>(p); return; }
    CURL_CASE(32, 8)
#ifdef SGC_SWEEP_BUILD
    CURL_CASE(32, 4)
    CURL_CASE(32, 16)
    CURL_CASE(64, 4)
    CURL_CASE(64, 8)
#endif
#undef CURL_CASE
    k_curl3<T, C, ADD, 32, 8><<<dim3((p.Nx + 31) / 32, (p.Ny + 7) / 8, grid.z), dim3(32, 8, 1), 0, st>>>(p);
}


// apply_divg! / apply_grad! on 3-D grids through the same staged z-march (k_curl3p MODE = DIVG / GRAD)
template <typename T, typename C, bool ADD, int MODE>
static int launch_divgrad_one(const CurlParams<T, C> &p, dim3 grid, cudaStream_t st) {
    constexpr int TX = 32, TY = 4, S = 6, MINB = 4;
    constexpr size_t smem = (size_t)S * 3 * (TY + 1) * (TX + 1) * sizeof(T);
    auto kern = k_curl3p<T, C, ADD, TX, TY, S, MINB, true, MODE>;
    static std::atomic<uint64_t> attr_devices{0};  // per instantiation: the devices the attribute has been set on
    CUDA_TRY(sgc_set_dynamic_smem(kern, smem, attr_devices));
    kern<<<grid, dim3(TX, TY, 1), smem, st>>>(p);
    return SGC_OK;
}

template <typename T, typename C>
int sgc_launch_divgrad3(const sgc_op *op, bool is_divg, T *G, const T *F, const void *const *ghost, bool add_mode,
                        int64_t k_begin, int64_t k_end, cudaStream_t st) {
    if (k_end <= k_begin) return SGC_OK;
    CurlParams<T, C> p{};
    const int64_t M = op->M, sz = op->N[0] * op->N[1];
    const Term *tz = nullptr;
    for (const Term &t : op->terms) {  // one ∂ term per axis
        const int d = t.axis;
        p.c[d] = (const C *)((const unsigned char *)op->arena + t.off_c);
        p.e[d] = t.e; p.e_real[d] = t.e_real; p.fwd[d] = t.fwd;
        p.bc_lo[d] = p.bc_hi[d] = t.bloch ? BC_BLOCH : BC_SYM;
        if (d == 2) tz = &t;
    }
    if (!tz) return sgc_fail(SGC_ERR_INVALID, "divergence / gradient without a z term");
    if (op->lo_iface) p.bc_lo[2] = BC_IFACE;
    if (op->hi_iface) p.bc_hi[2] = BC_IFACE;
    if (is_divg) {  // slot 0 = F_z (z-differenced), slot 1 = F_x (x-differenced), slot 2 = F_y (y-differenced)
        p.F[0] = F + 2 * M; p.F[1] = F; p.F[2] = F + M;
        p.G[0] = p.G[1] = p.G[2] = G;
    } else {        // slot 0 = f
        p.F[0] = p.F[1] = p.F[2] = F;
        for (int d = 0; d < 3; ++d) p.G[d] = G + d * M;
    }
    const T *g = ghost ? (const T *)ghost[tz->in] : nullptr;
    p.ghost[0] = g ? g : (tz->fwd ? p.F[0] : p.F[0] + (op->N[2] - 1) * sz);
    p.ghost[1] = p.ghost[0];
    p.Nx = (int)op->N[0]; p.Ny = (int)op->N[1]; p.Nz = (int)op->N[2];
    p.k_begin = (int)k_begin; p.k_end = (int)k_end;
    p.march = 12;
    const int64_t nk = k_end - k_begin;
    dim3 grid((unsigned)((p.Nx + 31) / 32), (unsigned)((p.Ny + 3) / 4), (unsigned)((nk + p.march - 1) / p.march));
    if (grid.y > 65535 || grid.z > 65535) return sgc_fail(SGC_ERR_INVALID, "grid too large for the staged divergence / gradient");
    int s;
    if (is_divg) s = add_mode ? launch_divgrad_one<T, C, true, CURL3_DIVG>(p, grid, st) : launch_divgrad_one<T, C, false, CURL3_DIVG>(p, grid, st);
    else s = add_mode ? launch_divgrad_one<T, C, true, CURL3_GRAD>(p, grid, st) : launch_divgrad_one<T, C, false, CURL3_GRAD>(p, grid, st);
    if (s) return s;
    return sgc_check_launch(is_divg ? "k_curl3p<DIVG>" : "k_curl3p<GRAD>");
}

// apply_mean! on 3-D grids through the same staged z-march (k_curl3p MODE = MEAN): one launch instead of three sweeps
template <typename C> static inline C narrow_mean(cplx v);
template <> inline double narrow_mean<double>(cplx v) { return v.re; }
template <> inline cplx narrow_mean<cplx>(cplx v) { return v; }

template <typename T, typename C>
int sgc_launch_mean3(const sgc_op *op, T *G, const T *F, bool add_mode, int64_t k_begin, int64_t k_end, cudaStream_t st) {
    if (k_end <= k_begin) return SGC_OK;
    CurlParams<T, C> p{};
    const int64_t M = op->M, sz = op->N[0] * op->N[1];
    const Term *tz = nullptr;
    for (const Term &t : op->terms) {  // one m̂ term per axis, component d averaged along axis d
        const int d = t.axis;
        p.c[d] = (const C *)((const unsigned char *)op->arena + t.off_c);
        p.m[d].w = (const C *)((const unsigned char *)op->arena + t.off_w);
        p.m[d].a2 = narrow_mean<C>(t.a2); p.m[d].a1 = narrow_mean<C>(t.a1); p.m[d].b0 = narrow_mean<C>(t.b0);
        p.m[d].w_hi = narrow_mean<C>(t.w_hi); p.m[d].w_lo = narrow_mean<C>(t.w_lo);
        p.m[d].weighted = (t.kind == KIND_MEAN_WEIGHTED) ? 1 : 0;
        p.e[d] = t.e; p.e_real[d] = t.e_real; p.fwd[d] = t.fwd;
        p.bc_lo[d] = p.bc_hi[d] = t.bloch ? BC_BLOCH : BC_SYM;
        if (d == 2) tz = &t;
    }
    if (!tz) return sgc_fail(SGC_ERR_INVALID, "mean without a z term");
    // slot 0 = F_z (z-averaged through the register carry), slot 1 = F_x (x-averaged), slot 2 = F_y (y-averaged)
    p.F[0] = F + 2 * M; p.F[1] = F; p.F[2] = F + M;
    p.G[0] = G + 2 * M; p.G[1] = G; p.G[2] = G + M;
    p.ghost[0] = tz->fwd ? p.F[0] : p.F[0] + (op->N[2] - 1) * sz;  // the wrap-around plane
    p.ghost[1] = p.ghost[0];
    p.Nx = (int)op->N[0]; p.Ny = (int)op->N[1]; p.Nz = (int)op->N[2];
    p.k_begin = (int)k_begin; p.k_end = (int)k_end;
    p.march = 12;
    const int64_t nk = k_end - k_begin;
    dim3 grid((unsigned)((p.Nx + 31) / 32), (unsigned)((p.Ny + 3) / 4), (unsigned)((nk + p.march - 1) / p.march));
    if (grid.y > 65535 || grid.z > 65535) return sgc_fail(SGC_ERR_INVALID, "grid too large for the staged mean");
    int s = add_mode ? launch_divgrad_one<T, C, true, CURL3_MEAN>(p, grid, st) : launch_divgrad_one<T, C, false, CURL3_MEAN>(p, grid, st);
    if (s) return s;
    return sgc_check_launch("k_curl3p<MEAN>");
}
template int sgc_launch_mean3<double, double>(const sgc_op *, double *, const double *, bool, int64_t, int64_t, cudaStream_t);
template int sgc_launch_mean3<cplx, double>(const sgc_op *, cplx *, const cplx *, bool, int64_t, int64_t, cudaStream_t);
template int sgc_launch_mean3<cplx, cplx>(const sgc_op *, cplx *, const cplx *, bool, int64_t, int64_t, cudaStream_t);

template int sgc_launch_divgrad3<double, double>(const sgc_op *, bool, double *, const double *, const void *const *, bool, int64_t, int64_t, cudaStream_t);
template int sgc_launch_divgrad3<cplx, double>(const sgc_op *, bool, cplx *, const cplx *, const void *const *, bool, int64_t, int64_t, cudaStream_t);
template int sgc_launch_divgrad3<cplx, cplx>(const sgc_op *, bool, cplx *, const cplx *, const void *const *, bool, int64_t, int64_t, cudaStream_t);

template int sgc_launch_curl3<double, double>(const sgc_op *, double *, const double *, const void *const *, bool, int64_t, int64_t, cudaStream_t);
template int sgc_launch_curl3<cplx, double>(const sgc_op *, cplx *, const cplx *, const void *const *, bool, int64_t, int64_t, cudaStream_t);
template int sgc_launch_curl3<cplx, cplx>(const sgc_op *, cplx *, const cplx *, const void *const *, bool, int64_t, int64_t, cudaStream_t);
