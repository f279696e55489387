// sgc_curlcurl.cu — launch planning and C-ABI entry points of the fused curl-of-curl
// (sgc_curlcurl.cuh): G OP C₂(C₁ F) in one pass, H never written to memory.
#include "sgc_internal.h"

#include <algorithm>
#include <cmath>
#include <type_traits>

#include "sgc_curlcurl.cuh"
#include "sgc_cc4.cuh"
#include "sgc_cc6.cuh"

using namespace sgc;

namespace {

#ifdef SGC_SWEEP_BUILD
constexpr bool kCcExperiments = true;
#else
constexpr bool kCcExperiments = false;
#endif

// ---- fourth generation (sgc_cc4.cuh): TMA-fed, warp-specialised ----------------------------------------------------
// cuTensorMapEncodeTiled comes from the driver; the library links the runtime only, so the entry point is looked up.
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
    static std::atomic<void *> cached{nullptr};
    void *f = cached.load(std::memory_order_acquire);
    if (!f) {
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        cached.store(f, std::memory_order_release);
    }
    return (EncodeTiledFn)f;
}

// one ComplexF64 component array (Nx, Ny, Nz) as a 3-D tensor of Float64 pairs; box = one (TX+2) x (TY+2) tile plane
int encode_component(CUtensorMap *tm, const cplx *base, int Nx, int Ny, int Nz, int box_x, int box_y) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return sgc_fail(SGC_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[3] = {(cuuint64_t)2 * Nx, (cuuint64_t)Ny, (cuuint64_t)Nz};
    const cuuint64_t strides[2] = {(cuuint64_t)Nx * 16, (cuuint64_t)Nx * Ny * 16};  // bytes, dims 1 and 2
    const cuuint32_t box[3] = {(cuuint32_t)2 * box_x, (cuuint32_t)box_y, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void *)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return sgc_fail(SGC_ERR_CUDA, "cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
    return SGC_OK;
}

// the whole ComplexF64 field (Nx, Ny, Nz, 3) as a 4-D tensor of Float64 pairs; box = (TX+2) x (TY+2) x 2 planes x 3 components
int encode_field4(CUtensorMap *tm, const cplx *base, int Nx, int Ny, int Nz, int box_x, int box_y) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return sgc_fail(SGC_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[4] = {(cuuint64_t)2 * Nx, (cuuint64_t)Ny, (cuuint64_t)Nz, 3};
    const cuuint64_t strides[3] = {(cuuint64_t)Nx * 16, (cuuint64_t)Nx * Ny * 16, (cuuint64_t)Nx * Ny * Nz * 16};
    const cuuint32_t box[4] = {(cuuint32_t)2 * box_x, (cuuint32_t)box_y, 2, 3};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, (void *)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return sgc_fail(SGC_ERR_CUDA, "cuTensorMapEncodeTiled (4-D) failed (CUresult %d)", (int)r);
    return SGC_OK;
}

template <typename C, bool ADD, int TY, int SE, int SH, bool FZ1, bool FZ2>
int launch_cc4_one(const CurlCurl4Params<C> &P, dim3 grid, cudaStream_t st) {
    constexpr size_t smem = cc4_smem_bytes(TY, SE, SH);
    auto kern = k_cc4<C, ADD, TY, SE, SH, FZ1, FZ2>;
    static std::atomic<uint64_t> attr_devices{0};
    CUDA_TRY(sgc_set_dynamic_smem(kern, smem, attr_devices));
    kern<<<grid, dim3(32 * (2 * TY + 3), 1, 1), smem, st>>>(P);
    return SGC_OK;
}
template <typename C, bool ADD, int TY, int SE, int SH>
int launch_cc4_dirs(const CurlCurl4Params<C> &P, dim3 grid, cudaStream_t st) {
    if (P.q.fwd1[2]) return P.q.fwd2[2] ? launch_cc4_one<C, ADD, TY, SE, SH, true, true>(P, grid, st)
                                        : launch_cc4_one<C, ADD, TY, SE, SH, true, false>(P, grid, st);
    return P.q.fwd2[2] ? launch_cc4_one<C, ADD, TY, SE, SH, false, true>(P, grid, st)
                       : launch_cc4_one<C, ADD, TY, SE, SH, false, false>(P, grid, st);
}
template <typename C, bool ADD>
int launch_cc4_tiles(const CurlCurl4Params<C> &P, int ty, int se, int sh, dim3 grid, cudaStream_t st) {
#ifdef SGC_SWEEP_BUILD
#define CC4_CASE(TY, SE, SH) if (ty == TY && se == SE && sh == SH) return launch_cc4_dirs<C, ADD, TY, SE, SH>(P, grid, st);
    CC4_CASE(6, 4, 3)
    CC4_CASE(8, 4, 3)
#undef CC4_CASE
#endif
    (void)P; (void)grid; (void)st;
    return sgc_fail(SGC_ERR_INVALID, "fused curl-of-curl (TMA): tile 32x%d with %d/%d stages is not built", ty, se, sh);
}

template <typename C, bool ADD, int TY, int SE, bool FZ1, bool FZ2>
int launch_cc5_one(const CurlCurl4Params<C> &P, dim3 grid, cudaStream_t st) {
    constexpr size_t smem = cc5_smem_bytes(TY, SE);
    auto kern = k_cc5<C, ADD, TY, SE, FZ1, FZ2>;
    static std::atomic<uint64_t> attr_devices{0};
    CUDA_TRY(sgc_set_dynamic_smem(kern, smem, attr_devices));
    kern<<<grid, dim3(32 * (TY + 1), 1, 1), smem, st>>>(P);
    return SGC_OK;
}
template <typename C, bool ADD, int TY, int SE>
int launch_cc5_dirs(const CurlCurl4Params<C> &P, dim3 grid, cudaStream_t st) {
    if (P.q.fwd1[2]) return P.q.fwd2[2] ? launch_cc5_one<C, ADD, TY, SE, true, true>(P, grid, st)
                                        : launch_cc5_one<C, ADD, TY, SE, true, false>(P, grid, st);
    return P.q.fwd2[2] ? launch_cc5_one<C, ADD, TY, SE, false, true>(P, grid, st)
                       : launch_cc5_one<C, ADD, TY, SE, false, false>(P, grid, st);
}
template <typename C, bool ADD>
int launch_cc5_tiles(const CurlCurl4Params<C> &P, int ty, int se, dim3 grid, cudaStream_t st) {
#ifdef SGC_SWEEP_BUILD
#define CC5_CASE(TY, SE) if (ty == TY && se == SE) return launch_cc5_dirs<C, ADD, TY, SE>(P, grid, st);
    CC5_CASE(7, 3)
    CC5_CASE(7, 4)
    CC5_CASE(5, 4)
#undef CC5_CASE
#endif
    (void)P; (void)grid; (void)st;
    return sgc_fail(SGC_ERR_INVALID, "fused curl-of-curl (v5): tile 32x%d with %d stages is not built", ty, se);
}

// ---- sixth generation (sgc_cc6.cuh): two planes per barrier interval; C₁ forward / C₂ backward in z only ----------------
template <typename C, bool ADD, int TY, int NP, int MINB, int SP>
int launch_cc6_one(const CurlCurl4Params<C> &P, dim3 grid, cudaStream_t st) {
    constexpr size_t smem = cc6_smem_bytes(TY, NP);
    auto kern = k_cc6<C, ADD, TY, NP, MINB, SP>;
    static std::atomic<uint64_t> attr_devices{0};
    CUDA_TRY(sgc_set_dynamic_smem(kern, smem, attr_devices));
    kern<<<grid, dim3(32 * (TY + 1), 1, 1), smem, st>>>(P);
    return SGC_OK;
}
// sp: steady state of boundary-free tiles — 0 two-phase code, 1 one straight-line block (H points first), 2 outputs first
template <typename C, bool ADD>
int launch_cc6_tiles(const CurlCurl4Params<C> &P, int ty, int np, int sp, dim3 grid, cudaStream_t st) {
#define CC6_CASE(TY, NP, MINB, SP) if (ty == TY && np == NP && sp == SP) return launch_cc6_one<C, ADD, TY, NP, MINB, SP>(P, grid, st);
    CC6_CASE(5, 3, 2, 2)
#ifdef SGC_SWEEP_BUILD
    CC6_CASE(5, 3, 2, 1)
    CC6_CASE(5, 3, 2, 0)
    CC6_CASE(11, 3, 1, 2)
    CC6_CASE(4, 4, 2, 2)
#endif
#undef CC6_CASE
    return sgc_fail(SGC_ERR_INVALID, "fused curl-of-curl (v6): tile 32x%d with %d pair stages, pipelining %d is not built", ty, np, sp);
}

// z-march that fills whole waves of one-CTA-per-SM: among the chunk counts with a march of at least 16 planes, the
// one whose last wave is fullest (ties: the longer march, i.e. fewer halo planes)
int cc4_march(int64_t tiles, int64_t nk) {
    int sms = 148;
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    double best = -1.0;
    int best_march = (int)std::min<int64_t>(nk, 64);
    for (int64_t chunks = 1; chunks <= nk; ++chunks) {
        const int64_t march = (nk + chunks - 1) / chunks;
        if (march < 16 && chunks > 1) break;
        const int64_t ctas = tiles * ((nk + march - 1) / march);
        const double waves = (double)ctas / sms;
        const double eff = waves / std::ceil(waves) * ((double)march / (march + 2));
        if (eff > best + 1e-9) { best = eff; best_march = (int)march; }
    }
    return best_march;
}

// z-march of the sixth generation (two CTAs per SM): an even number of planes between 24 and 48 whose last wave is
// fullest, weighted by the z-halo overhead (m + 2 planes computed for m written); measured flat within 1.5 % over 20 … 52
int cc6_march(int64_t tiles, int64_t nk) {
    if (nk <= 48) return (int)nk;
    int sms = 148;
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    double best = -1.0;
    int best_march = 32;
    for (int m = 24; m <= 48; m += 2) {
        const double waves = (double)(tiles * ((nk + m - 1) / m)) / (2.0 * sms);
        const double eff = waves / std::ceil(waves) * ((double)m / (m + 2));
        if (eff > best + 1e-9) { best = eff; best_march = m; }
    }
    return best_march;
}

template <typename T, typename C, bool ADD, int TX, int TY, int SE, int MINB>
int launch_cc_one(const CurlCurlParams<T, C> &p, dim3 grid, cudaStream_t st) {
    constexpr size_t smem = ((size_t)SE * 3 * (TX + 2) * (TY + 2) + (size_t)3 * 3 * (TX + 1) * (TY + 1)) * sizeof(T);
    auto kern = k_curlcurl3<T, C, ADD, TX, TY, SE, MINB>;
    static std::atomic<uint64_t> attr_devices{0};  // per instantiation: the devices the attribute has been set on
    CUDA_TRY(sgc_set_dynamic_smem(kern, smem, attr_devices));
    kern<<<grid, dim3(TX, TY, 1), smem, st>>>(p);
    return SGC_OK;
}

template <typename T, typename C, bool ADD>
int launch_cc_tiles(const CurlCurlParams<T, C> &p, int tx, int ty, int se, dim3 grid, cudaStream_t st) {
#define CC_CASE(TX, TY, SE, MINB) \
    if (tx == TX && ty == TY && se == SE) return launch_cc_one<T, C, ADD, TX, TY, SE, MINB>(p, grid, st);
    CC_CASE(32, 8, 3, 2)
#ifdef SGC_SWEEP_BUILD
    CC_CASE(32, 4, 4, 3)
    CC_CASE(64, 4, 3, 2)
#endif
#undef CC_CASE
    return sgc_fail(SGC_ERR_INVALID, "fused curl-of-curl: tile %dx%d with %d stages is not built", tx, ty, se);
}

// version 2 (warp roles): tile 32 x TY, compile-time z-directions
template <typename T, typename C, bool ADD, int TY, int SE, bool FZ1, bool FZ2>
int launch_ccw_one(const CurlCurlParams<T, C> &p, dim3 grid, cudaStream_t st) {
    constexpr size_t smem = ((size_t)SE * 3 * 34 * (TY + 2) + (size_t)3 * 3 * 33 * (TY + 1)) * sizeof(T);
    auto kern = k_curlcurl3w<T, C, ADD, TY, SE, FZ1, FZ2>;
    static std::atomic<uint64_t> attr_devices{0};  // per instantiation: the devices the attribute has been set on
    CUDA_TRY(sgc_set_dynamic_smem(kern, smem, attr_devices));
    kern<<<grid, dim3(32, TY + 1, 1), smem, st>>>(p);
    return SGC_OK;
}
template <typename T, typename C, bool ADD, int TY, int SE>
int launch_ccw_dirs(const CurlCurlParams<T, C> &p, dim3 grid, cudaStream_t st) {
    if (p.fwd1[2]) return p.fwd2[2] ? launch_ccw_one<T, C, ADD, TY, SE, true, true>(p, grid, st)
                                    : launch_ccw_one<T, C, ADD, TY, SE, true, false>(p, grid, st);
    return p.fwd2[2] ? launch_ccw_one<T, C, ADD, TY, SE, false, true>(p, grid, st)
                     : launch_ccw_one<T, C, ADD, TY, SE, false, false>(p, grid, st);
}
template <typename T, typename C, bool ADD>
int launch_ccw_tiles(const CurlCurlParams<T, C> &p, int ty, int se, dim3 grid, cudaStream_t st) {
#define CCW_CASE(TY, SE) if (ty == TY && se == SE) return launch_ccw_dirs<T, C, ADD, TY, SE>(p, grid, st);
    CCW_CASE(7, 3)
    CCW_CASE(7, 4)
    CCW_CASE(4, 4)
#ifdef SGC_SWEEP_BUILD
    CCW_CASE(8, 3)
    CCW_CASE(4, 3)
#endif
#undef CCW_CASE
    return sgc_fail(SGC_ERR_INVALID, "fused curl-of-curl: tile 32x%d with %d stages is not built", ty, se);
}

struct AxisInfo {
    int fwd[3], bloch[3], e_real[3];
    cplx e[3];
};
AxisInfo axis_info(const sgc_op *op) {
    AxisInfo a{};
    for (int d = 0; d < 3; ++d) { a.fwd[d] = 0; a.bloch[d] = 1; a.e[d] = cplx{1, 0}; a.e_real[d] = 1; }
    for (const Term &t : op->terms)
        if (t.kind == KIND_PARTIAL) { a.fwd[t.axis] = t.fwd; a.bloch[t.axis] = t.bloch; a.e[t.axis] = t.e; a.e_real[t.axis] = t.e_real; }
    return a;
}

template <typename C> C narrow_c(cplx v);
template <> double narrow_c<double>(cplx v) { return v.re; }
template <> cplx narrow_c<cplx>(cplx v) { return v; }

template <typename T, typename C>
int launch_cc(const sgc_op *c2, const sgc_op *c1, T *G, const T *F, bool add_mode, const void *const *ghost_lo,
              const void *const *ghost_hi, int64_t k_begin, int64_t k_end, cudaStream_t st) {
    if (k_end <= k_begin) return SGC_OK;
    CurlCurlParams<T, C> p{};
    const int64_t M = c1->M;
    const AxisInfo a1 = axis_info(c1), a2 = axis_info(c2);
    for (int d = 0; d < 3; ++d) {
        p.F[d] = F + d * M;
        p.G[d] = G + d * M;
        p.c1[d] = (const C *)((const unsigned char *)c1->arena + c1->fused_off[d]);
        p.c2[d] = (const C *)((const unsigned char *)c2->arena + c2->fused_off[d]);
        p.e1[d] = a1.e[d]; p.e1_real[d] = a1.e_real[d]; p.fwd1[d] = a1.fwd[d];
        p.e2[d] = a2.e[d]; p.e2_real[d] = a2.e_real[d]; p.fwd2[d] = a2.fwd[d];
        p.bc1_lo[d] = p.bc1_hi[d] = a1.bloch[d] ? BC_BLOCH : BC_SYM;
        p.bc2_lo[d] = p.bc2_hi[d] = a2.bloch[d] ? BC_BLOCH : BC_SYM;
        p.ghost_lo[d] = ghost_lo ? (const T *)ghost_lo[d] : nullptr;
        p.ghost_hi[d] = ghost_hi ? (const T *)ghost_hi[d] : nullptr;
    }
    if (c1->lo_iface) { p.bc1_lo[2] = BC_IFACE; p.bc2_lo[2] = BC_IFACE; }
    if (c1->hi_iface) { p.bc1_hi[2] = BC_IFACE; p.bc2_hi[2] = BC_IFACE; }
    p.c1z_lo = narrow_c<C>(c1->cz_ghost[0]);
    p.c1z_hi = narrow_c<C>(c1->cz_ghost[1]);
    p.Nx = (int)c1->N[0]; p.Ny = (int)c1->N[1]; p.Nz = (int)c1->N[2];
    p.k_begin = (int)k_begin; p.k_end = (int)k_end;
    // variant bits 0-3: 0 = default (sixth generation where it applies, else the third), 2 = third generation (warp roles),
    // 1 = first generation (two-pass H positions), 4 / 5 / 6 = k_cc4 / k_cc5 / k_cc6 where they apply
    const int version = c2->variant & 0xF;
    const int tx = (version == 1 && c2->tile_x) ? c2->tile_x : 32;
    const int ty = c2->tile_y ? c2->tile_y : (version == 1 ? 8 : 7);  // 32x7 outputs + 1 halo row = 8 warps, 2 CTAs/SM at 128 registers
    const int se = ((c2->variant >> 4) & 0xF) ? ((c2->variant >> 4) & 0xF) : (sizeof(T) == 8 && version != 1 ? 4 : 3);  // Float64: 3 CTAs/SM, deeper ring
    const int64_t nk = k_end - k_begin;
    p.march = c2->march > 0 ? c2->march : (nk >= 256 ? 64 : 32);  // z-halo overhead (n+2)/n planes vs. number of CTAs
    dim3 grid((unsigned)((p.Nx + tx - 1) / tx), (unsigned)((p.Ny + ty - 1) / ty), (unsigned)((nk + p.march - 1) / p.march));
    if (grid.y > 65535 || grid.z > 65535) return sgc_fail(SGC_ERR_INVALID, "grid too large for the fused curl-of-curl tiling");
    if ((int64_t)p.Nx * p.Ny >= (int64_t)0x7fffffff) return sgc_fail(SGC_ERR_INVALID, "plane too large for the fused curl-of-curl");
    // version 4: the TMA-fed warp-specialised experiment (sgc_cc4.cuh) where it applies — ComplexF64 fields, a grid that
    // holds a whole tile, no Bloch wrap in x / y and no slab ghost planes (the copy engine cannot fetch those cells);
    // everything else, and version 4 where it does not apply, runs the third generation
    if constexpr (std::is_same<T, cplx>::value) {
        int ty4 = (c2->tile_y == 8) ? 8 : 6;
        const bool wraps_xy = a1.bloch[0] || a1.bloch[1] || a2.bloch[0] || a2.bloch[1];
        const bool ghosts = (ghost_lo && ghost_lo[0]) || (ghost_hi && ghost_hi[0]) || c1->lo_iface || c1->hi_iface;
        // sixth generation (sgc_cc6.cuh) — the DEFAULT where it applies: all-symmetric box, C₁ forward / C₂ backward in z,
        // no slab ghost planes, a grid that holds a whole tile
        const bool v6_ok = !wraps_xy && !ghosts && !a1.bloch[2] && !a2.bloch[2] && a1.fwd[2] && !a2.fwd[2] && p.Nx >= 34;
        if (v6_ok && (version == 6 || (version == 0 && (c2->tile_y == 0 || c2->tile_y == 5)))) {
            const int ty6 = c2->tile_y ? c2->tile_y : 5;
            const int np6 = ((c2->variant >> 4) & 0xF) ? ((c2->variant >> 4) & 0xF) : 3;
            if (p.Ny >= ty6 + 2) {
                CurlCurl4Params<C> P{};
                P.q = p;
                int s6 = encode_field4(&P.tm[0], p.F[0], p.Nx, p.Ny, p.Nz, 34, ty6 + 2);
                if (!s6) {
                    const int64_t tiles6 = (int64_t)((p.Nx + 31) / 32) * ((p.Ny + ty6 - 1) / ty6);
                    P.q.march = c2->march > 0 ? std::min<int>(c2->march, CC6_MAXMARCH) : cc6_march(tiles6, nk);
                    dim3 g6((unsigned)((p.Nx + 31) / 32), (unsigned)((p.Ny + ty6 - 1) / ty6), (unsigned)((nk + P.q.march - 1) / P.q.march));
                    if (g6.y > 65535 || g6.z > 65535) return sgc_fail(SGC_ERR_INVALID, "grid too large for the fused curl-of-curl tiling");
                    // variant bits 8-11, steady state of boundary-free tiles: 0 = default (one straight-line block, outputs first),
                    // 1 = two-phase code only, 2 = one block, H points first
                    const int vsp = (c2->variant >> 8) & 0xF;
                    const int sp6 = vsp == 1 ? 0 : (vsp == 2 ? 1 : 2);
                    s6 = add_mode ? launch_cc6_tiles<C, true>(P, ty6, np6, sp6, g6, st) : launch_cc6_tiles<C, false>(P, ty6, np6, sp6, g6, st);
                    if (s6) return s6;
                    return sgc_check_launch("k_cc6");
                }
                // no tensor-map support in this driver: fall through to the third generation
            }
        }
        // (k_cc4 / k_cc5 are measured experiments, built with `make EXTRA=-DSGC_SWEEP_BUILD`; the product library runs the third
        // generation when they are asked for)
        if (kCcExperiments && version == 5 && !wraps_xy && !ghosts && p.Nx >= 34 && p.Ny >= 9) {
            const int ty5 = (c2->tile_y == 5) ? 5 : 7;
            const int se5 = (((c2->variant >> 4) & 0xF) == 3 && ty5 == 7) ? 3 : 4;
            CurlCurl4Params<C> P{};
            P.q = p;
            int s5 = SGC_OK;
            for (int d = 0; d < 3 && !s5; ++d) s5 = encode_component(&P.tm[d], p.F[d], p.Nx, p.Ny, p.Nz, 34, ty5 + 2);
            if (!s5) {
                P.q.march = c2->march > 0 ? c2->march : (nk >= 256 ? 64 : 32);
                dim3 g5((unsigned)((p.Nx + 31) / 32), (unsigned)((p.Ny + ty5 - 1) / ty5), (unsigned)((nk + P.q.march - 1) / P.q.march));
                if (g5.y > 65535 || g5.z > 65535) return sgc_fail(SGC_ERR_INVALID, "grid too large for the fused curl-of-curl tiling");
                s5 = add_mode ? launch_cc5_tiles<C, true>(P, ty5, se5, g5, st) : launch_cc5_tiles<C, false>(P, ty5, se5, g5, st);
                if (s5) return s5;
                return sgc_check_launch("k_cc5");
            }
        }
        if (kCcExperiments && version == 4 && !wraps_xy && !ghosts && p.Nx >= 34 && p.Ny >= ty4 + 2) {
            CurlCurl4Params<C> P{};
            P.q = p;
            int s4 = SGC_OK;
            for (int d = 0; d < 3 && !s4; ++d) s4 = encode_component(&P.tm[d], p.F[d], p.Nx, p.Ny, p.Nz, 34, ty4 + 2);
            if (!s4) {
                const int64_t tiles = (int64_t)((p.Nx + 31) / 32) * ((p.Ny + ty4 - 1) / ty4);
                P.q.march = c2->march > 0 ? c2->march : cc4_march(tiles, nk);
                dim3 g4((unsigned)((p.Nx + 31) / 32), (unsigned)((p.Ny + ty4 - 1) / ty4), (unsigned)((nk + P.q.march - 1) / P.q.march));
                if (g4.y > 65535 || g4.z > 65535) return sgc_fail(SGC_ERR_INVALID, "grid too large for the fused curl-of-curl tiling");
                s4 = add_mode ? launch_cc4_tiles<C, true>(P, ty4, 4, 3, g4, st) : launch_cc4_tiles<C, false>(P, ty4, 4, 3, g4, st);
                if (s4) return s4;
                return sgc_check_launch("k_cc4");
            }
            // no tensor-map support in this driver: fall through to the third generation
        }
    }
    int s;
    if (version == 4 || version == 5 || version == 6) {  // k_cc4 / k_cc5 / k_cc6 asked for but not applicable (Float64 field, tiny grid): third generation with its own defaults
        const int ty3 = 7, se3 = sizeof(T) == 8 ? 4 : 3;
        p.march = nk >= 256 ? 64 : 32;
        dim3 g3((unsigned)((p.Nx + 31) / 32), (unsigned)((p.Ny + ty3 - 1) / ty3), (unsigned)((nk + p.march - 1) / p.march));
        s = add_mode ? launch_ccw_tiles<T, C, true>(p, ty3, se3, g3, st) : launch_ccw_tiles<T, C, false>(p, ty3, se3, g3, st);
        if (s) return s;
        return sgc_check_launch("k_curlcurl3");
    }
    if (version == 1) s = add_mode ? launch_cc_tiles<T, C, true>(p, tx, ty, se, grid, st) : launch_cc_tiles<T, C, false>(p, tx, ty, se, grid, st);
    else s = add_mode ? launch_ccw_tiles<T, C, true>(p, ty, se, grid, st) : launch_ccw_tiles<T, C, false>(p, ty, se, grid, st);
    if (s) return s;
    return sgc_check_launch("k_curlcurl3");
}

}  // namespace

extern "C" int sgc_op_set_slab_coef(sgc_op *op, const double *dzinv_below, const double *dzinv_above, int coef_complex) {
    REQUIRE(op, "op is NULL");
    REQUIRE(op->fused_curl, "slab coefficients apply to full 3-D curl operators");
    const void *v[2] = {dzinv_below, dzinv_above};
    for (int s = 0; s < 2; ++s) {
        HV d = v[s] ? vec_at(v[s], coef_complex, 0) : hv(1.0);
        HV c = hmul(op->alpha, d);  // (α * ∆z⁻¹) exactly as the tables
        if (c.cx && !op->fused_cc) return sgc_fail(SGC_ERR_INVALID, "complex ghost coefficient for an operator with real tables");
        op->cz_ghost[s] = to_cplx(c);
    }
    return SGC_OK;
}

extern "C" int sgc_op_apply_curlcurl_range(const sgc_op *c2, const sgc_op *c1, void *G, const void *F, int opmode,
                                           const void *const *ghost_lo, const void *const *ghost_hi,
                                           int64_t k_begin, int64_t k_end, void *stream) {
    REQUIRE(c1 && c2 && G && F, "NULL argument");
    REQUIRE(opmode == SGC_OP_SET || opmode == SGC_OP_ADD, "opmode must be SGC_OP_SET or SGC_OP_ADD");
    REQUIRE(c1->fused_curl && c2->fused_curl, "the fused curl-of-curl needs two full 3-D curl operators");
    REQUIRE(c1->dtype == c2->dtype, "the two curls must have the same field type");
    for (int d = 0; d < 3; ++d) REQUIRE(c1->N[d] == c2->N[d], "the two curls must act on the same grid");
    REQUIRE(c1->fused_cc == c2->fused_cc, "the two curls must both have real or both have complex coefficient tables");
    REQUIRE(c1->lo_iface == c2->lo_iface && c1->hi_iface == c2->hi_iface, "the two curls must share the slab interfaces");
    REQUIRE(0 <= k_begin && k_begin <= k_end && k_end <= c1->N[2], "plane range [%lld,%lld) outside 0..%lld",
            (long long)k_begin, (long long)k_end, (long long)c1->N[2]);
    REQUIRE(G != F, "G and F must not alias");
    const bool any_ghost = (ghost_lo && ghost_lo[0]) || (ghost_hi && ghost_hi[0]);
    if (ghost_lo && ghost_lo[0]) REQUIRE(ghost_lo[1] && ghost_lo[2], "ghost_lo needs all three components");
    if (ghost_hi && ghost_hi[0]) REQUIRE(ghost_hi[1] && ghost_hi[2], "ghost_hi needs all three components");
    if (c1->lo_iface || c1->hi_iface || any_ghost) {
        int fz1 = 0, fz2 = 0;
        for (const Term &t : c1->terms) if (t.kind == KIND_PARTIAL && t.axis == 2) fz1 = t.fwd;
        for (const Term &t : c2->terms) if (t.kind == KIND_PARTIAL && t.axis == 2) fz2 = t.fwd;
        REQUIRE(fz1 != fz2, "slab interfaces need opposite z-directions in the two curls (halo reach of one plane)");
        if (c1->lo_iface) REQUIRE(ghost_lo && ghost_lo[0] && ghost_lo[1] && ghost_lo[2], "ghost_lo planes are required at a lower interface");
        if (c1->hi_iface) REQUIRE(ghost_hi && ghost_hi[0] && ghost_hi[1] && ghost_hi[2], "ghost_hi planes are required at an upper interface");
    }
    cudaStream_t st = (cudaStream_t)stream;
    const bool add = opmode == SGC_OP_ADD;
    if (c1->dtype == SGC_F64)
        return launch_cc<double, double>(c2, c1, (double *)G, (const double *)F, add, ghost_lo, ghost_hi, k_begin, k_end, st);
    if (c1->fused_cc)
        return launch_cc<cplx, cplx>(c2, c1, (cplx *)G, (const cplx *)F, add, ghost_lo, ghost_hi, k_begin, k_end, st);
    return launch_cc<cplx, double>(c2, c1, (cplx *)G, (const cplx *)F, add, ghost_lo, ghost_hi, k_begin, k_end, st);
}

extern "C" int sgc_op_apply_curlcurl(const sgc_op *c2, const sgc_op *c1, void *G, const void *F, int opmode, void *stream) {
    REQUIRE(c1 && c2, "NULL argument");
    return sgc_op_apply_curlcurl_range(c2, c1, G, F, opmode, nullptr, nullptr, 0, c1->N[2], stream);
}
