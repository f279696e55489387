// sgc_fused.cuh — one-launch kernels for the small operator compositions (sm_100a).
//
//   apply_∂!     (matrixfree/differential/partial.jl:33-385, 2-D and 3-D arrays)   Gv OP ∂w Fu
//   apply_divg!  (matrixfree/differential/divergence.jl:41-66)   g   OP Σ_w ∂w F_w
//   apply_grad!  (matrixfree/differential/gradient.jl:41-59)     G_w OP ∂w f
//   2-D sub-curls of apply_curl! (matrixfree/differential/curl.jl:52-137 with cmp_shp of length 2):
//       TE   Gz = −∂yFx + ∂xFy      (cmp_out = [3], cmp_in = [1,2])
//       TM   Gx = +∂yFz, Gy = −∂xFz (cmp_out = [1,2], cmp_in = [3])
//
// The reference runs these as K separate apply_∂! sweeps; the divergence re-reads and re-writes
// g K times.  Here every input component is read once and every output component written once:
//
//   * a thread owns VEC consecutive x-elements (16 or 32 bytes) of R consecutive y-rows of one
//     z-plane; the y-neighbour rows are the thread's own registers (R+1 row loads for R outputs),
//   * the x-neighbour across the vector boundary is one extra element load that hits the L1 line
//     the lane neighbour fetches with the same instruction,
//   * the z-neighbour plane is a second coalesced stream that the 126 MB L2 serves (the plane was
//     read as the centre plane one plane-time earlier).
//
// Arithmetic: each term is the same expression, in the same operand order, as the single-term
// kernels (`term_eval`, sgc_kernels.cuh), and the terms of one output are accumulated in the
// reference's order — first step with the caller's OP, the following ones `+=` — so the result is
// bit-identical to the sweeps.
#pragma once
#include "sgc_kernels.cuh"

namespace sgc {

// A pattern is a compile-time list of (output component, input component, array axis) triples in
// the reference's execution order.  GATHER patterns have one output, SCATTER patterns one input.
struct PatDesc {
    int nt, nin, nout;
    int out[3], in[3], axis[3];
};
enum : int { PAT_GATHER_XY = 0, PAT_GATHER_YX = 1, PAT_GATHER_XYZ = 2, PAT_SCATTER_XY = 3, PAT_SCATTER_YX = 4, PAT_SCATTER_XYZ = 5,
              PAT_ONE_X = 6, PAT_ONE_Y = 7, PAT_ONE_Z = 8, PAT_COUNT = 9 };
template <int P> struct Pat;
template <> struct Pat<PAT_GATHER_XY>   { static constexpr PatDesc D{2, 2, 1, {0, 0, 0}, {0, 1, 0}, {0, 1, 0}}; };  // divg 2-D
template <> struct Pat<PAT_GATHER_YX>   { static constexpr PatDesc D{2, 2, 1, {0, 0, 0}, {0, 1, 0}, {1, 0, 0}}; };  // TE sub-curl
template <> struct Pat<PAT_GATHER_XYZ>  { static constexpr PatDesc D{3, 3, 1, {0, 0, 0}, {0, 1, 2}, {0, 1, 2}}; };  // divg 3-D
template <> struct Pat<PAT_SCATTER_XY>  { static constexpr PatDesc D{2, 1, 2, {0, 1, 0}, {0, 0, 0}, {0, 1, 0}}; };  // grad 2-D
template <> struct Pat<PAT_SCATTER_YX>  { static constexpr PatDesc D{2, 1, 2, {0, 1, 0}, {0, 0, 0}, {1, 0, 0}}; };  // TM sub-curl
template <> struct Pat<PAT_SCATTER_XYZ> { static constexpr PatDesc D{3, 1, 3, {0, 1, 2}, {0, 0, 0}, {0, 1, 2}}; };  // grad 3-D
template <> struct Pat<PAT_ONE_X>       { static constexpr PatDesc D{1, 1, 1, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}}; };  // apply_∂! along x
template <> struct Pat<PAT_ONE_Y>       { static constexpr PatDesc D{1, 1, 1, {0, 0, 0}, {0, 0, 0}, {1, 0, 0}}; };  // apply_∂! along y
template <> struct Pat<PAT_ONE_Z>       { static constexpr PatDesc D{1, 1, 1, {0, 0, 0}, {0, 0, 0}, {2, 0, 0}}; };  // apply_∂! along z

template <typename C>
struct FTerm {
    const C *c;        // (parity·α)·∆w⁻¹[i]
    cplx e;            // e⁻ⁱᵏᴸ
    int e_real;
    int fwd;
    int bc_lo, bc_hi;  // BC_BLOCH | BC_SYM | BC_IFACE (interfaces only along the last axis)
};

template <typename T, typename C>
struct FusedParams {
    const T *F[3];
    T *G[3];
    const T *ghost[3];  // per input component: the z-plane beyond the slab (wrap plane when local)
    FTerm<C> t[3];
    int64_t Nx, Ny, Nz;
    int64_t k_begin, k_end;
};

// ∂ term from register values; `nb` is the value at i+1 (forward) / i−1 (backward), RAW at a
// Bloch face or a slab interface.  Same branches and formulas as term_eval<…, KIND_PARTIAL>.
template <typename T, typename C>
SGC_D T partial_eval(const FTerm<C> &f, int64_t Nw, int64_t i, T cur, T nb) {
    if (f.fwd) {
        if (i == Nw - 1) {
            if (f.bc_hi == BC_SYM) return neg(mul(f.c[i], cur));                                   // partial.jl:84-87
            if (f.bc_hi == BC_BLOCH) return mul(f.c[i], sub(mul_e(f.e, nb, f.e_real), cur));       // :62-65
            return mul(f.c[i], sub(nb, cur));
        }
        if (i == 0 && f.bc_lo == BC_SYM) return mul(f.c[0], nb);                                   // :77-80
        return mul(f.c[i], sub(nb, cur));                                                          // :54-58
    }
    if (i == 0) {
        if (f.bc_lo == BC_SYM) return zero_of(cur);                                                // :184-188
        if (f.bc_lo == BC_BLOCH) return mul(f.c[0], sub(cur, div_e(nb, f.e, f.e_real)));           // :179-183
        return mul(f.c[0], sub(cur, nb));
    }
    return mul(f.c[i], sub(cur, nb));                                                              // :171-175
}

template <typename T, int VEC>
SGC_D void set_zero(T (&v)[VEC]) {
#pragma unroll
    for (int e = 0; e < VEC; ++e) v[e] = zero_of(T{});
}

// L1-allocating read-only loads: the x-neighbour element of a thread sits in the 128-byte line its
// lane neighbour fetches with the same instruction, so the second access is an L1 hit.
template <typename T, int VEC>
SGC_D void ldg_vec_cached(T (&v)[VEC], const T *ptr) {
    constexpr int CH = VEC * (int)sizeof(T) / 16;
    double *o = reinterpret_cast<double *>(v);
#pragma unroll
    for (int q = 0; q < CH; ++q) {
        double a, b;
        asm volatile("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(a), "=d"(b) : "l"(reinterpret_cast<const double2 *>(ptr) + q));
        o[2 * q] = a; o[2 * q + 1] = b;
    }
}
SGC_D cplx ldg_cached(const cplx *p) {
    double a, b;
    asm volatile("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(a), "=d"(b) : "l"(p));
    return cplx{a, b};
}
SGC_D double ldg_cached(const double *p) {
    double a;
    asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(a) : "l"(p));
    return a;
}
// VEC consecutive coefficients (VEC·sizeof(C) is 16 or 32 bytes for the (T, C) pairs that are built,
// or 8 bytes for one real coefficient of a complex field)
template <typename C, int VEC>
SGC_D void ld_coef(C (&c)[VEC], const C *ptr) {
#pragma unroll
    for (int e = 0; e < VEC; ++e) c[e] = ptr[e];
}

// index of the term acting along `axis` (−1: none) and its input component
constexpr int pat_term_of_axis(const PatDesc &D, int axis) {
    for (int t = 0; t < D.nt; ++t)
        if (D.axis[t] == axis) return t;
    return -1;
}
template <int P>
struct PatInfo {
    static constexpr PatDesc D = Pat<P>::D;
    static constexpr int tx = pat_term_of_axis(Pat<P>::D, 0), ty = pat_term_of_axis(Pat<P>::D, 1), tz = pat_term_of_axis(Pat<P>::D, 2);
    static constexpr int ux = tx >= 0 ? Pat<P>::D.in[tx >= 0 ? tx : 0] : 0;
    static constexpr int uy = ty >= 0 ? Pat<P>::D.in[ty >= 0 ? ty : 0] : 0;
    static constexpr int uz = tz >= 0 ? Pat<P>::D.in[tz >= 0 ? tz : 0] : 0;
    static constexpr bool has_z = tz >= 0;
};

// Thread = VEC consecutive x-elements × R consecutive y-rows of one z-plane.  Threads whose cells
// and neighbours touch no boundary take the straight-line path: every load is issued before the
// first use, every term is  (±c)·(nb − cur)  (a backward difference c·(cur − nb) equals
// (−c)·(nb − cur) bit for bit).  The other threads evaluate `partial_eval` per element.
// MODE 0: both paths in one kernel; 1: interior threads only; 2: boundary threads only (1 + 2 are
// launched back to back: the straight-line kernel then carries no boundary code or registers).
enum : int { FUSED_ALL = 0, FUSED_INTERIOR = 1, FUSED_BOUNDARY = 2 };
template <typename T, typename C, int P, bool ADD, int VEC, int R, int MINB = 1, int MODE = FUSED_ALL>
__global__ void __launch_bounds__(256, MINB) k_fused(const __grid_constant__ FusedParams<T, C> p) {
    constexpr PatDesc D = Pat<P>::D;
    constexpr bool SCATTER = (D.nin == 1);
    constexpr bool HAS_Z = PatInfo<P>::has_z;
    const int64_t Nx = p.Nx, Ny = p.Ny, Nz = p.Nz;
    const int64_t x = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    const int64_t j0 = ((int64_t)blockIdx.y * blockDim.y + threadIdx.y) * R;
    if (x >= Nx || j0 >= Ny) return;
    const int64_t plane = Nx * Ny;
    const T Z = zero_of(T{});
    const bool interior_xy = (x > 0) && (x + VEC < Nx) && (j0 > 0) && (j0 + R < Ny);

    // ---- general path helpers (boundary threads) ---------------------------------------------------
    auto x_term = [&](const FTerm<C> &f, const T *row, const T (&cur)[VEC], T (&out)[VEC]) {
        const int fwd = f.fwd;
        const int64_t nbi = fwd ? x + VEC : x - 1;
        T got = Z;
        if (nbi >= 0 && nbi < Nx) got = ldg_stream(row + nbi);
        else if ((fwd ? f.bc_hi : f.bc_lo) == BC_BLOCH) got = ldg_stream(row + (fwd ? 0 : Nx - 1));
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            const T n = fwd ? (e < VEC - 1 ? cur[e + 1 < VEC ? e + 1 : e] : got) : (e > 0 ? cur[e > 0 ? e - 1 : 0] : got);
            out[e] = partial_eval<T, C>(f, Nx, x + e, cur[e], n);
        }
    };
    // rows j0−1+fwd … j0+R−1+fwd of `base` (a plane) at column x; wrap row or zero outside the grid
    auto load_rows_y = [&](const FTerm<C> &f, const T *base, T (&V)[R + 1][VEC]) {
        const int fwd = f.fwd;
#pragma unroll
        for (int q = 0; q <= R; ++q) {
            const int64_t jj = j0 - 1 + fwd + q;
            const bool needed = (q == 0) || (j0 + q - 1 < Ny);
            if (needed && jj >= 0 && jj < Ny) ldg_vec_stream<T, VEC>(V[q], base + jj * Nx + x);
            else if (needed && jj < 0 && f.bc_lo == BC_BLOCH) ldg_vec_stream<T, VEC>(V[q], base + (Ny - 1) * Nx + x);
            else if (needed && jj >= Ny && f.bc_hi == BC_BLOCH) ldg_vec_stream<T, VEC>(V[q], base + x);
            else set_zero<T, VEC>(V[q]);
        }
    };
    auto load_row_z = [&](const FTerm<C> &f, int u, int64_t k, int64_t j, T (&nb)[VEC]) {
        const int64_t kk = f.fwd ? k + 1 : k - 1;
        if (kk >= 0 && kk < Nz) ldg_vec_stream<T, VEC>(nb, p.F[u] + kk * plane + j * Nx + x);
        else if ((f.fwd ? f.bc_hi : f.bc_lo) != BC_SYM) ldg_vec_plain<T, VEC>(nb, p.ghost[u] + j * Nx + x);
        else set_zero<T, VEC>(nb);
    };

    for (int64_t k = p.k_begin + blockIdx.z; k < p.k_end; k += gridDim.z) {
        const bool interior = interior_xy && (!HAS_Z || (k > 0 && k < Nz - 1));
        if (MODE == FUSED_INTERIOR && !interior) continue;
        if (MODE == FUSED_BOUNDARY && interior) continue;
        if (MODE != FUSED_BOUNDARY && interior) {
            // ======================= straight-line path ==============================================
            // per term: CUR rows, the neighbour data and the sign-folded coefficients, all loaded first
            T cur[D.nin][R][VEC];        // rows j0..j0+R−1 of every input component
            T ynb[R + 1][VEC];           // y-term: rows j0−1+fwd … (shares storage with cur when possible)
            T xnb[R];                    // x-term: the element beyond the vector
            T znb[R][VEC];               // z-term: rows of plane k±1
            C cx[VEC], cy[R], cz;
            T old[SCATTER ? 1 : R][VEC];
            constexpr int tx_ = PatInfo<P>::tx, ty_ = PatInfo<P>::ty, tz_ = PatInfo<P>::tz;
            // -- loads --------------------------------------------------------------------------------
            if constexpr (ty_ >= 0) {
                const FTerm<C> &f = p.t[ty_];
                const T *b = p.F[PatInfo<P>::uy] + k * plane + (j0 - 1 + f.fwd) * Nx + x;
#pragma unroll
                for (int q = 0; q <= R; ++q) {
                    if (SCATTER) ldg_vec_cached<T, VEC>(ynb[q], b + q * Nx);  // the x-term reads its edge element from these lines
                    else ldg_vec_stream<T, VEC>(ynb[q], b + q * Nx);
                }
#pragma unroll
                for (int r = 0; r < R; ++r) cy[r] = f.fwd ? f.c[j0 + r] : neg(f.c[j0 + r]);
            }
            if constexpr (tx_ >= 0) {
                const FTerm<C> &f = p.t[tx_];
                constexpr int u = PatInfo<P>::ux;
                const T *b = p.F[u] + k * plane + j0 * Nx;
                if (!(SCATTER && ty_ >= 0)) {
#pragma unroll
                    for (int r = 0; r < R; ++r) ldg_vec_cached<T, VEC>(cur[u][r], b + r * Nx + x);
                }
                const int64_t nbi = f.fwd ? x + VEC : x - 1;
#pragma unroll
                for (int r = 0; r < R; ++r) xnb[r] = ldg_cached(b + r * Nx + nbi);
                ld_coef<C, VEC>(cx, f.c + x);
                if (!f.fwd) {
#pragma unroll
                    for (int e = 0; e < VEC; ++e) cx[e] = neg(cx[e]);
                }
            }
            if constexpr (tz_ >= 0) {
                const FTerm<C> &f = p.t[tz_];
                constexpr int u = PatInfo<P>::uz;
                const T *b = p.F[u] + k * plane + j0 * Nx + x;
                if (!(SCATTER && ty_ >= 0)) {
#pragma unroll
                    for (int r = 0; r < R; ++r) ldg_vec_stream<T, VEC>(cur[u][r], b + r * Nx);
                }
                const T *bn = b + (f.fwd ? plane : -plane);
#pragma unroll
                for (int r = 0; r < R; ++r) ldg_vec_stream<T, VEC>(znb[r], bn + r * Nx);
                cz = f.fwd ? f.c[k] : neg(f.c[k]);
            }
            if (ADD && !SCATTER) {
#pragma unroll
                for (int r = 0; r < R; ++r) ldg_vec_plain<T, VEC>(old[r], p.G[0] + k * plane + (j0 + r) * Nx + x);
            }
            // -- arithmetic ---------------------------------------------------------------------------
            T acc[R][VEC];
#pragma unroll
            for (int t = 0; t < D.nt; ++t) {
                const int u = (D.axis[t] == 0) ? PatInfo<P>::ux : PatInfo<P>::uz;  // (unused for the y-term)
                const int yf = (ty_ >= 0) ? p.t[ty_ >= 0 ? ty_ : 0].fwd : 1;
                T val[R][VEC];
#pragma unroll
                for (int r = 0; r < R; ++r)
#pragma unroll
                    for (int e = 0; e < VEC; ++e) {
                        // centre value of this term's input component
                        T c0;
                        if (D.axis[t] == 1 || (SCATTER && ty_ >= 0)) c0 = yf ? ynb[r][e] : ynb[r + 1][e];
                        else c0 = cur[u][r][e];
                        if (D.axis[t] == 0) {
                            const int xf = p.t[t].fwd;
                            T n;
                            if (SCATTER && ty_ >= 0) {
                                const T up = yf ? ynb[r][e + 1 < VEC ? e + 1 : e] : ynb[r + 1][e + 1 < VEC ? e + 1 : e];
                                const T dn = yf ? ynb[r][e > 0 ? e - 1 : 0] : ynb[r + 1][e > 0 ? e - 1 : 0];
                                n = xf ? (e < VEC - 1 ? up : xnb[r]) : (e > 0 ? dn : xnb[r]);
                            } else {
                                n = xf ? (e < VEC - 1 ? cur[u][r][e + 1 < VEC ? e + 1 : e] : xnb[r])
                                       : (e > 0 ? cur[u][r][e > 0 ? e - 1 : 0] : xnb[r]);
                            }
                            val[r][e] = mul(cx[e], sub(n, c0));
                        } else if (D.axis[t] == 1) {
                            const T n = yf ? ynb[r + 1][e] : ynb[r][e];
                            val[r][e] = mul(cy[r], sub(n, c0));
                        } else {
                            val[r][e] = mul(cz, sub(znb[r][e], c0));
                        }
                    }
                if (SCATTER) {
                    T *g = p.G[D.out[t]] + k * plane + j0 * Nx + x;
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        if (ADD) {
                            T o2[VEC];
                            ldg_vec_plain<T, VEC>(o2, g + r * Nx);
#pragma unroll
                            for (int e = 0; e < VEC; ++e) val[r][e] = add(o2[e], val[r][e]);
                        }
                        stg_vec_stream<T, VEC>(g + r * Nx, val[r]);
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < R; ++r)
#pragma unroll
                        for (int e = 0; e < VEC; ++e)
                            acc[r][e] = (t == 0) ? (ADD ? add(old[r][e], val[r][e]) : val[r][e]) : add(acc[r][e], val[r][e]);
                }
            }
            if (!SCATTER) {
                T *g = p.G[0] + k * plane + j0 * Nx + x;
#pragma unroll
                for (int r = 0; r < R; ++r) stg_vec_stream<T, VEC>(g + r * Nx, acc[r]);
            }
            continue;
        }
        // =========================== boundary threads ================================================
        if (MODE == FUSED_INTERIOR) continue;
        if (!SCATTER) {
            T acc[R][VEC];
            T *g = p.G[0] + k * plane + x;
            if (ADD) {
#pragma unroll
                for (int r = 0; r < R; ++r)
                    if (j0 + r < Ny) ldg_vec_plain<T, VEC>(acc[r], g + (j0 + r) * Nx);
            }
#pragma unroll
            for (int t = 0; t < D.nt; ++t) {
                const FTerm<C> &f = p.t[t];
                const int u = D.in[t];
                const T *base = p.F[u] + k * plane;
                T val[R][VEC];
                if (D.axis[t] == 0) {
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        if (j0 + r >= Ny) continue;
                        T cur[VEC];
                        ldg_vec_stream<T, VEC>(cur, base + (j0 + r) * Nx + x);
                        x_term(f, base + (j0 + r) * Nx, cur, val[r]);
                    }
                } else if (D.axis[t] == 1) {
                    T V[R + 1][VEC];
                    load_rows_y(f, base, V);
#pragma unroll
                    for (int r = 0; r < R; ++r)
#pragma unroll
                        for (int e = 0; e < VEC; ++e)
                            val[r][e] = partial_eval<T, C>(f, Ny, j0 + r, f.fwd ? V[r][e] : V[r + 1][e], f.fwd ? V[r + 1][e] : V[r][e]);
                } else {
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        if (j0 + r >= Ny) continue;
                        T cur[VEC], nb[VEC];
                        ldg_vec_stream<T, VEC>(cur, base + (j0 + r) * Nx + x);
                        load_row_z(f, u, k, j0 + r, nb);
#pragma unroll
                        for (int e = 0; e < VEC; ++e) val[r][e] = partial_eval<T, C>(f, Nz, k, cur[e], nb[e]);
                    }
                }
#pragma unroll
                for (int r = 0; r < R; ++r)
#pragma unroll
                    for (int e = 0; e < VEC; ++e) acc[r][e] = (t == 0 && !ADD) ? val[r][e] : add(acc[r][e], val[r][e]);
            }
#pragma unroll
            for (int r = 0; r < R; ++r)
                if (j0 + r < Ny) stg_vec_stream<T, VEC>(g + (j0 + r) * Nx, acc[r]);
        } else {
            const T *base = p.F[0] + k * plane;
            int ty_ = 0;  // the y-term (every scatter pattern has exactly one)
#pragma unroll
            for (int t = 0; t < D.nt; ++t)
                if (D.axis[t] == 1) ty_ = t;
            const FTerm<C> &fy = p.t[ty_];
            T V[R + 1][VEC];
            load_rows_y(fy, base, V);
#pragma unroll
            for (int t = 0; t < D.nt; ++t) {
                const FTerm<C> &f = p.t[t];
                T *g = p.G[D.out[t]] + k * plane + x;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if (j0 + r >= Ny) continue;
                    T cur[VEC], val[VEC];
#pragma unroll
                    for (int e = 0; e < VEC; ++e) cur[e] = fy.fwd ? V[r][e] : V[r + 1][e];
                    if (D.axis[t] == 0) {
                        x_term(f, base + (j0 + r) * Nx, cur, val);
                    } else if (D.axis[t] == 1) {
#pragma unroll
                        for (int e = 0; e < VEC; ++e)
                            val[e] = partial_eval<T, C>(f, Ny, j0 + r, cur[e], f.fwd ? V[r + 1][e] : V[r][e]);
                    } else {
                        T nb[VEC];
                        load_row_z(f, 0, k, j0 + r, nb);
#pragma unroll
                        for (int e = 0; e < VEC; ++e) val[e] = partial_eval<T, C>(f, Nz, k, cur[e], nb[e]);
                    }
                    if (ADD) {
                        T old[VEC];
                        ldg_vec_plain<T, VEC>(old, g + (j0 + r) * Nx);
#pragma unroll
                        for (int e = 0; e < VEC; ++e) val[e] = add(old[e], val[e]);
                    }
                    stg_vec_stream<T, VEC>(g + (j0 + r) * Nx, val);
                }
            }
        }
    }
}

}  // namespace sgc
