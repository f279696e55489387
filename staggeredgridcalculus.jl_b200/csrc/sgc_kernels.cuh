// sgc_kernels.cuh — matrix-free stencil kernels (hand-written for sm_100a).
//
//   k_term      one block of an operator: Gv OP (∂w | m̂w)(Fu) for one component pair
//               (matrixfree/differential/partial.jl:33-451, matrixfree/mean.jl:102-922).
//   k_curl3     the fused full 3-D curl (matrixfree/differential/curl.jl:52-137): all six
//               ∂ blocks in ONE pass, z-march with register-resident planes, y-neighbours
//               through a shared-memory tile, x-neighbours through warp shuffles, all
//               boundary rules (Bloch phase, PEC/PMC symmetry, slab interfaces) folded in.
//
// Every formula keeps the reference's operand order; the library is
// built with -fmad=false so nothing is contracted.
#pragma once
#include "sgc_common.cuh"

namespace sgc {

enum : int { KIND_PARTIAL = 0, KIND_MEAN_ARITH = 1, KIND_MEAN_WEIGHTED = 2 };
enum : int { BC_BLOCH = 0, BC_SYM = 1, BC_IFACE = 2 };

// ---------------------------------------------------------------------------------------------
// Generic single-term kernel
// ---------------------------------------------------------------------------------------------
template <typename T, typename C>
struct TermParams {
    T *G;             // output component
    const T *F;       // input component
    const T *ghost;   // plane replacing the wrap-around plane (never null; aliases F when local)
    const C *c;       // ∂: α·∆w⁻¹[i]          m̂ weighted: (0.5α)·∆w′⁻¹[i]     (length Nw)
    const C *w;       // m̂ weighted: ∆w[i]                                     (length Nw)
    C a2, a1;         // m̂ arithmetic: 0.5α, α
    C b0;             // m̂ weighted, backward+symmetry: α·∆w′⁻¹[1]
    C w_hi, w_lo;     // m̂ weighted: ∆w of the ghost plane above / below
    cplx e;           // e⁻ⁱᵏᴸ
    int e_real;       // e is a Julia Real
    int fwd;
    int bc_lo, bc_hi; // BC_BLOCH | BC_SYM | BC_IFACE at i = 0 / i = Nw-1
    int64_t inner;    // product of the dims faster than the operator axis
    int64_t Nw;       // length of the operator axis
    int64_t outer;    // product of the slower dims
    int64_t i_begin, i_end;  // restriction along the operator axis   (range applies here if it is the last dim)
    int64_t o_begin, o_end;  // restriction along the outer index
    int64_t ghost_stride_o;  // element stride of `ghost` per outer index (0 when outer is not used)
};

template <typename T, typename C, int KIND>
SGC_D T term_value(const TermParams<T, C> &p, int64_t x, int64_t i, int64_t o) {
    const int64_t Nw = p.Nw;
    const int64_t base = (o * Nw) * p.inner + x;
    const T cur = ldg_stream(p.F + base + i * p.inner);
    if (p.fwd) {
        if (i == Nw - 1) {  // positive end
            if (p.bc_hi == BC_SYM) {
                if (KIND == KIND_PARTIAL) return neg(mul(p.c[i], cur));             // partial.jl:84-87
                if (KIND == KIND_MEAN_ARITH) return mul(p.a2, cur);                 // mean.jl:149-151
                return mul(mul(p.c[i], p.w[i]), cur);                               // mean.jl:546-549
            }
            T g = ldg_plain(p.ghost + o * p.ghost_stride_o + x);
            if (p.bc_hi == BC_BLOCH) {
                if (KIND == KIND_PARTIAL) return mul(p.c[i], sub(mul_e(p.e, g, p.e_real), cur));    // :62-65
                if (KIND == KIND_MEAN_ARITH) return mul(p.a2, add(mul_e(p.e, g, p.e_real), cur));   // :129-131
                return mul(p.c[i], add(mul(p.w_hi, g), mul(p.w[i], cur)));  // w_hi = ∆w[1]*e      // :524-527
            }
            if (KIND == KIND_PARTIAL) return mul(p.c[i], sub(g, cur));
            if (KIND == KIND_MEAN_ARITH) return mul(p.a2, add(g, cur));
            return mul(p.c[i], add(mul(p.w_hi, g), mul(p.w[i], cur)));
        }
        const T nb = ldg_stream(p.F + base + (i + 1) * p.inner);
        if (i == 0 && p.bc_lo == BC_SYM) {  // negative end, input forced to zero
            if (KIND == KIND_PARTIAL) return mul(p.c[0], nb);                        // partial.jl:77-80
            if (KIND == KIND_MEAN_ARITH) return mul(p.a2, nb);                       // mean.jl:143-145
            return mul(mul(p.c[0], p.w[1]), nb);                                     // mean.jl:539-542
        }
        if (KIND == KIND_PARTIAL) return mul(p.c[i], sub(nb, cur));                  // partial.jl:54-58
        if (KIND == KIND_MEAN_ARITH) return mul(p.a2, add(nb, cur));                 // mean.jl:121-125
        return mul(p.c[i], add(mul(p.w[i + 1], nb), mul(p.w[i], cur)));              // mean.jl:516-520
    }
    // backward
    if (i == 0) {
        if (p.bc_lo == BC_SYM) {
            if (KIND == KIND_PARTIAL) return zero_of(cur);                           // partial.jl:184-188
            if (KIND == KIND_MEAN_ARITH) return mul(p.a1, cur);                      // mean.jl:241-245
            return mul(mul(p.b0, p.w[0]), cur);                                      // mean.jl:646-650
        }
        T g = ldg_plain(p.ghost + o * p.ghost_stride_o + x);
        if (p.bc_lo == BC_BLOCH) {
            if (KIND == KIND_PARTIAL) return mul(p.c[0], sub(cur, div_e(g, p.e, p.e_real)));       // :179-183
            if (KIND == KIND_MEAN_ARITH) return mul(p.a2, add(cur, div_e(g, p.e, p.e_real)));      // :237-240
            return mul(p.c[0], add(mul(p.w[0], cur), div_e(mul(p.w_lo, g), p.e, p.e_real)));       // :641-645
        }
        if (KIND == KIND_PARTIAL) return mul(p.c[0], sub(cur, g));
        if (KIND == KIND_MEAN_ARITH) return mul(p.a2, add(cur, g));
        return mul(p.c[0], add(mul(p.w[0], cur), mul(p.w_lo, g)));
    }
    const T nb = ldg_stream(p.F + base + (i - 1) * p.inner);
    if (KIND == KIND_PARTIAL) return mul(p.c[i], sub(cur, nb));                      // partial.jl:171-175
    if (KIND == KIND_MEAN_ARITH) return mul(p.a2, add(cur, nb));                     // mean.jl:229-233
    return mul(p.c[i], add(mul(p.w[i], cur), mul(p.w[i - 1], nb)));                  // mean.jl:633-637
}

// Thread mapping.  The three logical indices are x (fastest, `inner` long), i (operator axis)
// and o (outer).  XAXIS = true is the case inner == 1 (operator along the contiguous axis):
// threadIdx.x runs along i and (blockIdx.y, threadIdx.y) along o.  Otherwise threadIdx.x runs
// along x, (blockIdx.y, threadIdx.y) along i and blockIdx.z along o.
template <typename T, typename C, int KIND, bool ADD, bool XAXIS>
__global__ void __launch_bounds__(256) k_term(const __grid_constant__ TermParams<T, C> p) {
    if (XAXIS) {
        const int64_t i = p.i_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (i >= p.i_end) return;
        for (int64_t o = p.o_begin + (int64_t)blockIdx.y * blockDim.y + threadIdx.y; o < p.o_end;
             o += (int64_t)gridDim.y * blockDim.y) {
            T v = term_value<T, C, KIND>(p, 0, i, o);
            T *g = p.G + o * p.Nw + i;
            if (ADD) v = add(ldg_plain(g), v);
            stg_stream(g, v);
        }
    } else {
        const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (x >= p.inner) return;
        for (int64_t o = p.o_begin + blockIdx.z; o < p.o_end; o += gridDim.z) {
            for (int64_t i = p.i_begin + (int64_t)blockIdx.y * blockDim.y + threadIdx.y; i < p.i_end;
                 i += (int64_t)gridDim.y * blockDim.y) {
                T v = term_value<T, C, KIND>(p, x, i, o);
                T *g = p.G + (o * p.Nw + i) * p.inner + x;
                if (ADD) v = add(ldg_plain(g), v);
                stg_stream(g, v);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Vectorised single-term kernels (the fast path of k_term)
// ---------------------------------------------------------------------------------------------
// Same formulas, evaluated from values already in registers: `nb` is the value at i+1 (forward)
// or i−1 (backward); at the physical / slab face it is the RAW ghost value.
// ci = c[i], wi = w[i], wn = w[i+1] (forward) / w[i−1] (backward) — every table entry the formulas of
// element i touch (c[0], w[0], w[1] in the boundary branches are c[i], w[i], w[i+1] at i = 0).
template <typename T, typename C, int KIND>
SGC_D T term_eval_c(const TermParams<T, C> &p, int64_t i, T cur, T nb, C ci, C wi, C wn) {
    const int64_t Nw = p.Nw;
    if (p.fwd) {
        if (i == Nw - 1) {
            if (p.bc_hi == BC_SYM) {
                if (KIND == KIND_PARTIAL) return neg(mul(ci, cur));
                if (KIND == KIND_MEAN_ARITH) return mul(p.a2, cur);
                return mul(mul(ci, wi), cur);
            }
            if (p.bc_hi == BC_BLOCH) {
                if (KIND == KIND_PARTIAL) return mul(ci, sub(mul_e(p.e, nb, p.e_real), cur));
                if (KIND == KIND_MEAN_ARITH) return mul(p.a2, add(mul_e(p.e, nb, p.e_real), cur));
                return mul(ci, add(mul(p.w_hi, nb), mul(wi, cur)));
            }
            if (KIND == KIND_PARTIAL) return mul(ci, sub(nb, cur));
            if (KIND == KIND_MEAN_ARITH) return mul(p.a2, add(nb, cur));
            return mul(ci, add(mul(p.w_hi, nb), mul(wi, cur)));
        }
        if (i == 0 && p.bc_lo == BC_SYM) {
            if (KIND == KIND_PARTIAL) return mul(ci, nb);
            if (KIND == KIND_MEAN_ARITH) return mul(p.a2, nb);
            return mul(mul(ci, wn), nb);
        }
        if (KIND == KIND_PARTIAL) return mul(ci, sub(nb, cur));
        if (KIND == KIND_MEAN_ARITH) return mul(p.a2, add(nb, cur));
        return mul(ci, add(mul(wn, nb), mul(wi, cur)));
    }
    if (i == 0) {
        if (p.bc_lo == BC_SYM) {
            if (KIND == KIND_PARTIAL) return zero_of(cur);
            if (KIND == KIND_MEAN_ARITH) return mul(p.a1, cur);
            return mul(mul(p.b0, wi), cur);
        }
        if (p.bc_lo == BC_BLOCH) {
            if (KIND == KIND_PARTIAL) return mul(ci, sub(cur, div_e(nb, p.e, p.e_real)));
            if (KIND == KIND_MEAN_ARITH) return mul(p.a2, add(cur, div_e(nb, p.e, p.e_real)));
            return mul(ci, add(mul(wi, cur), div_e(mul(p.w_lo, nb), p.e, p.e_real)));
        }
        if (KIND == KIND_PARTIAL) return mul(ci, sub(cur, nb));
        if (KIND == KIND_MEAN_ARITH) return mul(p.a2, add(cur, nb));
        return mul(ci, add(mul(wi, cur), mul(p.w_lo, nb)));
    }
    if (KIND == KIND_PARTIAL) return mul(ci, sub(cur, nb));
    if (KIND == KIND_MEAN_ARITH) return mul(p.a2, add(cur, nb));
    return mul(ci, add(mul(wi, cur), mul(wn, nb)));
}
template <typename T, typename C, int KIND>
SGC_D T term_eval(const TermParams<T, C> &p, int64_t i, T cur, T nb) {
    const int64_t Nw = p.Nw;
    C ci{}, wi{}, wn{};
    if (KIND != KIND_MEAN_ARITH) ci = p.c[i];
    if (KIND == KIND_MEAN_WEIGHTED) {
        wi = p.w[i];
        const int64_t in = p.fwd ? i + 1 : i - 1;
        if (in >= 0 && in < Nw) wn = p.w[in];
    }
    return term_eval_c<T, C, KIND>(p, i, cur, nb, ci, wi, wn);
}

// VEC consecutive elements (VEC·sizeof(T) is a multiple of 16 bytes, address 16-byte aligned)
template <typename T, int VEC>
SGC_D void ldg_vec_stream(T (&v)[VEC], const T *ptr) {
    constexpr int CH = VEC * (int)sizeof(T) / 16;
    double2 t[CH];
#pragma unroll
    for (int q = 0; q < CH; ++q)
        asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(t[q].x), "=d"(t[q].y) : "l"(reinterpret_cast<const double2 *>(ptr) + q));
    const double *d = reinterpret_cast<const double *>(t);
    double *o = reinterpret_cast<double *>(v);
#pragma unroll
    for (int q = 0; q < 2 * CH; ++q) o[q] = d[q];
}
template <typename T, int VEC>
SGC_D void ldg_vec_plain(T (&v)[VEC], const T *ptr) {
    constexpr int CH = VEC * (int)sizeof(T) / 16;
    double *o = reinterpret_cast<double *>(v);
#pragma unroll
    for (int q = 0; q < CH; ++q) {
        const double2 t = reinterpret_cast<const double2 *>(ptr)[q];
        o[2 * q] = t.x; o[2 * q + 1] = t.y;
    }
}
template <typename T, int VEC>
SGC_D void stg_vec_stream(T *ptr, const T (&v)[VEC]) {
    constexpr int CH = VEC * (int)sizeof(T) / 16;
    const double *d = reinterpret_cast<const double *>(v);
#pragma unroll
    for (int q = 0; q < CH; ++q)
        asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1,%2};" ::"l"(reinterpret_cast<double2 *>(ptr) + q), "d"(d[2 * q]), "d"(d[2 * q + 1]) : "memory");
}

// Operator along the contiguous axis (inner == 1): each thread owns VEC consecutive elements of R
// rows; the neighbour across the vector boundary comes from the adjacent lane by shuffle, the
// warp-edge lane fetches it (or the wrap / ghost element) itself.
template <typename T, typename C, int KIND, bool ADD, int VEC, int R>
__global__ void __launch_bounds__(256) k_term_x(const __grid_constant__ TermParams<T, C> p) {
    const int64_t Nw = p.Nw;
    const int64_t iv = ((int64_t)blockIdx.x * 256 + threadIdx.x) * VEC;
    const int lane = threadIdx.x & 31;
    const bool active = iv < Nw;
    const int fwd = p.fwd;
    const bool edge = fwd ? (lane == 31 || iv + VEC >= Nw) : (lane == 0);
    const int64_t nbi = fwd ? iv + VEC : iv - 1;  // index of the element beyond my vector
    const bool nb_inside = (nbi >= 0 && nbi < Nw);
    const bool nb_ghost = fwd ? (p.bc_hi != BC_SYM) : (p.bc_lo != BC_SYM);
    const T Z = zero_of(T{});
    // the table entries of my VEC elements are the same for every row: load them once
    C ci[VEC], wi[VEC], wn[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
        ci[e] = C{}; wi[e] = C{}; wn[e] = C{};
        const int64_t i = iv + e, in = fwd ? i + 1 : i - 1;
        if (active && i < Nw) {
            if (KIND != KIND_MEAN_ARITH) ci[e] = p.c[i];
            if (KIND == KIND_MEAN_WEIGHTED) {
                wi[e] = p.w[i];
                if (in >= 0 && in < Nw) wn[e] = p.w[in];
            }
        }
    }
    for (int64_t o0 = p.o_begin + (int64_t)blockIdx.y * R; o0 < p.o_end; o0 += (int64_t)gridDim.y * R) {
        T cur[R][VEC];
        T nb[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const bool live = active && (o0 + r < p.o_end);
            if (live) ldg_vec_stream<T, VEC>(cur[r], p.F + (o0 + r) * Nw + iv);
            else {
#pragma unroll
                for (int e = 0; e < VEC; ++e) cur[r][e] = Z;
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const bool live = active && (o0 + r < p.o_end);
            T got = fwd ? shfl_down1(cur[r][0]) : shfl_up1(cur[r][VEC - 1]);
            if (edge && live) {
                if (nb_inside) got = ldg_stream(p.F + (o0 + r) * Nw + nbi);
                else got = nb_ghost ? ldg_plain(p.ghost + (o0 + r) * p.ghost_stride_o) : Z;
            }
            nb[r] = got;
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (!(active && (o0 + r < p.o_end))) continue;
            T out[VEC];
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                const T n = fwd ? (e < VEC - 1 ? cur[r][e + 1 < VEC ? e + 1 : e] : nb[r]) : (e > 0 ? cur[r][e > 0 ? e - 1 : 0] : nb[r]);
                out[e] = term_eval_c<T, C, KIND>(p, iv + e, cur[r][e], n, ci[e], wi[e], wn[e]);
            }
            T *g = p.G + (o0 + r) * Nw + iv;
            if (ADD) {
                T old[VEC];
                ldg_vec_plain<T, VEC>(old, g);
#pragma unroll
                for (int e = 0; e < VEC; ++e) out[e] = add(old[e], out[e]);
            }
            stg_vec_stream<T, VEC>(g, out);
        }
    }
}

// Operator along a strided axis (inner > 1): each thread owns VEC consecutive x of R consecutive
// rows along the operator axis and keeps them in registers, so R outputs cost R+1 row loads.
// (accumulating calls run with R = 2, and they and the weighted means with three resident CTAs per SM — measured; the arithmetic mean
// and ∂ `=` forms are faster unconstrained —, the read-modify-write loads of G issued together with
// the loads of F: 8192² Float64 m̂y `+=` 0.32 → see profiles/r2_time_2d_vec.jsonl)
template <typename T, typename C, int KIND, bool ADD, int VEC, int R>
__global__ void __launch_bounds__(256, ((R <= 2 || KIND == KIND_MEAN_WEIGHTED) ? 3 : 0)) k_term_s(const __grid_constant__ TermParams<T, C> p) {
    const int64_t Nw = p.Nw, inner = p.inner;
    const int64_t x = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    if (x >= inner) return;
    const int fwd = p.fwd;
    const T Z = zero_of(T{});
    for (int64_t o = p.o_begin + blockIdx.z; o < p.o_end; o += gridDim.z) {
        const T *Fo = p.F + o * Nw * inner + x;
        for (int64_t i0 = p.i_begin + ((int64_t)blockIdx.y * blockDim.y + threadIdx.y) * R; i0 < p.i_end;
             i0 += (int64_t)gridDim.y * blockDim.y * R) {
            // rows i0-1+fwd .. i0+R-1+fwd
            T V[R + 1][VEC];
#pragma unroll
            for (int q = 0; q <= R; ++q) {
                const int64_t ii = i0 - 1 + fwd + q;
                const bool needed = (q == 0) || (i0 + q - 1 < p.i_end);  // cur of output row q-… or nb of its neighbour
                if (needed && ii >= 0 && ii < Nw) ldg_vec_stream<T, VEC>(V[q], Fo + ii * inner);
                else if (needed && ((ii < 0 && p.bc_lo != BC_SYM) || (ii >= Nw && p.bc_hi != BC_SYM)))
                    ldg_vec_plain<T, VEC>(V[q], p.ghost + o * p.ghost_stride_o + x);
                else {
#pragma unroll
                    for (int e = 0; e < VEC; ++e) V[q][e] = Z;
                }
            }
            T old[ADD ? R : 1][VEC];  // read-modify-write: the old values of G travel with the loads of F, not behind the arithmetic
            if (ADD) {
#pragma unroll
                for (int r = 0; r < R; ++r)
                    if (i0 + r < p.i_end) ldg_vec_plain<T, VEC>(old[ADD ? r : 0], p.G + (o * Nw + i0 + r) * inner + x);
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int64_t i = i0 + r;
                if (i >= p.i_end) break;
                // the table entries of row i, once for the whole vector
                C ci{}, wi{}, wn{};
                if (KIND != KIND_MEAN_ARITH) ci = p.c[i];
                if (KIND == KIND_MEAN_WEIGHTED) {
                    wi = p.w[i];
                    const int64_t in = fwd ? i + 1 : i - 1;
                    if (in >= 0 && in < Nw) wn = p.w[in];
                }
                T out[VEC];
#pragma unroll
                for (int e = 0; e < VEC; ++e)
                    out[e] = term_eval_c<T, C, KIND>(p, i, fwd ? V[r][e] : V[r + 1][e], fwd ? V[r + 1][e] : V[r][e], ci, wi, wn);
                T *g = p.G + (o * Nw + i) * inner + x;
                if (ADD) {
#pragma unroll
                    for (int e = 0; e < VEC; ++e) out[e] = add(old[ADD ? r : 0][e], out[e]);
                }
                stg_vec_stream<T, VEC>(g, out);
            }
        }
    }
}

// Gv .= 0   (matrixfree/differential/curl.jl:77) restricted to a contiguous range
template <typename T>
__global__ void __launch_bounds__(256) k_fill_zero(T *G, int64_t n) {
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n;
         q += (int64_t)gridDim.x * blockDim.x)
        stg_stream(G + q, zero_of(T{}));
}

// ---------------------------------------------------------------------------------------------
// Fused full 3-D curl
// ---------------------------------------------------------------------------------------------
// slots of a rank's control block (int64 each; sgc_comm.cu allocates and peer-maps it)
enum : int { SGC_CTRL_FLAG_LO = 0, SGC_CTRL_FLAG_HI = 1, SGC_CTRL_EPOCH = 2, SGC_CTRL_TICKETS = 3 /* 2 x uint32 */,
             SGC_CTRL_SIDES_DONE = 4 /* uint32 */, SGC_CTRL_BAR_COUNT = 5, SGC_CTRL_TIMEOUTS = 6, SGC_CTRL_BAR_SLOTS = 8 /* + world */ };
// a neighbour that never arrives (ranks issuing different numbers of synchronised applies, a rank that died) must not
// hang the GPU: after this many polls (seconds) the waiter counts a timeout in its control block and goes on
constexpr unsigned SGC_SPIN_LIMIT = 1u << 23;

// apply_mean! through the staged z-march (k_curl3p MODE = CURL3_MEAN): the m̂ coefficients of one axis
// (the fields of TermParams, matrixfree/mean.jl:102-922)
template <typename C>
struct MeanAxis {
    const C *w;        // weighted: ∆w[i]
    C a2, a1;          // arithmetic: 0.5α, α
    C b0;              // weighted, backward + symmetric: α·∆w′⁻¹[1]
    C w_hi, w_lo;      // weighted: ∆w of the plane beyond the upper / lower end (Bloch phase folded into w_hi)
    int weighted;
};

template <typename T, typename C>
struct CurlParams {
    const T *F[3];       // Fx, Fy, Fz  (each Nx*Ny*Nz)
    T *G[3];             // Gx, Gy, Gz
    const T *ghost[2];   // z-ghost planes of Fx, Fy (Nx*Ny each); alias F's own wrap plane when local
    const C *c[3];       // α·∆x⁻¹[i], α·∆y⁻¹[j], α·∆z⁻¹[k]
    cplx e[3];
    int e_real[3];
    int fwd[3];
    int bc_lo[3], bc_hi[3];
    int Nx, Ny, Nz;
    int k_begin, k_end;  // planes to produce
    int march;           // planes per CTA
    // neighbour synchronisation folded into the kernel (z-slabs in peer memory, no barrier kernel), see
    // sgc_comm.cu: the CTAs that touch the first (last) plane of the slab wait until the lower (upper) neighbour
    // has finished the boundary CTAs of its previous synchronised apply (the count that neighbour wrote into MY
    // control block >= my own count of finished applies); per side, the last boundary CTA to finish publishes
    // count + 1 into that neighbour's control block, and the last side to finish advances my own count.  The count
    // lives in device memory (sync_ctrl[SGC_CTRL_EPOCH]), so the launch is the same every time and can be replayed
    // from a CUDA graph.  Interior CTAs neither wait nor count.  The z-chunks are rotated by one so that the two
    // boundary chunks are scheduled LAST: a whole kernel duration of slack against neighbour skew, and the peer
    // loads of the ghost plane overlap the interior.  sync_ctrl == nullptr: off.
    long long *sync_ctrl;            // this rank's control block (SGC_CTRL_* slots)
    long long *sync_post[2];         // peer addresses: the lower neighbour's FLAG_HI slot, the upper neighbour's FLAG_LO slot
    int sync_side[2];                // this rank has a neighbour on the lower / upper side
    MeanAxis<C> m[3];                // MODE = CURL3_MEAN only (c[d] then holds 0.5α·∆w′⁻¹ of a weighted mean)
};

// m̂ value at a cell that no boundary rule touches (mean.jl:121-125 / :229-233 arithmetic, :516-520 / :633-637 weighted;
// the forward and backward forms differ only in the order of the two summands, which IEEE addition does not see)
template <typename T, typename C>
SGC_D T mean_plain(bool weighted, C a2, C c, C w_cur, C w_nb, T cur, T nb) {
    if (!weighted) return mul(a2, add(nb, cur));
    return mul(c, add(mul(w_nb, nb), mul(w_cur, cur)));
}
// m̂ value with the boundary rules: the branches and formulas of term_value<…, KIND_MEAN_*> with the operands in
// registers; `nb` is the RAW neighbour (the wrap-around value at a Bloch face, unused at a symmetric one)
template <typename T, typename C>
SGC_D T mean_eval(const CurlParams<T, C> &p, int d, int i, int Nw, T cur, T nb) {
    const MeanAxis<C> &m = p.m[d];
    const C *c = p.c[d];
    const bool W = m.weighted != 0;
    if (p.fwd[d]) {
        if (i == Nw - 1) {
            if (p.bc_hi[d] == BC_SYM) return W ? mul(mul(c[i], m.w[i]), cur) : mul(m.a2, cur);                      // :546-549 / :149-151
            if (W) return mul(c[i], add(mul(m.w_hi, nb), mul(m.w[i], cur)));                                        // :524-527
            if (p.bc_hi[d] == BC_BLOCH) return mul(m.a2, add(mul_e(p.e[d], nb, p.e_real[d]), cur));                  // :129-131
            return mul(m.a2, add(nb, cur));
        }
        if (i == 0 && p.bc_lo[d] == BC_SYM) return W ? mul(mul(c[0], m.w[1]), nb) : mul(m.a2, nb);                   // :539-542 / :143-145
        return W ? mul(c[i], add(mul(m.w[i + 1], nb), mul(m.w[i], cur))) : mul(m.a2, add(nb, cur));
    }
    if (i == 0) {
        if (p.bc_lo[d] == BC_SYM) return W ? mul(mul(m.b0, m.w[0]), cur) : mul(m.a1, cur);                           // :646-650 / :241-245
        if (p.bc_lo[d] == BC_BLOCH)
            return W ? mul(c[0], add(mul(m.w[0], cur), div_e(mul(m.w_lo, nb), p.e[d], p.e_real[d])))                 // :641-645
                     : mul(m.a2, add(cur, div_e(nb, p.e[d], p.e_real[d])));                                          // :237-240
        return W ? mul(c[0], add(mul(m.w[0], cur), mul(m.w_lo, nb))) : mul(m.a2, add(cur, nb));
    }
    return W ? mul(c[i], add(mul(m.w[i], cur), mul(m.w[i - 1], nb))) : mul(m.a2, add(cur, nb));
}

SGC_D long long ld_acquire_sys(const long long *p) {
    long long v;
    asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
SGC_D void st_release_sys(long long *p, long long v) {
    asm volatile("st.release.sys.global.s64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
SGC_D long long ld_volatile_s64(const long long *p) {
    long long v;
    asm volatile("ld.volatile.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
SGC_D void st_volatile_s64(long long *p, long long v) {
    asm volatile("st.volatile.global.s64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// wait until *flag >= want (written by a peer over NVLink); bounded, see SGC_SPIN_LIMIT
SGC_D void wait_flag_sys(const long long *flag, long long want, long long *ctrl, unsigned ns) {
    unsigned polls = 0;
    while (ld_acquire_sys(flag) < want) {
        __nanosleep(ns);
        if (++polls > SGC_SPIN_LIMIT) {
            atomicAdd(reinterpret_cast<unsigned long long *>(ctrl + SGC_CTRL_TIMEOUTS), 1ull);
            break;
        }
    }
}

// One CTA owns a TX×TY column of cells and marches `march` planes up in z.
//   * Fx, Fy: two planes live in registers (`lo`, `hi`); ∂z = cz[k]·(hi − lo) for both forward
//     (lo = plane k, hi = k+1) and backward (lo = k−1, hi = k) differences, so every plane
//     of Fx, Fy is loaded from HBM exactly once per CTA (+1 at the start of the march).
//   * y-neighbours of Fx and Fz come from a (TY+1)×TX shared-memory tile (double-buffered,
//     one __syncthreads per plane); the extra row is the only y-halo traffic.
//   * x-neighbours of Fy and Fz come from the neighbouring lane (__shfl); the lane at the edge
//     of the warp loads the single missing element itself.
//   * loads for plane k+1 are issued before plane k is computed (register prefetch).
template <typename T, typename C, bool ADD, int TX, int TY>
__global__ void __launch_bounds__(TX *TY, (TX * TY <= 256) ? 2 : 1) k_curl3(const __grid_constant__ CurlParams<T, C> p) {
    static_assert(TX % 32 == 0, "a warp must lie inside one row");
    __shared__ T sFx[2][TY + 1][TX];
    __shared__ T sFz[2][TY + 1][TX];

    const int tx = threadIdx.x, ty = threadIdx.y;
    const int lane = tx & 31;
    const int Nx = p.Nx, Ny = p.Ny, Nz = p.Nz;
    const int i = blockIdx.x * TX + tx;
    const int j = blockIdx.y * TY + ty;
    const bool valid = (i < Nx) && (j < Ny);
    const int64_t sy = Nx, sz = (int64_t)Nx * Ny;
    const int64_t ij = (int64_t)j * sy + i;

    const int fx = p.fwd[0], fy = p.fwd[1], fz = p.fwd[2];

    // ---- per-thread x-edge description (who supplies the x-neighbour) ------------------------
    // forward: neighbour is i+1, taken from lane+1; backward: i-1 from lane-1.
    // xload: this lane must fetch its neighbour from memory at index xnb (warp edge or wrap).
    // xkind: 0 plain neighbour, 1 Bloch wrap (apply e), 2 symmetric ghost / zero term.
    int xnb, xkind = 0;
    bool xload;
    if (fx) {
        const bool last = (i == Nx - 1);
        xnb = last ? 0 : i + 1;
        xkind = last ? (p.bc_hi[0] == BC_BLOCH ? 1 : 2) : 0;
        xload = valid && (last ? (xkind == 1) : (lane == 31));
    } else {
        const bool first = (i == 0);
        xnb = first ? Nx - 1 : i - 1;
        xkind = first ? (p.bc_lo[0] == BC_BLOCH ? 1 : 2) : 0;
        xload = valid && (first ? (xkind == 1) : (lane == 0));
    }
    const bool x_lo_zero = fx && (i == 0) && (p.bc_lo[0] == BC_SYM);  // forward symmetric: Fu[1] := 0
    const int64_t xnb_off = (int64_t)j * sy + xnb;

    // ---- per-thread y-edge description ------------------------------------------------------------
    int ykind = 0;  // 0 plain, 1 Bloch wrap, 2 symmetric
    if (fy) {
        if (j == Ny - 1) ykind = (p.bc_hi[1] == BC_BLOCH) ? 1 : 2;
    } else {
        if (j == 0) ykind = (p.bc_lo[1] == BC_BLOCH) ? 1 : 2;
    }
    const bool y_lo_zero = fy && (j == 0) && (p.bc_lo[1] == BC_SYM);
    // row of the shared tile holding this thread's y-neighbour: tile row r holds grid row
    // j0 + r (forward, halo = row TY) or j0 - 1 + r (backward, halo = row 0).
    const int srow_self = fy ? ty : ty + 1;
    const int srow_nb = fy ? ty + 1 : ty;
    // halo row loader: the first TX*2 threads (linear id) fetch the halo row of Fx and Fz.
    const int lin = ty * TX + tx;
    const bool is_halo_loader = lin < 2 * TX;
    const int hcomp = lin / TX;  // 0 → Fx, 1 → Fz
    const int hx = lin % TX;
    const int hi_ = blockIdx.x * TX + hx;
    int hj;  // grid row of the halo
    bool hload = false;
    if (is_halo_loader && hi_ < Nx) {
        if (fy) {
            hj = blockIdx.y * TY + TY;  // may exceed the last row when the tile is ragged
            const int jlast = min(blockIdx.y * TY + TY, Ny) - 1;  // last valid row of this tile
            hj = jlast + 1;
            if (hj >= Ny) { hj = 0; hload = (p.bc_hi[1] == BC_BLOCH); }
            else hload = true;
        } else {
            hj = blockIdx.y * TY - 1;
            if (hj < 0) { hj = Ny - 1; hload = (p.bc_lo[1] == BC_BLOCH); }
            else hload = true;
        }
    }
    // In a ragged forward tile the halo row sits right after the last valid row.
    const int hrow = fy ? (min(blockIdx.y * TY + TY, Ny) - blockIdx.y * TY) : 0;
    const int64_t h_off = hload ? ((int64_t)hj * sy + hi_) : 0;
    const T *hsrc = (hcomp == 0) ? p.F[0] : p.F[2];

    const C cxi = valid ? p.c[0][i] : C{};
    const C cyj = valid ? p.c[1][j] : C{};

    const int k0 = p.k_begin + blockIdx.z * p.march;
    const int k1 = min(k0 + p.march, p.k_end);
    if (k0 >= k1) return;

    const T Z = zero_of(T{});

    // plane fetch helpers -----------------------------------------------------------------------
    // "hi" plane index for output plane k: forward k+1, backward k.   "lo": forward k, backward k-1.
    // Out-of-range hi/lo planes come from the ghost pointers (wrap plane or neighbour rank).
    auto load_xy_plane = [&](int comp, int kp) -> T {  // comp 0/1 = Fx/Fy, raw value of plane kp
        if (!valid) return Z;
        if (kp < 0 || kp >= Nz) return ldg_plain(p.ghost[comp] + ij);
        return ldg_stream(p.F[comp] + (int64_t)kp * sz + ij);
    };

    // prologue: lo planes of Fx, Fy for k0
    T xlo = load_xy_plane(0, fz ? k0 : k0 - 1);
    T ylo = load_xy_plane(1, fz ? k0 : k0 - 1);
    // prefetch registers for plane k0
    T xhi_n = load_xy_plane(0, fz ? k0 + 1 : k0);
    T yhi_n = load_xy_plane(1, fz ? k0 + 1 : k0);
    T zc_n = valid ? ldg_stream(p.F[2] + (int64_t)k0 * sz + ij) : Z;
    T yx_n = Z, zx_n = Z;   // x-edge elements of Fy, Fz (centre plane)
    T h_n = Z;              // halo-row element (Fx or Fz, centre plane)
    {
        const int kc = k0;  // centre plane
        if (xload) {
            // Fy centre plane: forward-z centre is plane k (== lo), backward-z centre is plane k (== hi)
            yx_n = ldg_stream(p.F[1] + (int64_t)kc * sz + xnb_off);
            zx_n = ldg_stream(p.F[2] + (int64_t)kc * sz + xnb_off);
        }
        if (hload) h_n = ldg_stream(hsrc + (int64_t)kc * sz + h_off);
    }

    int buf = 0;
    for (int k = k0; k < k1; ++k, buf ^= 1) {
        const T xhi = xhi_n, yhi = yhi_n, zc = zc_n, yxe = yx_n, zxe = zx_n, he = h_n;
        // ---- issue the loads of plane k+1 ------------------------------------------------------
        if (k + 1 < k1) {
            xhi_n = load_xy_plane(0, fz ? k + 2 : k + 1);
            yhi_n = load_xy_plane(1, fz ? k + 2 : k + 1);
            zc_n = valid ? ldg_stream(p.F[2] + (int64_t)(k + 1) * sz + ij) : Z;
            if (xload) {
                yx_n = ldg_stream(p.F[1] + (int64_t)(k + 1) * sz + xnb_off);
                zx_n = ldg_stream(p.F[2] + (int64_t)(k + 1) * sz + xnb_off);
            }
            if (hload) h_n = ldg_stream(hsrc + (int64_t)(k + 1) * sz + h_off);
        }
        // centre-plane values of Fx, Fy
        const T xc = fz ? xlo : xhi;
        const T yc = fz ? ylo : yhi;

        // ---- publish Fx, Fz centre values for the y-neighbour exchange ---------------------------
        if (valid) {  // rows beyond the grid stay unwritten: the halo row may live there
            sFx[buf][srow_self][tx] = xc;
            sFz[buf][srow_self][tx] = zc;
        }
        if (is_halo_loader) {
            if (hcomp == 0) sFx[buf][hrow][hx] = he;
            else sFz[buf][hrow][hx] = he;
        }
        __syncthreads();

        // ---- z terms:  cz[k]·(hi − lo) with boundary rules ----------------------------------------
        const C czk = p.c[2][k];
        T dzx, dzy;  // ∂z Fx, ∂z Fy (already multiplied by α·∆z⁻¹[k])
        {
            T xh = xhi, yh = yhi, xl = xlo, yl = ylo;
            bool zero_term = false;
            if (fz) {
                if (k == Nz - 1) {
                    if (p.bc_hi[2] == BC_BLOCH) { xh = mul_e(p.e[2], xh, p.e_real[2]); yh = mul_e(p.e[2], yh, p.e_real[2]); }
                    else if (p.bc_hi[2] == BC_SYM) { xh = Z; yh = Z; }
                }
                if (k == 0 && p.bc_lo[2] == BC_SYM) { xl = Z; yl = Z; }
            } else if (k == 0) {
                if (p.bc_lo[2] == BC_BLOCH) { xl = div_e(xl, p.e[2], p.e_real[2]); yl = div_e(yl, p.e[2], p.e_real[2]); }
                else if (p.bc_lo[2] == BC_SYM) zero_term = true;
            }
            dzx = zero_term ? Z : mul(czk, sub(xh, xl));
            dzy = zero_term ? Z : mul(czk, sub(yh, yl));
        }

        // ---- y terms: neighbours from shared memory ------------------------------------------------
        T dyx, dyz;  // ∂y Fx, ∂y Fz
        {
            T xn = sFx[buf][srow_nb][tx];
            T zn = sFz[buf][srow_nb][tx];
            T xh, xl, zh, zl;
            bool zero_term = false;
            if (fy) {
                xh = xn; zh = zn; xl = xc; zl = zc;
                if (ykind == 1) { xh = mul_e(p.e[1], xh, p.e_real[1]); zh = mul_e(p.e[1], zh, p.e_real[1]); }
                else if (ykind == 2) { xh = Z; zh = Z; }
                if (y_lo_zero) { xl = Z; zl = Z; }
            } else {
                xh = xc; zh = zc; xl = xn; zl = zn;
                if (ykind == 1) { xl = div_e(xl, p.e[1], p.e_real[1]); zl = div_e(zl, p.e[1], p.e_real[1]); }
                else if (ykind == 2) zero_term = true;
            }
            dyx = zero_term ? Z : mul(cyj, sub(xh, xl));
            dyz = zero_term ? Z : mul(cyj, sub(zh, zl));
        }

        // ---- x terms: neighbours from the adjacent lane --------------------------------------------
        T dxy, dxz;  // ∂x Fy, ∂x Fz
        {
            T yn = fx ? shfl_down1(yc) : shfl_up1(yc);
            T zn = fx ? shfl_down1(zc) : shfl_up1(zc);
            if (xload) { yn = yxe; zn = zxe; }
            T yh, yl, zh, zl;
            bool zero_term = false;
            if (fx) {
                yh = yn; zh = zn; yl = yc; zl = zc;
                if (xkind == 1) { yh = mul_e(p.e[0], yh, p.e_real[0]); zh = mul_e(p.e[0], zh, p.e_real[0]); }
                else if (xkind == 2) { yh = Z; zh = Z; }
                if (x_lo_zero) { yl = Z; zl = Z; }
            } else {
                yh = yc; zh = zc; yl = yn; zl = zn;
                if (xkind == 1) { yl = div_e(yl, p.e[0], p.e_real[0]); zl = div_e(zl, p.e[0], p.e_real[0]); }
                else if (xkind == 2) zero_term = true;
            }
            dxy = zero_term ? Z : mul(cxi, sub(yh, yl));
            dxz = zero_term ? Z : mul(cxi, sub(zh, zl));
        }

        // ---- combine in the reference's accumulation order (curl.jl:66-91) ------------------------
        //   Gx: −∂zFy then +∂yFz ;  Gy: +∂zFx then −∂xFz ;  Gz: −∂yFx then +∂xFy
        if (valid) {
            const int64_t o = (int64_t)k * sz + ij;
            T gx, gy, gz;
            if (ADD) {
                gx = add(add(ldg_plain(p.G[0] + o), neg(dzy)), dyz);
                gy = add(add(ldg_plain(p.G[1] + o), dzx), neg(dxz));
                gz = add(add(ldg_plain(p.G[2] + o), neg(dyx)), dxy);
            } else {
                gx = add(add(Z, neg(dzy)), dyz);  // Gv .= 0 ; += (−∂zFy) ; += ∂yFz
                gy = add(dzx, neg(dxz));
                gz = add(neg(dyx), dxy);
            }
            stg_stream(p.G[0] + o, gx);
            stg_stream(p.G[1] + o, gy);
            stg_stream(p.G[2] + o, gz);
        }
        // rotate planes
        xlo = xhi;
        ylo = yhi;
    }
}


// ---------------------------------------------------------------------------------------------
// Fused full 3-D curl, version 2: z-planes staged through a shared-memory ring by cp.async
// ---------------------------------------------------------------------------------------------
// Same mathematics and the same CurlParams as k_curl3.  The difference is how HBM latency is
// hidden: instead of holding the next plane in registers (which costs occupancy), every z-plane
// of the CTA's (TX+1)×(TY+1) tile — own cells plus the one-cell x/y halos — is copied
// global→shared by `cp.async` into a ring of S stages, S−2 planes ahead of the plane being
// computed.  Bytes in flight are then bounded by shared memory, not by registers.
//
//   plane slot j (j = 0..n) holds grid plane  base + j,  base = k0 (forward ∂z) or k0−1 (backward);
//   output plane k0+t uses  lo = slot t,  hi = slot t+1,  centre = lo (forward) / hi (backward).
//   Fz and the x/y halos are only needed on centre planes and are only copied for those.
//
// One __syncthreads per plane.  x-neighbours still travel by warp shuffle; only the lane at the
// warp / tile edge reads the halo column from shared memory.

SGC_D void cp_async_elem(cplx *dst_smem, const cplx *src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
}
SGC_D void cp_async_elem(double *dst_smem, const double *src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src) : "memory");
}
SGC_D void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
SGC_D void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// General (boundary-aware) evaluation of one cell of plane k from the staged planes.  This is
// the code path of k_curl3 with the neighbours coming from shared memory.
template <typename T, typename C>
struct CurlCellBC {
    int xkind, ykind;        // 0 plain, 1 Bloch wrap, 2 symmetric
    bool x_lo_zero, y_lo_zero;
};

// MODE selects which of the staged data and of the six ∂ terms are used:
//   CURL3_CURL  G = ∇×F                                   (slots 0,1,2 = Fx, Fy, Fz)
//   CURL3_DIVG  g = ∂xF_x + ∂yF_y + ∂zF_z  (apply_divg!, divergence.jl:41-66): slot 0 = F_z (the z-differenced
//               slot), slot 1 = F_x (x-differenced), slot 2 = F_y (y-differenced); one output, terms in x,y,z order
//   CURL3_GRAD  G_w = ∂w f                  (apply_grad!, gradient.jl:41-59): slot 0 = f, differenced along all axes
//   CURL3_MEAN  G_w = m̂_w F_w               (apply_mean!, mean.jl:47-89): slots as for the divergence — F_z is averaged along z
//               through the register carry, F_x along x, F_y along y — and three outputs (G[0] = G_z, G[1] = G_x, G[2] = G_y)
enum : int { CURL3_CURL = 0, CURL3_DIVG = 1, CURL3_GRAD = 2, CURL3_MEAN = 3 };
template <typename T, typename C, bool ADD, int TX, int TY, int S, int MINB, bool XSHFL, int MODE = CURL3_CURL>
__global__ void __launch_bounds__(TX *TY, MINB) k_curl3p(const __grid_constant__ CurlParams<T, C> p) {
    static_assert(TX % 32 == 0, "a warp must lie inside one row");
    static_assert(TX * TY >= 2 * TX + 2 * TY, "not enough threads for the halo copies");
    static_assert(S >= 3, "need at least 3 stages");
    constexpr int PITCH = TX + 1;
    constexpr int COMP = (TY + 1) * PITCH;
    constexpr int STAGE = 3 * COMP;
    constexpr bool DIVLIKE = (MODE == CURL3_DIVG || MODE == CURL3_MEAN);  // staging of the divergence
    extern __shared__ __align__(16) unsigned char sgc_smem_raw[];
    T *const sm = reinterpret_cast<T *>(sgc_smem_raw);

    const int tx = threadIdx.x, ty = threadIdx.y;
    const int lane = tx & 31;
    const int Nx = p.Nx, Ny = p.Ny, Nz = p.Nz;
    const int i0 = blockIdx.x * TX, j0 = blockIdx.y * TY;
    const int i = i0 + tx, j = j0 + ty;
    const bool valid = (i < Nx) && (j < Ny);
    const int nvx = min(TX, Nx - i0), nvy = min(TY, Ny - j0);  // valid extent of this tile
    const int64_t sy = Nx, sz = (int64_t)Nx * Ny;
    const int64_t ij = (int64_t)j * sy + i;
    const int fx = p.fwd[0], fy = p.fwd[1], fz = p.fwd[2];
    // a tile that touches no x/y boundary and is full needs no boundary logic at all
    const bool tile_plain = (i0 > 0) && (i0 + TX < Nx) && (j0 > 0) && (j0 + TY < Ny);

    // shared-tile coordinates: the halo row/column sits right after the last valid row/column
    // (forward) or at index 0 (backward)
    const int self = (ty + (fy ? 0 : 1)) * PITCH + (tx + (fx ? 0 : 1));
    const int ynb = self + (fy ? PITCH : -PITCH);
    const int xnb = self + (fx ? 1 : -1);

    // ---- x / y boundary description of this thread (identical to k_curl3) ------------------------
    int xkind = 0, ykind = 0;  // 0 plain, 1 Bloch wrap, 2 symmetric
    bool x_lo_zero = false, y_lo_zero = false;
    if (!tile_plain) {
        if (fx) { if (i == Nx - 1) xkind = (p.bc_hi[0] == BC_BLOCH) ? 1 : 2; }
        else    { if (i == 0) xkind = (p.bc_lo[0] == BC_BLOCH) ? 1 : 2; }
        if (fy) { if (j == Ny - 1) ykind = (p.bc_hi[1] == BC_BLOCH) ? 1 : 2; }
        else    { if (j == 0) ykind = (p.bc_lo[1] == BC_BLOCH) ? 1 : 2; }
        x_lo_zero = fx && (i == 0) && (p.bc_lo[0] == BC_SYM);
        y_lo_zero = fy && (j == 0) && (p.bc_lo[1] == BC_SYM);
    }
    // the lane whose x-neighbour is not in its own warp reads it from the shared tile
    const bool xedge = fx ? (lane == 31 || tx == nvx - 1) : (lane == 0);

    // ---- halo copy duty of this thread (at most one element per plane) ---------------------------
    const int lin = ty * TX + tx;
    int hcomp = -1, hdst = 0;
    int64_t hsrc = 0;
    if (lin < 2 * TX) {  // y-halo row of Fx (lin < TX) and Fz
        const int x = (lin < TX) ? lin : lin - TX;
        if (x < nvx) {
            int hj = fy ? j0 + nvy : j0 - 1;
            bool ok = true;
            if (hj >= Ny) { hj = 0; ok = (p.bc_hi[1] == BC_BLOCH); }
            if (hj < 0) { hj = Ny - 1; ok = (p.bc_lo[1] == BC_BLOCH); }
            if (ok) {
                hcomp = (lin < TX) ? 0 : 2;  // y-halo row of slot 0 and slot 2
                if (DIVLIKE && hcomp == 0) hcomp = -1;  // divergence: only slot 2 (F_y) is y-differenced
                if (MODE == CURL3_GRAD && hcomp == 2) hcomp = -1;  // gradient: only slot 0 (f)
                hdst = (fy ? nvy : 0) * PITCH + (x + (fx ? 0 : 1));
                hsrc = (int64_t)hj * sy + (i0 + x);
            }
        }
    } else if (lin < 2 * TX + 2 * TY) {  // x-halo column of Fy and Fz
        const int q = lin - 2 * TX;
        const int y = (q < TY) ? q : q - TY;
        if (y < nvy) {
            int hi = fx ? i0 + nvx : i0 - 1;
            bool ok = true;
            if (hi >= Nx) { hi = 0; ok = (p.bc_hi[0] == BC_BLOCH); }
            if (hi < 0) { hi = Nx - 1; ok = (p.bc_lo[0] == BC_BLOCH); }
            if (ok) {
                hcomp = (q < TY) ? 1 : 2;  // x-halo column of slot 1 and slot 2
                if (DIVLIKE && hcomp == 2) hcomp = -1;  // divergence: only slot 1 (F_x) is x-differenced
                if (MODE == CURL3_GRAD) hcomp = (q < TY) ? 0 : -1;  // gradient: the x-halo of slot 0 (f)
                hdst = (y + (fy ? 0 : 1)) * PITCH + (fx ? nvx : 0);
                hsrc = (int64_t)(j0 + y) * sy + hi;
            }
        }
    }
    const T *const hbase = (hcomp >= 0) ? p.F[hcomp] + hsrc : nullptr;
    T *const hsm = sm + (hcomp >= 0 ? hcomp * COMP + hdst : 0);

    // (arithmetic means carry no coefficient tables)
    const bool has_cx = (MODE != CURL3_MEAN) || p.m[0].weighted, has_cy = (MODE != CURL3_MEAN) || p.m[1].weighted;
    const C cxi = (valid && has_cx) ? p.c[0][i] : C{};
    const C cyj = (valid && has_cy) ? p.c[1][j] : C{};
    // weighted means of an interior cell: ∆x[i], ∆x[i±1], ∆y[j], ∆y[j±1]
    C mwx = C{}, mwxn = C{}, mwy = C{}, mwyn = C{};
    if (MODE == CURL3_MEAN && valid) {
        if (p.m[0].weighted) { mwx = p.m[0].w[i]; mwxn = p.m[0].w[min(max(i + (fx ? 1 : -1), 0), Nx - 1)]; }
        if (p.m[1].weighted) { mwy = p.m[1].w[j]; mwyn = p.m[1].w[min(max(j + (fy ? 1 : -1), 0), Ny - 1)]; }
    }
    // backward difference c·(cur − nb) == (−c)·(nb − cur) exactly: fold the direction into the sign
    const C cxs = fx ? cxi : neg(cxi);
    const C cys = fy ? cyj : neg(cyj);

    // synchronised applies run the two boundary chunks last (chunk order 1, 2, ..., nzc−1, 0)
    const int zchunk = p.sync_ctrl ? ((blockIdx.z + 1 == gridDim.z) ? 0 : (int)blockIdx.z + 1) : (int)blockIdx.z;
    const int k0 = p.k_begin + zchunk * p.march;
    const int k1 = min(k0 + p.march, p.k_end);
    if (k0 >= k1) return;
    const int n = k1 - k0;              // planes to produce
    const int base = fz ? k0 : k0 - 1;  // grid plane of slot 0
    const T Z = zero_of(T{});

    if (p.sync_ctrl && (k0 == 0 || k1 == Nz)) {
        // this CTA reads a neighbour's plane or overwrites a plane a neighbour may still be reading: the
        // neighbour on that side must have finished the boundary CTAs of its previous apply
        if (lin == 0) {
            const long long done = ld_volatile_s64(p.sync_ctrl + SGC_CTRL_EPOCH);  // my finished applies
            if (k0 == 0 && p.sync_side[0])
                wait_flag_sys(p.sync_ctrl + SGC_CTRL_FLAG_LO, done, p.sync_ctrl, 64);
            if (k1 == Nz && p.sync_side[1])
                wait_flag_sys(p.sync_ctrl + SGC_CTRL_FLAG_HI, done, p.sync_ctrl, 64);
        }
        __syncthreads();
    }

    // running state of the producer side: slot index, its ring stage, its grid plane offset
    int jj_issue = 0;
    int stage_issue = 0;
    int64_t off_issue = (int64_t)base * sz;  // element offset of the plane of slot jj_issue
    auto issue = [&]() {
        if (jj_issue <= n) {
            T *st = sm + stage_issue * STAGE;
            const int kp = base + jj_issue;
            const bool centre = fz ? (jj_issue < n) : (jj_issue >= 1);
            if (kp >= 0 && kp < Nz) {
                if (valid) {
                    cp_async_elem(st + self, p.F[0] + off_issue + ij);
                    if (MODE == CURL3_CURL || (DIVLIKE && centre)) cp_async_elem(st + COMP + self, p.F[1] + off_issue + ij);
                    if (MODE != CURL3_GRAD && centre) cp_async_elem(st + 2 * COMP + self, p.F[2] + off_issue + ij);
                }
                if (centre && hcomp >= 0) cp_async_elem(hsm + stage_issue * STAGE, hbase + off_issue);
            } else if (valid) {  // plane outside the slab: wrap-around plane or the neighbour rank's plane
                cp_async_elem(st + self, p.ghost[0] + ij);
                if (MODE == CURL3_CURL) cp_async_elem(st + COMP + self, p.ghost[1] + ij);
            }
        }
        cp_async_commit();
        ++jj_issue;
        off_issue += sz;
        stage_issue = (stage_issue + 1 == S) ? 0 : stage_issue + 1;
    };

#pragma unroll
    for (int q = 0; q < S - 1; ++q) issue();

    T xlo = Z, ylo = Z;
    int stage_lo = 0;
    int64_t off_out = (int64_t)k0 * sz + ij;
    for (int t = 0; t < n; ++t, off_out += sz) {
        const int k = k0 + t;
        T g0 = Z, g1 = Z, g2 = Z;
        if (ADD && valid) {  // issue the read-modify-write loads early
            g0 = ldg_plain(p.G[0] + off_out);
            if (MODE != CURL3_DIVG) { g1 = ldg_plain(p.G[1] + off_out); g2 = ldg_plain(p.G[2] + off_out); }
        }
        cp_async_wait<S - 3>();  // slots ≤ t+1 have landed (for this thread) ...
        __syncthreads();         // ... and for everybody; everybody has also finished plane t−1
        issue();                 // refill the stage that plane t−1 used

        const int stage_hi = (stage_lo + 1 == S) ? 0 : stage_lo + 1;
        const T *slo = sm + stage_lo * STAGE;
        const T *shi = sm + stage_hi * STAGE;
        const T *sc = fz ? slo : shi;
        stage_lo = stage_hi;
        if (t == 0) { xlo = slo[self]; ylo = slo[COMP + self]; }
        const T xhi = shi[self], yhi = shi[COMP + self];
        const T xc = fz ? xlo : xhi;
        // slot 1 is staged on centre planes only for the divergence: read it there, not through the carry
        const T yc = (MODE == CURL3_GRAD) ? Z : (DIVLIKE ? sc[COMP + self] : (fz ? ylo : yhi));
        const T zc = (MODE == CURL3_GRAD) ? Z : sc[2 * COMP + self];
        const C czk = p.c[2][k];
        T gx, gy, gz;

        if constexpr (MODE == CURL3_MEAN) {
            // slot 0 = F_z averaged along z (register carry), slot 1 = F_x along x, slot 2 = F_y along y
            const T zcur = fz ? xlo : xhi, znb = fz ? xhi : xlo;
            const T xnbv = sc[COMP + xnb], ynbv = sc[2 * COMP + ynb];
            T r0, r1, r2;
            if (tile_plain && k > 0 && k < Nz - 1) {
                const bool wz = p.m[2].weighted != 0;
                const C mcz = wz ? p.c[2][k] : C{}, mwz = wz ? p.m[2].w[k] : C{}, mwzn = wz ? p.m[2].w[fz ? k + 1 : k - 1] : C{};
                r0 = mean_plain<T, C>(wz, p.m[2].a2, mcz, mwz, mwzn, zcur, znb);
                r1 = mean_plain<T, C>(p.m[0].weighted != 0, p.m[0].a2, cxi, mwx, mwxn, yc, xnbv);
                r2 = mean_plain<T, C>(p.m[1].weighted != 0, p.m[1].a2, cyj, mwy, mwyn, zc, ynbv);
            } else {
                r0 = mean_eval<T, C>(p, 2, k, Nz, zcur, znb);
                r1 = mean_eval<T, C>(p, 0, valid ? i : 0, Nx, yc, xnbv);
                r2 = mean_eval<T, C>(p, 1, valid ? j : 0, Ny, zc, ynbv);
            }
            gx = ADD ? add(g0, r0) : r0;
            gy = ADD ? add(g1, r1) : r1;
            gz = ADD ? add(g2, r2) : r2;
        } else if (tile_plain && k > 0 && k < Nz - 1) {
            // ---- fast path: no boundary in reach; every term is c·(nb − cur) ------------------------
            if (MODE == CURL3_CURL) {
                const T xn = sc[ynb], zn = sc[2 * COMP + ynb];
                T yn, zx;
                if (XSHFL) {
                    yn = fx ? shfl_down1(yc) : shfl_up1(yc);
                    zx = fx ? shfl_down1(zc) : shfl_up1(zc);
                    if (xedge) { yn = sc[COMP + xnb]; zx = sc[2 * COMP + xnb]; }
                } else {
                    yn = sc[COMP + xnb];
                    zx = sc[2 * COMP + xnb];
                }
                const T dzx = mul(czk, sub(xhi, xlo));
                const T dzy = mul(czk, sub(yhi, ylo));
                const T dyx = mul(cys, sub(xn, xc));
                const T dyz = mul(cys, sub(zn, zc));
                const T dxy = mul(cxs, sub(yn, yc));
                const T dxz = mul(cxs, sub(zx, zc));
                if (ADD) {
                    gx = add(add(g0, neg(dzy)), dyz);
                    gy = add(add(g1, dzx), neg(dxz));
                    gz = add(add(g2, neg(dyx)), dxy);
                } else {
                    gx = add(add(Z, neg(dzy)), dyz);
                    gy = add(dzx, neg(dxz));
                    gz = add(neg(dyx), dxy);
                }
            } else if (MODE == CURL3_DIVG) {
                const T dz = mul(czk, sub(xhi, xlo));                 // ∂z of slot 0 (F_z)
                const T dy = mul(cys, sub(sc[2 * COMP + ynb], zc));    // ∂y of slot 2 (F_y)
                const T dx = mul(cxs, sub(sc[COMP + xnb], yc));        // ∂x of slot 1 (F_x)
                gx = ADD ? add(add(add(g0, dx), dy), dz) : add(add(dx, dy), dz);  // divergence.jl:52-63: x, then y, then z
                gy = Z; gz = Z;
            } else {
                gx = mul(cxs, sub(sc[xnb], xc));   // ∂x f
                gy = mul(cys, sub(sc[ynb], xc));   // ∂y f
                gz = mul(czk, sub(xhi, xlo));      // ∂z f
                if (ADD) { gx = add(g0, gx); gy = add(g1, gy); gz = add(g2, gz); }
            }
        } else {
            // ---- general path: boundary rules folded in (same code as k_curl3) -----------------------
            // z term of slot 0 (and slot 1), y terms of slot 0 / slot 2, x terms of slot 1 / slot 2 (slot 0 for GRAD)
            T dzx, dzy;
            {
                T xh = xhi, yh = yhi, xl = xlo, yl = ylo;
                bool zero_term = false;
                if (fz) {
                    if (k == Nz - 1) {
                        if (p.bc_hi[2] == BC_BLOCH) { xh = mul_e(p.e[2], xh, p.e_real[2]); yh = mul_e(p.e[2], yh, p.e_real[2]); }
                        else if (p.bc_hi[2] == BC_SYM) { xh = Z; yh = Z; }
                    }
                    if (k == 0 && p.bc_lo[2] == BC_SYM) { xl = Z; yl = Z; }
                } else if (k == 0) {
                    if (p.bc_lo[2] == BC_BLOCH) { xl = div_e(xl, p.e[2], p.e_real[2]); yl = div_e(yl, p.e[2], p.e_real[2]); }
                    else if (p.bc_lo[2] == BC_SYM) zero_term = true;
                }
                dzx = zero_term ? Z : mul(czk, sub(xh, xl));
                dzy = zero_term ? Z : mul(czk, sub(yh, yl));
            }
            T dyx, dyz;
            {
                const T xn = sc[ynb], zn = sc[2 * COMP + ynb];
                T xh, xl, zh, zl;
                bool zero_term = false;
                if (fy) {
                    xh = xn; zh = zn; xl = xc; zl = zc;
                    if (ykind == 1) { xh = mul_e(p.e[1], xh, p.e_real[1]); zh = mul_e(p.e[1], zh, p.e_real[1]); }
                    else if (ykind == 2) { xh = Z; zh = Z; }
                    if (y_lo_zero) { xl = Z; zl = Z; }
                } else {
                    xh = xc; zh = zc; xl = xn; zl = zn;
                    if (ykind == 1) { xl = div_e(xl, p.e[1], p.e_real[1]); zl = div_e(zl, p.e[1], p.e_real[1]); }
                    else if (ykind == 2) zero_term = true;
                }
                dyx = zero_term ? Z : mul(cyj, sub(xh, xl));
                dyz = zero_term ? Z : mul(cyj, sub(zh, zl));
            }
            T dxy, dxz;
            {
                // first operand pair: slot 1 (slot 0 for the gradient); second: slot 2
                const T c1v = (MODE == CURL3_GRAD) ? xc : yc;
                T yn = (MODE == CURL3_GRAD) ? sc[xnb] : sc[COMP + xnb], zn = sc[2 * COMP + xnb];
                if (XSHFL && MODE == CURL3_CURL) {
                    const T ys = fx ? shfl_down1(yc) : shfl_up1(yc);
                    const T zs = fx ? shfl_down1(zc) : shfl_up1(zc);
                    if (!xedge) { yn = ys; zn = zs; }
                }
                T yh, yl, zh, zl;
                bool zero_term = false;
                if (fx) {
                    yh = yn; zh = zn; yl = c1v; zl = zc;
                    if (xkind == 1) { yh = mul_e(p.e[0], yh, p.e_real[0]); zh = mul_e(p.e[0], zh, p.e_real[0]); }
                    else if (xkind == 2) { yh = Z; zh = Z; }
                    if (x_lo_zero) { yl = Z; zl = Z; }
                } else {
                    yh = c1v; zh = zc; yl = yn; zl = zn;
                    if (xkind == 1) { yl = div_e(yl, p.e[0], p.e_real[0]); zl = div_e(zl, p.e[0], p.e_real[0]); }
                    else if (xkind == 2) zero_term = true;
                }
                dxy = zero_term ? Z : mul(cxi, sub(yh, yl));
                dxz = zero_term ? Z : mul(cxi, sub(zh, zl));
            }
            if (MODE == CURL3_CURL) {
                if (ADD) {
                    gx = add(add(g0, neg(dzy)), dyz);
                    gy = add(add(g1, dzx), neg(dxz));
                    gz = add(add(g2, neg(dyx)), dxy);
                } else {
                    gx = add(add(Z, neg(dzy)), dyz);
                    gy = add(dzx, neg(dxz));
                    gz = add(neg(dyx), dxy);
                }
            } else if (MODE == CURL3_DIVG) {
                gx = ADD ? add(add(add(g0, dxy), dyz), dzx) : add(add(dxy, dyz), dzx);
                gy = Z; gz = Z;
            } else {
                gx = dxy; gy = dyx; gz = dzx;
                if (ADD) { gx = add(g0, gx); gy = add(g1, gy); gz = add(g2, gz); }
            }
        }
        if (valid) {
            stg_stream(p.G[0] + off_out, gx);
            if (MODE != CURL3_DIVG) {
                stg_stream(p.G[1] + off_out, gy);
                stg_stream(p.G[2] + off_out, gz);
            }
        }
        xlo = xhi;
        ylo = yhi;
    }
    cp_async_wait<0>();
    if (p.sync_ctrl && (k0 == 0 || k1 == Nz)) {
        // per side: the last boundary CTA to finish tells that neighbour that this rank's apply has written (and
        // no longer reads) the planes at that interface; the last side advances this rank's own count
        __syncthreads();
        if (lin == 0) {
            __threadfence();
            const unsigned side_total = gridDim.x * gridDim.y;
            const unsigned nsides = (unsigned)(p.sync_side[0] + p.sync_side[1]);
            unsigned *const tickets = reinterpret_cast<unsigned *>(p.sync_ctrl + SGC_CTRL_TICKETS);
            unsigned *const sides_done = reinterpret_cast<unsigned *>(p.sync_ctrl + SGC_CTRL_SIDES_DONE);
            const long long done = ld_volatile_s64(p.sync_ctrl + SGC_CTRL_EPOCH);  // unchanged until all sides are through
#pragma unroll
            for (int sd = 0; sd < 2; ++sd) {
                const bool mine = sd == 0 ? (k0 == 0) : (k1 == Nz);
                if (mine && p.sync_side[sd] && atomicAdd(tickets + sd, 1u) == side_total - 1) {
                    tickets[sd] = 0;
                    __threadfence_system();
                    st_release_sys(p.sync_post[sd], done + 1);
                    if (atomicAdd(sides_done, 1u) == nsides - 1) {
                        *sides_done = 0;
                        __threadfence();
                        st_volatile_s64(p.sync_ctrl + SGC_CTRL_EPOCH, done + 1);
                    }
                }
            }
        }
    }
}

}  // namespace sgc
