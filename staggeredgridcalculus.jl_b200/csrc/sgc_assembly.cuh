// sgc_assembly.cuh — direct CSC assembly of the staggered-grid operators (sm_100a).
//
// The reference builds COO triplets with O(M) temporaries and hands them to
// `dropzeros!(sparse(I, J, V, m, n))` (matrix/differential/partial.jl:87-314, curl.jl:50-104,
// matrix/mean.jl:72-377).  Every block of every operator has the same shape — a diagonal and
// one off-diagonal shifted by ±1 along one axis, with values that depend only on the index
// along that axis — so one column of the final matrix holds at most 2 entries per block of its
// block-column.  The kernels below therefore work per COLUMN:
//
//   column → its ≤ 2·Kout candidate triplets (the same triplets the reference emits for that
//   column, duplicates summed in the reference's order, value==0 dropped) → count
//   → prefix scan (per CTA in-kernel, across CTAs a device scan of the CTA totals) = colptr
//   → the triplets are sorted by row inside the column segment (a register sorting network:
//   the segmented sort of a ≤ 6-element segment) → staged in shared memory → written to
//   rowval / nzval with fully coalesced stores.
//
// Index arithmetic is 64-bit throughout (512³ curl has 1.6e9 nonzeros).
#pragma once
#include "sgc_common.cuh"

namespace sgc {

constexpr int ASM_MAX_BLOCKS = 9;
constexpr int ASM_MAX_ENTRIES = 6;  // 2 per block, ≤ 3 blocks per block-column

template <typename T>
struct AsmBlock {
    const T *V0;    // diagonal value by row index along `axis`          (parity folded in)
    const T *Vs;    // off-diagonal value by ROW index along `axis`      (parity, Bloch, symmetry folded in)
    int nv, nu;     // block row / block column (0-based positions in cmp_out / cmp_in)
    int axis;       // array dimension of the shift
    int ns;         // +1 forward, -1 backward, 0: diagonal-only block (permutation)
};

// One entry of a column whose cell lies in the interior of the grid (no wrap-around, nothing dropped), in ascending
// row order: row = (Kout·c | c) + rowadd, value = V[sub[axis] + dsub].  The order of a column's entries only depends
// on the blocks' shifts and component positions there, so the host sorts them once (sgc_asm.cu: asm_params).
template <typename T>
struct AsmFastEntry {
    const T *V;
    int64_t rowadd;
    int axis, dsub;
};

template <typename T>
struct AsmParams {
    AsmBlock<T> blk[ASM_MAX_BLOCKS];
    int nblk;
    int Kout, Kin;
    int cmpfirst;
    int64_t N[3];
    int64_t M;
    int64_t ncol;                 // Kin*M
    int64_t row_begin, row_end;   // global rows kept (0-based, half-open)
    int64_t index_base;
    // cell-major kernels: the blocks of each block column, and the group geometry
    int nblk_of[3];
    int blk_of[3][3];
    int cols_per_group;           // Kin for component-major ordering (a group = one grid cell), else 1
    int64_t ngroups;              // M (component-major) or Kin*M (block ordering: group = (nu, cell))
    // closed-form entry offsets (all rows kept): the per-column count is separable,
    // count = a(i) + b(j) + c(k), so the number of entries before a cell is a closed form of the
    // prefix sums PA, PB, PC.  Set s = nu (block ordering) or 3 (all block columns summed).
    // row partition, component-major ordering: only the columns of cells within one z-plane (the
    // longest stencil reach) of the kept rows' cells, or in the first / last plane (Bloch wrap), can
    // hold entries; CTAs entirely outside [near_lo, near_hi) and the wrap planes skip the entry work
    int skip_enabled;
    int64_t near_lo, near_hi, wrap_span;
    const int *cnt_tab[3][3];     // [nu][axis] → int[N_axis]: separable per-column counts (valid when all rows are kept)
    int prefix_enabled;
    const int64_t *prefix[4][3];  // [set][axis] → int64[N_axis + 1], exclusive prefix sums
    int64_t set_base[4];          // entries before the first group of the set
    // column partition: groups [cta_first·BLOCK, g_end) are processed; outputs are re-based to the block
    int64_t cta_first, g_end, col_base, entry_base;
    // interior cells of CTAs whose columns are all full: entries straight from the host-sorted descriptors
    int fast_enabled;
    unsigned in_lo[3], in_span[3];  // interior ⇔ sub[a] − in_lo[a] ≤ in_span[a] (unsigned) on every axis
    AsmFastEntry<T> fast[3][ASM_MAX_ENTRIES];
};

// entries stored before group g (= before its first column) when all rows are kept
template <typename T, typename IDX>
SGC_D int64_t prefix_entries_before(const AsmParams<T> &p, int64_t g) {
    int set = 3;
    int64_t c = g;
    if (!(p.cols_per_group > 1 || p.Kin == 1)) {
        set = (g >= 2 * p.M) ? 2 : (g >= p.M ? 1 : 0);
        c = g - (int64_t)set * p.M;
    }
    const int64_t Nx = p.N[0], Ny = p.N[1];
    const IDX tq = (IDX)c / (IDX)Nx;  // (32-bit divisions when the matrix is small enough)
    const int64_t t = (int64_t)tq, i = c - t * Nx, k = (int64_t)(tq / (IDX)Ny), j = t - k * Ny;
    const int64_t *PA = p.prefix[set][0], *PB = p.prefix[set][1], *PC = p.prefix[set][2];
    const int64_t SA = PA[Nx], SB = PB[Ny];
    const int64_t bj = PB[j + 1] - PB[j], ck = PC[k + 1] - PC[k];
    return p.set_base[set] + k * (Ny * SA + Nx * SB) + Nx * Ny * PC[k]   // whole planes below k
           + j * SA + Nx * PB[j] + j * Nx * ck                          // whole rows below j in plane k
           + PA[i] + i * (bj + ck);                                      // cells before i in row (j,k)
}

SGC_D bool is_zero(double v) { return v == 0.0; }
SGC_D bool is_zero(cplx v) { return v.re == 0.0 && v.im == 0.0; }

// The (≤ 2) triplets block B contributes to the column of grid cell c: value, LOCAL row of the
// (block of the) matrix, and validity (nonzero value, row inside the partition).
// R is the row-arithmetic type: int32_t when the matrix has fewer than 2^31 rows and columns.
template <typename T, typename IDX, typename R>
SGC_D void block_entries(const AsmParams<T> &p, const AsmBlock<T> &B, IDX c, const IDX sub[3], T v[2], R r[2],
                         bool ok[2]) {
    const int a = B.axis;
    const R Nw = (R)p.N[a];
    const R cw = (R)sub[a];
    const R stride = a == 0 ? (R)1 : (a == 1 ? (R)p.N[0] : (R)(p.N[0] * p.N[1]));
    ok[0] = true;
    ok[1] = false;
    r[0] = (R)c;
    r[1] = (R)c;
    v[1] = zero_of(T{});
    if (B.ns == 0) {  // diagonal-only (create_πcmp)
        v[0] = B.V0[0];
    } else if (Nw == 1) {
        // diagonal and off-diagonal triplets coincide; `sparse` sums them in input order
        v[0] = add(B.V0[0], B.Vs[0]);
    } else {
        // off-diagonal triplet of row r lands in column r shifted by +ns  ⇒  r = c − ns
        R rw = cw - (R)B.ns;
        if (rw < 0) rw += Nw;
        if (rw >= Nw) rw -= Nw;
        v[0] = B.V0[cw];
        v[1] = B.Vs[rw];
        r[1] = (R)c + (rw - cw) * stride;
        ok[1] = true;
    }
    const R rb = (R)p.row_begin, re = (R)p.row_end;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const R grow = p.cmpfirst ? (R)p.Kout * r[q] + (R)B.nv : r[q] + (R)p.M * (R)B.nv;
        ok[q] = ok[q] && !is_zero(v[q]) && grow >= rb && grow < re;  // dropzeros!, partition
        r[q] = grow - rb;
    }
}

template <typename R, typename T>
SGC_D void cmpswap(R &ra, T &va, R &rb, T &vb) {
    if (rb < ra) {
        const R tr = ra; ra = rb; rb = tr;
        const T tv = va; va = vb; vb = tv;
    }
}

// ---------------------------------------------------------------------------------------------
// Cell-major assembly kernels (the fast path)
// ---------------------------------------------------------------------------------------------
// A "group" is the set of matrix columns one thread owns.  With the default component-major
// ordering (order_cmpfirst = true) the Kin columns of a grid cell are consecutive, so a group is
// a cell and a CTA's output is one contiguous range of colptr / rowval / nzval; with block
// ordering a group is a single column (nu, cell).  All threads of a warp then walk the same
// blocks (no divergence), subscripts come from two 32-bit divisions per thread, and the tables
// V₀/Vₛ are read with lane-contiguous or warp-uniform indices.
//
//   pass 1  k_asm_group<COUNT>  : per-CTA entry totals (nothing else is written)
//           cub::DeviceScan over the CTA totals (≤ a few MB)
//   pass 2  k_asm_group<FILL…>  : per-thread counts → block scan → colptr, and the sorted entries
//                                 staged in shared memory → coalesced rowval / nzval stores.
// HBM traffic is the algorithmic minimum: colptr, rowval, nzval are each written once and nothing
// of size O(nnz) is ever read.
enum : int { ASM_COUNT = 0, ASM_COLPTR = 1, ASM_FILL = 2, ASM_FILL_ALL = 3 };

template <typename T, typename IDX>
SGC_D void group_decode(const AsmParams<T> &p, int64_t g, int &nu0, IDX &c, IDX sub[3]) {
    if (p.cols_per_group > 1 || p.Kin == 1) {
        nu0 = 0;
        c = (IDX)g;
    } else {  // block ordering: g = nu*M + c, no division needed
        nu0 = (g >= 2 * p.M) ? 2 : (g >= p.M ? 1 : 0);
        c = (IDX)(g - (int64_t)nu0 * p.M);
    }
    const IDX n0 = (IDX)p.N[0], n1 = (IDX)p.N[1];
    const IDX t = c / n0;
    sub[0] = c - t * n0;
    sub[2] = t / n1;
    sub[1] = t - sub[2] * n1;
}

template <typename T, typename IDX, typename R>
SGC_D int column_count_of(const AsmParams<T> &p, int nu, IDX c, const IDX sub[3]) {
    int cnt = 0;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        if (q < p.nblk_of[nu]) {
            T v[2];
            R r[2];
            bool ok[2];
            block_entries<T, IDX, R>(p, p.blk[p.blk_of[nu][q]], c, sub, v, r, ok);
            cnt += (int)ok[0] + (int)ok[1];
        }
    }
    return cnt;
}

// Sorting networks over (row, value) pairs held in registers; invalid slots carry the largest
// row and sink to the end.  E = 2, 4 or 6 inputs.
template <int E, typename R, typename T>
SGC_D void sort_entries(R (&rows)[E], T (&vals)[E]) {
    if (E == 2) {
        cmpswap(rows[0], vals[0], rows[1], vals[1]);
    } else if (E == 4) {
        cmpswap(rows[0], vals[0], rows[1], vals[1]); cmpswap(rows[2], vals[2], rows[3], vals[3]);
        cmpswap(rows[0], vals[0], rows[2], vals[2]); cmpswap(rows[1], vals[1], rows[3], vals[3]);
        cmpswap(rows[1], vals[1], rows[2], vals[2]);
    } else {
        cmpswap(rows[0], vals[0], rows[5], vals[5]); cmpswap(rows[1], vals[1], rows[3], vals[3]); cmpswap(rows[2], vals[2], rows[4], vals[4]);
        cmpswap(rows[1], vals[1], rows[2], vals[2]); cmpswap(rows[3], vals[3], rows[4], vals[4]);
        cmpswap(rows[0], vals[0], rows[3], vals[3]); cmpswap(rows[2], vals[2], rows[5], vals[5]);
        cmpswap(rows[0], vals[0], rows[1], vals[1]); cmpswap(rows[2], vals[2], rows[3], vals[3]); cmpswap(rows[4], vals[4], rows[5], vals[5]);
        cmpswap(rows[1], vals[1], rows[2], vals[2]); cmpswap(rows[3], vals[3], rows[4], vals[4]);
    }
}

// Entries of column (nu, c) in registers, UNSORTED, plus the rank each one takes in ascending row
// order (invalid slots carry the largest row and rank last); returns the count.  Ranking by counting
// (E(E−1)/2 comparisons) instead of a sorting network keeps the values — up to 4 registers each — in
// place: they are stored straight to their sorted shared-memory slot.  R is the row type: 32-bit
// when the matrix has fewer than 2^31 rows (half the register and ALU cost), else 64-bit.
template <typename T, typename IDX, typename R, int E>
SGC_D int column_entries_of(const AsmParams<T> &p, int nu, IDX c, const IDX sub[3], R (&rows)[E], T (&vals)[E], int (&rank)[E]) {
    const R INVALID = (R)(sizeof(R) == 4 ? 0x7fffffff : 0x7fffffffffffffffLL);
#pragma unroll
    for (int q = 0; q < E; ++q) { rows[q] = INVALID; vals[q] = zero_of(T{}); }
    int cnt = 0;
#pragma unroll
    for (int q = 0; q < E / 2; ++q) {
        if (q < p.nblk_of[nu]) {
            T v[2];
            R r[2];
            bool ok[2];
            block_entries<T, IDX, R>(p, p.blk[p.blk_of[nu][q]], c, sub, v, r, ok);
            cnt += (int)ok[0] + (int)ok[1];
            if (ok[0]) { rows[2 * q] = r[0]; vals[2 * q] = v[0]; }
            if (ok[1]) { rows[2 * q + 1] = r[1]; vals[2 * q + 1] = v[1]; }
        }
    }
    if (sizeof(T) > 8) {  // wide values: rank, do not move them
#pragma unroll
        for (int e = 0; e < E; ++e) {
            int rk = 0;
#pragma unroll
            for (int f = 0; f < E; ++f)
                if (f != e) rk += (rows[f] < rows[e] || (rows[f] == rows[e] && f < e)) ? 1 : 0;
            rank[e] = rk;
        }
    } else {              // 8-byte values: the sorting network is cheaper (measured), slots are then compile-time
        sort_entries<E>(rows, vals);
#pragma unroll
        for (int e = 0; e < E; ++e) rank[e] = e;
    }
    return cnt;
}

// block-wide exclusive scan of one int per thread (BLOCK ≤ 1024, power of two multiple of 32)
template <int BLOCK>
SGC_D int block_exclusive_scan(int v, int *s_warp, int &total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += n;
    }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int w = (lane < BLOCK / 32) ? s_warp[lane] : 0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, w, d);
            if (lane >= d) w += n;
        }
        s_warp[32 + lane] = w;  // inclusive warp prefix
    }
    __syncthreads();
    total = s_warp[32 + BLOCK / 32 - 1];
    return incl - v + (wid ? s_warp[32 + wid - 1] : 0);
}

// NCOLS = columns per group (1..3), E = entry slots per column (2, 4 or 6).  In the FILL modes
// the group's entries are computed ONCE into registers (NCOLS·E ≤ 12 slots), then the block scan
// of the per-thread totals gives their position.
template <typename T, typename IDX, typename R, int BLOCK, int MODE, int NCOLS, int E>
__global__ void __launch_bounds__(BLOCK) k_asm_group(const __grid_constant__ AsmParams<T> p,
                                                     int64_t *__restrict__ cta_total,        // COUNT: out
                                                     const int64_t *__restrict__ cta_offset,  // COLPTR / FILL_ALL: in
                                                     const int64_t *__restrict__ colptr_in,   // FILL: in
                                                     int64_t *__restrict__ colptr_out,        // COLPTR / FILL_ALL: out
                                                     int64_t *__restrict__ rowval, T *__restrict__ nzval) {
    __shared__ int s_warp[64];
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr bool WRITES = (MODE == ASM_FILL || MODE == ASM_FILL_ALL);
    const int64_t g0 = ((int64_t)blockIdx.x + p.cta_first) * BLOCK;
    const int64_t g = g0 + threadIdx.x;
    if (p.skip_enabled) {
        const int64_t g1 = min(g0 + (int64_t)BLOCK, p.ngroups);
        const bool in_near = (g1 > p.near_lo) && (g0 < p.near_hi);
        const bool in_wrap = (g0 < p.wrap_span) || (g1 > p.ngroups - p.wrap_span);
        if (!in_near && !in_wrap) {  // no column of this CTA reaches a kept row
            if (MODE == ASM_COUNT) {
                if (threadIdx.x == 0) cta_total[blockIdx.x] = 0;
            } else if (MODE == ASM_COLPTR || MODE == ASM_FILL_ALL) {
                const int64_t v = cta_offset[blockIdx.x] + p.index_base;
                const int64_t col0s = g0 * NCOLS;
                for (int q = threadIdx.x; q < (int)(g1 - g0) * NCOLS; q += BLOCK) colptr_out[col0s + q] = v;
            }
            return;
        }
    }
    const bool live = g < p.g_end;
    int nu0 = 0;
    IDX c = 0, sub[3] = {0, 0, 0};
    if (live) group_decode<T, IDX>(p, g, nu0, c, sub);

    // 8-byte values: all entries of the group are produced once, up front, sorted by the network and held in
    // registers across the scan (fewest instructions; the kernel is issue-bound).  Wider values: per-column counts
    // first (from the separable tables when all rows are kept) and the entries one column at a time when they are
    // staged, ranked instead of moved — 56 instead of 96 registers for ComplexF64.
    constexpr bool LATE = sizeof(T) > 8;
    constexpr bool HOLD = WRITES && !LATE;
    // closed-form offsets available (all rows kept): the counts come from the separable tables for every value type, and
    // CTAs whose columns are all full emit the entries of interior cells from the host-sorted descriptors (p.fast) —
    // no block walk, no zero / range tests, no sorting; cells on a grid face keep the general path
    const bool hold = HOLD && !p.fast_enabled;
    int cnt[NCOLS];
    R hrows[HOLD ? NCOLS : 1][E];
    T hvals[HOLD ? NCOLS : 1][E];
    int mine = 0;
#pragma unroll
    for (int q = 0; q < NCOLS; ++q) {
        cnt[q] = 0;
        if (live) {
            if (HOLD && hold) {
                int rk[E];
                cnt[q] = column_entries_of<T, IDX, R, E>(p, nu0 + q, c, sub, hrows[HOLD ? q : 0], hvals[HOLD ? q : 0], rk);
            } else if (p.prefix_enabled) {
                cnt[q] = p.cnt_tab[nu0 + q][0][sub[0]] + p.cnt_tab[nu0 + q][1][sub[1]] + p.cnt_tab[nu0 + q][2][sub[2]];
            } else {
                cnt[q] = column_count_of<T, IDX, R>(p, nu0 + q, c, sub);
            }
        }
        mine += cnt[q];
    }
    int total;
    const int off0 = block_exclusive_scan<BLOCK>(mine, s_warp, total);
    if (MODE == ASM_COUNT) {
        if (threadIdx.x == 0) cta_total[blockIdx.x] = total;
        return;
    }
    // absolute (0-based) position of this CTA's first entry
    const int64_t col0 = g0 * NCOLS - p.col_base;  // (column index within the block of columns being assembled)
    const int64_t ncol_out = p.g_end * NCOLS - p.col_base;
    int64_t base;
    if (MODE == ASM_FILL) base = colptr_in[col0] - p.index_base;
    else if (p.prefix_enabled) base = prefix_entries_before<T, IDX>(p, g0) - p.entry_base;  // closed form: no count pass, no scan
    else base = cta_offset[blockIdx.x];
    constexpr int PADK = NCOLS * E;  // entry slots per thread
    if (total == BLOCK * PADK) {
        // ---- regular CTA: every column holds exactly E entries (no boundary drop in reach) --------
        // colptr is an arithmetic sequence; a thread's PADK entries are staged at tid·(PADK+1) + slot
        // (odd pitch: conflict-free) with compile-time offsets and leave in lane-contiguous stores.
        const int64_t ib = p.index_base;
        if (MODE == ASM_COLPTR || MODE == ASM_FILL_ALL) {
#pragma unroll
            for (int q = 0; q < NCOLS; ++q) {
                const int c = threadIdx.x + q * BLOCK;
                colptr_out[col0 + c] = base + ib + (int64_t)E * c;
            }
            if (g == p.g_end - 1) colptr_out[ncol_out] = base + ib + total;
            if (MODE == ASM_COLPTR) return;
        }
        if (WRITES) {
            constexpr int SLOTS = BLOCK * (PADK + 1);
            T *s_vals = reinterpret_cast<T *>(smem_raw);
            R *s_rows = reinterpret_cast<R *>(smem_raw + sizeof(T) * (size_t)SLOTS);
            const int s0 = threadIdx.x * (PADK + 1);
            const bool fast = p.fast_enabled && ((unsigned)sub[0] - p.in_lo[0] <= p.in_span[0]) && ((unsigned)sub[1] - p.in_lo[1] <= p.in_span[1]) &&
                              ((unsigned)sub[2] - p.in_lo[2] <= p.in_span[2]);
            const R rowbase = p.cmpfirst ? (R)p.Kout * (R)c : (R)c;
#pragma unroll
            for (int q = 0; q < NCOLS; ++q) {
                if (fast) {
#pragma unroll
                    for (int e = 0; e < E; ++e) {
                        const AsmFastEntry<T> &f = p.fast[nu0 + q][e];
                        const int64_t at = (int64_t)(f.axis == 0 ? sub[0] : (f.axis == 1 ? sub[1] : sub[2])) + f.dsub;
                        s_rows[s0 + q * E + e] = rowbase + (R)f.rowadd;
                        s_vals[s0 + q * E + e] = f.V[at];
                    }
                } else if (HOLD && hold) {
#pragma unroll
                    for (int e = 0; e < E; ++e) {
                        s_rows[s0 + q * E + e] = hrows[HOLD ? q : 0][e];
                        s_vals[s0 + q * E + e] = hvals[HOLD ? q : 0][e];
                    }
                } else {
                    R rows[E];
                    T vals[E];
                    int rank[E];
                    column_entries_of<T, IDX, R, E>(p, nu0 + q, c, sub, rows, vals, rank);
#pragma unroll
                    for (int e = 0; e < E; ++e) {  // entry e goes to its rank within the column
                        const int slot = s0 + q * E + rank[e];
                        s_rows[slot] = rows[e];
                        s_vals[slot] = vals[e];
                    }
                }
            }
            __syncthreads();
            int64_t *rv = rowval + base + threadIdx.x;
            T *nz = nzval + base + threadIdx.x;
#pragma unroll
            for (int it = 0; it < PADK; ++it) {
                const int q = threadIdx.x + it * BLOCK;
                const int slot = q + q / PADK;
                rv[it * BLOCK] = (int64_t)s_rows[slot] + ib;
                stg_stream(nz + it * BLOCK, s_vals[slot]);
            }
        }
        return;
    }
    if (MODE == ASM_COLPTR || MODE == ASM_FILL_ALL) {
        // colptr of the CTA's columns: staged so that the stores are lane-contiguous
        __shared__ int64_t s_colptr[BLOCK * NCOLS];
        if (live) {
            int64_t o = base + off0 + p.index_base;
#pragma unroll
            for (int q = 0; q < NCOLS; ++q) { s_colptr[threadIdx.x * NCOLS + q] = o; o += cnt[q]; }
            if (g == p.g_end - 1) colptr_out[ncol_out] = o;  // one-past-the-end entry
        }
        __syncthreads();
        const int64_t ncols_cta = min((int64_t)BLOCK * NCOLS, ncol_out - col0);
        for (int q = threadIdx.x; q < ncols_cta; q += BLOCK) colptr_out[col0 + q] = s_colptr[q];
        if (MODE == ASM_COLPTR) return;
    }
    if (WRITES) {
        // Entry q of the CTA is staged at slot q + q/PADK (the same layout as the regular case)
        constexpr int SLOTS = BLOCK * (PADK + 1);
        T *s_vals = reinterpret_cast<T *>(smem_raw);
        R *s_rows = reinterpret_cast<R *>(smem_raw + sizeof(T) * (size_t)SLOTS);
        if (live) {
            int o = off0;
#pragma unroll
            for (int q = 0; q < NCOLS; ++q) {
                if (HOLD && hold) {
#pragma unroll
                    for (int e = 0; e < E; ++e)
                        if (e < cnt[q]) {
                            const int slot = (o + e) + (o + e) / PADK;
                            s_rows[slot] = hrows[HOLD ? q : 0][e];
                            s_vals[slot] = hvals[HOLD ? q : 0][e];
                        }
                } else {
                    R rows[E];
                    T vals[E];
                    int rank[E];
                    column_entries_of<T, IDX, R, E>(p, nu0 + q, c, sub, rows, vals, rank);
#pragma unroll
                    for (int e = 0; e < E; ++e)
                        if (rank[e] < cnt[q]) {  // valid entries rank before the invalid ones
                            const int at = o + rank[e];
                            const int slot = at + at / PADK;
                            s_rows[slot] = rows[e];
                            s_vals[slot] = vals[e];
                        }
                }
                o += cnt[q];
            }
        }
        __syncthreads();
        const int64_t ib = p.index_base;
        for (int q = threadIdx.x; q < total; q += BLOCK) {
            const int slot = q + q / PADK;
            rowval[base + q] = (int64_t)s_rows[slot] + ib;
            stg_stream(nzval + base + q, s_vals[slot]);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Warp-granular one-pass assembly for the closed-form case (all rows kept; also a block of columns).
// The closed form gives the output offset of ANY group, so nothing couples the warps of a CTA: each warp takes 32
// consecutive groups, derives its own base offset, and stages its entries in a private slice of shared memory —
// no CTA-wide barrier, no block scan (k_asm_group: 3 barriers per CTA, 22 % of its stall samples).  A warp whose
// columns are all full (`__all_sync`) writes an arithmetic colptr and emits interior cells from the host-sorted
// descriptors; any other warp runs the general entry path behind an intra-warp shuffle scan.
// ---------------------------------------------------------------------------------------------
// CELLS = false: a thread owns the NCOLS = Kin columns of one grid cell (component-major ordering).
// CELLS = true : a thread owns NCOLS consecutive columns of ONE block column (block ordering, or a single-column operator:
//                create_∂, create_m̂, create_grad) — consecutive cells along x, Nx a multiple of NCOLS so that they never leave
//                the row; the per-warp offset arithmetic and the per-thread subscript divisions are then shared by
//                NCOLS·E entries instead of E (create_∂ 384³: 0.75 → see profiles/r2_time_asm_cells.jsonl).
template <typename T, typename IDX, typename R, int BLOCK, int NCOLS, int E, bool CELLS>
__global__ void __launch_bounds__(BLOCK) k_asm_warp(const __grid_constant__ AsmParams<T> p, int64_t *__restrict__ colptr_out,
                                                    int64_t *__restrict__ rowval, T *__restrict__ nzval) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int PADK = NCOLS * E;
    constexpr int WSLOTS = 32 * (PADK + 1);
    constexpr int NWARP = BLOCK / 32;
    constexpr int CPG = CELLS ? NCOLS : 1;  // matrix columns per group index of AsmParams (CELLS: a group there is ONE column)
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    T *s_vals = reinterpret_cast<T *>(smem_raw) + wid * WSLOTS;
    R *s_rows = reinterpret_cast<R *>(smem_raw + sizeof(T) * (size_t)WSLOTS * NWARP) + wid * WSLOTS;
    // thread index in units of NCOLS columns; g_end of the parameters counts groups of cols_per_group columns
    const int64_t t_end = CELLS ? p.g_end / NCOLS : p.g_end;
    const int64_t t0 = (CELLS ? (int64_t)blockIdx.x : (int64_t)blockIdx.x + p.cta_first) * BLOCK + wid * 32;  // the warp's first thread index
    if (t0 >= t_end) return;
    const int64_t t = t0 + lane;
    const bool live = t < t_end;
    int nu0 = 0;
    IDX c = 0, sub[3] = {0, 0, 0};
    if (live) group_decode<T, IDX>(p, t * CPG, nu0, c, sub);
    int cnt[NCOLS];
    int mine = 0;
#pragma unroll
    for (int q = 0; q < NCOLS; ++q) {
        if (CELLS) cnt[q] = live ? p.cnt_tab[nu0][0][sub[0] + q] + p.cnt_tab[nu0][1][sub[1]] + p.cnt_tab[nu0][2][sub[2]] : 0;
        else cnt[q] = live ? p.cnt_tab[nu0 + q][0][sub[0]] + p.cnt_tab[nu0 + q][1][sub[1]] + p.cnt_tab[nu0 + q][2][sub[2]] : 0;
        mine += cnt[q];
    }
    const int64_t ib = p.index_base;
    const int64_t base = prefix_entries_before<T, IDX>(p, t0 * CPG) - p.entry_base;
    const int64_t col0 = t0 * NCOLS - p.col_base;
    const int64_t ncol_out = t_end * NCOLS - p.col_base;
    // column q of this thread: block column, cell and subscripts
    auto col_of = [&](int q, int &nu, IDX &cc, IDX (&ss)[3]) {
        nu = CELLS ? nu0 : nu0 + q;
        cc = CELLS ? c + (IDX)q : c;
        ss[0] = CELLS ? sub[0] + (IDX)q : sub[0]; ss[1] = sub[1]; ss[2] = sub[2];
    };
    if (__all_sync(0xffffffffu, mine == PADK)) {
        // ---- every column of the warp holds exactly E entries ----------------------------------------------------
        // (thread-major staging: thread `lane` owns the slots lane·(PADK+1) …, i.e. the warp's columns lane·NCOLS + q)
#pragma unroll
        for (int q = 0; q < NCOLS; ++q) {
            const int cc = lane + q * 32;
            colptr_out[col0 + cc] = base + ib + (int64_t)E * cc;
        }
        if (t == t_end - 1) colptr_out[ncol_out] = base + ib + 32 * PADK;
        const int s0 = lane * (PADK + 1);
        const bool in_yz = p.fast_enabled && ((unsigned)sub[1] - p.in_lo[1] <= p.in_span[1]) && ((unsigned)sub[2] - p.in_lo[2] <= p.in_span[2]);
#pragma unroll
        for (int q = 0; q < NCOLS; ++q) {
            int nu;
            IDX cc, ss[3];
            col_of(q, nu, cc, ss);
            if (in_yz && ((unsigned)ss[0] - p.in_lo[0] <= p.in_span[0])) {
                const R rowbase = p.cmpfirst ? (R)p.Kout * (R)cc : (R)cc;
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const AsmFastEntry<T> &f = p.fast[nu][e];
                    const int64_t at = (int64_t)(f.axis == 0 ? ss[0] : (f.axis == 1 ? ss[1] : ss[2])) + f.dsub;
                    s_rows[s0 + q * E + e] = rowbase + (R)f.rowadd;
                    s_vals[s0 + q * E + e] = f.V[at];
                }
            } else {
                R rows[E];
                T vals[E];
                int rank[E];
                column_entries_of<T, IDX, R, E>(p, nu, cc, ss, rows, vals, rank);
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const int slot = s0 + q * E + rank[e];
                    s_rows[slot] = rows[e];
                    s_vals[slot] = vals[e];
                }
            }
        }
        __syncwarp();
        int64_t *rv = rowval + base + lane;
        T *nz = nzval + base + lane;
#pragma unroll
        for (int it = 0; it < PADK; ++it) {
            const int q = lane + it * 32;
            const int slot = q + q / PADK;
            rv[it * 32] = (int64_t)s_rows[slot] + ib;
            stg_stream(nz + it * 32, s_vals[slot]);
        }
        return;
    }
    // ---- general warp: entries dropped at a symmetric face, duplicates on an axis of length 1, or a ragged end ----
    int incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += n;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    const int off0 = incl - mine;
    if (live) {
        int64_t o = base + off0 + ib;
        int at0 = off0;
#pragma unroll
        for (int q = 0; q < NCOLS; ++q) {
            colptr_out[col0 + (int64_t)lane * NCOLS + q] = o;
            o += cnt[q];
            int nu;
            IDX cc, ss[3];
            col_of(q, nu, cc, ss);
            R rows[E];
            T vals[E];
            int rank[E];
            column_entries_of<T, IDX, R, E>(p, nu, cc, ss, rows, vals, rank);
#pragma unroll
            for (int e = 0; e < E; ++e)
                if (rank[e] < cnt[q]) {  // valid entries rank before the invalid ones
                    const int at = at0 + rank[e];
                    const int slot = at + at / PADK;
                    s_rows[slot] = rows[e];
                    s_vals[slot] = vals[e];
                }
            at0 += cnt[q];
        }
        if (t == t_end - 1) colptr_out[ncol_out] = o;  // one-past-the-end entry
    }
    __syncwarp();
    for (int q = lane; q < total; q += 32) {
        const int slot = q + q / PADK;
        rowval[base + q] = (int64_t)s_rows[slot] + ib;
        stg_stream(nzval + base + q, s_vals[slot]);
    }
}

// Count pass when all rows are kept (no row partition): the number of stored entries of a column
// is separable, count(nu, c) = A[nu][0][i] + A[nu][1][j] + A[nu][2][k], where A[nu][a][·] counts
// the nonzero diagonal / off-diagonal values of the blocks of block column nu that act along
// axis a (host-built from the same V₀/Vₛ tables, O(ΣN)).  The kernel only sums three small table
// entries per column and reduces them per CTA.
struct AsmCountTables {
    const int *A[3][3];  // [nu][axis] → int[N_axis]
};

template <typename IDX, int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_asm_count_fast(AsmCountTables t, int cols_per_group, int Kin, int64_t M,
                                                          int64_t ngroups, IDX n0, IDX n1, int64_t *__restrict__ cta_total) {
    __shared__ int s_warp[64];
    const int64_t g = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    int mine = 0;
    if (g < ngroups) {
        int nu0 = 0;
        IDX c;
        if (cols_per_group > 1 || Kin == 1) c = (IDX)g;
        else { nu0 = (g >= 2 * M) ? 2 : (g >= M ? 1 : 0); c = (IDX)(g - (int64_t)nu0 * M); }
        const IDX q = c / n0, i = c - q * n0, k = q / n1, j = q - k * n1;
        for (int nu = nu0; nu < nu0 + cols_per_group; ++nu) mine += t.A[nu][0][i] + t.A[nu][1][j] + t.A[nu][2][k];
    }
    int total;
    block_exclusive_scan<BLOCK>(mine, s_warp, total);
    if (threadIdx.x == 0) cta_total[blockIdx.x] = total;
}

// y += A x   (CSC, one thread per column) — test utility mirroring `mul!(g, Curl, f)`.
SGC_D void atomic_add(double *p, double v) { atomicAdd(p, v); }
SGC_D void atomic_add(cplx *p, cplx v) {
    atomicAdd(&p->re, v.re);
    atomicAdd(&p->im, v.im);
}
template <typename T>
__global__ void k_csc_matvec(int64_t n, const int64_t *colptr, const int64_t *rowval, const T *nzval,
                             int64_t base, const T *x, T *y) {
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (int64_t)gridDim.x * blockDim.x) {
        const T xc = x[c];
        for (int64_t q = colptr[c] - base; q < colptr[c + 1] - base; ++q)
            atomic_add(y + (rowval[q] - base), mul(nzval[q], xc));
    }
}

}  // namespace sgc
