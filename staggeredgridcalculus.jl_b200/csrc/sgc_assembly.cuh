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
//   → device-wide exclusive prefix scan (colptr) → the triplets are sorted by row inside the
//   column segment (a register sorting network: the segmented sort of a ≤ 6-element segment)
//   → staged in shared memory → written to rowval / nzval with fully coalesced stores.
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
};

SGC_D bool is_zero(double v) { return v == 0.0; }
SGC_D bool is_zero(cplx v) { return v.re == 0.0 && v.im == 0.0; }

// Subscripts of grid cell c and the block column of matrix column `col`.
template <typename T, typename IDX>
SGC_D void column_decode(const AsmParams<T> &p, int64_t col, int &nu, IDX &c, IDX sub[3]) {
    if (p.cmpfirst) {
        const IDX kin = (IDX)p.Kin;
        c = kin == 3 ? (IDX)col / 3 : (kin == 2 ? (IDX)col / 2 : (IDX)col);
        nu = (int)((IDX)col - c * kin);
    } else {
        nu = (int)((IDX)col / (IDX)p.M);
        c = (IDX)col - (IDX)nu * (IDX)p.M;
    }
    const IDX n0 = (IDX)p.N[0], n1 = (IDX)p.N[1];
    const IDX t = c / n0;
    sub[0] = c - t * n0;
    sub[2] = t / n1;
    sub[1] = t - sub[2] * n1;
}

// The (≤ 2) triplets block B contributes to the column of grid cell c: value, LOCAL row of the
// (block of the) matrix, and validity (nonzero value, row inside the partition).
template <typename T, typename IDX>
SGC_D void block_entries(const AsmParams<T> &p, const AsmBlock<T> &B, IDX c, const IDX sub[3], T v[2], int64_t r[2],
                         bool ok[2]) {
    const int a = B.axis;
    const int64_t Nw = p.N[a];
    const int64_t cw = (int64_t)sub[a];
    const int64_t stride = a == 0 ? 1 : (a == 1 ? p.N[0] : p.N[0] * p.N[1]);
    ok[0] = true;
    ok[1] = false;
    r[0] = (int64_t)c;
    r[1] = (int64_t)c;
    v[1] = zero_of(T{});
    if (B.ns == 0) {  // diagonal-only (create_πcmp)
        v[0] = B.V0[0];
    } else if (Nw == 1) {
        // diagonal and off-diagonal triplets coincide; `sparse` sums them in input order
        v[0] = add(B.V0[0], B.Vs[0]);
    } else {
        // off-diagonal triplet of row r lands in column r shifted by +ns  ⇒  r = c − ns
        int64_t rw = cw - B.ns;
        if (rw < 0) rw += Nw;
        if (rw >= Nw) rw -= Nw;
        v[0] = B.V0[cw];
        v[1] = B.Vs[rw];
        r[1] = (int64_t)c + (rw - cw) * stride;
        ok[1] = true;
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int64_t grow = p.cmpfirst ? (int64_t)p.Kout * r[q] + B.nv : r[q] + p.M * B.nv;
        ok[q] = ok[q] && !is_zero(v[q]) && grow >= p.row_begin && grow < p.row_end;  // dropzeros!, partition
        r[q] = grow - p.row_begin;
    }
}

// Number of stored entries of one column.
template <typename T, typename IDX>
SGC_D int column_count(const AsmParams<T> &p, int64_t col) {
    int nu;
    IDX c, sub[3];
    column_decode<T, IDX>(p, col, nu, c, sub);
    int cnt = 0;
    for (int b = 0; b < p.nblk; ++b) {
        const AsmBlock<T> &B = p.blk[b];
        if (B.nu != nu) continue;
        T v[2];
        int64_t r[2];
        bool ok[2];
        block_entries<T, IDX>(p, B, c, sub, v, r, ok);
        cnt += (int)ok[0] + (int)ok[1];
    }
    return cnt;
}

template <typename T>
SGC_D void cmpswap(int64_t &ra, T &va, int64_t &rb, T &vb) {
    if (rb < ra) {
        const int64_t tr = ra; ra = rb; rb = tr;
        const T tv = va; va = vb; vb = tv;
    }
}

// The entries of one column sorted by row, in registers: invalid slots carry row = INT64_MAX and
// sink to the end of a fixed 6-input sorting network (the "segmented sort" of a ≤ 6-entry
// column segment).  Returns the count.
template <typename T, typename IDX>
SGC_D int column_entries(const AsmParams<T> &p, int64_t col, int64_t (&rows)[ASM_MAX_ENTRIES], T (&vals)[ASM_MAX_ENTRIES]) {
    int nu;
    IDX c, sub[3];
    column_decode<T, IDX>(p, col, nu, c, sub);
    constexpr int64_t INVALID = 0x7fffffffffffffffLL;
#pragma unroll
    for (int q = 0; q < ASM_MAX_ENTRIES; ++q) { rows[q] = INVALID; vals[q] = zero_of(T{}); }
    int cnt = 0, slot = 0;  // slot: which pair of the register array the next matching block fills
    for (int b = 0; b < p.nblk; ++b) {
        const AsmBlock<T> &B = p.blk[b];
        if (B.nu != nu) continue;
        T v[2];
        int64_t r[2];
        bool ok[2];
        block_entries<T, IDX>(p, B, c, sub, v, r, ok);
        cnt += (int)ok[0] + (int)ok[1];
        // static register indexing: the block-column has at most 3 blocks
#pragma unroll
        for (int s3 = 0; s3 < 3; ++s3)
            if (slot == s3) {
                if (ok[0]) { rows[2 * s3] = r[0]; vals[2 * s3] = v[0]; }
                if (ok[1]) { rows[2 * s3 + 1] = r[1]; vals[2 * s3 + 1] = v[1]; }
            }
        ++slot;
    }
    // 6-input sorting network (12 compare-exchanges)
    cmpswap(rows[0], vals[0], rows[5], vals[5]); cmpswap(rows[1], vals[1], rows[3], vals[3]); cmpswap(rows[2], vals[2], rows[4], vals[4]);
    cmpswap(rows[1], vals[1], rows[2], vals[2]); cmpswap(rows[3], vals[3], rows[4], vals[4]);
    cmpswap(rows[0], vals[0], rows[3], vals[3]); cmpswap(rows[2], vals[2], rows[5], vals[5]);
    cmpswap(rows[0], vals[0], rows[1], vals[1]); cmpswap(rows[2], vals[2], rows[3], vals[3]); cmpswap(rows[4], vals[4], rows[5], vals[5]);
    cmpswap(rows[1], vals[1], rows[2], vals[2]); cmpswap(rows[3], vals[3], rows[4], vals[4]);
    return cnt;
}

// Functor for the scan's input iterator: number of stored entries of column `col`
// (0 for the one-past-the-end column so that the scan also produces colptr[n]).
template <typename T, typename IDX>
struct ColumnCount {
    AsmParams<T> p;
    __device__ int64_t operator()(int64_t col) const {
        if (col >= p.ncol) return 0;
        return (int64_t)column_count<T, IDX>(p, col);
    }
};

// Fill: one thread per column, entries staged through shared memory so that the global
// stores of rowval / nzval are contiguous per CTA.
template <typename T, typename IDX, int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_asm_fill(const __grid_constant__ AsmParams<T> p,
                                                    const int64_t *__restrict__ colptr,
                                                    int64_t *__restrict__ rowval, T *__restrict__ nzval,
                                                    int max_per_col) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *s_vals = reinterpret_cast<T *>(smem_raw);
    int64_t *s_rows = reinterpret_cast<int64_t *>(smem_raw + sizeof(T) * (size_t)BLOCK * max_per_col);

    const int64_t col0 = (int64_t)blockIdx.x * BLOCK;
    const int64_t col = col0 + threadIdx.x;
    const int64_t col_end = min(col0 + (int64_t)BLOCK, p.ncol);
    const int64_t base = colptr[col0] - p.index_base;
    const int64_t total = colptr[col_end] - p.index_base - base;

    if (col < p.ncol) {
        int64_t rows[ASM_MAX_ENTRIES];
        T vals[ASM_MAX_ENTRIES];
        const int cnt = column_entries<T, IDX>(p, col, rows, vals);
        const int64_t off = colptr[col] - p.index_base - base;
#pragma unroll
        for (int q = 0; q < ASM_MAX_ENTRIES; ++q)
            if (q < cnt) {
                s_rows[off + q] = rows[q] + p.index_base;
                s_vals[off + q] = vals[q];
            }
    }
    __syncthreads();
    for (int64_t q = threadIdx.x; q < total; q += BLOCK) {
        rowval[base + q] = s_rows[q];
        stg_stream(nzval + base + q, s_vals[q]);
    }
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
