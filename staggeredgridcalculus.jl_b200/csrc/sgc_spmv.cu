// sgc_spmv.cu — y = A·x for an assembled SparseMatrixCSC on the device: the `mul!(g, Curl, f)` the reference's tests
// use as the cross-check of the matrix-free operators (test/curl.jl:55, test/partial.jl:67-70) and the step after
// assembly in a solver.
//
// CSC is the wrong way round for a product that writes y: a column-parallel kernel has to scatter with atomics (the
// round-1 utility, 169 G atomics/s).  The matrix is applied many times for one assembly, so the transpose structure
// is built ONCE (sgc_spmv_create): the entries are stably sorted by row (cub radix sort), which gives for every row
// the positions of its entries in the CSC arrays in ascending column order, plus their column indices.  The product
// (sgc_spmv_apply) is then a row-parallel GATHER: y[r] = Σ nzval[perm[q]]·x[col[q]] accumulated from 0 in ascending
// column order — the order of Julia's column sweep — with no atomics; for the banded operators of this library the
// gathered nzval / x entries of neighbouring rows are neighbours in memory, so HBM sees every array once.
#include "sgc_internal.h"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

using namespace sgc;

struct sgc_spmv {
    int64_t m = 0, n = 0, nnz = 0;
    int index_base = 1;
    int64_t *rowptr = nullptr;   // m + 1, 0-based positions into perm / tcol
    uint32_t *perm = nullptr;    // nnz: position of the entry in the CSC arrays
    uint32_t *tcol = nullptr;    // nnz: its (0-based) column
};

namespace {

__global__ void k_expand_cols(int64_t n, const int64_t *__restrict__ colptr, const int64_t *__restrict__ rowval, int64_t base,
                              uint32_t *__restrict__ cols, uint32_t *__restrict__ rows, uint32_t *__restrict__ ids,
                              unsigned long long *__restrict__ rowcount) {
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (int64_t)gridDim.x * blockDim.x) {
        for (int64_t q = colptr[c] - base; q < colptr[c + 1] - base; ++q) {
            const uint32_t r = (uint32_t)(rowval[q] - base);
            cols[q] = (uint32_t)c;
            rows[q] = r;
            ids[q] = (uint32_t)q;
            atomicAdd(rowcount + r, 1ull);
        }
    }
}

__global__ void k_gather_u32(int64_t nnz, const uint32_t *__restrict__ src, const uint32_t *__restrict__ idx, uint32_t *__restrict__ dst) {
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nnz; q += (int64_t)gridDim.x * blockDim.x) dst[q] = src[idx[q]];
}

template <typename T>
__global__ void __launch_bounds__(256) k_spmv_rows(int64_t m, const int64_t *__restrict__ rowptr, const uint32_t *__restrict__ perm,
                                                   const uint32_t *__restrict__ tcol, const T *__restrict__ nz, const T *__restrict__ x,
                                                   T *__restrict__ y) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    const int64_t q0 = rowptr[r], q1 = rowptr[r + 1];
    T s = zero_of(T{});
    for (int64_t q = q0; q < q1; ++q) s = add(s, mul(nz[perm[q]], x[tcol[q]]));  // y[r] += A[r,c]·x[c], c ascending
    y[r] = s;
}

}  // namespace

extern "C" int sgc_spmv_destroy(sgc_spmv *p) {
    if (!p) return SGC_OK;
    if (p->rowptr) cudaFree(p->rowptr);
    if (p->perm) cudaFree(p->perm);
    if (p->tcol) cudaFree(p->tcol);
    delete p;
    return SGC_OK;
}

extern "C" int sgc_spmv_create(sgc_spmv **plan, int64_t m, int64_t n, const int64_t *colptr, const int64_t *rowval, int index_base,
                               void *stream) {
    REQUIRE(plan && colptr && rowval, "NULL argument");
    REQUIRE(index_base == 0 || index_base == 1, "index_base must be 0 or 1");
    REQUIRE(m >= 0 && n >= 0 && m < (int64_t)0xffffffffLL && n < (int64_t)0xffffffffLL, "matrix dimensions must fit 32 bits");
    *plan = nullptr;
    cudaStream_t st = (cudaStream_t)stream;
    int64_t last = 0;
    CUDA_TRY(cudaMemcpyAsync(&last, colptr + n, sizeof last, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    const int64_t nnz = last - index_base;
    REQUIRE(nnz >= 0 && nnz < (int64_t)0xffffffffLL, "nnz = %lld does not fit 32 bits", (long long)nnz);
    sgc_spmv *p = new sgc_spmv();
    p->m = m; p->n = n; p->nnz = nnz; p->index_base = index_base;
    uint32_t *cols = nullptr, *rows = nullptr, *ids = nullptr, *rows2 = nullptr;
    void *tmp = nullptr, *tmp2 = nullptr;
    auto fail_all = [&](int code, const char *what, cudaError_t e) {
        for (void *q : {(void *)cols, (void *)rows, (void *)ids, (void *)rows2, tmp, tmp2}) if (q) cudaFree(q);
        sgc_spmv_destroy(p);
        return sgc_fail(code, "sgc_spmv_create: %s: %s", what, cudaGetErrorString(e));
    };
#define SPMV_TRY(expr, what) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return fail_all(_e == cudaErrorMemoryAllocation ? SGC_ERR_NOMEM : SGC_ERR_CUDA, what, _e); } while (0)
    const size_t nn = (size_t)std::max<int64_t>(nnz, 1);
    SPMV_TRY(cudaMalloc((void **)&p->rowptr, sizeof(int64_t) * (size_t)(m + 2)), "rowptr");
    SPMV_TRY(cudaMalloc((void **)&p->perm, sizeof(uint32_t) * nn), "perm");
    SPMV_TRY(cudaMalloc((void **)&p->tcol, sizeof(uint32_t) * nn), "tcol");
    SPMV_TRY(cudaMalloc((void **)&cols, sizeof(uint32_t) * nn), "cols");
    SPMV_TRY(cudaMalloc((void **)&rows, sizeof(uint32_t) * nn), "rows");
    SPMV_TRY(cudaMalloc((void **)&ids, sizeof(uint32_t) * nn), "ids");
    SPMV_TRY(cudaMalloc((void **)&rows2, sizeof(uint32_t) * nn), "sorted rows");
    // per-row counts accumulate in rowptr[1 .. m], the exclusive scan over [0 .. m] then gives the row starts
    SPMV_TRY(cudaMemsetAsync(p->rowptr, 0, sizeof(int64_t) * (size_t)(m + 2), st), "memset");
    if (n > 0 && nnz > 0) {
        const unsigned blocks = (unsigned)std::min<int64_t>((n + 255) / 256, 148 * 16);
        k_expand_cols<<<blocks, 256, 0, st>>>(n, colptr, rowval, index_base, cols, rows, ids, (unsigned long long *)p->rowptr);
        { int s = sgc_check_launch("k_expand_cols"); if (s) { fail_all(SGC_ERR_CUDA, "launch", cudaSuccess); return s; } }
        int bits = 1;
        while (bits < 32 && ((int64_t)1 << bits) < m) ++bits;
        size_t need = 0;
        SPMV_TRY(cub::DeviceRadixSort::SortPairs(nullptr, need, rows, rows2, ids, p->perm, (int)nnz, 0, bits, st), "sort size");
        SPMV_TRY(cudaMalloc(&tmp, need), "sort scratch");
        SPMV_TRY(cub::DeviceRadixSort::SortPairs(tmp, need, rows, rows2, ids, p->perm, (int)nnz, 0, bits, st), "sort");  // stable: columns stay ascending
        const unsigned gblocks = (unsigned)std::min<int64_t>((nnz + 255) / 256, 148 * 32);
        k_gather_u32<<<gblocks, 256, 0, st>>>(nnz, cols, p->perm, p->tcol);
        { int s = sgc_check_launch("k_gather_u32"); if (s) { fail_all(SGC_ERR_CUDA, "launch", cudaSuccess); return s; } }
        sgc_count_launches(1);
    }
    {   // rowptr[r] = Σ_{r' < r} count[r']: the counts sit in rowptr[0 .. m−1] (row r counted at index r), scan in place over m+1
        size_t need = 0;
        SPMV_TRY(cub::DeviceScan::ExclusiveSum(nullptr, need, p->rowptr, p->rowptr, (int)(m + 1), st), "scan size");
        SPMV_TRY(cudaMalloc(&tmp2, need), "scan scratch");
        SPMV_TRY(cub::DeviceScan::ExclusiveSum(tmp2, need, p->rowptr, p->rowptr, (int)(m + 1), st), "scan");
        sgc_count_launches(1);
    }
    SPMV_TRY(cudaStreamSynchronize(st), "sync");
#undef SPMV_TRY
    for (void *q : {(void *)cols, (void *)rows, (void *)ids, (void *)rows2, tmp, tmp2}) if (q) cudaFree(q);
    *plan = p;
    return SGC_OK;
}

extern "C" int sgc_spmv_apply(const sgc_spmv *p, const void *nzval, int val_dtype, const void *x, void *y, void *stream) {
    REQUIRE(p && nzval && x && y, "NULL argument");
    REQUIRE(val_dtype == SGC_F64 || val_dtype == SGC_C128, "val_dtype must be SGC_F64 or SGC_C128");
    if (p->m == 0) return SGC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned blocks = (unsigned)((p->m + 255) / 256);
    if (val_dtype == SGC_C128)
        k_spmv_rows<cplx><<<blocks, 256, 0, st>>>(p->m, p->rowptr, p->perm, p->tcol, (const cplx *)nzval, (const cplx *)x, (cplx *)y);
    else
        k_spmv_rows<double><<<blocks, 256, 0, st>>>(p->m, p->rowptr, p->perm, p->tcol, (const double *)nzval, (const double *)x, (double *)y);
    return sgc_check_launch("k_spmv_rows");
}
