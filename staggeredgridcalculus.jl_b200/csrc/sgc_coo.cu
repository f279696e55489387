// sgc_coo.cu — the generic assembly path: COO triplets → sort → CSC.
//
// This is the device restatement of what the reference does literally: every block emits its
// 2M triplets `[diag; off-diag]` in the reference's order (matrix/differential/partial.jl:303-311,
// curl.jl:89-98), the triplets are sorted by (column,row) with a STABLE radix sort so that
// duplicates stay in input order, duplicates are summed, stored zeros are dropped
// (`dropzeros!(sparse(...))`), and colptr is produced by a search over the sorted keys.
// It needs ~48 bytes of scratch per triplet and several passes, so it is the cross-check for
// the closed-form kernels of sgc_assembly.cuh, not the fast path.
#include "sgc_internal.h"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_reduce.cuh>
#include <cub/device/device_select.cuh>

#include "sgc_assembly.cuh"

using namespace sgc;

namespace {

template <typename T>
__global__ void k_emit_coo(AsmParams<T> p, uint64_t *keys, T *vals, int64_t m_rows_total) {
    // one thread per (block, half, r): half 0 = diagonal triplet, 1 = off-diagonal triplet
    const int64_t per_blk = 2 * p.M;
    const int64_t total = per_blk * p.nblk;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
        const int b = (int)(q / per_blk);
        const int64_t w = q - (int64_t)b * per_blk;
        const int half = w >= p.M;
        const int64_t r = half ? w - p.M : w;
        const AsmBlock<T> &B = p.blk[b];
        int64_t sub[3];
        int64_t t = r / p.N[0];
        sub[0] = r - t * p.N[0];
        sub[2] = t / p.N[1];
        sub[1] = t - sub[2] * p.N[1];
        const int64_t stride[3] = {1, p.N[0], p.N[0] * p.N[1]};
        const int a = B.axis;
        const int64_t Nw = p.N[a], rw = sub[a];
        int64_t c = r;
        T v;
        if (B.ns == 0) {
            v = half ? zero_of(T{}) : B.V0[0];
        } else if (!half) {
            v = B.V0[rw];
        } else {
            int64_t cw = rw + B.ns;  // Jₛ = circshift(I, -ns·ŵ): column of the off-diagonal entry of row r
            if (cw < 0) cw += Nw;
            if (cw >= Nw) cw -= Nw;
            c = r + (cw - rw) * stride[a];
            v = B.Vs[rw];
        }
        const int64_t grow = p.cmpfirst ? (int64_t)p.Kout * r + B.nv : r + p.M * B.nv;
        const int64_t gcol = p.cmpfirst ? (int64_t)p.Kin * c + B.nu : c + p.M * B.nu;
        keys[q] = (uint64_t)gcol * (uint64_t)m_rows_total + (uint64_t)grow;
        vals[q] = v;
    }
}

template <typename T>
struct SumOp {
    __device__ T operator()(const T &a, const T &b) const { return add(a, b); }
};

// keep[i] = value != 0 and row inside the partition
template <typename T>
__global__ void k_flag(const uint64_t *keys, const T *vals, int64_t n, uint64_t m_total, int64_t row_begin,
                       int64_t row_end, unsigned char *keep) {
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = (int64_t)(keys[q] % m_total);
        keep[q] = (!is_zero(vals[q]) && row >= row_begin && row < row_end) ? 1 : 0;
    }
}

template <typename T>
__global__ void k_finalize(const uint64_t *keys, const T *vals, int64_t nnz, uint64_t m_total, int64_t row_begin,
                           int64_t base, int64_t ncol, int64_t *colptr, int64_t *rowval, T *nzval) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nnz; q += stride) {
        rowval[q] = (int64_t)(keys[q] % m_total) - row_begin + base;
        nzval[q] = vals[q];
    }
    // colptr[c] = base + number of kept keys with column < c   (lower bound on c·m_total)
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c <= ncol; c += stride) {
        const uint64_t target = (uint64_t)c * m_total;
        int64_t lo = 0, hi = nnz;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (keys[mid] < target) lo = mid + 1;
            else hi = mid;
        }
        colptr[c] = lo + base;
    }
}

struct Scratch {
    std::vector<void *> ptrs;
    ~Scratch() { for (void *p : ptrs) cudaFree(p); }
    template <typename U> int alloc(U **p, size_t n) {
        void *q = nullptr;
        if (cudaMalloc(&q, (n ? n : 1) * sizeof(U)) != cudaSuccess) return sgc_fail(SGC_ERR_NOMEM, "out of device memory in the COO path");
        ptrs.push_back(q);
        *p = (U *)q;
        return SGC_OK;
    }
};

template <typename T>
int coo_sort_typed(sgc_asm *a, int64_t *colptr, int index_base, int64_t *rowval, T *nzval, int64_t cap,
                   int64_t *nnz_out, cudaStream_t st) {
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
    const uint64_t m_total = (uint64_t)a->Kout * (uint64_t)a->M;
    const int64_t E = 2 * a->M * p.nblk;
    REQUIRE(E < (int64_t)0x7fffffff, "COO cross-check path is limited to 2^31 triplets");
    Scratch S;
    uint64_t *k_in, *k_out, *k_uni, *k_sel;
    T *v_in, *v_out, *v_uni, *v_sel;
    unsigned char *keep;
    int64_t *d_num;
    int s;
    if ((s = S.alloc(&k_in, E)) || (s = S.alloc(&k_out, E)) || (s = S.alloc(&k_uni, E)) || (s = S.alloc(&k_sel, E)) ||
        (s = S.alloc(&v_in, E)) || (s = S.alloc(&v_out, E)) || (s = S.alloc(&v_uni, E)) || (s = S.alloc(&v_sel, E)) ||
        (s = S.alloc(&keep, E)) || (s = S.alloc(&d_num, 2)))
        return s;
    if (E == 0) {
        CUDA_TRY(cudaMemsetAsync(colptr, 0, sizeof(int64_t) * (p.ncol + 1), st));
    }
    const unsigned blocks = (unsigned)std::min<int64_t>((E + 255) / 256 + 1, 148 * 32);
    k_emit_coo<T><<<blocks, 256, 0, st>>>(p, k_in, v_in, (int64_t)m_total);
    if ((s = sgc_check_launch("k_emit_coo"))) return s;

    size_t tmp_bytes = 0, need = 0;
    void *tmp = nullptr;
    // stable sort by key = col·m + row
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, need, k_in, k_out, v_in, v_out, (int)E, 0, 64, st));
    tmp_bytes = need;
    size_t need2 = 0, need3 = 0;
    CUDA_TRY(cub::DeviceReduce::ReduceByKey(nullptr, need2, k_out, k_uni, v_out, v_uni, d_num, SumOp<T>(), (int)E, st));
    CUDA_TRY(cub::DeviceSelect::Flagged(nullptr, need3, k_uni, keep, k_sel, d_num + 1, (int)E, st));
    tmp_bytes = std::max(tmp_bytes, std::max(need2, need3));
    { unsigned char *t8; if ((s = S.alloc(&t8, tmp_bytes))) return s; tmp = t8; }
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k_in, k_out, v_in, v_out, (int)E, 0, 64, st));
    // sum duplicates (input order is preserved by the stable sort)
    CUDA_TRY(cub::DeviceReduce::ReduceByKey(tmp, tmp_bytes, k_out, k_uni, v_out, v_uni, d_num, SumOp<T>(), (int)E, st));
    int64_t h_num[2] = {0, 0};
    {
        int n_unique = 0;
        // ReduceByKey writes an int-sized count when num_items is int
        CUDA_TRY(cudaMemcpyAsync(&n_unique, d_num, sizeof(int), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        h_num[0] = n_unique;
    }
    const int64_t U = h_num[0];
    // dropzeros! and the row partition
    k_flag<T><<<blocks, 256, 0, st>>>(k_uni, v_uni, U, m_total, a->row_begin, a->row_end, keep);
    if ((s = sgc_check_launch("k_flag"))) return s;
    CUDA_TRY(cub::DeviceSelect::Flagged(tmp, tmp_bytes, k_uni, keep, k_sel, (int *)(d_num + 1), (int)U, st));
    CUDA_TRY(cub::DeviceSelect::Flagged(tmp, tmp_bytes, v_uni, keep, v_sel, (int *)(d_num + 1), (int)U, st));
    int n_sel = 0;
    CUDA_TRY(cudaMemcpyAsync(&n_sel, d_num + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    sgc_count_launches(8);
    const int64_t nnz = n_sel;
    if (nnz_out) *nnz_out = nnz;
    if (nnz > cap) return sgc_fail(SGC_ERR_INVALID, "nnz = %lld exceeds the capacity %lld of rowval/nzval", (long long)nnz, (long long)cap);
    const unsigned fblocks = (unsigned)std::min<int64_t>((std::max(nnz, p.ncol + 1) + 255) / 256, 148 * 32);
    k_finalize<T><<<fblocks, 256, 0, st>>>(k_sel, v_sel, nnz, m_total, a->row_begin, index_base, p.ncol, colptr, rowval, nzval);
    if ((s = sgc_check_launch("k_finalize"))) return s;
    CUDA_TRY(cudaStreamSynchronize(st));  // scratch is freed on return
    return SGC_OK;
}

}  // namespace

extern "C" int sgc_asm_coo_sort(sgc_asm *a, int64_t *colptr, int index_base, int64_t *rowval, void *nzval,
                                int64_t nnz_capacity, int64_t *nnz, void *stream) {
    if (a) a->last_stream = (cudaStream_t)stream;
    REQUIRE(a && colptr && ((rowval && nzval) || nnz_capacity == 0), "NULL argument");
    REQUIRE(index_base == 0 || index_base == 1, "index_base must be 0 or 1");
    if (a->cx) return coo_sort_typed<cplx>(a, colptr, index_base, rowval, (cplx *)nzval, nnz_capacity, nnz, (cudaStream_t)stream);
    return coo_sort_typed<double>(a, colptr, index_base, rowval, (double *)nzval, nnz_capacity, nnz, (cudaStream_t)stream);
}
