// sgc_comm.cu — the multi-GPU context of libsgc_b200: peer-mapped device memory for the
// z-slabs of a decomposed grid, a device-side barrier of all ranks, and the per-rank control block through which
// the fused curl kernel synchronises with its two z-neighbours (sgc_kernels.cuh, CurlParams::sync_*).
//
// Nothing here depends on torch or NCCL.  Two ways to reach the neighbours' memory:
//   * one process per GPU (torchrun, MPI, Distributed.jl): every rank allocates with sgc_peer_alloc, which returns a
//     64-byte CUDA IPC handle next to the pointer; the host exchanges the handles by whatever transport it has (an
//     all-gather of 64 bytes per rank) and maps the neighbours' allocations with sgc_peer_open.  The memory is plain
//     cudaMalloc memory, so the local kernels see exactly what they see on one GPU;
//   * one process driving several GPUs (a Julia session): sgc_comm_init_local enables peer access between the
//     devices, after which every cudaMalloc pointer is valid on every device (unified addressing) and no handle
//     has to travel.
// Memory that was peer-mapped by somebody else (e.g. torch symmetric memory) can be used too:
// sgc_comm_connect_ptrs takes the already-mapped control blocks.
#include "sgc_internal.h"

#include "sgc_kernels.cuh"

using namespace sgc;

static_assert(SGC_IPC_HANDLE_BYTES == sizeof(cudaIpcMemHandle_t), "SGC_IPC_HANDLE_BYTES must match cudaIpcMemHandle_t");

struct sgc_comm {
    int world = 1, rank = 0, device = 0;
    bool local_mode = false;                 // one process, several devices
    bool own_ctrl = false;                   // ctrl was allocated by sgc_comm_init
    long long *ctrl = nullptr;               // this rank's control block (device memory)
    std::vector<long long *> peer_ctrl;      // [world]; peer_ctrl[rank] == ctrl
    std::vector<void *> opened;              // IPC mappings to close
    bool connected = false;
    long long **d_table = nullptr;           // device copy of peer_ctrl (built on the first barrier)
};

namespace {

// One CTA, `world` threads.  Thread j tells rank j "rank `rank` has arrived at barrier #n" (release store into
// rank j's slot for me) and waits until rank j has told me the same.  n comes from device memory, so the launch
// is identical every time (CUDA-graph replay).  Everything earlier in the stream on every rank is visible to
// everything later in the stream on every rank.
__global__ void k_peer_barrier(long long *ctrl, long long *const *peer_ctrl, int rank, int world) {
    const int j = threadIdx.x;
    const long long n = ld_volatile_s64(ctrl + SGC_CTRL_BAR_COUNT) + 1;
    __threadfence_system();
    if (j < world && j != rank) {
        st_release_sys(peer_ctrl[j] + SGC_CTRL_BAR_SLOTS + rank, n);
        wait_flag_sys(ctrl + SGC_CTRL_BAR_SLOTS + j, n, ctrl, 32);
    }
    __syncthreads();
    if (j == 0) {
        st_volatile_s64(ctrl + SGC_CTRL_BAR_COUNT, n);
        __threadfence_system();
    }
}

int ctrl_bytes(int world) { return (int)sizeof(long long) * (SGC_CTRL_BAR_SLOTS + world + 8); }

}  // namespace

extern "C" int sgc_comm_ctrl_bytes(int world) { return ctrl_bytes(world > 0 ? world : 1); }

extern "C" int sgc_comm_init(sgc_comm **comm_out, int world, int rank) {
    REQUIRE(comm_out, "comm is NULL");
    *comm_out = nullptr;
    REQUIRE(world >= 1 && world <= 1024 && rank >= 0 && rank < world, "rank %d outside a world of %d", rank, world);
    sgc_comm *c = new sgc_comm();
    c->world = world; c->rank = rank;
    cudaError_t e = cudaGetDevice(&c->device);
    if (e == cudaSuccess) e = cudaMalloc((void **)&c->ctrl, ctrl_bytes(world));
    if (e == cudaSuccess) e = cudaMemset(c->ctrl, 0, ctrl_bytes(world));
    if (e != cudaSuccess) { delete c; return sgc_fail(SGC_ERR_CUDA, "sgc_comm_init: %s", cudaGetErrorString(e)); }
    c->own_ctrl = true;
    c->peer_ctrl.assign(world, nullptr);
    c->peer_ctrl[rank] = c->ctrl;
    c->connected = (world == 1);
    *comm_out = c;
    return SGC_OK;
}

extern "C" int sgc_comm_handle(sgc_comm *c, void *handle_out) {
    REQUIRE(c && handle_out, "NULL argument");
    REQUIRE(c->own_ctrl, "this context uses an externally mapped control block");
    cudaIpcMemHandle_t h;
    CUDA_TRY(cudaIpcGetMemHandle(&h, c->ctrl));
    memcpy(handle_out, &h, sizeof h);
    return SGC_OK;
}

extern "C" int sgc_comm_connect(sgc_comm *c, const void *handles) {
    REQUIRE(c && handles, "NULL argument");
    REQUIRE(!c->local_mode, "a local (single-process) context is connected at creation");
    for (int r = 0; r < c->world; ++r) {
        if (r == c->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const unsigned char *)handles + (size_t)r * SGC_IPC_HANDLE_BYTES, sizeof h);
        void *p = nullptr;
        CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        c->opened.push_back(p);
        c->peer_ctrl[r] = (long long *)p;
    }
    c->connected = true;
    return SGC_OK;
}

extern "C" int sgc_comm_connect_ptrs(sgc_comm *c, void *my_ctrl, void *const *peer_ctrl) {
    REQUIRE(c && my_ctrl && peer_ctrl, "NULL argument");
    if (c->own_ctrl && c->ctrl) cudaFree(c->ctrl);
    c->own_ctrl = false;
    c->ctrl = (long long *)my_ctrl;
    for (int r = 0; r < c->world; ++r) c->peer_ctrl[r] = (long long *)peer_ctrl[r];
    c->peer_ctrl[c->rank] = c->ctrl;
    CUDA_TRY(cudaMemset(c->ctrl, 0, ctrl_bytes(c->world)));
    c->connected = true;
    return SGC_OK;
}

extern "C" int sgc_comm_init_local(sgc_comm **comms, int ngpus) {
    REQUIRE(comms && ngpus >= 1, "bad arguments");
    int have = 0, prev = 0;
    CUDA_TRY(cudaGetDeviceCount(&have));
    REQUIRE(ngpus <= have, "%d GPUs requested, %d visible", ngpus, have);
    CUDA_TRY(cudaGetDevice(&prev));
    for (int r = 0; r < ngpus; ++r) comms[r] = nullptr;
    for (int r = 0; r < ngpus; ++r) {
        CUDA_TRY(cudaSetDevice(r));
        for (int q = 0; q < ngpus; ++q) {
            if (q == r) continue;
            cudaError_t e = cudaDeviceEnablePeerAccess(q, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); continue; }
            if (e != cudaSuccess) { cudaSetDevice(prev); return sgc_fail(SGC_ERR_CUDA, "no peer access %d -> %d: %s", r, q, cudaGetErrorString(e)); }
        }
        int s = sgc_comm_init(&comms[r], ngpus, r);
        if (s) { cudaSetDevice(prev); return s; }
        comms[r]->local_mode = true;
    }
    for (int r = 0; r < ngpus; ++r) {
        for (int q = 0; q < ngpus; ++q) comms[r]->peer_ctrl[q] = comms[q]->ctrl;
        comms[r]->connected = true;
    }
    CUDA_TRY(cudaSetDevice(prev));
    return SGC_OK;
}

extern "C" int sgc_comm_rank(const sgc_comm *c, int *rank, int *world) {
    REQUIRE(c, "comm is NULL");
    if (rank) *rank = c->rank;
    if (world) *world = c->world;
    return SGC_OK;
}

extern "C" int sgc_peer_alloc(sgc_comm *c, size_t bytes, void **dptr, void *handle_out) {
    REQUIRE(c && dptr, "NULL argument");
    *dptr = nullptr;
    int prev = 0;
    CUDA_TRY(cudaGetDevice(&prev));
    if (prev != c->device) CUDA_TRY(cudaSetDevice(c->device));
    cudaError_t e = cudaMalloc(dptr, bytes ? bytes : 1);
    if (e == cudaSuccess && handle_out) {
        if (c->local_mode) memset(handle_out, 0, SGC_IPC_HANDLE_BYTES);
        else {
            cudaIpcMemHandle_t h;
            e = cudaIpcGetMemHandle(&h, *dptr);
            if (e == cudaSuccess) memcpy(handle_out, &h, sizeof h);
        }
    }
    if (prev != c->device) cudaSetDevice(prev);
    if (e == cudaErrorMemoryAllocation) { cudaGetLastError(); return sgc_fail(SGC_ERR_NOMEM, "sgc_peer_alloc: out of device memory (%zu bytes)", bytes); }
    if (e != cudaSuccess) return sgc_fail(SGC_ERR_CUDA, "sgc_peer_alloc: %s", cudaGetErrorString(e));
    return SGC_OK;
}

extern "C" int sgc_peer_open(sgc_comm *c, const void *handle, void **peer_dptr) {
    REQUIRE(c && handle && peer_dptr, "NULL argument");
    REQUIRE(!c->local_mode, "single-process contexts address peer memory directly; no handle to open");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof h);
    CUDA_TRY(cudaIpcOpenMemHandle(peer_dptr, h, cudaIpcMemLazyEnablePeerAccess));
    c->opened.push_back(*peer_dptr);
    return SGC_OK;
}

extern "C" int sgc_peer_close(sgc_comm *c, void *peer_dptr) {
    REQUIRE(c, "comm is NULL");
    for (size_t q = 0; q < c->opened.size(); ++q)
        if (c->opened[q] == peer_dptr) {
            c->opened.erase(c->opened.begin() + q);
            CUDA_TRY(cudaIpcCloseMemHandle(peer_dptr));
            return SGC_OK;
        }
    return sgc_fail(SGC_ERR_INVALID, "pointer was not opened through this context");
}

extern "C" int sgc_peer_free(sgc_comm *c, void *dptr) {
    REQUIRE(c, "comm is NULL");
    CUDA_TRY(cudaFree(dptr));
    return SGC_OK;
}

extern "C" int sgc_peer_barrier(sgc_comm *c, void *stream) {
    REQUIRE(c, "comm is NULL");
    if (c->world == 1) return SGC_OK;
    REQUIRE(c->connected, "sgc_comm_connect has not been called");
    if (!c->d_table) {
        CUDA_TRY(cudaMalloc((void **)&c->d_table, sizeof(long long *) * c->world));
        CUDA_TRY(cudaMemcpy(c->d_table, c->peer_ctrl.data(), sizeof(long long *) * c->world, cudaMemcpyHostToDevice));
    }
    long long **table = c->d_table;
    const int threads = (c->world + 31) / 32 * 32;
    k_peer_barrier<<<1, threads, 0, (cudaStream_t)stream>>>(c->ctrl, table, c->rank, c->world);
    return sgc_check_launch("k_peer_barrier");
}

extern "C" int sgc_comm_destroy(sgc_comm *c) {
    if (!c) return SGC_OK;
    for (void *p : c->opened) cudaIpcCloseMemHandle(p);
    if (c->own_ctrl && c->ctrl) cudaFree(c->ctrl);
    if (c->d_table) cudaFree(c->d_table);
    delete c;
    return SGC_OK;
}

// Finished synchronised applies of this rank (host read-back; synchronises the stream) — tests and diagnostics.
extern "C" int sgc_comm_epoch(sgc_comm *c, int64_t *epoch, void *stream) {
    REQUIRE(c && epoch, "NULL argument");
    long long v = 0;
    CUDA_TRY(cudaMemcpyAsync(&v, c->ctrl + SGC_CTRL_EPOCH, sizeof v, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    *epoch = v;
    return SGC_OK;
}

// Waits on a neighbour that gave up after SGC_SPIN_LIMIT polls (0 in a correct run; host read-back, synchronises).
extern "C" int sgc_comm_timeouts(sgc_comm *c, int64_t *count, void *stream) {
    REQUIRE(c && count, "NULL argument");
    long long v = 0;
    CUDA_TRY(cudaMemcpyAsync(&v, c->ctrl + SGC_CTRL_TIMEOUTS, sizeof v, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    *count = v;
    return SGC_OK;
}

// ---- operators on a slab: neighbour synchronisation inside the fused curl kernel ---------------------------------
extern "C" int sgc_op_attach_comm(sgc_op *op, sgc_comm *c, int lo_rank, int hi_rank) {
    REQUIRE(op && c, "NULL argument");
    REQUIRE(op->fused_curl, "neighbour synchronisation is built into the fused 3-D curl kernel only");
    REQUIRE(c->connected, "sgc_comm_connect has not been called");
    REQUIRE(lo_rank < c->world && hi_rank < c->world, "neighbour rank outside the world");
    op->sync_ctrl = c->ctrl;
    op->sync_side[0] = lo_rank >= 0 ? 1 : 0;
    op->sync_side[1] = hi_rank >= 0 ? 1 : 0;
    // I am the UPPER neighbour of lo_rank (its FLAG_HI slot is mine to write) and the LOWER neighbour of hi_rank
    op->sync_post[0] = lo_rank >= 0 ? c->peer_ctrl[lo_rank] + SGC_CTRL_FLAG_HI : nullptr;
    op->sync_post[1] = hi_rank >= 0 ? c->peer_ctrl[hi_rank] + SGC_CTRL_FLAG_LO : nullptr;
    return SGC_OK;
}

extern "C" int sgc_op_apply_synced(const sgc_op *op, void *G, const void *F, int opmode, const void *const *ghost,
                                   void *stream) {
    REQUIRE(op, "op is NULL");
    REQUIRE(op->fused_curl && op->sync_ctrl, "call sgc_op_attach_comm on a full 3-D curl operator first");
    op->sync_on = true;
    int s = sgc_op_apply_range(op, G, F, opmode, ghost, 0, op->N[op->ndim - 1], stream);
    op->sync_on = false;
    return s;
}
