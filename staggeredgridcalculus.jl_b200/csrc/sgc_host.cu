// sgc_host.cu — the host-array entry points: the strict drop-in form of the reference's calls, whose arguments
// are Julia `Array`s in host memory (apply_curl! matrixfree/differential/curl.jl:52-63, apply_∂! partial.jl:33-43 …).
//
// The path is bound by the PCIe link, not by the kernels, so everything here is about keeping the link busy in both
// directions at once:
//   * the caller's arrays are page-locked IN PLACE (cudaHostRegister) the first time they are seen and the
//     registration is cached, so that every later call DMA-copies straight out of / into the caller's memory at the
//     pinned rate — no staging copy through an intermediate buffer, no extra pass over host DRAM;
//   * the input streams in z-chunks on one stream, the kernels run on each chunk's planes as soon as the planes
//     their stencil reaches have arrived (sgc_op_apply_range), and the result streams out on a third stream: H2D,
//     compute and D2H overlap, a pass costs about max(H2D, D2H) instead of their sum.
#include "sgc_internal.h"

#include <algorithm>
#include <map>
#include <mutex>

using namespace sgc;

struct sgc_hostpipe {
    int device = -1;
    cudaStream_t s_in = nullptr, s_cmp = nullptr, s_out = nullptr;
    void *buf[3] = {nullptr, nullptr, nullptr};  // input field, output field, intermediate field
    size_t bytes[3] = {0, 0, 0};
    std::vector<cudaEvent_t> ev;
    size_t ev_used = 0;
};

void sgc_hostpipe_destroy(sgc_hostpipe *hp) {
    if (!hp) return;
    for (cudaEvent_t e : hp->ev) cudaEventDestroy(e);
    for (int q = 0; q < 3; ++q) if (hp->buf[q]) cudaFree(hp->buf[q]);
    if (hp->s_in) cudaStreamDestroy(hp->s_in);
    if (hp->s_cmp) cudaStreamDestroy(hp->s_cmp);
    if (hp->s_out) cudaStreamDestroy(hp->s_out);
    delete hp;
}

namespace {

// ---- page-locking of caller memory, cached ---------------------------------------------------------------------
std::mutex g_reg_mutex;
std::map<uintptr_t, size_t> g_registered;  // base → bytes
int g_reg_mode = 1;                        // 0: never (pageable copies), 1: register on first sight and keep

bool is_registered_locked(uintptr_t p, size_t bytes) {
    auto it = g_registered.upper_bound(p);
    if (it == g_registered.begin()) return false;
    --it;
    return it->first <= p && p + bytes <= it->first + it->second;
}

// Pins [ptr, ptr+bytes) unless it already is (by us, or by the caller: cudaMallocHost / pinned tensors).  Failure to
// pin is not an error: the copies then go through the driver's pageable path.
void ensure_pinned(const void *ptr, size_t bytes) {
    if (g_reg_mode == 0 || !ptr || bytes == 0) return;
    std::lock_guard<std::mutex> lock(g_reg_mutex);
    const uintptr_t p = (uintptr_t)ptr;
    if (is_registered_locked(p, bytes)) return;
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, ptr) == cudaSuccess && at.type == cudaMemoryTypeHost) return;  // pinned already
    cudaGetLastError();
    // a stale or overlapping entry (the caller freed and re-allocated): drop what overlaps, then pin afresh
    for (auto it = g_registered.begin(); it != g_registered.end();) {
        if (it->first < p + bytes && p < it->first + it->second) {
            cudaHostUnregister((void *)it->first);
            it = g_registered.erase(it);
        } else ++it;
    }
    cudaGetLastError();
    if (cudaHostRegister((void *)ptr, bytes, cudaHostRegisterDefault) == cudaSuccess) g_registered[p] = bytes;
    else cudaGetLastError();
}

int pipe_of(const sgc_op *op, sgc_hostpipe **out) {
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (op->hp && op->hp->device != dev) { sgc_hostpipe_destroy(op->hp); op->hp = nullptr; }
    if (!op->hp) {
        sgc_hostpipe *hp = new sgc_hostpipe();
        hp->device = dev;
        op->hp = hp;
        CUDA_TRY(cudaStreamCreateWithFlags(&hp->s_in, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&hp->s_cmp, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&hp->s_out, cudaStreamNonBlocking));
    }
    op->hp->ev_used = 0;
    *out = op->hp;
    return SGC_OK;
}

int need_buf(sgc_hostpipe *hp, int q, size_t bytes) {
    if (hp->bytes[q] >= bytes) return SGC_OK;
    if (hp->buf[q]) { cudaFree(hp->buf[q]); hp->buf[q] = nullptr; hp->bytes[q] = 0; }
    cudaError_t e = cudaMalloc(&hp->buf[q], bytes);
    if (e == cudaErrorMemoryAllocation) { cudaGetLastError(); return sgc_fail(SGC_ERR_NOMEM, "out of device memory for the host-array staging buffers (%zu bytes)", bytes); }
    if (e != cudaSuccess) return sgc_fail(SGC_ERR_CUDA, "cudaMalloc failed: %s", cudaGetErrorString(e));
    hp->bytes[q] = bytes;
    return SGC_OK;
}

// `to` waits for everything queued on `from` so far
int chain(sgc_hostpipe *hp, cudaStream_t from, cudaStream_t to) {
    if (hp->ev_used == hp->ev.size()) {
        cudaEvent_t e;
        CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        hp->ev.push_back(e);
    }
    cudaEvent_t e = hp->ev[hp->ev_used++];
    CUDA_TRY(cudaEventRecord(e, from));
    CUDA_TRY(cudaStreamWaitEvent(to, e, 0));
    return SGC_OK;
}

// planes [a, b) of each of `ncomp` components (each range is contiguous in a Julia array)
int copy_planes(void *dst, const void *src, int ncomp, int64_t M, int64_t plane, size_t es, int64_t a, int64_t b,
                cudaMemcpyKind kind, cudaStream_t st) {
    if (b <= a) return SGC_OK;
    const size_t n = (size_t)(b - a) * plane * es;
    for (int w = 0; w < ncomp; ++w) {
        const size_t off = ((size_t)w * M + (size_t)a * plane) * es;
        CUDA_TRY(cudaMemcpyAsync((unsigned char *)dst + off, (const unsigned char *)src + off, n, kind, st));
    }
    return SGC_OK;
}

struct LastAxis { int reach; bool bloch; };  // stencil reach along the last array dimension: +1, −1 or 0
LastAxis last_axis_of(const sgc_op *op) {
    LastAxis r{0, false};
    for (const Term &t : op->terms)
        if (t.kind >= 0 && t.axis == op->ndim - 1) { r.reach = t.fwd ? 1 : -1; r.bloch = t.bloch != 0; }
    return r;
}

int default_chunks(int64_t Nz, size_t bytes) {
    if (bytes < ((size_t)8 << 20)) return 1;  // too small to be worth splitting
    return (int)std::min<int64_t>(Nz, 16);
}

#define TRY(expr) do { int _s = (expr); if (_s) return _s; } while (0)

}  // namespace

extern "C" int sgc_host_set_registration(int mode) {
    REQUIRE(mode == 0 || mode == 1, "mode must be 0 (never page-lock caller memory) or 1 (page-lock on first use, cached)");
    std::lock_guard<std::mutex> lock(g_reg_mutex);
    g_reg_mode = mode;
    return SGC_OK;
}

extern "C" int sgc_host_register(const void *ptr, size_t bytes) {
    REQUIRE(ptr, "ptr is NULL");
    ensure_pinned(ptr, bytes);
    return SGC_OK;
}

extern "C" int sgc_host_unregister(const void *ptr) {
    std::lock_guard<std::mutex> lock(g_reg_mutex);
    if (!ptr) {  // everything
        for (auto &kv : g_registered) cudaHostUnregister((void *)kv.first);
        g_registered.clear();
        cudaGetLastError();
        return SGC_OK;
    }
    auto it = g_registered.find((uintptr_t)ptr);
    if (it != g_registered.end()) {
        cudaHostUnregister((void *)it->first);
        g_registered.erase(it);
        cudaGetLastError();
    }
    return SGC_OK;
}

// One operator on host arrays, z-chunked and full-duplex.  After chunk [a, b) of the input has arrived, the planes
// whose stencil reaches only arrived planes are computed and sent back:
//   no term along the last axis      : [a, b)
//   forward  (needs plane k+1)       : [a−1, b−1), the last chunk up to N (plane N−1 reads plane 0 or nothing)
//   backward (needs plane k−1)       : [a, b); with a Bloch boundary plane 0 reads plane N−1 and is finished last.
extern "C" int sgc_op_apply_host_chunked(const sgc_op *op, void *G_host, const void *F_host, int opmode, int nchunks) {
    REQUIRE(op && G_host && F_host, "NULL argument");
    REQUIRE(opmode == SGC_OP_SET || opmode == SGC_OP_ADD, "opmode must be SGC_OP_SET or SGC_OP_ADD");
    const size_t es = op->dtype == SGC_C128 ? 16 : 8;
    const int64_t M = op->M, Nz = op->N[op->ndim - 1], plane = M / Nz;
    const size_t bF = es * (size_t)M * op->n_in, bG = es * (size_t)M * op->n_out;
    sgc_hostpipe *hp = nullptr;
    TRY(pipe_of(op, &hp));
    TRY(need_buf(hp, 0, bF));
    TRY(need_buf(hp, 1, bG));
    ensure_pinned(F_host, bF);
    ensure_pinned(G_host, bG);
    void *dF = hp->buf[0], *dG = hp->buf[1];
    const LastAxis la = last_axis_of(op);
    if (nchunks <= 0) nchunks = default_chunks(Nz, bF + bG);
    nchunks = (int)std::max<int64_t>(1, std::min<int64_t>(nchunks, Nz));
    if (Nz < 2) nchunks = 1;
    const bool add = opmode == SGC_OP_ADD;
    const bool defer0 = (la.reach < 0 && la.bloch && nchunks > 1);
    for (int c = 0; c < nchunks; ++c) {
        const int64_t a = Nz * c / nchunks, b = Nz * (c + 1) / nchunks;
        TRY(copy_planes(dF, F_host, op->n_in, M, plane, es, a, b, cudaMemcpyHostToDevice, hp->s_in));
        if (add) TRY(copy_planes(dG, G_host, op->n_out, M, plane, es, a, b, cudaMemcpyHostToDevice, hp->s_in));
        TRY(chain(hp, hp->s_in, hp->s_cmp));
        int64_t kb = a, ke = b;
        if (la.reach > 0) { kb = c > 0 ? a - 1 : 0; ke = (c == nchunks - 1) ? Nz : b - 1; }
        else if (defer0 && c == 0) kb = 1;
        if (ke <= kb) continue;
        TRY(sgc_op_apply_range(op, dG, dF, opmode, nullptr, kb, ke, hp->s_cmp));
        TRY(chain(hp, hp->s_cmp, hp->s_out));
        TRY(copy_planes(G_host, dG, op->n_out, M, plane, es, kb, ke, cudaMemcpyDeviceToHost, hp->s_out));
    }
    if (defer0) {
        TRY(sgc_op_apply_range(op, dG, dF, opmode, nullptr, 0, 1, hp->s_cmp));
        TRY(chain(hp, hp->s_cmp, hp->s_out));
        TRY(copy_planes(G_host, dG, op->n_out, M, plane, es, 0, 1, cudaMemcpyDeviceToHost, hp->s_out));
    }
    CUDA_TRY(cudaStreamSynchronize(hp->s_out));
    CUDA_TRY(cudaStreamSynchronize(hp->s_cmp));
    return SGC_OK;
}

extern "C" int sgc_op_apply_host(const sgc_op *op, void *G_host, const void *F_host, int opmode) {
    return sgc_op_apply_host_chunked(op, G_host, F_host, opmode, 0);
}

// G OP C2(C1 F) on host arrays (two apply_curl!, the intermediate field stays on the device).  With a forward-z first
// curl and a backward-z second one — the E → H → E′ pass of a Maxwell operator — the pass is z-chunked: after chunk
// [a, b) of F has arrived, H = C1 F exists on [a−1, b−1) (H[k] reads F[k+1]) and G on the same planes (G[k] reads
// H[k−1]).  With a Bloch z boundary G[0] reads H[N−1] and is finished last.  Other direction pairings run serially.
extern "C" int sgc_op_apply_curlcurl_host(const sgc_op *c2, const sgc_op *c1, void *G_host, const void *F_host, int opmode,
                                          int nchunks) {
    REQUIRE(c1 && c2 && G_host && F_host, "NULL argument");
    REQUIRE(opmode == SGC_OP_SET || opmode == SGC_OP_ADD, "opmode must be SGC_OP_SET or SGC_OP_ADD");
    REQUIRE(c1->fused_curl && c2->fused_curl, "the curl-of-curl needs two full 3-D curl operators");
    REQUIRE(c1->dtype == c2->dtype, "the two curls must have the same field type");
    for (int d = 0; d < 3; ++d) REQUIRE(c1->N[d] == c2->N[d], "the two curls must act on the same grid");
    REQUIRE(!c1->lo_iface && !c1->hi_iface && !c2->lo_iface && !c2->hi_iface, "host-array passes run on whole grids, not slabs");
    const size_t es = c1->dtype == SGC_C128 ? 16 : 8;
    const int64_t M = c1->M, Nz = c1->N[2], plane = M / Nz;
    const size_t bytes = es * (size_t)M * 3;
    sgc_hostpipe *hp = nullptr;
    TRY(pipe_of(c2, &hp));
    for (int q = 0; q < 3; ++q) TRY(need_buf(hp, q, bytes));
    ensure_pinned(F_host, bytes);
    ensure_pinned(G_host, bytes);
    void *dF = hp->buf[0], *dG = hp->buf[1], *dH = hp->buf[2];
    const LastAxis l1 = last_axis_of(c1), l2 = last_axis_of(c2);
    const bool add = opmode == SGC_OP_ADD;
    if (nchunks <= 0) nchunks = default_chunks(Nz, 2 * bytes);
    nchunks = (int)std::max<int64_t>(1, std::min<int64_t>(nchunks, Nz));
    if (!(l1.reach > 0 && l2.reach < 0) || Nz < 2) nchunks = 1;
    if (nchunks == 1) {  // serial: everything in, two curls, everything out
        TRY(copy_planes(dF, F_host, 3, M, plane, es, 0, Nz, cudaMemcpyHostToDevice, hp->s_cmp));
        if (add) TRY(copy_planes(dG, G_host, 3, M, plane, es, 0, Nz, cudaMemcpyHostToDevice, hp->s_cmp));
        TRY(sgc_op_apply(c1, dH, dF, SGC_OP_SET, hp->s_cmp));
        TRY(sgc_op_apply(c2, dG, dH, opmode, hp->s_cmp));
        TRY(copy_planes(G_host, dG, 3, M, plane, es, 0, Nz, cudaMemcpyDeviceToHost, hp->s_cmp));
        CUDA_TRY(cudaStreamSynchronize(hp->s_cmp));
        return SGC_OK;
    }
    const bool defer0 = l2.bloch;
    for (int c = 0; c < nchunks; ++c) {
        const int64_t a = Nz * c / nchunks, b = Nz * (c + 1) / nchunks;
        TRY(copy_planes(dF, F_host, 3, M, plane, es, a, b, cudaMemcpyHostToDevice, hp->s_in));
        if (add) TRY(copy_planes(dG, G_host, 3, M, plane, es, a, b, cudaMemcpyHostToDevice, hp->s_in));
        TRY(chain(hp, hp->s_in, hp->s_cmp));
        const int64_t kb = c > 0 ? a - 1 : 0, ke = (c == nchunks - 1) ? Nz : b - 1;
        if (ke <= kb) continue;
        TRY(sgc_op_apply_range(c1, dH, dF, SGC_OP_SET, nullptr, kb, ke, hp->s_cmp));
        const int64_t kb2 = (defer0 && kb == 0) ? 1 : kb;
        if (ke <= kb2) continue;
        TRY(sgc_op_apply_range(c2, dG, dH, opmode, nullptr, kb2, ke, hp->s_cmp));
        TRY(chain(hp, hp->s_cmp, hp->s_out));
        TRY(copy_planes(G_host, dG, 3, M, plane, es, kb2, ke, cudaMemcpyDeviceToHost, hp->s_out));
    }
    if (defer0) {
        TRY(sgc_op_apply_range(c2, dG, dH, opmode, nullptr, 0, 1, hp->s_cmp));
        TRY(chain(hp, hp->s_cmp, hp->s_out));
        TRY(copy_planes(G_host, dG, 3, M, plane, es, 0, 1, cudaMemcpyDeviceToHost, hp->s_out));
    }
    CUDA_TRY(cudaStreamSynchronize(hp->s_out));
    CUDA_TRY(cudaStreamSynchronize(hp->s_cmp));
    return SGC_OK;
}
