/* sgc_b200.h — C ABI of libsgc_b200.so: the B200-native replacement for the two data-parallel
 * hot paths of wsshin/StaggeredGridCalculus.jl (matrix-free staggered-grid operators and
 * sparse-operator assembly).
 *
 * The reference has no FFI: its boundary is Julia multiple dispatch on the concrete methods
 * cited next to each entry point below (paths relative to the reference's src/).  A Julia shim
 * (staggeredgridcalculus.jl_b200/julia/SGCB200.jl) adds methods with the
 * same signatures on a device-array type and forwards them here through `ccall`; the Python
 * `ctypes` mirror in staggeredgridcalculus.jl_b200/ binds exactly the same symbols.
 *
 * Conventions
 *  - Plain C types only.  Every function returns an `int` status (SGC_OK == 0); the message of
 *    the last failure on the calling thread is returned by sgc_last_error().
 *  - Field arrays (`G`, `F`, `Gv`, `Fu`, `g`, `f`) are DEVICE pointers to column-major Julia
 *    arrays: F[i,j,k,w] with the component index w last (slowest), i.e. K contiguous
 *    component planes of prod(N) elements.  dtype SGC_F64 = Float64, SGC_C128 = ComplexF64
 *    (re,im interleaved).  The `*_host` variants take HOST pointers and do the copies.
 *  - Per-axis coefficient vectors (∆w⁻¹, ∆w, ∆w′⁻¹) are HOST pointers to N[w] elements of
 *    Float64 (`coef_complex == 0`) or ComplexF64 (`coef_complex == 1`); they are O(ΣN) and are
 *    turned into resident device tables when an operator handle is created.
 *  - Scalars e⁻ⁱᵏᴸ and α are passed as (re, im) pairs of doubles.
 *  - Axis numbers `nw`, components `cmp_*` are 1-based, exactly as in the reference.
 *  - `op`: SGC_OP_SET is Val(:(=)), SGC_OP_ADD is Val(:(+=))  (matrixfree/matrixfree.jl:6-8).
 *  - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream).  Calls are
 *    asynchronous with respect to the host unless stated otherwise.
 *  - No CPU fallback exists: without a CUDA device every compute entry point fails with
 *    SGC_ERR_CUDA.
 */
#ifndef SGC_B200_H
#define SGC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SGC_ABI_VERSION 2

/* status codes */
#define SGC_OK 0
#define SGC_ERR_INVALID 1 /* the reference's @assert / @error conditions */
#define SGC_ERR_CUDA 2    /* CUDA runtime failure, or no device */
#define SGC_ERR_INEXACT 3 /* complex α / e / ∆ with a Float64 field (Julia InexactError) */
#define SGC_ERR_NOMEM 4

/* element types */
#define SGC_F64 0
#define SGC_C128 1

/* set_or_add! */
#define SGC_OP_SET 0
#define SGC_OP_ADD 1

/* ---------------------------------------------------------------------------------------- */
/* library / device plumbing                                                                */
/* ---------------------------------------------------------------------------------------- */
int sgc_abi_version(void);
/* bit 0: development build with every tile / ring-depth variant of the tuning sweeps (`make EXTRA=-DSGC_SWEEP_BUILD`);
 * the product library carries the defaults and a few alternatives only. */
int sgc_build_flags(void);
const char *sgc_last_error(void);
int sgc_device_count(int *count);
int sgc_set_device(int device);
int sgc_malloc(void **dptr, size_t bytes);
int sgc_free(void *dptr);
int sgc_malloc_host(void **hptr, size_t bytes); /* pinned */
/* Test hook: out[q] = z[q] / w[q] evaluated on the HOST by the library's own complex-division routine (Julia
 * Base's scaled robust algorithm, the source the kernels compile); n (re, im) pairs. */
int sgc_debug_cdiv(const double *z, const double *w, double *out, int64_t n);
/* Test hook: out[q] = inv(w[q]) by the library's host routine for `e^(-1.0)` (Julia Base's scaled robust_cinv,
 * with the multiply-add NOT fused), used for backward Bloch blocks of the assembly. */
int sgc_debug_cinv(const double *w, double *out, int64_t n);
/* pinned AND write-combined: for buffers the host only writes and the GPU reads (H2D sources) */
int sgc_malloc_host_wc(void **hptr, size_t bytes);
int sgc_free_host(void *hptr);
int sgc_memcpy_h2d(void *dst, const void *src, size_t bytes, void *stream);
int sgc_memcpy_d2h(void *dst, const void *src, size_t bytes, void *stream);
int sgc_memset_zero(void *dptr, size_t bytes, void *stream);
int sgc_stream_create(void **stream);
int sgc_stream_destroy(void *stream);
int sgc_stream_sync(void *stream);
/* Number of kernels this library has launched on the calling process so far (bench.py's
 * `gpu_launches` evidence). */
int64_t sgc_launch_count(void);

/* ---------------------------------------------------------------------------------------- */
/* Matrix-free operators: operator handles                                                  */
/* An sgc_op owns the device-resident per-axis tables (α·∆w⁻¹ etc.) of one operator so that  */
/* repeated application (solver iterations, CUDA-graph capture) does no host work.           */
/* ---------------------------------------------------------------------------------------- */
typedef struct sgc_op sgc_op;

/* apply_∂!  — matrixfree/differential/partial.jl:33 (3-D), :241 (2-D), :388 (1-D); the
 * scalar-∆ wrapper :11-21 is `fill` on the caller's side. */
int sgc_op_create_partial(sgc_op **op, int dtype, int ndim, const int64_t *N, int nw, int isfwd,
                          const void *dwinv, int coef_complex, int isbloch, const double *e,
                          const double *alpha);

/* apply_m̂!  — matrixfree/mean.jl:102/295/432 (arithmetic: dw == dwpinv == NULL) and
 * :493/705/856 (weighted). */
int sgc_op_create_mhat(sgc_op **op, int dtype, int ndim, const int64_t *N, int nw, int isfwd,
                       const void *dw, const void *dwpinv, int coef_complex, int isbloch,
                       const double *e, const double *alpha);

/* apply_curl!  — matrixfree/differential/curl.jl:52-137 (wrappers :12-49).  `dlinv[d]` is the
 * vector of array dimension d (length N[d]); `e` holds ndim (re,im) pairs; cmp_shp has ndim
 * entries, cmp_out Kout, cmp_in Kin.  The full 3-D curl (cmp_shp = cmp_out = cmp_in = 1:3)
 * runs as ONE fused kernel; sub-curls run as per-block sweeps. */
int sgc_op_create_curl(sgc_op **op, int dtype, int ndim, const int64_t *N, const int *isfwd,
                       const void *const *dlinv, int coef_complex, const int *isbloch,
                       const double *e, const int *cmp_shp, int Kout, const int *cmp_out, int Kin,
                       const int *cmp_in, const double *alpha);

/* apply_divg!  — matrixfree/differential/divergence.jl:41-66.  F has ndim components. */
int sgc_op_create_divg(sgc_op **op, int dtype, int ndim, const int64_t *N, const int *isfwd,
                       const void *const *dlinv, int coef_complex, const int *isbloch,
                       const double *e, const double *alpha);

/* apply_grad!  — matrixfree/differential/gradient.jl:41-59.  G has ndim components. */
int sgc_op_create_grad(sgc_op **op, int dtype, int ndim, const int64_t *N, const int *isfwd,
                       const void *const *dlinv, int coef_complex, const int *isbloch,
                       const double *e, const double *alpha);

/* apply_mean!  — matrixfree/mean.jl:47-65 (arithmetic: dl == dlpinv == NULL), :69-89. */
int sgc_op_create_mean(sgc_op **op, int dtype, int ndim, const int64_t *N, const int *isfwd,
                       const void *const *dl, const void *const *dlpinv, int coef_complex,
                       const int *isbloch, const double *e, const double *alpha);

/* z-slab mode (multi-GPU): N given at creation is the LOCAL slab; the last
 * array dimension is the decomposed one.  `lo_interface` / `hi_interface` != 0 mark the slab
 * face as an interior interface (the neighbour plane comes from another rank, no boundary
 * condition is applied there) instead of a physical boundary. */
int sgc_op_set_slab(sgc_op *op, int lo_interface, int hi_interface);
/* Weighted means along the slab axis: dw of the neighbour's plane just below / above the slab (one element
 * each, re[,im]; the global wrap-around entries at Bloch faces), which the boundary formulas of
 * matrixfree/mean.jl:516-527 / :633-645 read instead of the local table's first / last entry. */
int sgc_op_set_slab_weights(sgc_op *op, const double *dw_below, const double *dw_above, int coef_complex);

/* Kernel tuning knobs of the fused curl (0 keeps the default): tile size, z-march length, and
 * `variant`: low 4 bits select the kernel (0 = cp.async shared-memory ring [default], 1 =
 * register-prefetch z-march, 2 = unfused per-block sweeps, the reference's own structure;
 * for single-block operators 3 = force the scalar k_term kernel instead of the vectorised ones);
 * bits 4-7 give the number of ring stages for kernel 0 (0 = default 4), bits 8-11 the
 * CTAs-per-SM register target it was compiled for (0 = default).  apply_mean! on a 3-D grid: 0 = one launch of
 * the staged z-march, 2 / 3 = the three sweeps.
 * On the SECOND operator of sgc_op_apply_curlcurl the low 4 bits select the fused curl-of-curl kernel: 0 = default
 * (k_cc6 — TMA-fed, two z-planes per barrier interval — where it applies: ComplexF64, all-symmetric box, forward then
 * backward in z, no slab ghosts; otherwise the third generation), 2 = third generation (warp roles, cp.async ring),
 * 1 = first generation, 4 / 5 = the TMA experiments k_cc4 / k_cc5, 6 = k_cc6 or, where it does not apply, the third
 * generation; bits 4-7 = ring stages (k_cc6: pairs of planes). */
int sgc_op_set_tuning(sgc_op *op, int tile_x, int tile_y, int march_z, int variant);

/* G OP α·(operator)(F).  Returns immediately; work is queued on `stream`. */
int sgc_op_apply(const sgc_op *op, void *G, const void *F, int opmode, void *stream);

/* Same, restricted to the planes [k_begin, k_end) of the last array dimension, with the
 * ghost plane(s) of the input taken from `ghost[c]` (one DEVICE pointer per input component
 * c, each to prod(N[0..ndim-2]) elements; NULL where the operator needs none or where the
 * plane lives in F itself).  For a physical Bloch face the ghost plane holds the RAW values
 * of the wrap-around plane; the kernel applies e⁻ⁱᵏᴸ (forward) or divides by it (backward)
 * exactly as partial.jl:140 / :226 do.  `ghost` pointers may be peer (NVLink) addresses. */
int sgc_op_apply_range(const sgc_op *op, void *G, const void *F, int opmode,
                       const void *const *ghost, int64_t k_begin, int64_t k_end, void *stream);

int sgc_op_destroy(sgc_op *op);

/* Number of kernel launches one sgc_op_apply of this operator issues. */
int sgc_op_launches(const sgc_op *op, int opmode);

/* ---------------------------------------------------------------------------------------- */
/* Matrix-free operators: one-shot calls with the reference's argument lists.  The handle     */
/* behind a call is cached (keyed by every parameter, coefficient vectors by content; 32      */
/* entries, least recently used evicted), so a repeated call only queues its kernel and       */
/* returns without synchronising.                                                             */
/* ---------------------------------------------------------------------------------------- */
int sgc_oneshot_cache_clear(void);
int sgc_apply_partial(void *Gv, const void *Fu, int opmode, int dtype, int ndim, const int64_t *N,
                      int nw, int isfwd, const void *dwinv, int coef_complex, int isbloch,
                      const double *e, const double *alpha, void *stream);
int sgc_apply_mhat(void *Gv, const void *Fu, int opmode, int dtype, int ndim, const int64_t *N,
                   int nw, int isfwd, const void *dw, const void *dwpinv, int coef_complex,
                   int isbloch, const double *e, const double *alpha, void *stream);
int sgc_apply_curl(void *G, const void *F, int opmode, int dtype, int ndim, const int64_t *N,
                   const int *isfwd, const void *const *dlinv, int coef_complex,
                   const int *isbloch, const double *e, const int *cmp_shp, int Kout,
                   const int *cmp_out, int Kin, const int *cmp_in, const double *alpha,
                   void *stream);
int sgc_apply_divg(void *g, const void *F, int opmode, int dtype, int ndim, const int64_t *N,
                   const int *isfwd, const void *const *dlinv, int coef_complex,
                   const int *isbloch, const double *e, const double *alpha, void *stream);
int sgc_apply_grad(void *G, const void *f, int opmode, int dtype, int ndim, const int64_t *N,
                   const int *isfwd, const void *const *dlinv, int coef_complex,
                   const int *isbloch, const double *e, const double *alpha, void *stream);
int sgc_apply_mean(void *G, const void *F, int opmode, int dtype, int ndim, const int64_t *N,
                   const int *isfwd, const void *const *dl, const void *const *dlpinv,
                   int coef_complex, const int *isbloch, const double *e, const double *alpha,
                   void *stream);

/* ---------------------------------------------------------------------------------------- */
/* Multi-GPU context: z-slabs of one grid on the GPUs of one node.       */
/* apply_curl! (matrixfree/differential/curl.jl:52-63) and apply_∂! (partial.jl:33-43) on a  */
/* slab need ONE plane from ONE z-neighbour; the fused kernel loads it straight from that     */
/* neighbour's memory over NVLink (`ghost` pointers of sgc_op_apply_range), so all the        */
/* library has to provide is peer-mapped memory and ordering between the ranks' kernels.      */
/* No torch, no NCCL: plain CUDA IPC (one process per GPU) or peer access (one process).      */
/* ---------------------------------------------------------------------------------------- */
typedef struct sgc_comm sgc_comm;
#define SGC_IPC_HANDLE_BYTES 64

/* One process per GPU: create the context of `rank` on the CURRENT device, fetch the handle of its control
 * block, exchange the handles between the processes (any transport: 64 bytes per rank, rank-major), connect. */
int sgc_comm_init(sgc_comm **comm, int world, int rank);
int sgc_comm_handle(sgc_comm *comm, void *handle_out /* SGC_IPC_HANDLE_BYTES */);
int sgc_comm_connect(sgc_comm *comm, const void *handles /* world x SGC_IPC_HANDLE_BYTES */);
/* Same, for control blocks somebody else has peer-mapped already (e.g. symmetric-memory allocators):
 * my_ctrl / peer_ctrl[r] point to sgc_comm_ctrl_bytes(world) bytes each; my_ctrl is zeroed. */
int sgc_comm_connect_ptrs(sgc_comm *comm, void *my_ctrl, void *const *peer_ctrl);
int sgc_comm_ctrl_bytes(int world);
/* One process driving `ngpus` devices (0..ngpus-1): enables peer access between all pairs and returns one
 * connected context per device in comms[0..ngpus-1]; afterwards every sgc_peer_alloc / sgc_malloc pointer is
 * valid on every device and sgc_peer_open is not needed. */
int sgc_comm_init_local(sgc_comm **comms, int ngpus);
int sgc_comm_rank(const sgc_comm *comm, int *rank, int *world);
int sgc_comm_destroy(sgc_comm *comm);

/* Peer-mappable device memory on this context's device (cudaMalloc), with its IPC handle (handle_out may be
 * NULL).  sgc_peer_open maps a neighbour's allocation into this process and returns the address this rank's
 * kernels use for it (the base of the neighbour's allocation); sgc_peer_close unmaps it. */
int sgc_peer_alloc(sgc_comm *comm, size_t bytes, void **dptr, void *handle_out /* SGC_IPC_HANDLE_BYTES */);
int sgc_peer_open(sgc_comm *comm, const void *handle, void **peer_dptr);
int sgc_peer_close(sgc_comm *comm, void *peer_dptr);
int sgc_peer_free(sgc_comm *comm, void *dptr);

/* Device-side barrier of all ranks, queued on `stream` (one small kernel: every rank writes its arrival into
 * every peer's control block and waits for theirs).  Everything queued before it on any rank is visible to
 * everything queued after it on every rank.  All ranks must call it the same number of times.  Graph-capturable. */
int sgc_peer_barrier(sgc_comm *comm, void *stream);
/* Synchronised applies finished by this rank so far (reads the control block; synchronises `stream`). */
int sgc_comm_epoch(sgc_comm *comm, int64_t *epoch, void *stream);
/* Device-side waits on a neighbour (barrier kernel, synchronised applies) are bounded: a waiter that polled for
 * seconds without seeing its neighbour counts a timeout in the control block and goes on, so that ranks issuing
 * different numbers of synchronised calls (or a rank that died) cannot hang the GPU.  0 in a correct run; anything
 * else means the results since the last check are not valid.  Reads the control block; synchronises `stream`. */
int sgc_comm_timeouts(sgc_comm *comm, int64_t *count, void *stream);

/* Neighbour synchronisation folded into the fused 3-D curl kernel — no barrier kernel between applies.
 * sgc_op_attach_comm names the ranks that own the slab below / above this one (-1: none; for a Bloch z
 * boundary the ring wraps).  sgc_op_apply_synced then runs a whole-slab apply in which only the CTAs that
 * touch the first (last) plane take part: they are scheduled LAST, wait until the lower (upper) neighbour has
 * finished the boundary CTAs of its previous synchronised apply, and per side the last one to finish
 * publishes this rank's new count into that neighbour's control block over NVLink.  Interior CTAs neither
 * wait nor count.  The count lives in device memory, so the call sequence can be captured in a CUDA graph
 * and replayed.  Contract: all ranks issue the same sequence of synchronised applies on peer-mapped fields,
 * nothing else modifies those fields in between, and the inputs of the first apply are complete on all
 * ranks (one sgc_peer_barrier).  After anything else has touched the fields: sgc_peer_barrier again. */
int sgc_op_attach_comm(sgc_op *op, sgc_comm *comm, int lo_rank, int hi_rank);
int sgc_op_apply_synced(const sgc_op *op, void *G, const void *F, int opmode, const void *const *ghost,
                        void *stream);

/* Fused curl-of-curl  G OP C2(C1 F)  in one pass over memory: the intermediate field
 * H = C1 F stays in shared memory (96 instead of 192 bytes per ComplexF64 cell).  Bit-identical
 * to sgc_op_apply(c1, H, F, SET) followed by sgc_op_apply(c2, G, H, opmode); replaces the pair
 * of apply_curl! calls either side of H in a Maxwell solve (matrixfree/differential/curl.jl:52-137
 * twice; matrix form `Cv*Cu`, test/curl.jl:123-153).  c1, c2: full 3-D curl operators on the same
 * grid, field type and coefficient type.  Tuning is taken from c2 (tile_x, tile_y, march_z,
 * variant bits 4-7 = E-ring stages). */
int sgc_op_apply_curlcurl(const sgc_op *c2, const sgc_op *c1, void *G, const void *F, int opmode,
                          void *stream);
/* z-slab form: planes [k_begin,k_end); at an interface (sgc_op_set_slab on BOTH operators)
 * ghost_lo[c] / ghost_hi[c] are DEVICE (or peer) pointers to plane -1 / plane Nz of input
 * component c (all three components, whole planes), and sgc_op_set_slab_coef(c1, ...) must have
 * supplied the neighbours' dz^-1 entries.  Needs opposite z-directions in c1 and c2. */
int sgc_op_apply_curlcurl_range(const sgc_op *c2, const sgc_op *c1, void *G, const void *F, int opmode,
                                const void *const *ghost_lo, const void *const *ghost_hi,
                                int64_t k_begin, int64_t k_end, void *stream);
/* dz^-1 of the plane below the slab (global plane k0-1) and above it (global plane k1), one
 * element each (re[,im]); NULL = 1.0.  Stored pre-multiplied by alpha like the tables. */
int sgc_op_set_slab_coef(sgc_op *op, const double *dzinv_below, const double *dzinv_above, int coef_complex);

/* Host-array entry points (the strict drop-in form: Julia `Array`s in host memory; PCIe-bound).  `G_host` /
 * `F_host` hold n_out / n_in components of prod(N) elements each.  The caller's arrays are page-locked in place the
 * first time they are seen (cudaHostRegister, cached — see sgc_host_set_registration), the input streams in
 * z-chunks, the kernels run on each chunk's planes as soon as the planes their stencil reaches have arrived, and
 * the result streams out on a third stream, so that H2D, compute and D2H overlap.  The call returns when G_host
 * is complete.  sgc_op_apply_host = apply_∂! / apply_m̂! / apply_curl! / apply_divg! / apply_grad! / apply_mean!
 * with host arrays (matrixfree/differential/partial.jl:33-43, curl.jl:52-63, …); nchunks <= 0 picks a default. */
int sgc_op_apply_host(const sgc_op *op, void *G_host, const void *F_host, int opmode);
int sgc_op_apply_host_chunked(const sgc_op *op, void *G_host, const void *F_host, int opmode, int nchunks);
/* G OP C2(C1 F) on host arrays: the E -> H -> E' pass of two apply_curl! (curl.jl:52-137 twice), the intermediate
 * field stays on the device.  Forward-z C1 with backward-z C2 is z-chunked and full-duplex (Bloch z: plane 0 of
 * the result is finished last); other pairings run serially. */
int sgc_op_apply_curlcurl_host(const sgc_op *c2, const sgc_op *c1, void *G_host, const void *F_host, int opmode,
                               int nchunks);
/* Page-locking of caller memory.  mode 1 (default): an array is registered the first time a *_host entry point
 * sees it and stays registered (later calls copy at the pinned rate); the caller must call sgc_host_unregister
 * before freeing such an array (the Julia shim does so in a finalizer).  mode 0: never register — pageable copies.
 * sgc_host_register pins explicitly; sgc_host_unregister(NULL) drops every registration. */
int sgc_host_set_registration(int mode);
int sgc_host_register(const void *ptr, size_t bytes);
int sgc_host_unregister(const void *ptr);

/* ---------------------------------------------------------------------------------------- */
/* Device-side producers of the per-axis vectors: PML-stretched spacings      */
/* s∆l = s(l)·∆l of create_stretched_∆l / create_sfactor / calc_sfactor (src/pml.jl:94-153) and   */
/* their inverses (invert_∆l, src/util.jl:42-48), so that a frequency sweep never round-trips     */
/* through the host.  One sgc_axis = one axis of a Grid (grid.jl:125-207), primal or dual set.    */
/* ---------------------------------------------------------------------------------------- */
typedef struct sgc_axis sgc_axis;
/* l[i]: locations of the n points (HOST), dl[i]: their spacings (HOST); lpml_neg / lpml_pos: PML interface
 * locations, Lpml_neg / Lpml_pos: PML thicknesses (pml_loc, pml.jl:165-174); pml: {m, R, κmax, amax, ma} or NULL for
 * PMLParam() = {4, exp(-16), 1, 0, 4}.  The ω-independent profiles σ(d), κ(d), a(d) are tabulated once (host
 * log / pow) and kept on the device. */
int sgc_axis_create(sgc_axis **ax, int64_t n, const double *l, const double *dl, double lpml_neg, double lpml_pos,
                    double Lpml_neg, double Lpml_pos, const double *pml);
int sgc_axis_destroy(sgc_axis *ax);
/* s∆l and / or (s∆l)⁻¹ at angular frequency ω into DEVICE arrays of n ComplexF64 (either may be NULL). */
int sgc_axis_stretch(const sgc_axis *ax, double omega, void *sdl_device, void *sdl_inv_device, void *stream);
/* Rebuild the coefficient tables of an operator handle for a new ω on the device: for every ∂ term along array
 * dimension d, c[i] = (parity·α)·(s∆l_d[i])⁻¹ from axes[d].  For ∂ / curl / divergence / gradient operators created
 * with ComplexF64 coefficient vectors of the same lengths; queued on `stream`, no host round-trip. */
int sgc_op_set_axes(sgc_op *op, const sgc_axis *const *axes, double omega, void *stream);

/* ---------------------------------------------------------------------------------------- */
/* Sparse assembly: create_* → SparseMatrixCSC{T,Int64}                                      */
/* Two steps, so that the caller owns the output arrays (Julia Vectors or device buffers):   */
/*   sgc_asm_create_*  describe the operator (host work O(ΣN));                               */
/*   sgc_asm_colptr    count + prefix-scan on the device → colptr and nnz;                    */
/*   sgc_asm_fill      write rowval / nzval (rows ascending in each column, duplicates        */
/*                     summed, stored zeros dropped: `dropzeros!(sparse(I,J,V,m,n))`).        */
/* ---------------------------------------------------------------------------------------- */
typedef struct sgc_asm sgc_asm;

/* create_∂  — matrix/differential/partial.jl:45-60, create_∂info :87-314 */
int sgc_asm_create_partial(sgc_asm **a, int ndim, const int64_t *N, int nw, int isfwd,
                           const void *dwinv, int coef_complex, int isbloch, const double *e,
                           int e_complex);
/* create_m̂  — matrix/mean.jl:124-138, create_minfo :141-377 (dw == dwpinv == NULL: ones) */
int sgc_asm_create_mhat(sgc_asm **a, int ndim, const int64_t *N, int nw, int isfwd, const void *dw,
                        const void *dwpinv, int coef_complex, int isbloch, const double *e,
                        int e_complex);
/* create_curl  — matrix/differential/curl.jl:50-104 */
int sgc_asm_create_curl(sgc_asm **a, int ndim, const int64_t *N, const int *isfwd,
                        const void *const *dlinv, int coef_complex, const int *isbloch,
                        const double *e, int e_complex, const int *cmp_shp, int Kout,
                        const int *cmp_out, int Kin, const int *cmp_in, int order_cmpfirst);
/* create_divg / create_grad  — matrix/differential/divergence.jl:37-70, gradient.jl:37-70 */
int sgc_asm_create_divg(sgc_asm **a, int ndim, const int64_t *N, const int *isfwd,
                        const void *const *dlinv, int coef_complex, const int *isbloch,
                        const double *e, int e_complex, int order_cmpfirst);
int sgc_asm_create_grad(sgc_asm **a, int ndim, const int64_t *N, const int *isfwd,
                        const void *const *dlinv, int coef_complex, const int *isbloch,
                        const double *e, int e_complex, int order_cmpfirst);
/* create_mean  — matrix/mean.jl:72-111 */
int sgc_asm_create_mean(sgc_asm **a, int ndim, const int64_t *N, const int *isfwd,
                        const void *const *dl, const void *const *dlpinv, int coef_complex,
                        const int *isbloch, const double *e, int e_complex, int order_cmpfirst);
/* create_πcmp  — matrix/permutation.jl:15-40.  `scale` holds Kf Float64 or ComplexF64. */
int sgc_asm_create_picmp(sgc_asm **a, int ndim, const int64_t *N, int Kf, const int *permute,
                         const void *scale, int scale_complex, int order_cmpfirst);

/* Row partition for multi-GPU assembly: keep only the global rows
 * [row_begin, row_end) (0-based, half-open) and emit them re-based to the block.  The vertical
 * concatenation of the blocks of a partition of 0..m equals the global matrix. */
int sgc_asm_set_row_range(sgc_asm *a, int64_t row_begin, int64_t row_end);

/* Column partition (the scalable multi-GPU form): assemble only the columns [col_begin, col_end) (0-based,
 * half-open) as a stand-alone m x (col_end - col_begin) CSC block with GLOBAL row indices — the horizontal
 * concatenation of the blocks of a partition of 0..n is the global matrix, i.e. the global rowval / nzval are the
 * plain concatenation of the blocks' arrays (matrix/differential/curl.jl:103) and the global colptr is each
 * block's colptr shifted by the entries before it.  Every rank does exactly its share of the single-GPU work:
 * the closed-form offsets hold for any column range, so there is no count pass, scan or read-back.  Range ends
 * must be multiples of 128 * (order_cmpfirst ? Kin : 1) columns (or n); sgc_asm_col_partition returns the range of
 * part `part` of `nparts` balanced, aligned parts.  Cannot be combined with a row range. */
int sgc_asm_col_partition(const sgc_asm *a, int part, int nparts, int64_t *col_begin, int64_t *col_end);
int sgc_asm_set_col_range(sgc_asm *a, int64_t col_begin, int64_t col_end);

/* m, n of the (block of the) matrix and its value type (SGC_F64 / SGC_C128 =
 * promote_type of the inputs, curl.jl:59). */
int sgc_asm_shape(const sgc_asm *a, int64_t *m, int64_t *n, int *val_dtype);

/* Fast path (one pass over the output): sgc_asm_count returns nnz so that the caller can
 * allocate; sgc_asm_fill_all then writes colptr, rowval and nzval (all DEVICE pointers) in a
 * single kernel — every output byte is written once and nothing of size O(nnz) is read.
 * When all rows are kept, the number of stored entries per column is separable over the grid
 * axes, so nnz and every CTA's output offset are closed forms of O(ΣN) prefix tables: no count
 * kernel, no scan, no synchronisation.  With a row partition sgc_asm_count runs a count pass
 * (per-CTA totals + device scan) and synchronises the stream. */
int sgc_asm_count(sgc_asm *a, int64_t *nnz, void *stream);
int sgc_asm_fill_all(sgc_asm *a, int64_t *colptr, int index_base, int64_t *rowval, void *nzval,
                     void *stream);
/* Test switches (the tests run every path and require identical output).  force_count_pass bit 0: use the count
 * kernel + device scan even where the closed form applies; bit 1: interior cells take the general entry path
 * (block walk + sort) instead of the host-sorted descriptors; bit 2: the CTA-granular fill kernel (block scan) instead
 * of the warp-granular one where the closed form applies; bit 3: one column per thread where the warp-granular kernel
 * would give a thread several consecutive cells. */
int sgc_asm_set_tuning(sgc_asm *a, int force_count_pass);

/* Two-step form with a separate colptr pass (kept for callers that want the structure first).
 * colptr: DEVICE pointer to n+1 Int64.  index_base 1 gives Julia's 1-based arrays, 0 gives
 * C/scipy.  *nnz (host) is written after the scan has completed (this call synchronises). */
int sgc_asm_colptr(sgc_asm *a, int64_t *colptr, int index_base, int64_t *nnz, void *stream);

/* rowval: DEVICE pointer to nnz Int64; nzval: DEVICE pointer to nnz values of val_dtype. */
int sgc_asm_fill(const sgc_asm *a, const int64_t *colptr, int index_base, int64_t *rowval,
                 void *nzval, void *stream);

/* Generic cross-check path (north-star wording): emit the reference's COO triplets per block
 * on the device, radix-sort by (column,row), sum duplicates, drop zeros.  Same outputs as the
 * two calls above; slower; used by the tests to validate the closed-form kernels. */
int sgc_asm_coo_sort(sgc_asm *a, int64_t *colptr, int index_base, int64_t *rowval, void *nzval,
                     int64_t nnz_capacity, int64_t *nnz, void *stream);

/* Host-array convenience: colptr/rowval/nzval are HOST pointers (call sgc_asm_nnz_host first
 * to size rowval / nzval). */
int sgc_asm_nnz_host(sgc_asm *a, int64_t *nnz);
int sgc_asm_fill_host(sgc_asm *a, int64_t *colptr, int index_base, int64_t *rowval, void *nzval);

int sgc_asm_destroy(sgc_asm *a);

/* y = A·x for an assembled CSC matrix on the device: `mul!(g, Curl, f)` (test/curl.jl:55, test/partial.jl:67-70).
 * Plan form (the tuned one): sgc_spmv_create sorts the entries by row once (stable radix sort: within a row the
 * columns stay ascending, the order of Julia's column sweep) and keeps, per row, the positions of its entries in the
 * CSC arrays and their columns; sgc_spmv_apply is then a row-parallel gather without atomics, which may be repeated
 * with new nzval / x.  colptr / rowval are only read during creation.  Dimensions and nnz must fit 32 bits. */
typedef struct sgc_spmv sgc_spmv;
int sgc_spmv_create(sgc_spmv **plan, int64_t m, int64_t n, const int64_t *colptr, const int64_t *rowval, int index_base,
                    void *stream);
int sgc_spmv_apply(const sgc_spmv *plan, const void *nzval, int val_dtype, const void *x, void *y, void *stream);
int sgc_spmv_destroy(sgc_spmv *plan);
/* One-shot column-parallel form (one thread per column, atomics on y): no plan, slower. */
int sgc_csc_matvec(int64_t m, int64_t n, const int64_t *colptr, const int64_t *rowval,
                   const void *nzval, int index_base, int val_dtype, const void *x, void *y,
                   void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SGC_B200_H */
