"""Host-side mirror of the reference's matrix-free API (src/matrixfree), over the C ABI.

Function names, argument order, defaults and error behaviour follow the Julia methods cited in
each docstring; `op` is `"="` / `"+="` for `Val(:(=))` / `Val(:(+=))`, axis and component
numbers are 1-based.  Arrays are `DeviceArray`s (HBM).  Every function mutates its first
argument and returns None, like the reference.  `Operator` keeps the device-resident tables
of one operator alive for repeated application (solver loops, benchmarks, slab mode).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import lib, check
from .device import (DeviceArray, coef_vector, coef_vectors, ints, int64s, opcode, scalar2, scalars2,
                     stream_handle, require_cuda)

__all__ = ["Operator", "apply_partial", "apply_mhat", "apply_curl", "apply_divg", "apply_grad", "apply_mean"]


def _together(*vs):
    """Promote several optional coefficient vectors to one common element type."""
    cplx = any(v is not None and np.iscomplexobj(v) for v in vs)
    return [None if v is None else np.asarray(v, dtype=np.complex128 if cplx else np.float64) for v in vs]


class Operator:
    """A matrix-free operator with its coefficient tables resident on the device (`sgc_op`)."""

    def __init__(self, handle, dtype_code, N, n_out, n_in, keep):
        self._h = handle
        self.dtype_code = dtype_code
        self.N = tuple(N)
        self.n_out, self.n_in = n_out, n_in
        self._keep = keep

    # ---- constructors (one per reference operator) -----------------------------------------
    @classmethod
    def partial(cls, dtype_code, N, nw, isfwd, dwinv=1.0, isbloch=True, e=1.0, alpha=1.0):
        require_cuda()
        N = tuple(int(n) for n in N)
        if not 1 <= nw <= len(N):
            raise _lib.SgcError(_lib.SGC_ERR_INVALID, f"nw = {nw} out of range")
        p, keep, cc = coef_vector(dwinv, N[nw - 1])
        h = C.c_void_p()
        check(lib.sgc_op_create_partial(C.byref(h), dtype_code, len(N), int64s(N), nw, int(isfwd), p, cc,
                                        int(isbloch), scalar2(e), scalar2(alpha)))
        return cls(h, dtype_code, N, 1, 1, keep)

    @classmethod
    def mhat(cls, dtype_code, N, nw, isfwd, dw=None, dwpinv=None, isbloch=True, e=1.0, alpha=1.0):
        require_cuda()
        N = tuple(int(n) for n in N)
        if not 1 <= nw <= len(N):
            raise _lib.SgcError(_lib.SGC_ERR_INVALID, f"nw = {nw} out of range")
        if dw is not None:
            dw = np.full(N[nw - 1], dw) if np.ndim(dw) == 0 else dw
            dwpinv = np.full(N[nw - 1], dwpinv) if np.ndim(dwpinv) == 0 else dwpinv
        dw, dwpinv = _together(dw, dwpinv)
        p1, k1, c1 = coef_vector(dw, N[nw - 1])
        p2, k2, c2 = coef_vector(dwpinv, N[nw - 1])
        h = C.c_void_p()
        check(lib.sgc_op_create_mhat(C.byref(h), dtype_code, len(N), int64s(N), nw, int(isfwd), p1, p2, c1 or c2,
                                     int(isbloch), scalar2(e), scalar2(alpha)))
        o = cls(h, dtype_code, N, 1, 1, (k1, k2))
        o.weighted_axis = nw if dw is not None else None  # (the slab driver refuses weighted means along the slab axis)
        return o

    @classmethod
    def curl(cls, dtype_code, N, isfwd, dlinv=None, isbloch=None, e=None, cmp_shp=None, cmp_out=None,
             cmp_in=None, alpha=1.0):
        require_cuda()
        N = tuple(int(n) for n in N)
        K = len(N)
        dlinv = [1.0] * K if dlinv is None else dlinv
        isbloch = [True] * K if isbloch is None else isbloch
        e = [1.0] * K if e is None else e
        cmp_shp = list(range(1, K + 1)) if cmp_shp is None else list(cmp_shp)
        cmp_out = list(range(1, K + 1)) if cmp_out is None else list(cmp_out)
        cmp_in = list(range(1, K + 1)) if cmp_in is None else list(cmp_in)
        arr, keep, cc = coef_vectors(dlinv, N)
        h = C.c_void_p()
        check(lib.sgc_op_create_curl(C.byref(h), dtype_code, K, int64s(N), ints(isfwd), arr, cc, ints(isbloch),
                                     scalars2(e), ints(cmp_shp), len(cmp_out), ints(cmp_out), len(cmp_in),
                                     ints(cmp_in), scalar2(alpha)))
        o = cls(h, dtype_code, N, len(cmp_out), len(cmp_in), keep)
        o.isfwd, o.isbloch = [bool(b) for b in isfwd], [bool(b) for b in isbloch]  # (host pipelines check their plan against these)
        return o

    @classmethod
    def divg(cls, dtype_code, N, isfwd, dlinv=None, isbloch=None, e=None, alpha=1.0):
        return cls._divgrad(lib.sgc_op_create_divg, dtype_code, N, isfwd, dlinv, isbloch, e, alpha, 1, len(N))

    @classmethod
    def grad(cls, dtype_code, N, isfwd, dlinv=None, isbloch=None, e=None, alpha=1.0):
        return cls._divgrad(lib.sgc_op_create_grad, dtype_code, N, isfwd, dlinv, isbloch, e, alpha, len(N), 1)

    @classmethod
    def _divgrad(cls, fn, dtype_code, N, isfwd, dlinv, isbloch, e, alpha, n_out, n_in):
        require_cuda()
        N = tuple(int(n) for n in N)
        K = len(N)
        dlinv = [1.0] * K if dlinv is None else dlinv
        isbloch = [True] * K if isbloch is None else isbloch
        e = [1.0] * K if e is None else e
        arr, keep, cc = coef_vectors(dlinv, N)
        h = C.c_void_p()
        check(fn(C.byref(h), dtype_code, K, int64s(N), ints(isfwd), arr, cc, ints(isbloch), scalars2(e),
                 scalar2(alpha)))
        return cls(h, dtype_code, N, n_out, n_in, keep)

    @classmethod
    def mean(cls, dtype_code, N, isfwd, dl=None, dlpinv=None, isbloch=None, e=None, alpha=1.0):
        require_cuda()
        N = tuple(int(n) for n in N)
        K = len(N)
        isbloch = [True] * K if isbloch is None else isbloch
        e = [1.0] * K if e is None else e
        if dl is not None:
            both = list(dl) + list(dlpinv)
            arr, keep, cc = coef_vectors(both, N + N)
            a1 = (C.c_void_p * K)(*[arr[i] for i in range(K)])
            a2 = (C.c_void_p * K)(*[arr[K + i] for i in range(K)])
        else:
            a1 = a2 = None
            keep, cc = [], 0
        h = C.c_void_p()
        check(lib.sgc_op_create_mean(C.byref(h), dtype_code, K, int64s(N), ints(isfwd), a1, a2, cc, ints(isbloch),
                                     scalars2(e), scalar2(alpha)))
        return cls(h, dtype_code, N, K, K, keep)

    # ---- use ------------------------------------------------------------------------------------
    def _check_arrays(self, G: DeviceArray, F: DeviceArray):
        M = int(np.prod(self.N, dtype=np.int64))
        if G.dtype_code != self.dtype_code or F.dtype_code != self.dtype_code:
            raise _lib.SgcError(_lib.SGC_ERR_INVALID, "array element type differs from the operator's")
        if G.t.numel() != M * self.n_out or F.t.numel() != M * self.n_in:  # @assert size(Gv)==size(Fu)
            raise _lib.SgcError(_lib.SGC_ERR_INVALID,
                                f"array sizes {G.shape} / {F.shape} do not match grid {self.N} with "
                                f"{self.n_out} output and {self.n_in} input components")

    def apply(self, G: DeviceArray, F: DeviceArray, op="=", stream=None):
        self._check_arrays(G, F)
        check(lib.sgc_op_apply(self._h, C.c_void_p(G.ptr), C.c_void_p(F.ptr), opcode(op), stream_handle(stream)))

    def apply_range(self, G, F, op, k_begin, k_end, ghost=None, stream=None):
        """Planes [k_begin, k_end) of the last array dimension; `ghost[c]` = device pointer (int)
        of the ghost plane of input component c, or None."""
        self._check_arrays(G, F)
        garr = None
        if ghost is not None:
            garr = (C.c_void_p * self.n_in)(*[None if g is None else int(g) for g in ghost])
        check(lib.sgc_op_apply_range(self._h, C.c_void_p(G.ptr), C.c_void_p(F.ptr), opcode(op), garr,
                                     int(k_begin), int(k_end), stream_handle(stream)))

    def attach_comm(self, comm_handle, lo_rank, hi_rank):
        """Neighbour synchronisation inside the fused curl kernel (`sgc_op_attach_comm`): the ranks owning the slab
        below / above this one, or None."""
        check(lib.sgc_op_attach_comm(self._h, comm_handle, -1 if lo_rank is None else int(lo_rank),
                                     -1 if hi_rank is None else int(hi_rank)))

    def apply_synced(self, G, F, op, ghost, stream=None):
        """One whole-slab apply whose neighbour synchronisation happens inside the kernel."""
        self._check_arrays(G, F)
        garr = None
        if ghost is not None:
            garr = (C.c_void_p * self.n_in)(*[None if g is None else int(g) for g in ghost])
        check(lib.sgc_op_apply_synced(self._h, C.c_void_p(G.ptr), C.c_void_p(F.ptr), opcode(op), garr,
                                      stream_handle(stream)))

    def apply_after(self, first: "Operator", G: DeviceArray, F: DeviceArray, op="=", k_begin=None, k_end=None,
                    ghost_lo=None, ghost_hi=None, stream=None):
        """Fused curl-of-curl `G OP self(first(F))` in one pass over memory (the intermediate field
        never leaves shared memory); bit-identical to `first.apply(H, F, "=")` followed by
        `self.apply(G, H, op)`.  Both must be full 3-D curl operators on the same grid.  With slab
        interfaces, `ghost_lo[c]` / `ghost_hi[c]` are device pointers (ints) of plane −1 / plane Nz
        of input component c."""
        first._check_arrays(F, F)
        self._check_arrays(G, G)
        Nz = self.N[-1]
        k_begin = 0 if k_begin is None else int(k_begin)
        k_end = Nz if k_end is None else int(k_end)
        glo = None if ghost_lo is None else (C.c_void_p * 3)(*[None if g is None else int(g) for g in ghost_lo])
        ghi = None if ghost_hi is None else (C.c_void_p * 3)(*[None if g is None else int(g) for g in ghost_hi])
        check(lib.sgc_op_apply_curlcurl_range(self._h, first._h, C.c_void_p(G.ptr), C.c_void_p(F.ptr), opcode(op), glo, ghi,
                                              k_begin, k_end, stream_handle(stream)))

    def set_slab_coef(self, dzinv_below=None, dzinv_above=None):
        """∆z⁻¹ of the planes just below / above this rank's slab (the neighbours' entries), needed by
        the fused curl-of-curl at slab interfaces."""
        vals = [v for v in (dzinv_below, dzinv_above) if v is not None]
        cc = int(any(np.iscomplexobj(np.asarray(v)) for v in vals))
        def pack(v):
            if v is None:
                return None
            z = complex(v)
            return (C.c_double * 2)(z.real, z.imag) if cc else (C.c_double * 1)(float(np.real(v)))
        a, b = pack(dzinv_below), pack(dzinv_above)
        check(lib.sgc_op_set_slab_coef(self._h, a, b, cc))

    def apply_host(self, G: np.ndarray, F: np.ndarray, op="=", nchunks=0):
        """Host arrays in, host arrays out (column-major numpy arrays, G modified in place): `sgc_op_apply_host` —
        the arrays are page-locked in place on first use, the pass is z-chunked and full-duplex."""
        self._check_host(G, self.n_out)
        self._check_host(F, self.n_in)
        check(lib.sgc_op_apply_host_chunked(self._h, C.c_void_p(G.ctypes.data), C.c_void_p(F.ctypes.data), opcode(op), int(nchunks)))

    def apply_after_host(self, first: "Operator", G: np.ndarray, F: np.ndarray, op="=", nchunks=0):
        """`G OP self(first(F))` on host arrays in ONE library call (`sgc_op_apply_curlcurl_host`)."""
        self._check_host(G, 3)
        first._check_host(F, 3)
        check(lib.sgc_op_apply_curlcurl_host(self._h, first._h, C.c_void_p(G.ctypes.data), C.c_void_p(F.ctypes.data),
                                             opcode(op), int(nchunks)))

    def _check_host(self, A: np.ndarray, ncomp):
        want = np.complex128 if self.dtype_code == _lib.SGC_C128 else np.float64
        M = int(np.prod(self.N, dtype=np.int64))
        if A.dtype != want or A.size != M * ncomp or not (A.flags.f_contiguous or A.ndim == 1):
            raise _lib.SgcError(_lib.SGC_ERR_INVALID, f"host array {A.shape} {A.dtype} does not match grid {self.N} x {ncomp} "
                                f"components of {np.dtype(want).name} in column-major order")
        _pin_for_lifetime(A)

    def set_slab(self, lo_interface: bool, hi_interface: bool):
        check(lib.sgc_op_set_slab(self._h, int(lo_interface), int(hi_interface)))

    def set_slab_weights(self, dw_below, dw_above):
        """∆w of the neighbour rank's plane just below / above this slab (weighted means along the slab axis)."""
        cc = int(any(np.iscomplexobj(np.asarray(v)) for v in (dw_below, dw_above)))
        def pack(v):
            z = complex(v)
            return (C.c_double * 2)(z.real, z.imag) if cc else (C.c_double * 1)(float(np.real(v)))
        check(lib.sgc_op_set_slab_weights(self._h, pack(dw_below), pack(dw_above), cc))
        self.weighted_axis = None  # the slab driver may now take it

    def set_tuning(self, tile_x=0, tile_y=0, march_z=0, variant=0):
        check(lib.sgc_op_set_tuning(self._h, tile_x, tile_y, march_z, variant))

    def launches(self, op="=") -> int:
        return lib.sgc_op_launches(self._h, opcode(op))

    def destroy(self):
        if self._h is not None and self._h.value:
            lib.sgc_op_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


_PINNED = {}  # data pointer of the owning array -> weakref.finalize (drops the registration before the memory is freed)


def _pin_for_lifetime(A: np.ndarray):
    """The library page-locks host arrays in place and caches the registration (`sgc_host_register`); the
    registration must be dropped before the memory is freed.  Tie it to the lifetime of the array that OWNS the
    memory: its finaliser calls `sgc_host_unregister`."""
    import weakref
    owner = A
    while isinstance(getattr(owner, "base", None), np.ndarray):
        owner = owner.base
    if getattr(owner, "base", None) is not None or not isinstance(owner, np.ndarray):
        return  # memory owned by something else (mmap, torch, ...): leave it to the caller / pageable path
    ptr = owner.ctypes.data
    f = _PINNED.get(ptr)
    if f is not None and f.alive:
        return
    lib.sgc_host_register(C.c_void_p(ptr), owner.nbytes)
    _PINNED[ptr] = weakref.finalize(owner, _unpin, ptr)


def _unpin(ptr):
    _PINNED.pop(ptr, None)
    try:
        lib.sgc_host_unregister(C.c_void_p(ptr))
    except Exception:
        pass


def _sync_and_destroy(op: Operator, stream):
    import torch
    (stream or torch.cuda.current_stream()).synchronize()
    op.destroy()


def apply_partial(Gv: DeviceArray, Fu: DeviceArray, op, nw, isfwd, dwinv=1.0, isbloch=True, e=1.0, *, alpha=1.0,
                  stream=None):
    """`apply_∂!(Gv, Fu, Val(OP), nw, isfwd, ∆w⁻¹, isbloch, e⁻ⁱᵏᴸ; α)` —
    matrixfree/differential/partial.jl:11-21 (scalar ∆), :33 / :241 / :388 (3-D/2-D/1-D)."""
    if Gv.shape != Fu.shape:  # @assert size(Gv)==size(Fu)
        raise _lib.SgcError(_lib.SGC_ERR_INVALID, f"size(Gv) = {Gv.shape} != size(Fu) = {Fu.shape}")
    N = Fu.shape
    if not 1 <= nw <= len(N):
        raise _lib.SgcError(_lib.SGC_ERR_INVALID, f"nw = {nw} out of range")
    p, keep, cc = coef_vector(dwinv, N[nw - 1])
    check(lib.sgc_apply_partial(C.c_void_p(Gv.ptr), C.c_void_p(Fu.ptr), opcode(op), Fu.dtype_code, len(N), int64s(N),
                                nw, int(isfwd), p, cc, int(isbloch), scalar2(e), scalar2(alpha),
                                stream_handle(stream)))
    return None


def apply_mhat(Gv: DeviceArray, Fu: DeviceArray, op, nw, isfwd, dw=None, dwpinv=None, isbloch=True, e=1.0, *,
               alpha=1.0, stream=None):
    """`apply_m̂!` — matrixfree/mean.jl:102/295/432 (arithmetic: omit `dw`, `dwpinv`) and
    :493/705/856 (weighted)."""
    if Gv.shape != Fu.shape:
        raise _lib.SgcError(_lib.SGC_ERR_INVALID, f"size(Gv) = {Gv.shape} != size(Fu) = {Fu.shape}")
    o = Operator.mhat(Fu.dtype_code, Fu.shape, nw, isfwd, dw, dwpinv, isbloch, e, alpha)
    o.apply(Gv, Fu, op, stream)
    _sync_and_destroy(o, stream)
    return None


def apply_curl(G: DeviceArray, F: DeviceArray, op, isfwd, dlinv=None, isbloch=None, e=None, *, cmp_shp=None,
               cmp_out=None, cmp_in=None, alpha=1.0, stream=None):
    """`apply_curl!(G, F, Val(OP), isfwd, ∆l⁻¹, isbloch, e⁻ⁱᵏᴸ; cmp_shp, cmp_out, cmp_in, α)` —
    matrixfree/differential/curl.jl:12-24, :27-49, :52-137."""
    K = G.ndim - 1
    if F.ndim != K + 1 or G.shape[:K] != F.shape[:K]:
        raise _lib.SgcError(_lib.SGC_ERR_INVALID, f"G {G.shape} and F {F.shape} are not fields on one grid")
    cmp_out = list(range(1, G.shape[K] + 1)) if cmp_out is None else list(cmp_out)  # curl.jl:20-21
    cmp_in = list(range(1, F.shape[K] + 1)) if cmp_in is None else list(cmp_in)
    if len(cmp_out) != G.shape[K] or len(cmp_in) != F.shape[K]:
        raise _lib.SgcError(_lib.SGC_ERR_INVALID, "cmp_out / cmp_in do not match the component dimension")
    o = Operator.curl(F.dtype_code, G.shape[:K], isfwd, dlinv, isbloch, e, cmp_shp, cmp_out, cmp_in, alpha)
    o.apply(G, F, op, stream)
    _sync_and_destroy(o, stream)
    return None


def apply_curlcurl(G: DeviceArray, F: DeviceArray, op, isfwd1, dlinv1, isfwd2, dlinv2, isbloch=None, e=None, *,
                   alpha1=1.0, alpha2=1.0, stream=None):
    """`G OP ∇₂×(∇₁×F)`: the pair `apply_curl!(H, F, Val(:(=)), isfwd1, ∆l₁⁻¹, …)`,
    `apply_curl!(G, H, Val(OP), isfwd2, ∆l₂⁻¹, …)` (matrixfree/differential/curl.jl:52-137 twice) in
    one pass over memory, H never materialised; results are bit-identical to the two calls."""
    if G.shape != F.shape or G.ndim != 4 or G.shape[3] != 3:
        raise _lib.SgcError(_lib.SGC_ERR_INVALID, f"G {G.shape} / F {F.shape} are not 3-component fields on one 3-D grid")
    N = G.shape[:3]
    c1 = Operator.curl(F.dtype_code, N, isfwd1, dlinv1, isbloch, e, alpha=alpha1)
    c2 = Operator.curl(F.dtype_code, N, isfwd2, dlinv2, isbloch, e, alpha=alpha2)
    c2.apply_after(c1, G, F, op, stream=stream)
    _sync_and_destroy(c1, stream)
    c2.destroy()
    return None


def apply_divg(g: DeviceArray, F: DeviceArray, op, isfwd, dlinv=None, isbloch=None, e=None, *, alpha=1.0,
               stream=None):
    """`apply_divg!` — matrixfree/differential/divergence.jl:17-66."""
    K = g.ndim
    if F.ndim != K + 1 or F.shape != g.shape + (K,):  # @assert K₊₁==K+1
        raise _lib.SgcError(_lib.SGC_ERR_INVALID, f"F {F.shape} is not a {K}-component field on grid {g.shape}")
    o = Operator.divg(F.dtype_code, g.shape, isfwd, dlinv, isbloch, e, alpha)
    o.apply(g, F, op, stream)
    _sync_and_destroy(o, stream)
    return None


def apply_grad(G: DeviceArray, f: DeviceArray, op, isfwd, dlinv=None, isbloch=None, e=None, *, alpha=1.0,
               stream=None):
    """`apply_grad!` — matrixfree/differential/gradient.jl:17-59."""
    K = f.ndim
    if G.ndim != K + 1 or G.shape != f.shape + (K,):
        raise _lib.SgcError(_lib.SGC_ERR_INVALID, f"G {G.shape} is not a {K}-component field on grid {f.shape}")
    o = Operator.grad(f.dtype_code, f.shape, isfwd, dlinv, isbloch, e, alpha)
    o.apply(G, f, op, stream)
    _sync_and_destroy(o, stream)
    return None


def apply_mean(G: DeviceArray, F: DeviceArray, op, isfwd, dl=None, dlpinv=None, isbloch=None, e=None, *, alpha=1.0,
               stream=None):
    """`apply_mean!` — matrixfree/mean.jl:21-29 / :47-65 (arithmetic), :32-42 / :69-89 (weighted)."""
    K = G.ndim - 1
    if G.shape != F.shape or G.shape[K] != K:
        raise _lib.SgcError(_lib.SGC_ERR_INVALID, f"G {G.shape} / F {F.shape} are not {K}-component fields")
    o = Operator.mean(F.dtype_code, G.shape[:K], isfwd, dl, dlpinv, isbloch, e, alpha)
    o.apply(G, F, op, stream)
    _sync_and_destroy(o, stream)
    return None
