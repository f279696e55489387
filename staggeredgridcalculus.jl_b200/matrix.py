"""Host-side mirror of the reference's sparse-assembly API (src/matrix), over the C ABI.

Every `create_*` returns a `DeviceCSC`: the three arrays of a `SparseMatrixCSC{T,Int64}`
(`colptr`, `rowval`, `nzval`; 1-based by default, exactly what Julia's constructor takes)
resident in HBM.  `T` follows the reference's `promote_type` rule: real inputs give a real
matrix (matrix/differential/curl.jl:14-21, :59).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import lib, check
from .device import coef_vector, coef_vectors, ints, int64s, scalar2, scalars2, stream_handle, require_cuda
from .matrixfree import _together

__all__ = ["DeviceCSC", "create_partial", "create_mhat", "create_mean", "create_curl", "create_divg",
           "create_grad", "create_picmp"]


class DeviceCSC:
    """`SparseMatrixCSC{Tv,Int64}` on the device."""

    def __init__(self, m, n, colptr, rowval, nzval, index_base):
        self.m, self.n = int(m), int(n)
        self.colptr, self.rowval, self.nzval = colptr, rowval, nzval
        self.index_base = int(index_base)

    @property
    def nnz(self) -> int:
        return int(self.rowval.numel())

    def to_host(self):
        return self.colptr.cpu().numpy(), self.rowval.cpu().numpy(), self.nzval.cpu().numpy()

    def to_scipy(self):
        import scipy.sparse as sp
        cp, rv, nz = self.to_host()
        return sp.csc_matrix((nz, rv - self.index_base, cp - self.index_base), shape=(self.m, self.n))

    def to_pinned_host(self):
        """colptr / rowval / nzval copied into page-locked host tensors (what a Julia caller wraps as
        `SparseMatrixCSC(m, n, colptr, rowval, nzval)` without a further copy)."""
        outs = []
        for t in (self.colptr, self.rowval, self.nzval):
            h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            h.copy_(t, non_blocking=True)
            outs.append(h)
        torch.cuda.current_stream().synchronize()
        return tuple(outs)

    def save_npz(self, path):
        """On-disk CSC dump for cross-language parity checks (numpy `.npz`: m, n, index_base, colptr, rowval,
        nzval — Julia reads it with NPZ.jl, scipy with `load_npz_csc`)."""
        cp, rv, nz = self.to_host()
        np.savez(path, m=self.m, n=self.n, index_base=self.index_base, colptr=cp, rowval=rv, nzval=nz)

    def mul(self, x: torch.Tensor, y: "torch.Tensor | None" = None) -> torch.Tensor:
        """`mul!(y, A, x)` on the device through the row-sorted plan (`sgc_spmv_*`): the plan is built on first use and
        kept; every product is a gather without atomics, accumulated per row in ascending column order."""
        code = _lib.SGC_C128 if self.nzval.dtype == torch.complex128 else _lib.SGC_F64
        if getattr(self, "_plan", None) is None:
            h = C.c_void_p()
            check(lib.sgc_spmv_create(C.byref(h), self.m, self.n, C.c_void_p(self.colptr.data_ptr()), C.c_void_p(self.rowval.data_ptr()),
                                      self.index_base, stream_handle()))
            self._plan = h
        x = x.to(self.nzval.dtype).contiguous()
        if y is None:
            y = torch.empty(self.m, dtype=self.nzval.dtype, device=x.device)
        check(lib.sgc_spmv_apply(self._plan, C.c_void_p(self.nzval.data_ptr()), code, C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()),
                                 stream_handle()))
        return y

    def __del__(self):
        try:
            if getattr(self, "_plan", None) is not None:
                lib.sgc_spmv_destroy(self._plan)
                self._plan = None
        except Exception:
            pass

    def matvec(self, x: torch.Tensor) -> torch.Tensor:
        """`A * x` on the device, column-parallel with atomics (no plan; the reference's cross-check `mul!(g, A, f)`)."""
        code = _lib.SGC_C128 if self.nzval.dtype == torch.complex128 else _lib.SGC_F64
        x = x.to(self.nzval.dtype).contiguous()
        y = torch.empty(self.m, dtype=self.nzval.dtype, device=x.device)
        check(lib.sgc_csc_matvec(self.m, self.n, C.c_void_p(self.colptr.data_ptr()), C.c_void_p(self.rowval.data_ptr()),
                                 C.c_void_p(self.nzval.data_ptr()), self.index_base, code, C.c_void_p(x.data_ptr()),
                                 C.c_void_p(y.data_ptr()), stream_handle()))
        return y


def load_npz_csc(path):
    """Inverse of `DeviceCSC.save_npz` on the host: a scipy `csc_matrix`."""
    import scipy.sparse as sp
    d = np.load(path)
    b = int(d["index_base"])
    return sp.csc_matrix((d["nzval"], d["rowval"] - b, d["colptr"] - b), shape=(int(d["m"]), int(d["n"])))


class _Asm:
    def __init__(self, h, keep):
        self.h, self.keep = h, keep

    def __del__(self):
        if self.h is not None and self.h.value:
            lib.sgc_asm_destroy(self.h)
            self.h = None


def col_partition(a: "_Asm", part, nparts):
    """Aligned, balanced column range of part `part` of `nparts` (`sgc_asm_col_partition`)."""
    c0, c1 = C.c_int64(), C.c_int64()
    check(lib.sgc_asm_col_partition(a.h, int(part), int(nparts), C.byref(c0), C.byref(c1)))
    return c0.value, c1.value


def _finish(a: _Asm, index_base, row_range, method, device, stream=None) -> DeviceCSC:
    if isinstance(row_range, dict):  # {"cols": (part, nparts)} or {"cols": (col_begin, col_end, "abs")}: a block of COLUMNS
        spec = row_range["cols"]
        c0, c1 = (spec[0], spec[1]) if len(spec) == 3 else col_partition(a, spec[0], spec[1])
        check(lib.sgc_asm_set_col_range(a.h, int(c0), int(c1)))
        row_range = None
    if row_range is not None:
        check(lib.sgc_asm_set_row_range(a.h, int(row_range[0]), int(row_range[1])))
    m, n, vt = C.c_int64(), C.c_int64(), C.c_int()
    check(lib.sgc_asm_shape(a.h, C.byref(m), C.byref(n), C.byref(vt)))
    vdt = torch.complex128 if vt.value == _lib.SGC_C128 else torch.float64
    dev = device or "cuda"
    colptr = torch.empty(n.value + 1, dtype=torch.int64, device=dev)
    nnz = C.c_int64()
    st = stream_handle(stream)
    if method == "counted":  # the count kernel + device scan instead of the closed-form offsets
        check(lib.sgc_asm_set_tuning(a.h, 1))
    if method == "general":  # interior cells through the general entry path (block walk + sort), not the host-sorted descriptors
        check(lib.sgc_asm_set_tuning(a.h, 2))
    if method == "cta":  # the CTA-granular fill kernel (block scan) instead of the warp-granular one
        check(lib.sgc_asm_set_tuning(a.h, 4))
    if method == "onecol":  # the warp-granular kernel with one column per thread (no several-cells-per-thread form)
        check(lib.sgc_asm_set_tuning(a.h, 8))
    # nnz (closed form, or the count pass for a row partition) so that rowval / nzval can be allocated
    check(lib.sgc_asm_count(a.h, C.byref(nnz), st))
    rowval = torch.empty(nnz.value, dtype=torch.int64, device=dev)
    nzval = torch.empty(nnz.value, dtype=vdt, device=dev)
    if method in ("direct", "counted", "general", "cta", "onecol"):  # one pass: colptr + rowval + nzval
        check(lib.sgc_asm_fill_all(a.h, C.c_void_p(colptr.data_ptr()), index_base, C.c_void_p(rowval.data_ptr()),
                                   C.c_void_p(nzval.data_ptr()), st))
    elif method == "twostep":  # colptr first, then rowval / nzval from the given colptr
        nnz2 = C.c_int64()
        check(lib.sgc_asm_colptr(a.h, C.c_void_p(colptr.data_ptr()), index_base, C.byref(nnz2), st))
        assert nnz2.value == nnz.value
        check(lib.sgc_asm_fill(a.h, C.c_void_p(colptr.data_ptr()), index_base, C.c_void_p(rowval.data_ptr()),
                               C.c_void_p(nzval.data_ptr()), st))
    elif method == "coo":  # generic cross-check: COO triplets → stable radix sort → sum → drop zeros
        nnz2 = C.c_int64()
        check(lib.sgc_asm_coo_sort(a.h, C.c_void_p(colptr.data_ptr()), index_base, C.c_void_p(rowval.data_ptr()),
                                   C.c_void_p(nzval.data_ptr()), nnz.value, C.byref(nnz2), st))
        if nnz2.value != nnz.value:
            raise _lib.SgcError(_lib.SGC_ERR_INVALID, f"COO path found nnz = {nnz2.value}, closed form {nnz.value}")
    else:
        raise ValueError(method)
    return DeviceCSC(m.value, n.value, colptr, rowval, nzval, index_base)


def _defaults(K, isbloch, e):
    isbloch = [True] * K if isbloch is None else list(isbloch)
    e = [1.0] * K if e is None else e
    e_arr = np.asarray(e)
    return isbloch, list(e_arr), int(np.iscomplexobj(e_arr))


def create_partial(nw, isfwd, N, dwinv=1.0, isbloch=True, e=1.0, *, index_base=1, row_range=None,
                   method="direct", device=None) -> DeviceCSC:
    """`create_∂(nw, isfwd, N, ∆w⁻¹, isbloch, e⁻ⁱᵏᴸ)` — matrix/differential/partial.jl:45-60."""
    require_cuda()
    N = tuple(int(n) for n in N)
    if not 1 <= nw <= len(N):
        raise _lib.SgcError(_lib.SGC_ERR_INVALID, f"nw = {nw} out of range")
    p, keep, cc = coef_vector(dwinv, N[nw - 1])
    h = C.c_void_p()
    check(lib.sgc_asm_create_partial(C.byref(h), len(N), int64s(N), nw, int(isfwd), p, cc, int(isbloch), scalar2(e),
                                     int(np.iscomplexobj(np.asarray(e)))))
    return _finish(_Asm(h, keep), index_base, row_range, method, device)


def create_mhat(nw, isfwd, N, dw=None, dwpinv=None, isbloch=True, e=1.0, *, index_base=1, row_range=None,
                method="direct", device=None) -> DeviceCSC:
    """`create_m̂(nw, isfwd, N, ∆w, ∆w′⁻¹, isbloch, e⁻ⁱᵏᴸ)` — matrix/mean.jl:124-138."""
    require_cuda()
    N = tuple(int(n) for n in N)
    if not 1 <= nw <= len(N):
        raise _lib.SgcError(_lib.SGC_ERR_INVALID, f"nw = {nw} out of range")
    if dw is None:  # ∆w = ones(N[nw]); ∆w′⁻¹ = ones(N[nw])   (mean.jl:129)
        dw, dwpinv = np.ones(N[nw - 1]), np.ones(N[nw - 1])
    dw, dwpinv = _together(dw, dwpinv)
    p1, k1, c1 = coef_vector(dw, N[nw - 1])
    p2, k2, c2 = coef_vector(dwpinv, N[nw - 1])
    h = C.c_void_p()
    check(lib.sgc_asm_create_mhat(C.byref(h), len(N), int64s(N), nw, int(isfwd), p1, p2, c1 or c2, int(isbloch),
                                  scalar2(e), int(np.iscomplexobj(np.asarray(e)))))
    return _finish(_Asm(h, (k1, k2)), index_base, row_range, method, device)


def _grid_from(dlinv, N, K):
    if N is None:
        N = tuple(len(v) for v in dlinv)
    N = tuple(int(n) for n in N)
    dlinv = [1.0] * K if dlinv is None else dlinv
    return N, dlinv


def create_curl(isfwd, dlinv=None, isbloch=None, e=None, *, N=None, cmp_shp=None, cmp_out=None, cmp_in=None,
                order_cmpfirst=True, index_base=1, row_range=None, method="direct", device=None) -> DeviceCSC:
    """`create_curl(isfwd, ∆l⁻¹, isbloch, e⁻ⁱᵏᴸ; cmp_shp, cmp_out, cmp_in, order_cmpfirst)` —
    matrix/differential/curl.jl:4-36 (wrappers; pass `N` with scalar ∆l⁻¹) and :50-104."""
    require_cuda()
    K = len(isfwd)
    N, dlinv = _grid_from(dlinv, N, K)
    isbloch, e, ec = _defaults(K, isbloch, e)
    cmp_shp = list(range(1, K + 1)) if cmp_shp is None else list(cmp_shp)
    cmp_out = list(range(1, K + 1)) if cmp_out is None else list(cmp_out)
    cmp_in = list(range(1, K + 1)) if cmp_in is None else list(cmp_in)
    arr, keep, cc = coef_vectors(dlinv, N)
    h = C.c_void_p()
    check(lib.sgc_asm_create_curl(C.byref(h), K, int64s(N), ints(isfwd), arr, cc, ints(isbloch), scalars2(e), ec,
                                  ints(cmp_shp), len(cmp_out), ints(cmp_out), len(cmp_in), ints(cmp_in),
                                  int(order_cmpfirst)))
    return _finish(_Asm(h, keep), index_base, row_range, method, device)


def _divgrad(fn, isfwd, dlinv, isbloch, e, N, order_cmpfirst, index_base, row_range, method, device):
    require_cuda()
    K = len(isfwd)
    N, dlinv = _grid_from(dlinv, N, K)
    isbloch, e, ec = _defaults(K, isbloch, e)
    arr, keep, cc = coef_vectors(dlinv, N)
    h = C.c_void_p()
    check(fn(C.byref(h), K, int64s(N), ints(isfwd), arr, cc, ints(isbloch), scalars2(e), ec, int(order_cmpfirst)))
    return _finish(_Asm(h, keep), index_base, row_range, method, device)


def create_divg(isfwd, dlinv=None, isbloch=None, e=None, *, N=None, order_cmpfirst=True, index_base=1,
                row_range=None, method="direct", device=None) -> DeviceCSC:
    """`create_divg` — matrix/differential/divergence.jl:12-70."""
    return _divgrad(lib.sgc_asm_create_divg, isfwd, dlinv, isbloch, e, N, order_cmpfirst, index_base, row_range,
                    method, device)


def create_grad(isfwd, dlinv=None, isbloch=None, e=None, *, N=None, order_cmpfirst=True, index_base=1,
                row_range=None, method="direct", device=None) -> DeviceCSC:
    """`create_grad` — matrix/differential/gradient.jl:12-70."""
    return _divgrad(lib.sgc_asm_create_grad, isfwd, dlinv, isbloch, e, N, order_cmpfirst, index_base, row_range,
                    method, device)


def create_mean(isfwd, dl=None, dlpinv=None, isbloch=None, e=None, *, N=None, order_cmpfirst=True, index_base=1,
                row_range=None, method="direct", device=None) -> DeviceCSC:
    """`create_mean` — matrix/mean.jl:48-63 (wrappers; arithmetic: pass `N`), :72-111."""
    require_cuda()
    K = len(isfwd)
    if dl is None:  # fill.(1.0, N)   (mean.jl:53)
        dl = [np.ones(int(n)) for n in N]
        dlpinv = [np.ones(int(n)) for n in N]
    N = tuple(len(v) for v in dl)
    isbloch, e, ec = _defaults(K, isbloch, e)
    arr, keep, cc = coef_vectors(list(dl) + list(dlpinv), N + N)
    a1 = (C.c_void_p * K)(*[arr[i] for i in range(K)])
    a2 = (C.c_void_p * K)(*[arr[K + i] for i in range(K)])
    h = C.c_void_p()
    check(lib.sgc_asm_create_mean(C.byref(h), K, int64s(N), ints(isfwd), a1, a2, cc, ints(isbloch), scalars2(e), ec,
                                  int(order_cmpfirst)))
    return _finish(_Asm(h, keep), index_base, row_range, method, device)


def create_picmp(N, permute, *, scale=None, order_cmpfirst=True, index_base=1, row_range=None, method="direct",
                 device=None) -> DeviceCSC:
    """`create_πcmp(N, permute; scale, order_cmpfirst)` — matrix/permutation.jl:7-40."""
    require_cuda()
    N = tuple(int(n) for n in N)
    Kf = len(permute)
    scale = np.ones(Kf) if scale is None else np.asarray(scale)
    sc = np.iscomplexobj(scale)
    k = np.ascontiguousarray(scale, dtype=np.complex128 if sc else np.float64)
    h = C.c_void_p()
    check(lib.sgc_asm_create_picmp(C.byref(h), len(N), int64s(N), Kf, ints(permute), C.c_void_p(k.ctypes.data),
                                   int(sc), int(order_cmpfirst)))
    return _finish(_Asm(h, k), index_base, row_range, method, device)
