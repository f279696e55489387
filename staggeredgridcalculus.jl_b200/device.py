"""Device arrays with Julia's (column-major) shape semantics, and argument marshalling.

PyTorch is used only as plumbing here: it owns the HBM allocations and the CUDA streams whose
raw handles are passed through the C ABI.  No torch operator ever touches the field data.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

_NP2CODE = {np.dtype(np.float64): _lib.SGC_F64, np.dtype(np.complex128): _lib.SGC_C128}
_TORCH = {_lib.SGC_F64: torch.float64, _lib.SGC_C128: torch.complex128}


def require_cuda():
    if not torch.cuda.is_available():
        raise _lib.SgcError(_lib.SGC_ERR_CUDA, "no CUDA device: libsgc_b200 has no CPU fallback")


class DeviceArray:
    """`Array{Float64|ComplexF64,N}` living in HBM, indexed and laid out like the Julia array
    (first index fastest).  The analogue of the shim's `DeviceArray{T,N}`."""

    def __init__(self, tensor: torch.Tensor, shape):
        assert tensor.dim() == 1 and tensor.is_contiguous()
        self.t = tensor
        self.shape = tuple(int(s) for s in shape)
        assert int(np.prod(self.shape, dtype=np.int64)) == tensor.numel()

    # ---- construction -------------------------------------------------------------------
    @classmethod
    def empty(cls, shape, dtype=np.complex128, device=None):
        require_cuda()
        code = _NP2CODE[np.dtype(dtype)]
        n = int(np.prod(shape, dtype=np.int64))
        return cls(torch.empty(n, dtype=_TORCH[code], device=device or "cuda"), shape)

    @classmethod
    def zeros(cls, shape, dtype=np.complex128, device=None):
        a = cls.empty(shape, dtype, device)
        a.t.zero_()
        return a

    @classmethod
    def from_numpy(cls, a: np.ndarray, device=None):
        require_cuda()
        a = np.asarray(a)
        if a.dtype not in _NP2CODE:
            a = a.astype(np.complex128 if np.iscomplexobj(a) else np.float64)
        flat = np.ascontiguousarray(a.reshape(-1, order="F"))
        return cls(torch.from_numpy(flat).to(device or "cuda"), a.shape)

    # ---- access -------------------------------------------------------------------------
    def numpy(self) -> np.ndarray:
        return self.t.cpu().numpy().reshape(self.shape, order="F")

    @property
    def dtype_code(self) -> int:
        return _lib.SGC_C128 if self.t.dtype == torch.complex128 else _lib.SGC_F64

    @property
    def ptr(self) -> int:
        return self.t.data_ptr()

    @property
    def ndim(self):
        return len(self.shape)

    def component(self, w: int) -> "DeviceArray":
        """`selectdim(A, ndims(A), w)` (1-based w): contiguous slice of the last dimension."""
        m = int(np.prod(self.shape[:-1], dtype=np.int64))
        return DeviceArray(self.t[(w - 1) * m:w * m], self.shape[:-1])

    def copy(self) -> "DeviceArray":
        return DeviceArray(self.t.clone(), self.shape)


# ---- marshalling helpers ---------------------------------------------------------------------
def scalar2(x):
    x = complex(x)
    return (C.c_double * 2)(x.real, x.imag)


def scalars2(xs):
    arr = (C.c_double * (2 * len(xs)))()
    for i, x in enumerate(xs):
        x = complex(x)
        arr[2 * i], arr[2 * i + 1] = x.real, x.imag
    return arr


def ints(xs):
    return (C.c_int * len(xs))(*[int(x) for x in xs])


def int64s(xs):
    return (C.c_int64 * len(xs))(*[int(x) for x in xs])


def is_complex_scalar(x) -> bool:
    return isinstance(x, (complex, np.complexfloating))


def coef_vectors(vecs, N):
    """Per-axis coefficient vectors → (void** array, keepalive, coef_complex).

    Scalars are expanded like the reference's wrappers (`fill.(∆l⁻¹, N)`,
    matrixfree/differential/curl.jl:24); mixed real/complex tuples are promoted to complex,
    which is what the concrete methods' `promote_type` does (matrix/differential/curl.jl:59)."""
    if vecs is None:
        return None, [], 0
    out = []
    for v, n in zip(vecs, N):
        v = np.full(int(n), v) if np.ndim(v) == 0 else np.asarray(v)
        if len(v) != n:
            raise _lib.SgcError(_lib.SGC_ERR_INVALID, f"coefficient vector of length {len(v)} for an axis of length {n}")
        out.append(v)
    cplx = any(np.iscomplexobj(v) for v in out)
    keep = [np.ascontiguousarray(v, dtype=np.complex128 if cplx else np.float64) for v in out]
    arr = (C.c_void_p * len(keep))(*[k.ctypes.data for k in keep])
    return arr, keep, int(cplx)


def coef_vector(v, n):
    if v is None:
        return None, None, 0
    v = np.full(int(n), v) if np.ndim(v) == 0 else np.asarray(v)
    if len(v) != n:
        raise _lib.SgcError(_lib.SGC_ERR_INVALID, f"coefficient vector of length {len(v)} for an axis of length {n}")
    cplx = np.iscomplexobj(v)
    k = np.ascontiguousarray(v, dtype=np.complex128 if cplx else np.float64)
    return C.c_void_p(k.ctypes.data), k, int(cplx)


def stream_handle(stream=None):
    if stream is None:
        stream = torch.cuda.current_stream()
    if isinstance(stream, torch.cuda.Stream):
        return C.c_void_p(stream.cuda_stream)
    return C.c_void_p(int(stream))


def opcode(op) -> int:
    """`Val(:(=))` / `Val(:(+=))`."""
    if op in ("=", _lib.SGC_OP_SET):
        return _lib.SGC_OP_SET
    if op in ("+=", _lib.SGC_OP_ADD):
        return _lib.SGC_OP_ADD
    raise _lib.SgcError(_lib.SGC_ERR_INVALID, f"op must be '=' or '+=', got {op!r}")
