"""ctypes wrapper of oracle/libsgc_oracle.so (the C restatement of the reference's CPU
algorithms).  TEST INFRASTRUCTURE / CPU BASELINE ONLY — see oracle/sgc_oracle_c.c."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "libsgc_oracle.so")


class _CX(C.Structure):
    _fields_ = [("re", C.c_double), ("im", C.c_double)]


def _load():
    if not os.path.exists(_PATH):
        subprocess.run(["make", "-C", _HERE], check=True)
    lib = C.CDLL(_PATH)
    i64 = C.c_int64
    lib.sgc_ref_apply_partial_c128.restype = None
    lib.sgc_ref_apply_partial_c128.argtypes = [C.c_void_p, C.c_void_p, i64, i64, i64, C.c_int, C.c_int, C.c_void_p,
                                               C.c_int, _CX, _CX, C.c_int]
    lib.sgc_ref_apply_curl_c128.restype = None
    lib.sgc_ref_apply_curl_c128.argtypes = [C.c_void_p, C.c_void_p, i64, i64, i64, C.POINTER(C.c_int), C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.c_void_p, _CX, C.c_int]
    lib.sgc_ref_create_curl_f64.restype = i64
    lib.sgc_ref_create_curl_f64.argtypes = [i64, i64, i64, C.POINTER(C.c_int), C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.POINTER(C.c_int), C.c_void_p, C.c_int, C.POINTER(C.c_void_p),
                                            C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
    lib.sgc_ref_free.argtypes = [C.c_void_p]
    lib.sgc_ref_max_threads.restype = C.c_int
    lib.sgc_ref_set_threads.argtypes = [C.c_int]
    return lib


lib = _load()


def _cx(z):
    z = complex(z)
    return _CX(z.real, z.imag)


def _c128(v):
    return np.ascontiguousarray(v, dtype=np.complex128)


def max_threads() -> int:
    return int(lib.sgc_ref_max_threads())


def set_threads(n: int):
    lib.sgc_ref_set_threads(int(n))


def apply_partial(Gv, Fu, op, nw, isfwd, dwinv, isbloch, e=1.0, alpha=1.0):
    """3-D ComplexF64 `apply_∂!`; Gv, Fu column-major (Nx,Ny,Nz) complex128 arrays."""
    assert Gv.flags.f_contiguous and Fu.flags.f_contiguous and Gv.dtype == np.complex128
    Nx, Ny, Nz = Fu.shape
    d = _c128(dwinv)
    lib.sgc_ref_apply_partial_c128(Gv.ctypes.data, Fu.ctypes.data, Nx, Ny, Nz, nw, int(isfwd), d.ctypes.data,
                                   int(isbloch), _cx(e), _cx(alpha), int(op == "+="))


def apply_curl(G, F, op, isfwd, dlinv, isbloch, e, alpha=1.0):
    """Full 3-D ComplexF64 `apply_curl!`; G, F column-major (Nx,Ny,Nz,3)."""
    assert G.flags.f_contiguous and F.flags.f_contiguous and G.dtype == np.complex128 and F.dtype == np.complex128
    Nx, Ny, Nz, _ = F.shape
    d = [_c128(v) for v in dlinv]
    ee = _c128(np.asarray(e))
    fw = (C.c_int * 3)(*[int(b) for b in isfwd])
    bl = (C.c_int * 3)(*[int(b) for b in isbloch])
    lib.sgc_ref_apply_curl_c128(G.ctypes.data, F.ctypes.data, Nx, Ny, Nz, fw, d[0].ctypes.data, d[1].ctypes.data,
                                d[2].ctypes.data, bl, ee.ctypes.data, _cx(alpha), int(op == "+="))


def create_curl(isfwd, dlinv, isbloch, e, order_cmpfirst=True):
    """Float64 `create_curl`; returns (colptr, rowval, nzval) 1-based numpy arrays."""
    d = [np.ascontiguousarray(v, dtype=np.float64) for v in dlinv]
    Nx, Ny, Nz = (len(v) for v in d)
    ee = np.ascontiguousarray(e, dtype=np.float64)
    fw = (C.c_int * 3)(*[int(b) for b in isfwd])
    bl = (C.c_int * 3)(*[int(b) for b in isbloch])
    cp, rv, nz = C.c_void_p(), C.c_void_p(), C.c_void_p()
    nnz = lib.sgc_ref_create_curl_f64(Nx, Ny, Nz, fw, d[0].ctypes.data, d[1].ctypes.data, d[2].ctypes.data, bl,
                                      ee.ctypes.data, int(order_cmpfirst), C.byref(cp), C.byref(rv), C.byref(nz))
    if nnz < 0:
        raise MemoryError("C oracle create_curl ran out of memory")
    n = 3 * Nx * Ny * Nz
    colptr = np.ctypeslib.as_array(C.cast(cp, C.POINTER(C.c_int64)), shape=(n + 1,)).copy()
    rowval = np.ctypeslib.as_array(C.cast(rv, C.POINTER(C.c_int64)), shape=(max(nnz, 1),))[:nnz].copy()
    nzval = np.ctypeslib.as_array(C.cast(nz, C.POINTER(C.c_double)), shape=(max(nnz, 1),))[:nnz].copy()
    for p in (cp, rv, nz):
        lib.sgc_ref_free(p)
    return colptr, rowval, nzval


def time_create_curl(isfwd, dlinv, isbloch, e, order_cmpfirst=True):
    """Run create_curl, free the result, return (seconds, nnz) — for the CPU baseline."""
    import time
    d = [np.ascontiguousarray(v, dtype=np.float64) for v in dlinv]
    Nx, Ny, Nz = (len(v) for v in d)
    ee = np.ascontiguousarray(e, dtype=np.float64)
    fw = (C.c_int * 3)(*[int(b) for b in isfwd])
    bl = (C.c_int * 3)(*[int(b) for b in isbloch])
    cp, rv, nz = C.c_void_p(), C.c_void_p(), C.c_void_p()
    t0 = time.perf_counter()
    nnz = lib.sgc_ref_create_curl_f64(Nx, Ny, Nz, fw, d[0].ctypes.data, d[1].ctypes.data, d[2].ctypes.data, bl,
                                      ee.ctypes.data, int(order_cmpfirst), C.byref(cp), C.byref(rv), C.byref(nz))
    dt = time.perf_counter() - t0
    for p in (cp, rv, nz):
        lib.sgc_ref_free(p)
    return dt, int(nnz)
