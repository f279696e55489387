"""ctypes binding of libsgc_b200.so (include/sgc_b200.h).

The library is the product; this module only loads it and declares the prototypes.  There is
no fallback of any kind: if the shared library has not been built, importing the package
raises, and every compute entry point fails with SGC_ERR_CUDA on a box without a GPU.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libsgc_b200.so")

SGC_OK, SGC_ERR_INVALID, SGC_ERR_CUDA, SGC_ERR_INEXACT, SGC_ERR_NOMEM = 0, 1, 2, 3, 4
SGC_F64, SGC_C128 = 0, 1
SGC_OP_SET, SGC_OP_ADD = 0, 1
SGC_IPC_HANDLE_BYTES = 64


class SgcError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libsgc_b200 error {code}: {msg}")
        self.code = code


class InexactError(SgcError, TypeError):
    """Julia's InexactError: a complex α / e⁻ⁱᵏᴸ / ∆ with a Float64 field."""


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "or `make -C staggeredgridcalculus.jl_b200/csrc`.  There is no CPU fallback.")

lib = C.CDLL(LIB_PATH)

i64p = C.POINTER(C.c_int64)
intp = C.POINTER(C.c_int)
dblp = C.POINTER(C.c_double)
vpp = C.POINTER(C.c_void_p)

_PROTOS = {
    "sgc_abi_version": (C.c_int, []),
    "sgc_build_flags": (C.c_int, []),
    "sgc_oneshot_cache_clear": (C.c_int, []),
    "sgc_last_error": (C.c_char_p, []),
    "sgc_device_count": (C.c_int, [intp]),
    "sgc_set_device": (C.c_int, [C.c_int]),
    "sgc_malloc": (C.c_int, [vpp, C.c_size_t]),
    "sgc_free": (C.c_int, [C.c_void_p]),
    "sgc_malloc_host": (C.c_int, [vpp, C.c_size_t]),
    "sgc_debug_cinv": (C.c_int, [dblp, dblp, C.c_int64]),
    "sgc_debug_cdiv": (C.c_int, [dblp, dblp, dblp, C.c_int64]),
    "sgc_malloc_host_wc": (C.c_int, [vpp, C.c_size_t]),
    "sgc_free_host": (C.c_int, [C.c_void_p]),
    "sgc_memcpy_h2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "sgc_memcpy_d2h": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "sgc_memset_zero": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p]),
    "sgc_stream_create": (C.c_int, [vpp]),
    "sgc_stream_destroy": (C.c_int, [C.c_void_p]),
    "sgc_stream_sync": (C.c_int, [C.c_void_p]),
    "sgc_launch_count": (C.c_int64, []),
    "sgc_op_create_partial": (C.c_int, [vpp, C.c_int, C.c_int, i64p, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                        C.c_int, dblp, dblp]),
    "sgc_op_create_mhat": (C.c_int, [vpp, C.c_int, C.c_int, i64p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                     C.c_int, C.c_int, dblp, dblp]),
    "sgc_op_create_curl": (C.c_int, [vpp, C.c_int, C.c_int, i64p, intp, vpp, C.c_int, intp, dblp, intp, C.c_int,
                                     intp, C.c_int, intp, dblp]),
    "sgc_op_create_divg": (C.c_int, [vpp, C.c_int, C.c_int, i64p, intp, vpp, C.c_int, intp, dblp, dblp]),
    "sgc_op_create_grad": (C.c_int, [vpp, C.c_int, C.c_int, i64p, intp, vpp, C.c_int, intp, dblp, dblp]),
    "sgc_op_create_mean": (C.c_int, [vpp, C.c_int, C.c_int, i64p, intp, vpp, vpp, C.c_int, intp, dblp, dblp]),
    "sgc_op_set_slab": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "sgc_op_set_slab_weights": (C.c_int, [C.c_void_p, dblp, dblp, C.c_int]),
    "sgc_op_set_tuning": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]),
    "sgc_op_apply": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "sgc_op_apply_range": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, vpp, C.c_int64, C.c_int64,
                                     C.c_void_p]),
    "sgc_op_destroy": (C.c_int, [C.c_void_p]),
    "sgc_op_launches": (C.c_int, [C.c_void_p, C.c_int]),
    "sgc_apply_partial": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, i64p, C.c_int, C.c_int,
                                    C.c_void_p, C.c_int, C.c_int, dblp, dblp, C.c_void_p]),
    "sgc_apply_mhat": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, i64p, C.c_int, C.c_int,
                                 C.c_void_p, C.c_void_p, C.c_int, C.c_int, dblp, dblp, C.c_void_p]),
    "sgc_apply_curl": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, i64p, intp, vpp, C.c_int,
                                 intp, dblp, intp, C.c_int, intp, C.c_int, intp, dblp, C.c_void_p]),
    "sgc_apply_divg": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, i64p, intp, vpp, C.c_int,
                                 intp, dblp, dblp, C.c_void_p]),
    "sgc_apply_grad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, i64p, intp, vpp, C.c_int,
                                 intp, dblp, dblp, C.c_void_p]),
    "sgc_apply_mean": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, i64p, intp, vpp, vpp,
                                 C.c_int, intp, dblp, dblp, C.c_void_p]),
    "sgc_comm_init": (C.c_int, [vpp, C.c_int, C.c_int]),
    "sgc_comm_handle": (C.c_int, [C.c_void_p, C.c_void_p]),
    "sgc_comm_connect": (C.c_int, [C.c_void_p, C.c_void_p]),
    "sgc_comm_connect_ptrs": (C.c_int, [C.c_void_p, C.c_void_p, vpp]),
    "sgc_comm_ctrl_bytes": (C.c_int, [C.c_int]),
    "sgc_comm_init_local": (C.c_int, [vpp, C.c_int]),
    "sgc_comm_rank": (C.c_int, [C.c_void_p, intp, intp]),
    "sgc_comm_destroy": (C.c_int, [C.c_void_p]),
    "sgc_peer_alloc": (C.c_int, [C.c_void_p, C.c_size_t, vpp, C.c_void_p]),
    "sgc_peer_open": (C.c_int, [C.c_void_p, C.c_void_p, vpp]),
    "sgc_peer_close": (C.c_int, [C.c_void_p, C.c_void_p]),
    "sgc_peer_free": (C.c_int, [C.c_void_p, C.c_void_p]),
    "sgc_peer_barrier": (C.c_int, [C.c_void_p, C.c_void_p]),
    "sgc_comm_epoch": (C.c_int, [C.c_void_p, i64p, C.c_void_p]),
    "sgc_comm_timeouts": (C.c_int, [C.c_void_p, i64p, C.c_void_p]),
    "sgc_op_attach_comm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "sgc_op_apply_synced": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, vpp, C.c_void_p]),
    "sgc_op_apply_curlcurl": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "sgc_op_apply_curlcurl_range": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, vpp, vpp,
                                              C.c_int64, C.c_int64, C.c_void_p]),
    "sgc_op_set_slab_coef": (C.c_int, [C.c_void_p, dblp, dblp, C.c_int]),
    "sgc_op_apply_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "sgc_op_apply_host_chunked": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "sgc_op_apply_curlcurl_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "sgc_host_set_registration": (C.c_int, [C.c_int]),
    "sgc_host_register": (C.c_int, [C.c_void_p, C.c_size_t]),
    "sgc_host_unregister": (C.c_int, [C.c_void_p]),
    "sgc_axis_create": (C.c_int, [vpp, C.c_int64, dblp, dblp, C.c_double, C.c_double, C.c_double, C.c_double, dblp]),
    "sgc_axis_destroy": (C.c_int, [C.c_void_p]),
    "sgc_axis_stretch": (C.c_int, [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sgc_op_set_axes": (C.c_int, [C.c_void_p, vpp, C.c_double, C.c_void_p]),
    "sgc_asm_create_partial": (C.c_int, [vpp, C.c_int, i64p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, dblp,
                                         C.c_int]),
    "sgc_asm_create_mhat": (C.c_int, [vpp, C.c_int, i64p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                      C.c_int, dblp, C.c_int]),
    "sgc_asm_create_curl": (C.c_int, [vpp, C.c_int, i64p, intp, vpp, C.c_int, intp, dblp, C.c_int, intp, C.c_int,
                                      intp, C.c_int, intp, C.c_int]),
    "sgc_asm_create_divg": (C.c_int, [vpp, C.c_int, i64p, intp, vpp, C.c_int, intp, dblp, C.c_int, C.c_int]),
    "sgc_asm_create_grad": (C.c_int, [vpp, C.c_int, i64p, intp, vpp, C.c_int, intp, dblp, C.c_int, C.c_int]),
    "sgc_asm_create_mean": (C.c_int, [vpp, C.c_int, i64p, intp, vpp, vpp, C.c_int, intp, dblp, C.c_int, C.c_int]),
    "sgc_asm_create_picmp": (C.c_int, [vpp, C.c_int, i64p, C.c_int, intp, C.c_void_p, C.c_int, C.c_int]),
    "sgc_asm_set_row_range": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64]),
    "sgc_asm_set_tuning": (C.c_int, [C.c_void_p, C.c_int]),
    "sgc_asm_col_partition": (C.c_int, [C.c_void_p, C.c_int, C.c_int, i64p, i64p]),
    "sgc_asm_set_col_range": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64]),
    "sgc_asm_shape": (C.c_int, [C.c_void_p, i64p, i64p, intp]),
    "sgc_asm_count": (C.c_int, [C.c_void_p, i64p, C.c_void_p]),
    "sgc_asm_fill_all": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sgc_asm_colptr": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, i64p, C.c_void_p]),
    "sgc_asm_fill": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sgc_asm_coo_sort": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, i64p,
                                   C.c_void_p]),
    "sgc_asm_nnz_host": (C.c_int, [C.c_void_p, i64p]),
    "sgc_asm_fill_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "sgc_asm_destroy": (C.c_int, [C.c_void_p]),
    "sgc_spmv_create": (C.c_int, [vpp, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "sgc_spmv_apply": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sgc_spmv_destroy": (C.c_int, [C.c_void_p]),
    "sgc_csc_matvec": (C.c_int, [C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                 C.c_void_p, C.c_void_p, C.c_void_p]),
}

for _name, (_res, _args) in _PROTOS.items():
    _f = getattr(lib, _name)  # AttributeError here means the .so is stale: rebuild
    _f.restype = _res
    _f.argtypes = _args


def check(status: int):
    if status == SGC_OK:
        return
    msg = lib.sgc_last_error().decode("utf-8", "replace")
    if status == SGC_ERR_INEXACT:
        raise InexactError(status, msg)
    if status == SGC_ERR_INVALID:
        raise SgcError(status, msg)
    raise SgcError(status, msg)
