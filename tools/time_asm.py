"""Phase timing of the CSC assembly (development tool)."""
import sys, os, json, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sgc_b200 as S
from sgc_b200._lib import lib, check
from sgc_b200.device import ints, int64s, scalars2, stream_handle

def run(n, cplx, isbloch, cmpfirst=1, tuning=0):
    N = (n, n, n); M = n ** 3
    h = C.c_void_p()
    e = [np.exp(-0.1j)] * 3 if cplx else [1.0] * 3
    check(lib.sgc_asm_create_curl(C.byref(h), 3, int64s(N), ints([1, 1, 1]), None, 0, ints(isbloch), scalars2(e), int(cplx),
                                  ints([1, 2, 3]), 3, ints([1, 2, 3]), 3, ints([1, 2, 3]), cmpfirst))
    check(lib.sgc_asm_set_tuning(h, tuning))  # 2: interior cells through the general entry path
    colptr = torch.empty(3 * M + 1, dtype=torch.int64, device="cuda")
    nnz = C.c_int64()
    st = stream_handle()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    check(lib.sgc_asm_count(h, C.byref(nnz), st))  # warm
    rowval = torch.empty(nnz.value, dtype=torch.int64, device="cuda")
    nzval = torch.empty(nnz.value, dtype=torch.complex128 if cplx else torch.float64, device="cuda")
    check(lib.sgc_asm_fill_all(h, C.c_void_p(colptr.data_ptr()), 1, C.c_void_p(rowval.data_ptr()), C.c_void_p(nzval.data_ptr()), st))
    torch.cuda.synchronize()
    reps = 5
    ev[0].record()
    for _ in range(reps):
        check(lib.sgc_asm_set_row_range(h, 0, 3 * M))  # invalidates the cached count
        check(lib.sgc_asm_count(h, C.byref(nnz), st))
    ev[1].record()
    for _ in range(reps):
        check(lib.sgc_asm_fill_all(h, C.c_void_p(colptr.data_ptr()), 1, C.c_void_p(rowval.data_ptr()), C.c_void_p(nzval.data_ptr()), st))
    ev[2].record()
    torch.cuda.synchronize()
    t_col, t_fill = ev[0].elapsed_time(ev[1]) / reps, ev[1].elapsed_time(ev[2]) / reps
    es = 16 if cplx else 8
    alg = nnz.value * (8 + es) + (3 * M + 1) * 8
    print(json.dumps(dict(n=n, tuning=tuning, cplx=cplx, bloch=isbloch, cmpfirst=cmpfirst, nnz=nnz.value, count_ms=round(t_col, 3), fill_all_ms=round(t_fill, 3),
                          total_ms=round(t_col + t_fill, 3), gnnz_s=round(nnz.value / (t_col + t_fill) / 1e6, 1),
                          alg_gbs=round(alg / (t_col + t_fill) / 1e6), fill_gbs=round(alg / t_fill / 1e6))), flush=True)
    lib.sgc_asm_destroy(h)

if __name__ == "__main__":
    for n in ([int(a) for a in sys.argv[1:]] or [256, 512]):
        for tuning in (0, 8):
            run(n, False, [1, 1, 1], tuning=tuning)
            run(n, True, [1, 1, 1], tuning=tuning)
            run(n, False, [1, 0, 0], tuning=tuning)
            run(n, False, [1, 1, 1], cmpfirst=0, tuning=tuning)
            run(n, True, [1, 1, 1], cmpfirst=0, tuning=tuning)
