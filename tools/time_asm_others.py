"""Fill pass of the single-column operators and block orderings at 384^3 Float64 (development tool): several cells per thread
(tuning 0) against one column per thread (tuning 8) in the warp-granular kernel."""
import sys, os, json, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sgc_b200 as S
from sgc_b200 import matrix as MX
from sgc_b200._lib import lib, check
from sgc_b200.device import stream_handle

def timeit(fn, iters=10, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters

n = int(sys.argv[1]) if len(sys.argv) > 1 else 384
Na = (n, n, n)
rng = np.random.default_rng(3)
da = [rng.uniform(0.5, 1.5, n) for _ in range(3)]
dc = [d + 1j * rng.uniform(-.2, .2, n) for d in da]
cases = (("create_partial_f64", lambda: S.create_partial(3, True, Na, da[2], True), 8),
         ("create_partial_c128", lambda: S.create_partial(1, False, Na, dc[0], True, np.exp(-0.3j)), 16),
         ("create_mhat_f64", lambda: S.create_mhat(2, False, Na, da[1], da[1], False), 8),
         ("create_grad_f64", lambda: S.create_grad([False] * 3, da, [True, True, False]), 8),
         ("create_curl_f64_block_ordering", lambda: S.create_curl([True] * 3, da, [True] * 3, order_cmpfirst=False), 8),
         ("create_mean_f64_block_ordering", lambda: S.create_mean([True, False, True], da, da, [True, True, False], order_cmpfirst=False), 8))
for name, fn, es in cases:
    A = fn()
    nnz, ncol = A.nnz, A.n
    finish = MX._finish
    try:
        MX._finish = lambda a, *args, **kw: a
        hnd = fn()
    finally:
        MX._finish = finish
    st, nn = stream_handle(None), C.c_int64()
    ref = [t.clone() for t in (A.colptr, A.rowval, A.nzval)]
    for tuning in (0, 8):
        check(lib.sgc_asm_set_tuning(hnd.h, tuning))
        check(lib.sgc_asm_count(hnd.h, C.byref(nn), st))
        fill = lambda: check(lib.sgc_asm_fill_all(hnd.h, C.c_void_p(A.colptr.data_ptr()), 1, C.c_void_p(A.rowval.data_ptr()),
                                                  C.c_void_p(A.nzval.data_ptr()), st))
        for t in (A.colptr, A.rowval, A.nzval): t.zero_()
        ms = timeit(fill)
        same = all(bool(torch.equal(x, y)) for x, y in zip(ref, (A.colptr, A.rowval, A.nzval)))
        alg = nnz * (8 + es) + (ncol + 1) * 8
        print(json.dumps(dict(op=name, n=n, tuning=tuning, nnz=nnz, ms_fill=round(ms, 4), gnnz_s=round(nnz / ms / 1e6, 1), gbs=round(alg / ms / 1e6), same=same)), flush=True)
    del A, hnd
