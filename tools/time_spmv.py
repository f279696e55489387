"""y = A x with the assembled 256^3 curl (CSC, one thread per column, atomics): development timing."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sgc_b200 as S
N = (256, 256, 256); M = 256 ** 3
for cplx in (False, True):
    e = [np.exp(-0.1j)] * 3 if cplx else [1.0] * 3
    A = S.create_curl([True] * 3, None, [True] * 3, e, N=N)
    x = torch.rand(3 * M, device="cuda", dtype=torch.float64)
    if cplx:
        x = torch.complex(x, torch.rand(3 * M, device="cuda", dtype=torch.float64))
    es = 16 if cplx else 8
    y = torch.empty(3 * M, dtype=x.dtype, device="cuda")
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record(); A.mul(x, y); t1.record(); torch.cuda.synchronize()
    plan_ms = t0.elapsed_time(t1)
    for name, fn, byt in (("column-parallel, atomics (sgc_csc_matvec)", lambda: A.matvec(x), A.nnz * (8 + es) + (3 * M + 1) * 8 + 2 * 3 * M * es),
                          ("row-sorted plan, gather (sgc_spmv_apply)", lambda: A.mul(x, y), A.nnz * (8 + es) + (3 * M + 1) * 8 + 2 * 3 * M * es)):
        for _ in range(2): fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5): fn()
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        print(json.dumps(dict(cplx=cplx, form=name, nnz=A.nnz, ms=round(ms, 3), gnnz_s=round(A.nnz / ms / 1e6, 1), gbs=round(byt / ms / 1e6),
                              frac_of_6554=round(byt / ms / 1e6 / 6553.9, 3), plan_and_first_product_ms=round(plan_ms, 1))), flush=True)
    same = bool(torch.allclose(A.matvec(x), y, rtol=1e-12, atol=1e-12))
    print(json.dumps(dict(cplx=cplx, forms_agree=same)), flush=True)
    del A, x, y
