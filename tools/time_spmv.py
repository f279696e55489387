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
    for _ in range(2): y = A.matvec(x)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): y = A.matvec(x)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    es = 16 if cplx else 8
    byt = A.nnz * (8 + es) + (3 * M + 1) * 8 + 2 * 3 * M * es
    print(json.dumps(dict(cplx=cplx, nnz=A.nnz, ms=round(ms, 3), gnnz_s=round(A.nnz / ms / 1e6, 1), gbs=round(byt / ms / 1e6))), flush=True)
    del A, x, y
