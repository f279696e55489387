"""Development sweep of apply_∂! along y on 8192^2 Float64 (k_fused, pattern ONE_Y), `=` and `+=`
(needs a library built with `make EXTRA=-DSGC_FUSED_SWEEP`)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sgc_b200 as S

def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters

cfgs = [(32, 4, 1), (32, 4, 2), (32, 2, 2), (32, 2, 3), (32, 2, 4), (16, 4, 2), (16, 4, 3), (16, 4, 4), (16, 2, 3), (16, 2, 4), (32, 1, 3), (32, 1, 4), (16, 1, 4)]
n = 8192
N, M = (n, n), n * n
rng = np.random.default_rng(0)
d = rng.uniform(0.5, 1.5, n)
F = S.DeviceArray(torch.rand(M, device="cuda", dtype=torch.float64), N)
G = S.DeviceArray.zeros(N, np.float64)
for name, fwd, op, B in (("partial_nw2_fwd_set", True, "=", 16), ("partial_nw2_bwd_add", False, "+=", 24), ("partial_nw2_fwd_add", True, "+=", 24)):
    o = S.Operator.partial(S.SGC_F64, N, 2, fwd, d, False)
    for vb, r, mb in cfgs + [(0, 0, 0)]:
        o.set_tuning(vb, r, mb, 0)
        try:
            ms = timeit(lambda: o.apply(G, F, op))
            print(json.dumps(dict(op=name, vecbytes=vb, rows=r, minb=mb, ms=round(ms, 4), gbs=round(B * M / ms / 1e6))), flush=True)
        except Exception as ex:
            print(json.dumps(dict(op=name, vecbytes=vb, rows=r, minb=mb, err=str(ex)[:80])), flush=True)
