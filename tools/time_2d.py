"""BASELINE config 3: 2-D 8192^2 Float64 ∂ / m̂ / TE-TM sub-curls with PEC/PMC BCs (development timing)."""
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

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
N = (n, n); M = n * n
rng = np.random.default_rng(0)
F = S.DeviceArray(torch.rand(M, device="cuda", dtype=torch.float64), N)
G = S.DeviceArray.zeros(N, np.float64)
d = rng.uniform(0.5, 1.5, n)
for nw in (1, 2):
    for fwd in (True, False):
        for op, B in (("=", 16), ("+=", 24)):
            o = S.Operator.partial(S.SGC_F64, N, nw, fwd, d, False)
            ms = timeit(lambda: o.apply(G, F, op))
            print(json.dumps(dict(op="partial", nw=nw, fwd=fwd, mode=op, ms=round(ms, 4), gcell=round(M / ms / 1e6, 1), gbs=round(B * M / ms / 1e6))), flush=True)
        o = S.Operator.mhat(S.SGC_F64, N, nw, fwd, None, None, False)
        ms = timeit(lambda: o.apply(G, F, "="))
        print(json.dumps(dict(op="mhat_arith", nw=nw, fwd=fwd, ms=round(ms, 4), gbs=round(16 * M / ms / 1e6))), flush=True)
        o = S.Operator.mhat(S.SGC_F64, N, nw, fwd, d, d, False)
        ms = timeit(lambda: o.apply(G, F, "="))
        print(json.dumps(dict(op="mhat_weighted", nw=nw, fwd=fwd, ms=round(ms, 4), gbs=round(16 * M / ms / 1e6))), flush=True)
# TE: Hz = curl of (Ex,Ey): cmp_out=[3], cmp_in=[1,2]; TM: (Ex,Ey) = curl of Hz
F2 = S.DeviceArray(torch.rand(2 * M, device="cuda", dtype=torch.float64), N + (2,))
G1 = S.DeviceArray.zeros(N + (1,), np.float64)
o = S.Operator.curl(S.SGC_F64, N, [True, True], [d, d], [False, False], [1.0, 1.0], cmp_shp=[1, 2], cmp_out=[3], cmp_in=[1, 2])
ms = timeit(lambda: o.apply(G1, F2, "="))
print(json.dumps(dict(op="TE 1x2 curl", launches=o.launches("="), ms=round(ms, 4), gbs_alg24=round(24 * M / ms / 1e6))), flush=True)
o = S.Operator.curl(S.SGC_F64, N, [False, False], [d, d], [False, False], [1.0, 1.0], cmp_shp=[1, 2], cmp_out=[1, 2], cmp_in=[3])
ms = timeit(lambda: o.apply(F2, G1, "="))
print(json.dumps(dict(op="TM 2x1 curl", launches=o.launches("="), ms=round(ms, 4), gbs_alg24=round(24 * M / ms / 1e6))), flush=True)
a, b = F.t, G.t
ms = timeit(lambda: b.copy_(a))
print(json.dumps(dict(copy_gbs=round(16 * M / ms / 1e6))))
