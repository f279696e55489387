"""Fused curl-of-curl vs two fused curls for Float64 fields at 320^3 (development timing)."""
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
n = 320; N = (n, n, n); M = n ** 3
rng = np.random.default_rng(0)
d1 = [rng.uniform(0.5, 1.5, n) for _ in range(3)]; d2 = [rng.uniform(0.5, 1.5, n) for _ in range(3)]
E = S.DeviceArray(torch.rand(3 * M, device="cuda", dtype=torch.float64), N + (3,))
H, G, Gref = (S.DeviceArray.zeros(N + (3,), np.float64) for _ in range(3))
c1 = S.Operator.curl(S.SGC_F64, N, [True] * 3, d1, [False] * 3, [1.0] * 3)
c2 = S.Operator.curl(S.SGC_F64, N, [False] * 3, d2, [False] * 3, [1.0] * 3)
def two():
    c1.apply(H, E, "="); c2.apply(Gref, H, "=")
ms2 = timeit(two)
print(json.dumps(dict(mode="two fused curls f64", ms=round(ms2, 4), gbs_96=round(96 * M / ms2 / 1e6))), flush=True)
for ty, se, march in ((7, 3, 64), (7, 4, 64), (7, 3, 32), (4, 3, 64), (4, 4, 64)):
    c2.set_tuning(32, ty, march, se << 4)
    ms = timeit(lambda: c2.apply_after(c1, G, E, "="))
    print(json.dumps(dict(ty=ty, se=se, march=march, ms=round(ms, 4), speedup=round(ms2 / ms, 3), gbs_48=round(48 * M / ms / 1e6), ok=bool(torch.equal(G.t, Gref.t)))), flush=True)
