"""Float64 / real-coefficient variants of the fused 3-D curl at 320^3 (development timing)."""
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
d = [rng.uniform(0.5, 1.5, n) for _ in range(3)]
F = S.DeviceArray(torch.rand(3 * M, device="cuda", dtype=torch.float64), N + (3,))
G = S.DeviceArray.zeros(N + (3,), np.float64)
o = S.Operator.curl(S.SGC_F64, N, [True] * 3, d, [False] * 3, [1.0] * 3)
for cfg in ((0, 0, 0, 0), (32, 4, 12, (6 << 4) | (4 << 8)), (32, 8, 12, (4 << 4) | (3 << 8)), (64, 4, 12, (4 << 4) | (3 << 8)), (32, 8, 12, (6 << 4) | (2 << 8)), (64, 8, 12, (3 << 4) | (2 << 8)), (32, 16, 12, (3 << 4) | (2 << 8))):
    try:
        o.set_tuning(*cfg)
        ms = timeit(lambda: o.apply(G, F, "="))
        print(json.dumps(dict(dtype="f64", cfg=cfg, ms=round(ms, 4), gbs=round(48 * M / ms / 1e6))), flush=True)
    except Exception as ex:
        print(json.dumps(dict(cfg=cfg, err=str(ex)[:80])), flush=True)
g = S.DeviceArray.zeros(N, np.float64)
dv = S.Operator.divg(S.SGC_F64, N, [True] * 3, d, [False] * 3, [1.0] * 3)
gr = S.Operator.grad(S.SGC_F64, N, [False] * 3, d, [False] * 3, [1.0] * 3)
for name, fn in (("divg f64", lambda: dv.apply(g, F, "=")), ("grad f64", lambda: gr.apply(G, g, "="))):
    ms = timeit(fn); print(json.dumps(dict(op=name, ms=round(ms, 4), gbs=round(32 * M / ms / 1e6))), flush=True)
dv.set_tuning(0, 0, 0, 6); gr.set_tuning(0, 0, 0, 6)
for name, fn in (("divg f64 row-register", lambda: dv.apply(g, F, "=")), ("grad f64 row-register", lambda: gr.apply(G, g, "="))):
    ms = timeit(fn); print(json.dumps(dict(op=name, ms=round(ms, 4), gbs=round(32 * M / ms / 1e6))), flush=True)
