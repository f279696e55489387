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
n = 8192; N = (n, n); M = n * n
rng = np.random.default_rng(0)
F = S.DeviceArray(torch.rand(M, device="cuda", dtype=torch.float64), N)
G = S.DeviceArray.zeros(N, np.float64)
d = rng.uniform(0.5, 1.5, n)
for variant in (0, 4, 5):
    for nw in (1, 2):
        for name, o, mode, B in (("partial=", S.Operator.partial(S.SGC_F64, N, nw, True, d, False), "=", 16),
                                 ("partial+=", S.Operator.partial(S.SGC_F64, N, nw, True, d, False), "+=", 24),
                                 ("mhat_arith", S.Operator.mhat(S.SGC_F64, N, nw, True, None, None, False), "=", 16),
                                 ("mhat_weighted", S.Operator.mhat(S.SGC_F64, N, nw, False, d, d, False), "=", 16)):
            o.set_tuning(0, 0, 0, variant)
            ms = timeit(lambda: o.apply(G, F, mode))
            print(json.dumps(dict(variant=variant, nw=nw, op=name, ms=round(ms, 4), gbs=round(B * M / ms / 1e6))), flush=True)
