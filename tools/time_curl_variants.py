"""Fused 3-D curl at 256^3: accumulate mode, real coefficients, Bloch (development timing)."""
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
n = 256; N = (n, n, n); M = n ** 3
rng = np.random.default_rng(0)
gen = torch.Generator(device="cuda").manual_seed(0)
F = S.DeviceArray(torch.complex(torch.rand(3 * M, generator=gen, device="cuda", dtype=torch.float64), torch.rand(3 * M, generator=gen, device="cuda", dtype=torch.float64)), N + (3,))
G = S.DeviceArray.zeros(N + (3,))
dc = [rng.uniform(0.5, 1.5, n) + 1j * rng.uniform(-.2, .2, n) for _ in range(3)]
dr = [rng.uniform(0.5, 1.5, n) for _ in range(3)]
for name, d, bl, e in (("complex coef, PEC", dc, [False] * 3, [1.0] * 3), ("real coef, PEC", dr, [False] * 3, [1.0] * 3),
                       ("complex coef, Bloch", dc, [True] * 3, [np.exp(-0.3j)] * 3)):
    for fwd in (True, False):
        o = S.Operator.curl(S.SGC_C128, N, [fwd] * 3, d, bl, e)
        for mode, B in (("=", 96), ("+=", 144)):
            ms = timeit(lambda: o.apply(G, F, mode))
            print(json.dumps(dict(case=name, fwd=fwd, mode=mode, ms=round(ms, 4), gbs=round(B * M / ms / 1e6))), flush=True)
