"""Fused 3-D curl at 256^3 ComplexF64 (PML-type coefficients, PEC box): x-neighbours from the shared tile (default) against by
warp shuffle (variant bit 12), `=`; alternating order, several repetitions (the later ones in the sustained, power-capped state)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sgc_b200 as S

def timeit(fn, iters=40, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters

n = 256
N = (n, n, n); M = n ** 3
E = S.DeviceArray(torch.view_as_complex(torch.rand(3 * M, 2, device="cuda", dtype=torch.float64)), N + (3,))
H, G = S.DeviceArray.zeros(N + (3,)), S.DeviceArray.zeros(N + (3,))
rng = np.random.default_rng(0)
d1 = [rng.uniform(0.5, 1.5, n) + 1j * rng.uniform(-.2, .2, n) for _ in range(3)]
d2 = [rng.uniform(0.5, 1.5, n) + 1j * rng.uniform(-.2, .2, n) for _ in range(3)]
c1 = S.Operator.curl(S.SGC_C128, N, [True] * 3, d1, [False] * 3, [1.0] * 3)
c2 = S.Operator.curl(S.SGC_C128, N, [False] * 3, d2, [False] * 3, [1.0] * 3)
def step():
    c1.apply(H, E, "="); c2.apply(G, H, "=")
for rep in range(4):
    for name, var in (("shuffle", 1 << 12), ("shared", 0)):
        c1.set_tuning(0, 0, 0, var); c2.set_tuning(0, 0, 0, var)
        ms = timeit(step)
        print(json.dumps(dict(rep=rep, x_neighbour=name, ms_two_curls=round(ms, 4), gbs=round(192 * M / ms / 1e6))), flush=True)
