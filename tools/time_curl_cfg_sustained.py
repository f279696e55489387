"""Headline step (two fused curls, 256^3 ComplexF64, PEC box) for the tile / ring configurations the product library carries, in the
sustained (power-capped) state: the configurations alternate, several repetitions of 40 steps each."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sgc_b200 as S

def timeit(fn, iters=40, warm=3):
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
for _ in range(300): step()  # into the sustained state
cfgs = [(0, 0, 0, 0), (32, 4, 12, (4 << 4)), (32, 8, 12, 0), (64, 4, 12, 0), (0, 0, 24, 0), (0, 0, 8, 0), (32, 8, 24, 0)]
for rep in range(3):
    for tx, ty, march, var in cfgs:
        c1.set_tuning(tx, ty, march, var); c2.set_tuning(tx, ty, march, var)
        ms = timeit(step)
        print(json.dumps(dict(rep=rep, tile=(tx, ty), march=march, variant=var, ms_two_curls=round(ms, 4), gbs=round(192 * M / ms / 1e6))), flush=True)
