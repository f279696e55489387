"""apply_mean! at 256^3 ComplexF64: one launch of the staged z-march against the reference's three sweeps (development tool)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sgc_b200 as S

def timeit(fn, iters=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
N = (n, n, n); M = n ** 3
F = S.DeviceArray(torch.view_as_complex(torch.rand(3 * M, 2, device="cuda", dtype=torch.float64)), N + (3,))
G, G2 = S.DeviceArray.zeros(N + (3,)), S.DeviceArray.zeros(N + (3,))
rng = np.random.default_rng(3)
d3 = [rng.uniform(0.5, 1.5, m) + 1j * rng.uniform(-.2, .2, m) for m in N]
for name, mk in (("arith_fwd_sym", lambda: S.Operator.mean(S.SGC_C128, N, [True] * 3, None, None, [False] * 3, [1.0] * 3)),
                 ("weighted_bwd_sym", lambda: S.Operator.mean(S.SGC_C128, N, [False] * 3, d3, d3, [False] * 3, [1.0] * 3)),
                 ("weighted_fwd_bloch", lambda: S.Operator.mean(S.SGC_C128, N, [True] * 3, d3, d3, [True] * 3, [np.exp(-0.3j)] * 3))):
    o = mk()
    for op in ("=", "+="):
        o.set_tuning(0, 0, 0, 0)
        ms1 = timeit(lambda: o.apply(G, F, op))
        G.t.zero_(); G2.t.zero_()
        o.apply(G, F, op)
        o.set_tuning(0, 0, 0, 2)
        ms3 = timeit(lambda: o.apply(G2, F, op))
        G2.t.zero_()
        o.apply(G2, F, op)
        same = bool(torch.equal(torch.view_as_real(G.t), torch.view_as_real(G2.t)))
        B = 96 if op == "=" else 144
        print(json.dumps(dict(case=name, op=op, launches=o.launches(op), ms_one_launch=round(ms1, 4), ms_sweeps=round(ms3, 4), gbs=round(B * M / ms1 / 1e6),
                              gbs_sweeps=round(B * M / ms3 / 1e6), bit_identical=same)), flush=True)
