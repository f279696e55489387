"""2-D 8192^2 Float64 single-term operators with the vector width forced (variant 4: 16-byte, 5: 32-byte) against the default."""
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

n = 8192
MARCH = int(sys.argv[1]) if len(sys.argv) > 1 else 0  # x-axis kernel: CTAs per SM (0 = default)
N = (n, n); M = n * n
rng = np.random.default_rng(0)
F = S.DeviceArray(torch.rand(M, device="cuda", dtype=torch.float64), N)
G = S.DeviceArray.zeros(N, np.float64)
d = rng.uniform(0.5, 1.5, n)
for nw in (1, 2):
    for name, mk in (("partial", lambda fwd: S.Operator.partial(S.SGC_F64, N, nw, fwd, d, False)),
                     ("mhat_arith", lambda fwd: S.Operator.mhat(S.SGC_F64, N, nw, fwd, None, None, False)),
                     ("mhat_weighted", lambda fwd: S.Operator.mhat(S.SGC_F64, N, nw, fwd, d, d, False))):
        for fwd in (True, False):
            o = mk(fwd)
            for op, B in (("=", 16), ("+=", 24)):
                row = dict(op=name, nw=nw, fwd=fwd, mode=op)
                for v in (0, 4, 5):
                    o.set_tuning(0, 0, MARCH, v)
                    ms = timeit(lambda: o.apply(G, F, op))
                    row[f"v{v}_ms"] = round(ms, 4); row[f"v{v}_gbs"] = round(B * M / ms / 1e6)
                print(json.dumps(row), flush=True)
