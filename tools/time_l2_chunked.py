"""Hypothesis test: curl-of-curl as two apply_curl! per z-chunk, so that the intermediate field is re-read from L2
instead of HBM (full-size H: its dirty lines are still written back once).  CUDA-graph replay of the chunk sequence."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sgc_b200 as S

n = 256
N = (n, n, n); M = n ** 3
gen = torch.Generator(device="cuda").manual_seed(0)
E = S.DeviceArray(torch.complex(torch.rand(3 * M, generator=gen, device="cuda", dtype=torch.float64) - 0.5,
                                torch.rand(3 * M, generator=gen, device="cuda", dtype=torch.float64) - 0.5), N + (3,))
H, G, Gref = (S.DeviceArray.zeros(N + (3,)) for _ in range(3))
rng = np.random.default_rng(0)
d1 = [rng.uniform(0.5, 1.5, n) + 1j * rng.uniform(-.2, .2, n) for _ in range(3)]
d2 = [rng.uniform(0.5, 1.5, n) + 1j * rng.uniform(-.2, .2, n) for _ in range(3)]
c1 = S.Operator.curl(S.SGC_C128, N, [True] * 3, d1, [False] * 3, [1.0] * 3)
c2 = S.Operator.curl(S.SGC_C128, N, [False] * 3, d2, [False] * 3, [1.0] * 3)

def timeit(fn, iters=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters

c1.apply(H, E, "="); c2.apply(Gref, H, "=")
ms2 = timeit(lambda: (c1.apply(H, E, "="), c2.apply(Gref, H, "=")))
print(json.dumps(dict(mode="two whole-grid curls", ms=round(ms2, 4))), flush=True)
for chunk in (4, 8, 12, 16, 24, 32, 64):
    for march in (0, 4, 8):
        if march > chunk: continue
        c1.set_tuning(0, 0, march, 0); c2.set_tuning(0, 0, march, 0)
        def seq():
            for a in range(0, n, chunk):
                b = min(a + chunk, n)
                c1.apply_range(H, E, "=", a, b)                       # H[a,b) needs E up to plane b (resident)
                c2.apply_range(G, H, "=", a, b)                       # E'[k] needs H[k-1], H[k]: H[a-1] is from the previous chunk
        G.t.zero_()
        side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            seq(); torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side, capture_error_mode="thread_local"):
                seq()
        torch.cuda.current_stream().wait_stream(side)
        ms = timeit(g.replay)
        ok = bool(torch.equal(torch.view_as_real(G.t), torch.view_as_real(Gref.t)))
        print(json.dumps(dict(mode="chunked, graph", chunk=chunk, march=march or 12, launches=2 * ((n + chunk - 1) // chunk), ms=round(ms, 4),
                              speedup=round(ms2 / ms, 3), ok=ok)), flush=True)
