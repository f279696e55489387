"""Development sweep of the one-launch composition kernel (needs a library built with
`make EXTRA=-DSGC_FUSED_SWEEP`): TE / TM sub-curls at 8192^2 Float64 and 4096^2... ComplexF64,
3-D divergence / gradient at 256^3."""
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

cfgs = [(32, 4, 1), (32, 4, 2), (32, 2, 2), (32, 2, 3), (32, 2, 4), (16, 4, 2), (16, 4, 3), (16, 4, 4), (16, 2, 3), (16, 2, 4), (32, 1, 3), (32, 1, 4), (16, 1, 4)]
rng = np.random.default_rng(0)

def field(shape, cplx):
    n = int(np.prod(shape))
    t = torch.rand(n, device="cuda", dtype=torch.float64)
    if cplx:
        t = torch.complex(t, torch.rand(n, device="cuda", dtype=torch.float64))
    return S.DeviceArray(t, shape)

def run(tag, cplx, N, mk_ops):
    es = 16 if cplx else 8
    M = int(np.prod(N))
    for name, o, gshape, fshape, units in mk_ops:
        F = field(fshape, cplx); G = field(gshape, cplx); ref = field(gshape, cplx)
        o.set_tuning(0, 0, 0, 2); o.apply(ref, F, "=")
        for vb, r, mb in cfgs:
            o.set_tuning(vb, r, mb, 0)
            row = dict(case=tag, op=name, vecbytes=vb, rows=r, minb=mb)
            try:
                ms = timeit(lambda: o.apply(G, F, "="))
                row.update(ms=round(ms, 4), gbs=round(units * es * M / ms / 1e6), ok=bool(torch.equal(torch.view_as_real(G.t) if cplx else G.t, torch.view_as_real(ref.t) if cplx else ref.t)))
            except Exception as ex:
                row["err"] = str(ex)[:80]
            print(json.dumps(row), flush=True)
        o.set_tuning(0, 0, 0, 0)
        ms = timeit(lambda: o.apply(G, F, "="))
        print(json.dumps(dict(case=tag, op=name, default=True, ms=round(ms, 4), gbs=round(units * es * M / ms / 1e6))), flush=True)

for cplx, n in ((False, 8192), (True, 4096)):
    N = (n, n)
    dt = S.SGC_C128 if cplx else S.SGC_F64
    d = rng.uniform(0.5, 1.5, n) + (1j * rng.uniform(-.2, .2, n) if cplx else 0)
    te = S.Operator.curl(dt, N, [True, True], [d, d], [False, False], [1.0, 1.0], cmp_shp=[1, 2], cmp_out=[3], cmp_in=[1, 2])
    tm = S.Operator.curl(dt, N, [False, False], [d, d], [False, False], [1.0, 1.0], cmp_shp=[1, 2], cmp_out=[1, 2], cmp_in=[3])
    run(f"2d_{'c128' if cplx else 'f64'}_{n}", cplx, N, [("TE", te, N + (1,), N + (2,), 3), ("TM", tm, N + (2,), N + (1,), 3)])
for cplx, n in ((True, 256), (False, 320)):
    N = (n, n, n)
    dt = S.SGC_C128 if cplx else S.SGC_F64
    d = rng.uniform(0.5, 1.5, n) + (1j * rng.uniform(-.2, .2, n) if cplx else 0)
    dv = S.Operator.divg(dt, N, [True] * 3, [d] * 3, [False] * 3, [1.0] * 3)
    gr = S.Operator.grad(dt, N, [False] * 3, [d] * 3, [False] * 3, [1.0] * 3)
    run(f"3d_{'c128' if cplx else 'f64'}_{n}", cplx, N, [("divg", dv, N, N + (3,), 4), ("grad", gr, N + (3,), N, 4)])
