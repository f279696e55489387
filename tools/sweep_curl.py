"""Tile / stage / z-march sweep of the fused curl kernels (development tool)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import sgc_b200 as S

def time_op(o, G, F, op, iters=10, warm=3):
    for _ in range(warm):
        o.apply(G, F, op)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(iters):
        o.apply(G, F, op)
    t1.record()
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) / iters

V2 = [(32, 4, 4, 4), (32, 4, 5, 4), (32, 4, 6, 4), (32, 4, 7, 4), (32, 4, 8, 3), (32, 4, 3, 4), (32, 8, 3, 3), (32, 8, 4, 3), (32, 8, 4, 2), (32, 8, 6, 2), (32, 16, 3, 2), (32, 16, 4, 1),
      (64, 4, 4, 3), (64, 4, 4, 2), (64, 8, 3, 2)]

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    marches = [int(a) for a in sys.argv[2].split(",")] if len(sys.argv) > 2 else [16, 32, 64, 0]
    N = (n, n, n)
    M = n ** 3
    gen = torch.Generator(device="cuda").manual_seed(0)
    F = S.DeviceArray(torch.complex(torch.rand(3 * M, generator=gen, device="cuda", dtype=torch.float64) - 0.5,
                                    torch.rand(3 * M, generator=gen, device="cuda", dtype=torch.float64) - 0.5), N + (3,))
    G = S.DeviceArray.zeros(N + (3,))
    Gref = S.DeviceArray.zeros(N + (3,))
    rng = np.random.default_rng(0)
    for cc, fw, bl in ((True, True, False), (True, False, True), (False, True, True)):
        dl = [rng.uniform(0.5, 1.5, n) + (1j * rng.uniform(-.2, .2, n) if cc else 0) for _ in range(3)]
        o = S.Operator.curl(S.SGC_C128, N, [fw] * 3, dl, [bl] * 3, [np.exp(-0.1j)] * 3)
        o.set_tuning(0, 0, 0, 2)
        o.apply(Gref, F, "=")
        best = None
        for (tx, ty, st, mb) in V2:
          for xs in (0, 1):
            for march in marches:
                o.set_tuning(tx, ty, march, 0 | (st << 4) | (mb << 8) | (xs << 12))
                G.t.zero_()
                ms = time_op(o, G, F, "=")
                ok = bool(torch.equal(G.t, Gref.t))
                r = dict(cc=cc, fwd=fw, bloch=bl, k="v2", tile=(tx, ty), stages=st, minb=mb, xsm=xs, march=march, ms=round(ms, 4),
                         gbs=round(96 * M / ms / 1e6), ok=ok)
                print(json.dumps(r), flush=True)
                if ok and (best is None or ms < best["ms"]):
                    best = r
        for tile in ((32, 8), (64, 4)):
            o.set_tuning(tile[0], tile[1], 32, 1)
            ms = time_op(o, G, F, "=")
            print(json.dumps(dict(cc=cc, k="v1", tile=tile, march=32, ms=round(ms, 4), gbs=round(96 * M / ms / 1e6),
                                  ok=bool(torch.equal(G.t, Gref.t)))), flush=True)
        print("BEST", json.dumps(best), flush=True)
    a, b = F.t, G.t
    for _ in range(3): b.copy_(a)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(10): b.copy_(a)
    t1.record(); torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / 10
    print(json.dumps(dict(copy_gbs=2 * a.numel() * 16 / ms / 1e6)))

if __name__ == "__main__":
    main()
