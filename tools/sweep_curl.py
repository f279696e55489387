"""Tile / z-march sweep of the fused curl kernel (development tool; prints one line per config)."""
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

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    N = (n, n, n)
    M = n ** 3
    gen = torch.Generator(device="cuda").manual_seed(0)
    F = S.DeviceArray(torch.complex(torch.rand(3 * M, generator=gen, device="cuda", dtype=torch.float64) - 0.5,
                                    torch.rand(3 * M, generator=gen, device="cuda", dtype=torch.float64) - 0.5), N + (3,))
    G = S.DeviceArray.zeros(N + (3,))
    rng = np.random.default_rng(0)
    res = []
    for cc in (False, True):
        dl = [rng.uniform(0.5, 1.5, n) + (1j * rng.uniform(-.2, .2, n) if cc else 0) for _ in range(3)]
        o = S.Operator.curl(S.SGC_C128, N, [True] * 3, dl, [True] * 3, [np.exp(-0.1j)] * 3)
        for tile in ((32, 4), (32, 8), (32, 16), (64, 4), (64, 8)):
            for march in (8, 16, 32, 64, 0):
                o.set_tuning(tile[0], tile[1], march, 0)
                ms = time_op(o, G, F, "=")
                gbs = 96 * M / ms / 1e6
                res.append(dict(cc=cc, tile=tile, march=march, ms=ms, gbs=gbs, gcell=M / ms / 1e6))
                print(json.dumps(res[-1]), flush=True)
        o.set_tuning(0, 0, 0, 2)
        ms = time_op(o, G, F, "=")
        print(json.dumps(dict(cc=cc, variant="unfused-7-launches", ms=ms, gcell=M / ms / 1e6)), flush=True)
    # copy bandwidth reference on this box: G <- F
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
