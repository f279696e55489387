"""Fused curl-of-curl vs two fused curls at 256^3 ComplexF64 (development sweep)."""
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

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
N = (n, n, n); M = n ** 3
gen = torch.Generator(device="cuda").manual_seed(0)
def rnd():
    return S.DeviceArray(torch.complex(torch.rand(3 * M, generator=gen, device="cuda", dtype=torch.float64) - 0.5,
                                       torch.rand(3 * M, generator=gen, device="cuda", dtype=torch.float64) - 0.5), N + (3,))
E, H, G, Gref = rnd(), S.DeviceArray.zeros(N + (3,)), S.DeviceArray.zeros(N + (3,)), S.DeviceArray.zeros(N + (3,))
rng = np.random.default_rng(0)
d1 = [rng.uniform(0.5, 1.5, n) + 1j * rng.uniform(-.2, .2, n) for _ in range(3)]
d2 = [rng.uniform(0.5, 1.5, n) + 1j * rng.uniform(-.2, .2, n) for _ in range(3)]
BCS = {'pec': [[False] * 3], 'bloch': [[True] * 3], 'both': [[False] * 3, [True] * 3]}[sys.argv[3] if len(sys.argv) > 3 else 'both']
for bl in BCS:
    c1 = S.Operator.curl(S.SGC_C128, N, [True] * 3, d1, bl, [np.exp(-0.3j)] * 3)
    c2 = S.Operator.curl(S.SGC_C128, N, [False] * 3, d2, bl, [np.exp(-0.3j)] * 3)
    def two():
        c1.apply(H, E, "="); c2.apply(Gref, H, "=")
    ms2 = timeit(two)
    print(json.dumps(dict(bloch=bl[0], mode="two fused curls", ms=round(ms2, 4), gcell_pairs=round(M / ms2 / 1e6, 2), gbs_192=round(192 * M / ms2 / 1e6))), flush=True)
    # ver 4: TMA-fed warp-specialised kernel (se = E stages | H stages << 4), ver 2: third generation (warp roles, cp.async)
    cfgs = [(0, 7, 4, 5, m) for m in (0, 32, 128)] + [(0, 7, 3, 5, 0), (0, 5, 4, 5, 0), (0, 5, 4, 5, 32)] + [(0, 6, 0, 4, 0), (32, 7, 3, 2, 64)]
    if len(sys.argv) > 2 and sys.argv[2] == "v6":  # ver 6: two planes per barrier interval (se = E pair stages)
        # se low nibble = E pair stages, high nibble: 0 = default (one block, outputs first), 1 = two-phase code only, 2 = one block, H points first
        cfgs = [(0, 5, 3, 6, 0), (0, 5, 3 + 32, 6, 0), (0, 5, 3, 6, 32), (0, 5, 3 + 32, 6, 32), (0, 5, 3 + 16, 6, 32), (0, 5, 4, 5, 32)]
    if len(sys.argv) > 2 and sys.argv[2] == "v6march":
        cfgs = [(0, 5, 3, 6, m) for m in (20, 22, 24, 26, 28, 32, 38, 44, 52, 64)]
    if len(sys.argv) > 2 and sys.argv[2] == "quick":
        cfgs = [(0, 6, 0, 4, 0), (32, 7, 3, 2, 64)]
    for tx, ty, se, ver, march in cfgs:
        for _ in (0,):
            c2.set_tuning(tx, ty, march, (se << 4) | ver)
            try:
                ms = timeit(lambda: c2.apply_after(c1, G, E, "="))
                ok = bool(torch.equal(torch.view_as_real(G.t), torch.view_as_real(Gref.t)))
                print(json.dumps(dict(bloch=bl[0], ver=ver, tx=tx, ty=ty, se=se, march=march, ms=round(ms, 4), speedup=round(ms2 / ms, 3),
                                      gbs_96=round(96 * M / ms / 1e6), ok=ok, ndiff=int((torch.view_as_real(G.t) != torch.view_as_real(Gref.t)).sum()))), flush=True)
            except Exception as ex:
                print(json.dumps(dict(tx=tx, ty=ty, se=se, march=march, err=str(ex)[:100])), flush=True)
    c2.set_tuning(0, 0, 0, 0)
