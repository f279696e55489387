"""Small driver for ncu: a few launches of the fused curl at 256^3 with PML-like complex coefficients."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sgc_b200 as S
n = 256
N = (n, n, n); M = n ** 3
tx, ty, march, st, mb, kern = [int(a) for a in (sys.argv[1:7] if len(sys.argv) >= 7 else "32 4 16 4 4 0".split())]
gen = torch.Generator(device="cuda").manual_seed(0)
F = S.DeviceArray(torch.complex(torch.rand(3 * M, generator=gen, device="cuda", dtype=torch.float64) - 0.5,
                                torch.rand(3 * M, generator=gen, device="cuda", dtype=torch.float64) - 0.5), N + (3,))
G = S.DeviceArray.zeros(N + (3,))
rng = np.random.default_rng(0)
dl = [rng.uniform(0.5, 1.5, n) + 1j * rng.uniform(-.2, .2, n) for _ in range(3)]
o = S.Operator.curl(S.SGC_C128, N, [True] * 3, dl, [False] * 3, [1.0] * 3)
o.set_tuning(tx, ty, march, kern | (st << 4) | (mb << 8))
for _ in range(5):
    o.apply(G, F, "=")
torch.cuda.synchronize()
print("done", S.launch_count())
