"""ncu driver: a few launches of the fused curl-of-curl at 256^3 (PML-like complex coefficients, PEC box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sgc_b200 as S
n = 256
N = (n, n, n); M = n ** 3
tx, ty, march, se = [int(a) for a in (sys.argv[1:5] if len(sys.argv) >= 5 else "0 0 0 0".split())]
gen = torch.Generator(device="cuda").manual_seed(0)
E = S.DeviceArray(torch.complex(torch.rand(3 * M, generator=gen, device="cuda", dtype=torch.float64) - 0.5,
                                torch.rand(3 * M, generator=gen, device="cuda", dtype=torch.float64) - 0.5), N + (3,))
G = S.DeviceArray.zeros(N + (3,))
rng = np.random.default_rng(0)
d1 = [rng.uniform(0.5, 1.5, n) + 1j * rng.uniform(-.2, .2, n) for _ in range(3)]
d2 = [rng.uniform(0.5, 1.5, n) + 1j * rng.uniform(-.2, .2, n) for _ in range(3)]
c1 = S.Operator.curl(S.SGC_C128, N, [True] * 3, d1, [False] * 3, [1.0] * 3)
c2 = S.Operator.curl(S.SGC_C128, N, [False] * 3, d2, [False] * 3, [1.0] * 3)
c2.set_tuning(tx, ty, march, (se << 4) | (int(sys.argv[5]) if len(sys.argv) > 5 else 0))  # se may carry the H-ring stages in its high nibble
for _ in range(4):
    c2.apply_after(c1, G, E, "=")
torch.cuda.synchronize()
print("done", S.launch_count())
