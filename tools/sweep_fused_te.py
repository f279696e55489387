"""ncu driver: the TE sub-curl (k_fused) at 8192^2 Float64."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sgc_b200 as S
n = 8192; N = (n, n); M = n * n
d = np.random.default_rng(0).uniform(0.5, 1.5, n)
F2 = S.DeviceArray(torch.rand(2 * M, device="cuda", dtype=torch.float64), N + (2,))
G1 = S.DeviceArray.zeros(N + (1,), np.float64)
te = S.Operator.curl(S.SGC_F64, N, [True, True], [d, d], [False, False], [1.0, 1.0], cmp_shp=[1, 2], cmp_out=[3], cmp_in=[1, 2])
for _ in range(6):
    te.apply(G1, F2, "=")
torch.cuda.synchronize()
print("done")
