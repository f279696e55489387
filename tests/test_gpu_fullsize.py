"""BASELINE.json configs at their OWN sizes against the oracle (not invariants): configs[1] 256³ PML-stretched
curl-of-curl, configs[2] 8192² Float64 ∂ / m̂ / TE / TM with PEC/PMC boundaries, configs[3]'s create_curl at
128³–192³ against the C restatement of the reference's COO → `sparse` → `dropzeros!` path.

The reference's own checks of the same facts: test/curl.jl:54-57 (matrix-free against the matrix), test/partial.jl:64-70,
test/mean.jl:83-89.  Tolerance (north star): colptr / rowval bit-exact, values within 4 ulp / 1e-13 — the tests
additionally require bit equality where both sides execute the same IEEE operations in the same order.
"""
import importlib.util
import os

import numpy as np
import pytest

from oracle import sgc_oracle as O
from tests.util import assert_parity

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def S():
    import sgc_b200
    return sgc_b200


def _bench_module():
    spec = importlib.util.spec_from_file_location("sgc_bench_inputs", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("bc", ["pec", "bloch"])
def test_config1_256cubed_pml_curlcurl_matches_c_oracle(S, bc):
    """configs[1]: the exact inputs bench.py times (256³ nonuniform grid, PML-stretched complex ∆l⁻¹), E → H → E′
    with two apply_curl!, and the fused one-pass curl-of-curl, against the C oracle on the same E.  The Bloch
    variant adds phases on every axis (division by e at the backward wrap)."""
    from oracle import c_oracle as CO
    B = _bench_module()
    dprim, ddual = B.build_inputs(256)
    N = (256, 256, 256)
    rng = np.random.default_rng(20260)
    E = np.asfortranarray(rng.uniform(-1, 1, N + (3,)) + 1j * rng.uniform(-1, 1, N + (3,)))
    bl = [bc == "bloch"] * 3
    e = list(np.exp(-2j * np.pi * np.array([0.1, 0.2, 0.3]))) if bc == "bloch" else [1.0] * 3
    Href, E2ref = np.zeros_like(E), np.zeros_like(E)
    CO.apply_curl(Href, E, "=", [True] * 3, ddual, bl, e)
    CO.apply_curl(E2ref, Href, "=", [False] * 3, dprim, bl, e)
    c1 = S.Operator.curl(S.SGC_C128, N, [True] * 3, ddual, bl, e)
    c2 = S.Operator.curl(S.SGC_C128, N, [False] * 3, dprim, bl, e)
    Ed = S.DeviceArray.from_numpy(E)
    Hd, E2d = S.DeviceArray.empty(N + (3,)), S.DeviceArray.empty(N + (3,))
    c1.apply(Hd, Ed, "=")
    c2.apply(E2d, Hd, "=")
    assert_parity(Hd.numpy(), Href, f"256³ {bc} H = C_dual E", exact=True)
    assert_parity(E2d.numpy(), E2ref, f"256³ {bc} E' = C_prim H", exact=True)
    E2d.t.zero_()
    c2.apply_after(c1, E2d, Ed, "=")
    assert_parity(E2d.numpy(), E2ref, f"256³ {bc} fused curl-of-curl", exact=True)


def test_config2_8192sq_f64_operators_match_oracle(S):
    """configs[2]: 2-D 8192² Float64 apply_∂! / apply_m̂! (arithmetic and weighted) and the TE / TM sub-curls with
    symmetric (PEC/PMC) boundaries, '=' and '+=', against the numpy oracle at full size."""
    n = 8192
    N = (n, n)
    rng = np.random.default_rng(3)
    d = rng.uniform(0.5, 1.5, n)
    dp = rng.uniform(0.5, 1.5, n)
    F = np.asfortranarray(rng.uniform(-1, 1, N))
    G0 = np.asfortranarray(rng.uniform(-1, 1, N))
    Fd = S.DeviceArray.from_numpy(F)
    for nw in (1, 2):
        for isfwd in (True, False):
            for op in ("=", "+="):
                ref = G0.copy(order="F")
                O.apply_partial(ref, F, op, nw, isfwd, d, False)
                Gd = S.DeviceArray.from_numpy(G0)
                S.apply_partial(Gd, Fd, op, nw, isfwd, d, False)
                assert_parity(Gd.numpy(), ref, f"8192² ∂ nw={nw} fwd={isfwd} {op}", exact=True)
            ref = G0.copy(order="F")
            O.apply_mhat(ref, F, "=", nw, isfwd, None, None, False)
            Gd = S.DeviceArray.from_numpy(G0)
            S.apply_mhat(Gd, Fd, "=", nw, isfwd, None, None, False)
            assert_parity(Gd.numpy(), ref, f"8192² m̂ arithmetic nw={nw} fwd={isfwd}", exact=True)
            ref = G0.copy(order="F")
            O.apply_mhat(ref, F, "+=", nw, isfwd, d, dp, False)
            Gd = S.DeviceArray.from_numpy(G0)
            S.apply_mhat(Gd, Fd, "+=", nw, isfwd, d, dp, False)
            assert_parity(Gd.numpy(), ref, f"8192² m̂ weighted nw={nw} fwd={isfwd}", exact=True)
    # TE: Hz from (Ex, Ey); TM: (Ex, Ey) from Hz   (test/curl.jl:67-92)
    F2 = np.asfortranarray(rng.uniform(-1, 1, N + (2,)))
    F1 = np.asfortranarray(rng.uniform(-1, 1, N + (1,)))
    for isfwd in ([True, True], [False, False]):
        ref = np.zeros(N + (1,), order="F")
        O.apply_curl(ref, F2, "=", isfwd, [d, dp], [False, False], [1.0, 1.0], cmp_shp=[1, 2], cmp_out=[3], cmp_in=[1, 2])
        Gd = S.DeviceArray.zeros(N + (1,), np.float64)
        S.apply_curl(Gd, S.DeviceArray.from_numpy(F2), "=", isfwd, [d, dp], [False, False], [1.0, 1.0], cmp_shp=[1, 2],
                     cmp_out=[3], cmp_in=[1, 2])
        assert_parity(Gd.numpy(), ref, f"8192² TE sub-curl fwd={isfwd}", exact=True)
        ref = np.zeros(N + (2,), order="F")
        O.apply_curl(ref, F1, "=", isfwd, [d, dp], [False, False], [1.0, 1.0], cmp_shp=[1, 2], cmp_out=[1, 2], cmp_in=[3])
        Gd = S.DeviceArray.zeros(N + (2,), np.float64)
        S.apply_curl(Gd, S.DeviceArray.from_numpy(F1), "=", isfwd, [d, dp], [False, False], [1.0, 1.0], cmp_shp=[1, 2],
                     cmp_out=[1, 2], cmp_in=[3])
        assert_parity(Gd.numpy(), ref, f"8192² TM sub-curl fwd={isfwd}", exact=True)


@pytest.mark.parametrize("n,isbloch,order", [(128, [True, True, True], True), (128, [True, False, False], False),
                                              (192, [True, True, True], True)])
def test_config3_create_curl_matches_c_reference_path(S, n, isbloch, order):
    """configs[3] family: Float64 create_curl at 128³ / 192³ (25–85 M nnz), all-Bloch and a mixed boundary whose
    symmetric axes drop entries, both orderings, against the C restatement of the reference's own path
    (create_∂info temporaries → COO → `sparse` → `dropzeros!`, oracle/sgc_oracle_c.c)."""
    from oracle import c_oracle as CO
    rng = np.random.default_rng(n)
    dl = [rng.uniform(0.5, 1.5, n) for _ in range(3)]
    isfwd = [True, False, True]
    e = [1.0, 1.0, 1.0]
    cp, rv, nz = CO.create_curl(isfwd, dl, isbloch, e, order_cmpfirst=order)
    A = S.create_curl(isfwd, dl, isbloch, e, order_cmpfirst=order)
    gcp, grv, gnz = A.to_host()
    assert np.array_equal(gcp, cp), "colptr differs from the reference path"
    assert np.array_equal(grv, rv), "rowval differs from the reference path"
    assert_parity(gnz, nz, f"create_curl {n}³ nzval", exact=True)
