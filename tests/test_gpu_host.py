"""The host-array entry points of the C ABI (sgc_op_apply_host[_chunked], sgc_op_apply_curlcurl_host): the strict drop-in
form of the reference's calls with Julia `Array`s (apply_∂! partial.jl:33-43, apply_curl! curl.jl:52-63, …).  Every
chunking of the z-pipelined path must reproduce the oracle bit for bit, for every stencil reach along the last axis
(forward / backward / none), Bloch (deferred wrap plane) and symmetric boundaries, `=` and `+=`."""
import itertools

import numpy as np
import pytest

from oracle import sgc_oracle as O
from tests.util import assert_parity, rand_coef, rand_field

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S():
    import sgc_b200
    return sgc_b200


@pytest.mark.parametrize("nchunks", [1, 2, 5, 64])
def test_apply_host_chunked_matches_oracle(S, nchunks):
    rng = np.random.default_rng(40 + nchunks)
    N = (19, 12, 14)
    dl = [rand_coef(rng, n, True) for n in N]
    e = [np.exp(-0.3j), np.exp(0.5j), np.exp(-0.9j)]
    F3 = rand_field(rng, N + (3,), True)
    f1 = rand_field(rng, N, True)
    for fz, bz, op in itertools.product((True, False), (True, False), ("=", "+=")):
        isfwd, isbloch = [True, False, fz], [False, True, bz]
        tag = f"fwdz={fz} blochz={bz} {op} chunks={nchunks}"
        # curl (reach ±1 along z through two of its six terms)
        G = rand_field(rng, N + (3,), True)
        ref = G.copy(order="F")
        O.apply_curl(ref, F3, op, isfwd, dl, isbloch, e)
        o = S.Operator.curl(S.SGC_C128, N, isfwd, dl, isbloch, e)
        o.apply_host(G, F3, op, nchunks=nchunks)
        assert_parity(G, ref, f"apply_host curl {tag}", exact=True)
        # ∂z, ∂x (no reach along the last axis), m̂z
        for nw in (3, 1):
            g = rand_field(rng, N, True)
            ref = g.copy(order="F")
            O.apply_partial(ref, f1, op, nw, isfwd[nw - 1], dl[nw - 1], isbloch[nw - 1], e[nw - 1])
            o = S.Operator.partial(S.SGC_C128, N, nw, isfwd[nw - 1], dl[nw - 1], isbloch[nw - 1], e[nw - 1])
            o.apply_host(g, f1, op, nchunks=nchunks)
            assert_parity(g, ref, f"apply_host ∂{nw} {tag}", exact=True)
        g = rand_field(rng, N, True)
        ref = g.copy(order="F")
        O.apply_mhat(ref, f1, op, 3, fz, dl[2], dl[2], bz, e[2])
        o = S.Operator.mhat(S.SGC_C128, N, 3, fz, dl[2], dl[2], bz, e[2])
        o.apply_host(g, f1, op, nchunks=nchunks)
        assert_parity(g, ref, f"apply_host m̂z {tag}")
        # divergence (vector in, scalar out) and gradient
        g = rand_field(rng, N, True)
        ref = g.copy(order="F")
        O.apply_divg(ref, F3, op, isfwd, dl, isbloch, e)
        S.Operator.divg(S.SGC_C128, N, isfwd, dl, isbloch, e).apply_host(g, F3, op, nchunks=nchunks)
        assert_parity(g, ref, f"apply_host divg {tag}", exact=True)
        G = rand_field(rng, N + (3,), True)
        ref = G.copy(order="F")
        O.apply_grad(ref, f1, op, isfwd, dl, isbloch, e)
        S.Operator.grad(S.SGC_C128, N, isfwd, dl, isbloch, e).apply_host(G, f1, op, nchunks=nchunks)
        assert_parity(G, ref, f"apply_host grad {tag}", exact=True)


def test_apply_host_2d_and_1d_float64(S):
    rng = np.random.default_rng(44)
    for N in ((37, 29), (53,)):
        F = rand_field(rng, N, False)
        nw = len(N)
        d = rand_coef(rng, N[nw - 1], False)
        for isfwd, isbloch, op in itertools.product((True, False), (True, False), ("=", "+=")):
            g = rand_field(rng, N, False)
            ref = g.copy(order="F")
            O.apply_partial(ref, F, op, nw, isfwd, d, isbloch)
            S.Operator.partial(S.SGC_F64, N, nw, isfwd, d, isbloch).apply_host(g, F, op, nchunks=3)
            assert_parity(g, ref, f"apply_host {len(N)}-D ∂ fwd={isfwd} bloch={isbloch} {op}", exact=True)


@pytest.mark.parametrize("nchunks", [0, 1, 3, 7])
def test_curlcurl_host_matches_two_oracle_curls(S, nchunks):
    rng = np.random.default_rng(50 + nchunks)
    N = (33, 17, 21)
    d1 = [rand_coef(rng, n, True) for n in N]
    d2 = [rand_coef(rng, n, True) for n in N]
    e = [np.exp(-0.3j), np.exp(0.5j), np.exp(-0.9j)]
    F = rand_field(rng, N + (3,), True)
    for (f1z, f2z), bz, op in itertools.product(((True, False), (False, True), (True, True)), (True, False), ("=", "+=")):
        isfwd1, isfwd2, isbloch = [True, True, f1z], [False, False, f2z], [True, False, bz]
        G = rand_field(rng, N + (3,), True)
        H, ref = np.zeros_like(F), G.copy(order="F")
        O.apply_curl(H, F, "=", isfwd1, d1, isbloch, e)
        O.apply_curl(ref, H, op, isfwd2, d2, isbloch, e)
        c1 = S.Operator.curl(S.SGC_C128, N, isfwd1, d1, isbloch, e)
        c2 = S.Operator.curl(S.SGC_C128, N, isfwd2, d2, isbloch, e)
        c2.apply_after_host(c1, G, F, op, nchunks=nchunks)
        assert_parity(G, ref, f"curlcurl_host z-dirs ({f1z},{f2z}) blochz={bz} {op} chunks={nchunks}", exact=True)


def test_host_registration_follows_array_lifetime(S):
    """The Python host ties the cached page-locking to the owning array: dropping the array drops the registration
    (a new array that lands on the same address is registered afresh and gives correct results)."""
    import gc
    rng = np.random.default_rng(60)
    N = (64, 64, 48)
    d = [rand_coef(rng, n, False) for n in N]
    o = S.Operator.curl(S.SGC_C128, N, [True] * 3, d, [True] * 3, [1.0] * 3)
    for _ in range(4):
        F = rand_field(rng, N + (3,), True)
        G = np.zeros_like(F)
        ref = np.zeros_like(F)
        O.apply_curl(ref, F, "=", [True] * 3, d, [True] * 3, [1.0] * 3)
        o.apply_host(G, F, "=")
        assert_parity(G, ref, "apply_host after re-allocation", exact=True)
        del F, G
        gc.collect()


def test_device_pml_vectors_and_operator_refresh(S):
    """s∆l / (s∆l)⁻¹ produced on the device (`sgc_axis_stretch`) against the host restatement of
    create_stretched_∆l / invert_∆l (gridgen.py; src/pml.jl:94-153, src/util.jl:42-48) within 4 ulp, and an operator
    whose tables were rebuilt on the device for a new ω (`sgc_op_set_axes`) against one created from the host vectors."""
    from sgc_b200 import gridgen as GG
    rng = np.random.default_rng(70)
    N = (48, 40, 36)
    lprim = [GG.graded_lprim(n, rng) for n in N]
    for isbloch in ([False] * 3, [True, False, True]):
        grid = GG.Grid(lprim, isbloch)
        Npml = ([6, 5, 0], [4, 6, 7])
        for which in (0, 1):
            axes = GG.DeviceAxes(grid, Npml, which)
            for omega in (2 * np.pi / 9.3, 0.37):
                sdl = GG.create_stretched_dl(omega, grid, Npml)
                inv = GG.invert_dl(sdl)
                got = axes.stretch(omega)
                for k in range(3):
                    assert_parity(got[k][0].cpu().numpy(), sdl[which][k], f"s∆l axis {k} set {which} ω={omega:.3f}")
                    assert_parity(got[k][1].cpu().numpy(), inv[which][k], f"(s∆l)⁻¹ axis {k} set {which} ω={omega:.3f}")
            # operator refresh: create at ω₀, rebuild on the device for ω₁, compare with an operator created at ω₁
            w0, w1 = 0.5, 2 * np.pi / 7.1
            isfwd = [True, False, True]
            e = [np.exp(-0.3j), 1.0, np.exp(0.9j)]
            inv0 = GG.invert_dl(GG.create_stretched_dl(w0, grid, Npml))[which]
            inv1 = [t[1].cpu().numpy() for t in axes.stretch(w1)]
            F = rand_field(rng, N + (3,), True)
            for maker in (lambda d: S.Operator.curl(S.SGC_C128, N, isfwd, d, isbloch, e, alpha=0.5 - 0.25j),
                          lambda d: S.Operator.divg(S.SGC_C128, N, isfwd, d, isbloch, e),
                          lambda d: S.Operator.grad(S.SGC_C128, N, isfwd, d, isbloch, e, alpha=-2.0)):
                o = maker(inv0)
                axes.refresh(o, w1)
                ref = maker(inv1)
                nout, nin = o.n_out, o.n_in
                Fd = S.DeviceArray.from_numpy(F if nin == 3 else np.asfortranarray(F[..., 0]))
                shp = N + (3,) if nout == 3 else N
                G1, G2 = S.DeviceArray.zeros(shp), S.DeviceArray.zeros(shp)
                o.apply(G1, Fd, "=")
                ref.apply(G2, Fd, "=")
                assert_parity(G1.numpy(), G2.numpy(), "operator refreshed on the device vs created from the same vectors", exact=True)
