"""GPU parity tests of the matrix-free operators: CUDA library (through the C ABI) vs oracle.

Same seeded inputs on both sides; sizes chosen so that the numpy oracle finishes in seconds.
The expected agreement is bit-for-bit (identical IEEE operations in identical order); the
north-star tolerance (4 ulp / 1e-13) is the hard limit, `exact=True` marks the cases where
bit equality is additionally REQUIRED.
"""
import itertools

import numpy as np
import pytest

from oracle import sgc_oracle as O
from tests.util import assert_parity, rand_field, rand_coef

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S():
    import sgc_b200
    return sgc_b200


def dev(S, a):
    return S.DeviceArray.from_numpy(a)


# ------------------------------------------------------------------------------------------------
# apply_∂!
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N", [(10,), (8, 9), (8, 9, 10), (33, 5, 3), (1, 4, 2), (70, 1, 1)])
@pytest.mark.parametrize("cplx", [False, True])
def test_apply_partial_matches_oracle(S, N, cplx):
    """test/partial.jl:67-70 widened: every axis, fwd/bwd, Bloch/symmetric, '=' and '+='."""
    rng = np.random.default_rng(100 + len(N) + 7 * cplx)
    Fu = rand_field(rng, N, cplx)
    G0 = rand_field(rng, N, cplx)
    for nw in range(1, len(N) + 1):
        for isfwd, isbloch, op in itertools.product((True, False), (True, False), ("=", "+=")):
            if isfwd and not isbloch and N[nw - 1] < 2:
                continue
            dwinv = rand_coef(rng, N[nw - 1], False)
            Gref = G0.copy(order="F")
            O.apply_partial(Gref, Fu, op, nw, isfwd, dwinv, isbloch)
            Gd = dev(S, G0)
            S.apply_partial(Gd, dev(S, Fu), op, nw, isfwd, dwinv, isbloch)
            assert_parity(Gd.numpy(), Gref, f"∂ N={N} nw={nw} fwd={isfwd} bloch={isbloch} {op}", exact=True)


@pytest.mark.parametrize("N", [(9,), (6, 7), (5, 6, 7)])
def test_apply_partial_complex_everything(S, N):
    """PML-stretched (complex) ∆w⁻¹, complex Bloch phase, complex α, scalar ∆ wrapper."""
    rng = np.random.default_rng(7)
    Fu = rand_field(rng, N, True)
    G0 = rand_field(rng, N, True)
    for nw in range(1, len(N) + 1):
        for isfwd, isbloch, op in itertools.product((True, False), (True, False), ("=", "+=")):
            for dwinv, e, alpha in [(rand_coef(rng, N[nw - 1], True), np.exp(-0.7j), 1.0),
                                    (rand_coef(rng, N[nw - 1], False), np.exp(0.3j), 0.5 - 0.25j),
                                    (rand_coef(rng, N[nw - 1], True), -1.0, 2.0),
                                    (0.75, np.exp(-1.1j), -1.5)]:
                Gref = G0.copy(order="F")
                O.apply_partial(Gref, Fu, op, nw, isfwd, dwinv, isbloch, e, alpha=alpha)
                Gd = dev(S, G0)
                S.apply_partial(Gd, dev(S, Fu), op, nw, isfwd, dwinv, isbloch, e, alpha=alpha)
                assert_parity(Gd.numpy(), Gref, f"∂ complex N={N} nw={nw} fwd={isfwd} bloch={isbloch} {op}")


@pytest.mark.parametrize("N", [(512,), (260, 5), (256, 3, 2), (6, 260), (4, 6, 36), (130, 7, 3), (1024, 2)])
@pytest.mark.parametrize("cplx", [False, True])
def test_vectorised_term_kernels_match_oracle_and_scalar_kernel(S, N, cplx):
    """Sizes that take the vectorised k_term_x / k_term_s kernels (16- and 32-byte vectors, shuffle
    across the vector boundary, register-marched rows): every kind, axis, direction and boundary,
    against the oracle and bit-for-bit against the scalar kernel (variant 3)."""
    rng = np.random.default_rng(150 + len(N) + 3 * cplx)
    Fu = rand_field(rng, N, cplx)
    G0 = rand_field(rng, N, cplx)
    Fd = dev(S, Fu)
    code = S.SGC_C128 if cplx else S.SGC_F64
    for nw in range(1, len(N) + 1):
        n = N[nw - 1]
        for isfwd, isbloch, op in itertools.product((True, False), (True, False), ("=", "+=")):
            e = np.exp(-0.4j) if cplx else -1.0
            alpha = (0.5 + 0.5j) if cplx else 1.25
            for kind in ("partial", "mhat", "mhat_w"):
                if kind == "partial":
                    dwinv = rand_coef(rng, n, cplx)
                    ref = G0.copy(order="F")
                    O.apply_partial(ref, Fu, op, nw, isfwd, dwinv, isbloch, e, alpha=alpha)
                    o = S.Operator.partial(code, N, nw, isfwd, dwinv, isbloch, e, alpha)
                else:
                    dw = rand_coef(rng, n, False) if kind == "mhat_w" else None
                    dwp = rand_coef(rng, n, cplx) if kind == "mhat_w" else None
                    ref = G0.copy(order="F")
                    O.apply_mhat(ref, Fu, op, nw, isfwd, dw, dwp, isbloch, e, alpha=alpha)
                    o = S.Operator.mhat(code, N, nw, isfwd, dw, dwp, isbloch, e, alpha)
                Ga, Gb = dev(S, G0), dev(S, G0)
                o.apply(Ga, Fd, op)
                o.set_tuning(0, 0, 0, 3)
                o.apply(Gb, Fd, op)
                what = f"{kind} N={N} nw={nw} fwd={isfwd} bloch={isbloch} {op}"
                assert_parity(Ga.numpy(), Gb.numpy(), what + " (vector vs scalar)", exact=True)
                assert_parity(Ga.numpy(), ref, what)
                o.destroy()


def test_apply_partial_errors(S):
    Fu = dev(S, np.zeros((4, 5)))
    with pytest.raises(S.SgcError):  # @assert size(Gv)==size(Fu)          partial.jl:251
        S.apply_partial(dev(S, np.zeros((4, 6))), Fu, "=", 1, True)
    with pytest.raises(S.SgcError):  # @assert 1≤nw≤K
        S.apply_partial(dev(S, np.zeros((4, 5))), Fu, "=", 3, True)
    with pytest.raises(S.SgcError):  # @assert size(Fu,nw)==length(∆w⁻¹)   partial.jl:253
        S.apply_partial(dev(S, np.zeros((4, 5))), Fu, "=", 1, True, np.ones(5))
    with pytest.raises(S.InexactError):  # complex value into a Float64 array
        S.apply_partial(dev(S, np.zeros((4, 5))), Fu, "=", 1, True, 1.0, True, np.exp(1j))
    with pytest.raises(S.SgcError):  # forward symmetric needs Fu[2]
        S.apply_partial(dev(S, np.zeros((1, 5))), dev(S, np.zeros((1, 5))), "=", 1, True, 1.0, False)


# ------------------------------------------------------------------------------------------------
# apply_m̂!, apply_mean!
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N", [(10,), (8, 9), (3, 4, 5), (34, 3, 2)])
@pytest.mark.parametrize("cplx", [False, True])
def test_apply_mhat_matches_oracle(S, N, cplx):
    """test/mean.jl:86-89 widened to the arithmetic variant, complex data and '+='."""
    rng = np.random.default_rng(200 + len(N))
    Fu = rand_field(rng, N, cplx)
    G0 = rand_field(rng, N, cplx)
    for nw in range(1, len(N) + 1):
        for isfwd, isbloch, op, weighted in itertools.product((True, False), (True, False), ("=", "+="), (False, True)):
            if isfwd and not isbloch and N[nw - 1] < 2:
                continue
            dw = rand_coef(rng, N[nw - 1], False) if weighted else None
            dwp = rand_coef(rng, N[nw - 1], cplx) if weighted else None
            e = np.exp(-0.4j) if cplx else -1.0
            alpha = (0.5 + 0.5j) if cplx else 1.25
            Gref = G0.copy(order="F")
            O.apply_mhat(Gref, Fu, op, nw, isfwd, dw, dwp, isbloch, e, alpha=alpha)
            Gd = dev(S, G0)
            S.apply_mhat(Gd, dev(S, Fu), op, nw, isfwd, dw, dwp, isbloch, e, alpha=alpha)
            assert_parity(Gd.numpy(), Gref, f"m̂ N={N} nw={nw} fwd={isfwd} bloch={isbloch} {op} w={weighted}")


@pytest.mark.parametrize("N", [(10,), (8, 9), (3, 4, 5)])
def test_apply_mean_matches_oracle(S, N):
    """test/mean.jl:100-102."""
    rng = np.random.default_rng(300)
    K = len(N)
    F = rand_field(rng, N + (K,), True)
    G0 = rand_field(rng, N + (K,), True)
    for isfwd, isbloch, op, weighted in itertools.product((True, False), (True, False), ("=", "+="), (False, True)):
        dl = [rand_coef(rng, n, False) for n in N] if weighted else None
        dlp = [rand_coef(rng, n, False) for n in N] if weighted else None
        e = [np.exp(-0.2j * (k + 1)) for k in range(K)]
        Gref = G0.copy(order="F")
        O.apply_mean(Gref, F, op, [isfwd] * K, dl, dlp, [isbloch] * K, e)
        Gd = dev(S, G0)
        S.apply_mean(Gd, dev(S, F), op, [isfwd] * K, dl, dlp, [isbloch] * K, e)
        assert_parity(Gd.numpy(), Gref, f"mean N={N} fwd={isfwd} bloch={isbloch} {op} w={weighted}")


# ------------------------------------------------------------------------------------------------
# apply_curl!  (fused kernel for the full 3-D curl)
# ------------------------------------------------------------------------------------------------
CURL_SIZES = [(3, 4, 5), (37, 19, 11), (64, 16, 9), (5, 3, 2), (33, 9, 1), (1, 1, 4), (2, 2, 2)]


@pytest.mark.parametrize("N", CURL_SIZES)
@pytest.mark.parametrize("isfwd", list(itertools.product((False, True), repeat=3)))
def test_apply_curl_matches_oracle(S, N, isfwd):
    """test/curl.jl:54-57: nonuniform ∆l⁻¹, isbloch = [T,F,F], random complex Bloch phases;
    plus all-Bloch and all-symmetric, '=' and '+='."""
    rng = np.random.default_rng(400 + sum(N))
    F = rand_field(rng, N + (3,), True)
    G0 = rand_field(rng, N + (3,), True)
    dlinv = [rand_coef(rng, n, False) for n in N]
    e = rng.uniform(-1, 1, 3) + 1j * rng.uniform(-1, 1, 3)
    for isbloch in ([True, False, False], [True, True, True], [False, False, False], [False, True, True]):
        if any(f and not b and n < 2 for f, b, n in zip(isfwd, isbloch, N)):
            continue
        for op in ("=", "+="):
            Gref = G0.copy(order="F")
            O.apply_curl(Gref, F, op, isfwd, dlinv, isbloch, e)
            Gd = dev(S, G0)
            S.apply_curl(Gd, dev(S, F), op, isfwd, dlinv, isbloch, e)
            assert_parity(Gd.numpy(), Gref, f"curl N={N} fwd={isfwd} bloch={isbloch} {op}", exact=True)


@pytest.mark.parametrize("cplx_field,cplx_coef,alpha", [(False, False, 1.0), (False, False, -2.5),
                                                        (True, True, 1.0), (True, False, 0.3 - 0.8j),
                                                        (True, True, 1.5 + 0.5j)])
def test_apply_curl_dtypes_and_alpha(S, cplx_field, cplx_coef, alpha):
    """Float64 fields, PML-stretched complex ∆l⁻¹, α ≠ 1 (gaps listed in SURVEY.md §4)."""
    rng = np.random.default_rng(500)
    N = (21, 10, 7)
    F = rand_field(rng, N + (3,), cplx_field)
    G0 = rand_field(rng, N + (3,), cplx_field)
    dlinv = [rand_coef(rng, n, cplx_coef) for n in N]
    e = [np.exp(-0.1j), np.exp(-0.2j), np.exp(-0.3j)] if cplx_field else [1.0, -1.0, 1.0]
    for isfwd in ([True, True, True], [False, False, False], [True, False, True]):
        for isbloch in ([True, True, True], [False, True, False]):
            for op in ("=", "+="):
                Gref = G0.copy(order="F")
                O.apply_curl(Gref, F, op, isfwd, dlinv, isbloch, e, alpha=alpha)
                Gd = dev(S, G0)
                S.apply_curl(Gd, dev(S, F), op, isfwd, dlinv, isbloch, e, alpha=alpha)
                assert_parity(Gd.numpy(), Gref, f"curl dtypes fwd={isfwd} bloch={isbloch} {op}")


# every instantiated configuration of the fused kernels: (tile_x, tile_y, stages, CTAs/SM) of the
# cp.async ring kernel (variant 0) with shuffle / shared-memory x-neighbours, and the
# register-prefetch kernel (variant 1)
def _sweep_build():
    try:
        import sgc_b200
        return bool(sgc_b200._lib.lib.sgc_build_flags() & 1)
    except Exception:
        return False


# the product library carries four ring configurations of the fused curl (and one tile of its first generation); the
# whole matrix of the tuning sweeps only exists in a `make EXTRA=-DSGC_SWEEP_BUILD` library
RING_CONFIGS = [(32, 4, 6, 4), (32, 4, 4, 4), (32, 8, 4, 3), (64, 4, 4, 3)]
V1_TILES = [(32, 8)]
if _sweep_build():
    RING_CONFIGS += [(32, 4, 3, 4), (32, 4, 5, 4), (32, 4, 7, 4), (32, 4, 8, 3), (32, 8, 3, 3), (32, 8, 4, 2), (32, 8, 6, 2), (32, 16, 3, 2),
                     (32, 16, 4, 1), (64, 4, 4, 2), (64, 8, 3, 2)]
    V1_TILES += [(32, 4), (32, 16), (64, 4), (64, 8)]
FUSED_VARIANTS = [(tx, ty, 0 | (st << 4) | (mb << 8) | (xs << 12)) for (tx, ty, st, mb) in RING_CONFIGS for xs in (0, 1)]
FUSED_VARIANTS += [(tx, ty, 1) for (tx, ty) in V1_TILES]


@pytest.mark.parametrize("cfg", FUSED_VARIANTS)
@pytest.mark.parametrize("march", [1, 3, 64])
def test_fused_curl_tilings_agree_with_unfused_sweeps(S, cfg, march):
    """Every kernel variant / tiling / ring depth / z-march length of the fused curl gives the
    bits of the per-block sweeps (variant 2 = the reference's own six-sweep structure on the GPU)."""
    rng = np.random.default_rng(600)
    N = (70, 21, 13)
    F = dev(S, rand_field(rng, N + (3,), True))
    G0 = rand_field(rng, N + (3,), True)
    dlinv = [rand_coef(rng, n, True) for n in N]
    e = [np.exp(-0.5j), np.exp(0.25j), np.exp(-1.0j)]
    for isfwd, isbloch in [([True, True, True], [True, True, True]), ([False, False, False], [True, True, True]),
                           ([True, False, True], [False, True, False]), ([False, True, False], [False, False, True])]:
        for op in ("=", "+="):
            o = S.Operator.curl(S.SGC_C128, N, isfwd, dlinv, isbloch, e, alpha=0.75)
            Ga, Gb = dev(S, G0), dev(S, G0)
            o.set_tuning(cfg[0], cfg[1], march, cfg[2])
            o.apply(Ga, F, op)
            o.set_tuning(0, 0, 0, 2)
            o.apply(Gb, F, op)
            assert o.launches(op) == (7 if op == "=" else 6)
            assert_parity(Ga.numpy(), Gb.numpy(), f"fused {cfg} march={march} fwd={isfwd} {op}", exact=True)
            o.destroy()


@pytest.mark.parametrize("N", [(200, 150, 40), (97, 130, 33)])
def test_fused_curl_interior_fast_path_on_larger_grids(S, N):
    """Grids with many interior tiles (the boundary-free fast path) next to boundary tiles, all
    BC kinds, Float64 and ComplexF64, against the per-block sweeps bit-for-bit."""
    rng = np.random.default_rng(601)
    for cplx in (True, False):
        F = dev(S, rand_field(rng, N + (3,), cplx))
        G0 = rand_field(rng, N + (3,), cplx)
        dlinv = [rand_coef(rng, n, cplx) for n in N]
        e = [np.exp(-0.5j), np.exp(0.25j), np.exp(-1.0j)] if cplx else [1.0, -1.0, 1.0]
        for isfwd in itertools.product((False, True), repeat=3):
            for isbloch in ([True, True, True], [False, False, False]):
                o = S.Operator.curl(S.SGC_C128 if cplx else S.SGC_F64, N, isfwd, dlinv, isbloch, e)
                for op in ("=", "+="):
                    Ga, Gb = dev(S, G0), dev(S, G0)
                    o.set_tuning(0, 0, 0, 0)
                    o.apply(Ga, F, op)
                    o.set_tuning(0, 0, 0, 2)
                    o.apply(Gb, F, op)
                    assert_parity(Ga.numpy(), Gb.numpy(), f"fast path N={N} cplx={cplx} fwd={isfwd} bloch={isbloch} {op}",
                                  exact=True)
                o.destroy()


def test_apply_subcurls_match_oracle(S):
    """test/curl.jl:62-119: 1×1 curls in 1-D, 1×2 / 2×1 (TE/TM) curls in 2-D."""
    rng = np.random.default_rng(700)
    N = (3, 4, 5)
    for isfwd in itertools.product((False, True), repeat=3):
        dlinv = [rand_coef(rng, n, False) for n in N]
        isbloch = [True, False, False]
        e = rng.uniform(-1, 1, 3) + 1j * rng.uniform(-1, 1, 3)
        for cmp_shp in ([1], [2], [3], [1, 2], [2, 3], [3, 1]):
            if len(cmp_shp) == 1:
                cmp_outs = [[c] for c in (1, 2, 3) if c != cmp_shp[0]]
            else:
                cmp_outs = [cmp_shp, [c for c in (1, 2, 3) if c not in cmp_shp]]
            for cmp_out in cmp_outs:
                cmp_in = [6 - cmp_shp[0] - cmp_out[0]] if len(cmp_shp) == 1 else [c for c in (1, 2, 3) if c not in cmp_out]
                Nc = tuple(N[c - 1] for c in cmp_shp)
                sel = lambda v: [v[c - 1] for c in cmp_shp]
                Fc = rand_field(rng, Nc + (len(cmp_in),), True)
                Gref = np.zeros(Nc + (len(cmp_out),), dtype=complex, order="F")
                O.apply_curl(Gref, Fc, "=", sel(isfwd), sel(dlinv), sel(isbloch), sel(e), cmp_shp=cmp_shp,
                             cmp_out=cmp_out, cmp_in=cmp_in)
                Gd = dev(S, rand_field(rng, Nc + (len(cmp_out),), True))
                S.apply_curl(Gd, dev(S, Fc), "=", sel(isfwd), sel(dlinv), sel(isbloch), sel(e), cmp_shp=cmp_shp,
                             cmp_out=cmp_out, cmp_in=cmp_in)
                assert_parity(Gd.numpy(), Gref, f"subcurl shp={cmp_shp} out={cmp_out} in={cmp_in}", exact=True)
    with pytest.raises(S.SgcError):  # curl.jl:85: cmp_shp lacks the needed axis
        S.apply_curl(dev(S, np.zeros((3, 4, 1), complex)), dev(S, np.zeros((3, 4, 1), complex)), "=", [True, True],
                     cmp_shp=[1, 1], cmp_out=[3], cmp_in=[1])


# ------------------------------------------------------------------------------------------------
# apply_divg!, apply_grad!
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N", [(7,), (4, 5), (6, 5, 4)])
def test_apply_divg_grad_match_oracle(S, N):
    """test/divergence.jl:48-56, test/gradient.jl:48-56."""
    rng = np.random.default_rng(800)
    K = len(N)
    F = rand_field(rng, N + (K,), True)
    f = rand_field(rng, N, True)
    g0 = rand_field(rng, N, True)
    G0 = rand_field(rng, N + (K,), True)
    for isfwd in itertools.product((False, True), repeat=K):
        dlinv = [rand_coef(rng, n, False) for n in N]
        isbloch = [True] + [False] * (K - 1)
        e = rng.uniform(-1, 1, K) + 1j * rng.uniform(-1, 1, K)
        for op in ("=", "+="):
            gref = g0.copy(order="F")
            O.apply_divg(gref, F, op, isfwd, dlinv, isbloch, e, alpha=0.5)
            gd = dev(S, g0)
            S.apply_divg(gd, dev(S, F), op, isfwd, dlinv, isbloch, e, alpha=0.5)
            assert_parity(gd.numpy(), gref, f"divg N={N} fwd={isfwd} {op}", exact=True)
            Gref = G0.copy(order="F")
            O.apply_grad(Gref, f, op, isfwd, dlinv, isbloch, e, alpha=-1.0)
            Gd = dev(S, G0)
            S.apply_grad(Gd, dev(S, f), op, isfwd, dlinv, isbloch, e, alpha=-1.0)
            assert_parity(Gd.numpy(), Gref, f"grad N={N} fwd={isfwd} {op}", exact=True)


# ------------------------------------------------------------------------------------------------
# one-launch compositions (k_fused): divg, grad, 2-D TE / TM sub-curls
# ------------------------------------------------------------------------------------------------
FUSED_COMPOSITION_SIZES = [(66, 9), (64, 70), (8, 260), (34, 6, 5), (128, 5, 3), (36, 37, 2), (4, 3, 9), (40, 13, 6)]


def _compositions(S, N, rng, cplx_field, cplx_coef, isfwd, isbloch, e, alpha):
    """(name, build-operator, oracle-call, output shape, input shape) for every composition on N."""
    K = len(N)
    dt = S.SGC_C128 if cplx_field else S.SGC_F64
    dlinv = [rand_coef(rng, n, cplx_coef) for n in N]
    out = [("divg", lambda: S.Operator.divg(dt, N, isfwd, dlinv, isbloch, e, alpha=alpha),
            lambda G, F, op: O.apply_divg(G, F, op, isfwd, dlinv, isbloch, e, alpha=alpha), N, N + (K,)),
           ("grad", lambda: S.Operator.grad(dt, N, isfwd, dlinv, isbloch, e, alpha=alpha),
            lambda G, F, op: O.apply_grad(G, F, op, isfwd, dlinv, isbloch, e, alpha=alpha), N + (K,), N)]
    if K == 2:
        for name, shp, co, ci in (("TE", [1, 2], [3], [1, 2]), ("TM", [1, 2], [1, 2], [3]), ("TE'", [2, 1], [3], [2, 1]),
                                  ("TM'", [3, 1], [3, 1], [2])):
            out.append((name, lambda shp=shp, co=co, ci=ci: S.Operator.curl(dt, N, isfwd, dlinv, isbloch, e, cmp_shp=shp, cmp_out=co, cmp_in=ci, alpha=alpha),
                        lambda G, F, op, shp=shp, co=co, ci=ci: O.apply_curl(G, F, op, isfwd, dlinv, isbloch, e, cmp_shp=shp, cmp_out=co, cmp_in=ci, alpha=alpha),
                        N + (len(co),), N + (len(ci),)))
    return out


@pytest.mark.parametrize("N", FUSED_COMPOSITION_SIZES)
@pytest.mark.parametrize("cplx_field,cplx_coef", [(False, False), (True, False), (True, True)])
def test_fused_compositions_match_oracle_and_sweeps(S, N, cplx_field, cplx_coef):
    """One launch per divergence / gradient / TE / TM curl, bit-identical to the oracle and to the
    library's own sweep path (variant 2 = the reference's structure), for every direction and
    boundary combination, `=` and `+=`."""
    rng = np.random.default_rng(4100 + len(N) + 2 * cplx_field + cplx_coef)
    K = len(N)
    fused_seen = 0
    for isfwd in itertools.product((False, True), repeat=K):
        for isbloch in ([True] * K, [False] * K, [True, False, True][:K]):
            e = list(np.exp(-1j * rng.uniform(0, 3, K))) if cplx_field else [1.0, -1.0, 1.0][:K]
            alpha = (0.75 - 0.5j) if cplx_coef else -1.25
            for name, build, oracle, gshape, fshape in _compositions(S, N, rng, cplx_field, cplx_coef, list(isfwd), isbloch, e, alpha):
                F = rand_field(rng, fshape, cplx_field)
                G0 = rand_field(rng, gshape, cplx_field)
                o = build()
                for op in ("=", "+="):
                    Gref = G0.copy(order="F")
                    oracle(Gref, F, op)
                    Ga, Gb = dev(S, G0), dev(S, G0)
                    o.set_tuning(0, 0, 0, 0)
                    n0 = S.launch_count()
                    o.apply(Ga, dev(S, F), op)
                    fused_seen += (S.launch_count() - n0 == 1)
                    o.set_tuning(0, 0, 0, 2)
                    o.apply(Gb, dev(S, F), op)
                    what = f"{name} N={N} fwd={isfwd} bloch={isbloch} {op}"
                    assert_parity(Ga.numpy(), Gref, what + " (fused vs oracle)", exact=True)
                    assert_parity(Ga.numpy(), Gb.numpy(), what + " (fused vs sweeps)", exact=True)
                    if K == 3:  # the 3-D forms default to the staged z-march; variant 6 = the row-register kernel
                        Gc = dev(S, G0)
                        o.set_tuning(0, 0, 0, 6)
                        o.apply(Gc, dev(S, F), op)
                        assert_parity(Gc.numpy(), Gref, what + " (row-register kernel vs oracle)", exact=True)
                o.destroy()
    assert fused_seen > 0, "the one-launch path never ran"


def test_fused_divg_grad_slab_split_equals_whole(S):
    """z-slab interfaces (ghost planes) through the fused divergence / gradient kernels."""
    rng = np.random.default_rng(4200)
    N, ks = (18, 7, 9), 4
    dlinv = [rand_coef(rng, n, True) for n in N]
    plane = N[0] * N[1] * 16
    for isfwd_z, isbloch_z in itertools.product((True, False), repeat=2):
        isfwd, isbloch, e = [False, True, isfwd_z], [True, False, isbloch_z], [np.exp(0.4j), 1.0, np.exp(-1.1j)]
        for kind, fshape, gshape in (("divg", N + (3,), N), ("grad", N, N + (3,))):
            F, G0 = rand_field(rng, fshape, True), rand_field(rng, gshape, True)
            mk = getattr(S.Operator, kind)
            whole = mk(S.SGC_C128, N, isfwd, dlinv, isbloch, e)
            Gw = dev(S, G0)
            whole.apply(Gw, dev(S, F), "+=")
            ref, out = Gw.numpy(), np.empty_like(G0)
            slabs = [(0, ks), (ks, N[2])]
            cut = lambda A, a, b: np.asfortranarray(A[:, :, a:b] if A.ndim == 3 else A[:, :, a:b, :])
            Fd = [dev(S, cut(F, a, b)) for a, b in slabs]
            for r, (a, b) in enumerate(slabs):
                Nl = (N[0], N[1], b - a)
                o = mk(S.SGC_C128, Nl, isfwd, [dlinv[0], dlinv[1], dlinv[2][a:b]], isbloch, e)
                o.set_slab(lo_interface=(r == 1), hi_interface=(r == 0))
                other, No = Fd[1 - r], slabs[1 - r][1] - slabs[1 - r][0]
                kpl = 0 if isfwd_z else No - 1
                ncomp = 3 if kind == "divg" else 1
                ghost = [(other.component(c + 1).ptr if other.ndim == 4 else other.ptr) + kpl * plane for c in range(ncomp)]
                Gd = dev(S, cut(G0, a, b))
                kb = (b - a - 1) if isfwd_z else 0
                for k0, k1 in ((0, kb), (kb, kb + 1), (kb + 1, b - a)):
                    o.apply_range(Gd, Fd[r], "+=", k0, k1, ghost)
                if out.ndim == 3:
                    out[:, :, a:b] = Gd.numpy()
                else:
                    out[:, :, a:b, :] = Gd.numpy()
                o.destroy()
            assert_parity(out, ref, f"fused {kind} slab split fwd_z={isfwd_z} bloch_z={isbloch_z}", exact=True)


# ------------------------------------------------------------------------------------------------
# fused curl-of-curl (k_curlcurl3): one pass, H never written — bit-identical to two apply_curl!
# ------------------------------------------------------------------------------------------------
# (tile_x, tile_y, march, ring stages, kernel version): version 0 / 2 = third generation (warp roles, cp.async; the
# default), 1 = first generation, 4 = the TMA-fed warp-specialised experiment k_cc4 where it applies (ComplexF64, no Bloch
# wrap in x / y, no slab ghosts; tile_y 6 or 8), else the third generation; 5 = k_cc5; 6 = k_cc6 (two planes per barrier interval;
# C1 forward / C2 backward in z, ring stages = E pair stages), else the third generation.  k_cc4 / k_cc5 are measured experiments that
# only a `make EXTRA=-DSGC_SWEEP_BUILD` library carries; the product library answers versions 4 / 5 with the third generation, so
# these tilings then exercise that routing.
CURLCURL_TILINGS = [(0, 0, 0, 0, 0), (32, 7, 3, 3, 2), (32, 7, 5, 4, 2), (32, 4, 2, 4, 2), (32, 7, 1, 3, 2), (32, 4, 7, 4, 2),
                    (32, 8, 3, 3, 1), (32, 8, 2, 3, 1), (32, 8, 6, 3, 1),
                    (0, 6, 3, 0, 4), (0, 8, 1, 0, 4), (0, 6, 0, 0, 4), (0, 7, 3, 4, 5), (0, 7, 1, 3, 5), (0, 5, 2, 4, 5), (0, 7, 0, 0, 5),
                    (0, 5, 0, 3, 6), (0, 5, 3, 3, 6), (0, 5, 8, 3, 6), (0, 5, 5, 3, 6)]


def _curlcurl_case(S, rng, N, cplx_field, cplx_coef, isfwd1, isfwd2, isbloch, tilings, what):
    F = rand_field(rng, N + (3,), cplx_field)
    G0 = rand_field(rng, N + (3,), cplx_field)
    d1 = [rand_coef(rng, n, cplx_coef) for n in N]
    d2 = [rand_coef(rng, n, cplx_coef) for n in N]
    e = list(np.exp(-1j * rng.uniform(0, 3, 3))) if cplx_field else [1.0, -1.0, 1.0]
    dt = S.SGC_C128 if cplx_field else S.SGC_F64
    a1, a2 = ((0.5 + 0.25j), (-1.5 + 0.5j)) if cplx_coef else (0.5, -1.5)
    c1 = S.Operator.curl(dt, N, isfwd1, d1, isbloch, e, alpha=a1)
    c2 = S.Operator.curl(dt, N, isfwd2, d2, isbloch, e, alpha=a2)
    Fd = dev(S, F)
    for op in ("=", "+="):
        H = S.DeviceArray.zeros(N + (3,), np.complex128 if cplx_field else np.float64)
        Gref = dev(S, G0)
        c1.apply(H, Fd, "=")
        c2.apply(Gref, H, op)
        ref = Gref.numpy()
        for (tx, ty, march, se, ver) in tilings:
            c2.set_tuning(tx, ty, march, (se << 4) | ver)
            Gd = dev(S, G0)
            n0 = S.launch_count()
            c2.apply_after(c1, Gd, Fd, op)
            assert S.launch_count() == n0 + 1
            assert_parity(Gd.numpy(), ref, f"{what} {op} tile=({tx},{ty}) march={march} stages={se} v{ver}", exact=True)
        c2.set_tuning(0, 0, 0, 0)
    # and against the oracle (two apply_curl! of the restatement)
    Href = np.zeros_like(F)
    O.apply_curl(Href, F, "=", isfwd1, d1, isbloch, e, alpha=a1)
    Gor = G0.copy(order="F")
    O.apply_curl(Gor, Href, "+=", isfwd2, d2, isbloch, e, alpha=a2)
    Gd = dev(S, G0)
    c2.apply_after(c1, Gd, Fd, "+=")
    assert_parity(Gd.numpy(), Gor, what + " vs oracle", exact=True)
    c1.destroy(); c2.destroy()


@pytest.mark.parametrize("N", [(40, 24, 12), (33, 9, 5), (70, 19, 7), (5, 3, 2), (64, 1, 3), (1, 1, 1)])
@pytest.mark.parametrize("cplx_field,cplx_coef", [(True, True), (True, False), (False, False)])
def test_fused_curlcurl_equals_two_curls(S, N, cplx_field, cplx_coef):
    rng = np.random.default_rng(5100 + sum(N) + 2 * cplx_field + cplx_coef)
    dirs = [([True] * 3, [False] * 3), ([False] * 3, [True] * 3), ([True, False, True], [True, True, False]),
            ([False, True, False], [False, True, False])]
    for isfwd1, isfwd2 in dirs:
        for isbloch in ([True] * 3, [False] * 3, [True, False, True], [False, True, False]):
            if not all(isbloch[d] or N[d] >= 2 or not (isfwd1[d] or isfwd2[d]) for d in range(3)):
                continue  # forward + symmetric on an axis of length 1 is rejected (reads Fu[2])
            _curlcurl_case(S, rng, N, cplx_field, cplx_coef, isfwd1, isfwd2, isbloch, CURLCURL_TILINGS[:3] + CURLCURL_TILINGS[6:7] + CURLCURL_TILINGS[9:10] + CURLCURL_TILINGS[12:13] + CURLCURL_TILINGS[16:18],
                           f"curlcurl N={N} fwd1={isfwd1} fwd2={isfwd2} bloch={isbloch}")


@pytest.mark.parametrize("tiling", CURLCURL_TILINGS)
def test_fused_curlcurl_tilings_on_a_larger_grid(S, tiling):
    """Interior (boundary-free) tiles, ragged tiles and every built tile shape."""
    rng = np.random.default_rng(5200)
    for N, isbloch in (((200, 75, 21), [False] * 3), ((97, 66, 40), [True, True, False])):
        for isfwd1, isfwd2 in (([True] * 3, [False] * 3), ([False, True, False], [True, False, True])):
            _curlcurl_case(S, rng, N, True, True, isfwd1, isfwd2, isbloch, [tiling], f"curlcurl N={N} tiling={tiling}")


def test_fused_curlcurl_slab_split_equals_whole(S):
    """Two z-slabs with ghost planes (what the multi-GPU driver does over peer memory): interior
    interfaces, and physical Bloch faces whose wrap-around partner is the other slab."""
    rng = np.random.default_rng(5300)
    N, ks = (37, 18, 11), 4
    F, G0 = rand_field(rng, N + (3,), True), rand_field(rng, N + (3,), True)
    d1, d2 = [rand_coef(rng, n, True) for n in N], [rand_coef(rng, n, True) for n in N]
    plane = N[0] * N[1] * 16
    for fz1, isbloch_z in itertools.product((True, False), repeat=2):
        isfwd1, isfwd2 = [True, False, fz1], [False, True, not fz1]
        isbloch, e = [True, False, isbloch_z], [np.exp(0.3j), 1.0, np.exp(-0.7j)]
        w1 = S.Operator.curl(S.SGC_C128, N, isfwd1, d1, isbloch, e, alpha=0.5 - 0.25j)
        w2 = S.Operator.curl(S.SGC_C128, N, isfwd2, d2, isbloch, e)
        Gw = dev(S, G0)
        w2.apply_after(w1, Gw, dev(S, F), "+=")
        ref, out = Gw.numpy(), np.empty_like(G0)
        slabs = [(0, ks), (ks, N[2])]
        Fd = [dev(S, np.asfortranarray(F[:, :, a:b, :])) for a, b in slabs]
        for r, (a, b) in enumerate(slabs):
            Nl = (N[0], N[1], b - a)
            c1 = S.Operator.curl(S.SGC_C128, Nl, isfwd1, [d1[0], d1[1], d1[2][a:b]], isbloch, e, alpha=0.5 - 0.25j)
            c2 = S.Operator.curl(S.SGC_C128, Nl, isfwd2, [d2[0], d2[1], d2[2][a:b]], isbloch, e)
            for c in (c1, c2):
                c.set_slab(lo_interface=(r == 1), hi_interface=(r == 0))
            other, No = Fd[1 - r], slabs[1 - r][1] - slabs[1 - r][0]
            # with two slabs the other slab is both the lower and the upper neighbour (ring)
            below = [other.component(c + 1).ptr + (No - 1) * plane for c in range(3)]
            above = [other.component(c + 1).ptr for c in range(3)]
            need_lo = (r == 1) or isbloch_z   # interface, or a physical Bloch face (wrap partner = other slab)
            need_hi = (r == 0) or isbloch_z
            c1.set_slab_coef(dzinv_below=d1[2][(a - 1) % N[2]], dzinv_above=d1[2][b % N[2]])
            Gd = dev(S, np.asfortranarray(G0[:, :, a:b, :]))
            kmid = (b - a) // 2
            for k0, k1 in ((0, kmid), (kmid, b - a)):  # two plane ranges, as a chunked driver would issue them
                c2.apply_after(c1, Gd, Fd[r], "+=", k_begin=k0, k_end=k1, ghost_lo=below if need_lo else None,
                               ghost_hi=above if need_hi else None)
            out[:, :, a:b, :] = Gd.numpy()
            c1.destroy(); c2.destroy()
        assert_parity(out, ref, f"curlcurl slab split fz1={fz1} bloch_z={isbloch_z}", exact=True)
        w1.destroy(); w2.destroy()


# ------------------------------------------------------------------------------------------------
# z-slab mode on one GPU: two slabs with ghost planes reproduce the undivided result
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("isfwd_z,isbloch_z", [(True, True), (False, True), (True, False), (False, False)])
def test_slab_split_with_ghost_planes_equals_whole(S, isfwd_z, isbloch_z):
    rng = np.random.default_rng(900)
    N = (19, 11, 12)
    ks = 5  # slab 0: planes 0..4, slab 1: planes 5..11
    F = rand_field(rng, N + (3,), True)
    G0 = rand_field(rng, N + (3,), True)
    dlinv = [rand_coef(rng, n, True) for n in N]
    isfwd = [True, False, isfwd_z]
    isbloch = [True, False, isbloch_z]
    e = [np.exp(-0.3j), 1.0, np.exp(0.9j)]
    whole = S.Operator.curl(S.SGC_C128, N, isfwd, dlinv, isbloch, e)
    Gw = dev(S, G0)
    whole.apply(Gw, dev(S, F), "+=")
    ref = Gw.numpy()
    out = np.empty_like(ref)
    slabs = [(0, ks), (ks, N[2])]
    Fd = [dev(S, np.asfortranarray(F[:, :, a:b, :])) for a, b in slabs]
    plane = N[0] * N[1] * 16
    for r, (a, b) in enumerate(slabs):
        Nl = (N[0], N[1], b - a)
        o = S.Operator.curl(S.SGC_C128, Nl, isfwd, [dlinv[0], dlinv[1], dlinv[2][a:b]], isbloch, e)
        o.set_slab(lo_interface=(r == 1), hi_interface=(r == 0))
        other = Fd[1 - r]
        No = slabs[1 - r][1] - slabs[1 - r][0]
        ghost = []
        for c in range(2):  # Fx, Fy need the z-neighbour plane
            comp = other.component(c + 1)
            # forward: plane above my top = first plane of the other slab (rank 0) or the wrap plane 0
            # of the global grid (rank 1); backward: plane below my bottom = last plane of the other slab
            kplane = 0 if isfwd_z else No - 1
            ghost.append(comp.ptr + kplane * plane)
        ghost.append(None)
        Gd = dev(S, np.asfortranarray(G0[:, :, a:b, :]))
        # interior planes and the boundary plane in two launches, as the overlapped halo path does
        kb = (b - a - 1) if isfwd_z else 0
        for k0, k1 in ((0, kb), (kb, kb + 1), (kb + 1, b - a)):
            o.apply_range(Gd, Fd[r], "+=", k0, k1, ghost)
        out[:, :, a:b, :] = Gd.numpy()
        o.destroy()
    assert_parity(out, ref, f"slab split fwd_z={isfwd_z} bloch_z={isbloch_z}", exact=True)


# ------------------------------------------------------------------------------------------------
# the other boundary forms: one-shot C entry points and host arrays
# ------------------------------------------------------------------------------------------------
def test_one_shot_c_entry_points_and_host_arrays(S):
    import ctypes as C
    from sgc_b200 import _lib
    from sgc_b200.device import coef_vectors, ints, int64s, scalar2, scalars2, stream_handle
    rng = np.random.default_rng(1000)
    N = (9, 8, 7)
    F = rand_field(rng, N + (3,), True)
    dlinv = [rand_coef(rng, n, False) for n in N]
    isfwd, isbloch, e = [True, False, True], [True, True, False], [np.exp(-0.3j), np.exp(0.2j), 1.0]
    Gref = np.zeros_like(F)
    O.apply_curl(Gref, F, "=", isfwd, dlinv, isbloch, e)
    arr, keep, cc = coef_vectors(dlinv, N)
    Fd, Gd = dev(S, F), dev(S, np.zeros_like(F))
    _lib.check(_lib.lib.sgc_apply_curl(C.c_void_p(Gd.ptr), C.c_void_p(Fd.ptr), 0, S.SGC_C128, 3, int64s(N),
                                       ints(isfwd), arr, cc, ints(isbloch), scalars2(e), ints([1, 2, 3]), 3,
                                       ints([1, 2, 3]), 3, ints([1, 2, 3]), scalar2(1.0), stream_handle()))
    assert_parity(Gd.numpy(), Gref, "sgc_apply_curl one-shot", exact=True)
    # divergence / gradient / mean one-shots
    gref = np.zeros(N, dtype=complex, order="F")
    O.apply_divg(gref, F, "=", isfwd, dlinv, isbloch, e)
    gd = dev(S, np.zeros(N, dtype=complex))
    _lib.check(_lib.lib.sgc_apply_divg(C.c_void_p(gd.ptr), C.c_void_p(Fd.ptr), 0, S.SGC_C128, 3, int64s(N),
                                       ints(isfwd), arr, cc, ints(isbloch), scalars2(e), scalar2(1.0), stream_handle()))
    assert_parity(gd.numpy(), gref, "sgc_apply_divg one-shot", exact=True)
    f = rand_field(rng, N, True)
    Gref = np.zeros_like(F)
    O.apply_grad(Gref, f, "=", isfwd, dlinv, isbloch, e)
    _lib.check(_lib.lib.sgc_apply_grad(C.c_void_p(Gd.ptr), C.c_void_p(dev(S, f).ptr), 0, S.SGC_C128, 3, int64s(N),
                                       ints(isfwd), arr, cc, ints(isbloch), scalars2(e), scalar2(1.0), stream_handle()))
    assert_parity(Gd.numpy(), Gref, "sgc_apply_grad one-shot", exact=True)
    Gref = np.zeros_like(F)
    O.apply_mean(Gref, F, "=", isfwd, None, None, isbloch, e)
    _lib.check(_lib.lib.sgc_apply_mean(C.c_void_p(Gd.ptr), C.c_void_p(Fd.ptr), 0, S.SGC_C128, 3, int64s(N),
                                       ints(isfwd), None, None, 0, ints(isbloch), scalars2(e), scalar2(1.0),
                                       stream_handle()))
    assert_parity(Gd.numpy(), Gref, "sgc_apply_mean one-shot")
    fu = np.asfortranarray(F[..., 0])
    gref = np.zeros(N, dtype=complex, order="F")
    O.apply_mhat(gref, fu, "=", 2, False, None, None, True, e[1])
    _lib.check(_lib.lib.sgc_apply_mhat(C.c_void_p(gd.ptr), C.c_void_p(dev(S, fu).ptr), 0, S.SGC_C128, 3, int64s(N),
                                       2, 0, None, None, 0, 1, scalar2(e[1]), scalar2(1.0), stream_handle()))
    assert_parity(gd.numpy(), gref, "sgc_apply_mhat one-shot")
    # host arrays through sgc_op_apply_host (the strict drop-in form for Julia `Array`s)
    o = S.Operator.curl(S.SGC_C128, N, isfwd, dlinv, isbloch, e)
    Gh = rand_field(rng, N + (3,), True)
    Gref = Gh.copy(order="F")
    O.apply_curl(Gref, F, "+=", isfwd, dlinv, isbloch, e)
    o.apply_host(Gh, F, "+=")
    assert_parity(Gh, Gref, "apply_host", exact=True)
    o.destroy()


def test_full_size_properties_256(S):
    """BASELINE config 2 size (256³, ComplexF64): size-independent properties instead of the
    oracle — linearity, curl of a constant field vanishes under periodic BCs, div(curl) == 0
    to rounding on a uniform grid (test/divergence.jl:59-74), '+=' into zeros == '='."""
    import torch
    N = (256, 256, 256)
    M = N[0] * N[1] * N[2]
    gen = torch.Generator(device="cuda").manual_seed(1)
    F = S.DeviceArray(torch.complex(torch.rand(3 * M, generator=gen, device="cuda", dtype=torch.float64) - 0.5,
                                    torch.rand(3 * M, generator=gen, device="cuda", dtype=torch.float64) - 0.5), N + (3,))
    isfwd = [True, True, True]
    curl = S.Operator.curl(S.SGC_C128, N, isfwd)
    G = S.DeviceArray.empty(N + (3,))
    curl.apply(G, F, "=")
    # constant field → zero curl
    Cst = S.DeviceArray(torch.full((3 * M,), 0.25 - 0.5j, dtype=torch.complex128, device="cuda"), N + (3,))
    Gc = S.DeviceArray.empty(N + (3,))
    curl.apply(Gc, Cst, "=")
    assert float(Gc.t.abs().max()) == 0.0
    # div(curl F) == 0 exactly on the uniform periodic grid
    divg = S.Operator.divg(S.SGC_C128, N, isfwd)
    d = S.DeviceArray.empty(N)
    divg.apply(d, G, "=")
    scale = float(G.t.abs().max())
    assert float(d.t.abs().max()) <= 16 * np.finfo(float).eps * scale
    # '+=' into zeros reproduces '=' bit-for-bit; a second '+=' doubles it up to rounding
    G2 = S.DeviceArray.zeros(N + (3,))
    curl.apply(G2, F, "+=")
    assert torch.equal(G2.t, G.t)
    curl.apply(G2, F, "+=")
    assert float((G2.t - 2 * G.t).abs().max()) <= 8 * np.finfo(float).eps * scale
    # fused vs unfused sweeps at full size
    curl.set_tuning(0, 0, 0, 2)
    G3 = S.DeviceArray.empty(N + (3,))
    curl.apply(G3, F, "=")
    assert torch.equal(G3.t, G.t)
