"""The reference's own test-suite, ported to pin the oracle (CPU only, no GPU).

Each test cites the reference test it restates.  Matrix comparisons are exact (`==`),
matrix-free vs matrix comparisons use Julia's `≈` (rtol = sqrt(eps)), as in the reference.
"""
import itertools

import numpy as np
import pytest

from oracle import sgc_oracle as O
from tests.naive import naive_partial, naive_mhat, vec, blk2cmp_perm

RTOL = np.sqrt(np.finfo(np.float64).eps)


def jl_isapprox(a, b):
    return np.linalg.norm(a - b) <= RTOL * max(np.linalg.norm(a), np.linalg.norm(b))


@pytest.mark.parametrize("N", [(10,), (8, 9), (8, 9, 10)])
def test_create_partial_and_apply_partial(N):
    """test/partial.jl:3-77 (1-D), :79-153 (2-D), :155-229 (3-D)."""
    rng = np.random.default_rng(1)
    M = int(np.prod(N))
    Fu = np.asfortranarray(rng.random(N))
    for nw in range(1, len(N) + 1):
        dwinv = rng.random(N[nw - 1])
        for ns, isbloch in itertools.product((-1, 1), (True, False)):
            D = naive_partial(N, nw, ns, dwinv, isbloch)
            C = O.create_partial(nw, ns == 1, N, dwinv, isbloch)
            assert np.array_equal(C.todense(), D)  # test/partial.jl:64
            Gv = np.empty_like(Fu)
            O.apply_partial(Gv, Fu, "=", nw, ns == 1, dwinv, isbloch)
            assert jl_isapprox(vec(Gv), D @ vec(Fu))  # test/partial.jl:67-70


def _curl_blk(dx, dy, dz, M):
    Z = np.zeros((M, M), dtype=dx.dtype)
    return np.block([[Z, -dz, dy], [dz, Z, -dx], [-dy, dx, Z]])


@pytest.mark.parametrize("isfwd", list(itertools.product((False, True), repeat=3)))
def test_create_curl_and_apply_curl(isfwd):
    """test/curl.jl:12-121."""
    rng = np.random.default_rng(2)
    N = (3, 4, 5)
    M = int(np.prod(N))
    r = blk2cmp_perm(M)
    F = np.asfortranarray(rng.random(N + (3,)) + 1j * rng.random(N + (3,)))

    # uniform grid, Bloch boundaries                                  test/curl.jl:14-31
    d = [O.create_partial(nw, isfwd[nw - 1], N).todense() for nw in (1, 2, 3)]
    Curl = O.create_curl(isfwd, None, N=N, order_cmpfirst=False).todense()
    assert Curl.shape == (3 * M, 3 * M)
    assert np.all(np.any(Curl != 0, axis=0)) and np.all(np.any(Curl != 0, axis=1))
    assert np.all(Curl.sum(axis=1) == 0)
    assert np.array_equal(Curl, _curl_blk(*d, M))

    # nonuniform grid, general boundaries                             test/curl.jl:33-57
    dlinv = [rng.random(n) for n in N]
    isbloch = [True, False, False]
    e = rng.random(3) + 1j * rng.random(3)
    d = [O.create_partial(nw, isfwd[nw - 1], N, dlinv[nw - 1], isbloch[nw - 1], e[nw - 1]).todense()
         for nw in (1, 2, 3)]
    Curl = O.create_curl(isfwd, dlinv, isbloch, e, order_cmpfirst=False).todense()
    assert np.array_equal(Curl, _curl_blk(*d, M))
    Curl_cmp = O.create_curl(isfwd, dlinv, isbloch, e, order_cmpfirst=True).todense()
    assert np.array_equal(Curl_cmp, Curl[np.ix_(r, r)])  # test/curl.jl:50-51
    G = np.empty_like(F)
    O.apply_curl(G, F, "=", isfwd, dlinv, isbloch, e)
    assert jl_isapprox(vec(G), Curl @ vec(F))

    # 1×1 curls in 1-D and 1×2 / 2×1 curls in 2-D                     test/curl.jl:62-119
    nw_blk = [[0, -3, 2], [3, 0, -1], [-2, 1, 0]]
    for cmp_shp in ([1], [2], [3], [1, 2], [2, 3], [3, 1]):
        if len(cmp_shp) == 1:
            cmp_outs = [[c] for c in (1, 2, 3) if c != cmp_shp[0]]
        else:
            cmp_outs = [cmp_shp, [c for c in (1, 2, 3) if c not in cmp_shp]]
        for cmp_out in cmp_outs:
            if len(cmp_shp) == 1:
                cmp_in = [6 - cmp_shp[0] - cmp_out[0]]
            else:
                cmp_in = [c for c in (1, 2, 3) if c not in cmp_out]
            Ncmp = tuple(N[c - 1] for c in cmp_shp)
            Mcmp = int(np.prod(Ncmp))
            sel = lambda v: [v[c - 1] for c in cmp_shp]
            Cs = O.create_curl(sel(isfwd), sel(dlinv), sel(isbloch), sel(e), cmp_shp=cmp_shp,
                               cmp_out=cmp_out, cmp_in=cmp_in, order_cmpfirst=False).todense()
            rows = []
            for nv in cmp_out:
                row = []
                for nu in cmp_in:
                    nw = nw_blk[nv - 1][nu - 1]
                    if nw == 0:
                        row.append(np.zeros((Mcmp, Mcmp), dtype=complex))
                    else:
                        sn, nw = np.sign(nw), abs(nw)
                        ind_nw = cmp_shp.index(nw) + 1
                        row.append(sn * O.create_partial(ind_nw, isfwd[nw - 1], Ncmp, dlinv[nw - 1],
                                                         isbloch[nw - 1], e[nw - 1]).todense())
                rows.append(row)
            assert np.array_equal(Cs, np.block(rows))
            Fc = np.asfortranarray(rng.random(Ncmp + (len(cmp_in),)) + 1j * rng.random(Ncmp + (len(cmp_in),)))
            Gc = np.empty(Ncmp + (len(cmp_out),), dtype=complex, order="F")
            O.apply_curl(Gc, Fc, "=", sel(isfwd), sel(dlinv), sel(isbloch), sel(e),
                         cmp_shp=cmp_shp, cmp_out=cmp_out, cmp_in=cmp_in)
            assert jl_isapprox(vec(Gc), Cs @ vec(Fc))


@pytest.mark.parametrize("isfwd", [[True, True, True], [True, False, True]])
def test_curl_of_curl(isfwd):
    """test/curl.jl:123-153 and :155-186."""
    N = (3, 4, 5)
    ones = [np.ones(n) for n in N]
    nisfwd = [not b for b in isfwd]
    isbloch = [True] * 3
    Cu = O.create_curl(isfwd, ones, isbloch, np.ones(3), order_cmpfirst=False).todense()
    Cv = O.create_curl(nisfwd, ones, isbloch, np.ones(3), order_cmpfirst=False).todense()
    A = Cv @ Cu
    assert np.all(np.diag(A) == 4)
    assert np.all((A != 0).sum(axis=1) == 13)
    assert np.array_equal(A, A.conj().T)
    B = A - 4 * np.eye(A.shape[0])
    assert np.all(np.abs(B[B != 0]) == 1)


@pytest.mark.parametrize("isfwd", list(itertools.product((False, True), repeat=2)))
def test_create_divg_and_apply_divg(isfwd):
    """test/divergence.jl:12-57."""
    rng = np.random.default_rng(3)
    N = (4, 5)
    M = int(np.prod(N))
    r = blk2cmp_perm(M, 2)
    F = np.asfortranarray(rng.random(N + (2,)) + 1j * rng.random(N + (2,)))
    Divg = O.create_divg(isfwd, None, N=N, order_cmpfirst=False).todense()
    assert Divg.shape == (M, 2 * M)
    assert np.all(Divg.sum(axis=0) == 0) and np.all(Divg.sum(axis=1) == 0)
    assert np.all(np.abs(Divg).sum(axis=0) == 2) and np.all(np.abs(Divg).sum(axis=1) == 4)
    d = [O.create_partial(nw, isfwd[nw - 1], N).todense() for nw in (1, 2)]
    assert np.array_equal(Divg, np.hstack(d))
    dlinv = [rng.random(n) for n in N]
    isbloch = [True, False]
    e = rng.random(2) + 1j * rng.random(2)
    Divg = O.create_divg(isfwd, dlinv, isbloch, e, order_cmpfirst=False).todense()
    d = [O.create_partial(nw, isfwd[nw - 1], N, dlinv[nw - 1], isbloch[nw - 1], e[nw - 1]).todense()
         for nw in (1, 2)]
    assert np.array_equal(Divg, np.hstack(d))
    Dc = O.create_divg(isfwd, dlinv, isbloch, e, order_cmpfirst=True).todense()
    assert np.array_equal(Dc, Divg[:, r])
    g = np.empty(N, dtype=complex, order="F")
    O.apply_divg(g, F, "=", isfwd, dlinv, isbloch, e)
    assert jl_isapprox(vec(g), Divg @ vec(F))
    assert jl_isapprox(vec(g), Dc @ vec(F)[r])


@pytest.mark.parametrize("isfwd", list(itertools.product((False, True), repeat=2)))
def test_create_grad_and_apply_grad(isfwd):
    """test/gradient.jl:12-57."""
    rng = np.random.default_rng(4)
    N = (4, 5)
    M = int(np.prod(N))
    r = blk2cmp_perm(M, 2)
    f = np.asfortranarray(rng.random(N) + 1j * rng.random(N))
    dlinv = [rng.random(n) for n in N]
    isbloch = [True, False]
    e = rng.random(2) + 1j * rng.random(2)
    Grad = O.create_grad(isfwd, dlinv, isbloch, e, order_cmpfirst=False).todense()
    d = [O.create_partial(nw, isfwd[nw - 1], N, dlinv[nw - 1], isbloch[nw - 1], e[nw - 1]).todense()
         for nw in (1, 2)]
    assert np.array_equal(Grad, np.vstack(d))
    Gc = O.create_grad(isfwd, dlinv, isbloch, e, order_cmpfirst=True).todense()
    assert np.array_equal(Gc, Grad[r, :])
    G = np.empty(N + (2,), dtype=complex, order="F")
    O.apply_grad(G, f, "=", isfwd, dlinv, isbloch, e)
    assert jl_isapprox(vec(G), Grad @ vec(f))
    assert jl_isapprox(vec(G)[r], Gc @ vec(f))


def test_divg_of_curl_and_curl_of_grad_vanish():
    """test/divergence.jl:59-74 and test/gradient.jl:59-74."""
    N = (3, 4, 5)
    isfwd = [True, True, True]
    Curl = O.create_curl(isfwd, None, N=N, order_cmpfirst=False).todense()
    Divg = O.create_divg(isfwd, None, N=N, order_cmpfirst=False).todense()
    Grad = O.create_grad(isfwd, None, N=N, order_cmpfirst=False).todense()
    assert np.all(Divg @ Curl == 0)
    assert np.all(Curl @ Grad == 0)


@pytest.mark.parametrize("N", [(10,), (8, 9), (3, 4, 5)])
def test_create_mean_and_apply_mean(N):
    """test/mean.jl:3-104 (1-D), :106-207 (2-D), :209-310 (3-D; reference uses (3,4,5))."""
    rng = np.random.default_rng(5)
    K = len(N)
    M = int(np.prod(N))
    Fu = np.asfortranarray(rng.random(N))
    F = np.asfortranarray(rng.random(N + (K,)) + 1j * rng.random(N + (K,)))
    for isfwd, isbloch in itertools.product((True, False), (True, False)):
        dl = [rng.random(n) for n in N]
        dlp = [rng.random(n) for n in N]
        Mdiag = np.zeros((K * M, K * M))
        for nw in range(1, K + 1):
            Mw = naive_mhat(N, nw, isfwd, dl[nw - 1], dlp[nw - 1], isbloch)
            C = O.create_mhat(nw, isfwd, N, dl[nw - 1], dlp[nw - 1], isbloch)
            assert np.array_equal(C.todense(), Mw)  # test/mean.jl:83
            Gv = np.empty_like(Fu)
            O.apply_mhat(Gv, Fu, "=", nw, isfwd, dl[nw - 1], dlp[nw - 1], isbloch)
            assert jl_isapprox(vec(Gv), Mw @ vec(Fu))  # :86-89
            Mdiag[(nw - 1) * M:nw * M, (nw - 1) * M:nw * M] = Mw
        Cm = O.create_mean([isfwd] * K, dl, dlp, [isbloch] * K, order_cmpfirst=False)
        assert np.array_equal(Cm.todense(), Mdiag)  # :97
        G = np.empty_like(F)
        O.apply_mean(G, F, "=", [isfwd] * K, dl, dlp, [isbloch] * K)
        assert jl_isapprox(vec(G), Mdiag @ vec(F))  # :100-102
        # arithmetic variant (not tested by the reference): equals weighted with ∆ = 1
        Ga = np.empty_like(F)
        O.apply_mean(Ga, F, "=", [isfwd] * K, None, None, [isbloch] * K)
        ones = [np.ones(n) for n in N]
        Gw = np.empty_like(F)
        O.apply_mean(Gw, F, "=", [isfwd] * K, ones, ones, [isbloch] * K)
        assert np.array_equal(Ga, Gw)


def test_create_picmp():
    """test/permutation.jl:7-35."""
    N = (3, 4)
    M = 12
    permute, scale = [3, 1, 2], [-1, 1, -1]
    P = O.create_picmp(N, permute, scale=scale, order_cmpfirst=False).todense()
    P2 = np.zeros((3 * M, 3 * M))
    for nw in range(1, 4):
        ib, jb = permute[nw - 1], nw
        P2[(ib - 1) * M:ib * M, (jb - 1) * M:jb * M] = scale[nw - 1] * np.eye(M)
    assert np.array_equal(P, P2)
    Pc = O.create_picmp(N, permute, scale=scale).todense()
    p = np.zeros((3, 3))
    for nw in range(1, 4):
        p[permute[nw - 1] - 1, nw - 1] = scale[nw - 1]
    assert np.array_equal(Pc, np.kron(np.eye(M), p))


# ---- gaps in the reference's suite that SURVEY.md §4 lists: covered here for the oracle ----
@pytest.mark.parametrize("N", [(7,), (5, 6), (4, 5, 6)])
def test_partial_with_bloch_phase_alpha_and_accumulate(N):
    rng = np.random.default_rng(6)
    M = int(np.prod(N))
    Fu = np.asfortranarray(rng.random(N) + 1j * rng.random(N))
    G0 = np.asfortranarray(rng.random(N) + 1j * rng.random(N))
    alpha = 0.7 - 0.2j
    for nw in range(1, len(N) + 1):
        dwinv = rng.random(N[nw - 1]) + 1j * rng.random(N[nw - 1])  # PML-stretched
        e = np.exp(-1j * 0.37 * nw)
        for ns in (-1, 1):
            D = naive_partial(N, nw, ns, dwinv, True, dtype=complex, e=e)
            C = O.create_partial(nw, ns == 1, N, dwinv, True, e)
            assert np.allclose(C.todense(), D, rtol=1e-14, atol=0)
            assert np.array_equal(C.todense() != 0, D != 0)
            G = G0.copy(order="F")
            O.apply_partial(G, Fu, "+=", nw, ns == 1, dwinv, True, e, alpha=alpha)
            assert jl_isapprox(vec(G), vec(G0) + alpha * (D @ vec(Fu)))


def test_csc_structure_is_canonical():
    """Rows strictly ascending within every column, no stored zeros, 1-based, Int64."""
    rng = np.random.default_rng(7)
    N = (4, 1, 5)  # includes a length-1 axis: diag and off-diag COO entries coincide
    dlinv = [rng.random(n) for n in N]
    for isbloch in ([True, True, True], [False, True, False]):
        C = O.create_curl([True, False, True], dlinv, isbloch, [1.0, 1.0, 1.0])
        assert C.colptr.dtype == np.int64 and C.rowval.dtype == np.int64
        assert C.colptr[0] == 1 and C.colptr[-1] == C.nnz + 1
        for c in range(C.n):
            rows = C.rowval[C.colptr[c] - 1:C.colptr[c + 1] - 1]
            assert np.all(np.diff(rows) > 0)
        assert np.all(C.nzval != 0)
    # Bloch with e = 1 along the length-1 axis: -∆ + ∆ == 0 exactly ⇒ the whole ∂y block drops
    C1 = O.create_partial(2, True, N, dlinv[1], True, 1.0)
    assert C1.nnz == 0


def test_nnz_counts_closed_form():
    """nnz(∂w) = 2M (Bloch) or 2M − 2M/Nw (symmetry)."""
    N = (4, 5, 6)
    M = 120
    for nw in (1, 2, 3):
        for isfwd in (True, False):
            assert O.create_partial(nw, isfwd, N, 1.0, True).nnz == 2 * M
            assert O.create_partial(nw, isfwd, N, 1.0, False).nnz == 2 * M - 2 * M // N[nw - 1]
            assert O.create_mhat(nw, isfwd, N, None, None, True).nnz == 2 * M
        assert O.create_mhat(nw, True, N, None, None, False).nnz == 2 * M - 2 * M // N[nw - 1]
        assert O.create_mhat(nw, False, N, None, None, False).nnz == 2 * M - M // N[nw - 1]


# ------------------------------------------------------------------------------------------------
# Julia Base scalar semantics outside /root/reference: complex division
# ------------------------------------------------------------------------------------------------
# The ten "hard" divisions of Baudin & Smith, "A robust complex division in Scilab" (arXiv:1210.4539), which Julia's own
# test-suite uses for `/(::ComplexF64, ::ComplexF64)` (`harddivs` in test/complex.jl): (z, w, z/w).
HARD_DIVISIONS = [
    ((1.0, 1.0), (1.0, 2.0 ** 1023), (2.0 ** -1023, -2.0 ** -1023)),
    ((1.0, 1.0), (2.0 ** -1023, 2.0 ** -1023), (2.0 ** 1023, 0.0)),
    ((2.0 ** 1023, 2.0 ** -1023), (2.0 ** 677, 2.0 ** -677), (2.0 ** 346, -2.0 ** -1008)),
    ((2.0 ** 1023, 2.0 ** 1023), (1.0, 1.0), (2.0 ** 1023, 0.0)),
    ((2.0 ** 1020, 2.0 ** -844), (2.0 ** 656, 2.0 ** -780), (2.0 ** 364, -2.0 ** -1072)),
    ((2.0 ** -71, 2.0 ** 1021), (2.0 ** 1001, 2.0 ** -323), (2.0 ** -1072, 2.0 ** 20)),
    ((2.0 ** -347, 2.0 ** -54), (2.0 ** -1037, 2.0 ** -1058), (3.898125604559113300e289, 8.174961907852353577e295)),
    ((2.0 ** -1074, 2.0 ** -1074), (2.0 ** -1073, 2.0 ** -1074), (0.6, 0.2)),
    ((2.0 ** 1015, 2.0 ** -989), (2.0 ** 1023, 2.0 ** 1023), (0.001953125, -0.001953125)),
    ((2.0 ** -622, 2.0 ** -1071), (2.0 ** -343, 2.0 ** -798), (1.02951151789360578e-84, 6.97145987515076231e-220)),
]


def _ulps(a, b):
    return 0.0 if a == b else abs(a - b) / np.spacing(abs(b))


def test_restated_julia_complex_division_on_published_hard_cases():
    """The oracle's restatement of Base's robust division reproduces the published known answers (nine exactly, the
    denormal case #8 to one ulp — the accuracy the algorithm's authors report for it)."""
    for n, (z, w, expect) in enumerate(HARD_DIVISIONS, 1):
        r = O._robust_cdiv(O.JV(np.array([z[0]]), np.array([z[1]])), O.JV(np.array([w[0]]), np.array([w[1]])))
        err = max(_ulps(float(r.re[0]), expect[0]), _ulps(float(r.im[0]), expect[1]))
        assert err <= (1.0 if n == 8 else 0.0), f"hard division #{n}: {float(r.re[0])!r}, {float(r.im[0])!r} vs {expect} ({err} ulp)"


def test_library_complex_division_equals_the_oracle_bit_for_bit():
    """The CUDA library's division routine (host evaluation of the very source the kernels compile) against the oracle:
    the hard cases and 10^5 random operand pairs over the whole exponent range, bit for bit."""
    import ctypes as C
    sgc = pytest.importorskip("sgc_b200")
    lib = sgc._lib.lib
    rng = np.random.default_rng(99)
    n = 100000
    def rnd():
        return rng.uniform(-1, 1, n) * 2.0 ** rng.integers(-1000, 1000, n)
    z = np.stack([rnd(), rnd()], axis=1)
    w = np.stack([rnd(), rnd()], axis=1)
    zh = np.array([[a[0], a[1]] for a, _, _ in HARD_DIVISIONS]); wh = np.array([[b[0], b[1]] for _, b, _ in HARD_DIVISIONS])
    z, w = np.ascontiguousarray(np.vstack([zh, z])), np.ascontiguousarray(np.vstack([wh, w]))
    out = np.empty_like(z)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    assert lib.sgc_debug_cdiv(dp(z), dp(w), dp(out), len(z)) == 0
    with np.errstate(all="ignore"):
        r = O._robust_cdiv(O.JV(z[:, 0].copy(), z[:, 1].copy()), O.JV(w[:, 0].copy(), w[:, 1].copy()))
    assert np.array_equal(out[:, 0], np.asarray(r.re), equal_nan=True)
    assert np.array_equal(out[:, 1], np.asarray(r.im), equal_nan=True)


def test_library_complex_inverse_equals_the_oracle_bit_for_bit():
    """`inv(e⁻ⁱᵏᴸ)` (matrix/differential/partial.jl:217 for backward blocks): the library's host routine against the
    oracle's restatement of Base's robust_cinv on 10^5 random arguments over the whole exponent range."""
    import ctypes as C
    sgc = pytest.importorskip("sgc_b200")
    lib = sgc._lib.lib
    rng = np.random.default_rng(98)
    n = 100000
    w = np.ascontiguousarray(np.stack([rng.uniform(-1, 1, n) * 2.0 ** rng.integers(-1020, 1020, n),
                                       rng.uniform(-1, 1, n) * 2.0 ** rng.integers(-1020, 1020, n)], axis=1))
    out = np.empty_like(w)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    assert lib.sgc_debug_cinv(dp(w), dp(out), n) == 0
    with np.errstate(all="ignore"):
        r = O.jl_inv(O.JV(w[:, 0].copy(), w[:, 1].copy()))
    assert np.array_equal(out[:, 0], np.asarray(r.re), equal_nan=True)
    assert np.array_equal(out[:, 1], np.asarray(r.im), equal_nan=True)
