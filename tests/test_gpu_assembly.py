"""GPU parity tests of the sparse assembly: CUDA library (through the C ABI) vs oracle.

`colptr` / `rowval` must match the oracle BIT-EXACTLY (north star); `nzval` is expected to be
bit-identical too (same IEEE operations), with 4 ulp / 1e-13 as the hard limit.
"""
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


def assert_csc_equal(dev_csc, ref: O.CSC, what, exact_vals=True):
    cp, rv, nz = dev_csc.to_host()
    assert (dev_csc.m, dev_csc.n) == (ref.m, ref.n), f"{what}: shape"
    assert cp.dtype == np.int64 and rv.dtype == np.int64
    assert np.array_equal(cp, ref.colptr), f"{what}: colptr differs"
    assert np.array_equal(rv, ref.rowval), f"{what}: rowval differs"
    assert nz.dtype == ref.nzval.dtype, f"{what}: eltype {nz.dtype} vs {ref.nzval.dtype}"
    assert_parity(nz, ref.nzval, f"{what}: nzval", exact=exact_vals)


@pytest.mark.parametrize("N", [(10,), (8, 9), (8, 9, 10), (1, 5, 3), (4, 1, 1)])
def test_create_partial_matches_oracle(S, N):
    """test/partial.jl:64 (exact `==` against the naive matrix, via the pinned oracle)."""
    rng = np.random.default_rng(10)
    for nw in range(1, len(N) + 1):
        for isfwd, isbloch in itertools.product((True, False), (True, False)):
            for dwinv, e in [(rand_coef(rng, N[nw - 1], False), 1.0),
                             (rand_coef(rng, N[nw - 1], False), np.exp(-0.3j)),
                             (rand_coef(rng, N[nw - 1], True), np.exp(0.8j)),
                             (2.0, -1.0)]:
                ref = O.create_partial(nw, isfwd, N, dwinv, isbloch, e)
                for method in ("direct", "counted", "twostep", "coo"):
                    got = S.create_partial(nw, isfwd, N, dwinv, isbloch, e, method=method)
                    assert_csc_equal(got, ref, f"create_∂ N={N} nw={nw} fwd={isfwd} bloch={isbloch} {method}")


@pytest.mark.parametrize("N", [(10,), (8, 9), (3, 4, 5), (1, 6, 2)])
def test_create_mhat_and_mean_match_oracle(S, N):
    """test/mean.jl:83, :97."""
    rng = np.random.default_rng(11)
    K = len(N)
    for isfwd, isbloch in itertools.product((True, False), (True, False)):
        dl = [rand_coef(rng, n, False) for n in N]
        dlp = [rand_coef(rng, n, False) for n in N]
        for nw in range(1, K + 1):
            for e in (1.0, np.exp(-0.6j)):
                ref = O.create_mhat(nw, isfwd, N, dl[nw - 1], dlp[nw - 1], isbloch, e)
                got = S.create_mhat(nw, isfwd, N, dl[nw - 1], dlp[nw - 1], isbloch, e)
                assert_csc_equal(got, ref, f"create_m̂ N={N} nw={nw} fwd={isfwd} bloch={isbloch}")
            ref = O.create_mhat(nw, isfwd, N, None, None, isbloch)
            got = S.create_mhat(nw, isfwd, N, None, None, isbloch, method="coo")
            assert_csc_equal(got, ref, f"create_m̂ arithmetic N={N} nw={nw}")
        for cmpfirst in (True, False):
            ref = O.create_mean([isfwd] * K, dl, dlp, [isbloch] * K, order_cmpfirst=cmpfirst)
            got = S.create_mean([isfwd] * K, dl, dlp, [isbloch] * K, order_cmpfirst=cmpfirst)
            assert_csc_equal(got, ref, f"create_mean N={N} fwd={isfwd} bloch={isbloch} cmpfirst={cmpfirst}")
        ref = O.create_mean([isfwd] * K, None, None, [isbloch] * K, N=N)
        got = S.create_mean([isfwd] * K, None, None, [isbloch] * K, N=N)
        assert_csc_equal(got, ref, f"create_mean arithmetic N={N}")


@pytest.mark.parametrize("N", [(3, 4, 5), (9, 7, 6), (1, 4, 3), (5, 1, 1)])
@pytest.mark.parametrize("isfwd", list(itertools.product((False, True), repeat=3)))
def test_create_curl_matches_oracle(S, N, isfwd):
    """test/curl.jl:31, :47, :51 — uniform/Bloch, nonuniform/general BCs, both orderings."""
    rng = np.random.default_rng(12)
    dlinv = [rand_coef(rng, n, False) for n in N]
    dlinv_c = [rand_coef(rng, n, True) for n in N]
    e = rng.uniform(-1, 1, 3) + 1j * rng.uniform(-1, 1, 3)
    cases = [(None, None, None), (dlinv, [True, False, False], e), (dlinv, [True, True, True], [1.0, 1.0, 1.0]),
             (dlinv_c, [False, True, False], e), (dlinv, [False, False, False], [1.0, 1.0, 1.0])]
    for dl, isbloch, ee in cases:
        for cmpfirst in (True, False):
            ref = O.create_curl(isfwd, dl, isbloch, ee, N=N if dl is None else None, order_cmpfirst=cmpfirst)
            for method in ("direct", "general", "cta", "counted", "twostep", "coo"):
                got = S.create_curl(isfwd, dl, isbloch, ee, N=N, order_cmpfirst=cmpfirst, method=method)
                assert_csc_equal(got, ref, f"create_curl N={N} fwd={isfwd} bloch={isbloch} cmpfirst={cmpfirst} {method}")


def test_create_subcurls_divg_grad_picmp_match_oracle(S):
    """test/curl.jl:62-107, test/divergence.jl:29-45, test/gradient.jl:29-45, test/permutation.jl."""
    rng = np.random.default_rng(13)
    N = (3, 4, 5)
    isfwd = [True, False, True]
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
            sel = lambda v: [v[c - 1] for c in cmp_shp]
            for cmpfirst in (True, False):
                ref = O.create_curl(sel(isfwd), sel(dlinv), sel(isbloch), sel(e), cmp_shp=cmp_shp, cmp_out=cmp_out,
                                    cmp_in=cmp_in, order_cmpfirst=cmpfirst)
                got = S.create_curl(sel(isfwd), sel(dlinv), sel(isbloch), sel(e), cmp_shp=cmp_shp, cmp_out=cmp_out,
                                    cmp_in=cmp_in, order_cmpfirst=cmpfirst)
                assert_csc_equal(got, ref, f"subcurl shp={cmp_shp} out={cmp_out} in={cmp_in}")
    for K, Nk in ((1, (6,)), (2, (4, 5)), (3, (3, 4, 5))):
        for fw in itertools.product((False, True), repeat=K):
            dl = [rand_coef(rng, n, False) for n in Nk]
            ib = ([True, False, False])[:K]
            ee = e[:K]
            for cmpfirst in (True, False):
                for method in ("direct", "counted", "twostep", "coo"):
                    assert_csc_equal(S.create_divg(fw, dl, ib, ee, order_cmpfirst=cmpfirst, method=method),
                                     O.create_divg(fw, dl, ib, ee, order_cmpfirst=cmpfirst), f"divg K={K} {fw} {method}")
                    assert_csc_equal(S.create_grad(fw, dl, ib, ee, order_cmpfirst=cmpfirst, method=method),
                                     O.create_grad(fw, dl, ib, ee, order_cmpfirst=cmpfirst), f"grad K={K} {fw} {method}")
    for cmpfirst in (True, False):
        for scale in ([-1, 1, -1], [0.5, 0.0, 2.0], [1j, 2.0, -1.0]):
            for method in ("direct", "counted", "twostep", "coo"):
                assert_csc_equal(S.create_picmp((3, 4), [3, 1, 2], scale=scale, order_cmpfirst=cmpfirst, method=method),
                                 O.create_picmp((3, 4), [3, 1, 2], scale=scale, order_cmpfirst=cmpfirst),
                                 f"πcmp cmpfirst={cmpfirst} scale={scale} {method}")


@pytest.mark.parametrize("N", [(140, 9, 7), (33, 40, 5), (130, 3, 3), (64, 6, 5)])
def test_interior_descriptor_path_equals_general_path(S, N):
    """CTAs whose columns are all full emit interior cells from host-sorted descriptors; cells on a grid face and every
    other CTA walk the blocks and sort.  Both must give the same matrix (and the oracle's) for every operator shape."""
    rng = np.random.default_rng(77)
    dl = [rand_coef(rng, n, True) for n in N]
    e = np.exp(-1j * rng.uniform(0, 3, 3))
    for isfwd in ([True, False, True], [False, True, False]):
        for isbloch in ([True] * 3, [True, False, True]):
            for cmpfirst in (True, False):
                ref = O.create_curl(isfwd, dl, isbloch, e, order_cmpfirst=cmpfirst)
                for method in ("direct", "general", "cta", "onecol"):
                    got = S.create_curl(isfwd, dl, isbloch, e, N=N, order_cmpfirst=cmpfirst, method=method)
                    assert_csc_equal(got, ref, f"create_curl N={N} fwd={isfwd} bloch={isbloch} cmpfirst={cmpfirst} {method}")
            for name, mk, mo in (("divg", S.create_divg, O.create_divg), ("grad", S.create_grad, O.create_grad)):
                for cmpfirst in (True, False):
                    ref = mo(isfwd, dl, isbloch, e, order_cmpfirst=cmpfirst)
                    for method in ("direct", "general", "cta", "onecol"):
                        got = mk(isfwd, dl, isbloch, e, order_cmpfirst=cmpfirst, method=method)
                        assert_csc_equal(got, ref, f"create_{name} N={N} fwd={isfwd} bloch={isbloch} cmpfirst={cmpfirst} {method}")
            ref = O.create_mean(isfwd, dl, dl, isbloch, e)
            for method in ("direct", "general", "cta", "onecol"):
                assert_csc_equal(S.create_mean(isfwd, dl, dl, isbloch, e, method=method), ref, f"create_mean N={N} {method}")
            for nw in (1, 2, 3):
                ref = O.create_partial(nw, isfwd[nw - 1], N, dl[nw - 1], isbloch[nw - 1], e[nw - 1])
                for method in ("direct", "general", "cta", "onecol"):
                    assert_csc_equal(S.create_partial(nw, isfwd[nw - 1], N, dl[nw - 1], isbloch[nw - 1], e[nw - 1], method=method), ref,
                                     f"create_∂ N={N} nw={nw} {method}")
                ref = O.create_mhat(nw, isfwd[nw - 1], N, dl[nw - 1], dl[nw - 1], isbloch[nw - 1], e[nw - 1])
                for method in ("direct", "general", "cta", "onecol"):
                    assert_csc_equal(S.create_mhat(nw, isfwd[nw - 1], N, dl[nw - 1], dl[nw - 1], isbloch[nw - 1], e[nw - 1], method=method), ref,
                                     f"create_m̂ N={N} nw={nw} {method}")


def test_zero_coefficients_are_dropped(S):
    """`dropzeros!` is value based: zeros in ∆w⁻¹ and e = 1 on a length-1 axis vanish."""
    N = (5, 1, 4)
    dl = [np.array([1.0, 0.0, 2.0, 0.0, 3.0]), np.array([0.7]), np.array([1.0, -0.0, 1.0, 1.0])]
    for isfwd in ([True, True, True], [False, True, False]):
        ref = O.create_curl(isfwd, dl, [True, True, True], [1.0, 1.0, 1.0])
        got = S.create_curl(isfwd, dl, [True, True, True], [1.0, 1.0, 1.0])
        assert_csc_equal(got, ref, f"zeros fwd={isfwd}")
        got = S.create_curl(isfwd, dl, [True, True, True], [1.0, 1.0, 1.0], method="coo")
        assert_csc_equal(got, ref, f"zeros coo fwd={isfwd}")


def test_row_partition_concatenates_to_global(S):
    """per-rank row blocks stacked vertically equal the global matrix."""
    import scipy.sparse as sp
    rng = np.random.default_rng(14)
    N = (6, 5, 4)
    M = 120
    dlinv = [rand_coef(rng, n, False) for n in N]
    e = [np.exp(-0.1j), np.exp(-0.2j), np.exp(-0.3j)]
    for cmpfirst in (True, False):
        for isbloch in ([True, True, True], [False, True, False]):
            full = S.create_curl([True, False, True], dlinv, isbloch, e, order_cmpfirst=cmpfirst, index_base=0).to_scipy()
            for nparts in (2, 3, 8):
                cuts = [round(i * 3 * M / nparts) for i in range(nparts + 1)]
                blocks = []
                for a, b in zip(cuts[:-1], cuts[1:]):
                    blk = S.create_curl([True, False, True], dlinv, isbloch, e, order_cmpfirst=cmpfirst, index_base=0,
                                        row_range=(a, b))
                    assert blk.m == b - a and blk.n == 3 * M
                    cp, rv, _ = blk.to_host()
                    for c in range(0, blk.n, 7):
                        assert np.all(np.diff(rv[cp[c]:cp[c + 1]]) > 0)
                    blocks.append(blk.to_scipy())
                stacked = sp.vstack(blocks).tocsc()
                stacked.sort_indices()
                assert np.array_equal(stacked.indptr, full.indptr)
                assert np.array_equal(stacked.indices, full.indices)
                assert np.array_equal(stacked.data, full.data)


def test_row_partition_on_a_tall_grid_skips_far_columns(S):
    """Per-rank blocks of a grid with many z-planes: most column CTAs of a rank hold no entry of its rows
    and are skipped; the blocks must still stack to the global matrix (all-Bloch incl. the z wrap-around,
    and a symmetric case), for curl, divergence and gradient."""
    import scipy.sparse as sp
    rng = np.random.default_rng(16)
    N = (8, 4, 64)
    M = int(np.prod(N))
    dlinv = [rand_coef(rng, n, True) for n in N]
    e = [np.exp(-0.1j), np.exp(-0.2j), np.exp(-0.3j)]
    makers = [("curl", lambda **kw: S.create_curl([True, False, True], dlinv, kw.pop("isbloch"), e, index_base=0, **kw), 3 * M),
              ("divg", lambda **kw: S.create_divg([False, True, True], dlinv, kw.pop("isbloch"), e, index_base=0, **kw), M),
              ("grad", lambda **kw: S.create_grad([True, True, False], dlinv, kw.pop("isbloch"), e, index_base=0, **kw), 3 * M)]
    for name, make, nrows in makers:
        for isbloch in ([True, True, True], [True, False, False]):
            full = make(isbloch=isbloch).to_scipy()
            for nparts in (4, 7):
                cuts = [round(i * nrows / nparts) for i in range(nparts + 1)]
                blocks = [make(isbloch=isbloch, row_range=(a, b)).to_scipy() for a, b in zip(cuts[:-1], cuts[1:])]
                stacked = sp.vstack(blocks).tocsc()
                stacked.sort_indices()
                what = f"{name} bloch={isbloch} parts={nparts}"
                assert np.array_equal(stacked.indptr, full.indptr), what
                assert np.array_equal(stacked.indices, full.indices), what
                assert np.array_equal(stacked.data, full.data), what


def test_column_partition_blocks_concatenate_to_the_global_arrays(S):
    """The scalable multi-GPU form (`sgc_asm_set_col_range`): the block of columns of part p is a stand-alone CSC
    matrix with global row indices, and the GLOBAL rowval / nzval are the plain concatenation of the blocks' arrays
    (colptr: each block's shifted by the entries before it) — bit for bit, both orderings, Bloch and symmetric
    boundaries, real and complex values, curl / divergence / gradient / mean, 128³ included."""
    rng = np.random.default_rng(18)
    cases = []
    for N in ((40, 12, 9), (128, 128, 128)):
        dl = [rand_coef(rng, n, False) for n in N]
        dlc = [rand_coef(rng, n, True) for n in N]
        e = [np.exp(-0.1j), np.exp(-0.2j), np.exp(-0.3j)]
        big = N[0] == 128
        cases.append((N, "curl f64", lambda N=N, dl=dl, **kw: S.create_curl([True, False, True], dl, [True, True, True], [1.0] * 3, **kw)))
        if big:
            continue
        cases.append((N, "curl c128 mixed BC", lambda N=N, dlc=dlc, e=e, **kw: S.create_curl([True, False, True], dlc, [False, True, False], e, **kw)))
        cases.append((N, "curl block ordering", lambda N=N, dl=dl, e=e, **kw: S.create_curl([False, True, True], dl, [True, False, True], e, order_cmpfirst=False, **kw)))
        cases.append((N, "divg", lambda N=N, dl=dl, e=e, **kw: S.create_divg([False, True, True], dl, [True, False, True], e, **kw)))
        cases.append((N, "grad", lambda N=N, dl=dl, e=e, **kw: S.create_grad([True, True, False], dl, [False, True, True], e, **kw)))
        cases.append((N, "mean", lambda N=N, dl=dl, **kw: S.create_mean([True, False, True], dl, dl, [True, False, True], **kw)))
    for N, name, make in cases:
        full = make()
        cp, rv, nz = full.to_host()
        for nparts in ((8,) if N[0] == 128 else (1, 2, 3, 8)):
            cps, rvs, nzs, ncols, off = [], [], [], 0, 0
            for part in range(nparts):
                blk = make(row_range={"cols": (part, nparts)})
                assert blk.m == full.m
                bcp, brv, bnz = blk.to_host()
                assert bcp[0] == 1 and bcp[-1] == 1 + len(brv)
                cps.append(bcp[:-1] + off)
                rvs.append(brv); nzs.append(bnz)
                off += len(brv)
                ncols += blk.n
            what = f"{name} {N} in {nparts} column blocks"
            assert ncols == full.n, what
            assert np.array_equal(np.concatenate(cps + [np.array([1 + off])]), cp), what + ": colptr"
            assert np.array_equal(np.concatenate(rvs), rv), what + ": rowval"
            assert np.array_equal(np.concatenate(nzs), nz), what + ": nzval"


@pytest.mark.parametrize("isfwd", [[True, True, True], [True, False, True]])
def test_curl_of_curl_invariants_with_device_matrices(S, isfwd):
    """test/curl.jl:123-186 with the matrices assembled on the GPU: `Cv*Cu` has diagonal 4, 13 nonzeros per row, is
    Hermitian with ±1 off-diagonals; each block of Cu equals −(block of Cv)ᴴ; and `Cv*(Cu*f)` equals the fused
    curl-of-curl kernel applied to f (block ordering)."""
    import torch
    N = (3, 4, 5)
    M = 60
    ones = [np.ones(n) for n in N]
    nisfwd = [not b for b in isfwd]
    Cu = S.create_curl(isfwd, ones, [True, False, False], np.ones(3), order_cmpfirst=False, index_base=0).to_scipy().toarray()
    Cv = S.create_curl(nisfwd, ones, [True, False, False], np.ones(3), order_cmpfirst=False, index_base=0).to_scipy().toarray()
    for i in range(3):
        for j in ((i + 1) % 3, (i + 2) % 3):
            assert np.array_equal(-Cv[i * M:(i + 1) * M, j * M:(j + 1) * M].conj().T, Cu[i * M:(i + 1) * M, j * M:(j + 1) * M])
    Au = S.create_curl(isfwd, ones, [True] * 3, np.ones(3), order_cmpfirst=False)
    Av = S.create_curl(nisfwd, ones, [True] * 3, np.ones(3), order_cmpfirst=False)
    A = Av.to_scipy().toarray() @ Au.to_scipy().toarray()
    assert np.all(np.diag(A) == 4)
    assert np.all((A != 0).sum(axis=1) == 13)
    assert np.array_equal(A, A.conj().T)
    B = A - 4 * np.eye(A.shape[0])
    assert np.all(np.abs(B[B != 0]) == 1)
    rng = np.random.default_rng(18)
    f = rand_field(rng, N + (3,), False)
    fd = S.DeviceArray.from_numpy(f)
    g = S.DeviceArray.zeros(N + (3,), np.float64)
    S.apply_curlcurl(g, fd, "=", isfwd, ones, nisfwd, ones, [True] * 3, [1.0] * 3)
    via_matrices = Av.matvec(Au.matvec(fd.t))
    assert torch.allclose(via_matrices, g.t, rtol=1e-13, atol=1e-13)


def test_csc_hand_off_formats(S, tmp_path):
    """pinned host arrays for the `SparseMatrixCSC(m,n,colptr,rowval,nzval)` constructor and the
    on-disk `.npz` dump, both equal to the oracle's matrix."""
    from sgc_b200.matrix import load_npz_csc
    rng = np.random.default_rng(17)
    N = (5, 4, 3)
    dlinv = [rand_coef(rng, n, True) for n in N]
    e = [np.exp(-0.1j), 1.0, np.exp(0.3j)]
    A = S.create_curl([True, False, True], dlinv, [True, False, True], e)
    R = O.create_curl([True, False, True], dlinv, [True, False, True], e)
    cp, rv, nz = A.to_pinned_host()
    assert cp.is_pinned() and rv.is_pinned() and nz.is_pinned()
    assert np.array_equal(cp.numpy(), R.colptr) and np.array_equal(rv.numpy(), R.rowval) and np.array_equal(nz.numpy(), R.nzval)
    A.save_npz(tmp_path / "curl.npz")
    B = load_npz_csc(tmp_path / "curl.npz")
    assert np.array_equal(B.indptr, R.colptr - 1) and np.array_equal(B.indices, R.rowval - 1) and np.array_equal(B.data, R.nzval)


def test_matrix_times_field_equals_matrix_free(S):
    """test/curl.jl:54-57: `mul!(g, Curl, f)` ≈ `apply_curl!` (block ordering), on the device."""
    import torch
    rng = np.random.default_rng(15)
    N = (12, 10, 9)
    F = rand_field(rng, N + (3,), True)
    dlinv = [rand_coef(rng, n, True) for n in N]
    isbloch = [True, False, True]
    e = [np.exp(-0.4j), 1.0, np.exp(0.6j)]
    for isfwd in ([True, True, True], [False, True, False]):
        A = S.create_curl(isfwd, dlinv, isbloch, e, order_cmpfirst=False)
        Fd = S.DeviceArray.from_numpy(F)
        g = A.matvec(Fd.t)
        G = S.DeviceArray.empty(N + (3,))
        S.apply_curl(G, Fd, "=", isfwd, dlinv, isbloch, e)
        assert torch.allclose(g, G.t, rtol=1e-12, atol=1e-12)


def test_full_size_curl_assembly_properties(S):
    """BASELINE config 4 structure at 256³ (402 M nonzeros at 512³ is exercised by bench.py):
    closed-form nnz, 4 entries per column, strictly ascending rows, zero row sums, ±∆⁻¹ values,
    and equality of the component-major matrix with the permuted block-ordered one."""
    import torch
    N = (256, 256, 256)
    M = N[0] * N[1] * N[2]
    A = S.create_curl([True, True, True], None, N=N, order_cmpfirst=True, index_base=1)
    assert A.nzval.dtype == torch.float64  # real inputs → real matrix (curl.jl:14-21)
    assert A.nnz == 12 * M
    assert int(A.colptr[0]) == 1 and int(A.colptr[-1]) == 12 * M + 1
    assert bool((A.colptr[1:] - A.colptr[:-1] == 4).all())
    rv = A.rowval.view(-1, 4)
    assert bool((rv[:, 1:] > rv[:, :-1]).all())
    assert bool((A.nzval.abs() == 1).all())
    ones = torch.ones(3 * M, dtype=torch.float64, device="cuda")
    assert float(A.matvec(ones).abs().max()) == 0.0  # Curl * ones == 0   (test/curl.jl:29)
    # symmetric BCs drop 2M/Nw entries per ∂ block
    B = S.create_curl([True, True, True], None, [True, False, False], None, N=N)
    assert B.nnz == 12 * M - 2 * (2 * M // N[1]) - 2 * (2 * M // N[2])
    # Curl_cmpfirst == Curl[r, r]   (test/curl.jl:50-51) checked through A·x on a random x
    x = torch.rand(3 * M, dtype=torch.float64, device="cuda")
    Ab = S.create_curl([True, True, True], None, N=N, order_cmpfirst=False)
    r = torch.arange(3 * M, device="cuda").view(3, M).t().reshape(-1)  # cmp-major position → block index
    yb = Ab.matvec(x)
    yc = A.matvec(x[r])
    assert torch.allclose(yc, yb[r], rtol=0, atol=1e-12)


def test_spmv_plan_matches_the_matrix_free_operators(S):
    """`mul!(g, Curl, f)` against `apply_curl!` — the reference's own cross-check (test/curl.jl:54-57, ≈) — through the
    row-sorted SpMV plan: equal within the north-star tolerance (the two sides group the same products differently),
    and bit-identical to a host evaluation that accumulates each row in ascending column order; also ∂, divergence,
    gradient, and a matrix with empty rows (symmetric boundary)."""
    import torch
    rng = np.random.default_rng(19)
    N = (24, 17, 11)
    M = int(np.prod(N))
    dl = [rand_coef(rng, n, True) for n in N]
    e = [np.exp(-0.3j), np.exp(0.5j), np.exp(-0.9j)]
    for isbloch in ([True, True, True], [False, True, False]):
        isfwd = [True, False, True]
        A = S.create_curl(isfwd, dl, isbloch, e, order_cmpfirst=False)
        F = rand_field(rng, N + (3,), True)
        x = torch.from_numpy(F.reshape(-1, order="F").copy()).cuda()
        y = A.mul(x).cpu().numpy()
        G = np.zeros_like(F)
        O.apply_curl(G, F, "=", isfwd, dl, isbloch, e)
        scale = np.abs(G).max()
        assert np.abs(y - G.reshape(-1, order="F")).max() <= 64 * np.finfo(float).eps * scale, "mul!(g, Curl, f) vs apply_curl!"
        # exact: rows accumulated in ascending column order from zero
        sp = A.to_scipy().tocsr()
        sp.sort_indices()
        xs = x.cpu().numpy()
        ref = np.zeros(3 * M, dtype=complex)
        for r in range(0, 3 * M, 97):
            s = 0j
            for q in range(sp.indptr[r], sp.indptr[r + 1]):
                a, b = sp.data[q], xs[sp.indices[q]]
                s = complex(s.real + (a.real * b.real - a.imag * b.imag), s.imag + (a.real * b.imag + a.imag * b.real))
            ref[r] = s
        assert np.array_equal(y[::97], ref[::97]), "row sums are not accumulated in ascending column order"
        assert np.allclose(A.matvec(x).cpu().numpy(), y, rtol=1e-13, atol=1e-13 * scale)  # the column-parallel utility agrees
    # real matrices: ∂, divergence (rectangular M x 3M), gradient (3M x M)
    dr = [rand_coef(rng, n, False) for n in N]
    f = rand_field(rng, N, False)
    F3 = rand_field(rng, N + (3,), False)
    A = S.create_divg([True, True, False], dr, [True, False, True])
    g = np.zeros(N, order="F")
    O.apply_divg(g, F3, "=", [True, True, False], dr, [True, False, True], [1.0] * 3)
    # the matrix uses component-major column ordering (order_cmpfirst=true): x[3·cell + w]
    xin = np.ascontiguousarray(np.moveaxis(F3, -1, 0).reshape(3, -1, order="F").T).reshape(-1)
    y = A.mul(torch.from_numpy(xin).cuda()).cpu().numpy()
    assert np.abs(y - g.reshape(-1, order="F")).max() <= 64 * np.finfo(float).eps * np.abs(g).max()
    A = S.create_grad([False, True, True], dr, [True, True, False], order_cmpfirst=False)
    Gr = np.zeros(N + (3,), order="F")
    O.apply_grad(Gr, f, "=", [False, True, True], dr, [True, True, False], [1.0] * 3)
    y = A.mul(torch.from_numpy(f.reshape(-1, order="F").copy()).cuda()).cpu().numpy()
    assert np.abs(y - Gr.reshape(-1, order="F")).max() <= 64 * np.finfo(float).eps * np.abs(Gr).max()
