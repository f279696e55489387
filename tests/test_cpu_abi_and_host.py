"""CPU-only checks: the C-ABI library loads and exports every declared symbol, the host logic
rejects bad arguments, the two oracles agree bit-for-bit, and the input producers (Grid / PML)
reproduce the reference's closed forms.  No GPU compute is attempted here."""
import ctypes
import itertools
import os
import re

import numpy as np
import pytest

from oracle import sgc_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_symbol_the_header_declares():
    import sgc_b200
    hdr = open(os.path.join(ROOT, "include", "sgc_b200.h"), encoding="utf-8").read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(sgc_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 45
    lib = ctypes.CDLL(sgc_b200.LIB_PATH)
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, f"libsgc_b200.so lacks {missing}"
    assert sorted(sgc_b200._lib._PROTOS) == declared, "the ctypes binding and the header disagree"
    assert lib.sgc_abi_version() == 2


def test_no_cpu_fallback_product_path_fails_loudly_without_gpu():
    import torch
    import sgc_b200
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(sgc_b200.SgcError):
        sgc_b200.DeviceArray.zeros((4, 4))
    with pytest.raises(sgc_b200.SgcError):
        sgc_b200.create_curl([True] * 3, None, N=(4, 4, 4))
    # the product package never imports the oracle
    import sys
    pkg = os.path.join(ROOT, "staggeredgridcalculus.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                src = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert "import oracle" not in src and "from oracle" not in src and "sgc_oracle" not in src, f


def test_host_argument_checks_do_not_need_a_gpu():
    """Creation-time validation (the reference's @assert / @error) happens before any CUDA call
    that needs a device."""
    import sgc_b200
    from sgc_b200 import _lib
    from sgc_b200.device import int64s, scalar2
    h = ctypes.c_void_p()
    N = int64s((4, 5))
    one = scalar2(1.0)
    assert _lib.lib.sgc_op_create_partial(ctypes.byref(h), 0, 2, N, 3, 1, None, 0, 1, one, one) == _lib.SGC_ERR_INVALID
    assert b"nw" in _lib.lib.sgc_last_error()
    assert _lib.lib.sgc_op_create_partial(ctypes.byref(h), 7, 2, N, 1, 1, None, 0, 1, one, one) == _lib.SGC_ERR_INVALID
    assert _lib.lib.sgc_op_create_partial(ctypes.byref(h), 0, 2, N, 1, 1, None, 0, 1, scalar2(1j), one) == _lib.SGC_ERR_INEXACT
    assert _lib.lib.sgc_op_create_partial(ctypes.byref(h), 0, 2, int64s((1, 5)), 1, 1, None, 0, 0, one, one) == _lib.SGC_ERR_INVALID


@pytest.mark.parametrize("isfwd", list(itertools.product((False, True), repeat=3)))
def test_c_oracle_equals_numpy_oracle_bit_for_bit(isfwd):
    from oracle import c_oracle as CO
    rng = np.random.default_rng(42)
    N = (7, 6, 5)
    F = np.asfortranarray(rng.uniform(-1, 1, N + (3,)) + 1j * rng.uniform(-1, 1, N + (3,)))
    G0 = np.asfortranarray(rng.uniform(-1, 1, N + (3,)) + 1j * rng.uniform(-1, 1, N + (3,)))
    dl = [rng.uniform(0.5, 1.5, n) + 1j * rng.uniform(-0.3, 0.3, n) for n in N]
    e = np.exp(-1j * np.array([0.3, 0.7, 1.1]))
    for isbloch in ([True, True, True], [False, True, False], [False, False, False]):
        for op in ("=", "+="):
            A, B = G0.copy(order="F"), G0.copy(order="F")
            O.apply_curl(A, F, op, isfwd, dl, isbloch, e, alpha=0.5 - 0.25j)
            CO.apply_curl(B, F, op, isfwd, dl, isbloch, e, alpha=0.5 - 0.25j)
            assert np.array_equal(A, B)
        dr = [rng.uniform(0.5, 1.5, n) for n in N]
        er = [1.0, -1.0, 1.0]
        for cmpfirst in (True, False):
            R = O.create_curl(isfwd, dr, isbloch, er, order_cmpfirst=cmpfirst)
            cp, rv, nz = CO.create_curl(isfwd, dr, isbloch, er, cmpfirst)
            assert np.array_equal(cp, R.colptr) and np.array_equal(rv, R.rowval) and np.array_equal(nz, R.nzval)


def test_gridgen_reproduces_the_pml_closed_form():
    """test/pml.jl:8-52: s-factor == f(l) for the default PMLParam, s∆l == s .* ∆l."""
    from sgc_b200 import gridgen as GG
    rng = np.random.default_rng(3)
    Nn, Np, N = 5, 7, 10
    Mtot = N + Nn + Np
    dldual = rng.random(Mtot) + 0.1
    L = dldual.sum()
    lprim = np.cumsum(np.concatenate([[-L / 2], dldual]))
    g = GG.Grid((lprim,), [True])
    assert g.N == (Mtot,)
    assert np.array_equal(g.dl[1][0], np.diff(lprim))
    ldual = (lprim[:-1] + lprim[1:]) / 2
    dlprim = np.concatenate([[ldual[0] + L - ldual[-1]], np.diff(ldual)])
    assert np.allclose(g.dl[0][0], dlprim, rtol=1e-13, atol=0)  # test/pml.jl:51 uses ≈ here
    omega = 0.83
    pml = GG.PMLParam()
    lp_n, lp_p = lprim[Nn], lprim[-1 - Np]
    Ln, Lp = lp_n - lprim[0], lprim[-1] - lp_p
    sm_n, sm_p = -(pml.m + 1) * np.log(pml.R) / 2 / Ln, -(pml.m + 1) * np.log(pml.R) / 2 / Lp

    def f(l):
        if l <= lp_n:
            return 1 + sm_n * ((lp_n - l) / Ln) ** pml.m / (1j * omega)
        if l >= lp_p:
            return 1 + sm_p * ((l - lp_p) / Lp) ** pml.m / (1j * omega)
        return 1.0 + 0j

    sprim = np.array([f(l) for l in lprim[:-1]])
    sdual = np.array([f(l) for l in ldual])
    s = GG.create_stretched_dl(omega, g, ([Nn], [Np]))
    assert np.allclose(s[0][0], sprim * g.dl[0][0], rtol=1e-13, atol=0)
    assert np.allclose(s[1][0], sdual * g.dl[1][0], rtol=1e-13, atol=0)
    inv = GG.invert_dl(s)
    assert np.array_equal(inv[0][0], 1.0 / s[0][0])
    # symmetric grid: ghost dual point mirrored about the boundary (grid.jl:141)
    g2 = GG.Grid((np.array([0.0, 1.0, 3.0, 4.0]),), [False])
    assert np.array_equal(g2.dl[0][0], np.array([1.0, 1.5, 1.5]))


@pytest.mark.parametrize("nz,nchunks", [(256, 16), (7, 3), (5, 8), (1, 4), (33, 1)])
@pytest.mark.parametrize("lo_iface,hi_iface", [(False, False), (True, False), (False, True), (True, True)])
def test_host_pipeline_plan_covers_every_plane_once(nz, nchunks, lo_iface, hi_iface):
    """The z-chunked host pipeline (hostpipe.pipeline_plan): every plane of E is copied in once, every plane of H and
    of E′ is computed exactly once, and no step uses data that has not arrived / been computed yet."""
    import importlib.util
    import os
    import sys
    import types
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    # hostpipe imports torch and the library binding; only the pure planning function is exercised here
    src = open(os.path.join(here, "staggeredgridcalculus.jl_b200", "hostpipe.py")).read()
    start = src.index("def pipeline_plan")
    end = src.index("\nclass HostCurlCurl")
    ns = {}
    exec(src[start:end], ns)
    steps, tail = ns["pipeline_plan"](nz, nchunks, lo_iface, hi_iface)
    arrived = 0
    H_done, E2_done = [0] * nz, [0] * nz
    for st in steps:
        a, b = st["h2d"]
        assert a == arrived and b > a
        arrived = b
        kb, ke = st["H"]
        for k in range(kb, max(kb, ke)):
            assert k + 1 < arrived or k == nz - 1 and arrived == nz and not hi_iface, "H[k] needs E[k+1]"
            H_done[k] += 1
        kb2, ke2 = st["E2"]
        for k in range(kb2, max(kb2, ke2)):
            assert H_done[k] == 1 and (k == 0 or H_done[k - 1] == 1), "E'[k] needs H[k-1], H[k]"
            E2_done[k] += 1
    assert arrived == nz
    if tail["H"]:
        for k in range(*tail["H"]):
            H_done[k] += 1
    for rng_ in tail["E2"]:
        for k in range(*rng_):
            E2_done[k] += 1
    assert H_done == [1] * nz and E2_done == [1] * nz


def test_product_never_touches_the_oracle_and_binding_matches_header():
    """Static checks of the boundary: (1) nothing under the product package (Python, C++/CUDA, Julia shim) imports,
    includes, links or names the oracle; (2) every prototype of include/sgc_b200.h has a ctypes declaration in
    `_lib.py` with the same number of parameters."""
    import os
    import re
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(here, "staggeredgridcalculus.jl_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")) or f == "Makefile":
                text = open(os.path.join(root, f), encoding="utf-8").read()
                assert not re.search(r"\boracle\b|sgc_oracle|libsgc_oracle", text), f"{f} refers to the oracle"
    header = open(os.path.join(here, "include", "sgc_b200.h"), encoding="utf-8").read()
    header = re.sub(r"/\*.*?\*/", " ", header, flags=re.S)
    protos = re.findall(r"\b(?:int|int64_t|const char \*)\s*(sgc_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", header, flags=re.S)
    assert len(protos) >= 50
    import sgc_b200
    table = sgc_b200._lib._PROTOS
    for name, params in protos:
        assert name in table, f"{name} is declared in the header but has no ctypes prototype"
        params = params.strip()
        n = 0 if params in ("", "void") else len(params.split(","))
        assert len(table[name][1]) == n, f"{name}: header has {n} parameters, _lib.py declares {len(table[name][1])}"


def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the reference's CPU algorithm, C/OpenMP port) runs without a GPU and prints exactly
    one JSON line carrying the keys of the bench contract."""
    import json
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(here, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--slab", "64,64,64"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-500:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
