"""Consumes tests/golden/julia/{manifest.json,data.bin} — inputs and outputs of the REAL reference, written by
tools/make_julia_golden.jl on a machine that has Julia — and compares (CPU) the oracle and (GPU) the CUDA library
against them: sparse structure bit-exact, values within 4 ulp / 1e-13 (the north-star tolerance).

No Julia exists in the build image, so the files are absent there: the tests then SKIP with the message
"parity unpinned", which is the honest state of the oracle.  The loader and the comparison code are
exercised on CPU by a self-made manifest in the same format (`test_manifest_roundtrip`)."""
import json
import os

import numpy as np
import pytest

from oracle import sgc_oracle as O
from tests.util import assert_parity

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "julia")
UNPINNED = ("parity unpinned: tests/golden/julia/manifest.json is absent (no Julia in this image); "
            "run `julia --project=<reference> tools/make_julia_golden.jl` to pin the oracle to the reference")
_DT = {"f64": np.float64, "c128": np.complex128, "i64": np.int64, "bool": np.uint8}


class Golden:
    def __init__(self, root):
        with open(os.path.join(root, "manifest.json")) as f:
            self.manifest = json.load(f)
        self.blob = np.memmap(os.path.join(root, "data.bin"), dtype=np.uint8, mode="r")

    def arr(self, d):
        dt = np.dtype(_DT[d["dtype"]])
        n = int(np.prod(d["shape"], dtype=np.int64)) if d["shape"] else 1
        a = np.frombuffer(self.blob, dtype=dt, count=n, offset=d["offset"]).reshape(d["shape"], order="F")
        return np.asfortranarray(a.astype(bool) if d["dtype"] == "bool" else a)

    @staticmethod
    def num(d):
        return complex(d["re"], d["im"]) if d["complex"] else float(d["re"])

    def csc(self, d):
        return d["m"], d["n"], self.arr(d["colptr"]), self.arr(d["rowval"]), self.arr(d["nzval"])

    def cases(self, fn):
        return [c for c in self.manifest["cases"] if c["fn"] == fn]


def load():
    if not os.path.exists(os.path.join(GOLDEN, "manifest.json")):
        pytest.skip(UNPINNED)
    return Golden(GOLDEN)


# ---- the implementations under comparison, behind one calling convention ---------------------------------------------
class OracleImpl:
    name = "oracle"

    def run_apply(self, fn, G0, F, args, kwargs):
        G = G0.copy(order="F")
        getattr(O, fn)(G, F, *args, **kwargs)
        return G

    def run_create(self, fn, args, kwargs):
        A = getattr(O, fn)(*args, **kwargs)
        return A.m, A.n, A.colptr, A.rowval, A.nzval


class GpuImpl:
    name = "libsgc_b200"

    def __init__(self):
        import sgc_b200
        self.S = sgc_b200

    def run_apply(self, fn, G0, F, args, kwargs):
        S = self.S
        Gd = S.DeviceArray.from_numpy(G0)
        getattr(S, fn)(Gd, S.DeviceArray.from_numpy(F), *args, **kwargs)
        return Gd.numpy()

    def run_create(self, fn, args, kwargs):
        A = getattr(self.S, fn)(*args, **kwargs)
        cp, rv, nz = A.to_host()
        return A.m, A.n, cp, rv, nz


def check_all(g: Golden, impl):
    n = 0
    for c in g.cases("apply_partial"):
        got = impl.run_apply("apply_partial", g.arr(c["G0"]), g.arr(c["F"]), (c["op"], c["nw"], c["isfwd"], g.arr(c["dwinv"]), c["isbloch"], g.num(c["e"])),
                             dict(alpha=g.num(c["alpha"])))
        assert_parity(got, g.arr(c["G"]), f"{impl.name} apply_∂! vs Julia {c['N']} nw={c['nw']} fwd={c['isfwd']} bloch={c['isbloch']} {c['op']}")
        n += 1
    for c in g.cases("apply_mhat") + g.cases("apply_mhat_weighted"):
        w = c["fn"].endswith("weighted")
        args = (c["op"], c["nw"], c["isfwd"], g.arr(c["dw"]) if w else None, g.arr(c["dwpinv"]) if w else None, c["isbloch"], g.num(c["e"]))
        got = impl.run_apply("apply_mhat", g.arr(c["G0"]), g.arr(c["F"]), args, dict(alpha=g.num(c["alpha"])))
        assert_parity(got, g.arr(c["G"]), f"{impl.name} apply_m̂! vs Julia {c['N']} nw={c['nw']} weighted={w} {c['op']}")
        n += 1
    for c in g.cases("apply_curl"):
        got = impl.run_apply("apply_curl", g.arr(c["G0"]), g.arr(c["F"]),
                             (c["op"], c["isfwd"], [g.arr(d) for d in c["dlinv"]], c["isbloch"], [g.num(e) for e in c["e"]]), dict(alpha=g.num(c["alpha"])))
        assert_parity(got, g.arr(c["G"]), f"{impl.name} apply_curl! vs Julia fwd={c['isfwd']} bloch={c['isbloch']} {c['op']}")
        n += 1
    for c in g.cases("apply_curl_sub"):
        F = g.arr(c["F"])
        G0 = np.zeros(tuple(c["N"]) + (len(c["cmp_out"]),), dtype=F.dtype, order="F")
        got = impl.run_apply("apply_curl", G0, F, ("=", c["isfwd"], [g.arr(d) for d in c["dlinv"]], c["isbloch"], [1.0, 1.0]),
                             dict(cmp_shp=[1, 2], cmp_out=c["cmp_out"], cmp_in=c["cmp_in"]))
        assert_parity(got, g.arr(c["G"]), f"{impl.name} sub-curl vs Julia out={c['cmp_out']} in={c['cmp_in']}")
        n += 1
    for fn in ("apply_divg", "apply_grad"):
        for c in g.cases(fn):
            got = impl.run_apply(fn, g.arr(c["G0"]), g.arr(c["F"]), (c["op"], c["isfwd"], [g.arr(d) for d in c["dlinv"]], c["isbloch"], [g.num(e) for e in c["e"]]), {})
            assert_parity(got, g.arr(c["G"]), f"{impl.name} {fn}! vs Julia fwd={c['isfwd']} bloch={c['isbloch']} {c['op']}")
            n += 1
    for c in g.cases("apply_mean_weighted"):
        got = impl.run_apply("apply_mean", g.arr(c["G0"]), g.arr(c["F"]),
                             (c["op"], c["isfwd"], [g.arr(d) for d in c["dl"]], [g.arr(d) for d in c["dlpinv"]], c["isbloch"], [g.num(e) for e in c["e"]]), {})
        assert_parity(got, g.arr(c["G"]), f"{impl.name} apply_mean! vs Julia fwd={c['isfwd']} bloch={c['isbloch']} {c['op']}")
        n += 1

    def same_csc(got, c, what):
        m, ncol, cp, rv, nz = g.csc(c["A"])
        assert (got[0], got[1]) == (m, ncol), what + ": shape"
        assert np.array_equal(got[2], cp), what + ": colptr differs from Julia's"
        assert np.array_equal(got[3], rv), what + ": rowval differs from Julia's"
        assert got[4].dtype == nz.dtype, what + f": eltype {got[4].dtype} vs Julia {nz.dtype}"
        assert_parity(got[4], nz, what + ": nzval")

    for c in g.cases("create_partial"):
        same_csc(impl.run_create("create_partial", (c["nw"], c["isfwd"], c["N"], g.arr(c["dwinv"]), c["isbloch"], g.num(c["e"])), {}), c, f"{impl.name} create_∂ {c['N']}")
        n += 1
    for c in g.cases("create_mhat"):
        same_csc(impl.run_create("create_mhat", (c["nw"], c["isfwd"], c["N"], g.arr(c["dw"]), g.arr(c["dwpinv"]), c["isbloch"], g.num(c["e"])), {}), c,
                 f"{impl.name} create_m̂ {c['N']}")
        n += 1
    for fn in ("create_curl", "create_divg", "create_grad"):
        for c in g.cases(fn):
            same_csc(impl.run_create(fn, (c["isfwd"], [g.arr(d) for d in c["dlinv"]], c["isbloch"], [g.num(e) for e in c["e"]]),
                                     dict(order_cmpfirst=c["order_cmpfirst"])), c, f"{impl.name} {fn} {c['N']} cmpfirst={c['order_cmpfirst']}")
            n += 1
    for c in g.cases("create_mean"):
        same_csc(impl.run_create("create_mean", (c["isfwd"], [g.arr(d) for d in c["dl"]], [g.arr(d) for d in c["dlinv"]], c["isbloch"], [g.num(e) for e in c["e"]]),
                                 dict(order_cmpfirst=c["order_cmpfirst"])), c, f"{impl.name} create_mean {c['N']}")
        n += 1
    for c in g.cases("create_picmp"):
        same_csc(impl.run_create("create_picmp", (c["N"], c["permute"]), dict(scale=g.arr(c["scale"]), order_cmpfirst=c["order_cmpfirst"])), c,
                 f"{impl.name} create_πcmp {c['N']}")
        n += 1
    return n


def test_oracle_matches_julia_golden():
    """CPU: the numpy restatement against the reference's own outputs (closes the oracle pin when the files exist)."""
    g = load()
    assert check_all(g, OracleImpl()) > 0
    for c in g.cases("pml"):  # grid / PML producers (src/grid.jl, src/pml.jl, src/util.jl) as restated in gridgen.py
        import sgc_b200.gridgen as GG  # noqa: F401  (numpy only)
        grid = GG.Grid([g.arr(d) for d in c["lprim"]], c["isbloch"])
        sdl = GG.create_stretched_dl(float(c["omega"]), grid, (c["npml_neg"], c["npml_pos"]))
        inv = GG.invert_dl(sdl)
        for k in range(3):
            assert_parity(sdl[0][k], g.arr(c["sdl_prim"][k]), "create_stretched_∆l primal")
            assert_parity(sdl[1][k], g.arr(c["sdl_dual"][k]), "create_stretched_∆l dual")
            assert_parity(inv[0][k], g.arr(c["inv_prim"][k]), "invert_∆l primal")
            assert_parity(inv[1][k], g.arr(c["inv_dual"][k]), "invert_∆l dual")


@pytest.mark.gpu
def test_library_matches_julia_golden():
    """GPU: libsgc_b200 through the C ABI against the reference's own outputs."""
    g = load()
    assert check_all(g, GpuImpl()) > 0


def test_manifest_roundtrip(tmp_path):
    """The golden-file format and the comparison driver, exercised without Julia: a manifest written here in the
    generator's format, with the ORACLE's outputs in place of Julia's, must load and compare clean (and a corrupted
    value must be caught)."""
    rng = np.random.default_rng(1)
    blob = bytearray()

    def put(a):
        a = np.asfortranarray(a)
        code = {np.dtype(np.float64): "f64", np.dtype(np.complex128): "c128", np.dtype(np.int64): "i64"}[a.dtype]
        d = {"dtype": code, "shape": list(a.shape), "offset": len(blob)}
        blob.extend(a.tobytes(order="F"))
        return d

    def num(x):
        return {"re": float(np.real(x)), "im": float(np.imag(x)), "complex": bool(np.iscomplexobj(x))}

    N = (5, 4, 3)
    F = rng.uniform(-1, 1, N + (3,)) + 1j * rng.uniform(-1, 1, N + (3,))
    G0 = rng.uniform(-1, 1, N + (3,)) + 1j * rng.uniform(-1, 1, N + (3,))
    dl = [rng.uniform(0.5, 1.5, n) + 1j * rng.uniform(-0.5, 0.5, n) for n in N]
    e = [np.exp(-0.3j), np.exp(0.5j), np.exp(-0.9j)]
    G = np.asfortranarray(G0.copy())
    O.apply_curl(G, np.asfortranarray(F), "+=", [True, False, True], dl, [True, False, True], e, alpha=0.5 - 0.25j)
    A = O.create_curl([True, False, True], dl, [True, False, True], e)
    cases = [{"fn": "apply_curl", "N": list(N), "op": "+=", "isfwd": [True, False, True], "isbloch": [True, False, True], "e": [num(x) for x in e],
              "alpha": num(0.5 - 0.25j), "dlinv": [put(d) for d in dl], "F": put(F), "G0": put(G0), "G": put(G)},
             {"fn": "create_curl", "N": list(N), "isfwd": [True, False, True], "isbloch": [True, False, True], "e": [num(x) for x in e], "order_cmpfirst": True,
              "dlinv": [put(d) for d in dl], "A": {"m": A.m, "n": A.n, "colptr": put(A.colptr), "rowval": put(A.rowval), "nzval": put(A.nzval)}}]
    (tmp_path / "manifest.json").write_text(json.dumps({"generator": "test", "cases": cases}))
    (tmp_path / "data.bin").write_bytes(bytes(blob))
    g = Golden(str(tmp_path))
    assert check_all(g, OracleImpl()) == 2
    # a wrong reference value is caught
    bad = bytearray(blob)
    off = cases[0]["G"]["offset"]
    bad[off:off + 8] = np.float64(123.0).tobytes()
    (tmp_path / "data.bin").write_bytes(bytes(bad))
    with pytest.raises(AssertionError):
        check_all(Golden(str(tmp_path)), OracleImpl())
