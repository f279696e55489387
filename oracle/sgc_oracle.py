"""CPU oracle for the two hot paths of StaggeredGridCalculus.jl — TEST INFRASTRUCTURE ONLY.

This file is a *restatement* (not a copy) of the reference's algorithms in numpy, written so
that every floating-point operation happens in the order Julia would perform it.  It is the
checker for the CUDA library: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  Nothing under
``staggeredgridcalculus.jl_b200/`` imports it, and the product path has no CPU fallback.

Pinning status
--------------
* The reference is pure Julia and no Julia runtime exists in this image, so the oracle cannot
  be run against the reference itself.  It is pinned instead against **every known-answer
  construction in the reference's own test-suite** (ported in ``tests/test_oracle_*.py``):
  the entry-by-entry naive matrices of ``test/partial.jl:22-61`` / ``test/mean.jl:35-80``
  (exact ``==``), the block identities and the component-major reorder map of
  ``test/curl.jl:19-51``, the invariants ``row sums == 0``, ``diag(Cv*Cu) == 4``, 13 nnz/row,
  Hermitian (``test/curl.jl:29,147-152``), ``Divg*Curl == 0`` (``test/divergence.jl:73``),
  ``Curl*Grad == 0`` (``test/gradient.jl:73``) and matrix-free ≈ matrix
  (``test/curl.jl:54-57``).  The reference holds no stored golden vectors.
* **Parity unpinned at the ulp level**: Julia Base's scalar semantics (complex ``*``, ``/``,
  ``inv``, no FMA contraction) are restated below from the published Base algorithms
  (``base/complex.jl``) and cannot be verified bit-for-bit here.  What can be anchored is: the
  complex division reproduces the ten published hard cases of Baudin & Smith (arXiv:1210.4539; the
  ``harddivs`` vectors of Julia's own ``test/complex.jl``) — nine exactly, the denormal case to one ulp —
  and the CUDA library's division routine equals this one bit for bit on those and on 10^5 random operand
  pairs over the whole exponent range (``tests/test_oracle_reference_suite.py``).

Conventions
-----------
Arrays are numpy arrays indexed exactly like the Julia arrays (``F[i,j,k,w]``, 0-based here);
memory order is irrelevant to the oracle.  ``op`` is ``"="`` or ``"+="`` (Julia
``Val(:(=))`` / ``Val(:(+=))``).  Axis arguments ``nw`` are 1-based as in the reference.
Complex arithmetic is carried out on separate real/imaginary float64 arrays (class ``JV``) so
that no library (numpy SIMD loops, libm) can contract or reorder the operations.
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "JV", "jl_inv", "jl_pow_pm1",
    "apply_partial", "apply_mhat", "apply_curl", "apply_divg", "apply_grad", "apply_mean",
    "create_partial_info", "create_minfo", "create_partial", "create_mhat", "create_mean",
    "create_curl", "create_divg", "create_grad", "create_picmp",
    "sparse_dropzeros", "CSC", "CURL_BLK",
]

_F = np.float64


# =============================================================================================
# Julia scalar semantics (Julia Base complex.jl)
# =============================================================================================
def _bcast(x, like):
    """Broadcast `x` against `like` without touching its bits (keeps the sign of zero)."""
    return np.broadcast_to(x, np.broadcast(x, like).shape)


class JV:
    """A real (im is None) or complex Float64 value/array with Julia's arithmetic rules.

    Real*Complex scales both parts; Complex*Complex is the 4-multiply/2-add form; division
    by a complex uses Base's scaled robust algorithm; nothing is ever fused.
    """

    __slots__ = ("re", "im")

    def __init__(self, re, im=None):
        self.re = np.asarray(re, dtype=_F)
        self.im = None if im is None else np.asarray(im, dtype=_F)

    # ---- construction / extraction -----------------------------------------------------
    @staticmethod
    def of(x) -> "JV":
        if isinstance(x, JV):
            return x
        a = np.asarray(x)
        if np.iscomplexobj(a):
            return JV(a.real.astype(_F), a.imag.astype(_F))
        return JV(a.astype(_F))

    @property
    def iscomplex(self) -> bool:
        return self.im is not None

    def to_numpy(self):
        if self.im is None:
            return np.array(self.re, dtype=_F)
        out = np.empty(np.broadcast(self.re, self.im).shape, dtype=np.complex128)
        out.real = self.re
        out.imag = self.im
        return out

    def __getitem__(self, idx) -> "JV":
        return JV(self.re[idx], None if self.im is None else self.im[idx])

    def reshape(self, shape) -> "JV":
        return JV(self.re.reshape(shape), None if self.im is None else self.im.reshape(shape))

    @property
    def shape(self):
        return self.re.shape

    # ---- arithmetic --------------------------------------------------------------------
    def __neg__(self):
        return JV(-self.re, None if self.im is None else -self.im)

    def __add__(self, o):
        o = JV.of(o)
        if self.im is None and o.im is None:
            return JV(self.re + o.re)
        if self.im is None:  # +(x::Real, z::Complex) = Complex(x + real(z), imag(z))
            return JV(self.re + o.re, _bcast(o.im, self.re))
        if o.im is None:
            return JV(self.re + o.re, _bcast(self.im, o.re))
        return JV(self.re + o.re, self.im + o.im)

    def __sub__(self, o):
        o = JV.of(o)
        if self.im is None and o.im is None:
            return JV(self.re - o.re)
        if self.im is None:  # -(x::Real, z::Complex) = Complex(x - real(z), -imag(z))
            return JV(self.re - o.re, _bcast(-o.im, self.re))
        if o.im is None:  # -(z::Complex, x::Real) = Complex(real(z) - x, imag(z))
            return JV(self.re - o.re, _bcast(self.im, o.re))
        return JV(self.re - o.re, self.im - o.im)

    def __mul__(self, o):
        o = JV.of(o)
        if self.im is None and o.im is None:
            return JV(self.re * o.re)
        if self.im is None:  # *(x::Real, z::Complex) = Complex(x * real(z), x * imag(z))
            return JV(self.re * o.re, self.re * o.im)
        if o.im is None:
            return JV(o.re * self.re, o.re * self.im)
        # *(z::Complex, w::Complex) = Complex(zr*wr - zi*wi, zr*wi + zi*wr)
        return JV(self.re * o.re - self.im * o.im, self.re * o.im + self.im * o.re)

    __rmul__ = lambda self, o: JV.of(o).__mul__(self)
    __radd__ = lambda self, o: JV.of(o).__add__(self)
    __rsub__ = lambda self, o: JV.of(o).__sub__(self)

    def __truediv__(self, o):
        o = JV.of(o)
        if o.im is None:
            if self.im is None:
                return JV(self.re / o.re)
            return JV(self.re / o.re, self.im / o.re)  # /(z::Complex, x::Real)
        if self.im is None:  # /(a::Real, z::Complex) = a * inv(z)
            return self * jl_inv(o)
        return _robust_cdiv(self, o)


def _robust_cdiv2(a, b, c, d, r, t):
    # Base.robust_cdiv2 (complex.jl); element-wise with the same branch structure.
    with np.errstate(all="ignore"):
        br = b * r
        v_rnz = np.where(br != 0, (a + br) * t, a * t + (b * t) * r)
        v_rz = (a + d * (b / c)) * t
    return np.where(r != 0, v_rnz, v_rz)


def _robust_cdiv1(a, b, c, d):
    with np.errstate(all="ignore"):
        r = d / c
        t = 1.0 / (c + d * r)
    p = _robust_cdiv2(a, b, c, d, r, t)
    q = _robust_cdiv2(b, -a, c, d, r, t)
    return p, q


def _robust_cdiv(z: JV, w: JV) -> JV:
    """`/(z::ComplexF64, w::ComplexF64)` of Julia Base (Baudin & Smith robust division)."""
    a, b = np.broadcast_arrays(z.re, z.im)
    c, d = np.broadcast_arrays(w.re, w.im)
    shp = np.broadcast(a, c).shape
    a = np.broadcast_to(a, shp).astype(_F).copy()
    b = np.broadcast_to(b, shp).astype(_F).copy()
    c = np.broadcast_to(c, shp).astype(_F).copy()
    d = np.broadcast_to(d, shp).astype(_F).copy()
    half, two = 0.5, 2.0
    ab = np.maximum(np.abs(a), np.abs(b))
    cd = np.maximum(np.abs(c), np.abs(d))
    ov = np.finfo(_F).max
    un = np.finfo(_F).tiny
    eps = np.finfo(_F).eps
    bs = two / (eps * eps)
    s = np.ones(shp, dtype=_F)
    m = ab >= half * ov
    a[m] *= half; b[m] *= half; s[m] *= two
    m = cd >= half * ov
    c[m] *= half; d[m] *= half; s[m] *= half
    m = ab <= un * two / eps
    a[m] *= bs; b[m] *= bs; s[m] /= bs
    m = cd <= un * two / eps
    c[m] *= bs; d[m] *= bs; s[m] *= bs
    p1, q1 = _robust_cdiv1(a, b, c, d)
    p2, q2 = _robust_cdiv1(b, a, d, c)
    sel = np.abs(d) <= np.abs(c)
    p = np.where(sel, p1, p2)
    q = np.where(sel, q1, -q2)
    return JV(p * s, q * s)


def jl_inv(w) -> JV:
    """`inv(w)`; for ComplexF64 the scaled `robust_cinv` algorithm of Julia Base."""
    w = JV.of(w)
    if w.im is None:
        return JV(1.0 / w.re)
    c = np.array(w.re, dtype=_F, copy=True)
    d = np.array(w.im, dtype=_F, copy=True)
    c, d = np.broadcast_arrays(c, d)
    c = c.copy(); d = d.copy()
    absc, absd = np.abs(c), np.abs(d)
    cd = np.where(absc > absd, absc, absd)
    eps = np.finfo(_F).eps
    bs = 2.0 / (eps * eps)
    s = np.ones_like(c)
    big = cd >= np.finfo(_F).max / 2
    small = (~big) & (cd <= 2 * np.finfo(_F).tiny / eps)
    c[big] *= 0.5; d[big] *= 0.5; s[big] = 0.5
    c[small] *= bs; d[small] *= bs; s[small] = bs

    # Base's robust_cinv writes `muladd(d, r, c)`, which LLVM contracts to one FMA where the host has the
    # instruction and leaves as a multiply and an add where it has not.  Which of the two the reference's own
    # machine produces cannot be determined here; this restatement (and the CUDA library's host routine, which the
    # tests hold bit-identical to it) evaluates it unfused.  The two differ by at most one rounding of `p`, i.e.
    # within the 4-ulp budget of the north star, and only in `inv(e⁻ⁱᵏᴸ)` of backward Bloch blocks.
    def cinv(cc, dd):  # robust_cinv: r = d/c; p = inv(muladd(d, r, c)); q = -r*p
        with np.errstate(all="ignore"):
            r = dd / cc
            p = 1.0 / (dd * r + cc)
            q = -r * p
        return p, q

    p1, q1 = cinv(c, d)
    q2, p2 = cinv(-d, -c)
    sel = absd <= absc
    p = np.where(sel, p1, p2)
    q = np.where(sel, q1, q2)
    return JV(p * s, q * s)


def jl_pow_pm1(e, ns: float) -> JV:
    """`e⁻ⁱᵏᴸ^ns` for `ns = ±1.0` (matrix/differential/partial.jl:217): integer-valued
    exponent ⇒ power-by-squaring ⇒ `e` itself or `inv(e)`."""
    e = JV.of(e)
    return e if ns > 0 else jl_inv(e)


# =============================================================================================
# Matrix-free operators  (src/matrixfree)
# =============================================================================================
def _store(G: np.ndarray, idx, val: JV, op: str):
    """`set_or_add!` (matrixfree/matrixfree.jl:6-8): `G[idx] = val` or `G[idx] += val`."""
    if op == "+=":
        val = JV.of(G[idx]) + val
    elif op != "=":
        raise ValueError(f"op must be '=' or '+=', got {op!r}")
    if np.iscomplexobj(G):
        if val.im is None:
            G[idx] = np.broadcast_to(val.re, G[idx].shape)
        else:
            tmp = np.empty(G[idx].shape, dtype=np.complex128)
            tmp.real = val.re
            tmp.imag = val.im
            G[idx] = tmp
    else:
        if val.im is not None and np.any(val.im != 0):
            raise TypeError("InexactError: cannot store a complex value into a Float64 array")
        G[idx] = np.broadcast_to(val.re, G[idx].shape)


def _col(v: JV, n: int, ndim: int) -> JV:
    """Align a length-n vector with the leading axis of an ndim-array (for broadcasting)."""
    return v.reshape((n,) + (1,) * (ndim - 1))


def apply_partial(Gv, Fu, op, nw, isfwd, dwinv=1.0, isbloch=True, e=1.0, alpha=1.0):
    """`apply_∂!` — matrixfree/differential/partial.jl:33-237 (3-D), :241-385 (2-D),
    :388-451 (1-D); scalar-∆ wrapper :11-21.  Formulas per SURVEY.md Appendix A.1."""
    assert Gv.shape == Fu.shape
    K = Fu.ndim
    assert 1 <= nw <= K
    a = nw - 1
    G = np.moveaxis(Gv, a, 0)
    F = JV.of(np.moveaxis(Fu, a, 0))
    Nw = F.shape[0]
    if np.ndim(dwinv) == 0:
        dwinv = np.full(Nw, dwinv)  # fill(∆w⁻¹, N[nw])  (partial.jl:21)
    assert len(dwinv) == Nw
    al = JV.of(alpha)
    e = JV.of(e)
    d = JV.of(dwinv)
    c = _col(al * d, Nw, K)  # (α * ∆w⁻¹[i]) for every i
    sl = lambda lo, hi: slice(lo, hi)
    if isfwd:
        if isbloch:
            if Nw > 1:  # i = 1..Nw-1                                     partial.jl:54-58
                _store(G, sl(0, Nw - 1), c[0:Nw - 1] * (F[1:Nw] - F[0:Nw - 1]), op)
            beta = al * d[Nw - 1]  # β = α * ∆w⁻¹[Nw]                         partial.jl:62-65
            _store(G, sl(Nw - 1, Nw), beta * (e * F[0:1] - F[Nw - 1:Nw]), op)
        else:
            if Nw < 2:
                raise IndexError("forward symmetric ∂ reads Fu[2]: needs N[nw] ≥ 2 (partial.jl:79)")
            if Nw > 2:  # i = 2..Nw-1                                     partial.jl:69-73
                _store(G, sl(1, Nw - 1), c[1:Nw - 1] * (F[2:Nw] - F[1:Nw - 1]), op)
            beta = al * d[0]  # partial.jl:77-80
            _store(G, sl(0, 1), beta * F[1:2], op)
            beta = (-al) * d[Nw - 1]  # β = -α * ∆w⁻¹[Nw]                     partial.jl:84-87
            _store(G, sl(Nw - 1, Nw), beta * F[Nw - 1:Nw], op)
    else:
        if Nw > 1:  # i = 2..Nw                                           partial.jl:171-175
            _store(G, sl(1, Nw), c[1:Nw] * (F[1:Nw] - F[0:Nw - 1]), op)
        if isbloch:  # partial.jl:179-183
            beta = al * d[0]
            _store(G, sl(0, 1), beta * (F[0:1] - F[Nw - 1:Nw] / e), op)
        else:  # partial.jl:184-188
            _store(G, sl(0, 1), JV(np.zeros(())), op)
    return None


def apply_mhat(Gv, Fu, op, nw, isfwd, dw=None, dwpinv=None, isbloch=True, e=1.0, alpha=1.0):
    """`apply_m̂!` — matrixfree/mean.jl:102-489 (arithmetic; `dw is None`) and :493-922
    (weighted).  Formulas per SURVEY.md Appendix A.2."""
    assert Gv.shape == Fu.shape
    K = Fu.ndim
    a = nw - 1
    G = np.moveaxis(Gv, a, 0)
    F = JV.of(np.moveaxis(Fu, a, 0))
    Nw = F.shape[0]
    al = JV.of(alpha)
    a2 = JV(0.5) * al  # α2 = 0.5α                                           mean.jl:114, 509
    e = JV.of(e)
    sl = lambda lo, hi: slice(lo, hi)
    if dw is None:
        assert dwpinv is None
        if isfwd:
            if isbloch:
                if Nw > 1:  # mean.jl:121-125
                    _store(G, sl(0, Nw - 1), a2 * (F[1:Nw] + F[0:Nw - 1]), op)
                _store(G, sl(Nw - 1, Nw), a2 * (e * F[0:1] + F[Nw - 1:Nw]), op)  # :129-131
            else:
                if Nw < 2:
                    raise IndexError("forward symmetric m̂ reads Fu[2]: needs N[nw] ≥ 2")
                if Nw > 2:  # mean.jl:135-139
                    _store(G, sl(1, Nw - 1), a2 * (F[2:Nw] + F[1:Nw - 1]), op)
                _store(G, sl(0, 1), a2 * F[1:2], op)  # :143-145
                _store(G, sl(Nw - 1, Nw), a2 * F[Nw - 1:Nw], op)  # :149-151
        else:
            if Nw > 1:  # mean.jl:229-233
                _store(G, sl(1, Nw), a2 * (F[1:Nw] + F[0:Nw - 1]), op)
            if isbloch:  # :237-240
                _store(G, sl(0, 1), a2 * (F[0:1] + F[Nw - 1:Nw] / e), op)
            else:  # :241-245  (α, not α2)
                _store(G, sl(0, 1), al * F[0:1], op)
        return None

    assert len(dw) == Nw and len(dwpinv) == Nw
    w = JV.of(dw)
    wp = JV.of(dwpinv)
    W = _col(w, Nw, K)
    c = _col(a2 * wp, Nw, K)  # (α2 * ∆w′⁻¹[i])
    if isfwd:
        if isbloch:
            if Nw > 1:  # mean.jl:516-520
                _store(G, sl(0, Nw - 1),
                       c[0:Nw - 1] * (W[1:Nw] * F[1:Nw] + W[0:Nw - 1] * F[0:Nw - 1]), op)
            beta = a2 * wp[Nw - 1]  # :524-527   β * (∆w[1]*e*Fu[1] + ∆w[Nw]*Fu[Nw])
            _store(G, sl(Nw - 1, Nw), beta * ((w[0] * e) * F[0:1] + w[Nw - 1] * F[Nw - 1:Nw]), op)
        else:
            if Nw < 2:
                raise IndexError("forward symmetric m̂ reads Fu[2]: needs N[nw] ≥ 2")
            if Nw > 2:  # :531-535
                _store(G, sl(1, Nw - 1),
                       c[1:Nw - 1] * (W[2:Nw] * F[2:Nw] + W[1:Nw - 1] * F[1:Nw - 1]), op)
            beta = a2 * wp[0]  # :539-542   β * ∆w[2]*Fu[2]  ==  (β*∆w[2])*Fu[2]
            _store(G, sl(0, 1), (beta * w[1]) * F[1:2], op)
            beta = a2 * wp[Nw - 1]  # :546-549
            _store(G, sl(Nw - 1, Nw), (beta * w[Nw - 1]) * F[Nw - 1:Nw], op)
    else:
        if Nw > 1:  # :633-637
            _store(G, sl(1, Nw), c[1:Nw] * (W[1:Nw] * F[1:Nw] + W[0:Nw - 1] * F[0:Nw - 1]), op)
        if isbloch:  # :641-645   β * (∆w[1]*Fu[1] + ∆w[Nw]*Fu[Nw]/e)
            beta = a2 * wp[0]
            _store(G, sl(0, 1), beta * (w[0] * F[0:1] + (w[Nw - 1] * F[Nw - 1:Nw]) / e), op)
        else:  # :646-650   β = α * ∆w′⁻¹[1];  β * ∆w[1]*Fu[1]
            beta = al * wp[0]
            _store(G, sl(0, 1), (beta * w[0]) * F[0:1], op)
    return None


# CURL_BLK[nv][nu] (row nv, column nu), matrix/differential/curl.jl:46-48 (the literal there
# lists columns):   [ 0 -1 +1 ;  +1 0 -1 ;  -1 +1 0 ]
CURL_BLK = ((0, -1, 1), (1, 0, -1), (-1, 1, 0))


def _vecs(dlinv, N):
    """Scalar-∆ wrappers: `fill.(∆l⁻¹, N)` (matrixfree/differential/curl.jl:24)."""
    return [np.full(n, v) if np.ndim(v) == 0 else np.asarray(v) for v, n in zip(dlinv, N)]


def apply_curl(G, F, op, isfwd, dlinv=None, isbloch=None, e=None, *, cmp_shp=None, cmp_out=None,
               cmp_in=None, alpha=1.0):
    """`apply_curl!` — matrixfree/differential/curl.jl:52-137 (+ wrappers :12-49)."""
    K = G.ndim - 1
    assert F.ndim == K + 1
    N = G.shape[:K]
    dlinv = _vecs([1.0] * K if dlinv is None else dlinv, N)
    isbloch = [True] * K if isbloch is None else isbloch
    e = [1.0] * K if e is None else e
    cmp_shp = list(range(1, K + 1)) if cmp_shp is None else list(cmp_shp)
    cmp_out = list(range(1, G.shape[K] + 1)) if cmp_out is None else list(cmp_out)
    cmp_in = list(range(1, F.shape[K] + 1)) if cmp_in is None else list(cmp_in)
    for ind_nv, nv in enumerate(cmp_out):
        Gv = G[..., ind_nv]
        cur = op
        for ind_nu, nu in enumerate(cmp_in):
            parity = CURL_BLK[nv - 1][nu - 1]
            if parity == 0:  # curl.jl:75-81
                if cur == "=":
                    Gv[...] = 0
                    cur = "+="
                continue
            nw = 6 - nv - nu
            if cmp_shp.count(nw) != 1:  # curl.jl:85 (@error)
                raise ValueError(f"cmp_shp = {cmp_shp} does not have one and only one nw = {nw}")
            ind_nw = cmp_shp.index(nw) + 1
            Fu = F[..., ind_nu]
            apply_partial(Gv, Fu, cur, ind_nw, isfwd[ind_nw - 1], dlinv[ind_nw - 1],
                          isbloch[ind_nw - 1], e[ind_nw - 1],
                          alpha=JV(float(parity)) * JV.of(alpha))  # α = parity*α   curl.jl:89
            cur = "+="
    return None


def apply_divg(g, F, op, isfwd, dlinv=None, isbloch=None, e=None, *, alpha=1.0):
    """`apply_divg!` — matrixfree/differential/divergence.jl:41-66."""
    K = g.ndim
    assert F.ndim == K + 1
    dlinv = _vecs([1.0] * K if dlinv is None else dlinv, g.shape)
    isbloch = [True] * K if isbloch is None else isbloch
    e = [1.0] * K if e is None else e
    apply_partial(g, F[..., 0], op, 1, isfwd[0], dlinv[0], isbloch[0], e[0], alpha=alpha)
    for nw in range(2, K + 1):
        apply_partial(g, F[..., nw - 1], "+=", nw, isfwd[nw - 1], dlinv[nw - 1], isbloch[nw - 1],
                      e[nw - 1], alpha=alpha)
    return None


def apply_grad(G, f, op, isfwd, dlinv=None, isbloch=None, e=None, *, alpha=1.0):
    """`apply_grad!` — matrixfree/differential/gradient.jl:41-59."""
    K = f.ndim
    assert G.ndim == K + 1
    dlinv = _vecs([1.0] * K if dlinv is None else dlinv, f.shape)
    isbloch = [True] * K if isbloch is None else isbloch
    e = [1.0] * K if e is None else e
    for nw in range(1, K + 1):
        apply_partial(G[..., nw - 1], f, op, nw, isfwd[nw - 1], dlinv[nw - 1], isbloch[nw - 1],
                      e[nw - 1], alpha=alpha)
    return None


def apply_mean(G, F, op, isfwd, dl=None, dlpinv=None, isbloch=None, e=None, *, alpha=1.0):
    """`apply_mean!` — matrixfree/mean.jl:47-65 (arithmetic), :69-89 (weighted)."""
    K = G.ndim - 1
    isbloch = [True] * K if isbloch is None else isbloch
    e = [1.0] * K if e is None else e
    for nw in range(1, K + 1):
        apply_mhat(G[..., nw - 1], F[..., nw - 1], op, nw, isfwd[nw - 1],
                   None if dl is None else dl[nw - 1], None if dlpinv is None else dlpinv[nw - 1],
                   isbloch[nw - 1], e[nw - 1], alpha=alpha)
    return None


# =============================================================================================
# Sparse assembly  (src/matrix) — COO exactly as the reference builds it, then the stdlib
# contract `dropzeros!(sparse(I, J, V, m, n))`.
# =============================================================================================
class CSC:
    """`SparseMatrixCSC{Tv,Int64}`: 1-based `colptr` (n+1), `rowval`, `nzval`."""

    def __init__(self, m, n, colptr, rowval, nzval):
        self.m, self.n = int(m), int(n)
        self.colptr = np.asarray(colptr, dtype=np.int64)
        self.rowval = np.asarray(rowval, dtype=np.int64)
        self.nzval = np.asarray(nzval)

    @property
    def nnz(self):
        return int(self.colptr[-1] - 1)

    def to_scipy(self):
        import scipy.sparse as sp
        return sp.csc_matrix((self.nzval, self.rowval - 1, self.colptr - 1), shape=(self.m, self.n))

    def todense(self):
        return self.to_scipy().toarray()


def sparse_dropzeros(I, J, V: JV, m, n) -> CSC:
    """`dropzeros!(sparse(I, J, V, m, n))` (Julia stdlib SparseArrays; call sites
    matrix/differential/partial.jl:60, curl.jl:103, divergence.jl:69, gradient.jl:69,
    matrix/mean.jl:110,138, matrix/permutation.jl:40): duplicates summed in input order,
    rows ascending within each column, then every stored value `== 0` removed."""
    I = np.asarray(I, dtype=np.int64)
    J = np.asarray(J, dtype=np.int64)
    order = np.lexsort((I, J))  # stable: primary J (column), secondary I (row)
    I, J = I[order], J[order]
    vre = np.broadcast_to(V.re, I.shape)[order]
    vim = None if V.im is None else np.broadcast_to(V.im, I.shape)[order]
    if len(I):
        first = np.ones(len(I), dtype=bool)
        first[1:] = (I[1:] != I[:-1]) | (J[1:] != J[:-1])
        starts = np.flatnonzero(first)
        I, J = I[starts], J[starts]
        vre = np.add.reduceat(vre, starts)
        if vim is not None:
            vim = np.add.reduceat(vim, starts)
    keep = (vre != 0) if vim is None else ((vre != 0) | (vim != 0))
    I, J, vre = I[keep], J[keep], vre[keep]
    if vim is not None:
        vim = vim[keep]
    counts = np.bincount(J - 1, minlength=n).astype(np.int64)
    colptr = np.empty(n + 1, dtype=np.int64)
    colptr[0] = 1
    np.cumsum(counts, out=colptr[1:])
    colptr[1:] += 1
    nz = JV(vre, vim).to_numpy()
    return CSC(m, n, colptr, I, nz)


def _promote(*xs):
    return any(JV.of(x).iscomplex for x in xs)


def _as_T(v: JV, cplx: bool) -> JV:
    if cplx and v.im is None:
        return JV(v.re, np.zeros_like(v.re))
    return v


def _shifted_index_arrays(N, nw, ns):
    """I = reshape(collect(1:M), N);  Jₛ = circshift(I, -ns*ŵ)   (partial.jl:118-122)."""
    M = int(np.prod(N))
    I = np.arange(1, M + 1, dtype=np.int64).reshape(N, order="F")
    Js = np.roll(I, -int(ns), axis=nw - 1)
    return I, Js


def create_partial_info(nw, isfwd, N, dwinv, isbloch, e):
    """`create_∂info` — matrix/differential/partial.jl:87-314.  Returns (Is, Js, Vs: JV)."""
    N = tuple(int(n) for n in N)
    K = len(N)
    M = int(np.prod(N))
    Nw = N[nw - 1]
    ns = 1.0 if isfwd else -1.0
    I, Js = _shifted_index_arrays(N, nw, ns)
    cplx = _promote(dwinv, e)  # T = promote_type(eltype(∆w⁻¹), eltype(e⁻ⁱᵏᴸ))     :136
    sizew = tuple(Nw if d == nw - 1 else 1 for d in range(K))
    dW = JV.of(np.asarray(dwinv)).reshape(sizew)
    ones = _as_T(JV(np.ones(N)), cplx)
    V0 = (JV(-ns) * ones) * dW  # V₀ = -ns .* ones(T, N) .* ∆W⁻¹                    :137
    Vs = (JV(ns) * ones) * dW  # Vₛ =  ns .* ones(T, N) .* ∆W⁻¹                     :138
    V0 = _as_T(JV(np.array(np.broadcast_to(V0.re, N)),
                  None if V0.im is None else np.array(np.broadcast_to(V0.im, N))), cplx)
    Vs = _as_T(JV(np.array(np.broadcast_to(Vs.re, N)),
                  None if Vs.im is None else np.array(np.broadcast_to(Vs.im, N))), cplx)

    def slab(iw):
        return tuple(slice(iw, iw + 1) if d == nw - 1 else slice(None) for d in range(K))

    if isbloch:  # Vₛ[slab] .*= e⁻ⁱᵏᴸ^ns                                           :216-217
        sb = slab(Nw - 1 if isfwd else 0)
        new = Vs[sb] * jl_pow_pm1(e, ns)
        Vs.re[sb] = new.re
        if cplx:
            Vs.im[sb] = new.im if new.im is not None else 0.0
    else:  # V₀[r_w == 1] .= 0 ; Vₛ[slab] .= 0                                       :290-299
        sb = slab(0)
        V0.re[sb] = 0.0
        if cplx:
            V0.im[sb] = 0.0
        sb = slab(Nw - 1 if isfwd else 0)
        Vs.re[sb] = 0.0
        if cplx:
            Vs.im[sb] = 0.0
    fl = lambda A: A.reshape(M, order="F")
    i, js = fl(I), fl(Js)
    Is = np.concatenate([i, i])  # :309-311
    Jc = np.concatenate([i, js])
    Vre = np.concatenate([fl(V0.re), fl(Vs.re)])
    Vim = np.concatenate([fl(V0.im), fl(Vs.im)]) if cplx else None
    return Is, Jc, JV(Vre, Vim)


def create_minfo(nw, isfwd, N, dw, dwpinv, isbloch, e):
    """`create_minfo` — matrix/mean.jl:141-377."""
    N = tuple(int(n) for n in N)
    K = len(N)
    M = int(np.prod(N))
    Nw = N[nw - 1]
    ns = 1.0 if isfwd else -1.0
    I, Js = _shifted_index_arrays(N, nw, ns)
    cplx = _promote(dw, dwpinv, e)  # mean.jl:192
    sizew = tuple(Nw if d == nw - 1 else 1 for d in range(K))
    W = JV.of(np.asarray(dw)).reshape(sizew)
    Wp = JV.of(np.asarray(dwpinv)).reshape(sizew)
    half = _as_T(JV(np.full(N, 0.5)), cplx)
    V0 = (half * W) * Wp  # fill(T(0.5), N) .* ∆W .* ∆W′⁻¹                            :193
    Vs = half * W  # :195
    Vs = JV(np.roll(Vs.re, -int(ns), axis=nw - 1),
            None if Vs.im is None else np.roll(Vs.im, -int(ns), axis=nw - 1))  # circshift :196
    Vs = Vs * Wp  # :197
    V0 = _as_T(JV(np.array(V0.re), None if V0.im is None else np.array(V0.im)), cplx)
    Vs = _as_T(JV(np.array(Vs.re), None if Vs.im is None else np.array(Vs.im)), cplx)

    def slab(iw):
        return tuple(slice(iw, iw + 1) if d == nw - 1 else slice(None) for d in range(K))

    if isbloch:  # :273-274
        sb = slab(Nw - 1 if isfwd else 0)
        new = Vs[sb] * jl_pow_pm1(e, ns)
        Vs.re[sb] = new.re
        if cplx:
            Vs.im[sb] = new.im if new.im is not None else 0.0
    else:  # :350-356
        sb = slab(0)
        val = 0.0 if isfwd else 2.0
        new = V0[sb] * JV(val)
        V0.re[sb] = new.re
        if cplx:
            V0.im[sb] = new.im
        sb = slab(Nw - 1 if isfwd else 0)
        Vs.re[sb] = 0.0
        if cplx:
            Vs.im[sb] = 0.0
    fl = lambda A: A.reshape(M, order="F")
    i, js = fl(I), fl(Js)
    Is = np.concatenate([i, i])
    Jc = np.concatenate([i, js])
    Vre = np.concatenate([fl(V0.re), fl(Vs.re)])
    Vim = np.concatenate([fl(V0.im), fl(Vs.im)]) if cplx else None
    return Is, Jc, JV(Vre, Vim)


def create_partial(nw, isfwd, N, dwinv=1.0, isbloch=True, e=1.0) -> CSC:
    """`create_∂` — matrix/differential/partial.jl:45-60."""
    N = tuple(int(n) for n in N)
    if np.ndim(dwinv) == 0:
        dwinv = np.full(N[nw - 1], dwinv)
    M = int(np.prod(N))
    return sparse_dropzeros(*create_partial_info(nw, isfwd, N, dwinv, isbloch, e), M, M)


def create_mhat(nw, isfwd, N, dw=None, dwpinv=None, isbloch=True, e=1.0) -> CSC:
    """`create_m̂` — matrix/mean.jl:124-138."""
    N = tuple(int(n) for n in N)
    if dw is None:
        dw = np.ones(N[nw - 1]); dwpinv = np.ones(N[nw - 1])
    M = int(np.prod(N))
    return sparse_dropzeros(*create_minfo(nw, isfwd, N, dw, dwpinv, isbloch, e), M, M)


def _cat(parts_I, parts_J, parts_V, cplx):
    I = np.concatenate(parts_I) if parts_I else np.zeros(0, np.int64)
    J = np.concatenate(parts_J) if parts_J else np.zeros(0, np.int64)
    if not parts_V:
        return I, J, JV(np.zeros(0), np.zeros(0) if cplx else None)
    re = np.concatenate([v.re for v in parts_V])
    im = np.concatenate([_as_T(v, True).im for v in parts_V]) if cplx else None
    return I, J, JV(re, im)


def create_curl(isfwd, dlinv, isbloch=None, e=None, *, N=None, cmp_shp=None, cmp_out=None,
                cmp_in=None, order_cmpfirst=True) -> CSC:
    """`create_curl` — matrix/differential/curl.jl:50-104 (+ wrappers :4-36).  With scalar
    `dlinv` entries pass `N` (the wrapper `fill.(∆l⁻¹, (N...,))`, curl.jl:22)."""
    K = len(isfwd)
    if N is not None:
        dlinv = _vecs([1.0] * K if dlinv is None else dlinv, N)
    dlinv = [np.asarray(v) for v in dlinv]
    isbloch = [True] * K if isbloch is None else isbloch
    e = [1.0] * K if e is None else e
    cmp_shp = list(range(1, K + 1)) if cmp_shp is None else list(cmp_shp)
    cmp_out = list(range(1, K + 1)) if cmp_out is None else list(cmp_out)
    cmp_in = list(range(1, K + 1)) if cmp_in is None else list(cmp_in)
    Kout, Kin = len(cmp_out), len(cmp_in)
    Ng = tuple(len(v) for v in dlinv)
    M = int(np.prod(Ng))
    cplx = _promote(*dlinv) or _promote(np.asarray(e))  # curl.jl:59
    pI, pJ, pV = [], [], []
    for ind_nv, nv in enumerate(cmp_out, start=1):
        istr, ioff = (Kout, ind_nv - Kout) if order_cmpfirst else (1, M * (ind_nv - 1))
        for ind_nu, nu in enumerate(cmp_in, start=1):
            parity = CURL_BLK[nv - 1][nu - 1]
            if parity == 0:
                continue
            jstr, joff = (Kin, ind_nu - Kin) if order_cmpfirst else (1, M * (ind_nu - 1))
            nw = 6 - nv - nu
            if cmp_shp.count(nw) != 1:
                raise ValueError(f"cmp_shp = {cmp_shp} does not have one and only one nw = {nw}")
            ind_nw = cmp_shp.index(nw) + 1
            I, J, V = create_partial_info(ind_nw, isfwd[ind_nw - 1], Ng, dlinv[ind_nw - 1],
                                          isbloch[ind_nw - 1], e[ind_nw - 1])
            pI.append(istr * I + ioff)  # curl.jl:89-91
            pJ.append(jstr * J + joff)
            pV.append(V * JV(float(parity)))
    return sparse_dropzeros(*_cat(pI, pJ, pV, cplx), Kout * M, Kin * M)


def create_divg(isfwd, dlinv, isbloch=None, e=None, *, N=None, order_cmpfirst=True) -> CSC:
    """`create_divg` — matrix/differential/divergence.jl:37-70."""
    K = len(isfwd)
    if N is not None:
        dlinv = _vecs([1.0] * K if dlinv is None else dlinv, N)
    dlinv = [np.asarray(v) for v in dlinv]
    isbloch = [True] * K if isbloch is None else isbloch
    e = [1.0] * K if e is None else e
    Ng = tuple(len(v) for v in dlinv)
    M = int(np.prod(Ng))
    cplx = _promote(*dlinv) or _promote(np.asarray(e))
    pI, pJ, pV = [], [], []
    for nw in range(1, K + 1):
        I, J, V = create_partial_info(nw, isfwd[nw - 1], Ng, dlinv[nw - 1], isbloch[nw - 1], e[nw - 1])
        jstr, joff = (K, nw - K) if order_cmpfirst else (1, M * (nw - 1))
        pI.append(I); pJ.append(jstr * J + joff); pV.append(V)
    return sparse_dropzeros(*_cat(pI, pJ, pV, cplx), M, K * M)


def create_grad(isfwd, dlinv, isbloch=None, e=None, *, N=None, order_cmpfirst=True) -> CSC:
    """`create_grad` — matrix/differential/gradient.jl:37-70."""
    K = len(isfwd)
    if N is not None:
        dlinv = _vecs([1.0] * K if dlinv is None else dlinv, N)
    dlinv = [np.asarray(v) for v in dlinv]
    isbloch = [True] * K if isbloch is None else isbloch
    e = [1.0] * K if e is None else e
    Ng = tuple(len(v) for v in dlinv)
    M = int(np.prod(Ng))
    cplx = _promote(*dlinv) or _promote(np.asarray(e))
    pI, pJ, pV = [], [], []
    for nw in range(1, K + 1):
        I, J, V = create_partial_info(nw, isfwd[nw - 1], Ng, dlinv[nw - 1], isbloch[nw - 1], e[nw - 1])
        istr, ioff = (K, nw - K) if order_cmpfirst else (1, M * (nw - 1))
        pI.append(istr * I + ioff); pJ.append(J); pV.append(V)
    return sparse_dropzeros(*_cat(pI, pJ, pV, cplx), K * M, M)


def create_mean(isfwd, dl=None, dlpinv=None, isbloch=None, e=None, *, N=None,
                order_cmpfirst=True) -> CSC:
    """`create_mean` — matrix/mean.jl:72-111 (arithmetic wrapper :48-53: ∆l = ∆l′⁻¹ = ones)."""
    K = len(isfwd)
    if dl is None:
        dl = [np.ones(n) for n in N]; dlpinv = [np.ones(n) for n in N]
    dl = [np.asarray(v) for v in dl]; dlpinv = [np.asarray(v) for v in dlpinv]
    isbloch = [True] * K if isbloch is None else isbloch
    e = [1.0] * K if e is None else e
    Ng = tuple(len(v) for v in dl)
    M = int(np.prod(Ng))
    cplx = _promote(*dl) or _promote(*dlpinv) or _promote(np.asarray(e))
    pI, pJ, pV = [], [], []
    for nw in range(1, K + 1):
        I, J, V = create_minfo(nw, isfwd[nw - 1], Ng, dl[nw - 1], dlpinv[nw - 1], isbloch[nw - 1], e[nw - 1])
        istr, ioff = (K, nw - K) if order_cmpfirst else (1, M * (nw - 1))
        pI.append(istr * I + ioff); pJ.append(istr * J + ioff); pV.append(V)
    return sparse_dropzeros(*_cat(pI, pJ, pV, cplx), K * M, K * M)


def create_picmp(N, permute, *, scale=None, order_cmpfirst=True) -> CSC:
    """`create_πcmp` — matrix/permutation.jl:15-40."""
    M = int(np.prod(N))
    Kf = len(permute)
    scale = np.ones(Kf) if scale is None else np.asarray(scale)
    cplx = np.iscomplexobj(scale)
    pI, pJ, pV = [], [], []
    r = np.arange(1, M + 1, dtype=np.int64)
    for nw in range(1, Kf + 1):
        nv = permute[nw - 1]
        istr, ioff = (Kf, nv - Kf) if order_cmpfirst else (1, M * (nv - 1))
        jstr, joff = (Kf, nw - Kf) if order_cmpfirst else (1, M * (nw - 1))
        pI.append(istr * r + ioff); pJ.append(jstr * r + joff)
        pV.append(JV.of(np.full(M, scale[nw - 1])))
    return sparse_dropzeros(*_cat(pI, pJ, pV, cplx), Kf * M, Kf * M)
