"""Shared helpers of the parity tests."""
import numpy as np

# north_star tolerance: "nzval and stencil outputs must agree within 4 ulp / 1e-13 relative
# for Float64 and ComplexF64".  Both implementations execute the same IEEE operations in the
# same order, so the tests additionally report (and, where stated, require) bit equality.
MAX_ULP = 4
RTOL = 1e-13


def _parts(a):
    a = np.asarray(a)
    if np.iscomplexobj(a):
        return np.stack([a.real, a.imag], axis=-1)
    return a.astype(np.float64)[..., None]


def max_ulp_error(got, ref):
    """Largest component-wise distance in units of the ulp of the reference ELEMENT's
    magnitude (|z| for complex), treating +0.0 == -0.0."""
    g, r = _parts(got), _parts(ref)
    mag = np.sqrt((r ** 2).sum(axis=-1, keepdims=True))
    mag = np.maximum(mag, np.finfo(np.float64).tiny)
    ulp = np.spacing(mag)
    return float((np.abs(g - r) / ulp).max()) if g.size else 0.0


def assert_parity(got, ref, what="", exact=False):
    got, ref = np.asarray(got), np.asarray(ref)
    assert got.shape == ref.shape, f"{what}: shape {got.shape} vs {ref.shape}"
    assert np.all(np.isfinite(_parts(got))), f"{what}: non-finite values"
    if np.array_equal(got, ref):  # bit-exact up to the sign of zero
        return 0.0
    err = max_ulp_error(got, ref)
    scale = np.abs(ref).max() if ref.size else 0.0
    worst = np.abs(got - ref).max()
    assert not exact, f"{what}: not bit-identical (max {err:.2f} ulp, abs {worst:.3e})"
    assert err <= MAX_ULP or worst <= RTOL * scale, (
        f"{what}: max error {err:.2f} ulp (abs {worst:.3e}, scale {scale:.3e}) exceeds {MAX_ULP} ulp / {RTOL} rel")
    return err


def rand_field(rng, shape, cplx):
    a = rng.uniform(-1, 1, size=shape)
    if cplx:
        a = a + 1j * rng.uniform(-1, 1, size=shape)
    return np.asfortranarray(a)


def rand_coef(rng, n, cplx):
    v = rng.uniform(0.5, 1.5, size=n)
    if cplx:
        v = v + 1j * rng.uniform(-0.5, 0.5, size=n)
    return v
