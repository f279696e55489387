"""Host-side producers of the hot path's per-axis inputs: `Grid`, `create_stretched_∆l`,
`invert_∆l`.

These O(ΣN) set-up routines are OUT OF SCOPE for the CUDA path and
stay on the host; they are restated here only so that tests and `bench.py` can build the
"256³ nonuniform grid with PML stretch" inputs of BASELINE.json without Julia.
"""
from __future__ import annotations

import numpy as np

__all__ = ["Grid", "PMLParam", "calc_sfactor", "create_stretched_dl", "invert_dl", "graded_lprim"]


class Grid:
    """`Grid(lprim, isbloch)` — src/grid.jl:125-207 (only N, L, l, ∆l, bounds are kept)."""

    def __init__(self, lprim, isbloch):
        lprim = [np.asarray(l, dtype=np.float64).copy() for l in lprim]
        for l in lprim:
            if np.any(np.diff(l) < 0):
                raise ValueError("all entry vectors of lprim must be sorted")  # grid.jl:127
        self.isbloch = [bool(b) for b in isbloch]
        ldual = [(l[:-1] + l[1:]) / 2 for l in lprim]  # movingavg, util.jl:50-64
        self.bounds = ([l[0] for l in lprim], [l[-1] for l in lprim])
        self.L = [b1 - b0 for b0, b1 in zip(*self.bounds)]
        for k in range(len(lprim)):  # grid.jl:140-143
            l0 = ldual[k][-1] - self.L[k] if self.isbloch[k] else 2 * self.bounds[0][k] - ldual[k][0]
            ldual[k] = np.concatenate([[l0], ldual[k]])
        dlprim = [np.diff(l) for l in ldual]  # ∆l at primal points   grid.jl:149
        dldual = [np.diff(l) for l in lprim]  # ∆l at dual points     grid.jl:150
        self.dl = (dlprim, dldual)
        self.N = tuple(len(d) for d in dlprim)
        self.l = ([l[:-1] for l in lprim], [l[1:] for l in ldual])  # grid.jl:198-202


class PMLParam:
    """src/pml.jl:66-73."""

    def __init__(self):
        self.m, self.R, self.kmax, self.amax, self.ma = 4.0, np.exp(-16), 1.0, 0.0, 4.0


def calc_sfactor(omega, d, Lpml, pml: PMLParam):
    """`calc_sfactor` — src/pml.jl:143-153:  s = κ + σ/(a + iω)."""
    smax = -(pml.m + 1) * np.log(pml.R) / (2 * Lpml)
    sigma = smax * (d / Lpml) ** pml.m
    kappa = 1 + (pml.kmax - 1) * (d / Lpml) ** pml.m
    a = pml.amax * (1 - d / Lpml) ** pml.ma
    return kappa + sigma / (a + 1j * omega)


def create_stretched_dl(omega, grid: Grid, Npml, pml: PMLParam | None = None):
    """`create_stretched_∆l(ω, grid, Npml, pml)` — src/pml.jl:94-140.  Returns (s∆lprim, s∆ldual)."""
    pml = pml or PMLParam()
    K = len(grid.N)
    out = ([], [])
    for k in range(K):
        lprim = grid.l[0][k]
        full = np.concatenate([lprim, [grid.bounds[1][k]]])
        nn, npos = int(Npml[0][k]), int(Npml[1][k])
        lpml_n = full[nn]  # v[1+n]                                     pml.jl:170
        lpml_p = grid.bounds[1][k] if npos == 0 else lprim[len(lprim) - npos]  # v[end+1-n]
        Ln, Lp = lpml_n - grid.bounds[0][k], grid.bounds[1][k] - lpml_p
        for ngt in (0, 1):
            r = grid.l[ngt][k]
            s = np.ones(len(r), dtype=np.complex128)
            for i, ri in enumerate(r):  # pml.jl:128-133
                if ri < lpml_n:
                    s[i] = calc_sfactor(omega, lpml_n - ri, Ln, pml)
                if ri > lpml_p:
                    s[i] = calc_sfactor(omega, ri - lpml_p, Lp, pml)
            out[ngt].append(s * grid.dl[ngt][k])
    return out


def invert_dl(dl):
    """`invert_∆l` — src/util.jl:42-48."""
    return ([1.0 / v for v in dl[0]], [1.0 / v for v in dl[1]])


def graded_lprim(n, rng, ratio=1.05):
    """Seeded, smoothly graded primal grid with n cells (neighbouring-cell ratio ≤ `ratio`)."""
    steps = rng.uniform(-np.log(ratio), np.log(ratio), size=n)
    d = np.exp(np.clip(np.cumsum(steps) * 0.2, -1.0, 1.0))
    for i in range(1, n):  # enforce the ratio bound
        d[i] = min(max(d[i], d[i - 1] / ratio), d[i - 1] * ratio)
    return np.concatenate([[0.0], np.cumsum(d)])
