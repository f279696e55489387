"""Host-side producers of the hot path's per-axis inputs: `Grid`, `create_stretched_∆l`,
`invert_∆l`.

These O(ΣN) set-up routines are OUT OF SCOPE for the CUDA path and
stay on the host; they are restated here only so that tests and `bench.py` can build the
"256³ nonuniform grid with PML stretch" inputs of BASELINE.json without Julia.
"""
from __future__ import annotations

import numpy as np

__all__ = ["Grid", "PMLParam", "calc_sfactor", "create_stretched_dl", "invert_dl", "graded_lprim", "pml_loc", "DeviceAxes"]


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


def pml_loc(grid: Grid, Npml):
    """`pml_loc` — src/pml.jl:165-174: ((lpml₋[k]), (lpml₊[k])), ((Lpml₋[k]), (Lpml₊[k]))."""
    lp_n, lp_p, L_n, L_p = [], [], [], []
    for k in range(len(grid.N)):
        lprim = grid.l[0][k]
        full = np.concatenate([lprim, [grid.bounds[1][k]]])
        nn, npos = int(Npml[0][k]), int(Npml[1][k])
        a = full[nn]
        b = grid.bounds[1][k] if npos == 0 else lprim[len(lprim) - npos]
        lp_n.append(a); lp_p.append(b)
        L_n.append(a - grid.bounds[0][k]); L_p.append(grid.bounds[1][k] - b)
    return (lp_n, lp_p), (L_n, L_p)


class DeviceAxes:
    """The per-axis vectors of a Grid with PML, produced ON THE DEVICE (`sgc_axis_*`, csrc/sgc_grid.cu): `stretch(ω)`
    gives s∆l / (s∆l)⁻¹ as device tensors, `refresh(op, ω)` rebuilds an operator handle's coefficient tables for a new
    ω without a host round-trip (frequency sweeps).  `which` = 0 for the primal set (∆l at primal points), 1 dual."""

    def __init__(self, grid: Grid, Npml, which: int, pml: "PMLParam | None" = None):
        import ctypes as C
        from ._lib import lib, check
        self._lib, self._check, self._C = lib, check, C
        (lp_n, lp_p), (L_n, L_p) = pml_loc(grid, Npml)
        par = None
        if pml is not None:
            par = (C.c_double * 5)(pml.m, pml.R, pml.kmax, pml.amax, pml.ma)
        self.h, self.n = [], []
        for k in range(len(grid.N)):
            l = np.ascontiguousarray(grid.l[which][k], dtype=np.float64)
            dl = np.ascontiguousarray(grid.dl[which][k], dtype=np.float64)
            h = C.c_void_p()
            check(lib.sgc_axis_create(C.byref(h), len(l), l.ctypes.data_as(C.POINTER(C.c_double)), dl.ctypes.data_as(C.POINTER(C.c_double)),
                                      float(lp_n[k]), float(lp_p[k]), float(L_n[k]), float(L_p[k]), par))
            self.h.append(h)
            self.n.append(len(l))

    def stretch(self, omega, stream=None):
        """[(s∆l_k, (s∆l_k)⁻¹) for every axis k] as complex128 device tensors."""
        import torch
        from .device import stream_handle
        C = self._C
        out = []
        for h, n in zip(self.h, self.n):
            a = torch.empty(n, dtype=torch.complex128, device="cuda")
            b = torch.empty(n, dtype=torch.complex128, device="cuda")
            self._check(self._lib.sgc_axis_stretch(h, float(omega), C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), stream_handle(stream)))
            out.append((a, b))
        return out

    def refresh(self, op, omega, stream=None):
        from .device import stream_handle
        C = self._C
        arr = (C.c_void_p * len(self.h))(*[h.value for h in self.h])
        self._check(self._lib.sgc_op_set_axes(op._h, arr, float(omega), stream_handle(stream)))

    def __del__(self):
        try:
            for h in self.h:
                self._lib.sgc_axis_destroy(h)
        except Exception:
            pass


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
