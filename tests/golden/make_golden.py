"""Generates tests/golden/*.npz — small known-answer fixtures for the hot path.

The reference is pure Julia and cannot be run in this image, and its test-suite stores no golden
vectors (every test builds a naive matrix in-line, with unseeded rand).  These fixtures are
therefore produced by the *reference's own known-answer constructions*, re-implemented
independently of the oracle and of the CUDA library in tests/naive.py (entry-by-entry matrices
of test/partial.jl:22-61 and test/mean.jl:35-80, the block identity of test/curl.jl:19-21), on
seeded inputs:

  matrices  : dense ∂w / m̂w / Curl — compared with `==` (test/partial.jl:64, test/mean.jl:83,
              test/curl.jl:47)
  apply     : (naive matrix) * f — compared with `≈` (rtol √eps), as test/partial.jl:67-70 does

Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.naive import naive_partial, naive_mhat, vec  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    rng = np.random.default_rng(20260)
    N = (4, 3, 5)
    M = int(np.prod(N))
    dl = [rng.uniform(0.5, 1.5, n) for n in N]          # ∆l⁻¹ (and ∆l for m̂)
    dlp = [rng.uniform(0.5, 1.5, n) for n in N]         # ∆l′⁻¹
    e = np.exp(-2j * np.pi * np.array([0.1, 0.2, 0.3]))  # Bloch phases of BASELINE config 1
    F = rng.uniform(-1, 1, N + (3,)) + 1j * rng.uniform(-1, 1, N + (3,))
    data = dict(N=np.array(N), dl0=dl[0], dl1=dl[1], dl2=dl[2], dlp0=dlp[0], dlp1=dlp[1], dlp2=dlp[2], e=e, F=F)
    for nw in (1, 2, 3):
        for ns in (-1, 1):
            for isbloch in (True, False):
                key = f"nw{nw}_ns{ns}_b{int(isbloch)}"
                D = naive_partial(N, nw, ns, dl[nw - 1], isbloch, dtype=complex if isbloch else float,
                                  e=e[nw - 1] if isbloch else 1.0)
                data["partial_" + key] = D
                data["partial_apply_" + key] = D @ vec(F[..., 0])
                Mw = naive_mhat(N, nw, ns == 1, dl[nw - 1], dlp[nw - 1], isbloch)  # e = 1
                data["mhat_" + key] = Mw
                data["mhat_apply_" + key] = Mw @ vec(F[..., 1])
    # curl = [0 -∂z ∂y; ∂z 0 -∂x; -∂y ∂x 0]  (test/curl.jl:19-21), block ordering, mixed BCs
    isbloch = [True, False, True]
    for tag, isfwd in (("fff", (True, True, True)), ("bbb", (False, False, False)), ("fbf", (True, False, True))):
        d = [naive_partial(N, nw, 1 if isfwd[nw - 1] else -1, dl[nw - 1], isbloch[nw - 1], dtype=complex,
                           e=e[nw - 1] if isbloch[nw - 1] else 1.0) for nw in (1, 2, 3)]
        Z = np.zeros((M, M), dtype=complex)
        C = np.block([[Z, -d[2], d[1]], [d[2], Z, -d[0]], [-d[1], d[0], Z]])
        data["curl_" + tag] = C
        data["curl_apply_" + tag] = C @ vec(F)
    np.savez_compressed(os.path.join(OUT, "sgc_golden_4x3x5.npz"), **data)
    print("wrote", os.path.join(OUT, "sgc_golden_4x3x5.npz"), len(data), "arrays")


if __name__ == "__main__":
    main()
