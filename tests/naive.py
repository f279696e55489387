"""Independent entry-by-entry constructions, ported from the reference's own tests.

These are the *known answers* the reference's test-suite checks its operators against
(`test/partial.jl:22-61`, `test/mean.jl:35-80`); they share no code with `oracle/` or with
the CUDA library.  Dense numpy matrices, 0-based linear indices in column-major order
(Julia `LinearIndices`).
"""
import numpy as np


def sub2ind(N, sub):
    ind, stride = 0, 1
    for n, s in zip(N, sub):
        ind += s * stride
        stride *= n
    return ind


def ind2sub(N, ind):
    sub = []
    for n in N:
        sub.append(ind % n)
        ind //= n
    return sub


def naive_partial(N, nw, ns, dwinv, isbloch, dtype=np.float64, e=1.0):
    """test/partial.jl:22-61 (1-D), :98-137 (2-D), :174-213 (3-D).  `nw` 1-based, ns = ±1.
    `e` extends the construction to a Bloch phase ≠ 1 the way create_∂ is specified
    (ghost value Fu[Nw+1] = e·Fu[1], Fu[0] = Fu[Nw]/e)."""
    M = int(np.prod(N))
    a = nw - 1
    Nw = N[a]
    D = np.zeros((M, M), dtype=dtype)
    for ind in range(M):
        sub = ind2sub(N, ind)
        d = dwinv[sub[a]]
        D[ind, ind] = -ns * d
        sub2 = list(sub)
        wrapped = False
        if ns == 1:
            if sub2[a] == Nw - 1:
                sub2[a] = 0; wrapped = True
            else:
                sub2[a] += 1
        else:
            if sub2[a] == 0:
                sub2[a] = Nw - 1; wrapped = True
            else:
                sub2[a] -= 1
        v = ns * d
        if wrapped and e != 1.0:
            v = v * (e if ns == 1 else 1.0 / e)
        D[ind, sub2ind(N, sub2)] += v
    if not isbloch:
        mask = np.ones(N, dtype=np.float64, order="F")
        sl = [slice(None)] * len(N)
        sl[a] = 0
        mask[tuple(sl)] = 0
        m = mask.reshape(M, order="F")
        if ns == 1:
            D = D * m[None, :]  # ∂ws * spdiagm(Mask): zero the input at the boundary
        else:
            D = D * m[:, None]  # spdiagm(Mask) * ∂ws: zero the output at the boundary
    return D


def naive_mhat(N, nw, isfwd, dw, dwpinv, isbloch, dtype=np.float64):
    """test/mean.jl:35-80 (1-D) and the 2-D/3-D analogues."""
    M = int(np.prod(N))
    a = nw - 1
    Nw = N[a]
    Mw = np.zeros((M, M), dtype=dtype)
    for ind_on in range(M):
        sub_on = ind2sub(N, ind_on)
        iw = sub_on[a]
        Mw[ind_on, ind_on] = 0.5 * dw[iw] * dwpinv[iw]
        sub_off = list(sub_on)
        if isfwd:
            sub_off[a] = 0 if sub_off[a] == Nw - 1 else sub_off[a] + 1
        else:
            if sub_off[a] == 0:
                if isbloch:
                    sub_off[a] = Nw - 1
            else:
                sub_off[a] -= 1
        ind_off = sub2ind(N, sub_off)
        if (not isfwd) and (not isbloch) and sub_on[a] == 0:
            Mw[ind_on, ind_off] = dw[sub_off[a]] * dwpinv[iw]  # doubled diagonal entry
        else:
            Mw[ind_on, ind_off] = 0.5 * dw[sub_off[a]] * dwpinv[iw]
    if isfwd and not isbloch:
        mask = np.ones(N, dtype=np.float64, order="F")
        sl = [slice(None)] * len(N)
        sl[a] = 0
        mask[tuple(sl)] = 0
        Mw = Mw * mask.reshape(M, order="F")[None, :]
    return Mw


def vec(A):
    """Julia `A[:]` (column-major flattening)."""
    return np.asarray(A).reshape(-1, order="F")


def blk2cmp_perm(M, K=3):
    """r = reshape(collect(1:KM), M, K)'[:]  (test/curl.jl:5), 0-based."""
    return np.arange(K * M).reshape(K, M).T.reshape(-1)
