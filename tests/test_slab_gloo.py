"""world_size-2 (and 3) CPU tests of the z-slab halo plan (`sgc_b200.slab`), gloo backend.

The product's SlabCurl drives the exchange (who sends which plane to whom, interior/boundary
split); the compute backend is replaced by an oracle-based one, so the test checks that the
distributed result equals the undivided oracle result BIT-FOR-BIT for every z boundary
condition.  The CUDA backend of the same class is covered on GPUs by tests/test_gpu_slab.py.
"""
import contextlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import sgc_oracle as O


class OracleBackend:
    """Stands in for CudaBackend: planes [k0,k1) of the local slab computed with the oracle on
    the slab extended by its ghost plane (see the comments for how each face is emulated)."""

    def __init__(self, dtype_code, N, isfwd, dlinv, isbloch, e, alpha, lo_iface, hi_iface):
        self.N, self.isfwd, self.dlinv, self.isbloch, self.e, self.alpha = tuple(N), list(isfwd), dlinv, list(isbloch), list(e), alpha
        self.lo_iface, self.hi_iface = lo_iface, hi_iface
        self.plane = N[0] * N[1]

    def alloc_ghost(self):
        return torch.zeros(2, self.plane, dtype=torch.complex128)

    def plane_view(self, F, comp, k):
        M = self.plane * self.N[2]
        return F[comp * M + k * self.plane: comp * M + (k + 1) * self.plane]

    @contextlib.contextmanager
    def comm_begin(self):
        yield

    def comm_join(self):
        pass

    def apply_range(self, G, F, op, k0, k1, ghost):
        if k1 <= k0:
            return
        Nx, Ny, Nz = self.N
        Fl = F.numpy().reshape((Nx, Ny, Nz, 3), order="F")
        Gl = G.numpy().reshape((Nx, Ny, Nz, 3), order="F")
        fz, bz = self.isfwd[2], self.isbloch[2]
        ez = O.JV.of(self.e[2])
        gh = np.zeros((Nx, Ny, 1, 3), dtype=complex)
        if ghost is not None:
            for c in range(2):
                gh[:, :, 0, c] = ghost[c].numpy().reshape((Nx, Ny), order="F")
        isbloch = list(self.isbloch)
        if fz:
            if self.hi_iface:
                pass  # raw neighbour plane
            elif bz:  # physical Bloch face: ghost = e·(wrap plane)                     partial.jl:140
                g = ez * O.JV.of(gh if ghost is not None else Fl[:, :, 0:1, :])
                gh = g.to_numpy()
            else:  # physical symmetric face: ghost field is zero                          partial.jl:153-162
                gh[...] = 0
            ext = np.concatenate([Fl, gh], axis=2)
            dz = np.concatenate([self.dlinv[2], [1.0]])
            # the lo face rule (Fu[1] := 0) only exists on a physical symmetric lo face
            isbloch[2] = True if (self.lo_iface or bz) else False
            Gext = np.concatenate([Gl, np.zeros_like(gh)], axis=2)
            O.apply_curl(Gext, ext, op, self.isfwd, [self.dlinv[0], self.dlinv[1], dz], isbloch, [self.e[0], self.e[1], 1.0],
                         alpha=self.alpha)
            Gl[:, :, k0:k1, :] = Gext[:, :, k0:k1, :]
        else:
            if self.lo_iface:
                pass
            elif bz:  # ghost = (wrap plane)/e                                             partial.jl:226
                g = O.JV.of(gh if ghost is not None else Fl[:, :, Nz - 1:Nz, :]) / ez
                gh = g.to_numpy()
            else:  # symmetric: the z-term is zero at the boundary ⇒ ghost := own plane   partial.jl:229-231
                gh = Fl[:, :, 0:1, :].copy()
            ext = np.concatenate([gh, Fl], axis=2)
            dz = np.concatenate([[1.0], self.dlinv[2]])
            isbloch[2] = True
            Gext = np.concatenate([np.zeros_like(gh), Gl], axis=2)
            O.apply_curl(Gext, ext, op, self.isfwd, [self.dlinv[0], self.dlinv[1], dz], isbloch, [self.e[0], self.e[1], 1.0],
                         alpha=self.alpha)
            Gl[:, :, k0:k1, :] = Gext[:, :, k0 + 1:k1 + 1, :]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from sgc_b200.slab import SlabCurl, partition_planes
        rng = np.random.default_rng(77)  # same stream on every rank: identical global data
        N = (6, 5, 11)
        F = np.asfortranarray(rng.uniform(-1, 1, N + (3,)) + 1j * rng.uniform(-1, 1, N + (3,)))
        G0 = np.asfortranarray(rng.uniform(-1, 1, N + (3,)) + 1j * rng.uniform(-1, 1, N + (3,)))
        dl = [rng.uniform(0.5, 1.5, n) + 1j * rng.uniform(-0.2, 0.2, n) for n in N]
        e = [np.exp(-0.3j), np.exp(0.5j), np.exp(-0.9j)]
        k0, k1 = partition_planes(N[2], world)[rank]
        ok = True
        for fz in (True, False):
            for bz in (True, False):
                for op in ("=", "+="):
                    isfwd, isbloch = [True, False, fz], [True, False, bz]
                    ref = G0.copy(order="F")
                    O.apply_curl(ref, F, op, isfwd, dl, isbloch, e, alpha=0.75)
                    Nl = (N[0], N[1], k1 - k0)
                    Fl = torch.from_numpy(np.ascontiguousarray(F[:, :, k0:k1, :].reshape(-1, order="F")))
                    Gl = torch.from_numpy(np.ascontiguousarray(G0[:, :, k0:k1, :].reshape(-1, order="F")))
                    sc = SlabCurl(1, Nl, isfwd, [dl[0], dl[1], dl[2][k0:k1]], isbloch, e, rank, world, alpha=0.75,
                                  backend_cls=OracleBackend)
                    sc.apply(Gl, Fl, op)
                    got = Gl.numpy().reshape(Nl + (3,), order="F")
                    ok = ok and np.array_equal(got, ref[:, :, k0:k1, :])
        results[rank] = ok
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_slab_curl_halo_plan_matches_undivided_oracle(world):
    port = _free_port()
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, port, results), nprocs=world, join=True)
    assert all(results.get(r) is True for r in range(world)), dict(results)


def test_halo_plan_table():
    from sgc_b200.slab import halo_plan, partition_planes, row_range_for_rank
    assert partition_planes(10, 3) == [(0, 3), (3, 6), (6, 10)]
    assert [row_range_for_rank(10, r, 4) for r in range(4)] == [(0, 2), (2, 5), (5, 7), (7, 10)]
    p = halo_plan(0, 4, True, True)
    assert (p["send_to"], p["recv_from"], p["lo_interface"], p["hi_interface"]) == (3, 1, False, True)
    p = halo_plan(3, 4, True, False)
    assert (p["send_to"], p["recv_from"], p["hi_interface"]) == (2, None, False)
    p = halo_plan(0, 4, False, False)
    assert (p["send_to"], p["recv_from"]) == (1, None)
    p = halo_plan(0, 1, True, True)
    assert (p["send_to"], p["recv_from"]) == (None, None)
