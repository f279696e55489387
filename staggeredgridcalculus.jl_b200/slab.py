"""z-slab decomposition of the matrix-free curl across the GPUs of one node.

One process per GPU.  Each rank owns a contiguous slab of z-planes of every field component.
Only ∂z couples slabs, and only for Fx and Fy (Gx ∋ −∂zFy, Gy ∋ +∂zFx), so one apply needs ONE
z-plane of TWO components from ONE neighbour:

    forward  ∂z : the plane above my top  = plane 0   of rank r+1  (wraps to rank 0 for Bloch)
    backward ∂z : the plane below my bottom = last plane of rank r−1 (wraps to the last rank)

The plane travels with `torch.distributed` point-to-point ops (NCCL over NVLink on GPUs, gloo in
the CPU tests) on a side stream while the interior planes are computed; the single boundary
plane is computed after the halo has landed.  The physical boundary condition stays inside the
kernel: for a Bloch face the ghost plane holds the RAW values of the wrap-around plane and the
kernel applies e⁻ⁱᵏᴸ (forward) or divides by it (backward) exactly as
matrixfree/differential/partial.jl:140 / :226 do; a symmetric face needs no message.

The compute backend is pluggable so that the plan (who sends which plane to whom, and what is
computed when) can be exercised on CPU with world_size 2 (tests/test_slab_gloo.py); the default
backend is the CUDA library and nothing else ships.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .device import DeviceArray
from .matrixfree import Operator

__all__ = ["SlabCurl", "CudaBackend", "halo_plan", "partition_planes", "row_range_for_rank"]


def partition_planes(nz_total: int, world: int):
    """Contiguous, balanced plane ranges [(k0, k1), ...] of the global z extent."""
    cuts = [(nz_total * r) // world for r in range(world + 1)]
    return list(zip(cuts[:-1], cuts[1:]))


def row_range_for_rank(n_rows: int, rank: int, world: int):
    """Row partition of a KM-row matrix for per-GPU CSC blocks (`create_*(…, row_range=…)`)."""
    return (n_rows * rank) // world, (n_rows * (rank + 1)) // world


def halo_plan(rank: int, world: int, isfwd_z: bool, isbloch_z: bool):
    """Who I send my boundary plane to and who I receive my ghost plane from (None = nobody).

    Returns dict(send_to, send_plane ('first'|'last'), recv_from, lo_interface, hi_interface,
    boundary ('top'|'bottom'))."""
    last = world - 1
    if isfwd_z:  # my plane 0 goes down to rank-1; my ghost (above my top) comes from rank+1
        send_to = rank - 1 if rank > 0 else (last if isbloch_z else None)
        recv_from = rank + 1 if rank < last else (0 if isbloch_z else None)
        send_plane, boundary = "first", "top"
    else:  # my last plane goes up to rank+1; my ghost (below my bottom) comes from rank-1
        send_to = rank + 1 if rank < last else (0 if isbloch_z else None)
        recv_from = rank - 1 if rank > 0 else (last if isbloch_z else None)
        send_plane, boundary = "last", "bottom"
    if world == 1:
        send_to = recv_from = None  # the wrap plane is local: the kernel reads it in place
    return dict(send_to=send_to, send_plane=send_plane, recv_from=recv_from, lo_interface=rank > 0,
                hi_interface=rank < last, boundary=boundary)


class CudaBackend:
    """The product backend: libsgc_b200 kernels, NCCL halo on a side stream."""

    def __init__(self, dtype_code, N, isfwd, dlinv, isbloch, e, alpha, lo_iface, hi_iface):
        self.N = tuple(N)
        self.op = Operator.curl(dtype_code, N, isfwd, dlinv, isbloch, e, alpha=alpha)
        self.op.set_slab(lo_iface, hi_iface)
        self.dtype = torch.complex128 if dtype_code == _lib.SGC_C128 else torch.float64
        self.plane = self.N[0] * self.N[1]
        self.comm_stream = torch.cuda.Stream()
        self.esize = 16 if dtype_code == _lib.SGC_C128 else 8

    def alloc_ghost(self):
        return torch.empty(2, self.plane, dtype=self.dtype, device="cuda")

    def plane_view(self, F: DeviceArray, comp: int, k: int):
        M = self.plane * self.N[2]
        return F.t[comp * M + k * self.plane: comp * M + (k + 1) * self.plane]

    def apply_range(self, G, F, op, k0, k1, ghost):
        gp = None
        if ghost is not None:
            gp = [ghost[0].data_ptr(), ghost[1].data_ptr(), None]
        self.op.apply_range(G, F, op, k0, k1, gp)

    # stream plumbing -----------------------------------------------------------------------
    def comm_begin(self):
        self.comm_stream.wait_stream(torch.cuda.current_stream())
        return torch.cuda.stream(self.comm_stream)

    def comm_join(self):
        torch.cuda.current_stream().wait_stream(self.comm_stream)


class SlabCurl:
    """`apply_curl!` on a z-slab-decomposed grid (full 3-D curl, any dtype the library takes).

    `N` and `dlinv[2]` are LOCAL (this rank's slab); `isbloch`/`e` describe the GLOBAL boundary."""

    def __init__(self, dtype_code, N, isfwd, dlinv, isbloch, e, rank, world, alpha=1.0, backend_cls=CudaBackend,
                 group=None):
        self.rank, self.world, self.group = rank, world, group
        self.N = tuple(int(n) for n in N)
        self.plan = halo_plan(rank, world, bool(isfwd[2]), bool(isbloch[2]))
        self.backend = backend_cls(dtype_code, self.N, isfwd, dlinv, isbloch, e, alpha, self.plan["lo_interface"],
                                   self.plan["hi_interface"])
        self.ghost = self.backend.alloc_ghost() if self.plan["recv_from"] is not None else None

    def apply(self, G, F, op="="):
        Nz = self.N[2]
        plan, be = self.plan, self.backend
        if plan["send_to"] is None and plan["recv_from"] is None:
            be.apply_range(G, F, op, 0, Nz, None)  # single GPU, or an isolated symmetric slab
            return
        ksend = 0 if plan["send_plane"] == "first" else Nz - 1
        kb = Nz - 1 if plan["boundary"] == "top" else 0
        with be.comm_begin():
            ops = []
            if plan["send_to"] is not None:
                for c in range(2):
                    ops.append(dist.P2POp(dist.isend, be.plane_view(F, c, ksend), plan["send_to"], self.group))
            if plan["recv_from"] is not None:
                for c in range(2):
                    ops.append(dist.P2POp(dist.irecv, self.ghost[c], plan["recv_from"], self.group))
            reqs = dist.batch_isend_irecv(ops) if ops else []
            for r in reqs:
                r.wait()
        # interior planes overlap with the halo
        if plan["boundary"] == "top":
            be.apply_range(G, F, op, 0, kb, None)
        else:
            be.apply_range(G, F, op, kb + 1, Nz, None)
        be.comm_join()
        be.apply_range(G, F, op, kb, kb + 1, self.ghost)
