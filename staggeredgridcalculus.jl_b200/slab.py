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

Two halo transports exist.  `SlabCurl(...)` alone uses NCCL point-to-point messages as described
above.  `SlabCurl(..., arena=PeerArena(...))` is the NVLink-native form: the fields live in
symmetric (peer-mapped) memory, the ghost pointer handed to the fused kernel IS the neighbour's
plane, and the stencil kernel pulls its halo over NVLink with ordinary loads while it computes —
one kernel per apply, no message, no staging buffer, no interior/boundary split; a device-side
barrier on the stream orders successive applies.  Measured at 2×B200 (256³ per GPU): 0.583 ms per
curl-of-curl step vs 0.631 ms with NCCL messages and 0.560 ms on one GPU.

The compute backend is pluggable so that the plan (who sends which plane to whom, and what is
computed when) can be exercised on CPU with world_size 2 (tests/test_slab_gloo.py); the default
backend is the CUDA library and nothing else ships.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

import ctypes as C

from . import _lib
from ._lib import lib, check
from .device import DeviceArray
from .matrixfree import Operator

__all__ = ["SlabOp", "SlabCurlCurl", "SlabCurl", "CudaBackend", "PeerArena", "halo_plan", "partition_planes", "row_range_for_rank"]


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


class _RawCuda:
    """A raw device allocation seen through `__cuda_array_interface__` (torch only wraps it; the library owns it)."""

    def __init__(self, ptr, n_f64):
        self.__cuda_array_interface__ = {"shape": (int(n_f64),), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 3, "strides": None}


class PeerArena:
    """Peer-mapped device memory for the fields of a slab-decomposed problem plus the library's multi-GPU context
    (`sgc_comm`, csrc/sgc_comm.cu).  Every rank allocates the same number of bytes with `sgc_peer_alloc`
    (plain cudaMalloc memory + a CUDA IPC handle), the 64-byte handles travel through one `all_gather`, and each
    rank maps the others' allocations with `sgc_peer_open` — torch.distributed is only the courier of the
    handles.  Fields are carved out of the arena with `array()`; all ranks must carve the same sequence so that
    offsets agree.  `backend="symm"` (or a failing IPC mapping under `"auto"`) uses
    `torch.distributed._symmetric_memory` for the mapping instead; barrier and neighbour flags are the
    library's either way."""

    def __init__(self, nbytes, group=None, backend="auto"):
        group = group or dist.group.WORLD
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.dev = torch.cuda.current_device()
        nbytes = (int(nbytes) + 255) // 256 * 256
        self.nbytes = nbytes
        self.comm = C.c_void_p()
        check(lib.sgc_comm_init(C.byref(self.comm), self.world, self.rank))
        self.backend = None
        if backend in ("auto", "ipc"):
            err = self._init_ipc(nbytes)   # collective: every rank walks through the same all_gather / all_reduce
            if err is None:
                self.backend = "ipc"
            elif backend == "ipc":
                raise _lib.SgcError(_lib.SGC_ERR_CUDA, f"CUDA IPC peer mapping failed: {err}")
            else:
                import sys
                sys.stderr.write(f"[sgc] CUDA IPC peer mapping unavailable ({err[:160]}); using symmetric memory for the mapping\n")
        if self.backend is None:
            self._init_symm(nbytes)
            self.backend = "symm"
        self.used = 0

    # ---- mappings ---------------------------------------------------------------------------------------------
    def _all_gather_bytes(self, payload: bytes):
        t = torch.frombuffer(bytearray(payload), dtype=torch.uint8).cuda()
        out = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(out, t, group=self.group)
        return [bytes(o.cpu().numpy().tobytes()) for o in out]

    def _all_min(self, ok: bool) -> bool:
        flag = torch.tensor([1 if ok else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        return bool(int(flag.item()))

    def _init_ipc(self, nbytes):
        """Returns None on success, else a description of what failed (on ANY rank)."""
        H = _lib.SGC_IPC_HANDLE_BYTES
        hc, hb, base, err = (C.c_ubyte * H)(), (C.c_ubyte * H)(), C.c_void_p(), None
        try:
            check(lib.sgc_comm_handle(self.comm, hc))
            check(lib.sgc_peer_alloc(self.comm, nbytes, C.byref(base), hb))
        except Exception as ex:
            err = f"rank {self.rank}: {ex}"
        got = self._all_gather_bytes(bytes([0 if err is None else 1]) + bytes(hc) + bytes(hb))
        if any(g[0] for g in got):
            if base.value:
                lib.sgc_free(base)
            return err or "another rank could not export its allocation"
        self.base = int(base.value)
        self.peer_base = [self.base] * self.world
        try:
            check(lib.sgc_comm_connect(self.comm, b"".join(g[1:1 + H] for g in got)))
            for r, g in enumerate(got):
                if r != self.rank:
                    pp = C.c_void_p()
                    check(lib.sgc_peer_open(self.comm, g[1 + H:], C.byref(pp)))
                    self.peer_base[r] = int(pp.value)
            self.buf = torch.as_tensor(_RawCuda(self.base, nbytes // 8), device=f"cuda:{self.dev}")
        except Exception as ex:
            err = f"rank {self.rank}: {ex}"
        if not self._all_min(err is None):
            lib.sgc_comm_destroy(self.comm)
            lib.sgc_free(base)
            self.comm = C.c_void_p()
            check(lib.sgc_comm_init(C.byref(self.comm), self.world, self.rank))
            return err or "another rank could not map its neighbours"
        self._owned = True
        return None

    def _init_symm(self, nbytes):
        import torch.distributed._symmetric_memory as symm
        cb = int(lib.sgc_comm_ctrl_bytes(self.world))
        cb = (cb + 255) // 256 * 256
        self.buf_all = symm.empty((nbytes + cb) // 8, dtype=torch.float64, device=f"cuda:{self.dev}")
        self.hdl = symm.rendezvous(self.buf_all, self.group.group_name)
        ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        arr = (C.c_void_p * self.world)(*ptrs)
        check(lib.sgc_comm_connect_ptrs(self.comm, C.c_void_p(ptrs[self.rank]), arr))
        self.buf = self.buf_all[cb // 8:]
        self.base = self.buf.data_ptr()
        self.peer_base = [p + cb for p in ptrs]
        self._owned = False
        dist.barrier(group=self.group)  # every control block is zeroed before anyone signals

    # ---- carving ------------------------------------------------------------------------------------------------
    def array(self, shape, dtype=np.complex128) -> DeviceArray:
        n = int(np.prod(shape, dtype=np.int64))
        es = 16 if np.dtype(dtype) == np.complex128 else 8
        off = (self.used + 255) // 256 * 256
        if off + n * es > self.nbytes:
            raise _lib.SgcError(_lib.SGC_ERR_NOMEM, "PeerArena exhausted")
        self.used = off + n * es
        flat = self.buf[off // 8: off // 8 + n * es // 8]
        t = torch.view_as_complex(flat.view(-1, 2)) if es == 16 else flat
        a = DeviceArray(t, shape)
        a._arena_keepalive = self
        return a

    def peer_ptr(self, arr: DeviceArray, peer: int) -> int:
        off = arr.ptr - self.base
        if not 0 <= off < self.nbytes:
            raise _lib.SgcError(_lib.SGC_ERR_INVALID, "the field does not live in this PeerArena")
        return self.peer_base[peer] + off

    def barrier(self, stream=None):
        """Device-side barrier of all ranks (`sgc_peer_barrier`), ordered on the current stream."""
        from .device import stream_handle
        check(lib.sgc_peer_barrier(self.comm, stream_handle(stream)))

    def epoch(self) -> int:
        from .device import stream_handle
        v = C.c_int64()
        check(lib.sgc_comm_epoch(self.comm, C.byref(v), stream_handle(None)))
        return int(v.value)

    def timeouts(self) -> int:
        """Device-side waits on a neighbour that gave up (`sgc_comm_timeouts`): 0 in a correct run."""
        from .device import stream_handle
        v = C.c_int64()
        check(lib.sgc_comm_timeouts(self.comm, C.byref(v), stream_handle(None)))
        return int(v.value)

    def assert_same(self, value, what="slab shape"):
        """The peer paths address the neighbour's planes with the LOCAL slab extents: all ranks must agree."""
        vals = [None] * self.world
        dist.all_gather_object(vals, tuple(value), group=self.group)
        if any(v != vals[0] for v in vals):
            raise _lib.SgcError(_lib.SGC_ERR_INVALID, f"{what} differs between ranks ({vals}): the peer-memory halo "
                                "paths need equal slabs (pad nz to a multiple of the world size, or use the NCCL transport)")

    def close(self):
        if getattr(self, "comm", None) is not None and self.comm.value:
            torch.cuda.synchronize()
            base = self.base if self._owned else None
            lib.sgc_comm_destroy(self.comm)   # closes the IPC mappings it opened
            self.comm = C.c_void_p()
            if base:
                lib.sgc_free(C.c_void_p(base))


class SlabCurl:
    """`apply_curl!` on a z-slab-decomposed grid (full 3-D curl, any dtype the library takes).

    `N` and `dlinv[2]` are LOCAL (this rank's slab); `isbloch`/`e` describe the GLOBAL boundary.
    With `arena` the halo plane is read straight from the neighbour's memory by the fused kernel
    (every rank's `F` must then be the same carve-out of the arena, and slabs of equal size)."""

    def __init__(self, dtype_code, N, isfwd, dlinv, isbloch, e, rank, world, alpha=1.0, backend_cls=CudaBackend,
                 group=None, arena: "PeerArena | None" = None, sync="barrier"):
        """`sync` (peer-memory transport only): "barrier" = a device-side barrier of all ranks before every apply
        (safe with anything else touching the fields in between); "flags" = neighbour synchronisation inside the
        stencil kernel (all ranks must then run the same sequence of applies on the arena's fields and nothing
        else may modify those fields in between)."""
        self.rank, self.world, self.group = rank, world, group
        self.sync = sync
        self.bloch_z = bool(isbloch[2])
        self.N = tuple(int(n) for n in N)
        self.plan = halo_plan(rank, world, bool(isfwd[2]), bool(isbloch[2]))
        self.backend = backend_cls(dtype_code, self.N, isfwd, dlinv, isbloch, e, alpha, self.plan["lo_interface"],
                                   self.plan["hi_interface"])
        self.arena = arena if world > 1 else None
        self.esize = 16 if dtype_code == _lib.SGC_C128 else 8
        self.ghost = None
        if self.arena is None and self.plan["recv_from"] is not None:
            self.ghost = self.backend.alloc_ghost()
        if self.arena is not None:
            self.arena.assert_same(self.N)

    def _setup_flag_sync(self):
        """Neighbour ranks along z (ring if the global z boundary is Bloch): `sgc_op_attach_comm`."""
        r, w = self.rank, self.world
        lo = r - 1 if r > 0 else (w - 1 if self.bloch_z and w > 1 else None)
        hi = r + 1 if r < w - 1 else (0 if self.bloch_z and w > 1 else None)
        self.backend.op.attach_comm(self.arena.comm, lo, hi)
        self._flag_sync_ready = True

    def resync(self):
        """After anything other than synchronised applies has touched the fields (host copies, other kernels):
        one all-rank barrier re-establishes the precondition of the in-kernel neighbour synchronisation."""
        if self.arena is not None:
            self.arena.barrier()

    def _apply_peer(self, G, F, op):
        """One fused kernel; the ghost pointers are the neighbour's planes (NVLink loads)."""
        Nz, plan = self.N[2], self.plan
        plane = self.N[0] * self.N[1]
        M = plane * Nz
        ghost = None
        if plan["recv_from"] is not None:
            base = self.arena.peer_ptr(F, plan["recv_from"])
            kpl = 0 if plan["boundary"] == "top" else Nz - 1
            ghost = [base + (c * M + kpl * plane) * self.esize for c in range(2)] + [None]
        if self.sync == "flags":
            # neighbour synchronisation inside the kernel: boundary CTAs wait for the neighbours' previous apply,
            # the last CTA publishes this one — no barrier kernel, the interior never waits
            if not getattr(self, "_flag_sync_ready", False):
                self._setup_flag_sync()
            if not getattr(self.arena, "_synced_once", False):
                self.arena.barrier()  # the inputs of the very first synchronised apply
                self.arena._synced_once = True
            self.backend.op.apply_synced(G, F, op, ghost)
            return
        self.arena.barrier()  # every rank's F is complete, and nobody still reads what G overwrites
        self.backend.op.apply_range(G, F, op, 0, Nz, ghost)

    def apply(self, G, F, op="="):
        Nz = self.N[2]
        plan, be = self.plan, self.backend
        if self.arena is not None:
            return self._apply_peer(G, F, op)
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


class SlabCurlCurl:
    """Fused curl-of-curl `G OP C₂(C₁ F)` on a z-slab-decomposed grid: ONE kernel per rank and pass, the
    intermediate field never leaves shared memory, and the planes −1 and Nz of F that the slab's edge needs are
    read straight from the neighbours' memory (`PeerArena`) — at interior interfaces and, for a Bloch z boundary,
    at the physical faces whose wrap-around partner is another rank.  Needs opposite z-directions in the two
    curls (halo reach of one plane) and slabs of equal size.

    `N`, `dlinv1[2]`, `dlinv2[2]` are LOCAL; `dz1_below` / `dz1_above` are the entries of the GLOBAL ∆z₁⁻¹ just
    below / above this rank's slab (wrapping around for Bloch; None at a symmetric physical face)."""

    def __init__(self, dtype_code, N, isfwd1, dlinv1, isfwd2, dlinv2, isbloch, e, rank, world, arena: "PeerArena | None",
                 dz1_below=None, dz1_above=None, alpha1=1.0, alpha2=1.0):
        if bool(isfwd1[2]) == bool(isfwd2[2]) and world > 1:
            raise _lib.SgcError(_lib.SGC_ERR_INVALID, "slab-decomposed curl-of-curl needs opposite z-directions in the two curls")
        self.rank, self.world, self.arena = rank, world, arena if world > 1 else None
        self.N = tuple(int(n) for n in N)
        self.esize = 16 if dtype_code == _lib.SGC_C128 else 8
        self.c1 = Operator.curl(dtype_code, self.N, isfwd1, dlinv1, isbloch, e, alpha=alpha1)
        self.c2 = Operator.curl(dtype_code, self.N, isfwd2, dlinv2, isbloch, e, alpha=alpha2)
        last = world - 1
        for c in (self.c1, self.c2):
            c.set_slab(lo_interface=rank > 0, hi_interface=rank < last)
        ring = bool(isbloch[2]) and world > 1
        self.lo = rank - 1 if rank > 0 else (last if ring else None)   # who owns plane −1
        self.hi = rank + 1 if rank < last else (0 if ring else None)    # who owns plane Nz
        if world > 1:
            if (self.lo is not None and dz1_below is None) or (self.hi is not None and dz1_above is None):
                raise _lib.SgcError(_lib.SGC_ERR_INVALID, "dz1_below / dz1_above are needed wherever a neighbour rank exists")
            self.c1.set_slab_coef(dz1_below if self.lo is not None else None, dz1_above if self.hi is not None else None)
        if self.arena is not None:
            self.arena.assert_same(self.N)

    def apply(self, G, F, op="="):
        if self.arena is None:
            self.c2.apply_after(self.c1, G, F, op)
            return
        plane = self.N[0] * self.N[1]
        M = plane * self.N[2]
        self.arena.barrier()  # every rank's F is complete, and nobody still reads what G overwrites
        glo = ghi = None
        if self.lo is not None:
            base = self.arena.peer_ptr(F, self.lo)
            glo = [base + (c * M + (self.N[2] - 1) * plane) * self.esize for c in range(3)]
        if self.hi is not None:
            base = self.arena.peer_ptr(F, self.hi)
            ghi = [base + c * M * self.esize for c in range(3)]
        self.c2.apply_after(self.c1, G, F, op, ghost_lo=glo, ghost_hi=ghi)


class SlabOp:
    """Any matrix-free operator of the library (`Operator.partial / mhat / mean / divg / grad / curl`) on a grid that
    is slab-decomposed along its LAST axis, with the one-plane halo read from the neighbour's memory (`PeerArena`):
    `apply_∂!` / `apply_m̂!` along the slab axis need one plane of their input, `apply_divg!` one plane of the last
    component, `apply_grad!` one plane of f; operators acting along other axes need nothing.

    `op` is built on the LOCAL grid with the local slice of the slab-axis vectors; `isfwd_last` / `isbloch_last`
    are the direction and the GLOBAL boundary type of the term(s) acting along the slab axis (None when the
    operator has no such term); `ghost_comps` lists the input components those terms read."""

    def __init__(self, op: Operator, isfwd_last, isbloch_last, ghost_comps, rank, world, arena: "PeerArena | None"):
        if world > 1 and getattr(op, "weighted_axis", None) == len(op.N):
            # the boundary weights ∆w[1] / ∆w[N] of mean.jl:516-527, :633-645 belong to the neighbour rank's slice
            raise _lib.SgcError(_lib.SGC_ERR_INVALID, "a weighted mean along the slab axis needs the neighbours' ∆w: "
                                "call op.set_slab_weights(dw_below, dw_above) first")
        self.op, self.rank, self.world = op, rank, world
        self.arena = arena if world > 1 else None
        self.ghost_comps = tuple(ghost_comps) if isfwd_last is not None else ()
        self.plan = halo_plan(rank, world, bool(isfwd_last), bool(isbloch_last)) if isfwd_last is not None else None
        op.set_slab(lo_interface=rank > 0, hi_interface=rank < world - 1)
        self.esize = 16 if op.dtype_code == _lib.SGC_C128 else 8
        if self.arena is not None:
            self.arena.assert_same(op.N)

    def apply(self, G, F, op="="):
        N = self.op.N
        n_last = N[-1]
        if self.arena is None:
            self.op.apply_range(G, F, op, 0, n_last, None)
            return
        M = int(np.prod(N, dtype=np.int64))
        plane = M // n_last
        self.arena.barrier()  # every rank's F is complete, and nobody still reads what G overwrites
        ghost = None
        if self.plan is not None and self.plan["recv_from"] is not None:
            base = self.arena.peer_ptr(F, self.plan["recv_from"])
            kpl = 0 if self.plan["boundary"] == "top" else n_last - 1
            ghost = [base + (c * M + kpl * plane) * self.esize if c in self.ghost_comps else None for c in range(self.op.n_in)]
        self.op.apply_range(G, F, op, 0, n_last, ghost)
