"""Host-array driver for a curl-of-curl pass: z-chunked, full-duplex pipelined.

The strict drop-in use of the reference passes host `Array`s; on a GPU that path is bound by the
PCIe link (≈55 GB/s each way on the B200 boxes), not by the kernels.  This helper hides as much of
that as the link allows: the input field streams in z-chunks on one stream, the two fused curls
run on each chunk's planes as soon as their z-neighbours have arrived (`sgc_op_apply_range`), and
the result streams out on a third stream — H2D, compute and D2H overlap, so a pass costs about
max(H2D, D2H) instead of their sum.

Dependencies (forward ∂z in the first curl, backward in the second, non-periodic z):
    H[k]  needs E[k], E[k+1]      → after chunk [a,b) arrived:  H planes  [a−1, b−1)   (last chunk: up to Nz)
    E'[k] needs H[k], H[k−1]      → the same plane range, immediately afterwards.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import lib, check
from .device import DeviceArray
from .matrixfree import Operator

__all__ = ["HostCurlCurl", "HostSlabCurlCurl", "bind_to_gpu_numa_node", "pipeline_plan"]


def pipeline_plan(nz: int, nchunks: int, lo_iface: bool = False, hi_iface: bool = False):
    """Plane ranges of the z-chunked curl-of-curl pipeline (forward C₁, backward C₂, non-periodic slab faces).

    Returns (steps, tail): `steps[c] = dict(h2d=(a, b), H=(kb, ke), E2=(kb2, ke))` — after chunk [a, b) of E has
    arrived, H is computed on [kb, ke) (it needs E[k+1]) and E′ on [kb2, ke) (it needs H[k−1]), and E′[kb2:ke] is
    copied out; empty ranges have kb >= ke.  `tail = dict(H=..., E2=[...])` lists the planes that need a
    neighbour rank's data and are finished after the peer barrier: H[nz−1] with an upper interface, then E′[0]
    with a lower interface and E′[nz−1] with an upper interface.  Every plane of H and of E′ appears exactly once."""
    nchunks = max(1, min(int(nchunks), int(nz)))
    cuts = [(nz * c) // nchunks for c in range(nchunks + 1)]
    steps = []
    for c, (a, b) in enumerate(zip(cuts[:-1], cuts[1:])):
        kb = a - 1 if c > 0 else 0
        ke = (nz - 1 if hi_iface else nz) if c == nchunks - 1 else b - 1
        kb2 = max(kb, 1) if lo_iface else kb
        steps.append(dict(h2d=(a, b), H=(kb, ke), E2=(kb2, ke)))
    tail = dict(H=(nz - 1, nz) if hi_iface else None,
                E2=([(0, 1)] if lo_iface else []) + ([(nz - 1, nz)] if hi_iface and not (lo_iface and nz == 1) else []))
    return steps, tail


class HostCurlCurl:
    def __init__(self, curl_fwd: Operator, curl_bwd: Operator, N, nchunks=8):
        # the plane plan below is only valid for: first curl forward in z, second backward, non-periodic z
        f1, f2, bl = getattr(curl_fwd, "isfwd", None), getattr(curl_bwd, "isfwd", None), getattr(curl_fwd, "isbloch", None)
        if f1 is None or f2 is None or not f1[2] or f2[2] or bl[2] or getattr(curl_bwd, "isbloch", bl)[2]:
            raise _lib.SgcError(_lib.SGC_ERR_INVALID, "HostCurlCurl streams z-chunks for a forward-z first curl, a backward-z "
                                "second curl and a non-periodic z boundary; use Operator.apply_host for other combinations")
        self.c1, self.c2 = curl_fwd, curl_bwd
        self.N = tuple(int(n) for n in N)
        self.plane = self.N[0] * self.N[1]
        self.M = self.plane * self.N[2]
        nz = self.N[2]
        nchunks = max(1, min(nchunks, nz))
        cuts = [(nz * c) // nchunks for c in range(nchunks + 1)]
        self.chunks = list(zip(cuts[:-1], cuts[1:]))
        self.s_in, self.s_cmp, self.s_out = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
        self.E = DeviceArray.empty(self.N + (3,))
        self.H = DeviceArray.empty(self.N + (3,))
        self.E2 = DeviceArray.empty(self.N + (3,))

    def _copy_planes(self, fn, dev: DeviceArray, host: torch.Tensor, a, b, stream):
        """planes [a,b) of each of the 3 components (each range is contiguous)."""
        n = (b - a) * self.plane * 16
        if n <= 0:
            return
        for w in range(3):
            off = (w * self.M + a * self.plane) * 16
            d, h = C.c_void_p(dev.ptr + off), C.c_void_p(host.data_ptr() + off)
            if fn is lib.sgc_memcpy_h2d:
                check(fn(d, h, n, C.c_void_p(stream.cuda_stream)))
            else:
                check(fn(h, d, n, C.c_void_p(stream.cuda_stream)))

    def run(self, E_host: torch.Tensor, Out_host: torch.Tensor):
        """E_host, Out_host: pinned complex128 tensors of 3·M elements (column-major fields)."""
        nz = self.N[2]
        last = len(self.chunks) - 1
        start = torch.cuda.Event()
        start.record(torch.cuda.current_stream())
        for s in (self.s_in, self.s_cmp, self.s_out):
            s.wait_event(start)
        for c, (a, b) in enumerate(self.chunks):
            self._copy_planes(lib.sgc_memcpy_h2d, self.E, E_host, a, b, self.s_in)
            arrived = torch.cuda.Event()
            arrived.record(self.s_in)
            self.s_cmp.wait_event(arrived)
            kb = a - 1 if c > 0 else 0
            ke = nz if c == last else b - 1
            if ke > kb:
                self.c1.apply_range(self.H, self.E, "=", kb, ke, None, stream=self.s_cmp)
                self.c2.apply_range(self.E2, self.H, "=", kb, ke, None, stream=self.s_cmp)
                done = torch.cuda.Event()
                done.record(self.s_cmp)
                self.s_out.wait_event(done)
                self._copy_planes(lib.sgc_memcpy_d2h, self.E2, Out_host, kb, ke, self.s_out)
        self.s_out.synchronize()


def bind_to_gpu_numa_node(device_index: int):
    """Pin this process to the CPUs of the NUMA node the GPU hangs off (sysfs), so that pinned host
    buffers allocated afterwards are first-touched on that node and the PCIe copies do not cross the
    socket interconnect.  One process per GPU (torchrun) otherwise lands on arbitrary cores.
    Returns the CPU list it bound to, or None when the topology is not exposed."""
    import os
    try:
        pr = torch.cuda.get_device_properties(device_index)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/local_cpulist") as f:
            txt = f.read().strip()
        cpus = set()
        for part in txt.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return sorted(cpus)
    except Exception:
        return None


class HostSlabCurlCurl:
    """The same z-chunked duplex pipeline for one rank of a z-slab-decomposed grid (N > 1 GPUs, fields in
    a `PeerArena`).  Planes whose stencil stays inside the slab stream through exactly as on one GPU;
    the two planes that need a neighbour's data are finished at the end, after a device-side barrier
    (the neighbour's E has arrived / its H is complete), and copied out last:

        forward C₁ : H[Nz−1] needs E[0] of rank+1          (hi interface)
        backward C₂: E′[0]   needs H[Nz−1] of rank−1       (lo interface)
    """

    def __init__(self, slab_fwd, slab_bwd, E: DeviceArray, H: DeviceArray, E2: DeviceArray, nchunks=16):
        self.s1, self.s2 = slab_fwd, slab_bwd
        assert slab_fwd.arena is not None and slab_fwd.plan["boundary"] == "top" and slab_bwd.plan["boundary"] == "bottom"
        self.N = slab_fwd.N
        self.plane = self.N[0] * self.N[1]
        self.M = self.plane * self.N[2]
        nz = self.N[2]
        nchunks = max(1, min(nchunks, nz))
        cuts = [(nz * c) // nchunks for c in range(nchunks + 1)]
        self.chunks = list(zip(cuts[:-1], cuts[1:]))
        self.s_in, self.s_cmp, self.s_out = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
        self.E, self.H, self.E2 = E, H, E2

    _copy_planes = HostCurlCurl._copy_planes

    def _ghost(self, slab, F):
        plan, Nz = slab.plan, self.N[2]
        if plan["recv_from"] is None:
            return None
        base = slab.arena.peer_ptr(F, plan["recv_from"])
        kpl = 0 if plan["boundary"] == "top" else Nz - 1
        return [base + (c * self.M + kpl * self.plane) * 16 for c in range(2)] + [None]

    def run(self, E_host: torch.Tensor, Out_host: torch.Tensor):
        nz = self.N[2]
        hi_iface = self.s1.plan["recv_from"] is not None   # my top H plane needs the upper neighbour's E
        lo_iface = self.s2.plan["recv_from"] is not None   # my bottom E′ plane needs the lower neighbour's H
        c1, c2 = self.s1.backend.op, self.s2.backend.op
        steps, tail = pipeline_plan(nz, len(self.chunks), lo_iface, hi_iface)
        start = torch.cuda.Event()
        start.record(torch.cuda.current_stream())
        for s in (self.s_in, self.s_cmp, self.s_out):
            s.wait_event(start)
        for st in steps:
            a, b = st["h2d"]
            self._copy_planes(lib.sgc_memcpy_h2d, self.E, E_host, a, b, self.s_in)
            arrived = torch.cuda.Event()
            arrived.record(self.s_in)
            self.s_cmp.wait_event(arrived)
            (kb, ke), (kb2, ke2) = st["H"], st["E2"]
            if ke > kb:
                c1.apply_range(self.H, self.E, "=", kb, ke, None, stream=self.s_cmp)
            if ke2 > kb2:
                c2.apply_range(self.E2, self.H, "=", kb2, ke2, None, stream=self.s_cmp)
                done = torch.cuda.Event()
                done.record(self.s_cmp)
                self.s_out.wait_event(done)
                self._copy_planes(lib.sgc_memcpy_d2h, self.E2, Out_host, kb2, ke2, self.s_out)
        # the interface planes: every rank's E has landed → H top plane; every rank's H is complete → E′
        with torch.cuda.stream(self.s_cmp):
            self.s1.arena.barrier()
            if tail["H"]:
                c1.apply_range(self.H, self.E, "=", tail["H"][0], tail["H"][1], self._ghost(self.s1, self.E), stream=self.s_cmp)
            self.s1.arena.barrier()
            for (k0, k1) in tail["E2"]:
                gh = self._ghost(self.s2, self.H) if (k0 == 0 and lo_iface) else None
                c2.apply_range(self.E2, self.H, "=", k0, k1, gh, stream=self.s_cmp)
            self.s1.arena.barrier()  # nobody overwrites E / H of the next pass while a neighbour still reads them
        done = torch.cuda.Event()
        done.record(self.s_cmp)
        self.s_out.wait_event(done)
        for (k0, k1) in tail["E2"]:
            self._copy_planes(lib.sgc_memcpy_d2h, self.E2, Out_host, k0, k1, self.s_out)
        self.s_out.synchronize()
