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

from ._lib import lib, check
from .device import DeviceArray
from .matrixfree import Operator

__all__ = ["HostCurlCurl"]


class HostCurlCurl:
    def __init__(self, curl_fwd: Operator, curl_bwd: Operator, N, nchunks=8):
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
