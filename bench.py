#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native StaggeredGridCalculus hot path.

Metric (BASELINE.json): curl apply in Gcell-updates/s, ComplexF64, next to the HBM roofline.
Workload at N GPUs (BASELINE.json configs[1], z-slab decomposed for N > 1, weak scaling):
a 256×256×(256·N) nonuniform Yee grid with PML-stretched (complex) ∆l, symmetric (PEC) outer
boundaries, and one matrix-free curl-of-curl pass per step,

        H = C_dual E        (apply_curl!, isfwd = [T,T,T], ∆l⁻¹ at dual points)
        E' = C_prim H       (apply_curl!, isfwd = [F,F,F], ∆l⁻¹ at primal points),

i.e. TWO curl applications = 2·M cell-updates per step, fields resident in HBM (`value`), and
the same pass driven from pinned host arrays with the copies inside the timed region (`e2e`).

  python bench.py [--gpus N] [--steps K] [--warmup W]          # our arm (CUDA library)
  python bench.py --impl reference [...]                        # the reference's CPU algorithm
                                                                # (C/OpenMP port, oracle/)
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NLOCAL = (256, 256, 256)   # per-GPU slab of the headline workload (BASELINE configs[1])
CFG5 = (1024, 1024, 128)   # per-GPU slab of BASELINE configs[4]: 1024^3 over 8 GPUs
NPML = 10
ALG_BYTES_PER_CELL = 96    # one curl, OP '=': read 3 + write 3 ComplexF64 components


def load_gridgen():
    """gridgen.py as a standalone module (numpy only): the reference arm builds its inputs without importing the
    package, i.e. without loading libsgc_b200.so."""
    import importlib.util
    path = os.path.join(ROOT, "staggeredgridcalculus.jl_b200", "gridgen.py")
    spec = importlib.util.spec_from_file_location("sgc_gridgen_standalone", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def build_inputs(nz_total, seed=20260, nlocal=None):
    """Seeded nonuniform grid + PML stretch (host, O(ΣN)); returns ∆l⁻¹ at (primal, dual) points."""
    GG = load_gridgen()
    nl = nlocal or NLOCAL
    rng = np.random.default_rng(seed)
    lprim = [GG.graded_lprim(nl[0], rng), GG.graded_lprim(nl[1], rng), GG.graded_lprim(nz_total, rng)]
    grid = GG.Grid(lprim, [False, False, False])
    mean_cell = np.mean([np.mean(np.diff(l)) for l in lprim])
    omega = 2 * np.pi / (20 * mean_cell)
    sdl = GG.create_stretched_dl(omega, grid, ([NPML] * 3, [NPML] * 3))
    return GG.invert_dl(sdl)


def ulp_report(got, ref):
    """(max error in ulp of the reference element's magnitude, number of elements that differ); +0.0 == -0.0.
    Same metric as tests/util.py::max_ulp_error, evaluated on the differing elements only."""
    if np.array_equal(got, ref):
        return 0.0, 0
    bad = got != ref
    g, r = got[bad], ref[bad]
    mag = np.maximum(np.abs(r), np.finfo(np.float64).tiny)
    ulp = np.spacing(mag)
    err = np.maximum(np.abs(g.real - r.real), np.abs(g.imag - r.imag)) / ulp
    return float(err.max()), int(bad.sum())


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons while the timed region runs (pynvml)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self.stop_flag:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((time.perf_counter(), mhz, r))
            except Exception:
                pass
            time.sleep(0.002)
        self.names = names

    def result(self, t0=None, t1=None):
        """Median SM clock and the throttle reasons seen between the host times t0 and t1 (the timed region; the thread
        is started before the barriers in front of it so that its set-up cost is not inside it)."""
        self.stop_flag = True
        self.join(timeout=1.0)
        inside = [q for q in self.samples if (t0 is None or q[0] >= t0) and (t1 is None or q[0] <= t1)] or self.samples
        if not inside:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": []}
        reasons = set()
        for _, _, r in inside:
            for bit, name in getattr(self, "names", {}).items():
                if r & bit:
                    reasons.add(name)
        return {"sm_mhz": float(np.median([q[1] for q in inside])), "sm_max_mhz": self.max_mhz, "reasons": sorted(reasons),
                "samples": len(inside)}


def cpu_reference_rate(dlinv_prim, dlinv_dual, nz, target_seconds, threads=None):
    """Times the C/OpenMP restatement of the reference's algorithm (oracle/) on the host cores:
    curl-of-curl on a 256×256×nz ComplexF64 grid.  Returns (Gcell/s, threads, description)."""
    from oracle import c_oracle as CO
    threads = threads or (len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count())
    CO.set_threads(threads)
    N = (NLOCAL[0], NLOCAL[1], nz)
    M = N[0] * N[1] * N[2]
    rng = np.random.default_rng(1)
    E = np.asfortranarray(rng.uniform(-1, 1, N + (3,)) + 1j * rng.uniform(-1, 1, N + (3,)))
    H = np.zeros_like(E)
    E2 = np.zeros_like(E)
    dd = [dlinv_dual[0], dlinv_dual[1], dlinv_dual[2][:nz]]
    dp = [dlinv_prim[0], dlinv_prim[1], dlinv_prim[2][:nz]]
    bl, e = [False] * 3, [1.0] * 3

    def step():
        CO.apply_curl(H, E, "=", [True] * 3, dd, bl, e)
        CO.apply_curl(E2, H, "=", [False] * 3, dp, bl, e)

    step()  # warm-up (page faults)
    n, t0 = 0, time.perf_counter()
    while True:
        step()
        n += 1
        dt = time.perf_counter() - t0
        if dt >= target_seconds or n >= 2000:
            break
    rate = 2 * M * n / dt / 1e9
    return rate, threads, f"{n} curl-of-curl passes on 256x256x{nz} ComplexF64 (same inputs family), {dt:.1f} s wall"


def run_reference(args, rank, world):
    if rank != 0:
        return
    nz = NLOCAL[2]  # one rank's full 256^3 grid of the workload; a step is one curl-of-curl pass over it
    dprim, ddual = build_inputs(NLOCAL[2] * max(1, args.gpus))
    # the driver's K/W apply to "steps" of the bounded sample
    from oracle import c_oracle as CO
    threads = len(os.sched_getaffinity(0))
    CO.set_threads(threads)
    N = (NLOCAL[0], NLOCAL[1], nz)
    M = N[0] * N[1] * N[2]
    rng = np.random.default_rng(1)
    E = np.asfortranarray(rng.uniform(-1, 1, N + (3,)) + 1j * rng.uniform(-1, 1, N + (3,)))
    H, E2 = np.zeros_like(E), np.zeros_like(E)
    dd = [ddual[0], ddual[1], ddual[2][:nz]]
    dp = [dprim[0], dprim[1], dprim[2][:nz]]
    bl, e = [False] * 3, [1.0] * 3

    def step():
        CO.apply_curl(H, E, "=", [True] * 3, dd, bl, e)
        CO.apply_curl(E2, H, "=", [False] * 3, dp, bl, e)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = 2 * M * args.steps / dt / 1e9
    sample = f"{args.steps} curl-of-curl passes on the 256x256x{nz} grid of one rank, {dt:.2f} s wall"
    out = {
        "impl": "reference", "metric": "curl_apply_gcell_per_s", "value": val, "unit": "Gcell/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "c128", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": val, "unit": "Gcell/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "Gcell/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "C/OpenMP restatement of the reference's six-sweep apply_curl! (Julia is not available in this image)",
    }
    print(json.dumps(out), flush=True)


def workload_config(n, nl=None):
    nl = nl or NLOCAL
    fld_mb = nl[0] * nl[1] * nl[2] * 3 * 16 / 1e6
    return {"workload": f"{nl[0]}x{nl[1]}x{nl[2] * n} nonuniform Yee grid, PML-stretched complex dl (Npml=10), PEC outer "
                        f"boundaries, ComplexF64 curl-of-curl E->H->E' = 2 apply_curl! per step"
                        + (f", z-slab decomposed over {n} GPUs ({nl[0]}x{nl[1]}x{nl[2]} per GPU) with one-plane halos" if n > 1 else ""),
            "cells_per_gpu": nl[0] * nl[1] * nl[2], "curl_applies_per_step": 2,
            "l2_policy": f"inputs_larger_than_L2 (3 fields x {fld_mb:.0f} MB per GPU vs 126 MB L2)",
            "parallelism": f"zslab{n}"}


class Leg:
    """One weak-scaling leg of the benchmark: a per-GPU slab `nl` of a grid that is z-slab decomposed over `world`
    ranks, the two curl operators of the curl-of-curl pass, and the three fields E, H, E'."""

    def __init__(self, S, torch, dist, nl, rank, world, args, seed_shift=0):
        from sgc_b200.slab import SlabCurl
        self.S, self.torch, self.dist = S, torch, dist
        self.nl, self.rank, self.world = tuple(nl), rank, world
        self.M = nl[0] * nl[1] * nl[2]
        self.nz_total = nl[2] * world
        self.dprim, self.ddual = build_inputs(self.nz_total, nlocal=nl)
        self.k0, self.k1 = rank * nl[2], (rank + 1) * nl[2]
        k0, k1, M = self.k0, self.k1, self.M
        bl, e = [False] * 3, [1.0] * 3
        # N > 1: fields live in peer-mapped memory (sgc_comm, CUDA IPC) and the fused kernel reads its one-plane
        # halo straight from the neighbour over NVLink; NCCL send/recv is the fallback transport.
        self.arena, self.halo = None, "none"
        if world > 1 and args.halo == "peer":
            try:
                from sgc_b200.slab import PeerArena
                self.arena = PeerArena(3 * 3 * M * 16 + 4096)
                self.halo = f"peer-memory loads inside the stencil kernel (sgc_comm, mapping: {self.arena.backend}, NVLink)"
            except Exception as ex:
                sys.stderr.write(f"[bench] peer memory unavailable ({type(ex).__name__}: {str(ex)[:160]}); using NCCL halos\n")
        if world > 1 and self.arena is None:
            self.halo = "NCCL send/recv on a side stream, overlapped with the interior planes"
        self.sync = args.sync if self.arena is not None else "barrier"
        if self.arena is not None:
            self.halo += ("; neighbour synchronisation inside the kernel (flags over NVLink, boundary chunks last), no barrier kernel"
                          if self.sync == "flags" else "; sgc_peer_barrier kernel before every apply")
        self.c_dual = SlabCurl(S.SGC_C128, nl, [True] * 3, [self.ddual[0], self.ddual[1], self.ddual[2][k0:k1]], bl, e, rank, world,
                               arena=self.arena, sync=self.sync)
        self.c_prim = SlabCurl(S.SGC_C128, nl, [False] * 3, [self.dprim[0], self.dprim[1], self.dprim[2][k0:k1]], bl, e, rank, world,
                               arena=self.arena, sync=self.sync)
        gen = torch.Generator(device="cuda").manual_seed(1234 + rank + seed_shift)

        def rnd():
            return torch.complex(torch.rand(3 * M, generator=gen, device="cuda", dtype=torch.float64) * 2 - 1,
                                 torch.rand(3 * M, generator=gen, device="cuda", dtype=torch.float64) * 2 - 1)
        if self.arena is not None:
            self.E, self.H, self.E2 = (self.arena.array(self.nl + (3,)) for _ in range(3))
            self.E.t.copy_(rnd()); self.H.t.zero_(); self.E2.t.zero_()
        else:
            self.E = S.DeviceArray(rnd(), self.nl + (3,))
            self.H = S.DeviceArray.zeros(self.nl + (3,))
            self.E2 = S.DeviceArray.zeros(self.nl + (3,))
        self.barrier()

    def step(self):
        self.c_dual.apply(self.H, self.E, "=")
        self.c_prim.apply(self.E2, self.H, "=")

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def resync(self):
        """Something other than the synchronised applies touched the fields (or read them on the host)."""
        self.barrier()
        if self.arena is not None:
            self.arena.barrier()

    # ---- parity: this rank's planes against the C oracle (reference algorithm on the host) ------------------------
    def parity(self, ranges=None):
        """Runs one pass and compares E' on the local plane ranges [a, b) (default: the whole slab) with the C
        restatement of the reference (six sweeps per apply_curl!, oracle/sgc_oracle_c.c) evaluated on the same E:
        the oracle sees the planes [a−2, b+2) of the GLOBAL grid (the neighbours' edge planes are gathered), so
        slab interfaces are checked against a computation that knows nothing about slabs."""
        torch, dist = self.torch, self.dist
        from oracle import c_oracle as CO
        CO.set_threads(max(1, len(os.sched_getaffinity(0)) // self.world))
        nx, ny, nz = self.nl
        plane = nx * ny
        self.resync()
        self.step()
        self.resync()
        E = self.E.numpy()
        got = self.E2.numpy()
        Et = self.E.t.view(3, nz, plane)
        edge = torch.stack([Et[:, :2, :], Et[:, nz - 2:, :]]).contiguous()  # (lo/hi side, comp, 2 planes, plane)
        edges = [edge]
        if self.world > 1:
            edges = [torch.empty_like(edge) for _ in range(self.world)]
            dist.all_gather(edges, edge)

        def nb_planes(r, side):  # (nx, ny, 2, 3) planes at the lower / upper end of rank r's slab
            a = edges[r][side].cpu().numpy()  # (3, 2, plane)
            return np.asfortranarray(np.transpose(a.reshape(3, 2, ny, nx), (3, 2, 1, 0)))

        worst, ndiff, cells = 0.0, 0, 0
        for (a, b) in (ranges or [(0, nz)]):
            ga, gb = self.k0 + a, self.k0 + b
            lo, hi = max(ga - 2, 0), min(gb + 2, self.nz_total)
            parts = []
            if lo < self.k0:
                parts.append(nb_planes(self.rank - 1, 1)[:, :, 2 - (self.k0 - lo):, :])
            parts.append(E[:, :, max(lo, self.k0) - self.k0:min(hi, self.k1) - self.k0, :])
            if hi > self.k1:
                parts.append(nb_planes(self.rank + 1, 0)[:, :, :hi - self.k1, :])
            Ex = np.asfortranarray(np.concatenate(parts, axis=2))
            Hx, E2x = np.zeros_like(Ex), np.zeros_like(Ex)
            dd = [self.ddual[0], self.ddual[1], self.ddual[2][lo:hi]]
            dp = [self.dprim[0], self.dprim[1], self.dprim[2][lo:hi]]
            CO.apply_curl(Hx, Ex, "=", [True] * 3, dd, [False] * 3, [1.0] * 3)
            CO.apply_curl(E2x, Hx, "=", [False] * 3, dp, [False] * 3, [1.0] * 3)
            err, nd = ulp_report(got[:, :, a:b, :], E2x[:, :, ga - lo:gb - lo, :])
            worst, ndiff, cells = max(worst, err), ndiff + nd, cells + (b - a) * plane
        t = torch.tensor([worst, -float(ndiff), -float(cells)], device="cuda", dtype=torch.float64)
        if self.world > 1:
            tmax = t.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            worst, ndiff, cells = float(tmax[0]), int(-t[1]), int(-t[2])
        return {"against": "C restatement of the reference's apply_curl! (oracle/sgc_oracle_c.c) on the same E, all ranks",
                "max_ulp": worst, "cells": cells, "elements_differing": ndiff, "tolerance_ulp": 4, "ok": bool(worst <= 4.0)}

    # ---- timing ---------------------------------------------------------------------------------------------------
    def time(self, steps, warmup, use_graph, local_rank):
        torch, dist, S = self.torch, self.dist, self.S
        self.resync()
        t_w = time.perf_counter()
        for _ in range(warmup):
            self.step()
        self.barrier()
        # W steps are a floor, not a duration: after the oracle comparison (seconds of CPU work, idle GPUs and NVLink
        # links) a few milliseconds of stepping do not bring clocks and links back up — keep stepping for ~0.2 s
        per_step = max((time.perf_counter() - t_w) / max(warmup, 1), 1e-5)
        n_more = min(2000, int(0.2 / per_step))
        if self.world > 1:
            # every rank MUST run the same number of synchronised applies (a rank that runs one more spins in its
            # boundary CTAs for a neighbour that never comes): agree on the largest count
            cnt = torch.tensor([n_more], device="cuda", dtype=torch.int64)
            dist.all_reduce(cnt, op=dist.ReduceOp.MAX)
            n_more = int(cnt.item())
        for _ in range(n_more):
            self.step()
        self.barrier()
        graph, launches_per_step, run_step = None, None, self.step
        if use_graph:
            try:
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    g = torch.cuda.CUDAGraph()
                    l0 = S.launch_count()
                    with torch.cuda.graph(g, stream=side, capture_error_mode="thread_local"):
                        self.step()
                    launches_per_step = S.launch_count() - l0
                torch.cuda.current_stream().wait_stream(side)
                graph = g
            except Exception as ex:  # capture unsupported for some op: stay eager
                sys.stderr.write(f"[bench] CUDA-graph capture failed ({type(ex).__name__}: {str(ex)[:160]}); running eagerly\n")
                graph = None
                torch.cuda.synchronize()
            ok = torch.tensor([1 if graph is not None else 0], device="cuda")
            if self.world > 1:  # all ranks must take the same path
                dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 0:
                graph = None
            if graph is not None:
                run_step = graph.replay
                for _ in range(max(2, warmup)):
                    run_step()
        # the clock sampler (NVML initialisation: milliseconds, different on every rank) starts BEFORE the barriers: nothing
        # but the event record sits between the device-side barrier and the first step, so no rank starts its timed
        # region late and makes its neighbours' boundary CTAs wait inside theirs
        launches0 = S.launch_count()
        sampler = ClockSampler(local_rank)
        sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        if self.arena is not None:
            # the host barrier releases the ranks up to a millisecond apart; a device-side barrier in front of the
            # first event lines the GPUs up, so that no rank's timed region starts with a wait for a late neighbour
            self.arena.barrier()
        ev0.record()
        t_host0 = time.perf_counter()
        for _ in range(steps):
            run_step()
        host_enqueue_ms = (time.perf_counter() - t_host0) * 1e3 / steps
        ev1.record()
        ev1.synchronize()
        t_host1 = time.perf_counter()
        self.barrier()
        ms_mine = ev0.elapsed_time(ev1) / steps
        clocks = sampler.result(t_host0, t_host1)
        launches = (launches_per_step * steps) if graph is not None else (S.launch_count() - launches0)
        per_rank = [ms_mine]
        if self.world > 1:
            t = torch.tensor([ms_mine], device="cuda", dtype=torch.float64)
            allt = [torch.empty_like(t) for _ in range(self.world)]
            dist.all_gather(allt, t)
            per_rank = [float(x.item()) for x in allt]
        ms_step = max(per_rank)
        return {"ms_step": ms_step, "per_rank_ms": [round(x, 4) for x in per_rank], "clocks": clocks, "launches": int(launches),
                "graph": graph is not None, "host_enqueue_ms": host_enqueue_ms,
                "value": 2 * self.M * self.world / (ms_step * 1e-3) / 1e9}

    def close(self):
        self.barrier()
        self.E = self.H = self.E2 = None
        self.c_dual = self.c_prim = None
        if self.arena is not None:
            self.arena.close()
            self.arena = None
        self.torch.cuda.empty_cache()


def small_bloch_parity(S, torch, dist, rank, world):
    """A small all-Bloch grid (phases on every axis, z ring across the ranks): each rank's slab of E' = C₂(C₁ E),
    through the same peer-memory / in-kernel-synchronised path as the timed step, against the C oracle on the whole
    grid — exercises the wrap-around halo (raw plane from the far rank, phase applied in the kernel)."""
    from oracle import c_oracle as CO
    from sgc_b200.slab import SlabCurl, PeerArena
    CO.set_threads(max(1, len(os.sched_getaffinity(0)) // world))
    nl = (64, 48, 8)
    N = (nl[0], nl[1], nl[2] * world)
    rng = np.random.default_rng(77)
    E = np.asfortranarray(rng.uniform(-1, 1, N + (3,)) + 1j * rng.uniform(-1, 1, N + (3,)))
    d1 = [rng.uniform(0.5, 1.5, n) + 1j * rng.uniform(-0.2, 0.2, n) for n in N]
    d2 = [rng.uniform(0.5, 1.5, n) + 1j * rng.uniform(-0.2, 0.2, n) for n in N]
    e = list(np.exp(-2j * np.pi * np.array([0.1, 0.2, 0.3])))
    bl = [True] * 3
    H, ref = np.zeros_like(E), np.zeros_like(E)
    CO.apply_curl(H, E, "=", [True] * 3, d1, bl, e)
    CO.apply_curl(ref, H, "=", [False] * 3, d2, bl, e)
    k0, k1 = rank * nl[2], (rank + 1) * nl[2]
    arena = PeerArena(3 * 3 * nl[0] * nl[1] * nl[2] * 16 + 4096) if world > 1 else None
    sync = "flags" if arena is not None else "barrier"
    c1 = SlabCurl(S.SGC_C128, nl, [True] * 3, [d1[0], d1[1], d1[2][k0:k1]], bl, e, rank, world, arena=arena, sync=sync)
    c2 = SlabCurl(S.SGC_C128, nl, [False] * 3, [d2[0], d2[1], d2[2][k0:k1]], bl, e, rank, world, arena=arena, sync=sync)
    loc = S.DeviceArray.from_numpy(np.asfortranarray(E[:, :, k0:k1, :]))
    if arena is not None:
        A, B, Cc = (arena.array(nl + (3,)) for _ in range(3))
        A.t.copy_(loc.t)
    else:
        A, B, Cc = loc, S.DeviceArray.zeros(nl + (3,)), S.DeviceArray.zeros(nl + (3,))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    for _ in range(2):  # twice: the second pass overwrites planes the neighbours read in the first (WAR ordering)
        c1.apply(B, A, "=")
        c2.apply(Cc, B, "=")
    torch.cuda.synchronize()
    err, nd = ulp_report(Cc.numpy(), ref[:, :, k0:k1, :])
    t = torch.tensor([err], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
        c1 = c2 = None
        arena.close()
    return {"grid": f"{N[0]}x{N[1]}x{N[2]} all-Bloch with phases, z ring over {world} rank(s)", "max_ulp": float(t.item()),
            "ok": bool(float(t.item()) <= 4.0)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--graph", default="auto", choices=["auto", "on", "off"],
                    help="replay the step from a CUDA graph (auto: for N > 1 with in-kernel synchronisation)")
    ap.add_argument("--halo", default="peer", choices=["peer", "nccl"])
    ap.add_argument("--sync", default="flags", choices=["flags", "barrier"],
                    help="peer-memory halos: neighbour flags inside the stencil kernel, or a barrier kernel before every apply")
    ap.add_argument("--slab", default="256,256,256", help="per-GPU slab Nx,Ny,Nz (default: BASELINE configs[1])")
    ap.add_argument("--no-cfg5", action="store_true", help="skip the BASELINE configs[4] leg (1024x1024x128 per GPU)")
    ap.add_argument("--no-parity", action="store_true", help="side experiments only: skip the oracle comparison")
    ap.add_argument("--no-assembly", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="side experiments only: skip the host-array leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-chunks", type=int, default=16, help="z-chunks of the host-array pipeline")
    ap.add_argument("--no-extras", action="store_true", help="skip the BASELINE configs[2] (2-D operators) and other side measurements")
    ap.add_argument("--serial-e2e", action="store_true", help="N > 1: H2D, step, D2H without the chunked pipeline")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    global NLOCAL
    NLOCAL = tuple(int(v) for v in args.slab.split(","))

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    # keep stdout for the ONE JSON line: whatever native libraries print while the job sets up (NCCL prints its
    # version banner to stdout on some boxes) is sent to stderr; file descriptor 1 is restored for the result line
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    import sgc_b200 as S

    assert torch.cuda.is_available(), "bench.py needs a GPU: the product path has no CPU fallback"
    torch.cuda.set_device(local_rank)
    from sgc_b200.hostpipe import bind_to_gpu_numa_node
    numa_cpus = bind_to_gpu_numa_node(local_rank) if world > 1 else None  # pinned buffers on the GPU's own node
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime
        # (a hung collective should end the run in minutes, not after NCCL's default 10)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(seconds=240))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    leg = Leg(S, torch, dist, NLOCAL, rank, world, args)
    M = leg.M
    E, H, E2, c_dual, c_prim, arena = leg.E, leg.H, leg.E2, leg.c_dual, leg.c_prim, leg.arena

    # ---- parity first: the timed configuration against the oracle, on every rank ---------------------------------
    parity = None
    if not args.no_parity:
        parity = leg.parity()
        parity["bloch_wrap_case"] = small_bloch_parity(S, torch, dist, rank, world)
        parity["ok"] = bool(parity["ok"] and parity["bloch_wrap_case"]["ok"])

    use_graph = args.graph == "on" or (args.graph == "auto" and world > 1 and arena is not None and leg.sync == "flags")
    tm = leg.time(args.steps, args.warmup, use_graph, local_rank)
    if leg.arena is not None and leg.arena.timeouts() != 0:  # a device-side wait on a neighbour gave up: the numbers mean nothing
        print(f"rank {rank}: {leg.arena.timeouts()} neighbour waits timed out inside the synchronised applies", file=sys.stderr)
        sys.exit(3)
    ms_step, value = tm["ms_step"], tm["value"]

    # ---- end-to-end: pinned host arrays → H2D → two curls → D2H, every step ----------------------
    nbytes = 3 * M * 16
    e2e_val, e2e_s, e2e_steps, e2e_mode = None, 0.0, 1, "skipped (--no-e2e)"
    if not args.no_e2e:
        c_dual.sync = c_prim.sync = "barrier"  # host copies touch the fields between applies: full barriers
        E_host = torch.empty(3 * M, dtype=torch.complex128).pin_memory()
        E_host.copy_(E.t)
        Out_host = torch.empty(3 * M, dtype=torch.complex128).pin_memory()
        from sgc_b200._lib import lib, check
        import ctypes as C
        st = torch.cuda.current_stream()

        def e2e_step_serial():
            check(lib.sgc_memcpy_h2d(C.c_void_p(E.ptr), C.c_void_p(E_host.data_ptr()), nbytes, C.c_void_p(st.cuda_stream)))
            leg.step()
            check(lib.sgc_memcpy_d2h(C.c_void_p(Out_host.data_ptr()), C.c_void_p(E2.ptr), nbytes, C.c_void_p(st.cuda_stream)))
            check(lib.sgc_stream_sync(C.c_void_p(st.cuda_stream)))

        leg.resync()
        e2e_step, e2e_mode, e2e_first_ms = e2e_step_serial, "serial: H2D, 2 curls, D2H", None
        if world == 1:
            # ONE library call per step (sgc_op_apply_curlcurl_host: the z-chunked full-duplex pipeline — H2D | 2 fused
            # curls per chunk | D2H on three streams — lives inside the library).  Timed with host arrays in pinned
            # memory from the library's allocator (sgc_malloc_host), and, as a second figure, with ordinary pageable
            # numpy arrays, which the library page-locks in place on first sight (cached).
            import ctypes as C2
            e2e_step_serial()
            ref_np = Out_host.numpy().copy()
            E_np = E_host.numpy().copy()          # plain numpy arrays: NOT pinned by the caller
            E_host = Out_host = None
            op1, op2 = c_dual.backend.op, c_prim.backend.op
            Out_np = np.zeros_like(E_np)
            t0 = time.perf_counter()
            op2.apply_after_host(op1, Out_np, E_np, "=", nchunks=args.e2e_chunks)
            e2e_first_ms = (time.perf_counter() - t0) * 1e3
            if not np.array_equal(Out_np, ref_np):
                raise RuntimeError("sgc_op_apply_curlcurl_host disagrees with the serial host path")
            for _ in range(2):
                op2.apply_after_host(op1, Out_np, E_np, "=", nchunks=args.e2e_chunks)
            t0 = time.perf_counter()
            for _ in range(5):
                op2.apply_after_host(op1, Out_np, E_np, "=", nchunks=args.e2e_chunks)
            e2e_pageable_ms = (time.perf_counter() - t0) * 1e3 / 5
            lib.sgc_host_unregister(None)
            Out_np = None

            def pinned(n):  # numpy view of library-pinned host memory
                p = C2.c_void_p()
                check(lib.sgc_malloc_host(C2.byref(p), n * 16))
                return np.ctypeslib.as_array(C2.cast(p, C2.POINTER(C2.c_double)), shape=(2 * n,)).view(np.complex128), p
            E_pin, pE = pinned(3 * M)
            Out_pin, pO = pinned(3 * M)
            E_pin[:] = E_np
            Out_pin[:] = 0
            E_np = None
            op2.apply_after_host(op1, Out_pin, E_pin, "=", nchunks=args.e2e_chunks)
            if not np.array_equal(Out_pin, ref_np):
                raise RuntimeError("sgc_op_apply_curlcurl_host (pinned arrays) disagrees with the serial host path")
            e2e_step = lambda: op2.apply_after_host(op1, Out_pin, E_pin, "=", nchunks=args.e2e_chunks)
            e2e_mode = (f"one C-ABI call per step (sgc_op_apply_curlcurl_host), host arrays in sgc_malloc_host memory; z-chunked "
                        f"({args.e2e_chunks}) full-duplex pipeline inside the library; with ordinary pageable numpy arrays "
                        f"(page-locked in place by the library, first call {e2e_first_ms:.0f} ms) {e2e_pageable_ms:.2f} ms per step")
        elif arena is not None and not args.serial_e2e:  # the same pipeline per rank; interface planes after a peer barrier
            from sgc_b200.hostpipe import HostSlabCurlCurl
            pipe = HostSlabCurlCurl(c_dual, c_prim, E, H, E2, nchunks=args.e2e_chunks)
            e2e_step_serial()
            ref_out = Out_host.clone()
            Out_host.zero_()
            pipe.run(E_host, Out_host)
            if not torch.equal(ref_out, Out_host):
                raise RuntimeError("pipelined slab host path disagrees with the serial host path")
            e2e_step = lambda: pipe.run(E_host, Out_host)
            e2e_mode = (f"per-rank z-chunked ({args.e2e_chunks}) full-duplex pipeline, interface planes after a peer barrier; "
                        f"process bound to the GPU's NUMA node ({'yes' if numa_cpus else 'topology not exposed'})")

        e2e_steps = max(3, min(args.steps, 10))
        for _ in range(2):
            e2e_step()
        leg.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        leg.barrier()
        e2e_s = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        e2e_val = 2 * M * world * e2e_steps / e2e_s / 1e9
        E_host = Out_host = None
        if world == 1:
            E_pin = Out_pin = e2e_step = None
            lib.sgc_free_host(pE); lib.sgc_free_host(pO)

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))

    other = None
    if world == 1 and not args.no_extras and rank == 0:
        other = bench_other_configs(S, torch, peak, c_dual.backend.op, c_prim.backend.op, E, E2, ms_step)
    halo_txt, main_sync = leg.halo, leg.sync
    E = H = E2 = c_dual = c_prim = arena = None
    leg.close()

    # ---- BASELINE configs[4]: the 1024^3 grid's per-GPU slab (1024x1024x128), same step, no host leg --------------
    cfg5 = None
    if not args.no_cfg5 and NLOCAL == (256, 256, 256):
        try:
            leg5 = Leg(S, torch, dist, CFG5, rank, world, args, seed_shift=100)
            nz5 = CFG5[2]
            p5 = None if args.no_parity else leg5.parity([(0, 3), (nz5 // 2 - 2, nz5 // 2 + 2), (nz5 - 3, nz5)])
            use_graph5 = args.graph == "on" or (args.graph == "auto" and world > 1 and leg5.arena is not None and leg5.sync == "flags")
            t5 = leg5.time(max(5, min(args.steps, 20)), 3, use_graph5, local_rank)
            k5 = t5["ms_step"] / 2
            a5 = ALG_BYTES_PER_CELL * leg5.M / (k5 * 1e-3) / 1e9
            cfg5 = {"workload": f"BASELINE configs[4]: {CFG5[0]}x{CFG5[1]}x{CFG5[2] * world} ComplexF64 curl-of-curl, {CFG5[0]}x{CFG5[1]}x{CFG5[2]} slab per GPU",
                    "value": t5["value"], "unit": "Gcell/s", "ms_per_step": t5["ms_step"], "per_rank_ms": t5["per_rank_ms"],
                    "per_gpu_gbs": a5, "per_gpu_frac": a5 / peak, "cuda_graph": t5["graph"], "parity": p5}
            if p5 is not None and parity is not None:
                parity["ok"] = bool(parity["ok"] and p5["ok"])
            leg5.close()
        except Exception as ex:
            cfg5 = {"error": f"{type(ex).__name__}: {str(ex)[:200]}"}

    # BASELINE configs[3] at N > 1: the 512^3 create_curl partitioned into per-GPU CSC blocks
    asm_part = None
    if world > 1 and not args.no_assembly:
        asm_part = bench_assembly_partitioned(S, torch, dist, rank, world)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        if parity is not None and not parity["ok"]:
            sys.exit(1)
        return

    kernel_ms = ms_step / 2  # the step is exactly two launches of the fused curl kernel (+ halo planes for N>1)
    achieved = ALG_BYTES_PER_CELL * M / (kernel_ms * 1e-3) / 1e9
    out = {
        "metric": "curl_apply_gcell_per_s", "value": value, "unit": "Gcell/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "c128", "data": "synthetic", "config": workload_config(world),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "kernel": "k_curl3p<cplx,cplx,ADD=0,32x4 tile,6-stage cp.async ring>",
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "fallback 6650 (of fallback)",
                     "frac_of_nominal_8TBs": achieved / 8000.0,
                     "algorithmic_bytes_per_launch": ALG_BYTES_PER_CELL * M},
        "clocks": tm["clocks"],
        "e2e": {"value": e2e_val, "unit": "Gcell/s", "h2d_bytes_per_step": nbytes if e2e_val else 0, "d2h_bytes_per_step": nbytes if e2e_val else 0,
                "ms_per_step": e2e_s / e2e_steps * 1e3, "steps": e2e_steps, "mode": e2e_mode},
        "gpu_launches": tm["launches"],
        "cuda_graph": tm["graph"],
        "per_rank_ms": tm["per_rank_ms"],
        "halo_transport": halo_txt,
        "host_enqueue_ms_per_step": tm["host_enqueue_ms"],
    }
    tr = os.path.join(ROOT, "profiles", "curl3_traffic.json")
    if world == 1 and os.path.exists(tr) and NLOCAL == (256, 256, 256):  # the ncu capture is of the 1-GPU 256^3 launch
        try:
            out["roofline"]["traffic"] = json.load(open(tr)).get("dram_bytes_per_launch")
        except Exception:
            pass
    # (order: the long side measurements first, the headline's companions last — a tail of the line keeps them)
    if other is not None:
        out["other_configs"] = other
    if world == 1 and not args.no_cpu_baseline:
        dprim, ddual = build_inputs(NLOCAL[2])
        rate, threads, sample = cpu_reference_rate(dprim, ddual, NLOCAL[2], args.cpu_seconds)
        out["cpu_baseline"] = {"value": rate, "unit": "Gcell/s", "cores": threads, "kind": "port", "sample": sample}
    if world == 1 and not args.no_assembly:
        out["assembly_c128"] = bench_assembly(S, torch, peak, cplx=True)
        out["assembly"] = bench_assembly(S, torch, peak)
        if not args.no_cpu_baseline and "error" not in out["assembly"]:
            # the reference's assembly algorithm (create_∂info temporaries → COO → serial counting-sort `sparse` →
            # `dropzeros!`; C port, single thread like the stdlib) on a 128³ sample of the same matrix family
            try:
                from oracle import c_oracle as CO
                n = 128
                dt, nnz_cpu = CO.time_create_curl([True] * 3, [np.ones(n)] * 3, [True] * 3, [1.0] * 3)
                out["assembly"]["cpu_baseline"] = {"value": nnz_cpu / dt, "unit": "nnz/s", "cores": 1, "kind": "port",
                                                   "sample": f"create_curl {n}^3 Float64 all-Bloch ({nnz_cpu} nnz) in {dt:.2f} s"}
            except Exception as ex:
                out["assembly"]["cpu_baseline"] = {"error": str(ex)[:120]}
    if world > 1 and not args.no_assembly:
        out["assembly"] = asm_part
    if cfg5 is not None:
        out["cfg5"] = cfg5
    if other is not None and "curlcurl_fused" in other:  # the step as one launch: next to the headline it accelerates
        out["curlcurl_fused"] = other.pop("curlcurl_fused")
        if parity is not None and not out["curlcurl_fused"].get("bit_identical_to_two_launches", False):
            parity["ok"] = False
    out["parity"] = parity
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    print(json.dumps(out), flush=True)
    if world > 1:
        os.dup2(2, 1)
        dist.barrier()
        dist.destroy_process_group()
    if parity is not None and not parity["ok"]:
        sys.exit(1)


def bench_assembly(S, torch, peak, cplx=False):
    """Second half of the metric: create_curl CSC assembly in nnz/s (BASELINE.json configs[3]:
    512³, Float64, uniform, all-Bloch, nnz = 12·512³ = 1 610 612 736), through the C ABI with
    caller-owned output arrays (what the Julia shim does): count pass + scan + nnz read-back, then
    the single fill pass that writes colptr, rowval and nzval."""
    import ctypes as C
    from sgc_b200._lib import lib, check
    from sgc_b200.device import ints, int64s, scalars2, stream_handle
    n = 512
    N = (n, n, n)
    M = n ** 3
    res = {}
    try:
        h = C.c_void_p()
        e = list(np.exp(-2j * np.pi * np.array([0.1, 0.2, 0.3]))) if cplx else [1.0] * 3
        check(lib.sgc_asm_create_curl(C.byref(h), 3, int64s(N), ints([1, 1, 1]), None, 0, ints([1, 1, 1]),
                                      scalars2(e), int(cplx), ints([1, 2, 3]), 3, ints([1, 2, 3]), 3, ints([1, 2, 3]), 1))
        st = stream_handle()
        nnz = C.c_int64()
        check(lib.sgc_asm_count(h, C.byref(nnz), st))
        colptr = torch.empty(3 * M + 1, dtype=torch.int64, device="cuda")
        rowval = torch.empty(nnz.value, dtype=torch.int64, device="cuda")
        nzval = torch.empty(nnz.value, dtype=torch.complex128 if cplx else torch.float64, device="cuda")

        def once():
            check(lib.sgc_asm_set_row_range(h, 0, 3 * M))  # drops the cached count: time the whole assembly
            check(lib.sgc_asm_count(h, C.byref(nnz), st))
            check(lib.sgc_asm_fill_all(h, C.c_void_p(colptr.data_ptr()), 1, C.c_void_p(rowval.data_ptr()),
                                       C.c_void_p(nzval.data_ptr()), st))
        for _ in range(2):
            once()
        torch.cuda.synchronize()
        reps = 5
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(reps):
            once()
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / reps
        ok = bool((colptr[1:] - colptr[:-1] == 4).all()) and int(colptr[-1]) == nnz.value + 1
        lib.sgc_asm_destroy(h)
        alg = nnz.value * (24 if cplx else 16) + (3 * M + 1) * 8
        res = {"workload": f"create_curl 512^3 {'ComplexF64 (Bloch phases)' if cplx else 'Float64'} uniform all-Bloch, order_cmpfirst=true; "
                           f"colptr+rowval+nzval on device ({alg / 1e9:.1f} GB)",
               "nnz": nnz.value, "ms": ms, "nnz_per_s": nnz.value / (ms * 1e-3), "algorithmic_bytes": alg,
               "achieved_gbs": alg / (ms * 1e-3) / 1e9, "frac": alg / (ms * 1e-3) / 1e9 / peak, "structure_ok": ok,
               "includes": "nnz and per-CTA offsets in closed form (no count pass, no scan) + single fill pass writing colptr, rowval, nzval"}
        del colptr, rowval, nzval
        torch.cuda.empty_cache()
    except Exception as ex:  # out of memory on a shared box etc.
        res = {"error": str(ex)[:200]}
    return res


def bench_assembly_partitioned(S, torch, dist, rank, world):
    """512³ Float64 all-Bloch create_curl partitioned over the ranks, both ways: (rows) rank r assembles the CSC
    block of its contiguous row range (own colptr of length n+1; the blocks stack to the global matrix); (cols)
    rank r assembles its block of columns — the global rowval / nzval are the concatenation of the blocks' arrays
    and every rank does exactly 1/N of the single-GPU work.  No exchange; the time is the max over ranks, the rate
    the whole matrix's nnz over that time."""
    res = _bench_assembly_part(S, torch, dist, rank, world, by_cols=True)   # the scalable form first
    res["row_partition"] = _bench_assembly_part(S, torch, dist, rank, world, by_cols=False)
    return res


def _bench_assembly_part(S, torch, dist, rank, world, by_cols):
    import ctypes as C
    from sgc_b200._lib import lib, check
    from sgc_b200.device import ints, int64s, scalars2, stream_handle
    from sgc_b200.slab import row_range_for_rank
    n = 512
    N, M = (n, n, n), n ** 3
    try:
        h = C.c_void_p()
        check(lib.sgc_asm_create_curl(C.byref(h), 3, int64s(N), ints([1, 1, 1]), None, 0, ints([1, 1, 1]),
                                      scalars2([1.0] * 3), 0, ints([1, 2, 3]), 3, ints([1, 2, 3]), 3, ints([1, 2, 3]), 1))
        r0, r1 = row_range_for_rank(3 * M, rank, world)
        st = stream_handle()
        nnz = C.c_int64()
        ncols = 3 * M
        if by_cols:
            c0, c1 = C.c_int64(), C.c_int64()
            check(lib.sgc_asm_col_partition(h, rank, world, C.byref(c0), C.byref(c1)))
            check(lib.sgc_asm_set_col_range(h, c0.value, c1.value))
            ncols = c1.value - c0.value
        else:
            check(lib.sgc_asm_set_row_range(h, r0, r1))
        check(lib.sgc_asm_count(h, C.byref(nnz), st))
        colptr = torch.empty(ncols + 1, dtype=torch.int64, device="cuda")
        rowval = torch.empty(nnz.value, dtype=torch.int64, device="cuda")
        nzval = torch.empty(nnz.value, dtype=torch.float64, device="cuda")

        def once():
            if by_cols:
                check(lib.sgc_asm_set_col_range(h, c0.value, c1.value))  # drops the cached count: time the whole assembly
            else:
                check(lib.sgc_asm_set_row_range(h, r0, r1))
            check(lib.sgc_asm_count(h, C.byref(nnz), st))
            check(lib.sgc_asm_fill_all(h, C.c_void_p(colptr.data_ptr()), 1, C.c_void_p(rowval.data_ptr()),
                                       C.c_void_p(nzval.data_ptr()), st))
        for _ in range(2):
            once()
        torch.cuda.synchronize()
        dist.barrier()
        reps = 5
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(reps):
            once()
        ev1.record()
        torch.cuda.synchronize()
        t = torch.tensor([ev0.elapsed_time(ev1) / reps, float(nnz.value)], device="cuda", dtype=torch.float64)
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        lib.sgc_asm_destroy(h)
        ms, total = float(tmax[0]), int(t[1])
        if by_cols:
            return {"workload": f"create_curl 512^3 Float64 all-Bloch partitioned into {world} per-GPU CSC blocks of COLUMNS, no exchange "
                                f"(global rowval / nzval = concatenation of the blocks' arrays, matrix/differential/curl.jl:103)",
                    "nnz": total, "ms": ms, "nnz_per_s": total / (ms * 1e-3), "nnz_ok": total == 12 * M,
                    "includes": "per rank: closed-form nnz and offsets + single fill pass over its columns (max over ranks)"}
        return {"workload": f"create_curl 512^3 Float64 all-Bloch, row-partitioned into {world} per-GPU CSC blocks (no exchange)",
                "nnz": total, "ms": ms, "nnz_per_s": total / (ms * 1e-3), "nnz_ok": total == 12 * M,
                "includes": "per rank: count pass + device scan + nnz read-back + single fill pass (max over ranks)"}
    except Exception as ex:
        return {"error": str(ex)[:200]}


def _time(torch, fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def bench_other_configs(S, torch, peak, c_dual, c_prim, E, E2, ms_two_curls):
    """Side measurements (not the headline): the fused curl-of-curl on the headline grid, 3-D divergence / gradient /
    mean, BASELINE.json configs[2] — 2-D 8192² Float64 ∂ / m̂ / TE / TM operators with PEC/PMC (symmetric)
    boundaries — and the other create_* assemblies.  `frac` = ALGORITHMIC bytes / time / HBM peak."""
    res = {}
    r4 = lambda x: round(float(x), 4)
    try:
        M = NLOCAL[0] * NLOCAL[1] * NLOCAL[2]
        # the same step (BASELINE configs[1]: E -> H -> E') as ONE kernel: H lives in shared memory only, so a cell costs
        # 96 B instead of 192.  Checked here against the two-launch result of the same operators (which the parity leg
        # compared with the C restatement of the reference): the fused kernel must reproduce it bit for bit.
        Href, Gref = S.DeviceArray.zeros(NLOCAL + (3,)), S.DeviceArray.zeros(NLOCAL + (3,))
        c_dual.apply(Href, E, "="); c_prim.apply(Gref, Href, "=")
        c_prim.apply_after(c_dual, E2, E, "=")
        same = bool(torch.equal(torch.view_as_real(E2.t), torch.view_as_real(Gref.t)))
        del Href, Gref
        ms = _time(torch, lambda: c_prim.apply_after(c_dual, E2, E, "="), iters=50, warm=5)
        res["curlcurl_fused"] = {"what": "E' = C_prim(C_dual E) in ONE launch of k_cc6 (TMA-fed, two z-planes per barrier interval), H kept in shared memory",
                                 "ms": r4(ms), "ms_two_launches": r4(ms_two_curls), "speedup": r4(ms_two_curls / ms),
                                 "gcell_updates_per_s": r4(2 * M / ms / 1e6), "achieved_gbs_96B_per_cell": r4(96 * M / ms / 1e6),
                                 "frac_of_96B_per_cell": r4(96 * M / ms / 1e6 / peak), "bit_identical_to_two_launches": same}
        N3 = NLOCAL
        rng = np.random.default_rng(3)
        d3 = [rng.uniform(0.5, 1.5, n) + 1j * rng.uniform(-.2, .2, n) for n in N3]
        g = S.DeviceArray.zeros(N3)
        dv = S.Operator.divg(S.SGC_C128, N3, [True] * 3, d3, [False] * 3, [1.0] * 3)
        gr = S.Operator.grad(S.SGC_C128, N3, [False] * 3, d3, [False] * 3, [1.0] * 3)
        mn = S.Operator.mean(S.SGC_C128, N3, [True] * 3, None, None, [False] * 3, [1.0] * 3)
        mw = S.Operator.mean(S.SGC_C128, N3, [False] * 3, d3, d3, [False] * 3, [1.0] * 3)
        three_d = {}
        for name, fn, B in (("divg", lambda: dv.apply(g, E, "="), 64), ("grad", lambda: gr.apply(E2, g, "="), 64),
                            ("mean_arith", lambda: mn.apply(E2, E, "="), 96), ("mean_weighted", lambda: mw.apply(E2, E, "="), 96)):
            ms = _time(torch, fn)
            three_d[name] = {"ms": r4(ms), "frac": r4(B * M / ms / 1e6 / peak)}
        three_d["launches"] = {"divg": dv.launches("="), "grad": gr.launches("="), "mean": mn.launches("=")}
        res["ops_3d_c128"] = three_d
        del g
        n = 8192
        N2, M2 = (n, n), n * n
        d = rng.uniform(0.5, 1.5, n)
        F = S.DeviceArray(torch.rand(M2, device="cuda", dtype=torch.float64), N2)
        G = S.DeviceArray.zeros(N2, np.float64)
        F2 = S.DeviceArray(torch.rand(2 * M2, device="cuda", dtype=torch.float64), N2 + (2,))
        G1 = S.DeviceArray.zeros(N2 + (1,), np.float64)
        ops = []
        for nw in (1, 2):
            ops.append((f"partial_nw{nw}_fwd_set", S.Operator.partial(S.SGC_F64, N2, nw, True, d, False), G, F, "=", 16))
            ops.append((f"partial_nw{nw}_bwd_add", S.Operator.partial(S.SGC_F64, N2, nw, False, d, False), G, F, "+=", 24))
            ops.append((f"mhat_arith_nw{nw}", S.Operator.mhat(S.SGC_F64, N2, nw, True, None, None, False), G, F, "=", 16))
            ops.append((f"mhat_weighted_nw{nw}", S.Operator.mhat(S.SGC_F64, N2, nw, False, d, d, False), G, F, "=", 16))
        ops.append(("TE_curl_1x2", S.Operator.curl(S.SGC_F64, N2, [True, True], [d, d], [False, False], [1.0, 1.0], cmp_shp=[1, 2],
                                                   cmp_out=[3], cmp_in=[1, 2]), G1, F2, "=", 24))
        ops.append(("TM_curl_2x1", S.Operator.curl(S.SGC_F64, N2, [False, False], [d, d], [False, False], [1.0, 1.0], cmp_shp=[1, 2],
                                                   cmp_out=[1, 2], cmp_in=[3]), F2, G1, "=", 24))
        two_d = {}
        for name, o, Gd, Fd, mode, B in ops:
            ms = _time(torch, lambda: o.apply(Gd, Fd, mode))
            two_d[name] = {"ms": r4(ms), "frac": r4(B * M2 / ms / 1e6 / peak)}
        res["ops_2d_8192_f64_pec_pmc"] = two_d
        del F, G, F2, G1, ops
        torch.cuda.empty_cache()
        # the other assemblies (Float64, 384³ grid, all-Bloch with a symmetric z for the drop path), whole create_* call
        na = 384
        Na = (na, na, na)
        da = [rng.uniform(0.5, 1.5, na) for _ in range(3)]
        asm = {}
        cases = (("create_partial", lambda: S.create_partial(3, True, Na, da[2], True)),
                 ("create_mhat", lambda: S.create_mhat(2, False, Na, da[1], da[1], False)),
                 ("create_mean", lambda: S.create_mean([True, False, True], da, da, [True, True, False])),
                 ("create_divg", lambda: S.create_divg([True] * 3, da, [True, True, False])),
                 ("create_grad", lambda: S.create_grad([False] * 3, da, [True, True, False])))
        import ctypes as C
        from sgc_b200 import matrix as MX
        from sgc_b200._lib import lib as _lib_, check as _check_
        from sgc_b200.device import stream_handle as _sh_
        for name, fn in cases:
            A = fn()
            nnz, ncol = A.nnz, A.n
            ms = _time(torch, fn, iters=5, warm=1)  # the whole reference-shaped call: tables, handle, output allocation, kernel
            alg = nnz * 16 + (ncol + 1) * 8
            # the device work alone: one fill pass (colptr + rowval + nzval) on a held handle into the same arrays
            finish = MX._finish
            try:
                MX._finish = lambda a, *args, **kw: a
                hnd = fn()
            finally:
                MX._finish = finish
            st, nn = _sh_(None), C.c_int64()
            _check_(_lib_.sgc_asm_count(hnd.h, C.byref(nn), st))
            fill = lambda: _check_(_lib_.sgc_asm_fill_all(hnd.h, C.c_void_p(A.colptr.data_ptr()), 1, C.c_void_p(A.rowval.data_ptr()),
                                                          C.c_void_p(A.nzval.data_ptr()), st))
            ms_fill = _time(torch, fill, iters=10, warm=2)
            asm[name] = {"ms_whole_call": r4(ms), "ms_fill": r4(ms_fill), "gnnz_per_s": r4(nnz / ms_fill / 1e6), "frac": r4(alg / ms_fill / 1e6 / peak),
                         "frac_whole_call": r4(alg / ms / 1e6 / peak)}
            del A, hnd
        res["assembly_others_384_f64"] = asm
        # call latency of the reference-shaped one-shot entry point (handle cached inside the library) against a held
        # handle, on a 64³ grid where the kernel itself takes ~10 µs (BASELINE configs[0] size)
        import ctypes as C
        import time as _t
        from sgc_b200._lib import lib, check
        from sgc_b200.device import ints, int64s, scalars2, scalar2, stream_handle, coef_vectors
        n6 = (64, 64, 64)
        d6 = [rng.uniform(0.5, 1.5, 64) + 1j * rng.uniform(-.2, .2, 64) for _ in range(3)]
        F6 = S.DeviceArray(torch.view_as_complex(torch.rand(3 * 64 ** 3, 2, device="cuda", dtype=torch.float64)), n6 + (3,))
        G6 = S.DeviceArray.zeros(n6 + (3,))
        arr, keep, cc = coef_vectors(d6, n6)
        e6 = list(np.exp(-2j * np.pi * np.array([0.1, 0.2, 0.3])))
        args = (C.c_void_p(G6.ptr), C.c_void_p(F6.ptr), 0, S.SGC_C128, 3, int64s(n6), ints([1, 1, 1]), arr, cc, ints([1, 1, 1]), scalars2(e6),
                ints([1, 2, 3]), 3, ints([1, 2, 3]), 3, ints([1, 2, 3]), scalar2(1.0), stream_handle())
        lib.sgc_oneshot_cache_clear()
        torch.cuda.synchronize()
        t0 = _t.perf_counter(); check(lib.sgc_apply_curl(*args)); torch.cuda.synchronize(); first_us = (_t.perf_counter() - t0) * 1e6
        reps = 300
        t0 = _t.perf_counter()
        for _ in range(reps):
            check(lib.sgc_apply_curl(*args))
        torch.cuda.synchronize()
        cached_us = (_t.perf_counter() - t0) * 1e6 / reps
        o6 = S.Operator.curl(S.SGC_C128, n6, [True] * 3, d6, [True] * 3, e6)
        o6.apply(G6, F6, "=")
        torch.cuda.synchronize()
        t0 = _t.perf_counter()
        for _ in range(reps):
            o6.apply(G6, F6, "=")
        torch.cuda.synchronize()
        handle_us = (_t.perf_counter() - t0) * 1e6 / reps
        res["call_latency_us_64cubed_curl"] = {"one_shot_first_call": r4(first_us), "one_shot_cached": r4(cached_us), "held_handle": r4(handle_us)}
    except Exception as ex:
        res["error"] = str(ex)[:200]
    return res


if __name__ == "__main__":
    main()
