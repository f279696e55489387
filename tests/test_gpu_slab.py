"""Multi-GPU z-slab curl with NCCL halo exchange (needs ≥ 2 GPUs: `gpurun --gpus 2`).
Each rank's slab result must equal the corresponding planes of the single-GPU result
bit-for-bit, for Bloch and symmetric z boundaries, forward and backward ∂z."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, results, use_arena=False):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import sgc_b200 as S
        from sgc_b200.slab import SlabCurl, partition_planes
        rng = np.random.default_rng(5)
        N = (45, 23, 4 * world + 3)
        F = np.asfortranarray(rng.uniform(-1, 1, N + (3,)) + 1j * rng.uniform(-1, 1, N + (3,)))
        G0 = np.asfortranarray(rng.uniform(-1, 1, N + (3,)) + 1j * rng.uniform(-1, 1, N + (3,)))
        dl = [rng.uniform(0.5, 1.5, n) + 1j * rng.uniform(-0.2, 0.2, n) for n in N]
        e = [np.exp(-0.3j), np.exp(0.5j), np.exp(-0.9j)]
        k0, k1 = partition_planes(N[2], world)[rank]
        ok = True
        arena = None
        if use_arena:
            from sgc_b200.slab import PeerArena
            N = (45, 23, 5 * world)  # the peer path needs slabs of equal size
            F = np.asfortranarray(rng.uniform(-1, 1, N + (3,)) + 1j * rng.uniform(-1, 1, N + (3,)))
            G0 = np.asfortranarray(rng.uniform(-1, 1, N + (3,)) + 1j * rng.uniform(-1, 1, N + (3,)))
            dl = [rng.uniform(0.5, 1.5, n) + 1j * rng.uniform(-0.2, 0.2, n) for n in N]
            k0, k1 = partition_planes(N[2], world)[rank]
            arena = PeerArena(2 * 3 * N[0] * N[1] * (k1 - k0) * 16 + 1024)
            Ga, Fa = arena.array((N[0], N[1], k1 - k0, 3)), arena.array((N[0], N[1], k1 - k0, 3))
        for fz in (True, False):
            for bz in (True, False):
                for op in ("=", "+="):
                    isfwd, isbloch = [False, True, fz], [True, False, bz]
                    whole = S.Operator.curl(S.SGC_C128, N, isfwd, dl, isbloch, e)
                    Gw = S.DeviceArray.from_numpy(G0)
                    whole.apply(Gw, S.DeviceArray.from_numpy(F), op)
                    ref = Gw.numpy()[:, :, k0:k1, :]
                    Nl = (N[0], N[1], k1 - k0)
                    sc = SlabCurl(S.SGC_C128, Nl, isfwd, [dl[0], dl[1], dl[2][k0:k1]], isbloch, e, rank, world, arena=arena)
                    Gl = S.DeviceArray.from_numpy(np.asfortranarray(G0[:, :, k0:k1, :]))
                    Fl = S.DeviceArray.from_numpy(np.asfortranarray(F[:, :, k0:k1, :]))
                    if arena is not None:
                        arena.barrier()  # nobody may still be reading Fa from the previous case
                        Ga.t.copy_(Gl.t); Fa.t.copy_(Fl.t)
                        Gl, Fl = Ga, Fa
                    for _ in range(2 if op == "=" else 1):  # repeated application reuses the ghost buffers
                        sc.apply(Gl, Fl, op)
                    torch.cuda.synchronize()
                    ok = ok and np.array_equal(Gl.numpy(), ref)
        results[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("use_arena", [False, True], ids=["nccl-messages", "peer-memory-loads"])
def test_slab_curl_matches_single_gpu(use_arena):
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), results, use_arena), nprocs=world, join=True)
    assert all(results.get(r) is True for r in range(world)), dict(results)


def _worker_flags(rank, world, port, results):
    """A chain of alternating forward / backward curls (E → H → E → …, ping-pong between two arena fields) with the
    neighbour synchronisation inside the kernel (`sync="flags"`): every rank's slab of the final field must equal
    the single-GPU chain bit for bit — which fails if a halo is read before it is complete (RAW) or a plane is
    overwritten while a neighbour still reads it (WAR)."""
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import sgc_b200 as S
        from sgc_b200.slab import SlabCurl, PeerArena, partition_planes
        rng = np.random.default_rng(6)
        N = (64, 40, 6 * world)
        E0 = np.asfortranarray(rng.uniform(-1, 1, N + (3,)) + 1j * rng.uniform(-1, 1, N + (3,)))
        d1 = [rng.uniform(0.5, 1.5, n) + 1j * rng.uniform(-0.2, 0.2, n) for n in N]
        d2 = [rng.uniform(0.5, 1.5, n) + 1j * rng.uniform(-0.2, 0.2, n) for n in N]
        e = [np.exp(-0.3j), np.exp(0.5j), np.exp(-0.9j)]
        k0, k1 = partition_planes(N[2], world)[rank]
        Nl = (N[0], N[1], k1 - k0)
        ok = True
        arena = PeerArena(2 * 3 * N[0] * N[1] * (k1 - k0) * 16 + 1024)
        A, B = arena.array(Nl + (3,)), arena.array(Nl + (3,))
        iters = 6
        for bz in (True, False):
            isbloch = [True, False, bz]
            w1 = S.Operator.curl(S.SGC_C128, N, [True] * 3, d1, isbloch, e, alpha=0.25)
            w2 = S.Operator.curl(S.SGC_C128, N, [False] * 3, d2, isbloch, e, alpha=0.25)
            Ew, Hw = S.DeviceArray.from_numpy(E0), S.DeviceArray.zeros(N + (3,))
            for _ in range(iters):
                w1.apply(Hw, Ew, "=")
                w2.apply(Ew, Hw, "=")
            ref = Ew.numpy()[:, :, k0:k1, :]
            c1 = SlabCurl(S.SGC_C128, Nl, [True] * 3, [d1[0], d1[1], d1[2][k0:k1]], isbloch, e, rank, world, alpha=0.25,
                          arena=arena, sync="flags")
            c2 = SlabCurl(S.SGC_C128, Nl, [False] * 3, [d2[0], d2[1], d2[2][k0:k1]], isbloch, e, rank, world, alpha=0.25,
                          arena=arena, sync="flags")
            arena.barrier()  # nobody still reads A / B of the previous case
            A.t.copy_(S.DeviceArray.from_numpy(np.asfortranarray(E0[:, :, k0:k1, :])).t)
            B.t.zero_()
            arena.barrier()  # inputs complete everywhere before the flag-synchronised chain starts
            for _ in range(iters):
                c1.apply(B, A, "=")
                c2.apply(A, B, "=")
            torch.cuda.synchronize()
            ok = ok and np.array_equal(A.numpy(), ref)
        ok = ok and arena.timeouts() == 0  # no device-side wait on a neighbour gave up
        results[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_slab_flag_sync_chain_matches_single_gpu():
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker_flags, args=(world, _free_port(), results), nprocs=world, join=True)
    assert all(results.get(r) is True for r in range(world)), dict(results)


def _worker_curlcurl(rank, world, port, results):
    """Fused curl-of-curl across ranks (SlabCurlCurl: one kernel per rank, E planes −1 / Nz over peer memory) against
    two single-GPU curls, for both z-direction pairings, Bloch (ring) and symmetric z boundaries, `=` and `+=`."""
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import sgc_b200 as S
        from sgc_b200.slab import SlabCurlCurl, PeerArena, partition_planes
        rng = np.random.default_rng(7)
        N = (70, 37, 5 * world)
        F = np.asfortranarray(rng.uniform(-1, 1, N + (3,)) + 1j * rng.uniform(-1, 1, N + (3,)))
        G0 = np.asfortranarray(rng.uniform(-1, 1, N + (3,)) + 1j * rng.uniform(-1, 1, N + (3,)))
        d1 = [rng.uniform(0.5, 1.5, n) + 1j * rng.uniform(-0.2, 0.2, n) for n in N]
        d2 = [rng.uniform(0.5, 1.5, n) + 1j * rng.uniform(-0.2, 0.2, n) for n in N]
        e = [np.exp(-0.3j), np.exp(0.5j), np.exp(-0.9j)]
        k0, k1 = partition_planes(N[2], world)[rank]
        Nl = (N[0], N[1], k1 - k0)
        arena = PeerArena(2 * 3 * N[0] * N[1] * (k1 - k0) * 16 + 1024)
        Ga, Fa = arena.array(Nl + (3,)), arena.array(Nl + (3,))
        ok = True
        for fz1 in (True, False):
            for bz in (True, False):
                for op in ("=", "+="):
                    isfwd1, isfwd2, isbloch = [True, False, fz1], [False, True, not fz1], [True, False, bz]
                    w1 = S.Operator.curl(S.SGC_C128, N, isfwd1, d1, isbloch, e, alpha=0.5)
                    w2 = S.Operator.curl(S.SGC_C128, N, isfwd2, d2, isbloch, e)
                    H = S.DeviceArray.zeros(N + (3,))
                    Gw = S.DeviceArray.from_numpy(G0)
                    w1.apply(H, S.DeviceArray.from_numpy(F), "=")
                    w2.apply(Gw, H, op)
                    ref = Gw.numpy()[:, :, k0:k1, :]
                    cc = SlabCurlCurl(S.SGC_C128, Nl, isfwd1, [d1[0], d1[1], d1[2][k0:k1]], isfwd2, [d2[0], d2[1], d2[2][k0:k1]],
                                      isbloch, e, rank, world, arena, dz1_below=d1[2][(k0 - 1) % N[2]], dz1_above=d1[2][k1 % N[2]],
                                      alpha1=0.5)
                    arena.barrier()  # nobody may still be reading Fa from the previous case
                    Ga.t.copy_(S.DeviceArray.from_numpy(np.asfortranarray(G0[:, :, k0:k1, :])).t)
                    Fa.t.copy_(S.DeviceArray.from_numpy(np.asfortranarray(F[:, :, k0:k1, :])).t)
                    cc.apply(Ga, Fa, op)
                    torch.cuda.synchronize()
                    ok = ok and np.array_equal(Ga.numpy(), ref)
        results[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_slab_fused_curlcurl_matches_single_gpu():
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker_curlcurl, args=(world, _free_port(), results), nprocs=world, join=True)
    assert all(results.get(r) is True for r in range(world)), dict(results)


def _weighted_mz(S, n, fz, wl, dl, bz, ez, w_global, k0, k1, is_slab):
    o = S.Operator.mhat(S.SGC_C128, n, 3, fz, wl, dl, bz, ez)
    if is_slab:  # the neighbours' ∆w just below / above the slab (wrapping around the global grid)
        o.set_slab_weights(w_global[(k0 - 1) % len(w_global)], w_global[k1 % len(w_global)])
    return o


def _worker_slabop(rank, world, port, results):
    """`SlabOp`: ∂z, m̂z, divergence, gradient and (no halo) ∂x, weighted m̂y on a z-slab-decomposed grid over peer memory,
    against the single-GPU operators, forward / backward, Bloch (ring) / symmetric, `=` and `+=`."""
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import sgc_b200 as S
        from sgc_b200.slab import SlabOp, PeerArena, partition_planes
        rng = np.random.default_rng(8)
        N = (40, 18, 4 * world)
        k0, k1 = partition_planes(N[2], world)[rank]
        Nl = (N[0], N[1], k1 - k0)
        Ml = Nl[0] * Nl[1] * Nl[2]
        d = [rng.uniform(0.5, 1.5, n) + 1j * rng.uniform(-0.2, 0.2, n) for n in N]
        w = [rng.uniform(0.5, 1.5, n) for n in N]
        e = [np.exp(-0.3j), np.exp(0.5j), np.exp(-0.9j)]
        arena = PeerArena(2 * 3 * Ml * 16 + 1024)
        cplx = lambda shape: np.asfortranarray(rng.uniform(-1, 1, shape) + 1j * rng.uniform(-1, 1, shape))
        loc = lambda v: [v[0], v[1], v[2][k0:k1]]
        cut = lambda A: np.asfortranarray(A[:, :, k0:k1] if A.ndim == 3 else A[:, :, k0:k1, :])
        ok = True
        for fz in (True, False):
            for bz in (True, False):
                bl = [True, False, bz]
                fw = [False, True, fz]
                cases = [
                    ("dz", N, N, lambda n, v: S.Operator.partial(S.SGC_C128, n, 3, fz, v(d)[2], bz, e[2], alpha=0.5), fz, bz, [0]),
                    ("dx", N, N, lambda n, v: S.Operator.partial(S.SGC_C128, n, 1, fz, v(d)[0], True, e[0]), None, None, []),
                    ("mz", N, N, lambda n, v: S.Operator.mhat(S.SGC_C128, n, 3, fz, None, None, bz, e[2], alpha=2.0), fz, bz, [0]),
                    ("mz weighted", N, N, lambda n, v: _weighted_mz(S, n, fz, v(w)[2], v(d)[2], bz, e[2], w[2], k0, k1, n != N), fz, bz, [0]),
                    ("my weighted", N, N, lambda n, v: S.Operator.mhat(S.SGC_C128, n, 2, fz, v(w)[1], v(d)[1], False), None, None, []),
                    ("divg", N, N + (3,), lambda n, v: S.Operator.divg(S.SGC_C128, n, fw, v(d), bl, e), fz, bz, [2]),
                    ("grad", N + (3,), N, lambda n, v: S.Operator.grad(S.SGC_C128, n, fw, v(d), bl, e), fz, bz, [0]),
                ]
                for name, gshape, fshape, mk, f_last, b_last, gc in cases:
                    F, G0 = cplx(fshape), cplx(gshape)
                    for op in ("=", "+="):
                        whole = mk(N, lambda v: v)
                        Gw = S.DeviceArray.from_numpy(G0)
                        whole.apply(Gw, S.DeviceArray.from_numpy(F), op)
                        ref = cut(Gw.numpy())
                        so = SlabOp(mk(Nl, loc), f_last, b_last, gc, rank, world, arena)
                        arena.barrier()  # nobody may still be reading the arena fields of the previous case
                        arena.used = 64  # re-carve (after the flags) for this case's shapes; all ranks do the same
                        Ga, Fa = arena.array(cut(G0).shape), arena.array(cut(F).shape)
                        Ga.t.copy_(S.DeviceArray.from_numpy(cut(G0)).t); Fa.t.copy_(S.DeviceArray.from_numpy(cut(F)).t)
                        so.apply(Ga, Fa, op)
                        torch.cuda.synchronize()
                        good = np.array_equal(Ga.numpy(), ref)
                        if not good:
                            results[f"fail{rank}"] = f"{name} fz={fz} bz={bz} {op}"
                        ok = ok and good
        results[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_slab_other_operators_match_single_gpu():
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker_slabop, args=(world, _free_port(), results), nprocs=world, join=True)
    assert all(results.get(r) is True for r in range(world)), dict(results)
