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
