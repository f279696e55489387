"""Where does a z-slab curl-of-curl step spend its time at N GPUs?  Per-rank CUDA-event timings of the same
step (two fused curls on a 256³ slab per GPU, bench.py's inputs) under successively more coupling:

  solo         every rank runs the single-GPU operators on its own slab in ordinary memory (no neighbours at all)
  peer-nosync  fields in the peer-mapped arena, ghost planes read from the neighbours, NO ordering between ranks
               (results invalid — timing only: the cost of the mapping and of the NVLink halo loads)
  barrier      sgc_peer_barrier kernel before every apply
  flags        neighbour synchronisation inside the kernel, boundary chunks last (eager launches)
  flags-graph  the same step replayed from a CUDA graph

  torchrun --nproc-per-node N tools/slab_timeline.py [--steps 200] [--slab 256,256,256]
Prints one JSON line per mode on rank 0: per-rank ms per step, and the max (what bench.py reports).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--slab", default="256,256,256")
    ap.add_argument("--modes", default="solo,peer-nosync,barrier,flags,flags-graph")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import bench as B
    import sgc_b200 as S
    from sgc_b200.slab import SlabCurl, PeerArena

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    nl = tuple(int(v) for v in args.slab.split(","))
    M = nl[0] * nl[1] * nl[2]
    dprim, ddual = B.build_inputs(nl[2] * world, nlocal=nl)
    k0, k1 = rank * nl[2], (rank + 1) * nl[2]
    bl, e = [False] * 3, [1.0] * 3
    dd = [ddual[0], ddual[1], ddual[2][k0:k1]]
    dp = [dprim[0], dprim[1], dprim[2][k0:k1]]
    arena = PeerArena(3 * 3 * M * 16 + 4096) if world > 1 else None

    def rnd():
        return torch.complex(torch.rand(3 * M, device="cuda", dtype=torch.float64) * 2 - 1,
                             torch.rand(3 * M, device="cuda", dtype=torch.float64) * 2 - 1)

    def sync_all():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    def measure(step, graph=False):
        for _ in range(5):
            step()
        sync_all()
        run = step
        if graph:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=side, capture_error_mode="thread_local"):
                    step()
            torch.cuda.current_stream().wait_stream(side)
            run = g.replay
            run(); run()
            sync_all()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(args.steps):
            run()
        ev1.record()
        sync_all()
        t = torch.tensor([ev0.elapsed_time(ev1) / args.steps], device="cuda", dtype=torch.float64)
        allt = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        return [round(float(x.item()), 4) for x in allt]

    out = []
    for mode in args.modes.split(","):
        if mode == "solo":
            E = S.DeviceArray(rnd(), nl + (3,))
            H, E2 = S.DeviceArray.zeros(nl + (3,)), S.DeviceArray.zeros(nl + (3,))
            c1 = S.Operator.curl(S.SGC_C128, nl, [True] * 3, dd, bl, e)
            c2 = S.Operator.curl(S.SGC_C128, nl, [False] * 3, dp, bl, e)
            per = measure(lambda: (c1.apply(H, E, "="), c2.apply(E2, H, "=")))
            del E, H, E2
        elif arena is None:
            continue
        else:
            arena.used = 0
            E, H, E2 = (arena.array(nl + (3,)) for _ in range(3))
            E.t.copy_(rnd())
            sync = "flags" if mode.startswith("flags") else "barrier"
            s1 = SlabCurl(S.SGC_C128, nl, [True] * 3, dd, bl, e, rank, world, arena=arena, sync=sync)
            s2 = SlabCurl(S.SGC_C128, nl, [False] * 3, dp, bl, e, rank, world, arena=arena, sync=sync)
            if mode == "peer-nosync":
                def ghost(s, F):
                    p = s.plan
                    if p["recv_from"] is None:
                        return None
                    base = arena.peer_ptr(F, p["recv_from"])
                    kpl = 0 if p["boundary"] == "top" else nl[2] - 1
                    plane = nl[0] * nl[1]
                    return [base + (c * M + kpl * plane) * 16 for c in range(2)] + [None]
                g1, g2 = ghost(s1, E), ghost(s2, H)
                per = measure(lambda: (s1.backend.op.apply_range(H, E, "=", 0, nl[2], g1),
                                       s2.backend.op.apply_range(E2, H, "=", 0, nl[2], g2)))
            else:
                sync_all()
                arena.barrier()
                per = measure(lambda: (s1.apply(H, E, "="), s2.apply(E2, H, "=")), graph=(mode == "flags-graph"))
            del E, H, E2, s1, s2
        torch.cuda.empty_cache()
        out.append({"mode": mode, "n_gpus": world, "slab": list(nl), "steps": args.steps, "ms_per_step_max": max(per),
                    "ms_per_step_min": min(per), "per_rank_ms": per})
        if rank == 0:
            print(json.dumps(out[-1]), flush=True)
    sync_all()
    if arena is not None:
        arena.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
