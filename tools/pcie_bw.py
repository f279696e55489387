"""Host <-> device copy bandwidth per GPU when all ranks copy at once (torchrun; evidence for the multi-GPU host-array leg).
Pinned host buffers, 512 MiB per direction, H2D alone, D2H alone, both at once on two streams; per-rank GB/s and the node total."""
import os, json, torch, torch.distributed as dist

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
nbytes = 512 << 20
h_in, h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory(), torch.empty(nbytes, dtype=torch.uint8).pin_memory()
d_in, d_out = torch.empty(nbytes, dtype=torch.uint8, device="cuda"), torch.empty(nbytes, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

def run(h2d, d2h, reps=6):
    def once():
        if h2d:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    once(); torch.cuda.synchronize()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): once()
    s1.synchronize(); s2.synchronize()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    return nbytes / ms / 1e6  # GB/s per direction

res = {}
for name, (h, d) in (("h2d_alone", (True, False)), ("d2h_alone", (False, True)), ("both_at_once_per_direction", (True, True))):
    v = torch.tensor([run(h, d)], device="cuda", dtype=torch.float64)
    if world > 1:
        allv = [torch.zeros_like(v) for _ in range(world)]
        dist.all_gather(allv, v)
        vals = [float(x) for x in allv]
    else:
        vals = [float(v)]
    res[name] = {"per_rank_gbs": [round(x, 1) for x in vals], "node_total_gbs": round(sum(vals), 1)}
if rank == 0:
    print(json.dumps({"gpus": world, "bytes_per_copy": nbytes, **res}))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
