"""Experiment: z-slab curl with peer-memory halos through torch symmetric memory (run under torchrun)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm
import sgc_b200 as S

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
N = (256, 256, 256); M = N[0] * N[1] * N[2]; plane = N[0] * N[1]
t0 = time.time()
buf = symm.empty(3 * 3 * M * 2, dtype=torch.float64, device=f"cuda:{lr}")   # E, H, E2 as (re,im) doubles
hdl = symm.rendezvous(buf, dist.group.WORLD.group_name)
print(rank, "rendezvous ok", time.time() - t0, [hex(p) for p in hdl.buffer_ptrs], flush=True)
cbuf = torch.view_as_complex(buf.view(-1, 2))
E = S.DeviceArray(cbuf[0:3 * M], N + (3,)); H = S.DeviceArray(cbuf[3 * M:6 * M], N + (3,)); E2 = S.DeviceArray(cbuf[6 * M:9 * M], N + (3,))
gen = torch.Generator(device="cuda").manual_seed(7 + rank)
E.t.copy_(torch.complex(torch.rand(3 * M, generator=gen, device="cuda", dtype=torch.float64), torch.rand(3 * M, generator=gen, device="cuda", dtype=torch.float64)))
rng = np.random.default_rng(0)
dlz = rng.uniform(0.5, 1.5, N[2] * world) + 1j * rng.uniform(-.2, .2, N[2] * world)
dlx = rng.uniform(0.5, 1.5, N[0]) + 0j; dly = rng.uniform(0.5, 1.5, N[1]) + 0j
k0 = rank * N[2]
isb = [False, False, True]; e = [1.0, 1.0, np.exp(-0.4j)]
from sgc_b200.slab import SlabCurl, halo_plan
def make(isfwd):
    o = S.Operator.curl(S.SGC_C128, N, [isfwd] * 3, [dlx, dly, dlz[k0:k0 + N[2]]], isb, e)
    p = halo_plan(rank, world, isfwd, True)
    o.set_slab(p["lo_interface"], p["hi_interface"])
    return o, p
def peer_ptr(arr, peer):
    return hdl.buffer_ptrs[peer] + (arr.ptr - hdl.buffer_ptrs[rank])
def apply_p2p(o, p, G, F, op="="):
    hdl.barrier(channel=0)
    gh = None
    if p["recv_from"] is not None:
        base = peer_ptr(F, p["recv_from"])
        kpl = 0 if p["boundary"] == "top" else N[2] - 1
        gh = [base + (c * M + kpl * plane) * 16 for c in range(2)] + [None]
    o.apply_range(G, F, op, 0, N[2], gh)
of, pf = make(True); ob, pb = make(False)
# reference: NCCL path
scf = SlabCurl(S.SGC_C128, N, [True] * 3, [dlx, dly, dlz[k0:k0 + N[2]]], isb, e, rank, world)
Href = S.DeviceArray.zeros(N + (3,))
scf.apply(Href, E, "="); torch.cuda.synchronize(); dist.barrier()
apply_p2p(of, pf, H, E); torch.cuda.synchronize()
print(rank, "p2p == nccl:", bool(torch.equal(H.t, Href.t)), flush=True)
def step():
    apply_p2p(of, pf, H, E); apply_p2p(ob, pb, E2, H)
for _ in range(5): step()
torch.cuda.synchronize(); dist.barrier()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record(); th = time.perf_counter()
for _ in range(50): step()
host = (time.perf_counter() - th) / 50 * 1e3
ev1.record(); torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1) / 50
t = torch.tensor([ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0: print("p2p step ms (max over ranks)", float(t), "host enqueue ms", host, "Gcell/s", 2 * M * world / float(t) / 1e6, flush=True)
dist.barrier(); dist.destroy_process_group()
