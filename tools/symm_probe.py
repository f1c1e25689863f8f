"""Probe: torch symmetric memory (peer buffers + device-side barrier) between the ranks of one node."""
import os, time
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
n = 7_000_000
buf = symm_mem.empty(n, dtype=torch.float32, device=dev)
hdl = symm_mem.rendezvous(buf, dist.group.WORLD.group_name)
print(rank, "rendezvous ok", type(hdl).__name__, "multicast_ptr", getattr(hdl, "multicast_ptr", None), flush=True)
buf.fill_(float(rank + 1))
torch.cuda.synchronize(); dist.barrier()
local = torch.empty(n, dtype=torch.float32, device=dev)
side = torch.cuda.Stream()
src = hdl.get_buffer(0, (n,), torch.float32)
for it in range(3):
    with torch.cuda.stream(side):
        hdl.barrier(channel=0)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        local.copy_(src, non_blocking=True)
        e1.record()
        hdl.barrier(channel=0)
    side.synchronize()
    print(rank, "pull", n * 4 / 1e6, "MB in", e0.elapsed_time(e1), "ms; value", float(local[123]), flush=True)
# chunked two-stage exchange: rank r pulls chunk r from rank 0, then the others' chunks from their owners
chunk = n // world
bufs = [hdl.get_buffer(r, (n,), torch.float32) for r in range(world)]
with torch.cuda.stream(side):
    hdl.barrier(channel=0)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    if rank != 0:
        buf[rank * chunk:(rank + 1) * chunk].copy_(bufs[0][rank * chunk:(rank + 1) * chunk], non_blocking=True)
    hdl.barrier(channel=0)
    for r in range(1, world):
        if r != rank and rank != 0:
            buf[r * chunk:(r + 1) * chunk].copy_(bufs[r][r * chunk:(r + 1) * chunk], non_blocking=True)
    if rank != 0:
        buf[:chunk].copy_(bufs[0][:chunk], non_blocking=True)
    e1.record()
    hdl.barrier(channel=0)
side.synchronize()
print(rank, "two-stage", e0.elapsed_time(e1), "ms", float(buf[5]), float(buf[n - 5]), flush=True)
dist.destroy_process_group()
