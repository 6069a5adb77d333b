"""Probe: torch symmetric memory (peer-mapped buffers) + copy-engine pulls over NVLink.  torchrun --nproc-per-node N."""
import os
import time

import torch
import torch.distributed as dist

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
import torch.distributed._symmetric_memory as symm_mem  # noqa: E402

n = 512 * 1024 * 1024 // 4 * 4  # 2 GiB of float32
t = symm_mem.empty(n, dtype=torch.float32, device=dev)
t.fill_(float(rank + 1))
hdl = symm_mem.rendezvous(t, group=dist.group.WORLD)
dist.barrier()
torch.cuda.synchronize()
dst = torch.empty(n, dtype=torch.float32, device=dev)
streams = [torch.cuda.Stream() for _ in range(world)]
for rep in range(3):
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    total = 0
    for k in range(1, world):
        src = (rank - k) % world
        peer = hdl.get_buffer(src, (n,), torch.float32)
        with torch.cuda.stream(streams[k % len(streams)] if rep == 2 else torch.cuda.current_stream()):
            dst.copy_(peer, non_blocking=True)
        total += n * 4
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    ok = float(dst[0].item())
    if rank == 0:
        print(f"rep {rep}: pulled {total / 1e9:.1f} GB from {world - 1} peers in {dt * 1e3:.1f} ms = {total / dt / 1e9:.0f} GB/s (last value {ok})", flush=True)
dist.barrier()
dist.destroy_process_group()
