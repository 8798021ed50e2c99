"""Host -> device copy ceiling of the box: N concurrent pinned cudaMemcpyAsync streams, one per GPU.

  python scripts/h2d_ceiling.py                      one process drives GPUs 0..n-1 for n = 1, 2, 4, 8 (as available)
  torchrun --nproc-per-node N scripts/h2d_ceiling.py  one process per GPU (each rank allocates -- first-touches -- its
                                                      own pinned buffer), all ranks copy at the same time

Answers VERDICT r1 weak #8: is the ~160-185 GB/s aggregate that the 8-GPU end-to-end bench saw the host's ceiling
(PCIe root complexes / memory bandwidth of the virtualised host) or an artefact of how bench.py allocates?
Plain cudaMemcpyAsync per copy (torch's non_blocking copy_ from pinned memory)."""
import json
import os
import sys
import time

import torch

GB = 1 << 30
SIZE = int(os.environ.get("H2D_BYTES", 2 * GB))
REPS = int(os.environ.get("H2D_REPS", 8))


def run_single(n):
    hosts = [torch.empty(SIZE, dtype=torch.uint8, pin_memory=True) for _ in range(n)]
    devs = [torch.empty(SIZE, dtype=torch.uint8, device=f"cuda:{i}") for i in range(n)]
    streams = [torch.cuda.Stream(device=i) for i in range(n)]
    for h in hosts:
        h.fill_(1)                                     # touch the pages
    for i in range(n):                                 # warm-up
        with torch.cuda.stream(streams[i]):
            devs[i].copy_(hosts[i], non_blocking=True)
    for i in range(n):
        torch.cuda.synchronize(i)
    t0 = time.perf_counter()
    for _ in range(REPS):
        for i in range(n):
            with torch.cuda.stream(streams[i]):
                devs[i].copy_(hosts[i], non_blocking=True)
    for i in range(n):
        torch.cuda.synchronize(i)
    dt = time.perf_counter() - t0
    return n * REPS * SIZE / dt / 1e9


def run_rank():
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    h = torch.empty(SIZE, dtype=torch.uint8, pin_memory=True)
    h.fill_(1)
    d = torch.empty(SIZE, dtype=torch.uint8, device="cuda")
    d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(REPS):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    per_rank = REPS * SIZE / float(dt.item()) / 1e9
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"mode": "one process per GPU", "n_gpus": world, "bytes_per_copy": SIZE, "copies": REPS,
                          "aggregate_GBps": world * REPS * SIZE / float(dt.item()) / 1e9, "rank0_GBps": per_rank}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    if "RANK" in os.environ and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        run_rank()
    else:
        ng = torch.cuda.device_count()
        for n in (1, 2, 4, 8):
            if n <= ng:
                print(json.dumps({"mode": "one process, one stream per GPU", "n_gpus": n, "bytes_per_copy": SIZE,
                                  "copies": REPS, "aggregate_GBps": run_single(n)}), flush=True)
