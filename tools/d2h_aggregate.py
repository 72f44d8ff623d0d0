"""Aggregate device-to-host bandwidth of the box: every rank copies a page-locked 26 MB buffer (the value-dependent
output bytes of one c2 evaluation) from its GPU to host memory, all ranks at the same time, with plain
cudaMemcpyAsync (torch .copy_ on pinned memory).  The host-delivered path of N concurrent evaluators cannot beat
N x 26 MB / this rate.  Launch: python -m torch.distributed.run --nproc-per-node N tools/d2h_aggregate.py"""
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nbytes = 26 * 1000 * 1000
src = torch.empty(nbytes // 8, dtype=torch.float64, device="cuda")
dst = torch.empty(nbytes // 8, dtype=torch.float64).pin_memory()
for reps in (200,):
    for _ in range(10):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"{world} GPUs x {reps} copies of {nbytes / 1e6:.0f} MB, all concurrent: slowest rank {t.item():.4f} s -> "
              f"{nbytes * reps / t.item() * 1e-9:.1f} GB/s per GPU, {world * nbytes * reps / t.item() * 1e-9:.1f} GB/s aggregate; "
              f"floor of one 26 MB delivery per rank = {t.item() / reps * 1e3:.3f} ms", flush=True)
if world > 1:
    dist.destroy_process_group()
