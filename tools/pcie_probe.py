"""Aggregate pinned D2H / H2D rate of N GPUs copying at the same time (torchrun): the platform ceiling for the e2e figure."""
import os
import time
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
n = 320 << 20
hp = torch.empty(n, dtype=torch.uint8).pin_memory()
dp = torch.empty(n, dtype=torch.uint8, device=dev)
hq = torch.empty(n // 4, dtype=torch.uint8).pin_memory()
dq = torch.empty(n // 4, dtype=torch.uint8, device=dev)


def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    if world > 1: dist.barrier()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / reps
    if world > 1:
        x = torch.tensor([dt], dtype=torch.float64, device=dev); dist.all_reduce(x, op=dist.ReduceOp.MAX); dt = float(x.item())
    return dt


d2h = timed(lambda: hp.copy_(dp, non_blocking=True))
h2d = timed(lambda: dq.copy_(hq, non_blocking=True))
if rank == 0:
    print(f"{world} GPU(s) at once: D2H {n / d2h / 1e9:.1f} GB/s per GPU ({world * n / d2h / 1e9:.1f} aggregate), "
          f"H2D {n / 4 / h2d / 1e9:.1f} GB/s per GPU; host cores {os.cpu_count()}")
if world > 1:
    dist.destroy_process_group()
