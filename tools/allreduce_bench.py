#!/usr/bin/env python
"""Latency of the gradient all-reduce alone (NCCL, ncclAvg, fp32) at the buffer sizes of the benchmark models, device-timed, max over
ranks.  torchrun --nproc-per-node N tools/allreduce_bench.py   (NCCL_ALGO / NCCL_PROTO from the environment)"""
import os
import torch
import torch.distributed as dist

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
out = []
for numel in (200_000, 1_170_000, 4_000_000):
    buf = torch.randn(numel, device=dev)
    for _ in range(20):
        dist.all_reduce(buf, op=dist.ReduceOp.AVG)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        dist.all_reduce(buf, op=dist.ReduceOp.AVG)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 200 * 1e3], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out.append(f"{numel * 4 / 1e6:.1f} MB: {float(t):.1f} us")
if rank == 0:
    print(f"world {dist.get_world_size()} NCCL_ALGO={os.environ.get('NCCL_ALGO', '-')} NCCL_PROTO={os.environ.get('NCCL_PROTO', '-')} "
          f"NCCL_NVLS_ENABLE={os.environ.get('NCCL_NVLS_ENABLE', '-')}: " + ", ".join(out), flush=True)
dist.destroy_process_group()
