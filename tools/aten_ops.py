#!/usr/bin/env python
"""Which ATen operators (i.e. eager PyTorch kernels, not libphc_b200.so) run inside one training step, with shapes and the Python
call site — the residue that the fused C-ABI calls have not absorbed.     python tools/aten_ops.py [workload] [steps]"""
import collections, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import ProfilerActivity, profile
from phc_gnn_b200 import graph
from phc_gnn_b200.nn import PHMSkipConnectAdd
from phc_gnn_b200.synthetic import make_batch, workloads
from phc_gnn_b200.train import TrainStep

name = sys.argv[1] if len(sys.argv) > 1 else "ppa"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
wl = workloads(4)[name]
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = PHMSkipConnectAdd(**wl.model).to(dev)
step = TrainStep(model, wl)
model.train()
batches = [make_batch(wl, seed=i).to(dev) for i in range(4)]
for i in range(6):
    graph.clear_cache(); step(batches[i % 4])
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True, with_stack=True) as prof:
    for i in range(steps):
        graph.clear_cache(); step(batches[i % 4])
    torch.cuda.synchronize()
rows = []
for ev in prof.key_averages(group_by_input_shape=True, group_by_stack_n=8):
    dt = getattr(ev, "self_device_time_total", 0) or getattr(ev, "self_cuda_time_total", 0)
    if not ev.key.startswith("aten::") or dt <= 0:
        continue
    site = ""
    for fr in (ev.stack or []):
        if "phc_gnn_b200" in fr or "bench.py" in fr or "tools/" in fr:
            site = fr.split("/")[-1]
            break
    rows.append((ev.count / steps, dt / ev.count, ev.key, str(ev.input_shapes)[:70], site[:80]))
print(f"# {name}: aten ops with their own device time, per step (over {steps} steps)")
for r in sorted(rows, key=lambda r: -r[0]):
    print(f"{r[0]:6.1f} x {r[1]:6.1f} us  {r[2]:30s} {r[3]:72s} {r[4]}")
