#!/usr/bin/env python
"""In-situ kernel durations of the training step (CUPTI via torch.profiler: warm caches, real clocks, kernels
back to back) — complements the serialised cold-cache ncu launch list.

    python tools/kprof.py [workload] [steps]  > gpurun_out/kprof.txt
"""
import collections, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import ProfilerActivity, profile
from phc_gnn_b200 import graph
from phc_gnn_b200.nn import PHMSkipConnectAdd
from phc_gnn_b200.synthetic import make_batch, workloads
from phc_gnn_b200.train import TrainStep

name = sys.argv[1] if len(sys.argv) > 1 else "ppa"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
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
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(steps):
    graph.clear_cache(); step(batches[i % 4])
e1.record(); torch.cuda.synchronize()
print(f"# {name}: {e0.elapsed_time(e1) / steps:.3f} ms/step unprofiled")
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(steps):
        graph.clear_cache(); step(batches[i % 4])
    torch.cuda.synchronize()


def short(s):
    s = s.replace("(anonymous namespace)::", "").replace("void ", "")
    m = re.search(r"native::([A-Za-z_0-9]+)", s)
    if s.startswith("at::") and m:
        return "at::" + m.group(1)
    m = re.search(r"([A-Za-z_0-9:]+)\s*(<|\()", s)
    return m.group(1) if m else s[:60]


agg = collections.defaultdict(lambda: [0, 0.0])
spans = []
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        dur = ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
        k = short(ev.name)
        agg[k][0] += 1
        agg[k][1] += dur
        spans.append((ev.time_range.start, ev.time_range.end, k))
tot = sum(v[1] for v in agg.values())
spans.sort()
busy, cur_s, cur_e = 0.0, None, None
gaps = collections.defaultdict(lambda: [0, 0.0])
prev = None
for s, e, k in spans:
    if prev is not None and s > prev[1]:
        gaps[prev[2] + " -> " + k][0] += 1
        gaps[prev[2] + " -> " + k][1] += s - prev[1]
    prev = (s, e, k) if prev is None or e >= prev[1] else prev
    if cur_e is None or s > cur_e:
        if cur_e is not None:
            busy += cur_e - cur_s
        cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
if cur_e is not None:
    busy += cur_e - cur_s
wall = spans[-1][1] - spans[0][0] if spans else 0.0
print(f"# kernel time {tot / steps / 1e3:.3f} ms/step; device busy {busy / steps / 1e3:.3f} ms/step of {wall / steps / 1e3:.3f} ms/step wall (profiled)")
print(f"{'us/step':>10} {'share':>6} {'n/step':>7} {'avg us':>8}  kernel")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1] / steps:10.1f} {100 * v[1] / tot:5.1f}% {v[0] / steps:7.1f} {v[1] / v[0]:8.1f}  {k}")
print(f"\n# idle gaps between consecutive kernels (us/step), top 25 of {sum(v[1] for v in gaps.values()) / steps:.1f} us/step")
for k, v in sorted(gaps.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"{v[1] / steps:10.1f} {v[0] / steps:7.1f} {v[1] / v[0]:8.1f}  {k}")
