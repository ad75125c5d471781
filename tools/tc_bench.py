#!/usr/bin/env python
"""Micro-benchmark of PHMLinear (one shape) — used under ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from phc_gnn_b200 import ops
n, F, M, prec = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 10
DEV = "cuda:0"
x = torch.randn(M, F, device=DEV, requires_grad=True); A = torch.randn(n, n, n, device=DEV, requires_grad=True)
W = torch.randn(n, F // n, F // n, device=DEV, requires_grad=True); b = torch.randn(F, device=DEV, requires_grad=True)
gy = torch.randn(M, F, device=DEV)
for _ in range(2):
    y = ops.phm_linear(x, A, W, b, precision=prec); y.backward(gy)
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
e[0].record()
for _ in range(reps):
    y = ops.phm_linear(x, A, W, b, precision=prec)
e[1].record()
for _ in range(reps):
    y.backward(gy, retain_graph=True)
e[2].record()
torch.cuda.synchronize()
f, bw = e[0].elapsed_time(e[1]) / reps, e[1].elapsed_time(e[2]) / reps
fl = 2.0 * M * F * F
print(f"n={n} F={F} M={M} prec={prec}: fwd {f*1e3:.1f} us ({fl/f/1e9:.1f} TFLOP/s)  bwd {bw*1e3:.1f} us ({2*fl/bw/1e9:.1f} TFLOP/s)")
if os.environ.get("TC_PROF"):
    import ctypes
    from phc_gnn_b200 import _lib
    lib = _lib.load()
    buf = torch.zeros(148 * 8, dtype=torch.int64, device=DEV)
    lib.phc_debug_set_tc_profile.argtypes = [ctypes.c_void_p]
    lib.phc_debug_set_tc_profile(buf.data_ptr())
    y = ops.phm_linear(x, A, W, b, precision=prec)
    torch.cuda.synchronize()
    lib.phc_debug_set_tc_profile(None)
    r = buf.view(148, 8).float().cpu()
    names = ["prod wait_rfull", "prod work", "mma wait_tempty", "mma wait_afull", "mma wait_bfull", "drain (incl. wait)", "kernel total", "prod wait_aempty"]
    for i, nm in enumerate(names):
        print(f"  {nm:18s} mean {r[:, i].mean():12.0f}  min {r[:, i].min():12.0f}  max {r[:, i].max():12.0f} cycles")
